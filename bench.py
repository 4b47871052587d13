#!/usr/bin/env python
"""
bench.py -- headline benchmark: X+Z shots decoded per second, LP118_0 lifted-product code, normalised min-sum,
layered schedule, 50 iterations, OSD off, depolarizing p = 0.05, 10^6 shots per step per GPU (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one resident batch: decode the X errors (plan on Hz), decode the Z errors
(plan on Hx), classify + count (simulator.py:244-304 of the reference), followed by the one collective of the path,
an all-reduce of the int64[10] counters.  Prints ONE JSON line on rank 0.

  value     : shots/s with the bit-packed syndromes/errors already in HBM (device sampler, untimed), CUDA events,
              max over ranks.
  e2e       : same metric through the host-buffer C-ABI call (qldpc_decode_host): pinned host syndromes in,
              error estimates + iteration counts + convergence flags back in pinned host memory, copies inside
              the timed region.
  roofline  : the dominant kernel (ms_decode_kernel); achieved = algorithmic bytes (16 B per edge-iteration for the
              layered schedule, SURVEY.md section 8d, + bit-packed I/O) / its CUDA-event time, against the measured HBM copy
              bandwidth of MEASURED_PEAKS.json.  The kernel keeps its state in shared memory, so real DRAM traffic is
              far below the algorithmic figure; `traffic` comes from profiles/ when an ncu capture exists.
  cpu_baseline / --impl reference : the reference's own CPU implementation of the path on this box's host cores, on a bounded
              sample of the same workload: the UNMODIFIED reference (decoders.MS_decoder called as simulator.py:278-279 does,
              pip-installed into baseline/_ref by baseline/install_reference.sh; one process per core) -- kind "reference" --
              and, beside it, the oracle's C port of the same decoder (OpenMP over shots) -- "port".  Without baseline/_ref
              only the port is timed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CODE = "LP118_0"
P_DEPOL = 0.05
DEC_ITERS = 50
SCHEDULE = "L"
SHOTS_PER_STEP = 1_000_000
BYTES_PER_EDGE_ITER = 16.0          # layered / serial: c2v read+write, posterior read+write
CPU_SAMPLE_SHOTS = 262144         # cpu_baseline leg, C port: ~3.5 s wall on 16 host threads (~1 core-minute)
REF_STEP_SHOTS = 65536            # --impl reference without baseline/_ref (C port): ~0.9 s per step on 16 host threads
REF_PY_SHOTS_PER_CORE = 8         # unmodified reference: ~70 ms per X+Z shot and core => ~0.6 s per step
METRIC = "shots/sec decoded (X+Z), LP118_0 MS-layered 50 it"
UNIT = "shots/s"


def workload_name(shots):
    return (f"{CODE} lifted-product code [[544,80]] (Hx,Hz 240x544, 1920 edges each), min-sum layered (13/11 layers), "
            f"{DEC_ITERS} iterations, OSD off, depolarizing p={P_DEPOL}, {shots} shots per step per GPU")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def pin_to_gpu_numa_node(index):
    """Run this rank's host thread on the CPUs of the NUMA node its GPU hangs off (8 ranks share one box)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def recorded_on_chip():
    """Issue-slot and shared-memory utilisation of the dominant kernel from the committed ncu capture (profiles/on_chip.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "on_chip.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        busy = [x for x in sm if x >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------
def cpu_decode_rate(shots, seed, threads, repeat=1):
    from oracle import oracle
    from qldpcsim_b200 import pcmlibrary, sampler
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(CODE)]
    rec = sampler.sample_record(Hx, Hz, P_DEPOL, shots, seed=seed)
    sy_z, sy_x, _, _ = oracle.split_record(rec, Hz.shape[0], Hx.shape[0], Hx.shape[1])
    lX, lZ = oracle.schedule_layers(Hx, Hz, SCHEDULE)
    gz, gx = oracle.Graph(Hz), oracle.Graph(Hx)
    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        gz.decode("MS", sy_z, p=P_DEPOL / 3, max_iter=DEC_ITERS, layers=lX, n_threads=threads)     # simulator.py:278
        gx.decode("MS", sy_x, p=P_DEPOL / 3, max_iter=DEC_ITERS, layers=lZ, n_threads=threads)     # simulator.py:279
        times.append(time.perf_counter() - t0)
    return times


# ---- the unmodified reference (baseline/_ref), one process per core ------------------------------------------------
_REF = {}


def _ref_init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_loader
    from qldpcsim_b200 import pcmlibrary
    _REF["dec"] = ref_loader.load("decoders")
    layerize = ref_loader.load_layerize()
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(CODE)]          # what simulator.load_matrix returns (:35)
    _REF["Hx"], _REF["Hz"] = Hx, Hz
    _REF["lX"], _REF["lZ"] = layerize(Hx, serial=False), layerize(Hz, serial=False)   # simulator.py:230-234, schedule 'L'


def _ref_chunk(job):
    sy_z, sy_x = job
    dec, it = _REF["dec"], 0
    for a, b in zip(sy_z, sy_x):
        _, i1 = dec.MS_decoder(_REF["Hz"], a.astype(int), p=P_DEPOL / 3, max_iter=DEC_ITERS, layers=_REF["lX"], OSDorder=-1)   # :278
        _, i2 = dec.MS_decoder(_REF["Hx"], b.astype(int), p=P_DEPOL / 3, max_iter=DEC_ITERS, layers=_REF["lZ"], OSDorder=-1)   # :279
        it += i1 + i2
    return it


def reference_available():
    from oracle import ref_loader
    return ref_loader.available()


class ReferencePool:
    """Times the unmodified reference's MS decoder on `shots` shots per call, spread over `procs` worker processes."""

    def __init__(self, procs):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs, initializer=_ref_init)
        self.pool.map(_ref_chunk, [(np.zeros((0, 1)), np.zeros((0, 1)))] * procs)          # workers up, reference imported

    def rate(self, shots, seed):
        from oracle import oracle
        from qldpcsim_b200 import pcmlibrary, sampler
        Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(CODE)]
        rec = sampler.sample_record(Hx, Hz, P_DEPOL, shots, seed=seed)
        sy_z, sy_x, _, _ = oracle.split_record(rec, Hz.shape[0], Hx.shape[0], Hx.shape[1])
        jobs = [(sy_z[a:a + 2], sy_x[a:a + 2]) for a in range(0, shots, 2)]      # small jobs: a non-converging shot costs 20x a typical one
        t0 = time.perf_counter()
        self.pool.map(_ref_chunk, jobs, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.terminate()


def config_block(shots):
    """The keys both arms print (same workload, same labels)."""
    return {"workload": workload_name(shots), "decType": "MS", "decSchedule": "L", "decIterations": DEC_ITERS, "p": P_DEPOL,
            "shots_per_step_per_gpu": shots}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from oracle import oracle
    oracle.build()
    if reference_available():
        step_shots = max(200, REF_PY_SHOTS_PER_CORE * cores)
        pool = ReferencePool(cores)
        times = [pool.rate(step_shots, 1234 + k) for k in range(args.warmup + args.steps)][args.warmup:]
        pool.close()
        kind = "reference"
        sample = (f"{step_shots} shots of that workload per step x {len(times)} steps, the unmodified reference (baseline/_ref, "
                  f"decoders.MS_decoder called as simulator.py:278-279), {cores} worker processes")
    else:
        step_shots = REF_STEP_SHOTS
        times = cpu_decode_rate(step_shots, 1234, cores, repeat=args.warmup + args.steps)[args.warmup:]
        kind = "port"
        sample = (f"{step_shots} shots per step x {len(times)} steps, oracle/qldpc_oracle.c (C port of decoders.py MS_decoder, OpenMP "
                  f"over shots); baseline/_ref is absent")
    total = sum(times)
    value = step_shots * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {**config_block(SHOTS_PER_STEP), "sample_shots_per_step": step_shots},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from qldpcsim_b200 import _lib, bitpack, pcmlibrary, simulator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    shots = args.shots
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(CODE)]
    pipe = simulator.Pipeline(Hx, Hz, P_DEPOL, "MS", DEC_ITERS, SCHEDULE, device=local)
    E = pipe.decX.pcm.nnz
    n_batches = min(4, max(1, args.steps))
    # resident synthetic batches: global shot index = (step slot * world + rank) * shots + s
    batches = [pipe.sample_device(shots, seed=20261018, first_shot=(b * world + rank) * shots) for b in range(n_batches)]
    torch.cuda.synchronize(dev)

    lib = _lib.lib()
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    nw, mzw, mxw = bitpack.words(pipe.n), bitpack.words(pipe.m_z), bitpack.words(pipe.m_x)
    outX = (torch.empty((shots, nw), dtype=torch.int32, device=dev), torch.empty(shots, dtype=torch.int32, device=dev), None, None)
    outZ = (torch.empty((shots, nw), dtype=torch.int32, device=dev), torch.empty(shots, dtype=torch.int32, device=dev), None, None)

    def step(b, ev=None):
        synz, synx, errx, errz = batches[b % n_batches]
        st = torch.cuda.current_stream(dev)
        if ev:
            ev[0].record(st)
        pipe.decX.decode_packed(synz, out=outX)
        if ev:
            ev[1].record(st)
        pipe.decZ.decode_packed(synx, out=outZ)
        if ev:
            ev[2].record(st)
        _lib.check(lib.qldpc_classify(pipe.decX.handle, pipe.decZ.handle, errx.data_ptr(), errz.data_ptr(), outX[0].data_ptr(),
                                      outZ[0].data_ptr(), synz.data_ptr(), synx.data_ptr(), outX[1].data_ptr(), outZ[1].data_ptr(),
                                      shots, counters.data_ptr(), st.cuda_stream))
        if world > 1:
            dist.all_reduce(counters)          # the path's only collective (80 bytes); counters are re-zeroed per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clk = ClockSampler(local) if rank == 0 else None     # started before the warm-up: nvidia-smi needs ~0.5 s to deliver its first line
    for w in range(args.warmup):
        counters.zero_()
        step(w)
    barrier()
    pipe.decX.work_done(reset=True)
    pipe.decZ.work_done(reset=True)
    launches0 = lib.qldpc_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step_counters = []
    barrier()
    t_begin.record(torch.cuda.current_stream(dev))
    for k in range(args.steps):
        counters.zero_()
        step(k, evs[k])
        per_step_counters.append(counters.clone())
    t_end.record(torch.cuda.current_stream(dev))
    barrier()
    elapsed_ms = t_begin.elapsed_time(t_end)
    launches = lib.qldpc_launch_count() - launches0
    executed = (pipe.decX.work_done() + pipe.decZ.work_done()) / args.steps      # messages actually computed per step (this rank)
    clocks = clk.stop() if clk else None
    tX = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    tZ = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    cs = torch.stack(per_step_counters).cpu().numpy().astype(np.float64)
    # counters were all-reduced: per-rank share for the roofline (weak scaling, identical distributions)
    itX, itZ = cs[:, _lib.CNT_ITERS_X].mean() / world, cs[:, _lib.CNT_ITERS_Z].mean() / world
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = shots * world * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: the same step through the host-buffer C-ABI call (qldpc_simulate_host): the bit-packed record in pinned host
    # memory in, both decodes + classification on the device, the counters back on the host; copies inside the timed region
    pin_to_gpu_numa_node(local)
    host = [torch.empty(tuple(x.shape), dtype=torch.int32).pin_memory() for x in batches[0]]      # syn_z, syn_x, errX, errZ
    for h, d in zip(host, batches[0]):
        h.copy_(d.cpu())
    h_cnt = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64).pin_memory()

    def e2e_step():
        pipe.run_host(host[0], host[1], host[2], host[3], counters=h_cnt)

    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = shots * world * e2e_steps / float(te.item())
    # the host path must give the counters of the device-resident path on the same batch
    dev_cnt = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    pipe.run(*batches[0], counters=dev_cnt)
    torch.cuda.synchronize(dev)
    assert torch.equal(h_cnt, dev_cnt.cpu()), f"host path differs from device path: {h_cnt.tolist()} vs {dev_cnt.cpu().tolist()}"

    if rank == 0:
        peak, peak_src = measured_peak()
        io_bytes = shots * (4 * mzw + 4 * nw + 4) + shots * (4 * mxw + 4 * nw + 4)
        alg_bytes = (itX + itZ) * E * BYTES_PER_EDGE_ITER + io_bytes
        achieved = alg_bytes / ((tX + tZ) * 1e-3) / 1e9
        tr = recorded_traffic()
        oc = recorded_on_chip()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {**config_block(shots), "sharding": f"shots x {world} ranks, counters all-reduced (NCCL)" if world > 1 else "single GPU",
                       "l2": "inputs+outputs per step (~345 MB) exceed the 126 MB L2; 4 resident batches cycled",
                       "avg_iters_X": itX / shots, "avg_iters_Z": itZ / shots,
                       "edge_iterations_per_s": (itX + itZ) * E * world / (elapsed_ms / args.steps * 1e-3),
                       "edge_updates_credited": (itX + itZ) * E, "edge_updates_executed": executed},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": shots * 4 * (mzw + mxw + 2 * nw),
                    "d2h_bytes_per_step": 8 * _lib.NUM_COUNTERS,
                    "api": "Pipeline.run_host -> qldpc_simulate_host: pinned host record [sy_z | sy_x | errX | errZ] in, decode X + decode Z + "
                           "classification (simulator.py:244-304), int64[10] counters back on the host"},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "kernel": "ms_decode_kernel<8,5,3,24,1>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "bytes_per_edge_iteration": BYTES_PER_EDGE_ITER, "kernel_ms_per_step": tX + tZ,
                         "kernel_share_of_step": (tX + tZ) / (elapsed_ms / args.steps),
                         "traffic": tr.get("dram_bytes_per_launch") if tr else None,
                         "frac_executed": (executed * BYTES_PER_EDGE_ITER + io_bytes) / ((tX + tZ) * 1e-3) / 1e9 / peak,
                         "on_chip": oc,
                         "note": "achieved / frac credit whole iterations (SURVEY.md section 8d: E x returned iterations); a decode that "
                                 "converges inside an iteration executes less -- frac_executed counts the messages the kernel actually "
                                 "computed (device counter).  The message state stays in shared memory, so DRAM traffic is ~3 orders of "
                                 "magnitude below the algorithmic bytes and the HBM model is notional: the real bound is on chip "
                                 "(on_chip: issue slots and shared-memory wavefronts of the same kernel, from the committed ncu capture)"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            from oracle import oracle
            oracle.build()
            cores = os.cpu_count() or 1
            tcpu = cpu_decode_rate(CPU_SAMPLE_SHOTS, 1234, cores)[0]
            port = {"value": CPU_SAMPLE_SHOTS / tcpu, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{CPU_SAMPLE_SHOTS} shots of the same workload (host sampler seed 1234), "
                              f"oracle/qldpc_oracle.c MS decode X+Z, OpenMP {cores} threads"}
            if reference_available():
                n_ref = max(256, 16 * cores)                    # ~70 ms per shot and core => ~1-2 s wall
                pool = ReferencePool(cores)
                tref = pool.rate(n_ref, 1234)
                pool.close()
                line["cpu_baseline"] = {"value": n_ref / tref, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{n_ref} shots of the same workload (host sampler seed 1234), the unmodified reference "
                                                  f"(baseline/_ref, decoders.MS_decoder X+Z called as simulator.py:278-279), {cores} worker processes",
                                        "port": port}
            else:
                line["cpu_baseline"] = port
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shots", type=int, default=SHOTS_PER_STEP, help="shots per step per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
