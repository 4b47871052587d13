#!/usr/bin/env python
"""BASELINE.json config 4: qc_ldpc_tanner + bicycle codes, min-sum layered, 10^8 shots sharded over the GPUs of one node,
counters reduced with one NCCL all-reduce.

    python benchmarks/cfg4_scaling.py --code T --shots 100000000                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        benchmarks/cfg4_scaling.py --code T --shots 100000000                         (8 GPUs)

Runs the public driver simulate_p(..., sampler_kind='device'): each rank samples and decodes its contiguous range of global
shot indices; the sampler is keyed by the global shot index, so the counters must be IDENTICAL for every number of GPUs
(SURVEY.md section 8e acceptance test) -- rank 0 prints them with the wall time (sampling + decoding + classification +
reduction, CUDA-synchronised, max over ranks)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qldpcsim_b200 import pcmlibrary, simulator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="T")
    ap.add_argument("--shots", type=int, default=100_000_000)
    ap.add_argument("--p", type=float, default=0.03)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--classes", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    Hx, Hz = pcmlibrary.by_name(a.code)
    kw = dict(shots=a.shots, decType="MS", decIterations=a.iters, decSchedule="L", sampler_kind="device", rngSeed=2026,
              device=local, classes=a.classes)
    simulator.simulate_p(Hx, Hz, a.p, **{**kw, "shots": min(a.shots, 200_000 * world)})        # warm-up (plans, NCCL)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    res = simulator.simulate_p(Hx, Hz, a.p, **kw)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"config": f"cfg4 {a.code} MS-L {a.iters} it p={a.p}", "n_gpus": world, "shots": a.shots,
                          "seconds": round(float(dt.item()), 3), "shots_per_s": round(a.shots / float(dt.item()), 1),
                          "result": res}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
