"""
Monte-Carlo driver -- B200 drop-in for the reference's `simulator` module (simulator.py:167-373).

`simulate_p` reproduces the shot loop of simulator.py:244-304 as four batched device passes:
    sample (or take a caller-supplied record) -> decode X errors on Hz -> decode Z errors on Hx -> classify + count.
The reference's wiring is kept verbatim (simulator.py:270-282): MS/BP get the prior p/3, decIterations and the layer
lists; the partition built from Hx (`layersX`) is used for the decode on Hz and vice versa; MS gets OSDorder, BP
never does; NG and BF get neither iterations nor schedule (BF therefore always runs its default 50 iterations).

Shots are independent, so with torch.distributed initialised the shots are sharded over the ranks (one process per
GPU) and the only collective is one all-reduce(SUM) of the int64[10] outcome counters.
"""
from __future__ import annotations

import argparse
from typing import Optional, Tuple

import numpy as np

from . import _lib, bitpack, sampler
from .decoders import Decoder
from .pcm import load_matrix, logical_operators, schedule_layers  # noqa: F401  (load_matrix re-exported like the reference)

DEFAULT_SEED = 0x5EED


# ---------------------------------------------------------------------------------------------------------
# sharding helpers (pure host logic, covered by the gloo tests)
# ---------------------------------------------------------------------------------------------------------
def shard_range(shots: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous global shot range [lo, hi) of `rank`: GPU g decodes [g*S/G, (g+1)*S/G)."""
    return (shots * rank) // world, (shots * (rank + 1)) // world


def dist_info() -> Tuple[int, int]:
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def reduce_counters(counters):
    """all_reduce(SUM) of the int64[10] counter vector over the default process group (NCCL on GPU tensors, gloo on
    CPU tensors).  No-op without an initialised process group."""
    _, world = dist_info()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def counters_to_result(c, shots: int, classes: bool = False) -> dict:
    """The dict simulate_p returns (simulator.py:308-315); with `classes` also the four outcome classes of the
    reference's README.md:15-22, which its own counters cannot separate (extension)."""
    c = [int(x) for x in c]
    extra = {}
    if classes:
        extra = {"outcome_exact": c[_lib.CNT_EXACT], "outcome_degenerate": c[_lib.CNT_TRUE_DEGEN],
                 "outcome_logical_error": c[_lib.CNT_LOGICAL], "outcome_decoder_failure": c[_lib.CNT_FAIL_ANY]}
    return {**{
        "DecFailures_X": c[_lib.CNT_FAIL_X],
        "DecFailures_Z": c[_lib.CNT_FAIL_Z],
        "decSuccessExact": c[_lib.CNT_EXACT],
        "decSuccessDegen": c[_lib.CNT_DEGEN],
        "Avg_number_of_iterations_X": c[_lib.CNT_ITERS_X] / float(shots),
        "Avg_number_of_iterations_Z": c[_lib.CNT_ITERS_Z] / float(shots),
    }, **extra}


# ---------------------------------------------------------------------------------------------------------
# the batched pipeline
# ---------------------------------------------------------------------------------------------------------
class Pipeline:
    """Device pipeline for one (Hx, Hz, p, decoder configuration): two decode plans + sampler + classifier."""

    def __init__(self, Hx, Hz, p: float, decType: str = "MS", decIterations: int = 99, decSchedule: str = "F",
                 OSDorder: int = -1, device: Optional[int] = None, logicals: bool = False, kernel: str = "auto"):
        Hx = (np.asarray(Hx) % 2).astype(np.int8)
        Hz = (np.asarray(Hz) % 2).astype(np.int8)
        if Hx.shape[1] != Hz.shape[1]:
            raise ValueError("Hx and Hz must have the same number of columns (physical qubits).")   # simulator.py:67
        if decType not in ("NG", "BF", "MS", "BP"):
            raise ValueError("Unrecognized decoder type.")                                           # simulator.py:284
        layersX, layersZ = schedule_layers(Hx, Hz, decSchedule)                                      # simulator.py:228-236
        self.Hx, self.Hz, self.p = Hx, Hz, float(p)
        self.m_x, self.n = Hx.shape
        self.m_z = Hz.shape[0]
        kw = {"device": device}
        if decType == "NG":                                                                          # simulator.py:272-273
            self.decX, self.decZ = Decoder(Hz, "NG", **kw), Decoder(Hx, "NG", **kw)
        elif decType == "BF":                                                                        # simulator.py:275-276
            self.decX, self.decZ = Decoder(Hz, "BF", max_iter=50, **kw), Decoder(Hx, "BF", max_iter=50, **kw)
        elif decType == "MS":                                                                        # simulator.py:278-279
            self.decX = Decoder(Hz, "MS", p=p / 3, max_iter=decIterations, layers=layersX, OSDorder=OSDorder, kernel=kernel, **kw)
            self.decZ = Decoder(Hx, "MS", p=p / 3, max_iter=decIterations, layers=layersZ, OSDorder=OSDorder, kernel=kernel, **kw)
        else:                                                                                        # simulator.py:281-282
            self.decX = Decoder(Hz, "BP", p=p / 3, max_iter=decIterations, layers=layersX, **kw)
            self.decZ = Decoder(Hx, "BP", p=p / 3, max_iter=decIterations, layers=layersZ, **kw)
        import torch
        self.torch = torch
        self.device = torch.device("cuda", self.decX.device)
        self.logicals = bool(logicals)
        if logicals:
            # true outcome classes (README.md:15-22): the X residual is tested against the logical Z operators (attached to
            # the plan on Hz), the Z residual against the logical X operators (plan on Hx)
            Lx, Lz = logical_operators(Hx, Hz)
            for dec, L in ((self.decX, Lz), (self.decZ, Lx)):
                rows = np.ascontiguousarray(bitpack.pack_rows(L.astype(bool))) if L.shape[0] else np.zeros((0, bitpack.words(self.n)), np.uint32)
                _lib.check(_lib.lib().qldpc_plan_set_logicals(dec.handle, rows.ctypes.data if rows.size else None, int(L.shape[0])))
            self.k = int(Lx.shape[0])

    # -- inputs -------------------------------------------------------------------------------------------
    def upload_record(self, record: np.ndarray):
        """Host bool record [sy_z | sy_x | errX | errZ] (simulator.py:249-252) -> bit-packed device tensors."""
        t = self.torch
        r = np.asarray(record)
        mz, mx, n = self.m_z, self.m_x, self.n
        parts = (r[:, :mz], r[:, mz:mz + mx], r[:, mz + mx:mz + mx + n], r[:, mz + mx + n:mz + mx + 2 * n])
        return tuple(t.from_numpy(bitpack.pack_rows(x).view(np.int32)).to(self.device) for x in parts)

    def sample_device(self, shots: int, seed: int, first_shot: int = 0):
        """On-device depolarizing sampler (qldpc_sample): returns (syn_z, syn_x, errX, errZ) bit-packed tensors."""
        t = self.torch
        nw, mzw, mxw = bitpack.words(self.n), bitpack.words(self.m_z), bitpack.words(self.m_x)
        errx = t.empty((shots, nw), dtype=t.int32, device=self.device)
        errz = t.empty((shots, nw), dtype=t.int32, device=self.device)
        synz = t.empty((shots, mzw), dtype=t.int32, device=self.device)
        synx = t.empty((shots, mxw), dtype=t.int32, device=self.device)
        st = t.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().qldpc_sample(self.decX.handle, self.decZ.handle, self.p, int(seed) & (2**64 - 1), int(first_shot),
                                          int(shots), errx.data_ptr(), errz.data_ptr(), synz.data_ptr(), synx.data_ptr(), st))
        return synz, synx, errx, errz

    # -- one pass -----------------------------------------------------------------------------------------
    def run(self, synz, synx, errx, errz, counters=None, keep: bool = False, mid_event=None):
        """decode X, decode Z, classify.  Returns the int64[10] device counter tensor (accumulated into `counters`)."""
        t = self.torch
        if counters is None:
            counters = t.zeros(_lib.NUM_COUNTERS, dtype=t.int64, device=self.device)
        shots = synz.shape[0]
        ex, itx, cvx, _ = self.decX.decode_packed(synz, want_converged=keep)
        ez, itz, cvz, _ = self.decZ.decode_packed(synx, want_converged=keep)
        if mid_event is not None:
            mid_event.record(t.cuda.current_stream(self.device))
        st = t.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().qldpc_classify(self.decX.handle, self.decZ.handle, errx.data_ptr(), errz.data_ptr(),
                                            ex.data_ptr(), ez.data_ptr(), synz.data_ptr(), synx.data_ptr(),
                                            itx.data_ptr(), itz.data_ptr(), shots, counters.data_ptr(), st))
        if keep:
            self.last = {"eX": ex, "eZ": ez, "itX": itx, "itZ": itz, "convX": cvx, "convZ": cvz}
        return counters


    def run_host(self, syn_z, syn_x, err_x, err_z, counters=None):
        """The whole shot loop on HOST buffers (qldpc_simulate_host): bit-packed int32 host tensors / arrays of the record's four
        column groups in, the int64[10] outcome counters out (a pinned host tensor, returned).  Copies, both decodes and the
        classification are pipelined chunk by chunk inside the library."""
        t = self.torch
        if counters is None:
            counters = t.zeros(_lib.NUM_COUNTERS, dtype=t.int64)
        ptr = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        _lib.check(_lib.lib().qldpc_simulate_host(self.decX.handle, self.decZ.handle, ptr(syn_z), ptr(syn_x), ptr(err_x), ptr(err_z),
                                                 int(syn_z.shape[0]), ptr(counters)))
        return counters

    def decode_host(self, syn_z, syn_x, out_x, out_z):
        """Host-buffer decode of both error types (qldpc_decode_host, pinned or pageable memory): syn_* are int32 host tensors
        (shots, words(m)), out_* = (ehat int32 (shots, words(n)), iters int32 (shots,), converged uint8 (shots,)).  The two
        decodes are independent plans with their own streams and staging buffers; they are issued from two host threads so
        that the copy-in of one overlaps the kernels of the other and the tail of a chunk of one decode (a few long-running
        shots keep their SMs) is filled by CTAs of the other."""
        import threading
        L = _lib.lib()
        shots = int(syn_z.shape[0])
        errs = []

        def work(dec, hs, ho):
            try:
                _lib.check(L.qldpc_decode_host(dec.handle, hs.data_ptr(), shots, ho[0].data_ptr(), ho[1].data_ptr(),
                                               ho[2].data_ptr() if ho[2] is not None else None, None))
            except Exception as e:          # re-raised in the caller's thread
                errs.append(e)
        th = threading.Thread(target=work, args=(self.decZ, syn_x, out_z))
        th.start()
        work(self.decX, syn_z, out_x)
        th.join()
        if errs:
            raise errs[0]


def simulate_p(Hx: np.ndarray, Hz: np.ndarray, p: float, shots: int = 1000, decType: str = "MS",
               decIterations: int = 99, decSchedule: str = "F", OSDorder: int = -1, rngSeed: Optional[int] = None,
               *, record: Optional[np.ndarray] = None, sampler_kind: str = "host", chunk: int = 1 << 20,
               device: Optional[int] = None, details: bool = False, classes: bool = False) -> dict:
    """Batched equivalent of simulator.py:167-315; returns the same six-key dict.

    record       : optional bool array (shots, m_z + m_x + 2n) = [sy_z | sy_x | errX | errZ]; when given it replaces
                   the sampler (this is how parity with the reference is checked: same record on both sides).
    sampler_kind : 'host'  -- NumPy sampler of sampler.py (bit-identical to the oracle's batches, seeded by rngSeed)
                   'device'-- Philox sampler on the GPU keyed by the global shot index (for 10^6+ shots)
    Unlike the reference (whose rngSeed never reaches Stim, simulator.py:187-188) runs are reproducible.
    With torch.distributed initialised each rank processes its shard_range and counters are all-reduced.
    classes      : also count the four outcome classes of README.md:15-22 (exact / degenerate / logical error / decoder
                   failure) through the code's logical operators; adds the keys outcome_* to the returned dict.
    """
    pipe = Pipeline(Hx, Hz, p, decType, decIterations, decSchedule, OSDorder, device, logicals=classes)
    t = pipe.torch
    rank, world = dist_info()
    lo, hi = shard_range(shots, rank, world)
    seed = DEFAULT_SEED if rngSeed is None else int(rngSeed)
    counters = t.zeros(_lib.NUM_COUNTERS, dtype=t.int64, device=pipe.device)
    kept = []
    if record is not None:
        rec = np.asarray(record)
        if rec.shape[0] != shots:
            raise ValueError("record must hold one row per shot")
        for s0 in range(lo, hi, chunk):
            s1 = min(hi, s0 + chunk)
            pipe.run(*pipe.upload_record(rec[s0:s1]), counters=counters, keep=details)
            if details:
                kept.append(pipe.last)
    elif sampler_kind == "host":
        # the host stream is defined over the whole batch; every rank draws it and keeps its slice
        rec = sampler.sample_record(pipe.Hx, pipe.Hz, p, shots, seed=seed)
        for s0 in range(lo, hi, chunk):
            s1 = min(hi, s0 + chunk)
            pipe.run(*pipe.upload_record(rec[s0:s1]), counters=counters, keep=details)
            if details:
                kept.append(pipe.last)
    elif sampler_kind == "device":
        for s0 in range(lo, hi, chunk):
            s1 = min(hi, s0 + chunk)
            pipe.run(*pipe.sample_device(s1 - s0, seed, first_shot=s0), counters=counters, keep=details)
            if details:
                kept.append(pipe.last)
    else:
        raise ValueError("sampler_kind must be 'host' or 'device'")
    reduce_counters(counters)
    t.cuda.synchronize(pipe.device)
    res = counters_to_result(counters.cpu().numpy(), shots, classes=classes)
    if details:
        res["_details"] = {k: t.cat([d[k] for d in kept]).cpu().numpy() for k in kept[0]} if kept else {}
        res["_counters"] = counters.cpu().numpy()
    return res


def format_results(p, results, shots: int) -> str:
    """The result table of simulator.py:342-347."""
    lines = ['\n                             ===          SIMULATION RESULTS          ===\n',
             '   Depolarizing probability | qBlock error rate | Decoding failures (X,Z) | Average iterations (X,Z)',
             '----------------------------+-------------------+-------------------------+---------------------------']
    for pT, r in zip(p, results):
        qbler = 1. - (r['decSuccessExact'] + r['decSuccessDegen']) / shots
        lines.append(f"         {pT:10.2e}         |     {qbler:7.2e}      |       {r['DecFailures_X']:5},{r['DecFailures_Z']:5}       "
                     f"|      {r['Avg_number_of_iterations_X']:5.2f}, {r['Avg_number_of_iterations_Z']:5.2f}")
    if results and "outcome_exact" in results[0]:
        lines += ['', '   Outcome classes (README: perfect match / degenerate / logical error / decoder failure)',
                  '   Depolarizing probability |   exact   | degenerate | logical error | decoder failure | logical+failure rate',
                  '----------------------------+-----------+------------+---------------+-----------------+----------------------']
        for pT, r in zip(p, results):
            bad = (r['outcome_logical_error'] + r['outcome_decoder_failure']) / shots
            lines.append(f"         {pT:10.2e}         | {r['outcome_exact']:9} | {r['outcome_degenerate']:10} | {r['outcome_logical_error']:13} "
                         f"| {r['outcome_decoder_failure']:15} |       {bad:7.2e}")
    return "\n".join(lines)


def simulate(HxFile: str, HzFile: str, p, shots: int = 1000, decType: str = 'MS', decIterations: int = 99,
             decSchedule: str = 'F', OSDorder: int = -1, rngSeed: Optional[int] = None, **kw):
    """Same signature and printed table as simulator.py:319-347 (returns None; extra keyword arguments go to simulate_p)."""
    Hx = load_matrix(HxFile)
    Hz = load_matrix(HzFile)
    assert max(p) <= 1. and min(p) >= 0.
    results = [simulate_p(Hx, Hz, p=pT, shots=shots, rngSeed=rngSeed, decType=decType, decIterations=decIterations,
                          decSchedule=decSchedule, OSDorder=OSDorder, **kw) for pT in p]
    if dist_info()[0] == 0:
        print(format_results(p, results, shots))


def main(argv=None):
    """CLI with the reference's flags (simulator.py:351-373)."""
    parser = argparse.ArgumentParser(description="B200 batched QC-LDPC depolarizing-channel simulator.")
    parser.add_argument("--Hx", required=True, help="Path to Hx parity-check matrix (.npy).")
    parser.add_argument("--Hz", required=True, help="Path to Hz parity-check matrix (.npy).")
    parser.add_argument("--p", type=float, nargs='+', required=True, help="Depolarizing probability.")
    parser.add_argument("--shots", type=int, default=1000, help="Number of Monte Carlo shots.")
    parser.add_argument("--rngSeed", type=int, default=None, help="RNG seed.")
    parser.add_argument("--decType", choices=['NG', 'BF', 'MS', 'BP'], default='MS',
                        help="Decoder type: [NG] Naive Greedy; [MS] Min-Sum; [BP] Belief Propagation.")
    parser.add_argument("--decIterations", type=int, default=99, help="Number of decoding iterations.")
    parser.add_argument("--decSchedule", choices=['F', 'L', 'S'], default='F',
                        help="Decoder scheduling method: [F] flooding; [L] layered; [S] serial.")
    parser.add_argument("--OSDorder", type=int, default=-1, help="Ordered Statistics Decoding order.")
    parser.add_argument("--sampler", choices=['host', 'device'], default='host', help="Where shots are drawn.")
    parser.add_argument("--classes", action="store_true", help="Also count exact / degenerate / logical-error / failure outcomes.")
    args = parser.parse_args(argv)
    print('\n   Command line arguments:')
    print(args)
    print('')
    simulate(HxFile=args.Hx, HzFile=args.Hz, p=args.p, shots=args.shots, decType=args.decType,
             decIterations=args.decIterations, decSchedule=args.decSchedule, OSDorder=args.OSDorder,
             rngSeed=args.rngSeed, sampler_kind=args.sampler, classes=args.classes)


if __name__ == "__main__":
    main()
