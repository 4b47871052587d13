#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations (and the other decoders) on one GPU, device-resident inputs.

    python benchmarks/configs.py [--quick]        -> one JSON line per configuration

Each configuration is timed with CUDA events around Pipeline.run (decode X, decode Z, classify) after one warm-up
pass; inputs come from the on-device sampler.  This is a secondary harness: the headline number is bench.py's.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from qldpcsim_b200 import _lib, pcmlibrary, simulator  # noqa: E402

CONFIGS = [
    # name, code, decType, sched, p, iters, OSD, shots
    ("cfg0 Steane MS-F", "steane", "MS", "F", 0.05, 50, -1, 4_000_000),
    ("cfg1 LP04_0 MS-L", "LP04_0", "MS", "L", 0.05, 50, -1, 1_000_000),
    ("headline LP118_0 MS-L p=.01", "LP118_0", "MS", "L", 0.01, 50, -1, 1_000_000),
    ("headline LP118_0 MS-L p=.02", "LP118_0", "MS", "L", 0.02, 50, -1, 1_000_000),
    ("headline LP118_0 MS-L p=.05", "LP118_0", "MS", "L", 0.05, 50, -1, 1_000_000),
    ("headline LP118_0 MS-L p=.10", "LP118_0", "MS", "L", 0.10, 50, -1, 400_000),
    ("LP118_0 MS-F p=.05", "LP118_0", "MS", "F", 0.05, 50, -1, 1_000_000),
    ("LP118_2 MS-L p=.05", "LP118_2", "MS", "L", 0.05, 50, -1, 400_000),
    ("cfg2 LP118_0 BP-F 100it p=.02", "LP118_0", "BP", "F", 0.02, 100, -1, 400_000),
    ("cfg2 LP118_0 BP-F 100it p=.05", "LP118_0", "BP", "F", 0.05, 100, -1, 200_000),
    ("cfg2 LP118_0 BP-F 100it p=.10", "LP118_0", "BP", "F", 0.10, 100, -1, 40_000),
    ("cfg3 LP118_2 MS-S p=.05", "LP118_2", "MS", "S", 0.05, 50, -1, 1_000_000),
    ("cfg3 LP118_2 MS-S + OSD-10 p=.05", "LP118_2", "MS", "S", 0.05, 50, 10, 100_000),
    ("LP118_0 MS-L + OSD-0 p=.10", "LP118_0", "MS", "L", 0.10, 50, 0, 200_000),
    ("cfg4 Tanner MS-L p=.03", "T", "MS", "L", 0.03, 50, -1, 400_000),
    ("cfg4 bicycle MS-L p=.03", "bicycle", "MS", "L", 0.03, 50, -1, 1_000_000),
    ("cfg3 LP118_2 MS-S p=.05 (unmerged layers)", "LP118_2", "MS", "S", 0.05, 50, -1, 200_000, "plain"),
    ("LP118_0 MS-S p=.05", "LP118_0", "MS", "S", 0.05, 50, -1, 1_000_000),
    ("LP04_0 MS-S p=.05", "LP04_0", "MS", "S", 0.05, 50, -1, 1_000_000),
    ("Tanner MS-S p=.03", "T", "MS", "S", 0.03, 50, -1, 400_000),
    ("LP118_0 NG p=.02", "LP118_0", "NG", "F", 0.02, 50, -1, 400_000),
    ("LP118_0 BF p=.02", "LP118_0", "BF", "F", 0.02, 50, -1, 400_000),
]


def peak_gbs():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = peak_gbs()
    for cfg in CONFIGS:
        name, code, dt, sched, p, iters, osd, shots = cfg[:8]
        kernel = cfg[8] if len(cfg) > 8 else "auto"
        if a.only and a.only not in name:
            continue
        if a.quick:
            shots = max(10_000, shots // 10)
        Hx, Hz = pcmlibrary.by_name(code)
        pipe = simulator.Pipeline(Hx, Hz, p, dt, iters, sched, osd, kernel=kernel)
        inp = pipe.sample_device(shots, 11, 0)
        pipe.run(*inp)
        torch.cuda.synchronize(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        c = pipe.run(*inp, mid_event=ev[1])          # ev[1]: after the two decodes, before the classifier
        ev[2].record()
        torch.cuda.synchronize(dev)
        ms, ms_dec = ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[1])
        c = c.cpu().numpy()
        Ex, Ez, n = pipe.decX.pcm.nnz, pipe.decZ.pcm.nnz, pipe.n
        work = (c[_lib.CNT_ITERS_X] * Ex + c[_lib.CNT_ITERS_Z] * Ez)
        roof = None
        if dt in ("MS", "BP"):
            # 16 B per edge-iteration (layered / serial), 8 + 8 n/E (flooding), binary32 state; the
            # sum-product decoder keeps binary64 state (x2); + bit-packed I/O
            b_ei = (16.0 if sched != "F" else 8.0 + 8.0 * n / (0.5 * (Ex + Ez))) * (2.0 if dt == "BP" else 1.0)
            io = shots * 4 * (2 * bitpackwords(n) + bitpackwords(pipe.m_z) + bitpackwords(pipe.m_x) + 2)
            model = float(work) * b_ei + io
            roof = {"bound": "hbm", "bytes_per_edge_iteration": round(b_ei, 3), "model_bytes": model, "decode_ms": round(ms_dec, 3),
                    "achieved_gbs": round(model / (ms_dec * 1e-3) / 1e9, 1), "peak_gbs": peak, "frac": round(model / (ms_dec * 1e-3) / 1e9 / peak, 4)}
        ix = pipe.decX.info()
        print(json.dumps({"config": name, "shots": shots, "ms": round(ms, 3), "shots_per_s": round(shots / ms * 1e3, 1),
                          "avg_iters_X": round(c[_lib.CNT_ITERS_X] / shots, 3), "avg_iters_Z": round(c[_lib.CNT_ITERS_Z] / shots, 3),
                          "fail_X": int(c[0]), "fail_Z": int(c[1]), "exact": int(c[2]),
                          "edge_iterations_per_s": float(work) / (ms * 1e-3) if dt in ("MS", "BP") else None,
                          "shots_per_cta": ix["shots_per_cta"], "warps_per_shot": ix["warps_per_shot"],
                          "steps_per_iteration": ix["steps_per_iteration"], "n_layers": ix["n_layers"], "edges": Ex + Ez,
                          "roofline": roof}), flush=True)
        del pipe


def bitpackwords(nbits):
    return (nbits + 31) // 32


if __name__ == "__main__":
    main()
