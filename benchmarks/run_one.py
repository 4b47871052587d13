#!/usr/bin/env python
"""One configuration, one warm-up pass and one timed pass (sample on the device, decode X, decode Z, classify): the command the
per-kernel ncu captures of profiles/capture_kernels.sh wrap.

    python benchmarks/run_one.py --code LP118_0 --dec BP --sched F --p 0.05 --iters 100 --shots 200000
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from qldpcsim_b200 import _lib, pcmlibrary, simulator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="LP118_0")
    ap.add_argument("--dec", default="MS")
    ap.add_argument("--sched", default="L")
    ap.add_argument("--p", type=float, default=0.05)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--osd", type=int, default=-1)
    ap.add_argument("--shots", type=int, default=1_000_000)
    ap.add_argument("--kernel", default="auto")
    a = ap.parse_args()
    Hx, Hz = pcmlibrary.by_name(a.code)
    pipe = simulator.Pipeline(Hx, Hz, a.p, a.dec, a.iters, a.sched, a.osd, kernel=a.kernel)
    inp = pipe.sample_device(a.shots, 11, 0)
    pipe.run(*inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = pipe.run(*inp).cpu().numpy()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"code": a.code, "dec": a.dec, "sched": a.sched, "p": a.p, "iters": a.iters, "osd": a.osd, "shots": a.shots,
                      "ms": ms, "shots_per_s": a.shots / ms * 1e3, "iters_X": int(c[_lib.CNT_ITERS_X]), "iters_Z": int(c[_lib.CNT_ITERS_Z]),
                      "fail_X": int(c[0]), "fail_Z": int(c[1]), "info": pipe.decX.info()}))


if __name__ == "__main__":
    main()
