#!/usr/bin/env python
"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): every kernel family once on small batches.

    compute-sanitizer --tool memcheck  python benchmarks/sanitize_case.py
    compute-sanitizer --tool racecheck python benchmarks/sanitize_case.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from qldpcsim_b200 import pcmlibrary, sampler, simulator  # noqa: E402

CASES = [("LP118_0", "MS", "L", 0.08, 300, -1), ("LP04_0", "MS", "L", 0.10, 300, 0), ("LP118_0", "BP", "F", 0.06, 120, -1),
         ("LP04_0", "MS", "S", 0.06, 300, -1), ("bicycle", "MS", "L", 0.03, 200, -1), ("LP04_0", "NG", "F", 0.03, 300, -1),
         ("LP04_0", "BF", "F", 0.03, 300, -1), ("steane", "MS", "F", 0.1, 300, 1)]
for code, dt, sched, p, shots, osd in CASES:
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=5)
    r = simulator.simulate_p(Hx, Hz, p, shots=shots, decType=dt, decIterations=12, decSchedule=sched, OSDorder=osd, record=rec,
                             classes=True)
    print(code, dt, sched, {k: v for k, v in r.items() if not k.startswith("Avg")}, flush=True)
r = simulator.simulate_p(Hx, Hz, 0.05, shots=500, decType="MS", decIterations=10, decSchedule="F", sampler_kind="device", rngSeed=3)
print("device sampler", r["decSuccessExact"])
