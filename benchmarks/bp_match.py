#!/usr/bin/env python
"""Per-shot match rate of the GPU sum-product decoder against the goldens produced by the unmodified reference
(tests/golden/*_BP_*.npz) and against the CPU oracle on larger seeded batches; the bar is 99.9 % of shots."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import golden_names, load_golden  # noqa: E402
from oracle import oracle  # noqa: E402
from qldpcsim_b200 import bitpack, pcmlibrary, sampler, simulator  # noqa: E402

for name in golden_names():
    if "_BP_" not in name:
        continue
    g = load_golden(name)
    r = simulator.simulate_p(g["Hx"], g["Hz"], float(g["p"]), shots=int(g["shots"]), decType="BP", decIterations=int(g["decIterations"]),
                             decSchedule=str(g["sched"]), record=g["rec"], details=True)
    n = g["n"]
    eX = bitpack.unpack_rows(r["_details"]["eX"].view(np.uint32), n)
    eZ = bitpack.unpack_rows(r["_details"]["eZ"].view(np.uint32), n)
    mx = ((eX == g["eX_ref"]).all(1) & (r["_details"]["itX"] == g["itX"])).mean()
    mz = ((eZ == g["eZ_ref"]).all(1) & (r["_details"]["itZ"] == g["itZ"])).mean()
    print(f"golden {name:22s} shots {int(g['shots']):5d}  match X {mx:.5f}  Z {mz:.5f}", flush=True)
for code, sched, p, shots, iters in (("LP118_0", "F", 0.05, 20000, 100), ("LP118_0", "F", 0.10, 4000, 100), ("LP04_0", "L", 0.08, 20000, 30),
                                     ("bicycle", "F", 0.03, 10000, 50)):
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=777)
    want = oracle.simulate_p(Hx, Hz, rec, p, decType="BP", decIterations=iters, decSchedule=sched, details=True)["_details"]
    got = simulator.simulate_p(Hx, Hz, p, shots=shots, decType="BP", decIterations=iters, decSchedule=sched, record=rec, details=True)["_details"]
    n = Hx.shape[1]
    eX = bitpack.unpack_rows(got["eX"].view(np.uint32), n)
    eZ = bitpack.unpack_rows(got["eZ"].view(np.uint32), n)
    mx = ((eX == want["eX"]).all(1) & (got["itX"] == want["itX"])).mean()
    mz = ((eZ == want["eZ"]).all(1) & (got["itZ"] == want["itZ"])).mean()
    print(f"oracle {code:8s} BP-{sched} p={p} shots {shots:6d}  match X {mx:.5f}  Z {mz:.5f}", flush=True)
