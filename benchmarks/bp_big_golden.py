#!/usr/bin/env python
"""GPU sum-product decoder against the 1600-decode reference golden (tests/golden/big_LP118_0_BP_F_p05_X.npz)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import oracle  # noqa: E402
from qldpcsim_b200 import bitpack, pcm, pcmlibrary  # noqa: E402
from qldpcsim_b200.decoders import Decoder  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "big_LP118_0_BP_F_p05_X.npz"))
Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name("LP118_0")]
m, n = Hz.shape
syn = bitpack.unpack_rows(g["syn"], m).astype(np.uint8)
e_ref = bitpack.unpack_rows(g["e"], n)
it_ref = g["it"]
lX, _ = pcm.schedule_layers(Hx, Hz, "F")
out = Decoder(Hz, "BP", p=float(g["p"]) / 3, max_iter=int(g["decIterations"]), layers=lX).decode(syn)
same = (out["e_hat"] == e_ref).all(1) & (out["iters"] == it_ref)
unconv = it_ref == int(g["decIterations"])
fast = it_ref <= 20
print(f"GPU vs reference: match {same.mean():.5f} ({(~same).sum()} mismatches, {((~same) & unconv).sum()} on reference-unconverged decodes); "
      f"decodes converging within 20 iterations: {same[fast].mean():.5f} of {fast.sum()}")
o = oracle.Graph(Hz).decode("BP", syn, p=float(g["p"]) / 3, max_iter=int(g["decIterations"]), layers=lX)
so = (o["e_hat"] == e_ref).all(1) & (o["iters"] == it_ref)
print(f"oracle (glibc) vs reference: match {so.mean():.5f} ({(~so).sum()} mismatches)")
fail_ref = ((e_ref.astype(np.int64) @ Hz.T.astype(np.int64)) % 2 != syn).any(1).mean()
fail_gpu = ((out["e_hat"].astype(np.int64) @ Hz.T.astype(np.int64)) % 2 != syn).any(1).mean()
print(f"failure rate reference {fail_ref:.5f}  GPU {fail_gpu:.5f}")
