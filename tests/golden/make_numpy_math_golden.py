#!/usr/bin/env python
"""Golden values of np.tanh / np.arctanh (float64) as evaluated by the NumPy of the build container (2.3.5, x86-64, AVX512_SKX
dispatch: its own SIMD tanh kernel and Intel SVML's atanh8_ha) -- the two transcendental functions of the reference's
sum-product decoder (decoders.py:254-259).  oracle/npymath.h and qldpcsim_b200/csrc/npymath.cuh restate those routines; this
fixture pins the restatement on hosts whose NumPy dispatches differently (no AVX-512: np.arctanh falls back to libm there).

    python tests/golden/make_numpy_math_golden.py
"""
import os

import numpy as np

feat = np._core._multiarray_umath.__cpu_features__
assert feat.get("AVX512_SKX"), "generate on a host where NumPy dispatches to AVX512_SKX (the reference goldens were made on one)"
rng = np.random.default_rng(20261018)
N = 6000
xt = np.concatenate([
    rng.uniform(-25, 25, N), np.sign(rng.uniform(-1, 1, N)) * 10.0 ** rng.uniform(-12, 1.5, N),
    np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-310, -1e-310, 5e-324, 0.125, 0.1875, 0.25, 24.0, 23.999, 1e300, 19.0, 22.0, 30.0]),
    np.ldexp(1.0 + np.arange(32) / 32.0, rng.integers(-4, 5, 32)).astype(np.float64),
])
xa = np.concatenate([
    rng.uniform(-1, 1, N), np.sign(rng.uniform(-1, 1, N)) * (1 - 10.0 ** rng.uniform(-16, 0, N)),
    np.sign(rng.uniform(-1, 1, N // 2)) * 10.0 ** rng.uniform(-300, 0, N // 2),
    np.array([0.0, -0.0, 1.0, -1.0, 1.5, -2.0, np.nan, np.inf, 1 - 1e-9, -(1 - 1e-9), 1 - 2.0 ** -53, 0.5, -0.5, 2.0 ** -1022, 5e-324]),
])
with np.errstate(all="ignore"):
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "numpy_tanh_arctanh.npz"),
                        numpy_version=np.__version__, x_tanh=xt, y_tanh=np.tanh(xt), x_arctanh=xa, y_arctanh=np.arctanh(xa))
print("wrote", len(xt), len(xa))
