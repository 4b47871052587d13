"""The oracle's restatement of NumPy's float64 tanh / arctanh (oracle/npymath.h) is bit-identical to NumPy: against values
NumPy produced in the build container (tests/golden/numpy_tanh_arctanh.npz) and, where the running NumPy takes the same
dispatch path (AVX512_SKX), against NumPy live on fresh arguments.  The sum-product decoder of the reference calls these two
functions per edge (decoders.py:254-259); with them the oracle -- and the CUDA kernel, which uses the same algorithms and
tables -- reproduces the reference bit for bit."""
import os

import numpy as np

from conftest import GOLDEN_DIR
from oracle import oracle


def _same(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))


def test_against_numpy_golden():
    g = np.load(os.path.join(GOLDEN_DIR, "numpy_tanh_arctanh.npz"))
    assert _same(oracle.npy_tanh(g["x_tanh"]), g["y_tanh"]).all()
    assert _same(oracle.npy_arctanh(g["x_arctanh"]), g["y_arctanh"]).all()
    # the fixture is not vacuous: libm differs from NumPy on a good part of it
    import math
    diff_t = sum(math.tanh(x) != y for x, y in zip(g["x_tanh"][:4000], g["y_tanh"][:4000]))
    assert diff_t > 400


def test_against_live_numpy_when_same_dispatch():
    feat = np._core._multiarray_umath.__cpu_features__
    rng = np.random.default_rng(5)
    xt = np.concatenate([rng.uniform(-30, 30, 400000), np.sign(rng.uniform(-1, 1, 400000)) * 10.0 ** rng.uniform(-300, 2, 400000)])
    if feat.get("AVX512_SKX") or feat.get("AVX2"):            # np.tanh: NumPy's own kernel on both dispatch targets (FMA)
        assert _same(oracle.npy_tanh(xt), np.tanh(xt)).all()
    if feat.get("AVX512_SKX"):                                  # np.arctanh: SVML only with AVX-512
        xa = np.concatenate([rng.uniform(-1, 1, 400000), np.sign(rng.uniform(-1, 1, 400000)) * (1 - 10.0 ** rng.uniform(-16, 0, 400000))])
        assert _same(oracle.npy_arctanh(xa), np.arctanh(xa)).all()


def test_tables_are_the_same_in_oracle_and_product():
    """The CUDA kernel (qldpcsim_b200/csrc/npymath.cuh) and the oracle include separate copies of the table file."""
    root = os.path.dirname(GOLDEN_DIR.rstrip("/"))
    root = os.path.dirname(root)
    a = open(os.path.join(root, "oracle", "npymath_tables.inc")).read()
    b = open(os.path.join(root, "qldpcsim_b200", "csrc", "npymath_tables.inc")).read()
    assert a == b
