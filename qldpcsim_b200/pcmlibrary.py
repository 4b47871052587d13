"""
Library of CSS parity-check-matrix pairs (Hx, Hz) -- same generator names and outputs as the reference's
PCMlibrary (PCMlibrary.py:25-203), re-implemented from the code definitions.

Quasi-cyclic families are described by a base matrix of circulant shifts (-1 = empty block) and a lift L;
block (i, j) with shift s has its 1 of row r at column (r + s) mod L (PCMlibrary.py:129-138).  Besides the
dense matrices the module exposes the QC description itself (`qc_base`) so the PCM compiler does not have
to rediscover it.

Every generator returns dense int64 0/1 arrays (the reference's bicycle generator returns float64 with the
same values).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# Protograph shift tables of the lifted-product families (Quantum 6, 767 (2022)); data as listed in
# PCMlibrary.py:143-189.  Each entry: (lift L, 3 x n_b base matrix of shifts).
_LP_TABLE: Dict[Tuple[str, int], Tuple[int, Tuple[Tuple[int, ...], ...]]] = {
    ("LP04", 0): (7, ((0, 0, 0, 0), (0, 1, 2, 5), (0, 6, 3, 1))),
    ("LP04", 1): (9, ((0, 0, 0, 0), (0, 1, 6, 7), (0, 4, 5, 2))),
    ("LP04", 2): (17, ((0, 0, 0, 0), (0, 1, 2, 11), (0, 8, 12, 13))),
    ("LP04", 3): (19, ((0, 0, 0, 0), (0, 2, 6, 9), (0, 16, 7, 11))),
    ("LP118", 0): (16, ((0, 0, 0, 0, 0), (0, 2, 4, 7, 11), (0, 3, 10, 14, 15))),
    ("LP118", 1): (21, ((0, 0, 0, 0, 0), (0, 4, 5, 7, 17), (0, 14, 18, 12, 11))),
    ("LP118", 2): (30, ((0, 0, 0, 0, 0), (0, 2, 14, 24, 25), (0, 16, 11, 14, 13))),
}
# Tanner's (3,5) QC code, L = 31 (PCMlibrary.py:96-100)
_TANNER = (31, ((1, 2, 4, 8, 16), (5, 10, 20, 9, 18), (25, 19, 7, 14, 28)))


def lift(base: np.ndarray, L: int) -> np.ndarray:
    """Expand a matrix of circulant shifts (-1 = zero block) into a dense 0/1 matrix."""
    base = np.asarray(base, dtype=np.int64)
    mb, nb = base.shape
    H = np.zeros((mb * L, nb * L), dtype=np.int64)
    bi, bj = np.nonzero(base >= 0)
    r = np.arange(L)
    for i, j in zip(bi, bj):
        H[i * L + r, j * L + (r + base[i, j]) % L] = 1
    return H


def _hypergraph_product_bases(B: np.ndarray, L: int) -> Tuple[np.ndarray, np.ndarray]:
    """Base (shift) matrices of the lifted product of B with its conjugate transpose.

    Bx = [ B (x) I_nb | I_mb (x) B* ],  Bz = [ I_nb (x) B | B* (x) I_mb ],  B* = L - B^T
    (PCMlibrary.py:102-105, 191-194).  Shifts are kept exactly as the reference produces them (B* entries
    lie in 1..L; a shift of L is the identity).
    """
    B = np.asarray(B, dtype=np.int64)
    mb, nb = B.shape
    Bt = L - B.T

    def kron_shift(A, left_identity: int = 0, right_identity: int = 0):
        if right_identity:
            a_r, a_c = A.shape
            out = -np.ones((a_r * right_identity, a_c * right_identity), dtype=np.int64)
            for d in range(right_identity):
                out[d::right_identity, d::right_identity] = A
            return out
        a_r, a_c = A.shape
        out = -np.ones((a_r * left_identity, a_c * left_identity), dtype=np.int64)
        for d in range(left_identity):
            out[d * a_r:(d + 1) * a_r, d * a_c:(d + 1) * a_c] = A
        return out

    Bx = np.concatenate([kron_shift(B, right_identity=nb), kron_shift(Bt, left_identity=mb)], axis=1)
    Bz = np.concatenate([kron_shift(B, left_identity=nb), kron_shift(Bt, right_identity=mb)], axis=1)
    return Bx, Bz


def qc_base(name: str, index: int = 0):
    """QC description of a library code: (Bx, Bz, L) with shift matrices (-1 = empty block)."""
    if name in ("T", "tanner"):
        L, B = _TANNER
    else:
        if (name, index) not in _LP_TABLE:
            if name not in ("LP04", "LP118"):
                raise ValueError("qc_ldpc_lifted_codes: unrecognized code family.")
            raise ValueError(f"qc_ldpc_lifted_codes: index out of bounds for code family {name}.")
        L, B = _LP_TABLE[(name, index)]
    Bx, Bz = _hypergraph_product_bases(np.array(B), L)
    return Bx, Bz, L


def shor_code():
    """[[9,1,3]] Shor code (PCMlibrary.py:25-48): Hz = six ZZ pair checks, Hx = two weight-6 X checks."""
    Hz = np.zeros((6, 9), dtype=np.int64)
    for blk in range(3):
        for t in range(2):
            Hz[2 * blk + t, 3 * blk + t: 3 * blk + t + 2] = 1
    Hx = np.zeros((2, 9), dtype=np.int64)
    Hx[0, 0:6] = 1
    Hx[1, 3:9] = 1
    return Hx, Hz


def steane_code():
    """[[7,1,3]] Steane code (PCMlibrary.py:51-62): both matrices are the [7,4] Hamming check matrix."""
    H = np.array([[1, 0, 0, 1, 0, 1, 1],
                  [0, 1, 0, 1, 1, 0, 1],
                  [0, 0, 1, 0, 1, 1, 1]], dtype=np.int64)
    return H.copy(), H.copy()


def bicycle_code():
    """MacKay bicycle code from the size-73 perfect difference set (PCMlibrary.py:66-77): H = [C | C^T]."""
    N = 73
    support = np.array([2, 8, 15, 19, 20, 34, 42, 44, 72])
    C = np.zeros((N, N), dtype=np.int64)
    rows = np.arange(N)
    for s in support:
        C[rows, (rows + s) % N] = 1          # row i is the first row rolled right by i
    H0 = np.concatenate([C, C.T], axis=1)
    return H0.copy(), H0.copy()


def qc_ldpc_tanner_code():
    """Lifted product of Tanner's (3,5) QC-LDPC code, L = 31 (PCMlibrary.py:81-113)."""
    Bx, Bz, L = qc_base("T")
    return lift(Bx, L), lift(Bz, L)


def qc_ldpc_lifted_code(family: str = "LP04", index: int = 0):
    """Lifted-product codes LP04_{0..3}, LP118_{0..2} (PCMlibrary.py:119-203)."""
    Bx, Bz, L = qc_base(family, index)
    return lift(Bx, L), lift(Bz, L)


def by_name(name: str):
    """Convenience: 'steane', 'shor', 'bicycle', 'T', 'LP04_0' ... 'LP118_2' -> (Hx, Hz)."""
    if name == "steane":
        return steane_code()
    if name == "shor":
        return shor_code()
    if name == "bicycle":
        return bicycle_code()
    if name in ("T", "tanner"):
        return qc_ldpc_tanner_code()
    fam, idx = name.rsplit("_", 1)
    return qc_ldpc_lifted_code(fam, int(idx))
