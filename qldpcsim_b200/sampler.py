"""
Host-side depolarizing sampler: the stand-in for the reference's Stim circuit + sampler
(simulator.py:43-160, 196-197).

The Stim circuit of the reference prepares a codeword, applies PAULI_CHANNEL_1(p/3, p/3, p/3) to every data
qubit (simulator.py:107) and reads out, per shot, the Z-check syndrome of the X-component, the X-check
syndrome of the Z-component, and the two error components themselves, in the record order
[sy_z | sy_x | errX | errZ] (simulator.py:141-144, parsed at :249-252).  That record is a deterministic
function of i.i.d. per-qubit Paulis, which is what this module draws (definition fixed by SURVEY.md section 8d so
that the oracle and the GPU decode identical batches):

    u = default_rng(seed).random((shots, n));  X = u < p/3;  Y = p/3 <= u < 2p/3;  Z = 2p/3 <= u < p
    errX = X | Y;  errZ = Z | Y;  sy_z = Hz errX mod 2;  sy_x = Hx errZ mod 2
"""
from __future__ import annotations

import numpy as np


def sample_errors(n: int, p: float, shots: int, seed=None, chunk: int = 1 << 16):
    """(errX, errZ) as uint8 arrays (shots, n).  Row-chunked; the stream equals one big rng.random((shots, n))."""
    rng = np.random.default_rng(seed)
    errX = np.empty((shots, n), dtype=np.uint8)
    errZ = np.empty((shots, n), dtype=np.uint8)
    for s0 in range(0, shots, chunk):
        s1 = min(shots, s0 + chunk)
        u = rng.random((s1 - s0, n))
        x = u < p / 3
        y = (p / 3 <= u) & (u < 2 * p / 3)
        z = (2 * p / 3 <= u) & (u < p)
        errX[s0:s1] = x | y
        errZ[s0:s1] = z | y
    return errX, errZ


def syndromes(H: np.ndarray, err: np.ndarray) -> np.ndarray:
    """H err^T mod 2 per shot, uint8 (shots, m)."""
    Hf = (np.asarray(H) % 2).astype(np.float32)
    out = np.empty((err.shape[0], Hf.shape[0]), dtype=np.uint8)
    step = 1 << 15
    for s0 in range(0, err.shape[0], step):         # float32 matmul is exact here (row weights << 2^24)
        blk = err[s0:s0 + step].astype(np.float32) @ Hf.T
        out[s0:s0 + step] = blk.astype(np.int64) & 1
    return out


def sample_record(Hx: np.ndarray, Hz: np.ndarray, p: float, shots: int, seed=None) -> np.ndarray:
    """bool record (shots, m_z + m_x + 2n) = [sy_z | sy_x | errX | errZ], the layout simulate_p parses."""
    n = Hx.shape[1]
    errX, errZ = sample_errors(n, p, shots, seed)
    sy_z = syndromes(Hz, errX)
    sy_x = syndromes(Hx, errZ)
    return np.concatenate([sy_z, sy_x, errX, errZ], axis=1).astype(bool)
