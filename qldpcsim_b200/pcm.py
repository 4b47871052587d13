"""
PCM compiler (host side): dense parity-check matrix -> what the CUDA library's plan builder consumes.

 * `load_matrix`     : the reference's input format (simulator.py:20-35): .npy or whitespace text, reduced
                       mod 2, int8.
 * `compile_pcm`     : CSR edge list in np.where(H) order (ascending check, then ascending variable --
                       decoders.py:224), CSC, degree statistics.
 * `layerize`        : the check partition of the layered / serial schedules (simulator.py:212-224).
 * `schedule_layers` : (layersX, layersZ) as the driver builds them (simulator.py:228-236).
 * `logical_operators`: bases of the logical X / Z operators of the CSS code (what the true outcome classes of
                       README.md:15-22 need; the reference never computes them).
The device-side tables (message layout, check-slot assignment, per-layer variable groups, runs of layers with disjoint
variable sets merged into one step) are derived from the CSR + layers by qldpc_plan_create in csrc/ (see include/qldpc_b200.h).

Quasi-cyclic structure (PCMlibrary.py:129-138, 195-201) is exploited STRUCTURALLY by the plan builder -- the stable
descending-degree renumbering keeps circulant blocks contiguous (bank-conflict-free lane groups), and the single-check layers of
a circulant block row are what the merged steps of the serial schedule collapse (450 -> 16 steps on LP118_2) -- but NOT by
computing edge addresses from (block, shift) on the device: DESIGN.md section 9 records why that trade loses (it adds integer
instructions to an issue-bound kernel and frees too little shared memory for one more resident shot on any code of the library).
A circulant detector that existed for that purpose was removed; `pcmlibrary.qc_base` still exposes (Bx, Bz, L) of the generators.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


def load_matrix(path: str) -> np.ndarray:
    """Binary matrix from .npy or whitespace-separated 0/1 text; returns (mat % 2) as int8 (simulator.py:20-35)."""
    if path.endswith(".npy"):
        mat = np.load(path)
    else:
        rows = []
        with open(path, "rt") as f:
            for line in f:
                line = line.strip()
                if line:
                    rows.append([int(tok) for tok in line.split()])
        mat = np.array(rows, dtype=int)
    return (mat % 2).astype(np.int8)


@dataclass
class CompiledPCM:
    H: np.ndarray           # int8 dense 0/1
    m: int
    n: int
    nnz: int
    row_ptr: np.ndarray     # int32 (m+1)
    col_idx: np.ndarray     # int32 (nnz)   CSR
    col_ptr: np.ndarray     # int32 (n+1)
    row_idx: np.ndarray     # int32 (nnz)   CSC, ascending check per variable
    row_weight_max: int
    col_weight_max: int
    _cache: dict = field(default_factory=dict, repr=False)


def compile_pcm(H: np.ndarray, find_qc: bool = False) -> CompiledPCM:
    """`find_qc` is accepted for compatibility and ignored (see the module docstring)."""
    H = (np.asarray(H) % 2).astype(np.int8)
    if H.ndim != 2:
        raise ValueError("parity-check matrix must be 2-D")
    m, n = H.shape
    chk, var = np.nonzero(H)                                 # row-major: the reference's edge order
    nnz = int(chk.size)
    row_ptr = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(np.bincount(chk, minlength=m), out=row_ptr[1:])
    order = np.lexsort((chk, var))                           # by variable, then check
    col_ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(var, minlength=n), out=col_ptr[1:])
    rw = np.diff(row_ptr)
    cw = np.diff(col_ptr)
    return CompiledPCM(H=H, m=m, n=n, nnz=nnz, row_ptr=row_ptr, col_idx=var.astype(np.int32),
                       col_ptr=col_ptr, row_idx=chk[order].astype(np.int32),
                       row_weight_max=int(rw.max(initial=0)), col_weight_max=int(cw.max(initial=0)))


def layerize(H: np.ndarray, serial: bool = False) -> List[np.ndarray]:
    """Greedy partition of the checks into runs of consecutive rows with pairwise disjoint supports
    (single rows when `serial`) -- the layers of simulator.py:212-224.

    A run is extended row by row; the row that would give some column a second 1 starts the next run."""
    H = np.asarray(H)
    m = H.shape[0]
    support = H != 0
    cuts = [0]
    used = np.zeros(H.shape[1], dtype=bool)
    for i in range(m):
        row = support[i]
        if i > cuts[-1] and (serial or (used & row).any()):
            cuts.append(i)
            used = row.copy()
        else:
            used |= row
    cuts.append(m)
    return [np.arange(a, b) for a, b in zip(cuts[:-1], cuts[1:])]


def schedule_layers(Hx: np.ndarray, Hz: np.ndarray, decSchedule: str) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """(layersX, layersZ) as simulator.py:228-236: 'F' one layer with every check, 'L' layerize, 'S' serial."""
    if decSchedule == "F":
        return [np.arange(Hx.shape[0])], [np.arange(Hz.shape[0])]
    if decSchedule in ("L", "S"):
        s = decSchedule == "S"
        return layerize(Hx, serial=s), layerize(Hz, serial=s)
    raise ValueError("Unrecognized decoder scheduling option.")


def flatten_layers(layers: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    ptr = np.zeros(len(layers) + 1, dtype=np.int32)
    if len(layers):
        ptr[1:] = np.cumsum([len(l) for l in layers])
        idx = np.concatenate([np.asarray(l, dtype=np.int32).ravel() for l in layers]).astype(np.int32)
    else:
        idx = np.zeros(0, dtype=np.int32)
    return ptr, np.ascontiguousarray(idx)


# ---------------------------------------------------------------------------------------------------------
# logical operators (extension: the reference's README defines the outcome classes, its code cannot separate them)
# ---------------------------------------------------------------------------------------------------------
def _rows_to_ints(M: np.ndarray) -> List[int]:
    M = (np.asarray(M) % 2).astype(np.uint8)
    return [int.from_bytes(np.packbits(r, bitorder="little").tobytes(), "little") for r in M]


def _ints_to_rows(vals: Sequence[int], n: int) -> np.ndarray:
    out = np.zeros((len(vals), n), dtype=np.int8)
    nbytes = (n + 7) // 8
    for i, v in enumerate(vals):
        out[i] = np.unpackbits(np.frombuffer(v.to_bytes(nbytes, "little"), dtype=np.uint8), bitorder="little")[:n]
    return out


class _GF2Span:
    """Incremental row space over GF(2) on Python integers (bit j = column j): reduced basis keyed by leading bit."""

    def __init__(self):
        self.piv = {}

    def reduce(self, v: int) -> int:
        while v:
            h = v.bit_length() - 1
            b = self.piv.get(h)
            if b is None:
                return v
            v ^= b
        return 0

    def add(self, v: int) -> bool:
        v = self.reduce(v)
        if not v:
            return False
        self.piv[v.bit_length() - 1] = v
        return True


def gf2_nullspace(M: np.ndarray) -> List[int]:
    """Basis of {v : M v = 0 (mod 2)} as integers (same space as gf2math.nullSpace, gf2math.py:12-50; own algorithm:
    eliminate [M^T | I] and read the identity part of the rows whose M^T part vanished)."""
    M = (np.asarray(M) % 2).astype(np.uint8)
    m, n = M.shape
    cols = _rows_to_ints(M.T)                     # n integers of m bits
    rows = [(c << n) | (1 << j) for j, c in enumerate(cols)]    # high part: column of M, low part: e_j
    piv = {}
    null = []
    for v in rows:
        while v >> n:
            h = v.bit_length() - 1
            b = piv.get(h)
            if b is None:
                piv[h] = v
                v = 0
                break
            v ^= b
        if v:
            null.append(v & ((1 << n) - 1))
    return null


def logical_operators(Hx: np.ndarray, Hz: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(Lx, Lz): k x n bases of the logical X operators ker(Hz)/rowspace(Hx) and the logical Z operators
    ker(Hx)/rowspace(Hz), k = n - rank(Hx) - rank(Hz).

    A residual X error d with Hz d = 0 is a stabiliser iff Lz d = 0; a residual Z error d with Hx d = 0 iff Lx d = 0."""
    Hx = (np.asarray(Hx) % 2).astype(np.uint8)
    Hz = (np.asarray(Hz) % 2).astype(np.uint8)
    n = Hx.shape[1]
    out = []
    for stab, other in ((Hx, Hz), (Hz, Hx)):          # Lx: in ker(Hz), independent of rowspace(Hx); then Lz
        span = _GF2Span()
        for v in _rows_to_ints(stab):
            span.add(v)
        logical = [v for v in gf2_nullspace(other) if span.add(v)]
        out.append(_ints_to_rows(logical, n))
    return out[0], out[1]
