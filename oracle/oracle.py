"""
CPU oracle for the qLDPCsim decoder path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/qldpc_oracle.c plus NumPy restatements of the driver-side logic of the
reference (layer partition, X/Z wiring, outcome counters).  Nothing under qldpcsim_b200/ may import
this module; it is the checker used by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs.

Parity status: PINNED against outputs of the unmodified reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.

Reference citations are relative to the upstream repository (albertogp71/qLDPCsim v0.2.2).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqldpc_oracle.so")
_lib = None

DEC_TYPES = {"NG": 0, "BF": 1, "MS": 2, "BP": 3}


def build(force: bool = False) -> str:
    """Compile qldpc_oracle.c with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("qldpc_oracle.c", "npymath.h", "npymath_tables.inc")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "libqldpc_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_double
    lib.orc_graph_create.restype = vp
    lib.orc_graph_create.argtypes = [vp, i32, i32]
    lib.orc_graph_destroy.argtypes = [vp]
    lib.orc_graph_edges.argtypes = [vp]
    lib.orc_decode_batch.restype = i32
    lib.orc_decode_batch.argtypes = [vp, i32, vp, i64, f64, i32, vp, vp, i32, f64, f64, i32, vp, vp, vp, vp, vp, i32]
    lib.orc_osd.restype = i32
    lib.orc_osd.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.orc_osd_reliability.argtypes = [vp, i32, vp]
    lib.orc_npy_tanh.argtypes = [vp, vp, i64]
    lib.orc_npy_arctanh.argtypes = [vp, vp, i64]
    lib.orc_rank.restype = i32
    lib.orc_rank.argtypes = [vp, i32, i32]
    _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------------------
# Driver-side restatements
# ---------------------------------------------------------------------------------------------
def layerize(H: np.ndarray, serial: bool = False) -> List[np.ndarray]:
    """Greedy contiguous layer partition -- restates simulator.py:212-224.

    Grow a window of consecutive checks while no column of the window has weight > 1 (and, for the
    serial schedule, while it holds a single check); close the layer just before the offending row.
    """
    H = np.asarray(H)
    m = H.shape[0]
    layers = []
    lo = 0
    colw = np.zeros(H.shape[1], dtype=np.int64)
    for hi in range(m):
        nxt = colw + (H[hi] != 0)
        if (nxt.max(initial=0) > 1) or (serial and hi > lo):
            layers.append(np.arange(lo, hi))
            lo = hi
            colw = (H[hi] != 0).astype(np.int64)
        else:
            colw = nxt
    layers.append(np.arange(lo, m))
    return layers


def schedule_layers(Hx: np.ndarray, Hz: np.ndarray, decSchedule: str):
    """(layersX, layersZ) exactly as simulator.py:228-236 builds them."""
    if decSchedule == "F":
        return [np.arange(Hx.shape[0])], [np.arange(Hz.shape[0])]
    if decSchedule in ("L", "S"):
        return layerize(Hx, serial=decSchedule == "S"), layerize(Hz, serial=decSchedule == "S")
    raise ValueError("Unrecognized decoder scheduling option.")


def prior_llr(p: float, eps: float = 1e-9) -> float:
    """decoders.py:147 / :232 -- float64 log computed with NumPy."""
    return float(np.log((1 - p) / max(p, eps)))


def osd_perm_numpy(posteriorLLRs: np.ndarray) -> np.ndarray:
    """The permutation of decoders.py:320-325, computed with NumPy itself (unstable argsort)."""
    sat = np.where(np.abs(posteriorLLRs) < 100.0, posteriorLLRs, 100.0 * np.sign(posteriorLLRs))
    prob = 1.0 / (1.0 + np.exp(sat))
    rel = np.where(prob > 0.5, prob, 1 - prob)
    return np.argsort(rel).astype(np.int32), rel


# ---------------------------------------------------------------------------------------------
# Decoder oracle
# ---------------------------------------------------------------------------------------------
class Graph:
    """Tanner graph of one parity-check matrix (dense 0/1 array)."""

    def __init__(self, H: np.ndarray):
        lib = _load()
        self.H = np.ascontiguousarray((np.asarray(H) % 2).astype(np.uint8))
        self.m, self.n = self.H.shape
        self._g = lib.orc_graph_create(_ptr(self.H), self.m, self.n)
        self.E = lib.orc_graph_edges(self._g)

    def __del__(self):
        try:
            if self._g:
                _load().orc_graph_destroy(self._g)
                self._g = None
        except Exception:
            pass

    def decode(self, decType: str, syndromes: np.ndarray, *, p: float = 0.0, max_iter: int = 50,
               layers: Optional[Sequence[np.ndarray]] = None, beta: float = 0.75, eps: float = 1e-9,
               OSDorder: int = -1, osd_perm: Optional[np.ndarray] = None, want_posterior: bool = False,
               n_threads: int = 0):
        """Decode a batch.  syndromes: (shots, m) 0/1.  Returns dict(e_hat, iters, converged[, posterior]).

        Call conventions follow simulator.py:270-282: NG takes nothing, BF takes max_iter (the driver
        leaves it at 50), MS/BP take p (already divided by 3 by the caller), max_iter and layers.
        """
        lib = _load()
        syn = np.ascontiguousarray(np.asarray(syndromes).astype(np.uint8) & 1)
        if syn.ndim == 1:
            syn = syn[None, :]
        shots = syn.shape[0]
        assert syn.shape[1] == self.m
        if layers is None:
            layers = [np.arange(self.m)]
        lptr = np.zeros(len(layers) + 1, dtype=np.int32)
        lptr[1:] = np.cumsum([len(l) for l in layers])
        lidx = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int32) for l in layers])
                                    if len(layers) else np.zeros(0, np.int32)).astype(np.int32)
        e_out = np.zeros((shots, self.n), dtype=np.uint8)
        iters = np.zeros(shots, dtype=np.int32)
        conv = np.zeros(shots, dtype=np.uint8)
        post = np.zeros((shots, self.n), dtype=np.float64) if want_posterior else None
        perm = None
        if osd_perm is not None:
            perm = np.ascontiguousarray(np.asarray(osd_perm, dtype=np.int32).reshape(shots, self.n))
        L = prior_llr(p, eps) if decType in ("MS", "BP") else 0.0
        rc = lib.orc_decode_batch(self._g, DEC_TYPES[decType], _ptr(syn), shots, L, int(max_iter), _ptr(lptr),
                                  _ptr(lidx), len(layers), float(beta), float(eps), int(OSDorder), _ptr(perm),
                                  _ptr(e_out), _ptr(iters), _ptr(conv), _ptr(post), int(n_threads))
        if rc != 0:
            raise RuntimeError(f"oracle decode failed rc={rc}")
        out = {"e_hat": e_out, "iters": iters, "converged": conv.astype(bool)}
        if want_posterior:
            out["posterior"] = post
        return out

    def osd(self, e_hat: np.ndarray, syndrome: np.ndarray, posteriorLLRs: np.ndarray, order: int = 0,
            perm: Optional[np.ndarray] = None) -> np.ndarray:
        """decoders.py:299-370 for one shot.  Returns a new uint8 vector."""
        lib = _load()
        e = np.ascontiguousarray(np.asarray(e_hat).astype(np.uint8) & 1).copy()
        s = np.ascontiguousarray(np.asarray(syndrome).astype(np.uint8) & 1)
        llr = np.ascontiguousarray(np.asarray(posteriorLLRs, dtype=np.float64))
        pp = None if perm is None else np.ascontiguousarray(np.asarray(perm, dtype=np.int32))
        rc = lib.orc_osd(self._g, _ptr(e), _ptr(s), _ptr(llr), int(order), _ptr(pp))
        if rc != 0:
            raise RuntimeError(f"oracle OSD failed rc={rc}")
        return e


def npy_tanh(x: np.ndarray) -> np.ndarray:
    """np.tanh as NumPy 2.3.5 evaluates it for float64 (npymath.h: simd_tanh_f64 restated)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _load().orc_npy_tanh(_ptr(x), _ptr(y), x.size)
    return y


def npy_arctanh(x: np.ndarray) -> np.ndarray:
    """np.arctanh as NumPy 2.3.5 evaluates it for float64 on AVX-512 hosts (npymath.h: __svml_atanh8_ha restated)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _load().orc_npy_arctanh(_ptr(x), _ptr(y), x.size)
    return y


def gf2_rank(A: np.ndarray) -> int:
    """gf2math.py:91-135."""
    A = np.ascontiguousarray((np.asarray(A) % 2).astype(np.uint8))
    return int(_load().orc_rank(_ptr(A), A.shape[0], A.shape[1]))


def split_record(record: np.ndarray, m_z: int, m_x: int, n: int):
    """simulator.py:249-252 -- record columns are [sy_z | sy_x | errX | errZ]."""
    r = np.asarray(record)
    return (r[:, :m_z], r[:, m_z:m_z + m_x], r[:, m_z + m_x:m_z + m_x + n], r[:, m_z + m_x + n:m_z + m_x + 2 * n])


def simulate_p(Hx: np.ndarray, Hz: np.ndarray, record: np.ndarray, p: float, decType: str = "MS",
               decIterations: int = 99, decSchedule: str = "F", OSDorder: int = -1, n_threads: int = 0,
               details: bool = False) -> dict:
    """Restates the shot loop of simulator.py:244-315 on a caller-supplied measurement record.

    Wiring (simulator.py:270-282): the X-error estimate comes from (Hz, sy_z) and uses layersX (built
    from Hx!); the Z-error estimate from (Hx, sy_x) with layersZ; MS/BP get p/3 and decIterations, MS
    also OSDorder, BP never; NG/BF get neither (BF therefore always runs max_iter=50).
    """
    Hx = (np.asarray(Hx) % 2).astype(np.int8)
    Hz = (np.asarray(Hz) % 2).astype(np.int8)
    m_x, n = Hx.shape
    m_z = Hz.shape[0]
    sy_z, sy_x, errX, errZ = split_record(record, m_z, m_x, n)
    shots = record.shape[0]
    layersX, layersZ = schedule_layers(Hx, Hz, decSchedule)
    gz, gx = Graph(Hz), Graph(Hx)
    if decType == "NG":
        dx = gz.decode("NG", sy_z, n_threads=n_threads)
        dz = gx.decode("NG", sy_x, n_threads=n_threads)
    elif decType == "BF":
        dx = gz.decode("BF", sy_z, max_iter=50, n_threads=n_threads)
        dz = gx.decode("BF", sy_x, max_iter=50, n_threads=n_threads)
    elif decType == "MS":
        dx = gz.decode("MS", sy_z, p=p / 3, max_iter=decIterations, layers=layersX, OSDorder=OSDorder, n_threads=n_threads)
        dz = gx.decode("MS", sy_x, p=p / 3, max_iter=decIterations, layers=layersZ, OSDorder=OSDorder, n_threads=n_threads)
    elif decType == "BP":
        dx = gz.decode("BP", sy_z, p=p / 3, max_iter=decIterations, layers=layersX, n_threads=n_threads)
        dz = gx.decode("BP", sy_x, p=p / 3, max_iter=decIterations, layers=layersZ, n_threads=n_threads)
    else:
        raise ValueError("Unrecognized decoder type.")
    eX, eZ = dx["e_hat"], dz["e_hat"]
    exact = np.all(eX == errX, axis=1) & np.all(eZ == errZ, axis=1)                       # :294-295
    dX = (errX.astype(np.int64) ^ eX)
    dZ = (errZ.astype(np.int64) ^ eZ)

    def imat(a, Ht):
        # integer matrix product through BLAS: the entries are counts <= n < 2^24, exactly representable in binary32
        return np.rint(a.astype(np.float32) @ Ht.astype(np.float32)).astype(np.int64)
    degen = (~exact) & np.all(imat(dX, Hz.T) == 0, axis=1) & np.all(imat(dZ, Hx.T) == 0, axis=1)   # :296-298 (no mod 2)
    failX = np.any(imat(eX, Hz.T) % 2 != sy_z, axis=1)                                    # :300-301
    failZ = np.any(imat(eZ, Hx.T) % 2 != sy_x, axis=1)                                    # :302-303
    res = {
        "DecFailures_X": int(failX.sum()),
        "DecFailures_Z": int(failZ.sum()),
        "decSuccessExact": int(exact.sum()),
        "decSuccessDegen": int(degen.sum()),
        "Avg_number_of_iterations_X": float(dx["iters"].sum()) / float(shots),
        "Avg_number_of_iterations_Z": float(dz["iters"].sum()) / float(shots),
    }
    if details:
        res["_details"] = {"eX": eX, "eZ": eZ, "itX": dx["iters"], "itZ": dz["iters"], "convX": dx["converged"],
                           "convZ": dz["converged"], "exact": exact, "failX": failX, "failZ": failZ}
    return res


def outcome_classes(Hx: np.ndarray, Hz: np.ndarray, errX, errZ, eX, eZ) -> dict:
    """The four outcome classes of the reference's README.md:15-22 (its code does not implement them: simulator.py:296-298
    is vacuous).  Checker for the library's extension counters; deliberately a DIFFERENT method from the library's
    (which tests the residual against logical operators): here a residual with matching syndromes is a stabiliser iff
    appending it to the stabiliser matrix does not raise the GF(2) rank (gf2math.rank semantics, gf2math.py:91-135)."""
    Hx = (np.asarray(Hx) % 2).astype(np.uint8)
    Hz = (np.asarray(Hz) % 2).astype(np.uint8)
    dX = (np.asarray(errX).astype(np.uint8) ^ np.asarray(eX).astype(np.uint8)) & 1
    dZ = (np.asarray(errZ).astype(np.uint8) ^ np.asarray(eZ).astype(np.uint8)) & 1
    rx, rz = gf2_rank(Hx), gf2_rank(Hz)
    out = {"exact": 0, "degenerate": 0, "logical_error": 0, "decoder_failure": 0}
    fail = ((dX.astype(np.int64) @ Hz.T.astype(np.int64)) % 2).any(axis=1) | ((dZ.astype(np.int64) @ Hx.T.astype(np.int64)) % 2).any(axis=1)
    for s in range(dX.shape[0]):
        if fail[s]:
            out["decoder_failure"] += 1
        elif not dX[s].any() and not dZ[s].any():
            out["exact"] += 1
        else:
            in_x = (not dX[s].any()) or gf2_rank(np.vstack([Hx, dX[s:s + 1]])) == rx      # X residual in rowspace(Hx)
            in_z = (not dZ[s].any()) or gf2_rank(np.vstack([Hz, dZ[s:s + 1]])) == rz
            out["degenerate" if (in_x and in_z) else "logical_error"] += 1
    return out
