"""
Decoders for quantum LDPC codes -- B200 drop-in for the reference's `decoders` module.

Two levels:

 * `Decoder`            : one device plan per (parity-check matrix, decoder configuration); decodes BATCHES of
                          syndromes with hand-written sm_100a kernels through the C ABI (include/qldpc_b200.h).
 * `NG_decoder`, `BF_decoder`, `MS_decoder`, `BP_decoder`, `OSDdec`
                        : the reference's per-shot functions with identical signatures, return types and dtypes
                          (decoders.py:27, :74, :110-117, :189-195, :299-304).  They run a batch of one through a
                          cached `Decoder`, so code written against the reference works unchanged.

There is no CPU implementation in this package: without the CUDA library and a GPU every call raises.
"""
from __future__ import annotations

import ctypes
import hashlib
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, bitpack
from .pcm import CompiledPCM, compile_pcm, flatten_layers


class LayersRequired(AttributeError, ValueError):
    """The reference's `layers=None` default dereferences np.range, which does not exist (decoders.py:144, :221):
    it raises AttributeError.  A layer list is therefore mandatory for MS/BP."""


def prior_llr(p: float, eps: float = 1e-9) -> float:
    """log((1-p)/max(p, eps)) evaluated in float64 with NumPy, as decoders.py:147 / :232 do."""
    with np.errstate(divide="ignore"):
        return float(np.log((1 - p) / max(p, eps)))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.QldpcError("no CUDA device: qldpcsim_b200 has no CPU fallback")
    return torch


class Decoder:
    """Batched syndrome decoder bound to one parity-check matrix and one configuration."""

    def __init__(self, H, decType: str, *, p: Optional[float] = None, max_iter: int = 50,
                 layers: Optional[Sequence[np.ndarray]] = None, beta: float = 0.75, OSDorder: int = -1,
                 eps: float = 1e-9, device: Optional[int] = None, kernel: str = "auto"):
        """kernel (MS only): 'auto' | 'plain' (warp per shot and layer per step: never merge runs of layers with disjoint
        variable sets into one step, never use the eight-lanes-per-shot kernel; for A/B measurements -- results are identical)."""
        if decType not in _lib.DEC_TYPES:
            raise ValueError("Unrecognized decoder type.")
        self.pcm: CompiledPCM = H if isinstance(H, CompiledPCM) else compile_pcm(H, find_qc=False)
        self.decType = decType
        self.m, self.n = self.pcm.m, self.pcm.n
        self.mw, self.nw = bitpack.words(self.m), bitpack.words(self.n)
        self.max_iter = int(max_iter)
        self.OSDorder = int(OSDorder)
        torch = _torch()
        self.device = torch.cuda.current_device() if device is None else int(device)
        iterative = decType in ("MS", "BP")
        if iterative:
            if layers is None:
                raise LayersRequired("module 'numpy' has no attribute 'range' -- pass `layers` (decoders.py:144)")
            if p is None:
                raise TypeError("MS/BP need the prior error probability p")
            lptr, lidx = flatten_layers(layers)
            if lidx.size and (lidx.min() < 0 or lidx.max() >= self.m):
                raise IndexError("layer holds a check index outside the matrix (reference: IndexError at decoders.py:156)")
        else:
            lptr, lidx = np.zeros(1, np.int32), np.zeros(0, np.int32)
        self._keep = (self.pcm.row_ptr, self.pcm.col_idx, lptr, lidx)
        g = _lib.Graph(m=self.m, n=self.n, nnz=self.pcm.nnz,
                       row_ptr=self.pcm.row_ptr.ctypes.data, col_idx=self.pcm.col_idx.ctypes.data,
                       n_layers=len(lptr) - 1 if iterative else 0, layer_ptr=lptr.ctypes.data, layer_chk=lidx.ctypes.data)
        self.prior = prior_llr(p, eps) if iterative else 0.0
        o = _lib.Opts(dec_type=_lib.DEC_TYPES[decType], max_iter=self.max_iter, prior_llr=self.prior, beta=float(beta),
                      eps=float(eps), osd_order=self.OSDorder, reserved={"auto": 0, "plain": 3}[kernel])
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().qldpc_plan_create(ctypes.byref(g), ctypes.byref(o), self.device, ctypes.byref(self._h)))

    # ------------------------------------------------------------------------------------------ lifetime
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.lib().qldpc_plan_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        L = _lib.lib()
        keys = ["m", "n", "nnz", "n_layers", "grid", "threads", "smem_bytes", "shots_per_cta", "row_weight_max",
                "col_weight_max", "rank", "reserved", "plan_wavefronts", "plan_wavefronts_ideal", "state_bytes", "logical_k",
                "steps_per_iteration", "warps_per_shot"]
        return {k: int(L.qldpc_plan_info(self._h, i)) for i, k in enumerate(keys)}

    def work_done(self, reset: bool = False) -> int:
        """Check-to-variable messages the min-sum kernel actually computed since plan creation / the last reset (synchronises)."""
        return int(_lib.lib().qldpc_plan_work(self._h, 1 if reset else 0))

    # ------------------------------------------------------------------------------------------ device API
    def decode_packed(self, syn_bits, *, want_converged: bool = True, want_llr: bool = False, out=None):
        """Device tensors in / out.  syn_bits: int32 CUDA tensor (shots, words(m)).
        Returns (ehat_bits int32 (shots, words(n)), iters int32 (shots,), converged uint8 or None, llr or None)."""
        torch = _torch()
        assert syn_bits.is_cuda and syn_bits.dtype == torch.int32 and syn_bits.is_contiguous()
        shots = syn_bits.shape[0]
        assert syn_bits.numel() == shots * self.mw
        dev = syn_bits.device
        if out is None:
            ehat = torch.empty((shots, self.nw), dtype=torch.int32, device=dev)
            iters = torch.empty((shots,), dtype=torch.int32, device=dev)
            conv = torch.empty((shots,), dtype=torch.uint8, device=dev) if want_converged else None
            llr = torch.empty((shots, self.n), dtype=torch.float64, device=dev) if want_llr else None
        else:
            ehat, iters, conv, llr = out
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().qldpc_decode(self._h, syn_bits.data_ptr(), shots, ehat.data_ptr(), iters.data_ptr(),
                                          conv.data_ptr() if conv is not None else None,
                                          llr.data_ptr() if llr is not None else None, stream))
        return ehat, iters, conv, llr

    def osd_packed(self, ehat_bits, syn_bits, llr, order: int = 0, perm=None):
        """In-place OSD on device tensors (decoders.py:299-370)."""
        torch = _torch()
        shots = ehat_bits.shape[0]
        stream = torch.cuda.current_stream(ehat_bits.device).cuda_stream
        _lib.check(_lib.lib().qldpc_osd(self._h, ehat_bits.data_ptr(), syn_bits.data_ptr(), llr.data_ptr(),
                                       perm.data_ptr() if perm is not None else None, shots, int(order), stream))
        return ehat_bits

    # ------------------------------------------------------------------------------------------ host API
    def decode_host_packed(self, syn_bits: np.ndarray, *, want_llr: bool = False, out=None):
        """Bit-packed host arrays in / out through qldpc_decode_host (copies overlapped with kernels)."""
        syn_bits = np.ascontiguousarray(syn_bits, dtype=np.uint32)
        shots = syn_bits.shape[0]
        if out is None:
            ehat = np.empty((shots, self.nw), dtype=np.uint32)
            iters = np.empty((shots,), dtype=np.int32)
            conv = np.empty((shots,), dtype=np.uint8)
            llr = np.empty((shots, self.n), dtype=np.float64) if want_llr else None
        else:
            ehat, iters, conv, llr = out
        _lib.check(_lib.lib().qldpc_decode_host(self._h, syn_bits.ctypes.data, shots, ehat.ctypes.data, iters.ctypes.data,
                                               conv.ctypes.data, llr.ctypes.data if llr is not None else None))
        return ehat, iters, conv, llr

    def decode(self, syndromes, *, want_llr: bool = False):
        """syndromes: (shots, m) 0/1 array.  Returns dict(e_hat uint8 (shots, n), iters int32, converged bool[, posterior])."""
        syn = np.asarray(syndromes)
        if syn.ndim == 1:
            syn = syn[None, :]
        if syn.shape[1] != self.m:
            raise ValueError(f"syndrome length {syn.shape[1]} != number of checks {self.m}")
        ehat, iters, conv, llr = self.decode_host_packed(bitpack.pack_rows(syn), want_llr=want_llr)
        out = {"e_hat": bitpack.unpack_rows(ehat, self.n), "iters": iters, "converged": conv.astype(bool)}
        if want_llr:
            out["posterior"] = llr
        return out


# ---------------------------------------------------------------------------------------------------------
# Reference-signature per-shot functions
# ---------------------------------------------------------------------------------------------------------
_PLAN_CACHE: dict = {}
_PLAN_CACHE_MAX = 32


def _cached_decoder(H: np.ndarray, decType: str, **cfg) -> Decoder:
    Hc = np.ascontiguousarray((np.asarray(H) % 2).astype(np.int8))
    layers = cfg.get("layers")
    lkey = None if layers is None else hashlib.sha1(b"|".join(np.asarray(l, dtype=np.int64).tobytes() for l in layers)).hexdigest()
    key = (hashlib.sha1(Hc.tobytes()).hexdigest(), Hc.shape, decType, lkey,
           tuple(sorted((k, v) for k, v in cfg.items() if k != "layers")))
    d = _PLAN_CACHE.get(key)
    if d is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).close()
        d = Decoder(Hc, decType, **cfg)
        _PLAN_CACHE[key] = d
    return d


def _empty(H, syndrome):
    H = np.asarray(H)
    return H.size == 0 or np.asarray(syndrome).size == 0


def NG_decoder(H: np.ndarray, syndrome: np.ndarray):
    """Naive greedy decoder (decoders.py:27-66).  Returns (est int8 (n,), steps)."""
    d = _cached_decoder(H, "NG")
    r = d.decode(syndrome)
    return r["e_hat"][0].astype(np.int8), int(r["iters"][0])


def BF_decoder(H: np.ndarray, syndrome: np.ndarray, max_iter: int = 50):
    """Bit-flipping decoder (decoders.py:74-102).  Returns (e_hat bool (n,), iterations)."""
    if _empty(H, syndrome):
        return np.zeros(np.asarray(H).shape[1] if np.asarray(H).size else 0, dtype=np.int8)     # decoders.py:86-87
    d = _cached_decoder(H, "BF", max_iter=max_iter)
    r = d.decode(syndrome)
    return r["e_hat"][0].astype(bool), int(r["iters"][0])


def MS_decoder(H: np.ndarray, syndrome: np.ndarray, p: float, max_iter: int = 99, layers: list = None,
               beta: float = 0.75, OSDorder: int = -1, eps: float = 1e-9):
    """Normalised min-sum with optional OSD (decoders.py:110-182).  Returns (e_hat int8 (n,), iterations)."""
    if _empty(H, syndrome):
        return np.zeros(np.asarray(H).shape[1] if np.asarray(H).size else 0, dtype=np.int8)     # decoders.py:138-139
    d = _cached_decoder(H, "MS", p=float(p), max_iter=int(max_iter), layers=layers, beta=float(beta),
                        OSDorder=int(OSDorder), eps=float(eps))
    r = d.decode(syndrome)
    return r["e_hat"][0].astype(np.int8), int(r["iters"][0])


def BP_decoder(H: np.ndarray, syndrome: np.ndarray, p: float, max_iter: int = 99, layers: list = None,
               OSDorder: int = -1, eps: float = 1e-9):
    """Sum-product BP with optional OSD (decoders.py:189-290).  Returns (e_hat int64 (n,), iterations)."""
    if _empty(H, syndrome):
        return np.zeros(np.asarray(H).shape[1] if np.asarray(H).size else 0, dtype=np.int8)     # decoders.py:215-216
    d = _cached_decoder(H, "BP", p=float(p), max_iter=int(max_iter), layers=layers, OSDorder=int(OSDorder), eps=float(eps))
    r = d.decode(syndrome)
    return r["e_hat"][0].astype(np.int64), int(r["iters"][0])


def OSDdec(H: np.ndarray, e_hat: np.ndarray, syndrome: np.ndarray, posteriorLLRs: np.ndarray, order: int = 0,
           perm: Optional[np.ndarray] = None) -> np.ndarray:
    """OSD post-decoding (decoders.py:299-370).  Mutates and returns `e_hat` like the reference (:368-370).

    `perm` (extension): column order to use instead of the library's stable sort by (reliability, index); the
    reference's own order comes from NumPy's unstable argsort and is platform-defined on ties."""
    torch = _torch()
    d = _cached_decoder(H, "NG")          # OSD only needs the matrix tables of a plan
    dev = torch.device("cuda", d.device)
    eb = torch.from_numpy(bitpack.pack_rows(np.asarray(e_hat)).view(np.int32)).to(dev)
    sb = torch.from_numpy(bitpack.pack_rows(np.asarray(syndrome)).view(np.int32)).to(dev)
    llr = torch.from_numpy(np.ascontiguousarray(np.asarray(posteriorLLRs, dtype=np.float64)).reshape(1, -1)).to(dev)
    pm = None if perm is None else torch.from_numpy(np.ascontiguousarray(np.asarray(perm, dtype=np.int32)).reshape(1, -1)).to(dev)
    d.osd_packed(eb, sb, llr, order, pm)
    torch.cuda.synchronize(dev)
    out = bitpack.unpack_rows(eb.cpu().numpy().view(np.uint32), d.n)[0]
    e_hat[...] = out.astype(e_hat.dtype)
    return e_hat
