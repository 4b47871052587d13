"""The C-ABI shared library loads and exports every symbol include/qldpc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def test_header_symbols_exported():
    from qldpcsim_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "qldpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(qldpc_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in the header but not exported"
    assert sorted(declared) == sorted(_lib.EXPORTS)
    lib = _lib.lib()
    assert lib.qldpc_abi_version() == _lib.ABI_VERSION == 2
    assert lib.qldpc_words(1) == 1 and lib.qldpc_words(32) == 1 and lib.qldpc_words(33) == 2 and lib.qldpc_words(544) == 17
    assert lib.qldpc_launch_count() == 0


def test_struct_layout_matches_header():
    from qldpcsim_b200 import _lib
    assert ctypes.sizeof(_lib.Opts) == 40            # 2 x int32, 3 x double, 2 x int32
    assert _lib.Graph.row_ptr.offset == 16 and _lib.Graph.layer_ptr.offset == 40


def test_plan_create_argument_errors_do_not_need_a_gpu():
    """Invalid arguments are rejected before any CUDA call."""
    import numpy as np
    from qldpcsim_b200 import _lib
    lib = _lib.lib()
    h = ctypes.c_void_p()
    rp = np.array([0, 2], np.int32)
    ci = np.array([1, 0], np.int32)                  # not ascending
    g = _lib.Graph(m=1, n=2, nnz=2, row_ptr=rp.ctypes.data, col_idx=ci.ctypes.data, n_layers=0, layer_ptr=None, layer_chk=None)
    o = _lib.Opts(dec_type=7, max_iter=1, prior_llr=1.0, beta=0.75, eps=1e-9, osd_order=-1, reserved=0)
    assert lib.qldpc_plan_create(ctypes.byref(g), ctypes.byref(o), 0, ctypes.byref(h)) == -1
    assert b"decoder type" in lib.qldpc_last_error()
    o.dec_type = _lib.MS
    assert lib.qldpc_plan_create(ctypes.byref(g), ctypes.byref(o), 0, ctypes.byref(h)) == -1     # MS without layers
    assert b"layer" in lib.qldpc_last_error()
    o.dec_type = _lib.NG
    assert lib.qldpc_plan_create(ctypes.byref(g), ctypes.byref(o), 0, ctypes.byref(h)) == -1     # bad CSR
    assert b"ascending" in lib.qldpc_last_error()
