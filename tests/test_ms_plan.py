"""Host-side layout planner of the min-sum kernel (qldpcsim_b200/csrc/ms_plan.h), exercised on the CPU through the small C
wrapper csrc/ms_plan_test.cpp (compiled here with g++): the renumbering is a permutation, every edge of every check gets
exactly one slot, every layer's variable groups cover its variables exactly once, and the modelled shared-memory wavefront
count stays close to the conflict-free bound on the layered lifted-product configurations of BASELINE.json."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from qldpcsim_b200 import pcm, pcmlibrary

SRC = os.path.join(ROOT, "qldpcsim_b200", "csrc", "ms_plan_test.cpp")


@pytest.fixture(scope="module")
def planner(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("msplan") / "libmsplan.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, SRC], check=True)
    return ctypes.CDLL(so)


def kernel_shape(H):
    dc, cw = int(H.sum(1).max()), H.sum(0)
    dv = int(cw.max())
    dci = [s for s in (4, 8, 18, 32) if s >= dc][0]
    dvi = [s for s in (4, 5, 9, 16) if s >= dv][0]
    fast = 3 if dvi <= 5 else (dvi if dvi <= 9 else 0)
    dmin = fast if (fast > 0 and int(cw.min()) >= fast) else 0
    return dci, dvi, dmin


def run_planner(lib, H, layers, search=1, packed=0):
    c = pcm.compile_pcm(H, find_qc=False)
    lp, lc = pcm.flatten_layers(layers)
    m, n = H.shape
    dci, dvi, dmin = kernel_shape(H)
    perm = np.zeros(n, np.int32)
    slot_edge = np.zeros(m * dci, np.int32)
    lvar_ptr = np.zeros(len(layers) + 1, np.int32)
    lvar = np.zeros(65536, np.uint32)
    stats = np.zeros(6, np.int64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.ms_plan_probe(m, n, c.nnz, P(c.row_ptr), P(c.col_idx), len(layers), P(lp), P(lc), dci, dvi, dmin, search,
                           P(perm), P(slot_edge), P(lvar_ptr), P(lvar), lvar.size, P(stats), packed)
    assert rc == 0
    return c, dci, perm, slot_edge.reshape(m, dci), lvar_ptr, lvar[:stats[4]], stats


CASES = [("steane", "F"), ("steane", "L"), ("shor", "S"), ("LP04_0", "L"), ("LP04_0", "F"), ("LP118_0", "L"), ("LP118_0", "S"),
         ("LP118_2", "L"), ("T", "L"), ("bicycle", "L")]


@pytest.mark.parametrize("code,sched", CASES)
def test_layout_is_consistent(planner, code, sched):
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    if code == "shor":
        pairs = [(Hz, pcm.layerize(Hz, serial=True)), (Hx, pcm.layerize(Hx, serial=True))]
    else:
        lX, lZ = pcm.schedule_layers(Hx, Hz, sched)
        pairs = [(Hz, lX), (Hx, lZ)]            # the reference's cross-wiring (simulator.py:278-282)
    for H, layers in pairs:
        c, dci, perm, slot_edge, lvar_ptr, lvar, stats = run_planner(planner, H, layers)
        n = H.shape[1]
        assert sorted(perm.tolist()) == list(range(n))
        cw = H.sum(0)
        assert (np.diff(cw[np.argsort(perm)]) <= 0).all(), "renumbering must be by descending column weight"
        for i in range(H.shape[0]):
            e = slot_edge[i][slot_edge[i] >= 0]
            assert sorted(e.tolist()) == list(range(c.row_ptr[i], c.row_ptr[i + 1])), "every edge of a check exactly once"
        for l, checks in enumerate(layers):
            ent = lvar[lvar_ptr[l]:lvar_ptr[l + 1]]
            assert ent.size % 32 == 0
            js = np.concatenate([(ent & 0xFFFF) // 4, (ent >> 16) // 4])
            js = js[js != n]
            want = np.unique(perm[np.nonzero(H[checks].any(0))[0]])
            assert sorted(js.tolist()) == want.tolist(), "variable groups cover the layer's variables exactly once"
        assert stats[0] >= stats[1] > 0 and stats[0] <= stats[5]


@pytest.mark.parametrize("code,bound", [("LP118_0", 1.12), ("LP118_2", 1.10), ("T", 1.10)])
def test_layered_lifted_codes_are_nearly_conflict_free(planner, code, bound):
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    lX, lZ = pcm.schedule_layers(Hx, Hz, "L")
    for H, layers in ((Hz, lX), (Hx, lZ)):
        stats = run_planner(planner, H, layers)[-1]
        assert stats[0] <= bound * stats[1], (code, stats[0], stats[1])


def test_random_irregular_graph(planner):
    rng = np.random.default_rng(7)
    H = (rng.random((40, 90)) < 0.06).astype(np.int8)
    H[:, H.sum(0) == 0] = 0
    H[0, :3] = 1
    layers = [np.arange(0, 13), np.arange(13, 14), np.arange(14, 40)]
    c, dci, perm, slot_edge, lvar_ptr, lvar, stats = run_planner(planner, H, layers)
    assert sorted(perm.tolist()) == list(range(90))
    for i in range(40):
        e = slot_edge[i][slot_edge[i] >= 0]
        assert sorted(e.tolist()) == list(range(c.row_ptr[i], c.row_ptr[i + 1]))


@pytest.mark.parametrize("code", ["bicycle", "LP04_0", "steane"])
def test_eight_lane_layout(planner, code):
    """Layout of the eight-lane kernel (ms_sub_kernel.cuh; every layer a single check): the searched renumbering stays a
    permutation by descending column weight, every edge of a check sits in exactly one (trip, lane) cell, and on the bicycle
    code (18 variables per check, three trips of eight) the modelled shared-memory wavefronts come within 10 % of the
    conflict-free bound (identity numbering: 1.55 x)."""
    import time
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    for H in (Hz, Hx):
        c = pcm.compile_pcm(H, find_qc=False)
        m, n = H.shape
        dcs = max(8, (int(H.sum(1).max()) + 7) & ~7)
        dv = int(H.sum(0).max())
        dvi, dmin = (9, 9) if (dcs == 24 and dv == 9 and int(H.sum(0).min()) == 9) else (16, 0)
        perm = np.zeros(n, np.int32)
        cell = np.zeros(m * dcs, np.int32)
        stats = np.zeros(4, np.int64)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        t0 = time.time()
        assert planner.ms_plan_probe_sub8(m, n, c.nnz, P(c.row_ptr), P(c.col_idx), dcs, dvi, dmin, 1, P(perm), P(cell), P(stats)) == 0
        assert time.time() - t0 < 5.0
        assert sorted(perm.tolist()) == list(range(n))
        cw = H.sum(0)
        assert (np.diff(cw[np.argsort(perm)]) <= 0).all()
        cell = cell.reshape(m, dcs)
        for i in range(m):
            e = cell[i][cell[i] >= 0]
            assert sorted(e.tolist()) == list(range(c.row_ptr[i], c.row_ptr[i + 1]))
        assert stats[1] <= stats[0] <= stats[3]
        if code == "bicycle":
            assert stats[3] > 1.4 * stats[1] and stats[0] <= 1.10 * stats[1], stats


@pytest.mark.parametrize("code,sched", [("LP118_0", "L"), ("LP118_0", "F"), ("LP04_0", "L"), ("T", "L")])
def test_packed_regions(planner, code, sched):
    """Packed layout (regions of the c2v array on multiples of 16 instead of 32 words: the 21st resident shot of LP118_0): the
    array shrinks or stays, every edge keeps exactly one slot, and the modelled shared-memory wavefronts -- the planner's bank
    model follows the per-region rotation -- stay within 10 % of the aligned layout's."""
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    lX, lZ = pcm.schedule_layers(Hx, Hz, sched)
    for H, layers in ((Hz, lX), (Hx, lZ)):
        c, dci, perm0, se0, _, _, st0 = run_planner(planner, H, layers, 1, 0)
        c, dci, perm1, se1, _, _, st1 = run_planner(planner, H, layers, 1, 1)
        assert st1[3] <= st0[3] and st1[3] % 4 == 0
        assert st1[0] <= 1.10 * st0[0]
        assert sorted(perm1.tolist()) == list(range(H.shape[1]))
        for i in range(H.shape[0]):
            e = se1[i][se1[i] >= 0]
            assert sorted(e.tolist()) == list(range(c.row_ptr[i], c.row_ptr[i + 1]))
        if code == "LP118_0":
            assert st0[3] == 2048 and st1[3] == 1968
