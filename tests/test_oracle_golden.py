"""Pins the CPU oracle (oracle/qldpc_oracle.c) on the golden outputs of the unmodified reference."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import oracle


def _run(g, H, sy, layers, **extra):
    dt = g["decType"]
    gr = oracle.Graph(H)
    if dt == "NG":
        return gr.decode("NG", sy)
    if dt == "BF":
        return gr.decode("BF", sy, max_iter=50)
    kw = dict(p=float(g["p"]) / 3, max_iter=int(g["decIterations"]), layers=layers)
    kw.update(extra)
    return gr.decode(dt, sy, **kw)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    lX, lZ = oracle.schedule_layers(g["Hx"], g["Hz"], g["sched"])
    osd = int(g["OSDorder"])
    for which, (H, sy, lay, e_ref, it_ref) in enumerate(((g["Hz"], g["sy_z"], lX, g["eX_ref"], g["itX"]),
                                                         (g["Hx"], g["sy_x"], lZ, g["eZ_ref"], g["itZ"]))):
        extra = {}
        if osd >= 0:
            # reproduce the reference's (unstable) np.argsort order from the stored permutations
            perm = np.tile(np.arange(g["n"], dtype=np.int32), (len(sy), 1))
            sel = g["osd_which"] == which
            perm[g["osd_shot"][sel]] = g["osd_perm"][sel]
            extra = dict(OSDorder=osd, osd_perm=perm, want_posterior=True)
        o = _run(g, H, sy, lay, **extra)
        same = (o["e_hat"] == e_ref).all(axis=1) & (o["iters"] == it_ref)
        # every decoder bit-exact, sum-product included (np.tanh / np.arctanh restated in oracle/npymath.h)
        assert same.all(), f"{name}: oracle differs from the reference on shots {np.nonzero(~same)[0][:8]}"
        if osd >= 0:
            sel = np.nonzero(g["osd_which"] == which)[0]
            assert np.array_equal(o["posterior"][g["osd_shot"][sel]], g["osd_llr"][sel]), "posterior must be bit-identical"


@pytest.mark.parametrize("name", [n for n in golden_names() if "OSD" in n])
def test_oracle_osd_stable_order_on_tie_free_shots(name):
    g = load_golden(name)
    n = g["n"]
    from qldpcsim_b200 import bitpack
    for k in range(len(g["osd_shot"])):
        which, s = int(g["osd_which"][k]), int(g["osd_shot"][k])
        H = g["Hz"] if which == 0 else g["Hx"]
        sy = (g["sy_z"] if which == 0 else g["sy_x"])[s]
        ref = (g["eX_ref"] if which == 0 else g["eZ_ref"])[s]
        perm_np, rel = oracle.osd_perm_numpy(g["osd_llr"][k])
        e_in = bitpack.unpack_rows(g["osd_e_in"][k:k + 1], n)[0]
        out = oracle.Graph(H).osd(e_in, sy, g["osd_llr"][k], int(g["OSDorder"]))
        assert ((out.astype(np.int64) @ H.T.astype(np.int64)) % 2 == sy).all()
        if len(np.unique(rel)) == n:
            assert np.array_equal(out, ref)


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith(("steane_MS_F", "LP04_0_MS_L_p0", "LP04_0_NG", "LP04_0_BF",
                                                                              "steane_BF", "steane_NG", "shor", "LP118_0_MS_L_p05"))])
def test_oracle_counters(name):
    """The restated shot loop (classification, counters, X/Z wiring) against the reference's own simulate_p."""
    g = load_golden(name)
    if "counters" not in g:
        pytest.skip("no simulate_p counters stored")
    r = oracle.simulate_p(g["Hx"], g["Hz"], g["rec"], float(g["p"]), decType=g["decType"],
                          decIterations=int(g["decIterations"]), decSchedule=g["sched"], OSDorder=int(g["OSDorder"]))
    shots = int(g["shots"])
    got = [r["DecFailures_X"], r["DecFailures_Z"], r["decSuccessExact"], r["decSuccessDegen"],
           round(r["Avg_number_of_iterations_X"] * shots), round(r["Avg_number_of_iterations_Z"] * shots)]
    assert got == list(g["counters"])


def test_survey_golden_table():
    """SURVEY.md section 4 table (produced independently in the survey session) -- same sampler, same counters."""
    want = {"steane_MS_F_p01": (0, 0, 988, 0, 1024, 1024), "steane_MS_F_p02": (0, 0, 981, 0, 1039, 1039),
            "steane_MS_F_p05": (0, 0, 942, 0, 1099, 1095), "steane_MS_F_p10": (0, 0, 839, 0, 1163, 1141),
            "steane_BP_F_p05": (0, 0, 942, 0, 1099, 1095), "steane_BF_p05": (23, 24, 803, 0, 6194, 6145),
            "steane_NG_p05": (0, 0, 973, 0, 228, 224), "LP04_0_MS_L_p02": (0, 0, 200, 0, 244, 237),
            "LP04_0_MS_L_p05": (4, 3, 188, 0, 618, 554), "LP04_0_NG_p02": (7, 3, 90, 0, 2690, 1297),
            "LP04_0_BF_p02": (59, 59, 20, 0, 3334, 3089)}
    for name, c in want.items():
        assert tuple(load_golden(name)["counters"]) == c, name


@pytest.mark.parametrize("tag", ["02", "05", "10"])
def test_bp_big_reference_goldens_bit_exact(tag):
    """1600 LP118_0 BP-F decodes of the unmodified reference per depolarizing probability (the ends and the middle of BASELINE
    config 2's sweep; tests/golden/make_bp_golden.py).  With NumPy's own tanh / arctanh restated (oracle/npymath.h) the oracle
    is bit-identical on every one of them -- including the 964 decodes at p = 0.10 that never converge.  (With glibc's
    functions, round 1, the match was 99.69 % at p = 0.05 and 99.81 % at p = 0.10.)"""
    import os
    from conftest import GOLDEN_DIR
    from qldpcsim_b200 import bitpack, pcm, pcmlibrary
    g = np.load(os.path.join(GOLDEN_DIR, f"big_LP118_0_BP_F_p{tag}_X.npz"))
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name("LP118_0")]
    m, n = Hz.shape
    syn = bitpack.unpack_rows(g["syn"], m).astype(np.uint8)
    e_ref, it_ref, iters = bitpack.unpack_rows(g["e"], n), g["it"], int(g["decIterations"])
    lX, _ = pcm.schedule_layers(Hx, Hz, "F")
    o = oracle.Graph(Hz).decode("BP", syn, p=float(g["p"]) / 3, max_iter=iters, layers=lX)
    same = (o["e_hat"] == e_ref).all(1) & (o["iters"] == it_ref)
    assert same.all(), f"p=0.{tag}: {int((~same).sum())} of {len(same)} decodes differ from the reference"
