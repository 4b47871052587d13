"""
GPU parity tests (run on the B200 box): the CUDA decoders, called through the C ABI, against
  (1) the golden outputs of the unmodified reference (tests/golden/*.npz), every case;
  (2) the CPU oracle on larger seeded batches;
  (3) size-independent properties at benchmark scale.
Bars (BASELINE.json north_star): NG / BF / OSD / outcome counters bit-exact; MS hard decisions and iteration
counts equal on >= 99.99 % of shots; BP on >= 99.9 % of shots.  What is asserted is stronger: MS and BP bit-exact too.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

MS_BAR = 0.9999
BP_BAR = 0.999


def _decoders_for(g, cuda_device):
    from qldpcsim_b200.decoders import Decoder
    from qldpcsim_b200.pcm import schedule_layers
    dt, it, osd = g["decType"], int(g["decIterations"]), int(g["OSDorder"])
    if dt in ("MS", "BP"):
        lX, lZ = schedule_layers(g["Hx"], g["Hz"], g["sched"])
        kw = dict(p=float(g["p"]) / 3, max_iter=it)
        if dt == "MS":
            kw["OSDorder"] = osd
        return Decoder(g["Hz"], dt, layers=lX, **kw), Decoder(g["Hx"], dt, layers=lZ, **kw)
    if dt == "BF":
        return Decoder(g["Hz"], "BF", max_iter=50), Decoder(g["Hx"], "BF", max_iter=50)
    return Decoder(g["Hz"], "NG"), Decoder(g["Hx"], "NG")


def _match_fraction(out, e_ref, it_ref):
    same = np.all(out["e_hat"] == e_ref, axis=1) & (out["iters"] == it_ref)
    return same.mean(), same


@pytest.mark.parametrize("name", [n for n in golden_names() if "OSD" not in n])
def test_golden_decoders(name, cuda_device):
    g = load_golden(name)
    dX, dZ = _decoders_for(g, cuda_device)
    oX = dX.decode(g["sy_z"])
    oZ = dZ.decode(g["sy_x"])
    fx, sx = _match_fraction(oX, g["eX_ref"], g["itX"])
    fz, sz = _match_fraction(oZ, g["eZ_ref"], g["itZ"])
    dt = g["decType"]
    if dt in ("NG", "BF"):
        assert fx == 1.0 and fz == 1.0, f"{name}: integer decoder must be bit-exact ({fx}, {fz})"
    elif dt == "MS":
        # the arithmetic is specified to the bit, so on these small sets every shot must agree
        assert fx == 1.0 and fz == 1.0, f"{name}: MS mismatch X {np.nonzero(~sx)[0][:5]} Z {np.nonzero(~sz)[0][:5]}"
    else:
        # np.tanh / np.arctanh are evaluated with NumPy's own algorithms: bit-exact like the other decoders
        assert fx == 1.0 and fz == 1.0, f"{name}: BP mismatch X {np.nonzero(~sx)[0][:5]} Z {np.nonzero(~sz)[0][:5]}"


@pytest.mark.parametrize("name", [n for n in golden_names() if "OSD" in n])
def test_golden_osd(name, cuda_device):
    """MS + OSD: bit-exact when the reference's own column order (np.argsort output, stored) is supplied;
    with the library's stable order every tie-free shot must still be bit-exact and every shot must satisfy
    its syndrome."""
    import torch
    from qldpcsim_b200 import bitpack
    from qldpcsim_b200.decoders import Decoder
    from qldpcsim_b200.pcm import schedule_layers
    g = load_golden(name)
    lX, lZ = schedule_layers(g["Hx"], g["Hz"], g["sched"])
    it, order, p = int(g["decIterations"]), int(g["OSDorder"]), float(g["p"])
    n = g["n"]
    for which, (H, sy, lay, e_ref, it_ref) in enumerate(((g["Hz"], g["sy_z"], lX, g["eX_ref"], g["itX"]),
                                                         (g["Hx"], g["sy_x"], lZ, g["eZ_ref"], g["itZ"]))):
        # (a) plain MS, posterior out, then OSD with the reference permutation
        d = Decoder(H, "MS", p=p / 3, max_iter=it, layers=lay, OSDorder=-1)
        o = d.decode(sy, want_llr=True)
        assert np.array_equal(o["iters"], it_ref)
        sel = np.nonzero(g["osd_which"] == which)[0]
        shots_osd = g["osd_shot"][sel]
        assert np.array_equal(np.sort(shots_osd), np.nonzero(~o["converged"])[0]), "unconverged set differs"
        # posterior handed to OSD must be bit-identical to the reference's (float64)
        assert np.array_equal(o["posterior"][shots_osd], g["osd_llr"][sel])
        e_in = bitpack.unpack_rows(g["osd_e_in"][sel], n)
        assert np.array_equal(o["e_hat"][shots_osd], e_in)
        dev = torch.device("cuda", d.device)
        eb = torch.from_numpy(bitpack.pack_rows(e_in).view(np.int32)).to(dev)
        sb = torch.from_numpy(bitpack.pack_rows(sy[shots_osd]).view(np.int32)).to(dev)
        llr = torch.from_numpy(np.ascontiguousarray(g["osd_llr"][sel])).to(dev)
        perm = torch.from_numpy(np.ascontiguousarray(g["osd_perm"][sel].astype(np.int32))).to(dev)
        d.osd_packed(eb, sb, llr, order, perm)
        got = bitpack.unpack_rows(eb.cpu().numpy().view(np.uint32), n)
        assert np.array_equal(got, e_ref[shots_osd]), f"{name}: OSD with reference perm not bit-exact"
        # (b) fused path, library's stable order
        d2 = Decoder(H, "MS", p=p / 3, max_iter=it, layers=lay, OSDorder=order)
        o2 = d2.decode(sy)
        assert np.array_equal(o2["iters"], it_ref)
        conv = o["converged"]
        assert np.array_equal(o2["e_hat"][conv], e_ref[conv])
        syn_ok = ((o2["e_hat"].astype(np.int64) @ H.T.astype(np.int64)) % 2 == sy).all(axis=1)
        assert syn_ok.all(), "OSD output must reproduce the syndrome"
        # tie-free shots: stable order == any order
        for k, s in zip(sel, shots_osd):
            sat = np.where(np.abs(g["osd_llr"][k]) < 100.0, g["osd_llr"][k], 100.0 * np.sign(g["osd_llr"][k]))
            prob = 1. / (1. + np.exp(sat))
            rel = np.where(prob > 0.5, prob, 1 - prob)
            if len(np.unique(rel)) == n:
                assert np.array_equal(o2["e_hat"][s], e_ref[s]), f"{name}: tie-free shot {s} differs"


@pytest.mark.parametrize("name", [n for n in golden_names()])
def test_golden_counters(name, cuda_device):
    """simulate_p on the stored record reproduces the reference's own simulate_p counters (where stored) and the
    counters recomputed from the reference's per-shot outputs (all cases)."""
    from qldpcsim_b200 import simulator
    g = load_golden(name)
    if "OSD" in name or g["decType"] == "BP":
        pytest.skip("tolerance-based cases are covered per shot")
    if g["code"] == "shor" and g["decType"] in ("MS", "BP"):
        pytest.skip("unusable in the reference")
    r = simulator.simulate_p(g["Hx"], g["Hz"], float(g["p"]), shots=int(g["shots"]), decType=g["decType"],
                             decIterations=int(g["decIterations"]), decSchedule=g["sched"], OSDorder=int(g["OSDorder"]),
                             record=g["rec"])
    Hx, Hz = g["Hx"].astype(np.int64), g["Hz"].astype(np.int64)
    exact = (g["eX_ref"] == g["errX"]).all(axis=1) & (g["eZ_ref"] == g["errZ"]).all(axis=1)
    failX = (((g["eX_ref"].astype(np.int64) @ Hz.T) % 2) != g["sy_z"]).any(axis=1)
    failZ = (((g["eZ_ref"].astype(np.int64) @ Hx.T) % 2) != g["sy_x"]).any(axis=1)
    shots = int(g["shots"])
    assert r["DecFailures_X"] == int(failX.sum()) and r["DecFailures_Z"] == int(failZ.sum())
    assert r["decSuccessExact"] == int(exact.sum())
    assert r["decSuccessDegen"] == 0
    assert round(r["Avg_number_of_iterations_X"] * shots) == int(g["itX"].sum())
    assert round(r["Avg_number_of_iterations_Z"] * shots) == int(g["itZ"].sum())
    if "counters" in g:
        c = g["counters"]
        assert [r["DecFailures_X"], r["DecFailures_Z"], r["decSuccessExact"], r["decSuccessDegen"]] == list(c[:4])


# ---------------------------------------------------------------------------------------------------------
# larger seeded batches against the CPU oracle
# ---------------------------------------------------------------------------------------------------------
ORACLE_CASES = [
    # code, decType, sched, p, shots, iters
    ("steane", "MS", "F", 0.10, 20000, 50),
    ("LP04_0", "MS", "L", 0.05, 20000, 50),
    ("LP04_0", "MS", "F", 0.08, 5000, 50),
    ("LP04_0", "MS", "S", 0.05, 2000, 50),
    ("LP118_0", "MS", "L", 0.05, 10000, 50),
    ("LP118_0", "MS", "L", 0.05, 100000, 50),      # the headline configuration, 10^5 shots bit for bit
    ("LP118_2", "MS", "S", 0.05, 70000, 50),       # batch large enough for the automatic lane-per-shot path + hand-over
    ("LP118_0", "MS", "L", 0.10, 3000, 50),
    ("LP118_0", "MS", "F", 0.05, 4000, 50),
    ("LP118_2", "MS", "L", 0.05, 2000, 50),
    ("LP118_2", "MS", "S", 0.05, 300, 50),
    ("T", "MS", "L", 0.04, 2000, 50),
    ("bicycle", "MS", "L", 0.03, 4000, 50),
    ("LP04_0", "NG", "F", 0.03, 5000, 50),
    ("LP118_0", "NG", "F", 0.02, 2000, 50),
    ("LP04_0", "BF", "F", 0.02, 5000, 50),
    ("LP118_0", "BF", "F", 0.02, 2000, 50),
    ("shor", "NG", "F", 0.05, 5000, 50),
    ("shor", "BF", "F", 0.05, 5000, 50),
    ("LP04_0", "BP", "F", 0.05, 3000, 100),
    ("LP04_0", "BP", "L", 0.08, 2000, 30),
    ("LP118_0", "BP", "F", 0.05, 1500, 100),
    ("bicycle", "BP", "F", 0.03, 1500, 50),
]


@pytest.mark.parametrize("code,decType,sched,p,shots,iters", ORACLE_CASES)
def test_against_oracle(code, decType, sched, p, shots, iters, cuda_device):
    from oracle import oracle
    from qldpcsim_b200 import pcmlibrary, sampler, simulator
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=4321)
    want = oracle.simulate_p(Hx, Hz, rec, p, decType=decType, decIterations=iters, decSchedule=sched, details=True)
    got = simulator.simulate_p(Hx, Hz, p, shots=shots, decType=decType, decIterations=iters, decSchedule=sched,
                               record=rec, details=True)
    wd, gd = want["_details"], got["_details"]
    n = Hx.shape[1]
    from qldpcsim_b200 import bitpack
    eX = bitpack.unpack_rows(gd["eX"].view(np.uint32), n)
    eZ = bitpack.unpack_rows(gd["eZ"].view(np.uint32), n)
    sameX = (eX == wd["eX"]).all(axis=1) & (gd["itX"] == wd["itX"])
    sameZ = (eZ == wd["eZ"]).all(axis=1) & (gd["itZ"] == wd["itZ"])
    if decType in ("NG", "BF"):
        assert sameX.all() and sameZ.all()
        for k in ("DecFailures_X", "DecFailures_Z", "decSuccessExact", "decSuccessDegen",
                  "Avg_number_of_iterations_X", "Avg_number_of_iterations_Z"):
            assert got[k] == want[k], k
    elif decType == "MS":
        assert sameX.mean() >= MS_BAR and sameZ.mean() >= MS_BAR, (sameX.mean(), sameZ.mean())
        assert sameX.all() and sameZ.all(), "MS is specified to the bit; any mismatch is a bug"
        for k in ("DecFailures_X", "DecFailures_Z", "decSuccessExact", "decSuccessDegen"):
            assert got[k] == want[k], k
    else:
        assert sameX.mean() >= BP_BAR and sameZ.mean() >= BP_BAR, (sameX.mean(), sameZ.mean())
        assert sameX.all() and sameZ.all(), "BP uses NumPy's tanh / arctanh algorithms on both sides: any mismatch is a bug"
        for k in ("DecFailures_X", "DecFailures_Z", "decSuccessExact", "decSuccessDegen"):
            assert got[k] == want[k], k


# ---------------------------------------------------------------------------------------------------------
# OSD at BASELINE-config size: n > 992 columns runs osd_kernel<2> (two words per lane), which none of the reference
# goldens small enough to generate in bulk reaches.  The oracle (orc_osd, pinned on the reference goldens incl. the
# LP118_2 / Tanner ones) orders columns like the library (stable), so EVERY shot must be bit-exact, ties included.
# ---------------------------------------------------------------------------------------------------------
OSD_BIG_CASES = [
    # code, sched, p, shots, iters, order, min unconverged decodes (X + Z)
    ("LP118_2", "S", 0.05, 700, 3, 0, 200),
    ("LP118_2", "S", 0.05, 700, 3, 1, 200),
    ("LP118_2", "S", 0.05, 700, 3, 10, 200),
    ("T", "L", 0.05, 700, 3, 0, 200),
    ("T", "L", 0.05, 700, 3, 1, 200),
    ("T", "L", 0.05, 700, 3, 10, 200),
    ("LP118_2", "S", 0.05, 70000, 50, 10, 100),     # BASELINE config 3 as stated: MS serial, 50 iterations, OSD order 10
]


@pytest.mark.parametrize("code,sched,p,shots,iters,order,min_unconv", OSD_BIG_CASES)
def test_osd_config3_size_against_oracle(code, sched, p, shots, iters, order, min_unconv, cuda_device):
    from oracle import oracle
    from qldpcsim_b200 import bitpack, pcmlibrary, sampler, simulator
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    n = Hx.shape[1]
    assert n > 992, "the case must run the two-words-per-lane OSD kernel"
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=2468)
    want = oracle.simulate_p(Hx, Hz, rec, p, decType="MS", decIterations=iters, decSchedule=sched, OSDorder=order, details=True)
    got = simulator.simulate_p(Hx, Hz, p, shots=shots, decType="MS", decIterations=iters, decSchedule=sched, OSDorder=order,
                               record=rec, details=True)
    wd, gd = want["_details"], got["_details"]
    unconv = int((~wd["convX"]).sum() + (~wd["convZ"]).sum())
    assert unconv >= min_unconv, f"only {unconv} unconverged decodes reach OSD"
    eX = bitpack.unpack_rows(gd["eX"].view(np.uint32), n)
    eZ = bitpack.unpack_rows(gd["eZ"].view(np.uint32), n)
    assert np.array_equal(gd["itX"], wd["itX"]) and np.array_equal(gd["itZ"], wd["itZ"])
    badX, badZ = np.nonzero((eX != wd["eX"]).any(axis=1))[0], np.nonzero((eZ != wd["eZ"]).any(axis=1))[0]
    assert len(badX) == 0 and len(badZ) == 0, f"OSD-{order} differs from the oracle on shots X {badX[:5]} Z {badZ[:5]}"
    # every shot reproduces its syndrome after OSD (rank(H) basis columns always exist)
    assert got["DecFailures_X"] == 0 and got["DecFailures_Z"] == 0
    assert got["decSuccessExact"] == want["decSuccessExact"]
    if order >= 2:
        # App. B-8: the reference's order loop aliases its buffers, so order >= 2 returns the order-0 vector
        got0 = simulator.simulate_p(Hx, Hz, p, shots=min(shots, 2000), decType="MS", decIterations=iters, decSchedule=sched,
                                    OSDorder=0, record=rec[:2000], details=True)["_details"]
        assert np.array_equal(got0["eX"], gd["eX"][:2000]) and np.array_equal(got0["eZ"], gd["eZ"][:2000])


@pytest.mark.parametrize("code,rank_x,rank_z", [("steane", 3, 3), ("shor", 2, 6), ("LP04_0", 78, 78), ("LP118_0", 232, 232),
                                                ("LP118_2", 442, 442), ("T", 457, 457), ("bicycle", 55, 55)])
def test_plan_rank_matches_survey(code, rank_x, rank_z, cuda_device):
    """GF(2) rank computed at plan creation (gf2math.py:91-135) = SURVEY.md App. C; OSD stops its column walk there."""
    from oracle import oracle
    from qldpcsim_b200 import pcmlibrary
    from qldpcsim_b200.decoders import Decoder
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    assert Decoder(Hx, "NG").info()["rank"] == rank_x == oracle.gf2_rank(Hx)
    assert Decoder(Hz, "NG").info()["rank"] == rank_z == oracle.gf2_rank(Hz)


def test_reference_signature_functions(cuda_device):
    """The five per-shot functions keep the reference's signatures, return types and dtypes."""
    from qldpcsim_b200 import decoders as D
    g = load_golden("LP04_0_MS_L_p05")
    from qldpcsim_b200.pcm import schedule_layers
    lX, _ = schedule_layers(g["Hx"], g["Hz"], "L")
    s = 3
    e, it = D.MS_decoder(g["Hz"], g["sy_z"][s].astype(int), p=0.05 / 3, max_iter=50, layers=lX)
    assert e.dtype == np.int8 and isinstance(it, int) and np.array_equal(e, g["eX_ref"][s]) and it == g["itX"][s]
    e, it = D.BP_decoder(g["Hz"], g["sy_z"][s].astype(int), p=0.05 / 3, max_iter=50, layers=lX)
    assert e.dtype == np.int64 and e.shape == (g["n"],)
    e, it = D.BF_decoder(g["Hz"], g["sy_z"][s].astype(int))
    assert e.dtype == np.bool_
    e, it = D.NG_decoder(g["Hz"], g["sy_z"][s].astype(int))
    assert e.dtype == np.int8
    with pytest.raises(AttributeError):
        D.MS_decoder(g["Hz"], g["sy_z"][s].astype(int), p=0.01)        # layers=None is unusable in the reference too
    out = D.MS_decoder(np.zeros((0, 5)), np.zeros(0), p=0.01, layers=[])
    assert isinstance(out, np.ndarray) and out.size == 0                # decoders.py:138-139: bare array


def test_osddec_function(cuda_device):
    from qldpcsim_b200 import bitpack, decoders as D
    g = load_golden("LP04_0_MS_L_OSD0_p10")
    k = 0
    H = g["Hz"] if g["osd_which"][k] == 0 else g["Hx"]
    sy = (g["sy_z"] if g["osd_which"][k] == 0 else g["sy_x"])[g["osd_shot"][k]]
    ref = (g["eX_ref"] if g["osd_which"][k] == 0 else g["eZ_ref"])[g["osd_shot"][k]]
    e = bitpack.unpack_rows(g["osd_e_in"][k:k + 1], g["n"])[0].astype(np.int8)
    out = D.OSDdec(H, e, sy.astype(int), g["osd_llr"][k], 0, perm=g["osd_perm"][k])
    assert out is e and np.array_equal(e, ref)


# ---------------------------------------------------------------------------------------------------------
# device sampler and properties at scale
# ---------------------------------------------------------------------------------------------------------
def _philox_host(c0, c1, c2, c3, k0, k1):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    mask = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & mask, p1 & mask, ((p0 >> 32) ^ c3 ^ k1) & mask, p0 & mask
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    return c0, c1, c2, c3


def _check_device_sampler(Hx, Hz, p, seed, first, shots):
    """qldpc_sample against a pure-Python restatement of its definition (Philox4x32-10, thresholds k p/3 2^32)."""
    from qldpcsim_b200 import bitpack, simulator
    pipe = simulator.Pipeline(Hx, Hz, p, "NG")
    synz, synx, errx, errz = [t.cpu().numpy().view(np.uint32) for t in pipe.sample_device(shots, seed, first)]
    n = Hx.shape[1]
    thr = [min(int((k + 1) * p / 3.0 * 4294967296.0 + 0.5), 4294967296) for k in range(3)]
    eX = np.zeros((shots, n), np.uint8)
    eZ = np.zeros((shots, n), np.uint8)
    for s in range(shots):
        gs = first + s
        for q in range(n):
            w, bit = q // 32, q % 32
            r = _philox_host(gs & 0xFFFFFFFF, gs >> 32, w * 8 + bit // 4, 0, seed & 0xFFFFFFFF, seed >> 32)[bit % 4]
            X, Y, Z = r < thr[0], thr[0] <= r < thr[1], thr[1] <= r < thr[2]
            eX[s, q], eZ[s, q] = X or Y, Z or Y
    assert np.array_equal(bitpack.unpack_rows(errx, n), eX) and np.array_equal(bitpack.unpack_rows(errz, n), eZ)
    assert np.array_equal(bitpack.unpack_rows(synz, Hz.shape[0]), (eX.astype(np.int64) @ Hz.T.astype(np.int64)) % 2)
    assert np.array_equal(bitpack.unpack_rows(synx, Hx.shape[0]), (eZ.astype(np.int64) @ Hx.T.astype(np.int64)) % 2)


def test_device_sampler_matches_host_restatement(cuda_device):
    from qldpcsim_b200 import pcmlibrary
    Hx, Hz = pcmlibrary.by_name("LP04_0")
    _check_device_sampler(Hx, Hz, 0.07, 0x1234ABCD5678, 1000, 6)


def big_sparse_matrix(seed=5, m=2200, n=4000, rw=30):
    """Synthetic matrix beyond the sizes of the reference's library: more than 1024 checks (row-wise sampler / classifier,
    dense bit-flipping kernel) and more than 65535 edges (naive-greedy tables read through L1 instead of shared memory)."""
    rng = np.random.default_rng(seed)
    H = np.zeros((m, n), np.int8)
    for i in range(m):
        H[i, rng.choice(n, rw, replace=False)] = 1
    return H


def test_more_than_1024_checks(cuda_device):
    from oracle import oracle
    from qldpcsim_b200 import bitpack, sampler, simulator
    Hx, Hz = big_sparse_matrix(5), big_sparse_matrix(6)
    n = Hx.shape[1]
    _check_device_sampler(Hx, Hz, 0.01, 0xBEEF, 123456789012, 2)                       # sample_kernel<0>
    shots, p = 40, 0.004
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=31)
    for dt in ("NG", "BF"):                                                             # ng_decode_kernel<false>, bf_decode_kernel
        want = oracle.simulate_p(Hx, Hz, rec, p, decType=dt, details=True)
        got = simulator.simulate_p(Hx, Hz, p, shots=shots, decType=dt, record=rec, details=True)     # classify_kernel<0, .>
        eX = bitpack.unpack_rows(got["_details"]["eX"].view(np.uint32), n)
        eZ = bitpack.unpack_rows(got["_details"]["eZ"].view(np.uint32), n)
        assert np.array_equal(eX, want["_details"]["eX"]) and np.array_equal(eZ, want["_details"]["eZ"]), dt
        assert np.array_equal(got["_details"]["itX"], want["_details"]["itX"]), dt
        for k in ("DecFailures_X", "DecFailures_Z", "decSuccessExact", "decSuccessDegen"):
            assert got[k] == want[k], (dt, k)
        assert 0 < want["DecFailures_X"] + want["DecFailures_Z"] or dt == "NG"


def test_device_sampler_statistics_and_sharding(cuda_device):
    from qldpcsim_b200 import bitpack, pcmlibrary, simulator
    Hx, Hz = pcmlibrary.by_name("LP118_0")
    p, shots = 0.06, 200000
    pipe = simulator.Pipeline(Hx, Hz, p, "NG")
    full = [t.cpu().numpy() for t in pipe.sample_device(shots, 99, 0)]
    a = [t.cpu().numpy() for t in pipe.sample_device(shots // 2, 99, 0)]
    b = [t.cpu().numpy() for t in pipe.sample_device(shots - shots // 2, 99, shots // 2)]
    for f, x, y in zip(full, a, b):
        assert np.array_equal(f, np.concatenate([x, y])), "batch must not depend on the sharding"
    n = Hx.shape[1]
    eX = bitpack.unpack_rows(full[2].view(np.uint32), n)
    eZ = bitpack.unpack_rows(full[3].view(np.uint32), n)
    N = shots * n
    for est, want in ((eX.mean(), 2 * p / 3), (eZ.mean(), 2 * p / 3), ((eX & eZ).mean(), p / 3)):
        assert abs(est - want) < 5 * np.sqrt(want * (1 - want) / N)


def test_full_size_properties(cuda_device):
    """BASELINE-size batch (10^6 shots, LP118_0 MS-L 50 it): converged <=> syndrome reproduced, iteration range,
    counters independent of chunking, exact + failures consistent."""
    import torch
    from qldpcsim_b200 import pcmlibrary, simulator
    Hx, Hz = pcmlibrary.by_name("LP118_0")
    shots, p = 1_000_000, 0.05
    pipe = simulator.Pipeline(Hx, Hz, p, "MS", 50, "L")
    synz, synx, errx, errz = pipe.sample_device(shots, 7, 0)
    c1 = pipe.run(synz, synx, errx, errz, keep=True).cpu().numpy()
    last = pipe.last
    conv = last["convX"].bool()
    it = last["itX"]
    assert int(it.min()) >= 1 and int(it.max()) <= 50
    assert bool((it[~conv] == 50).all())
    # failures counted by the classifier (fresh syndrome computation) == unconverged shots of the decoder
    assert c1[0] == int((~conv).sum().item()) and c1[1] == int((~last["convZ"].bool()).sum().item())
    assert c1[6] == shots and c1[4] == int(it.sum().item())
    # chunking invariance (what multi-GPU sharding relies on)
    from qldpcsim_b200 import _lib
    c2 = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=pipe.device)
    for lo in range(0, shots, 300_000):
        hi = min(shots, lo + 300_000)
        pipe.run(synz[lo:hi], synx[lo:hi], errx[lo:hi], errz[lo:hi], counters=c2)
    assert np.array_equal(c1, c2.cpu().numpy())
    # exact matches cannot be failures
    assert c1[2] + max(c1[0], c1[1]) <= shots


CLASS_CASES = [("steane", "MS", "F", 0.08, 4000), ("shor", "NG", "F", 0.08, 4000), ("LP04_0", "MS", "L", 0.08, 3000),
               ("LP04_0", "NG", "F", 0.03, 3000), ("LP118_0", "MS", "L", 0.09, 1500), ("bicycle", "BF", "F", 0.02, 1500)]


@pytest.mark.parametrize("code,decType,sched,p,shots", CLASS_CASES)
def test_outcome_classes(code, decType, sched, p, shots, cuda_device):
    """Extension counters (README.md:15-22: exact / degenerate / logical error / decoder failure) against an independent
    CPU restatement (rank test instead of logical operators); the reference-compatible counters are unaffected."""
    from oracle import oracle
    from qldpcsim_b200 import bitpack, pcmlibrary, sampler, simulator
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=99)
    plain = simulator.simulate_p(Hx, Hz, p, shots=shots, decType=decType, decIterations=20, decSchedule=sched, record=rec)
    got = simulator.simulate_p(Hx, Hz, p, shots=shots, decType=decType, decIterations=20, decSchedule=sched, record=rec,
                               details=True, classes=True)
    for k in plain:
        assert got[k] == plain[k], k
    n = Hx.shape[1]
    eX = bitpack.unpack_rows(got["_details"]["eX"].view(np.uint32), n)
    eZ = bitpack.unpack_rows(got["_details"]["eZ"].view(np.uint32), n)
    _, _, errX, errZ = oracle.split_record(rec, Hz.shape[0], Hx.shape[0], n)
    want = oracle.outcome_classes(Hx, Hz, errX, errZ, eX, eZ)
    assert got["outcome_exact"] == want["exact"] == got["decSuccessExact"]
    assert got["outcome_degenerate"] == want["degenerate"]
    assert got["outcome_logical_error"] == want["logical_error"]
    assert got["outcome_decoder_failure"] == want["decoder_failure"]
    assert sum(want.values()) == shots


@pytest.mark.parametrize("tag", ["02", "05", "10"])
def test_bp_big_reference_goldens_bit_exact(tag, cuda_device):
    """GPU sum-product decoder against 1600 decodes of the unmodified reference per depolarizing probability (BASELINE config
    2's code, decoder and iteration count; p = 0.02 / 0.05 / 0.10).  The kernel evaluates np.tanh / np.arctanh with NumPy's
    own algorithms (csrc/npymath.cuh), so every decode -- converged or not -- must be bit-identical."""
    import os
    from conftest import GOLDEN_DIR
    from qldpcsim_b200 import bitpack, pcm, pcmlibrary
    from qldpcsim_b200.decoders import Decoder
    g = np.load(os.path.join(GOLDEN_DIR, f"big_LP118_0_BP_F_p{tag}_X.npz"))
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name("LP118_0")]
    m, n = Hz.shape
    syn = bitpack.unpack_rows(g["syn"], m).astype(np.uint8)
    e_ref, it_ref, iters = bitpack.unpack_rows(g["e"], n), g["it"], int(g["decIterations"])
    lX, _ = pcm.schedule_layers(Hx, Hz, "F")
    out = Decoder(Hz, "BP", p=float(g["p"]) / 3, max_iter=iters, layers=lX).decode(syn)
    same = (out["e_hat"] == e_ref).all(1) & (out["iters"] == it_ref)
    assert same.all(), f"p=0.{tag}: {int((~same).sum())} of {len(same)} decodes differ from the reference"


def test_osd_shared_memory_kernel_still_exact(cuda_device):
    """The register-resident OSD kernel serves every code of the reference's library; the shared-memory formulation remains for
    larger matrices (more than 512 checks or 1088 columns).  QLDPC_OSD_KERNEL=s forces it: the reference OSD goldens (LP04_0,
    LP118_0 and -- two words per lane -- LP118_2 / Tanner) must stay bit-exact through it (environment knobs are read once per
    process, hence the sub-process)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, QLDPC_OSD_KERNEL="s")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu", "-k",
                        "test_golden_osd or (test_osd_config3_size_against_oracle and 700)"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("dt,sched,osd,shots", [("MS", "L", -1, 5 * (1 << 18) + 7), ("MS", "L", 0, 6000), ("BP", "F", -1, 6000), ("NG", "F", -1, 6000)])
def test_simulate_host_matches_device_path(dt, sched, osd, shots, cuda_device):
    """qldpc_simulate_host (host record in, counters out; chunks of 2^19 shots double buffered on two compute streams behind a copy stream) gives the counters of the
    device-resident path (decode X, decode Z, classify) on the same batch -- several chunks in flight, OSD plans, other decoders."""
    import torch
    from qldpcsim_b200 import pcmlibrary, simulator
    Hx, Hz = pcmlibrary.by_name("LP04_0")
    pipe = simulator.Pipeline(Hx, Hz, 0.08, dt, 12, sched, osd, logicals=True)
    batch = pipe.sample_device(shots, 3, 0)
    want = pipe.run(*batch).cpu()
    host = [b.cpu().contiguous() for b in batch]
    got = pipe.run_host(*host)
    torch.cuda.synchronize()
    assert torch.equal(got, want), (got.tolist(), want.tolist())
    assert int(got[6]) == shots
    # pageable NumPy arrays are accepted as well, and the call can be repeated on the same plans
    got2 = pipe.run_host(*[h.numpy() for h in host])
    assert torch.equal(got2, want)
