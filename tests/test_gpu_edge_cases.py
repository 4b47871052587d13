"""Edge cases on the GPU against the CPU oracle: random irregular matrices (empty rows/columns, ragged layers, repeated
and missing checks in the layer list), every instantiated kernel shape, extreme priors, tiny and empty batches."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cmp(H, syn, decType, cuda_device, **kw):
    from oracle import oracle
    from qldpcsim_b200.decoders import Decoder
    okw = dict(kw)
    want = oracle.Graph(H).decode(decType, syn, **okw)
    got = Decoder(H, decType, **kw).decode(syn)
    return want, got


def _random_case(rng, m, n, density, max_rw=None):
    H = (rng.random((m, n)) < density).astype(np.int8)
    if max_rw:
        for i in range(m):
            idx = np.nonzero(H[i])[0]
            if len(idx) > max_rw:
                H[i, rng.choice(idx, len(idx) - max_rw, replace=False)] = 0
    e = (rng.random((200, n)) < 0.08).astype(np.int64)
    syn = (e @ H.T.astype(np.int64)) % 2
    return H, syn.astype(np.uint8)


@pytest.mark.parametrize("seed", range(12))
def test_ms_random_irregular(seed, cuda_device):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(2, 40)), int(rng.integers(3, 70))
    H, syn = _random_case(rng, m, n, rng.choice([0.05, 0.15, 0.3]), max_rw=30)
    # layer list: random contiguous split, sometimes shuffled, sometimes with a check repeated and one left out
    cuts = sorted(set([0, m] + list(rng.integers(0, m + 1, size=rng.integers(0, 6)))))
    layers = [np.arange(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    if seed % 3 == 1:
        rng.shuffle(layers)
    if seed % 3 == 2 and m > 3:
        layers = layers + [np.array([0, m - 1])]
        layers[0] = layers[0][1:]
    if (H.sum(axis=0).max(initial=0) > 16):
        pytest.skip("column weight above the instantiated kernels")
    p = float(rng.choice([0.01, 0.1, 0.3]))
    want, got = _cmp(H, syn, "MS", cuda_device, p=p, max_iter=int(rng.integers(1, 12)), layers=layers)
    assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["converged"], want["converged"])


@pytest.mark.parametrize("rw", [1, 3, 4, 5, 8, 9, 17, 18, 19, 31, 32])
def test_ms_every_row_weight_shape(rw, cuda_device):
    """Row weights on both sides of every instantiated DC (4, 8, 18, 32), regular and ragged."""
    rng = np.random.default_rng(rw)
    m, n = 12, 96
    for ragged in (False, True):
        H = np.zeros((m, n), np.int8)
        for i in range(m):
            w = rw if not ragged else int(rng.integers(1, rw + 1))
            H[i, rng.choice(n, w, replace=False)] = 1
        e = (rng.random((300, n)) < 0.05).astype(np.int64)
        syn = ((e @ H.T.astype(np.int64)) % 2).astype(np.uint8)
        for layers in ([np.arange(m)], [np.arange(0, 5), np.arange(5, 12)], [np.array([i]) for i in range(m)]):
            want, got = _cmp(H, syn, "MS", cuda_device, p=0.02, max_iter=8, layers=layers)
            assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"]), (rw, ragged, len(layers))


@pytest.mark.parametrize("cw", [1, 4, 5, 6, 9, 10, 16])
def test_ms_every_column_weight_shape(cw, cuda_device):
    rng = np.random.default_rng(100 + cw)
    m, n = 40, 24
    H = np.zeros((m, n), np.int8)
    for j in range(n):
        H[rng.choice(m, cw, replace=False), j] = 1
    e = (rng.random((200, n)) < 0.1).astype(np.int64)
    syn = ((e @ H.T.astype(np.int64)) % 2).astype(np.uint8)
    want, got = _cmp(H, syn, "MS", cuda_device, p=0.05, max_iter=6, layers=[np.arange(0, 20), np.arange(20, 40)])
    assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])


def test_ms_extreme_priors(cuda_device):
    """p > 1/2 (negative prior: untouched variables decide 1), p -> 0 (eps clamp), p = 1/2 (zero prior)."""
    from qldpcsim_b200 import pcmlibrary
    Hx, _ = pcmlibrary.by_name("LP04_0")
    rng = np.random.default_rng(3)
    e = (rng.random((300, Hx.shape[1])) < 0.05).astype(np.int64)
    syn = ((e @ Hx.T) % 2).astype(np.uint8)
    from qldpcsim_b200.pcm import layerize
    lay = layerize(Hx)
    for p in (0.8, 0.5, 1e-12, 0.0, 0.999999):
        want, got = _cmp((Hx % 2).astype(np.int8), syn, "MS", cuda_device, p=p, max_iter=5, layers=lay)
        assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"]), p


def test_ms_other_beta_and_max_iter_edge(cuda_device):
    from qldpcsim_b200 import pcmlibrary
    Hx, _ = pcmlibrary.by_name("LP04_0")
    H = (Hx % 2).astype(np.int8)
    rng = np.random.default_rng(4)
    e = (rng.random((200, H.shape[1])) < 0.06).astype(np.int64)
    syn = ((e @ H.T.astype(np.int64)) % 2).astype(np.uint8)
    for beta, it in ((1.0, 3), (0.5, 1), (0.9, 2)):
        want, got = _cmp(H, syn, "MS", cuda_device, p=0.02, max_iter=it, layers=[np.arange(H.shape[0])], beta=beta)
        assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])


@pytest.mark.parametrize("beta", [-0.75, 1.0 / 3.0, 0.1, 1.7, 0.0])
def test_ms_rounded_minimum_rule(beta, cuda_device):
    """The kernel takes min / second min on the binary32-ROUNDED magnitudes (see ms_check_phase); the reference takes them
    in binary64 before rounding.  Many iterations at a high error rate make the messages diverge until distinct binary64
    magnitudes round to the same binary32 value -- the case in which the two formulations could differ if the argument in
    the kernel header were wrong -- and non-dyadic or negative normalisations exercise the rounding of the product."""
    from qldpcsim_b200 import pcmlibrary
    from qldpcsim_b200.pcm import layerize
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name("LP04_0")]
    rng = np.random.default_rng(11)
    e = (rng.random((1500, Hz.shape[1])) < 0.09).astype(np.int64)
    syn = ((e @ Hz.T.astype(np.int64)) % 2).astype(np.uint8)
    want, got = _cmp(Hz, syn, "MS", cuda_device, p=0.03, max_iter=40, layers=layerize(Hx), beta=beta)
    assert (want["iters"] == 40).sum() > 50, "the case must contain non-converging shots"
    assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["converged"], want["converged"])


@pytest.mark.parametrize("seed", range(6))
def test_bp_bf_ng_random_irregular(seed, cuda_device):
    rng = np.random.default_rng(50 + seed)
    m, n = int(rng.integers(2, 30)), int(rng.integers(3, 50))
    H, syn = _random_case(rng, m, n, 0.15, max_rw=30)
    for dt, kw in (("NG", {}), ("BF", {"max_iter": 50}), ("BF", {"max_iter": 3})):
        want, got = _cmp(H, syn, dt, cuda_device, **kw)
        assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"]), dt
        assert np.array_equal(got["converged"], want["converged"]), dt
    layers = [np.arange(0, m // 2), np.arange(m // 2, m)]
    want, got = _cmp(H, syn, "BP", cuda_device, p=0.05, max_iter=6, layers=layers)
    assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])   # same tanh / arctanh algorithms


def test_batch_sizes(cuda_device):
    """0, 1, and a batch that does not fill one CTA; output buffers untouched beyond the batch."""
    from oracle import oracle
    from qldpcsim_b200 import pcmlibrary
    from qldpcsim_b200.decoders import Decoder
    H = (pcmlibrary.steane_code()[0] % 2).astype(np.int8)
    d = Decoder(H, "MS", p=0.05, max_iter=50, layers=[np.arange(3)])
    out = d.decode(np.zeros((0, 3), np.uint8))
    assert out["e_hat"].shape == (0, 7) and out["iters"].shape == (0,)
    syn = np.array([[1, 0, 1]], np.uint8)
    want = oracle.Graph(H).decode("MS", syn, p=0.05, max_iter=50, layers=[np.arange(3)])
    got = d.decode(syn[0])                                  # 1-D syndrome accepted
    assert np.array_equal(got["e_hat"], want["e_hat"]) and got["iters"][0] == want["iters"][0]
    with pytest.raises(ValueError):
        d.decode(np.zeros((2, 4), np.uint8))                # wrong syndrome length


def test_plan_errors(cuda_device):
    from qldpcsim_b200 import _lib
    from qldpcsim_b200.decoders import Decoder
    H = np.zeros((3, 5), np.int8)
    H[0, :] = 1
    with pytest.raises(IndexError):
        Decoder(H, "MS", p=0.1, layers=[np.array([0, 3])])          # check index outside the matrix
    with pytest.raises(ValueError):
        Decoder(H, "XX")
    big = np.ones((40, 40), np.int8)                                 # row weight 40 > 32
    with pytest.raises(_lib.QldpcError):
        Decoder(big, "MS", p=0.1, layers=[np.arange(40)])


def test_integration_md_stub_runs(cuda_device):
    """The ctypes binding printed in INTEGRATION.md is executable as written (library path substituted)."""
    import os
    import re
    from conftest import ROOT, load_golden
    from qldpcsim_b200 import _lib
    from qldpcsim_b200.pcm import schedule_layers
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.findall(r"```python\n(import ctypes, numpy as np.*?)```", text, flags=re.S)[0]
    block = block.replace('ctypes.CDLL("libqldpc_b200.so")', f'ctypes.CDLL("{_lib.LIB_PATH}")')
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    g = load_golden("LP04_0_MS_L_p05")
    lX, _ = schedule_layers(g["Hx"], g["Hz"], "L")
    e, it = ns["decode_batch"](g["Hz"], g["sy_z"], "MS", p=0.05 / 3, max_iter=50, layers=lX)
    assert np.array_equal(e, g["eX_ref"]) and np.array_equal(it, g["itX"])
    e, it = ns["decode_batch"](g["Hz"], g["sy_z"], "NG")
    assert e.shape == g["eX_ref"].shape


POSTERIOR_CASES = [("steane", "S", 0.1, 2000), ("LP04_0", "S", 0.05, 600), ("LP04_0", "L", 0.08, 600), ("LP118_0", "L", 0.05, 400),
                   ("LP118_0", "F", 0.05, 200), ("LP118_2", "S", 0.05, 100), ("bicycle", "L", 0.03, 300), ("T", "S", 0.03, 70)]


@pytest.mark.parametrize("code,sched,p,shots", POSTERIOR_CASES)
def test_ms_posterior_matches_oracle(code, sched, p, shots, cuda_device):
    """Estimates, iteration counts, convergence flags and the float64 posterior (what decoders.py:179-180 hands to OSDdec) of
    every shot on serial, layered and flooding schedules."""
    from oracle import oracle
    from qldpcsim_b200 import pcmlibrary, sampler
    from qldpcsim_b200.decoders import Decoder
    from qldpcsim_b200.pcm import schedule_layers
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=77)
    sy_z = rec[:, :Hz.shape[0]]
    lX, _ = schedule_layers(Hx, Hz, sched)
    want = oracle.Graph(Hz).decode("MS", sy_z, p=p / 3, max_iter=20, layers=lX, want_posterior=True)
    for kernel in ("auto", "plain"):
        got = Decoder(Hz, "MS", p=p / 3, max_iter=20, layers=lX, kernel=kernel).decode(sy_z, want_llr=True)
        assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])
        assert np.array_equal(got["converged"], want["converged"])
        assert np.array_equal(got["posterior"], want["posterior"])


MERGE_CASES = [("LP04_0", "S", 0.06, 3000, 50), ("LP118_0", "S", 0.06, 1500, 50), ("LP118_2", "S", 0.05, 1500, 50), ("T", "S", 0.04, 800, 50),
               ("LP118_2", "S", 0.09, 300, 8), ("steane", "S", 0.1, 2000, 50)]


@pytest.mark.parametrize("code,sched,p,shots,iters", MERGE_CASES)
def test_ms_merged_steps_match_oracle(code, sched, p, shots, iters, cuda_device):
    """Serial schedule: runs of single-check layers with disjoint variable sets are executed as one step of the kernel (SPEC
    instances) and committed sub-layer by sub-layer only near convergence.  Estimates, iteration counts (which encode the layer
    at which the reference's per-layer test fires, decoders.py:175-176) and the float64 posterior of EVERY shot -- converged in
    the middle of a run or not -- must equal the oracle's and the unmerged kernel's."""
    from oracle import oracle
    from qldpcsim_b200 import pcmlibrary, sampler
    from qldpcsim_b200.decoders import Decoder
    from qldpcsim_b200.pcm import schedule_layers
    Hx, Hz = [(h % 2).astype(np.int8) for h in pcmlibrary.by_name(code)]
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=13)
    for H, sy, lay in ((Hz, rec[:, :Hz.shape[0]], schedule_layers(Hx, Hz, sched)[0]),
                       (Hx, rec[:, Hz.shape[0]:Hz.shape[0] + Hx.shape[0]], schedule_layers(Hx, Hz, sched)[1])):
        want = oracle.Graph(H).decode("MS", sy, p=p / 3, max_iter=iters, layers=lay, want_posterior=True)
        d = Decoder(H, "MS", p=p / 3, max_iter=iters, layers=lay, kernel="auto")
        info = d.info()
        if code != "steane":
            assert info["steps_per_iteration"] * 4 < info["n_layers"], info       # merging is active
        got = d.decode(sy, want_llr=True)
        assert np.array_equal(got["iters"], want["iters"])
        assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["converged"], want["converged"])
        assert np.array_equal(got["posterior"], want["posterior"])
        plain = Decoder(H, "MS", p=p / 3, max_iter=iters, layers=lay, kernel="plain")
        assert plain.info()["steps_per_iteration"] == plain.info()["n_layers"]
        gp = plain.decode(sy, want_llr=True)
        assert np.array_equal(gp["e_hat"], want["e_hat"]) and np.array_equal(gp["iters"], want["iters"])
        assert np.array_equal(gp["posterior"], want["posterior"])


def test_decode_host_many_chunks_serial(cuda_device):
    """qldpc_decode_host keeps two chunks in flight on two streams; their scratch (work counters, staging buffers, hand-over
    lists) must be per slot.  Serial-schedule plan, more than two chunks, against the device-resident path."""
    import torch
    from qldpcsim_b200 import pcmlibrary, simulator
    Hx, Hz = pcmlibrary.by_name("LP118_2")
    shots = 3 * (1 << 18) + 4321
    pipe = simulator.Pipeline(Hx, Hz, 0.05, "MS", 50, "S")
    synz, _, _, _ = pipe.sample_device(shots, 5, 0)
    e_dev, it_dev, cv_dev, _ = pipe.decX.decode_packed(synz)
    e_h, it_h, cv_h, _ = pipe.decX.decode_host_packed(synz.cpu().numpy().view(np.uint32))
    torch.cuda.synchronize()
    assert np.array_equal(e_h.view(np.int32), e_dev.cpu().numpy()) and np.array_equal(it_h, it_dev.cpu().numpy())
    assert np.array_equal(cv_h, cv_dev.cpu().numpy())


@pytest.mark.parametrize("m,n,rw,cwmax", [(50, 60, 7, 16), (100, 120, 14, 16), (73, 146, 18, 9), (30, 64, 30, 16), (140, 80, 8, 16)])
def test_ms_single_overlapping_check_layers(m, n, rw, cwmax, cuda_device):
    """Schedules whose every layer is one check that overlaps its neighbours run on the eight-lane kernel (ms_sub_kernel.cuh):
    every row-weight class (8 / 16 / 24 / 32 cells), one to five parity words, shots that never converge and keep flipping
    (the flip path: per-variable parity masks with the three-word instance, check lists otherwise) -- estimates, iteration
    counts, convergence flags and the float64 posterior of every shot against the oracle."""
    from oracle import oracle
    from qldpcsim_b200.decoders import Decoder
    rng = np.random.default_rng(1000 * m + rw)
    H = np.zeros((m, n), np.int8)
    if (m, n, rw) == (73, 146, 18):       # the shape of the bicycle code (regular: row weight 18, column weight 9) with a random graph
        for blk in range(2):
            first = rng.choice(73, 9, replace=False)
            for i in range(73):
                H[i, 73 * blk + (first + i) % 73] = 1
    else:
        colw = np.zeros(n, int)
        for i in range(m):
            ok = np.nonzero(colw < cwmax)[0]
            w = min(rw if i % 3 else max(1, rw - 2), len(ok))
            idx = rng.choice(ok, w, replace=False)
            H[i, idx] = 1
            colw[idx] += 1
    layers = [np.array([i]) for i in rng.permutation(m)]
    e = (rng.random((600, n)) < 0.07).astype(np.int64)
    syn = ((e @ H.T.astype(np.int64)) % 2).astype(np.uint8)
    syn[:40] = rng.integers(0, 2, size=(40, m))          # mostly undecodable: oscillating decisions until the iteration limit
    want = oracle.Graph(H).decode("MS", syn, p=0.03, max_iter=12, layers=layers, want_posterior=True)
    dec = Decoder(H, "MS", p=0.03, max_iter=12, layers=layers)
    got = dec.decode(syn, want_llr=True)
    assert dec.info()["shots_per_cta"] % 4 == 0 and dec.info()["steps_per_iteration"] == m
    assert not want["converged"].all() and want["converged"].any()
    assert np.array_equal(got["e_hat"], want["e_hat"]) and np.array_equal(got["iters"], want["iters"])
    assert np.array_equal(got["converged"], want["converged"])
    assert np.array_equal(got["posterior"], want["posterior"])
