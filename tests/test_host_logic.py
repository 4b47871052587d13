"""Host-side logic: code library, PCM compiler, layer partition, bit packing, sampler, sharding, result table."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from qldpcsim_b200 import bitpack, pcm, pcmlibrary, sampler

ALL_CODES = ["steane", "shor", "LP04_0", "LP04_1", "LP04_2", "LP04_3", "LP118_0", "LP118_1", "LP118_2", "T", "bicycle"]


@pytest.mark.parametrize("code", ALL_CODES)
def test_library_matches_reference_matrices(code):
    z = np.load(os.path.join(GOLDEN_DIR, "codes.npz"))
    for tag, H in zip("xz", pcmlibrary.by_name(code)):
        shape = tuple(z[f"{code}_H{tag}_shape"])
        r, c = z[f"{code}_H{tag}_rc"].astype(np.int64)
        want = np.zeros(shape, np.int64)
        want[r, c] = 1
        assert H.shape == shape and np.array_equal(H, want)
    Hx, Hz = pcmlibrary.by_name(code)
    assert not ((Hx @ Hz.T) % 2).any(), "CSS condition"


def test_library_errors():
    with pytest.raises(ValueError):
        pcmlibrary.qc_ldpc_lifted_code("LP04", 7)
    with pytest.raises(ValueError):
        pcmlibrary.qc_ldpc_lifted_code("LPXX", 0)


@pytest.mark.parametrize("code", ["steane", "LP04_0", "LP118_0", "bicycle", "shor"])
def test_compile_pcm_roundtrip(code):
    for H in pcmlibrary.by_name(code):
        c = pcm.compile_pcm(H)
        dense = np.zeros((c.m, c.n), np.int8)
        for i in range(c.m):
            cols = c.col_idx[c.row_ptr[i]:c.row_ptr[i + 1]]
            assert (np.diff(cols) > 0).all()
            dense[i, cols] = 1
        assert np.array_equal(dense, H % 2)
        dense2 = np.zeros_like(dense)
        for j in range(c.n):
            rows = c.row_idx[c.col_ptr[j]:c.col_ptr[j + 1]]
            assert (np.diff(rows) > 0).all()
            dense2[rows, j] = 1
        assert np.array_equal(dense2, dense)
        chk, var = np.where(H)                      # the reference's edge order (decoders.py:224)
        assert np.array_equal(var, c.col_idx)
        assert c.row_weight_max == H.sum(axis=1).max() and c.col_weight_max == H.sum(axis=0).max()


def test_layerize_structure():
    Hx, Hz = pcmlibrary.by_name("LP118_0")
    lx, lz = pcm.schedule_layers(Hx, Hz, "L")
    assert [len(l) for l in lx] == [16, 16, 16, 16, 32, 16, 16, 16, 32, 16, 16, 16, 16]
    assert [len(l) for l in lz] == [16, 16, 32, 16, 32, 16, 32, 16, 32, 16, 16]
    for H, layers in ((Hx, lx), (Hz, lz)):
        assert np.array_equal(np.concatenate(layers), np.arange(H.shape[0]))
        for l in layers:
            assert H[l].sum(axis=0).max() <= 1          # disjoint supports inside a layer
    sx, sz = pcm.schedule_layers(Hx, Hz, "S")
    assert len(sx) == Hx.shape[0] and all(len(l) == 1 for l in sx)
    fx, fz = pcm.schedule_layers(Hx, Hz, "F")
    assert len(fx) == 1 and len(fx[0]) == Hx.shape[0]
    with pytest.raises(ValueError):
        pcm.schedule_layers(Hx, Hz, "Q")
    # the partition handed to the decode on Hz is built from Hx (simulator.py:230-234, 278-282): not column-disjoint there
    assert max(Hz[l].sum(axis=0).max() for l in lx) > 1


def test_layerize_random_against_oracle():
    from oracle import oracle
    rng = np.random.default_rng(5)
    for _ in range(40):
        m, n = rng.integers(1, 30), rng.integers(1, 40)
        H = (rng.random((m, n)) < 0.12).astype(np.int8)
        for serial in (False, True):
            a, b = pcm.layerize(H, serial), oracle.layerize(H, serial)
            assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_bitpack_roundtrip_and_layout():
    rng = np.random.default_rng(0)
    for cols in (1, 7, 31, 32, 33, 64, 175, 544, 1054):
        bits = rng.integers(0, 2, (5, cols), dtype=np.uint8)
        w = bitpack.pack_rows(bits)
        assert w.dtype == np.uint32 and w.shape == (5, bitpack.words(cols))
        assert np.array_equal(bitpack.unpack_rows(w, cols), bits)
        j = cols - 1
        assert ((w[:, j // 32] >> np.uint32(j % 32)) & 1).astype(np.uint8).tolist() == bits[:, j].tolist()
    assert bitpack.pack_rows(np.zeros((0, 10))).shape == (0, 1)


def test_sampler_definition():
    Hx, Hz = pcmlibrary.by_name("LP04_0")
    p, shots, n = 0.05, 300, 175
    rec = sampler.sample_record(Hx, Hz, p, shots, seed=1234)
    u = np.random.default_rng(1234).random((shots, n))
    X, Y, Z = u < p / 3, (p / 3 <= u) & (u < 2 * p / 3), (2 * p / 3 <= u) & (u < p)
    eX, eZ = X | Y, Z | Y
    want = np.concatenate([(eX @ Hz.T) % 2, (eZ @ Hx.T) % 2, eX, eZ], axis=1).astype(bool)
    assert np.array_equal(rec, want)
    a, _ = sampler.sample_errors(n, p, 1000, seed=3, chunk=64)
    b, _ = sampler.sample_errors(n, p, 1000, seed=3, chunk=1 << 16)
    assert np.array_equal(a, b)


def test_load_matrix(tmp_path):
    H = pcmlibrary.steane_code()[0]
    np.save(tmp_path / "h.npy", H * 3)                       # reduced mod 2 on load
    txt = tmp_path / "h.txt"
    txt.write_text("\n".join(" ".join(str(v) for v in row) for row in H) + "\n\n")
    for f in (str(tmp_path / "h.npy"), str(txt)):
        M = pcm.load_matrix(f)
        assert M.dtype == np.int8 and np.array_equal(M, H % 2)


def test_shard_range_and_result_dict():
    from qldpcsim_b200 import simulator
    for shots in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 3, 4, 8):
            r = [simulator.shard_range(shots, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == shots
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    res = simulator.counters_to_result([3, 4, 90, 0, 250, 260, 100, 0], 100)
    assert res == {"DecFailures_X": 3, "DecFailures_Z": 4, "decSuccessExact": 90, "decSuccessDegen": 0,
                   "Avg_number_of_iterations_X": 2.5, "Avg_number_of_iterations_Z": 2.6}
    txt = simulator.format_results([0.05], [res], 100)
    assert "SIMULATION RESULTS" in txt and "1.00e-01" in txt and "    3,    4" in txt


def test_no_cpu_fallback():
    """Without a GPU the decoder entry points must raise, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qldpcsim_b200 import _lib, decoders
    H = pcmlibrary.steane_code()[0]
    with pytest.raises(_lib.QldpcError):
        decoders.NG_decoder(H, np.zeros(3, int))
    with pytest.raises(_lib.QldpcError):
        decoders.Decoder(H, "MS", p=0.01, layers=[np.arange(3)])


@pytest.mark.parametrize("code,k", [("steane", 1), ("shor", 1), ("LP04_0", 19), ("LP118_0", 80), ("bicycle", 36)])
def test_logical_operators(code, k):
    """Bases of the logical operators used by the true outcome classes (extension, README.md:15-22): k = n - rank(Hx) - rank(Hz)
    operators of each type, commuting with the stabilisers of the other type, independent of the stabilisers of their own type,
    and pairing non-degenerately with each other."""
    Hx, Hz = [(h % 2).astype(np.int64) for h in pcmlibrary.by_name(code)]
    Lx, Lz = pcm.logical_operators(Hx, Hz)
    assert Lx.shape == (k, Hx.shape[1]) and Lz.shape == (k, Hx.shape[1])
    assert not ((Hz @ Lx.T) % 2).any() and not ((Hx @ Lz.T) % 2).any()

    def rank(M):
        span = pcm._GF2Span()
        return sum(span.add(v) for v in pcm._rows_to_ints(M))
    assert rank(np.vstack([Hx, Lx])) == rank(Hx) + k and rank(np.vstack([Hz, Lz])) == rank(Hz) + k
    assert rank((Lx @ Lz.T) % 2) == k
    # null space: every basis vector is annihilated, dimension n - rank
    N = pcm._ints_to_rows(pcm.gf2_nullspace(Hz), Hz.shape[1]).astype(np.int64)
    assert not ((Hz @ N.T) % 2).any() and N.shape[0] == Hz.shape[1] - rank(Hz)
