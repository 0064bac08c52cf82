"""The oracle (CPU restatement) must reproduce what the UNMODIFIED reference
returned for the committed inputs (tests/golden, written by
oracle/make_golden.py in the build container).  This is what pins the oracle.

The comparison is exact where the same torch build ran both (sequence and
values), and margin-aware otherwise: the reference's conv1d/FFT kernels may
round differently on another host CPU, so a step whose recorded top-2 margin is
below 1e-5 is allowed to differ and ends the sequence comparison for that
signal (BASELINE.json north_star parity rule)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mp_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SC_CASES = sorted(glob.glob(os.path.join(GOLDEN, "sc_*.npz")))
MARGIN = 1e-5
RTOL = 1e-4


def compare_trace(g, atom, pos, val, residual):
    """atom/pos/val are (S,B) from the implementation under test."""
    steps, batch = g["atom"].shape
    for j in range(batch):
        for s in range(steps):
            if g["margin"][s, j] <= MARGIN:
                break  # legitimately ambiguous: sequences may diverge from here
            assert (int(atom[s, j]), int(pos[s, j])) == (int(g["atom"][s, j]), int(g["pos"][s, j])), (s, j)
            assert abs(abs(float(val[s, j])) - float(g["absval"][s, j])) <= RTOL * max(float(g["absval"][s, j]), 1e-12)
        else:
            e_ref = float((g["residual"][j].astype(np.float64) ** 2).sum())
            e_new = float((np.asarray(residual[j], dtype=np.float64) ** 2).sum())
            assert abs(e_new - e_ref) <= RTOL * max(e_ref, 1e-12)


@pytest.mark.parametrize("path", SC_CASES, ids=[os.path.basename(p)[:-4] for p in SC_CASES])
def test_sparse_code_matches_reference(path):
    g = np.load(path)
    name = os.path.basename(path)
    kw = {}
    if "fftpath" in name:
        kw["approx"] = int(g["signal"].shape[-1])
    if "lcn" in name:
        kw["local_contrast_norm"] = True
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    tr = O.greedy_pursuit(sig, d, int(g["steps"]), **kw)
    compare_trace(g, tr.atom.numpy(), tr.pos.numpy(), tr.val.numpy(), tr.residual.numpy())
    # grouped/flattened return order (first-seen atom order) and decode
    flat, scatter, residual = O.sparse_code(sig, d, int(g["steps"]), flatten=True, return_residual=True, **kw)
    if (g["margin"] > MARGIN).all():
        order = np.array([(ai, j, int(p)) for ai, j, p, a in flat], dtype=np.int64)
        assert np.array_equal(order, g["flat_order"])
        recon = scatter(tuple(sig.shape), flat)
        np.testing.assert_allclose(recon.numpy(), g["recon"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(residual.numpy(), g["residual"], rtol=1e-5, atol=1e-6)


def test_zero_signal_picks_first_atom_first_position():
    g = np.load(os.path.join(GOLDEN, "sc_zero_b1_n128_k4_a16.npz"))
    assert (g["atom"] == 0).all() and (g["pos"] == 0).all() and (g["absval"] == 0).all()
    tr = O.greedy_pursuit(torch.from_numpy(g["signal"]), torch.from_numpy(g["d"]), 3)
    assert (tr.atom == 0).all() and (tr.pos == 0).all() and (tr.val == 0).all()


def test_correlation_helpers():
    g = np.load(os.path.join(GOLDEN, "corr_helpers.npz"))
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    np.testing.assert_allclose(O.correlate_direct(sig, d).numpy(), g["torch_conv"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.correlate_fft(sig, d).numpy(), g["fft_full"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.correlate_fft(sig, d, approx=slice(3, 40)).numpy(), g["fft_slice"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.correlate_fft(sig, d, approx=17).numpy(), g["fft_topk"], rtol=1e-5, atol=1e-6)
    # the two correlation forms agree with each other (reference's compare_conv, modules/conv.py:75-84)
    np.testing.assert_allclose(g["fft_full"], g["torch_conv"], atol=2e-5)


def test_nary_fft_convolution():
    g = np.load(os.path.join(GOLDEN, "fft_convolve_nary.npz"))
    a, b, c = (torch.from_numpy(g[k]) for k in "abc")
    np.testing.assert_allclose(O.convolve_fft(a, b).numpy(), g["two"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.convolve_fft(a, b, c).numpy(), g["three"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(O.convolve_fft(a, b, norm="ortho").numpy(), g["two_ortho"], rtol=1e-5, atol=1e-5)


def test_feature_map_and_selection_helpers():
    g = np.load(os.path.join(GOLDEN, "feature_map.npz"))
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    fm, res = O.sparse_feature_map(sig, d, n_steps=9, return_residual=True)
    np.testing.assert_allclose(fm.numpy(), g["fm"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res.numpy(), g["residual"], rtol=1e-5, atol=1e-6)
    x = torch.from_numpy(g["x"])
    sparse, packed, context = O.sparsify2(x, n_to_keep=3)
    assert np.array_equal(sparse.numpy(), g["sparse"])
    assert np.array_equal(packed.numpy(), g["packed"])
    assert np.array_equal(context.numpy(), g["context"])
    np.testing.assert_allclose(O.soft_dirac(x.reshape(2, -1)).numpy(), g["soft_dirac"], atol=1e-7)


def test_dictionary_learning_step():
    g = np.load(os.path.join(GOLDEN, "dictionary_learning.npz"))
    learned = O.dictionary_learning_step(torch.from_numpy(g["signal"]), torch.from_numpy(g["d"]).clone(), n_steps=8)
    np.testing.assert_allclose(learned.numpy(), g["learned"], rtol=1e-5, atol=1e-6)


def test_band_split_and_merge():
    g = np.load(os.path.join(GOLDEN, "band_split.npz"))
    x = torch.from_numpy(g["x"])
    split = O.band_split(x, 256)
    assert list(split.keys()) == list(g["sizes"])
    for size, band in split.items():
        np.testing.assert_allclose(band.numpy(), g[f"band_{size}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.band_merge(split, 2048).numpy(), g["merged"], rtol=1e-5, atol=1e-6)


def test_multiband_codec():
    g = np.load(os.path.join(GOLDEN, "multiband.npz"))
    x = torch.from_numpy(g["x"])
    sizes = [int(s) for s in g["sizes"]]
    specs = [O.BandOracle(size, torch.from_numpy(g[f"d_{size}"]), is_lowest_band=(i == 0))
             for i, size in enumerate(sizes)]
    model = O.MultibandOracle(specs, n_samples=2048)
    enc = model.encode(x, int(g["steps"]))
    for size in sizes:
        got = np.array([(ai, j, int(p)) for ai, j, p, _ in enc[size][0]])
        assert np.array_equal(got, g[f"events_{size}"]), size
    flat = model.flattened_event_tuples(enc)
    assert np.array_equal(np.array([e[0] for e in flat]), g["flat_atom"])
    assert np.array_equal(np.array([e[1] for e in flat]), g["flat_batch"])
    np.testing.assert_allclose(np.array([float(e[2]) for e in flat]), g["flat_time"])
    np.testing.assert_allclose(np.array([float(e[3]) for e in flat]), g["flat_amp"], rtol=1e-5)
    decoded = model.decode(model.hierarchical_event_tuples(flat, enc))
    np.testing.assert_allclose(decoded.numpy(), g["decoded"], rtol=1e-4, atol=1e-5)
    recon, _ = model.recon(x, int(g["steps"]))
    np.testing.assert_allclose(recon.numpy(), g["recon"], rtol=1e-4, atol=1e-5)


def test_mp_forward():
    g = np.load(os.path.join(GOLDEN, "mp_forward.npz"))
    ch = O.mp_forward(torch.from_numpy(g["atoms"]), torch.from_numpy(g["audio"]), 256, 5)
    np.testing.assert_allclose(ch.numpy(), g["channels"], rtol=1e-5, atol=1e-7)


# --------------------------------------------------------------------------
# headline shapes: inputs regenerate from their seeds; the oracle follows the reference there too
# --------------------------------------------------------------------------
def _headline_prefix(name, steps):
    """First `steps` iterations of the oracle at a headline shape against the reference's recorded trace."""
    from headline_inputs import load_single
    g, d, sig = load_single(name)
    tr = O.greedy_pursuit(sig, d, steps)
    assert np.array_equal(tr.atom.numpy(), g["atom"][:steps]) and np.array_equal(tr.pos.numpy(), g["pos"][:steps])
    np.testing.assert_allclose(tr.val.numpy(), g["val"][:steps], rtol=1e-5)
    return g, tr


@pytest.mark.parametrize("name,steps", [("hl_c2_k512_a1024_n32768_b4_s64", 4), ("hl_long_k64_a4096_n32768_b2_s24", 24),
                                        ("hl_long_k32_a8192_n32768_b2_s24", 24), ("hl_long_k16_a16384_n65536_b1_s16", 16)])
def test_headline_inputs_and_oracle_prefix(name, steps):
    g, tr = _headline_prefix(name, steps)
    if steps == int(g["steps"]):
        np.testing.assert_allclose(tr.residual.numpy().reshape(g["residual"].shape), g["residual"], rtol=1e-5, atol=1e-6)


def test_headline_inputs_regenerate():
    """The two large cases are not re-run on the CPU (minutes): their inputs must regenerate bit-for-bit, and the
    recorded traces must be binding (every step above the 1e-5 margin)."""
    from headline_inputs import HEADLINE_SINGLE, load_multiband, load_single
    for name in HEADLINE_SINGLE:
        g, d, sig = load_single(name)
        assert (g["margin"] > MARGIN).mean() >= 0.9
    g, x, dicts, bands = load_multiband()
    for size in (int(s) for s in g["sizes"]):
        assert (g[f"margin_{size}"] > MARGIN).mean() >= 0.9
