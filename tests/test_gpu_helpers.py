"""GPU parity of the helpers either side of the pursuit -- FFT convolution
(modules/fft.py), correlation helpers (modules/conv.py), band split / merge
(modules/decompose.py), the multi-band codec (modules/multibanddict.py),
sparsify2 (modules/sparse.py) and mp.py's forward -- against the golden
vectors the unmodified reference produced and against the CPU oracle.
Floating point: relative 1e-4 of the result's scale unless stated."""
import os

import numpy as np
import pytest
import torch

import matching_pursuit_b200 as mpb
from matching_pursuit_b200 import conv as mconv, decompose as mdec, fft as mfft, mp as mmp
from oracle import mp_oracle as O
from parity import MARGIN

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


def close(got, want, rel=1e-4):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= rel * scale, (np.abs(got - want).max(), scale)


def test_golden_nary_fft_convolution():
    g = np.load(os.path.join(GOLDEN, "fft_convolve_nary.npz"))
    a, b, c = (torch.from_numpy(g[k]).to(DEV) for k in "abc")
    close(mfft.fft_convolve(a, b).cpu(), g["two"])
    close(mfft.fft_convolve(a, b, c).cpu(), g["three"])             # 3 operands alias at period 2n, as the reference
    close(mfft.fft_convolve(a, b, norm="ortho").cpu(), g["two_ortho"])
    out = mfft.fft_convolve(a.cpu(), b.cpu())                        # host tensors in -> host tensors out
    assert not out.is_cuda
    close(out, g["two"])


@pytest.mark.parametrize("n", [256, 1000, 4096, 2 ** 15, 2 ** 16])
def test_fft_convolution_against_oracle(n):
    g = torch.Generator().manual_seed(n)
    a = torch.randn(2, 1, n, generator=g)
    b = torch.randn(1, 3, n, generator=g) * torch.exp(-torch.arange(n) / 200.0)
    want = O.convolve_fft(a.double(), b.double())
    close(mfft.fft_convolve(a.to(DEV), b.to(DEV)).cpu(), want, rel=2e-5)
    want_corr = torch.fft.irfft(torch.fft.rfft(torch.nn.functional.pad(a.double(), (0, n))) *
                                torch.conj(torch.fft.rfft(torch.nn.functional.pad(b.double(), (0, n)))))[..., :n]
    close(mfft.fft_correlate(a.to(DEV), b.to(DEV)).cpu(), want_corr, rel=2e-5)


def test_golden_correlation_helpers():
    g = np.load(os.path.join(GOLDEN, "corr_helpers.npz"))
    sig, d = torch.from_numpy(g["signal"]).to(DEV), torch.from_numpy(g["d"]).to(DEV)
    close(mconv.torch_conv(sig, d).cpu(), g["torch_conv"], rel=1e-5)
    close(mconv.fft_convolve(sig, d).cpu(), g["fft_full"], rel=1e-5)
    close(mconv.fft_convolve(sig, d, approx=200).cpu(), g["fft_full"], rel=1e-5)
    # atoms are used as given, not normalised
    close(mconv.torch_conv(sig, 3.0 * d).cpu(), 3.0 * g["torch_conv"], rel=1e-5)
    # band-limited product over the bins of the length N+A transform (modules/conv.py:24-29)
    close(mconv.fft_convolve(sig, d, approx=slice(3, 40)).cpu(), g["fft_slice"], rel=1e-5)
    with pytest.raises(NotImplementedError):
        mconv.fft_convolve(sig, d, approx=17)       # defective in the reference


@pytest.mark.parametrize("slce", [slice(3, 40), slice(None, None, 3), slice(-20, None), slice(0, None),
                                  slice(7, 7)], ids=str)
def test_band_limit_against_torch_fft(slce):
    """mpb200_band_limit == irfft(mask(rfft(pad(x, L)))) for an even L that is not a power of two."""
    torch.manual_seed(5)
    n, a = 3000, 100
    x = torch.randn(3, n)
    spec = torch.fft.rfft(torch.nn.functional.pad(x.double(), (0, a)), dim=-1)
    kept = torch.zeros_like(spec)
    kept[..., slce] = spec[..., slce]
    want = torch.fft.irfft(kept, n=n + a, dim=-1)
    got = mpb.engine.band_limit(x.to(DEV), n + a, slce).cpu().double()
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))
    with pytest.raises(mpb.MpbError):
        mpb.engine.band_limit(x.to(DEV), n + a + 1, slce)     # odd transform length
    with pytest.raises(ValueError):
        mpb.engine.band_limit(x.to(DEV), n + a, slice(100, 5, -2))    # torch rejects such a slice in the reference too


def test_sparse_code_band_limited_against_oracle():
    """sparse_code(approx=slice): the reference's per-step band-limited FFT correlation
    (modules/matchingpursuit.py:278-280 -> modules/conv.py:24-29)."""
    from parity import compare_with_oracle_trace
    k, a, n, b, s = 24, 64, 2048, 2, 12
    d = O.make_dictionary(k, a, seed=21)
    sig = O.make_planted_signals(d, b, n, 8, seed=22)
    slce = slice(0, 400)
    tr = O.greedy_pursuit(sig, d, s, approx=slce, want_margin=True)
    flat, scatter, residual = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, approx=slce, flatten=True,
                                              return_residual=True)
    atom, pos, val = mpb.sparse_code_arrays(sig.to(DEV), d.to(DEV), s, approx=slce)[:3]
    checked = compare_with_oracle_trace(tr, atom.cpu().numpy(), pos.cpu().numpy(), val.cpu().numpy(),
                                        residual.cpu().numpy()[:, 0])
    assert checked > 0


def test_golden_band_split_and_merge():
    g = np.load(os.path.join(GOLDEN, "band_split.npz"))
    x = torch.from_numpy(g["x"]).to(DEV)
    split = mdec.fft_frequency_decompose(x, 256)
    assert list(split.keys()) == list(g["sizes"])
    for size, band in split.items():
        assert band.shape == (2, 1, size)
        close(band.cpu(), g[f"band_{size}"], rel=2e-5)
    close(mdec.fft_frequency_recompose(split, 2048).cpu(), g["merged"], rel=2e-5)


def test_band_split_large_against_oracle():
    x = O.make_noise_signals(2, 2 ** 16, seed=3)
    want = O.band_split(x, 2048)
    got = mdec.fft_frequency_decompose(x.to(DEV), 2048)
    assert list(got.keys()) == list(want.keys()) == [2048, 4096, 8192, 16384, 32768, 65536]
    for size in want:
        close(got[size].cpu(), want[size], rel=2e-5)
    close(mdec.fft_frequency_recompose(got, 2 ** 16).cpu(), O.band_merge(want, 2 ** 16), rel=2e-5)


def test_golden_sparsify2():
    g = np.load(os.path.join(GOLDEN, "feature_map.npz"))
    x = torch.from_numpy(g["x"]).to(DEV)
    sparse, packed, context = mpb.sparsify2(x, n_to_keep=3)
    assert np.array_equal(sparse.cpu().numpy(), g["sparse"])
    assert np.array_equal(packed.cpu().numpy(), g["packed"])
    assert np.array_equal(context.cpu().numpy(), g["context"])


def test_golden_multiband_codec():
    g = np.load(os.path.join(GOLDEN, "multiband.npz"))
    x = torch.from_numpy(g["x"]).to(DEV)
    sizes = [int(s) for s in g["sizes"]]
    k, a, steps = int(g["k"]), int(g["a"]), int(g["steps"])
    specs = []
    for i, size in enumerate(sizes):
        spec = mpb.BandSpec(size, k, a, device=DEV, signal_samples=2048, is_lowest_band=(i == 0))
        spec.d = torch.from_numpy(g[f"d_{size}"]).to(DEV)
        specs.append(spec)
    model = mpb.MultibandDictionaryLearning(specs, n_samples=2048)
    # margins of the per-band pursuits, from the oracle
    split = O.band_split(torch.from_numpy(g["x"]), 256)
    safe = all((O.greedy_pursuit(split[s], torch.from_numpy(g[f"d_{s}"]), steps, want_margin=True).margin.numpy()
                > MARGIN).all() for s in sizes)
    enc = model.encode(x, steps)
    assert list(enc.keys()) == sizes
    if not safe:
        pytest.skip("an ambiguous step in one of the bands")
    for size in sizes:
        got = np.array([(ai, j, int(p)) for ai, j, p, _ in enc[size][0]])
        assert np.array_equal(got, g[f"events_{size}"]), size
    flat = model.flattened_event_tuples(enc)
    assert np.array_equal(np.array([e[0] for e in flat]), g["flat_atom"])
    assert np.array_equal(np.array([e[1] for e in flat]), g["flat_batch"])
    np.testing.assert_allclose(np.array([float(e[2]) for e in flat]), g["flat_time"])
    np.testing.assert_allclose(np.array([float(e[3]) for e in flat]), g["flat_amp"], rtol=1e-4)
    decoded = model.decode(model.hierarchical_event_tuples(flat, enc))
    close(decoded.cpu(), g["decoded"], rel=1e-4)
    recon, events = model.recon(x, steps)
    close(recon.cpu(), g["recon"], rel=1e-4)
    assert list(events.keys()) == sizes


def test_golden_mp_forward():
    g = np.load(os.path.join(GOLDEN, "mp_forward.npz"))
    m = mmp.MatchingPursuit(n_atoms=8, atom_samples=32, n_samples=256, n_iterations=5).to(DEV)
    with torch.no_grad():
        m.atoms.copy_(torch.from_numpy(g["atoms"]))
        ch = m(torch.from_numpy(g["audio"]).to(DEV))
    assert ch.shape == (2, 5, 256)
    close(ch.cpu(), g["channels"], rel=1e-4)
    ch2 = m(torch.from_numpy(g["audio"]).to(DEV))        # grad enabled: same values, graph attached (fixed-index form)
    assert ch2.requires_grad
    close(ch2.detach().cpu(), g["channels"], rel=1e-4)


def test_multiband_config4_shapes():
    """BASELINE configs[3]: 6 bands x 1024 atoms on 2^16-sample signals (A=128, 16 steps here)."""
    n, k, a, steps, b = 2 ** 16, 1024, 128, 16, 2
    sizes = [2048 * 2 ** i for i in range(6)]
    specs = [mpb.BandSpec(s, k, a, device=DEV, signal_samples=n, is_lowest_band=(i == 0))
             for i, s in enumerate(sizes)]
    g = torch.Generator().manual_seed(0)
    for i, spec in enumerate(specs):
        spec.d = O.make_dictionary(k, a, seed=10 + i).to(DEV)
    model = mpb.MultibandDictionaryLearning(specs, n_samples=n)
    x = O.make_noise_signals(b, n, seed=5).to(DEV)
    enc = model.encode(x, steps)
    assert all(len(enc[s][0]) == b * steps for s in sizes)
    recon, events = model.recon(x, steps)
    assert recon.shape == (b, 1, n)
    # the coded energy can only lower the residual of each band: ||band - recon_band|| <= ||band||
    split = mdec.fft_frequency_decompose(x, 2048)
    for s in sizes:
        r, _, _ = specs[sizes.index(s)].recon(split[s], steps)
        assert float((split[s] - r).norm()) < float(split[s].norm())


def test_differentiable_reevaluation_matches_reference_gradients():
    """Gradients of a loss on (values, residual) with respect to the dictionary and the signal: the CUDA pursuit
    + PyTorch re-evaluation on fixed indices against the reference's dense loop restated with CPU autograd
    (modules/matchingpursuit.py:269-328: conv1d map, torch.max over the flattened map, scatter, subtract)."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    k, a, n, b, s = 12, 32, 512, 2, 10
    d0 = torch.zeros(k, a).uniform_(-1, 1)
    sig0 = O.make_planted_signals(O.unit_norm(d0), b, n, 6, seed=9)

    def reference_loss(sig, d):
        du = d / (torch.norm(d, dim=-1, keepdim=True) + 1e-8)
        residual, vals, seq = sig.clone(), [], []
        for _ in range(s):
            fm = F.conv1d(F.pad(residual, (0, a)), du.view(k, 1, a))[..., :n]
            value, index = torch.max(fm.reshape(b, -1), dim=-1)
            atom, pos = index // n, index % n
            sparse = torch.zeros(b, 1, n + a)
            for j in range(b):
                sparse[j, 0, pos[j]: pos[j] + a] = du[atom[j]] * value[j]
            residual = residual - sparse[..., :n]
            vals.append(value)
            seq.append((atom.clone(), pos.clone()))
        return (residual ** 2).sum() + torch.stack(vals).sum(), seq

    sig_r, d_r = sig0.clone().requires_grad_(True), d0.clone().requires_grad_(True)
    loss_r, seq = reference_loss(sig_r, d_r)
    loss_r.backward()

    sig_g, d_g = sig0.clone().to(DEV).requires_grad_(True), d0.clone().to(DEV).requires_grad_(True)
    atom, pos, val, residual = mpb.autograd.sparse_code_differentiable(sig_g, d_g, s)
    for step, (ra, rp) in enumerate(seq):
        assert torch.equal(atom[:, step].cpu(), ra) and torch.equal(pos[:, step].cpu(), rp)
    loss_g = (residual ** 2).sum() + val.sum()
    loss_g.backward()
    assert abs(float(loss_g.detach()) - float(loss_r.detach())) <= 1e-4 * abs(float(loss_r.detach()))
    close(d_g.grad.cpu(), d_r.grad, rel=1e-3)
    close(sig_g.grad.cpu(), sig_r.grad, rel=1e-3)
