"""Write tests/golden/*.npz by running the UNMODIFIED reference (container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  Every file stores the exact input arrays (so no RNG
has to be reproduced on another machine) and what the reference returned for
them.  The true iteration order and the top-2 margin of every step are
recorded through the reference's own ``visit_key_point`` hook
(modules/matchingpursuit.py:323-324), because ``flatten=True`` output is
grouped by atom.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import mp_oracle as O
from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_sparse_code(ref, signal, d, steps, **kw):
    b, _, n = signal.shape
    seq = []

    def visit(fm, ai, p, a):
        top2 = torch.topk(fm.reshape(-1).double(), 2)[0]
        margin = float((top2[0] - top2[1]) / top2[0].abs().clamp_min(1e-300))
        seq.append((ai, int(p), float(a.norm()), margin, float(fm[ai, int(p)])))

    flat, scatter, residual = ref.matchingpursuit.sparse_code(
        signal, d, n_steps=steps, flatten=True, return_residual=True, visit_key_point=visit, **kw)
    seq = np.array(seq, dtype=np.float64).reshape(steps, b, 5)
    # value of each event = <scaled atom, unit atom>; recover it signed from the returned atoms
    du = ref.normalization.unit_norm(d)
    order = np.array([(ai, j, int(p)) for ai, j, p, a in flat], dtype=np.int64)
    flat_val = np.array([float((a.view(-1) * du[ai]).sum()) for ai, j, p, a in flat], dtype=np.float64)
    recon = scatter(tuple(signal.shape), flat)
    return dict(signal=signal.numpy(), d=d.numpy(), steps=np.int64(steps),
                atom=seq[..., 0].astype(np.int64), pos=seq[..., 1].astype(np.int64),
                absval=seq[..., 2].astype(np.float32), margin=seq[..., 3], val=seq[..., 4].astype(np.float32),
                flat_order=order, flat_val=flat_val.astype(np.float32),
                residual=residual.detach().numpy(), recon=recon.detach().numpy())


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(1234)
    cases = {}

    # 1. white-noise signals, direct conv1d path
    d = O.make_dictionary(16, 64, seed=0)
    sig = O.make_noise_signals(2, 1024, seed=2)
    cases["sc_noise_b2_n1024_k16_a64"] = run_sparse_code(ref, sig, d, 24)
    # 2. same through the reference FFT correlation (approx = N -> full product)
    cases["sc_noise_fftpath"] = run_sparse_code(ref, sig, d, 24, approx=1024)
    # 3. planted atoms, healthy margins
    d = O.make_dictionary(64, 128, seed=0)
    sig = O.make_planted_signals(d, 3, 4096, 24, seed=1)
    cases["sc_planted_b3_n4096_k64_a128"] = run_sparse_code(ref, sig, d, 32)
    # 4. ragged sizes: odd atom count, N not a multiple of anything, un-normalised dictionary
    g = torch.Generator().manual_seed(7)
    d = torch.randn(7, 50, generator=g) * 3.0
    sig = torch.randn(2, 1, 1000, generator=g)
    cases["sc_ragged_b2_n1000_k7_a50"] = run_sparse_code(ref, sig, d, 20)
    # 5. right-edge overhang: energy concentrated in the last samples
    d = O.make_dictionary(8, 32, seed=3)
    sig = torch.zeros(1, 1, 256)
    sig[0, 0, -20:] = torch.linspace(1, 2, 20)
    sig[0, 0, 5:37] += 0.5 * d[2]
    cases["sc_edge_b1_n256_k8_a32"] = run_sparse_code(ref, sig, d, 12)
    # 6. all-zero signal: every step must return (atom 0, position 0, value 0)
    d = O.make_dictionary(4, 16, seed=4)
    cases["sc_zero_b1_n128_k4_a16"] = run_sparse_code(ref, torch.zeros(1, 1, 128), d, 3)
    # 7. atom as long as the signal, and a single atom
    d = O.make_dictionary(1, 64, seed=5)
    sig = O.make_noise_signals(1, 64, seed=6)
    cases["sc_single_atom_full_length"] = run_sparse_code(ref, sig, d, 6)
    # 8. local contrast norm selection (matchingpursuit.py:286-296)
    d = O.make_dictionary(12, 32, seed=8)
    sig = O.make_noise_signals(2, 512, seed=9)
    cases["sc_lcn_b2_n512_k12_a32"] = run_sparse_code(ref, sig, d, 10, local_contrast_norm=True)
    for name, payload in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)

    # correlation helpers
    d = O.make_dictionary(6, 24, seed=10)
    sig = O.make_noise_signals(2, 200, seed=11)
    np.savez_compressed(
        os.path.join(OUT, "corr_helpers.npz"), signal=sig.numpy(), d=d.numpy(),
        torch_conv=ref.conv.torch_conv(sig, d).numpy(),
        fft_full=ref.conv.fft_convolve(sig, d).numpy(),
        fft_slice=ref.conv.fft_convolve(sig, d, approx=slice(3, 40)).numpy(),
        fft_topk=ref.conv.fft_convolve(sig, d, approx=17).numpy())
    a = torch.randn(2, 3, 96, generator=g)
    bb = torch.randn(1, 3, 96, generator=g)
    cc = torch.randn(2, 1, 96, generator=g)
    np.savez_compressed(
        os.path.join(OUT, "fft_convolve_nary.npz"), a=a.numpy(), b=bb.numpy(), c=cc.numpy(),
        two=ref.fft.fft_convolve(a, bb).numpy(), three=ref.fft.fft_convolve(a, bb, cc).numpy(),
        two_ortho=ref.fft.fft_convolve(a, bb, norm="ortho").numpy())

    # sparse_feature_map, sparsify2, soft_dirac
    d = O.make_dictionary(10, 20, seed=12)
    sig = O.make_noise_signals(2, 300, seed=13)
    fm, res = ref.matchingpursuit.sparse_feature_map(sig, d, n_steps=9, return_residual=True)
    x = torch.randn(2, 5, 40, generator=g)
    s2 = ref.sparse.sparsify2(x, n_to_keep=3)
    np.savez_compressed(
        os.path.join(OUT, "feature_map.npz"), signal=sig.numpy(), d=d.numpy(), fm=fm.detach().numpy(),
        residual=res.detach().numpy(), x=x.numpy(), sparse=s2[0].numpy(), packed=s2[1].numpy(),
        context=s2[2].numpy(), soft_dirac=ref.sparse.soft_dirac(x.reshape(2, -1)).numpy())

    # dictionary learning step (caller of the path; stays PyTorch in the product)
    d = O.make_dictionary(9, 24, seed=14)
    sig = O.make_planted_signals(d, 2, 400, 6, seed=15)
    learned = ref.matchingpursuit.dictionary_learning_step(sig, d.clone(), n_steps=8)
    np.savez_compressed(os.path.join(OUT, "dictionary_learning.npz"), signal=sig.numpy(), d=d.numpy(),
                        learned=learned.detach().numpy())

    # band split / merge and multi-band coding
    x = O.make_noise_signals(2, 2048, seed=16)
    split = ref.decompose.fft_frequency_decompose(x, 256)
    merged = ref.decompose.fft_frequency_recompose(split, 2048)
    payload = dict(x=x.numpy(), merged=merged.numpy(), sizes=np.array(list(split.keys())))
    for size, band in split.items():
        payload[f"band_{size}"] = band.numpy()
    np.savez_compressed(os.path.join(OUT, "band_split.npz"), **payload)

    sizes, k, a, steps = [256, 512, 1024, 2048], 16, 32, 6
    specs = []
    for i, size in enumerate(sizes):
        spec = ref.multibanddict.BandSpec(size, k, a, device=torch.device("cpu"),
                                          signal_samples=2048, is_lowest_band=(i == 0))
        spec.d = O.make_dictionary(k, a, seed=20 + i)
        specs.append(spec)
    model = ref.multibanddict.MultibandDictionaryLearning(specs, n_samples=2048)
    enc = model.encode(x, steps)
    flat = model.flattened_event_tuples(enc)
    hier = model.hierarchical_event_tuples(flat, enc)
    decoded = model.decode(hier)
    recon, _ = model.recon(x, steps)
    payload = dict(x=x.numpy(), sizes=np.array(sizes), k=np.int64(k), a=np.int64(a), steps=np.int64(steps),
                   decoded=decoded.detach().numpy(), recon=recon.detach().numpy(),
                   flat_atom=np.array([e[0] for e in flat]), flat_batch=np.array([e[1] for e in flat]),
                   flat_time=np.array([float(e[2]) for e in flat]),
                   flat_amp=np.array([float(e[3]) for e in flat], dtype=np.float32))
    for i, size in enumerate(sizes):
        payload[f"d_{size}"] = specs[i].d.numpy()
        payload[f"events_{size}"] = np.array([(ai, j, int(p)) for ai, j, p, _ in enc[size][0]])
    np.savez_compressed(os.path.join(OUT, "multiband.npz"), **payload)

    # mp.py forward (class body executed from the reference file)
    m = ref.MatchingPursuit(n_atoms=8, atom_samples=32, n_samples=256, n_iterations=5)
    audio = O.make_noise_signals(2, 256, seed=30)
    with torch.no_grad():
        ch = m.forward(audio)
    np.savez_compressed(os.path.join(OUT, "mp_forward.npz"), atoms=m.atoms.detach().numpy(),
                        audio=audio.numpy(), channels=ch.numpy())
    # gradients through the reference's own code (callers that train through the path): sparse_coding_loss
    # (modules/matchingpursuit.py:128-146), sparse_code with grad (:229-345) and mp.py forward (:50-67)
    gg = torch.Generator().manual_seed(41)
    d = (torch.randn(10, 20, generator=gg) * 0.7).requires_grad_(True)
    target = O.make_planted_signals(O.unit_norm(d.detach()), 2, 300, 5, seed=42)
    recon = (target + 0.1 * torch.randn(2, 1, 300, generator=gg)).requires_grad_(True)
    loss = ref.matchingpursuit.sparse_coding_loss(recon, target, d, n_steps=6)
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "grad_sparse_coding_loss.npz"), d=d.detach().numpy(),
                        target=target.numpy(), recon=recon.detach().numpy(), steps=np.int64(6),
                        loss=np.float64(loss.item()), grad_recon=recon.grad.numpy(), grad_d=d.grad.numpy())

    d = (torch.randn(12, 32, generator=gg)).requires_grad_(True)
    sig = O.make_planted_signals(O.unit_norm(d.detach()), 2, 512, 6, seed=43).requires_grad_(True)
    w = torch.randn(32, generator=gg)
    flat, scatter, residual = ref.matchingpursuit.sparse_code(sig, d, n_steps=8, flatten=True, return_residual=True)
    loss = (residual ** 2).sum() + sum((a.view(-1) * w).sum() for _, _, _, a in flat)
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "grad_sparse_code.npz"), d=d.detach().numpy(), signal=sig.detach().numpy(),
                        w=w.numpy(), steps=np.int64(8), loss=np.float64(loss.item()),
                        order=np.array([(ai, j, int(p)) for ai, j, p, _ in flat]),
                        grad_signal=sig.grad.numpy(), grad_d=d.grad.numpy())

    m = ref.MatchingPursuit(n_atoms=8, atom_samples=32, n_samples=256, n_iterations=5)
    with torch.no_grad():
        m.atoms.copy_(torch.randn(1, 8, 32, generator=gg) * 0.3)
    audio = O.make_noise_signals(2, 256, seed=44).requires_grad_(True)
    w = torch.randn(2, 5, 256, generator=gg)
    ch = m.forward(audio)
    loss = (ch * w).sum()
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "grad_mp_forward.npz"), atoms=m.atoms.detach().numpy(),
                        audio=audio.detach().numpy(), w=w.numpy(), channels=ch.detach().numpy(),
                        loss=np.float64(loss.item()), grad_atoms=m.atoms.grad.numpy(), grad_audio=audio.grad.numpy())
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
