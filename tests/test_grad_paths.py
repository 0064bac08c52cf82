"""Gradient-requiring callers of the path: when an input requires grad the drop-ins do not
raise -- the engine finds the events and PyTorch re-evaluates them on fixed indices with the graph attached, or,
where the gradient needs the dense map (the soft-max straight-through term of sparse_feature_map), the loop runs as
PyTorch ops.  Checked against gradients the UNMODIFIED reference produced (tests/golden/grad_*.npz, written by
oracle/make_golden.py): modules/matchingpursuit.py:128-146 (sparse_coding_loss), :229-345 (sparse_code) and
mp.py:50-67 (MatchingPursuit.forward).

The CPU tests exercise the PyTorch formulations directly (indices from the oracle); the GPU tests go through the
drop-in entry points (indices from the CUDA engine)."""
import os

import numpy as np
import pytest
import torch

import matching_pursuit_b200 as mpb
from matching_pursuit_b200 import autograd as ag
from oracle import mp_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(got, want, rel=2e-3):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max() + 1e-12
    assert np.abs(got - want).max() <= rel * scale, (np.abs(got - want).max(), scale)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


# ---------------------------------------------------------------------------- CPU: the PyTorch formulations
def test_dense_pursuit_reproduces_sparse_coding_loss_gradients():
    g = load("grad_sparse_coding_loss")
    d = torch.from_numpy(g["d"]).requires_grad_(True)
    recon = torch.from_numpy(g["recon"]).requires_grad_(True)
    target = torch.from_numpy(g["target"])
    steps = int(g["steps"])
    _, _, _, _, _, r_map = ag.dense_pursuit(recon, d, steps, straight_through=True)
    with torch.no_grad():
        t_map = O.sparse_feature_map(target, d.detach(), steps)
    mx = max(r_map.max().item(), t_map.max().item())
    loss = torch.nn.functional.binary_cross_entropy(r_map / mx, t_map / mx)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    close(recon.grad, g["grad_recon"])
    close(d.grad, g["grad_d"])


def test_dense_pursuit_reproduces_sparse_code_gradients():
    g = load("grad_sparse_code")
    d = torch.from_numpy(g["d"]).requires_grad_(True)
    sig = torch.from_numpy(g["signal"]).requires_grad_(True)
    w = torch.from_numpy(g["w"])
    atom, pos, val, residual, du, _ = ag.dense_pursuit(sig, d, int(g["steps"]))
    loss = (residual ** 2).sum() + ((du[atom] * val[..., None]) * w).sum()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    close(sig.grad, g["grad_signal"])
    close(d.grad, g["grad_d"])


def test_fixed_index_forward_reproduces_mp_gradients():
    g = load("grad_mp_forward")
    atoms = torch.from_numpy(g["atoms"]).requires_grad_(True)
    audio = torch.from_numpy(g["audio"]).requires_grad_(True)
    w = torch.from_numpy(g["w"])
    n, s = audio.shape[-1], w.shape[1]
    # the (atom, time) sequence: from the oracle's restatement of the loop (the engine supplies it on the GPU)
    with torch.no_grad():
        k, a = atoms.shape[1], atoms.shape[2]
        padded = torch.cat([atoms.detach(), torch.zeros(1, k, n - a)], dim=-1)
        residual = audio.detach()
        ks, ts = [], []
        for _ in range(s):
            spec = O.convolve_fft(residual, padded)
            idx = spec.reshape(spec.shape[0], -1).argmax(-1)
            ks.append(idx // n); ts.append(idx % n)
            _, time, atom = O.sparsify2(spec, n_to_keep=1)
            residual = residual - O.convolve_fft(atom @ padded, time)
    ch = ag.fixed_index_forward(atoms, audio, torch.stack(ks, 1), torch.stack(ts, 1), n)
    close(ch.detach(), g["channels"], rel=1e-4)
    loss = (ch * w).sum()
    loss.backward()
    close(atoms.grad, g["grad_atoms"])
    close(audio.grad, g["grad_audio"])


def test_differentiable_correlation_forms_agree():
    sig = O.make_noise_signals(2, 200, seed=3)
    du = O.make_dictionary(5, 24, seed=4)
    close(ag.correlation_map(sig, du), O.correlate_direct(sig, du), rel=1e-5)
    close(ag.correlation_map(sig, du, approx=200), O.correlate_fft(sig, du), rel=1e-5)
    close(ag.correlation_map(sig, du, approx=slice(3, 40)), O.correlate_fft(sig, du, approx=slice(3, 40)), rel=1e-5)


# ---------------------------------------------------------------------------- GPU: through the drop-ins
@pytest.mark.gpu
def test_sparse_coding_loss_dropin_gradients():
    g = load("grad_sparse_coding_loss")
    d = torch.from_numpy(g["d"]).cuda().requires_grad_(True)
    recon = torch.from_numpy(g["recon"]).cuda().requires_grad_(True)
    target = torch.from_numpy(g["target"]).cuda()
    loss = mpb.matchingpursuit.sparse_coding_loss(recon, target, d, n_steps=int(g["steps"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    close(recon.grad.cpu(), g["grad_recon"])
    close(d.grad.cpu(), g["grad_d"])


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["cuda", "cpu"])
def test_sparse_code_dropin_gradients(where):
    g = load("grad_sparse_code")
    d = torch.from_numpy(g["d"]).to(where).requires_grad_(True)
    sig = torch.from_numpy(g["signal"]).to(where).requires_grad_(True)
    w = torch.from_numpy(g["w"]).to(where)
    flat, scatter, residual = mpb.sparse_code(sig, d, n_steps=int(g["steps"]), flatten=True, return_residual=True)
    assert np.array_equal(np.array([(ai, j, int(p)) for ai, j, p, _ in flat]), g["order"])
    loss = (residual ** 2).sum() + sum((a.view(-1) * w).sum() for _, _, _, a in flat)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    close(sig.grad.cpu(), g["grad_signal"])
    close(d.grad.cpu(), g["grad_d"])
    # the array-level form hands back values and residual with the graph attached as well
    atom, pos, val, res = mpb.sparse_code_arrays(sig, d, int(g["steps"]))
    assert val.requires_grad and res.requires_grad and atom.dtype == torch.int32


@pytest.mark.gpu
def test_mp_forward_dropin_gradients():
    g = load("grad_mp_forward")
    m = mpb.mp.MatchingPursuit(n_atoms=8, atom_samples=32, n_samples=256, n_iterations=5).cuda()
    with torch.no_grad():
        m.atoms.copy_(torch.from_numpy(g["atoms"]))
    audio = torch.from_numpy(g["audio"]).cuda().requires_grad_(True)
    w = torch.from_numpy(g["w"]).cuda()
    ch = m.forward(audio)
    close(ch.detach().cpu(), g["channels"], rel=1e-4)
    (ch * w).sum().backward()
    close(m.atoms.grad.cpu(), g["grad_atoms"])
    close(audio.grad.cpu(), g["grad_audio"])
    with torch.no_grad():                                   # the forward-only path gives the same channels
        close(m.forward(audio.detach()).cpu(), g["channels"], rel=1e-4)
