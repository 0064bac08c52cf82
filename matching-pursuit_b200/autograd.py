"""Differentiable re-evaluation of a pursuit ("dictionary learning stays in PyTorch", BASELINE.json).

The greedy selection is not differentiable in the reference either: ``torch.max`` over the flattened map
(modules/matchingpursuit.py:298-303) passes gradient only to the selected entry
``fm[b, k, p] = sum_i pad(residual)[b, p + i] * d_unit[k, i]``.  So the CUDA engine finds the events
``(atom, position)`` without autograd, and the values, the scaled atoms and the residual are then recomputed with
ordinary PyTorch ops on those FIXED indices -- S small gathers/scatters of A samples per signal instead of S dense
correlations -- which reproduces the reference's forward values and its gradients with respect to the signal
and the dictionary (callers: ``sparse_coding_loss`` :128-146, ``mp.py`` training).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .engine import Plan
from .matchingpursuit import sparse_code_arrays


def sparse_code_differentiable(signal: torch.Tensor, d: torch.Tensor, n_steps: int = 100, *, mode: str = "auto",
                               plan: Optional[Plan] = None):
    """``signal`` (B,1,N), ``d`` (K,A); either may require grad.  Returns ``(atom int64 (B,S), pos int64 (B,S),
    val (B,S), residual (B,1,N))``; ``val`` and ``residual`` carry the autograd graph."""
    b, _, n = signal.shape
    k, a = d.shape[0], d.shape[-1]
    with torch.no_grad():
        atom, pos, _, _ = sparse_code_arrays(signal.detach(), d.detach().reshape(k, a), n_steps, mode=mode, plan=plan)
    atom, pos = atom.to(signal.device).long(), pos.to(signal.device).long()
    d2 = d.reshape(k, a)
    du = d2 / (torch.norm(d2, dim=-1, keepdim=True) + 1e-8)            # modules/normalization.py:4-6
    residual = signal.reshape(b, n)
    offs = torch.arange(a, device=signal.device)
    vals = []
    for s in range(n_steps):
        idx = pos[:, s, None] + offs                                     # (B, A); beyond N: zero padding (:275)
        seg = torch.gather(F.pad(residual, (0, a)), 1, idx)
        atoms_k = du[atom[:, s]]                                         # (B, A)
        v = (seg * atoms_k).sum(-1)                                      # fm[b, k, p]                     (:277, 299)
        upd = torch.zeros(b, n + a, device=signal.device, dtype=residual.dtype).scatter_add(1, idx, v[:, None] * atoms_k)
        residual = residual - upd[:, :n]                                 # truncated at the right edge     (:33-56, 328)
        vals.append(v)
    val = torch.stack(vals, dim=1) if vals else residual.new_zeros(b, 0)
    return atom, pos, val, residual.view(b, 1, n)


# --------------------------------------------------------------------------
# Dense differentiable loop: callers that need gradients through something only the dense map can give them
# (the soft-max straight-through term of sparse_feature_map, per-step callbacks, a caller-supplied map, LCN or a
# band-limited map under autograd).  Plain PyTorch ops on the caller's device with the autograd graph attached --
# "dictionary learning stays in PyTorch" (BASELINE.json); SURVEY.md 8b prescribes exactly this dispatch.
# --------------------------------------------------------------------------
def correlation_map(residual: torch.Tensor, du: torch.Tensor, approx=None) -> torch.Tensor:
    """Differentiable fm[b,k,t] = sum_i pad(residual)[b,t+i] * du[k,i] (modules/matchingpursuit.py:275-277); with
    ``approx`` the FFT form of modules/conv.py:11-53: circular product at length N+A with the reversed padded
    atoms, optional bin mask, rolled by one and cropped."""
    b, _, n = residual.shape
    k, a = du.shape
    if approx is None:
        return F.conv1d(F.pad(residual, (0, a)), du.view(k, 1, a))[..., :n]
    if isinstance(approx, int) and not isinstance(approx, bool) and approx < n:
        raise NotImplementedError("approx=int<N is defective in the reference and is not part of the engine")
    length = n + a
    sig_spec = torch.fft.rfft(F.pad(residual, (0, a)), dim=-1)
    atom_spec = torch.fft.rfft(torch.flip(F.pad(du, (0, length - a)), dims=(-1,)), dim=-1)[None]
    if isinstance(approx, slice):
        keep = torch.zeros(sig_spec.shape[-1], device=residual.device, dtype=torch.bool)
        keep[approx] = True
        prod = torch.where(keep, sig_spec * atom_spec, torch.zeros((), dtype=sig_spec.dtype, device=residual.device))
    else:
        prod = sig_spec * atom_spec
    return torch.roll(torch.fft.irfft(prod, n=length, dim=-1), 1, dims=-1)[..., :n]


def dense_pursuit(signal: torch.Tensor, d: torch.Tensor, n_steps: int, *, approx=None, local_contrast_norm=False,
                  compute_feature_map=None, on_map=None, on_select=None, straight_through=False):
    """The reference loop with the graph attached.  Returns ``(atom int64 (B,S), pos int64 (B,S), val (B,S),
    residual (B,1,N), du (K,A), accumulated)``; ``accumulated`` is the dense sum of ``soft_dirac(f) * f`` over the
    steps when ``straight_through`` (modules/matchingpursuit.py:100-101), else None."""
    b, _, n = signal.shape
    k, a = d.shape[0], d.shape[-1]
    d2 = d.reshape(k, a)
    du = d2 / (torch.norm(d2, dim=-1, keepdim=True) + 1e-8)                 # modules/normalization.py:4-6
    residual = signal.reshape(b, 1, n)
    offs = torch.arange(a, device=signal.device)
    atoms, poss, vals = [], [], []
    acc = torch.zeros(b, k, n, device=signal.device, dtype=signal.dtype) if straight_through else None
    for step in range(n_steps):
        fm = compute_feature_map(residual, du) if compute_feature_map is not None else \
            correlation_map(residual, du, approx)
        if on_map is not None:
            on_map(step, fm, du)
        flat = fm.reshape(b, -1)
        if straight_through:
            soft = torch.softmax(flat, dim=-1)
            hard = torch.zeros_like(soft).scatter(-1, soft.argmax(dim=-1, keepdim=True), 1.0)
            acc = acc + ((soft + (hard - soft).detach()) * flat).view(b, k, n)
        sel = flat
        if local_contrast_norm:                                              # :286-296
            plane = fm.view(b, 1, k, n)
            sel = (plane - F.avg_pool2d(plane, (9, 9), (1, 1), (4, 4))).reshape(b, -1)
        idx = sel.argmax(dim=-1, keepdim=True)
        v = torch.gather(flat, -1, idx).view(b)                              # gradient reaches the selected entry only
        ai, pp = (idx // n).view(b), (idx % n).view(b)
        scaled = du[ai] * v[:, None]
        if on_select is not None:
            on_select(step, fm, ai, pp, v, du)
        where = pp[:, None] + offs                                           # beyond N: truncated (:33-56)
        upd = torch.zeros(b, n + a, device=signal.device, dtype=residual.dtype).scatter_add(1, where, scaled)
        residual = residual - upd[:, None, :n]
        atoms.append(ai); poss.append(pp); vals.append(v)
    stack = (lambda xs, dt: torch.stack(xs, dim=1) if xs else torch.zeros(b, 0, device=signal.device, dtype=dt))
    return (stack(atoms, torch.int64), stack(poss, torch.int64), stack(vals, signal.dtype), residual, du, acc)


def fixed_index_forward(atoms: torch.Tensor, audio: torch.Tensor, k_idx: torch.Tensor, t_idx: torch.Tensor,
                        n_samples: int) -> torch.Tensor:
    """Differentiable re-evaluation of ``mp.py::MatchingPursuit.forward`` (mp.py:50-67) on the (atom, time)
    indices the engine found: per step the selected entry of the zero-padded convolution map
    ``v = sum_i atoms[k0, i] * residual[t0 - i]`` and the subtracted channel ``v^2 * atoms[k0]`` placed at t0 and
    truncated at N -- A multiply-adds per signal and step instead of two dense FFT convolutions.
    ``atoms`` (1,K,A) (the parameter), ``audio`` (B,1,N), ``k_idx``/``t_idx`` int64 (B,S).  Returns (B,S,N)."""
    b = audio.shape[0]
    a = atoms.shape[-1]
    n = n_samples
    table = atoms.reshape(-1, a)
    residual = audio.reshape(b, n)
    offs = torch.arange(a, device=audio.device)
    out = []
    for s in range(k_idx.shape[1]):
        k0, t0 = k_idx[:, s], t_idx[:, s]
        src = t0[:, None] - offs                                             # residual sample under atom tap i
        seg = torch.gather(residual, 1, src.clamp_min(0)) * (src >= 0)
        atom = table[k0]                                                     # (B, A)
        v = (seg * atom).sum(-1)
        where = t0[:, None] + offs
        step = torch.zeros(b, n + a, device=audio.device, dtype=residual.dtype).scatter_add(
            1, where, (v * v)[:, None] * atom)[:, :n]
        residual = residual - step
        out.append(step)
    return torch.stack(out, dim=1) if out else audio.new_zeros(b, 0, n)
