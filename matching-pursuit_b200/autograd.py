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
