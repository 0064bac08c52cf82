"""Counterpart of ``mp.py::MatchingPursuit`` (reference lines 32-67): the
differentiable MP-like loop whose feature map is a zero-padded CONVOLUTION of
the residual with the raw (un-normalised, mp.py:43-48) atoms, whose top-1 entry
picks (atom k0, time t0, value v), and whose subtracted channel is
``conv(v*atom_k0, v*delta_t0)`` = ``v^2 * atom_k0`` shifted to t0 and truncated
at N (mp.py:59-65).

Every arithmetic step of the forward loop is a library kernel: the convolution
map is the engine's correlation (``mpb200_correlate``) of the residual,
left-padded by A-1 zeros, with the REVERSED atoms -- whose spectra the plan
caches, where the reference re-transforms the static padded dictionary in
every iteration (mp.py:60) -- then ``mpb200_select_dense`` and
``mpb200_scatter_rows``.  When gradients are wanted (``mp.py::train``) the
engine only finds the (atom, time) indices and PyTorch re-evaluates the
selected values and channels on those fixed indices with the graph attached
(``autograd.fixed_index_forward``): the selection itself carries no gradient
in the reference either (top-k indices)."""
from __future__ import annotations

import torch
from torch import nn

from . import engine


class MatchingPursuit(nn.Module):
    def __init__(self, n_atoms: int, atom_samples: int, n_samples: int, n_iterations: int):
        super().__init__()
        self.n_atoms = n_atoms
        self.atom_samples = atom_samples
        self.n_samples = n_samples
        self.n_iterations = n_iterations
        self.atoms = nn.Parameter(torch.zeros(1, n_atoms, atom_samples).uniform_(-0.01, 0.01))    # mp.py:40

    @property
    def normalized_atoms(self):
        """Zero-padded to n_samples and -- despite the name -- NOT normalised (mp.py:43-48)."""
        pad = torch.zeros(1, self.n_atoms, self.n_samples - self.atom_samples, device=self.atoms.device)
        return torch.cat([self.atoms, pad], dim=-1)

    def _select(self, audio: torch.Tensor, want_channels: bool):
        """The forward loop on the engine, without autograd: returns (atom (B,S) int64, time (B,S) int64,
        channels (B,S,N) or None) on the work device."""
        from .matchingpursuit import get_plan
        batch = audio.shape[0]
        work = audio.device if audio.is_cuda else engine._require_cuda(None)
        n, a, k = self.n_samples, self.atom_samples, self.n_atoms
        atoms = engine._dev_f32(self.atoms.detach(), work).view(k, a)
        # conv[t] = sum_i atoms[i] x[t-i] = corr([0^(A-1), x], reversed atoms)[t]
        plan = get_plan(k, a, n + a - 1, batch, work, "recorrelate")
        with plan:
            plan.set_dictionary(torch.flip(atoms, dims=(-1,)), normalize=False)
            padded = torch.zeros(batch, n + a - 1, device=work)
            padded[:, a - 1:] = engine._dev_f32(audio, work).view(batch, n)
            residual = padded[:, a - 1:]                                          # view: updates land in `padded`
            channels = torch.zeros(batch, self.n_iterations, n, device=work) if want_channels else None
            rows = torch.arange(batch, device=work, dtype=torch.int32)
            ks, ts = [], []
            for i in range(self.n_iterations):
                spec = plan.correlate(padded)[..., :n].contiguous()               # mp.py:60  (B, K, N)
                best = engine.select_dense(spec)                                  # mp.py:61  top-1 of sparsify2
                v, kk, p = engine.unpack_best(best)
                scaled = engine.gather_atoms(atoms, kk, v * v)                    # mp.py:62-63: value applied twice
                buf = torch.zeros(batch, n, device=work)
                engine.scatter_rows(buf, scaled, rows, p)
                residual -= buf                                                   # mp.py:64
                if want_channels:
                    channels[:, i, :] = buf                                       # mp.py:65
                ks.append(kk.long()); ts.append(p.long())
        stack = (lambda xs: torch.stack(xs, dim=1) if xs else torch.zeros(batch, 0, dtype=torch.int64, device=work))
        return stack(ks), stack(ts), channels

    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        out_dev = audio.device
        if torch.is_grad_enabled() and (self.atoms.requires_grad or audio.requires_grad):
            from .autograd import fixed_index_forward
            with torch.no_grad():
                k_idx, t_idx, _ = self._select(audio, want_channels=False)
            return fixed_index_forward(self.atoms, audio, k_idx.to(out_dev), t_idx.to(out_dev), self.n_samples)
        with torch.no_grad():
            _, _, channels = self._select(audio, want_channels=True)
        return channels.to(out_dev)
