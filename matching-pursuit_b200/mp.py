"""Counterpart of ``mp.py::MatchingPursuit`` (reference lines 32-67), forward
only: the differentiable MP-like loop whose feature map is a zero-padded
CONVOLUTION of the residual with the raw (un-normalised, mp.py:43-48) atoms,
whose top-1 entry picks (atom k0, time t0, value v), and whose subtracted
channel is ``conv(v*atom_k0, v*delta_t0)`` = ``v^2 * atom_k0`` shifted to t0 and
truncated at N (mp.py:59-65).

Every arithmetic step is a library kernel: the convolution map
(``mpb200_fft_convolve``), the selection (``mpb200_select_dense``) and the
placement of the scaled atom (``mpb200_scatter_rows``).  Training
(``mp.py::train``, Adam + ``iterative_loss``) stays in PyTorch with the
reference module; this class raises if gradients are requested."""
from __future__ import annotations

import torch
from torch import nn

from . import engine
from ._lib import MpbError
from .fft import fft_convolve


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

    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and (self.atoms.requires_grad or audio.requires_grad):
            raise MpbError("matching_pursuit_b200.mp.MatchingPursuit is forward-only; wrap the call in "
                           "torch.no_grad() (training stays with the reference module in PyTorch)")
        batch, _, time = audio.shape
        out_dev = audio.device
        work = audio.device if audio.is_cuda else engine._require_cuda(None)
        n, a = self.n_samples, self.atom_samples
        atoms = engine._dev_f32(self.atoms.detach(), work).view(self.n_atoms, a)
        na = torch.cat([atoms, torch.zeros(self.n_atoms, n - a, device=work)], dim=-1).view(1, self.n_atoms, n)
        residual = engine._dev_f32(audio, work).clone().view(batch, n)
        channels = torch.zeros(batch, self.n_iterations, n, device=work)
        rows = torch.arange(batch, device=work, dtype=torch.int32)
        for i in range(self.n_iterations):
            spec = fft_convolve(residual.view(batch, 1, n), na)               # mp.py:60  (B, K, N)
            best = engine.select_dense(spec)                                  # mp.py:61  top-1 of sparsify2
            v, k, p = engine.unpack_best(best)
            scaled = engine.gather_atoms(atoms, k, v * v)                     # mp.py:62-63: value applied twice
            step = channels[:, i, :]                                          # strided view: scatter into a copy
            buf = torch.zeros(batch, n, device=work)
            engine.scatter_rows(buf, scaled, rows, p)
            engine.scatter_rows(residual, engine.gather_atoms(atoms, k, -(v * v)), rows, p)   # mp.py:64: r - v^2*atom
            step.copy_(buf)                                                   # mp.py:65
        return channels.to(out_dev)
