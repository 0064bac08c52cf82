"""Counterparts of the reference's ``modules/conv.py``: the atom x signal
correlation map in its direct (``torch_conv``, lines 4-9) and FFT
(``fft_convolve(signal, atoms, approx)``, lines 11-53) forms.  Both produce

    fm[b, k, t] = sum_i pad(signal)[b, t + i] * atoms[k, i],   t in [0, N)

and both are served by the engine's overlap-save correlation
(``mpb200_correlate``), which has no use for the choice of form."""
from __future__ import annotations

import torch

from . import engine
from .matchingpursuit import get_plan


def _dense_map(signal: torch.Tensor, atoms: torch.Tensor) -> torch.Tensor:
    if atoms.dim() != 2:
        raise ValueError("atoms must be (n_atoms, atom_size)")
    b, n = signal.shape[0], signal.shape[-1]
    out_dev = signal.device
    work = signal.device if signal.is_cuda else engine._require_cuda(None)
    plan = get_plan(atoms.shape[0], atoms.shape[1], n, b, work, "recorrelate")
    plan.set_dictionary(atoms, normalize=False)        # the helpers correlate with the atoms AS GIVEN
    return plan.correlate(engine._dev_f32(signal, work, (b, n))).to(out_dev)


def torch_conv(signal, atom):
    """modules/conv.py:4-9."""
    return _dense_map(signal, atom)


def fft_convolve(signal, atoms, approx=None):
    """modules/conv.py:11-53.  ``approx=None`` or ``int >= n_samples`` is the full product
    (:48-49).  ``approx=slice`` (band-limited product over the bins of a length N+A transform,
    :24-29) and ``approx=int < n_samples`` (:30-47, defective in the reference: only atom 0 is
    populated) are not served by the engine."""
    n = signal.shape[-1]
    if isinstance(approx, slice) or (isinstance(approx, int) and not isinstance(approx, bool) and approx < n):
        raise NotImplementedError("approx=slice / approx=int<N are not part of the engine yet")
    return _dense_map(signal, atoms)
