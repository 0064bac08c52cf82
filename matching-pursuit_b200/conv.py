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
    if atoms.shape[1] > engine.MAX_PLAN_ATOM:          # long atoms: correlated in parts (include/mpb200.h, mpb200_fold_parts)
        parts, n_parts = engine.split_long_atoms(engine._dev_f32(atoms, work))
        plan = get_plan(parts.shape[0], parts.shape[1], n, b, work, "recorrelate")
        with plan:
            plan.set_dictionary(parts, normalize=False)
            return engine.correlate_long(engine._dev_f32(signal, work, (b, n)), plan, atoms.shape[0], n_parts).to(out_dev)
    plan = get_plan(atoms.shape[0], atoms.shape[1], n, b, work, "recorrelate")
    plan.set_dictionary(atoms, normalize=False)        # the helpers correlate with the atoms AS GIVEN
    return plan.correlate(engine._dev_f32(signal, work, (b, n))).to(out_dev)


def torch_conv(signal, atom):
    """modules/conv.py:4-9."""
    return _dense_map(signal, atom)


def band_limited_map(signal2d: torch.Tensor, plan_long, n: int, slce: slice) -> torch.Tensor:
    """modules/conv.py:24-29 with the engine's kernels: the product of the masked signal spectrum with the
    atom spectra, inverted at length L = N + A, rolled and cropped (:51-53), equals the ordinary correlation
    of y = irfft(mask(rfft(pad(signal, L)))) with the atoms, because t + i < L never wraps.  ``plan_long`` is
    a plan for signals of L samples holding the atoms; returns (B, K, N)."""
    y = engine.band_limit(signal2d, plan_long.n_samples, slce)
    return plan_long.correlate(y)[..., :n].contiguous()


def fft_convolve(signal, atoms, approx=None):
    """modules/conv.py:11-53.  ``approx=None`` or ``int >= n_samples`` is the full product (:48-49);
    ``approx=slice`` keeps only those rfft bins of the length N+A transform (:24-29).
    ``approx=int < n_samples`` (:30-47) is defective in the reference (only atom 0 is populated,
    SURVEY.md 8 a3) and is not served by the engine."""
    n = signal.shape[-1]
    if isinstance(approx, int) and not isinstance(approx, bool) and approx < n:
        raise NotImplementedError("approx=int<N (top-k spectral bins, defective in the reference) is not part "
                                  "of the engine")
    if isinstance(approx, slice):
        if atoms.dim() != 2:
            raise ValueError("atoms must be (n_atoms, atom_size)")
        b = signal.shape[0]
        out_dev = signal.device
        work = signal.device if signal.is_cuda else engine._require_cuda(None)
        plan = get_plan(atoms.shape[0], atoms.shape[1], n + atoms.shape[1], b, work, "recorrelate")
        plan.set_dictionary(atoms, normalize=False)
        return band_limited_map(engine._dev_f32(signal, work, (b, n)), plan, n, approx).to(out_dev)
    return _dense_map(signal, atoms)
