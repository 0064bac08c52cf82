"""Counterparts of the reference's ``modules/decompose.py``: octave band split
(``fft_frequency_decompose``, lines 5-33), zero-stuffing resample
(``fft_resample``, lines 36-73) and merge (``fft_frequency_recompose``, lines
76-82), computed by ``mpb200_spectral_band`` (forward real FFT, band transfer
and inverse FFT on the device; lengths must be powers of two >= 256)."""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import engine
from ._lib import check, lib


def spectral_band(x: torch.Tensor, n_out: int, bin_lo: int, bin_hi: int) -> torch.Tensor:
    """``irfft(rfft(x, 'ortho')[bins lo..hi) kept], n=n_out, 'ortho')`` over the last dimension."""
    n_in = x.shape[-1]
    out_dev = x.device
    work = x.device if x.is_cuda else engine._require_cuda(None)
    x2 = engine._dev_f32(x, work).reshape(-1, n_in)
    out = torch.empty(x2.shape[0], n_out, device=work, dtype=torch.float32)
    with torch.cuda.device(work):
        check(lib().mpb200_spectral_band(engine._ptr(x2), x2.shape[0], n_in, engine._ptr(out), n_out, bin_lo, bin_hi,
                                         engine._stream_ptr(work)), "mpb200_spectral_band")
    return out.view(*x.shape[:-1], n_out).to(out_dev)


def fft_frequency_decompose(x, min_size):
    """{size: band} for size = min, 2*min, ..., N: the lowest band keeps bins [0, size/2], the
    others [size/4, size/2] (modules/decompose.py:5-33)."""
    output = OrderedDict()
    size = min_size
    while size <= x.shape[-1]:
        lo = 0 if size == min_size else size // 4
        output[size] = spectral_band(x, size, lo, size // 2 + 1)
        size *= 2
    return output


def fft_resample(x, desired_size, is_lowest_band):
    """modules/decompose.py:36-73 (the ``tukey(alpha=0)`` window there is identically one)."""
    n_coeffs = x.shape[-1] // 2 + 1
    lo = 0 if is_lowest_band else n_coeffs // 2
    return spectral_band(x, desired_size, lo, n_coeffs)


def fft_frequency_recompose(d, desired_size):
    """Sum of the resampled bands (modules/decompose.py:76-82)."""
    first = min(d.keys())
    total = None
    for size, band in d.items():
        r = fft_resample(band, desired_size, size == first)
        total = r if total is None else total + r
    return total
