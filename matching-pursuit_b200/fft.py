"""Counterpart of the reference's ``modules/fft.py::fft_convolve`` (lines
23-35; identical arithmetic at ``modules/transfer.py:548-569`` with
``correlation=False``): N-ary zero-padded FFT convolution with broadcasting
over the leading dimensions, computed by ``mpb200_fft_convolve``."""
from __future__ import annotations

import ctypes as C

import torch

from . import engine
from ._lib import MpbError, check, lib

MAX_OPERANDS = 4


def n_fft_coeffs(size: int) -> int:
    """modules/fft.py:6-7."""
    return size // 2 + 1


def _convolve(args, norm, conj_mask: int) -> torch.Tensor:
    if not 1 <= len(args) <= MAX_OPERANDS:
        raise MpbError(f"fft_convolve takes 1..{MAX_OPERANDS} operands, got {len(args)}")
    n = args[0].shape[-1]
    if any(a.shape[-1] != n for a in args):
        # the reference pads each operand to twice ITS OWN length and then fails to multiply the spectra
        raise RuntimeError("fft_convolve: operands must share their last dimension")
    out_dev = args[0].device
    work = args[0].device if args[0].is_cuda else engine._require_cuda(None)
    lead = torch.broadcast_shapes(*[tuple(a.shape[:-1]) for a in args])
    rows_out = 1
    for s in lead:
        rows_out *= s
    ops, maps, counts = [], [], []
    for a in args:
        a2 = engine._dev_f32(a, work).reshape(-1, n)
        ops.append(a2)
        counts.append(a2.shape[0])
        if tuple(a.shape[:-1]) == tuple(lead):
            maps.append(None)
        else:
            idx = torch.arange(a2.shape[0], device=work, dtype=torch.int32).reshape(a.shape[:-1])
            maps.append(idx.expand(lead).reshape(-1).contiguous())
    k = len(args)
    length = 2 * n                       # the reference's transform length (modules/fft.py:28)
    scale = {None: 1.0, "backward": 1.0, "ortho": float(length) ** ((1 - k) / 2.0),
             "forward": float(length) ** (1 - k)}[norm]
    out = torch.empty(rows_out, n, device=work, dtype=torch.float32)
    op_ptrs = (C.c_void_p * k)(*[o.data_ptr() for o in ops])
    map_ptrs = (C.c_void_p * k)(*[0 if m is None else m.data_ptr() for m in maps])
    rows = (C.c_int32 * k)(*counts)
    with torch.cuda.device(work):
        check(lib().mpb200_fft_convolve(op_ptrs, map_ptrs, rows, k, rows_out, n, conj_mask, C.c_float(scale),
                                        engine._ptr(out), engine._stream_ptr(work)), "mpb200_fft_convolve")
    return out.view(*lead, n).to(out_dev)


def fft_convolve(*args, norm=None) -> torch.Tensor:
    """``irfft(prod_i rfft(pad(x_i, 2n)))[..., :n]`` (modules/fft.py:23-35)."""
    return _convolve(args, norm, 0)


def fft_correlate(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """``transfer.fft_convolve(a, b, correlation=True)``: the spectrum of ``b`` is conjugated
    (modules/transfer.py:548-569)."""
    return _convolve((a, b), None, 0b10)
