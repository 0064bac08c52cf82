"""Counterpart of ``modules/sparse.py::sparsify2`` (lines 46-89): top-k over
the flattened (channel, time) plane, returned as three dense one-hot-like
tensors that hold the selected VALUE.  The selection is the library's dense
argmax (``mpb200_select_dense``); k > 1 repeats it with the winners masked."""
from __future__ import annotations

import torch

from . import engine


def top_entries(x: torch.Tensor, n_to_keep: int):
    """(values (B,k) float32, channel (B,k) int64, time (B,k) int64), descending."""
    b, c, t = x.shape
    work = x.device if x.is_cuda else engine._require_cuda(None)
    cur = engine._dev_f32(x, work)
    if n_to_keep > 1:
        cur = cur.clone()
    vals, chans, times = [], [], []
    rows = torch.arange(b, device=work)
    for i in range(n_to_keep):
        best = engine.select_dense(cur)
        v, k, p = engine.unpack_best(best)
        vals.append(v.clone()); chans.append(k.long()); times.append(p.long())
        if i + 1 < n_to_keep:
            cur[rows, k.long(), p.long()] = float("-inf")
    return torch.stack(vals, 1), torch.stack(chans, 1), torch.stack(times, 1)


def sparsify2(x: torch.Tensor, n_to_keep: int = 8):
    """modules/sparse.py:46-89: ``sparse`` (B,C,T), ``packed`` (B,k,T), ``context`` (B,k,C)."""
    b, c, t = x.shape
    out_dev = x.device
    vals, ch, tm = top_entries(x, n_to_keep)
    work = vals.device
    rows = torch.arange(b, device=work).view(-1, 1).expand(-1, n_to_keep)
    slot = torch.arange(n_to_keep, device=work).view(1, -1).expand(b, -1)
    sparse = torch.zeros(b, c, t, device=work)
    sparse[rows, ch, tm] = vals
    packed = torch.zeros(b, n_to_keep, t, device=work)
    packed[rows, slot, tm] = vals
    context = torch.zeros(b, n_to_keep, c, device=work)
    context[rows, slot, ch] = vals
    return sparse.to(out_dev), packed.to(out_dev), context.to(out_dev)
