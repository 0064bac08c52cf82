"""CPU oracle for the greedy matching-pursuit hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on CPU and with the same third-party arithmetic the
reference uses (``torch`` -- ``F.conv1d``, ``torch.fft``, ``torch.max`` -- the
only library that carries arithmetic on this path),
what JohnVinyard/matching-pursuit computes on its sparse-coding path.  It is
NOT product code: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity pinning: the reference ships no test, golden vector or fixture for this
path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF executed in the build container: ``oracle/make_golden.py``
imports the unmodified reference from ``/root/reference`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks that this
restatement reproduces those files.

Structure differs from the reference on purpose: the greedy loop is stated
once (``greedy_pursuit``) and records a step-major *trace* of
``(atom, position, value)``; every reference return convention is derived from
that trace.  Each function cites the reference lines it follows (paths are
relative to ``/root/reference``).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from functools import reduce
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch.nn import functional as F


# --------------------------------------------------------------------------
# normalisation  (modules/normalization.py:4-15)
# --------------------------------------------------------------------------

def unit_norm(x: torch.Tensor, dim: int = -1, epsilon: float = 1e-8) -> torch.Tensor:
    """``x / (||x||_2 + eps)`` along ``dim`` -- eps is added to the norm, not
    clamped (modules/normalization.py:4-6)."""
    return x / (torch.norm(x, dim=dim, keepdim=True) + epsilon)


def max_norm(x: torch.Tensor, dim: int = -1, epsilon: float = 1e-8) -> torch.Tensor:
    """``x / (max|x| + eps)`` along ``dim`` (modules/normalization.py:9-15)."""
    peak = torch.max(torch.abs(x), dim=dim, keepdim=True)[0]
    return x / (peak + epsilon)


# --------------------------------------------------------------------------
# atom x residual correlation
# --------------------------------------------------------------------------

def correlate_direct(signal: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """fm[b,k,t] = sum_i pad(signal)[b,t+i] * d[k,i], t in [0,N)
    (modules/conv.py:4-9; modules/matchingpursuit.py:275-277, 90-92).
    ``signal`` is (B,1,N); ``d`` is (K,A).  The signal is right-padded with A
    zeros, so atoms overhanging the right edge see zeros."""
    n = signal.shape[-1]
    k, a = d.shape
    return F.conv1d(F.pad(signal, (0, a)), d.view(k, 1, a))[..., :n]


def correlate_fft(signal: torch.Tensor, atoms: torch.Tensor, approx=None) -> torch.Tensor:
    """Same map through a length-(N+A) circular product with the flipped,
    right-padded atoms, rolled by +1 and cropped (modules/conv.py:11-53).

    ``approx`` follows the reference: ``None`` or ``int >= N`` -> full product
    (conv.py:48-49); ``slice`` -> only those rfft bins are kept
    (conv.py:24-29); ``int < N`` -> per-signal top-``approx`` magnitude bins
    (conv.py:30-47).  The last mode gathers the atom spectrum with indices of
    shape (B,1,approx) and therefore only ever populates atom 0; that defect
    is the reference's and is restated faithfully."""
    batch, n = signal.shape[0], signal.shape[-1]
    k, a = atoms.shape
    sig_p = F.pad(signal, (0, a))
    total = sig_p.shape[-1]
    atoms_p = F.pad(atoms, (0, total - a))
    s_spec = torch.fft.rfft(sig_p, dim=-1)
    a_spec = torch.fft.rfft(torch.flip(atoms_p, dims=(-1,)), dim=-1)[None, ...]
    if isinstance(approx, slice):
        prod = torch.zeros(batch, k, s_spec.shape[-1], dtype=s_spec.dtype, device=signal.device)
        prod[..., approx] = s_spec[..., approx] * a_spec[..., approx]
    elif isinstance(approx, int) and approx < n:
        prod = torch.zeros(batch, k, s_spec.shape[-1], dtype=s_spec.dtype, device=signal.device)
        _, idx = torch.topk(torch.abs(s_spec), k=approx, dim=-1)
        picked_sig = torch.gather(s_spec, dim=-1, index=idx)
        picked_atom = torch.gather(a_spec.repeat(batch, 1, 1), dim=-1, index=idx)
        prod = torch.scatter(prod, dim=-1, index=idx, src=picked_sig * picked_atom)
    else:
        prod = s_spec * a_spec
    fm = torch.roll(torch.fft.irfft(prod, dim=-1), 1, dims=(-1,))
    return fm[..., :n]


def convolve_fft(*args: torch.Tensor, norm: Optional[str] = None) -> torch.Tensor:
    """N-ary zero-padded FFT *convolution*: every argument is right-padded to
    twice its own last dimension, spectra are multiplied left to right, and
    the result is cropped to ``args[0].shape[-1]`` (modules/fft.py:23-35;
    identical arithmetic at modules/transfer.py:548-569 with
    ``correlation=False``)."""
    n = args[0].shape[-1]
    specs = [torch.fft.rfft(F.pad(x, (0, x.shape[-1])), dim=-1, norm=norm) for x in args]
    prod = reduce(lambda acc, cur: acc * cur, specs[1:], specs[0])
    return torch.fft.irfft(prod, dim=-1, norm=norm)[..., :n]


# --------------------------------------------------------------------------
# straight-through one-hot and top-k selection  (modules/sparse.py:29-89)
# --------------------------------------------------------------------------

def soft_dirac(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """Forward value: one-hot at argmax of softmax(x); backward: softmax
    (modules/sparse.py:29-43)."""
    soft = torch.softmax(x, dim=dim)
    _, idx = torch.max(soft, dim=dim, keepdim=True)
    hard = torch.scatter(torch.zeros_like(soft), dim, idx, 1.0)
    return soft + (hard - soft).detach()


def sparsify2(x: torch.Tensor, n_to_keep: int = 8):
    """top-k over the flattened (C*T) axis; returns ``sparse`` (B,C,T),
    ``packed`` (B,k,T) and ``context`` (B,k,C), each holding the selected
    VALUE at the selected coordinate (modules/sparse.py:46-89)."""
    b, c, t = x.shape
    flat = x.reshape(b, -1)
    vals, idx = torch.topk(flat, k=n_to_keep, dim=-1)
    chan, time = idx // t, idx % t
    slot = torch.arange(n_to_keep, device=x.device)[None, :]
    sparse = torch.scatter(torch.zeros_like(flat), -1, idx, vals).view(b, c, t)
    context = torch.scatter(torch.zeros(b, n_to_keep * c, device=x.device), -1,
                            slot * c + chan, vals).view(b, n_to_keep, c)
    packed = torch.scatter(torch.zeros(b, n_to_keep * t, device=x.device), -1,
                           slot * t + time, vals).view(b, n_to_keep, t)
    return sparse, packed, context


# --------------------------------------------------------------------------
# decode: place scaled atoms back on a time axis
# --------------------------------------------------------------------------

def make_scatter(n_samples: int, atom_size: int, device=None) -> Callable:
    """Decoder closure (modules/matchingpursuit.py:20-58).

    ``scatter(x, events)``: ``x`` is a shape tuple (-> fresh float32 zeros on
    ``device``, the reference's global ``util.device``, :28) or a tensor; the
    buffer is padded to 3N, each event ``(atom, batch, pos, scaled_atom)`` is
    ADDED at ``N + pos`` (single channel, :48) or ASSIGNED to channel = that
    batch row's running event count (multi channel, :50), and the middle N
    samples are returned (:56) -- atoms overhanging the right edge are
    truncated."""
    dev = torch.device("cpu") if device is None else device

    def scatter(x, events):
        multi = 1
        if isinstance(x, tuple):
            x = torch.zeros(*x, device=dev)
            multi = x.shape[1]
        wide = torch.cat([torch.zeros_like(x), x, torch.zeros_like(x)], dim=-1)
        seen: Dict[int, int] = {}
        for _, j, p, a in events:
            lo = n_samples + int(p)
            hi = lo + atom_size
            ch = seen.get(j, 0)
            if multi == 1:
                wide[j, :, lo:hi] += a.view(-1, atom_size)
            else:
                wide[j, ch, lo:hi] = a.view(-1, atom_size)
            seen[j] = ch + 1
        return wide[..., n_samples:2 * n_samples]

    return scatter


# --------------------------------------------------------------------------
# the greedy loop, stated once
# --------------------------------------------------------------------------

@dataclass
class Trace:
    """Step-major record of one greedy pursuit.  ``atom``/``pos`` are int64
    (S,B); ``val`` float32 (S,B); ``margin`` float64 (S,B) top-2 relative
    margin of the selection map (nan when not requested); ``d_unit`` the
    normalised dictionary the events refer to; ``residual`` (B,1,N)."""
    atom: torch.Tensor
    pos: torch.Tensor
    val: torch.Tensor
    margin: torch.Tensor
    d_unit: torch.Tensor
    residual: torch.Tensor


def selection_map(fm: torch.Tensor, local_contrast_norm: bool) -> torch.Tensor:
    """The map the argmax runs on: the raw correlation, or with the opt-in
    9x9 box-filter mean removed over the (atom, time) plane
    (modules/matchingpursuit.py:286-292)."""
    if not local_contrast_norm:
        return fm
    b, k, n = fm.shape
    plane = fm.view(b, 1, k, n)
    return (plane - F.avg_pool2d(plane, (9, 9), (1, 1), (4, 4))).view(b, k, n)


def greedy_pursuit(signal: torch.Tensor, d: torch.Tensor, n_steps: int, approx=None,
                   local_contrast_norm: bool = False,
                   compute_feature_map: Optional[Callable] = None,
                   on_map: Optional[Callable] = None,
                   on_select: Optional[Callable] = None,
                   want_margin: bool = False) -> Trace:
    """n_steps iterations of: correlate every atom with the residual, take the
    SIGNED global maximum over (atom, position) -- first flat index on ties --
    and subtract ``value * unit_atom`` at that position, truncated at the
    right edge (modules/matchingpursuit.py:254-328).

    ``signal`` is (B,1,N) float32; ``d`` (K,A).  Neither is modified.
    ``on_map(step, fm)`` sees the dense (B,K,N) correlation map of each step;
    ``on_select(step, fm, atom (B,), pos (B,), value (B,), scaled (B,A))`` runs
    after the selection and before the residual update, like the reference's
    per-row callbacks (:311-324).
    """
    if signal.dim() != 3 or signal.shape[1] != 1:
        raise ValueError("oracle states the single-channel (B,1,N) case only "
                         "(the reference's multi-channel branch fails at matchingpursuit.py:50)")
    b, _, n = signal.shape
    k, a = d.shape[0], d.shape[-1]
    du = unit_norm(d, dim=-1)                                   # :254
    residual = signal.clone()                                   # :256
    atoms = torch.zeros(n_steps, b, dtype=torch.int64)
    poss = torch.zeros(n_steps, b, dtype=torch.int64)
    vals = torch.zeros(n_steps, b, dtype=torch.float32)
    margins = torch.full((n_steps, b), float("nan"), dtype=torch.float64)
    for step in range(n_steps):
        if compute_feature_map is not None:                     # :272-273
            fm = compute_feature_map(residual, du)
        elif approx is None:                                    # :274-277
            fm = correlate_direct(residual, du)
        else:                                                   # :278-280
            fm = correlate_fft(residual, du, approx=approx)
        if on_map is not None:
            on_map(step, fm)
        sel = selection_map(fm, local_contrast_norm).reshape(b, -1)
        _, flat = torch.max(sel, dim=-1, keepdim=True)          # :294 / :299
        value = torch.gather(fm.reshape(b, -1), -1, flat)       # :296 (== max value when no LCN)
        if want_margin:
            top2 = torch.topk(sel.double(), 2, dim=-1)[0]
            margins[step] = (top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-300)
        ai = (flat // n).view(b)                                # :302
        pp = (flat % n).view(b)                                 # :303
        scaled = du[ai] * value.view(b, 1)                      # :305  (B,A), one fp32 rounding
        if on_select is not None:
            on_select(step, fm, ai, pp, value.view(b), scaled)
        for j in range(b):                                      # :326-328 via :33-56
            p = int(pp[j])
            keep = min(a, n - p)
            residual[j, 0, p:p + keep] -= scaled[j, :keep]
        atoms[step], poss[step], vals[step] = ai, pp, value.view(b)
    return Trace(atoms, poss, vals, margins, du, residual)


def trace_events(tr: Trace) -> List[List[Tuple[int, int, torch.Tensor, torch.Tensor]]]:
    """Per-step event lists in the reference's tuple format
    ``(atom:int, batch:int, pos: int64 (1,1), scaled_atom: float32 (1,1,A))``
    (modules/matchingpursuit.py:311-321)."""
    steps, b = tr.atom.shape
    out = []
    for s in range(steps):
        row = []
        for j in range(b):
            ai = int(tr.atom[s, j])
            a = (tr.d_unit[ai] * tr.val[s, j]).view(1, 1, -1)
            row.append((ai, j, tr.pos[s, j].view(1, 1), a))
        out.append(row)
    return out


def group_by_atom(per_step) -> "OrderedDict[int, list]":
    """``instances`` of the reference: a dict keyed by atom index in
    FIRST-SEEN order, each value the events of that atom in step-major,
    batch-minor order (modules/matchingpursuit.py:261, 321)."""
    grouped: "OrderedDict[int, list]" = OrderedDict()
    for row in per_step:
        for ev in row:
            grouped.setdefault(ev[0], []).append(ev)
    return grouped


def flatten_groups(grouped) -> list:
    """Concatenate the per-atom lists (modules/matchingpursuit.py:61-65)."""
    flat = []
    for lst in grouped.values():
        flat.extend(lst)
    return flat


def sparse_code(signal, d, n_steps=100, device=None, approx=None, flatten=False,
                extract_atom_embedding=None, visit_key_point=None, return_residual=False,
                local_contrast_norm=False, return_sparse_feature_map=False,
                compute_feature_map=None, fft_convolution=False):
    """Every return convention of the reference's ``sparse_code``
    (modules/matchingpursuit.py:229-345), derived from one ``greedy_pursuit``.
    ``device`` and ``fft_convolution`` are accepted and ignored as in the
    reference (:233, :242)."""
    b, _, n = signal.shape
    k, a = d.shape[0], d.shape[-1]
    embeddings = []
    du = unit_norm(d, dim=-1)

    def on_map(step, fm):                                                   # :282-283
        embeddings.append(extract_atom_embedding(fm, du))

    def on_select(step, fm, ai, pp, value, scaled):                         # :323-324
        for j in range(b):
            visit_key_point(fm[j].view(k, n), int(ai[j]), pp[j].view(1), scaled[j].view(a))

    tr = greedy_pursuit(signal, d, n_steps, approx=approx,
                        local_contrast_norm=local_contrast_norm,
                        compute_feature_map=compute_feature_map,
                        on_map=on_map if extract_atom_embedding is not None else None,
                        on_select=on_select if visit_key_point is not None else None)
    per_step = trace_events(tr)
    scatter = make_scatter(n, a, device=signal.device)
    if extract_atom_embedding is not None:                                  # :332-333
        return embeddings, tr.residual
    grouped = group_by_atom(per_step)
    if not flatten:                                                         # :335-336
        return grouped, scatter
    flat = flatten_groups(grouped)
    if return_residual:                                                     # :337-339
        return flat, scatter, tr.residual
    if return_sparse_feature_map:                                           # :340-342, :266-267, :317-318
        sfm = torch.zeros(b, k, n, device=signal.device)
        for s in range(tr.atom.shape[0]):
            for j in range(b):
                sfm[j, int(tr.atom[s, j]), int(tr.pos[s, j])] += tr.val[s, j]
        return flat, scatter, sfm
    return flat, scatter                                                    # :343-345


def sparse_feature_map(signal, d, n_steps=100, device=None, approx=None, pooling=None,
                       return_residual=False):
    """Dense (B,K,N) accumulation of the winners (modules/matchingpursuit.py:68-125).
    Forward values only: ``soft_dirac(f) * f`` is ``f`` at the argmax and 0
    elsewhere (:100-101), so each step adds the winning value at its
    (atom, position).  The subtraction is the in-place truncated slice
    update of :108-120, which rounds ``d*v`` then subtracts -- the same two
    roundings as ``greedy_pursuit``."""
    sig = signal.view(signal.shape[0], 1, -1)
    b, _, n = sig.shape
    k = d.shape[0]
    tr = greedy_pursuit(sig, d, n_steps, approx=approx)
    fm = torch.zeros(b, k, n, device=device)
    for s in range(n_steps):
        for j in range(b):
            fm[j, int(tr.atom[s, j]), int(tr.pos[s, j])] += tr.val[s, j]
    return (fm, tr.residual) if return_residual else fm


def dictionary_learning_step(signal, d, n_steps=100, device=None, approx=None,
                             local_constrast_norm=False, compute_feature_map=None,
                             fft_convolution=False):
    """One dictionary update (modules/matchingpursuit.py:348-419): code the
    batch, then for every used atom (first-seen order) add its instances back
    to a running copy of the SIGNAL (:367 -- not the coding residual), replace
    the atom by the unit-normed sum of the segments under its instances
    (:400-406) and subtract the re-scaled new atom (:408-415)."""
    b, c, n = signal.shape
    a = d.shape[-1]
    d = unit_norm(d, dim=-1)
    running = signal.clone()
    grouped, scatter = sparse_code(signal, d, n_steps=n_steps, approx=approx,
                                   local_contrast_norm=local_constrast_norm,
                                   compute_feature_map=compute_feature_map)
    for index, inst in grouped.items():
        running = running + scatter(running.shape, inst)
        wide = torch.cat([torch.zeros_like(running), running, torch.zeros_like(running)], dim=-1)
        segs = torch.cat([wide[j, :, n + int(p): n + int(p) + a][None] for _, j, p, _ in inst], dim=0)
        new_atom = unit_norm(torch.sum(segs, dim=0).view(-1)).view(c, a)
        d[index] = new_atom
        rescaled = [(ai, j, p, new_atom * torch.norm(atom, dim=-1, keepdim=True))
                    for ai, j, p, atom in inst]
        running = running - scatter(running.shape, rescaled)
    return unit_norm(d, dim=-1)


# --------------------------------------------------------------------------
# octave band split / merge  (modules/decompose.py:5-82)
# --------------------------------------------------------------------------

def band_split(x: torch.Tensor, min_size: int) -> "OrderedDict[int, torch.Tensor]":
    """Ortho rfft, then for each size s = min, 2*min, ..., N keep bins
    [0, s/2] (lowest band) or [s/4, s/2] (others) and irfft at length s
    (modules/decompose.py:5-33)."""
    coeffs = torch.fft.rfft(x, norm="ortho")
    out: "OrderedDict[int, torch.Tensor]" = OrderedDict()
    size = min_size
    while size <= x.shape[-1]:
        part = coeffs[:, :, : size // 2 + 1]
        if size > min_size:
            mask = torch.zeros(part.shape[2], device=x.device)
            mask[size // 4: size // 2 + 1] = 1
            part = part * mask[None, None, :]
        out[size] = torch.fft.irfft(part, n=size, norm="ortho")
        size *= 2
    return out


def band_resample(x: torch.Tensor, desired_size: int, is_lowest_band: bool) -> torch.Tensor:
    """Zero-stuff a band's ortho spectrum into a longer one and invert
    (modules/decompose.py:36-73).  The reference multiplies by
    ``tukey(n, alpha=0)``, which is identically 1."""
    b, c, _ = x.shape
    coeffs = torch.fft.rfft(x, norm="ortho")
    nc = coeffs.shape[2]
    wide = torch.zeros(b, c, desired_size // 2 + 1, dtype=torch.complex64, device=x.device)
    if is_lowest_band:
        wide[:, :, :nc] = coeffs
    else:
        wide[:, :, nc // 2: nc] = coeffs[:, :, nc // 2:]
    return torch.fft.irfft(wide, n=desired_size, norm="ortho")


def band_merge(bands: Dict[int, torch.Tensor], desired_size: int) -> torch.Tensor:
    """Sum of the resampled bands (modules/decompose.py:76-82)."""
    lowest = min(bands.keys())
    return sum(band_resample(v, desired_size, s == lowest) for s, v in bands.items())


# --------------------------------------------------------------------------
# multi-band orchestration  (modules/multibanddict.py:53-279, 282-473)
# --------------------------------------------------------------------------

class BandOracle:
    """One band's dictionary and codec (modules/multibanddict.py:53-279).
    ``d`` is given explicitly so tests control the bits."""

    def __init__(self, size: int, d: torch.Tensor, slce: Optional[slice] = None,
                 local_contrast_norm: bool = False, is_lowest_band: bool = False):
        self.size = size
        self.n_atoms, self.atom_size = d.shape
        self.slce = slce
        self.local_contrast_norm = local_contrast_norm
        self.is_lowest_band = is_lowest_band
        self.d = unit_norm(d)                                               # :89-91

    def encode(self, batch, steps=16):                                      # :238-263
        flat, scatter = sparse_code(batch, self.d, steps, approx=self.slce, flatten=True,
                                    local_contrast_norm=self.local_contrast_norm)
        return flat, scatter, batch.shape

    def decode(self, shape, events, scatter):                               # :265-266
        return scatter(shape, events)

    def recon(self, batch, steps=16):                                       # :268-279
        events, scatter, shape = self.encode(batch, steps)
        return self.decode(shape, events, scatter), events, scatter

    def learn(self, batch, steps=16):                                       # :178-187
        d = dictionary_learning_step(batch, self.d, steps, approx=self.slce,
                                     local_constrast_norm=self.local_contrast_norm)
        self.d = unit_norm(d)
        return d

    def to_global(self, event, offset):                                     # :204-217
        ai, j, p, atom = event
        return (offset + ai, j, p / self.size, torch.norm(atom))

    def to_local(self, event, offset):                                      # :219-235
        gi, j, unit_time, amp = event
        li = gi - offset
        return (li, j, int(unit_time * self.size), self.d[li] * amp)


class MultibandOracle:
    """Per-band pursuit over an octave split (modules/multibanddict.py:282-473)."""

    def __init__(self, specs: Sequence[BandOracle], n_samples: int):
        self.bands = OrderedDict((s.size, s) for s in specs)
        self.min_size = min(s.size for s in specs)
        self.n_samples = n_samples
        counts = {s.n_atoms for s in specs}
        if len(counts) > 1:                                                 # :289-291
            raise ValueError("Only specs with equal atom counts is currently allowed")
        self.n_atoms = counts.pop()

    def encode(self, batch, steps):                                         # :399-404
        split = band_split(batch, self.min_size)
        return OrderedDict((size, band.encode(split[size], steps)) for size, band in self.bands.items())

    def flattened_event_tuples(self, encoding):                             # :410-422
        out, offset = [], 0
        for size, (events, _, _) in encoding.items():
            band = self.bands[size]
            out.extend(band.to_global(ev, offset) for ev in events)
            offset += band.n_atoms
        return out

    def hierarchical_event_tuples(self, flat, original):                    # :424-443
        per_band: "OrderedDict[int, list]" = OrderedDict()
        ordered = list(self.bands.values())
        for ev in flat:
            index = ev[0] // self.n_atoms                                   # :406-408
            band = ordered[index]
            per_band.setdefault(band.size, []).append(band.to_local(ev, index * self.n_atoms))
        return OrderedDict((size, (events, original[size][1], original[size][2]))
                           for size, events in per_band.items())

    def decode(self, d):                                                    # :446-458
        out = OrderedDict()
        for size, (events, scatter, shape) in d.items():
            out[size] = self.bands[size].decode(shape, events, scatter)
        return band_merge(out, self.n_samples)

    def recon(self, batch, steps=16):                                       # :460-473
        split = band_split(batch, self.min_size)
        recon_bands, events = OrderedDict(), OrderedDict()
        for size, band in self.bands.items():
            r, e, _ = band.recon(split[size], steps)
            recon_bands[size] = r
            events[size] = e
        return band_merge(recon_bands, batch.shape[-1]), events

    def learn(self, batch, steps=16):                                       # :394-397
        for size, band in band_split(batch, self.min_size).items():
            self.bands[size].learn(band, steps)


# --------------------------------------------------------------------------
# mp.py::MatchingPursuit.forward  (mp.py:32-67)
# --------------------------------------------------------------------------

def mp_forward(atoms: torch.Tensor, audio: torch.Tensor, n_samples: int, n_iterations: int) -> torch.Tensor:
    """Differentiable MP-like loop of ``mp.py``: the feature map is a zero-
    padded CONVOLUTION (not correlation) of the residual with the raw
    (un-normalised, mp.py:43-48) atoms right-padded to N; the top-1 entry
    picks (atom k0, time t0, value v); the subtracted channel is
    ``conv(v * atom_k0, v * delta_t0)`` = ``v^2 * atom_k0`` shifted to t0 and
    truncated at N (mp.py:59-65).  ``atoms`` is (1,K,A); returns (B,S,N)."""
    b = audio.shape[0]
    k, a = atoms.shape[1], atoms.shape[2]
    padded = torch.cat([atoms, torch.zeros(1, k, n_samples - a, device=atoms.device)], dim=-1)
    residual = audio
    channels = torch.zeros(b, n_iterations, n_samples, device=audio.device)
    for i in range(n_iterations):
        spec = convolve_fft(residual, padded)
        _, time, atom = sparsify2(spec, n_to_keep=1)
        step = convolve_fft(atom @ padded, time)
        residual = residual - step
        channels[:, i:i + 1, :] = step
    return channels


# --------------------------------------------------------------------------
# synthetic inputs  (reference initialisation at
# modules/multibanddict.py:89-91 and experiments/archive/e_2023_7_18/experiment.py:43-68)
# --------------------------------------------------------------------------

def make_dictionary(n_atoms: int, atom_size: int, seed: int = 0) -> torch.Tensor:
    """``zeros(K,A).uniform_(-1,1)`` then unit norm -- the reference's own
    dictionary initialisation."""
    g = torch.Generator().manual_seed(seed)
    return unit_norm(torch.zeros(n_atoms, atom_size).uniform_(-1, 1, generator=g))


def make_planted_signals(d_unit: torch.Tensor, batch: int, n_samples: int, n_events: int,
                         seed: int = 1, noise: float = 0.01) -> torch.Tensor:
    """Family P: each signal is a sum of ``n_events`` dictionary atoms at
    uniform positions with amplitudes U(0.5,1), plus N(0, noise^2), then
    max-normed per signal.  Returns (B,1,N) float32."""
    k, a = d_unit.shape
    out = torch.zeros(batch, 1, n_samples)
    for j in range(batch):
        g = torch.Generator().manual_seed(seed + j)
        idx = torch.randint(0, k, (n_events,), generator=g)
        pos = torch.randint(0, max(1, n_samples - a + 1), (n_events,), generator=g)
        amp = torch.zeros(n_events).uniform_(0.5, 1.0, generator=g)
        for e in range(n_events):
            p = int(pos[e])
            keep = min(a, n_samples - p)
            out[j, 0, p:p + keep] += amp[e] * d_unit[int(idx[e]), :keep]
        out[j, 0] += noise * torch.randn(n_samples, generator=g)
    return max_norm(out)


def make_noise_signals(batch: int, n_samples: int, seed: int = 2) -> torch.Tensor:
    """Family G: white gaussian noise, the adversarial small-margin case."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 1, n_samples, generator=g)
