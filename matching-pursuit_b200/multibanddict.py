"""Counterparts of the reference's ``modules/multibanddict.py``: ``BandSpec``
(lines 53-279) and ``MultibandDictionaryLearning`` (lines 282-473) -- one
greedy pursuit per octave band of an FFT band split.  Types follow the
reference (:12-16): ``LocalEventTuple = (atom:int, batch:int, position, atom
tensor)``, ``GlobalEventTuple = (global atom:int, batch:int, unit time,
amplitude)``, ``BandEncodingPackage = (events, scatter, shape)``.

The per-band pursuits are independent problems (they share nothing but the
step count): ``MultibandDictionaryLearning.encode`` enqueues all of them on the
CUDA engine, one stream per band, before it waits for any; the band split and
merge run through ``mpb200_spectral_band``.  Event lists carry the packed arrays
they were built from (``EventList.packed`` / ``GlobalEventList.packed``), and the
conversions between local and global tuples (:204-235, :410-443) are array
operations on them -- per-event Python work is limited to materialising the
reference's tuple format.  Dictionary
learning (``learn``) keeps its atom update in PyTorch.  The STFT loss features
of the reference module (``multiband_spectrogram*``, :19-49) are out of scope."""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from hashlib import sha256
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import engine
from .decompose import fft_frequency_decompose, fft_frequency_recompose, fft_resample
from .matchingpursuit import (EventList, _events_from_packed, build_scatter_segments, dictionary_learning_step,
                              sparse_code, sparse_code_start)

Shape = Tuple
LocalEventTuple = Tuple[int, int, int, torch.Tensor]
GlobalEventTuple = Tuple[int, int, float, float]
BandEncodingPackage = Tuple[List[LocalEventTuple], Callable, Shape]


class GlobalEventList(EventList):
    """A list of GlobalEventTuples ``(global atom, batch, unit time (1,1) tensor, amplitude 0-d tensor)`` that carries
    them as arrays: ``packed = (global atom int64 (E,), batch int64 (E,), unit time float32 (E,), amplitude float32
    (E,))`` in list order; like :class:`EventList` the tuples are built on first access."""

    def _fill(self) -> None:
        if self._lazy:
            self._lazy = False
            atom, batch, time, amp = self.packed
            list.extend(self, zip(atom.tolist(), batch.tolist(), time.view(-1, 1, 1).unbind(0), amp.unbind(0)))


class _LocalEventList(EventList):
    """Local tuples as ``to_local_tuple`` makes them (modules/multibanddict.py:219-235): the position is a Python
    int and the atom a plain (A,) tensor."""

    def _fill(self) -> None:
        if self._lazy:
            self._lazy = False
            local, batch, pos, rows = self.packed
            list.extend(self, zip(local.tolist(), batch.tolist(), pos.tolist(), rows.unbind(0)))


def _packed_local(events, atom_size: int):
    """(atom int64 (E,), batch int64 (E,), pos int64 (E,), rows float32 (E, A)) of a local event list: the arrays
    it carries when the engine built it, else stacked from its tuples."""
    packed = getattr(events, "packed", None)
    if packed is not None and len(events) == packed[0].numel():
        atom, batch, pos, rows = packed
        return atom, batch, pos.reshape(-1), rows
    if len(events) == 0:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z, z, torch.zeros(0, atom_size)
    rows = torch.stack([ev[3].reshape(atom_size) for ev in events])
    dev = rows.device
    return (torch.tensor([int(ev[0]) for ev in events], dtype=torch.int64, device=dev),
            torch.tensor([int(ev[1]) for ev in events], dtype=torch.int64, device=dev),
            torch.tensor([int(ev[2]) for ev in events], dtype=torch.int64, device=dev), rows)


def _packed_global(events):
    """(global atom, batch, unit time, amplitude) arrays of a global event list."""
    packed = getattr(events, "packed", None)
    if packed is not None and len(events) == packed[0].numel():
        return packed
    if len(events) == 0:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z, torch.zeros(0), torch.zeros(0)
    as_t = (lambda x: x.reshape(()) if isinstance(x, torch.Tensor) else torch.tensor(float(x)))
    time = torch.stack([as_t(ev[2]).float() for ev in events])
    amp = torch.stack([as_t(ev[3]).float().to(time.device) for ev in events])
    dev = time.device
    return (torch.tensor([int(ev[0]) for ev in events], dtype=torch.int64, device=dev),
            torch.tensor([int(ev[1]) for ev in events], dtype=torch.int64, device=dev), time, amp)


def _unique_first(x: torch.Tensor):
    """Distinct values of a 1-D integer tensor and the index of each one's first occurrence."""
    uniq, inverse = torch.unique(x, return_inverse=True)
    first = torch.full((uniq.numel(),), x.numel(), dtype=torch.int64, device=x.device)
    first = first.scatter_reduce(0, inverse, torch.arange(x.numel(), device=x.device), reduce="amin")
    return uniq, first


def _unit_norm(d: torch.Tensor) -> torch.Tensor:
    """modules/normalization.py:4-6 on whatever device ``d`` lives (CUDA kernel when it can)."""
    if d.is_cuda:
        return engine.unit_norm(d)
    return d / (torch.norm(d, dim=-1, keepdim=True) + 1e-8)


class BandSpec(object):
    """One band's dictionary and codec (modules/multibanddict.py:53-279)."""

    def __init__(self, size: int, n_atoms: int, atom_size: int, slce: Optional[slice] = None, device=None,
                 signal_samples: int = 0, samplerate=22050, local_contrast_norm: bool = False,
                 is_lowest_band: bool = False):
        self.is_lowest_band = is_lowest_band
        self.signal_samples = signal_samples
        self.size = size
        self.n_atoms = n_atoms
        self.atom_size = atom_size
        self.slce = slce
        self.device = device
        self.samplerate = samplerate
        self.local_contrast_norm = local_contrast_norm
        d = torch.zeros(n_atoms, atom_size, requires_grad=False).uniform_(-1, 1).to(device)    # :89-90
        self.d = _unit_norm(d)
        self._embeddings = None

    def __hash__(self):                                                      # :95-97
        return hash(sha256(self.d.data.cpu().numpy()).hexdigest())

    @property
    def n_samples_at_native_rate(self):                                      # :100-107
        return self.atom_size * (self.signal_samples // self.size)

    def resampled_atoms(self) -> torch.Tensor:                               # :109-115
        return fft_resample(self.d.view(self.n_atoms, 1, self.atom_size), self.n_samples_at_native_rate,
                            self.is_lowest_band)

    def shape(self, batch_size):                                             # :153-154
        return (batch_size, 1, self.size)

    @property
    def filename(self):                                                      # :156-158
        return f"band_{self.size}.dat"

    @property
    def scatter_func(self):                                                  # :160-162
        return build_scatter_segments(self.size, self.atom_size)

    def get_atom(self, index: int, norm):                                    # :164-165
        return self.d[index] * norm

    def load(self):                                                          # :167-173
        try:
            self.d = torch.load(self.filename)
        except IOError:
            print(f"failed to load {self.filename}")

    def store(self):                                                         # :175-176
        torch.save(self.d, self.filename)

    def learn(self, batch, steps=16):                                        # :178-187
        d = dictionary_learning_step(batch, self.d, steps, device=self.device, approx=self.slce,
                                     local_constrast_norm=self.local_contrast_norm)
        self.d = _unit_norm(d)
        return d

    def to_global_atom_index(self, index: int, offset: int) -> int:         # :189-190
        return offset + index

    def to_local_atom_index(self, index: int, offset: int) -> int:          # :192-193
        return index - offset

    def to_unit_time(self, sample_position):                                 # :195-196
        return sample_position / self.size

    def to_sample_time(self, unit_time) -> int:                              # :198-199
        return int(unit_time * self.size)

    def to_amplitude(self, scaled_atom: torch.Tensor):                       # :201-202
        return torch.norm(scaled_atom)

    def to_global_arrays(self, events, offset: int):
        """The whole event list of this band as global arrays (:204-217 for every event at once):
        (global atom, batch, unit time, amplitude)."""
        atom, batch, pos, rows = _packed_local(events, self.atom_size)
        return atom + offset, batch, pos.reshape(-1) / self.size, torch.norm(rows, dim=-1)

    def to_local_events(self, global_atom, batch, unit_time, amplitude, offset: int) -> EventList:
        """Global arrays of events of THIS band back to a local event list (:219-235 for every event at once):
        sample position = trunc(unit_time * size), scaled atom = d[local] * amplitude."""
        local = global_atom - offset
        pos = (unit_time * self.size).to(torch.int64)                        # int(): truncation towards zero
        d = self.d
        rows = d[local.to(d.device)] * amplitude.to(d.device).reshape(-1, 1)
        out = _LocalEventList()
        out.packed = (local, batch, pos, rows)
        out._lazy = local.numel() > 0
        return out

    def to_global_tuple(self, event: LocalEventTuple, offset: int) -> GlobalEventTuple:   # :204-217
        atom_index, batch, sample_pos, atom = event
        return (self.to_global_atom_index(atom_index, offset), batch, self.to_unit_time(sample_pos),
                self.to_amplitude(atom))

    def to_local_tuple(self, event: GlobalEventTuple, offset: int) -> LocalEventTuple:    # :219-235
        global_index, batch, unit_time, amplitude = event
        local_index = self.to_local_atom_index(global_index, offset)
        return (local_index, batch, self.to_sample_time(unit_time), self.get_atom(local_index, amplitude))

    def encode_start(self, batch, steps=16, extract_embeddings=None, out_device=None):
        """Enqueue this band's pursuit on the current stream; ``encode_finish`` collects it."""
        return sparse_code_start(batch, self.d, steps, device=self.device, approx=self.slce, flatten=True,
                                 extract_atom_embedding=extract_embeddings,
                                 local_contrast_norm=self.local_contrast_norm,
                                 out_device=out_device), batch.shape, bool(extract_embeddings)

    @staticmethod
    def encode_finish(started) -> BandEncodingPackage:
        job, shape, embeddings = started
        encoding = job.result()
        if embeddings:
            return encoding                                                  # (embeddings, residual), :250-252
        instances, scatter = encoding
        return instances, scatter, shape

    def encode(self, batch, steps=16, extract_embeddings=None) -> BandEncodingPackage:    # :238-263
        return self.encode_finish(self.encode_start(batch, steps, extract_embeddings))

    def decode(self, shape, all_instances, scatter):                         # :265-266
        return scatter(shape, all_instances)

    def recon(self, batch, steps=16):                                        # :268-279
        all_instances, scatter, shape = self.encode(batch, steps)
        return self.decode(shape, all_instances, scatter), all_instances, scatter


class MultibandDictionaryLearning(object):
    """Per-band pursuit over an octave band split (modules/multibanddict.py:282-473)."""

    def __init__(self, specs: List[BandSpec], n_samples: int):
        self.bands = OrderedDict((spec.size, spec) for spec in specs)
        self.min_size = min(spec.size for spec in specs)
        self.n_samples = n_samples
        n_atoms = set(spec.n_atoms for spec in specs)
        if len(n_atoms) > 1:                                                 # :289-291
            raise ValueError("Only specs with equal atom counts is currently allowed")
        self.n_atoms = list(n_atoms)[0]
        self._embeddings = None

    def __len__(self):                                                       # :300-301
        return len(self.bands)

    def event_count(self, iterations: int) -> int:                           # :303-304
        return len(self) * iterations

    def get_atom(self, size, index, norm):                                   # :345-346
        return self.bands[size].get_atom(index, norm)

    def size_at_index(self, index):                                          # :348-349
        return list(self.bands.keys())[index]

    def index_of_size(self, band_size):                                      # :351-354
        return [b.size for b in self.bands.values()].index(band_size)

    def shape_dict(self, batch_size):                                        # :356-357
        return {size: band.shape(batch_size) for size, band in self.bands.items()}

    @property
    def total_atoms(self):                                                   # :367-369
        return sum(v.n_atoms for v in self.bands.values())

    @property
    def band_dicts(self):                                                    # :371-373
        return {size: band.d for size, band in self.bands.items()}

    @property
    def band_sizes(self):                                                    # :375-377
        return list(self.bands.keys())

    def partial_decoding_dict(self, batch_size):                             # :379-384
        return {size: (build_scatter_segments(size, self.bands[size].atom_size), (batch_size, 1, size))
                for size in self.bands.keys()}

    def store(self):                                                         # :386-388
        for band in self.bands.values():
            band.store()

    def load(self):                                                          # :390-392
        for band in self.bands.values():
            band.load()

    def learn(self, batch, steps=16):                                        # :394-397
        bands = fft_frequency_decompose(batch, self.min_size)
        for size, band in bands.items():
            self.bands[size].learn(band, steps)

    def encode(self, batch, steps, extract_embeddings=None) -> Dict[int, BandEncodingPackage]:   # :399-404
        """The bands are independent pursuits: each is enqueued on its own stream behind the band split, and only
        then are the results collected (the reference's serial loop over the bands, run concurrently)."""
        dev = batch.device if batch.is_cuda else engine._require_cuda(None)
        # a host batch is uploaded ONCE and split on the device; the results go back to the batch's own device
        staged = batch if batch.is_cuda else batch.to(dev, non_blocking=True)
        bands = fft_frequency_decompose(staged, self.min_size)
        main = torch.cuda.current_stream(dev)
        started = OrderedDict()
        for i, (size, band) in enumerate(self.bands.items()):
            side = self._stream(dev, i)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                started[size] = band.encode_start(bands[size], steps, extract_embeddings, out_device=batch.device)
        out = OrderedDict((size, BandSpec.encode_finish(job)) for size, job in started.items())
        for i in range(len(self.bands)):
            main.wait_stream(self._stream(dev, i))
        return out

    def _stream(self, dev, i: int):
        key = (dev.index, i)
        if not hasattr(self, "_streams"):
            self._streams = {}
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=dev)
        return self._streams[key]

    def get_band_from_global_atom_index(self, index: int) -> Tuple[int, BandSpec]:    # :406-408
        band_index = index // self.n_atoms
        return band_index, list(self.bands.values())[band_index]

    def flattened_event_tuples(self, encoding: Dict[int, BandEncodingPackage]) -> List[GlobalEventTuple]:   # :410-422
        """Every band's events as ``(global atom, batch, unit time, amplitude)``, bands in encoding order with
        offsets n_atoms apart.  One array conversion per band; as in the reference the unit time is a (1,1) tensor
        and the amplitude a 0-d tensor."""
        parts, offset = [], 0
        for size, (events, _, _) in encoding.items():
            band = self.bands[size]
            parts.append(band.to_global_arrays(events, offset))
            offset += band.n_atoms
        out = GlobalEventList()
        if not parts:
            return out
        atom, batch, time, amp = (torch.cat([p[i] for p in parts]) for i in range(4))
        out.packed = (atom, batch, time, amp)
        out._lazy = atom.numel() > 0
        return out

    def hierarchical_event_tuples(self, encoding: List[GlobalEventTuple],
                                  original: Dict[int, BandEncodingPackage]) -> Dict[int, BandEncodingPackage]:   # :424-443
        """Back to per-band local event lists: bands keyed in order of their first event, events of a band in list
        order (the reference's defaultdict), each converted with its band's size and dictionary."""
        atom, batch, time, amp = _packed_global(encoding)
        final = OrderedDict()
        if atom.numel() == 0:
            return final
        band_of = torch.div(atom, self.n_atoms, rounding_mode="floor")                 # :406-408
        specs = list(self.bands.values())
        uniq, first = _unique_first(band_of)
        for band_index in uniq[torch.argsort(first)].tolist():                         # first-seen order
            pick = torch.nonzero(band_of == band_index).reshape(-1)                    # ascending = list order
            band = specs[band_index]
            events = band.to_local_events(atom[pick], batch[pick], time[pick], amp[pick], band_index * self.n_atoms)
            _, scatter, shape = original[band.size]
            final[band.size] = (events, scatter, shape)
        return final

    def decode(self, d, shapes=None):                                        # :446-458
        output = OrderedDict()
        for size, tup in d.items():
            if shapes is not None:
                all_instances, scatter, shape = tup, self.bands[size].scatter_func, shapes[size]
            else:
                all_instances, scatter, shape = tup
            output[size] = self.bands[size].decode(shape, all_instances, scatter)
        return fft_frequency_recompose(output, self.n_samples)

    def recon(self, batch, steps=16):                                        # :460-473
        bands = fft_frequency_decompose(batch, self.min_size)
        recon_bands, events = OrderedDict(), OrderedDict()
        for size in self.bands.keys():
            r, e, _ = self.bands[size].recon(bands[size], steps)
            recon_bands[size] = r
            events[size] = e
        return fft_frequency_recompose(recon_bands, batch.shape[-1]), events
