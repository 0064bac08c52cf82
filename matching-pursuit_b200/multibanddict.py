"""Counterparts of the reference's ``modules/multibanddict.py``: ``BandSpec``
(lines 53-279) and ``MultibandDictionaryLearning`` (lines 282-473) -- one
greedy pursuit per octave band of an FFT band split.  Types follow the
reference (:12-16): ``LocalEventTuple = (atom:int, batch:int, position, atom
tensor)``, ``GlobalEventTuple = (global atom:int, batch:int, unit time,
amplitude)``, ``BandEncodingPackage = (events, scatter, shape)``.

The per-band pursuits are independent problems (they share nothing but the
step count), so each runs on the CUDA engine through :func:`sparse_code`; the
band split and merge run through ``mpb200_spectral_band``.  Dictionary
learning (``learn``) keeps its atom update in PyTorch.  The STFT loss features
of the reference module (``multiband_spectrogram*``, :19-49) are out of scope."""
from __future__ import annotations

from collections import OrderedDict, defaultdict
from hashlib import sha256
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import engine
from .decompose import fft_frequency_decompose, fft_frequency_recompose, fft_resample
from .matchingpursuit import build_scatter_segments, dictionary_learning_step, sparse_code

Shape = Tuple
LocalEventTuple = Tuple[int, int, int, torch.Tensor]
GlobalEventTuple = Tuple[int, int, float, float]
BandEncodingPackage = Tuple[List[LocalEventTuple], Callable, Shape]


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

    def to_global_tuple(self, event: LocalEventTuple, offset: int) -> GlobalEventTuple:   # :204-217
        atom_index, batch, sample_pos, atom = event
        return (self.to_global_atom_index(atom_index, offset), batch, self.to_unit_time(sample_pos),
                self.to_amplitude(atom))

    def to_local_tuple(self, event: GlobalEventTuple, offset: int) -> LocalEventTuple:    # :219-235
        global_index, batch, unit_time, amplitude = event
        local_index = self.to_local_atom_index(global_index, offset)
        return (local_index, batch, self.to_sample_time(unit_time), self.get_atom(local_index, amplitude))

    def encode(self, batch, steps=16, extract_embeddings=None) -> BandEncodingPackage:    # :238-263
        encoding = sparse_code(batch, self.d, steps, device=self.device, approx=self.slce, flatten=True,
                               extract_atom_embedding=extract_embeddings,
                               local_contrast_norm=self.local_contrast_norm)
        if extract_embeddings:
            return encoding
        instances, scatter = encoding
        return instances, scatter, batch.shape

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
        bands = fft_frequency_decompose(batch, self.min_size)
        return OrderedDict((size, band.encode(bands[size], steps, extract_embeddings))
                           for size, band in self.bands.items())

    def get_band_from_global_atom_index(self, index: int) -> Tuple[int, BandSpec]:    # :406-408
        band_index = index // self.n_atoms
        return band_index, list(self.bands.values())[band_index]

    def flattened_event_tuples(self, encoding: Dict[int, BandEncodingPackage]) -> List[GlobalEventTuple]:   # :410-422
        output = []
        offset = 0
        for size, package in encoding.items():
            events, scatter, shape = package
            band = self.bands[size]
            for event in events:
                output.append(band.to_global_tuple(event, offset))
            offset += band.n_atoms
        return output

    def hierarchical_event_tuples(self, encoding: List[GlobalEventTuple],
                                  original: Dict[int, BandEncodingPackage]) -> Dict[int, BandEncodingPackage]:   # :424-443
        hierarchical = defaultdict(list)
        for event in encoding:
            global_index, batch, unit_time, amplitude = event
            index, band = self.get_band_from_global_atom_index(global_index)
            hierarchical[band.size].append(band.to_local_tuple(event, index * self.n_atoms))
        final = OrderedDict()
        for size, events in hierarchical.items():
            _, scatter, shape = original[size]
            final[size] = (events, scatter, shape)
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
