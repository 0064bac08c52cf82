"""Thin torch-side wrapper of an ``mpb200_plan_t`` (include/mpb200.h).

PyTorch is plumbing here: it owns device buffers and the current CUDA stream;
every computation is a call through the C ABI into the sm_100a kernels.  There
is no CPU path -- constructing a :class:`Plan` without the built library or
without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import MODES, MODE_NAMES, MpbError, PlanInfo, check, lib


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise MpbError("no CUDA device: matching-pursuit_b200 has no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise MpbError(f"device must be a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _dev_f32(x: torch.Tensor, device: torch.device, shape=None) -> torch.Tensor:
    x = x.detach()
    if x.dtype != torch.float32 or x.device != device or not x.is_contiguous():
        x = x.to(device=device, dtype=torch.float32).contiguous()
    return x if shape is None else x.view(*shape)


class Plan:
    """Workspaces + derived dictionary tables for signals of ``n_samples``,
    batches up to ``max_batch`` and a ``(n_atoms, atom_size)`` dictionary of
    which this plan owns atoms ``[atom_lo, atom_hi)``."""

    def __init__(self, n_atoms: int, atom_size: int, n_samples: int, max_batch: int, mode: str = "auto",
                 atom_range: Optional[Tuple[int, int]] = None, device=None, gram_budget_bytes: int = 0):
        self.device = _require_cuda(device)
        self._lib = lib()
        lo, hi = (0, n_atoms) if atom_range is None else atom_range
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_plan_create(C.byref(handle), n_atoms, atom_size, n_samples, max_batch,
                                               MODES[mode] if isinstance(mode, str) else int(mode), lo, hi,
                                               C.c_uint64(gram_budget_bytes)), "mpb200_plan_create")
        self._handle = handle
        self.pins = 0                # users that hold the plan across calls (plan caches must not close a pinned plan)
        self._finalizer = weakref.finalize(self, Plan._destroy, self._lib, handle)
        info = PlanInfo()
        check(self._lib.mpb200_plan_info_get(self._h, C.byref(info)), "mpb200_plan_info_get")
        self.info = info
        self.n_atoms, self.atom_size, self.n_samples, self.max_batch = n_atoms, atom_size, n_samples, max_batch
        self.atom_lo, self.atom_hi = info.atom_lo, info.atom_hi
        self.mode = MODE_NAMES[info.mode]
        self.fft_size, self.block, self.n_blocks = info.fft_size, info.block, info.n_blocks
        self.device_bytes = info.device_bytes
        self.resident_batch = info.resident_batch   # signals processed at once (sub-batch of the map modes)
        self.fft_size2 = info.fft_size2
        self.batch = 0
        self.dictionary_key = None   # set by callers that cache plans

    @staticmethod
    def _destroy(library, handle):
        if handle:
            library.mpb200_plan_destroy(handle)

    @property
    def _h(self):
        """The native handle; a closed plan raises instead of handing freed memory to the library."""
        if self._handle is None:
            raise MpbError("this plan has been closed (mpb200_plan_destroy has run); create a new one")
        return self._handle

    @property
    def closed(self) -> bool:
        return self._handle is None

    def close(self):
        self._finalizer()
        self._handle = None

    def __enter__(self):
        self.pins += 1
        return self

    def __exit__(self, *exc):
        self.pins -= 1
        return False

    def set_refresh_every(self, iterations: int) -> "Plan":
        """GRAM mode: re-correlate the whole map every ``iterations`` steps (0 = never)."""
        check(self._lib.mpb200_plan_set_option(self._h, _lib.OPT_REFRESH_EVERY, int(iterations)), "mpb200_plan_set_option")
        return self

    def set_fused_loop(self, on: bool) -> "Plan":
        """Windowed re-correlation mode: run the iteration loop of ``sparse_code`` as one cooperative launch when the
        shape allows it (MPB200_OPT_FUSED_LOOP, default on); off = two launches per iteration."""
        check(self._lib.mpb200_plan_set_option(self._h, _lib.OPT_FUSED_LOOP, int(bool(on))), "mpb200_plan_set_option")
        return self

    def set_position_free(self, on: bool) -> "Plan":
        """SGRAM mode: force the position-free block tables on or off (chosen automatically otherwise)."""
        check(self._lib.mpb200_plan_set_option(self._h, _lib.OPT_POSITION_FREE, int(bool(on))), "mpb200_plan_set_option")
        return self

    def set_local_contrast_norm(self, on: bool) -> "Plan":
        """GRAM / SGRAM mode: select on ``fm - avg_pool2d(fm, 9x9)`` incrementally (include/mpb200.h,
        MPB200_OPT_LOCAL_CONTRAST_NORM; modules/matchingpursuit.py:286-296)."""
        check(self._lib.mpb200_plan_set_option(self._h, _lib.OPT_LOCAL_CONTRAST_NORM, int(bool(on))), "mpb200_plan_set_option")
        self.local_contrast_norm = bool(on)
        return self

    # ---- per-kernel timing (bench aid) ----------------------------------
    def timing(self, enable: bool) -> None:
        check(self._lib.mpb200_plan_timing_enable(self._h, int(bool(enable))), "mpb200_plan_timing_enable")

    def timing_read(self) -> dict:
        """{'first_pass'|'apply'|'recorrelate'|'gram_update': (milliseconds, intervals)} since the last read."""
        ms = (C.c_double * 5)()
        cnt = (C.c_int64 * 5)()
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_plan_timing_read(self._h, ms, cnt), "mpb200_plan_timing_read")
        return {"first_pass": (ms[1], cnt[1]), "apply": (ms[2], cnt[2]), "recorrelate": (ms[3], cnt[3]),
                "gram_update": (ms[4], cnt[4])}

    # ---- dictionary -----------------------------------------------------
    def set_dictionary(self, d: torch.Tensor, normalize: bool = True) -> "Plan":
        """``d`` is (K, A) (or (K, 1, A)); it is unit-normed on the device
        exactly like modules/normalization.py:4-6 (unless ``normalize`` is
        False: atoms used as given) and not modified."""
        d = _dev_f32(d, self.device).reshape(d.shape[0], -1)
        if tuple(d.shape) != (self.n_atoms, self.atom_size):
            raise MpbError(f"dictionary shape {tuple(d.shape)} does not match the plan "
                           f"({self.n_atoms}, {self.atom_size})")
        with torch.cuda.device(self.device):
            fn = self._lib.mpb200_plan_set_dictionary if normalize else self._lib.mpb200_plan_set_dictionary_raw
            check(fn(self._h, _ptr(d), _stream_ptr(self.device)), "mpb200_plan_set_dictionary")
        self._keep = d  # stream-ordered use: keep alive until the next call replaces it
        return self

    def unit_dictionary(self) -> torch.Tensor:
        out = torch.empty(self.n_atoms, self.atom_size, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_plan_get_unit_dictionary(self._h, _ptr(out), _stream_ptr(self.device)),
                  "mpb200_plan_get_unit_dictionary")
        return out

    # ---- whole pursuit --------------------------------------------------
    def _check_signal(self, signal: torch.Tensor) -> Tuple[int, torch.Tensor]:
        b = signal.shape[0]
        if signal.numel() != b * self.n_samples:
            raise MpbError(f"signal of shape {tuple(signal.shape)} does not hold {self.n_samples} samples per row")
        if not 1 <= b <= self.max_batch:
            raise MpbError(f"batch {b} is outside [1, {self.max_batch}]")
        return b, signal

    def sparse_code(self, signal: torch.Tensor, n_steps: int, want_residual: bool = True):
        """Device-resident pursuit.  Returns ``(atom, pos, val, residual)``:
        int32 (B,S), int32 (B,S), float32 (B,S), float32 (B,N) or None."""
        b, _ = self._check_signal(signal)
        sig = _dev_f32(signal, self.device, (b, self.n_samples))
        atom = torch.empty(b, n_steps, device=self.device, dtype=torch.int32)
        pos = torch.empty(b, n_steps, device=self.device, dtype=torch.int32)
        val = torch.empty(b, n_steps, device=self.device, dtype=torch.float32)
        res = torch.empty(b, self.n_samples, device=self.device, dtype=torch.float32) if want_residual else None
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_sparse_code(self._h, _ptr(sig), b, n_steps, _ptr(res), _ptr(atom), _ptr(pos),
                                               _ptr(val), _stream_ptr(self.device)), "mpb200_sparse_code")
        self.batch = 0
        return atom, pos, val, res

    def sparse_code_host(self, signal: torch.Tensor, n_steps: int, want_residual: bool = True, out=None):
        """End-to-end entry for HOST tensors: H2D, pursuit, D2H, synchronised.
        ``out`` may hold preallocated (pinned) ``(atom, pos, val, residual)``."""
        b, _ = self._check_signal(signal)
        if signal.device.type != "cpu" or signal.dtype != torch.float32 or not signal.is_contiguous():
            signal = signal.detach().to("cpu", torch.float32).contiguous()
        if out is None:
            atom = torch.empty(b, n_steps, dtype=torch.int32)
            pos = torch.empty(b, n_steps, dtype=torch.int32)
            val = torch.empty(b, n_steps, dtype=torch.float32)
            res = torch.empty(b, self.n_samples, dtype=torch.float32) if want_residual else None
        else:
            atom, pos, val, res = out
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_sparse_code_host(self._h, _ptr(signal), b, n_steps, _ptr(res), _ptr(atom),
                                                    _ptr(pos), _ptr(val), _stream_ptr(self.device)),
                  "mpb200_sparse_code_host")
        self.batch = 0
        return atom, pos, val, res

    def correlate(self, signal: torch.Tensor) -> torch.Tensor:
        """Dense (B, owned atoms, N) correlation map of ``signal`` (B, N)."""
        b, _ = self._check_signal(signal)
        sig = _dev_f32(signal, self.device, (b, self.n_samples))
        fm = torch.empty(b, self.atom_hi - self.atom_lo, self.n_samples, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_correlate(self._h, _ptr(sig), b, _ptr(fm), _stream_ptr(self.device)),
                  "mpb200_correlate")
        self.batch = 0
        return fm

    # ---- step-wise interface (atom sharding, per-step callbacks) ---------
    def begin(self, signal: torch.Tensor) -> None:
        b, _ = self._check_signal(signal)
        sig = _dev_f32(signal, self.device, (b, self.n_samples))
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_begin(self._h, _ptr(sig), b, _stream_ptr(self.device)), "mpb200_begin")
        self.batch = b

    def local_best(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, 4) int32 view of ``mpb200_best`` records: value bits, atom, position, pad."""
        if out is None:
            out = torch.empty(self.batch, 4, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_local_best(self._h, _ptr(out), _stream_ptr(self.device)), "mpb200_local_best")
        return out

    def apply(self, winner: torch.Tensor) -> None:
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_apply(self._h, _ptr(winner), _stream_ptr(self.device)), "mpb200_apply")

    # ---- atom-sharded exchange over peer memory --------------------------
    def exchange_create(self, world: int, rank: int) -> bytes:
        """Allocate this rank's mailbox; returns its 64-byte CUDA IPC handle for the peers."""
        buf = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_exchange_create(self._h, world, rank, buf), "mpb200_exchange_create")
        return bytes(buf)

    def exchange_connect(self, handles: bytes) -> None:
        """``handles``: the ``world`` IPC handles in rank order (64 bytes each), other processes' mailboxes."""
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_exchange_connect(self._h, buf), "mpb200_exchange_connect")

    def exchange_mailbox(self) -> int:
        out = C.c_void_p()
        check(self._lib.mpb200_exchange_mailbox(self._h, C.byref(out)), "mpb200_exchange_mailbox")
        return out.value

    def exchange_connect_local(self, mailboxes) -> None:
        """Same-process form (several plans in one process): ``mailboxes[r]`` from :meth:`exchange_mailbox`."""
        arr = (C.c_void_p * len(mailboxes))(*mailboxes)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_exchange_connect_local(self._h, arr), "mpb200_exchange_connect_local")

    def exchange_disconnect(self) -> None:
        """Unmap the peers' mailboxes (collective teardown, see distributed.AtomShardedPursuit.close)."""
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_exchange_disconnect(self._h), "mpb200_exchange_disconnect")

    def exchange_timed_out(self) -> bool:
        flag = C.c_int(0)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_exchange_status(self._h, C.byref(flag)), "mpb200_exchange_status")
        return bool(flag.value)

    def residual(self) -> torch.Tensor:
        out = torch.empty(self.batch, self.n_samples, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self._lib.mpb200_residual(self._h, _ptr(out), _stream_ptr(self.device)), "mpb200_residual")
        return out


def reduce_best(cand: torch.Tensor, n_ranks: int, batch: int) -> torch.Tensor:
    """Global winner per signal from rank-major candidates ``(n_ranks*batch, 4)`` int32."""
    dev = cand.device
    out = torch.empty(batch, 4, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib().mpb200_reduce_best(_ptr(cand), n_ranks, batch, _ptr(out), _stream_ptr(dev)), "mpb200_reduce_best")
    return out


def unpack_best(rec: torch.Tensor):
    """(value float32, atom int32, position int32) columns of best records."""
    return rec[:, 0].view(torch.float32), rec[:, 1], rec[:, 2]


# ---- stateless helpers ------------------------------------------------------

def unit_norm(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Rows of a CUDA tensor divided by (||row||_2 + eps) -- modules/normalization.py:4-6."""
    dev = _require_cuda(x.device)
    x2 = _dev_f32(x, dev).reshape(-1, x.shape[-1])
    y = torch.empty_like(x2)
    with torch.cuda.device(dev):
        check(lib().mpb200_unit_norm(_ptr(x2), _ptr(y), x2.shape[0], x2.shape[1], C.c_float(eps), _stream_ptr(dev)),
              "mpb200_unit_norm")
    return y.view(x.shape)


def gather_atoms(d_unit: torch.Tensor, atom: torch.Tensor, val: torch.Tensor) -> torch.Tensor:
    """``val[e] * d_unit[atom[e]]`` as an (E, A) tensor (modules/matchingpursuit.py:305)."""
    dev = d_unit.device
    n = atom.numel()
    out = torch.empty(n, d_unit.shape[1], device=dev, dtype=torch.float32)
    if n:
        atom = atom.to(torch.int32).contiguous()
        val = val.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            check(lib().mpb200_gather_atoms(_ptr(out), _ptr(d_unit), d_unit.shape[0], d_unit.shape[1], _ptr(atom),
                                            _ptr(val), n, _stream_ptr(dev)), "mpb200_gather_atoms")
    return out


def _row_offsets(row_index: torch.Tensor, n_rows: int):
    """Stable sort of the events by destination row -> (order, int32 offsets (n_rows+1))."""
    order = torch.sort(row_index.to(torch.int64), stable=True)[1]
    counts = torch.bincount(row_index.to(torch.int64), minlength=n_rows)
    offsets = torch.zeros(n_rows + 1, device=row_index.device, dtype=torch.int32)
    offsets[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return order, offsets


def scatter_add(out: torch.Tensor, d_unit: torch.Tensor, atom, batch_index, pos, val) -> torch.Tensor:
    """``out[b, p:p+A] += val * d_unit[atom]`` for every event in list order,
    truncated at the right edge (modules/matchingpursuit.py:20-58, :48)."""
    dev = out.device
    b, n = out.shape[0], out.shape[-1]
    ne = atom.numel()
    if ne:
        i32 = lambda t: t.reshape(-1).to(device=dev, dtype=torch.int32)
        atom, batch_index, pos = i32(atom), i32(batch_index), i32(pos)
        val = val.reshape(-1).to(device=dev, dtype=torch.float32)
        order, offsets = _row_offsets(batch_index, b)
        atom, batch_index, pos, val = (t[order].contiguous() for t in (atom, batch_index, pos, val))
        with torch.cuda.device(dev):
            check(lib().mpb200_scatter_add(_ptr(out), b, n, _ptr(d_unit), d_unit.shape[0], d_unit.shape[1],
                                           _ptr(atom), _ptr(batch_index), _ptr(pos), _ptr(val), _ptr(offsets), ne,
                                           _stream_ptr(dev)), "mpb200_scatter_add")
    return out


def scatter_rows(out: torch.Tensor, rows: torch.Tensor, row_index, pos) -> torch.Tensor:
    """``out2d[row_index[e], pos[e]:pos[e]+A] += rows[e]`` in list order, truncated
    at the right edge; ``out`` is viewed as (-1, N) (modules/matchingpursuit.py:41-52)."""
    dev = out.device
    n = out.shape[-1]
    n_rows = out.numel() // n
    ne = rows.shape[0]
    if ne:
        row_index = row_index.reshape(-1).to(device=dev, dtype=torch.int32)
        pos = pos.reshape(-1).to(device=dev, dtype=torch.int32)
        order, offsets = _row_offsets(row_index, n_rows)
        rows = rows.to(device=dev, dtype=torch.float32)[order].contiguous()
        row_index, pos = row_index[order].contiguous(), pos[order].contiguous()
        with torch.cuda.device(dev):
            check(lib().mpb200_scatter_rows(_ptr(out), n_rows, n, _ptr(rows), rows.shape[1], _ptr(row_index),
                                            _ptr(pos), _ptr(offsets), ne, _stream_ptr(dev)), "mpb200_scatter_rows")
    return out


def select_dense(fm: torch.Tensor, atom_offset: int = 0, local_contrast_norm: bool = False) -> torch.Tensor:
    """Per-signal signed argmax of a dense (B, K, N) CUDA map with the
    reference tie-break -> (B, 4) int32 best records
    (modules/matchingpursuit.py:298-303; with ``local_contrast_norm`` :286-296)."""
    dev = _require_cuda(fm.device)
    b, k, n = fm.shape
    fm = _dev_f32(fm, dev)
    out = torch.empty(b, 4, device=dev, dtype=torch.int32)
    fn = lib().mpb200_select_lcn if local_contrast_norm else lib().mpb200_select_dense
    with torch.cuda.device(dev):
        check(fn(_ptr(fm), b, k, n, atom_offset, _ptr(out), _stream_ptr(dev)), "mpb200_select_dense")
    return out


def subtract(residual: torch.Tensor, d_unit: torch.Tensor, winner: torch.Tensor) -> torch.Tensor:
    """In place: ``residual[b, p:p+A] -= value * d_unit[atom]`` for one winner per
    signal (modules/matchingpursuit.py:326-328)."""
    dev = residual.device
    b, n = residual.shape[0], residual.shape[-1]
    with torch.cuda.device(dev):
        check(lib().mpb200_subtract(_ptr(residual), b, n, _ptr(d_unit), d_unit.shape[0], d_unit.shape[1],
                                    _ptr(winner), _stream_ptr(dev)), "mpb200_subtract")
    return residual


def band_limit(x: torch.Tensor, length: int, slce: slice) -> torch.Tensor:
    """``irfft(mask(rfft(pad(x, length))))`` per row of the CUDA tensor ``x`` (rows, n): the mask keeps the
    rfft bins ``slce`` selects and zeroes the others -- the spectral mask of modules/conv.py:24-29.
    ``length`` (= n_samples + atom_size there) must be even.  Returns (rows, length)."""
    dev = _require_cuda(x.device)
    x2 = _dev_f32(x, dev).reshape(-1, x.shape[-1])
    n_bins_total = length // 2 + 1
    if slce.step is not None and slce.step < 1:
        raise ValueError("step must be greater than zero")      # what torch says of such a slice in the reference
    bins = range(*slce.indices(n_bins_total))
    out = torch.empty(x2.shape[0], length, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().mpb200_band_limit(_ptr(x2), x2.shape[0], x2.shape[1], length, bins[0] if len(bins) else 0,
                                      bins.step if len(bins) else 1, len(bins), _ptr(out), _stream_ptr(dev)),
              "mpb200_band_limit")
    return out


MAX_PLAN_ATOM = 2560      # MPB200_MAX_PLAN_ATOM: longest atom one window transform of a plan can hold
LONG_PART = 2048          # long atoms are correlated as consecutive parts of this many samples


def split_long_atoms(atoms: torch.Tensor, part: int = LONG_PART):
    """(K, A) -> ((K*P, part) parts, P): part p of atom k in row k*P + p, the last part zero padded."""
    k, a = atoms.shape
    p = -(-a // part)
    padded = torch.zeros(k, p * part, device=atoms.device, dtype=torch.float32)
    padded[:, :a] = atoms
    return padded.view(k * p, part).contiguous(), p


def fold_parts(sub_map: torch.Tensor, n_atoms: int, n_parts: int, part_len: int) -> torch.Tensor:
    """``fm[b,k,t] = sum_p sub_map[b, k*P+p, t+p*L]`` (include/mpb200.h, mpb200_fold_parts)."""
    b, _, n = sub_map.shape
    dev = sub_map.device
    out = torch.empty(b, n_atoms, n, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().mpb200_fold_parts(_ptr(sub_map), b, n_atoms, n_parts, part_len, n, _ptr(out), _stream_ptr(dev)),
              "mpb200_fold_parts")
    return out


def correlate_long(signal2d: torch.Tensor, parts_plan: "Plan", n_atoms: int, n_parts: int,
                   max_bytes: int = 2 << 30) -> torch.Tensor:
    """Dense (B, K, N) correlation map of atoms longer than a plan can take: the plan holds the dictionary of their
    parts (see :func:`split_long_atoms`); signals are walked in slices whose parts' map stays under ``max_bytes``."""
    b, n = signal2d.shape
    per_signal = n_atoms * n_parts * n * 4
    step = max(1, min(b, max_bytes // max(per_signal, 1)))
    outs = []
    for b0 in range(0, b, step):
        sub = parts_plan.correlate(signal2d[b0:b0 + step])
        outs.append(fold_parts(sub, n_atoms, n_parts, parts_plan.atom_size))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)


def correlate_gemm(signal2d: torch.Tensor, atoms: torch.Tensor, spin_blocks: int = 0) -> Optional[torch.Tensor]:
    """The dense (B, K, N) correlation map of ``atoms`` (K, A), used as given, with ``signal2d`` (B, N) on the
    tensor cores: 3xTF32 Toeplitz GEMM (include/mpb200.h, mpb200_correlate_gemm).  ``spin_blocks`` > 0: tensor-pipe
    peak probe, returns None."""
    dev = _require_cuda(signal2d.device)
    sig = _dev_f32(signal2d, dev)
    d = _dev_f32(atoms, dev)
    b, n = sig.shape
    out = None if spin_blocks > 0 else torch.empty(b, d.shape[0], n, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().mpb200_correlate_gemm(_ptr(sig), b, n, _ptr(d), d.shape[0], d.shape[1], _ptr(out), int(spin_blocks),
                                          _stream_ptr(dev)), "mpb200_correlate_gemm")
    return out


def dictionary_update(running: torch.Tensor, d_unit: torch.Tensor, group_offsets: torch.Tensor, group_atom: torch.Tensor,
                      ev_batch: torch.Tensor, ev_pos: torch.Tensor, ev_rows: torch.Tensor) -> None:
    """In place: the atom-update loop of modules/matchingpursuit.py:391-417 over events grouped by atom in first-seen
    order (include/mpb200.h, mpb200_dictionary_update).  ``running`` (B, N) and ``d_unit`` (K, A) are updated."""
    dev = running.device
    i32 = lambda t: t.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    go, ga, eb, ep = i32(group_offsets), i32(group_atom), i32(ev_batch), i32(ev_pos)
    rows = ev_rows.to(device=dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        check(lib().mpb200_dictionary_update(_ptr(running), running.shape[0], running.shape[-1], _ptr(d_unit),
                                             d_unit.shape[0], d_unit.shape[1], _ptr(go), _ptr(ga), ga.numel(), _ptr(eb),
                                             _ptr(ep), _ptr(rows), eb.numel(), _stream_ptr(dev)),
              "mpb200_dictionary_update")
