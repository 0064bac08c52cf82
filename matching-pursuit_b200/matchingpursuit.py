"""Drop-in counterparts of the reference's ``modules/matchingpursuit.py``.

Same names, positional/keyword signatures and return structures as the
reference (paths below are relative to the reference tree):

* :func:`sparse_code`            -- modules/matchingpursuit.py:229-345
* :func:`sparse_feature_map`     -- modules/matchingpursuit.py:68-125
* :func:`build_scatter_segments` -- modules/matchingpursuit.py:20-58
* :func:`flatten_atom_dict`      -- modules/matchingpursuit.py:61-65
* :func:`dictionary_learning_step` -- modules/matchingpursuit.py:348-419
  (the caller; its only expensive call is ``sparse_code``)

The greedy loop itself runs in the CUDA library (``include/mpb200.h``); this
module only converts between the reference's Python data formats and the
packed ``(atom, position, value)`` arrays the library works on.  Keyword-only
extras (``mode``, ``plan``) select engine behaviour and do not exist in the
reference.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from . import engine
from ._lib import MpbError
from .engine import Plan

# --------------------------------------------------------------------------
# plan cache: plans are expensive to size (workspaces) but cheap to re-aim at a
# new dictionary, which the reference allows to change on every call.
# --------------------------------------------------------------------------
_PLAN_CACHE: "dict[tuple, Plan]" = {}
_PLAN_CACHE_MAX = 8              # e.g. the six bands of a multi-band codec stay resident
_PLAN_CACHE_FRACTION = 0.25      # of the device memory: older plans are closed before a new one is sized


def _evict_until(dev_index: int, max_bytes: int, max_entries: int) -> None:
    """Close this device's oldest idle plans until it holds at most `max_bytes` in at most `max_entries` cached
    plans.  A plan that somebody is using (``with plan:``, see :func:`_run`) is never closed, and other devices'
    plans are neither counted nor touched."""
    def mine():
        return [(key, p) for key, p in _PLAN_CACHE.items() if key[0] == dev_index]     # oldest first (insertion order)

    def over():
        held = mine()
        return len(held) > max_entries or sum(p.device_bytes for _, p in held) > max_bytes
    while over():
        idle = [key for key, p in mine() if p.pins == 0]
        if not idle:
            break
        _PLAN_CACHE.pop(idle[0]).close()


def get_plan(n_atoms: int, atom_size: int, n_samples: int, batch: int, device, mode: str = "auto") -> Plan:
    """A cached plan for this shape (most recently used last).  Callers that keep the plan across further
    ``get_plan`` calls hold it with ``with plan:`` so that making room for another plan cannot close it."""
    dev = engine._require_cuda(device)
    key = (dev.index, n_atoms, atom_size, n_samples, mode)
    plan = _PLAN_CACHE.pop(key, None)
    if plan is not None and (plan.closed or (plan.max_batch < batch and plan.pins == 0)):
        plan.close()
        plan = None
    if plan is not None and plan.max_batch < batch:               # too small but in use: leave it to its user
        plan = None
    if plan is None:
        total = torch.cuda.get_device_properties(dev).total_memory
        _evict_until(dev.index, int(_PLAN_CACHE_FRACTION * total), _PLAN_CACHE_MAX - 1)
        try:
            plan = _make_plan(n_atoms, atom_size, n_samples, batch, mode, dev)
        except MpbError:
            clear_plan_cache(dev.index)                           # make room (idle plans only) and try once more
            plan = _make_plan(n_atoms, atom_size, n_samples, batch, mode, dev)
    _PLAN_CACHE[key] = plan                                       # (re)inserted last: most recently used
    return plan


def _make_plan(n_atoms, atom_size, n_samples, batch, mode, dev) -> Plan:
    """``mode="lcn"``: a plan whose selection is the incremental local-contrast-norm one -- it needs a resident map,
    so a shape AUTO would code by windowed re-correlation is planned as SGRAM instead."""
    if mode != "lcn":
        return Plan(n_atoms, atom_size, n_samples, batch, mode=mode, device=dev)
    plan = Plan(n_atoms, atom_size, n_samples, batch, mode="auto", device=dev)
    if plan.mode not in ("gram", "sgram"):
        plan.close()
        plan = Plan(n_atoms, atom_size, n_samples, batch, mode="sgram", device=dev)
    return plan.set_local_contrast_norm(True)


def clear_plan_cache(dev_index: Optional[int] = None) -> None:
    """Close every idle cached plan (of one device, or of all)."""
    for key in [k for k, p in _PLAN_CACHE.items() if (dev_index is None or k[0] == dev_index) and p.pins == 0]:
        _PLAN_CACHE.pop(key).close()


def _work_device(signal: torch.Tensor, device=None) -> torch.device:
    """CUDA device the pursuit runs on: the signal's own, else ``device``, else the current one."""
    if signal.is_cuda:
        return signal.device
    return engine._require_cuda(device if (device is not None and torch.device(device).type == "cuda") else None)


# --------------------------------------------------------------------------
# events
# --------------------------------------------------------------------------
def _materialising(name):
    """A list method that first turns a lazy EventList into the real list of tuples."""
    base = getattr(list, name)

    def method(self, *args, **kwargs):
        self._fill()
        return base(self, *args, **kwargs)
    method.__name__ = name
    return method


class EventList(list):
    """A list of reference-format event tuples ``(atom:int, batch:int, pos: int64 (1,1), scaled_atom: float32
    (1,1,A))`` (modules/matchingpursuit.py:305-321) that carries the packed arrays it stands for: ``packed = (atom
    int64 (E,), batch int64 (E,), pos int64 (E,), rows float32 (E, A))`` in list order.

    The tuples are built ON FIRST ACCESS: decoders, the multi-band conversions and the dictionary update work on
    ``packed`` and never need them, and materialising B*S Python tuples with two tensor views each costs more than
    the pursuit for small dictionaries.  Until then ``len()`` answers from the arrays; any other list operation --
    iteration, indexing, slicing, comparison, mutation, concatenation, pickling -- materialises first, after which
    this is an ordinary list."""
    packed = None
    _lazy = False

    @classmethod
    def from_packed(cls, atom, batch_idx, pos, rows) -> "EventList":
        out = cls()
        out.packed = (atom, batch_idx, pos, rows)
        out._lazy = atom.numel() > 0
        return out

    def _fill(self) -> None:
        if self._lazy:
            self._lazy = False
            atom, batch_idx, pos, rows = self.packed
            n = atom.numel()
            # one unbind per array instead of two Python-level indexing calls per event
            list.extend(self, zip(atom.tolist(), batch_idx.tolist(), pos.view(n, 1, 1).unbind(0),
                                  rows.view(n, 1, 1, rows.shape[-1]).unbind(0)))

    def __len__(self):
        return self.packed[0].numel() if self._lazy else list.__len__(self)

    def __reduce_ex__(self, protocol):
        self._fill()
        return (list, (list(self),))


for _name in ("__iter__", "__getitem__", "__setitem__", "__delitem__", "__contains__", "__reversed__", "__add__",
              "__iadd__", "__mul__", "__rmul__", "__imul__", "__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__",
              "__repr__", "append", "extend", "insert", "pop", "remove", "index", "count", "sort", "reverse", "copy",
              "clear"):
    setattr(EventList, _name, _materialising(_name))
EventList.__hash__ = None


def _events_from_packed(atom: torch.Tensor, batch_idx: torch.Tensor, pos: torch.Tensor,
                        rows: torch.Tensor) -> EventList:
    """The event list of the given packed arrays (tuples in the order given, built on first access).
    ``atom``/``batch_idx``/``pos`` int64 (E,), ``rows`` (E, A) on the output device."""
    return EventList.from_packed(atom, batch_idx, pos, rows)


def _first_seen_grouping(atom_step_major: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Order of the reference's ``instances`` dict: atoms keyed in first-seen
    order, events of one atom in step-major/batch-minor order
    (modules/matchingpursuit.py:261, 321).  Returns (permutation of the event
    list, atoms in first-seen order)."""
    uniq, first = np.unique(atom_step_major, return_index=True)
    seen_order = uniq[np.argsort(first, kind="stable")]
    rank = np.empty(int(uniq.max()) + 1 if uniq.size else 0, dtype=np.int64)
    rank[seen_order] = np.arange(seen_order.size)
    perm = np.argsort(rank[atom_step_major], kind="stable")
    return perm, seen_order


def flatten_atom_dict(atom_dict):
    """Concatenate the per-atom event lists (modules/matchingpursuit.py:61-65)."""
    flat = []
    for events in atom_dict.values():
        flat.extend(events)
    return flat


# --------------------------------------------------------------------------
# decode
# --------------------------------------------------------------------------
def build_scatter_segments(n_samples, atom_size, device=None):
    """Decoder closure (modules/matchingpursuit.py:20-58).

    ``scatter_segments(x, inst)``: ``x`` is a shape tuple (fresh float32 zeros)
    or a tensor to add onto; every event ``(atom, batch, pos, scaled_atom)`` is
    added at ``pos`` (one channel) or assigned to channel = that batch row's
    running event count (several channels, shape-tuple form only); atoms that
    overhang the right edge are truncated.  The reference allocates on its
    global ``util.device``; here the buffer lives on ``device`` (default: the
    device of the events, else the current CUDA device)."""

    def scatter_segments(x, inst):
        inst = inst if isinstance(inst, list) else list(inst)
        packed = getattr(inst, "packed", None)
        if packed is not None and len(inst) == packed[0].numel():
            _, bidx, pos, rows = packed
        elif len(inst) == 0:
            bidx = pos = torch.zeros(0, dtype=torch.int64)
            rows = torch.zeros(0, atom_size)
        else:
            bidx = torch.tensor([int(ev[1]) for ev in inst], dtype=torch.int64)
            pos = torch.tensor([int(ev[2]) for ev in inst], dtype=torch.int64)
            rows = torch.cat([ev[3].reshape(1, atom_size) for ev in inst], dim=0)
        if isinstance(x, tuple):
            dev = device
            if dev is None:
                dev = rows.device if rows.is_cuda else engine._require_cuda(None)
            out_dev = torch.device(dev)
            work = engine._require_cuda(out_dev if out_dev.type == "cuda" else None)
            out = torch.zeros(*x, device=work, dtype=torch.float32)
            channels = out.shape[1]
            fresh = True
        else:
            out_dev = x.device
            work = engine._require_cuda(out_dev if out_dev.type == "cuda" else None)
            out = x.detach().to(device=work, dtype=torch.float32).clone().contiguous()
            channels = 1
            fresh = False
        n_ch = out.shape[1]
        if len(inst):
            bidx_w = bidx.to(work)
            if fresh and channels > 1:
                # channel = running count of this batch row's events, in list order (:45-52)
                b_host = bidx.cpu().numpy()
                ch = np.zeros(len(b_host), dtype=np.int64)
                seen: dict = {}
                for e, j in enumerate(b_host.tolist()):
                    ch[e] = seen.get(j, 0)
                    seen[j] = ch[e] + 1
                row_index = bidx_w * n_ch + torch.from_numpy(ch).to(work)
                engine.scatter_rows(out, rows.to(work), row_index, pos.to(work))
            elif n_ch == 1:
                engine.scatter_rows(out, rows.to(work), bidx_w, pos.to(work))
            else:
                # a tensor target with several channels: the reference broadcasts the add over them (:48)
                for c in range(n_ch):
                    engine.scatter_rows(out, rows.to(work), bidx_w * n_ch + c, pos.to(work))
        return out if out.device == out_dev else out.to(out_dev)

    return scatter_segments


# --------------------------------------------------------------------------
# the pursuit
# --------------------------------------------------------------------------
def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _pursuit_dense(sig2d: torch.Tensor, plan: Optional[Plan], n_steps: int, compute_feature_map, on_map, on_select,
                   local_contrast_norm: bool, du: Optional[torch.Tensor] = None):
    """Pursuit that materialises the dense (B,K,N) map every step -- the
    schedule of the reference loop, needed whenever a caller-supplied callback
    must see (or supply) that map (modules/matchingpursuit.py:272-273, 283,
    324) or the selection is not the plain maximum (:286-296).  Correlation,
    selection and subtraction are still library kernels."""
    b, n = sig2d.shape
    dev = sig2d.device
    du = plan.unit_dictionary() if du is None else du
    n_atoms = du.shape[0]
    residual = sig2d.clone()
    atom = torch.empty(n_steps, b, device=dev, dtype=torch.int32)
    pos = torch.empty(n_steps, b, device=dev, dtype=torch.int32)
    val = torch.empty(n_steps, b, device=dev, dtype=torch.float32)
    for step in range(n_steps):
        if compute_feature_map is not None:
            fm = compute_feature_map(residual.view(b, 1, n), du)
            fm = fm.to(device=dev, dtype=torch.float32).contiguous().view(b, n_atoms, n)
        else:
            fm = plan.correlate(residual)
        if on_map is not None:
            on_map(step, fm, du)
        best = engine.select_dense(fm, local_contrast_norm=local_contrast_norm)
        v, k, p = engine.unpack_best(best)
        atom[step], pos[step], val[step] = k, p, v
        if on_select is not None:
            on_select(step, fm, k, p, v, du)
        engine.subtract(residual, du, best)
    return atom.t().contiguous(), pos.t().contiguous(), val.t().contiguous(), residual, du


def _run(signal: torch.Tensor, d: torch.Tensor, n_steps: int, device, approx, mode: str, plan: Optional[Plan],
         dense_kwargs: Optional[dict] = None):
    """Common front end: returns packed (atom, pos, val) int32/int32/float32
    (B,S), the residual (B,N), the unit dictionary and the work device."""
    b = signal.shape[0]
    n = signal.shape[-1]
    k, a = d.shape[0], d.shape[-1]
    work = plan.device if plan is not None else _work_device(signal, device)
    sig2d = engine._dev_f32(signal, work, (b, n))
    if isinstance(approx, int) and not isinstance(approx, bool) and approx < n:
        raise NotImplementedError(
            "approx=int<N (top-k spectral bins, modules/conv.py:30-47) is defective in the reference (only atom 0 "
            "is populated) and is not part of the engine; pass approx=None, a slice, or approx>=n_samples")
    if plan is None and a > engine.MAX_PLAN_ATOM:
        return _run_long_atoms(sig2d, d, n_steps, approx, dense_kwargs, work)
    if plan is None:
        # the dense schedule (per-step callbacks, LCN, band-limited maps) only ever calls correlate(): a
        # re-correlation plan holds no resident map or Gram table that would go unused
        dense_path = dense_kwargs is not None or isinstance(approx, slice)
        plan = get_plan(k, a, n, b, work, "recorrelate" if (dense_path and mode == "auto") else mode)
        plan.set_dictionary(d)
    with plan:                       # held: further get_plan calls (below, or inside callbacks) must not evict it
        return _run_held(plan, sig2d, d, n_steps, approx, dense_kwargs, work)


def _run_long_atoms(sig2d: torch.Tensor, d: torch.Tensor, n_steps: int, approx, dense_kwargs, work):
    """Atoms longer than a plan's window transform (the reference runs 4096 ... 16384 samples,
    experiments/archive/e_2023_3_8/experiment.py:352-358, e_2023_12_18/experiment.py:22-24): the dense-map schedule
    with the correlation assembled from the atoms' 2048-sample parts (``mpb200_correlate`` on the dictionary of
    parts + ``mpb200_fold_parts``); selection and subtraction are the usual kernels."""
    b, n = sig2d.shape
    k, a = d.shape[0], d.shape[-1]
    if isinstance(approx, slice):
        raise NotImplementedError("approx=slice with atoms longer than %d samples" % engine.MAX_PLAN_ATOM)
    du = engine.unit_norm(engine._dev_f32(d, work).reshape(k, a))
    parts, n_parts = engine.split_long_atoms(du)
    plan_parts = get_plan(k * n_parts, parts.shape[1], n, b, work, "recorrelate")
    with plan_parts:
        plan_parts.set_dictionary(parts, normalize=False)
        kw = dict(dense_kwargs) if dense_kwargs is not None else dict(
            compute_feature_map=None, on_map=None, on_select=None, local_contrast_norm=False)
        if kw["compute_feature_map"] is None:
            kw["compute_feature_map"] = \
                lambda residual, du_: engine.correlate_long(residual.view(b, n), plan_parts, k, n_parts)
        atom, pos, val, residual, du = _pursuit_dense(sig2d, None, n_steps, du=du, **kw)
    return None, atom, pos, val, residual, du, work


def _run_held(plan: Plan, sig2d: torch.Tensor, d: torch.Tensor, n_steps: int, approx, dense_kwargs, work):
    b, n = sig2d.shape
    k, a = d.shape[0], d.shape[-1]
    if isinstance(approx, slice):
        # band-limited correlation (modules/conv.py:24-29): the mask acts on the whole length-(N+A) spectrum of
        # the residual, so every step changes the whole map -- the reference's recompute-per-step schedule it is
        # (:278-280).  A caller-supplied compute_feature_map takes precedence, as there (:272-273).
        from .conv import band_limited_map
        dense_kwargs = dict(dense_kwargs) if dense_kwargs is not None else dict(
            compute_feature_map=None, on_map=None, on_select=None, local_contrast_norm=False)
        if dense_kwargs["compute_feature_map"] is None:
            plan_long = get_plan(k, a, n + a, b, work, "recorrelate")
            plan_long.set_dictionary(d)
            dense_kwargs["compute_feature_map"] = \
                lambda residual, du: band_limited_map(residual.view(b, n), plan_long, n, approx)
            with plan_long:
                atom, pos, val, residual, du = _pursuit_dense(sig2d, plan, n_steps, **dense_kwargs)
            return plan, atom, pos, val, residual, du, work
    if dense_kwargs is not None:
        atom, pos, val, residual, du = _pursuit_dense(sig2d, plan, n_steps, **dense_kwargs)
    else:
        atom, pos, val, residual = plan.sparse_code(sig2d, n_steps, want_residual=True)
        du = None
    return plan, atom, pos, val, residual, du, work


class _SparseCodeJob:
    """One ``sparse_code`` call in two halves: the constructor enqueues the device work on the CURRENT stream
    (pursuit, scaled atoms, the copy of the atom sequence the host needs for the reference's grouping) and returns
    without waiting; :meth:`result` waits for it and builds the reference's return structures.  Callers with several
    independent problems (the bands of a multi-band codec) start all jobs, each on its own stream, before they
    collect any."""

    def __init__(self, signal, d, n_steps, device, approx, flatten, extract_atom_embedding, visit_key_point,
                 return_residual, local_contrast_norm, return_sparse_feature_map, compute_feature_map, mode, plan,
                 out_device=None):
        batch, channels, time = signal.shape
        if channels != 1:
            raise NotImplementedError("multi-channel signals fail in the reference's scatter_segments "
                                      "(modules/matchingpursuit.py:50); only (B,1,N) is supported")
        self.with_grad = with_grad = _needs_grad(signal, d)
        self.batch, self.n_samples, self.n_steps = batch, time, n_steps
        self.n_atoms, self.atom_size = n_atoms, atom_size = d.shape[0], d.shape[-1]
        # results live on the signal's device, unless the caller staged a host signal on the GPU itself and says
        # where they belong (`out_device`: the multi-band codec uploads a host batch once instead of once per band)
        self.out_dev = signal.device if out_device is None else torch.device(out_device)
        self.flatten, self.return_residual = flatten, return_residual
        self.return_sparse_feature_map = return_sparse_feature_map
        self.want_embeddings = extract_atom_embedding is not None
        self.embeddings = embeddings = []
        n_samples = time

        dense = None
        if (compute_feature_map is not None or extract_atom_embedding is not None or visit_key_point is not None
                or local_contrast_norm):
            def on_map(step, fm, du):
                embeddings.append(extract_atom_embedding(fm, du))            # :282-283

            def on_select(step, fm, k, p, v, du):
                scaled = du[k.long()] * v[:, None] if with_grad else engine.gather_atoms(du, k, v)
                k_host = k.tolist()
                for j in range(batch):
                    visit_key_point(fm[j].view(n_atoms, n_samples), k_host[j], p[j].view(1).to(torch.int64),
                                    scaled[j].view(atom_size))

            dense = dict(compute_feature_map=compute_feature_map,
                         on_map=on_map if extract_atom_embedding is not None else None,
                         on_select=on_select if visit_key_point is not None else None,
                         local_contrast_norm=bool(local_contrast_norm))

        # plain LCN selection (no per-step callback that must see the dense map): the engine's incremental form
        if (dense is not None and local_contrast_norm and compute_feature_map is None and extract_atom_embedding is None
                and visit_key_point is None and approx is None and not with_grad and plan is None
                and atom_size <= engine.MAX_PLAN_ATOM and mode == "auto"):
            try:
                get_plan(n_atoms, atom_size, n_samples, batch, _work_device(signal, device), "lcn")
                dense, mode = None, "lcn"
            except MpbError:
                pass                                                        # no resident map possible: dense schedule
        if with_grad:
            # Gradients are wanted: the greedy selection itself has none (torch.max passes gradient to the selected
            # entry only), so the engine finds the events and PyTorch re-evaluates values, scaled atoms and residual
            # on those fixed indices with the graph attached -- or, when a callback / LCN / a band-limited map needs
            # the dense map under autograd, the whole loop runs as PyTorch ops ("dictionary learning stays in
            # PyTorch", BASELINE.json).
            from . import autograd as _ag
            work = signal.device
            if dense is None and approx is None:
                atom, pos, val, residual = _ag.sparse_code_differentiable(signal, d, n_steps, mode=mode, plan=plan)
                d2 = d.reshape(n_atoms, atom_size)
                du = d2 / (torch.norm(d2, dim=-1, keepdim=True) + 1e-8)
            else:
                atom, pos, val, residual, du, _ = _ag.dense_pursuit(signal, d, n_steps, approx=approx, **(dense or {}))
        else:
            plan_, atom, pos, val, residual, du, work = _run(signal, d, n_steps, device, approx, mode, plan, dense)
            if du is None and not self.want_embeddings:
                du = plan_.unit_dictionary()
        self.work = work
        self.residual = residual
        if self.want_embeddings:
            return
        # events in the reference's order of creation: step-major, batch-minor
        self.atom_sm = atom.t().reshape(-1)
        self.val_sm = val.t().reshape(-1)
        self.pos_sm = pos.t().reshape(-1)
        self.rows = (du[self.atom_sm.long()] * self.val_sm[:, None]) if with_grad else \
            engine.gather_atoms(du, self.atom_sm, self.val_sm)                                            # :305
        if self.atom_sm.is_cuda:
            self.atom_host = torch.empty(self.atom_sm.shape, dtype=self.atom_sm.dtype, pin_memory=True)
            self.atom_host.copy_(self.atom_sm, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(torch.cuda.current_stream(self.atom_sm.device))
        else:
            self.atom_host, self.done = self.atom_sm, None

    def result(self):
        batch, n_samples, n_steps, out_dev, work = self.batch, self.n_samples, self.n_steps, self.out_dev, self.work
        n_atoms, atom_size = self.n_atoms, self.atom_size
        if getattr(self, "done", None) is not None:
            self.done.synchronize()
        residual = self.residual.view(batch, 1, n_samples).to(out_dev)
        if self.want_embeddings:                                             # :332-333
            return self.embeddings, residual
        scatter_segments = build_scatter_segments(n_samples, atom_size, device=out_dev)
        atom_sm, val_sm, pos_sm, rows = self.atom_sm, self.val_sm, self.pos_sm, self.rows
        batch_sm = torch.arange(batch, device=work, dtype=torch.int64).repeat(n_steps)
        atom_host = self.atom_host.numpy().astype(np.int64)
        perm, seen_order = _first_seen_grouping(atom_host) if atom_host.size else (np.zeros(0, np.int64), atom_host)
        perm_t = torch.from_numpy(perm).to(work)
        g_atom = atom_sm.to(torch.int64)[perm_t].to(out_dev)
        g_batch = batch_sm[perm_t].to(out_dev)
        g_pos = pos_sm.to(torch.int64)[perm_t].to(out_dev)
        g_rows = rows[perm_t].to(out_dev)
        flattened = _events_from_packed(g_atom, g_batch, g_pos, g_rows)

        if not self.flatten:                                                 # :335-336
            instances = defaultdict(list)
            counts = np.bincount(atom_host, minlength=n_atoms) if atom_host.size else np.zeros(n_atoms, np.int64)
            start = 0
            for ai in seen_order.tolist():
                c = int(counts[ai])
                instances[ai] = EventList.from_packed(g_atom[start:start + c], g_batch[start:start + c],
                                                      g_pos[start:start + c], g_rows[start:start + c])
                start += c
            return instances, scatter_segments
        if self.return_residual:                                             # :337-339
            return flattened, scatter_segments, residual
        if self.return_sparse_feature_map:                                   # :340-342, :317-318
            sfm = torch.zeros(batch, n_atoms, n_samples, device=work)
            sfm = sfm.index_put((batch_sm, atom_sm.to(torch.int64), pos_sm.to(torch.int64)), val_sm, accumulate=True)
            return flattened, scatter_segments, sfm.to(out_dev)
        return flattened, scatter_segments                                   # :343-345


def sparse_code(signal, d, n_steps=100, device=None, approx=None, flatten=False, extract_atom_embedding=None,
                visit_key_point=None, return_residual=False, local_contrast_norm=False,
                return_sparse_feature_map=False, compute_feature_map=None, fft_convolution=False, *,
                mode: str = "auto", plan: Optional[Plan] = None):
    """Greedy convolutional matching pursuit (modules/matchingpursuit.py:229-345).

    ``signal`` (B,1,N), ``d`` (K,A) or (K,1,A); neither is modified; the
    dictionary is unit-normed first (:254).  Exactly ``n_steps`` events per
    signal.  Return conventions follow the reference (:332-345).  ``device``
    and ``fft_convolution`` are accepted and ignored, as there.  Tensors in the
    results live on ``signal.device``."""
    return _SparseCodeJob(signal, d, n_steps, device, approx, flatten, extract_atom_embedding, visit_key_point,
                          return_residual, local_contrast_norm, return_sparse_feature_map, compute_feature_map, mode,
                          plan).result()


def sparse_code_start(signal, d, n_steps=100, device=None, approx=None, flatten=False, extract_atom_embedding=None,
                      visit_key_point=None, return_residual=False, local_contrast_norm=False,
                      return_sparse_feature_map=False, compute_feature_map=None, fft_convolution=False, *,
                      mode: str = "auto", plan: Optional[Plan] = None, out_device=None) -> _SparseCodeJob:
    """:func:`sparse_code` without the wait: the device work is enqueued on the current stream and the returned
    job's ``result()`` gives what :func:`sparse_code` returns (on ``out_device`` when given, else on the signal's)."""
    return _SparseCodeJob(signal, d, n_steps, device, approx, flatten, extract_atom_embedding, visit_key_point,
                          return_residual, local_contrast_norm, return_sparse_feature_map, compute_feature_map, mode,
                          plan, out_device=out_device)


def sparse_code_arrays(signal, d, n_steps=100, *, approx=None, mode: str = "auto", plan: Optional[Plan] = None,
                       device=None):
    """Array-level form of :func:`sparse_code` for large batches: returns
    ``(atom int32 (B,S), pos int32 (B,S), val float32 (B,S), residual (B,1,N))``
    on ``signal.device`` without building B*S Python tuples."""
    batch = signal.shape[0]
    n_samples = signal.shape[-1]
    if _needs_grad(signal, d):
        from .autograd import sparse_code_differentiable         # indices from the engine, values re-evaluated with the graph
        atom, pos, val, residual = sparse_code_differentiable(signal, d, n_steps, mode=mode, plan=plan)
        return atom.to(torch.int32), pos.to(torch.int32), val, residual
    out_dev = signal.device
    _, atom, pos, val, residual, _, _ = _run(signal, d, n_steps, device, approx, mode, plan)
    return atom.to(out_dev), pos.to(out_dev), val.to(out_dev), residual.view(batch, 1, n_samples).to(out_dev)


def sparse_feature_map(signal, d, n_steps=100, device=None, approx=None, pooling=None, return_residual=False, *,
                       mode: str = "auto", plan: Optional[Plan] = None):
    """Dense (B,K,N) accumulation of the winners (modules/matchingpursuit.py:68-125):
    each step adds the winning value at its (atom, position) -- the forward
    value of ``soft_dirac(f) * f`` (:100-101).  ``pooling`` is accepted and
    ignored as in the reference; ``device`` places the dense map (:84-85)."""
    b = signal.shape[0]
    sig = signal.reshape(b, 1, -1)
    n = sig.shape[-1]
    k = d.shape[0]
    if _needs_grad(signal, d):
        # soft_dirac's backward is the soft-max over the WHOLE flattened map (modules/sparse.py:29-43), so the
        # gradient of this function needs the dense map of every step: the loop runs as PyTorch ops with the graph
        # attached (callers: sparse_coding_loss :128-146; "dictionary learning stays in PyTorch")
        from . import autograd as _ag
        _, _, _, residual, _, fm = _ag.dense_pursuit(sig, d, n_steps, approx=approx, straight_through=True)
        fm = fm.to(torch.device(device)) if device is not None else fm
        return (fm, residual) if return_residual else fm
    _, atom, pos, val, residual, _, work = _run(sig, d, n_steps, device if device is not None else None, approx,
                                                mode, plan)
    fm = torch.zeros(b, k, n, device=work)
    rows = torch.arange(b, device=work, dtype=torch.int64).repeat_interleave(n_steps)
    fm.index_put_((rows, atom.reshape(-1).to(torch.int64), pos.reshape(-1).to(torch.int64)), val.reshape(-1),
                  accumulate=True)
    fm_dev = torch.device(device) if device is not None else signal.device
    fm = fm.to(fm_dev)
    if return_residual:
        return fm, residual.view(b, 1, n).to(signal.device)
    return fm


def sparse_coding_loss(recon, target, d, n_steps=100, device=None, approx=None, pooling=None):
    """modules/matchingpursuit.py:128-146: binary cross-entropy between the sparse feature maps of a reconstruction
    (with gradient) and of its target (without), both divided by the larger of their maxima.  As in the reference
    ``approx`` is accepted and not forwarded."""
    r_map = sparse_feature_map(recon, d, n_steps, device=device, pooling=pooling)
    with torch.no_grad():
        t_map = sparse_feature_map(target, d, n_steps, device=device, pooling=pooling)
    mx = max(r_map.max().item(), t_map.max().item())
    return torch.nn.functional.binary_cross_entropy(r_map / mx, t_map.to(r_map.device) / mx)


def dictionary_learning_step(signal, d, n_steps: int = 100, device=None, approx=None,
                             local_constrast_norm: bool = False, compute_feature_map=None, fft_convolution=False, *,
                             mode: str = "auto"):
    """One dictionary update (modules/matchingpursuit.py:348-419).  The coding pass is the CUDA pursuit; the atom
    update (:391-417) is one kernel launch (``mpb200_dictionary_update``): for every used atom, in first-seen order,
    its instances are added back to a running copy of the SIGNAL (:367 -- not the coding residual), the atom
    becomes the unit-normed sum of the segments under them (:400-406) and the re-scaled new atom is subtracted
    (:408-415)."""
    batch, channels, n_samples = signal.shape
    work = _work_device(signal, device)
    with torch.no_grad():
        d_new = engine.unit_norm(engine._dev_f32(d, work).reshape(d.shape[0], -1)).clone()
        running = engine._dev_f32(signal, work).clone().view(batch * channels, n_samples)
        flat, _ = sparse_code(signal, d, n_steps=n_steps, device=device, approx=approx, flatten=True,
                              local_contrast_norm=local_constrast_norm,
                              compute_feature_map=compute_feature_map, fft_convolution=fft_convolution, mode=mode)
        if len(flat):
            # the flattened list IS the grouped order: atoms in first-seen order, each atom's events together
            atom, bidx, pos, rows = (t.to(work) for t in flat.packed)
            starts = torch.nonzero(torch.cat([torch.ones(1, dtype=torch.bool, device=work), atom[1:] != atom[:-1]])).view(-1)
            offsets = torch.cat([starts, torch.tensor([atom.numel()], device=work)])
            engine.dictionary_update(running, d_new, offsets, atom[starts], bidx, pos, rows)
        out = engine.unit_norm(d_new).view(d.shape)                           # :417
    return out.to(d.device)
