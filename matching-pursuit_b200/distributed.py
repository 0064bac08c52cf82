"""Multi-GPU forms of the pursuit: one process per GPU, ``torch.distributed``
for the plumbing (NCCL over NVLink on the GPU box; the host logic is
backend-agnostic and is exercised with ``gloo`` on CPU in
``tests/test_distributed_host.py`` through the injectable engine below).

* **Batch sharding** (BASELINE configs[1]-[3]): signals are independent
  problems (the reference never mixes batch rows, modules/matchingpursuit.py:299,
  :311-328), so every rank codes its slice with a replicated dictionary and
  there is NO data-path collective; :func:`shard_batch` gives the slice and
  :func:`gather_results` concatenates the packed results at the end.
* **Atom sharding** (configs[4], one long signal, large dictionary): residual
  and dictionary are replicated, rank ``g`` owns atoms ``atom_range(K, G, g)``;
  per iteration each rank finds its local winner, the 16-byte ``mpb200_best``
  records are all-gathered, every rank reduces them with the reference
  tie-break (max value, then lowest atom, then lowest position) and applies
  the global winner locally.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def atom_range(n_atoms: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced atom shard of ``rank`` (first ``n_atoms % world`` ranks hold one more)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    if world > n_atoms:
        raise ValueError(f"cannot shard {n_atoms} atoms over {world} ranks: every rank must own at least one atom")
    base, extra = divmod(n_atoms, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice ``[lo, hi)`` of the batch for ``rank`` (may be empty)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """Concatenate per-rank result rows (shard_batch order) on every rank.
    The only collective of the batch-sharded form, after the pursuit."""
    world = dist.get_world_size(group)
    sizes = [shard_batch(batch, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


class PlanEngine:
    """The CUDA engine behind :class:`AtomShardedPursuit` (default)."""

    def __init__(self, n_atoms, atom_size, n_samples, batch, lo, hi, device=None, mode="auto"):
        from .engine import Plan, reduce_best
        self.plan = Plan(n_atoms, atom_size, n_samples, batch, mode=mode, atom_range=(lo, hi), device=device)
        self._reduce = reduce_best
        self.device = self.plan.device

    def set_dictionary(self, d):
        self.plan.set_dictionary(d)

    def begin(self, signal):
        self.plan.begin(signal)

    def local_best(self):
        return self.plan.local_best()

    def reduce(self, cand, n_ranks, batch):
        return self._reduce(cand, n_ranks, batch)

    def apply(self, winner):
        self.plan.apply(winner)

    def residual(self):
        return self.plan.residual()


class AtomShardedPursuit:
    """Greedy pursuit of a replicated batch of signals with the dictionary's
    atoms sharded over the ranks of ``group``.  Every rank ends with the same
    ``(atom, pos, val)`` (B, S) sequence and the same residual.

    ``engine`` must provide ``set_dictionary, begin, local_best, reduce, apply,
    residual`` with the semantics of ``include/mpb200.h`` (begin / local_best /
    reduce_best / apply / residual); the default is the CUDA :class:`PlanEngine`.
    """

    def __init__(self, n_atoms: int, atom_size: int, n_samples: int, batch: int, group=None, engine=None,
                 device=None, mode: str = "auto", exchange: str = "nccl"):
        """``exchange``: ``"nccl"`` -- per-step ``all_gather_into_tensor`` of the 16-byte records between
        ``local_best`` and ``apply`` (host-driven loop); ``"p2p"`` -- the exchange is fused into the kernel
        that applies the winner (peer-memory stores over NVLink into every rank's mailbox, include/mpb200.h
        ``mpb200_exchange_*``): the whole pursuit is one library call with no collective in the loop."""
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.exchange = exchange
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = atom_range(n_atoms, self.world, self.rank)
        self.batch = batch
        self.engine = engine if engine is not None else PlanEngine(n_atoms, atom_size, n_samples, batch,
                                                                   self.lo, self.hi, device=device, mode=mode)
        self.exchange_ms: List[float] = []
        if exchange == "p2p" and self.world > 1:
            plan = self.engine.plan
            mine = plan.exchange_create(self.world, self.rank)
            local = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(plan.device)
            every = torch.empty(self.world * 64, dtype=torch.uint8, device=plan.device)
            dist.all_gather_into_tensor(every, local, group=self.group)      # set-up only: trade the IPC handles
            plan.exchange_connect(bytes(every.cpu().numpy().tobytes()))
            dist.barrier(group=self.group)

    def set_dictionary(self, d: torch.Tensor) -> "AtomShardedPursuit":
        self.engine.set_dictionary(d)
        return self

    def close(self) -> None:
        """Collective teardown: every rank unmaps its peers' mailboxes BEFORE any rank frees its own (freeing
        memory that another process still has mapped through CUDA IPC is undefined)."""
        plan = getattr(self.engine, "plan", None)
        if plan is None:
            return
        if self.exchange == "p2p" and self.world > 1:
            plan.exchange_disconnect()
            dist.barrier(group=self.group)
        plan.close()

    def _exchange(self, local: torch.Tensor) -> torch.Tensor:
        """All-gather of the (B, 4) int32 records -> rank-major (world*B, 4)."""
        if self.world == 1:
            return local
        out = torch.empty((self.world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out

    def run(self, signal: torch.Tensor, n_steps: int, time_exchange: bool = False):
        """Returns ``(atom int32 (B,S), pos int32 (B,S), val float32 (B,S), residual (B,N))``."""
        eng = self.engine
        if self.exchange == "p2p" and self.world > 1:
            atom, pos, val, res = eng.plan.sparse_code(signal, n_steps, want_residual=True)
            if eng.plan.exchange_timed_out():          # synchronises; a peer did not deliver a record within 20 s
                from ._lib import MpbError
                raise MpbError("atom-sharded pursuit: a candidate record from another rank did not arrive in time; "
                               "the results of this call are not valid on this rank")
            return atom, pos, val, res
        eng.begin(signal)
        b = signal.shape[0]
        wins = []
        self.exchange_ms = []
        timed = time_exchange and signal.is_cuda
        for _ in range(n_steps):
            local = eng.local_best()
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            cand = self._exchange(local)
            if timed:
                e1.record()
                self._pending = getattr(self, "_pending", []) + [(e0, e1)]
            win = eng.reduce(cand, self.world, b)
            eng.apply(win)
            wins.append(win)
        if timed:
            torch.cuda.synchronize()
            self.exchange_ms = [a.elapsed_time(z) for a, z in self._pending]
            self._pending = []
        rec = torch.stack(wins, dim=1) if wins else torch.zeros(b, 0, 4, dtype=torch.int32, device=signal.device)
        val = rec[..., 0].contiguous().view(torch.float32)
        return rec[..., 1].contiguous(), rec[..., 2].contiguous(), val, eng.residual()
