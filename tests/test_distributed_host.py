"""Host-side logic of the multi-GPU forms, on CPU with ``gloo`` (world size 2).

The product's engine is CUDA-only; here the engine slot of
``AtomShardedPursuit`` is filled with a CPU stand-in built from the oracle
(test infrastructure), so what is exercised is the product's sharding
arithmetic and exchange protocol: atom ranges, the all-gather of 16-byte
records in rank-major order, the global tie-break and the per-rank apply."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import matching_pursuit_b200  # noqa: F401  (registers the package)
from matching_pursuit_b200 import distributed as D
from oracle import mp_oracle as O


class OracleEngine:
    """CPU stand-in with the semantics of mpb200_begin / local_best / reduce_best / apply / residual."""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def set_dictionary(self, d):
        self.du = O.unit_norm(d)

    def begin(self, signal):
        self.res = signal.clone().view(signal.shape[0], 1, -1)

    def local_best(self):
        b, _, n = self.res.shape
        fm = O.correlate_direct(self.res, self.du[self.lo:self.hi])
        v, i = fm.reshape(b, -1).max(-1)
        rec = torch.zeros(b, 4, dtype=torch.int32)
        rec[:, 0] = v.view(torch.int32)
        rec[:, 1] = (i // n).to(torch.int32) + self.lo
        rec[:, 2] = (i % n).to(torch.int32)
        return rec

    def reduce(self, cand, n_ranks, batch):
        cand = cand.view(n_ranks, batch, 4)
        out = cand[0].clone()
        for r in range(1, n_ranks):
            for b in range(batch):
                cv, wv = cand[r, b, 0:1].view(torch.float32).item(), out[b, 0:1].view(torch.float32).item()
                c, w = cand[r, b], out[b]
                if cv > wv or (cv == wv and (c[1] < w[1] or (c[1] == w[1] and c[2] < w[2]))):
                    out[b] = c
        return out

    def apply(self, win):
        a = self.du.shape[1]
        n = self.res.shape[-1]
        for b in range(win.shape[0]):
            v = win[b, 0:1].view(torch.float32).item()
            k, p = int(win[b, 1]), int(win[b, 2])
            keep = min(a, n - p)
            self.res[b, 0, p:p + keep] -= (self.du[k] * v)[:keep]

    def residual(self):
        return self.res[:, 0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, a, n, b, steps, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = O.make_dictionary(k, a, seed=3)
        sig = O.make_planted_signals(d, b, n, 8, seed=4)
        lo, hi = D.atom_range(k, world, rank)
        pursuit = D.AtomShardedPursuit(k, a, n, b, engine=OracleEngine(lo, hi)).set_dictionary(d)
        assert (pursuit.lo, pursuit.hi) == (lo, hi) and pursuit.world == world
        atom, pos, val, res = pursuit.run(sig.view(b, n), steps)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), atom=atom.numpy(), pos=pos.numpy(), val=val.numpy(),
                 res=res.numpy())
        # batch-sharded form: per-rank slices gathered back in order
        lo_b, hi_b = D.shard_batch(5, world, rank)
        local = torch.arange(lo_b, hi_b, dtype=torch.float32).view(-1, 1).repeat(1, 3)
        full = D.gather_results(local, 5)
        assert torch.equal(full[:, 0], torch.arange(5, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


def test_atom_sharded_protocol_world2(tmp_path):
    k, a, n, b, steps = 11, 32, 512, 3, 10
    port = _free_port()
    mp.spawn(_worker, args=(2, port, k, a, n, b, steps, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for key in ("atom", "pos", "val", "res"):
        assert np.array_equal(r0[key], r1[key]), key       # every rank ends with the same sequence
    d = O.make_dictionary(k, a, seed=3)
    sig = O.make_planted_signals(d, b, n, 8, seed=4)
    tr = O.greedy_pursuit(sig, d, steps)
    assert np.array_equal(r0["atom"].T, tr.atom.numpy()) and np.array_equal(r0["pos"].T, tr.pos.numpy())
    np.testing.assert_allclose(r0["val"].T, tr.val.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r0["res"], tr.residual.numpy()[:, 0], rtol=1e-5, atol=1e-6)


def test_shard_arithmetic():
    for k, w in [(16384, 8), (11, 2), (7, 7), (10, 4)]:
        ranges = [D.atom_range(k, w, r) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == k
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
    with pytest.raises(ValueError):
        D.atom_range(3, 4, 0)
    with pytest.raises(ValueError):
        D.atom_range(8, 2, 2)
    assert [D.shard_batch(1024, 8, r) for r in (0, 7)] == [(0, 128), (896, 1024)]
    assert [D.shard_batch(3, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
