"""Two-process, two-GPU checks of the atom-sharded pursuit (skipped on a single-GPU box): the exchange fused
into the pursuit over CUDA-IPC mailboxes and the NCCL form must both give every rank the oracle's sequence."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, exchange: str, out_dir: str, mode: str = "recorrelate"):
    import torch.distributed as dist
    import matching_pursuit_b200 as mpb  # noqa: F401
    from matching_pursuit_b200.distributed import AtomShardedPursuit
    from oracle import mp_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    k, a, n, b, s = 37, 128, 4096, 3, 24
    d = O.make_dictionary(k, a, seed=5)
    sig = O.make_planted_signals(d, b, n, 12, seed=6)
    pursuit = AtomShardedPursuit(k, a, n, b, device=dev, mode=mode, exchange=exchange).set_dictionary(d)
    atom, pos, val, res = pursuit.run(sig.to(dev), s)
    torch.cuda.synchronize()
    timed_out = pursuit.engine.plan.exchange_timed_out() if exchange == "p2p" else False
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), atom=atom.cpu().numpy(), pos=pos.cpu().numpy(),
             val=val.cpu().numpy(), res=res.cpu().numpy(), timed_out=timed_out)
    pursuit.close()                       # collective: peers unmap the mailboxes before anybody frees one
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["recorrelate", "sgram"])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_ranks_atom_sharded(exchange, mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import mp_oracle as O
    from parity import compare_with_oracle_trace
    mp.spawn(_worker, args=(2, _free_port(), exchange, str(tmp_path), mode), nprocs=2, join=True)
    k, a, n, b, s = 37, 128, 4096, 3, 24
    d = O.make_dictionary(k, a, seed=5)
    sig = O.make_planted_signals(d, b, n, 12, seed=6)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(2))
    assert not bool(r0["timed_out"]) and not bool(r1["timed_out"])
    for key in ("atom", "pos", "val", "res"):
        assert np.array_equal(r0[key], r1[key]), key
    assert compare_with_oracle_trace(tr, r0["atom"], r0["pos"], r0["val"], r0["res"]) >= 0.9 * s * b
