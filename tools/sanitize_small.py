"""Development aid (GPU box): small pursuits through the paths with hand-rolled synchronisation, for
`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py`: the one-launch cooperative
loop (grid barrier, shared-memory buffer reuse), SGRAM's k_delta (bulk copies, mbarriers) and GRAM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from oracle import mp_oracle as O  # noqa: E402

dev = "cuda:0"
for (k, a, n, b, s, mode) in [(24, 64, 2048, 2, 6, "recorrelate"), (7, 128, 4096, 1, 5, "recorrelate"),
                              (16, 128, 4096, 2, 5, "sgram"), (16, 128, 4096, 2, 5, "gram")]:
    d = O.make_dictionary(k, a, seed=1)
    sig = O.make_planted_signals(d, b, n, 4, seed=2)
    plan = mpb.Plan(k, a, n, b, mode=mode, device=dev).set_dictionary(d)
    out = plan.sparse_code(sig.to(dev), s)
    torch.cuda.synchronize()
    tr = O.greedy_pursuit(sig, d, s)
    same = (out[0].cpu().numpy() == tr.atom.numpy().T).mean()          # trace is step-major (S, B)
    print(mode, (k, a, n, b, s), "atoms equal to the oracle:", float(same))
    plan.close()
