"""One configs[3] band (1024 x 128 on 65536 samples, 16 signals) for an ncu launch list (development aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from oracle import mp_oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
dev = torch.device("cuda", 0)
k, a, b, s = 1024, 128, 16, 24
d = O.make_dictionary(k, a, seed=10).to(dev)
x = O.make_planted_signals(d.cpu(), b, n, 4 * s, seed=1).to(dev)
plan = mpb.Plan(k, a, n, b, mode=mode, device=dev).set_dictionary(d)
print(plan.mode, plan.block, plan.n_blocks)
for _ in range(2):
    plan.sparse_code(x.view(b, n), s)
torch.cuda.synchronize()
