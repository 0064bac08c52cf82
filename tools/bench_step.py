"""Device timing of the step-wise (atom-sharded) interface on one GPU: begin() and the per-iteration
local_best -> reduce -> apply chain, for a plan that owns `--owned` of `--atoms` atoms (emulates one rank of
BASELINE configs[4]).  Development aid."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from bench import make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--atoms", type=int, default=16384)
ap.add_argument("--owned", type=int, default=2048)
ap.add_argument("--atom-size", type=int, default=2048)
ap.add_argument("--samples", type=int, default=2 ** 20)
ap.add_argument("--iterations", type=int, default=256)
ap.add_argument("--mode", default="auto")
args = ap.parse_args()
dev = torch.device("cuda", 0)
d, sig = make_inputs(torch, mpb, dev, 1, args.samples, args.atoms, args.atom_size, 256, 1)
plan = mpb.Plan(args.atoms, args.atom_size, args.samples, 1, mode=args.mode, atom_range=(0, args.owned),
                device=dev).set_dictionary(d)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for rep in range(2):
    ev[0].record()
    plan.begin(sig)
    ev[1].record()
    for _ in range(args.iterations):
        plan.apply(mpb.reduce_best(plan.local_best(), 1, 1))
    ev[2].record()
    torch.cuda.synchronize()
print(json.dumps({"mode": plan.mode, "owned": args.owned, "begin_ms": ev[0].elapsed_time(ev[1]),
                  "us_per_iteration": 1e3 * ev[1].elapsed_time(ev[2]) / args.iterations}))
