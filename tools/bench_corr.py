"""Quick device timing of the pursuit kernels on a slice of BASELINE configs[2]
(4096 x 2048 dictionary, 2^15-sample signals): per-iteration milliseconds of the
window re-correlation kernel and the implied time per (window, atom-pair) transform.
Development aid; the judged numbers come from bench.py."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from bench import make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iterations", type=int, default=24)
ap.add_argument("--atoms", type=int, default=4096)
ap.add_argument("--atom-size", type=int, default=2048)
ap.add_argument("--samples", type=int, default=2 ** 15)
ap.add_argument("--mode", default="recorrelate")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
d, sig = make_inputs(torch, mpb, dev, args.batch, args.samples, args.atoms, args.atom_size, 64, 1)
plan = mpb.Plan(args.atoms, args.atom_size, args.samples, args.batch, mode=args.mode, device=dev).set_dictionary(d)
plan.sparse_code(sig, args.iterations)
torch.cuda.synchronize()
plan.timing(True)
for _ in range(args.reps):
    out = plan.sparse_code(sig, args.iterations)
t = plan.timing_read()
pairs = (args.atoms + 1) // 2
corr_ms = t["recorrelate"][0] / max(t["recorrelate"][1], 1)
upd_ms = t["gram_update"][0] / max(t["gram_update"][1], 1)
rb = plan.resident_batch
print(json.dumps({
    "mode": plan.mode, "batch": args.batch, "resident_batch": rb, "fft_size": plan.fft_size, "fft_size2": plan.fft_size2,
    "us_per_atom_step": 1e3 * (corr_ms + upd_ms + t["apply"][0] / max(t["apply"][1], 1)) / min(rb, args.batch),
    "recorrelate_ms_per_iteration": corr_ms,
    "ns_per_transform": 1e6 * corr_ms / (args.batch * pairs),
    "apply_ms": t["apply"][0] / max(t["apply"][1], 1),
    "gram_ms": t["gram_update"][0] / max(t["gram_update"][1], 1),
    "first_pass_ms": t["first_pass"][0] / max(t["first_pass"][1], 1),
    "checksum": [int(out[0].sum()), int(out[1].sum())],
}))
