"""Where the time of MultibandDictionaryLearning.encode goes at configs[3] (development aid)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from matching_pursuit_b200 import decompose as mdec  # noqa: E402
from oracle import mp_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
n, k, a, s, b = 2 ** 16, 1024, 128, 64, 16
sizes = [2048 * 2 ** i for i in range(6)]
specs = [mpb.BandSpec(sz, k, a, device=dev, signal_samples=n, is_lowest_band=(i == 0)) for i, sz in enumerate(sizes)]
for i, spec in enumerate(specs):
    spec.d = O.make_dictionary(k, a, seed=10 + i).to(dev)
model = mpb.MultibandDictionaryLearning(specs, n_samples=n)
x_host = O.make_planted_signals(O.make_dictionary(k, a, seed=0), b, n, 4 * s, seed=1).pin_memory()
x_dev = x_host.to(dev)


def t(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, 1e3 * (time.perf_counter() - t0))
    return round(best, 2)


print("decompose dev", t(lambda: mdec.fft_frequency_decompose(x_dev, sizes[0])))
print("decompose host", t(lambda: mdec.fft_frequency_decompose(x_host, sizes[0])))
split = mdec.fft_frequency_decompose(x_dev, sizes[0])
print("arrays sequential dev", t(lambda: [mpb.sparse_code_arrays(split[sz], sp.d, s) for sz, sp in zip(sizes, specs)]))
for sz, sp in zip(sizes, specs):
    print("  band", sz, "arrays", t(lambda: mpb.sparse_code_arrays(split[sz], sp.d, s)),
          "encode(dev)", t(lambda: sp.encode(split[sz], s)),
          "mode", mpb.get_plan(k, a, sz, b, dev, "auto").mode)
for sz, sp in zip(sizes, specs):
    row = {}
    for mode in ("sgram", "gram"):
        plan = mpb.Plan(k, a, sz, b, mode=mode, device=dev)
        t0 = time.perf_counter()
        plan.set_dictionary(sp.d)
        torch.cuda.synchronize()
        row[mode + "_tables_ms"] = round(1e3 * (time.perf_counter() - t0), 2)
        row[mode] = t(lambda: plan.sparse_code(split[sz].view(b, sz), s))
        plan.close()
    print("  band", sz, row)
print("encode dev input", t(lambda: model.encode(x_dev, s)))
print("encode host input", t(lambda: model.encode(x_host, s)))
enc = model.encode(x_dev, s)
print("flattened_event_tuples", t(lambda: model.flattened_event_tuples(enc)))
flat = model.flattened_event_tuples(enc)
print("hierarchical_event_tuples", t(lambda: model.hierarchical_event_tuples(flat, enc)))
hier = model.hierarchical_event_tuples(flat, enc)
print("decode", t(lambda: model.decode(hier)))
