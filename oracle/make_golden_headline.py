"""Pin the HEADLINE shapes to the unmodified reference (container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_headline [case ...]

TEST INFRASTRUCTURE ONLY.  ``oracle/make_golden.py`` stores whole input arrays,
which is fine up to 64 atoms; the BASELINE.json shapes have 32 MB dictionaries,
so these files store the inputs BY SEED (``oracle.mp_oracle.make_dictionary`` /
``make_planted_signals`` -- torch CPU generators, identical bits for the same
torch build) together with float64 checksums of the arrays the seeds produced
here, and what ``modules/matchingpursuit.py::sparse_code`` (:229-345, the
default conv1d correlation) returned for them: the true iteration order, the
SIGNED value and the top-2 margin of every step through the reference's own
``visit_key_point`` hook (:323-324), and the final residual.

A test regenerates the inputs from the seeds, checks the checksums (so RNG
drift fails loudly instead of comparing different problems) and compares the
CUDA path step by step, re-synchronising from the reference's own events after
an ambiguous step (tests/parity.py::run_with_resync).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from oracle import mp_oracle as O
from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def checksum(x: torch.Tensor) -> np.ndarray:
    """Two position-sensitive float64 sums: cheap, and any RNG difference moves them."""
    v = x.detach().reshape(-1).double()
    w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 8191.0 + 1.0
    return np.array([float(v.sum()), float((v * w).sum())], dtype=np.float64)


def record(ref, signal, d, steps, **kw):
    b, _, n = signal.shape
    rows = []
    t0 = time.time()

    def visit(fm, ai, p, a):
        top2 = torch.topk(fm.reshape(-1), 2)[0].double()
        margin = float((top2[0] - top2[1]) / top2[0].abs().clamp_min(1e-300))
        rows.append((ai, int(p), float(fm[ai, int(p)]), margin))
        if len(rows) % b == 0:
            print(f"    step {len(rows) // b}/{steps}  {time.time() - t0:.0f} s", flush=True)

    with torch.no_grad():
        flat, scatter, residual = ref.matchingpursuit.sparse_code(
            signal, d, n_steps=steps, flatten=True, return_residual=True, visit_key_point=visit, **kw)
    seq = np.array(rows, dtype=np.float64).reshape(steps, b, 4)
    return dict(atom=seq[..., 0].astype(np.int64), pos=seq[..., 1].astype(np.int64),
                val=seq[..., 2].astype(np.float32), margin=seq[..., 3],
                residual=residual.detach().numpy().reshape(b, n))


def case_single(ref, name, k, a, n, b, steps, planted, d_seed, s_seed):
    d = O.make_dictionary(k, a, seed=d_seed)
    sig = O.make_planted_signals(d, b, n, planted, seed=s_seed)
    out = record(ref, sig, d, steps)
    out.update(k=np.int64(k), a=np.int64(a), n=np.int64(n), b=np.int64(b), steps=np.int64(steps),
               planted=np.int64(planted), d_seed=np.int64(d_seed), s_seed=np.int64(s_seed),
               d_checksum=checksum(d), signal_checksum=checksum(sig))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_multiband(ref, name, sizes, k, a, n, b, steps):
    """configs[3]: the reference's own band split (modules/decompose.py:5-33) of planted signals, then per band the
    call BandSpec.encode makes (modules/multibanddict.py:238-247) with the trace hook attached."""
    x = O.make_planted_signals(O.make_dictionary(k, a, seed=0), b, n, 4 * steps, seed=1)
    split = ref.decompose.fft_frequency_decompose(x, sizes[0])
    out = dict(k=np.int64(k), a=np.int64(a), n=np.int64(n), b=np.int64(b), steps=np.int64(steps),
               sizes=np.array(sizes), x_checksum=checksum(x))
    for i, size in enumerate(sizes):
        d = O.make_dictionary(k, a, seed=10 + i)
        print(f"  band {size}", flush=True)
        r = record(ref, split[size], d, steps)
        out[f"band_checksum_{size}"] = checksum(split[size])
        for key, v in r.items():
            out[f"{key}_{size}"] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


CASES = {
    # BASELINE configs[2] shape: 4096 x 2048 dictionary on 2^15 samples
    "hl_c3_k4096_a2048_n32768_b2_s64": lambda ref, nm: case_single(ref, nm, 4096, 2048, 2 ** 15, 2, 64, 48, 0, 1),
    # BASELINE configs[1] shape: 512 x 1024 dictionary on 2^15 samples
    "hl_c2_k512_a1024_n32768_b4_s64": lambda ref, nm: case_single(ref, nm, 512, 1024, 2 ** 15, 4, 64, 48, 0, 1),
    # BASELINE configs[3]: six bands of 1024 x 128 on 2^16 samples
    "hl_c4_6bands_k1024_a128_n65536_b2_s16": lambda ref, nm: case_multiband(
        ref, nm, [2048 * 2 ** i for i in range(6)], 1024, 128, 2 ** 16, 2, 16),
    # one rank's shard of BASELINE configs[4]: 2048 atoms x 2048 samples on 2^18 samples
    "hl_c5shard_k2048_a2048_n262144_b1_s16": lambda ref, nm: case_single(ref, nm, 2048, 2048, 2 ** 18, 1, 16, 12, 5, 3),
    # long atoms (experiments/archive/e_2023_3_8/experiment.py:352-358, e_2023_12_18/experiment.py:22-24)
    "hl_long_k64_a4096_n32768_b2_s24": lambda ref, nm: case_single(ref, nm, 64, 4096, 2 ** 15, 2, 24, 16, 6, 7),
    "hl_long_k32_a8192_n32768_b2_s24": lambda ref, nm: case_single(ref, nm, 32, 8192, 2 ** 15, 2, 24, 16, 8, 9),
    "hl_long_k16_a16384_n65536_b1_s16": lambda ref, nm: case_single(ref, nm, 16, 16384, 2 ** 16, 1, 16, 10, 10, 11),
}


def main(argv):
    ref = ref_loader.load()
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    for name in (argv or list(CASES)):
        t0 = time.time()
        print(name, flush=True)
        CASES[name](ref, name)
        print(f"  wrote {name}.npz in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
