"""FFT route against the tensor-core (tcgen05 3xTF32 Toeplitz GEMM) route of the dense correlation, and the TF32
tensor-pipe peak of this GPU measured with the same instruction stream (profiles/r2_gemm_vs_fft.md).  Development /
measurement aid; run on a B200:  python tools/bench_gemm.py [--peak-only]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from oracle import mp_oracle as O  # noqa: E402  (inputs only)

ap = argparse.ArgumentParser()
ap.add_argument("--peak-only", action="store_true")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


# ---- TF32 tensor-pipe peak: 148 CTAs (one per SM), each issuing `spin` blocks of 12 MMAs (128 x 256 x 8, kind::tf32)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
spin = 4000
sig = torch.randn(1, 128 * sms, device=dev)
d = torch.randn(256, 32, device=dev)
ms = timed(lambda: mpb.engine.correlate_gemm(sig, d, spin_blocks=spin), args.reps)
flop = sms * (1 + spin) * 12 * 2 * 128 * 256 * 8
peak = flop / (ms / 1e3) / 1e12
print(json.dumps({"what": "tf32_tensor_pipe_peak", "ms": ms, "tflops": peak, "ctas": sms, "mma_blocks_per_cta": 1 + spin,
                  "note": "dense kind::tf32 tcgen05.mma issue rate from shared-memory operands, no staging; the fp32-"
                          "equivalent ceiling of the 3xTF32 route is a third of this"}), flush=True)
if args.peak_only:
    sys.exit(0)

SHAPES = [
    # (label, K, A, N, B)
    ("configs[3] band 2048", 1024, 128, 2048, 16),
    ("configs[3] band 65536", 1024, 128, 65536, 16),
    ("configs[0]", 512, 512, 2 ** 15, 1),
    ("configs[1]", 512, 1024, 2 ** 15, 16),
    ("configs[2] (headline)", 4096, 2048, 2 ** 15, 4),
    ("refresh window A=128 (255 positions)", 1024, 128, 255, 64),
    ("refresh window A=512 (1023 positions)", 512, 512, 1023, 64),
    ("refresh window A=2048 (4095 positions)", 4096, 2048, 4095, 16),
]
for label, k, a, n, b in SHAPES:
    dct = O.make_dictionary(k, a, seed=0).to(dev)
    x = torch.randn(b, n, device=dev)
    plan = mpb.Plan(k, a, n, b, mode="recorrelate", device=dev).set_dictionary(dct, normalize=False)
    ms_fft = timed(lambda: plan.correlate(x), args.reps)
    ms_gemm = timed(lambda: mpb.engine.correlate_gemm(x, dct), args.reps)
    fm_f, fm_g = plan.correlate(x), mpb.engine.correlate_gemm(x, dct)
    err = float((fm_f - fm_g).abs().max() / fm_f.abs().max())
    flops = 2.0 * b * k * n * a
    print(json.dumps({"shape": label, "K": k, "A": a, "N": n, "B": b, "fft_ms": ms_fft, "gemm_ms": ms_gemm,
                      "gemm_over_fft": ms_gemm / ms_fft, "fp32_equiv_tflops_gemm": flops / (ms_gemm / 1e3) / 1e12,
                      "fp32_equiv_tflops_fft": flops / (ms_fft / 1e3) / 1e12, "map_bytes": 4.0 * b * k * n,
                      "max_rel_diff": err}), flush=True)
    plan.close()
    del fm_f, fm_g
