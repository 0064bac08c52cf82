"""The tensor-core route of the correlation (tcgen05 3xTF32 Toeplitz GEMM, include/mpb200.h
mpb200_correlate_gemm) against the CPU oracle's direct correlation (modules/matchingpursuit.py:275-277): the dense
map within fp32-level tolerance, and the same argmax as the oracle wherever its top-2 margin exceeds 1e-5 -- the
reason the GEMM is split-precision 3xTF32 and not plain TF32 (BASELINE.json north_star (1))."""
import numpy as np
import pytest
import torch

import matching_pursuit_b200 as mpb
from oracle import mp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    # (K, A, N, B)
    (40, 128, 1000, 2),        # partial atom tile, partial position tile
    (300, 512, 3000, 1),       # two atom tiles
    (64, 2048, 4096, 1),       # headline atom length: 64 tap blocks
    (7, 50, 333, 3),           # ragged everything (scalar dictionary loads, zero-padded taps)
    (256, 32, 128, 1),         # exactly one tile, one tap block
    (1024, 128, 8192, 2),      # a configs[3] band
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "K%d_A%d_N%d_B%d" % s)
def test_gemm_map_matches_oracle(shape):
    k, a, n, b = shape
    d = O.make_dictionary(k, a, seed=k + a)
    sig = O.make_noise_signals(b, n, seed=n)
    want = O.correlate_direct(sig, d)
    got = mpb.engine.correlate_gemm(sig.view(b, n).to(DEV), d.to(DEV)).cpu()
    assert got.shape == want.shape
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale, float((got - want).abs().max()) / scale
    # argmax parity: same winner as the oracle wherever the margin is above threshold
    flat_w, flat_g = want.reshape(b, -1), got.reshape(b, -1)
    top2 = torch.topk(flat_w.double(), 2, dim=-1)[0]
    margin = (top2[:, 0] - top2[:, 1]) / top2[:, 0].abs()
    for j in range(b):
        if margin[j] > 1e-5:
            assert int(flat_g[j].argmax()) == int(flat_w[j].argmax())
    # and the FFT route of the engine agrees with it
    plan = mpb.Plan(k, a, n, b, mode="recorrelate", device=DEV).set_dictionary(d, normalize=False)
    fft = plan.correlate(sig.view(b, n).to(DEV)).cpu()
    assert float((got - fft).abs().max()) <= 2e-5 * scale


def test_plain_tf32_would_not_hold_parity():
    """What the split buys: a single-pass TF32 product (inputs truncated to 19 bits) is three orders of magnitude
    less accurate than the 3xTF32 map -- far above the 1e-5 margin rule."""
    k, a, n = 64, 512, 2048
    d = O.make_dictionary(k, a, seed=1)
    sig = O.make_noise_signals(1, n, seed=2)
    want = O.correlate_direct(sig.double(), d.double())
    got = mpb.engine.correlate_gemm(sig.view(1, n).to(DEV), d.to(DEV)).cpu().double()
    trunc = lambda t: (t.view(torch.int32) & -8192).view(torch.float32)
    tf32 = O.correlate_direct(trunc(sig.clone()).double(), trunc(d.clone()).double())
    err3 = float((got - want).abs().max()) / float(want.abs().max())
    err1 = float((tf32 - want).abs().max()) / float(want.abs().max())
    assert err3 < 1e-5 < err1, (err3, err1)
