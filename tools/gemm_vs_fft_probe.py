"""One dense correlation per route (FFT: k_window_fft + k_corr; tensor cores: k_corr_gemm) on three shapes, for an
`ncu --set full` capture (tensor-pipe and DRAM counters of both routes; profiles/r2_gemm_vs_fft.md).  Development aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import matching_pursuit_b200 as mpb  # noqa: E402
from oracle import mp_oracle as O  # noqa: E402  (inputs only)

dev = torch.device("cuda", 0)
for k, a, n, b in [(1024, 128, 65536, 4), (512, 512, 32768, 4), (4096, 2048, 32768, 1)]:
    d = O.make_dictionary(k, a, seed=0).to(dev)
    x = torch.randn(b, n, device=dev)
    plan = mpb.Plan(k, a, n, b, mode="recorrelate", device=dev).set_dictionary(d, normalize=False)
    torch.cuda.synchronize()
    f = plan.correlate(x)
    g = mpb.engine.correlate_gemm(x, d)
    torch.cuda.synchronize()
    print(k, a, n, b, float((f - g).abs().max() / f.abs().max()))
    plan.close()
