"""Inputs of the headline-shape golden files (tests/golden/hl_*.npz): regenerated from the seeds the files store
(torch CPU generators through oracle.mp_oracle, exactly what oracle/make_golden_headline.py did in the build
container) and verified against the stored checksums, so a drifting RNG fails loudly instead of silently
comparing different problems."""
import os

import numpy as np
import torch

from oracle import mp_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HEADLINE_SINGLE = ["hl_c3_k4096_a2048_n32768_b2_s64", "hl_c2_k512_a1024_n32768_b4_s64",
                   "hl_c5shard_k2048_a2048_n262144_b1_s16"]
HEADLINE_LONG = ["hl_long_k64_a4096_n32768_b2_s24", "hl_long_k32_a8192_n32768_b2_s24",
                 "hl_long_k16_a16384_n65536_b1_s16"]
HEADLINE_MULTIBAND = "hl_c4_6bands_k1024_a128_n65536_b2_s16"


def checksum(x: torch.Tensor) -> np.ndarray:
    v = x.detach().reshape(-1).double()
    w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 8191.0 + 1.0
    return np.array([float(v.sum()), float((v * w).sum())], dtype=np.float64)


def same_checksum(x: torch.Tensor, want: np.ndarray) -> bool:
    got = checksum(x)
    scale = float(x.detach().abs().double().sum()) * 8192.0 + 1e-30
    return bool(np.all(np.abs(got - want) <= 1e-9 * scale))


def load_single(name):
    """-> (golden arrays, dictionary (K, A), signals (B, 1, N))."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = O.make_dictionary(int(g["k"]), int(g["a"]), seed=int(g["d_seed"]))
    sig = O.make_planted_signals(d, int(g["b"]), int(g["n"]), int(g["planted"]), seed=int(g["s_seed"]))
    assert same_checksum(d, g["d_checksum"]), "dictionary regenerated from its seed differs from the recorded one"
    assert same_checksum(sig, g["signal_checksum"]), "signals regenerated from their seed differ from the recorded ones"
    return g, d, sig


def load_multiband():
    """-> (golden arrays, x (B, 1, N), {size: dictionary}, {size: band signals (B, 1, size)}); the bands are the
    oracle's restatement of the reference's band split, verified against the reference's recorded checksums."""
    g = np.load(os.path.join(GOLDEN, HEADLINE_MULTIBAND + ".npz"))
    k, a, n, b, steps = (int(g[key]) for key in ("k", "a", "n", "b", "steps"))
    sizes = [int(s) for s in g["sizes"]]
    x = O.make_planted_signals(O.make_dictionary(k, a, seed=0), b, n, 4 * steps, seed=1)
    assert same_checksum(x, g["x_checksum"])
    bands = O.band_split(x, sizes[0])
    dicts = {}
    for i, size in enumerate(sizes):
        dicts[size] = O.make_dictionary(k, a, seed=10 + i)
        assert same_checksum(bands[size], g[f"band_checksum_{size}"]), size
    return g, x, dicts, bands
