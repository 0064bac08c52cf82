"""GPU parity at the HEADLINE shapes, against the live reference.

tests/golden/hl_*.npz hold what the unmodified reference
(modules/matchingpursuit.py:229-345, default conv1d correlation) returned for
BASELINE.json's shapes -- configs[2] (4096 x 2048 on 2^15), configs[1]
(512 x 1024 on 2^15), configs[3] (six bands of 1024 x 128 on 2^16) and one
rank's shard of configs[4] (2048 x 2048 on 2^18) -- and for the long atoms of
the reference's own experiments (4096, 8192 and 16384 samples).  Inputs are
stored by seed with checksums (oracle/make_golden_headline.py); every schedule
that fits is checked step by step, restarting from the reference's residual
after an ambiguous step (parity.compare_with_resync), and at least 90 % of the
(signal, step) pairs must have been binding."""
import os

import numpy as np
import pytest
import torch

import matching_pursuit_b200 as mpb
from headline_inputs import GOLDEN, HEADLINE_LONG, HEADLINE_SINGLE, load_multiband, load_single
from oracle import mp_oracle as O
from parity import compare_with_resync

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def runner(d, n, max_batch, mode):
    plan = mpb.Plan(d.shape[0], d.shape[-1], n, max_batch, mode=mode, device=DEV).set_dictionary(d)
    assert plan.mode == mode

    def run(signals, steps):
        sig = torch.from_numpy(np.ascontiguousarray(signals, dtype=np.float32)).to(DEV)
        out = plan.sparse_code(sig, steps)
        torch.cuda.synchronize()
        return tuple(t.cpu().numpy() for t in out)

    return run, plan


def check(sig, d, g, mode, suffix=""):
    b, n = sig.shape[0], sig.shape[-1]
    run, plan = runner(d, n, b, mode)
    rep = compare_with_resync(run, sig.numpy().reshape(b, n), O.unit_norm(d).numpy(), g["atom" + suffix],
                              g["pos" + suffix], g["val" + suffix], g["margin" + suffix], g["residual" + suffix])
    plan.close()
    assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
    return rep


# the K^2 (2A-1) table of the 4096 x 2048 dictionary is 275 GB: no GRAM schedule for configs[2]
MODES = {"hl_c3_k4096_a2048_n32768_b2_s64": ["sgram", "recorrelate", "full"]}


@pytest.mark.parametrize("mode", ["sgram", "recorrelate", "gram", "full"])
@pytest.mark.parametrize("name", HEADLINE_SINGLE)
def test_headline_shape_against_reference(name, mode):
    if mode not in MODES.get(name, ["sgram", "recorrelate", "gram", "full"]):
        pytest.skip("the Gram table of this dictionary does not fit")
    g, d, sig = load_single(name)
    check(sig, d, g, mode)


@pytest.mark.parametrize("mode", ["sgram", "recorrelate", "gram", "full"])
def test_multiband_shapes_against_reference(mode):
    """configs[3]: each of the six bands (2048 ... 65536 samples, 1024 x 128 dictionaries) against the reference's
    per-band pursuit of its own band split; the engine's band split is checked against the same arrays."""
    g, x, dicts, bands = load_multiband()
    from matching_pursuit_b200 import decompose as mdec
    split = mdec.fft_frequency_decompose(x.to(DEV), int(g["sizes"][0]))
    for size in (int(s) for s in g["sizes"]):
        np.testing.assert_allclose(split[size].cpu().numpy(), bands[size].numpy(), rtol=1e-4, atol=2e-6)
        check(bands[size], dicts[size], g, mode, suffix=f"_{size}")


@pytest.mark.parametrize("name", HEADLINE_LONG)
def test_long_atoms_against_reference(name):
    """Atoms of 4096, 8192 and 16384 samples (experiments/archive/e_2023_3_8/experiment.py:352-358,
    e_2023_12_18/experiment.py:22-24): longer than a plan's window transform, coded through the drop-in entry
    point, which correlates them in 2048-sample parts."""
    g, d, sig = load_single(name)
    b, n = sig.shape[0], sig.shape[-1]

    def run(signals, steps):
        x = torch.from_numpy(np.ascontiguousarray(signals, dtype=np.float32)).to(DEV).view(signals.shape[0], 1, n)
        atom, pos, val, res = mpb.sparse_code_arrays(x, d.to(DEV), steps)
        return atom.cpu().numpy(), pos.cpu().numpy(), val.cpu().numpy(), res.cpu().numpy().reshape(-1, n)

    rep = compare_with_resync(run, sig.numpy().reshape(b, n), O.unit_norm(d).numpy(), g["atom"], g["pos"], g["val"],
                              g["margin"], g["residual"])
    assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
    # the reference-format entry point and the dense helper take long atoms too
    flat, scatter, residual = mpb.sparse_code(sig.to(DEV), d.to(DEV), 4, flatten=True, return_residual=True)
    recon = scatter(tuple(sig.shape), flat)
    np.testing.assert_allclose((recon + residual).cpu().numpy(), sig.numpy(), atol=2e-5)
    fm = mpb.conv.torch_conv(sig[:1].to(DEV), O.unit_norm(d).to(DEV))
    want = O.correlate_direct(sig[:1], O.unit_norm(d))
    np.testing.assert_allclose(fm.cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5 * float(want.abs().max()))
