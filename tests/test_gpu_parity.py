"""GPU parity tests: the CUDA path, called through the C ABI (ctypes binding ->
``libmpb200.so``), against the CPU oracle and the committed golden vectors that
the unmodified reference produced (tests/golden, oracle/make_golden.py).

Tolerances (BASELINE.json north_star): identical (atom, position) wherever the
top-2 relative margin exceeds 1e-5; amplitudes and residual energy within 1e-4
relative."""
import glob
import os

import numpy as np
import pytest
import torch

import matching_pursuit_b200 as mpb
from oracle import mp_oracle as O
from parity import MARGIN, RTOL, compare_trace, compare_with_oracle_trace, compare_with_resync, resync_against_trace

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SC_CASES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "sc_*.npz")))
DEV = "cuda:0"


def run_plan(signal, d, steps, mode="recorrelate", atom_range=None):
    b, _, n = signal.shape
    plan = mpb.Plan(d.shape[0], d.shape[1], n, b, mode=mode, device=DEV, atom_range=atom_range)
    plan.set_dictionary(d)
    atom, pos, val, res = plan.sparse_code(signal.to(DEV), steps)
    torch.cuda.synchronize()
    return atom.cpu().numpy(), pos.cpu().numpy(), val.cpu().numpy(), res.cpu().numpy()


def plan_runner(d, n, max_batch, mode, **kw):
    """``run(signals (b, N) numpy, steps)`` for parity.compare_with_resync: one plan, re-used for the restarts."""
    plan = mpb.Plan(d.shape[0], d.shape[-1], n, max_batch, mode=mode, device=DEV, **kw).set_dictionary(d)

    def run(signals, steps):
        sig = torch.from_numpy(np.ascontiguousarray(signals, dtype=np.float32)).to(DEV)
        out = plan.sparse_code(sig, steps)
        torch.cuda.synchronize()
        return tuple(t.cpu().numpy() for t in out)

    return run, plan


def check_against_trace(sig, d, tr, mode, min_fraction=0.0, **kw):
    """Every step of every signal against the oracle trace (restarting from the oracle's residual after an
    ambiguous step that was resolved differently); at least `min_fraction` of the (signal, step) pairs must have
    been binding (margin above 1e-5 and amplitude above the round-off floor)."""
    b, n = sig.shape[0], sig.shape[-1]
    run, plan = plan_runner(d, n, b, mode, **kw)
    rep = resync_against_trace(run, sig.numpy(), tr)
    plan.close()
    assert rep.checked == rep.eligible and rep.checked >= min_fraction * rep.total, rep
    return rep


# --------------------------------------------------------------------------
# golden vectors from the live reference
# --------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["recorrelate", "full", "gram", "sgram"])
@pytest.mark.parametrize("path", [p for p in SC_CASES if "lcn" not in p],
                         ids=[os.path.basename(p)[:-4] for p in SC_CASES if "lcn" not in p])
def test_golden_sparse_code(path, mode):
    g = np.load(path)
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    b, n = sig.shape[0], sig.shape[-1]
    run, plan = plan_runner(d, n, b, mode)
    rep = compare_with_resync(run, g["signal"].reshape(b, n), O.unit_norm(d).numpy(), g["atom"], g["pos"], g["val"],
                              g["margin"], g["residual"].reshape(b, n))
    plan.close()
    # every golden case but the all-zero signal and the 64-sample one holds a majority of binding steps
    floor = 0 if ("zero" in path or "single_atom" in path) else 0.5
    assert rep.checked == rep.eligible and rep.checked >= floor * rep.total, rep


def test_golden_zero_signal_is_exact():
    g = np.load(os.path.join(GOLDEN, "sc_zero_b1_n128_k4_a16.npz"))
    atom, pos, val, res = run_plan(torch.from_numpy(g["signal"]), torch.from_numpy(g["d"]), 3)
    assert (atom == 0).all() and (pos == 0).all() and (val == 0).all() and (res == 0).all()


def test_golden_local_contrast_norm():
    g = np.load(os.path.join(GOLDEN, "sc_lcn_b2_n512_k12_a32.npz"))
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    flat, scatter, residual = mpb.sparse_code(sig.to(DEV), d.to(DEV), int(g["steps"]), flatten=True,
                                              return_residual=True, local_contrast_norm=True)
    if (g["margin"] > MARGIN).all():
        order = np.array([(ai, j, int(p)) for ai, j, p, a in flat], dtype=np.int64)
        assert np.array_equal(order, g["flat_order"])
        np.testing.assert_allclose(residual.cpu().numpy(), g["residual"], rtol=1e-4, atol=2e-5)


def test_golden_dense_correlation():
    g = np.load(os.path.join(GOLDEN, "corr_helpers.npz"))
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    plan = mpb.Plan(d.shape[0], d.shape[1], sig.shape[-1], sig.shape[0], device=DEV, mode="recorrelate")
    plan.set_dictionary(d)
    fm = plan.correlate(sig.to(DEV).view(sig.shape[0], -1)).cpu().numpy()
    # the golden map was computed with the dictionary as given (already unit norm here)
    np.testing.assert_allclose(fm, g["torch_conv"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(plan.unit_dictionary().cpu().numpy(), O.unit_norm(d).numpy(), rtol=0, atol=1e-7)


def test_golden_sparse_feature_map():
    g = np.load(os.path.join(GOLDEN, "feature_map.npz"))
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    fm, res = mpb.sparse_feature_map(sig.to(DEV), d.to(DEV), n_steps=9, return_residual=True)
    fm, res = fm.cpu().numpy(), res.cpu().numpy()
    assert np.array_equal(fm != 0, g["fm"] != 0)
    np.testing.assert_allclose(fm, g["fm"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res, g["residual"], rtol=1e-4, atol=2e-6)


def test_golden_dictionary_learning_step():
    g = np.load(os.path.join(GOLDEN, "dictionary_learning.npz"))
    sig, d = torch.from_numpy(g["signal"]), torch.from_numpy(g["d"])
    learned = mpb.dictionary_learning_step(sig.to(DEV), d.to(DEV).clone(), n_steps=8)
    np.testing.assert_allclose(learned.cpu().numpy(), g["learned"], rtol=1e-4, atol=1e-5)


# --------------------------------------------------------------------------
# oracle on seeded inputs (sizes the oracle finishes in seconds)
# --------------------------------------------------------------------------
CASES = [
    # (K, A, N, B, S, family)
    (32, 64, 2048, 4, 40, "planted"),
    (33, 100, 3000, 3, 30, "noise"),        # odd atom count, ragged sizes
    (64, 256, 8192, 2, 48, "planted"),
    (16, 512, 4096, 2, 24, "noise"),        # M = 2048
    (8, 1024, 8192, 2, 16, "planted"),      # M = 4096
    (6, 2048, 16384, 1, 12, "planted"),     # M = 8192
    (5, 700, 700, 2, 8, "noise"),           # atom as long as the signal
    (9, 33, 1001, 2, 20, "noise"),          # odd everything: map rows padded for the bulk copies (SGRAM)
    (3, 2, 67, 3, 10, "noise"),             # tiny atoms, prime signal length
]


def make_case(k, a, n, b, s, family):
    d = O.make_dictionary(k, a, seed=k + a)
    if family == "planted":
        sig = O.make_planted_signals(d, b, n, max(4, s // 2), seed=n)
    else:
        sig = O.make_noise_signals(b, n, seed=n)
    return sig, d


@pytest.mark.parametrize("mode", ["recorrelate", "full", "gram", "sgram"])
@pytest.mark.parametrize("case", CASES, ids=[f"K{c[0]}_A{c[1]}_N{c[2]}_B{c[3]}_{c[5]}" for c in CASES])
def test_oracle_parity(case, mode):
    k, a, n, b, s, family = case
    sig, d = make_case(*case)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    rep = check_against_trace(sig, d, tr, mode, min_fraction=0.9)
    assert rep.checked > 0


def _fuzz_shapes():
    rng = np.random.default_rng(20261018)
    shapes = [(1, 1, 5, 1, 3), (1, 2, 2, 2, 2), (2, 3, 3, 1, 4), (3, 1, 64, 2, 6)]      # degenerate corners first
    for _ in range(28):
        a = int(rng.choice([2, 3, 5, 17, 31, 64, 100, 129, 255, 300]))
        n = int(rng.integers(a, 12 * a + 40))
        shapes.append((int(rng.integers(1, 41)), a, n, int(rng.integers(1, 5)), int(rng.integers(3, 16))))
    return shapes


@pytest.mark.parametrize("shape", _fuzz_shapes(), ids=lambda s: "K%d_A%d_N%d_B%d_S%d" % s)
def test_random_shapes_all_schedules_against_oracle(shape):
    """Seeded random shapes (single atoms, atoms of 1-3 samples, atoms as long as the signal, odd sizes) through
    all four schedules against the oracle; noise signals, so near-ties and truncated winners are common."""
    k, a, n, b, s = shape
    d = O.make_dictionary(k, a, seed=k * 131 + a)
    sig = O.make_noise_signals(b, n, seed=n * 7 + b)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    for mode in ("recorrelate", "full", "gram", "sgram"):
        atom, pos, val, res = run_plan(sig, d, s, mode)
        assert ((atom >= 0) & (atom < k)).all() and ((pos >= 0) & (pos < n)).all(), mode
        check_against_trace(sig, d, tr, mode)


@pytest.mark.parametrize("refresh", [0, 50])
def test_gram_mode_long_run_against_oracle(refresh):
    """Incremental Gram updates over 300 iterations (drift check), with and without periodic
    re-correlation; noise signals make a good share of the winners overhang the right edge,
    which exercises the FFT route inside GRAM mode."""
    k, a, n, b, s = 128, 256, 8192, 2, 300
    d = O.make_dictionary(k, a, seed=11)
    sig = torch.cat([O.make_planted_signals(d, 1, n, 150, seed=12), O.make_noise_signals(1, n, seed=13)], dim=0)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    run, plan = plan_runner(d, n, b, "gram")
    plan.set_refresh_every(refresh)
    assert plan.mode == "gram" and plan.info.gram_bytes == k * k * 2 * a * 4
    rep = resync_against_trace(run, sig.numpy(), tr)
    assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
    assert int((tr.pos + a > n).sum()) > 0      # the case does contain truncated winners


@pytest.mark.parametrize("refresh", [0, 50])
def test_sgram_mode_long_run_against_oracle(refresh):
    """Synthesised Gram rows (SGRAM) over 300 iterations, truncated winners included."""
    k, a, n, b, s = 128, 256, 8192, 2, 300
    d = O.make_dictionary(k, a, seed=11)
    sig = torch.cat([O.make_planted_signals(d, 1, n, 150, seed=12), O.make_noise_signals(1, n, seed=13)], dim=0)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    run, plan = plan_runner(d, n, b, "sgram")
    plan.set_refresh_every(refresh)
    assert plan.mode == "sgram" and plan.info.gram_bytes == 0 and plan.fft_size2 == 512
    rep = resync_against_trace(run, sig.numpy(), tr)
    assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep


@pytest.mark.parametrize("case", [(64, 1024, 8192, 3, 40), (12, 2048, 16384, 2, 24), (33, 700, 5000, 2, 30)],
                         ids=["A1024", "A2048", "A700"])
def test_sgram_position_free_tables_equal_exact_positions(case):
    """SGRAM's position-free block tables (on by default only for large batches) must give bit-identical events
    and residuals to the exact-position tables, and both must follow the oracle; noise in the second signal makes
    truncated winners (the FFT route that mixes exact positions into the tables) common."""
    k, a, n, b, s = case
    d = O.make_dictionary(k, a, seed=21)
    sig = torch.cat([O.make_planted_signals(d, b - 1, n, s // 2, seed=22), O.make_noise_signals(1, n, seed=23)], dim=0)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    outs = {}
    for on in (False, True):
        run, plan = plan_runner(d, n, b, "sgram")
        plan.set_position_free(on)
        outs[on] = run(sig.numpy().reshape(b, n), s)
        rep = resync_against_trace(run, sig.numpy(), tr)
        assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
        plan.close()
    for x, y in zip(outs[False], outs[True]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("case", [(512, 512, 32768, 1, 32), (33, 700, 5000, 3, 30), (64, 128, 4096, 2, 24),
                                  (16, 2048, 16384, 2, 12), (7, 64, 1500, 4, 9)],
                         ids=["configs0", "A700_odd_atoms", "A128", "A2048", "tiny"])
def test_fused_loop_equals_stream_ordered_loop(case):
    """Windowed re-correlation: the one-launch cooperative loop (k_pursue_fused, MPB200_OPT_FUSED_LOOP, the default
    when one CTA per (pair group, signal) is resident) must give bit-identical events and residuals to the
    stream-ordered loop of k_apply + k_corr per iteration, and both must follow the oracle.  A noise signal makes
    winners at the edges (truncated atoms, windows clipped at 0 and N) common."""
    k, a, n, b, s = case
    d = O.make_dictionary(k, a, seed=31)
    parts = [O.make_planted_signals(d, max(b - 1, 1), n, max(s // 2, 1), seed=32)]
    if b > 1:
        parts.append(O.make_noise_signals(1, n, seed=33))
    sig = torch.cat(parts, dim=0)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    outs, launches = {}, {}
    for on in (False, True):
        run, plan = plan_runner(d, n, b, "recorrelate")
        plan.set_fused_loop(on)
        before = mpb.lib().mpb200_launch_count()
        outs[on] = run(sig.numpy().reshape(b, n), s)
        launches[on] = mpb.lib().mpb200_launch_count() - before
        rep = resync_against_trace(run, sig.numpy(), tr)
        assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
        plan.close()
    for x, y in zip(outs[False], outs[True]):
        assert np.array_equal(x, y)
    assert launches[True] < launches[False] - (s - 2), launches     # the fused path was really taken: one launch for the loop


def test_fused_loop_falls_back_when_the_grid_does_not_fit():
    """More (pair group, signal) CTAs than the device holds at once: the option stays on, the stream-ordered loop runs
    (two launches per iteration) and the results are the oracle's."""
    k, a, n, b, s = 1024, 128, 2048, 4, 6           # 64 pair groups (+1 writer) x 4 signals > 148 SMs x 1 CTA
    d = O.make_dictionary(k, a, seed=41)
    sig = O.make_planted_signals(d, b, n, 4, seed=42)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    run, plan = plan_runner(d, n, b, "recorrelate")
    plan.set_fused_loop(True)
    before = mpb.lib().mpb200_launch_count()
    run(sig.numpy().reshape(b, n), s)
    assert mpb.lib().mpb200_launch_count() - before >= 2 * s - 1
    rep = resync_against_trace(run, sig.numpy(), tr)
    assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
    plan.close()


def test_sgram_two_cta_form_equals_three_cta_form_and_oracle():
    """k_delta's double-buffered two-CTA form (4096-point transforms, resident batches of >= 32 signals: pair spectrum
    in registers, next window prefetched into the second staging buffer) against the three-CTA form the same plan takes
    for a batch of 16 -- bit-identical events and residuals -- and against the oracle.  Noise signals make winners at
    both edges (windows clipped at 0, truncated atoms that take the FFT route and break the prefetch chain) common."""
    k, a, n, b, s = 10, 2048, 8192, 32, 10
    d = O.make_dictionary(k, a, seed=51)
    sig = torch.cat([O.make_planted_signals(d, b - 4, n, 5, seed=52), O.make_noise_signals(4, n, seed=53)], dim=0)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    run, plan = plan_runner(d, n, b, "sgram")
    assert plan.fft_size2 == 4096
    whole = run(sig.numpy().reshape(b, n), s)                       # one launch per iteration over 32 signals
    halves = [run(sig.numpy().reshape(b, n)[i:i + 16], s) for i in (0, 16)]
    for j, x in enumerate(whole):
        assert np.array_equal(x, np.concatenate([h[j] for h in halves], axis=0))
    rep = resync_against_trace(run, sig.numpy(), tr)
    assert rep.checked == rep.eligible and rep.checked >= 0.9 * rep.total, rep
    plan.close()


def test_sgram_sub_batches_equal_one_batch():
    """A resident-map budget that holds 3 of 7 signals: the batch is walked in balanced sub-batches
    and every signal gets the result it gets alone (signals are independent problems)."""
    k, a, n, b, s = 24, 128, 4096, 7, 20
    d = O.make_dictionary(k, a, seed=3)
    sig = O.make_planted_signals(d, b, n, 12, seed=4)
    whole = mpb.Plan(k, a, n, b, mode="sgram", device=DEV).set_dictionary(d)
    assert whole.resident_batch == b
    ref = [t.cpu() for t in whole.sparse_code(sig.to(DEV), s)]
    part = mpb.Plan(k, a, n, b, mode="sgram", device=DEV, gram_budget_bytes=3 * k * n * 4 + 1024).set_dictionary(d)
    assert part.resident_batch == 3
    got = [t.cpu() for t in part.sparse_code(sig.to(DEV), s)]
    for r, g in zip(ref, got):
        assert torch.equal(r, g)
    hatom, hpos, hval, hres = part.sparse_code_host(sig, s)
    assert torch.equal(hatom, ref[0]) and torch.equal(hpos, ref[1]) and torch.equal(hval, ref[2])
    assert torch.equal(hres, ref[3])
    fm = part.correlate(sig.to(DEV)[:, 0])
    assert torch.equal(fm.cpu(), whole.correlate(sig.to(DEV)[:, 0]).cpu())
    with pytest.raises(mpb.MpbError):
        part.begin(sig.to(DEV))      # the step-wise interface holds one resident batch


def test_auto_mode_resolution():
    assert mpb.Plan(512, 1024, 2 ** 15, 64, device=DEV).mode == "gram"          # BASELINE configs[1]
    assert mpb.Plan(512, 512, 2 ** 15, 1, device=DEV).mode == "recorrelate"     # too few work items for the map modes
    assert mpb.Plan(4096, 2048, 2 ** 15, 32, device=DEV).mode == "sgram"
    assert mpb.Plan(16384, 2048, 2 ** 20, 1, device=DEV, atom_range=(0, 2048)).mode == "sgram"   # one rank of configs[4]
    p = mpb.Plan(4096, 2048, 2 ** 15, 1024, device=DEV)                          # BASELINE configs[2]: 275 GB table
    assert p.mode == "sgram" and 32 <= p.resident_batch < 1024
    p.close()
    assert mpb.Plan(4096, 2048, 2 ** 15, 8, device=DEV, mode="recorrelate").mode == "recorrelate"


def test_config1_full_size():
    """BASELINE.json configs[0]: 1 x 2^15 samples, 512 atoms x 512, 32 iterations."""
    d = O.make_dictionary(512, 512, seed=0)
    sig = O.make_planted_signals(d, 1, 2 ** 15, 32, seed=1)
    tr = O.greedy_pursuit(sig, d, 32, want_margin=True)
    for mode in ("recorrelate", "sgram", "gram"):
        check_against_trace(sig, d, tr, mode, min_fraction=0.9)


def test_host_entry_equals_device_entry():
    sig, d = make_case(32, 64, 2048, 4, 40, "planted")
    plan = mpb.Plan(32, 64, 2048, 4, device=DEV, mode="recorrelate").set_dictionary(d)
    a0, p0, v0, r0 = plan.sparse_code(sig.to(DEV), 20)
    a1, p1, v1, r1 = plan.sparse_code_host(sig.view(4, -1), 20)
    assert torch.equal(a0.cpu(), a1) and torch.equal(p0.cpu(), p1) and torch.equal(v0.cpu(), v1)
    assert torch.equal(r0.cpu(), r1)
    # inputs are not modified, repeated calls are reproducible
    a2, p2, v2, r2 = plan.sparse_code(sig.to(DEV), 20)
    assert torch.equal(a0, a2) and torch.equal(v0, v2) and torch.equal(r0, r2)


def test_atom_sharded_stepwise_equals_oracle():
    """Atom sharding emulated on one GPU: R plans own disjoint atom ranges, their
    local winners are reduced with the reference tie-break, every plan applies
    the global winner."""
    k, a, n, b, s = 37, 128, 4096, 3, 24
    d = O.make_dictionary(k, a, seed=5)
    sig = O.make_planted_signals(d, b, n, 12, seed=6)
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    bounds = [0, 11, 24, 37]
    plans = [mpb.Plan(k, a, n, b, device=DEV, mode="recorrelate", atom_range=(lo, hi)).set_dictionary(d)
             for lo, hi in zip(bounds[:-1], bounds[1:])]
    for p in plans:
        p.begin(sig.to(DEV))
    atoms, poss, vals = [], [], []
    for _ in range(s):
        cand = torch.cat([p.local_best() for p in plans], dim=0)
        win = mpb.reduce_best(cand, len(plans), b)
        v, kk, pp = mpb.unpack_best(win)
        atoms.append(kk.cpu().numpy()); poss.append(pp.cpu().numpy()); vals.append(v.cpu().numpy())
        for p in plans:
            p.apply(win)
    res = plans[0].residual().cpu().numpy()
    for p in plans[1:]:
        assert np.array_equal(p.residual().cpu().numpy(), res)
    checked = compare_trace(tr.atom.numpy(), tr.pos.numpy(), tr.val.numpy(), tr.margin.numpy(), tr.residual.numpy()[:, 0],
                            np.stack(atoms), np.stack(poss), np.stack(vals), res)
    assert checked >= 0.9 * s * b


@pytest.mark.parametrize("mode", ["recorrelate", "sgram"])
def test_atom_sharded_fused_exchange_equals_oracle(mode):
    """The exchange fused into the pursuit (mailboxes in peer memory, mpb200_exchange_*), exercised in one
    process: three atom-sharded plans on three streams of one GPU trade their candidates through each
    other's mailboxes; every plan must report the oracle's sequence and the same residual."""
    k, a, n, b, s = 37, 128, 4096, 3, 24
    d = O.make_dictionary(k, a, seed=5)
    sig = O.make_planted_signals(d, b, n, 12, seed=6).to(DEV)
    tr = O.greedy_pursuit(sig.cpu(), d, s, want_margin=True)
    bounds = [0, 11, 24, 37]
    world = len(bounds) - 1
    plans = [mpb.Plan(k, a, n, b, device=DEV, mode=mode, atom_range=(lo, hi)).set_dictionary(d)
             for lo, hi in zip(bounds[:-1], bounds[1:])]
    with pytest.raises(mpb.MpbError):
        plans[0].sparse_code(sig, s)                 # sharded plan without a connected exchange
    for r, p in enumerate(plans):
        assert len(p.exchange_create(world, r)) == 64
    boxes = [p.exchange_mailbox() for p in plans]
    for p in plans:
        p.exchange_connect_local(boxes)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=DEV) for _ in plans]
    outs = []
    for rep in range(2):                              # twice: sequence numbers keep counting across calls
        outs = []
        for p, st in zip(plans, streams):
            with torch.cuda.stream(st):
                outs.append(p.sparse_code(sig, s))
        torch.cuda.synchronize()
    assert not any(p.exchange_timed_out() for p in plans)
    atom, pos, val, res = (t.cpu().numpy() for t in outs[0])
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert torch.equal(x, y)
    assert compare_with_oracle_trace(tr, atom, pos, val, res) >= 0.9 * s * b


# --------------------------------------------------------------------------
# drop-in return conventions
# --------------------------------------------------------------------------
def test_dropin_return_conventions():
    sig, d = make_case(32, 64, 2048, 4, 40, "planted")
    s = 24
    tr = O.greedy_pursuit(sig, d, s, want_margin=True)
    if not (tr.margin.numpy() > MARGIN).all():
        pytest.skip("ambiguous step in the seeded case")
    want_flat, want_scatter, want_res = O.sparse_code(sig, d, s, flatten=True, return_residual=True)
    flat, scatter, res = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, flatten=True, return_residual=True)
    assert [(ai, j, int(p)) for ai, j, p, _ in flat] == [(ai, j, int(p)) for ai, j, p, _ in want_flat]
    ai, j, p, a = flat[0]
    assert isinstance(ai, int) and isinstance(j, int) and p.shape == (1, 1) and p.dtype == torch.int64
    assert a.shape == (1, 1, 64) and a.dtype == torch.float32 and a.is_cuda
    for (_, _, _, a0), (_, _, _, a1) in zip(flat, want_flat):
        np.testing.assert_allclose(a0.cpu().numpy(), a1.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res.cpu().numpy(), want_res.numpy(), rtol=1e-4, atol=2e-6)
    recon = scatter(tuple(sig.shape), flat)
    np.testing.assert_allclose(recon.cpu().numpy(), want_scatter(tuple(sig.shape), want_flat).numpy(),
                               rtol=1e-4, atol=2e-6)
    # signal = decode + residual
    np.testing.assert_allclose((recon + res).cpu().numpy(), sig.numpy(), atol=5e-6)
    # grouped form: keys in first-seen order
    inst, _ = mpb.sparse_code(sig.to(DEV), d.to(DEV), s)
    want_inst, _ = O.sparse_code(sig, d, s)
    assert list(inst.keys()) == list(want_inst.keys())
    for key in inst:
        assert [(x[1], int(x[2])) for x in inst[key]] == [(x[1], int(x[2])) for x in want_inst[key]]
    # sparse feature map form
    _, _, sfm = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, flatten=True, return_sparse_feature_map=True)
    _, _, want_sfm = O.sparse_code(sig, d, s, flatten=True, return_sparse_feature_map=True)
    np.testing.assert_allclose(sfm.cpu().numpy(), want_sfm.numpy(), rtol=1e-4, atol=1e-6)
    # CPU tensors in -> CPU tensors out
    flat_c, _, res_c = mpb.sparse_code(sig, d, s, flatten=True, return_residual=True)
    assert not res_c.is_cuda and not flat_c[0][3].is_cuda
    assert [(ai, j, int(p)) for ai, j, p, _ in flat_c] == [(ai, j, int(p)) for ai, j, p, _ in want_flat]


def test_dropin_callbacks_see_the_dense_map():
    sig, d = make_case(16, 32, 512, 2, 10, "planted")
    s = 6
    seen, want_seen = [], []

    def visit(fm, ai, p, a):
        seen.append((ai, int(p), tuple(fm.shape), float(a.abs().max())))

    def want_visit(fm, ai, p, a):
        want_seen.append((ai, int(p), tuple(fm.shape), float(a.abs().max())))

    mpb.sparse_code(sig.to(DEV), d.to(DEV), s, visit_key_point=visit)
    O.sparse_code(sig, d, s, visit_key_point=want_visit)
    assert [x[:3] for x in seen] == [x[:3] for x in want_seen]
    np.testing.assert_allclose([x[3] for x in seen], [x[3] for x in want_seen], rtol=1e-4)
    # compute_feature_map seam: a callback that supplies the map (here: the library's own dense correlation)
    calls = []

    def cfm(residual, du):
        calls.append(tuple(residual.shape))
        plan = mpb.get_plan(du.shape[0], du.shape[1], residual.shape[-1], residual.shape[0], residual.device)
        plan.set_dictionary(du, normalize=False)
        return plan.correlate(residual.view(residual.shape[0], -1))

    flat, _ = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, flatten=True, compute_feature_map=cfm)
    want_flat, _ = O.sparse_code(sig, d, s, flatten=True)
    assert calls == [(2, 1, 512)] * s
    assert [(ai, j, int(p)) for ai, j, p, _ in flat] == [(ai, j, int(p)) for ai, j, p, _ in want_flat]
    emb, res = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, extract_atom_embedding=lambda fm, dd: fm.amax(dim=(1, 2)))
    want_emb, want_res = O.sparse_code(sig, d, s, extract_atom_embedding=lambda fm, dd: fm.amax(dim=(1, 2)))
    np.testing.assert_allclose(torch.stack(emb).cpu().numpy(), torch.stack(want_emb).numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res.cpu().numpy(), want_res.numpy(), rtol=1e-4, atol=2e-6)


def test_select_dense_tie_break_and_subtract():
    fm = torch.zeros(2, 3, 40, device=DEV)
    fm[0, 2, 7] = 5.0; fm[0, 1, 30] = 5.0; fm[0, 1, 31] = 5.0     # tie -> lowest atom, then lowest position
    fm[1] = -1.0; fm[1, 0, 0] = -3.0                               # all negative: signed max, first index among ties
    best = mpb.select_dense(fm)
    v, k, p = mpb.unpack_best(best)
    assert k.tolist() == [1, 0] and p.tolist() == [30, 1] and v.tolist() == [5.0, -1.0]
    d = O.make_dictionary(3, 8, seed=1).to(DEV)
    r = torch.ones(2, 40, device=DEV)
    want = r.clone()
    want[0, 30:38] -= 5.0 * d[1]
    want[1, 1:9] -= -1.0 * d[0]
    mpb.subtract(r, d, best)
    assert torch.equal(r, want)
    # right-edge truncation
    fm2 = torch.zeros(1, 3, 40, device=DEV); fm2[0, 2, 36] = 2.0
    best2 = mpb.select_dense(fm2)
    r2 = torch.zeros(1, 40, device=DEV)
    mpb.subtract(r2, d, best2)
    assert torch.equal(r2[0, 36:], -(2.0 * d[2, :4])) and (r2[0, :36] == 0).all()


def test_scatter_segments_modes():
    a = 16
    d = O.make_dictionary(4, a, seed=2)
    events = [(1, 0, torch.tensor([[5]]), (0.5 * d[1]).view(1, 1, a)),
              (3, 1, torch.tensor([[60]]), (2.0 * d[3]).view(1, 1, a)),     # overhangs the right edge of N=64
              (0, 0, torch.tensor([[10]]), (-1.0 * d[0]).view(1, 1, a))]
    want_scatter = O.make_scatter(64, a)
    scatter = mpb.build_scatter_segments(64, a, device=DEV)
    got = scatter((2, 1, 64), events)
    np.testing.assert_allclose(got.cpu().numpy(), want_scatter((2, 1, 64), events).numpy(), atol=1e-7)
    base = torch.randn(2, 1, 64)
    got = scatter(base.to(DEV), events)
    np.testing.assert_allclose(got.cpu().numpy(), want_scatter(base, events).numpy(), atol=1e-7)
    # one channel per event (modules/matchingpursuit.py:50)
    got = scatter((2, 3, 64), events)
    np.testing.assert_allclose(got.cpu().numpy(), want_scatter((2, 3, 64), events).numpy(), atol=1e-7)
    assert scatter((2, 1, 64), []).abs().sum().item() == 0


# --------------------------------------------------------------------------
# size-independent properties at sizes the oracle cannot reach
# --------------------------------------------------------------------------
def test_properties_at_scale():
    """512 atoms x 1024 samples, 16 x 2^15 signals, 128 steps: the events must
    explain the residual (signal = decode(events) + residual), every value must
    be the correlation of the atom with the residual it was picked from, and the
    residual energy must fall by value^2 each step (unit atoms, no truncation)."""
    k, a, n, b, s = 512, 1024, 2 ** 15, 16, 128
    d = O.make_dictionary(k, a, seed=0)
    sig = O.make_planted_signals(d, b, n, 64, seed=1)
    plan = mpb.Plan(k, a, n, b, device=DEV, mode="recorrelate").set_dictionary(d)
    atom, pos, val, res = plan.sparse_code(sig.to(DEV), s)
    du = plan.unit_dictionary()
    out = torch.zeros(b, n, device=DEV)
    rows = torch.arange(b, device=DEV).repeat_interleave(s)
    mpb.scatter_add(out, du, atom.reshape(-1), rows, pos.reshape(-1), val.reshape(-1))
    np.testing.assert_allclose((out + res).cpu().numpy(), sig.view(b, n).numpy(), atol=2e-5)
    # energy bookkeeping in float64 on the host
    e0 = (sig.view(b, n).double() ** 2).sum(-1)
    e1 = (res.cpu().double() ** 2).sum(-1)
    interior = (pos.cpu() + a <= n)
    drop = (val.cpu().double() ** 2 * interior).sum(-1)
    assert ((e0 - e1) >= 0.999 * drop - 1e-6).all()
    assert (val.cpu()[:, 0] > 0).all() and (val.cpu() >= -1e-6).all()
    # first step of every signal equals the oracle's (one dense correlation on the CPU is affordable)
    fm = O.correlate_direct(sig, O.unit_norm(d))
    v0, i0 = fm.reshape(b, -1).max(-1)
    assert torch.equal(i0 // n, atom[:, 0].cpu().long()) and torch.equal(i0 % n, pos[:, 0].cpu().long())
    np.testing.assert_allclose(val[:, 0].cpu().numpy(), v0.numpy(), rtol=1e-4)


def test_headline_shape_schedules_agree():
    """BASELINE configs[2] at its full dictionary, signal length and iteration count (4096 x 2048, 2^15,
    512 iterations; 4 of the 1024 signals): the CPU oracle needs minutes per signal here, so parity is
    carried by size-independent properties -- the incremental SGRAM schedule (synthesised Gram rows, 511
    accumulated fp32 map updates per position) must agree with windowed re-correlation, which recomputes
    every value it reads from the residual: same (atom, position) for the planted half of the run, values
    and residual energies within the north-star tolerance over the whole run, and the events must explain
    the residual."""
    k, a, n, b, s = 4096, 2048, 2 ** 15, 4, 512
    d = O.make_dictionary(k, a, seed=0)
    sig = O.make_planted_signals(d, b, n, 256, seed=1).to(DEV)
    runs = {}
    for mode in ("recorrelate", "sgram"):
        plan = mpb.Plan(k, a, n, b, device=DEV, mode=mode).set_dictionary(d)
        runs[mode] = [t.cpu() for t in plan.sparse_code(sig, s)]
        if mode == "sgram":
            du = plan.unit_dictionary()
        plan.close()
    (ra, rp, rv, rr), (sa, sp, sv, sr) = runs["recorrelate"], runs["sgram"]
    assert torch.equal(ra[:, :256], sa[:, :256]) and torch.equal(rp[:, :256], sp[:, :256])
    same = (ra == sa) & (rp == sp)
    assert float(same.float().mean()) > 0.95          # later steps may part ways at a near-tie, legitimately
    agree = same.cumprod(dim=1).bool()                # steps before the first divergence of each signal
    assert float(((rv - sv).abs() * agree).max()) <= RTOL * float(rv.abs().max())
    e_r, e_s = (rr.double() ** 2).sum(-1), (sr.double() ** 2).sum(-1)
    assert ((e_r - e_s).abs() <= 1e-3 * e_r).all()
    out = torch.zeros(b, n, device=DEV)
    rows = torch.arange(b, device=DEV).repeat_interleave(s)
    mpb.scatter_add(out, du, sa.reshape(-1).to(DEV), rows, sp.reshape(-1).to(DEV), sv.reshape(-1).to(DEV))
    np.testing.assert_allclose((out.cpu() + sr).numpy(), sig.view(b, n).cpu().numpy(), atol=5e-5)


@pytest.mark.parametrize("mode", ["recorrelate", "gram", "sgram", "full"])
def test_non_finite_input_is_memory_safe(mode):
    """NaN / Inf samples make the reference's result meaningless (torch.max returns NaN); the engine promises
    only that nothing is read or written out of bounds: events stay inside the dictionary and the signal, the
    launch sequence completes, and an untouched second signal is coded exactly as it is alone."""
    k, a, n, b, s = 16, 64, 2048, 2, 12
    d = O.make_dictionary(k, a, seed=1)
    sig = O.make_planted_signals(d, b, n, 6, seed=2)
    bad = sig.clone()
    bad[0, 0, 100] = float("nan")
    bad[0, 0, 900] = float("inf")
    atom, pos, val, res = run_plan(bad, d, s, mode)
    assert ((atom >= 0) & (atom < k)).all() and ((pos >= 0) & (pos < n)).all()
    ref = run_plan(sig[1:], d, s, mode)
    assert np.array_equal(atom[1:], ref[0]) and np.array_equal(pos[1:], ref[1]) and np.array_equal(val[1:], ref[2])


def test_dictionary_tables_follow_the_tensor_even_through_dot_data():
    """The plan skips rebuilding its tables when handed the same, unmodified dictionary tensor again -- and must
    not be fooled by in-place writes, including writes through `.data`, which do not bump torch's version counter."""
    k, a, n, b, s = 16, 64, 2048, 2, 10
    d = O.make_dictionary(k, a, seed=1).to(DEV)
    sig = O.make_planted_signals(d.cpu(), b, n, 6, seed=2).to(DEV)
    plan = mpb.Plan(k, a, n, b, device=DEV, mode="recorrelate")
    first = [t.cpu() for t in plan.set_dictionary(d).sparse_code(sig, s)]
    again = [t.cpu() for t in plan.set_dictionary(d).sparse_code(sig, s)]          # cached tables
    assert all(torch.equal(x, y) for x, y in zip(first, again))
    d2 = O.make_dictionary(k, a, seed=7).to(DEV)
    want = [t.cpu() for t in mpb.Plan(k, a, n, b, device=DEV, mode="recorrelate").set_dictionary(d2).sparse_code(sig, s)]
    d.data.copy_(d2)                                                                # same object, same _version
    got = [t.cpu() for t in plan.set_dictionary(d).sparse_code(sig, s)]
    assert all(torch.equal(x, y) for x, y in zip(want, got))
    d.copy_(O.make_dictionary(k, a, seed=1).to(DEV))                               # versioned in-place write
    back = [t.cpu() for t in plan.set_dictionary(d).sparse_code(sig, s)]
    assert all(torch.equal(x, y) for x, y in zip(first, back))


def test_errors():
    with pytest.raises(mpb.MpbError):
        mpb.Plan(4, 5000, 128, 1, device=DEV)            # window FFT longer than supported
    with pytest.raises(mpb.MpbError):
        mpb.Plan(4, 16, 128, 1, device=DEV, atom_range=(3, 2))
    plan = mpb.Plan(4, 16, 128, 2, device=DEV)
    with pytest.raises(mpb.MpbError):
        plan.sparse_code(torch.zeros(1, 128, device=DEV), 2)     # dictionary not set
    plan.set_dictionary(torch.randn(4, 16))
    with pytest.raises(mpb.MpbError):
        plan.sparse_code(torch.zeros(3, 128, device=DEV), 2)     # batch > max_batch
    with pytest.raises(mpb.MpbError):
        plan.set_dictionary(torch.randn(5, 16))
    with pytest.raises(ValueError):
        mpb.sparse_code(torch.zeros(128, device=DEV), torch.randn(4, 16))   # not (B,C,N), as the reference


# --------------------------------------------------------------------------
# incremental local-contrast-norm selection (modules/matchingpursuit.py:286-296)
# --------------------------------------------------------------------------
LCN_CASES = [
    # (K, A, N, B, S)
    (12, 32, 512, 2, 16),
    (40, 100, 3000, 2, 24),      # ragged sizes, blocks of 16 positions
    (64, 256, 8192, 2, 30),
    (5, 700, 2000, 1, 10),       # fewer atoms than the box is tall
    (512, 512, 2 ** 15, 1, 24),  # experiments/e_2023_7_20's shape (512 x 512 dictionary)
]


@pytest.mark.parametrize("mode", ["sgram", "gram"])
@pytest.mark.parametrize("case", LCN_CASES, ids=lambda c: "K%d_A%d_N%d_B%d_S%d" % c)
def test_incremental_lcn_against_oracle(case, mode):
    """The engine refreshes the normalised map only in the winner's window (+-4 columns) and keeps a second
    block/row-max hierarchy on it; the oracle recomputes avg_pool2d over the whole map every step.  Noise signals:
    truncated winners and near-ties are common.  The reported values are RAW map values (:296)."""
    k, a, n, b, s = case
    d = O.make_dictionary(k, a, seed=k + a)
    sig = torch.cat([O.make_planted_signals(d, 1, n, max(4, s // 2), seed=n)] +
                    ([O.make_noise_signals(b - 1, n, seed=n + 1)] if b > 1 else []), dim=0)
    tr = O.greedy_pursuit(sig, d, s, local_contrast_norm=True, want_margin=True)
    run, plan = plan_runner(d, n, b, mode)
    plan.set_local_contrast_norm(True)
    rep = resync_against_trace(run, sig.numpy(), tr)
    plan.close()
    assert rep.checked == rep.eligible and rep.checked >= 0.8 * rep.total, rep


def test_lcn_dropin_takes_the_incremental_path_and_matches_the_dense_schedule():
    k, a, n, b, s = 24, 64, 2048, 2, 14
    d = O.make_dictionary(k, a, seed=31)
    sig = O.make_planted_signals(d, b, n, 8, seed=32)
    launches = mpb.lib().mpb200_launch_count()
    flat, scatter, res = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, flatten=True, return_residual=True,
                                         local_contrast_norm=True)
    fast = mpb.lib().mpb200_launch_count() - launches
    seen = []
    launches = mpb.lib().mpb200_launch_count()
    flat2, _, res2 = mpb.sparse_code(sig.to(DEV), d.to(DEV), s, flatten=True, return_residual=True,
                                     local_contrast_norm=True, visit_key_point=lambda fm, ai, p, at: seen.append(ai))
    dense = mpb.lib().mpb200_launch_count() - launches
    assert len(seen) == b * s
    assert [(ai, j, int(p)) for ai, j, p, _ in flat] == [(ai, j, int(p)) for ai, j, p, _ in flat2]
    np.testing.assert_allclose(res.cpu().numpy(), res2.cpu().numpy(), rtol=1e-4, atol=2e-6)
    tr = O.greedy_pursuit(sig, d, s, local_contrast_norm=True, want_margin=True)
    if (tr.margin.numpy() > MARGIN).all():
        want = O.sparse_code(sig, d, s, flatten=True, local_contrast_norm=True)[0]
        assert [(ai, j, int(p)) for ai, j, p, _ in flat] == [(ai, j, int(p)) for ai, j, p, _ in want]
    assert fast < dense          # three launches per step instead of a dense map + 81-tap selection per step
