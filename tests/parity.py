"""Shared parity helpers (BASELINE.json north_star rule): identical
(atom, position) wherever the reference's top-2 relative margin exceeds 1e-5;
SIGNED amplitudes and residual energy within 1e-4 relative.

After a legitimately ambiguous step (margin <= 1e-5, or the reference's own
amplitude down at fp32 round-off of the signal scale, see ATOL_OF_PEAK) two
correct implementations may pick different atoms and their sequences diverge.
The comparison does not stop there: :func:`compare_with_resync` rebuilds the
reference's residual after that step from the reference's own events (the
update is ``fl(r - fl(v * d))``, modules/matchingpursuit.py:305, 328, so the
replay is exact) and restarts the implementation under test from it, so EVERY
step of every signal is compared and the residual energy is always checked
(SURVEY.md section 8c)."""
import numpy as np

MARGIN = 1e-5
RTOL = 1e-4
# Absolute floor of the amplitude check, as a fraction of the signal's LARGEST amplitude: once a signal is
# explained down to fp32 round-off (a 4-sample signal after 8 atoms: amplitudes of 4e-6 against 1.0), the
# reference's own correlation carries an absolute error of about 1e-7 of the signal scale, so a purely relative
# tolerance on such a value compares noise with noise.
ATOL_OF_PEAK = 1e-6


class ParityReport:
    """checked: (signal, step) pairs whose (atom, position) had to match and did; eligible: pairs with a margin
    above the threshold and an amplitude above the round-off floor; resyncs: restarts from the reference's
    residual after an ambiguous step that the implementation resolved differently."""

    def __init__(self):
        self.checked = self.eligible = self.total = self.resyncs = self.energy_checks = 0

    def __repr__(self):
        return (f"ParityReport(checked={self.checked}, eligible={self.eligible}, total={self.total}, "
                f"resyncs={self.resyncs}, energy_checks={self.energy_checks})")


def replay(residual, d_unit, atom, pos, val):
    """Apply events to a float32 residual (N,) exactly as the reference does: scaled = fl(d * v), then
    r = fl(r - scaled), truncated at the right edge (modules/matchingpursuit.py:305, :33-56, :328)."""
    r = np.array(residual, dtype=np.float32, copy=True)
    n, a = r.shape[0], d_unit.shape[1]
    for k, p, v in zip(atom, pos, val):
        keep = min(a, n - int(p))
        scaled = (d_unit[int(k), :keep] * np.float32(v)).astype(np.float32)
        r[int(p):int(p) + keep] = (r[int(p):int(p) + keep] - scaled).astype(np.float32)
    return r


def compare_with_resync(run, signal, d_unit, ref_atom, ref_pos, ref_val, ref_margin, ref_residual=None,
                        margin=MARGIN, rtol=RTOL):
    """``run(signals (b, N) float32 numpy, n_steps) -> (atom (b,S), pos (b,S), val (b,S), residual (b,N))`` numpy
    arrays of the implementation under test.  ``signal`` (B, N); ``d_unit`` (K, A) the unit-normed dictionary the
    events refer to; reference step arrays are (S, B), ``ref_val`` SIGNED; ``ref_residual`` (B, N) or None.
    Returns a :class:`ParityReport`."""
    signal = np.asarray(signal, dtype=np.float32).reshape(np.asarray(signal).shape[0], -1)
    d_unit = np.asarray(d_unit, dtype=np.float32)
    steps, batch = ref_atom.shape
    rep = ParityReport()
    rep.total = steps * batch
    floor = [ATOL_OF_PEAK * float(np.abs(ref_val[:, j]).max()) if steps else 0.0 for j in range(batch)]

    def must_match(s, j):
        return bool(ref_margin[s, j] > margin) and abs(float(ref_val[s, j])) > 16.0 * floor[j]

    def walk(j, s0, atom, pos, val):
        """Compare one signal's steps s0.. against the reference; returns the first step that diverged
        legitimately (ambiguous), or `steps` when the sequences agree to the end."""
        for i in range(atom.shape[0]):
            s = s0 + i
            got, want = (int(atom[i]), int(pos[i])), (int(ref_atom[s, j]), int(ref_pos[s, j]))
            if got != want:
                assert not must_match(s, j), (f"signal {j} step {s}: got {got}, want {want} "
                                              f"(margin {ref_margin[s, j]:.3g}, value {float(ref_val[s, j]):.6g})")
                return s
            want_v = float(ref_val[s, j])
            assert abs(float(val[i]) - want_v) <= rtol * max(abs(want_v), 1e-12) + floor[j], \
                (s, j, float(val[i]), want_v)
            if must_match(s, j):
                rep.checked += 1
        return steps

    for s in range(steps):
        for j in range(batch):
            rep.eligible += int(must_match(s, j))

    atom, pos, val, residual = run(signal, steps)
    pending = []
    for j in range(batch):
        stop = walk(j, 0, atom[j], pos[j], val[j])
        if stop < steps:
            pending.append((j, stop, signal[j]))
        elif ref_residual is not None:
            _check_energy(residual[j], ref_residual[j], rtol, j)
            rep.energy_checks += 1
    # signals that resolved an ambiguous step differently: restart them from the reference's residual after it
    while pending:
        j, stop, base = pending.pop()
        # `base` is the residual the compared segment started from; replay the reference's events of that segment
        s_from = getattr(base, "_start", 0)
        rep.resyncs += 1
        synced = replay(np.asarray(base), d_unit, ref_atom[s_from:stop + 1, j], ref_pos[s_from:stop + 1, j],
                        ref_val[s_from:stop + 1, j])
        remaining = steps - (stop + 1)
        if remaining == 0:
            if ref_residual is not None:
                _check_energy(synced, ref_residual[j], rtol, j)
                rep.energy_checks += 1
            continue
        a2, p2, v2, r2 = run(synced[None, :], remaining)
        nxt = walk(j, stop + 1, a2[0], p2[0], v2[0])
        if nxt < steps:
            pending.append((j, nxt, _Started(synced, stop + 1)))
        elif ref_residual is not None:
            _check_energy(r2[0], ref_residual[j], rtol, j)
            rep.energy_checks += 1
    assert rep.checked == rep.eligible, rep
    return rep


class _Started(np.ndarray):
    """A residual that remembers which step of the reference sequence it is the state before."""

    def __new__(cls, arr, start):
        obj = np.asarray(arr, dtype=np.float32).view(cls)
        obj._start = start
        return obj

    def __array_finalize__(self, obj):
        self._start = getattr(obj, "_start", 0)


def _check_energy(residual, ref_residual, rtol, j):
    e_ref = float((np.asarray(ref_residual, dtype=np.float64) ** 2).sum())
    e_new = float((np.asarray(residual, dtype=np.float64) ** 2).sum())
    assert abs(e_new - e_ref) <= rtol * max(e_ref, 1e-12), (j, e_new, e_ref)


def compare_trace(ref_atom, ref_pos, ref_val, ref_margin, ref_residual, atom, pos, val, residual,
                  margin=MARGIN, rtol=RTOL):
    """One-shot form for results that are already computed (no restart possible): all step arrays are (S, B),
    values SIGNED; residuals (B, ..., N).  A signal's sequence comparison ends at the first ambiguous step that
    was resolved differently; its residual energy is then not comparable.  Returns the number of (signal, step)
    pairs whose (atom, position) had to match and did.  Prefer :func:`compare_with_resync`."""
    steps, batch = ref_atom.shape
    checked = 0
    for j in range(batch):
        floor = ATOL_OF_PEAK * float(np.abs(ref_val[:, j]).max()) if steps else 0.0
        for s in range(steps):
            strict = bool(ref_margin[s, j] > margin) and abs(float(ref_val[s, j])) > 16.0 * floor
            got = (int(atom[s, j]), int(pos[s, j]))
            want = (int(ref_atom[s, j]), int(ref_pos[s, j]))
            if got != want:
                assert not strict, f"signal {j} step {s}: got {got}, want {want} (margin {ref_margin[s, j]:.3g})"
                break
            a = float(ref_val[s, j])
            assert abs(float(val[s, j]) - a) <= rtol * max(abs(a), 1e-12) + floor, (s, j, float(val[s, j]), a)
            checked += int(strict)
        else:
            _check_energy(np.asarray(residual[j]), np.asarray(ref_residual[j]), rtol, j)
    return checked


def compare_with_oracle_trace(tr, atom_bs, pos_bs, val_bs, residual, **kw):
    """``tr`` is an oracle Trace (step-major); atom_bs/pos_bs/val_bs are the
    library's (B, S) arrays."""
    return compare_trace(tr.atom.numpy(), tr.pos.numpy(), tr.val.numpy(), tr.margin.numpy(),
                         tr.residual.numpy(), np.asarray(atom_bs).T, np.asarray(pos_bs).T, np.asarray(val_bs).T,
                         np.asarray(residual), **kw)


def resync_against_trace(run, signal, tr, **kw):
    """:func:`compare_with_resync` against an oracle Trace (``signal`` (B,1,N) or (B,N))."""
    sig = np.asarray(signal, dtype=np.float32)
    b = sig.shape[0]
    return compare_with_resync(run, sig.reshape(b, -1), tr.d_unit.numpy(), tr.atom.numpy(), tr.pos.numpy(),
                               tr.val.numpy(), tr.margin.numpy(), tr.residual.numpy().reshape(b, -1), **kw)
