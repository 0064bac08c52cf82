"""Shared parity helpers (BASELINE.json north_star rule): identical
(atom, position) wherever the oracle's top-2 relative margin exceeds 1e-5;
amplitudes and residual energy within 1e-4 relative.  After a legitimately
ambiguous step the sequences may diverge, so the sequence comparison of that
signal stops there (its residual energy is then not comparable either).  The
comparison also stops once the oracle's own amplitude has fallen to fp32
round-off of the signal scale (see ATOL_OF_PEAK): tiny signals are explained
exactly after a few atoms and what is picked afterwards is noise."""
import numpy as np

MARGIN = 1e-5
RTOL = 1e-4
# Absolute floor of the amplitude check, as a fraction of the signal's LARGEST amplitude: once a signal is
# explained down to fp32 round-off (a 4-sample signal after 8 atoms: amplitudes of 4e-6 against 1.0), the
# reference's own correlation carries an absolute error of about 1e-7 of the signal scale, so a purely relative
# tolerance on such a value compares noise with noise.
ATOL_OF_PEAK = 1e-6


def compare_trace(ref_atom, ref_pos, ref_absval, ref_margin, ref_residual, atom, pos, val, residual,
                  margin=MARGIN, rtol=RTOL):
    """All step arrays are (S, B); residuals (B, ..., N).  Returns the number of
    (signal, step) pairs that were compared exactly."""
    steps, batch = ref_atom.shape
    checked = 0
    for j in range(batch):
        floor = ATOL_OF_PEAK * float(np.abs(ref_absval[:, j]).max()) if steps else 0.0
        for s in range(steps):
            if not ref_margin[s, j] > margin:
                break
            if float(ref_absval[s, j]) <= 16.0 * floor:
                break       # explained down to round-off: from here on every implementation picks among noise
            got = (int(atom[s, j]), int(pos[s, j]))
            want = (int(ref_atom[s, j]), int(ref_pos[s, j]))
            assert got == want, f"signal {j} step {s}: got {got}, want {want} (margin {ref_margin[s, j]:.3g})"
            a = float(ref_absval[s, j])
            assert abs(abs(float(val[s, j])) - a) <= rtol * max(a, 1e-12) + floor, (s, j, float(val[s, j]), a)
            checked += 1
        else:
            e_ref = float((np.asarray(ref_residual[j], dtype=np.float64) ** 2).sum())
            e_new = float((np.asarray(residual[j], dtype=np.float64) ** 2).sum())
            assert abs(e_new - e_ref) <= rtol * max(e_ref, 1e-12), (j, e_new, e_ref)
    return checked


def compare_with_oracle_trace(tr, atom_bs, pos_bs, val_bs, residual, **kw):
    """``tr`` is an oracle Trace (step-major); atom_bs/pos_bs/val_bs are the
    library's (B, S) arrays."""
    return compare_trace(tr.atom.numpy(), tr.pos.numpy(), tr.val.abs().numpy(), tr.margin.numpy(),
                         tr.residual.numpy(), np.asarray(atom_bs).T, np.asarray(pos_bs).T, np.asarray(val_bs).T,
                         np.asarray(residual), **kw)
