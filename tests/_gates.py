"""Shared parity gates of the GPU tests."""
import numpy as np


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def interaction_gate(ours, ref32, ref64, label=""):
    """Interactions I^(m)(i,j) = v[4k] + v[4k+3] - v[4k+1] - v[4k+2] are differences of four rewards of the size of the
    logits, so their fp32 rounding noise is set by the logit scale, not by |I|.  The gate is the north star's
    1e-3 * max|I^(m)| per order, or -- where the reference's OWN fp32 run sits further than that from a float64
    evaluation of the same network on the same masked clouds (tests/golden/interactions_f64.npz) -- twice that measured
    noise.  Prints all the numbers; returns (error, bound)."""
    ours, ref32, ref64 = (np.asarray(x, np.float64) for x in (ours, ref32, ref64))
    scale = np.abs(ref32).max()
    noise = np.abs(ref32 - ref64).max()
    err = np.abs(ours - ref32).max()
    err64 = np.abs(ours - ref64).max()
    bound = max(1e-3 * scale, 2.0 * noise)
    print("%s max|I| %.3e | ours-ref32 %.2e (%.1e of max|I|) | ours-f64 %.2e | ref32-f64 (reference's own noise) %.2e | "
          "bound %.2e (%s)" % (label, scale, err, err / max(scale, 1e-30), err64, noise, bound,
                               "1e-3 max|I|" if bound == 1e-3 * scale else "2 x reference noise"))
    return err, bound


def flip_audit(ours, ref, f64_of, label, tol=1e-3, flip_cap=6e-3, rate_cap=0.003):
    """Audit of a batch of DGCNN logits against the reference's fp32 logits of the same masked clouds.

    DGCNN rebuilds its kNN graph in feature space, so a near-tie between the k-th and (k+1)-th neighbour is decided by
    fp32 rounding; in a masked cloud the coincident points flip together and ONE cloud's logits jump by up to a few 1e-3
    of scale between two correct fp32 evaluations.  Every cloud further than 1e-4 from the reference is therefore
    re-evaluated in float64 (f64_of(indices) -> logits; oracle, same masked input) and must be explained: beyond `tol`
    from the reference is only accepted when one of the two fp32 runs sits within 1e-4 of float64 (the other one
    flipped), no flip may exceed `flip_cap`, and OUR flips must not be more frequent than the reference's own
    (or than `rate_cap` of the batch).  Returns the indices of the clouds beyond 1e-4 (for callers that gate derived
    quantities on the untouched clouds only)."""
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max()
    per = np.abs(ours - ref).max(1) / scale
    out = np.nonzero(per > 1e-4)[0]
    print("%s: %d clouds, vs reference: max %.2e median %.2e, above 1e-4: %d, above %.0e: %d"
          % (label, len(per), per.max(), np.median(per), len(out), tol, int((per > tol).sum())))
    assert np.median(per) <= 1e-5, "typical clouds must sit at fp32 noise"
    assert len(out) <= 0.2 * len(per), "too many clouds away from the reference"
    if len(out) == 0:
        return out
    f64 = np.asarray(f64_of(out), np.float64)
    o64 = np.abs(ours[out] - f64).max(1) / scale
    r64 = np.abs(ref[out] - f64).max(1) / scale
    ours_flips, ref_flips = int((o64 > tol).sum()), int((r64 > tol).sum())
    unexplained = 0
    for i, c in enumerate(out):
        if per[c] > tol:
            ok = o64[i] <= 1e-4 or r64[i] <= 1e-4
            unexplained += not ok
            print("  cloud %5d: ours-ref %.2e | ours-f64 %.2e | ref-f64 %.2e  %s" % (
                c, per[c], o64[i], r64[i], "reference flipped" if o64[i] <= 1e-4 else ("we flipped" if ok else "UNEXPLAINED")))
    print("%s: float64 audit of %d clouds: ours-f64 max %.2e median %.2e | reference-f64 max %.2e median %.2e | flips beyond "
          "%.0e: ours %d, reference %d" % (label, len(out), o64.max(), np.median(o64), r64.max(), np.median(r64), tol,
                                          ours_flips, ref_flips))
    assert unexplained == 0
    assert max(o64.max(), r64.max()) <= flip_cap
    assert ours_flips <= max(ref_flips, int(np.ceil(rate_cap * len(per))))
    # and below the flip size: we may not sit further from float64 than the reference's own fp32 run does
    ours_off, ref_off = int((o64 > 1e-4).sum()), int((r64 > 1e-4).sum())
    print("%s: clouds further than 1e-4 from float64: ours %d, reference %d" % (label, ours_off, ref_off))
    assert ours_off <= max(ref_off, int(np.ceil(rate_cap * len(per))))
    return out
