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
