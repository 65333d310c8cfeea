"""Adversarial inputs for the exactness guard of the resampler kernels (test infrastructure).

The reference truncates ``gain * sum`` toward zero (libllzfilter/llz_resample.c:590-601), so an output whose sum lies
within a few ulps of a non-zero integer is where two evaluation orders can disagree.  ``tune_output`` moves three
samples of one output's window until the reference-order sum is a few 1e-12 away from an integer -- on the requested
side of it -- by a meet-in-the-middle search over two of the samples (65536^2 candidates per value of the third).
"""
from __future__ import annotations

import numpy as np


def terms_resample(plan, o: int):
    """(coefficient, input index) pairs of output ``o`` in the reference's accumulation order (llz_resample.c:586-592)"""
    row = plan.bank[o % plan.L]
    base = (o * plan.M) // plan.L
    return [(float(row[k]), base - k) for k in range(plan.cols)]


def terms_decimate(plan, i: int):
    """llz_resample.c:467-473: m outer, k inner"""
    out = []
    for m in range(plan.M):
        row = plan.bank[m]
        for k in range(plan.cols):
            out.append((float(row[k]), i * plan.M + m + plan.M * k - plan.n))
    return out


def reference_sum(terms, x, gain: float) -> float:
    """the reference's own evaluation: products and sums rounded separately, in order, then the gain"""
    acc = 0.0
    n = len(x)
    for c, s in terms:
        xv = float(x[s]) if 0 <= s < n else 0.0
        acc = acc + xv * c
    return acc * gain


def tune_output(rng, terms, x, gain: float, side: int, tries: int = 48, amp: int = 30000):
    """Rewrite three samples of ``x`` (in place) so that reference_sum(terms) is as close as possible to a non-zero
    integer, at or above it (side >= 0) or just below it (side < 0).  Returns (value, distance to the integer)."""
    n = len(x)
    live = [(abs(c), c, s) for c, s in terms if 0 <= s < n and c != 0.0]
    live.sort(reverse=True)
    if len(live) < 3:
        return None
    vals = np.arange(-amp, amp + 1, dtype=np.float64)

    def max_gap(coef):
        """largest hole in the fractional parts coef*x can supply: a tap near a small rational (1/3, 1 - eps) supplies few"""
        f = np.sort(np.mod(coef * gain * vals, 1.0))
        return max(float(np.diff(f).max()), float(f[0] + 1.0 - f[-1]))

    # distinct magnitudes only: a symmetric prototype has its taps in equal pairs, and c*(xa + xb) is one sample's worth
    distinct = []
    for t in live[:12]:
        if all(abs(t[0] - u[0]) > 1e-9 * u[0] for u in distinct):
            distinct.append(t)
    cand = sorted(distinct, key=lambda t: max_gap(t[1]))
    if len(cand) < 3 or max_gap(cand[1][1]) > 1e-3:
        return None
    (_, ca, sa), (_, cb, sb), (_, cc, sc) = cand[0], cand[1], cand[2]
    fa = np.mod(ca * gain * vals, 1.0)
    order = np.argsort(fa)
    fa_sorted = fa[order]
    best = None
    for _ in range(tries):
        x[sc] = int(rng.integers(-amp, amp))
        x[sa] = 0
        x[sb] = 0
        rest = reference_sum(terms, x, gain)
        need = np.mod(-(rest + cb * gain * vals), 1.0)            # fractional part the a-term has to supply
        idx = np.searchsorted(fa_sorted, need)
        for cand in (idx % len(vals), (idx - 1) % len(vals)):
            d = np.abs(fa_sorted[cand] - need)
            d = np.minimum(d, 1.0 - d)
            j = int(np.argmin(d))
            if d[j] > 5e-10:
                continue
            xa, xb = int(vals[order[cand[j]]]), int(vals[j])
            x[sa], x[sb] = xa, xb
            v = reference_sum(terms, x, gain)
            nint = round(v)
            dist = v - nint
            if nint == 0 or abs(v) > 32000:
                continue
            ok_side = dist >= 0 if side >= 0 else dist < 0
            if ok_side and (best is None or abs(dist) < abs(best[1])):
                best = (v, dist, xa, xb, int(x[sc]))
    if best is None:
        return None
    x[sa], x[sb], x[sc] = best[2], best[3], best[4]
    return best[0], best[1]
