#!/usr/bin/env python
"""Rounding margin for the projected matcher bound, emulated in numpy float32 with the operation order
a kernel would use: c_i = sum_k (a_k - mu_k) * U_ki accumulated sequentially in fp32 (no FMA), r = sqrt(max(sum (a_k -
mu_k)^2 - sum c_i^2, 0)), lb = (((c0a-c0b)^2 + (c1a-c1b)^2) + (c2a-c2b)^2) + (ra-rb)^2, against the refine's fp32 distance
(sum of (a_k - b_k)^2 in slot order).  Reports max over sampled pairs of sqrt(lb) - sqrt(d2): the filter is exact iff
the acceptance test  sqrt(lb) <= sqrt(best) * (1 + rel) + abs  covers that excess.   Run from the repo root."""
import sys

import numpy as np

sys.path.insert(0, "fast-lio-sam-qn_b200")
sys.path.insert(0, ".")
from b200reg import synth  # noqa: E402
from oracle import oracle  # noqa: E402

F32 = np.float32


def feats(Fd, mu, U):
    mu, U = mu.astype(F32), U.astype(F32)
    n = len(Fd)
    c = np.zeros((n, 3), F32)
    s2 = np.zeros(n, F32)
    for k in range(33):
        x = (Fd[:, k] - mu[k]).astype(F32)
        s2 = (s2 + (x * x).astype(F32)).astype(F32)
        for i in range(3):
            c[:, i] = (c[:, i] + (x * U[k, i]).astype(F32)).astype(F32)
    cc = ((c[:, 0] * c[:, 0]).astype(F32) + (c[:, 1] * c[:, 1]).astype(F32)).astype(F32)
    cc = (cc + (c[:, 2] * c[:, 2]).astype(F32)).astype(F32)
    r = np.sqrt(np.maximum((s2 - cc).astype(F32), F32(0))).astype(F32)
    return c, r


def main():
    b = np.load("profiles/fpfh_pca_basis.npz")
    worst = 0.0
    for seed, vox in ((2000, 0.3), (2005, 0.2), (2003, 0.4)):
        s, d, _ = synth.make_pair(seed, 100000, 100000, mode="quatro", voxel=vox)
        _, _, fs = oracle.fpfh(s)
        _, _, fd = oracle.fpfh(d)
        A = fs[(fs != 0).any(1)][:6000]
        B = fd[(fd != 0).any(1)][:8000]
        ca, ra = feats(A, b["mu"], b["U"])
        cb, rb = feats(B, b["mu"], b["U"])
        # fp32 refine distance in slot order
        d2 = np.zeros((len(A), len(B)), F32)
        for k in range(33):
            e = (A[:, k][:, None] - B[:, k][None, :]).astype(F32)
            d2 = (d2 + (e * e).astype(F32)).astype(F32)
        lb = np.zeros_like(d2)
        for i in range(3):
            e = (ca[:, i][:, None] - cb[:, i][None, :]).astype(F32)
            lb = (lb + (e * e).astype(F32)).astype(F32)
        e = (ra[:, None] - rb[None, :]).astype(F32)
        lb = (lb + (e * e).astype(F32)).astype(F32)
        excess = np.sqrt(lb.astype(np.float64)) - np.sqrt(d2.astype(np.float64))
        near = d2 < 1e-2
        print("pair %d/%.1f: %d x %d pairs; max sqrt(lb)-sqrt(d2) = %.3e overall, %.3e among d2 < 1e-2 (%d pairs); "
              "mean pruning lb/d2 = %.3f" % (seed, vox, len(A), len(B), excess.max(), excess[near].max() if near.any() else 0, near.sum(),
                                             float((lb.astype(np.float64) / np.maximum(d2, 1e-30))[d2 > 1].mean())))
        worst = max(worst, excess.max())
    print("worst excess %.3e  -> acceptance test sqrt(lb) <= sqrt(best) * 1.00001 + %.1e has a %.0fx margin" % (worst, 2e-3, 2e-3 / worst))


if __name__ == "__main__":
    main()
