#!/usr/bin/env python
"""Groundwork for the next matcher step: a FIXED 3-direction orthonormal basis + mean for the
projected lower bound  |a-b|^2 >= sum_i (u_i.(a-b))^2 + (|r_a| - |r_b|)^2  (r = residual of a - mu outside span(u)).
Any orthonormal basis gives a true bound; this one is the PCA of oracle FPFH descriptors of one voxelised KITTI-shaped pair
and generalises to other pairs / voxel sizes (survivors of the final-bound test, fraction of all pairs):

    pair (seed, voxel)   true    fixed basis   own basis   block norms (today)
    2000, 0.3 (train)    2.56%   2.96%         2.96%       4.86%
    2005, 0.2            1.22%   1.55%         1.54%       3.76%
    2010, 0.3 (30k)      1.02%   1.53%         1.52%       4.41%
    2003, 0.4            2.31%   2.94%         2.92%       5.82%

Writes profiles/fpfh_pca_basis.npz (mu[33], U[33,3]).  Run from the repo root.
"""
import sys

import numpy as np

sys.path.insert(0, "fast-lio-sam-qn_b200")
sys.path.insert(0, ".")
from b200reg import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    s, d, _ = synth.make_pair(2000, 100000, 100000, mode="quatro", voxel=0.3)
    _, _, fs = oracle.fpfh(s)
    _, _, fd = oracle.fpfh(d)
    X = np.concatenate([fs, fd]).astype(np.float64)
    X = X[(X != 0).any(1)]
    mu = X.mean(0)
    w, V = np.linalg.eigh(np.cov(X.T))
    U = V[:, ::-1][:, :3]
    U, _ = np.linalg.qr(U)  # exactly orthonormal columns
    np.savez(os.path.join("profiles", "fpfh_pca_basis.npz"), mu=mu, U=U, explained=w[::-1][:3] / w.sum())
    print("explained variance of the 3 directions:", w[::-1][:3] / w.sum())


if __name__ == "__main__":
    import os
    main()
