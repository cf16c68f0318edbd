#!/usr/bin/env python
"""Could the LM loop's 1-NN searches after the first be skipped per point?  Replay of the oracle's pose sequence on 100k x
100k pairs: a point keeps its correspondence without a tree walk if  d1' < (sqrt(d2_prev) - delta)^2  where d1' is the new
distance to its previous nearest neighbour, d2_prev the previous SECOND-nearest squared distance and delta how far the
transformed point moved.  Prints the certified fraction per search and the fraction of 32-point warps / 128-point blocks in
which EVERY lane is certified (what a kernel without compaction would save).  Run from the repo root."""
import sys
sys.path.insert(0, 'fast-lio-sam-qn_b200'); sys.path.insert(0, '.')
import numpy as np
from scipy.spatial import cKDTree
from b200reg import synth
from oracle import oracle

for seed in (1000, 1003, 1010):
    src, dst, _ = synth.make_pair(seed, 100000, 100000)
    r = oracle.gicp_align(src, dst, want_trace=True)
    poses = [t[:16].reshape(4, 4) for t in r["trace"]] + [r["T"]]  # linearize poses, then the fitness pass at the final pose
    # device order of the source points: Morton order of the source cloud (warps hold spatial neighbours)
    q = np.clip(((src[:, :3] + 80.0) / 160.0 * 1024).astype(np.int64), 0, 1023)
    code = np.zeros(len(src), np.int64)
    for b in range(10):
        for d in range(3):
            code |= ((q[:, d] >> b) & 1) << (3 * b + d)
    order = np.argsort(code, kind="stable")
    P = src[order, :3].astype(np.float64)
    tree = cKDTree(dst[:, :3].astype(np.float64))
    prev = None
    print("seed %d: %d searches (%d linearize + fitness)" % (seed, len(poses), len(poses) - 1))
    for k, T in enumerate(poses):
        X = P @ T[:3, :3].T + T[:3, 3]
        d, idx = tree.query(X, k=2)
        if prev is not None:
            Xp, dp, ip = prev
            delta = np.linalg.norm(X - Xp, axis=1)
            d1n = np.linalg.norm(X - dst[ip[:, 0], :3], axis=1)
            ok = d1n < (dp[:, 1] - delta) * (1 - 1e-5) - 1e-6
            assert np.all(idx[ok, 0] == ip[ok, 0])
            w = ok[:len(ok) // 32 * 32].reshape(-1, 32).all(1).mean()
            b = ok[:len(ok) // 128 * 128].reshape(-1, 128).all(1).mean()
            print("  search %d: moved %.4f m median (max %.3f), certified %.1f%% of the points, %.1f%% of the warps, %.1f%% of the 128-blocks"
                  % (k, np.median(delta), delta.max(), 100 * ok.mean(), 100 * w, 100 * b))
        prev = (X, d, idx)
