#!/usr/bin/env python
"""Streaming replay of the 33-D matcher (k_feat_nn) on oracle descriptors with (a) bitwise-duplicate base descriptors
collapsed per 64-record tile (run head keeps the run's lowest original index: exact under the lowest-index tie rule) and
(b) the projected bound (3 fixed PCA directions + residual norm) instead of the block norms.  Counts what the kernel pays
for: refines, record tests, per-query tile visits, per-block tile loads.  Run from the repo root."""
import os
import sys
import zlib
sys.path.insert(0, 'fast-lio-sam-qn_b200'); sys.path.insert(0, '.')
import numpy as np
from b200reg import synth
from oracle import oracle

seed, voxel = (int(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (2000, 0.3)
s, d, T = synth.make_pair(seed, 100000, 100000, mode="quatro", voxel=voxel)
_, _, fs = oracle.fpfh(s); _, _, fd = oracle.fpfh(d)
bs = np.load("profiles/fpfh_pca_basis.npz"); mu, U = bs["mu"], bs["U"]


def proj(F):
    X = F.astype(np.float64) - mu; P = X @ U
    R = X - P @ U.T
    return np.concatenate([P, np.sqrt((R * R).sum(1))[:, None]], 1).astype(np.float32)


def blocknorm(F):
    return np.sqrt((F.astype(np.float64).reshape(-1, 3, 11) ** 2).sum(2)).astype(np.float32)


def morton(q, dims, bits):
    c = np.zeros(len(q), np.uint64)
    for dmn in range(dims):
        for b in range(bits):
            c |= ((q[:, dmn] >> b) & 1).astype(np.uint64) << np.uint64(dims * b + dmn)
    return c


SCALE = np.array([60.0, 40.0, 30.0, 60.0])  # half-ranges of the fixed quantiser (P1..P3 centred, r from 0)


def code_proj(N):
    x = N.astype(np.float64).copy()
    x[:, :3] = (x[:, :3] + SCALE[:3]) / (2 * SCALE[:3])
    x[:, 3] = x[:, 3] / SCALE[3]
    q = np.clip((x * 128).astype(np.int64), 0, 127)
    return morton(q, 4, 7)


def code_norm(N):
    q = np.minimum(1023, (N * 10.23).astype(np.int64))
    return morton(q, 3, 10)


def hash32(F):
    return np.array([zlib.crc32(r.tobytes()) for r in F], np.uint64)


def prep(F, feat, code, hbits):
    ok = (F != 0).any(1); F = F[ok]; orig = np.nonzero(ok)[0]
    N = feat(F); c = code(N)
    key = (c << np.uint64(hbits)) | (hash32(F) & np.uint64((1 << hbits) - 1)) if hbits else c
    o = np.argsort(key, kind='stable')
    return F[o], N[o], key[o], orig[o]


def run(name, feat, code, hbits, dedup, sqrt_margin, seed_init=False):
    Q, QN, qc, qo = prep(fs, feat, code, hbits); B, BN, bc, bo = prep(fd, feat, code, hbits)
    TILE = int(os.environ.get("TILE", "64")); thr2 = np.float32(35.0 ** 2)
    nt = (len(B) + TILE - 1) // TILE
    usable = np.ones(len(B), bool)
    if dedup:
        same = np.zeros(len(B), bool)
        same[1:] = (B[1:] == B[:-1]).all(1)
        same[::TILE] = False  # runs do not cross tiles
        usable = ~same
    big = np.float32(1e30)
    bmin = np.array([np.where(usable[t*TILE:(t+1)*TILE, None], BN[t*TILE:(t+1)*TILE], big).min(0) for t in range(nt)])
    bmax = np.array([np.where(usable[t*TILE:(t+1)*TILE, None], BN[t*TILE:(t+1)*TILE], -big).max(0) for t in range(nt)])
    best = np.full(len(Q), thr2, np.float32)
    if seed_init:  # one refine per query up front: the base record where the query's own key would sit
        pos = np.clip(np.searchsorted(bc, qc), 0, len(B) - 1)
        d0 = ((Q.astype(np.float64) - B[pos].astype(np.float64)) ** 2).sum(1).astype(np.float32)
        best = np.minimum(best, d0)
    tests = refines = tile_visits = loads = 0
    nblk = (len(Q) + 127) // 128
    Bd = B.astype(np.float64); Qd = Q.astype(np.float64)
    for blk in range(nblk):
        q0 = blk * 128; q1 = min(len(Q), q0 + 128)
        mid = qc[min(len(Q) - 1, q0 + 64)]
        t0 = min(nt - 1, np.searchsorted(bc, mid) // TILE)
        order = list(range(t0, nt)) + list(range(t0 - 1, -1, -1))
        bq = best[q0:q1]; qn = QN[q0:q1]
        for t in order:
            bound = (np.sqrt(bq) * 1.00001 + 2e-3) ** 2 if sqrt_margin else bq * 1.0001 + 1e-3
            e = np.maximum(np.maximum(bmin[t] - qn, qn - bmax[t]), 0); lb = (e * e).sum(1)
            need = lb <= bound
            if not need.any(): continue
            loads += 1
            idx = np.nonzero(need)[0]; tile_visits += len(idx)
            sl = slice(t * TILE, (t + 1) * TILE)
            bn = BN[sl]; us = usable[sl]
            tests += len(idx) * len(bn)
            er = qn[idx][:, None, :] - bn[None, :, :]; lbr = (er * er).sum(2)
            pas = (lbr <= bound[idx][:, None]) & us[None, :]
            refines += pas.sum()
            D = ((Qd[q0:q1][idx][:, None, :] - Bd[sl][None, :, :]) ** 2).sum(2)
            D = np.where(pas, D, np.inf)
            bq[idx] = np.minimum(bq[idx], D.min(1).astype(np.float32))
        best[q0:q1] = bq
    tot = len(Q) * len(B)
    print("%-44s base %5d (usable %5d)  refines %5.2fM (%.2f%%)  record tests %5.1fM  tile visits/query %5.1f  tile loads/block %5.1f of %d"
          % (name, len(B), usable.sum(), refines / 1e6, 100 * refines / tot, tests / 1e6, tile_visits / len(Q), loads / nblk, nt))
    out = np.zeros(len(fs), np.float32); out[qo] = best
    return out


print("pair seed %d voxel %.1f: %d x %d descriptors" % (seed, voxel, (fs != 0).any(1).sum(), (fd != 0).any(1).sum()))
b0 = run("block norms, 3-D Morton (today)", blocknorm, code_norm, 0, False, False)
b1 = run("block norms + per-tile dedup (hash 2 bits)", blocknorm, code_norm, 2, True, False)
b2 = run("projected, 4-D Morton, no dedup", proj, code_proj, 0, False, True)
b3 = run("projected + per-tile dedup (hash 4 bits)", proj, code_proj, 4, True, True)
b4 = run("  ... + one seed refine per query", proj, code_proj, 4, True, True, seed_init=True)
assert np.array_equal(b0, b1) and np.array_equal(b0, b2) and np.array_equal(b0, b3), "best distances differ"
print("best distances identical in all four")
