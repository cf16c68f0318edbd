#!/usr/bin/env python
"""CPU emulation of a WARP-SHARED 15-NN walk (one traversal per 32 Morton-consecutive queries) to size the design
before writing the kernel: node visits, leaf visits, points tested per lane, candidates collected per lane (stale
per-lane bounds, refreshed at every flush), flush rounds -- against the per-thread walk of emulate_knn_seeding.py.

    python profiles/emulate_knn_warp_shared.py [n_points] [n_warps] [leaf] [buf]
"""
import sys

import numpy as np

sys.path.insert(0, "fast-lio-sam-qn_b200")
sys.path.insert(0, "profiles")
import emulate_knn_seeding as base  # noqa: E402

K = 15


def box_d2_many(Q, lo, hi):
    d = np.maximum(np.maximum(lo - Q, Q - hi), 0.0)
    return (d * d).sum(1)


def warp_walk(P, nodes, root, q0, leafcap, buf, seed_window=True):
    n = len(P)
    lanes = np.arange(q0, min(q0 + 32, n))
    Q = P[lanes]
    L = len(lanes)
    best = [[] for _ in range(L)]  # sorted (d2, pos)
    bound = np.full(L, np.inf)
    win = []
    for li, i in enumerate(lanes):
        lo = max(0, i - K // 2)
        hi = min(n - 1, lo + K - 1)
        lo = max(0, hi - (K - 1))
        win.append((lo, hi))
        if seed_window:
            d = ((P[lo:hi + 1] - Q[li]) ** 2).sum(1)
            best[li] = sorted(zip(d.tolist(), range(lo, hi + 1)))[:K]
            if len(best[li]) == K:
                bound[li] = best[li][-1][0]
    pend = [[] for _ in range(L)]
    st = dict(nodes=0, leaves=0, pts=0, cand=0, rounds=0, round_slots=0, inserted=0)

    def flush():
        most = max(len(p) for p in pend)
        if most == 0:
            return
        st["rounds"] += 1
        st["round_slots"] += most
        for li in range(L):
            for d, pos in pend[li]:
                if len(best[li]) < K or (d, pos) < best[li][-1]:
                    best[li].append((d, pos))
                    best[li].sort()
                    del best[li][K:]
                    st["inserted"] += 1
            pend[li] = []
            if len(best[li]) == K:
                bound[li] = best[li][-1][0]

    stack = []
    ref = root
    dlane = np.zeros(L)
    while True:
        alive = bool((dlane <= bound).any())
        while alive and ref >= 0:
            lo0, hi0, r0, lo1, hi1, r1 = nodes[ref]
            st["nodes"] += 1
            d0, d1 = box_d2_many(Q, lo0, hi0), box_d2_many(Q, lo1, hi1)
            a0, a1 = bool((d0 <= bound).any()), bool((d1 <= bound).any())
            if d1.min() < d0.min():
                r0, r1, d0, d1, a0, a1 = r1, r0, d1, d0, a1, a0
            if a1:
                stack.append((r1, d1))
            ref, dlane, alive = r0, d0, a0
        if alive:
            c = -1 - ref
            a, cnt = c >> 4, c & 15 if leafcap <= 15 else None
            if leafcap > 15:
                a, cnt = c >> 8, c & 255
            st["leaves"] += 1
            st["pts"] += cnt
            for pos in range(a, a + cnt):
                d = ((Q - P[pos]) ** 2).sum(1)
                for li in range(L):
                    if d[li] <= bound[li] and not (seed_window and win[li][0] <= pos <= win[li][1]):
                        pend[li].append((float(d[li]), pos))
                        st["cand"] += 1
            if max(len(p) for p in pend) > buf - leafcap:
                flush()
        found = False
        while stack:
            ref, dlane = stack.pop()
            if (dlane <= bound).any():
                found = True
                break
        if not found:
            break
    flush()
    return st, [[p for _, p in b] for b in best]


def build_leaf(pts, leafcap):
    base.LEAF = leafcap
    if leafcap <= 15:
        return base.build(pts)
    # wider leaves: re-encode refs as (start << 8) | count
    P, nodes, root = None, None, None
    lo = pts.min(0)
    ext = (pts.max(0) - lo).max()
    q = np.clip(((pts - lo) * (1023.0 / ext)).astype(np.int64), 0, 1023)
    code = (base.expand10(q[:, 2]) << 2) | (base.expand10(q[:, 1]) << 1) | base.expand10(q[:, 0])
    order = np.argsort(code, kind="stable")
    P, code = pts[order], code[order].astype(np.int64)
    n = len(P)
    key = (code << 20) | np.arange(n, dtype=np.int64)
    nodes = []

    def rec(a, b):
        if b - a + 1 <= leafcap:
            return -1 - ((a << 8) | (b - a + 1)), P[a:b + 1].min(0), P[a:b + 1].max(0)
        bit = int(key[a] ^ key[b]).bit_length() - 1
        s = a + int(np.searchsorted((key[a:b + 1] >> bit) & 1, 1))
        idx = len(nodes)
        nodes.append(None)
        r0, lo0, hi0 = rec(a, s - 1)
        r1, lo1, hi1 = rec(s, b)
        nodes[idx] = (lo0, hi0, r0, lo1, hi1, r1)
        return idx, np.minimum(lo0, lo1), np.maximum(hi0, hi1)
    sys.setrecursionlimit(10000)
    root, _, _ = rec(0, n - 1)
    return P, nodes, root


def main():
    from b200reg import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    nw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    leafcap = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    buf = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    src, _, _ = synth.make_pair(1000, n, n)
    P, nodes, root = build_leaf(src[:, :3].astype(np.float64), leafcap)
    rng = np.random.default_rng(0)
    starts = (rng.integers(0, len(P) // 32 - 1, nw) * 32)
    tot = {}
    exact = True
    for s in starts:
        st, res = warp_walk(P, nodes, root, int(s), leafcap, buf)
        for k, v in st.items():
            tot[k] = tot.get(k, 0) + v
        for li in range(0, 32, 11):  # spot check against brute force
            d = ((P - P[s + li]) ** 2).sum(1)
            want = set(np.lexsort((np.arange(len(P)), d))[:K].tolist())
            exact &= (set(res[li]) == want)
    print("leaf<=%d buf=%d: per WARP: node visits %.1f, leaves %.1f, points scanned %.1f, flush rounds %.1f (slots %.1f); "
          "per LANE: candidates collected %.1f, inserted %.1f; exact=%s" %
          (leafcap, buf, tot["nodes"] / nw, tot["leaves"] / nw, tot["pts"] / nw, tot["rounds"] / nw, tot["round_slots"] / nw,
           tot["cand"] / nw / 32, tot["inserted"] / nw / 32, exact))


if __name__ == "__main__":
    main()
