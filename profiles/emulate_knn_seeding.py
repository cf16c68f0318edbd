#!/usr/bin/env python
"""CPU emulation of the 15-NN pass (k_covariance / knn.cuh) to study SEEDING strategies without a GPU.

Builds the same kind of tree as csrc/index_build.cu (30-bit Morton order, Karras prefix splits, sub-trees of <= 8 points
collapsed into leaves), walks it per query exactly like knn_search (near child first, far child stacked with its bound,
strict '>' pruning) and counts, per query: boxes tested, leaves visited, and candidates ACCEPTED into the result set after
seeding (each costs one K-long select chain on the GPU; the SASS profile attributes 47% of k_covariance's warp instructions
to that chain at 8.5 active lanes).

    python profiles/emulate_knn_seeding.py [n_points] [n_queries]
"""
import sys

import numpy as np

sys.path.insert(0, "fast-lio-sam-qn_b200")
K, LEAF = 15, 8


def expand10(v):
    v = v.astype(np.uint32)
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def build(pts):
    lo = pts.min(0)
    ext = (pts.max(0) - lo).max()
    q = np.clip(((pts - lo) * (1023.0 / ext)).astype(np.int64), 0, 1023)
    code = (expand10(q[:, 2]) << 2) | (expand10(q[:, 1]) << 1) | expand10(q[:, 0])
    order = np.argsort(code, kind="stable")
    P, code = pts[order], code[order].astype(np.int64)
    n = len(P)
    # key with index tie-break, as the kernel's delta()
    key = (code << 20) | np.arange(n, dtype=np.int64)
    nodes = []  # (lo, hi, child0, child1); child >= 0 internal, < 0 leaf -1-((start<<4)|count)

    def rec(a, b):  # [a, b] inclusive
        if b - a + 1 <= LEAF:
            return -1 - ((a << 4) | (b - a + 1)), P[a:b + 1].min(0), P[a:b + 1].max(0)
        diff = key[a] ^ key[b]
        bit = int(diff).bit_length() - 1
        # first index in [a, b] whose bit is set
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


def box_d2(q, lo, hi):
    d = np.maximum(np.maximum(lo - q, q - hi), 0.0)
    return float(d @ d)


def search(P, nodes, root, q, seeds, skip):
    """seeds: candidate positions inserted first (convergent on the GPU); returns counters and the result positions."""
    best = []  # sorted list of (d2, pos)

    def insert(pos):
        d2 = float(((P[pos] - q) ** 2).sum())
        if len(best) < K:
            best.append((d2, pos))
            best.sort()
            return True
        if (d2, pos) < best[-1]:
            best[-1] = (d2, pos)
            best.sort()
            return True
        return False

    seen = set()
    for s in seeds:
        if s not in seen:
            seen.add(s)
            insert(s)
    worst = lambda: best[-1][0] if len(best) == K else np.inf
    boxes = leaves = accepted = 0
    stack = []
    ref, dnode = root, 0.0
    while True:
        alive = not (dnode > worst())
        while alive and ref >= 0:
            lo0, hi0, r0, lo1, hi1, r1 = nodes[ref]
            d0, d1 = box_d2(q, lo0, hi0), box_d2(q, lo1, hi1)
            boxes += 2
            if d1 < d0:
                r0, r1, d0, d1 = r1, r0, d1, d0
            if not (d1 > worst()):
                stack.append((r1, d1))
            ref, dnode = r0, d0
            alive = not (d0 > worst())
        if alive:
            c = -1 - ref
            a, cnt = c >> 4, c & 15
            leaves += 1
            for pos in range(a, a + cnt):
                if pos in seen or pos in skip:
                    continue
                if insert(pos):
                    accepted += 1
        found = False
        while stack:
            ref, dnode = stack.pop()
            if not (dnode > worst()):
                found = True
                break
        if not found:
            break
    return boxes, leaves, accepted, [p for _, p in best]


def main():
    from b200reg import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    src, _, _ = synth.make_pair(1000, n, n)
    P, nodes, root = build(src[:, :3].astype(np.float64))
    n = len(P)
    rng = np.random.default_rng(0)
    starts = rng.integers(0, n - 64, nq // 32)
    qs = np.concatenate([np.arange(s, s + 32) for s in starts])  # whole warps of Morton-consecutive queries

    def window(i, w):
        lo = max(0, i - w // 2)
        hi = min(n - 1, lo + w - 1)
        lo = max(0, hi - (w - 1))
        return list(range(lo, hi + 1))

    results = {}
    strategies = {
        "morton window 15 (current)": lambda i: window(i, 15),
        "morton window 31": lambda i: window(i, 31),
        "morton window 63": lambda i: window(i, 63),
    }
    for name, fn in strategies.items():
        stats = np.array([search(P, nodes, root, P[i], fn(i), set())[:3] for i in qs])
        results[name] = stats.mean(0)
    # two-pass: every 8th point searched first (window seeds), the others seeded with the result set of the nearest
    # (in Morton order) pass-1 point plus their own window of 15
    coarse = {}
    for i in sorted(set((qs // 8) * 8) | set(np.minimum((qs // 8) * 8 + 8, n - 1))):
        coarse[i] = search(P, nodes, root, P[i], window(i, 15), set())[3]
    stats = []
    for i in qs:
        c = (i // 8) * 8
        c2 = min(c + 8, n - 1)
        cn = c if (i - c) <= (c2 - i) else c2
        stats.append(search(P, nodes, root, P[i], window(i, 15) + coarse[cn], set())[:3])
    results["two-pass: window 15 + 15-NN of the nearest 1-in-8 point"] = np.array(stats).mean(0)
    print("%d points, %d queries (whole warps); per query: boxes tested / leaves visited / candidates accepted after seeding" % (n, len(qs)))
    for name, s in results.items():
        print("  %-58s %6.1f %6.1f %6.1f" % (name, s[0], s[1], s[2]))


if __name__ == "__main__":
    main()
