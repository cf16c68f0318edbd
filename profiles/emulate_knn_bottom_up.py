#!/usr/bin/env python
"""CPU emulation: BOTTOM-UP exact k-NN on the LBVH (start at the leaf of a seed, climb; at each ancestor search the
sibling sub-tree top-down if its box can still hold a better point; stop as soon as the ball of the current worst
distance lies inside the ancestor's box) against the top-down walk of knn.cuh.  Counts box tests and leaves per query.

    python profiles/emulate_knn_bottom_up.py           (run from the repo root)
"""
import sys

import numpy as np

sys.path.insert(0, "profiles")
sys.path.insert(0, "fast-lio-sam-qn_b200")
import emulate_knn_seeding as E  # noqa: E402


def annotate(nodes, root, P):
    """parent / own box per node and per leaf; leaf lookup per point position."""
    n = len(P)
    parent, box, leaf_of = {}, {}, np.zeros(n, np.int64)

    def rec(ref, par):
        parent[ref] = par
        if ref < 0:
            c = -1 - ref
            a, cnt = c >> 4, c & 15
            leaf_of[a:a + cnt] = ref
            return
        lo0, hi0, r0, lo1, hi1, r1 = nodes[ref]
        box[r0], box[r1] = (lo0, hi0), (lo1, hi1)
        rec(r0, ref)
        rec(r1, ref)

    sys.setrecursionlimit(10000)
    box[root] = (P.min(0), P.max(0))
    rec(root, None)
    return parent, box, leaf_of


def bottom_up(P, nodes, root, parent, box, leaf_of, q, k, start_pos, seeds):
    best = []

    def insert(pos):
        d2 = float(((P[pos] - q) ** 2).sum())
        if len(best) < k:
            best.append((d2, pos)); best.sort(); return
        if (d2, pos) < best[-1]:
            best[-1] = (d2, pos); best.sort()

    seen = set()
    for s in seeds:
        if s not in seen:
            seen.add(s); insert(s)
    worst = lambda: best[-1][0] if len(best) == k else np.inf
    boxes = leaves = 0

    def leaf(ref):
        nonlocal leaves
        leaves += 1
        c = -1 - ref
        a, cnt = c >> 4, c & 15
        for pos in range(a, a + cnt):
            if pos not in seen:
                seen.add(pos); insert(pos)

    def top_down(ref, dnode):
        nonlocal boxes
        stack = [(ref, dnode)]
        while stack:
            ref, dnode = stack.pop()
            if dnode > worst():
                continue
            while ref >= 0:
                lo0, hi0, r0, lo1, hi1, r1 = nodes[ref]
                d0, d1 = E.box_d2(q, lo0, hi0), E.box_d2(q, lo1, hi1)
                boxes += 2
                if d1 < d0:
                    r0, r1, d0, d1 = r1, r0, d1, d0
                if not (d1 > worst()):
                    stack.append((r1, d1))
                ref, dnode = r0, d0
                if d0 > worst():
                    ref = None
                    break
            if ref is not None and ref < 0:
                leaf(ref)

    node = leaf_of[start_pos]
    leaf(node)
    while parent[node] is not None:
        lo, hi = box[node]
        boxes += 1  # ball-in-box test costs about one box test
        r = np.sqrt(worst())
        if np.all(q - r >= lo) and np.all(q + r <= hi):
            break
        par = parent[node]
        lo0, hi0, r0, lo1, hi1, r1 = nodes[par]
        sib, sb = (r1, (lo1, hi1)) if r0 == node else (r0, (lo0, hi0))
        d = E.box_d2(q, sb[0], sb[1])
        boxes += 1
        if not (d > worst()):
            if sib < 0:
                leaf(sib)
            else:
                top_down(sib, d)
        node = par
    return boxes, leaves, [p for _, p in best]


def main():
    from b200reg import synth
    src, dst, _ = synth.make_pair(1000, 100000, 100000)
    P, nodes, root = E.build(src[:, :3].astype(np.float64))
    parent, box, leaf_of = annotate(nodes, root, P)
    n = len(P)
    rng = np.random.default_rng(0)
    qs = np.concatenate([np.arange(s, s + 32) for s in rng.integers(0, n - 64, 24)])

    def window(i, w=15):
        lo = max(0, i - w // 2); hi = min(n - 1, lo + w - 1); lo = max(0, hi - (w - 1))
        return list(range(lo, hi + 1))

    E.K = 15
    td = np.array([E.search(P, nodes, root, P[i], window(i), set())[:2] for i in qs])
    bu, ok = [], True
    for i in qs:
        b, l, res = bottom_up(P, nodes, root, parent, box, leaf_of, P[i], 15, i, window(i))
        ok &= sorted(res) == sorted(E.search(P, nodes, root, P[i], window(i), set())[3])
        bu.append((b, l))
    print("15-NN of cloud points in their own cloud (Morton-window seeds): box tests / leaves per query")
    print("  top-down (knn.cuh today)   %6.1f %5.1f" % tuple(td.mean(0)))
    print("  bottom-up from own leaf    %6.1f %5.1f   identical results: %s" % (*np.mean(bu, 0), ok))
    # 1-NN of the (identity-guess) source points in the target, seeded by the previous query's answer
    T, tnodes, troot = E.build(dst[:, :3].astype(np.float64))
    tparent, tbox, tleaf = annotate(tnodes, troot, T)
    E.K = 1
    td1, bu1, ok = [], [], True
    for s in rng.integers(0, n - 64, 24):
        prev = None
        for i in range(s, s + 32):
            q = P[i]
            b, l, a, res = E.search(T, tnodes, troot, q, [prev] if prev is not None else [], set())
            td1.append((b, l))
            if prev is not None:
                b2, l2, res2 = bottom_up(T, tnodes, troot, tparent, tbox, tleaf, q, 1, prev, [prev])
                bu1.append((b2, l2))
                ok &= res2 == res
            prev = res[0]
    print("1-NN of source points in the target, seeded with the previous query's answer:")
    print("  top-down                   %6.1f %5.1f" % tuple(np.mean(td1, 0)))
    print("  bottom-up from seed's leaf %6.1f %5.1f   identical results: %s" % (*np.mean(bu1, 0), ok))


if __name__ == "__main__":
    main()
