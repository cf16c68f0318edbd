#!/usr/bin/env python
"""Per-pair LM scheduling, measured: a 16-pair icpAlignment batch in which ONE pair runs all 32 outer iterations (synthetic
seed 1104: GICP itself diverges on it, identically on the CPU oracle) against the same batch with a clean pair in its place.
Round 1 kept launching the blocks of the 15 finished pairs until the last pair was done.
    python profiles/straggler.py > gpurun_out/straggler.txt"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "fast-lio-sam-qn_b200"))
import torch  # noqa: E402
import b200reg  # noqa: E402
from b200reg import synth  # noqa: E402

clean = [synth.make_pair(1000 + i, 100000, 100000) for i in range(16)]
bad = synth.make_pair(1104, 100000, 100000)
ctx = b200reg.Context(0)


def run(pairs, reps=8):
    ds = [torch.from_numpy(p[0]).cuda() for p in pairs]
    dd = [torch.from_numpy(p[1]).cuda() for p in pairs]
    args = ([t.data_ptr() for t in ds], [t.shape[0] for t in ds], [t.data_ptr() for t in dd], [t.shape[0] for t in dd], 16, 1)
    res = ctx.icp_alignment_ptrs(*args)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        res = ctx.icp_alignment_ptrs(*args)
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), [r.n_linearize for r in res], [bool(r.converged) for r in res]


t_clean, nl, _ = run(clean)
t_bad, nl2, cv2 = run(clean[:15] + [bad])
print("clean 16-pair batch: %.2f ms (linearize passes %s)" % (t_clean, nl))
print("15 clean pairs + seed 1104: %.2f ms (linearize passes %s, converged %s)" % (t_bad, nl2, cv2))
print("ratio %.2f" % (t_bad / t_clean))
ctx.close()
