"""Where does a from-host step of the sequence workload spend its wall time?  (bench.py's e2e arm of configs[4])
    python profiles/diag_sequence_e2e.py [keyframes] [big_first]
big_first=1: run a raw 100k x 100k loop closure on the same context first (the state bench.py reaches the sequence in)."""
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

nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
big_first = len(sys.argv) > 2 and sys.argv[2] == "1"
ctx = b200reg.Context(0)
if big_first:
    s, d, _ = synth.make_pair(2000, 100000, 100000, mode="quatro")
    ctx.loop_closure([s, s], [d, d])
    ctx.loop_closure([s, s], [d, d])
B, pts = 16, 30000
seq = synth.make_sequence(5, nkf, pts_per_keyframe=pts, threads=8)
kf = ctx.keyframes()
for c, T, t in zip(seq["clouds"], seq["poses"], seq["stamps"]):
    kf.add(c, T, t)
ctx.synchronize()
cfg = b200reg.default_loop_config()
allq = np.arange(nkf, dtype=np.int32)
cl_all = kf.fetch_closest(allq, cfg.loop_detection_radius, cfg.loop_detection_timediff_threshold)
cand = allq[cl_all >= 0]
batches = [cand[i:i + B] for i in range(0, len(cand) - B + 1, B)]
pinned = [torch.from_numpy(seq["clouds"][q]).pin_memory() for q in cand[:B]]
print("keyframes", nkf, "candidates", len(cand), "batches", len(batches), "big_first", big_first)


def step(i, ingest, acc):
    q = batches[i % len(batches)]
    t0 = time.perf_counter()
    if ingest:
        for j, qq in enumerate(q):
            kf.add(pinned[j % len(pinned)].numpy(), seq["poses"][qq], seq["stamps"][qq])
    t1 = time.perf_counter()
    cl = kf.fetch_closest(q, cfg.loop_detection_radius, cfg.loop_detection_timediff_threshold)
    t2 = time.perf_counter()
    kf.perform_loop_closure(q, cl, cfg, raw=True)
    t3 = time.perf_counter()
    acc.append((t1 - t0, t2 - t1, t3 - t2))


for ingest in (False, True, False, True):
    acc = []
    for i in range(3):
        step(i, ingest, [])
    t0 = time.perf_counter()
    for i in range(40):
        step(i, ingest, acc)
    tot = time.perf_counter() - t0
    a = np.array(acc) * 1e3
    print("ingest=%-5s  %.2f ms per step: add %.2f (max %.2f)  fetch_closest %.2f (max %.2f)  perform_loop_closure %.2f (max %.2f)  -> %.0f attempts/s"
          % (ingest, 1e3 * tot / 40, a[:, 0].mean(), a[:, 0].max(), a[:, 1].mean(), a[:, 1].max(), a[:, 2].mean(), a[:, 2].max(), B * 40 / tot))
