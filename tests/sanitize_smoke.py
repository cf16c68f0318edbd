"""Small end-to-end pass over EVERY kernel of libb200reg.so, sized for compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck|initcheck python tests/sanitize_smoke.py`).  Logs: profiles/r0N/sanitizer_*.txt"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "fast-lio-sam-qn_b200"))


def main():
    import b200reg
    from b200reg import synth
    ctx = b200reg.Context(0)
    src, dst, _ = synth.make_pair(1000, 3000, 3500)
    r = ctx.icp_alignment([src, src[:900]], [dst, dst[:7]])          # gicp path, ragged batch, tiny target
    cs, cd = ctx.create_clouds([src, dst])
    ctx.covariances([cs, cd], 20)
    ctx.knn(cd, src[:300], 15)
    ctx.linearize(cs, cd, np.eye(4))
    ctx.compute_error(cs, cd, np.eye(4), np.eye(4))
    ctx.set_covariances(cs, ctx.get_covariances(cs))
    ctx.covariances([cd], 12)                                        # k < capacity: the heap's pop path
    ctx.transform_cloud(cs, np.eye(4, dtype=np.float32))
    cs.destroy(); cd.destroy()
    qs, qd, _ = synth.make_pair(2000, 20000, 20000, mode="quatro", voxel=0.3)
    res, qi = ctx.loop_closure([qs, qs[:50]], [qd, qd[:40]])       # quatro path incl. an invalid pair
    from b200reg import native
    adv = native.default_quatro_params()
    adv.use_optimized_matching = 0                                  # advancedMatching + the global-memory solver
    res, qi = ctx.loop_closure([qs, qs[:50], qd], [qd, qd[:40], qs], qparams=adv)
    seq = synth.make_sequence(7, 60, pts_per_keyframe=2500, spacing=7.0)
    kf = ctx.keyframes()
    for c, T, t in zip(seq["clouds"], seq["poses"], seq["stamps"]):
        kf.add(c, T, t)
    xyzi32 = np.zeros((len(seq["clouds"][0]), 8), np.float32)      # pcl::PointXYZI records through the repack kernel
    xyzi32[:, :3] = seq["clouds"][0][:, :3]
    xyzi32[:, 4] = seq["clouds"][0][:, 3]
    kf.add(xyzi32, seq["poses"][0], 0.0)
    kf.add_world(synth.to_map_frame(seq["clouds"][1], seq["poses"][1]), seq["poses"][1][:3, 3], [0.0, 0.0, 0.0, 1.0], 0.1)
    q = np.arange(60, dtype=np.int32)
    cl = kf.fetch_closest(q)
    sel = q[cl >= 0][:3]
    if len(sel):
        kf.perform_loop_closure(sel, cl[cl >= 0][:3])
        cfg = b200reg.default_loop_config()
        cfg.enable_quatro = 0
        kf.perform_loop_closure(sel[:1], cl[cl >= 0][:1], cfg)
    kf.destroy()
    batch = b200reg.Batch(0, depth=2)                                # batch driver: two contexts, two graphs
    t1 = batch.submit_icp([src.ctypes.data], [len(src)], [dst.ctypes.data], [len(dst)], 16, 0)
    t2 = batch.submit_icp([src.ctypes.data], [len(src)], [dst.ctypes.data], [len(dst)], 16, 0)
    batch.wait(t1), batch.wait(t2)
    batch.close()
    c2 = b200reg.Context(0)
    c2.comm_init(b200reg.comm_unique_id(), 0, 1)
    c2.allgather_results(ctx.icp_alignment([src[:900]], [dst[:800]], raw=True))
    c2.comm_destroy()
    c2.close()
    ctx.close()
    print("SANITIZE_SMOKE_DONE", r[0]["converged"], qi[0]["valid"], len(sel))


if __name__ == "__main__":
    main()
