"""GPU parity for the "next" rows: candidate search, cloud assembly (transformPcd + submap merge +
pcl::VoxelGrid) and the batched performLoopClosure driver, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seq(synth):
    return synth.make_sequence(7, 140, pts_per_keyframe=6000, spacing=3.0, speed=8.0)


@pytest.fixture(scope="module")
def store(ctx, seq):
    kf = ctx.keyframes()
    for c, T, t in zip(seq["clouds"], seq["poses"], seq["stamps"]):
        kf.add(c, T, t)
    yield kf
    kf.destroy()


def test_fetch_closest_equals_oracle(store, oracle, seq):
    pos = seq["poses"][:, :3, 3]
    queries = np.arange(len(pos), dtype=np.int32)
    got = store.fetch_closest(queries, 35.0, 30.0)
    want = np.array([oracle.fetch_closest(pos, seq["stamps"], q, 35.0, 30.0) for q in queries])
    assert np.array_equal(got, want)
    assert (want >= 0).sum() > 10 and (want < 0).sum() > 10


def _sorted_rows(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_assemble_quatro_and_submap_modes(ctx, store, oracle, seq):
    import b200reg
    pairs = [(137, 3), (139, 70), (120, 55)]
    src_idx = [p[0] for p in pairs]
    dst_idx = [p[1] for p in pairs]
    for quatro, submap in ((1, 0), (0, 0), (1, 1)):
        cfg = b200reg.default_loop_config()
        cfg.enable_quatro, cfg.enable_submap_matching = quatro, submap
        sc, dc = store.assemble(src_idx, dst_idx, cfg, n_keyframes=140)
        for k, (si, di) in enumerate(pairs):
            os_, od_ = oracle.set_src_and_dst_cloud(seq["clouds"], seq["poses"], si, di, submap_range=5, voxel_res=0.3,
                                                    enable_quatro=bool(quatro), enable_submap_matching=bool(submap), n_keyframes=140)
            for cloud, want in ((sc[k], os_), (dc[k], od_)):
                got = ctx.cloud_points(cloud)
                assert got.shape[0] == want.shape[0], "voxel count must match pcl::VoxelGrid's"
                # same voxel order (linear voxel index); centroids agree to fp32 summation-order noise
                assert np.abs(got - want[:, :3]).max() < 2e-4
        for c in sc + dc:
            c.destroy()


def test_perform_loop_closure_matches_oracle(ctx, store, oracle, synth, seq):
    import b200reg
    queries = np.array([137, 139, 20, 118], np.int32)
    closest = store.fetch_closest(queries)
    assert closest[2] == -1 and (closest[[0, 1, 3]] >= 0).all()
    res, qi = store.perform_loop_closure(queries, closest)
    assert not res[2]["valid"] and np.array_equal(res[2]["T"], np.eye(4))  # dummy output (loop_closure.cpp:201-204)
    for k in (0, 1, 3):
        q, c = int(queries[k]), int(closest[k])
        src, dst = oracle.set_src_and_dst_cloud(seq["clouds"], seq["poses"], q, c, n_keyframes=q + 1)
        if qi[k]["valid"]:
            o = oracle.coarse_to_fine(src, dst, quatro_T=qi[k]["T"])
            rot, tr = synth.se3_error(res[k]["T"], o["T"])
            assert rot < 3e-4 and tr < 3e-3, (k, rot, tr)  # inputs differ by the voxel-centroid fp32 noise (2e-4 m)
            # the loop closure undoes the accumulated odometry drift between the two keyframes
            drift = seq["true_poses"][c] @ np.linalg.inv(seq["poses"][c]) @ seq["poses"][q] @ np.linalg.inv(seq["true_poses"][q])
            # T maps src (map frame by odom) onto dst: T ~= D_c * inv(D_q) with D = odom * inv(true)
            Dq = seq["poses"][q] @ np.linalg.inv(seq["true_poses"][q])
            Dc = seq["poses"][c] @ np.linalg.inv(seq["true_poses"][c])
            rot, tr = synth.se3_error(res[k]["T"], Dc @ np.linalg.inv(Dq))
            assert rot < 2e-2 and tr < 0.3, ("ground truth", k, rot, tr)
    # GICP-only mode (scan-to-submap, loop_closure.cpp:94-105)
    cfg = b200reg.default_loop_config()
    cfg.enable_quatro = 0
    res2, _ = store.perform_loop_closure(queries[:2], closest[:2], cfg)
    for k in range(2):
        q, c = int(queries[k]), int(closest[k])
        src, dst = oracle.set_src_and_dst_cloud(seq["clouds"], seq["poses"], q, c, enable_quatro=False, n_keyframes=q + 1)
        o = oracle.gicp_align(src, dst)
        rot, tr = synth.se3_error(res2[k]["T"], o["T"])
        assert rot < 3e-4 and tr < 3e-3, (k, rot, tr)


def test_loop_factors_close_the_loop(ctx, store, oracle, synth, seq):
    """Result consumption (fast_lio_sam_qn.cpp:220-237): the factor records of a batch equal the oracle's, and applying
    the measurement brings the drifted latest keyframe back onto the truth relative to the matched one."""
    queries = np.array([137, 20, 118], np.int32)
    closest = store.fetch_closest(queries)
    raw, _ = store.perform_loop_closure(queries, closest, raw=True)
    facs = store.loop_factors(queries, closest, raw)
    assert facs[1]["to_idx"] == -1 and not facs[1]["valid"]  # no candidate: nothing for the graph
    for k in (0, 2):
        q, c = int(queries[k]), int(closest[k])
        f = facs[k]
        assert (f["from_idx"], f["to_idx"]) == (q, c) and f["valid"] == bool(raw[k].valid)
        T = np.array(raw[k].pose_between).reshape(4, 4)  # RegistrationOutput::pose_between_eig_ is what the factor is built from
        M, var = oracle.loop_factor(T, seq["poses"][q], seq["poses"][c], raw[k].fitness)
        assert np.abs(f["measurement"] - M).max() < 1e-12 and np.array_equal(f["variances"], var)
        # the measured relative pose is the TRUE relative pose of the two keyframes (drift removed)
        true_rel = np.linalg.inv(seq["true_poses"][q]) @ seq["true_poses"][c]
        rot, tr = synth.se3_error(f["measurement"], true_rel)
        assert rot < 2e-2 and tr < 0.3, (k, rot, tr)


def test_batched_queries_see_their_own_keyframe_count(ctx, store, oracle, seq):
    """A batch replays several loopTimerFunc ticks: the sub-map bound `i < keyframes.size() - 1` (loop_closure.cpp:72,79,100)
    is evaluated with size = query + 1 PER QUERY, so a batch equals the one-by-one replay -- also for the non-maximal
    queries, whose merged clouds must not reach past (or onto) their own query keyframe."""
    import b200reg
    cfg = b200reg.default_loop_config()
    cfg.enable_submap_matching = 1
    # closest keyframes chosen within the sub-map range of their query so that the bound actually bites
    queries = np.array([60, 100, 139], np.int32)
    closest = np.array([57, 97, 136], np.int32)
    sc, dc = store.assemble(queries, closest, cfg, n_keyframes=queries + 1)
    for k in range(3):
        q, c = int(queries[k]), int(closest[k])
        os_, od_ = oracle.set_src_and_dst_cloud(seq["clouds"], seq["poses"], q, c, submap_range=5, voxel_res=0.3,
                                                enable_quatro=True, enable_submap_matching=True, n_keyframes=q + 1)
        for cloud, want in ((sc[k], os_), (dc[k], od_)):
            got = ctx.cloud_points(cloud)
            assert got.shape[0] == want.shape[0], (k, got.shape, want.shape)
            assert np.abs(got - want[:, :3]).max() < 2e-4
    # the batch-wide bound (the old behaviour: size = max(query) + 1) gives DIFFERENT clouds for the non-maximal queries
    sc2, dc2 = store.assemble(queries, closest, cfg, n_keyframes=int(queries.max()) + 1)
    assert sc2[0].n != sc[0].n and sc2[2].n == sc[2].n
    # and the batched driver equals the one-by-one replay bit for bit
    res, _ = store.perform_loop_closure(queries, closest, cfg)
    for k in range(3):
        single, _ = store.perform_loop_closure(queries[k:k + 1], closest[k:k + 1], cfg)
        assert np.array_equal(res[k]["T"], single[0]["T"]) and res[k]["fitness"] == single[0]["fitness"]
    for c in sc + dc + sc2 + dc2:
        c.destroy()


def test_pose_pcd_ingest_matches_oracle(ctx, oracle, synth, seq):
    """Row f4: the PosePcd constructor (pose_pcd.hpp:21-43) as a device step -- the world-frame scan FAST-LIO publishes goes into
    the LiDAR frame with pose_eig_.inverse(), the pose comes from the odometry quaternion -- against the oracle, and a keyframe
    ingested that way registers exactly like one added in the LiDAR frame."""
    from b200reg import io as bio
    kf = ctx.keyframes()
    worst = 0.0
    for k in (3, 70, 137):
        T = seq["poses"][k]
        lidar = seq["clouds"][k]
        world = synth.to_map_frame(lidar, T)  # what FAST-LIO publishes (world frame)
        q = bio._rot_to_quat(T[:3, :3])
        q = q * 1.0000001  # odometry quaternions are never exactly unit: tf's setRotation divides by |q|^2
        idx = kf.add_world(world, T[:3, 3], q, seq["stamps"][k])
        got, pose, ts = kf.get(idx)
        want, opose = oracle.pose_pcd_ingest(world, T[:3, 3], q)
        assert np.abs(pose - opose).max() < 1e-15 and ts == seq["stamps"][k]
        assert np.array_equal(got[:, 3], want[:, 3])                       # intensity carried
        # two different 4x4 inverses (cofactors vs Gauss-Jordan) differ by ~1e-16 relative, i.e. ~1e-14 m in the translation:
        # the float results are at most one ulp apart (or 1e-12 m where a coordinate is nearly zero)
        ulp = np.maximum(np.spacing(np.maximum(np.abs(want[:, :3]), np.abs(got[:, :3])).astype(np.float32)), 1e-12)
        assert (np.abs(got[:, :3] - want[:, :3]) <= ulp).all()
        assert (got[:, :3] == want[:, :3]).mean() > 0.99  # measured 0.9988: cond([R|t]) ~ 1e4 amplifies the inverses' round-off
        worst = max(worst, np.abs(got[:, :3] - lidar[:, :3]).max())
    assert worst < 2e-5  # back in the LiDAR frame: the original scan up to the float round trip through the world frame
    kf.destroy()


def test_keyframe_store_slabs_and_reserve(ctx, synth):
    """Keyframes live in slabs apart from the scratch pool: a reservation that runs out rolls over to a new slab, and
    every keyframe reads back exactly as it was added, before and after registration calls used the scratch pool."""
    seq = synth.make_sequence(3, 9, pts_per_keyframe=4000, spacing=5.0)
    kf = ctx.keyframes()
    kf.reserve(2 * 4000 + 100)  # room for two keyframes and a bit: the third one opens a new slab
    for c, T, t in zip(seq["clouds"], seq["poses"], seq["stamps"]):
        kf.add(c, T, t)
    kf.reserve(0)
    q = np.array([8], np.int32)
    kf.perform_loop_closure(q, np.array([0], np.int32))
    for i, (c, T, t) in enumerate(zip(seq["clouds"], seq["poses"], seq["stamps"])):
        pts, pose, ts = kf.get(i)
        assert np.array_equal(pts, c[:, :4].astype(np.float32)) and np.array_equal(pose, T) and ts == t
    kf.destroy()
