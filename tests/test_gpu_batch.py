"""The batch driver and the collective behind the C ABI (b200reg_batch_*, b200reg_comm_*,
b200reg_allgather_results), from Python (ctypes) and from a C++ client."""
import ctypes as C
import json
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "fast-lio-sam-qn_b200", "csrc")


@pytest.fixture(scope="module")
def pairs(synth):
    return [synth.make_pair(1200 + i, 9000 + 600 * i, 10000 - 400 * i) for i in range(6)]


def _ptrs(arrs):
    return [a.ctypes.data for a in arrs], [len(a) for a in arrs]


def test_batch_driver_equals_direct_calls(ctx, pairs):
    import b200reg
    srcs = [np.ascontiguousarray(p[0]) for p in pairs]
    dsts = [np.ascontiguousarray(p[1]) for p in pairs]
    direct = ctx.icp_alignment(srcs, dsts)
    batch = b200reg.Batch(0, depth=3)
    assert batch.depth == 3
    sp, sn = _ptrs(srcs)
    dp, dn = _ptrs(dsts)
    # several jobs in flight at once: whole batch, halves, single pairs -- any split gives the same bytes per pair
    t_all = batch.submit_icp(sp, sn, dp, dn, 16, 0)
    t_a = batch.submit_icp(sp[:3], sn[:3], dp[:3], dn[:3], 16, 0)
    t_b = batch.submit_icp(sp[3:], sn[3:], dp[3:], dn[3:], 16, 0)
    singles = [batch.submit_icp(sp[i:i + 1], sn[i:i + 1], dp[i:i + 1], dn[i:i + 1], 16, 0) for i in range(6)]
    res_all, lat = batch.wait(t_all, want_latency=True)
    assert lat > 0
    halves = list(batch.wait(t_a)) + list(batch.wait(t_b))
    ones = [batch.wait(t)[0] for t in singles]
    for i in range(6):
        want = direct[i]
        for r in (res_all[i], halves[i], ones[i]):
            assert np.array_equal(np.array(r.T).reshape(4, 4), want["T"]) and r.fitness == want["fitness"]
            assert r.n_linearize == want["n_linearize"] and bool(r.converged) == want["converged"]
    assert batch.launch_count > 0
    # coarse-to-fine jobs through the same driver
    import b200reg as B
    q = [B.synth.make_pair(2000 + i, 30000, 30000, mode="quatro", voxel=0.3) for i in range(2)]
    qs, qd = [p[0] for p in q], [p[1] for p in q]
    dres, dqi = ctx.loop_closure(qs, qd)
    sp, sn = _ptrs(qs)
    dp, dn = _ptrs(qd)
    t = batch.submit_loop_closure(sp, sn, dp, dn, 16, 0)
    res, qi = batch.wait(t, want_quatro=True)
    for i in range(2):
        assert np.array_equal(np.array(res[i].T).reshape(4, 4), dres[i]["T"]) and qi[i].n_corr == dqi[i]["n_corr"]
        assert np.array_equal(np.array(res[i].pose_between).reshape(4, 4), dres[i]["pose_between"])
    batch.close()


def test_batch_driver_reports_job_errors(pairs):
    """A failing job does not take the driver down: its status and message come back from wait()."""
    import b200reg
    from b200reg import native
    batch = b200reg.Batch(0, depth=2)
    src, dst = np.ascontiguousarray(pairs[0][0]), np.ascontiguousarray(pairs[0][1])
    bad = batch.submit_icp([src.ctypes.data], [len(src)], [dst.ctypes.data], [len(dst)], 10, 0)  # stride not a multiple of 4
    good = batch.submit_icp([src.ctypes.data], [len(src)], [dst.ctypes.data], [len(dst)], 16, 0)
    with pytest.raises(b200reg.B200RegError) as e:
        batch.wait(bad)
    assert "stride" in str(e.value)
    assert batch.wait(good)[0].converged
    assert native.lib().b200reg_batch_wait(batch.h, C.c_int64(12345), None) == -1  # unknown ticket
    batch.close()


def test_allgather_world1_and_without_communicator(ctx, pairs):
    """NCCL through the C ABI on one GPU: world = 1 communicator, and the no-communicator copy path."""
    import b200reg
    res = ctx.icp_alignment([pairs[0][0], pairs[1][0]], [pairs[0][1], pairs[1][1]], raw=True)
    plain = ctx.allgather_results(res)  # no communicator: world 1, a copy
    assert bytes(plain) == bytes(res)
    c2 = b200reg.Context(0)
    c2.comm_init(b200reg.comm_unique_id(), 0, 1)
    assert c2.comm_world == 1
    got = c2.allgather_results(res)
    assert bytes(got) == bytes(res)
    with pytest.raises(b200reg.B200RegError):
        c2.comm_init(b200reg.comm_unique_id(), 0, 1)  # ESTATE: already has one
    c2.comm_destroy()
    c2.close()


def test_allgather_two_ranks_in_one_process(pairs):
    """Two ranks as host threads of ONE process, one context per GPU (also exercises the per-device kernel attributes:
    the Quatro solver's opt-in shared memory must be set on every device a context lives on)."""
    import torch
    import b200reg
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    uid = b200reg.comm_unique_id()
    q = b200reg.synth.make_pair(2000, 30000, 30000, mode="quatro", voxel=0.3)
    out = [None, None]
    err = []

    def rank_main(r):
        try:
            c = b200reg.Context(r)
            c.comm_init(uid, r, 2)
            s, d = pairs[r][0], pairs[r][1]
            res = c.icp_alignment([s, q[0]], [d, q[1]], raw=True)
            lc, _ = c.loop_closure([q[0]], [q[1]])  # Quatro kernels on device r
            allr = c.allgather_results(res)
            out[r] = (bytes(allr), bytes(res), lc[0]["T"].copy())
            c.comm_destroy()
            c.close()
        except Exception as e:  # noqa: BLE001
            err.append(e)
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not err, err
    assert out[0][0] == out[1][0], "every rank holds the same gathered bytes"
    assert out[0][0] == out[0][1] + out[1][1], "rank order"
    assert np.array_equal(out[0][2], out[1][2]), "the same pair registers identically on either GPU"


def test_cxx_batch_and_comm_client(tmp_path, ctx, pairs):
    import torch
    from b200reg.build import build_native
    build_native()
    exe = str(tmp_path / "bc_client")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-pthread", "-o", exe,
                           os.path.join(REPO, "tests", "cpp", "batch_comm_client.cpp"), "-L" + CSRC, "-lb200reg", "-Wl,-rpath," + CSRC])
    src, dst = np.ascontiguousarray(pairs[2][0]), np.ascontiguousarray(pairs[2][1])
    sp, dp = str(tmp_path / "s.bin"), str(tmp_path / "d.bin")
    src.tofile(sp)
    dst.tofile(dp)
    world = min(2, torch.cuda.device_count())
    txt = subprocess.check_output([exe, sp, dp, str(len(src)), str(len(dst)), str(world)], timeout=300).decode()
    out = json.loads([ln for ln in txt.splitlines() if ln.startswith("{")][-1])  # NCCL may print its version banner on stdout first
    assert out["status"] == 0 and out["identical"] and out["world"] == world
    want = ctx.icp_alignment([src], [dst])[0]
    assert np.array_equal(np.array(out["T"]).reshape(4, 4), want["T"]) and out["fitness"] == want["fitness"]
    assert out["latency_ms"] > 0
