"""The C++ facade (fast-lio-sam-qn_b200/host/) keeps nano_gicp::NanoGICP / quatro<T> source-compatible:
a client written with LoopClosure's call sequence compiles (CPU box) and returns the same numbers as the
Python binding of the same C ABI (GPU box)."""
import json
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "fast-lio-sam-qn_b200", "csrc")
HOST = os.path.join(REPO, "fast-lio-sam-qn_b200", "host")
CLIENT_SRC = os.path.join(REPO, "tests", "cpp", "loop_closure_client.cpp")


MOCK_PCL = os.path.join(REPO, "tests", "cpp", "mock_pcl")


def _build(tmp, pcl_config=False):
    """pcl_config: compile the facade's B200REG_HAVE_PCL configuration (real <pcl/...> / <Eigen/...> include paths) against the
    API-shaped stand-in headers of tests/cpp/mock_pcl -- this container has neither PCL nor Eigen."""
    from b200reg.build import build_native
    build_native()
    exe = os.path.join(tmp, "lc_client_pcl" if pcl_config else "lc_client")
    extra = ["-I" + MOCK_PCL, "-DB200REG_EXPECT_PCL"] if pcl_config else []
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + HOST] + extra + ["-o", exe, CLIENT_SRC,
                           "-L" + CSRC, "-lb200reg", "-Wl,-rpath," + CSRC])
    return exe


def test_facade_compiles_in_the_pcl_eigen_configuration(tmp_path):
    """The #ifdef B200REG_HAVE_PCL branch (pcl::PointCloud / Eigen::Matrix from their own headers, Eigen-typed covariance
    vectors, Eigen::Matrix<double,6,6> hessian) compiles and links."""
    assert os.path.exists(_build(str(tmp_path), pcl_config=True))


def test_facade_client_compiles_as_cxx14(tmp_path):
    """The reference builds with -std=c++14/17 (fast_lio_sam_qn/CMakeLists.txt:6); no GPU needed to link."""
    exe = _build(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("pcl_config", [False, True])
def test_facade_results_equal_python_binding(tmp_path, ctx, synth, pcl_config):
    exe = _build(str(tmp_path), pcl_config)
    for mode, seed in (("gicp", 1001), ("quatro", 2000)):
        src, dst, _ = synth.make_pair(seed, 6000, 7000, mode=mode)
        sp, dp = str(tmp_path / "s.bin"), str(tmp_path / "d.bin")
        src.tofile(sp)
        dst.tofile(dp)
        out = json.loads(subprocess.check_output([exe, sp, dp, mode]).decode())
        T = np.array(out["T"]).reshape(4, 4)
        if mode == "gicp":
            r = ctx.icp_alignment([src], [dst])[0]
            want = r["Tf"].astype(np.float64)  # getFinalTransformation().cast<double>()
        else:
            res, _ = ctx.loop_closure([src], [dst])
            r = res[0]
            want = r["T"]
        assert out["aligned"] == len(src)
        assert bool(out["converged"]) == (r["converged"] and r["fitness"] < 1.5)
        if out["valid"]:
            assert np.array_equal(T, want), (mode, np.abs(T - want).max())
            assert out["score"] == r["fitness"]


@pytest.mark.gpu
def test_facade_rest_of_the_class_surface(tmp_path, ctx, synth):
    """setSource/TargetCovariances with the reference's container type, getFinalHessian (nano_gicp.hpp:91-106,
    lsq_registration.hpp:88) through the C++ client, against the Python binding of the same ABI."""
    exe = _build(str(tmp_path))
    src, dst, _ = synth.make_pair(1001, 6000, 7000)
    sp, dp = str(tmp_path / "s.bin"), str(tmp_path / "d.bin")
    src.tofile(sp)
    dst.tofile(dp)
    out = json.loads(subprocess.check_output([exe, sp, dp, "surface"]).decode())
    assert out["valid"] == 1 and out["setters_roundtrip_same"] == 1 and out["scaled_covs_differ"] == 1 and out["cov_struct"] == 1
    r = ctx.icp_alignment([src], [dst])[0]
    H = np.array(out["H"]).reshape(6, 6)
    assert np.array_equal(H, r["final_hessian"]) and np.array_equal(H, H.T) and np.linalg.eigvalsh(H).min() > 0
    # the hessian is the H of the LAST linearize: one more linearize at the pose that pass started from reproduces it
    cs, ct = ctx.create_clouds([src, dst])
    ctx.covariances([cs, ct], 15)
    # user covariances through the C ABI: the oracle's own, original order -> the result equals the computed-covariance run
    ctx.set_covariances(cs, ctx.get_covariances(cs))
    ctx.set_covariances(ct, ctx.get_covariances(ct))
    r2 = ctx.gicp_align([cs], [ct])[0]
    assert np.array_equal(r2["T"], r["T"])
    cs.destroy(); ct.destroy()
