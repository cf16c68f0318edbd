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


def _build(tmp):
    from b200reg.build import build_native
    build_native()
    exe = os.path.join(tmp, "lc_client")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + HOST, "-o", exe, CLIENT_SRC,
                           "-L" + CSRC, "-lb200reg", "-Wl,-rpath," + CSRC])
    return exe


def test_facade_client_compiles_as_cxx14(tmp_path):
    """The reference builds with -std=c++14/17 (fast_lio_sam_qn/CMakeLists.txt:6); no GPU needed to link."""
    exe = _build(str(tmp_path))
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_facade_results_equal_python_binding(tmp_path, ctx, synth):
    exe = _build(str(tmp_path))
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
