"""Result consumption: the BetweenFactor record of an accepted loop closure.

fast_lio_sam_qn/src/fast_lio_sam_qn.cpp:220-231 -- pose_from = poseEigToGtsamPose(pose_between * latest.pose_corrected)
("take care of the order"), pose_to = poseEigToGtsamPose(closest.pose_corrected), measurement = pose_from.between(pose_to),
variances = score x 6.  The ABI entry point is host arithmetic only, so it is checked on the CPU box against the numpy
restatement in oracle/oracle.py and against first principles.
"""
import ctypes

import numpy as np


def _pose(rng):
    # roll / pitch / yaw rotation and a translation, built here so that the test does not depend on synth's conventions
    r, p, y = rng.uniform(-0.7, 0.7), rng.uniform(-0.7, 0.7), rng.uniform(-3.1, 3.1)
    cx, sx, cy, sy, cz, sz = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ \
        np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.uniform(-80, 80, 3)
    return T


def test_struct_layout():
    from b200reg import native
    assert ctypes.sizeof(native.LoopFactor) == 4 * 4 + 16 * 8 + 6 * 8


def test_loop_factor_matches_oracle_and_first_principles(oracle):
    from b200reg import native
    rng = np.random.default_rng(5)
    for _ in range(50):
        Tb, Pl, Pc = _pose(rng), _pose(rng), _pose(rng)
        Tb[:3, 3] *= 0.02  # a loop correction is small
        score = float(rng.uniform(0.01, 1.4))
        f = native.loop_factor_from_poses(Tb, Pl, Pc, score, True, 17, 3)
        M, var = oracle.loop_factor(Tb, Pl, Pc, score)
        assert f["from_idx"] == 17 and f["to_idx"] == 3 and f["valid"]
        assert np.abs(f["measurement"] - M).max() < 1e-12
        assert np.array_equal(f["variances"], var) and np.all(var == score)
        # exact rotations survive the roll/pitch/yaw round trip: between = inv(Tb @ Pl) @ Pc ...
        assert np.abs(f["measurement"] - np.linalg.inv(Tb @ Pl) @ Pc).max() < 1e-9
        # ... and the ORDER matters (fast_lio_sam_qn.cpp:224): Pl @ Tb is a different constraint
        assert np.abs(f["measurement"] - np.linalg.inv(Pl @ Tb) @ Pc).max() > 1e-3


def test_loop_factor_reorthonormalises_float_transforms(oracle):
    """pose_between_eig_ comes from a float matrix (getFinalTransformation().cast<double>()): not exactly orthonormal.
    poseEigToGtsamPose's RPY round trip projects it onto SO(3); the measurement is a proper rigid transform."""
    from b200reg import native
    rng = np.random.default_rng(6)
    Tb = _pose(rng).astype(np.float32).astype(np.float64)
    Pl, Pc = _pose(rng), _pose(rng)
    f = native.loop_factor_from_poses(Tb, Pl, Pc, 0.3)
    R = f["measurement"][:3, :3]
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14 and abs(np.linalg.det(R) - 1) < 1e-14
    assert np.array_equal(f["measurement"][3], [0, 0, 0, 1])
    M, _ = oracle.loop_factor(Tb, Pl, Pc, 0.3)
    assert np.abs(f["measurement"] - M).max() < 1e-12


def test_invalid_result_is_flagged_not_dropped():
    from b200reg import native
    f = native.loop_factor_from_poses(np.eye(4), np.eye(4), np.eye(4), 1.7976931348623157e308, False, 4, 1)
    assert not f["valid"] and f["from_idx"] == 4
