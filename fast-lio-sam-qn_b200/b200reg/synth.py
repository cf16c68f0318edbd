"""Synthetic KITTI-shaped scan pairs -- input manufacture only, not the hot path."""
import ctypes as C
import os
import subprocess

import numpy as np

_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
_LIB = None


def build_synth(force=False):
    so = os.path.join(_CSRC, "libb200synth.so")
    src = os.path.join(_CSRC, "synth.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fopenmp", "-shared", "-fPIC", src, "-o", so])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_synth())
    return _LIB


def se3(yaw=0.0, pitch=0.0, roll=0.0, t=(0.0, 0.0, 0.0)):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def scan(scene_seed, scan_seed, pose, n):
    """n x 4 float32 (x, y, z, intensity) in the SENSOR frame."""
    pose = np.ascontiguousarray(pose, np.float64)
    out = np.empty((n, 4), np.float32)
    rc = _lib().b200synth_scan(C.c_uint64(scene_seed), C.c_uint64(scan_seed), pose.ctypes.data_as(C.POINTER(C.c_double)),
                               n, out.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != n:
        raise RuntimeError("b200synth_scan failed: %d" % rc)
    return out


def to_map_frame(pts, T):
    """utilities.hpp:164-175 transformPcd: double math, cast to float; keeps intensity."""
    out = pts.copy()
    out[:, :3] = (pts[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


def voxelize(pts, leaf):
    """pcl::VoxelGrid semantics (utilities.hpp:38-63): centroid of ALL fields per occupied voxel,
    output ordered by the linear voxel index i + j*dx + k*dx*dy.  numpy stand-in used to shape realistic Quatro inputs."""
    p = np.asarray(pts, np.float32)
    inv = np.float32(1.0 / leaf)
    mn = np.floor(p[:, :3].min(0) * inv).astype(np.int64)
    mx = np.floor(p[:, :3].max(0) * inv).astype(np.int64)
    ijk = np.floor(p[:, :3] * inv).astype(np.int64) - mn
    dims = mx - mn + 1
    lin = ijk[:, 0] + ijk[:, 1] * dims[0] + ijk[:, 2] * dims[0] * dims[1]
    order = np.argsort(lin, kind="stable")
    lin_s = lin[order]
    starts = np.flatnonzero(np.r_[True, lin_s[1:] != lin_s[:-1]])
    counts = np.diff(np.r_[starts, len(lin_s)])
    sums = np.add.reduceat(p[order].astype(np.float32), starts, axis=0)
    return (sums / counts[:, None].astype(np.float32)).astype(np.float32)


def make_pair(pair_seed, n_src, n_tgt=None, mode="gicp", voxel=None):
    """One loop-closure candidate pair in the common map frame.

    dst = scan at true pose A; src = scan at true pose B = A * T_gt, placed in the map with an
    odometry pose that drifted by D (mode "gicp": <= 0.5 m / 2 deg; "quatro": <= 10 m / 15 deg yaw).
    Registration must return T ~= inv(D) (maps src onto dst).  Returns (src, dst, T_expected).
    """
    n_tgt = n_tgt or n_src
    rng = np.random.default_rng(pair_seed)
    A = se3(yaw=rng.uniform(-0.1, 0.1), t=(rng.uniform(-20, 20), rng.uniform(-1.5, 1.5), 1.73))
    Tgt = se3(yaw=np.deg2rad(rng.uniform(-15, 15)), pitch=np.deg2rad(rng.uniform(-1, 1)),
              roll=np.deg2rad(rng.uniform(-1, 1)), t=(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-0.2, 0.2)))
    B = A @ Tgt
    if mode == "gicp":
        D = se3(yaw=np.deg2rad(rng.uniform(-2, 2)), pitch=np.deg2rad(rng.uniform(-0.3, 0.3)),
                roll=np.deg2rad(rng.uniform(-0.3, 0.3)),
                t=(rng.uniform(-0.35, 0.35), rng.uniform(-0.35, 0.35), rng.uniform(-0.05, 0.05)))
    else:
        D = se3(yaw=np.deg2rad(rng.uniform(-15, 15)), t=(rng.uniform(-7, 7), rng.uniform(-7, 7), rng.uniform(-0.3, 0.3)))
    dst = to_map_frame(scan(pair_seed, 2 * pair_seed + 1, A, n_tgt), A)
    src = to_map_frame(scan(pair_seed, 2 * pair_seed + 2, B, n_src), D @ B)
    if voxel:  # setSrcAndDstCloud voxelises both clouds (loop_closure.cpp:107, config.yaml:16 voxel_resolution 0.3)
        src, dst = voxelize(src, voxel), voxelize(dst, voxel)
    return src, dst, np.linalg.inv(D)


def se3_error(T_est, T_ref):
    """(rotation angle [rad], translation norm [m]) of inv(T_ref) * T_est."""
    E = np.linalg.inv(T_ref) @ T_est
    c = np.clip((np.trace(E[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
    return float(np.arccos(c)), float(np.linalg.norm(E[:3, 3]))


def make_sequence(seed, n_keyframes, pts_per_keyframe=30000, spacing=0.8, speed=8.0, drift_xy=0.01, drift_yaw_deg=0.01, threads=1):
    """KITTI-05-shaped keyframe sequence: laps of a two-lane street (U-turns at both ends,
    411 m per lap) so that every place is revisited after > 30 s; one scan per keyframe in the LiDAR frame; odometry
    poses = true poses with a slow random-walk drift.  Returns dict(clouds, poses (n,4,4), true_poses, stamps)."""
    rng = np.random.default_rng(seed)
    half, r = 100.0, 1.8
    lap = 4 * half + 2 * np.pi * r

    def pose_at(sarc):
        u = sarc % lap
        if u < 2 * half:                      # +x lane
            x, y, yaw = -half + u, -r, 0.0
        elif u < 2 * half + np.pi * r:        # U-turn at x = +half
            a = (u - 2 * half) / r
            x, y, yaw = half + r * np.sin(a), -r * np.cos(a), a
        elif u < 4 * half + np.pi * r:        # -x lane
            x, y, yaw = half - (u - 2 * half - np.pi * r), r, np.pi
        else:                                 # U-turn at x = -half
            a = (u - 4 * half - np.pi * r) / r
            x, y, yaw = -half - r * np.sin(a), r * np.cos(a), np.pi + a
        return se3(yaw=yaw, t=(x, y, 1.73))

    poses, true_poses, stamps = [], [], []
    D = np.eye(4)
    for k in range(n_keyframes):
        Tt = pose_at(k * spacing)
        D = D @ se3(yaw=np.deg2rad(rng.normal(0, drift_yaw_deg)), t=(rng.normal(0, drift_xy), rng.normal(0, drift_xy), rng.normal(0, drift_xy * 0.1)))
        true_poses.append(Tt)
        poses.append(D @ Tt)
        stamps.append(k * spacing / speed)
    # the scans depend only on (seed, k, true pose): generating them on several threads gives the same clouds
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            clouds = list(ex.map(lambda k: scan(seed, 7919 * seed + k, true_poses[k], pts_per_keyframe), range(n_keyframes)))
    else:
        clouds = [scan(seed, 7919 * seed + k, true_poses[k], pts_per_keyframe) for k in range(n_keyframes)]
    return dict(clouds=clouds, poses=np.array(poses), true_poses=np.array(true_poses), stamps=np.array(stamps))
