"""On-disk formats the reference node writes, so that the engine can be driven from a saved run:

  poses_kitti.txt   one line per keyframe: the 3x4 row-major [R|t]           fast_lio_sam_qn/src/fast_lio_sam_qn.cpp:344-360
  poses_tum.txt     timestamp tx ty tz qx qy qz qw                           fast_lio_sam_qn.cpp:361-376
  pcd/%06d.pcd      the keyframe cloud in the LiDAR frame (x y z intensity)  fast_lio_sam_qn.cpp:349-352 (pcl::io::savePCDFileASCII)

Host-side I/O only; nothing here is on the hot path.
"""
import os

import numpy as np


def save_poses_kitti(path, poses):
    poses = np.asarray(poses, np.float64).reshape(-1, 4, 4)
    with open(path, "w") as f:
        for T in poses:
            f.write(" ".join("%.17g" % v for v in T[:3, :].reshape(-1)) + "\n")


def load_poses_kitti(path):
    rows = np.loadtxt(path, dtype=np.float64).reshape(-1, 12)
    out = np.tile(np.eye(4), (len(rows), 1, 1))
    out[:, :3, :] = rows.reshape(-1, 3, 4)
    return out


def _quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _rot_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = q
    return np.array([x, y, z, w])


def save_poses_tum(path, stamps, poses):
    poses = np.asarray(poses, np.float64).reshape(-1, 4, 4)
    with open(path, "w") as f:
        for t, T in zip(stamps, poses):
            q = _rot_to_quat(T[:3, :3])
            f.write("%.9f %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n" % (t, T[0, 3], T[1, 3], T[2, 3], q[0], q[1], q[2], q[3]))


def load_poses_tum(path):
    rows = np.loadtxt(path, dtype=np.float64).reshape(-1, 8)
    poses = np.tile(np.eye(4), (len(rows), 1, 1))
    for T, r in zip(poses, rows):
        T[:3, :3] = _quat_to_rot(r[4:8] / np.linalg.norm(r[4:8]))
        T[:3, 3] = r[1:4]
    return rows[:, 0].copy(), poses


def save_pcd(path, pts, binary=False):
    """pts (n, 4): x y z intensity.  ASCII like pcl::io::savePCDFileASCII, or binary."""
    pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n" % (n, n, "binary" if binary else "ascii"))
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            f.write(pts[:, :4].tobytes())
        else:
            for p in pts:
                f.write(("%.9g %.9g %.9g %.9g\n" % (p[0], p[1], p[2], p[3])).encode())


def load_pcd(path):
    """Reads x y z [intensity] float32 PCD files (ascii or binary, the two forms PCL writes for PointXYZI)."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, npts, data = [], [], [], [], 0, None
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if not line or line.startswith("#"):
                if not line:
                    break
                continue
            key, _, rest = line.partition(" ")
            if key == "FIELDS":
                fields = rest.split()
            elif key == "SIZE":
                sizes = [int(v) for v in rest.split()]
            elif key == "TYPE":
                types = rest.split()
            elif key == "COUNT":
                counts = [int(v) for v in rest.split()]
            elif key == "POINTS":
                npts = int(rest)
            elif key == "DATA":
                data = rest
                break
        counts = counts or [1] * len(fields)
        if data == "ascii":
            arr = np.loadtxt(f, dtype=np.float64).reshape(npts, -1)
        elif data == "binary":
            rec = sum(s * c for s, c in zip(sizes, counts))
            raw = np.frombuffer(f.read(rec * npts), dtype=np.uint8).reshape(npts, rec)
            cols, off = [], 0
            for s, t, c in zip(sizes, types, counts):
                dt = {("F", 4): np.float32, ("F", 8): np.float64, ("U", 4): np.uint32, ("I", 4): np.int32, ("U", 1): np.uint8,
                      ("U", 2): np.uint16, ("I", 2): np.int16, ("I", 1): np.int8}[(t, s)]
                cols.append(raw[:, off:off + s * c].copy().view(dt).reshape(npts, c).astype(np.float64))
                off += s * c
            arr = np.concatenate(cols, 1)
        else:
            raise ValueError("unsupported PCD DATA section: %r" % data)
    out = np.zeros((npts, 4), np.float32)
    for name, col in (("x", 0), ("y", 1), ("z", 2), ("intensity", 3)):
        if name in fields:
            out[:, col] = arr[:, fields.index(name)]
    return out


def save_run(directory, clouds, poses, stamps, binary=False):
    """The reference's save layout: <dir>/pcd/%06d.pcd, poses_kitti.txt, poses_tum.txt (fast_lio_sam_qn.cpp:327-413)."""
    os.makedirs(os.path.join(directory, "pcd"), exist_ok=True)
    for i, c in enumerate(clouds):
        save_pcd(os.path.join(directory, "pcd", "%06d.pcd" % i), c, binary=binary)
    save_poses_kitti(os.path.join(directory, "poses_kitti.txt"), poses)
    save_poses_tum(os.path.join(directory, "poses_tum.txt"), stamps, poses)


def load_run(directory):
    """-> dict(clouds, poses, stamps): feed it to Context.keyframes().add(...) to replay a saved reference run."""
    stamps, poses = load_poses_tum(os.path.join(directory, "poses_tum.txt"))
    kitti = os.path.join(directory, "poses_kitti.txt")
    if os.path.exists(kitti):
        poses = load_poses_kitti(kitti)  # full-precision rotation
    clouds = []
    i = 0
    while os.path.exists(os.path.join(directory, "pcd", "%06d.pcd" % i)):
        clouds.append(load_pcd(os.path.join(directory, "pcd", "%06d.pcd" % i)))
        i += 1
    return dict(clouds=clouds, poses=poses, stamps=stamps)
