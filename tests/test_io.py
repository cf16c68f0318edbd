"""Saved-run formats: KITTI / TUM pose files and PCD keyframe clouds round-trip."""
import numpy as np


def test_saved_run_roundtrip(tmp_path, synth):
    from b200reg import io
    seq = synth.make_sequence(3, 6, pts_per_keyframe=300, spacing=5.0)
    d = str(tmp_path / "run")
    io.save_run(d, seq["clouds"], seq["poses"], seq["stamps"], binary=False)
    back = io.load_run(d)
    assert len(back["clouds"]) == 6
    assert np.allclose(back["poses"], seq["poses"], atol=1e-14)
    assert np.allclose(back["stamps"], seq["stamps"], atol=1e-9)
    for a, b in zip(back["clouds"], seq["clouds"]):
        assert np.array_equal(a, b)  # %.9g round-trips fp32 exactly
    # binary PCD and the TUM quaternion path
    io.save_pcd(str(tmp_path / "b.pcd"), seq["clouds"][0], binary=True)
    assert np.array_equal(io.load_pcd(str(tmp_path / "b.pcd")), seq["clouds"][0])
    st, poses = io.load_poses_tum(d + "/poses_tum.txt")
    assert np.abs(poses - seq["poses"]).max() < 1e-12


def test_result_bag_roundtrip_and_layout(tmp_path, synth):
    """result.bag (fast_lio_sam_qn.cpp:377-394): rosbag 2.0 with /keyframe_pcd (PointCloud2 of pcl::PointXYZI records) and
    /keyframe_pose (PoseStamped, quaternion through the reference's RPY round trip), stamped with the keyframe time."""
    import struct
    from b200reg import io
    seq = synth.make_sequence(4, 5, pts_per_keyframe=200, spacing=5.0)
    stamps = seq["stamps"] + 1700000000.25  # wall-clock-like stamps: secs / nsecs split
    p = str(tmp_path / "result.bag")
    io.save_result_bag(p, seq["clouds"], seq["poses"], stamps, frame="map")
    raw = open(p, "rb").read()
    assert raw.startswith(b"#ROSBAG V2.0\n")
    # the bag header record is padded to 4096 bytes and points at the index section
    (hl,) = struct.unpack_from("<I", raw, 13)
    hdr = io._parse_fields(raw[17:17 + hl])
    assert hdr["op"] == b"\x03" and struct.unpack("<I", hdr["conn_count"])[0] == 2 and struct.unpack("<I", hdr["chunk_count"])[0] == 5
    (idx,) = struct.unpack("<Q", hdr["index_pos"])
    first = next(io._records(raw, idx))
    assert first[0]["op"] == b"\x07" and first[0]["topic"] == b"/keyframe_pcd"
    conn = io._parse_fields(first[1])
    assert conn["type"] == b"sensor_msgs/PointCloud2" and conn["md5sum"] == b"1158d486dd51d683ce2f1be655c3c181"
    (_, _, pos0) = next(io._records(raw, 4096))
    assert pos0 == 4096  # first chunk right after the padded header
    back = io.load_result_bag(p)
    assert back["frame"] == "map" and len(back["clouds"]) == 5
    assert np.allclose(back["stamps"], stamps, atol=1e-6)
    for a, b in zip(back["clouds"], seq["clouds"]):
        assert np.array_equal(a, b)
    assert np.abs(back["poses"] - seq["poses"]).max() < 1e-9
