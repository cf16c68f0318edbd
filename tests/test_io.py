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
