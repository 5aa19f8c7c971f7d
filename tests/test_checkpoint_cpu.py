"""CPU: the PLY checkpoint layout of the reference's save_ply (utils/point_cloud_utils.py:10-99)."""
import numpy as np


def test_ply_layout_and_round_trip(tmp_path):
    import gsb200  # noqa: F401
    from gsb200.utils.point_cloud_utils import VERTEX_DTYPE, load_ply, save_ply
    rng = np.random.default_rng(0)
    n = 37
    params = {"positions": rng.normal(size=(n, 3)).astype(np.float32),
              "scales": rng.uniform(0.001, 0.2, (n, 3)).astype(np.float32),
              "rotations": rng.normal(size=(n, 4)).astype(np.float32),
              "opacities": rng.uniform(0, 1, n).astype(np.float32),
              "shs": rng.normal(size=(n * 16, 3)).astype(np.float32)}
    path = tmp_path / "point_cloud" / "iteration_5" / "point_cloud.ply"
    save_ply(params, str(path), n)
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().strip().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    props = [ln.split() for ln in lines[3:]]
    names = [p[2] for p in props]
    assert names == (["x", "y", "z", "scale_0", "scale_1", "scale_2", "opacity", "rot_x", "rot_y", "rot_z", "rot_w",
                      "red", "green", "blue", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)])
    assert [p[1] for p in props] == ["float"] * 11 + ["uchar"] * 3 + ["float"] * 48
    assert VERTEX_DTYPE.itemsize == 239 and len(body) == 239 * n          # packed records
    v = np.frombuffer(body, dtype=VERTEX_DTYPE)
    sh = params["shs"].reshape(n, 16, 3)
    # f_rest_{3(j-1)+c} = coefficient j, channel c (coefficient-major interleaved)
    assert np.array_equal(v["f_rest_0"], sh[:, 1, 0]) and np.array_equal(v["f_rest_4"], sh[:, 2, 1])
    assert np.array_equal(v["f_rest_44"], sh[:, 15, 2])
    want_red = np.array([int(np.clip(np.clip(c + 0.5, 0.0, 1.0) * 255, 0, 255)) for c in sh[:, 0, 0]], dtype=np.uint8)
    assert np.array_equal(v["red"], want_red)
    back = load_ply(str(path))
    for k in params:
        assert np.array_equal(back[k], params[k]), k                      # raw float32: bit-exact round trip
