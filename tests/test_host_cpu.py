"""CPU-only tests of the host-side helpers of the Python mirror (no kernels are called)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_make_frame_keeps_row_major_matrices():
    import gsb200  # noqa: F401
    from gsb200 import _lib
    v = np.arange(16, dtype=np.float32).reshape(4, 4)
    p = v.T.copy()
    f = _lib.make_frame(v, torch.from_numpy(p), [1, 2, 3], 0.3, 0.4, 80, 60, (0.1, 0.2, 0.3), degree=2, clamped=False,
                        scale_modifier=1.5)
    assert list(f.view) == list(range(16))                      # forward.py:694: row-major flatten
    assert list(f.proj) == p.reshape(-1).tolist()
    assert (f.width, f.height, f.degree, f.clamped) == (80, 60, 2, 0)
    assert abs(f.scale_modifier - 1.5) < 1e-7 and abs(f.tan_fovy - 0.4) < 1e-7


@pytest.mark.parametrize("tag,size", [("sq", (800, 800)), ("hd", (1920, 1080))])
def test_load_camera_matches_reference_source(golden_dir, tag, size):
    """the camera struct.  tests/golden/ref_cameras.npz holds what the reference's own
    utils/camera_utils.py:8-91 load_camera returned (run unmodified by make_golden.py case_cameras) for all 100
    Lego train poses; ours must return the same bits for every matrix the rasterizer consumes -- world_to_camera
    (translation in the last row), full_proj_matrix, camera_center, tan_fov -- and for the malformed view_matrix
    render.py:112 uses; scene_extent restates train.py:233-257."""
    import os
    import gsb200  # noqa: F401
    from gsb200.utils.camera_utils import load_nerf_cameras, scene_extent
    g = np.load(os.path.join(golden_dir, "ref_cameras.npz"))
    W, H = size
    cams = load_nerf_cameras(W, H)
    assert len(cams) == g[f"{tag}_world_to_camera"].shape[0] == 100
    for key in ("world_to_camera", "full_proj_matrix", "camera_center", "view_matrix", "proj_matrix", "R", "T"):
        ours = np.stack([np.asarray(c[key]) for c in cams])
        want = g[f"{tag}_{key}"]
        assert ours.dtype == want.dtype and np.array_equal(ours, want), key
    for key in ("tan_fovx", "tan_fovy", "fx", "fy", "cx", "cy"):
        assert np.array_equal(np.array([c[key] for c in cams], dtype=np.float64), g[f"{tag}_{key}"]), key
    assert scene_extent(cams, 1.0) == float(g[f"{tag}_scene_extent"])
    assert all(c["width"] == W and c["height"] == H for c in cams)


def test_flat_gaussians_resize_keeps_base_and_alignment():
    """The trainer's state buffers are allocated for a capacity and re-laid out in place when densify changes the
    Gaussian count: the base address (what peers have mapped) stays, every segment starts on a 16-byte boundary for
    every count, the layout equals the C side's gsb_flat_layout rule, and a count beyond the capacity is refused."""
    import gsb200  # noqa: F401
    from gsb200 import train
    F = train.FlatGaussians(1000, "cpu", capacity=train.headroom(1000))
    base = F.store.data_ptr()
    assert F.capacity >= 1500 and F.flat.numel() == train.flat_layout(1000)[1]
    for n in (1000, 1, 3, 1499, F.capacity):
        F.resize(n)
        offs, total = train.flat_layout(n)
        assert F.n == n and F.flat.data_ptr() == base and F.flat.numel() == total
        for k in train.KEYS:
            v = F[k]
            assert (v.data_ptr() - base) == 4 * offs[k] and (v.data_ptr() - base) % 16 == 0
            assert v.numel() == n * train.WIDTH[k] and tuple(v.shape) == train.SHAPE[k](n)
        assert total <= F.store.numel()
    with pytest.raises(ValueError):
        F.resize(F.capacity + 1)
    # the layout rule of csrc/optimizer.cu gsb_flat_layout: widths 3, 3, 4, 1, 48, each segment padded to 4 floats
    offs, total = train.flat_layout(5)
    assert [offs[k] for k in train.KEYS] == [0, 16, 32, 52, 60] and total == 60 + 240
