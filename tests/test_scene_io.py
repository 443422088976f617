"""Scene reader / writer (spin-nerf_b200/scene_io.py) against the reference loader
DS_NeRF/load_llff.py:load_llff_data — live where /root/reference exists (the build container), and against golden outputs of
that loader committed under tests/golden/scene_io.npz (made by tests/golden/make_scene_golden.py) everywhere.  CPU only."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sio = importlib.import_module("spin-nerf_b200.scene_io")
from make_scene_golden import CASES, colmap_case, flatten_rays, reference_loader, rowsum, run_reference   # noqa: E402

HAVE_REF = os.path.isfile("/root/reference/DS_NeRF/load_llff.py")
NAMES = ["images", "poses", "bds", "render_poses", "i_test", "masks", "inpainted_depths", "mask_indices"]


def ours(scene_dir, kw):
    kw = dict(kw)
    ren = {"recenter": "recenter_poses", "spherify": "spherify_poses"}
    return sio.load_scene(scene_dir, **{ren.get(k, k): v for k, v in kw.items()})


@pytest.mark.parametrize("case", list(CASES))
def test_loader_matches_committed_reference_outputs(case, tmp_path):
    g = np.load(os.path.join(ROOT, "tests", "golden", "scene_io.npz"))
    skw, lkw = CASES[case]
    sio.synthetic_scene(str(tmp_path), **skw)
    images, poses, bds, render_poses, i_test, masks, depths, mask_indices = ours(str(tmp_path), lkw)
    assert np.array_equal(poses, g[f"{case}__poses"])
    assert np.array_equal(np.asarray(bds), g[f"{case}__bds"])
    assert np.array_equal(render_poses, g[f"{case}__render_poses"])
    assert i_test == int(g[f"{case}__i_test"]) and list(mask_indices) == list(g[f"{case}__mask_indices"])
    assert np.array_equal(rowsum(masks), g[f"{case}__masks_rowsum"])
    assert np.array_equal(rowsum(depths), g[f"{case}__depths_rowsum"])
    assert np.array_equal(rowsum(images.reshape(images.shape[0], images.shape[1], -1)), g[f"{case}__images_rowsum"])


@pytest.mark.skipif(not HAVE_REF, reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("kw", [
    dict(factor=2, lpips=True), dict(factor=2, lpips=False), dict(factor=2, lpips=True, prepare=True),
    dict(factor=2, lpips=False, spherify=True), dict(factor=2, lpips=False, bd_factor=None),
    dict(factor=2, lpips=True, recenter=False, spherify_hack=False)])
@pytest.mark.parametrize("with_depth", [True, False])
def test_loader_is_bit_exact_with_the_reference_loader(kw, with_depth, tmp_path):
    meta = sio.synthetic_scene(str(tmp_path), n_views=8, hw=(40, 56), factor=2, seed=11, n_unlabelled=2)
    if not with_depth:      # no depth directory: both loaders fall back to the label files (load_llff.py:117-121)
        import shutil
        shutil.rmtree(os.path.join(str(tmp_path), "images_2", "depth"))
    got = ours(str(tmp_path), kw)
    want = run_reference(reference_loader(), str(tmp_path), kw)
    for name, a, b in zip(NAMES, got, want):
        if name in ("i_test", "mask_indices"):
            assert (a == b) if name == "i_test" else (list(a) == list(b)), name
        else:
            a, b = np.asarray(a), np.asarray(b)
            assert a.shape == b.shape and a.dtype == b.dtype, (name, a.shape, b.shape, a.dtype, b.dtype)
            assert np.array_equal(a, b, equal_nan=True), (name, float(np.nanmax(np.abs(a - b))))


def test_write_then_load_round_trip(tmp_path):
    meta = sio.synthetic_scene(str(tmp_path), n_views=5, hw=(32, 48), factor=2, seed=2, n_unlabelled=1)
    images, poses, bds, render_poses, i_test, masks, depths, mask_indices = sio.load_scene(
        str(tmp_path), factor=2, recenter_poses=False, bd_factor=None, spherify_hack=False, lpips=False)
    assert np.allclose(poses[:, :, :4], meta["c2w"], atol=1e-6)                      # [right, up, back | position] restored
    assert np.allclose(poses[:, :, 4], [32, 48, meta["focal"] / 2])                  # hwf at the training resolution
    assert np.allclose(bds, meta["bounds"], rtol=1e-6)
    assert np.abs(images - meta["inpainted"]).max() <= 0.5 / 255 + 1e-6             # 8-bit PNG quantisation
    assert np.abs(depths - meta["depths"]).max() <= 0.5 / 255 + 1e-6
    assert mask_indices == [0, 1, 2, 3] and np.all(masks[4] == -1)                   # the unlabelled view is marked -1
    for i in range(4):                                                               # dilation only grows the object
        assert np.all(masks[i][meta["masks"][i] != 0] == 1) and masks[i].sum() > meta["masks"][i].sum()
    assert render_poses.shape == (120, 3, 5) and 0 <= i_test < 5


def test_poses_bounds_layout():
    c2w = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    pb = sio.poses_bounds(c2w, (756, 1008, 907.2), [[1, 2], [3, 4]])
    assert pb.shape == (2, 17)
    m = pb[0, :15].reshape(3, 5)
    assert np.array_equal(m[:, 0], -c2w[0, :, 1]) and np.array_equal(m[:, 1], c2w[0, :, 0])
    assert np.array_equal(m[:, 2], c2w[0, :, 2]) and np.array_equal(m[:, 3], c2w[0, :, 3])
    assert np.array_equal(m[:, 4], [756, 1008, 907.2]) and np.array_equal(pb[1, 15:], [3, 4])


def test_colmap_depth_rays_match_committed_reference_outputs(tmp_path):
    g = np.load(os.path.join(ROOT, "tests", "golden", "scene_io.npz"))
    colmap_case(sio, str(tmp_path))
    rays = sio.colmap_depth_rays(str(tmp_path), factor=2, bd_factor=.75)
    assert [len(r["depth"]) for r in rays] == list(g["colmap__counts"])
    for k, v in flatten_rays(rays).items():
        assert np.allclose(v, g[f"colmap__{k}"], rtol=1e-12, atol=0), k      # float64 dot products: summation order may differ


@pytest.mark.skipif(not HAVE_REF, reason="the reference checkout only exists in the build container")
@pytest.mark.parametrize("bd_factor", [.75, None])
def test_colmap_depth_rays_match_the_reference_loader(bd_factor, tmp_path):
    colmap_case(sio, str(tmp_path))
    got = sio.colmap_depth_rays(str(tmp_path), factor=2, bd_factor=bd_factor)
    want = reference_loader().load_colmap_depth(str(tmp_path), factor=2, bd_factor=bd_factor)
    assert len(got) == len(want) > 0
    for a, b in zip(got, want):
        for k in ("depth", "coord", "weight"):
            assert np.asarray(a[k]).shape == np.asarray(b[k]).shape
            assert np.allclose(a[k], b[k], rtol=1e-12, atol=0), k


def test_colmap_binary_round_trip(tmp_path):
    meta, pts = colmap_case(sio, str(tmp_path))
    images = sio.read_colmap_images(os.path.join(str(tmp_path), "sparse", "0", "images.bin"))
    points = sio.read_colmap_points(os.path.join(str(tmp_path), "sparse", "0", "points3D.bin"))
    assert [im["id"] for im in images] == list(range(1, 7)) and images[2]["name"] == "IMG_0002.png"
    assert np.array_equal(points["ids"], np.arange(1, 61)) and np.allclose(points["xyz"], pts)
    for k, im in enumerate(images):      # the stored world-to-camera pose inverts to the camera we wrote (COLMAP axes: y down, z forward)
        R = sio.quat_to_rot(im["qvec"])
        assert np.allclose(R.T, meta["c2w"][k][:, :3] * np.array([1., -1., -1.]), atol=1e-12)
        assert np.allclose(-R.T @ im["tvec"], meta["c2w"][k][:, 3], atol=1e-12)
        seen = im["point3D_ids"] > 0      # every matched keypoint is the pinhole projection of its point
        cam = (pts[im["point3D_ids"][seen] - 1] - meta["c2w"][k][:, 3]) @ R.T
        assert np.allclose(cam[:, :2] / cam[:, 2:3] * meta["focal"] + [64, 48], im["xys"][seen], atol=1e-9)


def test_quaternion_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(500):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        q = -q if q[0] < 0 else q
        R = sio.quat_to_rot(q)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
        assert np.allclose(sio.rot_to_quat(R), q, atol=1e-12)
