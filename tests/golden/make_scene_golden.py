"""Golden outputs of the UNMODIFIED reference loader (DS_NeRF/load_llff.py:load_llff_data) on the deterministic synthetic
scene of spin-nerf_b200/scene_io.py:synthetic_scene — run in the build container, where /root/reference exists:

    python tests/golden/make_scene_golden.py        -> tests/golden/scene_io.npz
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
CASES = {          # name -> (scene kwargs, loader kwargs)
    "default": (dict(n_views=9, hw=(48, 64), factor=2, seed=3, n_unlabelled=2), dict(factor=2, lpips=True)),
    "spherify": (dict(n_views=7, hw=(32, 40), factor=4, seed=5, n_unlabelled=0), dict(factor=4, spherify=True, lpips=False)),
    "prepare": (dict(n_views=6, hw=(32, 40), factor=2, seed=7, n_unlabelled=1), dict(factor=2, prepare=True, lpips=True, bd_factor=None,
                                                                                     recenter=False)),
}


def rowsum(x):
    """layout-independent digest: float64 sums over the last axis of a C-contiguous copy (the reference returns strided views)"""
    return np.ascontiguousarray(x, np.float64).sum(-1)


def colmap_case(sio, scene_dir):
    """deterministic sparse model on top of a synthetic scene: 60 points, some beyond the depth bounds, random errors"""
    meta = sio.synthetic_scene(scene_dir, n_views=6, hw=(48, 64), factor=2, seed=4, n_unlabelled=1)
    rng = np.random.default_rng(9)
    pts = np.stack([rng.uniform(-1.5, 1.5, 60), rng.uniform(-1.2, 1.2, 60), -rng.uniform(0.8, 12.0, 60)], 1)
    sio.write_colmap_model(scene_dir, meta["c2w"], meta["focal"], (96, 128), pts, rng.uniform(0.2, 2.0, 60), rng)
    return meta, pts


def flatten_rays(rays):
    return {k: np.concatenate([np.asarray(r[k]).reshape(len(r["depth"]), -1) for r in rays], 0) for k in ("depth", "coord", "weight")}


def reference_loader():
    """The unmodified DS_NeRF/load_llff.py with the cv2-backed imageio stand-in of spin-nerf_b200/compat (this image has no
    imageio; other tests may have left an empty stub module under that name)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("imageio", os.path.join(ROOT, "spin-nerf_b200", "compat", "imageio.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    sys.modules["imageio"] = shim
    if "/root/reference/DS_NeRF" not in sys.path:
        sys.path.insert(0, "/root/reference/DS_NeRF")
    mod = importlib.import_module("load_llff")
    mod.imageio = shim
    return mod


def run_reference(ref, scene_dir, kw):
    kw = dict(kw)
    args = types.SimpleNamespace(lpips=kw.pop("lpips"))
    return ref.load_llff_data(scene_dir, args=args, **kw)


if __name__ == "__main__":
    sio = importlib.import_module("spin-nerf_b200.scene_io")
    ref = reference_loader()
    out = {}
    for name, (skw, lkw) in CASES.items():
        d = tempfile.mkdtemp()
        sio.synthetic_scene(d, **skw)
        images, poses, bds, render_poses, i_test, masks, depths, mask_indices = run_reference(ref, d, lkw)
        out[f"{name}__poses"] = poses
        out[f"{name}__bds"] = np.asarray(bds)
        out[f"{name}__render_poses"] = render_poses
        out[f"{name}__i_test"] = np.int64(i_test)
        out[f"{name}__mask_indices"] = np.asarray(mask_indices, np.int64)
        out[f"{name}__masks_rowsum"] = rowsum(masks)                  # [N,H] — full masks would be bulky
        out[f"{name}__depths_rowsum"] = rowsum(depths)
        out[f"{name}__images_rowsum"] = rowsum(images.reshape(images.shape[0], images.shape[1], -1))
    d = tempfile.mkdtemp()
    colmap_case(sio, d)
    rays = ref.load_colmap_depth(d, factor=2, bd_factor=.75)
    out["colmap__counts"] = np.asarray([len(r["depth"]) for r in rays], np.int64)
    for k, v in flatten_rays(rays).items():
        out[f"colmap__{k}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "scene_io.npz"), **out)
    print("wrote", len(out), "arrays")
