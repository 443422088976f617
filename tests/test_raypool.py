"""Ray pools (spin-nerf_b200/raypool.py) against a restatement of the array pipeline of
DS_NeRF/run_nerf.py:1225-1318 on a synthetic scene; get_rays_np itself is checked against the imported reference where
/root/reference exists.  CPU only."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sio = importlib.import_module("spin-nerf_b200.scene_io")
rp = importlib.import_module("spin-nerf_b200.raypool")
HAVE_REF = os.path.isfile("/root/reference/DS_NeRF/run_nerf_helpers.py")


def reference_arrays(images, poses, hwf, masks, depths, i_train, prepare, train_gt):
    """run_nerf.py:1228-1318 restated with the reference's array shapes: [N, ro+rd+rgb, H, W, 3] stacks with the label as a
    4th component, transposed to [N,H,W,3,4], training views only, flattened; then the three boolean selections."""
    H, W, focal = hwf
    rays = np.stack([np.stack(rp.camera_rays(H, W, focal, p), 0) for p in poses[:, :3, :4]], 0)          # [N,2,H,W,3]

    def with_label(label):
        lab = np.repeat(label[:, None, :, :, None], 3, axis=1)                                             # [N,3,H,W,1]
        a = np.concatenate([np.concatenate([rays, images[:, None]], 1), lab], -1)                          # [N,3,H,W,4]
        a = np.transpose(a, [0, 2, 3, 1, 4])
        return np.stack([a[i] for i in i_train], 0).reshape(-1, 3, 4).astype(np.float32)
    rays_rgb, rays_inp = with_label(masks), with_label(depths)
    clf = rays_rgb.reshape(-1, 3, 4) if (train_gt or prepare) else rays_rgb[rays_rgb[:, :, 3] == 0].reshape(-1, 3, 4)
    inp = rays_inp[rays_rgb[:, :, 3] != 0].reshape(-1, 3, 4)
    rgb = rays_rgb if prepare else rays_rgb[rays_rgb[:, :, 3] == 1].reshape(-1, 3, 4)
    return clf, inp, rgb


@pytest.mark.parametrize("prepare,train_gt,lpips", [(False, False, True), (False, False, False), (False, True, True), (True, False, True)])
def test_pools_reproduce_the_reference_arrays(prepare, train_gt, lpips, tmp_path):
    sio.synthetic_scene(str(tmp_path), n_views=8, hw=(20, 28), factor=2, seed=6, n_unlabelled=1)
    images, poses, bds, render_poses, i_test, masks, depths, _ = sio.load_scene(str(tmp_path), factor=2, lpips=lpips, prepare=prepare)
    hwf = (int(poses[0, 0, 4]), int(poses[0, 1, 4]), poses[0, 2, 4])
    i_train = [i for i in range(len(images)) if i != i_test]
    pools = rp.build_ray_pools(images, poses, hwf, masks, depths, i_train, prepare=prepare, train_gt=train_gt)
    clf, inp, rgb = reference_arrays(images, poses, hwf, masks, depths, i_train, prepare, train_gt)
    for idx, ref, fourth in ((pools.idx_clf, clf, pools.label), (pools.idx_inp, inp, pools.disp), (pools.idx_rgb, rgb, pools.label)):
        assert len(idx) == len(ref)
        assert np.array_equal(pools.o[idx], ref[:, 0, :3]) and np.array_equal(pools.d[idx], ref[:, 1, :3])
        assert np.array_equal(pools.rgb[idx], ref[:, 2, :3])
        assert np.array_equal(fourth[idx], ref[:, 0, 3]) and np.array_equal(fourth[idx], ref[:, 2, 3])
    if not prepare and not train_gt:
        assert len(pools.idx_clf) + len(pools.idx_inp) == len(pools.label)            # a pixel is either background or not
        if lpips:                                                                     # only one view keeps label +1 (load_llff.py:161)
            per_view = len(pools.label) // len(i_train)
            assert len(np.unique(pools.idx_rgb // per_view)) == 1


def test_sampling_draws_each_group_from_its_own_set(tmp_path):
    sio.synthetic_scene(str(tmp_path), n_views=8, hw=(20, 28), factor=2, seed=8, n_unlabelled=1)
    images, poses, bds, _, i_test, masks, depths, _ = sio.load_scene(str(tmp_path), factor=2, lpips=True)
    hwf = (int(poses[0, 0, 4]), int(poses[0, 1, 4]), poses[0, 2, 4])
    kept = 8 - 5                                    # the view that keeps label +1 under --lpips (load_llff.py:161)
    i_train = [i for i in range(8) if i != i_test or i == kept]
    pools = rp.build_ray_pools(images, poses, hwf, masks, depths, i_train)
    idx = pools.sample_indices(64, np.random.default_rng(0))
    assert idx.shape == (3, 64)
    assert np.all(pools.label[idx[0]] == 0) and np.all(pools.label[idx[1]] == 1) and np.all(pools.label[idx[2]] != 0)
    assert len(np.unique(idx[0])) == 64                                                # without replacement while the set is large enough


def test_device_side_index_draw_respects_the_groups(tmp_path):
    import torch
    sio.synthetic_scene(str(tmp_path), n_views=8, hw=(20, 28), factor=2, seed=8, n_unlabelled=1)
    images, poses, bds, _, i_test, masks, depths, _ = sio.load_scene(str(tmp_path), factor=2, lpips=True)
    hwf = (int(poses[0, 0, 4]), int(poses[0, 1, 4]), poses[0, 2, 4])
    pools = rp.build_ray_pools(images, poses, hwf, masks, depths, list(range(8)))
    dev = pools.to("cpu")
    assert dev["pool_od"].shape == (2, len(pools.label), 3) and dev["idx_clf"].dtype == torch.int64
    g = torch.Generator(); g.manual_seed(0)
    idx = rp.draw_step_indices(dev, 256, g)
    assert idx.shape == (3, 256)
    lab = torch.from_numpy(pools.label)
    assert bool((lab[idx[0]] == 0).all()) and bool((lab[idx[1]] == 1).all()) and bool((lab[idx[2]] != 0).all())
    assert torch.equal(dev["rgb"][idx[0]], torch.from_numpy(pools.rgb)[idx[0]])


@pytest.mark.skipif(not HAVE_REF, reason="the reference checkout only exists in the build container")
def test_camera_rays_match_reference_get_rays_np():
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    H, _ = ref_loader.load()
    rng = np.random.default_rng(0)
    c2w = rng.normal(size=(3, 4)).astype(np.float32)
    o, d = rp.camera_rays(30, 41, 37.5, c2w)
    ro, rd = H.get_rays_np(30, 41, 37.5, c2w)
    assert np.array_equal(o, ro) and np.array_equal(d, rd)


def test_dropin_raydataset_batches_equal_the_references():
    """dropin/data.py: same batches as DS_NeRF/data.py under the trainer's DataLoader construction (run_nerf.py:1340-1348),
    fetched with one gather per batch instead of N_rand __getitem__ calls."""
    import importlib.util
    import time
    import torch
    from torch.utils.data import DataLoader
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("spn_dropin_data", os.path.join(root, "spin-nerf_b200", "dropin", "data.py"))
    ours = importlib.util.module_from_spec(spec); spec.loader.exec_module(ours)

    class RefRayDataset(torch.utils.data.Dataset):           # DS_NeRF/data.py:4-15 restated (3 lines of behaviour)
        def __init__(self, ray_data):
            self.rayData, self.length = ray_data, ray_data.shape[0]

        def __len__(self):
            return self.length

        def __getitem__(self, index):
            return torch.Tensor(self.rayData[index])
    ref_cls = RefRayDataset
    if os.path.isfile("/root/reference/DS_NeRF/data.py"):    # the real one where the reference checkout exists
        spec = importlib.util.spec_from_file_location("spn_ref_data", "/root/reference/DS_NeRF/data.py")
        ref_mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_mod)
        ref_cls = ref_mod.RayDataset
    rng = np.random.default_rng(0)
    for dtype, shape in ((np.float32, (5000, 3, 4)), (np.float64, (3001, 4, 3))):
        rays = rng.standard_normal(shape).astype(dtype)
        times = []
        outs = []
        for cls in (ref_cls, ours.RayDataset):
            it = iter(DataLoader(cls(rays), batch_size=1024, shuffle=True, num_workers=0,
                                 generator=torch.Generator(device="cpu").manual_seed(3)))
            t0 = time.perf_counter()
            outs.append([b for b in it])
            times.append(time.perf_counter() - t0)
        assert len(outs[0]) == len(outs[1]) == -(-shape[0] // 1024)
        for a, b in zip(*outs):
            assert a.dtype == b.dtype == torch.float32 and torch.equal(a, b)
    assert torch.equal(ours.RayDataset(rays)[5], ref_cls(rays)[5])
    print(f"RayDataset epoch: reference {times[0] * 1e3:.1f} ms, drop-in {times[1] * 1e3:.1f} ms")


def test_train_synthetic_example_dry_run():
    """tools/train_synthetic.py --dry_run --lpips: scene written, loaded, pools built, first step's indices and LPIPS
    patches drawn — everything of the example that precedes the first GPU call."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "train_synthetic.py"), "--dry_run", "--lpips", "--views", "8",
                        "--height", "48", "--width", "64"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "first step would use indices (3, 1024)" in p.stdout and "first LPIPS step would render views" in p.stdout


def test_sparse_depth_rays_reproduce_the_references_rays_depth():
    """raypool.sparse_depth_rays against run_nerf.py:1266-1300 restated with the reference's own get_rays_by_coord_np
    (imported when the checkout is present, else the drop-in's identical numpy function)."""
    rp_mod = importlib.import_module("spin-nerf_b200.raypool")
    rng = np.random.default_rng(5)
    H, W, focal, N = 20, 30, 27.0, 4
    poses = rng.standard_normal((N, 3, 5)).astype(np.float32)
    masks = (rng.uniform(0, 1, (N, H, W)) > 0.6).astype(np.float32)
    masks[2] = -masks[2]                                     # LPIPS views carry label -1 (load_llff.py:161)
    gts = []
    for i in range(N):
        m = int(rng.integers(5, 40))
        coord = np.stack([rng.uniform(0, W + 1.5, m), rng.uniform(0, H + 1.5, m)], -1)     # some fall outside: clamped (:1274-1279)
        gts.append({"coord": coord, "depth": rng.uniform(1, 8, m), "weight": rng.uniform(0, 2, m)})
    try:
        from oracle import ref_loader
        by_coord = ref_loader.load()[0].get_rays_by_coord_np if ref_loader.available() else None
    except Exception:
        by_coord = None
    if by_coord is None:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spin-nerf_b200", "dropin"))
        by_coord = importlib.import_module("run_nerf_helpers").get_rays_by_coord_np
    i_train = [0, 2, 3]
    for prepare in (False, True):
        ref_list = []
        for i in i_train:                                    # run_nerf.py:1270-1296
            g = {k: v.copy() for k, v in gts[i].items()}
            if not prepare:
                ind = [_ for _ in range(len(g['coord']))
                       if masks[i][min(int(g['coord'][_][1]), masks[i].shape[0] - 1)][min(int(g['coord'][_][0]), masks[i].shape[1] - 1)] == 0]
                g = {k: v[ind] for k, v in g.items()}
            rd = np.transpose(np.stack(by_coord(H, W, focal, poses[i, :3, :4], g['coord']), axis=0), [1, 0, 2])
            ref_list.append(np.concatenate([rd, np.repeat(g['depth'][:, None, None], 3, axis=2),
                                            np.repeat(g['weight'][:, None, None], 3, axis=2)], axis=1))
        ref = np.concatenate(ref_list, axis=0)
        rays, depth, weight = rp_mod.sparse_depth_rays(gts, poses, (H, W, focal), masks, i_train, prepare=prepare)
        assert rays.shape == (2, ref.shape[0], 3) and rays.dtype == np.float32
        np.testing.assert_array_equal(rays.transpose(1, 0, 2), ref[:, :2].astype(np.float32))
        np.testing.assert_array_equal(depth, ref[:, 2, 0].astype(np.float32))
        np.testing.assert_array_equal(weight, ref[:, 3, 0].astype(np.float32))


def test_run_nerf_fused_cli_dry_run(tmp_path):
    """tools/run_nerf_fused.py with a reference-style config file (the keys of DS_NeRF/configs/config.txt) on a synthetic scene
    with a COLMAP sparse model: config parsing, scene loading, the four ray pools, LPIPS sampler — up to the first GPU call."""
    import subprocess
    sio = importlib.import_module("spin-nerf_b200.scene_io")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scene = str(tmp_path / "scene")
    info = sio.synthetic_scene(scene, n_views=6, hw=(24, 32), factor=2, seed=1, n_unlabelled=1)
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-1, 1, 60), rng.uniform(-1, 1, 60), rng.uniform(-6, -2.5, 60)], 1)
    sio.write_colmap_model(scene, info["c2w"], info["focal"], (48, 64), pts, rng.uniform(0.2, 2.0, 60), rng)
    cfg = tmp_path / "config.txt"
    cfg.write_text("expname = t\nN_gt = 40\nbasedir = %s\ndataset_type = llff\nfactor = 2\nN_rand = 64\nN_samples = 64\n"
                   "N_importance = 64\nuse_viewdirs = True\nraw_noise_std =1e0\ncolmap_depth = True\ndepth_loss = True\n"
                   "depth_lambda = 0.1\nno_ndc = True\nlindisp = True\nrender_factor = 1\ni_feat = 2000\ni_video = 2000\n"
                   "feat_weight = 0.1\nlrate = 0.03\nlrate_decay = 10\nwhite_bkgd = True\n" % (tmp_path / "logs"))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "run_nerf_fused.py"), "--config", str(cfg), "--datadir", scene,
                        "--lpips", "--no_tcnn", "--N_gt", "0", "--dry_run"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    assert "sparse-depth rays" in p.stdout and "dry run: stopping before the first GPU call" in p.stdout
    written = (tmp_path / "logs" / "t" / "args.txt").read_text()
    assert "lrate = 0.03" in written and "colmap_depth = True" in written and "N_rand = 64" in written
    held = subprocess.run([sys.executable, os.path.join(root, "tools", "run_nerf_fused.py"), "--config", str(cfg), "--datadir", scene,
                           "--N_gt", "2", "--dry_run"], capture_output=True, text=True, timeout=300)
    assert held.returncode == 0 and "4 training views" in held.stdout          # 6 views, the first 2 held out
    none_left = subprocess.run([sys.executable, os.path.join(root, "tools", "run_nerf_fused.py"), "--config", str(cfg), "--datadir", scene,
                                "--dry_run"], capture_output=True, text=True, timeout=300)    # the config's N_gt = 40
    assert none_left.returncode != 0 and "leaves no training view" in (none_left.stdout + none_left.stderr)
    bad = subprocess.run([sys.executable, os.path.join(root, "tools", "run_nerf_fused.py"), "--config", str(cfg), "--datadir", scene,
                          "--N_gt", "0", "--sigma_loss", "--dry_run"], capture_output=True, text=True, timeout=300)
    assert bad.returncode != 0 and "outside the fused hot path" in (bad.stdout + bad.stderr)


def test_run_nerf_fused_cli_loop(tmp_path):
    """The whole loop of tools/run_nerf_fused.py under the call recorder (tests/fused_cli_driver.py): 3 steps with the
    sparse-depth group, checkpoint + video + test-set render at step 2, then a second run that resumes from the checkpoint."""
    import json
    import subprocess
    sio = importlib.import_module("spin-nerf_b200.scene_io")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scene, logs = str(tmp_path / "scene"), str(tmp_path / "logs")
    info = sio.synthetic_scene(scene, n_views=6, hw=(24, 32), factor=2, seed=1, n_unlabelled=1)
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-1, 1, 60), rng.uniform(-1, 1, 60), rng.uniform(-6, -2.5, 60)], 1)
    sio.write_colmap_model(scene, info["c2w"], info["focal"], (48, 64), pts, rng.uniform(0.2, 2.0, 60), rng)

    def run(*extra):
        p = subprocess.run([sys.executable, os.path.join(root, "tests", "fused_cli_driver.py"), scene, logs] + list(extra),
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
        return json.loads([l for l in p.stdout.splitlines() if l.startswith("CLI ")][-1][4:]), p.stdout
    res, out = run("--colmap_depth", "--depth_loss", "--N_iters", "3", "--i_weights", "2", "--i_video", "2", "--i_testset", "2")
    assert res["rc"] == 0 and out.count("[TRAIN] Iter:") == 3
    assert "000002.tar" in res["files"] and "t_000002_rgb.mp4" in res["files"] and "t_000002_disp.mp4" in res["files"]
    assert any(f.startswith("testset_000002/rgb/") for f in res["files"])
    c = res["counts"]
    assert c["spn_render_rays_bwd"] == 3 and c["spn_adam_step"] == 6 and c["spn_train_losses"] == 3
    assert c["spn_get_rays"] == 120 + 1                     # the spiral video and the held-out test view
    lama = str(tmp_path / "lama")
    res3, out3 = run("--prepare", "--no_reload", "--N_iters", "2", "--i_feat", "2", "--i_weights", "100", "--render_factor", "1",
                     "--lama_dir", lama)
    assert sorted(os.listdir(lama)) == ["img%03d.png" % j for j in range(6)] + ["label"] and len(os.listdir(os.path.join(lama, "label"))) == 6
    assert res3["counts"]["spn_train_losses"] == 2 and "disparity / mask pairs" in out3
    res2, out2 = run("--N_iters", "4", "--i_weights", "100", "--lpips", "--lpips_from", "3")
    assert res2["counts"]["spn_render_rays_bwd"] == 3       # step 3, step 4 and step 4's LPIPS patch chunk
    assert "Reloading from" in out2 and out2.count("[TRAIN] Iter:") == 2 and "[TRAIN] Iter: 3 " in out2      # resumed after step 2
