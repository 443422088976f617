"""The reference's own trainer, unmodified, over our drop-in `run_nerf_helpers`: see
tests/seam_driver.py.  Needs the reference checkout (build container); skipped where it is absent (GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isfile("/root/reference/DS_NeRF/run_nerf.py"), reason="reference checkout not present")
@pytest.mark.timeout(600)
def test_unmodified_reference_trainer_runs_over_the_dropin():
    iters = 2
    env = dict(os.environ, PYTHONSAFEPATH="1", OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_driver.py"), str(iters)], capture_output=True,
                       text=True, env=env, timeout=580)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("SEAM ")][-1]
    res = json.loads(line[5:])
    assert res["helpers_file"] == os.path.join(ROOT, "spin-nerf_b200", "dropin", "run_nerf_helpers.py")
    assert res["run_nerf_file"] == "/root/reference/DS_NeRF/run_nerf.py"
    assert res["data_file"] == os.path.join(ROOT, "spin-nerf_b200", "dropin", "data.py")       # batched RayDataset (row f1)
    assert res["load_llff_file"] == "/root/reference/DS_NeRF/load_llff.py"                     # everything else: the reference's
    assert res["nerf_class_module"].endswith("nerf") and "spin-nerf_b200" in res["nerf_class_module"]
    # one reference train step = three render() calls (run_nerf.py:1455-1470), each coarse + fine, then autograd backward
    render = ["spn_mlp_fwd_points", "spn_raw2outputs_fwd", "spn_sample_pdf_cdf", "spn_mlp_fwd_points", "spn_raw2outputs_fwd"]
    step = render * 3 + ["spn_raw2outputs_bwd", "spn_mlp_bwd"] * 6
    assert res["calls"] == step * iters, res["calls"]
    assert p.stdout.count("[TRAIN] Iter:") == iters          # the trainer's own progress line (run_nerf.py:1699-1701)


@pytest.mark.skipif(not os.path.isfile("/root/reference/DS_NeRF/run_nerf.py"), reason="reference checkout not present")
@pytest.mark.timeout(900)
def test_unmodified_reference_trainer_checkpoints_and_videos_over_the_dropin():
    """Same run with i_weights = i_video = i_testset = 2: the reference's checkpoint writer (run_nerf.py:1626-1636), its own
    render_path over the 120 spiral poses + mp4 export (:1638-1671) and the test-set render (:1682-1697) all go through the
    drop-in's get_rays / NeRF / raw2outputs / sample_pdf and the import shims."""
    env = dict(os.environ, PYTHONSAFEPATH="1", OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_driver.py"), "2", "2"], capture_output=True,
                       text=True, env=env, timeout=880)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("SEAM ")][-1][5:])
    files = res["files"]
    assert "000002.tar" in files and "args.txt" in files
    assert sum(f.endswith("rgb.mp4") for f in files) == 1 and sum(f.endswith("disp.mp4") for f in files) == 1
    assert any(f.startswith("testset_000002") and f.endswith(".png") for f in files)
    # checkpoint layout of run_nerf.py:1628-1633 with reference-shaped state_dicts (24 tensors per network, helpers:86-102)
    assert set(res["ckpt"]) >= {"global_step", "network_fn_state_dict", "network_fine_state_dict", "optimizer_state_dict"}
    for k in ("network_fn_state_dict", "network_fine_state_dict"):
        keys = res["ckpt"][k]
        assert len(keys) == 24 and "pts_linears.5.weight" in keys and "views_linears.0.bias" in keys and "alpha_linear.weight" in keys
    n_get_rays = res["calls"].count("spn_get_rays")
    assert n_get_rays == 120 + 1 + 1           # the spiral video, the held-out video frame, the test-set frame


@pytest.mark.skipif(not os.path.isfile("/root/reference/DS_NeRF/run_nerf.py"), reason="reference checkout not present")
@pytest.mark.timeout(900)
def test_unmodified_reference_trainer_lpips_branch_over_the_dropin():
    """`--lpips` for 305 iterations: the reference's perceptual-loss branch (run_nerf.py:1523-1561, iterations > 300: its own
    render_path with patch windows over our get_rays / NeRF / raw2outputs / sample_pdf, the `lpips` import shim) on top of
    the three render calls, differentiated by autograd through our backward ops."""
    env = dict(os.environ, PYTHONSAFEPATH="1", OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_driver.py"), "304", "100000", "--lpips"],
                       capture_output=True, text=True, env=env, timeout=880)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("SEAM ")][-1][5:])
    calls = res["calls"]
    lpips_iters = 4                                              # the loop runs i = 1 .. 304 (run_nerf.py:1359-1360); i > 300 -> 4
    assert calls.count("spn_get_rays") == 4 * lpips_iters        # lpips_batch_size patches per iteration, rays from c2w
    render = ["spn_mlp_fwd_points", "spn_raw2outputs_fwd", "spn_sample_pdf_cdf", "spn_mlp_fwd_points", "spn_raw2outputs_fwd"]
    plain_step = render * 3 + ["spn_raw2outputs_bwd", "spn_mlp_bwd"] * 6
    assert calls[:len(plain_step) * 300] == plain_step * 300
    tail = calls[len(plain_step) * 300:]
    # an LPIPS iteration: the step's three calls, per patch rays + one render, then backward: the step's six (coarse + fine of
    # three calls) and ONE per patch — weights are detached and z_samples carry no gradient, so only the fine network of a
    # patch is differentiated (run_nerf.py:1541-1549)
    lpips_step = render * 3 + (["spn_get_rays"] + render) * 4 + ["spn_raw2outputs_bwd", "spn_mlp_bwd"] * (6 + 4)
    assert tail == lpips_step * lpips_iters


@pytest.mark.skipif(not os.path.isfile("/root/reference/DS_NeRF/run_nerf.py"), reason="reference checkout not present")
@pytest.mark.timeout(600)
def test_unmodified_reference_trainer_sparse_depth_step_over_the_dropin():
    """`--colmap_depth --depth_loss` (the shipped config): the reference's own load_colmap_depth reads the COLMAP model that
    scene_io.write_colmap_model wrote, its train step makes FOUR render calls, and the depth term differentiates only the fine
    network of the fourth (7 backward pairs) — what Trainer.step's fourth ray group reproduces in one chunk."""
    env = dict(os.environ, PYTHONSAFEPATH="1", OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_driver.py"), "2", "100000", "--colmap_depth", "--depth_loss"],
                       capture_output=True, text=True, env=env, timeout=580)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("SEAM ")][-1][5:])
    render = ["spn_mlp_fwd_points", "spn_raw2outputs_fwd", "spn_sample_pdf_cdf", "spn_mlp_fwd_points", "spn_raw2outputs_fwd"]
    assert res["calls"] == (render * 4 + ["spn_raw2outputs_bwd", "spn_mlp_bwd"] * 7) * 2


@pytest.mark.skipif(not os.path.isfile("/root/reference/DS_NeRF/run_nerf.py"), reason="reference checkout not present")
@pytest.mark.timeout(600)
def test_unmodified_reference_render_only_over_the_dropin():
    """`--render_only --render_test` (run_nerf.py:1168-1220): the reference's own render_path with need_alpha and per-frame
    dumps over the drop-in's get_rays / NeRF / raw2outputs / sample_pdf, mp4 export through the imageio shim."""
    env = dict(os.environ, PYTHONSAFEPATH="1", OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_driver.py"), "2", "100000", "--render_only", "--render_test",
                        "--render_factor", "2"], capture_output=True, text=True, env=env, timeout=580)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("SEAM ")][-1][5:])
    d = "renderonly_test_000000/"
    for f in ("rgb.mp4", "disp.mp4", "intrinsics.txt", "rgb/000000.png", "images/000000.png", "depth/000000.npy", "disp/000000.npy",
              "weight/000000.npy", "z/000000.npy", "alpha/000000.npy", "pose/000000.txt"):
        assert d + f in res["files"], f
    assert res["calls"] == ["spn_get_rays", "spn_mlp_fwd_points", "spn_raw2outputs_fwd", "spn_sample_pdf_cdf", "spn_mlp_fwd_points",
                            "spn_raw2outputs_fwd"]                 # one held-out view, no optimisation step
