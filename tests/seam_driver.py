"""Helper of tests/test_reference_seam.py (run as a subprocess): the UNMODIFIED reference trainer
(/root/reference/DS_NeRF/run_nerf.py: config parsing, load_llff, create_nerf, ray-batch assembly, DataLoaders, three render
calls per step, losses, autograd backward, torch.optim.Adam, learning-rate decay) is executed for a few iterations with
`import run_nerf_helpers` resolving to OUR drop-in module, exactly by the module-resolution recipe of INTEGRATION.md
section 1, on a scene written by scene_io.synthetic_scene.  There is no GPU here, so the C library is replaced by the call
recorder of tests/test_host_glue_dry_run.py (values are garbage): what this proves is the SEAM — every name, signature,
tensor shape, parameter / state_dict / optimizer contract the reference trainer relies on — and which C entry points one
reference train step turns into.  Prints one JSON line."""
import importlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                    # noqa: E402
import test_host_glue_dry_run as dry            # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    every = sys.argv[2] if len(sys.argv) > 2 else "100000"      # i_video / i_testset / i_weights period
    extra = sys.argv[3:]                                        # further reference flags, e.g. --lpips
    rec = dry.Recorder()
    for m in (dry.L, dry.ops, dry.render_mod):
        m.lib = (lambda rec=rec: rec); m.ptr = dry._ptr; m.stream = (lambda: 0)
    # the recorder computes nothing: hand out zero-filled outputs instead of uninitialised memory, so that the run is
    # deterministic (a NaN disparity loss is skipped by the trainer, run_nerf.py:1520, which would change the call list)
    dry.ops._empty = lambda shape, like, dtype=torch.float32: torch.zeros(shape, device=like.device, dtype=dtype)
    dry.nerf_mod.NeRF._sync = lambda self: (self.flat_params(), torch.zeros(64, dtype=torch.uint8))
    sio = importlib.import_module("spin-nerf_b200.scene_io")
    scene = tempfile.mkdtemp(prefix="spn_scene_")
    info = sio.synthetic_scene(scene, n_views=8, hw=(24, 32), factor=2, seed=0, n_unlabelled=1)
    if "--colmap_depth" in extra:       # a COLMAP sparse model for load_colmap_depth (load_llff.py:448-501)
        import numpy as np
        rng = np.random.default_rng(0)
        pts = np.stack([rng.uniform(-1, 1, 80), rng.uniform(-1, 1, 80), rng.uniform(-6, -2.5, 80)], 1)
        sio.write_colmap_model(scene, info["c2w"], info["focal"], (48, 64), pts, rng.uniform(0.2, 2.0, 80), rng)
    work = tempfile.mkdtemp(prefix="spn_work_")
    os.chdir(work)
    os.makedirs("lama/LaMa_test_images", exist_ok=True)
    # drop-in and import shims first, the reference's own directory last
    sys.path[:0] = [os.path.join(ROOT, "spin-nerf_b200", "dropin"), os.path.join(ROOT, "spin-nerf_b200", "compat")]
    sys.path.append("/root/reference/DS_NeRF")
    torch.cuda.set_device = lambda *a, **k: None            # run_nerf.py:39 needs a driver
    sys.argv = ["run_nerf.py", "--expname", "t", "--datadir", scene, "--basedir", os.path.join(work, "logs"),
                "--dataset_type", "llff", "--factor", "2", "--N_rand", "32", "--N_samples", "8", "--N_importance", "8",
                "--use_viewdirs", "--raw_noise_std", "1.0", "--no_ndc", "--lindisp", "--white_bkgd", "--no_tcnn", "--N_gt", "0",
                "--N_iters", str(iters), "--i_video", every, "--i_testset", every, "--i_weights", every,
                "--i_feat", "100000", "--i_print", "1", "--chunk", "512", "--netchunk", "4096"] + extra
    import run_nerf
    helpers = sys.modules["run_nerf_helpers"]
    torch.autograd.set_detect_anomaly(False)                # values are garbage here; anomaly mode would trip on NaNs
    run_nerf.train()
    logdir = os.path.join(work, "logs", "t")
    files = sorted(os.path.relpath(os.path.join(r, f), logdir) for r, _, fs in os.walk(logdir) for f in fs)
    ckpt_keys = {}
    for f in files:
        if f.endswith(".tar"):
            ck = torch.load(os.path.join(logdir, f), map_location="cpu", weights_only=False)
            ckpt_keys = {k: (sorted(v.keys()) if isinstance(v, dict) and k.startswith("network") else None) for k, v in ck.items()}
    print("SEAM " + json.dumps({"files": files, "ckpt": ckpt_keys, "helpers_file": helpers.__file__, "run_nerf_file": run_nerf.__file__,
                                "nerf_class_module": run_nerf.NeRF.__module__, "data_file": sys.modules["data"].__file__,
                                "load_llff_file": sys.modules["load_llff"].__file__, "calls": rec.names()}))


if __name__ == "__main__":
    main()
