"""Helper of tests/test_raypool.py::test_run_nerf_fused_cli_loop (run as a subprocess): tools/run_nerf_fused.py's whole loop
on CPU tensors with the C library replaced by the call recorder of tests/test_host_glue_dry_run.py (zero-filled outputs) —
checks the Python of the loop (sampling, sparse-depth group, LPIPS branch, checkpoints, resume, video / test-set renders),
not numerics.  argv: scene dir, log dir, extra flags...  Prints 'CLI <json>'."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                    # noqa: E402
import test_host_glue_dry_run as dry            # noqa: E402


def main():
    scene, logs, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
    rec = dry.Recorder()
    for m in (dry.L, dry.ops, dry.render_mod):
        m.lib = (lambda rec=rec: rec); m.ptr = dry._ptr; m.stream = (lambda: 0)
    dry.ops._empty = lambda shape, like, dtype=torch.float32: torch.zeros(shape, device=like.device, dtype=dtype)
    dry.nerf_mod.NeRF._sync = lambda self: (self.flat_params(), torch.zeros(64, dtype=torch.uint8))
    spec = importlib.util.spec_from_file_location("run_nerf_fused", os.path.join(ROOT, "tools", "run_nerf_fused.py"))
    cli = importlib.util.module_from_spec(spec); spec.loader.exec_module(cli)
    rc = cli.main(["--expname", "t", "--basedir", logs, "--datadir", scene, "--factor", "2", "--N_rand", "16", "--N_samples", "8",
                   "--N_importance", "8", "--use_viewdirs", "--raw_noise_std", "1.0", "--no_ndc", "--lindisp", "--white_bkgd",
                   "--no_tcnn", "--device", "cpu", "--chunk", "256", "--i_print", "1"] + extra)
    logdir = os.path.join(logs, "t")
    files = sorted(os.path.relpath(os.path.join(r, f), logdir) for r, _, fs in os.walk(logdir) for f in fs)
    names = rec.names()
    print("CLI " + json.dumps({"rc": rc, "files": files, "counts": {k: names.count(k) for k in sorted(set(names))}}))


if __name__ == "__main__":
    main()
