"""Helper of tests/test_bench_dry_run.py (run as a subprocess): bench.py's own arm and its two other workloads executed on CPU
tensors with the C library replaced by the call recorder of tests/test_host_glue_dry_run.py and the CUDA runtime surface that
bench.py touches (events, synchronize, set_device, pinned memory, the "cuda" device, the CUDA-graph step) replaced by host
stand-ins, on a shrunken frame size.  Numbers are meaningless; what is checked is that every line of the benchmark's Python
runs and that the JSON line carries the contract's keys — the part of `bench.py` a build container without a GPU can break."""
import argparse
import sys
import time

import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                    # noqa: E402
import test_host_glue_dry_run as dry            # noqa: E402
import bench                                    # noqa: E402


def install_shims():
    rec = dry.Recorder()
    for m in (dry.L, dry.ops, dry.render_mod):
        m.lib = (lambda rec=rec: rec); m.ptr = dry._ptr; m.stream = (lambda: 0)
    dry.ops._empty = lambda shape, like, dtype=torch.float32: torch.zeros(shape, device=like.device, dtype=dtype)
    dry.nerf_mod.NeRF._sync = lambda self: (self.flat_params(), torch.zeros(64, dtype=torch.uint8))

    class Event:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self, *a):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3
    torch.cuda.Event = Event
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.set_device = lambda *a: None
    bench._device = lambda local: torch.device("cpu")
    torch.Tensor.pin_memory = lambda self: self
    real_empty = torch.empty
    torch.empty = lambda *a, **k: real_empty(*((1024,) if a and a[0] == 64 * 1024 * 1024 else a), **k)     # the L2-flush buffer

    def prof_read(kind, n, ms):
        n._obj.value, ms._obj.value = 2, 1.0
        return 0
    rec_get = dry.Recorder.__getattr__
    dry.Recorder.__getattr__ = lambda self, name: prof_read if name == "spn_profile_read" else rec_get(self, name)
    dry.trainer_mod.Trainer.step_graphed = lambda self, *ins: self.step(*ins)      # no CUDA graphs on the host
    bench._init_dist = lambda: (0, 1, 0, torch.device("cpu"))
    bench.H, bench.W, bench.FOCAL = 48, 64, 57.6


def main():
    install_shims()
    which = sys.argv[1]
    if which == "train":
        sys.argv = ["bench.py", "--steps", "2", "--warmup", "3", "--n_rand", "16", "--cpu_rays", "8", "--deadline", "0", "--no_other_workloads"]
        bench.hbm_write_gbs = lambda dev: 1000.0          # the live memset probe allocates 2 GiB
        bench.main()
    else:
        bench.run_other_workload(argparse.Namespace(workload=which, precision="bf16", warmup=1, steps=2, n_rand=16))


if __name__ == "__main__":
    main()
