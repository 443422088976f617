"""The only numbers the reference tree publishes (DS_NeRF/torchsearchsorted/README.md:58-88, for an op its hot path does not
call): batched searchsorted of 50000 x 1000 queries in 50000 x 300 sorted rows (0.391 ms on an unnamed GPU, second run) and
`examples/benchmark.py`'s 5000 x 300 / 5000 x 100 case.  Device-timed here for spinnerf_b200.ops.searchsorted (B200) next to
torch.searchsorted on the same tensors.   python tools/bench_searchsorted.py"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

spn = importlib.import_module("spin-nerf_b200")
dev = "cuda"


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rows, na, nv, published in ((50000, 300, 1000, "0.391 ms (README.md:72)"), (5000, 300, 100, "0.000796 s per 100 calls (README.md:88)")):
    a = torch.sort(torch.rand(rows, na, device=dev), dim=1)[0]
    v = torch.rand(rows, nv, device=dev)
    out = torch.empty(rows, nv, dtype=torch.long, device=dev)
    ours = timed(lambda: spn.ops.searchsorted(a, v, out))
    ref = timed(lambda: torch.searchsorted(a, v, out=out))
    assert torch.equal(spn.ops.searchsorted(a, v), torch.searchsorted(a, v))
    gb = (a.numel() * 4 + v.numel() * 4 + out.numel() * 8) / 1e9
    print(f"a[{rows}x{na}] v[{rows}x{nv}]: ours {ours:.3f} ms ({gb / ours * 1e3:.0f} GB/s algorithmic), torch.searchsorted {ref:.3f} ms; "
          f"reference publishes {published}")
