"""Micro-benchmark of the fused MLP kernels alone (device-timed). usage: python tools/bench_mlp.py [M] [train]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
spn = importlib.import_module("spin-nerf_b200")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
train = len(sys.argv) > 2 and sys.argv[2] == "train"
dev = "cuda"
net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net = net.seeded_init_(1).to(dev)
x6 = torch.randn(M, 6, device=dev)
flat, packed = net._sync()
stash = spn.ops.mlp_stash(M, spn.PREC_BF16, dev) if train else None
for _ in range(3):
    raw, _ = spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, stash)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    raw, _ = spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, stash)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = M * 1186816
print(f"mlp_fwd bf16 M={M} train={train}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s  {M/ms/1e3:.1f} Msamples/s")
