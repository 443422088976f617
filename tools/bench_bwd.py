"""Device-timed MLP backward kernels (dgrad / wgrad) for several sample counts (in-library CUDA events)."""
import ctypes
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
dev = "cuda"
net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net = net.seeded_init_(1).to(dev)
flat, packed = net._sync()
for M in [int(a) for a in sys.argv[1:]] or [8192, 65536, 131072, 1048576]:
    x6 = torch.randn(M, 6, device=dev); draw = torch.randn(M, 4, device=dev)
    stash = spn.ops.mlp_stash(M, spn.PREC_BF16, dev)
    ws = spn.ops.mlp_bwd_workspace(M, spn.PREC_BF16, dev)
    g = torch.zeros(spn.MLP_NPARAMS, device=dev)
    spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, stash)
    reps = 5
    for _ in range(2):
        spn.ops.mlp_backward(flat, packed, stash, draw, g, spn.PREC_BF16, workspace=ws)
    torch.cuda.synchronize()
    L.lib().spn_profile_enable(1)
    for _ in range(reps):
        spn.ops.mlp_backward(flat, packed, stash, draw, g, spn.PREC_BF16, workspace=ws)
    torch.cuda.synchronize()
    out = []
    for kind, name, flop in ((1, "dgrad", 2 * 557696), (2, "wgrad", 2 * 593408)):
        n, ms = ctypes.c_int(), ctypes.c_float()
        L.lib().spn_profile_read(kind, ctypes.byref(n), ctypes.byref(ms))
        t = ms.value / max(n.value, 1)
        out.append(f"{name} {t:8.3f} ms {M * flop / t / 1e9:7.1f} TF")
    L.lib().spn_profile_enable(0)
    print(f"M={M:8d} tiles={M // 128:5d}  " + "   ".join(out) + f"   stash {stash.numel() / 2**20:7.1f} MiB dstash+partials {ws.numel() / 2**20:7.1f} MiB")
    del stash, ws, x6, draw
