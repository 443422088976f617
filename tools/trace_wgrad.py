"""Per-CTA segment time of the wgrad kernel by unit (clock64; spn_tc_set_trace) — calibrates the CTA assignment."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 393216
net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net = net.seeded_init_(1).to(dev)
flat, packed = net._sync()
x6 = torch.randn(M, 6, device=dev); draw = torch.randn(M, 4, device=dev)
stash = spn.ops.mlp_stash(M, spn.PREC_BF16, dev); ws = spn.ops.mlp_bwd_workspace(M, spn.PREC_BF16, dev)
g = torch.zeros(spn.MLP_NPARAMS, device=dev)
spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, stash)
for _ in range(2):
    spn.ops.mlp_backward(flat, packed, stash, draw, g, spn.PREC_BF16, workspace=ws)
tr = torch.zeros(3 * 12 * 2 * 24, dtype=torch.int64, device=dev)
L.check(L.lib().spn_tc_set_trace(L.ptr(tr)))
spn.ops.mlp_backward(flat, packed, stash, draw, g, spn.PREC_BF16, workspace=ws)
torch.cuda.synchronize()
L.lib().spn_tc_set_trace(None)
t = tr.cpu().numpy()[:2 * 160].reshape(-1, 2)
t = t[t[:, 1] > 0]
print(f"M={M}: {len(t)} CTAs; max {t[:, 1].max()} cycles, mean {t[:, 1].mean():.0f} (balance {t[:, 1].mean() / t[:, 1].max():.2f})")
for u in range(12):
    c = t[t[:, 0] == u][:, 1]
    if len(c):
        print(f"unit {u:2d}: {len(c):3d} CTAs  cycles mean {c.mean():9.0f}  min {c.min():9d}  max {c.max():9d}")
