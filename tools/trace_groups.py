"""Weight-group timing of the CTA-pair forward kernel on a common clock (globaltimer, ns): when the leader issued the
last load of groups B / C, when the PEER saw its copy land, when the leader's wait for the group returned."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nerf_oracle as O
spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
M = 1 << 20
dev = "cuda"
net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in O.init_params(1).items()}); net = net.to(dev)
x6 = torch.randn(M, 6, device=dev)
flat, packed = net._sync()
for _ in range(2):
    spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, None)
tr = torch.zeros(3 * 12 * 2 * 24, dtype=torch.int64, device=dev)
L.check(L.lib().spn_tc_set_trace(L.ptr(tr)))
spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, None)
torch.cuda.synchronize()
L.lib().spn_tc_set_trace(None)
t = tr.cpu().numpy().reshape(3, 12, 2, 24)
it = 1
print("step | T0 act_rdy(ns, rel) | j3 issued | peer saw A,B,C land | leader B-wait done | T1 act_rdy | j5 issued | leader C-wait done")
base = t[it, 1, 0, 5]
for s in range(1, 10):
    r0, r1 = t[it, s, 0], t[it, s, 1]
    f = lambda x: int(x - base) if x > 0 else -1
    print(f"{s:4d} | {f(r0[5]):8d} | {f(r1[1]):8d} | {f(r1[12]):8d} {f(r1[13]):8d} {f(r1[14]):8d} | {f(r0[7]):8d} | {f(r1[5]):8d} | {f(r1[15]):8d} | {f(r1[7]):8d}")
