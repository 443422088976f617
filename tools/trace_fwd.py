"""Timeline of CTA 0 of the fused forward kernel (clock64 stamps, spn_tc_set_trace).
usage: python tools/trace_fwd.py [M] [train]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
train = len(sys.argv) > 2 and sys.argv[2] == "train"
dev = "cuda"
net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
net = net.seeded_init_(1).to(dev)
x6 = torch.randn(M, 6, device=dev)
flat, packed = net._sync()
stash = spn.ops.mlp_stash(M, spn.PREC_BF16, dev) if train else None
for _ in range(2):
    spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, stash)
tr = torch.zeros(3 * 12 * 2 * 24, dtype=torch.int64, device=dev)
L.check(L.lib().spn_tc_set_trace(L.ptr(tr)))
spn.ops.mlp_forward_points(flat, packed, x6, spn.PREC_BF16, stash)
torch.cuda.synchronize()
L.lib().spn_tc_set_trace(None)
t = tr.cpu().numpy().reshape(3, 12, 2, 24)
t0 = t[t > 0].min()
names = ["mma:act_ready", "mma:w_first", "mma:w_last", "mma:issued", "epi:acc_ready", "epi:synced", "epi:done0", "epi:done255"]
print("round step tile | " + " ".join(f"{n:>8s}" for n in ["act_rdy", "lf_ok", "plf_ok", "issued", "acc_rdy", "p_accrdy", "epi_done", "p_epidone"]) + " |  epi_len  mma_wait_w")
for it in range(3):
    for s in range(12):
        for tl in range(2):
            e = t[it, s, tl]
            rel = [(int(x - t0) if x > 0 else -1) for x in e]
            epi = rel[6] - rel[4] if rel[6] > 0 and rel[4] > 0 else -1
            ww = rel[2] - rel[0]
            blk = " ".join(f"{rel[8 + i] - rel[0]:5d}" for i in range(4) if rel[8 + i] > 0)
            prod = " ".join(f"{rel[12 + i] - rel[0]:6d}" for i in range(4) if rel[12 + i] > 0)
            print(f"{it:5d} {s:4d} {tl:4d} | " + " ".join(f"{x:8d}" for x in rel[:8]) + f" | {epi:6d} {ww:6d} | chunk ready (rel act_ready): {blk} | loads issued: {prod} | epi internals (rel acc_rdy): " + " ".join(f"{rel[16 + i] - rel[4]:5d}" for i in range(8) if rel[16 + i] > 0))
print("cycles per pair-round:", [int(t[i + 1, 0, 0, 0] - t[i, 0, 0, 0]) for i in range(2)])
