"""Timeline of CTA 0 of the TS forward kernel (clock64 stamps, spn_tc_set_trace).  usage: python tools/trace_fwd_ts.py [M] [train]
Per (round, layer's last step, tile slot): MMA thread: batch start / all MMAs issued for half a and b; epilogue warp 4 lane 0:
accumulator ready (a, b), half a packed, a_ready arrive."""
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
hdr = ["mma_a", "mma_b", "iss_a", "iss_b", "acc_a", "acc_b", "pack_a", "a_rdy"]
print("round step slot | " + " ".join(f"{n:>8s}" for n in hdr) + " | a: wait->issue  issue->acc   b: wait->issue  issue->acc   acc_b->a_rdy")
for it in range(3):
    for s in range(12):
        for tl in range(2):
            e = t[it, s, tl]
            if not (e[:8] > 0).any():
                continue
            rel = [(int(x - t0) if x > 0 else -1) for x in e[:8]]
            d = lambda a, b: (rel[b] - rel[a]) if rel[a] >= 0 and rel[b] >= 0 else -1
            ex = [(int(x - t0) if x > 0 else -1) for x in e[8:14]]
            r2 = rel + ex
            dd = lambda a, b: (r2[b] - r2[a]) if r2[a] >= 0 and r2[b] >= 0 else -1
            print(f"{it:5d} {s:4d} {tl:4d} | " + " ".join(f"{x:8d}" for x in rel) +
                  f" | {d(0, 2):8d} {d(2, 4):8d}   {d(1, 3):8d} {d(3, 5):8d}   {d(5, 7):8d}"
                  f" | half a: ld+free {dd(4, 9)} math {dd(9, 6)} | half b: st+ld+free {dd(5, 12)} math {dd(12, 13)} tail {dd(13, 7)}")
starts = [int(t[i, 0, 0, 0]) for i in range(3)]
print("cycles per pair-round:", [starts[i + 1] - starts[i] for i in range(2)])
