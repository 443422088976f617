import importlib, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.chdir(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import test_gpu_parity as T
spn = T.spn; ops = T.ops
net, p = T.make_net(11, spn.PREC_BF16)
g, x90, x6 = T._mlp_case()
m = x6.shape[0]
flat, packed = net._sync()
stash = ops.mlp_stash(m, spn.PREC_BF16, T.DEV); stash.zero_()
raw, _ = ops.mlp_forward_points(flat, packed, T.T(x6), spn.PREC_BF16, stash)
torch.cuda.synchronize()
emu, acts = T.mlp_forward_bf16(p, x6)
ntiles = (m + 127) // 128
TB = T.STASH_TILE_BYTES
names = [("xp", T.SA_ENC, 1, None)] + [(f"h{i}", None, None, i) for i in range(8)] + [("feat", None, None, 8), ("hv", T.SA_HV, 2, None), ("xd", T.SA_DENC, 1, None)]
for nm, a0, na, layer in names:
    got = T.decode_tiles(stash, TB, a0, na, ntiles, m) if layer is None else T.decode_x8(stash, layer, ntiles, m)
    ref = acts[nm]
    w = ref.shape[1]
    e = np.abs(got[:, :w] - ref)
    halves = [e[:, :w // 2].max(), e[:, w // 2:].max()]
    print(f"{nm:5s} max err {e.max():.4f}  cols lo/hi {halves[0]:.4f} {halves[1]:.4f}  rows T0/T1 {e[:128].max():.4f} {e[128:].max():.4f}  scale {np.abs(ref).max():.3f}")
print("raw err", np.abs(T.N(raw) - emu).max(0))
