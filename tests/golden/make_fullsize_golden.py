"""Generate tests/golden/fullsize_step.npz: ONE train step of the UNMODIFIED reference at BASELINE configs[1] size — three render()
calls of N_rand = 1024 rays each (run_nerf.py:1455-1470) on the LLFF-shaped 1008x756 camera of bench.py, coarse + fine 64 + 64
samples, lindisp / white_bkgd / use_viewdirs, deterministic sampling — with its losses and autograd parameter gradients
(build container only, ~1 minute of CPU):

    python tests/golden/make_fullsize_golden.py

Stored: the 3 x 1024 rays and targets, every rendered map of the three calls, loss / psnr, and per parameter tensor of both
networks sum|g| (float64) and every 997th gradient entry.  Weights are oracle.init_params(seed) as in the other goldens.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O      # noqa: E402
from oracle import ref_loader            # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
T = torch.from_numpy
H, W = 756, 1008
FOCAL = 0.9 * W
NEAR, FAR = 1.2, 8.0
N_RAND = 1024
SEED_C, SEED_F = 31, 32


def poses(n, seed=0):          # bench.py's camera arc
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        z = np.array([0.1, -0.05, 1.0]) + rng.standard_normal(3) * 0.05
        z /= np.linalg.norm(z)
        x = np.cross([0, 1, 0], z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        pos = np.array([rng.uniform(-.5, .5), rng.uniform(-.5, .5), 0.0])
        out.append(np.stack([x, y, z, pos], 1).astype(np.float32))
    return out


def problem():
    rng = np.random.default_rng(77)
    rays = []
    for c2w in poses(3):
        ro, rd = O.get_rays(H, W, FOCAL, c2w)
        sel = rng.choice(H * W, N_RAND, replace=False)
        rays.append(np.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0).astype(np.float32))
    target_clf = rng.random((N_RAND, 3), dtype=np.float32)
    target_s = rng.random((N_RAND, 3), dtype=np.float32)
    depth_inp = (0.2 + 0.4 * rng.random((N_RAND,), dtype=np.float32)).astype(np.float32)
    return rays, target_clf, target_s, depth_inp


def params():
    pc, pf = O.init_params(SEED_C), O.init_params(SEED_F)
    for p in (pc, pf):
        p["alpha_linear.bias"] = p["alpha_linear.bias"] + np.float32(1.0)      # so that transmittance actually decays
    return pc, pf


def main():
    torch.set_num_threads(8)
    H_, R_ = ref_loader.load()
    torch.autograd.set_detect_anomaly(False)
    rays, target_clf, target_s, depth_inp = problem()
    netc, netf = ref_loader.reference_nets(*params())
    e10, _ = H_.get_embedder(10, 0); e4, _ = H_.get_embedder(4, 0)
    qfn = lambda inputs, viewdirs, fn: R_.run_network(inputs, viewdirs, fn, embed_fn=e10, embeddirs_fn=e4, netchunk=65536)
    kw = dict(chunk=32768, retraw=True, use_viewdirs=True, network_query_fn=qfn, network_fn=netc, network_fine=netf, N_samples=64,
              N_importance=64, ndc=False, lindisp=True, white_bkgd=True, perturb=0., raw_noise_std=0., near=NEAR, far=FAR)
    out = {}
    rgb, disp, acc, depth, ex = R_.render(H, W, FOCAL, rays=T(rays[0]), **kw)
    rgb_c, disp_c, acc_c, depth_c, ex_c = R_.render(H, W, FOCAL, rays=T(rays[1]), detach_weights=True, **kw)
    rgb_i, disp_i, acc_i, depth_i, ex_i = R_.render(H, W, FOCAL, rays=T(rays[2]), **kw)
    for tag, (a, b, c, d, e) in {"clf": (rgb, disp, acc, depth, ex), "s": (rgb_c, disp_c, acc_c, depth_c, ex_c),
                                 "inp": (rgb_i, disp_i, acc_i, depth_i, ex_i)}.items():
        for k, v in dict(rgb=a, disp=b, acc=c, depth=d, rgb0=e["rgb0"], disp0=e["disp0"], acc0=e["acc0"], z_std=e["z_std"]).items():
            out[f"{tag}__{k}"] = v.detach().numpy()
        out[f"{tag}__weights_sub"] = e["weights"].detach().numpy()[::16]
        out[f"{tag}__z_vals_sub"] = e["z_vals"].detach().numpy()[::16]
    img_loss = H_.img2mse(rgb, T(target_clf))
    psnr = H_.mse2psnr(img_loss)
    img_loss = img_loss + H_.img2mse(rgb_c, T(target_s)) + H_.img2mse(ex_c["rgb0"], T(target_s))
    loss = img_loss + H_.img2mse(ex["rgb0"], T(target_clf))
    inp_loss = torch.nn.MSELoss()(disp_i, T(depth_inp)) + torch.nn.MSELoss()(ex_i["disp0"], T(depth_inp))
    assert not inp_loss.isnan()
    loss = loss + inp_loss
    loss.backward()
    for tag, net in (("c", netc), ("f", netf)):
        for k, v in net.named_parameters():
            g = v.grad.numpy()
            out[f"g_abs_{tag}__{k}"] = np.abs(g).sum(dtype=np.float64)
            out[f"g_sub_{tag}__{k}"] = g.reshape(-1)[::997].copy()
    np.savez_compressed(os.path.join(OUT, "fullsize_step.npz"), rays=np.stack(rays, 0), target_clf=target_clf, target_s=target_s,
                        depth_inp=depth_inp, loss=np.float64(loss.item()), psnr=np.float64(psnr.item()),
                        cfg=np.array([H, W, FOCAL, NEAR, FAR, N_RAND, SEED_C, SEED_F], np.float64), **out)
    print("wrote fullsize_step.npz: loss", loss.item(), "psnr", psnr.item(), "acc mean", float(acc.mean()))


if __name__ == "__main__":
    main()
