"""Generate tests/golden/render_variants.npz by running the UNMODIFIED reference render() (DS_NeRF/run_nerf.py:90-165) on
CPU for the call variants the main render golden does not cover (build container only):

    python tests/golden/make_variants_golden.py

  depths      rays + a depth column (12-column ray matrix, run_nerf.py:150-151)
  c2w_patch   rays generated from c2w inside render(), patch window (:117-123)
  staticcam   c2w_staticcam: rays from one camera, view directions from another (:128-133)
  rgb_net     network_fine = NeRF_RGB(alpha_model=...) (helpers:159-216): colour head over a frozen density provider
  no_coarse   network_fn=None: the coarse pass queries network_fine.alpha_model (run_nerf.py:680-686)
Weights come from oracle.init_params(seed) as in make_golden.py (11 coarse, 12 fine / density provider, 13 colour net).
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
HW = (12, 16, 14.4)


def poses():
    a = np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 0.0]], np.float32)
    c, s = np.cos(0.1), np.sin(0.1)
    b = np.array([[c, 0, s, -0.15], [0, 1, 0, 0.1], [-s, 0, c, 0.05]], np.float32)
    return a, b


def main():
    torch.set_num_threads(8)
    H_, R_ = ref_loader.load()
    pc, pf, pr = O.init_params(11), O.init_params(12), O.init_params(13)
    for p in (pc, pf):
        p["alpha_linear.bias"] = p["alpha_linear.bias"] + np.float32(1.0)
    netc, netf = ref_loader.reference_nets(pc, pf)
    rgb_net = H_.NeRF_RGB(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True, alpha_model=netf)
    sd = {k: T(v.copy()) for k, v in pr.items() if not k.startswith("alpha_linear")}
    sd.update({"alpha_model." + k: T(v.copy()) for k, v in pf.items()})
    rgb_net.load_state_dict(sd)
    e10, _ = H_.get_embedder(10, 0); e4, _ = H_.get_embedder(4, 0)
    qfn = lambda inputs, viewdirs, fn: R_.run_network(inputs, viewdirs, fn, embed_fn=e10, embeddirs_fn=e4, netchunk=65536)
    a, b = poses()
    H, W, f = HW
    ro, rd = H_.get_rays_np(H, W, f, a)
    rng = np.random.default_rng(77)
    sel = rng.choice(H * W, 40, replace=False)
    rays = np.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0).astype(np.float32)
    depths = rng.uniform(1, 8, 40).astype(np.float32)
    base = dict(chunk=32768, retraw=True, use_viewdirs=True, network_query_fn=qfn, N_samples=64, N_importance=64, ndc=False,
                lindisp=True, white_bkgd=True, perturb=0., raw_noise_std=0., near=1.2, far=8.0)
    cases = {
        "depths": dict(rays=T(rays), depths=T(depths), network_fn=netc, network_fine=netf),
        "c2w_patch": dict(c2w=T(a), patch=(3, 5, 6, 8), network_fn=netc, network_fine=netf),
        "staticcam": dict(c2w=T(a), c2w_staticcam=T(b), network_fn=netc, network_fine=netf),
        "rgb_net": dict(rays=T(rays), network_fn=netc, network_fine=rgb_net, need_alpha=True),
        "no_coarse": dict(rays=T(rays), network_fn=None, network_fine=rgb_net),
    }
    out = {"rays": rays, "depths": depths, "pose_a": a, "pose_b": b}
    for tag, kw in cases.items():
        with torch.no_grad():
            rgb, disp, acc, depth, ex = R_.render(H, W, f, **dict(base, **kw))
        out.update({f"{tag}__rgb": rgb, f"{tag}__disp": disp, f"{tag}__acc": acc, f"{tag}__depth": depth})
        for k, v in ex.items():
            if tag in ("c2w_patch", "staticcam") and k in ("raw", "weights", "z_vals"):
                continue                      # per-sample arrays of whole frames: the maps pin these cases
            out[f"{tag}__{k}"] = v
        print(tag, tuple(rgb.shape), sorted(ex))
    np.savez_compressed(os.path.join(OUT, "render_variants.npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
