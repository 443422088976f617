"""End-to-end example on a synthetic SPIn-NeRF-shaped scene: write the scene to disk (scene_io.write_scene layout), load it
(scene_io.load_scene), build the resident ray pools (raypool.build_ray_pools), upload once and train with the fused step
(Trainer.step_from_pool: device-side batch assembly, one chunk per step, tcgen05 kernels, flat Adam).

    python tools/train_synthetic.py --steps 200            # needs a B200; prints loss / PSNR / rays per second
    python tools/train_synthetic.py --dry_run              # CPU only: everything up to (not including) the first GPU step
"""
import argparse
import importlib
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default=None, help="scene directory (default: a fresh temporary synthetic scene)")
    ap.add_argument("--views", type=int, default=12)
    ap.add_argument("--height", type=int, default=96)
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--factor", type=int, default=2)
    ap.add_argument("--n_rand", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--dry_run", action="store_true")
    args = ap.parse_args()
    sio = importlib.import_module("spin-nerf_b200.scene_io")
    rp = importlib.import_module("spin-nerf_b200.raypool")
    scene = args.scene or tempfile.mkdtemp(prefix="spn_scene_")
    if args.scene is None:
        sio.synthetic_scene(scene, n_views=args.views, hw=(args.height, args.width), factor=args.factor, seed=0, n_unlabelled=1)
    images, poses, bds, render_poses, i_test, masks, depths, mask_indices = sio.load_scene(scene, factor=args.factor, lpips=True)
    hwf = (int(poses[0, 0, 4]), int(poses[0, 1, 4]), float(poses[0, 2, 4]))
    kept = len(images) - 5                                           # the view that keeps mask label +1 (load_llff.py:161)
    i_train = [i for i in range(len(images)) if i != i_test or i == kept]
    pools = rp.build_ray_pools(images, poses, hwf, masks, depths, i_train)
    near, far = float(bds.min() * .9), float(bds.max() * 1.)          # run_nerf.py:1006-1007 (no_ndc)
    print(f"scene {scene}: {len(images)} views {hwf[0]}x{hwf[1]}, hold-out {i_test}, pool {len(pools.label)} rays "
          f"(unmasked {len(pools.idx_clf)}, masked {len(pools.idx_rgb)}, inpainted {len(pools.idx_inp)}), near/far {near:.3f}/{far:.3f}")
    if args.dry_run:
        dev_pools = pools.to("cpu")
        idx = rp.draw_step_indices(dev_pools, args.n_rand)
        print("dry run: first step would use indices", tuple(idx.shape), "of a pool", tuple(dev_pools["pool_od"].shape))
        return
    spn = importlib.import_module("spin-nerf_b200")
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    nets = []
    for _ in range(2):                                               # coarse, fine: nn.Linear default init like create_nerf
        net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).to(dev)
        net.precision = spn.PREC_BF16
        nets.append(net)
    tr = trainer_mod.Trainer(nets[0], nets[1], lr=5e-4, N_samples=64, N_importance=64, lindisp=True, white_bkgd=True,
                             perturb=1.0, raw_noise_std=1.0, near=near, far=far, ndc=False, hwf=hwf)
    dev_pools = pools.to(dev)
    t0 = None
    for step in range(1, args.steps + 1):
        idx = rp.draw_step_indices(dev_pools, args.n_rand)
        loss, psnr = tr.step_from_pool(dev_pools["pool_od"], dev_pools["rgb"], dev_pools["disp"], idx)
        if step == 10:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        if step % 50 == 0 or step == args.steps:
            print(f"step {step:5d}  loss {float(loss):.4f}  psnr {float(psnr):.2f}")
    torch.cuda.synchronize()
    if t0 is not None and args.steps > 10:
        dt = time.perf_counter() - t0
        print(f"{3 * args.n_rand * (args.steps - 10) / dt / 1e3:.0f} k rays/s over {args.steps - 10} steps (wall clock, incl. index draws)")


if __name__ == "__main__":
    main()
