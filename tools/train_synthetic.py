"""End-to-end example on a synthetic SPIn-NeRF-shaped scene: write the scene to disk (scene_io.write_scene layout), load it
(scene_io.load_scene), build the resident ray pools (raypool.build_ray_pools), upload once and train with the fused step
(Trainer.step_from_pool: device-side batch assembly, one chunk per step, tcgen05 kernels, flat Adam).

    python tools/train_synthetic.py --steps 200            # needs a B200; prints loss / PSNR / rays per second
    python tools/train_synthetic.py --steps 400 --lpips --video out/   # + the perceptual-loss branch after `--lpips_from` steps
                                                           #   (Trainer.step_with_lpips) and a spiral video (render_path -> mp4)
    python tools/train_synthetic.py --dry_run --lpips      # CPU only: everything up to (not including) the first GPU step
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
    ap.add_argument("--lpips", action="store_true", help="perceptual-loss branch (run_nerf.py:1523-1561) with the stand-in LPIPS")
    ap.add_argument("--lpips_from", type=int, default=300, help="first step of the LPIPS branch (reference: i > 300)")
    ap.add_argument("--video", default=None, help="directory for a render_path video of the scene's spiral poses after training")
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
    sampler = None
    if args.lpips:       # masks of every view (the reference indexes them by training-view id), targets resized once
        lp = importlib.import_module("spin-nerf_b200.lpips_patch")
        sampler = lp.PatchSampler(hwf, masks != 0, images, i_train, lpips_render_factor=2, patch_len_factor=8, lpips_batch_size=4,
                                  device=None if args.dry_run else "cuda")
        print(f"lpips branch: patches of {sampler.patch_len[0]}x{sampler.patch_len[1]} on a {sampler.Hs}x{sampler.Ws} grid, "
              f"{sampler.batch_size} per step from step {args.lpips_from}")
    if args.dry_run:
        dev_pools = pools.to("cpu")
        idx = rp.draw_step_indices(dev_pools, args.n_rand)
        print("dry run: first step would use indices", tuple(idx.shape), "of a pool", tuple(dev_pools["pool_od"].shape))
        if sampler is not None:
            views, Xs, Ys = sampler.sample()
            print("dry run: first LPIPS step would render views", views, "at origins", list(zip(Xs, Ys)),
                  "against targets", [tuple(t.shape) for t in sampler.target_patches(views, Xs, Ys)])
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
    lpips_fn = importlib.import_module("spin-nerf_b200.compat.lpips").LPIPS().to(dev) if args.lpips else None
    t0 = None
    for step in range(1, args.steps + 1):
        idx = rp.draw_step_indices(dev_pools, args.n_rand)
        batch = (dev_pools["pool_od"], dev_pools["rgb"], dev_pools["disp"], idx)
        if sampler is not None and step > args.lpips_from:
            views, Xs, Ys = sampler.sample()
            patches = [(x, y) + tuple(sampler.patch_len) for x, y in zip(Xs, Ys)]
            loss, psnr = tr.step_with_lpips(batch, [poses[v, :3, :4] for v in views], patches, sampler.target_patches(views, Xs, Ys),
                                            lpips_fn, (sampler.Hs, sampler.Ws, sampler.focal_s), from_pool=True)
        else:
            loss, psnr = tr.step_from_pool(*batch)
        if step == 10:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        if step % 50 == 0 or step == args.steps:
            print(f"step {step:5d}  loss {float(loss):.4f}  psnr {float(psnr):.2f}")
    torch.cuda.synchronize()
    if t0 is not None and args.steps > 10:
        dt = time.perf_counter() - t0
        print(f"{3 * args.n_rand * (args.steps - 10) / dt / 1e3:.0f} k rays/s over {args.steps - 10} steps (wall clock, incl. index draws)")
    if args.video:       # run_nerf.py:1638-1671: spiral poses through render_path (async frame sink), mp4 export
        os.makedirs(args.video, exist_ok=True)
        kw = dict(network_query_fn=None, network_fn=nets[0], network_fine=nets[1], N_samples=64, N_importance=64, lindisp=True,
                  white_bkgd=True, perturb=0., raw_noise_std=0., use_viewdirs=True, ndc=False, near=near, far=far)
        t0 = time.perf_counter()
        rgbs, disps, _ = spn.render_path(render_poses[::4], list(hwf), 32768, kw, savedir=args.video, need_alpha=True)
        fio = spn.frame_io
        fio.write_video(os.path.join(args.video, "rgb.mp4"), rgbs)
        fio.write_video(os.path.join(args.video, "disp.mp4"), disps / np.nanmax(disps))
        print(f"video: {len(rgbs)} frames of {rgbs.shape[1]}x{rgbs.shape[2]} rendered and written to {args.video} in {time.perf_counter() - t0:.1f} s")


if __name__ == "__main__":
    main()
