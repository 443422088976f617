#!/usr/bin/env python
"""The SPIn-NeRF trainer on the fused B200 path, driven by the reference's own flags and config files.

    python tools/run_nerf_fused.py --config DS_NeRF/configs/config.txt --datadir data/statue --expname statue --factor 2 \\
        --N_gt 0 --N_iters 10001 [--lpips] [--prepare] [--render_only]

`DS_NeRF/run_nerf.py --no_tcnn` runs unmodified over the drop-in module; this is the other way to switch:
the same stages of `train()` (run_nerf.py:963-1703) for the flags of the hot path, with every stage on the fused path —
scene_io.load_scene (load_llff_data), resident ray pools + device-side sampling instead of four DataLoaders
(raypool, Trainer.step_from_pool), the step's three or four render calls as ONE chunk with analytic losses and the flat Adam
(Trainer.step), the `--lpips` branch (Trainer.step_with_lpips), checkpoints in the reference's layout (Trainer.checkpoint:
either trainer resumes the other's run), videos / test renders through render_path's asynchronous frame sink, and with
`--prepare` (stage A of the pipeline, README.md:58-67) the hand-over to the inpainter: every view's rendered disparity and
mask as `<lama_dir>/img%03d.png`, `<lama_dir>/label/img%03d.png` every `i_feat` iterations (run_nerf.py:1563-1609).
Under torchrun (one process per GPU) N_rand stays the GLOBAL batch of the reference: every rank draws the same indices and
renders its contiguous 1/W of each ray group, gradients are averaged over NVLink, videos are rendered frame-sharded; rank 0
logs, checkpoints and writes files.
Flags keep the reference's names and defaults (config_parser, run_nerf.py:740-925); flags outside this path
(--sigma_loss, tcnn, blender / DTU data, object removal, ...) are rejected rather than ignored.
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402


def config_parser():
    try:
        import configargparse
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "spin-nerf_b200", "compat"))
        import configargparse
    p = configargparse.ArgumentParser()
    A = p.add_argument
    A('--config', is_config_file=True, help='config file path')
    A("--expname", type=str); A("--basedir", type=str, default='./logs/'); A("--datadir", type=str, default='./data/llff/fern')
    A("--dataset_type", type=str, default='llff'); A("--factor", type=int, default=8); A("--llffhold", type=int, default=1000000)
    A("--N_rand", type=int, default=32 * 32 * 4); A("--N_samples", type=int, default=64); A("--N_importance", type=int, default=0)
    A("--N_iters", type=int, default=10001); A("--N_gt", type=int, default=0)
    A("--lrate", type=float, default=5e-4); A("--lrate_decay", type=int, default=250)
    A("--chunk", type=int, default=1024 * 32); A("--netchunk", type=int, default=1024 * 64)
    A("--perturb", type=float, default=1.); A("--raw_noise_std", type=float, default=0.)
    A("--use_viewdirs", action='store_true'); A("--white_bkgd", action='store_true'); A("--lindisp", action='store_true')
    A("--no_ndc", action='store_true'); A("--no_tcnn", action='store_true'); A("--prepare", action='store_true')
    A("--colmap_depth", action='store_true'); A("--depth_loss", action='store_true'); A("--depth_lambda", type=float, default=0.1)
    A("--lpips", action='store_true'); A("--lpips_render_factor", type=int, default=2); A("--patch_len_factor", type=int, default=8)
    A("--lpips_batch_size", type=int, default=4)
    A("--lpips_standin", action='store_true', help="allow spin-nerf_b200/compat/lpips.py (seeded random conv stack) when the lpips package is missing")
    A("--lpips_from", type=int, default=300, help="the LPIPS branch runs on iterations > this (hard-coded 300 in run_nerf.py:1523)")
    A("--no_reload", action='store_true'); A("--ft_path", type=str, default=None)
    A("--render_only", action='store_true'); A("--render_factor", type=int, default=0)
    A("--i_print", type=int, default=100); A("--i_weights", type=int, default=10000); A("--i_video", type=int, default=50000)
    A("--i_testset", type=int, default=50000); A("--i_feat", type=int, default=2000); A("--feat_weight", type=float, default=0.1)
    A("--precision", type=str, default="bf16", help="bf16 (tcgen05) | fp32 (CUDA cores, tight-parity mode)")
    A("--lama_dir", type=str, default='lama/LaMa_test_images',
      help="where --prepare leaves the rendered disparities and masks for the inpainter (run_nerf.py:1598-1609, relative to the cwd)")
    A("--dry_run", action='store_true', help="stop before the first GPU call (CPU-only check of config / data / pools)")
    A("--device", type=str, default="cuda:0", help="CUDA device (the library has no CPU implementation; tests drive the loop "
                                                    "with a call recorder in its place)")
    return p


def main(argv=None):
    args, unknown = config_parser().parse_known_args(argv)
    if unknown:
        raise SystemExit(f"run_nerf_fused: flags outside the fused hot path: {unknown}")
    if args.dataset_type != 'llff' or not args.use_viewdirs or args.N_importance <= 0:
        raise SystemExit("run_nerf_fused: the fused path is the LLFF / use_viewdirs / coarse+fine configuration")
    spn = importlib.import_module("spin-nerf_b200")
    sio, rp = importlib.import_module("spin-nerf_b200.scene_io"), importlib.import_module("spin-nerf_b200.raypool")
    lp = importlib.import_module("spin-nerf_b200.lpips_patch")
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
    dist_mod = importlib.import_module("spin-nerf_b200.dist")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not args.dry_run:
        torch.cuda.set_device(local)
        args.device = f"cuda:{local}"
        dist_mod.init_from_env()
    __import__("random").seed(0); np.random.seed(0)          # every rank must draw the same LPIPS views / patch origins
    say = print if rank == 0 else (lambda *a, **k: None)

    # ---- data (run_nerf.py:978-1012, 1225-1329)
    images, poses, bds, render_poses, i_test, masks, depths, mask_indices = sio.load_scene(
        args.datadir, factor=args.factor, prepare=args.prepare, lpips=args.lpips)
    hwf = (int(poses[0, 0, 4]), int(poses[0, 1, 4]), float(poses[0, 2, 4]))
    i_test = list(np.arange(images.shape[0])[::args.llffhold]) if args.llffhold > 0 else [int(i_test)]
    i_train = list(range(images.shape[0]))
    if args.N_gt > 0:        # the first N_gt views are held-out ground truth (run_nerf.py:1030-1035; the README runs use --N_gt 0)
        i_test, i_train = i_train[:args.N_gt], i_train[args.N_gt:]
        if not i_train:
            raise SystemExit(f"run_nerf_fused: --N_gt {args.N_gt} leaves no training view out of {images.shape[0]}")
    if args.no_ndc:
        near, far = float(np.ndarray.min(bds) * .9), float(np.ndarray.max(bds) * 1.)
    else:
        near, far = 0., 1.
    pools = rp.build_ray_pools(images, poses, hwf, masks, depths, i_train, prepare=args.prepare)
    depth_pool = None
    if args.colmap_depth and args.depth_loss:
        gts = sio.colmap_depth_rays(args.datadir, factor=args.factor, bd_factor=.75)
        depth_pool = rp.sparse_depth_rays(gts, poses, hwf, masks, i_train, prepare=args.prepare)
    sampler = None
    if args.lpips:
        sampler = lp.PatchSampler(hwf, masks != 0, images, i_train, args.lpips_render_factor, args.patch_len_factor,
                                  args.lpips_batch_size, device=None if args.dry_run else args.device)
    logdir = os.path.join(args.basedir, args.expname)
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, 'args.txt'), 'w') as f:                                     # run_nerf.py:1133-1137
        for arg in sorted(vars(args)):
            f.write('{} = {}\n'.format(arg, getattr(args, arg)))
    say(f"{len(images)} views ({len(i_train)} training views) {hwf[0]}x{hwf[1]}, near/far {near:.3f}/{far:.3f}, pool {len(pools.label)} rays (unmasked "
          f"{len(pools.idx_clf)}, masked {len(pools.idx_rgb)}, inpainted {len(pools.idx_inp)})"
          + (f", {depth_pool[1].shape[0]} sparse-depth rays" if depth_pool is not None else ""))
    if args.dry_run:
        print("dry run: stopping before the first GPU call")
        return 0

    # ---- model + trainer (create_nerf, run_nerf.py:380-496)
    dev = torch.device(args.device)
    nets = []
    for _ in range(2):
        net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).to(dev)
        net.precision = spn.PREC_BF16 if args.precision == "bf16" else spn.PREC_FP32
        nets.append(net)
    tr = trainer_mod.Trainer(nets[0], nets[1], lr=args.lrate, lrate_decay=args.lrate_decay, N_samples=args.N_samples,
                             N_importance=args.N_importance, lindisp=args.lindisp, white_bkgd=args.white_bkgd,
                             perturb=args.perturb, raw_noise_std=args.raw_noise_std, near=near, far=far, ndc=not args.no_ndc, hwf=hwf,
                             sharder=trainer_mod.RaySharder(rank, world))
    dist_mod.broadcast_parameters(nets)                      # torch's default init differs per process
    ckpts = [args.ft_path] if args.ft_path not in (None, 'None') else \
        [os.path.join(logdir, f) for f in sorted(os.listdir(logdir)) if 'tar' in f]
    if ckpts and not args.no_reload:
        say('Reloading from', ckpts[-1])
        tr.load_checkpoint(torch.load(ckpts[-1], map_location=dev, weights_only=False))
    start = tr.global_step
    test_kw = dict(network_query_fn=None, network_fn=nets[0], network_fine=nets[1], N_samples=args.N_samples,
                   N_importance=args.N_importance, lindisp=args.lindisp, white_bkgd=args.white_bkgd, perturb=0., raw_noise_std=0.,
                   use_viewdirs=True, ndc=not args.no_ndc, near=near, far=far)

    def video(tag, poses_, savedir=None, gt=None):
        if world > 1:        # frames dealt round-robin to the ranks, gathered over NVLink, handed to rank 0 (no per-frame dumps)
            rgbs, disps = spn.render_path_sharded(poses_, list(hwf), args.chunk, test_kw, render_factor=args.render_factor, dst=0)
            if rank != 0:
                return None
        else:
            rgbs, disps, _ = spn.render_path(poses_, list(hwf), args.chunk, test_kw, gt_imgs=gt, savedir=savedir,
                                             render_factor=args.render_factor, need_alpha=True)
        base = os.path.join(logdir, tag)
        spn.frame_io.write_video(base + 'rgb.mp4', rgbs)
        spn.frame_io.write_video(base + 'disp.mp4', disps / np.nanmax(disps))
        return rgbs

    if args.render_only:                                                                        # run_nerf.py:1168-1220
        out = os.path.join(logdir, 'renderonly_path_{:06d}'.format(start))
        os.makedirs(out, exist_ok=True)
        video(os.path.basename(out) + '_', render_poses, savedir=out)
        say('Done rendering', out)
        return 0

    # ---- optimisation loop (run_nerf.py:1360-1703)
    dev_pools = pools.to(dev)
    dpool = None if depth_pool is None else tuple(torch.from_numpy(a).to(dev) for a in depth_pool)
    lpips_fn = None
    if args.lpips:
        # LPIPS-VGG is a fixed input of this path (the pip package `lpips`, run_nerf.py:36,970-974).  Without it the run would
        # optimise a DIFFERENT perceptual loss, so the seeded stand-in network is only used when asked for explicitly.
        try:
            lpips_mod = importlib.import_module("lpips")
        except ImportError:
            if not args.lpips_standin:
                raise RuntimeError("--lpips needs the `lpips` package (LPIPS-VGG weights). It is not installed; pass --lpips_standin "
                                   "to run the branch with spin-nerf_b200/compat/lpips.py, a seeded random conv stack that exercises the "
                                   "same code path but is NOT the reference's perceptual loss.")
            lpips_mod = importlib.import_module("spin-nerf_b200.compat.lpips")
            if rank == 0:
                print("WARNING: --lpips_standin: the perceptual term is a seeded random conv stack, not LPIPS-VGG "
                      "(recorded in args.txt)", flush=True)
        lpips_fn = lpips_mod.LPIPS(net='vgg').to(dev)
    poses_t = torch.from_numpy(np.ascontiguousarray(poses[:, :3, :4])).float()
    t0, rays_done = time.perf_counter(), 0
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + start)    # same stream on every rank: the trainer shards the draw
    for i in range(start + 1, args.N_iters + 1):
        idx = rp.draw_step_indices(dev_pools, args.N_rand, generator=gen)
        kw = {}
        if dpool is not None:
            di = torch.randint(0, dpool[1].numel(), (args.N_rand,), device=dev, generator=gen)
            kw = dict(rays_depth=dpool[0][:, di], target_depth=dpool[1][di], depth_lambda=args.depth_lambda)
        if kw or args.prepare or (sampler is not None and i > args.lpips_from):
            if args.prepare:     # stage A renders no inpainted-disparity rays (run_nerf.py:1469-1473, 1515): an empty third group
                inp = (dev_pools["pool_od"][:, :0], dev_pools["disp"][:0])
            else:
                inp = (dev_pools["pool_od"][:, idx[2]], dev_pools["disp"][idx[2]])
            batch = (dev_pools["pool_od"][:, idx[0]], dev_pools["rgb"][idx[0]], dev_pools["pool_od"][:, idx[1]],
                     dev_pools["rgb"][idx[1]]) + inp
            loss, psnr = tr.step(*batch, _apply=False, **kw)
            if sampler is not None and i > args.lpips_from:                                     # run_nerf.py:1523
                views, Xs, Ys = sampler.sample()
                patches = [(x, y) + tuple(sampler.patch_len) for x, y in zip(Xs, Ys)]
                loss = loss + tr.lpips_patch_backward([poses_t[v] for v in views], patches, sampler.target_patches(views, Xs, Ys),
                                                      lpips_fn, (sampler.Hs, sampler.Ws, sampler.focal_s))
            tr.apply_gradients()
        else:
            loss, psnr = tr.step_from_pool(dev_pools["pool_od"], dev_pools["rgb"], dev_pools["disp"], idx)
        rays_done += (3 + (1 if kw else 0) - (1 if args.prepare else 0)) * args.N_rand
        if args.prepare and i % args.i_feat == 0:               # hand-over to the inpainter (run_nerf.py:1563-1609)
            rf = max(args.render_factor, 1)
            if world > 1:
                _, disps = spn.render_path_sharded(poses, list(hwf), args.chunk, test_kw, render_factor=args.render_factor, dst=0)
            else:
                _, disps, _ = spn.render_path(poses, list(hwf), args.chunk, test_kw, render_factor=args.render_factor)
            if rank == 0:
                import cv2
                os.makedirs(os.path.join(args.lama_dir, 'label'), exist_ok=True)
                for j in range(len(poses)):
                    cv2.imwrite(os.path.join(args.lama_dir, f'img{j:0>3}.png'), disps[j] * 255)
                    cv2.imwrite(os.path.join(args.lama_dir, 'label', f'img{j:0>3}.png'), masks[j][::rf, ::rf] * 255)
                print('Wrote', len(poses), 'disparity / mask pairs to', args.lama_dir)
        if i % args.i_weights == 0 and rank == 0:
            path = os.path.join(logdir, '{:06d}.tar'.format(i))
            torch.save(tr.checkpoint(), path)
            print('Saved checkpoints at', path)
        if args.i_video > 0 and i % args.i_video == 0:
            video('{}_{:06d}_'.format(args.expname, i), render_poses)
        if i % args.i_testset == 0 and len(i_test) > 0 and rank == 0:
            out = os.path.join(logdir, 'testset_{:06d}'.format(i))
            os.makedirs(out, exist_ok=True)
            spn.render_path(poses[i_test], list(hwf), args.chunk, test_kw, gt_imgs=images[i_test], savedir=out,
                            render_factor=args.render_factor)
        if tr.peer is not None and (i % args.i_weights == 0 or (i % args.i_testset == 0 and len(i_test) > 0)):
            torch.distributed.barrier()        # the peer-memory exchange polls device-side: nobody may enter the next step minutes early
        if i % args.i_print == 0 and rank == 0:
            print(f"[TRAIN] Iter: {i} Loss: {float(loss)}  PSNR: {float(psnr)}  "
                  f"({rays_done / (time.perf_counter() - t0) / 1e3:.0f} k rays/s wall clock)")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
