#!/usr/bin/env python
"""bench.py — rays/sec of the SPIn-NeRF train step (3 render calls -> losses -> backward -> Adam) on B200.

  python bench.py --gpus 1 --steps 20 --warmup 5          # our arm (one JSON line)
  python bench.py --impl reference --steps 3 --warmup 1   # the reference algorithm on the host CPUs (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...       # ray-sharded data parallel, NCCL grad all-reduce

Workload (BASELINE.json configs[1]): statue-shaped synthetic LLFF scene, 1008x756 (factor 2), 30 views,
coarse+fine N_samples=64 N_importance=64, N_rand=1024 rays per render call per GPU, no_ndc, lindisp,
white_bkgd, use_viewdirs, perturb=1, raw_noise_std=1, random-init weights (seed 0).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 756, 1008
FOCAL = 0.9 * W
N_VIEWS = 30
NEAR, FAR = 1.2, 8.0
FLOP_FWD, FLOP_BWD = 1186816, 2302208            # per MLP evaluation
EVALS_PER_RAY = 192                              # 64 coarse + 128 fine
RENDERS_PER_STEP = 3


def poses(n, seed=0):
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the train step on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_threads_best(n_rays=64):
    """Thread count for the CPU arm: all host threads unless fewer are measurably faster (on a 128-core box the 256-wide GEMMs
    of this workload ran 5x slower on 128 threads than on 16).  One probe step per candidate."""
    import torch
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    for c in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(c)
        _, t = cpu_train_steps(n_rays, 1, 1)
        if best_t is None or t < 0.9 * best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def cpu_train_steps(n_rays, steps, warmup, seed=0, anomaly=False):
    """The reference's formulation of the train step on the host cores: eager PyTorch CPU ops, autograd, torch.optim.Adam
    (oracle/torch_port.py — the PyTorch restatement pinned against the reference's goldens; the unmodified reference is
    Python + PyTorch too but cannot travel to the GPU box).  Returns (rays/s, MEDIAN seconds per step).  anomaly=True runs
    with autograd anomaly detection on, which the reference switches on globally at import (run_nerf_helpers.py:5)."""
    import torch
    from oracle import nerf_oracle as O
    from oracle import torch_port as TP
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    pc, pf = TP.make_params(O.init_params(1), "cpu"), TP.make_params(O.init_params(2), "cpu")
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=5e-4, betas=(0.9, 0.999))
    all_rays = [O.get_rays(H, W, FOCAL, p) for p in poses(4)]   # the scene's rays are resident before the timed steps, as on the GPU
    times = []
    torch.autograd.set_detect_anomaly(bool(anomaly))
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        batches = []
        for call in range(RENDERS_PER_STEP):
            ro, rd = all_rays[(it + call) % len(all_rays)]
            sel = rng.choice(H * W, n_rays, replace=False)
            rays = torch.from_numpy(np.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0))
            batches.append((rays, torch.rand((n_rays, 3) if call < 2 else (n_rays,), generator=g)))
        opt.zero_grad()
        loss, _ = TP.spin_step_loss(batches, pc, pf, NEAR, FAR, perturb=True, raw_noise_std=1.0)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    torch.autograd.set_detect_anomaly(False)
    med = float(np.median(times))
    return RENDERS_PER_STEP * n_rays / med, med


def gpu_reference_port(spn, dev, pool, rgb_pool, disp_pool, n_rand, steps=6, warmup=2):
    """The like-for-like baseline of SURVEY.md section 8d: the reference's own formulation of the step — eager PyTorch ops,
    fp32 GEMMs with TF32 off (torch's default, which the reference never changes), autograd, torch.optim.Adam — on the SAME
    GPU and workload.  The unmodified reference cannot travel to the GPU box, so this times its PyTorch restatement
    (oracle/torch_port.py, pinned on CPU against the reference's goldens); baseline only, never the product path."""
    import torch
    from oracle import torch_port as TP
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    nets = [spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).seeded_init_(s)
            for s in (1, 2)]
    pc, pf = ({k: v.detach().clone().to(dev).requires_grad_(True) for k, v in n.state_dict().items()} for n in nets)
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=5e-4, betas=(0.9, 0.999))
    g = torch.Generator(device=dev); g.manual_seed(1)
    M = pool.shape[1]
    evs = []
    for it in range(warmup + steps):
        idx = torch.randint(0, M, (3, n_rand), device=dev, generator=g)
        b = [(pool[:, idx[0]], rgb_pool[idx[0]]), (pool[:, idx[1]], rgb_pool[idx[1]]), (pool[:, idx[2]], disp_pool[idx[2]])]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad()
        loss, _ = TP.spin_step_loss(b, pc, pf, NEAR, FAR, perturb=True, raw_noise_std=1.0)
        loss.backward()
        opt.step()
        e1.record()
        if it >= warmup:
            evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    return {"value": RENDERS_PER_STEP * n_rand / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "kind": "port: PyTorch restatement of the reference step (eager ops, fp32 GEMMs with TF32 off, autograd, "
                    "torch.optim.Adam) on the same GPU and workload; device-timed", "final_loss": float(loss.detach())}


def psnr_vs_reference_port(spn, dev, prec, n_rays=1024):
    """The second half of BASELINE.json's metric ("PSNR vs ref") on a TRAINED scene: the analytic scene the unmodified reference
    was trained on for tests/golden/convergence.npz (tests/golden/make_convergence_golden.py: 4 views of a textured plane, 300
    steps, 8.9 -> 30.1 dB) is trained here with the benchmark's arithmetic (same initial weights, same ray batches), then
    (a) the training PSNR reached is put beside the reference's, and (b) this library's render of held-out rays with the
    trained weights is compared with the reference's fp32 formulation (oracle/torch_port.py, TF32 off) on the same weights.
    A checker leg: a failure costs only this key."""
    import importlib.util
    import torch
    from oracle import torch_port as TP
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
    gpath = os.path.join(ROOT, "tests", "golden", "make_convergence_golden.py")
    spec = importlib.util.spec_from_file_location("make_convergence_golden", gpath)
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "convergence.npz"))
    torch.backends.cuda.matmul.allow_tf32 = False
    ro, rd, rgb_t, disp_t, idx = gen.problem()
    nets = []
    for p in gen.params():
        net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        net = net.to(dev); net.precision = prec
        nets.append(net)
    tr = trainer_mod.Trainer(nets[0], nets[1], lr=gen.LR, lrate_decay=gen.DECAY, N_samples=64, N_importance=64, lindisp=True,
                             white_bkgd=True, perturb=0.0, raw_noise_std=0.0, near=gen.NEAR, far=gen.FAR)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pool = T(np.stack([ro, rd], 0)); rgb_pool, disp_pool = T(rgb_t), T(disp_t)
    psnrs = []
    for it in range(gen.K):
        _, ps = tr.step_from_pool(pool, rgb_pool, disp_pool, torch.from_numpy(idx[it]).to(dev))
        psnrs.append(ps)
    trained = float(torch.stack(psnrs)[-20:].mean())
    g = torch.Generator(device=dev); g.manual_seed(7)
    ix = torch.randint(0, pool.shape[1], (n_rays,), device=dev, generator=g)
    rays = pool[:, ix].contiguous()
    with torch.no_grad():
        rgb, disp, acc, depth, ex = spn.render(gen.H, gen.W, gen.FOCAL, chunk=32768, rays=rays, use_viewdirs=True, ndc=False,
                                               near=gen.NEAR, far=gen.FAR, network_query_fn=None, network_fn=nets[0],
                                               network_fine=nets[1], N_samples=64, N_importance=64, lindisp=True, white_bkgd=True,
                                               perturb=0., raw_noise_std=0., need_alpha=True)
        # what early ray termination could skip in the fine pass of a no-grad render (north_star): samples whose transmittance
        # has already dropped below a threshold (run_nerf_helpers.py:384: T = exclusive cumprod of 1 - alpha + 1e-10)
        Tr = torch.cumprod(torch.cat([torch.ones_like(ex["alpha"][:, :1]), 1.0 - ex["alpha"] + 1e-10], -1), -1)[:, :-1]
        ert = {f"T<{t:g}": float((Tr < t).float().mean()) for t in (1e-2, 1e-3, 1e-4)}
        pc, pf = ({k: v.detach().clone().float() for k, v in n.state_dict().items()} for n in nets)
        ref = TP.render_rays(rays[0], rays[1], gen.NEAR, gen.FAR, pc, pf, lindisp=True, white_bkgd=True)
        mse = float(torch.mean((rgb - ref["rgb_map"]) ** 2))
        mse0 = float(torch.mean((ex["rgb0"] - ref["rgb0"]) ** 2))
        mse_t = float(torch.mean((rgb - rgb_pool[ix]) ** 2)); mse_rt = float(torch.mean((ref["rgb_map"] - rgb_pool[ix]) ** 2))
    db = lambda m: float(-10.0 * np.log10(max(m, 1e-20)))
    return {"rgb_db": db(mse), "rgb0_db": db(mse0), "rays": n_rays,
            "trained_psnr_db": trained, "reference_trained_psnr_db": float(gold["psnr"][-20:].mean()), "train_steps": int(gen.K),
            "render_vs_target_db": db(mse_t), "reference_render_vs_target_db": db(mse_rt),
            "ert_skippable_fine_sample_frac": ert,
            "scene": "analytic 4-view scene of tests/golden/make_convergence_golden.py, trained 300 steps with the benchmark's arithmetic",
            "checker": "oracle/torch_port.py: fp32 PyTorch render of the same rays with the same trained weights, TF32 off; "
                       "reference_trained_psnr_db from the unmodified reference's own 300-step run (tests/golden/convergence.npz)"}


def hbm_write_gbs(dev):
    """pure-write HBM bandwidth measured live (torch memset of 2 GiB, best of 5): what bounds the stash writes of the training
    forward / dgrad kernels; MEASURED_PEAKS.json's hbm_gbs is a COPY (half reads, half writes) and overstates it"""
    import torch
    x = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x.zero_(); e1.record(); torch.cuda.synchronize()
        best = max(best, x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del x
    return best


def other_workloads(timeout=150):
    """BASELINE configs[2] and configs[4] at N = 1, each as its own short `bench.py --workload ...` run (fresh process, same JSON
    contract), so that the default line carries every single-GPU configuration; configs[3] (strong scaling) needs N > 1."""
    out = {}
    for name, extra in (("train_lpips", ["--steps", "6", "--warmup", "3"]), ("render", ["--steps", "3", "--warmup", "3"])):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", name, "--deadline", str(timeout)] + extra,
                               capture_output=True, text=True, timeout=timeout + 30, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK=os.environ.get("LOCAL_RANK", "0")))
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if not line:
                raise RuntimeError(f"no result line (rc {r.returncode}): {r.stderr.strip()[-240:]}")
            d = json.loads(line[-1])
            out[name] = {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches", "frames_per_sec", "e2e")}
            out[name]["workload"] = d["config"]["workload"]
            if d.get("roofline"):
                out[name]["mlp_kernels"] = d["roofline"].get("kernels")
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_rays
    cores = cpu_threads_best()
    # bounded sample: one probe step sizes the per-step ray count so that warmup + steps end within ~4 minutes
    _, probe = cpu_train_steps(n, 1, 0)
    budget = 240.0 / max(1, args.steps + args.warmup)
    while n > 8 and probe * (n / args.cpu_rays) > budget:
        n //= 2
    rps, sec = cpu_train_steps(n, args.steps, args.warmup)
    rps_anom, _ = cpu_train_steps(n, min(args.steps, 3), 1, anomaly=True)
    line = {"impl": "reference", "metric": "rays/sec (train-step)", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, n),
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "anomaly_mode_on_value": rps_anom,
                             "sample": f"{RENDERS_PER_STEP}x{n} rays per step (same scene/config), PyTorch CPU restatement of the reference train "
                                       "step (oracle/torch_port.py: eager ops, autograd, torch.optim.Adam), median step time; `value` with autograd anomaly "
                                       "detection off (the faster way), `anomaly_mode_on_value` with it on as the reference runs"},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus, n_rand):
    return {"workload": "statue-shaped synthetic LLFF 1008x756 (factor 2), 30 views, coarse+fine N_samples=64 N_importance=64, "
                        f"N_rand={n_rand} rays per render call per GPU, {RENDERS_PER_STEP} render calls per step, "
                        "no_ndc lindisp white_bkgd use_viewdirs perturb=1 raw_noise_std=1 (BASELINE configs[1])",
            "n_rand_per_gpu": n_rand, "renders_per_step": RENDERS_PER_STEP, "parallelism": f"ray-dp{n_gpus}",
            "l2": "ray pool 548 MB + per-step activation stash > 126 MB L2; L2 flushed (256 MB write) between timed steps"}



# ------------------------------------------------------------------------------------------------
# the other two GPU configurations of BASELINE.json (not the headline line; same JSON contract)
# ------------------------------------------------------------------------------------------------
def _device(local):
    import torch
    return torch.device("cuda", local)


def _init_dist():
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = _device(local)
    if world > 1:
        import datetime
        import torch.distributed as dist
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    return rank, world, local, dev


def _max_over_ranks(x, dev, world):
    import torch
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def _kernel_profile(L, steps):
    import ctypes
    prof = {}
    for kind, name in ((0, "mlp_fwd"), (1, "mlp_dgrad"), (2, "mlp_wgrad")):
        n_l, ms_l = ctypes.c_int(), ctypes.c_float()
        L.spn_profile_read(kind, ctypes.byref(n_l), ctypes.byref(ms_l))
        if n_l.value and ms_l.value > 0:
            prof[name] = {"launches_per_step": n_l.value / steps, "ms_per_step": ms_l.value / steps}
    return prof


def run_other_workload(args):
    import torch
    spn = importlib.import_module("spin-nerf_b200")
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
    rank, world, local, dev = _init_dist()
    prec = spn.PREC_BF16 if args.precision == "bf16" else spn.PREC_FP32
    nets = []
    for seed in (1, 2):
        net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net = net.seeded_init_(seed).to(dev); net.precision = prec
        nets.append(net)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    barrier = (lambda: torch.distributed.barrier()) if world > 1 else (lambda: None)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    L = spn._lib.lib()
    clocks = ClockSampler(local); clocks.start()
    all_poses = poses(N_VIEWS)

    if args.workload == "render":
        # ---- configs[4]: one full-resolution frame per step and GPU through render() in chunks of 32768 rays (run_nerf.py:767)
        kw = dict(network_query_fn=None, network_fn=nets[0], network_fine=nets[1], N_samples=64, N_importance=64, lindisp=True,
                  white_bkgd=True, perturb=0., raw_noise_std=0., use_viewdirs=True, ndc=False, near=NEAR, far=FAR)
        rays_per_step = H * W * world

        def frames(k, first):
            evs = []
            for i in range(k):
                flush.fill_(1.0)
                c2w = torch.from_numpy(all_poses[(first + i * world + rank) % N_VIEWS]).to(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                with torch.no_grad():
                    rgb, disp, acc, depth, _ = spn.render(H, W, FOCAL, chunk=32768, c2w=c2w, **kw)
                e1.record(); evs.append((e0, e1))
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs), rgb
        phase(f"render workload: {args.warmup} warm-up frames"); frames(args.warmup, 0)
        barrier(); torch.cuda.synchronize()
        L.spn_profile_enable(1); L.spn_launch_count(1); clocks.rows.clear()
        ms, rgb = frames(args.steps, args.warmup)
        torch.cuda.synchronize(); barrier()
        clk = clocks.stop(); launches = int(L.spn_launch_count(0)); prof = _kernel_profile(L, args.steps); L.spn_profile_enable(0)
        ms = _max_over_ranks(ms, dev, world)
        value = rays_per_step * args.steps / (ms * 1e-3)
        # e2e: the public video API with host poses in and host frames out (frames sharded over the ranks, gathered over NVLink)
        vid_poses = [all_poses[i % N_VIEWS] for i in range(args.steps * world)]
        spn.render_path_sharded(vid_poses[:world], [H, W, FOCAL], 32768, kw)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        rgbs, disps = spn.render_path_sharded(vid_poses, [H, W, FOCAL], 32768, kw)
        torch.cuda.synchronize(); barrier()
        e2e_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, dev, world)
        flop = FLOP_FWD * EVALS_PER_RAY * H * W
        dom = prof.get("mlp_fwd")
        tfl = flop / (dom["ms_per_step"] * 1e-3) / 1e12 if dom else None
        line = {"metric": "rays/sec (render_path)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if prec == spn.PREC_BF16 else "f32", "data": "synthetic",
                "config": {"workload": "full-resolution 1008x756 novel-view frames (render_path), coarse+fine 64+64 samples, chunks of "
                                       "32768 rays, one frame per step and GPU, frames sharded round-robin (BASELINE configs[4])",
                           "frames_per_step": world, "parallelism": f"frame-dp{world}",
                           "l2": "per-frame outputs 0.8 GB > 126 MB L2; L2 flushed (256 MB write) between timed frames"},
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": H * W * len(vid_poses) / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 48 * world,
                        "d2h_bytes_per_step": H * W * 16 * world, "ms_per_step": e2e_ms / args.steps,
                        "note": "render_path_sharded: host poses in, gathered host RGB+disparity video out; timed with the host clock "
                                "around the call (it ends in a device->host copy), max over ranks"},
                "roofline": {"kernel": "mlp_fwd", "bound": "tensor", "achieved": tfl, "peak": peak, "unit": "TFLOP/s",
                             "frac": tfl / peak if tfl else None, "traffic": None, "kernels": prof},
                "cpu_baseline": None, "frames_per_sec": world * args.steps / (ms * 1e-3),
                "finite": bool(np.isfinite(rgbs).all()) if rgbs is not None else None}
    else:
        # ---- configs[2]: --lpips step: 3 x N_rand=4096 rays + 4 patches of 47x63 at render factor 2 per step and GPU
        lp_mod = importlib.import_module("spin-nerf_b200.lpips_patch")
        lpips = importlib.import_module("spin-nerf_b200.compat.lpips").LPIPS().to(dev)     # fixed-input stand-in (no lpips package)
        n_rand = 4096 if args.n_rand == 1024 else args.n_rand
        tr = trainer_mod.Trainer(nets[0], nets[1], lr=5e-4, N_samples=64, N_importance=64, lindisp=True, white_bkgd=True,
                                 perturb=1.0, raw_noise_std=1.0, near=NEAR, far=FAR, ndc=False, hwf=(H, W, FOCAL),
                                 sharder=trainer_mod.RaySharder(rank, world))
        g = torch.Generator(device=dev); g.manual_seed(0)
        ro_all, rd_all = [], []
        for c2w in all_poses:
            ro, rd = spn.ops.get_rays(H, W, FOCAL, torch.from_numpy(c2w).to(dev))
            ro_all.append(ro.reshape(-1, 3)); rd_all.append(rd.reshape(-1, 3))
        pool = torch.stack([torch.cat(ro_all), torch.cat(rd_all)], 0)
        M = pool.shape[1]
        rgb_pool = torch.rand(M, 3, device=dev, generator=g); disp_pool = torch.rand(M, device=dev, generator=g)
        Hs, Ws, fs, plen = lp_mod.patch_geometry([H, W, FOCAL], 2, 8)
        B = 4 * world
        tgt_frames = torch.rand(N_VIEWS, 3, Hs, Ws, device=dev, generator=g) * 2 - 1
        rnd = __import__("random").Random(0)
        n_patch_rays = 4 * plen[0] * plen[1]
        rays_per_step = (RENDERS_PER_STEP * n_rand + n_patch_rays) * world

        def one_step():
            idx = torch.randint(0, M, (3, n_rand * world), device=dev, generator=g)
            views = [rnd.randrange(N_VIEWS) for _ in range(B)]
            patches = [(rnd.randint(0, Hs - plen[0]), rnd.randint(0, Ws - plen[1]), plen[0], plen[1]) for _ in range(B)]
            tg = [lp_mod.crop(tgt_frames, v, p[0], p[1], plen) for v, p in zip(views, patches)]
            return tr.step_with_lpips((pool, rgb_pool, disp_pool, idx), [all_poses[v] for v in views], patches, tg, lpips,
                                      (Hs, Ws, fs), from_pool=True)

        def steps(k):
            evs = []
            for _ in range(k):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); loss, _ = one_step(); e1.record(); evs.append((e0, e1))
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs), loss
        phase(f"train_lpips workload: {args.warmup} warm-up steps"); steps(args.warmup)
        barrier(); torch.cuda.synchronize()
        L.spn_profile_enable(1); L.spn_launch_count(1); clocks.rows.clear()
        ms, loss = steps(args.steps)
        torch.cuda.synchronize(); barrier()
        clk = clocks.stop(); launches = int(L.spn_launch_count(0)); prof = _kernel_profile(L, args.steps); L.spn_profile_enable(0)
        ms = _max_over_ranks(ms, dev, world)
        value = rays_per_step * args.steps / (ms * 1e-3)
        evals = (RENDERS_PER_STEP * n_rand + n_patch_rays) * EVALS_PER_RAY
        dom = max(prof, key=lambda k: prof[k]["ms_per_step"]) if prof else None
        line = {"metric": "rays/sec (train-step)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if prec == spn.PREC_BF16 else "f32", "data": "synthetic",
                "config": {"workload": f"--lpips train step: 3 x N_rand={n_rand} rays + 4 LPIPS patches of {plen[0]}x{plen[1]} rays at render "
                                       "factor 2 per step and GPU, LPIPS = frozen stand-in conv stack (fixed input), statue-shaped synthetic "
                                       "scene (BASELINE configs[2])",
                           "n_rand_per_gpu": n_rand, "patch_rays_per_gpu": n_patch_rays, "parallelism": f"ray-dp{world}",
                           "l2": "activation stash > 126 MB L2; L2 flushed (256 MB write) between timed steps"},
                "clocks": clk, "gpu_launches": launches, "e2e": None,
                "roofline": {"kernel": dom, "kernels": prof, "bound": "tensor", "peak": peak, "unit": "TFLOP/s", "traffic": None,
                             "achieved": None, "frac": None,
                             "step_mlp_flop_frac_of_peak": evals * (FLOP_FWD + FLOP_BWD) * args.steps / (ms * 1e-3) / 1e12 / peak},
                "cpu_baseline": None, "final_loss": float(loss)}
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
PARTIAL = {}     # rank 0's result line as far as it is known; the watchdog prints it if a later phase hangs
T_START = time.perf_counter()


def phase(msg):
    """stderr breadcrumb (rank 0 only): if a run ever stalls, the log says in which phase"""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write(f"[bench +{time.perf_counter() - T_START:6.1f}s] {msg}\n"); sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n_rand", type=int, default=1024)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, what the driver's 1/2/4/8 series measures): --n_rand rays per render call PER GPU; strong: "
                         "--n_rand_global rays per render call in total, split over the ranks (BASELINE configs[3]: 8192 over 8 GPUs)")
    ap.add_argument("--n_rand_global", type=int, default=8192)
    ap.add_argument("--cpu_rays", type=int, default=1024, help="rays per render call in the CPU sample (default: the GPU arm's N_rand, same config)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_other_workloads", action="store_true", help="skip the short configs[2] / configs[4] runs appended to the N = 1 line")
    ap.add_argument("--workload", default="train", choices=["train", "train_lpips", "render"],
                    help="train = BASELINE configs[1] (the headline line, default); train_lpips = configs[2] (N_rand=4096 + 4 LPIPS "
                         "patches of 47x63 per step and GPU); render = configs[4] (full 1008x756 frames through render_path, one "
                         "frame per step and GPU)")
    ap.add_argument("--deadline", type=float, default=420.0,
                    help="hard wall-clock limit in seconds: a watchdog thread ends the process (exit code 3) if the run has not "
                         "finished by then, so a hung collective or kernel can never hold the GPU box")
    args = ap.parse_args()
    if args.impl == "reference":
        args.deadline = max(args.deadline, 1200.0)       # host-only run: nothing to protect but the caller's patience
    if args.deadline > 0:
        def _watchdog():
            time.sleep(args.deadline)
            sys.stderr.write(f"bench.py: deadline of {args.deadline:.0f} s exceeded, aborting\n"); sys.stderr.flush()
            if PARTIAL and int(os.environ.get("RANK", "0")) == 0:       # report what was measured before the hang
                print(json.dumps(dict(PARTIAL, aborted=f"deadline {args.deadline:.0f} s exceeded after the device-timed region")),
                      flush=True)
            os._exit(3)
        threading.Thread(target=_watchdog, daemon=True).start()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    if args.workload != "train":
        return run_other_workload(args)

    import torch
    spn = importlib.import_module("spin-nerf_b200")
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = _device(local)
    if world > 1:
        import datetime
        import torch.distributed as dist
        # One 4.8 MB all-reduce per step is latency-bound either way, so the in-switch (NVLS) algorithm buys nothing here;
        # its multicast set-up is the most fragile part of NCCL initialisation inside containers, and an unexplained
        # 8-GPU hang cost this project a round's GPU budget.  Off unless the caller asks for it; collectives that stall
        # abort the job after 3 minutes instead of torch's default 10.
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        phase(f"init_process_group nccl world={world}")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    prec = spn.PREC_BF16 if args.precision == "bf16" else spn.PREC_FP32

    # ---- model: reference-shaped coarse + fine networks, seeded init (identical on every rank)
    nets = []
    for seed in (1, 2):
        net = spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        net = net.seeded_init_(seed).to(dev); net.precision = prec
        nets.append(net)
    n_rand = args.n_rand if args.scaling == "weak" else max(1, args.n_rand_global // world)
    sharder = trainer_mod.RaySharder(rank, world)
    tr = trainer_mod.Trainer(nets[0], nets[1], lr=5e-4, N_samples=64, N_importance=64, lindisp=True, white_bkgd=True,
                             perturb=1.0, raw_noise_std=1.0, near=NEAR, far=FAR, ndc=False, hwf=(H, W, FOCAL),
                             sharder=sharder)

    # ---- synthetic scene resident in HBM: all rays of all views (548 MB), targets, inpainted disparities
    g = torch.Generator(device=dev); g.manual_seed(0)
    ro_all, rd_all = [], []
    for c2w in poses(N_VIEWS):
        ro, rd = spn.ops.get_rays(H, W, FOCAL, torch.from_numpy(c2w).to(dev))
        ro_all.append(ro.reshape(-1, 3)); rd_all.append(rd.reshape(-1, 3))
    pool = torch.stack([torch.cat(ro_all), torch.cat(rd_all)], 0)            # [2, M, 3]
    M = pool.shape[1]
    rgb_pool = torch.rand(M, 3, device=dev, generator=g); disp_pool = torch.rand(M, device=dev, generator=g)
    global_n = n_rand * world

    def device_batches():
        idx = torch.randint(0, M, (3, global_n), device=dev, generator=g)
        return (pool[:, idx[0]], rgb_pool[idx[0]], pool[:, idx[1]], rgb_pool[idx[1]], pool[:, idx[2]], disp_pool[idx[2]])

    flush = torch.empty(64 * 1024 * 1024, device=dev)                         # 256 MB > 126 MB L2

    def pool_step():
        """device-resident sampler: draw the step's ray indices, gather + assemble the batch in one kernel"""
        return tr.step_from_pool(pool, rgb_pool, disp_pool, torch.randint(0, M, (3, global_n), device=dev, generator=g))

    def run_steps(k, batch_fn, timed):
        ms, evs = 0.0, []
        for _ in range(k):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss, psnr = tr.step(*batch_fn()) if batch_fn is not None else pool_step()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], loss

    barrier = (lambda: torch.distributed.barrier()) if world > 1 else (lambda: None)
    clocks = ClockSampler(local); clocks.start()         # nvidia-smi needs ~1 s to start streaming samples
    phase(f"scene resident ({M} rays); {args.warmup} warm-up steps")
    run_steps(args.warmup, None, False)
    barrier(); torch.cuda.synchronize()
    L = spn._lib.lib()
    L.spn_profile_enable(1); L.spn_launch_count(1)
    clocks.rows.clear()                                   # keep only samples taken during the timed region
    t_wall = time.perf_counter()
    phase(f"{args.steps} timed steps")
    times, loss = run_steps(args.steps, None, True)
    torch.cuda.synchronize(); barrier()
    wall = time.perf_counter() - t_wall
    clk = clocks.stop()
    launches = int(L.spn_launch_count(0))
    import ctypes
    prof = {}
    for kind, name in ((0, "mlp_fwd"), (1, "mlp_dgrad"), (2, "mlp_wgrad")):
        n_l, ms_l = ctypes.c_int(), ctypes.c_float()
        L.spn_profile_read(kind, ctypes.byref(n_l), ctypes.byref(ms_l))
        prof[name] = (n_l.value, ms_l.value)
    L.spn_profile_enable(0)
    step_ms = float(np.sum(times))
    if world > 1:
        t = torch.tensor([step_ms], device=dev); torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        step_ms = float(t.item())
    rays_per_step = RENDERS_PER_STEP * global_n
    value = rays_per_step * args.steps / (step_ms * 1e-3)
    PARTIAL.update({"metric": "rays/sec (train-step)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
                    "vs_baseline": None, "dtype": "bf16" if prec == spn.PREC_BF16 else "f32", "data": "synthetic",
                    "config": workload_config(world, n_rand), "gpu_launches": launches, "e2e": None})

    phase(f"device-timed value = {value:.0f} rays/s; e2e pass from pinned host batches")
    # ---- e2e: the public train-step API fed from pinned HOST memory, loss read back each step
    host = [tuple(t.cpu().pin_memory() for t in device_batches()) for _ in range(max(args.warmup, 3) + args.steps)]
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    it = iter(host)

    def host_batches():
        return next(it)        # pinned host tensors: step_graphed copies them H2D into its static inputs

    # N > 1: three graphs with the two eager all-reduces between them; the opt-in peer-memory exchange is not graph-captured
    graph_e2e = os.environ.get("SPN_GRAPH_E2E", "1") != "0" and tr.peer is None

    def e2e_steps(k):
        evs = []
        for _ in range(k):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if graph_e2e:
                loss, psnr = tr.step_graphed(*host_batches())     # H2D copies + CUDA-graph replay of the whole step
            else:   # SPN_GRAPH_E2E=0: eager launches
                loss, psnr = tr.step(*[t.to(dev, non_blocking=True) for t in host_batches()])
            _ = float(loss)                      # D2H of the step's result (synchronises)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]
    e2e_steps(max(args.warmup, 3))               # eager pass, capture, first replays
    barrier()
    e2e_ms = float(np.sum(e2e_steps(args.steps)))
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev); torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = rays_per_step * args.steps / (e2e_ms * 1e-3)

    phase("e2e done")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel (fused MLP forward, tcgen05): MLP FLOPs / CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # a timed region of seconds runs under the power cap (cuBLAS settles at ~1350 MHz there): sustained peak; a short region
    # (the default 100 x 3.3 ms) runs at boost clocks: the burst peak is the honest denominator
    long_run = step_ms >= 4000.0
    pk_key = "bf16_tflops_sustained" if long_run else "bf16_tflops"
    peak = float(peaks.get(pk_key, 1400.0 if long_run else 1590.0))
    peak_src = (f"MEASURED_PEAKS.json {pk_key} ({'timed region >= 4 s: power-capped steady state' if long_run else 'timed region < 4 s: boost clocks'})"
                if peaks else f"fallback {'1.4' if long_run else '1.59'} PFLOP/s (B200_PROFILING.md)")
    evals_per_rank_step = RENDERS_PER_STEP * n_rand * EVALS_PER_RAY
    fwd_n, fwd_ms = prof["mlp_fwd"]
    kern = {}
    for name, flop in (("mlp_fwd", FLOP_FWD), ("mlp_dgrad", 2 * 557696), ("mlp_wgrad", 2 * 593408)):
        n_l, ms_l = prof[name]
        if n_l and ms_l > 0:
            kern[name] = {"launches_per_step": n_l / args.steps, "ms_per_step": ms_l / args.steps,
                          "tflops": evals_per_rank_step * flop * args.steps / (ms_l * 1e-3) / 1e12}
    # per-kernel roofline: algorithmic FLOPs and algorithmic HBM bytes per MLP evaluation, each against its
    # measured peak; the kernel's bound is the larger of the two floors.  Bytes: forward writes the activation stash once
    # (397 312 B per 128-sample tile: 18 E4M3 atoms + 4 bf16 atoms + masks), dgrad reads the masks and writes the dstash
    # (38 bf16 atoms), wgrad reads every stash / dstash atom it needs once (20 + 38 atoms of 16 KB).
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_eval = {"mlp_fwd": 397312 / 128.0, "mlp_dgrad": (38 * 16384 + 9 * 128 * 32) / 128.0, "mlp_wgrad": 58 * 16384 / 128.0}
    try:
        wr_bw = hbm_write_gbs(dev)
    except Exception:
        wr_bw = None
    for name, k in kern.items():
        gbs = evals_per_rank_step * bytes_per_eval[name] / (k["ms_per_step"] * 1e-3) / 1e9
        t_frac, h_frac = k["tflops"] / peak, gbs / hbm_peak
        k.update(tensor_frac=t_frac, hbm_gbs=gbs, hbm_frac=h_frac)
        if h_frac >= t_frac:
            k.update(bound="hbm", achieved=gbs, peak=hbm_peak, unit="GB/s", frac=h_frac)
        else:
            k.update(bound="tensor", achieved=k["tflops"], peak=peak, unit="TFLOP/s", frac=t_frac)
        if name != "mlp_wgrad" and wr_bw:      # these two only WRITE: a pure-write stream reaches ~60 % of the copy figure on this part
            k.update(frac_of_measured_write_bw=gbs / wr_bw)
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(dom)
        traffic_src = f"static: profiles/traffic.json from ncu capture {tj.get('_source')} (dram bytes per launch, mean of the coarse and fine launch; not measured in this run)"
    except Exception:
        pass
    roofline = None
    if dom:
        d = kern[dom]
        roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                    "frac": d["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": ("MEASURED_PEAKS.json hbm_gbs (copy: read + write)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)") if d["bound"] == "hbm" else peak_src,
                    "hbm_write_gbs_measured_live": wr_bw,
                    "algorithmic_units": "per MLP evaluation: fwd 1 186 816 FLOP / 3 104 B written, dgrad 1 115 392 FLOP / 5 152 B, "
                                         "wgrad 1 186 816 FLOP / 7 424 B read (58 atoms of 16 KB per 128-sample tile)",
                    "kernels": kern,
                    "step_mlp_flop_frac_of_peak": evals_per_rank_step * (FLOP_FWD + FLOP_BWD) * args.steps / (step_ms * 1e-3) / 1e12 / peak}
    psnr = None
    if world == 1:
        phase("PSNR of the render against the fp32 reference formulation")
        try:
            psnr = psnr_vs_reference_port(spn, dev, prec)
        except Exception as e:
            psnr = {"error": f"{type(e).__name__}: {e}"[:300]}
    gpu_port = None
    if not args.no_cpu_baseline and world == 1:
        phase("gpu_reference_port (PyTorch restatement of the reference step on this GPU)")
        try:
            gpu_port = gpu_reference_port(spn, dev, pool, rgb_pool, disp_pool, n_rand)
        except Exception as e:                           # a baseline must never cost the run its result line
            gpu_port = {"error": f"{type(e).__name__}: {e}"[:300]}
    cpu = None
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N = 1 only
        phase("cpu_baseline sample (PyTorch restatement of the reference step on the host cores)")
        t0 = time.perf_counter()
        try:
            cores = cpu_threads_best()
            rps, sec = cpu_train_steps(args.cpu_rays, 5, 2)
            rps_anom, _ = cpu_train_steps(args.cpu_rays, 3, 1, anomaly=True)
            cpu = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "anomaly_mode_on_value": rps_anom,
                   "sample": f"median of 5 steps after 2 warm-ups of {RENDERS_PER_STEP}x{args.cpu_rays} rays, same scene/config, "
                             "PyTorch CPU restatement of the reference train step (oracle/torch_port.py); `value` with autograd "
                             "anomaly detection off, `anomaly_mode_on_value` with it on as the reference runs (run_nerf_helpers.py:5) "
                             f"({time.perf_counter() - t0:.1f} s)"}
        except Exception as e:                           # a baseline must never cost the run its result line
            cpu = {"error": f"{type(e).__name__}: {e}"[:300]}
    line = {"metric": "rays/sec (train-step)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16" if prec == spn.PREC_BF16 else "f32", "data": "synthetic",
            "config": workload_config(world, n_rand), "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "roofline": roofline, "cpu_baseline": cpu, "gpu_reference_port": gpu_port, "psnr_vs_ref": psnr, "n_rand_per_sec": value / RENDERS_PER_STEP,
            "wall_s_timed_region": wall, "final_loss": float(loss),
            "stash": "activations E4M3, gradients bf16 (DESIGN.md section 3)",
            "grad_exchange": ("peer-memory reduce-scatter / all-gather fused with Adam (SPN_P2P_ALLREDUCE=1)" if tr.peer is not None
                              else "NCCL all-reduce per network, the fine one overlapped with the coarse backward") if world > 1 else None}
    if world == 1 and not args.no_other_workloads:
        phase("other workloads (BASELINE configs[2], configs[4]) as short separate runs")
        del tr, pool, rgb_pool, disp_pool, flush
        torch.cuda.empty_cache()
        line["other_workloads"] = other_workloads()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
