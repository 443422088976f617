"""How much could early ray termination skip in the fine pass of a no-grad render (BASELINE north_star)?
Build-container study on the CPU, oracle side only (nothing here is on the product path):

    python tests/golden/ert_study.py [--steps 4000]

Trains the analytic 4-view scene of make_convergence_golden.py with the PyTorch restatement of the reference step
(oracle/torch_port.py, the same step formulation: three render calls, six MSE terms, Adam, lr decay) well past the 300 steps of
the matched-PSNR golden, and at a few checkpoints renders 512 held-out-order rays and counts the fine-pass samples whose
transmittance T_i = prod_{j<i} (1 - alpha_j + 1e-10) (run_nerf_helpers.py:384) is already below a threshold: those are the MLP
evaluations a front-to-back wavefront renderer could skip, and max(T) of the first skipped sample bounds the colour error.
Writes tests/golden/ert_study.txt.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import torch_port as TP          # noqa: E402
import make_convergence_golden as gen        # noqa: E402

T = torch.from_numpy


def skippable(out, thresholds=(1e-2, 1e-3, 1e-4)):
    w = out["weights"]
    trans = 1.0 - torch.cat([torch.zeros_like(w[:, :1]), torch.cumsum(w, -1)[:, :-1]], -1)     # T_i = 1 - sum_{j<i} w_j
    res = {}
    for t in thresholds:
        res[t] = float((trans < t).float().mean())
    first_done = (trans < 1e-2).float().argmax(-1)                                              # 0 when the ray never terminates
    done = (trans < 1e-2).any(-1)
    return res, float(done.float().mean()), float(first_done[done].float().mean()) if bool(done.any()) else float("nan"), float(out["acc_map"].mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--threads", type=int, default=16)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    ro, rd, rgb_t, disp_t, _ = gen.problem()
    rng = np.random.default_rng(4048)
    pc_np, pf_np = gen.params()
    pc, pf = TP.make_params(pc_np, "cpu"), TP.make_params(pf_np, "cpu")
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=gen.LR, betas=(0.9, 0.999))
    probe = np.random.default_rng(7).integers(0, ro.shape[0], 512)
    checkpoints = sorted({300, 1000, 2000, args.steps} & set(range(1, args.steps + 1)) | {args.steps})
    lines = ["# early-ray-termination study on the analytic scene (tests/golden/ert_study.py), CPU, oracle/torch_port.py",
             "# fine pass = 128 samples per ray (64 stratified in disparity + 64 importance samples); fractions over 512 probe rays",
             "# steps  psnr_dB  acc_mean  rays_reaching_T<1e-2  mean_index_of_first_skippable  skippable: T<1e-2  T<1e-3  T<1e-4"]
    t0 = time.time()
    for it in range(1, args.steps + 1):
        idx = rng.integers(0, ro.shape[0], (3, gen.N_RAND))
        rays = lambda g: T(np.stack([ro[idx[g]], rd[idx[g]]], 0))
        batches = [(rays(0), T(rgb_t[idx[0]])), (rays(1), T(rgb_t[idx[1]])), (rays(2), T(disp_t[idx[2]]))]
        opt.zero_grad()
        loss, psnr = TP.spin_step_loss(batches, pc, pf, gen.NEAR, gen.FAR)
        loss.backward()
        opt.step()
        for group in opt.param_groups:
            group["lr"] = gen.LR * (0.1 ** (it / (gen.DECAY * 1000)))
        if it % 100 == 0:
            print(f"step {it} loss {float(loss):.5f} psnr {float(psnr):.2f} ({time.time() - t0:.0f} s)", flush=True)
        if it in checkpoints:
            with torch.no_grad():
                out = TP.render_rays(T(ro[probe]), T(rd[probe]), gen.NEAR, gen.FAR, pc, pf)
                mse = float(torch.mean((out["rgb_map"] - T(rgb_t[probe])) ** 2))
            frac, done, first, acc = skippable(out)
            lines.append(f"{it:6d}  {-10 * np.log10(mse):7.2f}  {acc:8.3f}  {done:20.3f}  {first:29.1f}  {frac[1e-2]:17.3f}  {frac[1e-3]:6.3f}  {frac[1e-4]:6.3f}")
            print(lines[-1], flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ert_study.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
