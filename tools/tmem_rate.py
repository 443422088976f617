"""TMEM -> register drain rate probe (tcgen05.ld), alone and under concurrent MMAs (see spn_tc_tmem_ld_rate)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
reps = 200
for with_mma in (0, 1):
    for nw in (1, 2, 4, 8, 16):
        L.check(L.lib().spn_tc_tmem_ld_rate(nw, reps, with_mma, L.ptr(out), L.stream()))
        torch.cuda.synchronize()
        cyc, mmas = out.tolist()
        nbytes = nw * 32 * 128 * 4 * reps
        extra = f"  MMAs meanwhile: {mmas} -> {cyc / max(mmas, 1):.1f} cycles/MMA" if with_mma else ""
        print(f"warps={nw:2d} mma={with_mma}: {cyc / reps:8.1f} cycles per 128-column drain, {nbytes / cyc:6.1f} B/cycle/SM{extra}")
