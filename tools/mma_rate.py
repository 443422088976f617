"""tcgen05.mma issue-rate probe for K-major vs MN-major operands (see spn_tc_mma_rate)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for n in (256, 128, 64):
    for a in (0, 1):
        for b in (0, 1):
            for reps in (64, 1024):
                L.check(L.lib().spn_tc_mma_rate(a, b, n, reps, L.ptr(out), L.stream()))
                torch.cuda.synchronize()
                print(f"N={n} A={'MN' if a else 'K'} B={'MN' if b else 'K'} reps={reps}: {out.item()/reps:.1f} cycles/MMA")
