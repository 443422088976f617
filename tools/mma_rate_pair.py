"""Rate of CTA-pair tcgen05.mma (cta_group::2, M = 256): A from shared memory vs tensor memory, N, accumulator rotation,
concurrent epilogue-like TMEM traffic (spn_tc_mma_rate_pair)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for ts in (0, 1):
    for n, nacc in ((256, 1), (128, 1), (128, 2), (64, 1), (64, 4)):
        for ldw in (0, 16):
            for reps in (64, 1024):
                L.check(L.lib().spn_tc_mma_rate_pair(ts, n, reps, nacc, ldw, L.ptr(out), L.stream()))
                torch.cuda.synchronize()
                print(f"A={'TMEM' if ts else 'smem'} N={n} nacc={nacc} tmem_traffic_warps={ldw:2d} reps={reps:4d}: {out[0].item()/reps:6.1f} cycles/MMA")
