"""cp.async.bulk (1-D TMA bulk copy) throughput per SM vs copy size, copies in flight, issuing lanes, the kind of
mbarrier wait (suspending try_wait vs polling test_wait) and private vs shared completion barrier — spn_tc_bulk_rate."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
GRID = 148
out = torch.zeros(GRID, dtype=torch.int64, device="cuda")
nbytes = 64 << 20
src = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); src.fill_(1)
MODES = {0: "private barriers, try_wait", 1: "private barriers, polling", 2: "ONE barrier per slot for all lanes"}
for copy in (4096, 16384):
    for lanes, depth, mode in ((1, 4, 0), (8, 4 if copy == 4096 else 1, 0), (8, 4 if copy == 4096 else 1, 2), (8, 2 if copy == 4096 else 1, 2),
                               (4, 4 if copy == 4096 else 2, 2)):
        iters = max(64, (4 << 20) // copy)
        for rep in range(2):
            L.check(L.lib().spn_tc_bulk_rate(L.ptr(src), nbytes, copy, depth, iters, GRID, lanes + 100 * mode, L.ptr(out), L.stream()))
            torch.cuda.synchronize()
        cyc = out.float().mean().item()
        print(f"L2 64 MiB: copy={copy:6d} lanes={lanes:2d} depth={depth:2d} [{MODES[mode]}]: "
              f"{copy * iters * lanes / cyc:6.1f} B/cycle/SM  ({cyc / iters:7.0f} cycles per round)")
