"""cp.async.bulk (1-D TMA bulk copy) throughput per SM vs copy size, copies in flight and issuing lanes
(spn_tc_bulk_rate).  Finding on B200: one thread retires at most one bulk copy per ~640 cycles."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

spn = importlib.import_module("spin-nerf_b200")
L = spn._lib
GRID = 148
out = torch.zeros(GRID, dtype=torch.int64, device="cuda")
for label, nbytes in (("HBM 2 GiB", 2 << 30), ("L2 64 MiB", 64 << 20)):
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); src.fill_(1)
    for copy in (4096, 16384, 32768):
        for lanes, depth in ((1, 1), (1, 4), (2, 4), (4, 4), (8, 4), (8, 1), (16, 2)):
            if copy * depth * lanes > 200 * 1024:
                continue
            iters = max(64, (4 << 20) // copy)
            for rep in range(2):   # second pass is the warm one for the L2-sized buffer
                L.check(L.lib().spn_tc_bulk_rate(L.ptr(src), nbytes, copy, depth, iters, GRID, lanes, L.ptr(out), L.stream()))
                torch.cuda.synchronize()
            cyc = out.float().mean().item()
            print(f"{label}: copy={copy:6d} lanes={lanes:2d} depth={depth}: {copy * iters * lanes / cyc:6.1f} B/cycle/SM"
                  f"  ({cyc / iters:7.0f} cycles per copy per lane)")
    del src
