"""Two-or-more-GPU check of the opt-in peer-memory gradient exchange (csrc/peer_reduce.cu) against the NCCL path.  Run on
one box under a short timeout (the kernels trap after ~4 s without a peer, so a failure is an error, not a hang):

    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/check_peer_allreduce.py

First (NCCL path only) the W-rank run is compared with a single-process run on the full batches: with Trainer(seed=...) the
ranks consume the single-process random stream, so the parameter trajectories must agree up to summation order.  Then each
rank runs the same 4 ray-sharded train steps twice from identical weights — once ending in NCCL all-reduce + flat Adam,
once in the fused peer-memory kernels — and compares the two parameter trajectories (fp32 sums in a different order:
all but < 1 % of the parameters within 1e-5) and, for the peer path, that all ranks hold bit-identical replicas afterwards."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import torch            # noqa: E402
import torch.distributed as dist   # noqa: E402


def same_trajectory(a, b, what, lr=5e-4, steps=4):
    """Parameters after `steps` Adam steps from two runs whose gradients differ only by fp32 summation order: Adam divides by
    sqrt(v), so an element whose gradient is numerically zero may move by +-lr in either run; everything else agrees tightly."""
    d = (a - b).abs()
    frac = float((d > 1e-5).float().mean())
    assert frac < 0.03 and float(d.max()) <= 2 * steps * lr * 1.01, (what, frac, float(d.max()))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    dist.init_process_group("nccl", device_id=dev)
    spn = importlib.import_module("spin-nerf_b200")
    trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
    g = torch.Generator(device=dev); g.manual_seed(0)
    n = 64 * world
    c2w = torch.tensor([[1, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 0.0]], device=dev)
    ro, rd = spn.ops.get_rays(24, 32, 28.8, c2w)
    pool = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
    steps = []
    for _ in range(4):
        ix = torch.randint(0, pool.shape[1], (3, n), device=dev, generator=g)
        steps.append((pool[:, ix[0]], torch.rand(n, 3, device=dev, generator=g), pool[:, ix[1]], torch.rand(n, 3, device=dev, generator=g),
                      pool[:, ix[2]], torch.rand(n, device=dev, generator=g)))
    results = {}
    for mode in ("nccl", "peer"):
        os.environ["SPN_P2P_ALLREDUCE"] = "1" if mode == "peer" else "0"
        nets = [spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).seeded_init_(s).to(dev)
                for s in (1, 2)]
        tr = trainer_mod.Trainer(nets[0], nets[1], lr=5e-4, perturb=0.0, raw_noise_std=0.0, near=1.2, far=8.0,
                                 sharder=trainer_mod.RaySharder(rank, world))
        assert (tr.peer is not None) == (mode == "peer")
        losses = [float(tr.step(*b)[0]) for b in steps]
        torch.cuda.synchronize()
        results[mode] = (losses, nets[0].flat_params().clone(), nets[1].flat_params().clone())
        if tr.peer is not None:
            for p in (nets[0].flat_params(), nets[1].flat_params()):      # replicas stay bit-identical across ranks
                ref = p.clone(); dist.broadcast(ref, 0)
                assert torch.equal(ref, p), "replicas diverged"
            tr.peer.close()
    # ---- sharded == single process: with Trainer(seed=...) every rank consumes the single-process random stream, so W ranks on
    # 1/W of the rays each must walk the trajectory of one process on all rays (stochastic sampling and density noise ON)
    os.environ["SPN_P2P_ALLREDUCE"] = "0"
    traj = {}
    for mode, shard in (("sharded", trainer_mod.RaySharder(rank, world)), ("single", trainer_mod.RaySharder(0, 1))):
        nets = [spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).seeded_init_(s).to(dev)
                for s in (1, 2)]
        tr = trainer_mod.Trainer(nets[0], nets[1], lr=5e-4, perturb=1.0, raw_noise_std=1.0, near=1.2, far=8.0, sharder=shard, seed=5)
        for b in steps:
            tr.step(*b)
        torch.cuda.synchronize()
        traj[mode] = (nets[0].flat_params().clone(), nets[1].flat_params().clone())
    for a, b in zip(traj["sharded"], traj["single"]):
        same_trajectory(a, b, "sharded vs single-process parameters")
    for a, b in zip(results["nccl"][1:], results["peer"][1:]):
        same_trajectory(a, b, "peer-memory vs NCCL parameters")
    assert np.allclose(results["nccl"][0], results["peer"][0], rtol=1e-4)
    if rank == 0:
        print("peer-memory gradient exchange matches NCCL: losses", results["peer"][0])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
