"""N>1 host logic on CPU: world_size-2 gloo.  Each rank computes oracle gradients on its ray shard; the summed flat
vectors scaled by 1/W must equal the single-process full-batch gradients (what Trainer.apply_gradients relies on)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import nerf_oracle as O
from oracle import train_oracle as TO

dist_mod = importlib.import_module("spin-nerf_b200.dist")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem():
    rng = np.random.default_rng(0)
    n = 16
    c2w = np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 0.0]], np.float32)
    ro, rd = O.get_rays(12, 16, 14.4, c2w)
    sel = rng.choice(12 * 16, n, replace=False)
    rb = O.make_ray_batch(ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel], 1.2, 8.0)
    target = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    return rb, target


def _grads(rb, target, n_total):
    pc, pf = O.init_params(1), O.init_params(2)
    # global-mean MSE: gradient of sum over the shard divided by the GLOBAL element count
    g_out = lambda o: dict(rgb_map=(2.0 / (n_total * 3) * (o["rgb_map"] - target)).astype(np.float32),
                           rgb0=(2.0 / (n_total * 3) * (o["rgb0"] - target)).astype(np.float32))
    _, gc, gf = TO.render_with_grads(rb, pc, pf, 16, 16, True, True, g_out)
    return np.concatenate([gc[k].reshape(-1) for k, _ in O.PARAM_SHAPES] + [gf[k].reshape(-1) for k, _ in O.PARAM_SHAPES])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w, _ = dist_mod.init_from_env("gloo")
    sh = dist_mod.RaySharder(r, w)
    rb, target = _problem()
    lo, hi = sh.bounds(rb.shape[0])
    # per-rank MEAN over its shard (what each rank's Trainer computes) == global-sum-gradient * W
    flat = torch.from_numpy(_grads(rb[lo:hi], target[lo:hi], hi - lo))
    scale = dist_mod.allreduce_sum_([flat])
    if r == 0:
        np.save(out, (flat * scale).numpy())
    torch.distributed.destroy_process_group()


def _frames_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = dist_mod.init_from_env("gloo")
    n = 7                                                    # not a multiple of the world size: rank 1 has a spare row
    mine = dist_mod.frames_for_rank(n, r, w)
    per = (n + w - 1) // w
    local = torch.full((per, 2, 3, 4), -1.0)
    for j, i in enumerate(mine):
        local[j] = float(i) + torch.arange(24.0).reshape(2, 3, 4) / 64
    every = dist_mod.gather_frames(local, n, r, w)
    only0 = dist_mod.gather_frames(local, n, r, w, dst=0)
    assert (only0 is None) == (r != 0)
    if r == 0:
        np.save(out, torch.stack([every, only0]).numpy())
    else:
        assert every.shape[0] == n
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_frame_gather_restores_frame_order(tmp_path):
    """render_path_sharded's collective (config 5): round-robin frames come back in frame order, on all ranks or on one."""
    out = str(tmp_path / "frames.npy")
    mp.spawn(_frames_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    want = np.arange(7.0)[:, None, None, None] + np.arange(24.0).reshape(2, 3, 4) / 64
    np.testing.assert_array_equal(got[0], want.astype(np.float32)); np.testing.assert_array_equal(got[1], want.astype(np.float32))
    one = torch.arange(5.0)[:, None].repeat(1, 3)
    assert torch.equal(dist_mod.gather_frames(one, 5, 0, 1), one)


def _video_worker(rank, world, port, out):
    """render_path_sharded under gloo with the C library replaced by the call recorder (zero-filled outputs): the frame
    dealing, the per-rank device buffer, the gather and the hand-over, end to end in two processes."""
    import importlib
    import test_host_glue_dry_run as dry
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    rec = dry.Recorder()
    for m in (dry.L, dry.ops, dry.render_mod):
        m.lib = (lambda rec=rec: rec); m.ptr = dry._ptr; m.stream = (lambda: 0)
    dry.ops._empty = lambda shape, like, dtype=torch.float32: torch.zeros(shape, device=like.device, dtype=dtype)
    dry.nerf_mod.NeRF._sync = lambda self: (self.flat_params(), torch.zeros(64, dtype=torch.uint8))
    # zero-filled chunk outputs: chunk_forward allocates with torch.empty, make that deterministic too
    real_empty = torch.empty
    dry.render_mod.torch.empty = lambda *a, **k: real_empty(*a, **k).zero_()
    dist_mod.init_from_env("gloo")
    spn = importlib.import_module("spin-nerf_b200")
    nets = [spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True) for _ in range(2)]
    kw = dict(network_query_fn=None, network_fn=nets[0], network_fine=nets[1], N_samples=8, N_importance=8, lindisp=True,
              white_bkgd=True, perturb=0., raw_noise_std=0., use_viewdirs=True, ndc=False, near=1.2, far=8.0)
    poses = np.tile(np.eye(4, dtype=np.float32)[None, :3, :], (5, 1, 1))
    rgbs, disps = spn.render_path_sharded(poses, [12, 16, 14.4], 64, kw, render_factor=2)
    only = spn.render_path_sharded(poses, [12, 16, 14.4], 64, kw, dst=1)
    mine = dist_mod.frames_for_rank(5, rank, world)
    ok = (rgbs.shape == (5, 6, 8, 3) and disps.shape == (5, 6, 8) and rec.names().count("spn_get_rays") == 2 * len(mine)
          and ((only[0] is None) == (rank != 1)) and (rank != 1 or only[0].shape == (5, 12, 16, 3)))
    np.save(out + f".{rank}.npy", np.array([int(ok)]))
    torch.distributed.destroy_process_group()


def _trainer_worker(rank, world, port, out):
    """Trainer.step on two ranks over gloo with the call recorder: each rank renders its contiguous half of every ray group,
    the single gradient buffer is all-reduced once, Adam is told to average (grad_scale = 1 / world)."""
    import importlib
    import test_host_glue_dry_run as dry
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    rec = dry.Recorder()
    for m in (dry.L, dry.ops, dry.render_mod):
        m.lib = (lambda rec=rec: rec); m.ptr = dry._ptr; m.stream = (lambda: 0)
    dry.nerf_mod.NeRF._sync = lambda self: (self.flat_params(), torch.zeros(64, dtype=torch.uint8))
    dist_mod.init_from_env("gloo")
    tmod = dry.trainer_mod

    def fake_backward(cfg, k, net_c, net_f, g, gc, gf, scratch=None, ws=None, detach_range=(0, 0)):
        gc.fill_(float(rank + 1)); gf.fill_(float(10 * (rank + 1)))       # what a rank's backward would accumulate
    tmod.chunk_backward = fake_backward
    tr = dry._trainer(rank=rank, world=world)
    with torch.no_grad():                                   # replicas that start different (a per-rank torch seed) ...
        tr.net_c.flat_params().add_(float(rank)); tr.net_f.flat_params().add_(float(rank))
    dist_mod.broadcast_parameters([tr.net_c, tr.net_f])     # ... are made identical to rank 0's
    same = dry._trainer(rank=0, world=1)
    bcast_ok = torch.equal(tr.net_c.flat_params(), same.net_c.flat_params()) and torch.equal(tr.net_f.flat_params(), same.net_f.flat_params())
    tr.step(*dry._batch(8))
    cfg = rec.calls[1][1][0]._obj
    adam = [c for c in rec.calls if c[0] == "spn_adam_step"]
    ok = (bcast_ok and cfg.n_rays == 12 and len(adam) == 2 and abs(adam[0][1][10] - 0.5) < 1e-12           # grad_scale = 1 / world
          and float(tr.grads[0].min()) == float(tr.grads[0].max()) == 3.0                    # 1 + 2 summed over the ranks
          and float(tr.grads[1].min()) == float(tr.grads[1].max()) == 30.0 and tr.global_step == 1)
    np.save(out + f".{rank}.npy", np.array([int(ok)]))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_trainer_step_glue(tmp_path):
    out = str(tmp_path / "trainer")
    mp.spawn(_trainer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert all(int(np.load(out + f".{r}.npy")[0]) == 1 for r in range(2))


@pytest.mark.timeout(300)
def test_two_rank_sharded_video_render_glue(tmp_path):
    out = str(tmp_path / "video")
    mp.spawn(_video_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert all(int(np.load(out + f".{r}.npy")[0]) == 1 for r in range(2))


def test_sharder_partitions_exactly():
    for n in (1, 7, 1024, 8192):
        for w in (1, 2, 3, 8):
            b = [dist_mod.RaySharder(r, w).bounds(n) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    assert dist_mod.frames_for_rank(120, 3, 8) == list(range(3, 120, 8))
    assert sorted(sum((dist_mod.frames_for_rank(120, r, 8) for r in range(8)), [])) == list(range(120))


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    rb, target = _problem()
    want = _grads(rb, target, rb.shape[0])
    # fp32 summation order differs (per-shard partial sums vs one pass): compare at 1e-3 of the gradient scale
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-3 * np.abs(want).max())
