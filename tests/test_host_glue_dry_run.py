"""Dry run of the Python host glue on CPU tensors with the C library replaced by a recorder: every C entry point
returns 0 and touches nothing, so values are garbage — what is checked is the part that has no GPU in it: which
entry points a step calls and in which order, buffer shapes / contiguity at the boundary (the recorder keeps _lib.ptr's
contiguity check), ray-range bookkeeping, sharding of patches, pooled-buffer reuse.  The numerics of the same paths are
checked on the GPU (tests/test_gpu_parity.py, tests/test_zz_gpu_next_rows.py)."""
import ctypes
import importlib

import numpy as np
import pytest
import torch

spn = importlib.import_module("spin-nerf_b200")
L = importlib.import_module("spin-nerf_b200._lib")
ops = importlib.import_module("spin-nerf_b200.ops")
render_mod = importlib.import_module("spin-nerf_b200.render")
trainer_mod = importlib.import_module("spin-nerf_b200.trainer")
nerf_mod = importlib.import_module("spin-nerf_b200.nerf")


class Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("spn_"):
            raise AttributeError(name)
        if name in ("spn_mlp_stash_bytes", "spn_mlp_bwd_workspace_bytes", "spn_mlp_packed_bytes"):
            return lambda *a: 64
        if name == "spn_launch_count":
            return lambda *a: 0

        def fn(*a):
            self.calls.append((name, a))
            return 0
        return fn

    def names(self):
        return [c[0] for c in self.calls]


def _ptr(t):
    if t is None:
        return None
    if not t.is_contiguous():
        raise RuntimeError("expected a contiguous tensor")
    return t.data_ptr() if t.numel() else 256


@pytest.fixture
def rec(monkeypatch):
    r = Recorder()
    for m in (L, ops, render_mod):
        monkeypatch.setattr(m, "lib", lambda r=r: r, raising=True)
        monkeypatch.setattr(m, "ptr", _ptr, raising=True)
        monkeypatch.setattr(m, "stream", lambda: 0, raising=True)
    monkeypatch.setattr(nerf_mod.NeRF, "_sync", lambda self: (self.flat_params(), torch.zeros(64, dtype=torch.uint8)))
    return r


def _trainer(rank=0, world=1):
    nets = [spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).seeded_init_(s)
            for s in (1, 2)]
    return trainer_mod.Trainer(nets[0], nets[1], hwf=(24, 32, 28.8), sharder=trainer_mod.RaySharder(rank, world))


def _batch(n):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g)
    return (r(2, n, 3), r(n, 3), r(2, n, 3), r(n, 3), r(2, n, 3), r(n))


def test_step_call_sequence_and_ranges(rec):
    tr = _trainer()
    loss, psnr = tr.step(*_batch(8))
    assert rec.names() == ["spn_build_ray_batch", "spn_render_rays_fwd", "spn_train_losses", "spn_render_rays_bwd",
                           "spn_adam_step", "spn_adam_step"]
    cfg = rec.calls[1][1][0]._obj
    assert (cfg.n_rays, cfg.ncols, cfg.n_samples, cfg.n_importance) == (24, 11, 64, 64)
    assert cfg.flags & L.F_PERTURB and not cfg.flags & L.F_DETACH_WEIGHTS
    n1n2n3 = rec.calls[2][1][6:9]
    assert n1n2n3 == (8, 8, 8)
    gr = rec.calls[3][1][2]._obj
    assert (gr.detach_begin, gr.detach_end) == (8, 16)            # the masked rays of the kept view, run_nerf.py:1465
    assert tr.global_step == 1 and loss.shape == () and psnr.shape == ()


def test_step_shards_rays_per_rank(rec):
    tr = _trainer(rank=1, world=2)
    tr.apply_gradients = lambda: None                              # no process group in this dry run
    tr.step(*_batch(8))
    cfg = rec.calls[1][1][0]._obj
    assert cfg.n_rays == 12 and rec.calls[2][1][6:9] == (4, 4, 4)


def test_step_with_lpips_renders_patches_as_one_test_kwargs_chunk(rec):
    tr = _trainer()
    poses = torch.eye(4)[None, :3, :].repeat(4, 1, 1)
    patches = [(2, 3, 5, 6), (0, 0, 5, 6), (20, 28, 5, 6), (7, 1, 5, 6)]        # the third sticks out: clamped to 4x4
    shapes = [(5, 6), (5, 6), (4, 4), (5, 6)]
    targets = [torch.zeros(1, 3, h, w) for h, w in shapes]
    seen = []

    def lpips_fn(pred, target):
        seen.append(tuple(pred.shape)); assert pred.shape == target.shape
        return ((pred - target) ** 2).mean((1, 2, 3))
    loss, _ = tr.step_with_lpips(_batch(8), poses, patches, targets, lpips_fn, (24, 32, 28.8))
    assert rec.names() == (["spn_build_ray_batch", "spn_render_rays_fwd", "spn_train_losses", "spn_render_rays_bwd"] +
                           ["spn_get_rays"] * 4 + ["spn_build_ray_batch", "spn_render_rays_fwd", "spn_render_rays_bwd",
                                                   "spn_adam_step", "spn_adam_step"])
    assert seen == [(1, 3, h, w) for h, w in shapes]
    cfg = rec.calls[9][1][0]._obj
    assert cfg.n_rays == 30 * 3 + 16
    assert cfg.flags & L.F_DETACH_WEIGHTS and not cfg.flags & L.F_PERTURB and cfg.raw_noise_std == 0.0
    gr = rec.calls[10][1][2]._obj
    assert gr.g_rgb and not gr.g_rgb0 and not gr.g_disp and not gr.g_disp0      # only the fine rgb carries the LPIPS gradient
    assert tr.global_step == 1 and loss.shape == ()


def test_lpips_patches_are_dealt_round_robin_to_ranks(rec):
    poses = torch.eye(4)[None, :3, :].repeat(4, 1, 1)
    patches = [(2, 3, 5, 6)] * 4
    targets = [torch.full((1, 3, 5, 6), float(i)) for i in range(4)]
    got = []
    for rank in (0, 1):
        tr = _trainer(rank=rank, world=2)
        seen = []
        tr.lpips_patch_backward(poses, patches, targets, lambda p, t: (seen.append(float(t.mean())), (p * 0).mean((1, 2, 3)))[1],
                                (24, 32, 28.8), batch_size=4)
        got.append(seen)
    assert got == [[0.0, 2.0], [1.0, 3.0]]
    assert _trainer(rank=1, world=2).lpips_patch_backward(poses[:1], patches[:1], targets[:1], None, (24, 32, 28.8)) == 0


def test_step_from_pool_call_sequence(rec):
    tr = _trainer()
    M = 100
    pool = torch.rand(2, M, 3); rgb = torch.rand(M, 3); disp = torch.rand(M)
    idx = torch.randint(0, M, (3, 8))
    tr.step_from_pool(pool, rgb, disp, idx)
    assert rec.names() == ["spn_gather_ray_batch", "spn_render_rays_fwd", "spn_train_losses", "spn_render_rays_bwd",
                           "spn_adam_step", "spn_adam_step"]


def test_gradient_vectors_share_one_buffer():
    tr = _trainer()
    gc, gf = tr.grads
    assert gc.numel() == gf.numel() == spn.MLP_NPARAMS and gc.is_contiguous() and gf.is_contiguous()
    assert gc.data_ptr() == tr.grad_all.data_ptr() and (gf.data_ptr() - gc.data_ptr()) % 512 == 0
    assert gf.data_ptr() + 4 * gf.numel() <= tr.grad_all.data_ptr() + 4 * tr.grad_all.numel()


def test_step_with_sparse_depth_group(rec):
    """--colmap_depth --depth_loss: the fourth render call rides in the chunk as a fourth ray range with a depth_map gradient."""
    tr = _trainer()
    g = torch.Generator().manual_seed(1)
    loss, _ = tr.step(*_batch(8), rays_depth=torch.rand(2, 5, 3, generator=g), target_depth=torch.rand(5, generator=g), depth_lambda=0.1)
    assert rec.names() == ["spn_build_ray_batch", "spn_render_rays_fwd", "spn_train_losses", "spn_render_rays_bwd",
                           "spn_adam_step", "spn_adam_step"]
    assert rec.calls[1][1][0]._obj.n_rays == 29 and rec.calls[2][1][6:9] == (8, 8, 8)
    gr = rec.calls[3][1][2]._obj
    assert gr.g_depth and gr.g_rgb and (gr.detach_begin, gr.detach_end) == (8, 16)
    assert loss.shape == ()
    with pytest.raises(RuntimeError):
        tr.step(*_batch(8), rays_depth=torch.rand(2, 5, 3), target_depth=torch.rand(4))


def test_seeded_ranks_consume_the_single_process_random_stream(rec, monkeypatch):
    """Trainer(seed=...): the jitter / resampling / noise tensors of a 2-rank run are the rows of the single-process run's
    tensors that belong to each rank's rays (three ray groups, each split contiguously), so multi-GPU training differs from
    single-GPU training only by the gradient all-reduce's summation order."""
    seen = {}
    real = trainer_mod.chunk_forward

    def spy(tag):
        def f(opts, rays, net_c, net_f, t_rand=None, u=None, noise0=None, noise1=None, train=False, pool=None):
            seen[tag] = [t.clone() for t in (t_rand, u, noise0, noise1)]
            return real(opts, rays, net_c, net_f, t_rand, u, noise0, noise1, train, pool)
        return f
    batch = _batch(8)

    def run(tag, rank, world, from_pool):
        nets = [spn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).seeded_init_(s)
                for s in (1, 2)]
        tr = trainer_mod.Trainer(nets[0], nets[1], hwf=(24, 32, 28.8), sharder=trainer_mod.RaySharder(rank, world), seed=7)
        tr.apply_gradients = lambda: None
        monkeypatch.setattr(trainer_mod, "chunk_forward", spy(tag))
        if from_pool:
            tr.step_from_pool(torch.rand(2, 50, 3), torch.rand(50, 3), torch.rand(50), torch.randint(0, 50, (3, 8)))
        else:
            tr.step(*batch)
    for from_pool in (False, True):
        run("one", 0, 1, from_pool); run("r0", 0, 2, from_pool); run("r1", 1, 2, from_pool)
        rows = lambda r: torch.cat([torch.arange(8 * g + 4 * r, 8 * g + 4 * r + 4) for g in range(3)])
        for t_one, t0, t1 in zip(seen["one"], seen["r0"], seen["r1"]):
            assert t_one.shape[0] == 24 and t0.shape[0] == t1.shape[0] == 12
            assert torch.equal(t0, t_one[rows(0)]) and torch.equal(t1, t_one[rows(1)])


def test_searchsorted_wrapper_validates_like_the_references(rec):
    """ops.searchsorted keeps the argument checks of torchsearchsorted's Python wrapper (searchsorted.py:23-40)."""
    a, v = torch.sort(torch.rand(4, 9), 1)[0], torch.rand(4, 5)
    out = ops.searchsorted(a, v, side='right')
    assert out.dtype == torch.long and tuple(out.shape) == (4, 5)
    assert rec.calls[-1][0] == "spn_searchsorted" and rec.calls[-1][1][3:8] == (4, 4, 9, 5, 0)
    assert tuple(ops.searchsorted(a[:1], v).shape) == (4, 5) and tuple(ops.searchsorted(a, v[:1]).shape) == (4, 5)
    for bad in (lambda: ops.searchsorted(a[0], v), lambda: ops.searchsorted(a, v[0]), lambda: ops.searchsorted(a[:3], v),
                lambda: ops.searchsorted(a, v, out=torch.empty(4, 5)), lambda: ops.searchsorted(a, v, out=torch.empty(3, 5, dtype=torch.long))):
        with pytest.raises(AssertionError):
            bad()
