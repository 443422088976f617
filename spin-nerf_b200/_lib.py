"""ctypes binding of libspinnerf_b200.so (the C ABI declared in include/spinnerf_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Tensors cross the boundary as raw device pointers + sizes only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspinnerf_b200.so")

PREC_BF16, PREC_FP32 = 0, 1
F_LINDISP, F_WHITE_BKGD, F_DETACH_WEIGHTS, F_PERTURB, F_NEED_ALPHA = 1, 2, 4, 8, 16
MLP_NPARAMS = 595844

c_fp = C.c_void_p   # every device pointer travels as void*


class RenderCfg(C.Structure):
    _fields_ = [("n_rays", C.c_int), ("ncols", C.c_int), ("n_samples", C.c_int), ("n_importance", C.c_int),
                ("flags", C.c_int), ("precision", C.c_int), ("raw_noise_std", C.c_float)]


_IO_FIELDS = ["rays", "params_coarse", "params_fine", "packed_coarse", "packed_fine", "t_rand", "u", "noise0",
              "noise1", "rgb_map", "disp_map", "acc_map", "depth_map", "weights", "z_vals", "raw", "alpha", "alpha0",
              "rgb0", "disp0", "acc0", "z_std", "z_coarse", "raw_coarse", "stash_coarse", "stash_fine"]
_GRAD_FIELDS = ["g_rgb", "g_disp", "g_acc", "g_depth", "g_weights", "g_rgb0", "g_disp0", "g_acc0", "grads_coarse",
                "grads_fine", "d_raw_scratch", "workspace"]


class RenderIO(C.Structure):
    _fields_ = [(k, c_fp) for k in _IO_FIELDS]


class RenderGrads(C.Structure):
    _fields_ = [(k, c_fp) for k in _GRAD_FIELDS] + [("detach_begin", C.c_int), ("detach_end", C.c_int)]


_SIGS = {
    # name: (restype, argtypes)
    "spn_version": (C.c_int, []),
    "spn_last_error": (C.c_char_p, []),
    "spn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "spn_profile_enable": (C.c_int, [C.c_int]),
    "spn_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "spn_launch_count": (C.c_longlong, [C.c_int]),
    "spn_get_rays": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp]),
    "spn_ndc_rays": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "spn_build_ray_batch": (C.c_int, [C.c_int, c_fp, c_fp, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, c_fp, c_fp]),
    "spn_embed": (C.c_int, [c_fp, C.c_int64, C.c_int, c_fp, c_fp]),
    "spn_sample_z": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp]),
    "spn_raw2outputs_fwd": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int] + [c_fp] * 7),
    "spn_raw2outputs_bwd": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, C.c_int] + [c_fp] * 7),
    "spn_sample_pdf": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp]),
    "spn_sample_pdf_cdf": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp]),
    "spn_searchsorted": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "spn_merge_sorted": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp]),
    "spn_resample": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "spn_mlp_param_offsets": (C.c_int, [C.POINTER(C.c_int64)]),
    "spn_mlp_packed_bytes": (C.c_size_t, []),
    "spn_mlp_pack_weights": (C.c_int, [c_fp, c_fp, c_fp]),
    "spn_mlp_stash_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "spn_mlp_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "spn_mlp_fwd_points": (C.c_int, [c_fp, c_fp, c_fp, C.c_int64, c_fp, c_fp, C.c_int, c_fp]),
    "spn_mlp_fwd_rays": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, c_fp, C.c_int, C.c_int, c_fp, c_fp, C.c_int, c_fp]),
    "spn_mlp_bwd": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_int64, c_fp, c_fp, C.c_int, c_fp]),
    "spn_tc_selftest_gemm": (C.c_int, [c_fp, c_fp, c_fp, C.c_int, C.c_int, c_fp]),
    "spn_tc_set_trace": (C.c_int, [c_fp]),
    "spn_tc_tmem_ld_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, c_fp, c_fp]),
    "spn_tc_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp]),
    "spn_tc_e4m3_decode": (C.c_int, [c_fp, c_fp, C.c_int, c_fp]),
    "spn_tc_mma_rate_pair": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp]),
    "spn_tc_bulk_rate": (C.c_int, [c_fp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp]),
    "spn_gather_ray_batch": (C.c_int, [C.c_int, c_fp, c_fp, c_fp, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float,
                                       c_fp, c_fp, c_fp, C.c_int, c_fp, c_fp, c_fp]),
    "spn_train_losses": (C.c_int, [c_fp] * 6 + [C.c_int] * 3 + [c_fp] * 7),
    "spn_adam_tick": (C.c_int, [c_fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, c_fp]),
    "spn_adam_step_dev": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_int64, c_fp, C.c_float, C.c_float, C.c_float, C.c_float, c_fp]),
    "spn_adam_step": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, c_fp]),
    "spn_peer_region_bytes": (C.c_size_t, [C.c_int64]),
    "spn_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "spn_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "spn_peer_close": (C.c_int, [c_fp]),
    "spn_peer_free": (C.c_int, [c_fp]),
    "spn_peer_grad_ptr": (C.c_void_p, [c_fp]),
    "spn_peer_allreduce_adam": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint, C.c_int64, C.c_int64, C.c_int64]
                                + [c_fp] * 6 + [C.c_float] * 4 + [C.c_int, C.c_float, c_fp]),
    "spn_render_rays_fwd": (C.c_int, [C.POINTER(RenderCfg), C.POINTER(RenderIO), c_fp]),
    "spn_render_rays_bwd": (C.c_int, [C.POINTER(RenderCfg), C.POINTER(RenderIO), C.POINTER(RenderGrads), c_fp]),
    "spn_render_host": (C.c_int, [C.POINTER(RenderCfg), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(spin-nerf_b200/csrc/build.sh).  There is no CPU / PyTorch fallback for this path.")
        l = C.CDLL(os.environ.get("SPN_LIB_OVERRIDE", LIB_PATH))   # override: A/B timing of experimental builds only
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)      # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().spn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"spinnerf_b200 {what} failed (code {rc}): {msg}")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a contiguous CUDA float32/int64/uint8 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("spinnerf_b200: expected a CUDA tensor (this path has no CPU implementation)")
    if not t.is_contiguous():
        raise RuntimeError("spinnerf_b200: expected a contiguous tensor")
    # empty tensors have data_ptr() == 0; the C side rejects NULL for required buffers but never
    # dereferences anything when the element count is 0, so hand it an aligned non-null token
    return t.data_ptr() if t.numel() else 256


def f32(t):
    """Contiguous fp32 CUDA view/copy of t."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def param_offsets():
    off = (C.c_int64 * 25)()
    check(lib().spn_mlp_param_offsets(off), "spn_mlp_param_offsets")
    return list(off)
