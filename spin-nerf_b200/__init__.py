"""spinnerf_b200 — B200-native volumetric-rendering hot path of SPIn-NeRF (DS_NeRF/run_nerf.py
render -> batchify_rays -> render_rays), hand-written sm_100a CUDA behind a C ABI.

The directory name carries a hyphen (spin-nerf_b200), so import it with
    importlib.import_module("spin-nerf_b200")
or put  spin-nerf_b200/dropin  on PYTHONPATH and `import run_nerf_helpers` (the reference's seam).
"""
from . import _lib
from ._lib import PREC_BF16, PREC_FP32, LIB_PATH, MLP_NPARAMS
from . import ops
from .nerf import NeRF, NeRF_RGB, default_precision, set_default_precision
from .embed import Embedder, get_embedder
from .render import (render, render_rays, batchify_rays, render_path, render_path_sharded, render_rays_composed, batchify,
                     run_network, create_nerf)
from . import frame_io, lpips_patch

__all__ = ["ops", "NeRF", "NeRF_RGB", "render", "render_rays", "batchify_rays", "render_path", "render_path_sharded", "batchify", "run_network",
           "create_nerf", "get_embedder", "Embedder",
           "frame_io", "lpips_patch", "PREC_BF16",
           "PREC_FP32", "set_default_precision", "default_precision"]
__version__ = "0.1.0"
