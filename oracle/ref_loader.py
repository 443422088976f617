"""Import the UNMODIFIED reference (DS_NeRF/run_nerf.py + run_nerf_helpers.py) read-only.

Test infrastructure only (see oracle/nerf_oracle.py header).  Works in the build container
where /root/reference exists, or on the GPU box from baseline/_ref/DS_NeRF if the driver's
offline copy is there.  Nothing under tests -m gpu / smoke() / bench's own arm uses this.

Recipe: stub the modules run_nerf.py imports but the image lacks
(matplotlib, imageio, lpips, tinycudann, tkinter, configargparse), neutralise
torch.cuda.set_device(0) (run_nerf.py:39 raises without a driver) and switch autograd
anomaly mode back off (run_nerf_helpers.py:5 turns it on globally at import).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_CANDIDATES = [
    os.environ.get("SPN_REFERENCE_DIR", ""),
    "/root/reference/DS_NeRF",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "DS_NeRF"),
]


def reference_dir():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "run_nerf_helpers.py")):
            return c
    return None


def available() -> bool:
    return reference_dir() is not None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(anomaly: bool = False):
    """Returns (run_nerf_helpers, run_nerf) modules of the reference."""
    import torch

    d = reference_dir()
    if d is None:
        raise RuntimeError("reference not found (expected /root/reference/DS_NeRF)")
    if "spn_ref_run_nerf" in sys.modules:
        return sys.modules["spn_ref_helpers"], sys.modules["spn_ref_run_nerf"]
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("imageio"); _stub("tinycudann"); _stub("configargparse")
    _stub("lpips", LPIPS=object)
    _stub("tkinter")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("run_nerf_helpers", "run_nerf")}
    set_device = torch.cuda.set_device
    torch.cuda.set_device = lambda *_a, **_k: None
    sys.path.insert(0, d)
    try:
        for k in ("run_nerf_helpers", "run_nerf"):
            sys.modules.pop(k, None)
        helpers = importlib.import_module("run_nerf_helpers")
        run_nerf = importlib.import_module("run_nerf")
    finally:
        torch.cuda.set_device = set_device
        sys.path[:] = saved_path
        torch.autograd.set_detect_anomaly(anomaly)
    # keep them under private names so a later `import run_nerf_helpers` can resolve to OUR drop-in
    sys.modules["spn_ref_helpers"] = helpers
    sys.modules["spn_ref_run_nerf"] = run_nerf
    for k, v in saved_mods.items():
        if v is not None:
            sys.modules[k] = v
        else:
            sys.modules.pop(k, None)
    return helpers, run_nerf


def reference_nets(p_coarse: dict, p_fine: dict | None):
    """Reference NeRF modules (helpers:74-156) loaded with oracle-style param dicts."""
    import torch

    helpers, _ = load()
    nets = []
    for p in (p_coarse, p_fine):
        if p is None:
            nets.append(None)
            continue
        net = helpers.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27,
                           use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        nets.append(net)
    return nets
