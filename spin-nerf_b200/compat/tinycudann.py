"""Import-only stand-in: DS_NeRF/run_nerf.py:22 imports run_nerf_helpers_tcnn unconditionally; the hash-grid
model itself is out of scope (use --no_tcnn)."""
def __getattr__(name):
    raise RuntimeError("tinycudann is not available; run with --no_tcnn (the B200 path implements the 8x256 MLP)")
