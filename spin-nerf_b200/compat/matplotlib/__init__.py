"""No-op matplotlib stand-in (only pyplot sanity plots are used, DS_NeRF/run_nerf.py:1581-1597)."""
def use(*a, **k):
    pass
