"""Headless tkinter stand-in: DS_NeRF/run_nerf.py:928-960 starts a GUI thread unconditionally."""
class _W:
    def __init__(self, *a, **k): pass
    def __getattr__(self, name):
        return lambda *a, **k: None
Tk = Label = Button = Entry = Scale = Checkbutton = Frame = _W
class IntVar(_W):
    def get(self): return 0
class DoubleVar(_W):
    def get(self): return 0.0
class StringVar(_W):
    def get(self): return ""
HORIZONTAL = "horizontal"
END = "end"
