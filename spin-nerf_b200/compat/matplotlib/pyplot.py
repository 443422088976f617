class _Fig:
    def set_dpi(self, *a, **k):
        pass
def _noop(*a, **k):
    return None
def gcf():
    return _Fig()
subplot = imshow = savefig = clf = plot = xlabel = ylabel = colorbar = figure = close = title = _noop
