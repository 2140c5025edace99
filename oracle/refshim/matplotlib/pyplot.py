"""Stub of matplotlib.pyplot: only `cm.get_cmap(name, n)` is reached from the
reference's detect_grid (utils/util_cylinder.py:1731).  Colours only affect the
overlay image, never the JSON."""
import colorsys


class _Cmap:
    def __init__(self, n):
        self.n = max(int(n), 1)

    def __call__(self, i):
        r, g, b = colorsys.hsv_to_rgb((i % self.n) / self.n, 1.0, 1.0)
        return (r, g, b, 1.0)


class _CM:
    @staticmethod
    def get_cmap(name, n=256):
        return _Cmap(n)


cm = _CM()
