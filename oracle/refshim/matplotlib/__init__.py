"""Stub: the reference imports matplotlib.pyplot only for overlay colours
(utils/util_cylinder.py:8, :1729-1732).  Not installed in this image."""
