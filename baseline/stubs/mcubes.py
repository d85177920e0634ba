"""Shim for PyMCubes (not installable here: no network).  The reference imports it at module level
(models/neus_hint_model.py:6) but only extract_geometry calls it."""


def marching_cubes(*a, **k):
    raise ImportError("PyMCubes is not installed in this image; extract_geometry needs the real package")
