"""Shim for imageio (image dumps of the reference trainer); writes are no-ops, reads are unsupported."""


def imwrite(*a, **k):
    return None


def imread(*a, **k):
    raise ImportError("imageio is not installed in this image")
