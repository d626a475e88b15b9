"""Shim for torchinfo (imported at CamRaDepth.py:9, used only under __main__)."""


def summary(*a, **k):
    raise NotImplementedError("torchinfo shim")
