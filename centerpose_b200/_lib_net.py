"""ctypes signatures of the network-forward part of the C ABI (filled in as it grows)."""
from __future__ import annotations


def bind(L):
    pass
