"""Shim S2 (part) -- `pybullet_data.getDataPath()`; the shim world has no data files (TEST INFRASTRUCTURE)."""
import os


def getDataPath():
    return os.path.dirname(os.path.abspath(__file__))
