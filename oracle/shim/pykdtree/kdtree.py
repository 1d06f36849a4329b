"""Shim of pykdtree.kdtree.KDTree on scipy's cKDTree."""
import numpy as np
from scipy.spatial import cKDTree


class KDTree:
    def __init__(self, data, leafsize=16):
        self._t = cKDTree(np.ascontiguousarray(data), leafsize=leafsize)

    def query(self, q, k=1, **kw):
        d, i = self._t.query(np.ascontiguousarray(q), k=k, workers=-1)
        return d, i
