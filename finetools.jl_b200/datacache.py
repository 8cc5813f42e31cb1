"""DataCache: coefficient holder (reference: src/DataCacheModule.jl:65-111).  Only the constant constructor
(:77-89) can cross the C ABI -- there are no callbacks into host code from the device."""
import numpy as np


class DataCache:
    def __init__(self, data):
        if callable(data):
            raise ValueError("DataCache built from a function is not GPU-eligible; only constant caches cross the C ABI")
        self._cache = np.array(data, dtype=np.float64)  # copies, also densifies identity/diagonal inputs

    def size(self):
        return self._cache.shape

    def __call__(self, XYZ=None, tangents=None, feid=0, qpid=0):
        return self._cache

    @property
    def data(self):
        return self._cache
