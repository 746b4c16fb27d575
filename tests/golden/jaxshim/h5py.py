"""`import h5py` for the reference's loaders and writers in an image without h5py.  Reading goes through the repo's
pure-Python HDF5 reader (oracle/h5lite.py), enough for ``h5py.File(name, 'r')['segments'][:]`` (optimize/dataio.py:114-115).
Writing (``File(name, 'a')`` / ``create_group`` / ``create_dataset``, optimize/simulate.py:136-165) is CAPTURED, not encoded:
everything written to a path accumulates in ``WRITTEN[path]`` as {"group/sub/dataset": ndarray} for the fixture generator."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))
from oracle.h5lite import H5Lite  # noqa: E402


class _Dataset:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, key):
        return self._arr[key]

    def __array__(self, dtype=None, copy=None):
        return self._arr if dtype is None else self._arr.astype(dtype)


WRITTEN = {}


class _Group:
    def __init__(self, store, prefix):
        self._store, self._prefix = store, prefix

    def create_group(self, name):
        return _Group(self._store, self._prefix + name + "/")

    def create_dataset(self, name, data=None, **kw):
        import numpy as np
        self._store[self._prefix + name] = np.array(data)


class File(_Group):
    def __init__(self, name, mode="r"):
        if mode != "r":
            _Group.__init__(self, WRITTEN.setdefault(name, {}), "")
            self._f = None
            return
        self._f = H5Lite(name)

    def __getitem__(self, key):
        return _Dataset(self._f.read("/" + key.lstrip("/")))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
