"""Stand-in for h5py (absent from this image) over the repository's own HDF5 reader / writer
(learn_region_grow_b200/hdf5.py): ``f = File(name, 'r'); f['points'][:]; f.close()``
(/root/reference/learn_region_grow_util.py:11-20) and ``File(name, 'w').create_dataset(name, data=, dtype=[, compression='gzip',
compression_opts=4])`` (tools/generate_synthetic_rooms.py:112-115, stage_data.py:249-256).  Files written here are real
HDF5; files written by the real h5py with default settings are read directly.  (``.npz`` payloads written by the first
version of this stand-in are still opened for reading.)
"""
import numpy as np

from learn_region_grow_b200.hdf5 import Dataset, Group, Hdf5Error, Hdf5Unsupported, Reader, Writer, SIGNATURE  # noqa: F401

__version__ = '0.0+lrg_b200'


class _NpzFile:
    def __init__(self, name):
        self.filename, self.mode = name, 'r'
        with np.load(name, allow_pickle=False) as z:
            self._data = {k: z[k] for k in z.files}

    def __getitem__(self, key):
        return self._data[key]

    def __contains__(self, key):
        return key in self._data

    def keys(self):
        return self._data.keys()

    def close(self):
        self._data = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def File(name, mode='r', **kw):
    if mode == 'r':
        with open(name, 'rb') as f:
            magic = f.read(4)
        if magic == b'PK\x03\x04':
            return _NpzFile(name)
        return Reader(name)
    if mode in ('w', 'w-', 'x'):
        return Writer(name)
    raise Hdf5Unsupported("File mode %r (only 'r' and 'w')" % mode)
