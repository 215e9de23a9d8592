"""Stand-in for h5py (absent from this image) covering the reference's dataset access pattern
(/root/reference/learn_region_grow_util.py:11-20): ``f = File(name, 'r'|'w'); f['points'][:]; f.close()``
and ``create_dataset(name, data=..., dtype=...)`` (tools/generate_synthetic_rooms.py:112-115).

Storage is a numpy ``.npz`` archive written at the *same path* the HDF5 file would have; real HDF5 files
need the real h5py (SURVEY.md 8f-4 lists a native HDF5 reader as a later row).
"""
import numpy as np


class _Dataset:
    def __init__(self, arr):
        self._arr = arr
        self.shape = arr.shape
        self.dtype = arr.dtype

    def __getitem__(self, key):
        return self._arr[key]

    def __len__(self):
        return len(self._arr)


class File:
    def __init__(self, name, mode='r', **kw):
        self.filename, self.mode = name, mode
        self._data = {}
        if mode.startswith('r'):
            with open(name, 'rb') as f:
                magic = f.read(8)
            if magic.startswith(b'\x89HDF'):
                raise OSError('%s is a real HDF5 file; install h5py to read it (this stand-in reads .npz payloads)' % name)
            with np.load(name, allow_pickle=False) as z:
                self._data = {k: z[k] for k in z.files}

    def __getitem__(self, key):
        return _Dataset(self._data[key])

    def __contains__(self, key):
        return key in self._data

    def keys(self):
        return self._data.keys()

    def create_dataset(self, name, data=None, dtype=None, shape=None, **kw):
        arr = np.zeros(shape, dtype=dtype) if data is None else np.asarray(data, dtype=dtype)
        self._data[name] = arr
        return _Dataset(arr)

    def close(self):
        if not self.mode.startswith('r') and self._data is not None:
            with open(self.filename, 'wb') as f:
                np.savez(f, **self._data)
        self._data = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False
