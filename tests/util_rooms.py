"""Shared helpers for the parity tests."""
import os

import numpy as np

from conftest import GOLDEN


def golden_room(seed):
    g = np.load(os.path.join(GOLDEN, 'driver_trace_%d.npz' % seed))
    return g['points'], g['order']


def unpack_mask(words, n=512):
    bits = np.unpackbits(np.asarray(words, dtype='<u4').view(np.uint8), bitorder='little')
    return bits[:n].astype(bool)


def idx_crc(idx):
    idx = np.asarray(idx, dtype=np.uint64)
    return int(np.sum((np.arange(len(idx), dtype=np.uint64) + 1) * idx) & np.uint64(0xFFFFFFFF))
