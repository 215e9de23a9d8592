"""Per-room segmentation statistics on the device (test_region_grow.py:319-349) for host arrays: thin wrapper over
``lrg_segmentation_metrics`` (include/lrg_b200.h).  ``Engine.room_metrics`` scores the labels an engine holds."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ROOM_METRICS_DTYPE


def segmentation_metrics(obj_ids, cluster_labels, return_label2=False):
    """obj_ids / cluster_labels: lists of per-room integer arrays of equal lengths.  Returns one ROOM_METRICS_DTYPE row per
    room (nmi, ami, ars, prc, rcl, iou, ...) and optionally the per-room cluster_label2 arrays."""
    lib = _lib.lib()
    _lib.require_gpu()
    if len(obj_ids) != len(cluster_labels):
        raise ValueError('obj_ids and cluster_labels must list the same rooms')
    counts = [len(o) for o in obj_ids]
    for n, l in zip(counts, cluster_labels):
        if len(l) != n:
            raise ValueError('a room has %d object ids but %d cluster labels' % (n, len(l)))
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    total = int(off[-1])
    cat = lambda rooms: np.ascontiguousarray(np.concatenate([np.asarray(r).astype(np.int32) for r in rooms]) if total else np.zeros(0, np.int32))
    obj, lab = cat(obj_ids), cat(cluster_labels)
    out = np.zeros(len(counts), dtype=ROOM_METRICS_DTYPE)
    bufs = []
    try:
        def dev(nbytes):
            p = C.c_void_p()
            _lib.check(lib.lrg_malloc(C.byref(p), max(nbytes, 4)))
            bufs.append(p)
            return p
        d_obj, d_lab = dev(obj.nbytes), dev(lab.nbytes)
        d_l2 = dev(lab.nbytes) if return_label2 else None
        if total:
            _lib.check(lib.lrg_memcpy_h2d(d_obj, _lib.ptr(obj), obj.nbytes))
            _lib.check(lib.lrg_memcpy_h2d(d_lab, _lib.ptr(lab), lab.nbytes))
        _lib.check(lib.lrg_segmentation_metrics(len(counts), _lib.ptr(off), d_obj, d_lab, _lib.ptr(out), d_l2, None))
        if return_label2:
            l2 = np.zeros(total, np.int32)
            if total:
                _lib.check(lib.lrg_memcpy_d2h(_lib.ptr(l2), d_l2, l2.nbytes))
            return out, [l2[off[i]:off[i + 1]] for i in range(len(counts))]
        return out
    finally:
        for p in bufs:
            lib.lrg_free(p)
