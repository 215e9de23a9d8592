"""Python host of the sm_100a LRGNet engine: thin object wrappers over the C ABI (include/lrg_b200.h).

``Engine`` mirrors the life cycle of the reference's network object + session
(/root/reference/test_region_grow.py:86-94): construct with the LrgNet constructor arguments, ``load_weights`` from
a checkpoint or a dict of variables, ``forward`` on host arrays, and -- the part the reference does in Python --
``segment_rooms`` which grows every room on the device without a host round trip per step.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import GrowParams, LrgError, ROOM_METRICS_DTYPE, ROOM_STATS_DTYPE, STEP_TRACE_DTYPE


def channel_lists(lite):
    """learn_region_grow_util.py:77-85"""
    if lite == 0 or lite is None:
        return [64, 64, 64, 128, 512], [256, 128]
    if lite == 1:
        return [64, 64], [64]
    if lite == 2:
        return [64, 64, 256], [64, 64]
    raise ValueError('lite must be 0/None, 1 or 2')


def variable_shapes(feature_size=13, lite=0):
    """(name, shape) of the trainable variables in graph-construction order (util.py:106-162) = blob order."""
    conv, conv2 = channel_lists(lite)
    out = []
    for prefix in ('lrg_', 'lrg_neighbor_'):
        for i, c in enumerate(conv):
            out.append((prefix + 'kernel%d' % i, (1, feature_size if i == 0 else conv[i - 1], c)))
            out.append((prefix + 'bias%d' % i, (c,)))
    for prefix in ('lrg_add_', 'lrg_remove_'):
        for i, c in enumerate(conv2 + [2]):
            out.append((prefix + 'kernel%d' % i, (1, conv[-1] * 2 + conv[1] if i == 0 else conv2[i - 1], c)))
            out.append((prefix + 'bias%d' % i, (c,)))
    return out


def pack_weights(tensors, feature_size=13, lite=0):
    """dict of checkpoint variables -> the flat float32 blob lrg_engine_load_weights expects."""
    parts = []
    for name, shape in variable_shapes(feature_size, lite):
        if name not in tensors:
            raise KeyError('checkpoint has no variable %r' % name)
        a = np.asarray(tensors[name], dtype=np.float32)
        if a.shape != tuple(shape):
            raise ValueError('variable %s has shape %s, expected %s' % (name, a.shape, tuple(shape)))
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


class Engine:
    def __init__(self, batch_size=1, seq_len=1, num_inlier_points=512, num_neighbor_points=512, feature_size=13,
                 lite=0, device=0, forward_mode=None):
        self.lib = _lib.lib()
        _lib.require_gpu()
        self.B = batch_size * seq_len
        self.Ni, self.Nj, self.F = num_inlier_points, num_neighbor_points, feature_size
        self.lite = 0 if lite is None else lite
        self.device = device
        self._h = C.c_void_p()
        _lib.check(self.lib.lrg_engine_create(C.byref(self._h), device, feature_size, num_inlier_points,
                                              num_neighbor_points, self.lite, self.B))
        self._weights = None
        self._room_offsets = None
        if forward_mode is None:
            forward_mode = {'': 0, 'auto': 0, 'fma': 1, 'tensor': 2, 'tf32': 2, 'f16': 3}[os.environ.get('LRG_FORWARD_MODE', '').lower()]
        if forward_mode:
            self.set_forward_mode(forward_mode)

    def set_forward_mode(self, mode):
        """0 auto (3xFP16 tensor tiles, repeated with 3xTF32 if an activation leaves the fp16 range), 1 fp32-FMA kernels,
        2 tcgen05 3xTF32, 3 tcgen05 3xFP16 without the fallback (tensor modes: full model only)."""
        _lib.check(self.lib.lrg_engine_set_forward_mode(self._h, int(mode)))

    def range_overflow(self):
        """(overflow seen since the last check, synchronous calls repeated with 3xTF32 so far)."""
        over, fb = C.c_int(0), C.c_int(0)
        _lib.check(self.lib.lrg_engine_range_overflow(self._h, C.byref(over), C.byref(fb)))
        return bool(over.value), fb.value

    def forward_mode(self):
        return self.lib.lrg_engine_forward_mode(self._h)

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.lrg_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ------------------------------------------------------------------------------------
    def load_weights(self, tensors):
        blob = pack_weights(tensors, self.F, self.lite)
        n = self.lib.lrg_engine_weight_count(self._h)
        if blob.size != n:
            raise ValueError('weight blob has %d floats, engine expects %d' % (blob.size, n))
        _lib.check(self.lib.lrg_engine_load_weights(self._h, _lib.ptr(blob), blob.size))
        self._weights = {k: np.asarray(tensors[k], np.float32) for k, _ in variable_shapes(self.F, self.lite)}

    def load_checkpoint(self, prefix):
        from . import ckpt
        self.load_weights(ckpt.load_checkpoint(prefix))

    def variables(self):
        return dict(self._weights or {})

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, inlier, neighbor):
        """inlier (B,Ni,F), neighbor (B,Nj,F) float32 host arrays -> (add_output (B,Nj,2), remove_output (B,Ni,2))."""
        inlier = np.ascontiguousarray(inlier, dtype=np.float32)
        neighbor = np.ascontiguousarray(neighbor, dtype=np.float32)
        if inlier.ndim != 3 or inlier.shape[1:] != (self.Ni, self.F):
            raise ValueError('inlier_pl expects shape (B, %d, %d), got %s' % (self.Ni, self.F, inlier.shape))
        if neighbor.ndim != 3 or neighbor.shape[1:] != (self.Nj, self.F) or neighbor.shape[0] != inlier.shape[0]:
            raise ValueError('neighbor_pl expects shape (%d, %d, %d), got %s' % (inlier.shape[0], self.Nj, self.F, neighbor.shape))
        B = inlier.shape[0]
        add = np.empty((B, self.Nj, 2), np.float32)
        rmv = np.empty((B, self.Ni, 2), np.float32)
        _lib.check(self.lib.lrg_forward_host(self._h, B, _lib.ptr(inlier), _lib.ptr(neighbor), _lib.ptr(add), _lib.ptr(rmv)))
        return add, rmv

    def forward_device(self, d_inlier, d_neighbor, d_add, d_remove, B, stream=None):
        """Asynchronous forward on device pointers (ints or torch CUDA tensors)."""
        _lib.check(self.lib.lrg_forward_device(self._h, B, _lib.ptr(d_inlier), _lib.ptr(d_neighbor), _lib.ptr(d_add),
                                               _lib.ptr(d_remove), C.c_void_p(stream) if stream else None))

    # -- region growing -----------------------------------------------------------------------------
    @staticmethod
    def _concat(rooms_points, rooms_order):
        counts = [len(p) for p in rooms_points]
        offsets = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        if len(counts):
            points = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32) for p in rooms_points]), dtype=np.float32)
            order = np.ascontiguousarray(np.concatenate([np.asarray(o).astype(np.int32) for o in rooms_order]), dtype=np.int32)
        else:
            points = np.zeros((0, 13), np.float32)
            order = np.zeros(0, np.int32)
        return offsets, points, order

    def upload_rooms(self, rooms_points, rooms_order, resolution=0.1):
        """rooms_points: list of (N_r, F) float32 feature arrays (test_region_grow.py:165-172);
        rooms_order: list of (N_r,) seed orders = argsort(curvatures) (:183)."""
        offsets, points, order = self._concat(rooms_points, rooms_order)
        if points.shape[0] and points.shape[1] != self.F:
            raise ValueError('rooms have %d features, engine was built for %d' % (points.shape[1], self.F))
        self.upload_concatenated(offsets, points, order, resolution)
        return offsets

    def upload_concatenated(self, offsets, points, order, resolution=0.1):
        _lib.check(self.lib.lrg_rooms_upload(self._h, len(offsets) - 1, _lib.ptr(offsets), _lib.ptr(points),
                                             _lib.ptr(order), C.c_float(resolution)))
        self._room_offsets = np.array(offsets, dtype=np.int64)

    def upload_raw_rooms(self, raw_rooms, resolution=0.1):
        """raw_rooms: list of (N_r, C>=6) float32 arrays x y z r g b [...] (the rows of the reference's H5 files).  The
        features of test_region_grow.py:119-173 are prepared on the device.  Returns the equalised room offsets."""
        counts = [len(r) for r in raw_rooms]
        raw_off = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=raw_off[1:])
        ncols = raw_rooms[0].shape[1] if raw_rooms else 6
        raw = (np.ascontiguousarray(np.concatenate([np.asarray(r, np.float32) for r in raw_rooms]), dtype=np.float32)
               if raw_rooms else np.zeros((0, ncols), np.float32))
        return self.upload_raw_concatenated(raw_off, raw, resolution)

    def upload_raw_concatenated(self, raw_offsets, raw_points, resolution=0.1, n_cols=None):
        """raw_points: (sum Nr, C) float32 host array, or a device pointer / torch CUDA tensor of that layout (then ``n_cols``
        is taken from the tensor or must be given): the raw points are read where they are."""
        n_rooms = len(raw_offsets) - 1
        on_device = isinstance(raw_points, int) or (hasattr(raw_points, 'is_cuda') and raw_points.is_cuda)
        if n_cols is None:
            n_cols = int(raw_points.shape[1])
        fn = self.lib.lrg_rooms_upload_raw_device if on_device else self.lib.lrg_rooms_upload_raw
        _lib.check(fn(self._h, n_rooms, _lib.ptr(np.ascontiguousarray(raw_offsets, np.int64)), _lib.ptr(raw_points), int(n_cols),
                      C.c_float(resolution)))
        eq = np.zeros(n_rooms + 1, dtype=np.int64)
        _lib.check(self.lib.lrg_rooms_equalized_offsets(self._h, _lib.ptr(eq)))
        self._room_offsets = eq
        self._raw_offsets = np.array(raw_offsets, dtype=np.int64)
        return eq

    def prepared_features(self):
        """After upload_raw_rooms: dict(points (sum Neq, F), order, equalized_idx, unequalized_idx), concatenated over rooms."""
        te, tr = int(self._room_offsets[-1]), int(self._raw_offsets[-1])
        out = dict(points=np.zeros((te, self.F), np.float32), order=np.zeros(te, np.int32),
                   equalized_idx=np.zeros(te, np.int32), unequalized_idx=np.zeros(tr, np.int32))
        _lib.check(self.lib.lrg_rooms_features_download(self._h, _lib.ptr(out['points']), _lib.ptr(out['order']),
                                                        _lib.ptr(out['equalized_idx']), _lib.ptr(out['unequalized_idx'])))
        return out

    def prepare_ms(self):
        """Device time of the last raw upload (copy + feature preparation + packing), CUDA events on the engine stream."""
        ms = C.c_float(0)
        _lib.check(self.lib.lrg_last_prepare_ms(self._h, C.byref(ms)))
        return ms.value

    def prepare_launches(self):
        """Kernels the last upload launched (feature preparation incl. its sort passes, packing, spatial index)."""
        n = C.c_int(0)
        _lib.check(self.lib.lrg_last_prepare_launches(self._h, C.byref(n)))
        return n.value

    def raw_labels(self, filled=True):
        """cluster_label[unequalized_idx] per room (test_region_grow.py:366)."""
        tr = int(self._raw_offsets[-1])
        out = np.zeros(tr, dtype=np.int32)
        _lib.check(self.lib.lrg_labels_download_raw(self._h, _lib.ptr(out), 1 if filled else 0))
        return [out[self._raw_offsets[i]:self._raw_offsets[i + 1]] for i in range(len(self._raw_offsets) - 1)]

    def segment_raw_rooms(self, raw_rooms, resolution=0.1, **kw):
        """Raw rooms in, per-raw-point instance labels out: feature preparation, growing and fill all on the device."""
        self.upload_raw_rooms(raw_rooms, resolution)
        stats = self.segment_resident(resolution=resolution, **kw)
        return self.raw_labels(True), stats

    def make_params(self, resolution=0.1, cluster_threshold=10, seed=0, max_slots=0, max_steps_per_region=0,
                    room_id_base=0, trace_capacity=0, flags=0, num_restarts=0, beam_width=0, search_width=0, spec_lanes=0,
                    spec_top=0, spec_min_idle=0, scoring='np', spec_crit=0):
        """``num_restarts`` > 1 selects the random-restart driver (test_random_restart.py, NUM_RESTARTS, 'np' scoring);
        ``beam_width`` / ``search_width`` > 0 the beam-search driver (test_beam_search.py, BEAM_WIDTH, SEARCH_WIDTH) with
        ``scoring`` 'np' (region size, the default, :41) or 'ml' (accumulated log-probability, ``--scoring ml``, :46-47,263-264);
        ``spec_lanes`` > 1 grows that many regions of a room side by side with in-order commits (same labels as 1), in the
        ``spec_top`` rooms with the most work left while they hold at least 1 / ``spec_crit`` of the run's remaining work, and
        wherever ``spec_min_idle`` CTAs idle (0 = defaults, < 0 = all / never)."""
        if scoring not in ('np', 'ml'):
            raise ValueError("scoring must be 'np' or 'ml'")
        if scoring == 'ml':
            flags |= _lib.FLAG_SCORE_ML
        p = GrowParams(resolution, cluster_threshold, seed, max_slots, max_steps_per_region, room_id_base,
                       trace_capacity, flags, num_restarts, beam_width, search_width, spec_lanes, spec_top, spec_min_idle, spec_crit)
        return p

    def segment_resident(self, params=None, **kw):
        """Grow all uploaded rooms on the device; returns the per-room statistics (structured array)."""
        if self._room_offsets is None:
            raise LrgError('segment_resident: no rooms uploaded')
        params = params or self.make_params(**kw)
        stats = np.zeros(len(self._room_offsets) - 1, dtype=ROOM_STATS_DTYPE)
        _lib.check(self.lib.lrg_segment_resident(self._h, C.byref(params), _lib.ptr(stats)))
        return stats

    def labels(self, filled=True):
        total = int(self._room_offsets[-1])
        out = np.zeros(total, dtype=np.int32)
        _lib.check(self.lib.lrg_labels_download(self._h, _lib.ptr(out), 1 if filled else 0))
        return [out[self._room_offsets[i]:self._room_offsets[i + 1]] for i in range(len(self._room_offsets) - 1)]

    def room_metrics(self, obj_ids, raw=False, filled=True, return_label2=False):
        """NMI / AMI / ARS / PRC / RCL / IOU of every room (test_region_grow.py:319-349) against ground-truth object ids:
        ``obj_ids`` = list of per-room int arrays, one id per equalised point, or per raw point with ``raw=True`` (rooms
        uploaded raw; gathered with equalized_idx on the device, :136).  Returns a structured array (one row per room) and,
        with ``return_label2``, the per-room cluster_label2 arrays (:323,335,339-341)."""
        obj = np.ascontiguousarray(np.concatenate([np.asarray(o).astype(np.int32) for o in obj_ids]) if len(obj_ids) else np.zeros(0, np.int32))
        expect = int((self._raw_offsets if raw else self._room_offsets)[-1])
        if obj.size != expect:
            raise ValueError('obj_ids hold %d values, the uploaded rooms have %d %s points' % (obj.size, expect, 'raw' if raw else 'equalised'))
        n_rooms = len(self._room_offsets) - 1
        out = np.zeros(n_rooms, dtype=ROOM_METRICS_DTYPE)
        label2 = np.zeros(int(self._room_offsets[-1]), np.int32) if return_label2 else None
        _lib.check(self.lib.lrg_room_metrics(self._h, _lib.ptr(obj), 1 if raw else 0, 1 if filled else 0, _lib.ptr(out), _lib.ptr(label2)))
        if return_label2:
            return out, [label2[self._room_offsets[i]:self._room_offsets[i + 1]] for i in range(n_rooms)]
        return out

    def trace(self, room, capacity, lane=None):
        buf = np.zeros(capacity, dtype=STEP_TRACE_DTYPE)
        n = C.c_int(0)
        if lane is None:
            _lib.check(self.lib.lrg_trace_download(self._h, room, _lib.ptr(buf), capacity, C.byref(n)))
        else:
            _lib.check(self.lib.lrg_trace_download_lane(self._h, room, lane, _lib.ptr(buf), capacity, C.byref(n)))
        return buf[:min(n.value, capacity)], n.value

    def profile(self):
        g, f, fw = C.c_float(0), C.c_float(0), C.c_float(0)
        it, ln = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.lrg_last_segment_profile(self._h, C.byref(g), C.byref(f), C.byref(it), C.byref(ln), C.byref(fw)))
        k = (C.c_float * 4)()
        _lib.check(self.lib.lrg_last_kernel_times(self._h, C.byref(k)))
        pers = C.c_int(0)
        busy = (C.c_double * 4)()
        items = (C.c_int64 * 4)()
        _lib.check(self.lib.lrg_last_grow_profile(self._h, C.byref(pers), C.byref(busy), C.byref(items)))
        delay = (C.c_double * 4)()
        _lib.check(self.lib.lrg_last_grow_queue_delay(self._h, C.byref(delay)))
        return dict(grow_ms=g.value, fill_ms=f.value, iterations=it.value, kernel_launches=ln.value, forward_ms=fw.value,
                    step_kernel_ms=k[0], branch_kernel_ms=k[1], gproj_kernel_ms=k[2], head_kernel_ms=k[3],
                    persistent=bool(pers.value), queue_delay_ms=dict(zip(('step', 'branch', 'gproj', 'head'), delay)), busy_ms=dict(zip(('step', 'branch', 'gproj', 'head'), busy)),
                    items=dict(zip(('step', 'branch', 'gproj', 'head'), items)))

    def labels_device_ptr(self, filled=True):
        p = C.c_void_p()
        _lib.check(self.lib.lrg_labels_device_ptr(self._h, 1 if filled else 0, C.byref(p)))
        return p.value

    def segment_rooms(self, rooms_points, rooms_order, resolution=0.1, **kw):
        """End-to-end call on host arrays: upload, grow, fill, download.  Returns (labels per room, stats)."""
        offsets, points, order = self._concat(rooms_points, rooms_order)
        return self.segment_concatenated(offsets, points, order, resolution=resolution, **kw)

    def segment_concatenated(self, offsets, points, order, resolution=0.1, **kw):
        params = self.make_params(resolution=resolution, **kw)
        n_rooms = len(offsets) - 1
        labels = np.zeros(int(offsets[-1]), dtype=np.int32)
        stats = np.zeros(n_rooms, dtype=ROOM_STATS_DTYPE)
        _lib.check(self.lib.lrg_segment_rooms_host(self._h, n_rooms, _lib.ptr(offsets), _lib.ptr(points), _lib.ptr(order),
                                                   C.byref(params), _lib.ptr(labels), _lib.ptr(stats)))
        self._room_offsets = np.array(offsets, dtype=np.int64)
        return [labels[offsets[i]:offsets[i + 1]] for i in range(n_rooms)], stats
