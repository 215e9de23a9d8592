"""The PointNet++ tf_ops with the reference's Python signatures (/root/reference/tf_ops/*/tf_*.py), computed by the
sm_100a kernels of liblrg_b200.so.  Arguments may be numpy arrays (copied to the device and back) or torch CUDA tensors
(used in place, results returned as torch tensors on the same device).  Shape/attribute violations raise ValueError with
the wording of the reference's OP_REQUIRES checks."""
import ctypes as C

import numpy as np

from . import _lib


class _DeviceBuffer:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        _lib.check(_lib.lib().lrg_malloc(C.byref(self.ptr), max(int(nbytes), 1)))

    def free(self):
        if self.ptr:
            _lib.lib().lrg_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        # an op that raises between inp()/out() and finish() drops its _Call: the buffers go with it
        try:
            self.free()
        except Exception:
            pass


class _Call:
    """Marshals numpy / torch arguments to device pointers for one op call."""

    def __init__(self, *inputs):
        _lib.require_gpu()
        self.torch = any(hasattr(a, 'data_ptr') for a in inputs)
        self.bufs = []
        self.outs = []
        if self.torch:
            import torch
            self.t = torch
            self.dev = next(a.device for a in inputs if hasattr(a, 'data_ptr'))
            if self.dev.type != 'cuda':
                raise ValueError('torch arguments must live on a CUDA device')
            _lib.check(_lib.lib().lrg_set_device(self.dev.index or 0))

    def inp(self, a, dtype):
        if self.torch:
            tdt = self.t.float32 if dtype == np.float32 else self.t.int32
            t = self.t.as_tensor(a, device=self.dev).to(tdt).contiguous()
            self.bufs.append(t)
            return C.c_void_p(t.data_ptr())
        a = np.ascontiguousarray(a, dtype=dtype)
        buf = _DeviceBuffer(a.nbytes)
        self.bufs.append(buf)
        if a.nbytes:
            _lib.check(_lib.lib().lrg_memcpy_h2d(buf.ptr, _lib.ptr(a), a.nbytes))
        return buf.ptr

    def out(self, shape, dtype, zero=False, init=None):
        if self.torch:
            tdt = self.t.float32 if dtype == np.float32 else self.t.int32
            t = self.t.zeros(shape, dtype=tdt, device=self.dev) if (zero or init is None) else self.t.as_tensor(init, device=self.dev).to(tdt).contiguous().clone()
            self.outs.append(t)
            return C.c_void_p(t.data_ptr())
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        buf = _DeviceBuffer(n)
        self.bufs.append(buf)
        if init is not None:
            a = np.ascontiguousarray(init, dtype=dtype)
            _lib.check(_lib.lib().lrg_memcpy_h2d(buf.ptr, _lib.ptr(a), a.nbytes))
        elif n:
            _lib.check(_lib.lib().lrg_memset(buf.ptr, 0, n))
        self.outs.append((buf, shape, dtype))
        return buf.ptr

    def stream(self):
        if self.torch:
            return C.c_void_p(self.t.cuda.current_stream(self.dev).cuda_stream)
        return None

    def finish(self):
        res = []
        if self.torch:
            res = list(self.outs)
        else:
            _lib.check(_lib.lib().lrg_device_synchronize())
            for buf, shape, dtype in self.outs:
                a = np.empty(shape, dtype=dtype)
                if a.nbytes:
                    _lib.check(_lib.lib().lrg_memcpy_d2h(_lib.ptr(a), buf.ptr, a.nbytes))
                res.append(a)
            for b in self.bufs:
                b.free()
        return res[0] if len(res) == 1 else tuple(res)


def _shape(a):
    return tuple(a.shape)


def _require(cond, msg):
    if not cond:
        raise ValueError(msg)


# ----------------------------------------------------------------------------- tf_sampling.py
def farthest_point_sample(npoint, inp):
    """tf_sampling.py:48-56: inp (B,N,3) float32 -> (B,npoint) int32, first index always 0."""
    s = _shape(inp)
    _require(len(s) == 3 and s[2] == 3, 'FarthestPointSample expects (batch_size,num_points,3) inp shape')   # tf_sampling.cpp:105
    b, n, _ = s
    c = _Call(inp)
    d_inp = c.inp(inp, np.float32)
    d_tmp = c.out((32, n), np.float32) if n > 65536 else None     # (up to 65,536 points live in registers: one CTA or a cluster)
    d_out = c.out((b, npoint), np.int32)
    _lib.check(_lib.lib().lrg_farthest_point_sampling(b, n, npoint, d_inp, d_tmp, d_out, c.stream()))
    res = c.finish()
    return res[-1] if isinstance(res, tuple) else res


def sample_and_group(npoint, radius, nsample, xyz, points=None):
    """train_pointnet.py:113-123: ``new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))``, ball query around new_xyz,
    ``grouped_xyz = group_point(xyz, idx) - new_xyz``, ``new_points = concat(grouped_xyz, group_point(points, idx))`` -- two
    launches (lrg_sample_and_group).  Returns (new_xyz, new_points, idx, grouped_xyz) like the reference."""
    s = _shape(xyz)
    _require(len(s) == 3 and s[2] == 3, 'FarthestPointSample expects (batch_size,num_points,3) inp shape')   # tf_sampling.cpp:105
    b, n, _ = s
    ch = 0
    if points is not None:
        sp = _shape(points)
        _require(len(sp) == 3 and sp[0] == b and sp[1] == n, 'GroupPoint expects (batch_size, num_points, channel) points shape')   # tf_grouping.cpp:148
        ch = sp[2]
    c = _Call(*([xyz] + ([points] if points is not None else [])))
    d_xyz = c.inp(xyz, np.float32)
    d_pts = c.inp(points, np.float32) if ch else None
    d_tmp = c.out((32, n), np.float32) if n > 65536 else None
    d_fps = c.out((b, npoint), np.int32)
    d_new_xyz = c.out((b, npoint, 3), np.float32)
    d_new_points = c.out((b, npoint, nsample, 3 + ch), np.float32)
    d_idx = c.out((b, npoint, nsample), np.int32)
    d_cnt = c.out((b, npoint), np.int32)
    d_gxyz = c.out((b, npoint, nsample, 3), np.float32)
    _lib.check(_lib.lib().lrg_sample_and_group(b, n, npoint, C.c_float(radius), nsample, ch, d_xyz, d_pts, d_tmp, d_fps, d_new_xyz, d_new_points,
                                               d_idx, d_cnt, d_gxyz, c.stream()))
    res = c.finish()
    new_xyz, new_points, idx, _, grouped_xyz = res[-5:]
    return new_xyz, new_points, idx, grouped_xyz


def gather_point(inp, idx):
    """tf_sampling.py:29-37: inp (B,N,3), idx (B,M) -> (B,M,3)."""
    s, si = _shape(inp), _shape(idx)
    _require(len(s) == 3 and s[2] == 3, 'GatherPoint expects (batch_size,num_points,3) inp shape')          # tf_sampling.cpp:131
    _require(len(si) == 2 and si[0] == s[0], 'GatherPoint expects (batch_size,num_result) idx shape')         # :135
    c = _Call(inp, idx)
    d_inp, d_idx = c.inp(inp, np.float32), c.inp(idx, np.int32)
    d_out = c.out((s[0], si[1], 3), np.float32)
    _lib.check(_lib.lib().lrg_gather_point(s[0], s[1], si[1], d_inp, d_idx, d_out, c.stream()))
    return c.finish()


def gather_point_grad(inp, idx, out_g):
    """GatherPointGrad (tf_sampling.py:38-43): scatter-add of out_g (B,M,3) into zeros like inp."""
    s, si = _shape(inp), _shape(idx)
    _require(len(s) == 3 and s[2] == 3, 'GatherPointGradGpuOp expects (batch_size,num_points,3) inp')
    _require(_shape(out_g) == (s[0], si[1], 3), 'GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape')
    c = _Call(inp, idx, out_g)
    d_idx, d_g = c.inp(idx, np.int32), c.inp(out_g, np.float32)
    d_out = c.out(s, np.float32, zero=True)
    _lib.check(_lib.lib().lrg_scatter_add_point(s[0], s[1], si[1], d_g, d_idx, d_out, c.stream()))
    return c.finish()


def prob_sample(inp, inpr):
    """tf_sampling.py:13-21: inp (B, ncategory) float32 probabilities (any positive scale), inpr (B, npoints) float32 uniforms in
    [0,1) -> (B, npoints) int32: for every uniform the first category whose float32 cumulative sum reaches uniform * total
    (cumulative sums in the reference's association, so the indices agree with its kernel)."""
    _require(len(_shape(inp)) == 2, 'ProbSample expects (batch_size,num_choices) inp shape')                  # tf_sampling.cpp:76
    b, n = _shape(inp)
    _require(len(_shape(inpr)) == 2 and _shape(inpr)[0] == b, 'ProbSample expects (batch_size,num_points) inpr shape')   # :79
    _require(n > 0, 'ProbSample expects at least one choice')
    m = _shape(inpr)[1]
    c = _Call(inp, inpr)
    d_p, d_r = c.inp(inp, np.float32), c.inp(inpr, np.float32)
    d_tmp = c.out((b, n), np.float32)
    d_out = c.out((b, m), np.int32)
    _lib.check(_lib.lib().lrg_prob_sample(b, n, m, d_p, d_r, d_tmp, d_out, c.stream()))
    return c.finish()[1]


# ----------------------------------------------------------------------------- tf_grouping.py
def query_ball_point(radius, nsample, xyz1, xyz2):
    """tf_grouping.py:8-20: xyz1 (B,N,3) points, xyz2 (B,M,3) queries -> idx (B,M,nsample) int32, pts_cnt (B,M) int32.
    Rows with an empty ball keep idx = 0 here (the reference leaves them uninitialised)."""
    _require(radius > 0, 'QueryBallPoint expects positive radius')                                           # tf_grouping.cpp:71
    _require(nsample > 0, 'QueryBallPoint expects positive nsample')                                         # :74
    s1, s2 = _shape(xyz1), _shape(xyz2)
    _require(len(s1) == 3 and s1[2] == 3, 'QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.')    # :79
    _require(len(s2) == 3 and s2[2] == 3 and s2[0] == s1[0], 'QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.')
    c = _Call(xyz1, xyz2)
    d1, d2 = c.inp(xyz1, np.float32), c.inp(xyz2, np.float32)
    d_idx = c.out((s1[0], s2[1], nsample), np.int32, zero=True)
    d_cnt = c.out((s1[0], s2[1]), np.int32, zero=True)
    _lib.check(_lib.lib().lrg_query_ball_point(s1[0], s1[1], s2[1], float(radius), int(nsample), d1, d2, d_idx, d_cnt, c.stream()))
    return c.finish()


def select_top_k(k, dist):
    """tf_grouping.py:22-31: dist (B,M,N) -> (idx (B,M,N) int32, dist_out (B,M,N)); only the first k columns are sorted."""
    _require(k > 0, 'SelectionSort expects positive k')                                                      # tf_grouping.cpp:113
    s = _shape(dist)
    _require(len(s) == 3, 'SelectionSort expects (b,m,n) dist shape.')                                       # :118
    c = _Call(dist)
    d = c.inp(dist, np.float32)
    d_i = c.out(s, np.int32)
    d_o = c.out(s, np.float32)
    _lib.check(_lib.lib().lrg_selection_sort(s[0], s[2], s[1], int(k), d, d_i, d_o, c.stream()))
    return c.finish()


def group_point(points, idx):
    """tf_grouping.py:33-41: points (B,N,C), idx (B,M,nsample) -> (B,M,nsample,C)."""
    s, si = _shape(points), _shape(idx)
    _require(len(s) == 3, 'GroupPoint expects (batch_size, num_points, channel) points shape')               # tf_grouping.cpp:149
    _require(len(si) == 3 and si[0] == s[0], 'GroupPoint expects (batch_size, npoints, nsample) idx shape')  # :155
    c = _Call(points, idx)
    d_p, d_i = c.inp(points, np.float32), c.inp(idx, np.int32)
    d_o = c.out((s[0], si[1], si[2], s[2]), np.float32)
    _lib.check(_lib.lib().lrg_group_point(s[0], s[1], s[2], si[1], si[2], d_p, d_i, d_o, c.stream()))
    return c.finish()


def group_point_grad(points, idx, grad_out):
    """GroupPointGrad (tf_grouping.py:42-46)."""
    s, si = _shape(points), _shape(idx)
    _require(len(s) == 3 and len(si) == 3 and si[0] == s[0], 'GroupPointGrad expects (b,n,c) points and (b,m,nsample) idx')
    _require(_shape(grad_out) == (s[0], si[1], si[2], s[2]), 'GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape')
    c = _Call(points, idx, grad_out)
    d_i, d_g = c.inp(idx, np.int32), c.inp(grad_out, np.float32)
    d_o = c.out(s, np.float32, zero=True)
    _lib.check(_lib.lib().lrg_group_point_grad(s[0], s[1], s[2], si[1], si[2], d_g, d_i, d_o, c.stream()))
    return c.finish()


def knn_point(k, xyz1, xyz2):
    """tf_grouping.py:48-73: brute-force squared distances (B,M,N) then select_top_k; returns (val, idx) (B,M,k)."""
    s1, s2 = _shape(xyz1), _shape(xyz2)
    _require(len(s1) == 3 and len(s2) == 3 and s1[0] == s2[0] and s1[2] == s2[2], 'knn_point expects (b,n,c) xyz1 and (b,m,c) xyz2')
    b, n, ch = s1
    m = s2[1]
    c = _Call(xyz1, xyz2)
    d1, d2 = c.inp(xyz1, np.float32), c.inp(xyz2, np.float32)
    d_dist = c.out((b, m, n), np.float32)
    d_outi = c.out((b, m, n), np.int32)
    d_out = c.out((b, m, n), np.float32)
    _lib.check(_lib.lib().lrg_pairwise_sqdist(b, n, m, ch, d1, d2, d_dist, c.stream()))              # :66-68
    _lib.check(_lib.lib().lrg_selection_sort(b, n, m, k, d_dist, d_outi, d_out, c.stream()))         # :70
    _, outi, out = c.finish()
    return out[:, :, :k], outi[:, :, :k]                                                             # :71-72


# ----------------------------------------------------------------------------- tf_interpolate.py
def three_nn(xyz1, xyz2):
    """tf_interpolate.py:8-17: xyz1 (B,N,3) unknown, xyz2 (B,M,3) known -> dist (B,N,3) squared distances, idx (B,N,3)."""
    s1, s2 = _shape(xyz1), _shape(xyz2)
    _require(len(s1) == 3 and s1[2] == 3, 'ThreeNN expects (b,n,3) xyz1 shape.')                             # tf_interpolate.cpp:163
    _require(len(s2) == 3 and s2[2] == 3 and s2[0] == s1[0], 'ThreeNN expects (b,m,3) xyz2 shape.')          # :168
    c = _Call(xyz1, xyz2)
    d1, d2 = c.inp(xyz1, np.float32), c.inp(xyz2, np.float32)
    d_d = c.out(s1, np.float32)
    d_i = c.out(s1, np.int32)
    _lib.check(_lib.lib().lrg_three_nn(s1[0], s1[1], s2[1], d1, d2, d_d, d_i, c.stream()))
    return c.finish()


def three_interpolate(points, idx, weight):
    """tf_interpolate.py:19-28: points (B,M,C), idx (B,N,3), weight (B,N,3) -> (B,N,C)."""
    s, si = _shape(points), _shape(idx)
    _require(len(s) == 3, 'ThreeInterpolate expects (b,m,c) points shape')                                   # tf_interpolate.cpp:197
    _require(len(si) == 3 and si[0] == s[0] and si[2] == 3, 'ThreeInterpolate expects (b,n,3) idx shape')    # :201
    _require(_shape(weight) == si, 'ThreeInterpolate expects (b,n,3) weight shape')                          # :204
    c = _Call(points, idx, weight)
    d_p, d_i, d_w = c.inp(points, np.float32), c.inp(idx, np.int32), c.inp(weight, np.float32)
    d_o = c.out((s[0], si[1], s[2]), np.float32)
    _lib.check(_lib.lib().lrg_three_interpolate(s[0], s[1], s[2], si[1], d_p, d_i, d_w, d_o, c.stream()))
    return c.finish()


def three_interpolate_grad(points, idx, weight, grad_out):
    """ThreeInterpolateGrad (tf_interpolate.py:29-34): returns grad_points (B,M,C)."""
    s, si = _shape(points), _shape(idx)
    _require(len(s) == 3 and len(si) == 3 and si[2] == 3, 'ThreeInterpolateGrad expects (b,m,c) points and (b,n,3) idx')
    _require(_shape(grad_out) == (s[0], si[1], s[2]), 'ThreeInterpolateGrad expects (b,n,c) grad_out shape')
    c = _Call(points, idx, weight, grad_out)
    d_i, d_w, d_g = c.inp(idx, np.int32), c.inp(weight, np.float32), c.inp(grad_out, np.float32)
    d_o = c.out(s, np.float32, zero=True)
    _lib.check(_lib.lib().lrg_three_interpolate_grad(s[0], si[1], s[2], s[1], d_g, d_i, d_w, d_o, c.stream()))
    return c.finish()
