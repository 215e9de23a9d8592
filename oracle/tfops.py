"""ctypes front-ends of the CPU tf_ops oracle (oracle/tfops_oracle.c) and of the reference's own kernels compiled
into oracle/_ref (GPU only) -- TEST INFRASTRUCTURE (oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_cpu = None


def cpu():
    global _cpu
    if _cpu is None:
        path = os.path.join(HERE, '_build', 'liboracle_tfops.so')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, 'tfops_oracle.c')):
            subprocess.run(['make', '-s', '-C', HERE], check=True)
        _cpu = C.CDLL(path)
    return _cpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def farthest_point_sample(npoint, inp):
    inp = _f(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    tmp = np.zeros(n, np.float32)
    cpu().oracle_farthest_point_sampling(b, n, npoint, _p(inp), _p(tmp), _p(out))
    return out


def gather_point(inp, idx):
    inp, idx = _f(inp), _i(idx)
    out = np.zeros((inp.shape[0], idx.shape[1], 3), np.float32)
    cpu().oracle_gather_point(inp.shape[0], inp.shape[1], idx.shape[1], _p(inp), _p(idx), _p(out))
    return out


def gather_point_grad(inp, idx, out_g):
    inp, idx, out_g = _f(inp), _i(idx), _f(out_g)
    g = np.zeros_like(inp)
    cpu().oracle_scatter_add_point(inp.shape[0], inp.shape[1], idx.shape[1], _p(out_g), _p(idx), _p(g))
    return g


def query_ball_point(radius, nsample, xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    cnt = np.zeros((b, m), np.int32)
    cpu().oracle_query_ball_point(b, n, m, C.c_float(radius), nsample, _p(xyz1), _p(xyz2), _p(idx), _p(cnt))
    return idx, cnt


def group_point(points, idx):
    points, idx = _f(points), _i(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.zeros((b, m, ns, c), np.float32)
    cpu().oracle_group_point(b, n, c, m, ns, _p(points), _p(idx), _p(out))
    return out


def group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _f(points), _i(idx), _f(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = np.zeros_like(points)
    cpu().oracle_group_point_grad(b, n, c, m, ns, _p(grad_out), _p(idx), _p(g))
    return g


def select_top_k(k, dist):
    dist = _f(dist)
    b, m, n = dist.shape
    outi = np.zeros((b, m, n), np.int32)
    out = np.zeros((b, m, n), np.float32)
    cpu().oracle_selection_sort(b, n, m, k, _p(dist), _p(outi), _p(out))
    return outi, out


def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    dist = np.zeros((b, n, 3), np.float32)
    idx = np.zeros((b, n, 3), np.int32)
    cpu().oracle_three_nn(b, n, xyz2.shape[1], _p(xyz1), _p(xyz2), _p(dist), _p(idx))
    return dist, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f(points), _i(idx), _f(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.zeros((b, n, c), np.float32)
    cpu().oracle_three_interpolate(b, m, c, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(points, idx, weight, grad_out):
    points, idx, weight, grad_out = _f(points), _i(idx), _f(weight), _f(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    g = np.zeros_like(points)
    cpu().oracle_three_interpolate_grad(b, n, c, m, _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


# ----------------------------------------------------------------------------- the reference's own kernels (GPU box)
class ReferenceKernels:
    """The unmodified reference launchers from oracle/_ref (device pointers in, legacy default stream)."""

    def __init__(self):
        self.sampling = C.CDLL(os.path.join(HERE, '_ref', 'libref_sampling.so'))
        self.grouping = C.CDLL(os.path.join(HERE, '_ref', 'libref_grouping.so'))
        self.fps = self.sampling._Z29farthestpointsamplingLauncheriiiPKfPfPi
        self.gather = self.sampling._Z19gatherpointLauncheriiiPKfPKiPf
        self.query_ball = self.grouping._Z22queryBallPointLauncheriiifiPKfS0_PiS1_
        self.group = self.grouping._Z18groupPointLauncheriiiiiPKfPKiPf
        self.selection_sort = self.grouping._Z21selectionSortLauncheriiiiPKfPiPf
        for fn in (self.fps, self.gather, self.query_ball, self.group, self.selection_sort):
            fn.restype = None
        self.fps.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3
        self.gather.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3
        self.query_ball.argtypes = [C.c_int] * 3 + [C.c_float, C.c_int] + [C.c_void_p] * 4
        self.group.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
        self.selection_sort.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, '_ref', 'libref_sampling.so')) and \
            os.path.exists(os.path.join(HERE, '_ref', 'libref_grouping.so'))
