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


def prob_cumsum(inp):
    """The float32 cumulative sums of /root/reference/tf_ops/sampling/tf_sampling_g.cu:7-89 (cumsumKernel), restated per
    row in numpy float32 scalars with the kernel's association: chunks of 8192 (BlockSize*4, :8,14); inside a chunk groups of
    four with the prefixes v1, v1+v2, (v1+v2)+v3, (v3+v4)+(v1+v2) (:20-33; a short last group: running sums :35-43); an
    in-place up-sweep / down-sweep over the group totals (:46-67); element = (group prefix + totals before the group) +
    running sum (:69-81); compensated running sum across chunks (:82-85)."""
    inp = _f(inp)
    f = np.float32
    b, n = inp.shape
    out = np.zeros((b, n), np.float32)
    for i in range(b):
        running, running2 = f(0), f(0)
        for j in range(0, n, 8192):
            ln = min(n - j, 8192)
            n2 = (ln + 3) // 4
            buf4 = np.zeros(n2 * 4, np.float32)
            tot = np.zeros(n2, np.float32)
            for g in range(n2):
                k = 4 * g
                if k + 3 < ln:
                    v1, v2, v3, v4 = (f(x) for x in inp[i, j + k:j + k + 4])
                    v2 = f(v2 + v1)
                    v4 = f(v4 + v3)
                    v3 = f(v3 + v2)
                    v4 = f(v4 + v2)
                    buf4[k:k + 4] = (v1, v2, v3, v4)
                    tot[g] = v4
                else:
                    v = f(0)
                    for k2 in range(k, ln):
                        v = f(v + inp[i, j + k2])
                        buf4[k2] = v
                    buf4[ln:n2 * 4] = v
                    tot[g] = v
            u = 0
            while (2 << u) <= n2:                                   # up-sweep (:46-55)
                for k in range(n2 >> (u + 1)):
                    i1, i2 = (((k << 1) + 2) << u) - 1, (((k << 1) + 1) << u) - 1
                    tot[i1] = f(tot[i1] + tot[i2])
                u += 1
            u -= 1
            while u >= 0:                                           # down-sweep (:56-66)
                for k in range((n2 - (1 << u)) >> (u + 1)):
                    i1, i2 = (((k << 1) + 3) << u) - 1, (((k << 1) + 2) << u) - 1
                    tot[i1] = f(tot[i1] + tot[i2])
                u -= 1
            for g in range(1, n2):                                  # :68-76
                buf4[4 * g:4 * g + 4] = (buf4[4 * g:4 * g + 4] + tot[g - 1]).astype(np.float32)
            out[i, j:j + ln] = (buf4[:ln] + running).astype(np.float32)      # :78-80
            t = f(tot[n2 - 1] + running2)                           # :82-85
            r2 = f(running + t)
            running2 = f(t - f(r2 - running))
            running = r2
    return out


def prob_sample(inp, inpr):
    """tf_sampling.py:13-21 -> probsampleLauncher (tf_sampling_g.cu:198-201): cumsumKernel, then binarysearchKernel (:91-104)."""
    inp, inpr = _f(inp), _f(inpr)
    b, n = inp.shape
    m = inpr.shape[1]
    cum = prob_cumsum(inp)
    base = 1
    while base < n:
        base <<= 1
    out = np.zeros((b, m), np.int32)
    for i in range(b):
        for j in range(m):
            q = np.float32(inpr[i, j] * cum[i, n - 1])
            r = n - 1
            k = base
            while k >= 1:
                if r >= k and cum[i, r - k] >= q:
                    r -= k
                k >>= 1
            out[i, j] = r
    return out


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
        self.prob_sample = self.sampling._Z18probsampleLauncheriiiPKfS0_PfPi
        for fn in (self.fps, self.gather, self.query_ball, self.group, self.selection_sort, self.prob_sample):
            fn.restype = None
        self.fps.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3
        self.gather.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3
        self.query_ball.argtypes = [C.c_int] * 3 + [C.c_float, C.c_int] + [C.c_void_p] * 4
        self.group.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
        self.selection_sort.argtypes = [C.c_int] * 4 + [C.c_void_p] * 3
        self.prob_sample.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, '_ref', 'libref_sampling.so')) and \
            os.path.exists(os.path.join(HERE, '_ref', 'libref_grouping.so'))


class ReferenceInterpolate:
    """The reference's own three_nn / three_interpolate(+grad): /root/reference/tf_ops/3d_interpolation/tf_interpolate.cpp
    compiled UNMODIFIED with g++ against the stand-in TensorFlow headers of oracle/tf_stubs (oracle/build_ref.py) into
    oracle/_ref/libref_interpolate.so.  CPU code -- runs in the build container and on the GPU box alike.  ``*_loop`` call the
    plain functions (:60,107,131); ``run_kernel`` drives the registered OpKernel's Compute (shape checks included)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(HERE, '_ref', 'libref_interpolate.so'))
        self.lib.ref_run_kernel.restype = C.c_int

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, '_ref', 'libref_interpolate.so'))

    def three_nn(self, xyz1, xyz2):
        xyz1, xyz2 = _f(xyz1), _f(xyz2)
        b, n, _ = xyz1.shape
        dist, idx = np.zeros((b, n, 3), np.float32), np.zeros((b, n, 3), np.int32)
        self.lib.ref_threenn_cpu(b, n, xyz2.shape[1], _p(xyz1), _p(xyz2), _p(dist), _p(idx))
        return dist, idx

    def three_interpolate(self, points, idx, weight):
        points, idx, weight = _f(points), _i(idx), _f(weight)
        b, m, c = points.shape
        n = idx.shape[1]
        out = np.zeros((b, n, c), np.float32)
        self.lib.ref_threeinterpolate_cpu(b, m, c, n, _p(points), _p(idx), _p(weight), _p(out))
        return out

    def three_interpolate_grad(self, points, idx, weight, grad_out):
        points, idx, weight, grad_out = _f(points), _i(idx), _f(weight), _f(grad_out)
        b, m, c = points.shape
        n = idx.shape[1]
        g = np.zeros_like(points)
        self.lib.ref_threeinterpolate_grad_cpu(b, n, c, m, _p(grad_out), _p(idx), _p(weight), _p(g))
        return g

    def run_kernel(self, name, inputs, outputs):
        """inputs: list of float32 / int32 arrays; outputs: list of preallocated arrays.  Raises ValueError with the
        OpKernel's InvalidArgument message (OP_REQUIRES) when a shape check fails."""
        inputs = [np.ascontiguousarray(a) for a in inputs]
        n_in = len(inputs)
        ptrs = (C.c_void_p * n_in)(*[a.ctypes.data for a in inputs])
        ranks = (C.c_int * n_in)(*[a.ndim for a in inputs])
        shapes = (C.c_longlong * (4 * n_in))()
        for i, a in enumerate(inputs):
            for d, s in enumerate(a.shape):
                shapes[4 * i + d] = s
        optrs = (C.c_void_p * len(outputs))(*[a.ctypes.data for a in outputs])
        err = C.create_string_buffer(512)
        rc = self.lib.ref_run_kernel(name.encode(), n_in, ptrs, ranks, shapes, len(outputs), optrs, err, 512)
        if rc != 0:
            raise ValueError(err.value.decode())
        return outputs
