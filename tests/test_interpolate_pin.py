"""three_nn / three_interpolate pinned to the reference's OWN code: tf_ops/3d_interpolation/tf_interpolate.cpp compiled
unmodified against stand-in TensorFlow headers (oracle/build_ref.py -> oracle/_ref/libref_interpolate.so, CPU code).

* CPU: the C restatement oracle/tfops_oracle.c equals the reference's loops bit for bit (indices AND float values) on the
  PointNet++ call shapes (train_pointnet.py:187-190), with exact ties and with m < 3; the reference's OpKernels (Compute
  with its OP_REQUIRES shape checks, :163-168,197-206) give the same outputs and the InvalidArgument wording our host
  wrappers repeat.
* GPU: the CUDA kernels through the C ABI equal the reference's loops bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import build_ref, tfops as O

if os.path.isdir(build_ref.REF):
    build_ref.build_ref()
pytestmark = pytest.mark.skipif(not O.ReferenceInterpolate.available(), reason='oracle/_ref/libref_interpolate.so not built')

SHAPES = [(1, 64, 16, 512), (2, 256, 64, 256), (1, 1024, 256, 256), (1, 1024, 1024, 128), (2, 50, 2, 8), (1, 10, 1, 4), (3, 33, 3, 5)]


def _case(b, n, m, c):
    rng = np.random.RandomState(100 * n + m)
    x1, x2 = rng.rand(b, n, 3).astype(np.float32), rng.rand(b, m, 3).astype(np.float32)
    if m >= 16:
        x2[:, 5] = x2[:, 3]                              # exact ties: the earliest index must win (strict <, :75-92)
        x1[:, 0] = x2[:, 7]                              # a zero distance
    pts = rng.randn(b, m, c).astype(np.float32)
    go = rng.randn(b, n, c).astype(np.float32)
    return x1, x2, pts, go


def _weights(dist):
    d = np.maximum(np.where(np.isfinite(dist), dist, 1e10), 1e-10)          # train_pointnet.py:146-149
    return ((1.0 / d) / np.sum(1.0 / d, axis=2, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize('b,n,m,c', SHAPES)
def test_c_restatement_equals_the_reference_loops(b, n, m, c):
    R = O.ReferenceInterpolate()
    x1, x2, pts, go = _case(b, n, m, c)
    rd, ri = R.three_nn(x1, x2)
    od, oi = O.three_nn(x1, x2)
    np.testing.assert_array_equal(oi, ri)
    np.testing.assert_array_equal(od, rd)                # bit for bit, inf where m < 3 ((float)1e40, :66-68)
    w = _weights(rd)
    np.testing.assert_array_equal(O.three_interpolate(pts, ri, w), R.three_interpolate(pts, ri, w))
    np.testing.assert_array_equal(O.three_interpolate_grad(pts, ri, w, go), R.three_interpolate_grad(pts, ri, w, go))


def test_reference_opkernels_compute_and_shape_checks():
    R = O.ReferenceInterpolate()
    x1, x2, pts, go = _case(2, 256, 64, 32)
    rd, ri = R.three_nn(x1, x2)
    kd, ki = np.zeros_like(rd), np.zeros_like(ri)
    R.run_kernel('ThreeNN', [x1, x2], [kd, ki])
    assert np.array_equal(kd, rd) and np.array_equal(ki, ri)
    w = _weights(rd)
    ko = np.zeros((2, 256, 32), np.float32)
    R.run_kernel('ThreeInterpolate', [pts, ri, w], [ko])
    assert np.array_equal(ko, R.three_interpolate(pts, ri, w))
    kg = np.zeros_like(pts)
    R.run_kernel('ThreeInterpolateGrad', [pts, ri, w, go], [kg])
    assert np.array_equal(kg, R.three_interpolate_grad(pts, ri, w, go))
    with pytest.raises(ValueError, match=r'ThreeNN expects \(b,n,3\) xyz1 shape'):
        R.run_kernel('ThreeNN', [x1[:, :, :2], x2], [kd, ki])
    with pytest.raises(ValueError, match=r'ThreeNN expects \(b,m,3\) xyz2 shape'):
        R.run_kernel('ThreeNN', [x1, x2[0]], [kd, ki])
    with pytest.raises(ValueError, match=r'ThreeInterpolate expects \(b,m,c\) points shape'):
        R.run_kernel('ThreeInterpolate', [pts[0], ri, w], [ko])


@pytest.mark.gpu
@pytest.mark.parametrize('b,n,m,c', SHAPES)
def test_cuda_kernels_equal_the_reference_loops(b, n, m, c):
    from learn_region_grow_b200 import tfops as T
    R = O.ReferenceInterpolate()
    x1, x2, pts, go = _case(b, n, m, c)
    rd, ri = R.three_nn(x1, x2)
    dist, idx = T.three_nn(x1, x2)
    np.testing.assert_array_equal(idx, ri)
    np.testing.assert_array_equal(dist, rd)
    w = _weights(rd)
    np.testing.assert_array_equal(T.three_interpolate(pts, ri, w), R.three_interpolate(pts, ri, w))
    # the gradient is a scatter-add (atomics on the device): equal up to the order of the float additions
    np.testing.assert_allclose(T.three_interpolate_grad(pts, ri, w, go), R.three_interpolate_grad(pts, ri, w, go), rtol=1e-4, atol=1e-4)
