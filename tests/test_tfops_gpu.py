"""tf_ops kernels (C ABI, device pointers) vs the C oracle and vs the reference's own kernels compiled into oracle/_ref."""
import ctypes as C

import numpy as np
import pytest

from oracle import tfops as O

pytestmark = pytest.mark.gpu


def _clouds(b, n, seed, dup=False):
    rng = np.random.RandomState(seed)
    x = rng.rand(b, n, 3).astype(np.float32)
    if dup:      # exact ties: duplicated points and a regular grid
        x[:, n // 2:] = x[:, :n - n // 2]
        x[0] = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(max(1, n // 64) + 1), indexing='ij'), -1).reshape(-1, 3)[:n].astype(np.float32) / 8
    return x


@pytest.mark.parametrize('b,n,m,dup', [(1, 1024, 1024, False), (3, 1024, 256, False), (2, 256, 64, True), (2, 64, 16, False),
                                       (1, 5000, 512, False), (2, 700, 700, True), (1, 20000, 128, False), (4, 1, 1, False)])
def test_farthest_point_sample(b, n, m, dup):
    from learn_region_grow_b200 import tfops as T
    x = _clouds(b, n, n + m, dup)
    got = T.farthest_point_sample(m, x)
    ref = O.farthest_point_sample(m, x)
    assert got.dtype == np.int32 and got.shape == (b, m)
    np.testing.assert_array_equal(got, ref)
    assert np.all(got[:, 0] == 0)


def test_gather_and_grad():
    from learn_region_grow_b200 import tfops as T
    x = _clouds(3, 500, 1)
    idx = np.random.RandomState(2).randint(0, 500, (3, 77)).astype(np.int32)
    np.testing.assert_array_equal(T.gather_point(x, idx), O.gather_point(x, idx))
    g = np.random.RandomState(3).randn(3, 77, 3).astype(np.float32)
    np.testing.assert_allclose(T.gather_point_grad(x, idx, g), O.gather_point_grad(x, idx, g), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('b,n,m,radius,ns', [(2, 1024, 256, 0.1, 32), (1, 1024, 1024, 0.2, 32), (3, 256, 64, 0.4, 32), (2, 64, 16, 0.8, 32),
                                             (1, 333, 77, 0.05, 5), (32, 512, 128, 0.3, 64)])
def test_query_ball_and_group(b, n, m, radius, ns):
    from learn_region_grow_b200 import tfops as T
    x = _clouds(b, n, 10 + n)
    q = x[:, :m] if b != 1 or n != 333 else _clouds(b, m, 99)       # PointNet++ queries are a subset of the points; one case is not
    idx, cnt = T.query_ball_point(radius, ns, x, q)
    ridx, rcnt = O.query_ball_point(radius, ns, x, q)
    np.testing.assert_array_equal(cnt, rcnt)
    np.testing.assert_array_equal(idx, ridx)            # rows without hits stay 0 in both
    feat = np.random.RandomState(4).randn(b, n, 7).astype(np.float32)
    np.testing.assert_array_equal(T.group_point(feat, idx), O.group_point(feat, idx))
    feat4 = np.random.RandomState(5).randn(b, n, 64).astype(np.float32)
    np.testing.assert_array_equal(T.group_point(feat4, idx), O.group_point(feat4, idx))
    go = np.random.RandomState(6).randn(b, m, ns, 7).astype(np.float32)
    np.testing.assert_allclose(T.group_point_grad(feat, idx, go), O.group_point_grad(feat, idx, go), rtol=1e-4, atol=1e-4)


def test_selection_sort_known_answer_and_random():
    from learn_region_grow_b200 import tfops as T
    dist = (10 - np.arange(16)).reshape(2, 2, 4).astype(np.float32)        # tf_ops/grouping/test/selection_sort.cpp:68-78
    outi, out = T.select_top_k(3, dist)
    assert outi.reshape(-1).tolist() == [3, 2, 1, 0] * 4
    assert out.reshape(-1).tolist() == [7, 8, 9, 10, 3, 4, 5, 6, -1, 0, 1, 2, -5, -4, -3, -2]
    d = np.random.RandomState(0).randint(0, 20, (3, 50, 333)).astype(np.float32)     # many exact ties
    gi, go = T.select_top_k(16, d)
    ri, ro = O.select_top_k(16, d)
    np.testing.assert_array_equal(gi, ri)
    np.testing.assert_array_equal(go, ro)
    val, idx = T.knn_point(4, _clouds(1, 64, 1), _clouds(1, 16, 2))
    assert val.shape == (1, 16, 4) and np.all(np.diff(val, axis=-1) >= 0)


@pytest.mark.parametrize('b,n,m,c', [(1, 64, 16, 512), (2, 256, 64, 256), (1, 1024, 256, 256), (1, 1024, 1024, 128), (2, 50, 2, 8), (1, 10, 1, 4)])
def test_three_nn_and_interpolate(b, n, m, c):
    from learn_region_grow_b200 import tfops as T
    x1, x2 = _clouds(b, n, 20 + n), _clouds(b, m, 30 + m)
    if m >= 16:
        x2[:, 5] = x2[:, 3]                              # exact ties: earliest index must win
    dist, idx = T.three_nn(x1, x2)
    rdist, ridx = O.three_nn(x1, x2)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dist, rdist)           # inf where m < 3, like (float)1e40
    pts = np.random.RandomState(7).randn(b, m, c).astype(np.float32)
    d = np.maximum(np.where(np.isfinite(dist), dist, 1e10), 1e-10)          # train_pointnet.py:146-149
    w = ((1.0 / d) / np.sum(1.0 / d, axis=2, keepdims=True)).astype(np.float32)
    np.testing.assert_array_equal(T.three_interpolate(pts, idx, w), O.three_interpolate(pts, idx, w))
    go = np.random.RandomState(8).randn(b, n, c).astype(np.float32)
    np.testing.assert_allclose(T.three_interpolate_grad(pts, idx, w, go), O.three_interpolate_grad(pts, idx, w, go), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('b,n,m', [(1, 1, 5), (2, 3, 16), (3, 64, 100), (2, 1000, 513), (2, 8192, 300), (1, 8193, 300), (2, 20001, 1000)])
def test_prob_sample(b, n, m):
    """prob_sample: index-exact against the restated cumsum / binary search and, where oracle/_ref is built, against the
    reference's own probsampleLauncher on the same device buffers (the float32 cumulative sums bit for bit)."""
    import torch
    from learn_region_grow_b200 import tfops as T, _lib
    rng = np.random.RandomState(100 + n)
    p = (rng.rand(b, n) * rng.choice([1e-3, 1.0, 50.0], size=(b, n))).astype(np.float32)
    r = rng.rand(b, m).astype(np.float32)
    r[:, 0] = 0.0
    got = T.prob_sample(p, r)
    assert got.dtype == np.int32 and got.shape == (b, m)
    np.testing.assert_array_equal(got, O.prob_sample(p, r))
    tp, tr = torch.from_numpy(p).cuda(), torch.from_numpy(r).cuda()
    tmp, out = torch.zeros(b, n, device='cuda'), torch.zeros(b, m, dtype=torch.int32, device='cuda')
    _lib.check(_lib.lib().lrg_prob_sample(b, n, m, C.c_void_p(tp.data_ptr()), C.c_void_p(tr.data_ptr()), C.c_void_p(tmp.data_ptr()),
                                          C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(tmp.cpu().numpy(), O.prob_cumsum(p))          # the cumulative sums themselves, bit for bit
    assert torch.equal(T.prob_sample(tp, tr), out)
    if O.ReferenceKernels.available():
        R = O.ReferenceKernels()
        tmp2, out2 = torch.zeros(b, n, device='cuda'), torch.zeros(b, m, dtype=torch.int32, device='cuda')
        R.prob_sample(b, n, m, C.c_void_p(tp.data_ptr()), C.c_void_p(tr.data_ptr()), C.c_void_p(tmp2.data_ptr()), C.c_void_p(out2.data_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(tmp2, tmp) and torch.equal(out2, out)


def test_op_argument_errors():
    from learn_region_grow_b200 import tfops as T
    x = _clouds(1, 16, 0)
    with pytest.raises(ValueError, match='positive radius'):
        T.query_ball_point(0.0, 4, x, x)
    with pytest.raises(ValueError, match='positive nsample'):
        T.query_ball_point(0.1, 0, x, x)
    with pytest.raises(ValueError, match='FarthestPointSample expects'):
        T.farthest_point_sample(4, x[..., :2])
    with pytest.raises(ValueError, match='ProbSample expects'):
        T.prob_sample(np.zeros((2, 4), np.float32), np.zeros((3, 5), np.float32))
    with pytest.raises(ValueError, match='positive k'):
        T.select_top_k(0, np.zeros((1, 2, 3), np.float32))


def test_torch_tensors_in_place():
    import torch
    from learn_region_grow_b200 import tfops as T
    x = torch.rand(2, 512, 3, device='cuda')
    idx = T.farthest_point_sample(64, x)
    assert idx.is_cuda and idx.dtype == torch.int32
    np.testing.assert_array_equal(idx.cpu().numpy(), O.farthest_point_sample(64, x.cpu().numpy()))
    new_xyz = T.gather_point(x, idx)
    bi, cnt = T.query_ball_point(0.2, 16, x, new_xyz)
    ri, rc = O.query_ball_point(0.2, 16, x.cpu().numpy(), new_xyz.cpu().numpy())
    np.testing.assert_array_equal(bi.cpu().numpy(), ri)


@pytest.mark.skipif(not O.ReferenceKernels.available(), reason='oracle/_ref not built (python -m oracle.build_ref in the build container)')
def test_against_reference_kernels_themselves():
    """The unmodified tf_sampling_g.cu / tf_grouping_g.cu, compiled for sm_100a, on the same device buffers."""
    from learn_region_grow_b200 import _lib
    L = _lib.lib()
    R = O.ReferenceKernels()

    def dev(a):
        p = C.c_void_p()
        _lib.check(L.lrg_malloc(C.byref(p), max(a.nbytes, 4)))
        _lib.check(L.lrg_memcpy_h2d(p, _lib.ptr(a), a.nbytes))
        return p

    def host(p, shape, dtype):
        a = np.empty(shape, dtype)
        _lib.check(L.lrg_device_synchronize())
        _lib.check(L.lrg_memcpy_d2h(_lib.ptr(a), p, a.nbytes))
        return a

    for (b, n, m, dup) in [(2, 1024, 256, False), (1, 4000, 300, False), (2, 600, 600, True)]:
        x = _clouds(b, n, 40 + n, dup)
        dx = dev(x)
        tmp = dev(np.zeros((32, n), np.float32))
        o_ref, o_new = dev(np.zeros((b, m), np.int32)), dev(np.zeros((b, m), np.int32))
        R.fps(b, n, m, dx, tmp, o_ref)
        _lib.check(L.lrg_farthest_point_sampling(b, n, m, dx, tmp, o_new, None))
        ref = host(o_ref, (b, m), np.int32)
        np.testing.assert_array_equal(host(o_new, (b, m), np.int32), ref)
        np.testing.assert_array_equal(O.farthest_point_sample(m, x), ref)       # pins the C restatement too
        # ball query + group on the sampled centres
        q = O.gather_point(x, ref)
        dq = dev(q)
        ns = 32
        i_ref, i_new = dev(np.zeros((b, m, ns), np.int32)), dev(np.zeros((b, m, ns), np.int32))
        c_ref, c_new = dev(np.zeros((b, m), np.int32)), dev(np.zeros((b, m), np.int32))
        R.query_ball(b, n, m, 0.15, ns, dx, dq, i_ref, c_ref)
        _lib.check(L.lrg_query_ball_point(b, n, m, 0.15, ns, dx, dq, i_new, c_new, None))
        ridx = host(i_ref, (b, m, ns), np.int32)
        np.testing.assert_array_equal(host(i_new, (b, m, ns), np.int32), ridx)
        np.testing.assert_array_equal(host(c_new, (b, m), np.int32), host(c_ref, (b, m), np.int32))
        np.testing.assert_array_equal(O.query_ball_point(0.15, ns, x, q)[0], ridx)
        feat = np.random.RandomState(1).randn(b, n, 16).astype(np.float32)
        df = dev(feat)
        g_ref, g_new = dev(np.zeros((b, m, ns, 16), np.float32)), dev(np.zeros((b, m, ns, 16), np.float32))
        R.group(b, n, 16, m, ns, df, i_ref, g_ref)
        _lib.check(L.lrg_group_point(b, n, 16, m, ns, df, i_ref, g_new, None))
        np.testing.assert_array_equal(host(g_new, (b, m, ns, 16), np.float32), host(g_ref, (b, m, ns, 16), np.float32))
        for p in (dx, tmp, o_ref, o_new, dq, i_ref, i_new, c_ref, c_new, df, g_ref, g_new):
            L.lrg_free(p)


@pytest.mark.parametrize('b,n,m,dup', [(1, 9000, 96, False), (2, 16384, 64, True), (1, 30000, 128, False), (1, 65536, 48, False),
                                       (1, 70000, 32, False)])
def test_farthest_point_sample_large_clouds(b, n, m, dup):
    """8,193 .. 65,536 points: the thread-block-cluster kernel (2 / 4 / 8 CTAs per cloud, points in registers, argmax exchanged
    through distributed shared memory); above that the workspace kernel.  Index-exact against the C restatement (itself pinned
    to the reference's own kernel in test_against_reference_kernels_themselves)."""
    from learn_region_grow_b200 import tfops as T
    x = _clouds(b, n, n + m, dup)
    got = T.farthest_point_sample(m, x)
    np.testing.assert_array_equal(got, O.farthest_point_sample(m, x))


def test_cluster_fps_equals_the_one_cta_kernel_and_the_reference_kernel():
    """Forcing the cluster kernel onto small clouds (lrg_fps_set_cluster_min) gives the indices of the one-CTA kernel, ties
    included; and on a 20,000-point cloud the reference's own farthestpointsamplingKernel agrees where it is available."""
    from learn_region_grow_b200 import _lib, tfops as T
    L = _lib.lib()
    clouds = [(_clouds(2, 600, 7, True), 600), (_clouds(1, 3000, 8), 257), (_clouds(3, 8192, 9), 40)]
    one = [T.farthest_point_sample(m, x) for x, m in clouds]
    try:
        _lib.check(L.lrg_fps_set_cluster_min(256))
        for (x, m), ref in zip(clouds, one):
            np.testing.assert_array_equal(T.farthest_point_sample(m, x), ref)
    finally:
        _lib.check(L.lrg_fps_set_cluster_min(0))
    if O.ReferenceKernels.available():
        R = O.ReferenceKernels()
        x = _clouds(1, 20000, 11)
        p, tmp, out = C.c_void_p(), C.c_void_p(), C.c_void_p()
        for ptr, nbytes in ((p, x.nbytes), (tmp, 32 * 20000 * 4), (out, 200 * 4)):
            _lib.check(L.lrg_malloc(C.byref(ptr), nbytes))
        _lib.check(L.lrg_memcpy_h2d(p, _lib.ptr(x), x.nbytes))
        R.fps(1, 20000, 200, p, tmp, out)
        ref = np.empty((1, 200), np.int32)
        _lib.check(L.lrg_device_synchronize())
        _lib.check(L.lrg_memcpy_d2h(_lib.ptr(ref), out, ref.nbytes))
        np.testing.assert_array_equal(T.farthest_point_sample(200, x), ref)
        for ptr in (p, tmp, out):
            L.lrg_free(ptr)


@pytest.mark.parametrize('b,n,npoint,radius,ns,c', [(2, 1024, 256, 0.15, 32, 6), (1, 4096, 512, 0.1, 32, 0), (3, 500, 64, 0.3, 16, 13),
                                                    (1, 12000, 128, 0.05, 8, 4)])
def test_sample_and_group_fused(b, n, npoint, radius, ns, c):
    """lrg_sample_and_group (train_pointnet.py:113-123 in two launches) == the separate ops == the oracle's composition, bit for
    bit: new_xyz, idx, grouped_xyz (translation-normalised) and new_points = concat(grouped_xyz, grouped features)."""
    from learn_region_grow_b200 import tfops as T
    x = _clouds(b, n, 50 + n)
    feat = np.random.RandomState(3).randn(b, n, c).astype(np.float32) if c else None
    new_xyz, new_points, idx, grouped_xyz = T.sample_and_group(npoint, radius, ns, x, feat)
    # the reference's own sequence through the separate drop-in ops
    s_new = T.gather_point(x, T.farthest_point_sample(npoint, x))
    s_idx, s_cnt = T.query_ball_point(radius, ns, x, s_new)
    s_g = T.group_point(x, s_idx) - s_new[:, :, None, :]
    s_np = np.concatenate([s_g, T.group_point(feat, s_idx)], -1) if c else s_g
    # ... and on the CPU oracle
    o_new = O.gather_point(x, O.farthest_point_sample(npoint, x))
    o_idx, _ = O.query_ball_point(radius, ns, x, o_new)
    o_g = O.group_point(x, o_idx) - o_new[:, :, None, :]
    for got, sep, ora in ((new_xyz, s_new, o_new), (idx, s_idx, o_idx), (grouped_xyz, s_g, o_g)):
        np.testing.assert_array_equal(got, sep)
        np.testing.assert_array_equal(got, ora)
    np.testing.assert_array_equal(new_points, s_np)
    assert new_points.shape == (b, npoint, ns, 3 + c)


def test_knn_point_on_the_device():
    """tf_grouping.py:48-73: distance matrix (lrg_pairwise_sqdist) + selection sort, both on the device; numpy and torch inputs."""
    from learn_region_grow_b200 import tfops as T
    x1, x2 = _clouds(2, 300, 1), _clouds(2, 40, 2)
    val, idx = T.knn_point(5, x1, x2)
    d = ((x1[:, None, :, :] - x2[:, :, None, :]) ** 2)
    dist = (d[..., 0] + d[..., 1]) + d[..., 2]
    ri, ro = O.select_top_k(5, dist.astype(np.float32))
    np.testing.assert_array_equal(idx, ri[:, :, :5])
    np.testing.assert_array_equal(val, ro[:, :, :5])
    torch = pytest.importorskip('torch')
    tv, ti = T.knn_point(5, torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda())
    assert ti.is_cuda
    np.testing.assert_array_equal(ti.cpu().numpy(), idx)
    np.testing.assert_array_equal(tv.cpu().numpy(), val)
