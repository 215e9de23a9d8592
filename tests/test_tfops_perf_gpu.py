"""tf_ops kernels of liblrg_b200.so timed beside the reference's own kernels (tf_sampling_g.cu / tf_grouping_g.cu compiled
unmodified for sm_100a into oracle/_ref) on the same device buffers, at the call shapes of the reference's only consumer
(`benchmarks.py --mode pointnet2` -> train_pointnet.py:181-190: four set-abstraction levels, B = 1 at inference, B = 100 in
training).  three_nn / three_interpolate are CPU ops in the reference (tf_interpolate.cpp:60-127): their C restatement is
timed on the host.  CUDA events on the legacy default stream, 3 warm-up + 20 timed launches per op; the arrays are a few MB at
most, i.e. L2-resident by nature (stated, not flushed).  The test asserts equal outputs only; the table goes to stdout and,
when the directory exists, to gpurun_out/tfops_perf.txt (copied to profiles/ by hand)."""
import ctypes as C
import os
import time

import numpy as np
import pytest

from oracle import tfops as O

pytestmark = pytest.mark.gpu

LEVELS = [(1024, 1024, 0.1, 3), (1024, 256, 0.2, 64), (256, 64, 0.4, 128), (64, 16, 0.8, 256)]   # n, m, radius, channels grouped
NSAMPLE = 32


def _time(torch, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps          # us per launch


@pytest.mark.skipif(not O.ReferenceKernels.available(), reason='oracle/_ref not built (python -m oracle.build_ref in the build container)')
def test_tfops_beside_reference_kernels():
    import torch
    from learn_region_grow_b200 import _lib
    L = _lib.lib()
    R = O.ReferenceKernels()
    p = lambda t: C.c_void_p(t.data_ptr())
    lines = ['%-22s %5s %5s %5s %4s | %10s %10s %7s | %9s' % ('op', 'B', 'n', 'm', 'c', 'ours us', 'ref us', 'ratio', 'ours GB/s')]
    rng = np.random.RandomState(5)
    for B in (1, 100):
        for (n, m, radius, c) in LEVELS:
            xyz = torch.from_numpy(rng.rand(B, n, 3).astype(np.float32)).cuda()
            tmp = torch.zeros(32, n, device='cuda')
            o_new = torch.zeros(B, m, dtype=torch.int32, device='cuda')
            o_ref = torch.zeros_like(o_new)
            t_new = _time(torch, lambda: _lib.check(L.lrg_farthest_point_sampling(B, n, m, p(xyz), p(tmp), p(o_new), None)))
            t_ref = _time(torch, lambda: R.fps(B, n, m, p(xyz), p(tmp), p(o_ref)))
            assert torch.equal(o_new, o_ref)
            lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('farthest_point_sample', B, n, m, '-', t_new, t_ref, t_ref / t_new,
                                                                                   B * (n * 12 + m * 4) / t_new / 1e3))
            q_new, q_ref = torch.zeros(B, m, 3, device='cuda'), torch.zeros(B, m, 3, device='cuda')
            t_new = _time(torch, lambda: _lib.check(L.lrg_gather_point(B, n, m, p(xyz), p(o_ref), p(q_new), None)))
            t_ref = _time(torch, lambda: R.gather(B, n, m, p(xyz), p(o_ref), p(q_ref)))
            assert torch.equal(q_new, q_ref)
            lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('gather_point', B, n, m, '-', t_new, t_ref, t_ref / t_new,
                                                                                   B * m * 28 / t_new / 1e3))
            i_new = torch.zeros(B, m, NSAMPLE, dtype=torch.int32, device='cuda')
            i_ref = torch.zeros_like(i_new)
            c_new = torch.zeros(B, m, dtype=torch.int32, device='cuda')
            c_ref = torch.zeros_like(c_new)
            t_new = _time(torch, lambda: _lib.check(L.lrg_query_ball_point(B, n, m, radius, NSAMPLE, p(xyz), p(q_ref), p(i_new), p(c_new), None)))
            t_ref = _time(torch, lambda: R.query_ball(B, n, m, radius, NSAMPLE, p(xyz), p(q_ref), p(i_ref), p(c_ref)))
            assert torch.equal(c_new, c_ref)
            hit = c_ref > 0                                  # (rows without a hit are left untouched by both, tf_grouping_g.cu:3-36)
            assert torch.equal(i_new[hit], i_ref[hit])
            lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('query_ball_point', B, n, m, '-', t_new, t_ref, t_ref / t_new,
                                                                                   B * (n * 12 + m * (12 + 4 * NSAMPLE + 4)) / t_new / 1e3))
            feat = torch.from_numpy(rng.randn(B, n, c).astype(np.float32)).cuda()
            i_ok = torch.where(hit[..., None], i_ref, torch.zeros_like(i_ref)).contiguous()
            g_new = torch.zeros(B, m, NSAMPLE, c, device='cuda')
            g_ref = torch.zeros_like(g_new)
            t_new = _time(torch, lambda: _lib.check(L.lrg_group_point(B, n, c, m, NSAMPLE, p(feat), p(i_ok), p(g_new), None)))
            t_ref = _time(torch, lambda: R.group(B, n, c, m, NSAMPLE, p(feat), p(i_ok), p(g_ref)))
            assert torch.equal(g_new, g_ref)
            lines.append('%-22s %5d %5d %5d %4d | %10.1f %10.1f %7.2f | %9.2f' % ('group_point', B, n, m, c, t_new, t_ref, t_ref / t_new,
                                                                                   B * m * NSAMPLE * (4 + 8 * c) / t_new / 1e3))
            # sample_and_group (train_pointnet.py:113-123) as lrg_sample_and_group's two launches beside the reference's five
            # kernel launches (FPS, gather, ball query, group xyz, group features; its two elementwise graph ops not counted)
            gx_ref = torch.zeros(B, m, NSAMPLE, 3, device='cuda')
            np_new = torch.zeros(B, m, NSAMPLE, 3 + c, device='cuda')
            gx_new = torch.zeros(B, m, NSAMPLE, 3, device='cuda')
            nx_new = torch.zeros(B, m, 3, device='cuda')
            f_new = torch.zeros(B, m, dtype=torch.int32, device='cuda')

            def ref_chain():
                R.fps(B, n, m, p(xyz), p(tmp), p(o_ref))
                R.gather(B, n, m, p(xyz), p(o_ref), p(q_ref))
                R.query_ball(B, n, m, radius, NSAMPLE, p(xyz), p(q_ref), p(i_ref), p(c_ref))
                R.group(B, n, 3, m, NSAMPLE, p(xyz), p(i_ok), p(gx_ref))
                R.group(B, n, c, m, NSAMPLE, p(feat), p(i_ok), p(g_ref))
            i_new.zero_()
            t_new = _time(torch, lambda: _lib.check(L.lrg_sample_and_group(B, n, m, radius, NSAMPLE, c, p(xyz), p(feat), None, p(f_new), p(nx_new),
                                                                           p(np_new), p(i_new), p(c_new), p(gx_new), None)))
            t_ref = _time(torch, ref_chain)
            assert torch.equal(f_new, o_ref) and torch.equal(nx_new, q_ref) and torch.equal(i_new[hit], i_ref[hit])
            assert torch.equal(np_new[..., 3:][hit], g_ref[hit]) and torch.equal(gx_new[hit], (gx_ref - q_ref[:, :, None, :])[hit])
            lines.append('%-22s %5d %5d %5d %4d | %10.1f %10.1f %7.2f | %9.2f' % ('sample_and_group', B, n, m, c, t_new, t_ref, t_ref / t_new,
                                                                                   B * (n * (12 + 4 * c) + m * NSAMPLE * (4 + 4 * (3 + c))) / t_new / 1e3))
        if B == 1:
            # large clouds: one CTA up to 8,192 points, a cluster of 2 / 4 / 8 CTAs (distributed shared memory) up to 65,536
            for n, m in [(8192, 256), (16384, 256), (32768, 256), (65536, 256)]:
                xyz = torch.from_numpy(rng.rand(B, n, 3).astype(np.float32)).cuda()
                tmp = torch.zeros(32, n, device='cuda')
                o_new = torch.zeros(B, m, dtype=torch.int32, device='cuda')
                o_ref = torch.zeros_like(o_new)
                t_new = _time(torch, lambda: _lib.check(L.lrg_farthest_point_sampling(B, n, m, p(xyz), p(tmp), p(o_new), None)), reps=5, warm=1)
                t_ref = _time(torch, lambda: R.fps(B, n, m, p(xyz), p(tmp), p(o_ref)), reps=5, warm=1)
                assert torch.equal(o_new, o_ref)
                lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('fps (large cloud)', B, n, m, '-', t_new, t_ref, t_ref / t_new,
                                                                                       B * (n * 12 + m * 4) / t_new / 1e3))
                if n == 8192:
                    _lib.check(L.lrg_fps_set_cluster_min(4096))
                    t_cl = _time(torch, lambda: _lib.check(L.lrg_farthest_point_sampling(B, n, m, p(xyz), p(tmp), p(o_new), None)), reps=5, warm=1)
                    _lib.check(L.lrg_fps_set_cluster_min(0))
                    assert torch.equal(o_new, o_ref)
                    lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('fps (8192 as cluster)', B, n, m, '-', t_cl, t_ref, t_ref / t_cl,
                                                                                           B * (n * 12 + m * 4) / t_cl / 1e3))
        for (n, m) in [(1024, 1024), (20000, 4096)]:                      # prob_sample: n categories, m draws per row
            pr = torch.from_numpy(rng.rand(B, n).astype(np.float32)).cuda()
            un = torch.from_numpy(rng.rand(B, m).astype(np.float32)).cuda()
            tmp_new, tmp_ref = torch.zeros(B, n, device='cuda'), torch.zeros(B, n, device='cuda')
            s_new = torch.zeros(B, m, dtype=torch.int32, device='cuda')
            s_ref = torch.zeros_like(s_new)
            t_new = _time(torch, lambda: _lib.check(L.lrg_prob_sample(B, n, m, p(pr), p(un), p(tmp_new), p(s_new), None)))
            t_ref = _time(torch, lambda: R.prob_sample(B, n, m, p(pr), p(un), p(tmp_ref), p(s_ref)))
            assert torch.equal(tmp_new, tmp_ref) and torch.equal(s_new, s_ref)
            lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('prob_sample', B, n, m, '-', t_new, t_ref, t_ref / t_new,
                                                                                   B * (n * 8 + m * 8) / t_new / 1e3))
        # feature propagation: three_nn(xyz1 (n), xyz2 (m)) + three_interpolate -- CPU ops in the reference
        for (n, m, c) in [(64, 16, 512), (256, 64, 256), (1024, 256, 256), (1024, 1024, 128)]:
            x1 = rng.rand(B, n, 3).astype(np.float32)
            x2 = rng.rand(B, m, 3).astype(np.float32)
            pts = rng.randn(B, m, c).astype(np.float32)
            t0 = time.perf_counter()
            d_ref, k_ref = O.three_nn(x1, x2)
            t_nn_ref = (time.perf_counter() - t0) * 1e6
            w = (1.0 / np.maximum(d_ref, 1e-10))
            w = (w / w.sum(axis=2, keepdims=True)).astype(np.float32)
            t0 = time.perf_counter()
            out_ref = O.three_interpolate(pts, k_ref, w)
            t_it_ref = (time.perf_counter() - t0) * 1e6
            tx1, tx2, tp, tw = (torch.from_numpy(a).cuda() for a in (x1, x2, pts, w))
            d_new = torch.zeros(B, n, 3, device='cuda')
            k_new = torch.zeros(B, n, 3, dtype=torch.int32, device='cuda')
            t_new = _time(torch, lambda: _lib.check(L.lrg_three_nn(B, n, m, p(tx1), p(tx2), p(d_new), p(k_new), None)))
            np.testing.assert_array_equal(k_new.cpu().numpy(), k_ref)
            np.testing.assert_array_equal(d_new.cpu().numpy(), d_ref)
            lines.append('%-22s %5d %5d %5d %4s | %10.1f %10.1f %7.2f | %9.2f' % ('three_nn (ref: CPU)', B, n, m, '-', t_new, t_nn_ref, t_nn_ref / t_new,
                                                                                   B * (n * 36 + m * 12) / t_new / 1e3))
            o_new2 = torch.zeros(B, n, c, device='cuda')
            tk = torch.from_numpy(k_ref).cuda()
            t_new = _time(torch, lambda: _lib.check(L.lrg_three_interpolate(B, m, c, n, p(tp), p(tk), p(tw), p(o_new2), None)))
            np.testing.assert_array_equal(o_new2.cpu().numpy(), out_ref)
            lines.append('%-22s %5d %5d %5d %4d | %10.1f %10.1f %7.2f | %9.2f' % ('three_interp (ref: CPU)', B, n, m, c, t_new, t_it_ref, t_it_ref / t_new,
                                                                                   B * (n * (24 + 4 * c) + m * c * 4) / t_new / 1e3))
    report = '\n'.join(lines)
    print('\n' + report)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, 'tfops_perf.txt'), 'w') as f:
            f.write(report + '\n')
