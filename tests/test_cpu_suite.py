"""CPU-side checks: the C-ABI library loads and exports what include/lrg_b200.h declares, checkpoint I/O, the import
stand-ins, the oracle's known answers.  No compute call is made without a GPU."""
import os
import re
import sys

import numpy as np
import pytest

from oracle import feature_prep

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from learn_region_grow_b200 import _lib, build
    build.build()
    header = open(os.path.join(REPO, 'include', 'lrg_b200.h')).read()
    declared = set(re.findall(r'\b(lrg_[a-z0-9_]+)\s*\(', header))
    declared.discard('lrg_stream_t')
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), 'missing export %s' % name
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().lrg_version() == 100


def test_no_cpu_fallback_without_gpu():
    from learn_region_grow_b200 import _lib
    if _lib.lib().lrg_device_count() > 0:
        pytest.skip('a GPU is present')
    from learn_region_grow_b200.engine import Engine
    with pytest.raises(_lib.LrgError, match='no CPU fallback'):
        Engine()
    from learn_region_grow_b200 import tfops
    with pytest.raises(_lib.LrgError):
        tfops.farthest_point_sample(4, np.zeros((1, 8, 3), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, 'learn_region_grow_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), os.path.join(root, f)


def test_checkpoint_roundtrip(tmp_path, golden_weights):
    from learn_region_grow_b200 import ckpt
    prefix = str(tmp_path / 'models' / 'lrgnet_model5.ckpt')
    extra = dict(golden_weights)
    extra['Variable'] = np.array(70310, dtype=np.int32)
    ckpt.save_checkpoint(prefix, extra)
    back = ckpt.load_checkpoint(prefix)
    assert set(back) == set(extra)
    for k in extra:
        assert back[k].shape == np.asarray(extra[k]).shape and np.array_equal(back[k], extra[k])
    with pytest.raises(ValueError):
        open(prefix + '.index', 'wb').write(b'not a table' * 10)
        ckpt.read_index(prefix)


@pytest.mark.skipif(not os.path.exists('/root/reference/models/lrgnet_model5.ckpt.index'), reason='reference checkpoint not present')
def test_reader_on_the_shipped_checkpoint(golden_weights):
    from learn_region_grow_b200 import ckpt
    t = ckpt.load_checkpoint('/root/reference/models/lrgnet_model5.ckpt')
    for k, v in golden_weights.items():
        assert np.array_equal(t[k], v)
    # the stored crc32c of a tensor matches ours (masked, as the table format stores it)
    e = ckpt.read_index('/root/reference/models/lrgnet_model5.ckpt')['lrg_bias0']
    raw = open('/root/reference/models/lrgnet_model5.ckpt.data-00000-of-00001', 'rb').read()[e['offset']:e['offset'] + e['size']]
    assert ckpt._mask_crc(ckpt._crc32c(raw)) == e['crc32c']


def test_golden_weights_inventory(golden_weights):
    from learn_region_grow_b200.engine import variable_shapes, pack_weights
    shapes = variable_shapes(13, 0)
    assert len(shapes) == 32 and sum(int(np.prod(s)) for _, s in shapes) == 791044
    blob = pack_weights(golden_weights)
    assert blob.dtype == np.float32 and blob.size == 791044
    assert np.array_equal(blob[:13 * 64], golden_weights['lrg_kernel0'].reshape(-1))


def test_philox_known_answers():
    from oracle import philox
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:            # Random123 kat_vectors, philox4x32 10 rounds
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want
    u = philox.u32_to_unit_float(philox.draw_u32(5, 1, 2, 3, 1000))
    assert u.dtype == np.float32 and 0 <= u.min() and u.max() < 1


def test_philox_sampling_is_uniform_without_replacement():
    from oracle.lrg_driver import PhiloxRng
    rng = PhiloxRng(1)
    counts = np.zeros(2000)
    for step in range(200):
        rng.begin_step(0, step)
        s = rng.sample(2000, 512, 'inlier')
        assert len(np.unique(s)) == 512 and np.all(np.diff(s) > 0)
        counts[s] += 1
    assert abs(counts.mean() - 200 * 512 / 2000) < 1e-9 and counts.std() < 10
    rng.begin_step(0, 0)
    s = rng.sample(100, 512, 'neighbor')
    assert np.array_equal(s[:100], np.arange(100)) and s.max() < 100 and len(s) == 512


def test_oracle_tfops_known_answer():
    from oracle import tfops
    want = open(os.path.join(REPO, 'tests', 'golden', 'selection_sort_kat.txt')).read().strip().split('\n')
    dist = np.array([float(x) for x in want[0].split()], np.float32).reshape(2, 2, 4)       # output of the reference's own
    outi, out = tfops.select_top_k(3, dist)                                                   # selection_sort.cpp (golden)
    assert outi.reshape(-1).tolist() == [int(x) for x in want[-2].split()]
    assert out.reshape(-1).tolist() == [float(x) for x in want[-1].split()]
    x = np.random.RandomState(0).rand(2, 200, 3).astype(np.float32)
    idx = tfops.farthest_point_sample(50, x)
    assert np.all(idx[:, 0] == 0) and all(len(set(r)) == 50 for r in idx)
    # brute-force farthest point property
    for b in range(2):
        d = np.full(200, np.inf)
        for j in range(1, 50):
            d = np.minimum(d, ((x[b] - x[b, idx[b, j - 1]]) ** 2).sum(1))
            assert d[idx[b, j]] >= d.max() - 1e-6
    dist3, i3 = tfops.three_nn(x, x[:, :40])
    bf = ((x[:, :, None, :] - x[:, None, :40, :]) ** 2).sum(-1)
    assert np.array_equal(i3[..., 0], bf.argmin(-1))
    bi, cnt = tfops.query_ball_point(0.2, 8, x, x[:, :30])
    assert np.all(bi[:, :, 0] <= np.arange(30)) and cnt.max() <= 8 and cnt.min() >= 1


def test_oracle_prob_sample_properties():
    """oracle.tfops.prob_sample (restatement of cumsumKernel + binarysearchKernel, tf_sampling_g.cu:7-104): the float32
    cumulative sums are monotone and within float32 rounding of the exact sums; an index is the first category whose
    cumulative sum reaches uniform * total; one-hot rows return the hot category; multi-chunk rows (> 8192) carry over."""
    from oracle import tfops
    rng = np.random.RandomState(3)
    for n in (1, 5, 64, 1000, 8193, 20000):
        p = rng.rand(2, n).astype(np.float32)
        c = tfops.prob_cumsum(p)
        exact = np.cumsum(p.astype(np.float64), axis=1)
        assert np.all(np.diff(c, axis=1) >= 0) and np.abs(c - exact).max() <= 1e-6 * exact.max()
    p = rng.rand(3, 777).astype(np.float32)
    r = rng.rand(3, 200).astype(np.float32)
    idx = tfops.prob_sample(p, r)
    c = tfops.prob_cumsum(p)
    assert idx.dtype == np.int32 and idx.shape == (3, 200)
    for i in range(3):
        np.testing.assert_array_equal(idx[i], np.searchsorted(c[i], (r[i] * c[i, -1]).astype(np.float32), side='left'))
    hot = np.zeros((1, 100), np.float32)
    hot[0, 37] = 2.5
    assert set(tfops.prob_sample(hot, r[:1])[0].tolist()) == {37}
    freq = np.bincount(tfops.prob_sample(np.array([[1, 3]], np.float32), rng.rand(1, 4000).astype(np.float32))[0], minlength=2) / 4000.0
    assert abs(freq[1] - 0.75) < 0.03


def test_numpy_row_sum_order_assumption():
    """The fill kernel restates numpy's float32 row sum (pairwise: 8 accumulators then a sequential tail)."""
    rng = np.random.RandomState(0)
    a = (rng.randn(5000, 13) * rng.rand(5000, 1) * 100).astype(np.float32) ** 2
    want = np.sum(a, axis=1)
    f = np.float32
    got = ((a[:, 0] + a[:, 1]) + (a[:, 2] + a[:, 3])) + ((a[:, 4] + a[:, 5]) + (a[:, 6] + a[:, 7]))
    for c in range(8, 13):
        got = got + a[:, c]
    assert got.dtype == f and np.array_equal(got, want)
    for F in (6, 9, 12):
        b = a[:, :F].copy()
        want = np.sum(b, axis=1)
        if F < 8:
            got = np.zeros(len(b), f)
            for c in range(F):
                got = got + b[:, c]
        else:
            got = ((b[:, 0] + b[:, 1]) + (b[:, 2] + b[:, 3])) + ((b[:, 4] + b[:, 5]) + (b[:, 6] + b[:, 7]))
            for c in range(8, F):
                got = got + b[:, c]
        assert np.array_equal(got, want), F


def test_forward_oracle_fp32_vs_fp64(golden_weights):
    from oracle import lrg_forward
    rng = np.random.RandomState(0)
    a, b = rng.randn(2, 512, 13).astype(np.float32), rng.randn(2, 512, 13).astype(np.float32)
    add32, rmv32 = lrg_forward.forward(golden_weights, a, b)
    add64, rmv64 = lrg_forward.forward(golden_weights, a, b, dtype=np.float64)
    assert add32.shape == (2, 512, 2) and add32.dtype == np.float32
    scale = np.abs(add64).max()
    assert np.abs(add32 - add64).max() < 1e-5 * scale and np.abs(rmv32 - rmv64).max() < 1e-5 * np.abs(rmv64).max()
    # factored head (what the kernels compute) == tiled/concatenated head (what the graph says)
    W = {k: v.astype(np.float64) for k, v in golden_weights.items()}
    h = b.astype(np.float64)
    acts = []
    for i in range(5):
        h = np.maximum(h @ W['lrg_neighbor_kernel%d' % i][0] + W['lrg_neighbor_bias%d' % i], 0)
        acts.append(h)
    hi = a.astype(np.float64)
    for i in range(5):
        hi = np.maximum(hi @ W['lrg_kernel%d' % i][0] + W['lrg_bias%d' % i], 0)
    g = np.concatenate([hi.max(1), acts[-1].max(1)], 1)
    K0 = W['lrg_add_kernel0'][0]
    z = np.maximum((g @ K0[:1024])[:, None, :] + acts[1] @ K0[1024:] + W['lrg_add_bias0'], 0)
    z = np.maximum(z @ W['lrg_add_kernel1'][0] + W['lrg_add_bias1'], 0) @ W['lrg_add_kernel2'][0] + W['lrg_add_bias2']
    assert np.abs(z - add64).max() < 1e-9 * scale


def test_session_standin_contract():
    dropin = os.path.join(REPO, 'learn_region_grow_b200', 'dropin')
    sys.path.insert(0, dropin)
    try:
        sys.modules.pop('tensorflow', None)
        import tensorflow as tf

        class Net:
            def __init__(self):
                self.a = tf.Handle(self, 'a')
                self.x = tf.Handle(self, 'x')
                tf.register_net(self)
                self.loaded = None

            def _evaluate(self, names, feeds):
                return {'a': feeds['x'] * 2}

            def _load_variables(self, tensors):
                self.loaded = tensors

        tf.compat.v1.reset_default_graph()
        cfg = tf.compat.v1.ConfigProto()
        cfg.gpu_options.allow_growth = True
        cfg.allow_soft_placement = True
        sess = tf.compat.v1.Session(config=cfg)
        net = Net()
        assert sess.run([net.a], {net.x: np.arange(3)})[0].tolist() == [0, 2, 4]
        assert sess.run(net.a, {net.x: 4}) == 8
        with pytest.raises(TypeError):
            sess.run(['nope'])
        import tensorflow.compat.v1 as v1
        assert v1.Session is tf.compat.v1.Session
    finally:
        sys.path.remove(dropin)
        sys.modules.pop('tensorflow', None)
        sys.modules.pop('tensorflow.compat', None)
        sys.modules.pop('tensorflow.compat.v1', None)


def test_h5_standin_and_loadFromH5(tmp_path):
    standins = os.path.join(REPO, 'learn_region_grow_b200', 'dropin', 'standins')
    sys.path.append(standins)
    try:
        from learn_region_grow_b200 import io_util
        from tools import rooms
        rs = [rooms.generate_room(1000, n_raw=500, n_boxes=2), rooms.generate_room(1001, n_raw=700, n_boxes=2)]
        path = str(tmp_path / 's3dis_area5.h5')
        io_util.saveToH5(path, rs)
        pts, obj, cls = io_util.loadFromH5(path)
        assert [len(p) for p in pts] == [len(r) for r in rs] and pts[0].shape[1] == 6
        assert np.array_equal(obj[1], rs[1][:, 6].astype(int)) and np.array_equal(cls[0], rs[0][:, 7].astype(int))
        flat = io_util.loadFromH5(path, load_labels=False)
        assert np.array_equal(flat[1], rs[1])
        io_util.savePLY(str(tmp_path / 'a.ply'), np.c_[rs[0][:5, :3], np.zeros((5, 3))])
        assert open(tmp_path / 'a.ply').read().startswith('ply\nformat ascii 1.0\nelement vertex 5\n')
        io_util.savePCD(str(tmp_path / 'a.pcd'), np.c_[rs[0][:5, :3], np.ones((5, 3))])
        assert 'POINTS 5' in open(tmp_path / 'a.pcd').read()
    finally:
        sys.path.remove(standins)
        sys.modules.pop('h5py', None)


@pytest.mark.parametrize('seed', [1000, 1001])
def test_metrics_oracle_reproduces_the_reference_log(seed):
    """oracle/metrics.py is pinned by the statistics line the UNMODIFIED reference driver printed for the golden rooms
    (test_region_grow.py:349): same obj_id (raw column 6 at equalized_idx) and the reference's own final cluster labels."""
    from oracle import metrics as om
    from tools import rooms
    z = np.load(os.path.join(REPO, 'tests', 'golden', 'driver_trace_%d.npz' % seed), allow_pickle=True)
    room = z['room']
    f = feature_prep.prepare_features(room, 0.1)
    obj_id = room[f['equalized_idx'], 6].astype(int)
    assert len(obj_id) == len(z['cluster_label'])
    m = om.room_statistics(obj_id, z['cluster_label'])
    line = [l for l in str(z['log']).split('\n') if l.startswith('Area 5 room 0 NMI')][0]
    logged = dict(zip(('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou'), [float(x) for x in re.findall(r': (\d\.\d\d)', line)]))
    for k, v in logged.items():
        assert abs(m[k] - v) <= 0.005 + 1e-9, (k, m[k], v)


def test_abi_struct_layouts_match_the_header(tmp_path):
    """The ctypes / numpy mirrors of the structs that cross the C ABI have the layout gcc gives include/lrg_b200.h."""
    import ctypes
    import subprocess
    from learn_region_grow_b200 import _lib
    src = tmp_path / 'layout.c'
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "lrg_b200.h"
int main(void) {
  printf("LrgGrowParams %zu %zu %zu %zu %zu %zu\\n", sizeof(LrgGrowParams), offsetof(LrgGrowParams, seed), offsetof(LrgGrowParams, flags), offsetof(LrgGrowParams, num_restarts), offsetof(LrgGrowParams, beam_width), offsetof(LrgGrowParams, search_width));
  printf("LrgRoomStats %zu %zu\\n", sizeof(LrgRoomStats), offsetof(LrgRoomStats, stop_other));
  printf("LrgStepTrace %zu %zu %zu %zu\\n", sizeof(LrgStepTrace), offsetof(LrgStepTrace, center), offsetof(LrgStepTrace, neighbor_idx_crc), offsetof(LrgStepTrace, score));
  printf("LrgRoomMetrics %zu %zu %zu\\n", sizeof(LrgRoomMetrics), offsetof(LrgRoomMetrics, iou), offsetof(LrgRoomMetrics, gt_match));
  return 0;
}
''')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict((l.split()[0], [int(x) for x in l.split()[1:]]) for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    G = _lib.GrowParams
    assert out['LrgGrowParams'] == [ctypes.sizeof(G), G.seed.offset, G.flags.offset, G.num_restarts.offset, G.beam_width.offset, G.search_width.offset]
    assert out['LrgRoomStats'] == [_lib.ROOM_STATS_DTYPE.itemsize, _lib.ROOM_STATS_DTYPE.fields['stop_other'][1]]
    T = _lib.STEP_TRACE_DTYPE
    assert out['LrgStepTrace'] == [T.itemsize, T.fields['center'][1], T.fields['neighbor_idx_crc'][1], T.fields['score'][1]]
    M = _lib.ROOM_METRICS_DTYPE
    assert out['LrgRoomMetrics'] == [M.itemsize, M.fields['iou'][1], M.fields['gt_match'][1]]
