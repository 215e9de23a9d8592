"""On-device region-grow driver vs the CPU oracle (oracle/lrg_driver.py, itself pinned to the unmodified reference
driver by tests/test_oracle_driver.py), through the C ABI.

Small rooms: the device records every grow step; the oracle is re-driven step by step with the same Philox stream and
must agree on seeds, set sizes, medians, sampled indices, masks, stop reasons and final labels.  The masks are sampled
(u < confidence, test_region_grow.py:266-267), so a draw that lands within the forward tolerance of the confidence may
legitimately differ; such a draw is adopted from the device (and counted) so the trajectories stay comparable.

Full-size rooms: size-independent properties (determinism, slot-count invariance, label invariants).
"""
import numpy as np
import pytest

from oracle import feature_prep

from oracle import lrg_driver, lrg_forward
from util_rooms import golden_room, idx_crc, unpack_mask

pytestmark = pytest.mark.gpu

NEAR_TIE = 2e-4       # |u - confidence| below which a mask bit may differ (forward tolerance, tests/test_forward_gpu.py)
STOP_NAMES = {2: 'noexpand', 3: 'stuck', 4: 'maxsteps', 5: 'empty'}


@pytest.fixture(scope='module')
def engine(golden_weights):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    yield e
    e.close()


def _replay(points, order, weights, trace, n_steps, seed, room_id=0, resolution=0.1):
    """Re-drive the oracle with the device's trace.  Returns (grower, adopted near-tie bits)."""
    fwd = lambda a, b: lrg_forward.forward(weights, a, b)
    g = lrg_driver.RoomGrower(points, order, fwd, lrg_driver.PhiloxRng(seed), room_id=room_id, resolution=resolution)
    adopted = 0
    t = 0
    for seed_id in np.arange(len(points))[order]:
        if g.visited[seed_id]:
            continue
        g.begin_region(seed_id)
        while True:
            st = g.prepare_step()
            if st is None:
                break
            assert t < n_steps, 'oracle wants more steps than the device ran'
            rec = trace[t]
            assert rec['seed_point'] == seed_id and rec['step_in_region'] == g.steps
            assert rec['n_inlier'] == st['n_inlier'] and rec['n_neighbor'] == st['n_neighbor']
            np.testing.assert_array_equal(rec['center'][:13][[0, 1, 6, 7, 8, 9, 10, 11, 12]], st['center'][[0, 1, 6, 7, 8, 9, 10, 11, 12]])
            assert rec['inlier_idx_crc'] == idx_crc(st['inlier_idx']) and rec['neighbor_idx_crc'] == idx_crc(st['neighbor_idx'])
            add, rmv = fwd(st['inlier'], st['neighbor'])
            dev_add, dev_rmv = unpack_mask(rec['add_mask']), unpack_mask(rec['remove_mask'])
            # oracle's own masks, to bound the disagreement
            add_conf, rmv_conf = lrg_driver.confidence(add[0]), lrg_driver.confidence(rmv[0])
            rng = lrg_driver.PhiloxRng(seed)
            rng.begin_step(room_id, g.steps, 0, g.seed_id)
            u_add, u_rmv = rng.uniform(512, 'add'), rng.uniform(512, 'remove')
            for dev, conf, u in ((dev_add, add_conf, u_add), (dev_rmv, rmv_conf, u_rmv)):
                differ = dev != (u < conf)
                assert np.all(np.abs(u[differ] - conf[differ]) < NEAR_TIE), 'mask bit differs outside the near-tie band'
                adopted += int(differ.sum())
            reason = g.apply_step(add[0], rmv[0], add_mask=dev_add, rmv_mask=dev_rmv)
            assert STOP_NAMES.get(int(rec['stop_reason'])) == reason
            assert rec['size_after'] == int(g.currentMask.sum()) if reason is None else True
            t += 1
            if reason is not None:
                break
    assert t == n_steps
    return g, adopted


@pytest.mark.parametrize('room_seed,rng_seed', [(1000, 0), (1001, 12345)])
def test_device_driver_replays_on_oracle(engine, golden_weights, room_seed, rng_seed):
    points, order = golden_room(room_seed)
    engine.upload_rooms([points], [order], resolution=0.1)
    stats = engine.segment_resident(resolution=0.1, seed=rng_seed, trace_capacity=4096)
    trace, n_steps = engine.trace(0, 4096)
    assert n_steps == stats['grow_steps'][0] and n_steps <= 4096 and n_steps > 50
    g, adopted = _replay(points, order, golden_weights, trace, n_steps, rng_seed)
    assert adopted <= 3
    raw = engine.labels(filled=False)[0]
    np.testing.assert_array_equal(raw, g.cluster_label)                    # cluster_label before the fill (:176,214)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())  # after the nearest-neighbour fill (:308-316)
    assert stats['regions'][0] == len(g.regions) and stats['clusters'][0] == g.cluster_id - 1
    by_reason = {r: sum(1 for x in g.regions if x[3] == r) for r in ('noneighbor', 'noexpand', 'stuck')}
    assert (stats['stop_noneighbor'][0], stats['stop_noexpand'][0], stats['stop_stuck'][0]) == \
        (by_reason['noneighbor'], by_reason['noexpand'], by_reason['stuck'])


def test_multi_room_batch_matches_single_room_runs(engine, golden_weights):
    """Rooms are independent units: growing them together (any slot count) gives the labels of growing them alone."""
    from tools import rooms as R
    feats = [feature_prep.prepare_features(R.generate_room(1000 + i, n_raw=3000 + 1500 * i, n_boxes=5)) for i in range(4)]
    pts = [f['points'] for f in feats] + [np.zeros((0, 13), np.float32)]      # plus an empty room
    orders = [f['order'] for f in feats] + [np.zeros(0, np.int64)]
    together, stats = engine.segment_rooms(pts, orders, resolution=0.1, seed=7, max_slots=3)
    assert stats['n_points'].tolist() == [len(p) for p in pts]
    for i in range(len(pts)):
        alone, st1 = engine.segment_rooms([pts[i]], [orders[i]], resolution=0.1, seed=7, room_id_base=i)
        np.testing.assert_array_equal(alone[0], together[i])
        assert st1['grow_steps'][0] == stats['grow_steps'][i]
    again, _ = engine.segment_rooms(pts, orders, resolution=0.1, seed=7, max_slots=5)
    for a, b in zip(again, together):
        np.testing.assert_array_equal(a, b)
    other, _ = engine.segment_rooms(pts, orders, resolution=0.1, seed=8)
    assert any(not np.array_equal(a, b) for a, b in zip(other, together))      # the seed matters


def test_full_size_room_properties(engine):
    """S3DIS-shaped rooms (~20k raw points, BASELINE.json config): invariants that hold at any size."""
    from tools import rooms as R
    feats = [feature_prep.prepare_features(R.generate_room(1000 + i)) for i in range(3)]
    pts, orders = [f['points'] for f in feats], [f['order'] for f in feats]
    labels, stats = engine.segment_rooms(pts, orders, resolution=0.1, seed=0)
    engine_raw = engine.labels(filled=False)
    for i, (lab, raw) in enumerate(zip(labels, engine_raw)):
        assert lab.shape == (len(pts[i]),)
        assert lab.min() >= 1 and lab.max() <= stats['clusters'][i]            # every point labelled after the fill
        assert np.array_equal(lab[raw > 0], raw[raw > 0])                      # the fill never relabels
        sizes = np.bincount(raw)[1:]
        assert len(sizes) == stats['clusters'][i] and sizes.min() > 10         # cluster_threshold (:33)
        assert stats['regions'][i] == stats['stop_noneighbor'][i] + stats['stop_noexpand'][i] + stats['stop_stuck'][i] + stats['stop_other'][i]
        assert stats['grow_steps'][i] >= stats['regions'][i] - stats['stop_noneighbor'][i]
    # graph replay and direct launches are the same computation
    from learn_region_grow_b200 import _lib
    labels2, stats2 = engine.segment_rooms(pts, orders, resolution=0.1, seed=0, flags=_lib.FLAG_NO_GRAPH)
    for a, b in zip(labels, labels2):
        np.testing.assert_array_equal(a, b)
    assert stats2['grow_steps'].tolist() == stats['grow_steps'].tolist()


def test_resolution_and_threshold_parameters(engine, golden_weights):
    """--resolution (test_region_grow.py:63) and cluster_threshold (:33) reach the device."""
    from tools import rooms as R
    f = feature_prep.prepare_features(R.generate_room(1000, n_raw=6000, n_boxes=4), 0.3)      # equalised at 0.3 m: one point per voxel
    points, order = f['points'], f['order']
    labels, stats = engine.segment_rooms([points], [order], resolution=0.3, seed=1, cluster_threshold=25, trace_capacity=2048)
    trace, n_steps = engine.trace(0, 2048)
    fwd = lambda a, b: lrg_forward.forward(golden_weights, a, b)
    raw = engine.labels(filled=False)[0]
    sizes = np.bincount(raw)[1:]
    assert len(sizes) == 0 or sizes.min() > 25
    # the first region's first step sees the shell of the seed voxel at 0.3 m
    vox = lrg_driver.voxelize(points[:, :3], 0.3)
    seed = order[0]
    shell = np.all(np.abs(vox - vox[seed]) <= 1, axis=1)
    shell[seed] = False
    assert trace[0]['seed_point'] == seed and trace[0]['n_neighbor'] == shell.sum()


def test_large_room_replays_on_oracle(engine, golden_weights):
    """A room larger than one scan chunk (N > 16,384 state words) with a floor region of several thousand inliers:
    multi-chunk scans, the >1024 and >2048 median paths and full 512-of-n sampling, replayed step by step on the oracle."""
    from tools import rooms as R
    f = feature_prep.prepare_features(R.generate_room(4242, n_raw=70000, n_boxes=6, dims=np.array([9.0, 9.0, 2.4])))
    points, order = f['points'], f['order']
    assert len(points) > 16384
    engine.upload_rooms([points], [order], resolution=0.1)
    stats = engine.segment_resident(resolution=0.1, seed=5, trace_capacity=16384)
    trace, n_steps = engine.trace(0, 16384)
    assert n_steps == stats['grow_steps'][0] and n_steps <= 16384
    assert trace['n_inlier'].max() > 2048 and trace['n_neighbor'].max() >= 512
    g, adopted = _replay(points, order, golden_weights, trace, n_steps, 5)
    assert adopted <= 8
    np.testing.assert_array_equal(engine.labels(filled=False)[0], g.cluster_label)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())


def test_outdoor_shaped_scene_at_30cm_replays_on_oracle(engine, golden_weights):
    """BASELINE config 5's shape at a size the oracle finishes: a 36 x 36 m ground plane with boxes on it at resolution 0.3
    (tools/rooms.generate_outdoor_scene) -- the ground region reaches thousands of inliers (block median, box query for the
    inlier list, full 512-of-n sampling) and every shell scan goes through the spatial index; replayed step by step."""
    from tools import rooms as R
    f = feature_prep.prepare_features(R.generate_outdoor_scene(3100, n_raw=40000, extent=36.0, n_boxes=30), 0.3)
    points, order = f['points'], f['order']
    assert len(points) > 8192
    engine.upload_rooms([points], [order], resolution=0.3)
    stats = engine.segment_resident(resolution=0.3, seed=2, trace_capacity=16384)
    trace, n_steps = engine.trace(0, 16384)
    assert n_steps == stats['grow_steps'][0] and 200 < n_steps <= 16384
    assert trace['n_inlier'].max() > 2048
    g, adopted = _replay(points, order, golden_weights, trace, n_steps, 2, resolution=0.3)
    print('outdoor scene: %d points, %d steps, %d regions, largest region %d, %d near-tie bits adopted' %
          (len(points), n_steps, len(g.regions), int(trace['n_inlier'].max()), adopted))
    assert adopted <= 12
    np.testing.assert_array_equal(engine.labels(filled=False)[0], g.cluster_label)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())
    # the same scene with eight speculative lanes (the engine's default for large scenes): the same labels
    engine.segment_resident(resolution=0.3, seed=2, spec_lanes=8)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())


@pytest.mark.parametrize('room_index', [0, 26])
def test_bench_shaped_room_replays_on_oracle(engine, golden_weights, room_index):
    """The rooms bench.py grows (tools/rooms.generate_room(1000 + i) with its defaults: ~20 k raw points, 20-40 boxes; index 26
    is the longest chain of the 68-room workload) replayed step by step on the pinned oracle -- seeds, set sizes, medians,
    sampled indices, masks, stop reasons, labels before and after the fill."""
    from tools import rooms as R
    f = feature_prep.prepare_features(R.generate_room(1000 + room_index))
    points, order = f['points'], f['order']
    engine.upload_rooms([points], [order], resolution=0.1)
    stats = engine.segment_resident(resolution=0.1, seed=0, trace_capacity=8192, room_id_base=room_index)
    trace, n_steps = engine.trace(0, 8192)
    assert n_steps == stats['grow_steps'][0] and 1000 < n_steps <= 8192
    g, adopted = _replay(points, order, golden_weights, trace, n_steps, 0, room_id=room_index)
    print('bench room %d: %d points, %d steps, %d regions, %d near-tie bits adopted' % (room_index, len(points), n_steps, len(g.regions), adopted))
    assert adopted <= 12
    np.testing.assert_array_equal(engine.labels(filled=False)[0], g.cluster_label)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())


def test_scheduling_variants_give_identical_labels(engine):
    """Lock-step loop, persistent kernel, persistent kernel with the priority ring: same computation, same labels."""
    from learn_region_grow_b200 import _lib
    from tools import rooms as R
    feats = [feature_prep.prepare_features(R.generate_room(1100 + i, n_raw=6000 + 2000 * i, n_boxes=6)) for i in range(5)]
    pts, orders = [f['points'] for f in feats], [f['order'] for f in feats]
    ref, st0 = engine.segment_rooms(pts, orders, resolution=0.1, seed=3)
    assert engine.profile()['persistent']
    for flags in (_lib.FLAG_LOCKSTEP, _lib.FLAG_PRIORITY, _lib.FLAG_NO_GRAPH):
        got, st = engine.segment_rooms(pts, orders, resolution=0.1, seed=3, flags=flags)
        assert engine.profile()['persistent'] == (flags == _lib.FLAG_PRIORITY)
        for a, b in zip(ref, got):
            np.testing.assert_array_equal(a, b)
        assert st['grow_steps'].tolist() == st0['grow_steps'].tolist()


def test_room_span_limit(engine):
    """The packed state words hold 10 bits of voxel coordinate per axis: a room spanning more is rejected at upload."""
    from learn_region_grow_b200._lib import LrgError
    pts = np.zeros((4, 13), np.float32)
    pts[:, 0] = [0.0, 50.0, 100.0, 102.5]          # 1025 voxels at 0.1 m
    with pytest.raises(LrgError):
        engine.upload_rooms([pts], [np.arange(4)], resolution=0.1)
    pts[:, 0] = [0.0, 50.0, 100.0, 102.0]          # 1020 voxels: fine
    labels, stats = engine.segment_rooms([pts], [np.arange(4)], resolution=0.1, seed=0)
    assert stats['n_points'][0] == 4 and labels[0].shape == (4,)


def test_upload_checks_the_drivers_preconditions(engine):
    """The device applies add / remove masks by point, the reference by voxel (test_region_grow.py:282-287): the same thing
    only with one point per voxel (what the reference's equalisation guarantees, :125-136).  An upload that breaks this, or
    whose seed order is not a permutation, is refused instead of silently giving other labels."""
    from learn_region_grow_b200._lib import LrgError
    f = feature_prep.prepare_features(__import__('tools.rooms', fromlist=['x']).generate_room(1300, n_raw=3000, n_boxes=3, dims=np.array([3.0, 2.5, 2.2])))
    pts, order = f['points'], f['order']
    engine.upload_rooms([pts], [order], resolution=0.1)                      # the equalised room is fine
    with pytest.raises(LrgError, match='share a voxel'):
        engine.upload_rooms([pts], [order], resolution=0.3)                  # ... but not at a coarser resolution
    dup = pts.copy()
    dup[7, :3] = dup[3, :3]
    with pytest.raises(LrgError, match='share a voxel'):
        engine.upload_rooms([pts, dup], [order, order], resolution=0.1)
    for bad in (np.where(np.arange(len(order)) == 5, order[6], order), np.where(np.arange(len(order)) == 9, len(order), order),
                np.where(np.arange(len(order)) == 9, -1, order)):
        with pytest.raises(LrgError, match='permutation'):
            engine.upload_rooms([pts], [bad.astype(np.int32)], resolution=0.1)
    with pytest.raises(LrgError, match='resolution'):
        engine.upload_rooms([pts], [order], resolution=0.1)
        engine.segment_resident(resolution=0.2)
    labels, stats = engine.segment_rooms([pts], [order], resolution=0.1, seed=0)
    assert labels[0].min() >= 1


def test_projection_servers_give_identical_labels(engine):
    """The pooled projection answered by the server CTAs (weights resident in shared memory, the default of the persistent
    kernel) and by work items that stream the weights from L2 (LRG_FLAG_NO_PROJ_SERVERS) sum in the same order: identical
    labels; so do the other A/B switches of the persistent kernel (no tile splitting, head tiles after the projection)."""
    from tools import rooms as R
    from learn_region_grow_b200 import _lib
    feats = [feature_prep.prepare_features(R.generate_room(1200 + i, n_raw=5000 + 2500 * i, n_boxes=6)) for i in range(4)]
    pts, orders = [f['points'] for f in feats], [f['order'] for f in feats]
    noserv = _lib.FLAG_NO_PROJ_SERVERS
    ref, st0 = engine.segment_rooms(pts, orders, resolution=0.1, seed=11, flags=noserv, spec_lanes=1)
    assert engine.profile()['persistent'] and engine.profile()['items']['gproj'] == 8 * st0['grow_steps'].sum()
    for kw in ({}, dict(num_restarts=3), dict(beam_width=2, search_width=2)):
        a, sa = engine.segment_rooms(pts, orders, resolution=0.1, seed=11, flags=noserv, **kw)
        b, sb = engine.segment_rooms(pts, orders, resolution=0.1, seed=11, **kw)
        assert engine.profile()['persistent'] and engine.profile()['items']['gproj'] == 0
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
        assert sa['grow_steps'].tolist() == sb['grow_steps'].tolist()
    for flags in (_lib.FLAG_NO_TILE_SPLIT, noserv | _lib.FLAG_HEADS_AFTER_PROJ, noserv | _lib.FLAG_NO_TILE_SPLIT | _lib.FLAG_HEADS_AFTER_PROJ):
        c, sc = engine.segment_rooms(pts, orders, resolution=0.1, seed=11, flags=flags)
        assert engine.profile()['persistent']
        for x, y in zip(ref, c):
            np.testing.assert_array_equal(x, y)
        assert sc['grow_steps'].tolist() == st0['grow_steps'].tolist()
