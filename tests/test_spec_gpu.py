"""Speculative lanes (LrgGrowParams.spec_lanes > 1): several regions of ONE room grown side by side with commits strictly
in seed order (test_region_grow.py:183-217).  The result must be BIT-IDENTICAL to the plain driver (spec_lanes = 1), which
tests/test_driver_gpu.py replays step by step on the pinned oracle: labels before and after the fill, grow steps, regions,
clusters and stop reasons per room -- in the persistent kernel and in the lock-step loop, alone or many rooms together."""
import numpy as np
import pytest

from oracle import feature_prep, lrg_driver, lrg_forward
from util_rooms import golden_room

pytestmark = pytest.mark.gpu
SAME = ('n_points', 'grow_steps', 'regions', 'clusters', 'stop_noneighbor', 'stop_noexpand', 'stop_stuck', 'stop_other')


@pytest.fixture(scope='module')
def engine(golden_weights):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    yield e
    e.close()


def _run(engine, pts, orders, **kw):
    labels, stats = engine.segment_rooms(pts, orders, resolution=0.1, **kw)
    return [l.copy() for l in labels], [l.copy() for l in engine.labels(filled=False)], stats.copy()


def _same(a, b):
    for x, y in zip(a[0], b[0]):
        np.testing.assert_array_equal(x, y)
    for x, y in zip(a[1], b[1]):
        np.testing.assert_array_equal(x, y)
    for k in SAME:
        assert a[2][k].tolist() == b[2][k].tolist(), k


@pytest.mark.parametrize('lanes', [2, 3, 4, 8])
def test_golden_room_equals_oracle_and_plain_driver(engine, golden_weights, lanes):
    points, order = golden_room(1001)
    plain = _run(engine, [points], [order], seed=12345, spec_lanes=1)
    default = _run(engine, [points], [order], seed=12345)          # (engine default: 4 lanes in the persistent kernel)
    _same(plain, default)
    spec = _run(engine, [points], [order], seed=12345, spec_lanes=lanes)
    _same(plain, spec)
    assert plain[2]['spec_wasted_steps'][0] == 0
    g = lrg_driver.RoomGrower(points, order, lambda a, b: lrg_forward.forward(golden_weights, a, b), lrg_driver.PhiloxRng(12345))
    g.run()
    # (the oracle's masks may differ from the device's in near-tie bits; the step-by-step replay of the plain driver bounds
    # that -- here only the cheap check that both agree on nearly every point)
    assert np.mean(g.cluster_label == spec[1][0]) > 0.95
    st = spec[2][0]
    print('golden room 1001, %d lanes: %d committed steps, %d discarded (%d regrown, %d dropped)'
          % (lanes, st['grow_steps'], st['spec_wasted_steps'], st['spec_restarts'], st['spec_dropped']))


def test_many_rooms_and_scheduling_variants(engine):
    from learn_region_grow_b200 import _lib
    from tools import rooms as R
    feats = [feature_prep.prepare_features(R.generate_room(1500 + i, n_raw=4000 + 2500 * i, n_boxes=4 + i)) for i in range(5)]
    pts = [f['points'] for f in feats] + [np.zeros((0, 13), np.float32), feats[0]['points'][:7]]
    orders = [f['order'] for f in feats] + [np.zeros(0, np.int64), np.arange(7)]
    plain = _run(engine, pts, orders, seed=3, spec_lanes=1)
    for kw in (dict(spec_lanes=4), dict(spec_lanes=2, max_slots=4), dict(spec_lanes=4, flags=_lib.FLAG_LOCKSTEP), dict(spec_lanes=3, max_slots=3)):
        spec = _run(engine, pts, orders, seed=3, **kw)
        _same(plain, spec)
        assert engine.profile()['persistent'] == ('flags' not in kw)
    # deterministic in the labels although the schedule (and with it what gets discarded) is not
    a = _run(engine, pts, orders, seed=3, spec_lanes=4)
    b = _run(engine, pts, orders, seed=3, spec_lanes=4)
    _same(a, b)


def test_bench_shaped_room(engine):
    from tools import rooms as R
    f = feature_prep.prepare_features(R.generate_room(1026))
    plain = _run(engine, [f['points']], [f['order']], seed=0, room_id_base=26)
    t_plain = engine.profile()['grow_ms']
    for lanes in (2, 4):
        spec = _run(engine, [f['points']], [f['order']], seed=0, room_id_base=26, spec_lanes=lanes)
        t_spec = engine.profile()['grow_ms']
        _same(plain, spec)
        st = spec[2][0]
        print('bench room 26, %d lanes: %.1f ms vs %.1f ms plain; %d committed steps, %d discarded (%d regrown, %d dropped)'
              % (lanes, t_spec, t_plain, st['grow_steps'], st['spec_wasted_steps'], st['spec_restarts'], st['spec_dropped']))


def test_spec_lanes_exclude_the_local_search_drivers(engine):
    from learn_region_grow_b200._lib import LrgError
    points, order = golden_room(1000)
    with pytest.raises(LrgError):
        engine.segment_rooms([points], [order], resolution=0.1, spec_lanes=2, num_restarts=3)
    with pytest.raises(LrgError):
        engine.segment_rooms([points], [order], resolution=0.1, spec_lanes=2, beam_width=2, search_width=2)
    with pytest.raises(LrgError):
        engine.segment_rooms([points], [order], resolution=0.1, spec_lanes=17)
