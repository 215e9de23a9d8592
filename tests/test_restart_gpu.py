"""Random-restart driver on the device (/root/reference/test_random_restart.py) vs the oracle restatement
(oracle/lrg_driver.py RestartRoomGrower, itself pinned bit for bit to the unmodified reference script by
tests/test_oracle_driver.py), through the C ABI (LrgGrowParams.num_restarts).

The restarts of a seed run side by side on the device (one lane each, own Philox streams); the oracle grows them one
after the other with the same streams and is re-driven with the device's per-lane traces like tests/test_driver_gpu.py."""
import numpy as np
import pytest

from oracle import feature_prep

from learn_region_grow_b200 import _lib
from oracle import lrg_driver, lrg_forward
from util_rooms import golden_room, idx_crc, unpack_mask

pytestmark = pytest.mark.gpu

NEAR_TIE = 2e-4
STOP_NAMES = {2: 'noexpand', 3: 'stuck', 4: 'maxsteps', 5: 'empty'}


@pytest.fixture(scope='module')
def engine(golden_weights):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    yield e
    e.close()


def _replay(points, order, weights, traces, seed, R):
    fwd = lambda a, b: lrg_forward.forward(weights, a, b)
    g = lrg_driver.RestartRoomGrower(points, order, fwd, lrg_driver.PhiloxRng(seed), num_restarts=R)
    pos = [0] * R
    adopted = 0
    for seed_id in np.arange(len(points))[order]:
        if g.visited[seed_id]:
            continue
        g.seed_steps = 0
        for lane in range(R):
            g.lane = lane
            g.begin_region(seed_id)
            while True:
                st = g.prepare_step()
                if st is None:
                    break
                trace, n_steps = traces[lane]
                assert pos[lane] < n_steps, 'oracle wants more steps than lane %d ran' % lane
                rec = trace[pos[lane]]
                assert rec['seed_point'] == seed_id and rec['step_in_region'] == g.steps
                assert rec['n_inlier'] == st['n_inlier'] and rec['n_neighbor'] == st['n_neighbor']
                assert rec['inlier_idx_crc'] == idx_crc(st['inlier_idx']) and rec['neighbor_idx_crc'] == idx_crc(st['neighbor_idx'])
                add, rmv = fwd(st['inlier'], st['neighbor'])
                dev_add, dev_rmv = unpack_mask(rec['add_mask']), unpack_mask(rec['remove_mask'])
                add_conf, rmv_conf = lrg_driver.confidence(add[0]), lrg_driver.confidence(rmv[0])
                rng = lrg_driver.PhiloxRng(seed)
                rng.begin_step(0, g.steps, lane, g.seed_id)
                u_add, u_rmv = rng.uniform(512, 'add'), rng.uniform(512, 'remove')
                for dev, conf, u in ((dev_add, add_conf, u_add), (dev_rmv, rmv_conf, u_rmv)):
                    differ = dev != (u < conf)
                    assert np.all(np.abs(u[differ] - conf[differ]) < NEAR_TIE), 'mask bit differs outside the near-tie band'
                    adopted += int(differ.sum())
                reason = g.apply_step(add[0], rmv[0], add_mask=dev_add, rmv_mask=dev_rmv)
                assert STOP_NAMES.get(int(rec['stop_reason'])) == reason
                pos[lane] += 1
                if reason is not None:
                    break
    for lane in range(R):
        assert pos[lane] == traces[lane][1]
    return g, adopted


@pytest.mark.parametrize('R,flags', [(3, 0), (10, 0), (4, _lib.FLAG_LOCKSTEP)])
def test_restart_driver_replays_on_oracle(engine, golden_weights, R, flags):
    points, order = golden_room(1000)
    engine.upload_rooms([points], [order], resolution=0.1)
    stats = engine.segment_resident(resolution=0.1, seed=3, trace_capacity=2048, num_restarts=R, flags=flags)
    traces = [engine.trace(0, 2048, lane=l) for l in range(R)]
    assert sum(t[1] for t in traces) == stats['grow_steps'][0]
    g, adopted = _replay(points, order, golden_weights, traces, 3, R)
    assert adopted <= 3 * R
    np.testing.assert_array_equal(engine.labels(filled=False)[0], g.cluster_label)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())
    assert stats['regions'][0] == len(g.regions) and stats['clusters'][0] == g.cluster_id - 1
    by_reason = {r: sum(1 for x in g.lane_regions if x[3] == r) for r in ('noneighbor', 'noexpand', 'stuck')}
    assert (stats['stop_noneighbor'][0], stats['stop_noexpand'][0], stats['stop_stuck'][0]) == \
        (by_reason['noneighbor'], by_reason['noexpand'], by_reason['stuck'])


def test_restart_scheduling_invariance(engine):
    """Rooms stay independent units: any number of groups, the persistent kernel or the lock-step loop, rooms alone or
    together -- same labels.  num_restarts <= 1 is the plain driver."""
    from tools import rooms as Rm
    feats = [feature_prep.prepare_features(Rm.generate_room(1000 + i, n_raw=2500 + 1000 * i, n_boxes=4)) for i in range(3)]
    pts = [f['points'] for f in feats] + [np.zeros((0, 13), np.float32)]
    orders = [f['order'] for f in feats] + [np.zeros(0, np.int64)]
    ref, st = engine.segment_rooms(pts, orders, resolution=0.1, seed=5, num_restarts=5)
    assert st['n_points'].tolist() == [len(p) for p in pts]
    plain, st1 = engine.segment_rooms(pts, orders, resolution=0.1, seed=5)
    assert st['grow_steps'].sum() > 3 * st1['grow_steps'].sum()           # five restarts per seed
    for kw in (dict(max_slots=5), dict(max_slots=10), dict(flags=_lib.FLAG_LOCKSTEP), dict(flags=_lib.FLAG_LOCKSTEP | _lib.FLAG_NO_GRAPH, max_slots=5)):
        again, st2 = engine.segment_rooms(pts, orders, resolution=0.1, seed=5, num_restarts=5, **kw)
        for a, b in zip(again, ref):
            np.testing.assert_array_equal(a, b)
        assert st2['grow_steps'].tolist() == st['grow_steps'].tolist()
    for i in range(3):
        alone, _ = engine.segment_rooms([pts[i]], [orders[i]], resolution=0.1, seed=5, num_restarts=5, room_id_base=i)
        np.testing.assert_array_equal(alone[0], ref[i])
    one, _ = engine.segment_rooms(pts, orders, resolution=0.1, seed=5, num_restarts=1)
    for a, b in zip(one, plain):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(_lib.LrgError):
        engine.segment_rooms(pts, orders, resolution=0.1, seed=5, num_restarts=17)
    # restarts keep the largest region: labelled clusters are on average larger than with the plain driver
    assert all(l.min() >= 1 for l in ref[:3])
