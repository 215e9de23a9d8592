"""Beam-search driver on the device (/root/reference/test_beam_search.py) vs the oracle restatement
(oracle/lrg_driver.py BeamRoomGrower, itself pinned bit for bit to the unmodified reference script by
tests/test_oracle_driver.py), through the C ABI (LrgGrowParams.beam_width / search_width).

The BEAM_WIDTH x SEARCH_WIDTH expansions of a round run side by side on the device (one lane each, own Philox streams);
the oracle runs them one after the other with the same streams and is re-driven with the device's per-lane traces like
tests/test_driver_gpu.py: a sampled mask bit may differ from the oracle's only inside the near-tie band of the forward
tolerance, and such bits are adopted from the device (and counted)."""
import numpy as np
import pytest

from oracle import feature_prep

from learn_region_grow_b200 import _lib
from oracle import lrg_driver, lrg_forward
from util_rooms import golden_room, idx_crc, unpack_mask

pytestmark = pytest.mark.gpu

NEAR_TIE = 2e-4


@pytest.fixture(scope='module')
def engine(golden_weights):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    yield e
    e.close()


LOG_PROB_TOL = 4e-4   # |d log p / d logit| <= 1 per row and the sums are means over the rows: twice the 2e-4 logit tolerance


def _replay(points, order, weights, traces, seed, B, W, scoring='np', score_err=None):
    fwd = lambda a, b: lrg_forward.forward(weights, a, b)
    g = lrg_driver.BeamRoomGrower(points, order, fwd, lrg_driver.PhiloxRng(seed), beam_width=B, search_width=W, scoring=scoring)
    pos = [0] * (B * W)
    adopted = [0]

    def forced(st, lane):
        trace, n_steps = traces[lane]
        assert pos[lane] < n_steps, 'oracle wants more expansions than lane %d ran' % lane
        rec = trace[pos[lane]]
        pos[lane] += 1
        assert rec['seed_point'] == g.seed_id and rec['step_in_region'] == g.round
        assert rec['n_inlier'] == st['n_inlier'] and rec['n_neighbor'] == st['n_neighbor']
        np.testing.assert_array_equal(rec['center'][:13][[0, 1, 6, 7, 8, 9, 10, 11, 12]], st['center'][[0, 1, 6, 7, 8, 9, 10, 11, 12]])
        assert rec['inlier_idx_crc'] == idx_crc(st['inlier_idx']) and rec['neighbor_idx_crc'] == idx_crc(st['neighbor_idx'])
        add, rmv = fwd(st['inlier'], st['neighbor'])
        dev_add, dev_rmv = unpack_mask(rec['add_mask']), unpack_mask(rec['remove_mask'])
        add_conf, rmv_conf = lrg_driver.confidence(add[0]), lrg_driver.confidence(rmv[0])
        rng = lrg_driver.PhiloxRng(seed)
        rng.begin_step(0, g.round, lane, g.seed_id)
        u_add, u_rmv = rng.uniform(512, 'add'), rng.uniform(512, 'remove')
        for dev, conf, u in ((dev_add, add_conf, u_add), (dev_rmv, rmv_conf, u_rmv)):
            differ = dev != (u < conf)
            assert np.all(np.abs(u[differ] - conf[differ]) < NEAR_TIE), 'mask bit differs outside the near-tie band'
            adopted[0] += int(differ.sum())
        return add, rmv, dev_add, dev_rmv

    def adopt_score(lane, parent, parts, new_score):
        # 'ml': the device's log-probabilities must be the oracle's within the forward tolerance; its score must be EXACTLY
        # float32(parent + add + rmv) of its own parts (:264); the oracle then ranks with the device's number, so that a
        # near-tie between two candidates cannot send the two searches down different paths
        rec = traces[lane][0][pos[lane] - 1]
        lp = rec['log_prob'].astype(np.float32)
        for mine, theirs in zip(parts, lp):
            if np.isfinite(mine) or np.isfinite(theirs):
                assert abs(float(mine) - float(theirs)) <= LOG_PROB_TOL, (lane, parts, lp)
                score_err.append(abs(float(mine) - float(theirs)))
        assert np.float32(rec['score']) == np.float32(np.float32(np.float32(parent) + lp[0]) + lp[1]), (lane, parent, lp, rec['score'])
        return np.float32(rec['score'])

    if scoring == 'ml':
        forced.score = adopt_score
    g.run(forced)
    for lane in range(B * W):
        trace, n_steps = traces[lane]
        assert pos[lane] == n_steps
        mine = [x for x in g.lane_log if x[2] == lane]
        assert len(mine) == n_steps
        for rec, (seed_id, rnd, _, updated, size) in zip(trace, mine):
            assert (rec['seed_point'], rec['step_in_region'], rec['size_after']) == (seed_id, rnd, size)
            assert (int(rec['stop_reason']) == 0) == (updated and size > 0)
    return g, adopted[0]


@pytest.mark.parametrize('B,W,flags', [(3, 3, 0), (1, 1, 0), (2, 2, _lib.FLAG_LOCKSTEP)])
def test_beam_driver_replays_on_oracle(engine, golden_weights, B, W, flags):
    points, order = golden_room(1000)
    engine.upload_rooms([points], [order], resolution=0.1)
    stats = engine.segment_resident(resolution=0.1, seed=3, trace_capacity=2048, beam_width=B, search_width=W, flags=flags)
    assert engine.profile()['persistent'] == (flags == 0)
    traces = [engine.trace(0, 2048, lane=l) for l in range(B * W)]
    assert sum(t[1] for t in traces) == stats['grow_steps'][0] and stats['grow_steps'][0] > 50
    g, adopted = _replay(points, order, golden_weights, traces, 3, B, W)
    assert adopted <= 3 * B * W
    np.testing.assert_array_equal(engine.labels(filled=False)[0], g.cluster_label)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())
    assert stats['regions'][0] == len(g.regions) and stats['clusters'][0] == g.cluster_id - 1
    by_reason = {r: sum(1 for x in g.regions if x[3] == r) for r in ('exhausted', 'stuck')}
    assert (stats['stop_noexpand'][0], stats['stop_stuck'][0]) == (by_reason['exhausted'], by_reason['stuck'])


@pytest.mark.parametrize('B,W,flags', [(3, 3, 0), (2, 3, _lib.FLAG_LOCKSTEP)])
def test_beam_ml_scoring_replays_on_oracle(engine, golden_weights, B, W, flags):
    """``--scoring ml`` (test_beam_search.py:46-47,238-256,263-264; LRG_FLAG_SCORE_ML): candidates ranked by accumulated
    log-probability.  The oracle restatement of it is pinned bit for bit to the unmodified script
    (tests/test_oracle_driver.py::test_beam_ml_scoring_replays_reference_trace); here the device's per-expansion
    log-probabilities are held to it within LOG_PROB_TOL, its score arithmetic exactly, and -- ranking with the device's scores --
    the labels must be identical."""
    points, order = golden_room(1000)
    engine.upload_rooms([points], [order], resolution=0.1)
    stats = engine.segment_resident(resolution=0.1, seed=3, trace_capacity=2048, beam_width=B, search_width=W, flags=flags, scoring='ml')
    traces = [engine.trace(0, 2048, lane=l) for l in range(B * W)]
    assert sum(t[1] for t in traces) == stats['grow_steps'][0] and stats['grow_steps'][0] > 50
    errs = []
    g, adopted = _replay(points, order, golden_weights, traces, 3, B, W, scoring='ml', score_err=errs)
    assert adopted <= 3 * B * W and len(errs) > 100
    print('ml scoring: %d expansions scored, max |log-prob difference| %.2e, mean %.2e' % (len(errs) // 2, max(errs), np.mean(errs)))
    np.testing.assert_array_equal(engine.labels(filled=False)[0], g.cluster_label)
    np.testing.assert_array_equal(engine.labels(filled=True)[0], g.fill())
    assert stats['regions'][0] == len(g.regions) and stats['clusters'][0] == g.cluster_id - 1
    # the two scorings search differently (same seed, same streams)
    engine.segment_resident(resolution=0.1, seed=3, beam_width=B, search_width=W, flags=flags)
    assert not np.array_equal(engine.labels(filled=False)[0], g.cluster_label)
    with pytest.raises(_lib.LrgError):
        engine.segment_resident(resolution=0.1, seed=3, scoring='ml')                      # needs the beam-search driver
    with pytest.raises(_lib.LrgError):
        engine.segment_resident(resolution=0.1, seed=3, num_restarts=3, scoring='ml')      # broken in the reference (:196)


def test_beam_scheduling_invariance(engine):
    """Rooms stay independent units: any number of groups, the persistent kernel or the lock-step loop, rooms alone or
    together -- same labels."""
    from tools import rooms as Rm
    feats = [feature_prep.prepare_features(Rm.generate_room(1000 + i, n_raw=2500 + 1000 * i, n_boxes=4)) for i in range(3)]
    pts = [f['points'] for f in feats] + [np.zeros((0, 13), np.float32)]
    orders = [f['order'] for f in feats] + [np.zeros(0, np.int64)]
    ref, st = engine.segment_rooms(pts, orders, resolution=0.1, seed=5, beam_width=3, search_width=3)
    assert st['n_points'].tolist() == [len(p) for p in pts]
    plain, st1 = engine.segment_rooms(pts, orders, resolution=0.1, seed=5)
    assert st['grow_steps'].sum() > 2 * st1['grow_steps'].sum()           # up to nine expansions per round
    for kw in (dict(max_slots=9), dict(max_slots=18), dict(flags=_lib.FLAG_LOCKSTEP), dict(flags=_lib.FLAG_LOCKSTEP | _lib.FLAG_NO_GRAPH, max_slots=9)):
        again, st2 = engine.segment_rooms(pts, orders, resolution=0.1, seed=5, beam_width=3, search_width=3, **kw)
        for a, b in zip(again, ref):
            np.testing.assert_array_equal(a, b)
        assert st2['grow_steps'].tolist() == st['grow_steps'].tolist()
    for i in range(3):
        alone, _ = engine.segment_rooms([pts[i]], [orders[i]], resolution=0.1, seed=5, beam_width=3, search_width=3, room_id_base=i)
        np.testing.assert_array_equal(alone[0], ref[i])
    assert all(l.min() >= 1 for l in ref[:3])
    for bad in (dict(beam_width=3), dict(beam_width=5, search_width=4), dict(beam_width=3, search_width=3, num_restarts=2)):
        with pytest.raises(_lib.LrgError):
            engine.segment_rooms(pts, orders, resolution=0.1, seed=5, **bad)
    # a plain run afterwards is unaffected by the lane state of the beam runs
    one, _ = engine.segment_rooms(pts, orders, resolution=0.1, seed=5)
    for a, b in zip(one, plain):
        np.testing.assert_array_equal(a, b)
