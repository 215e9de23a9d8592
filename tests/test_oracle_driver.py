"""Pins oracle/lrg_driver.py against the UNMODIFIED reference driver (tests/golden/driver_trace_*.npz were produced by
running /root/reference/test_region_grow.py under import shims, oracle/make_golden.py)."""
import os
import zlib

import numpy as np
import pytest

from oracle import lrg_driver, lrg_forward
from conftest import GOLDEN


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


@pytest.mark.parametrize('seed', [1000, 1001])
def test_driver_replays_reference_trace(seed, golden_weights):
    g = np.load(os.path.join(GOLDEN, 'driver_trace_%d.npz' % seed))
    calls = []

    def fwd(inlier, neighbor):
        calls.append((_crc(inlier), _crc(neighbor)))
        add, rmv = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
        return add.astype(np.float32), rmv.astype(np.float32)

    grower = lrg_driver.RoomGrower(g['points'], g['order'], fwd, lrg_driver.NumpyLegacyRng(0), resolution=0.1)
    grower.run()
    assert len(calls) == len(g['inlier_crc'])
    assert [c[0] for c in calls] == [int(x) for x in g['inlier_crc']]        # every tile fed to Session.run, bit for bit
    assert [c[1] for c in calls] == [int(x) for x in g['neighbor_crc']]
    np.testing.assert_array_equal(grower.fill(), g['cluster_label'])          # final labels after the NN fill
    # region lines printed by the reference (room R target T CLS: step S n/m points ...) agree with the oracle's record
    lines = [l for l in str(g['log']).split('\n') if l.startswith('room ')]
    labelled = [r for r in grower.regions if r[4]]
    assert len(lines) == len(labelled)
    for line, (seed_id, steps, size, reason, _) in zip(lines, labelled):
        tok = line.split()
        assert int(tok[6]) == steps and int(tok[7].split('/')[0]) == size and tok[-1] == reason


def test_literal_update_loop_equals_vectorised(golden_weights):
    """The per-point python loop of test_region_grow.py:282-287 (used for the CPU timing) and its vectorised form."""
    g = np.load(os.path.join(GOLDEN, 'driver_trace_1000.npz'))
    fwd = lambda a, b: lrg_forward.forward(golden_weights, a, b, dtype=np.float64)
    out = []
    for literal in (False, True):
        gr = lrg_driver.RoomGrower(g['points'], g['order'], fwd, lrg_driver.PhiloxRng(4), literal_update=literal)
        gr.run()
        out.append((gr.cluster_label.copy(), gr.total_steps, list(gr.regions)))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1:] == out[1][1:]


def test_restart_driver_replays_reference_trace(golden_weights):
    """oracle RestartRoomGrower vs the UNMODIFIED /root/reference/test_random_restart.py (tests/golden/restart_trace_1000.npz):
    every tile fed to Session.run over all 10 restarts of every seed, the final labels and the printed region lines."""
    base = np.load(os.path.join(GOLDEN, 'driver_trace_1000.npz'))
    g = np.load(os.path.join(GOLDEN, 'restart_trace_1000.npz'))
    calls = []

    def fwd(inlier, neighbor):
        calls.append((_crc(inlier), _crc(neighbor)))
        add, rmv = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
        return add.astype(np.float32), rmv.astype(np.float32)

    grower = lrg_driver.RestartRoomGrower(base['points'], base['order'], fwd, lrg_driver.NumpyLegacyRng(0), resolution=0.1,
                                          num_restarts=int(g['num_restarts']))
    grower.run()
    assert len(calls) == len(g['inlier_crc'])
    assert [c[0] for c in calls] == [int(x) for x in g['inlier_crc']]
    assert [c[1] for c in calls] == [int(x) for x in g['neighbor_crc']]
    np.testing.assert_array_equal(grower.fill(), g['cluster_label'])
    lines = [l for l in str(g['log']).split('\n') if l.startswith('room ')]
    labelled = [r for r in grower.regions if r[4]]
    assert len(lines) == len(labelled)
    for line, (seed_id, steps, size, reason, _) in zip(lines, labelled):
        tok = line.split()
        assert int(tok[6]) == steps and int(tok[7].split('/')[0]) == size and tok[-1] == reason


def test_beam_driver_replays_reference_trace(golden_weights):
    """oracle BeamRoomGrower vs the UNMODIFIED /root/reference/test_beam_search.py (tests/golden/beam_trace_1000.npz, run
    with Python 2's list-returning ``range`` in the script's globals, oracle/make_golden.py): every tile fed to Session.run
    (BEAM_WIDTH x SEARCH_WIDTH expansions per round), the final labels and the printed region lines."""
    base = np.load(os.path.join(GOLDEN, 'driver_trace_1000.npz'))
    g = np.load(os.path.join(GOLDEN, 'beam_trace_1000.npz'))
    calls = []

    def fwd(inlier, neighbor):
        calls.append((_crc(inlier), _crc(neighbor)))
        add, rmv = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
        return add.astype(np.float32), rmv.astype(np.float32)

    grower = lrg_driver.BeamRoomGrower(base['points'], base['order'], fwd, lrg_driver.NumpyLegacyRng(0), resolution=0.1,
                                       beam_width=int(g['beam_width']), search_width=int(g['search_width']))
    grower.run()
    assert len(calls) == len(g['inlier_crc'])
    assert [c[0] for c in calls] == [int(x) for x in g['inlier_crc']]
    assert [c[1] for c in calls] == [int(x) for x in g['neighbor_crc']]
    np.testing.assert_array_equal(grower.fill(), g['cluster_label'])
    # room R target T CLS: step S n/m points IOU ...   (test_beam_search.py:281)
    lines = [l for l in str(g['log']).split('\n') if l.startswith('room ')]
    labelled = [r for r in grower.regions if r[4]]
    assert len(lines) == len(labelled)
    for line, (seed_id, steps, size, reason, _) in zip(lines, labelled):
        tok = line.split()
        assert int(tok[6]) == steps and int(tok[7].split('/')[0]) == size


def test_beam_ml_scoring_replays_reference_trace(golden_weights):
    """'ml' scoring (``--scoring ml``, test_beam_search.py:46-47,238-256,263-264): oracle BeamRoomGrower(scoring='ml') vs the
    UNMODIFIED script (tests/golden/beam_ml_trace_1000.npz, oracle/make_golden.py beam_ml).  The ranking by accumulated
    log-probability decides which masks survive a round, so every tile of the 2,781 Session.run calls, the final labels and
    the printed region lines only agree when the float32 scores order the candidates like the script's own."""
    base = np.load(os.path.join(GOLDEN, 'driver_trace_1000.npz'))
    g = np.load(os.path.join(GOLDEN, 'beam_ml_trace_1000.npz'))
    np_run = np.load(os.path.join(GOLDEN, 'beam_trace_1000.npz'))
    assert len(g['inlier_crc']) != len(np_run['inlier_crc'])              # the two scorings really search differently
    calls = []

    def fwd(inlier, neighbor):
        calls.append((_crc(inlier), _crc(neighbor)))
        add, rmv = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
        return add.astype(np.float32), rmv.astype(np.float32)

    grower = lrg_driver.BeamRoomGrower(base['points'], base['order'], fwd, lrg_driver.NumpyLegacyRng(0), resolution=0.1,
                                       beam_width=int(g['beam_width']), search_width=int(g['search_width']), scoring='ml')
    grower.run()
    assert len(calls) == len(g['inlier_crc'])
    assert [c[0] for c in calls] == [int(x) for x in g['inlier_crc']]
    assert [c[1] for c in calls] == [int(x) for x in g['neighbor_crc']]
    np.testing.assert_array_equal(grower.fill(), g['cluster_label'])
    lines = [l for l in str(g['log']).split('\n') if l.startswith('room ')]
    labelled = [r for r in grower.regions if r[4]]
    assert len(lines) == len(labelled)
    for line, (seed_id, steps, size, reason, _) in zip(lines, labelled):
        tok = line.split()
        assert int(tok[6]) == steps and int(tok[7].split('/')[0]) == size


def test_beam_driver_philox_lanes_and_forced_replay(golden_weights):
    """The Philox form of the beam-search oracle (expansion (q, s) of round r draws at lane q * SEARCH_WIDTH + s, step r): a
    run is deterministic, and re-driving it with the masks it sampled (the hook tests/test_beam_gpu.py feeds with the device's
    masks) reproduces it exactly -- the replay harness itself is sound."""
    base = np.load(os.path.join(GOLDEN, 'driver_trace_1000.npz'))
    fwd = lambda a, b: lrg_forward.forward(golden_weights, a, b)
    runs = []
    for _ in range(2):
        g = lrg_driver.BeamRoomGrower(base['points'], base['order'], fwd, lrg_driver.PhiloxRng(9), beam_width=2, search_width=2)
        g.trace = []
        g.run()
        runs.append(g)
    a, b = runs
    assert np.array_equal(a.cluster_label, b.cluster_label) and a.lane_log == b.lane_log and a.regions == b.regions
    assert a.total_steps == len(a.lane_log) == sum(a.lane_steps) and max(x[2] for x in a.lane_log) <= 3
    assert all(r[3] in ('stuck', 'exhausted') for r in a.regions) and a.visited.all()
    # lanes of one round share the parent: the candidates of a round differ only through their lane's streams
    masks = iter(a.trace)

    def forced(st, lane):
        rec = next(masks)
        assert rec['n_inlier'] == st['n_inlier'] and np.array_equal(rec['inlier_idx'], st['inlier_idx'])
        add, rmv = fwd(st['inlier'], st['neighbor'])
        return add, rmv, rec['add_mask'], rec['rmv_mask']

    c = lrg_driver.BeamRoomGrower(base['points'], base['order'], fwd, lrg_driver.PhiloxRng(9), beam_width=2, search_width=2)
    c.run(forced)
    assert next(masks, None) is None
    assert np.array_equal(c.cluster_label, a.cluster_label) and c.lane_log == a.lane_log
    other = lrg_driver.BeamRoomGrower(base['points'], base['order'], fwd, lrg_driver.PhiloxRng(10), beam_width=2, search_width=2)
    other.run()
    assert other.lane_log != a.lane_log                                      # the seed matters


@pytest.mark.parametrize('lanes,validate', [(1, 'early'), (3, 'early'), (3, 'commit')])
def test_speculative_lanes_keep_the_plain_result(lanes, validate, golden_weights):
    """Design study for intra-room parallelism (oracle.lrg_driver.SpeculativeRoomGrower): regions of one room grown
    side by side, committed in seed order, a lane started over when a commit lands inside the envelope it has looked at --
    labels and region records are exactly the plain driver's, also when a region is only validated at its own commit."""
    base = np.load(os.path.join(GOLDEN, 'driver_trace_1000.npz'))
    fwd = lambda a, b: lrg_forward.forward(golden_weights, a, b)
    ref = lrg_driver.RoomGrower(base['points'], base['order'], fwd, lrg_driver.PhiloxRng(4))
    ref.run()
    s = lrg_driver.SpeculativeRoomGrower(base['points'], base['order'], fwd, seed=4, lanes=lanes, validate=validate)
    np.testing.assert_array_equal(s.run(), ref.cluster_label)
    assert s.regions == ref.regions and s.useful == ref.total_steps
    if lanes == 1:
        assert s.ticks == ref.total_steps and s.wasted == 0
    else:
        assert s.ticks < ref.total_steps
