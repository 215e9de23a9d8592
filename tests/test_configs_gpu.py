"""BASELINE.json configs 3 and 5 as parity-test cases: variable-N ScanNet-shaped rooms and a Semantic-KITTI-shaped
scene at 0.3 m, through the raw-points C ABI.  At these sizes the oracle driver does not finish in seconds, so the checks are
the size-independent ones: the equalisation maps and exact feature columns against the host restatement of
test_region_grow.py:119-173, label invariants, cluster_label[unequalized_idx], determinism, independence of the slot count,
and the statistics against the scikit-learn oracle on the engine's own labels."""
import numpy as np
import pytest

from oracle import feature_prep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine(golden_weights):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    yield e
    e.close()


def _check_rooms(engine, raws, resolution, seed):
    from tools import rooms
    from oracle import metrics as om
    labels_raw, stats = engine.segment_raw_rooms(raws, resolution=resolution, seed=seed)
    f = engine.prepared_features()
    eq_off, raw_off = engine._room_offsets, engine._raw_offsets
    eq_labels, unfilled = engine.labels(True), engine.labels(False)
    obj = [r[:, 6].astype(np.int32) for r in raws]
    m = engine.room_metrics(obj, raw=True)
    for i, raw in enumerate(raws):
        host = feature_prep.prepare_features(raw, resolution)
        e0, e1 = eq_off[i], eq_off[i + 1]
        assert e1 - e0 == len(host['points']) == stats['n_points'][i]
        np.testing.assert_array_equal(f['equalized_idx'][e0:e1], host['equalized_idx'])
        np.testing.assert_array_equal(f['unequalized_idx'][raw_off[i]:raw_off[i + 1]], host['unequalized_idx'])
        np.testing.assert_array_equal(f['points'][e0:e1, :9], host['points'][:, :9])
        assert np.isclose(f['points'][e0:e1, 9:], host['points'][:, 9:], atol=1e-5).all(axis=1).mean() > 0.99
        lab = eq_labels[i]
        assert lab.min() >= 1 and lab.max() <= stats['clusters'][i]
        assert np.array_equal(lab[unfilled[i] > 0], unfilled[i][unfilled[i] > 0])              # the fill never relabels
        sizes = np.bincount(unfilled[i])[1:]
        assert sizes.min() > 10                                                                 # cluster_threshold (:33,213)
        np.testing.assert_array_equal(labels_raw[i], lab[host['unequalized_idx']])              # :366
        o = om.room_statistics(host['obj_id'], lab)
        for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou'):
            assert abs(m[i][k] - o[k]) <= 1e-9, (i, k)
    return labels_raw, stats


def test_scannet_shaped_variable_rooms(engine):
    """config 3: raw sizes drawn log-uniformly in [5 k, 60 k] (README.md:48-49), seeds 2000 + room."""
    from tools import rooms
    raws = rooms.generate_area(5, seed_base=2000, log_uniform=(5000, 60000))
    assert max(map(len, raws)) > 3 * min(map(len, raws))
    labels, stats = _check_rooms(engine, raws, 0.1, seed=0)
    again, stats2 = engine.segment_raw_rooms(raws, resolution=0.1, seed=0, max_slots=2)          # fewer slots than rooms
    for a, b in zip(again, labels):
        np.testing.assert_array_equal(a, b)
    assert stats2['grow_steps'].tolist() == stats['grow_steps'].tolist()


def test_kitti_shaped_scene_at_30cm(engine):
    """config 5: one outdoor scene (ground plane + vehicle / pole / building sized boxes), resolution 0.3
    (test_region_grow.py --resolution 0.3): six-figure point counts, multi-chunk scans, inlier sets far above the median's
    shared-memory capacity."""
    from tools import rooms
    scene = rooms.generate_outdoor_scene(3000, n_raw=150000, extent=70.0, n_boxes=80)
    labels, stats = _check_rooms(engine, [scene], 0.3, seed=1)
    assert stats['n_points'][0] > 60000 and stats['grow_steps'][0] > 1000
    again, stats2 = engine.segment_raw_rooms([scene], resolution=0.3, seed=1)
    np.testing.assert_array_equal(again[0], labels[0])
    assert stats2['grow_steps'][0] == stats['grow_steps'][0]


def test_spatial_index_changes_no_label(engine):
    """The Morton-ordered spatial index (shell scans of the grow steps, nearest-labelled-point search of the fill) against the
    whole-room scans / the all-pairs fill (LRG_FLAG_NO_SPATIAL_INDEX) on rooms of three shapes at two resolutions: unfilled and
    filled labels and the per-room statistics must be identical -- the fill's block bounds may prune, never decide."""
    from tools import rooms
    from learn_region_grow_b200 import _lib
    cases = [([rooms.generate_room(1000 + i)[:, :6] for i in range(3)], 0.1),
             ([rooms.generate_room(2000 + i, n_raw=n)[:, :6] for i, n in enumerate((5000, 35000))], 0.1),
             ([rooms.generate_outdoor_scene(3000)[:, :6]], 0.3)]
    for raws, res in cases:
        out = []
        for flags in (0, _lib.FLAG_NO_SPATIAL_INDEX):
            engine.upload_raw_rooms(raws, res)
            st = engine.segment_resident(resolution=res, seed=1, flags=flags)
            out.append((engine.labels(False), engine.labels(True), st, engine.profile()['fill_ms']))
        (u0, f0, s0, ms0), (u1, f1, s1, ms1) = out
        for a, b in zip(u0, u1):
            np.testing.assert_array_equal(a, b)
        for a, b in zip(f0, f1):
            np.testing.assert_array_equal(a, b)
        assert s0['grow_steps'].tolist() == s1['grow_steps'].tolist() and s0['regions'].tolist() == s1['regions'].tolist()
        assert all((u == 0).sum() > 0 for u in u0)                     # there was something to fill
        print('fill: %d points, %d unlabeled: %.2f ms through the index, %.2f ms all pairs' %
              (sum(len(u) for u in u0), sum(int((u == 0).sum()) for u in u0), ms0, ms1))
