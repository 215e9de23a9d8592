"""Device statistics (test_region_grow.py:319-349) against the oracle restatement (which calls scikit-learn,
like the reference) -- through the C ABI: lrg_segmentation_metrics (stateless) and lrg_room_metrics (engine labels)."""
import os

import numpy as np
import pytest

from oracle import feature_prep

from conftest import REPO

pytestmark = pytest.mark.gpu

TOL = 1e-9        # double-precision closed forms; the device lgamma differs from libm by a few ulp


def _check(m, o):
    for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou'):
        if np.isnan(o[k]):
            assert np.isnan(m[k]), k
        else:
            assert abs(m[k] - o[k]) <= TOL, (k, m[k], o[k])
    assert m['gt_match'] == o['gt_match'] and m['n_clusters'] == o['n_clusters'] and m['n_classes'] == o['n_classes']


def _random_room(rng, n, n_obj, n_clu, noise, zero_frac=0.0, obj_values=None):
    obj = rng.randint(0, n_obj, n)
    lab = (obj * 7 + 3) % n_clu + 1
    flip = rng.rand(n) < noise
    lab[flip] = rng.randint(1, n_clu + 1, flip.sum())
    lab[rng.rand(n) < zero_frac] = 0
    if obj_values is not None:
        obj = np.asarray(obj_values)[obj]
    return obj, lab


def test_metrics_match_sklearn_on_random_rooms():
    from learn_region_grow_b200 import metrics
    from oracle import metrics as om
    rng = np.random.RandomState(5)
    rooms_ = [_random_room(rng, 12000, 40, 55, 0.2),
              _random_room(rng, 3000, 5, 3, 0.5),
              _random_room(rng, 20000, 90, 120, 0.05, zero_frac=0.1),
              _random_room(rng, 500, 12, 30, 0.9),
              _random_room(rng, 7000, 25, 25, 0.0),                                  # a perfect segmentation
              _random_room(rng, 4000, 6, 9, 0.3, obj_values=[-5, 0, 3, 1000, 70000, 2]),   # sparse / negative object ids
              _random_room(rng, 64, 2, 2, 0.4)]
    out, l2 = metrics.segmentation_metrics([r[0] for r in rooms_], [r[1] for r in rooms_], return_label2=True)
    for i, (obj, lab) in enumerate(rooms_):
        o = om.room_statistics(obj, lab)
        _check(out[i], o)
        assert np.array_equal(l2[i], o['cluster_label2']), i
        assert out[i]['n_points'] == len(obj)


def test_metrics_degenerate_rooms():
    from learn_region_grow_b200 import metrics
    from oracle import metrics as om
    n = 1000
    rng = np.random.RandomState(2)
    cases = [(np.zeros(n, int) + 4, np.ones(n, int)),                 # one object, one cluster: NMI = AMI = ARS = 1
             (np.zeros(n, int), rng.randint(1, 5, n)),                # one object, several clusters: NMI = AMI = 0
             (rng.randint(0, 6, n), np.ones(n, int)),                 # one cluster
             (rng.randint(0, 6, n), np.zeros(n, int)),                # nothing labelled: cluster_label.max() == 0, PRC = nan
             (rng.randint(0, 3, n), rng.randint(0, 3, n) * 5 + 2),    # label values with gaps: absent ids count as detections
             (np.arange(n) // 100, np.arange(n) // 100 + 1)]          # tied object sizes
    out = metrics.segmentation_metrics([c[0] for c in cases], [c[1] for c in cases])
    for i, (obj, lab) in enumerate(cases):
        _check(out[i], om.room_statistics(obj, lab))
    assert out[0]['nmi'] == 1.0 and out[0]['ami'] == 1.0 and out[0]['ars'] == 1.0
    # an empty room and an empty list
    out = metrics.segmentation_metrics([np.zeros(0, int), cases[0][0]], [np.zeros(0, int), cases[0][1]])
    assert np.isnan(out[0]['nmi']) and out[0]['n_points'] == 0 and out[1]['nmi'] == 1.0
    assert len(metrics.segmentation_metrics([], [])) == 0
    with pytest.raises(ValueError):
        metrics.segmentation_metrics([np.zeros(3, int)], [np.zeros(4, int)])


def test_metrics_reproduce_the_reference_log():
    """The statistics line printed by the unmodified reference driver for the golden rooms (2 decimals)."""
    import re
    from learn_region_grow_b200 import metrics
    from tools import rooms
    objs, labs, logged = [], [], []
    for seed in (1000, 1001):
        z = np.load(os.path.join(REPO, 'tests', 'golden', 'driver_trace_%d.npz' % seed), allow_pickle=True)
        f = feature_prep.prepare_features(z['room'], 0.1)
        objs.append(z['room'][f['equalized_idx'], 6].astype(int))
        labs.append(z['cluster_label'])
        line = [l for l in str(z['log']).split('\n') if l.startswith('Area 5 room 0 NMI')][0]
        logged.append([float(x) for x in re.findall(r': (\d\.\d\d)', line)])
    out = metrics.segmentation_metrics(objs, labs)
    for i in range(2):
        for k, v in zip(('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou'), logged[i]):
            assert abs(out[i][k] - v) <= 0.005 + 1e-9, (i, k, out[i][k], v)


def test_engine_room_metrics_raw_and_equalised(golden_weights):
    """Engine labels scored on the device: raw object ids (gathered with equalized_idx, :136) and equalised ids agree with
    the oracle on the labels the engine produced."""
    from tools import rooms
    from learn_region_grow_b200.engine import Engine
    from oracle import metrics as om
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    raw = [rooms.generate_room(1000, n_raw=2500, n_boxes=4, dims=np.array([3.0, 2.5, 2.2])),
           rooms.generate_room(1001, n_raw=6000, n_boxes=8, dims=np.array([3.0, 2.5, 2.2]))]
    e.upload_raw_rooms(raw, 0.1)
    e.segment_resident(resolution=0.1, seed=0)
    labels = e.labels(True)
    unfilled = e.labels(False)
    feats = e.prepared_features()
    off = e._room_offsets
    obj_raw = [r[:, 6].astype(np.int32) for r in raw]
    m_raw, l2 = e.room_metrics(obj_raw, raw=True, return_label2=True)
    obj_eq = [obj_raw[i][feats['equalized_idx'][off[i]:off[i + 1]]] for i in range(2)]
    m_eq = e.room_metrics(obj_eq)
    m_unf = e.room_metrics(obj_eq, filled=False)
    for i in range(2):
        o = om.room_statistics(obj_eq[i], labels[i])
        _check(m_raw[i], o)
        _check(m_eq[i], o)
        assert np.array_equal(l2[i], o['cluster_label2'])
        _check(m_unf[i], om.room_statistics(obj_eq[i], unfilled[i]))
    with pytest.raises(ValueError):
        e.room_metrics([obj_raw[0]], raw=True)
    e.close()
