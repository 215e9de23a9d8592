"""CPU restatement of the statistics block of the reference driver -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/test_region_grow.py:319-349 statement by statement and, like the reference (:12-14), calls
scikit-learn for NMI / AMI / ARS.  One deviation, stated: the objects are ordered with ``numpy.argsort(count,
kind='stable')`` -- the reference's default argsort (:328) leaves the order of equal counts unspecified (it changes with
the numpy build), the engine and this oracle both use the stable order.

Parity: PINNED by the statistics line the unmodified reference driver printed for the two golden rooms
(tests/golden/driver_trace_*.npz ``log``; tests/test_cpu_suite.py::test_metrics_oracle_reproduces_the_reference_log).
"""
import numpy as np
from sklearn.metrics import adjusted_mutual_info_score, adjusted_rand_score, normalized_mutual_info_score


def room_statistics(obj_id, cluster_label):
    """-> dict(nmi, ami, ars, prc, rcl, iou, gt_match, n_clusters, n_classes, cluster_label2)."""
    obj_id = np.asarray(obj_id)
    cluster_label = np.asarray(cluster_label)
    gt_match = 0                                                                  # :320
    dt_match = np.zeros(cluster_label.max(), dtype=bool)                          # :322
    cluster_label2 = np.zeros(len(cluster_label), dtype=int)                      # :323
    room_iou = []
    unique_id, count = np.unique(obj_id, return_counts=True)                      # :325
    for k in range(len(unique_id)):
        i = unique_id[np.argsort(count, kind='stable')][::-1][k]                  # :327
        best_iou = 0
        for j in range(1, cluster_label.max() + 1):                               # :329
            if not dt_match[j - 1]:
                iou = 1.0 * np.sum(np.logical_and(obj_id == i, cluster_label == j)) / np.sum(np.logical_or(obj_id == i, cluster_label == j))
                best_iou = max(best_iou, iou)
                if iou > 0.5:                                                     # :333
                    dt_match[j - 1] = True
                    gt_match += 1
                    cluster_label2[cluster_label == j] = k + 1
                    break
        room_iou.append(best_iou)
    for j in range(1, cluster_label.max() + 1):                                   # :339
        if not dt_match[j - 1]:
            cluster_label2[cluster_label == j] = j + obj_id.max()
    with np.errstate(invalid='ignore', divide='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            prc = np.mean(dt_match)                                               # :342
    rcl = 1.0 * gt_match / len(set(obj_id))                                       # :343
    room_iou = np.mean(room_iou)                                                  # :344
    return dict(nmi=normalized_mutual_info_score(obj_id, cluster_label),          # :346-348
                ami=adjusted_mutual_info_score(obj_id, cluster_label),
                ars=adjusted_rand_score(obj_id, cluster_label),
                prc=float(prc), rcl=rcl, iou=float(room_iou), gt_match=gt_match, n_clusters=int(cluster_label.max()),
                n_classes=len(unique_id), cluster_label2=cluster_label2)
