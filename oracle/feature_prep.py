"""Host restatement of the reference's feature preparation -- TEST INFRASTRUCTURE (see oracle/__init__.py).

``prepare_features`` follows /root/reference/test_region_grow.py:119-173 (equalise to one point per voxel in first-seen order,
room-normalised coordinates, 27-cell covariance of the RAW points -> normal + curvature, curvature / max, seed order) with the
python loops vectorised; the summation order of the covariance (float32 products accumulated in float64 per voxel, voxels in
itertools.product order) is the reference's.

Parity: PINNED by the 13-D features and the seed order the UNMODIFIED reference driver computed for the two golden rooms
(tests/golden/driver_trace_*.npz ``points`` / ``order``; tests/test_rooms.py).  The CUDA feature preparation
(learn_region_grow_b200/csrc/lrg_featprep.cu) is checked against both (tests/test_featprep_gpu.py).
"""
import numpy as np


def _pack(vox):
    lo = vox.min(axis=0)
    span = (vox.max(axis=0) - lo + 3).astype(np.int64)      # +3: room for the -1/+1 neighbour offsets
    v = vox.astype(np.int64) - lo + 1
    return (v[:, 0] * span[1] + v[:, 1]) * span[2] + v[:, 2], span


def prepare_features(unequalized_points, resolution=0.1):
    """test_region_grow.py:119-173 -> dict(points (Neq,13) f32, equalized_idx, unequalized_idx, curvatures f64, order)."""
    raw = np.asarray(unequalized_points)
    xyz_raw = raw[:, :3].astype(np.float32)
    vox = np.round(xyz_raw / resolution).astype(np.int64)                       # :126
    key, span = _pack(vox)
    uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    appearance = np.argsort(first, kind='stable')                               # voxels in first-seen order (:127-129)
    rank = np.empty_like(appearance)
    rank[appearance] = np.arange(len(appearance))
    equalized_idx = first[appearance]
    unequalized_idx = rank[inverse]                                             # :130
    points = raw[equalized_idx]
    xyz = points[:, :3]
    rgb = points[:, 3:6]
    room_coordinates = (xyz - xyz.min(axis=0)) / (xyz.max(axis=0) - xyz.min(axis=0))   # :139

    # per-voxel sums of p and of the float32 outer products (:151-155), then 27-cell gather (:146-150)
    nvox = len(uniq)
    outer = (xyz_raw[:, :, None] * xyz_raw[:, None, :]).astype(np.float64).reshape(-1, 9)
    sumA = np.zeros((nvox, 9))
    sumB = np.zeros((nvox, 3))
    cnt = np.bincount(inverse, minlength=nvox).astype(np.float64)
    for c in range(9):
        sumA[:, c] = np.bincount(inverse, weights=outer[:, c], minlength=nvox)
    for c in range(3):
        sumB[:, c] = np.bincount(inverse, weights=xyz_raw[:, c].astype(np.float64), minlength=nvox)
    ekey = key[equalized_idx]
    accA = np.zeros((len(ekey), 9))
    accB = np.zeros((len(ekey), 3))
    accN = np.zeros(len(ekey))
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                q = ekey + (dx * span[1] + dy) * span[2] + dz
                pos = np.searchsorted(uniq, q)
                pos[pos >= nvox] = nvox - 1
                hit = uniq[pos] == q
                accA[hit] += sumA[pos[hit]]
                accB[hit] += sumB[pos[hit]]
                accN[hit] += cnt[pos[hit]]
    cov = accA.reshape(-1, 3, 3) / accN[:, None, None] - (accB[:, :, None] * accB[:, None, :]) / (accN ** 2)[:, None, None]
    U, S, V = np.linalg.svd(cov)                                                # :157-158
    normals = np.fabs(V[:, 2, :])
    curvatures = np.fabs(S[:, 2] / (S[:, 0] + S[:, 1] + S[:, 2]))               # :159-161
    curvatures = curvatures / curvatures.max()                                  # :162-163
    feats = np.hstack((xyz, room_coordinates, rgb, normals, curvatures.reshape(-1, 1))).astype(np.float32)   # :172
    return dict(points=feats, equalized_idx=equalized_idx, unequalized_idx=unequalized_idx,
                curvatures=curvatures, order=np.argsort(curvatures),            # :183
                obj_id=raw[equalized_idx, 6].astype(int) if raw.shape[1] > 6 else None,
                cls_id=raw[equalized_idx, 7].astype(int) if raw.shape[1] > 7 else None)
