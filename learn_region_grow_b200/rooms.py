"""Synthetic S3DIS-shaped rooms and the host-side feature preparation that feeds the grow engine.

* ``generate_room`` follows the reference's own synthetic generator (/root/reference/tools/generate_synthetic_rooms.py:35-99:
  six noisy planes, per-surface mean colour + gaussian colour jitter clipped to [-0.5, 0.5]) and adds axis-aligned
  box "furniture" so that a room yields tens of regions like the S3DIS logs.  Layout is the H5 layout
  of the reference datasets: (N, 8) float32 = x y z r g b obj_id cls_id (README.md:47-50).
* ``prepare_features`` is a vectorised restatement of /root/reference/test_region_grow.py:119-173 (equalise to one
  point per voxel, room-normalised coordinates, 27-cell covariance -> normal + curvature).  It runs on the host in
  this round (SURVEY.md 8f-1 marks the device version as the next row); the literal loop lives in
  oracle/feature_prep.py and the two are compared in tests/test_rooms.py.
"""
import numpy as np

ROOM_MIN = np.array([1.0619999, 1.0630007, 2.073])
ROOM_MAX = np.array([44.094, 46.835, 7.647])
ROOM_DIMENSIONS = np.array([5.133024, 5.169554, 3.0433161])
ROOM_VARIATION = np.array([4.2353425, 5.5636344, 0.58006])
COLOR_VARIATION = np.array([0.15274304, 0.15051211, 0.15046296])
XYZ_NOISE = 0.01


def _surface(rng, n, origin, u, v, obj_id, cls_id):
    """n points on the parallelogram origin + a*u + b*v with the reference's noise and colour model."""
    P = np.zeros((n, 8))
    a = rng.random_sample(n)[:, None]
    b = rng.random_sample(n)[:, None]
    P[:, :3] = origin + a * u + b * v
    P[:, :3] += rng.randn(n, 3) * XYZ_NOISE                       # generate_synthetic_rooms.py:46
    return P


def _colorize(rng, P):
    mean_color = rng.random_sample(3) - 0.5                        # :47
    P[:, 3:6] = mean_color + rng.randn(len(P), 3) * COLOR_VARIATION * 0.5
    P[:, 3:6] = np.clip(P[:, 3:6], -0.5, 0.5)                      # :49-50


def generate_room(seed, n_raw=20000, n_boxes=None, dims=None, max_dim=12.0):
    """One synthetic room, (N,8) float32, N ~= n_raw."""
    rng = np.random.RandomState(seed)
    if dims is None:
        dims = ROOM_DIMENSIONS + rng.randn(3) * ROOM_VARIATION     # :104-106
        dims = np.minimum(np.maximum(dims, np.maximum(ROOM_MIN, [2.0, 2.0, 2.073])), np.minimum(ROOM_MAX, max_dim))
    w, l, h = [float(x) for x in dims]
    if n_boxes is None:
        n_boxes = int(rng.randint(20, 41))
    # surfaces: (origin, u, v) ; ids 1..6 are floor, ceiling and the four walls like the reference (:58-98)
    surfaces = [
        ((0, 0, 0), (w, 0, 0), (0, l, 0)), ((0, 0, h), (w, 0, 0), (0, l, 0)),
        ((0, 0, 0), (w, 0, 0), (0, 0, h)), ((0, l, 0), (w, 0, 0), (0, 0, h)),
        ((0, 0, 0), (0, l, 0), (0, 0, h)), ((w, 0, 0), (0, l, 0), (0, 0, h)),
    ]
    objects = [[s] for s in surfaces]
    classes = [2, 1, 3, 3, 3, 3]
    for _ in range(n_boxes):
        sx, sy = rng.uniform(0.3, min(1.6, 0.45 * w)), rng.uniform(0.3, min(1.6, 0.45 * l))
        sz = rng.uniform(0.3, min(1.8, 0.8 * h))
        x0, y0 = rng.uniform(0.05, w - sx - 0.05), rng.uniform(0.05, l - sy - 0.05)
        z0 = 0.0 if rng.random_sample() < 0.8 else rng.uniform(0.0, h - sz)
        faces = [
            ((x0, y0, z0 + sz), (sx, 0, 0), (0, sy, 0)),                        # top
            ((x0, y0, z0), (sx, 0, 0), (0, 0, sz)), ((x0, y0 + sy, z0), (sx, 0, 0), (0, 0, sz)),
            ((x0, y0, z0), (0, sy, 0), (0, 0, sz)), ((x0 + sx, y0, z0), (0, sy, 0), (0, 0, sz)),
        ]
        objects.append(faces)
        classes.append(int(rng.randint(4, 13)))
    areas = [[np.linalg.norm(np.cross(u, v)) for (_, u, v) in faces] for faces in objects]
    total = sum(sum(a) for a in areas)
    out = []
    for oid, (faces, fa) in enumerate(zip(objects, areas)):
        parts = []
        for (o, u, v), a in zip(faces, fa):
            n = max(1, int(round(n_raw * a / total)))
            parts.append(_surface(rng, n, np.array(o, float), np.array(u, float), np.array(v, float), oid + 1, classes[oid]))
        P = np.vstack(parts)
        _colorize(rng, P)
        P[:, 6] = oid + 1
        P[:, 7] = classes[oid]
        out.append(P)
    room = np.vstack(out)
    room = room[rng.permutation(len(room))]       # scanners do not deliver points object by object
    return room.astype(np.float32)


def generate_outdoor_scene(seed, n_raw=300000, extent=100.0, n_boxes=None):
    """Semantic-KITTI-shaped stand-in: an extent x extent metre ground plane with vehicle- and
    pole-sized boxes on it, (N,8) float32, N ~= n_raw; meant to be segmented at resolution 0.3
    (/root/reference/stage_semantic_kitti.py aligns 20 scans and voxel-downsamples them at 0.1 m)."""
    rng = np.random.RandomState(seed)
    if n_boxes is None:
        n_boxes = int(rng.randint(150, 301))
    e = float(extent)
    objects = [[((0, 0, 0), (e, 0, 0), (0, e, 0))]]
    classes = [1]
    for _ in range(n_boxes):
        kind = rng.randint(3)
        sx, sy, sz = [(4.2, 1.8, 1.5), (0.4, 0.4, 6.0), (8.0, 6.0, 4.0)][kind] * rng.uniform(0.7, 1.3, 3)
        x0, y0 = rng.uniform(0.5, e - sx - 0.5), rng.uniform(0.5, e - sy - 0.5)
        objects.append([
            ((x0, y0, sz), (sx, 0, 0), (0, sy, 0)),
            ((x0, y0, 0), (sx, 0, 0), (0, 0, sz)), ((x0, y0 + sy, 0), (sx, 0, 0), (0, 0, sz)),
            ((x0, y0, 0), (0, sy, 0), (0, 0, sz)), ((x0 + sx, y0, 0), (0, sy, 0), (0, 0, sz)),
        ])
        classes.append(2 + kind)
    areas = [[np.linalg.norm(np.cross(u, v)) for (_, u, v) in faces] for faces in objects]
    total = sum(sum(a) for a in areas)
    out = []
    for oid, (faces, fa) in enumerate(zip(objects, areas)):
        parts = []
        for (o, u, v), a in zip(faces, fa):
            n = max(1, int(round(n_raw * a / total)))
            parts.append(_surface(rng, n, np.array(o, float), np.array(u, float), np.array(v, float), oid + 1, classes[oid]))
        P = np.vstack(parts)
        _colorize(rng, P)
        P[:, 6] = oid + 1
        P[:, 7] = classes[oid]
        out.append(P)
    scene = np.vstack(out)
    return scene[rng.permutation(len(scene))].astype(np.float32)


def generate_area(n_rooms, seed_base=1000, n_raw=20000, log_uniform=None):
    """List of rooms; ``log_uniform=(lo, hi)`` draws the raw size per room (ScanNet-shaped)."""
    rooms = []
    for r in range(n_rooms):
        n = n_raw
        if log_uniform is not None:
            rs = np.random.RandomState(seed_base + r)
            n = int(np.exp(rs.uniform(np.log(log_uniform[0]), np.log(log_uniform[1]))))
        rooms.append(generate_room(seed_base + r, n_raw=n))
    return rooms


# ----------------------------------------------------------------------------- feature preparation (host)
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
