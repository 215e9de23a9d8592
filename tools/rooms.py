"""Synthetic S3DIS- / ScanNet- / Semantic-KITTI-shaped inputs for tests, probes and the benchmark (no network: there are no
datasets here).  Nothing in this file is on the product path: the engine takes raw rows and prepares the features on the device
(csrc/lrg_featprep.cu); the host restatement of that preparation used by the tests lives in oracle/feature_prep.py.

* ``generate_room`` follows the reference's own synthetic generator (/root/reference/tools/generate_synthetic_rooms.py:35-99:
  six noisy planes, per-surface mean colour + gaussian colour jitter clipped to [-0.5, 0.5]) and adds axis-aligned
  box "furniture" so that a room yields tens of regions like the S3DIS logs.  Layout is the H5 layout
  of the reference datasets: (N, 8) float32 = x y z r g b obj_id cls_id (README.md:47-50).
"""
import numpy as np

ROOM_MIN = np.array([1.0619999, 1.0630007, 2.073])
ROOM_MAX = np.array([44.094, 46.835, 7.647])
ROOM_DIMENSIONS = np.array([5.133024, 5.169554, 3.0433161])
ROOM_VARIATION = np.array([4.2353425, 5.5636344, 0.58006])
COLOR_VARIATION = np.array([0.15274304, 0.15051211, 0.15046296])
XYZ_NOISE = 0.01


def _surface(rng, n, origin, u, v, obj_id, cls_id, xyz_noise=XYZ_NOISE):
    """n points on the parallelogram origin + a*u + b*v with the reference's noise and colour model."""
    P = np.zeros((n, 8))
    a = rng.random_sample(n)[:, None]
    b = rng.random_sample(n)[:, None]
    P[:, :3] = origin + a * u + b * v
    P[:, :3] += rng.randn(n, 3) * xyz_noise                       # generate_synthetic_rooms.py:46
    return P


def _colorize(rng, P, color_jitter=0.5):
    mean_color = rng.random_sample(3) - 0.5                        # :47
    P[:, 3:6] = mean_color + rng.randn(len(P), 3) * COLOR_VARIATION * color_jitter
    P[:, 3:6] = np.clip(P[:, 3:6], -0.5, 0.5)                      # :49-50


def generate_room(seed, n_raw=20000, n_boxes=None, dims=None, max_dim=12.0, color_jitter=0.5, xyz_noise=XYZ_NOISE):
    """One synthetic room, (N,8) float32, N ~= n_raw.  ``color_jitter`` scales the per-point colour noise (the reference's generator:
    0.5, :48), ``xyz_noise`` is the per-point position noise in metres (the reference: 0.01, :46)."""
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
            parts.append(_surface(rng, n, np.array(o, float), np.array(u, float), np.array(v, float), oid + 1, classes[oid], xyz_noise))
        P = np.vstack(parts)
        _colorize(rng, P, color_jitter)
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
