"""Dataset / point-cloud file helpers with the reference's names and argument meaning
(/root/reference/learn_region_grow_util.py:11-73).  Drivers do ``from learn_region_grow_util import *`` and call
``loadFromH5`` (test_region_grow.py:96-99) and ``savePLY`` (:371-374)."""
import numpy


def loadFromH5(filename, load_labels=True):
    """Split the concatenated ``points`` (sum N, C) dataset into rooms with ``count_room`` (util.py:11-31).
    With labels the last two columns are the object id and the class id."""
    import h5py
    f = h5py.File(filename, 'r')
    all_points = f['points'][:]
    count_room = f['count_room'][:]
    f.close()
    bounds = numpy.concatenate([[0], numpy.cumsum(count_room)])
    rooms = [all_points[bounds[i]:bounds[i + 1], :] for i in range(len(count_room))]
    if not load_labels:
        return rooms
    room = [r[:, :-2] for r in rooms]
    labels = [r[:, -2].astype(int) for r in rooms]
    class_labels = [r[:, -1].astype(int) for r in rooms]
    return room, labels, class_labels


def saveToH5(filename, rooms):
    """Inverse of loadFromH5(load_labels=False): rooms is a list of (N_r, C) arrays (README.md:47-50 layout)."""
    import h5py
    f = h5py.File(filename, 'w')
    f.create_dataset('points', data=numpy.vstack(rooms), dtype=numpy.float32)
    f.create_dataset('count_room', data=[len(r) for r in rooms], dtype=numpy.int32)
    f.close()


STAGED_KEYS = ('points', 'count', 'neighbor_points', 'neighbor_count', 'add', 'remove', 'steps', 'complete')


def saveStagedH5(filename, staged):
    """The staged-training layout of stage_data.py:249-256: per grow step the centred inlier set (``points`` rows split by
    ``count``) with its ``remove`` mask, the neighbour shell (``neighbor_points`` / ``neighbor_count``) with its ``add``
    mask, the step index within its region (``steps``) and the completeness of the region (``complete``); every dataset
    gzip level 4 like the reference's."""
    import h5py
    kinds = dict(points=numpy.float32, neighbor_points=numpy.float32, complete=numpy.float32)
    f = h5py.File(filename, 'w')
    for k in STAGED_KEYS:
        f.create_dataset(k, data=staged[k], compression='gzip', compression_opts=4, dtype=kinds.get(k, numpy.int32))
    f.close()


def loadStagedH5(filename, feature_size=None):
    """Inverse, split per step the way train_region_grow.py:82-110 consumes it: lists of per-step inlier / neighbour arrays
    (columns ``:feature_size``) and their remove / add masks, plus ``steps`` and ``complete``."""
    import h5py
    f = h5py.File(filename, 'r')
    d = {k: f[k][:] for k in STAGED_KEYS}
    f.close()
    out = dict(inlier_count=d['count'], neighbor_count=d['neighbor_count'], steps=d['steps'], complete=d['complete'],
               inlier_points=[], remove=[], neighbor_points=[], add=[])
    idp = 0
    for c in d['count']:
        out['inlier_points'].append(d['points'][idp:idp + c, :feature_size])
        out['remove'].append(d['remove'][idp:idp + c])
        idp += c
    idp = 0
    for c in d['neighbor_count']:
        out['neighbor_points'].append(d['neighbor_points'][idp:idp + c, :feature_size])
        out['add'].append(d['add'][idp:idp + c])
        idp += c
    return out


def savePCD(filename, points):
    """ASCII PCD v0.7 with packed rgb (util.py:33-55)."""
    if len(points) == 0:
        return
    n = len(points)
    lines = ['# .PCD v0.7 - Point Cloud Data file format', 'VERSION 0.7', 'FIELDS x y z rgb', 'SIZE 4 4 4 4',
             'TYPE F F F I', 'COUNT 1 1 1 1', 'WIDTH %d' % n, 'HEIGHT 1', 'VIEWPOINT 0 0 0 1 0 0 0',
             'POINTS %d' % n, 'DATA ascii']
    with open(filename, 'w') as f:
        f.write('\n'.join(lines) + '\n')
        for p in points:
            rgb = (int(p[3]) << 16) | (int(p[4]) << 8) | int(p[5])
            f.write('%f %f %f %d\n' % (p[0], p[1], p[2], rgb))
    print('Saved %d points to %s' % (n, filename))


def savePLY(filename, points):
    """ASCII PLY with uchar colours (util.py:57-73)."""
    header = ['ply', 'format ascii 1.0', 'element vertex %d' % len(points), 'property float x', 'property float y',
              'property float z', 'property uchar red', 'property uchar green', 'property uchar blue', 'end_header']
    with open(filename, 'w') as f:
        f.write('\n'.join(header) + '\n')
        for p in points:
            f.write('%f %f %f %d %d %d\n' % (p[0], p[1], p[2], p[3], p[4], p[5]))
    print('Saved to %s: (%d points)' % (filename, len(points)))
