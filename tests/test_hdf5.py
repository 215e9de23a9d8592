"""learn_region_grow_b200/hdf5.py -- the native HDF5 subset behind the h5py stand-in.

Pins: (i) a file written by the HDF5 library itself -- tests/golden/hdf5_library_written_matlab73.mat, a copy of
scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat (BSD-3, MATLAB 7.4 -v7.3 save = HDF5 1.6 behind a 512-byte user block;
the only library-written HDF5 file in this image): superblock v0, v1 object headers, symbol-table group, contiguous
float64; (ii) write -> read round trips of the reference's three layouts (``points`` + ``count_room``, learn_region_grow_util.py
:11-20 / tools/generate_synthetic_rooms.py:112-115; the gzip-4 Semantic-KITTI layout, stage_semantic_kitti.py:204-206; the
staged-training layout, stage_data.py:249-256), byte-level checks of the structures the writer emits, and the error paths.
No library-written chunked + deflate fixture exists offline: that branch of the reader is held by the writer only.
"""
import os
import struct
import sys
import zlib

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from learn_region_grow_b200 import hdf5, io_util  # noqa: E402

GOLDEN = os.path.join(REPO, 'tests', 'golden')


@pytest.fixture
def h5py_standin():
    standins = os.path.join(REPO, 'learn_region_grow_b200', 'dropin', 'standins')
    sys.path.insert(0, standins)
    sys.modules.pop('h5py', None)
    try:
        import h5py
        yield h5py
    finally:
        sys.path.remove(standins)
        sys.modules.pop('h5py', None)


def test_reads_a_file_written_by_the_hdf5_library():
    f = hdf5.File(os.path.join(GOLDEN, 'hdf5_library_written_matlab73.mat'), 'r')
    assert f.superblock_version == 0 and list(f.keys()) == ['testdouble']
    d = f['testdouble']
    assert d.shape == (9, 1) and d.dtype == np.float64 and d.chunks is None and d.compression is None
    # scipy's own expectation for this family of fixtures (test_mio.py: testdouble = linspace(0, 2 pi, 9))
    assert np.array_equal(d[:].ravel(), np.arange(0, np.pi * 2 + 1e-9, np.pi / 4)) or np.allclose(d[:].ravel(), np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)
    assert 'testdouble' in f and 'nothing' not in f
    with pytest.raises(KeyError):
        f['nothing']
    f.close()


@pytest.mark.parametrize('kw', [{}, dict(compression='gzip', compression_opts=4), dict(compression='gzip', chunks=(64, 8)),
                                dict(compression='gzip', shuffle=True), dict(chunks=True)])
def test_points_count_room_round_trip(tmp_path, h5py_standin, kw):
    rng = np.random.RandomState(1)
    pts = rng.randn(30011, 8).astype(np.float32)
    cnt = np.array([10000, 20000, 11], np.int32)
    path = str(tmp_path / 'rooms.h5')
    f = h5py_standin.File(path, 'w')
    f.create_dataset('points', data=pts, dtype=np.float32, **kw)
    f.create_dataset('count_room', data=cnt, dtype=np.int32, **{k: v for k, v in kw.items() if k != 'chunks'})
    f.close()
    assert open(path, 'rb').read(8) == hdf5.SIGNATURE
    g = h5py_standin.File(path, 'r')
    assert sorted(g.keys()) == ['count_room', 'points']
    p = g['points']
    assert p.shape == pts.shape and p.dtype == np.float32 and len(p) == len(pts)
    assert np.array_equal(p[:], pts) and np.array_equal(g['count_room'][:], cnt) and g['count_room'].dtype == np.int32
    assert np.array_equal(p[5:9, :3], pts[5:9, :3])
    if kw.get('compression'):
        assert p.compression == 'gzip' and p.compression_opts == 4 and p.chunks is not None and p.shuffle == bool(kw.get('shuffle'))
        if 'chunks' in kw:
            assert p.chunks == kw['chunks']                  # 469 chunks: a two-level chunk B-tree
    g.close()
    rooms = io_util.loadFromH5(path, load_labels=False)
    assert [len(r) for r in rooms] == list(cnt) and np.array_equal(np.vstack(rooms), pts)


def test_staged_training_layout_round_trip(tmp_path, h5py_standin):
    rng = np.random.RandomState(2)
    count = rng.randint(1, 600, 40).astype(np.int32)
    ncount = rng.randint(0, 600, 40).astype(np.int32)
    staged = dict(points=rng.randn(count.sum(), 13).astype(np.float32), count=count,
                  neighbor_points=rng.randn(ncount.sum(), 13).astype(np.float32), neighbor_count=ncount,
                  add=rng.randint(0, 2, ncount.sum()).astype(np.int32), remove=rng.randint(0, 2, count.sum()).astype(np.int32),
                  steps=rng.randint(0, 50, 40).astype(np.int32), complete=rng.rand(40).astype(np.float32))
    path = str(tmp_path / 'staged_area1.h5')
    io_util.saveStagedH5(path, staged)
    f = h5py_standin.File(path, 'r')
    assert sorted(f.keys()) == sorted(io_util.STAGED_KEYS)
    for k in io_util.STAGED_KEYS:
        assert np.array_equal(f[k][:], staged[k]) and f[k].compression == 'gzip' and f[k].compression_opts == 4, k
    out = io_util.loadStagedH5(path, feature_size=13)
    assert len(out['inlier_points']) == 40 and [len(a) for a in out['add']] == list(ncount)
    assert np.array_equal(np.vstack(out['inlier_points']), staged['points']) and np.array_equal(np.concatenate(out['remove']), staged['remove'])


def test_writer_emits_the_structures_of_the_format_specification(tmp_path):
    """Walk the file by hand (fixed offsets of the version-0 superblock, HDF5 File Format Specification III.A / III.D / IV.A)."""
    path = str(tmp_path / 'x.h5')
    w = hdf5.File(path, 'w')
    w.create_dataset('b', data=np.arange(6, dtype=np.int32).reshape(2, 3))
    w.create_dataset('a', data=np.arange(2000, dtype=np.float64), compression='gzip', chunks=(500,))
    w.close()
    d = open(path, 'rb').read()
    assert d[:8] == hdf5.SIGNATURE and d[8] == 0 and d[13] == 8 and d[14] == 8          # version 0, 8-byte offsets / lengths
    base, free, eof, drv = struct.unpack_from('<QQQQ', d, 24)
    assert base == 0 and free == hdf5.UNDEF and eof == len(d) and drv == hdf5.UNDEF
    name_off, root_oh, cache_type = struct.unpack_from('<QQI', d, 56)
    btree, heap = struct.unpack_from('<QQ', d, 80)
    assert cache_type == 1 and d[btree:btree + 4] == b'TREE' and d[heap:heap + 4] == b'HEAP' and root_oh % 8 == 0
    assert d[root_oh] == 1 and struct.unpack_from('<H', d, root_oh + 2)[0] == 1          # v1 object header, one message
    assert struct.unpack_from('<HH', d, root_oh + 16) == (0x11, 16)                      # symbol-table message
    assert struct.unpack_from('<QQ', d, root_oh + 24) == (btree, heap)
    ntype, level, used = struct.unpack_from('<BBH', d, btree + 4)
    snod = struct.unpack_from('<Q', d, btree + 24 + 8)[0]
    assert (ntype, level, used) == (0, 0, 1) and d[snod:snod + 4] == b'SNOD' and struct.unpack_from('<H', d, snod + 6)[0] == 2
    heap_data = struct.unpack_from('<Q', d, heap + 24)[0]
    names = []
    for i in range(2):
        off, oh = struct.unpack_from('<QQ', d, snod + 8 + 40 * i)
        names.append(d[heap_data + off:d.index(b'\0', heap_data + off)])
        assert d[oh] == 1
    assert names == [b'a', b'b']                                                          # entries sorted by name
    r = hdf5.File(path)
    a = r['a']
    assert a.chunks == (500,) and np.array_equal(a[:], np.arange(2000.0)) and np.array_equal(r['b'][:], np.arange(6).reshape(2, 3))
    # the chunk B-tree: 4 children, each a zlib stream of one 500-element chunk
    lay = r._layout(a._msgs)
    node = lay['addr']
    assert d[node:node + 4] == b'TREE' and struct.unpack_from('<BBH', d, node + 4) == (1, 0, 4)
    size, mask, off0, _, child = struct.unpack_from('<IIQQQ', d, node + 24 + 32)
    assert mask == 0 and off0 == 500 and np.array_equal(np.frombuffer(zlib.decompress(d[child:child + size]), np.float64), np.arange(500.0, 1000.0))


def test_other_superblocks_and_unsupported_features(tmp_path):
    path = str(tmp_path / 'x.h5')
    w = hdf5.File(path, 'w')
    w.create_dataset('v', data=np.arange(10, dtype=np.int64))
    with pytest.raises(ValueError):
        w.create_dataset('v', data=np.arange(3))
    with pytest.raises(hdf5.Hdf5Unsupported):
        w.create_dataset('s', data=np.array(['a', 'b']))
    with pytest.raises(hdf5.Hdf5Unsupported):
        w.create_dataset('z', data=np.arange(3), compression='lzf')
    w.close()
    d = bytearray(open(path, 'rb').read())
    # a user block: the same file behind 512 bytes of foreign data still opens (addresses are relative to the superblock)
    ub = str(tmp_path / 'ub.h5')
    open(ub, 'wb').write(b'MATLAB 7.3 MAT-file'.ljust(512, b' ') + bytes(d))
    assert np.array_equal(hdf5.File(ub)['v'][:], np.arange(10))
    # a version-2 superblock in front of the same objects (libver='latest' writes this one)
    root_oh = struct.unpack_from('<Q', d, 64)[0]
    sb2 = hdf5.SIGNATURE + struct.pack('<BBBB', 2, 8, 8, 0) + struct.pack('<QQQQ', 0, hdf5.UNDEF, len(d), root_oh) + b'\0\0\0\0'
    d2 = bytearray(d)
    d2[:len(sb2)] = sb2
    v2 = str(tmp_path / 'v2.h5')
    open(v2, 'wb').write(d2)
    r = hdf5.File(v2)
    assert r.superblock_version == 2 and np.array_equal(r['v'][:], np.arange(10))
    bad = str(tmp_path / 'bad.h5')
    open(bad, 'wb').write(b'not hdf5' * 100)
    with pytest.raises(OSError):
        hdf5.File(bad)
    with pytest.raises(hdf5.Hdf5Unsupported):
        hdf5.File(path, 'a')
    # big-endian and unsigned types decode to native arrays
    be = str(tmp_path / 'be.h5')
    w = hdf5.File(be, 'w')
    w.create_dataset('u', data=np.arange(5, dtype='>u2'))
    w.close()
    u = hdf5.File(be)['u']
    assert u.dtype == np.dtype('<u2') and np.array_equal(u[:], np.arange(5))


@pytest.mark.skipif(not os.path.exists('/root/reference/tools/generate_synthetic_rooms.py'), reason='reference tree not present (GPU box)')
def test_unmodified_reference_generator_writes_hdf5_we_read_back(tmp_path, monkeypatch):
    """tools/generate_synthetic_rooms.py of the reference, unchanged, under the drop-in: its ``h5py.File(...,'w')`` +
    ``create_dataset(..., compression='gzip', compression_opts=4)`` calls (:112-115,127-130) produce real HDF5 files that
    ``loadFromH5`` (learn_region_grow_util.py:11-31) splits into the rooms the script generated."""
    from learn_region_grow_b200 import run_reference
    (tmp_path / 'data').mkdir()
    monkeypatch.chdir(tmp_path)
    g = run_reference.run('/root/reference/tools/generate_synthetic_rooms.py', [])
    for name, n_rooms in (('synthetic_train', 20), ('synthetic_test', 5)):
        path = str(tmp_path / 'data' / (name + '.h5'))
        assert open(path, 'rb').read(8) == hdf5.SIGNATURE
        f = hdf5.File(path)
        assert f['points'].compression == 'gzip' and f['points'].compression_opts == 4 and f['points'].dtype == np.float32
        assert f['count_room'].shape == (n_rooms,) and f['count_room'].dtype == np.int32 and f['points'].shape[1] == 8
        pts, obj, cls = io_util_load(path)
        assert len(pts) == n_rooms and sum(len(p) for p in pts) == f['points'].shape[0] and pts[0].shape[1] == 6
    # the script's last `area` list is the test split it just wrote
    assert all(np.array_equal(a.astype(np.float32)[:, :6], p) for a, p in zip(g['area'], pts))


def io_util_load(path):
    standins = os.path.join(REPO, 'learn_region_grow_b200', 'dropin', 'standins')
    sys.path.insert(0, standins)
    sys.modules.pop('h5py', None)
    try:
        return io_util.loadFromH5(path)
    finally:
        sys.path.remove(standins)
        sys.modules.pop('h5py', None)
