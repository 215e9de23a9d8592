"""Native reader / writer for the subset of the HDF5 file format the reference's datasets use.

The reference opens its inputs with ``h5py.File(name, 'r')[key][:]`` (/root/reference/learn_region_grow_util.py:11-20,
train_region_grow.py:72-110) and writes them with ``create_dataset(name, data=, dtype=[, compression='gzip',
compression_opts=4])`` (tools/generate_synthetic_rooms.py:112-115, stage_data.py:249-256, stage_semantic_kitti.py:204-206).
h5py is absent from this image, so this module speaks the on-disk format itself (HDF5 File Format Specification 3.0):

reader   superblock v0/v1 (what h5py writes by default) and v2/v3 (``libver='latest'``), user blocks; object headers v1 and
         v2 with continuation blocks; old-style groups (symbol-table message -> v1 B-tree -> SNOD nodes -> local heap) and
         compact new-style groups (link messages); dataspace v1/v2; fixed-point and IEEE floating-point datatypes of either
         byte order; data layout v1-v3 compact / contiguous / chunked (v1 B-tree of chunks) and v4 single-chunk, implicit
         and fixed-array chunk indexes; filter pipeline v1/v2 with deflate, shuffle and fletcher32.
writer   a genuine HDF5 file (superblock v0, v1 object headers, one symbol-table group, contiguous datasets, or chunked +
         deflate datasets behind a v1 B-tree when ``compression='gzip'``) -- the structures the HDF5 library itself emits
         with default settings, so that files staged here open with the real h5py elsewhere.

Anything outside the subset (dense groups in a fractal heap, compound / variable-length / reference types, extensible
arrays, v2 B-tree chunk indexes, external links) raises ``Hdf5Unsupported`` naming what was met.  Pinned by
tests/test_hdf5.py against a file written by the HDF5 library itself (the MATLAB 7.3 fixture scipy ships) and by
write -> read round trips of the reference's three layouts.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(OSError):
    pass


class Hdf5Unsupported(Hdf5Error):
    pass


# ------------------------------------------------------------------------------------------------------------------ reader
class _Buf:
    """Little-endian cursor over the file bytes."""

    def __init__(self, data, pos=0):
        self.d, self.p = data, pos

    def u(self, n):
        v = int.from_bytes(self.d[self.p:self.p + n], 'little')
        self.p += n
        return v

    def raw(self, n):
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def skip(self, n):
        self.p += n


class Dataset:
    """What ``File[name]`` returns: ``shape``, ``dtype``, ``len()`` and numpy-style ``[...]`` reads (whole array decoded once)."""

    def __init__(self, f, name, msgs):
        self._f, self.name, self._msgs, self._arr = f, name, msgs, None
        self.shape, self.maxshape = f._dataspace(msgs)
        self.dtype = f._datatype(msgs)
        self.chunks, self.compression, self.compression_opts, self.shuffle, self.fletcher32 = f._storage_summary(msgs)

    def _load(self):
        if self._arr is None:
            self._arr = self._f._read_data(self._msgs, self.shape, self.dtype)
        return self._arr

    def __getitem__(self, key):
        return self._load()[key]

    def __len__(self):
        if not self.shape:
            raise TypeError('scalar dataset has no len()')
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = self._load()
        return a if dtype is None else a.astype(dtype)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def ndim(self):
        return len(self.shape)


class Group:
    def __init__(self, f, name, links):
        self._f, self.name, self._links = f, name, links

    def keys(self):
        return self._links.keys()

    def __iter__(self):
        return iter(self._links)

    def __len__(self):
        return len(self._links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, key):
        node = self
        for part in [p for p in key.split('/') if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError("Unable to open object (object '%s' doesn't exist)" % key)
            node = node._f._open(node._links[part], (node.name.rstrip('/') + '/' + part))
        return node

    def items(self):
        return [(k, self[k]) for k in self._links]


MSG_DATASPACE, MSG_LINKINFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK, MSG_LAYOUT, MSG_FILTERS, MSG_CONT, MSG_SYMTAB = (
    0x1, 0x2, 0x3, 0x4, 0x5, 0x6, 0x8, 0xB, 0x10, 0x11)


class Reader(Group):
    """Read-only view of one HDF5 file (``File(name, 'r')`` hands this out)."""

    def __init__(self, filename):
        with open(filename, 'rb') as fh:
            self._d = fh.read()
        self.filename, self.mode = filename, 'r'
        base = 0
        while self._d[base:base + 8] != SIGNATURE:          # a user block pushes the superblock to 512, 1024, 2048 ...
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self._d):
                raise Hdf5Error('%s: not an HDF5 file (no superblock signature)' % filename)
        b = _Buf(self._d, base + 8)
        ver = b.u(1)
        if ver in (0, 1):
            b.skip(4)                                       # free-space, root-entry, reserved, shared-header versions
            self._so, self._sl = b.u(1), b.u(1)
            b.skip(1 + 2 + 2 + 4 + (4 if ver == 1 else 0))  # reserved, group leaf K, group internal K, flags[, chunk K, reserved]
            self._base = b.u(self._so)
            b.skip(3 * self._so)                            # free-space info, end of file, driver info
            b.skip(self._so)                                # root entry: link name offset
            root = b.u(self._so)
        elif ver in (2, 3):
            self._so, self._sl = b.u(1), b.u(1)
            b.skip(1)
            self._base = b.u(self._so)
            b.skip(2 * self._so)                            # superblock extension, end of file
            root = b.u(self._so)
        else:
            raise Hdf5Unsupported('superblock version %d' % ver)
        if self._base == 0 and base:                        # HDF5 stores addresses relative to the base address; with a
            self._base = base                               # user block the library records base 0 and offsets from the superblock
        self.superblock_version = ver
        self._cache = {}
        root_obj = self._open(root, '/')
        if not isinstance(root_obj, Group):
            raise Hdf5Error('root object is not a group')
        Group.__init__(self, self, '/', root_obj._links)

    # h5py surface
    def close(self):
        self._d = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    # ---- object headers
    def _at(self, addr):
        return _Buf(self._d, self._base + addr)

    def _messages(self, addr):
        """All (type, flags, payload bytes) of the object header at addr, continuation blocks included."""
        b = self._at(addr)
        out = []
        if self._d[b.p:b.p + 4] == b'OHDR':
            b.skip(4)
            if b.u(1) != 2:
                raise Hdf5Unsupported('object header version')
            flags = b.u(1)
            if flags & 0x20:
                b.skip(16)
            if flags & 0x10:
                b.skip(4)
            size = b.u(1 << (flags & 3))
            blocks = [(b.p, size)]
            while blocks:
                p, n = blocks.pop(0)
                c = _Buf(self._d, p)
                while c.p + 4 <= p + n:
                    t, sz, fl = c.u(1), c.u(2), c.u(1)
                    if flags & 0x04:
                        c.skip(2)
                    body = c.raw(sz)
                    if t == MSG_CONT:
                        cb = _Buf(body)
                        ca, cl = cb.u(self._so), cb.u(self._sl)
                        if self._d[self._base + ca:self._base + ca + 4] != b'OCHK':
                            raise Hdf5Error('bad object header continuation')
                        blocks.append((self._base + ca + 4, cl - 8))
                    elif t:
                        out.append((t, fl, body))
            return out
        ver = b.u(1)
        if ver != 1:
            raise Hdf5Error('object header version %d at %d' % (ver, addr))
        b.skip(1)
        nmsg, _, size = b.u(2), b.u(4), b.u(4)
        b.skip(4)                                           # pad to 8
        blocks = [(b.p, size)]
        while blocks and len(out) < nmsg + 64:
            p, n = blocks.pop(0)
            c = _Buf(self._d, p)
            while c.p + 8 <= p + n:
                t, sz, fl = c.u(2), c.u(2), c.u(1)
                c.skip(3)
                body = c.raw(sz)
                if t == MSG_CONT:
                    cb = _Buf(body)
                    blocks.append((self._base + cb.u(self._so), cb.u(self._sl)))
                elif t:
                    out.append((t, fl, body))
        return out

    def _open(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        types = [t for t, _, _ in msgs]
        for t, fl, _ in msgs:
            if fl & 0x02:
                raise Hdf5Unsupported('shared header message (type 0x%x) in %s' % (t, name))
        if MSG_LAYOUT in types:
            obj = Dataset(self, name, msgs)
        elif MSG_SYMTAB in types:
            b = _Buf(dict((t, m) for t, _, m in msgs)[MSG_SYMTAB])
            obj = Group(self, name, self._symbol_table(b.u(self._so), b.u(self._so)))
        elif MSG_LINK in types or MSG_LINKINFO in types:
            links = {}
            for t, _, m in msgs:
                if t == MSG_LINKINFO:
                    b = _Buf(m)
                    b.skip(1)
                    fl = b.u(1)
                    if fl & 1:
                        b.skip(8)
                    if b.u(self._so) != UNDEF & ((1 << (8 * self._so)) - 1):
                        raise Hdf5Unsupported('densely stored group %s (fractal heap)' % name)
                if t == MSG_LINK:
                    b = _Buf(m)
                    if b.u(1) != 1:
                        raise Hdf5Unsupported('link message version')
                    fl = b.u(1)
                    ltype = b.u(1) if fl & 0x08 else 0
                    if fl & 0x04:
                        b.skip(8)
                    if fl & 0x10:
                        b.skip(1)
                    lname = bytes(b.raw(b.u(1 << (fl & 3)))).decode('utf-8')
                    if ltype != 0:
                        continue                            # soft / external links are not followed
                    links[lname] = b.u(self._so)
            obj = Group(self, name, links)
        else:
            obj = Group(self, name, {})
        self._cache[addr] = obj
        return obj

    def _symbol_table(self, btree, heap):
        h = self._at(heap)
        if h.raw(4) != b'HEAP':
            raise Hdf5Error('bad local heap')
        h.skip(4)
        h.skip(2 * self._sl)
        heap_data = self._base + h.u(self._so)
        links = {}

        def walk(addr):
            b = self._at(addr)
            sig = bytes(b.raw(4))
            if sig == b'TREE':
                ntype, level, used = b.u(1), b.u(1), b.u(2)
                if ntype != 0:
                    raise Hdf5Error('group B-tree expected')
                b.skip(2 * self._so)
                for _ in range(used):
                    b.skip(self._sl)
                    walk(b.u(self._so))
            elif sig == b'SNOD':
                b.skip(2)
                for _ in range(b.u(2)):
                    off, oh = b.u(self._so), b.u(self._so)
                    ctype = b.u(4)
                    b.skip(4 + 16)
                    if ctype == 2:
                        continue                            # symbolic link
                    p = heap_data + off
                    links[bytes(self._d[p:self._d.index(b'\0', p)]).decode('utf-8')] = oh
            else:
                raise Hdf5Error('bad group node signature %r' % sig)
        walk(btree)
        return links

    # ---- dataset messages
    @staticmethod
    def _get(msgs, t):
        for tt, _, m in msgs:
            if tt == t:
                return m
        return None

    def _dataspace(self, msgs):
        b = _Buf(self._get(msgs, MSG_DATASPACE))
        ver, rank, flags = b.u(1), b.u(1), b.u(1)
        if ver == 1:
            b.skip(5)
        elif ver == 2:
            if b.u(1) == 2:
                return None, None                           # null dataspace
        else:
            raise Hdf5Unsupported('dataspace version %d' % ver)
        dims = tuple(b.u(self._sl) for _ in range(rank))
        mx = tuple(None if v == UNDEF & ((1 << (8 * self._sl)) - 1) else v for v in (b.u(self._sl) for _ in range(rank))) if flags & 1 else dims
        return dims, mx

    def _datatype(self, msgs):
        b = _Buf(self._get(msgs, MSG_DATATYPE))
        head = b.u(4)
        cls, bits, size = head & 0xF, head >> 8, b.u(4)
        order = '>' if bits & 1 else '<'
        if cls == 0:
            return np.dtype('%s%s%d' % (order, 'i' if bits & 0x08 else 'u', size))
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Unsupported('%d-byte floating point' % size)
            return np.dtype('%sf%d' % (order, size))
        names = {2: 'time', 3: 'string', 4: 'bitfield', 5: 'opaque', 6: 'compound', 7: 'reference', 8: 'enum', 9: 'variable-length', 10: 'array'}
        if cls == 3:
            return np.dtype('S%d' % size)
        if cls in (4, 5):
            return np.dtype('V%d' % size) if cls == 5 else np.dtype('%su%d' % (order, size))
        raise Hdf5Unsupported('datatype class %s' % names.get(cls, cls))

    def _filters(self, msgs):
        m = self._get(msgs, MSG_FILTERS)
        if m is None:
            return []
        b = _Buf(m)
        ver, n = b.u(1), b.u(1)
        if ver == 1:
            b.skip(6)
        elif ver != 2:
            raise Hdf5Unsupported('filter pipeline version %d' % ver)
        out = []
        for _ in range(n):
            fid = b.u(2)
            nlen = b.u(2) if (ver == 1 or fid >= 256) else 0
            b.skip(2)
            nvals = b.u(2)
            b.skip(nlen if ver == 2 else (nlen + 7) // 8 * 8)
            vals = [b.u(4) for _ in range(nvals)]
            if ver == 1 and nvals & 1:
                b.skip(4)
            out.append((fid, vals))
        return out

    def _layout(self, msgs):
        """-> dict(kind=compact|contiguous|chunked, ...)."""
        b = _Buf(self._get(msgs, MSG_LAYOUT))
        ver = b.u(1)
        if ver in (1, 2):
            ndim, cls = b.u(1), b.u(1)
            b.skip(5)
            addr = b.u(self._so) if cls != 0 else None
            dims = [b.u(4) for _ in range(ndim)]
            if cls == 2:
                b.skip(4)
                return dict(kind='chunked', index='btree1', addr=addr, chunk=tuple(dims[:-1]) if ver == 2 else tuple(dims))
            if cls == 0:
                return dict(kind='compact', data=b.raw(b.u(4)))
            return dict(kind='contiguous', addr=addr, size=None)
        cls = b.u(1)
        if cls == 0:
            return dict(kind='compact', data=b.raw(b.u(2)))
        if cls == 1:
            return dict(kind='contiguous', addr=b.u(self._so), size=b.u(self._sl))
        if cls != 2:
            raise Hdf5Unsupported('layout class %d (virtual dataset)' % cls)
        if ver == 3:
            ndim = b.u(1)
            addr = b.u(self._so)
            dims = [b.u(4) for _ in range(ndim)]
            return dict(kind='chunked', index='btree1', addr=addr, chunk=tuple(dims[:-1]))
        if ver != 4:
            raise Hdf5Unsupported('layout version %d' % ver)
        flags, ndim, enc = b.u(1), b.u(1), b.u(1)
        dims = [b.u(enc) for _ in range(ndim)]
        idx = b.u(1)
        out = dict(kind='chunked', chunk=tuple(dims[:-1]))
        if idx == 1:
            out['index'] = 'single'
            if flags & 2:
                out['fsize'], out['fmask'] = b.u(self._sl), b.u(4)
        elif idx == 2:
            out['index'] = 'implicit'
        elif idx == 3:
            out['index'] = 'farray'
            b.skip(1)
        else:
            raise Hdf5Unsupported('chunk index type %d (%s)' % (idx, {4: 'extensible array', 5: 'v2 B-tree'}.get(idx, '?')))
        out['addr'] = b.u(self._so)
        return out

    def _storage_summary(self, msgs):
        lay = self._layout(msgs)
        fl = dict(self._filters(msgs))
        for fid in fl:
            if fid not in (1, 2, 3):
                raise Hdf5Unsupported('filter id %d' % fid)
        return (lay.get('chunk'), 'gzip' if 1 in fl else None, (fl[1][0] if 1 in fl and fl[1] else None), 2 in fl, 3 in fl)

    def _unfilter(self, raw, filters, mask, itemsize):
        for i in reversed(range(len(filters))):
            fid, vals = filters[i]
            if mask >> i & 1:
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                es = vals[0] if vals else itemsize
                n = len(raw) // es
                body = np.frombuffer(raw, np.uint8, n * es).reshape(es, n).T.tobytes()
                raw = body + bytes(raw[n * es:])
            elif fid == 3:
                raw = raw[:-4]                              # checksum trails the chunk; not verified
        return raw

    def _read_data(self, msgs, shape, dtype):
        if shape is None:
            return np.zeros((0,), dtype)
        lay = self._layout(msgs)
        n = int(np.prod(shape, dtype=np.int64))
        if lay['kind'] == 'compact':
            return np.frombuffer(bytes(lay['data']), dtype, n).reshape(shape).astype(dtype.newbyteorder('='))
        so_mask = (1 << (8 * self._so)) - 1
        if lay['kind'] == 'contiguous':
            if lay['addr'] == UNDEF & so_mask or n == 0:
                return np.zeros(shape, dtype.newbyteorder('='))      # never written: fill value
            return np.frombuffer(self._d, dtype, n, self._base + lay['addr']).reshape(shape).astype(dtype.newbyteorder('='))
        chunk, filters = lay['chunk'], self._filters(msgs)
        out = np.zeros(shape, dtype.newbyteorder('='))
        if lay['addr'] == UNDEF & so_mask or n == 0:
            return out
        csize = int(np.prod(chunk, dtype=np.int64)) * dtype.itemsize

        def place(offset, raw, mask=0):
            raw = self._unfilter(raw, filters, mask, dtype.itemsize) if filters else raw
            blk = np.frombuffer(bytes(raw[:csize]), dtype, csize // dtype.itemsize).reshape(chunk)
            sel_o = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offset, chunk, shape))
            sel_c = tuple(slice(0, s.stop - s.start) for s in sel_o)
            out[sel_o] = blk[sel_c]

        grid = [-(-s // c) for s, c in zip(shape, chunk)]
        if lay['index'] == 'btree1':
            self._walk_chunks(lay['addr'], len(shape), place)
        elif lay['index'] == 'single':
            p = self._base + lay['addr']
            place((0,) * len(shape), self._d[p:p + lay.get('fsize', csize)], lay.get('fmask', 0))
        elif lay['index'] == 'implicit':
            for i, off in enumerate(np.ndindex(*grid)):
                p = self._base + lay['addr'] + i * csize
                place(tuple(o * c for o, c in zip(off, chunk)), self._d[p:p + csize])
        elif lay['index'] == 'farray':
            self._walk_fixed_array(lay['addr'], grid, chunk, csize, bool(filters), place)
        return out

    def _walk_chunks(self, addr, rank, place):
        b = self._at(addr)
        if bytes(b.raw(4)) != b'TREE':
            raise Hdf5Error('bad chunk B-tree node')
        ntype, level, used = b.u(1), b.u(1), b.u(2)
        if ntype != 1:
            raise Hdf5Error('chunk B-tree expected')
        b.skip(2 * self._so)
        for _ in range(used):
            size, mask = b.u(4), b.u(4)
            offset = tuple(b.u(8) for _ in range(rank))
            b.skip(8)
            child = b.u(self._so)
            if level:
                self._walk_chunks(child, rank, place)
            else:
                place(offset, self._d[self._base + child:self._base + child + size], mask)

    def _walk_fixed_array(self, addr, grid, chunk, csize, filtered, place):
        b = self._at(addr)
        if bytes(b.raw(4)) != b'FAHD':
            raise Hdf5Error('bad fixed array header')
        b.skip(2)
        esize, page_bits = b.u(1), b.u(1)
        nent = b.u(self._sl)
        db = self._at(b.u(self._so))
        if bytes(db.raw(4)) != b'FADB':
            raise Hdf5Error('bad fixed array data block')
        db.skip(2 + self._so)
        if nent > (1 << page_bits):
            raise Hdf5Unsupported('paged fixed array chunk index')
        for i, off in enumerate(np.ndindex(*grid)):
            if i >= nent:
                break
            a = db.u(self._so)
            size, mask = (db.u(esize - self._so - 4), db.u(4)) if filtered else (csize, 0)
            if a != UNDEF & ((1 << (8 * self._so)) - 1):
                place(tuple(o * c for o, c in zip(off, chunk)), self._d[self._base + a:self._base + a + size], mask)


# ------------------------------------------------------------------------------------------------------------------ writer
def _guess_chunk(shape, itemsize):
    """Chunk shape in the spirit of h5py's auto-chunking (halve the axes round-robin until a chunk is 8 KB - 1 MB)."""
    chunk = [max(int(s), 1) for s in shape]
    nbytes = int(np.prod(chunk, dtype=np.int64)) * itemsize
    target = min(max(16384.0 * 2 ** np.log10(max(nbytes, 1) / (1024.0 * 1024)), 8192), 1048576)
    i = 0
    while True:
        cb = int(np.prod(chunk, dtype=np.int64)) * itemsize
        if (cb < target or abs(cb - target) / target < 0.5) and cb < 1048576:
            break
        if all(c == 1 for c in chunk):
            break
        chunk[i % len(chunk)] = int(np.ceil(chunk[i % len(chunk)] / 2.0))
        i += 1
    return tuple(chunk)


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _msg(t, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHB3x', t, len(body), flags) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in 'iu':
        head = 0x10 | 0 | ((0x08 if dt.kind == 'i' else 0) << 8)
        return struct.pack('<II', head, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        # class 1, version 1; bit field: little endian, mantissa normalisation 2 (msb implied) in bits 4-5, sign location in bits 8-15
        sign, eloc, esz, msz, bias = (31, 23, 8, 23, 127) if dt.itemsize == 4 else (63, 52, 11, 52, 1023)
        head = 0x11 | ((0x20 | (sign << 8)) << 8)
        return struct.pack('<II', head, dt.itemsize) + struct.pack('<HHBBBBI', 0, 8 * dt.itemsize, eloc, esz, 0, msz, bias)
    raise Hdf5Unsupported('writing dtype %s' % dt)


class Writer:
    """``File(name, 'w')``: datasets are collected by ``create_dataset`` and the file is laid out at ``close()``."""
    LEAF_K, INTERNAL_K, CHUNK_K = 16, 16, 32

    def __init__(self, filename):
        self.filename, self.mode = filename, 'w'
        self._sets = {}

    def create_dataset(self, name, shape=None, dtype=None, data=None, compression=None, compression_opts=None, chunks=None, shuffle=False, **kw):
        name = name.strip('/')
        if '/' in name:
            raise Hdf5Unsupported('nested groups on write')
        if name in self._sets:
            raise ValueError('Unable to create dataset (name already exists)')
        arr = np.zeros(shape, dtype or np.float32) if data is None else np.asarray(data, dtype=dtype)
        if arr.dtype.kind not in 'iuf' or (arr.dtype.kind == 'f' and arr.dtype.itemsize == 2):
            raise Hdf5Unsupported('writing dtype %s' % arr.dtype)
        if compression in (True, 'gzip') or isinstance(compression, int) and not isinstance(compression, bool):
            level = compression if isinstance(compression, int) and not isinstance(compression, bool) else (4 if compression_opts is None else int(compression_opts))
        elif compression is None:
            level = None
        else:
            raise Hdf5Unsupported('compression %r' % (compression,))
        if (level is not None or chunks) and arr.ndim == 0:
            raise TypeError('Scalar datasets do not support chunking/compression')
        if chunks is True or (chunks is None and level is not None):
            chunks = _guess_chunk(arr.shape, arr.dtype.itemsize)
        if chunks is not None and chunks is not False:
            chunks = tuple(int(c) for c in (chunks if isinstance(chunks, (tuple, list)) else (chunks,)))
            if len(chunks) != arr.ndim or any(c < 1 for c in chunks):
                raise ValueError('Chunk shape must be positive and match the dataset rank')
        else:
            chunks = None
        self._sets[name] = (np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder('<'))), chunks, level, bool(shuffle))
        return _Pending(arr)

    def __contains__(self, key):
        return key.strip('/') in self._sets

    def keys(self):
        return self._sets.keys()

    def __getitem__(self, key):
        return _Pending(self._sets[key.strip('/')][0])

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    def close(self):
        if self._sets is None:
            return
        out = bytearray(96)                                  # superblock v0 with 8-byte offsets and lengths, patched last

        def alloc(b):
            out.extend(b'\0' * (-len(out) % 8))
            a = len(out)
            out.extend(b)
            return a

        entries = []                                         # (name, object header address)
        for name in sorted(self._sets, key=lambda s: s.encode('utf-8')):
            arr, chunks, level, shuffle = self._sets[name]
            msgs = _msg(MSG_DATASPACE, struct.pack('<BBB5x', 1, arr.ndim, 1) + b''.join(struct.pack('<Q', s) for s in arr.shape) * 2)
            msgs += _msg(MSG_DATATYPE, _dtype_msg(arr.dtype), flags=1)
            msgs += _msg(MSG_FILL, struct.pack('<BBBBI', 2, 3 if chunks else 2, 2, 1, 0))
            if chunks:
                filters = ([(2, b'shuffle', [arr.dtype.itemsize])] if shuffle else []) + ([(1, b'deflate', [level])] if level is not None else [])
                if filters:
                    body = struct.pack('<BB6x', 1, len(filters))
                    for fid, fname, vals in filters:
                        nm = _pad8(fname + b'\0')
                        body += struct.pack('<HHHH', fid, len(nm), 1, len(vals)) + nm + b''.join(struct.pack('<I', v) for v in vals)
                        body += b'\0' * (4 if len(vals) & 1 else 0)
                    msgs += _msg(MSG_FILTERS, body, flags=1)
                root = self._write_chunks(arr, tuple(int(c) for c in chunks), level, shuffle, alloc, out)
                msgs += _msg(MSG_LAYOUT, struct.pack('<BBB', 3, 2, arr.ndim + 1) + struct.pack('<Q', root) +
                             b''.join(struct.pack('<I', c) for c in tuple(chunks) + (arr.dtype.itemsize,)))
                nmsg = 5 if filters else 4
            else:
                addr = alloc(arr.tobytes()) if arr.size else UNDEF
                msgs += _msg(MSG_LAYOUT, struct.pack('<BB', 3, 1) + struct.pack('<QQ', addr, arr.nbytes))
                nmsg = 4
            entries.append((name, alloc(struct.pack('<BxHII4x', 1, nmsg, 1, len(msgs)) + msgs)))

        heap = bytearray(8)                                  # offset 0 = the empty name the first B-tree key points at
        offs = []
        for name, _ in entries:
            offs.append(len(heap))
            heap.extend(_pad8(name.encode('utf-8') + b'\0'))
        free = len(heap)
        heap.extend(struct.pack('<QQ', 1, 32) + b'\0' * 16)  # one free block closes the segment (next = 1: last)
        heap_data = alloc(bytes(heap))
        heap_addr = alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free, heap_data))
        per = 2 * self.LEAF_K
        snods = []
        for i in range(0, max(len(entries), 1), per):
            part = list(zip(offs, entries))[i:i + per]
            body = b'SNOD' + struct.pack('<BxH', 1, len(part))
            for off, (_, oh) in part:
                body += struct.pack('<QQII16x', off, oh, 0, 0)
            body += b'\0' * (40 * (per - len(part)))
            snods.append((alloc(body), part[-1][0] if part else 0))
        if len(snods) > 2 * self.INTERNAL_K:
            raise Hdf5Unsupported('more than %d datasets in one file' % (per * 2 * self.INTERNAL_K))
        node = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF) + struct.pack('<Q', 0)
        for a, last in snods:
            node += struct.pack('<QQ', a, last)
        node += b'\0' * (16 * (2 * self.INTERNAL_K - len(snods)))
        btree = alloc(node)
        root_msgs = _msg(MSG_SYMTAB, struct.pack('<QQ', btree, heap_addr))
        root = alloc(struct.pack('<BxHII4x', 1, 1, 1, len(root_msgs)) + root_msgs)
        out.extend(b'\0' * (-len(out) % 8))
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(out), UNDEF)
        sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree, heap_addr)
        out[:96] = sb
        with open(self.filename, 'wb') as fh:
            fh.write(out)
        self._sets = None

    def _write_chunks(self, arr, chunk, level, shuffle, alloc, out):
        """Chunks in row-major grid order behind a v1 B-tree (node type 1); returns the root node's address."""
        rank, es = arr.ndim, arr.dtype.itemsize
        grid = [-(-s // c) for s, c in zip(arr.shape, chunk)]
        nodes = []                                           # ((filtered size, offset tuple), address) of this level's children
        for g in (np.ndindex(*grid) if arr.size else ()):
            off = tuple(i * c for i, c in zip(g, chunk))
            blk = np.zeros(chunk, arr.dtype)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(off, chunk, arr.shape))
            blk[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
            raw = blk.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
            if level is not None:
                raw = zlib.compress(raw, level)
            nodes.append(((len(raw), off), alloc(raw)))
        if not nodes:
            return UNDEF
        end_key = (0, tuple(g * c for g, c in zip(grid, chunk)))        # one chunk past the last in every dimension

        def key(k):
            return struct.pack('<II', k[0], 0) + b''.join(struct.pack('<Q', o) for o in k[1]) + struct.pack('<Q', 0)

        cap, level_no = 2 * self.CHUNK_K, 0
        node_bytes = 24 + (cap + 1) * (16 + 8 * rank) + cap * 8
        while True:
            groups = [nodes[i:i + cap] for i in range(0, len(nodes), cap)]
            addrs = [alloc(b'\0' * node_bytes) for _ in groups]          # siblings point at each other: place first, fill second
            for i, g in enumerate(groups):
                body = b'TREE' + struct.pack('<BBHQQ', 1, level_no, len(g), addrs[i - 1] if i else UNDEF,
                                             addrs[i + 1] if i + 1 < len(groups) else UNDEF)
                for k, a in g:
                    body += key(k) + struct.pack('<Q', a)
                body += key(groups[i + 1][0][0] if i + 1 < len(groups) else end_key)
                out[addrs[i]:addrs[i] + len(body)] = body
            nodes = [(g[0][0], a) for g, a in zip(groups, addrs)]
            if len(nodes) == 1:
                return nodes[0][1]
            level_no += 1


class _Pending:
    def __init__(self, arr):
        self._arr, self.shape, self.dtype = arr, arr.shape, arr.dtype

    def __getitem__(self, key):
        return self._arr[key]

    def __len__(self):
        return len(self._arr)


def File(name, mode='r', **kw):
    """``h5py.File`` for the two modes the reference uses: 'r' (learn_region_grow_util.py:12) and 'w' (stage_data.py:243)."""
    if mode == 'r':
        return Reader(name)
    if mode in ('w', 'w-', 'x'):
        return Writer(name)
    raise Hdf5Unsupported("File mode %r (only 'r' and 'w')" % mode)
