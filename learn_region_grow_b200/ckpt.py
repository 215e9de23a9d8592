"""TensorFlow checkpoint-V2 ("tensor bundle") reader and writer, no TensorFlow needed.

The reference restores its weights with ``tf.compat.v1.train.Saver().restore(sess, MODEL_PATH)``
(/root/reference/test_region_grow.py:92-93) from ``models/lrgnet_model5.ckpt.{index,data-00000-of-00001}``.
TensorFlow is not installable in this image, so the drop-in ``Saver`` shim reads the two files directly:

* ``.index`` is a LevelDB *table* (sorted string table): data blocks of prefix-compressed
  ``key -> value`` entries, an index block, and a 48-byte footer ending in the magic
  ``0xdb4775248b80fb57``.  The value of the empty key is a ``BundleHeaderProto``; every other value is a
  ``BundleEntryProto {1: dtype, 2: TensorShapeProto, 3: shard_id, 4: offset, 5: size, 6: crc32c}``.
* ``.data-00000-of-00001`` holds the raw little-endian tensor bytes at those offsets.

The writer emits the same layout (single uncompressed data block per ~4 KiB, masked crc32c) so a
weight blob can be turned back into files the reference's own ``Saver`` would accept.
"""
import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DT_FLOAT, _DT_INT32, _DT_INT64 = 1, 3, 9
_NP_OF_DT = {_DT_FLOAT: np.dtype('<f4'), _DT_INT32: np.dtype('<i4'), _DT_INT64: np.dtype('<i8'), 2: np.dtype('<f8')}
_DT_OF_NP = {np.dtype('<f4'): _DT_FLOAT, np.dtype('<i4'): _DT_INT32, np.dtype('<i8'): _DT_INT64, np.dtype('<f8'): 2}


# ----------------------------------------------------------------------------- varint / proto helpers
def _get_varint(buf, pos):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Minimal protobuf wire parser -> list of (field, wiretype, value)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.append((field, wt, v))
    return out


def _parse_entry(buf):
    ent = {'dtype': 0, 'shape': [], 'shard': 0, 'offset': 0, 'size': 0, 'crc32c': 0}
    for field, wt, v in _parse_proto(buf):
        if field == 1:
            ent['dtype'] = v
        elif field == 2:
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:  # Dim
                    size = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            size = v3
                    ent['shape'].append(size)
        elif field == 3:
            ent['shard'] = v
        elif field == 4:
            ent['offset'] = v
        elif field == 5:
            ent['size'] = v
        elif field == 6:
            ent['crc32c'] = struct.unpack('<I', v)[0]
    return ent


# ----------------------------------------------------------------------------- crc32c (Castagnoli), masked as LevelDB does
_CRC_TABLE = None


def _crc32c(data):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = [int(x) for x in tab]
    crc = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- table reader
def _read_block(buf, offset, size):
    block = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError('compressed table blocks are not supported (type %d)' % ctype)
    n_restarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b'', []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_index(prefix):
    """Return {variable name: BundleEntry dict} for checkpoint ``prefix`` (path without extension)."""
    with open(prefix + '.index', 'rb') as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack('<Q', buf[-8:])[0] != _MAGIC:
        raise ValueError('%s.index is not a checkpoint-V2 index (bad table magic)' % prefix)
    footer = buf[-48:]
    pos = 0
    _, pos = _get_varint(footer, pos)   # metaindex offset
    _, pos = _get_varint(footer, pos)   # metaindex size
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    entries = {}
    for _, handle in _read_block(buf, ioff, isize):
        boff, p = _get_varint(handle, 0)
        bsize, p = _get_varint(handle, p)
        for key, val in _read_block(buf, boff, bsize):
            if key == b'':
                continue  # BundleHeaderProto
            entries[key.decode()] = _parse_entry(val)
    return entries


def load_checkpoint(prefix, names=None):
    """Read tensors of a checkpoint-V2 bundle as numpy arrays: {name: ndarray}."""
    entries = read_index(prefix)
    out = {}
    shards = {}
    for name, ent in entries.items():
        if names is not None and name not in names:
            continue
        if ent['dtype'] not in _NP_OF_DT:
            continue
        path = '%s.data-%05d-of-%05d' % (prefix, ent['shard'], 1)
        if path not in shards:
            with open(path, 'rb') as f:
                shards[path] = f.read()
        raw = shards[path][ent['offset']:ent['offset'] + ent['size']]
        out[name] = np.frombuffer(raw, dtype=_NP_OF_DT[ent['dtype']]).reshape(ent['shape']).copy()
    return out


# ----------------------------------------------------------------------------- table writer
def _entry_proto(arr, offset):
    shape = b''
    for d in arr.shape:
        dim = b'\x08' + _put_varint(int(d))
        shape += b'\x12' + _put_varint(len(dim)) + dim
    raw = arr.tobytes()
    msg = b'\x08' + _put_varint(_DT_OF_NP[arr.dtype])
    msg += b'\x12' + _put_varint(len(shape)) + shape
    if offset:
        msg += b'\x20' + _put_varint(offset)
    msg += b'\x28' + _put_varint(len(raw))
    msg += b'\x35' + struct.pack('<I', _mask_crc(_crc32c(raw)))
    return msg


def _build_block(items, restart_interval=16):
    out = bytearray()
    restarts = []
    prev = b''
    for i, (key, val) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val))
        out += key[shared:] + val
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def save_checkpoint(prefix, tensors):
    """Write {name: ndarray} as a single-shard checkpoint-V2 bundle (``prefix.index`` + ``prefix.data-00000-of-00001``)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode())
    data = bytearray()
    items = []
    # BundleHeaderProto: num_shards=1, endianness LITTLE(0, default), version {producer: 1}
    items.append((b'', b'\x08\x01' + b'\x1a\x02\x08\x01'))
    for name in names:
        arr = np.asarray(tensors[name], order='C')
        arr = arr.astype(arr.dtype.newbyteorder('<'), copy=False)
        items.append((name.encode(), _entry_proto(arr, len(data))))
        data += arr.tobytes()
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))

    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)  # no compression
        out.extend(struct.pack('<I', _mask_crc(_crc32c(block + b'\x00'))))
        return off, len(block)

    # data blocks of ~4 KiB
    index_items = []
    cur, cur_size = [], 0
    for key, val in items:
        cur.append((key, val))
        cur_size += len(key) + len(val) + 3
        if cur_size >= 4096:
            off, size = emit(_build_block(cur))
            index_items.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
            cur, cur_size = [], 0
    if cur:
        off, size = emit(_build_block(cur))
        index_items.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
    moff, msize = emit(_build_block([]))
    ioff, isize = emit(_build_block(index_items, restart_interval=1))
    footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', _MAGIC)
    out.extend(footer)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
