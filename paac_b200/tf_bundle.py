"""TensorFlow-1.x checkpoint bundles (tensor_bundle "V2": ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``), read and
written without TensorFlow.

The reference checkpoints with ``tf.train.Saver`` (actor_learner.py:79-93, networks.py:122-135) and ships eight
``pretrained/*/checkpoints/-80000000.index`` files (the ``.data`` blobs are absent).  This module makes that on-disk
layout interoperable with the B200 learner: the flat parameter buffer and the RMSProp slots are stored under the
reference's variable names (``local_learning_1/conv1_weights``, ``.../OptimizerVariables``, ...), HWIO /
[in, out] layouts, float32 little-endian.

Format (restated from the published tensor_bundle / LevelDB-table descriptions; pinned by re-serialising the shipped
``.index`` files byte for byte, tests/test_tf_bundle.py):
  .index  = a LevelDB table: data block(s) of prefix-compressed (key, value) entries with a restart array, an empty
            metaindex block, an index block, a 48-byte footer (two block handles, padding, magic 0xdb4775248b80fb57);
            every block is followed by a 1-byte compression type (0) and a masked CRC-32C of block + type.
            key ""   -> BundleHeaderProto  {num_shards = 1, endianness = LITTLE, version {producer = 1}}
            key name -> BundleEntryProto   {dtype, shape, shard_id, offset, size, crc32c (masked, of the tensor bytes)}
  .data-00000-of-00001 = the tensors' raw bytes back to back in key order.
"""
import collections
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
DT_FLOAT = 1
RESTART_INTERVAL = 16
_MASK_DELTA = 0xa282ead8


# ---- CRC-32C (Castagnoli), masked as LevelDB / TensorFlow store it ------------------------------------
def _make_table():
    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    return table


_TABLE = _make_table()
_TABLE_NP = None


def _crc32c_bytes(data, crc=0):
    crc ^= 0xffffffff
    t = _TABLE
    for b in bytes(data):
        crc = t[(crc ^ b) & 0xff] ^ (crc >> 8)
    return crc ^ 0xffffffff


def _gf2_times(mat, vec):
    out, i = 0, 0
    while vec:
        if vec & 1:
            out ^= mat[i]
        vec >>= 1
        i += 1
    return out


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def _combine_matrix(len2):
    """The GF(2) operator that advances a CRC-32C register over len2 zero bytes (zlib's crc32_combine, Castagnoli)."""
    odd = [0x82F63B78] + [1 << n for n in range(31)]          # operator for one zero bit
    even = _gf2_square(odd)                                     # two zero bits
    odd = _gf2_square(even)                                     # four zero bits
    mat = None
    while True:
        even = _gf2_square(odd)                                 # first pass: one zero byte
        if len2 & 1:
            mat = even if mat is None else [_gf2_times(even, mat[n]) for n in range(32)]
        len2 >>= 1
        if not len2:
            break
        odd = _gf2_square(even)
        if len2 & 1:
            mat = odd if mat is None else [_gf2_times(odd, mat[n]) for n in range(32)]
        len2 >>= 1
        if not len2:
            break
    return mat


def crc32c(data, crc=0):
    """CRC-32C of a bytes-like object.  Large inputs (multi-megabyte weight tensors) are cut into equal chunks whose
    CRCs are computed side by side with NumPy table look-ups and then combined (crc(A || B) = M_len(B) crc(A) ^ crc(B))."""
    data = bytes(data)
    n = len(data)
    if n < (1 << 16):
        return _crc32c_bytes(data, crc)
    global _TABLE_NP
    if _TABLE_NP is None:
        _TABLE_NP = np.asarray(_TABLE, dtype=np.uint32)
    chunk = 2048
    k = n // chunk
    body = np.frombuffer(data, dtype=np.uint8, count=k * chunk).reshape(k, chunk)
    regs = np.full(k, 0xffffffff, dtype=np.uint32)
    for j in range(chunk):
        regs = _TABLE_NP[(regs ^ body[:, j]) & 0xff] ^ (regs >> 8)
    regs ^= 0xffffffff
    mat = _combine_matrix(chunk)
    for c in regs.tolist():
        crc = _gf2_times(mat, crc) ^ c
    return _crc32c_bytes(data[k * chunk:], crc) if n > k * chunk else crc


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + _MASK_DELTA) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ---- varints / minimal protobuf ------------------------------------------------------------------------
def _put_varint(out, v):
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7f) << shift
        if b < 0x80:
            return v, pos
        shift += 7


def _parse_proto(buf):
    """-> list of (field, wire_type, value); value is int (varint / fixed32) or bytes (length-delimited)."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        else:
            raise ValueError('unsupported wire type %d' % wt)
        out.append((field, wt, v))
    return out


Entry = collections.namedtuple('Entry', 'dtype shape shard_id offset size crc32c')


def _parse_entry(buf):
    dtype, shape, shard, offset, size, crc = 0, [], 0, 0, 0, 0
    for field, _, v in _parse_proto(buf):
        if field == 1:
            dtype = v
        elif field == 2:                      # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
            for f2, _, dim in _parse_proto(v):
                if f2 == 2:
                    sz = 0
                    for f3, _, x in _parse_proto(dim):
                        if f3 == 1:
                            sz = x
                    shape.append(sz)
        elif field == 3:
            shard = v
        elif field == 4:
            offset = v
        elif field == 5:
            size = v
        elif field == 6:
            crc = v
    return Entry(dtype, tuple(shape), shard, offset, size, crc)


def _encode_entry(e):
    out = bytearray()
    out += b'\x08'
    _put_varint(out, e.dtype)
    shape = bytearray()
    for d in e.shape:
        dim = bytearray(b'\x08')
        _put_varint(dim, d)
        shape += b'\x12'
        _put_varint(shape, len(dim))
        shape += dim
    out += b'\x12'
    _put_varint(out, len(shape))
    out += shape
    if e.shard_id:
        out += b'\x18'
        _put_varint(out, e.shard_id)
    if e.offset:
        out += b'\x20'
        _put_varint(out, e.offset)
    out += b'\x28'
    _put_varint(out, e.size)
    out += b'\x35' + struct.pack('<I', e.crc32c)
    return bytes(out)


HEADER_VALUE = b'\x08\x01\x1a\x02\x08\x01'      # num_shards = 1, (endianness LITTLE = default), version { producer = 1 }


# ---- LevelDB table ---------------------------------------------------------------------------------------
def _read_block(buf, offset, size):
    block = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from('<I', buf, offset + size + 1)[0]
    if ctype != 0:
        raise ValueError('compressed table blocks are not supported')
    if unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
        raise ValueError('table block checksum mismatch')
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    entries, pos, key = [], 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        entries.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return entries


def _build_block(items, restart_interval=RESTART_INTERVAL):
    out, restarts, last, count = bytearray(), [0], b'', 0
    for key, value in items:
        shared = 0
        if count < restart_interval:
            m = min(len(last), len(key))
            while shared < m and last[shared] == key[shared]:
                shared += 1
        else:
            restarts.append(len(out))
            count = 0
        _put_varint(out, shared)
        _put_varint(out, len(key) - shared)
        _put_varint(out, len(value))
        out += key[shared:] + value
        last = key
        count += 1
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _with_trailer(block):
    return block + b'\x00' + struct.pack('<I', mask_crc(crc32c(block + b'\x00')))


def _handle(offset, size):
    out = bytearray()
    _put_varint(out, offset)
    _put_varint(out, size)
    return bytes(out)


def _shortest_separator(key):
    """LevelDB's index key for the LAST data block: the shortest string >= key (FindShortSuccessor)."""
    for i, b in enumerate(key):
        if b != 0xff:
            return key[:i] + bytes([b + 1])
    return key


def read_index(path):
    """-> (header bytes, OrderedDict name -> Entry) of a ``.index`` file."""
    buf = open(path, 'rb').read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError('%s is not a tensor bundle index (bad magic)' % path)
    footer = buf[len(buf) - 48:]
    _, p = _get_varint(footer, 0)             # metaindex handle
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    entries = collections.OrderedDict()
    header = None
    for _, handle in _read_block(buf, ioff, isize):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        for key, value in _read_block(buf, boff, bsize):
            if key == b'':
                header = value
            else:
                entries[key.decode()] = _parse_entry(value)
    return header, entries


def serialize_index(entries):
    """Bytes of a ``.index`` table for an OrderedDict name -> Entry (single data block, as TF writes small bundles)."""
    items = [(b'', HEADER_VALUE)] + [(k.encode(), _encode_entry(e)) for k, e in sorted(entries.items())]
    data = _build_block(items)
    out = bytearray(_with_trailer(data))
    meta_off = len(out)
    meta = _build_block([])
    out += _with_trailer(meta)
    index_off = len(out)
    index = _build_block([(_shortest_separator(items[-1][0]), _handle(0, len(data)))], restart_interval=1)
    out += _with_trailer(index)
    footer = bytearray(_handle(meta_off, len(meta)) + _handle(index_off, len(index)))
    footer += b'\x00' * (40 - len(footer))
    footer += struct.pack('<Q', MAGIC)
    return bytes(out + footer)


def write_bundle(prefix, tensors):
    """tensors: mapping name -> float32 ndarray.  Writes ``prefix.index`` and ``prefix.data-00000-of-00001``."""
    entries, offset = collections.OrderedDict(), 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in sorted(tensors):
            a = np.ascontiguousarray(np.asarray(tensors[name], dtype='<f4'))
            raw = a.tobytes()
            f.write(raw)
            entries[name] = Entry(DT_FLOAT, tuple(int(d) for d in a.shape), 0, offset, len(raw), mask_crc(crc32c(raw)))
            offset += len(raw)
    with open(prefix + '.index', 'wb') as f:
        f.write(serialize_index(entries))
    return entries


def read_bundle(prefix, verify=True):
    """-> OrderedDict name -> float32 ndarray (HWIO / [in, out] as stored)."""
    _, entries = read_index(prefix + '.index')
    blob = open(prefix + '.data-00000-of-00001', 'rb').read()
    out = collections.OrderedDict()
    for name, e in entries.items():
        if e.dtype != DT_FLOAT or e.shard_id != 0:
            raise ValueError('%s: only float32 single-shard bundles are supported' % name)
        raw = blob[e.offset:e.offset + e.size]
        if len(raw) != e.size:
            raise ValueError('%s: data file is truncated' % name)
        if verify and unmask_crc(e.crc32c) != crc32c(raw):
            raise ValueError('%s: tensor checksum mismatch' % name)
        out[name] = np.frombuffer(raw, dtype='<f4').reshape(e.shape).copy()
    return out


def write_checkpoint_state(folder, name):
    """The ``checkpoint`` text file tf.train.latest_checkpoint reads."""
    with open(os.path.join(folder, 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
