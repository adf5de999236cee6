"""TensorFlow tensor-bundle (checkpoint) reader, no TensorFlow needed.

The reference saves / loads Keras weights as a TF tensor bundle
(deepxi/model.py:279-280 `load_weights(model_path/epoch-<e-1>/variables/variables)`,
:2377-2383 SaveWeights): `variables.index` is a LevelDB-format table mapping tensor names to
BundleEntryProto{dtype, shape, shard_id, offset, size, crc32c}; `variables.data-XXXXX-of-YYYYY`
hold the raw little-endian tensors.  This module parses the index (names, shapes, per-tensor
masked crc32c) and reads + verifies tensors when the shards exist.  In the reference tree the
weight shards are missing (.MISSING_LARGE_BLOBS), so `load_tensors` raises FileNotFoundError there
and callers fall back to `deepxi_b200.weights.synthetic_*`.
"""
import os
import struct
import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]; pos += 1
        out |= (b & 0x7f) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(buf, off, size):
    """Yield (key, value) of one uncompressed table block."""
    blk = buf[off:off + size]
    if buf[off + size] != 0:
        raise ValueError('compressed table blocks are not supported')
    n_restarts = struct.unpack_from('<I', blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]; pos += non_shared
        yield key, blk[pos:pos + vlen]; pos += vlen


def _proto_fields(buf):
    """Minimal protobuf wire decoder: yields (field_no, wire_type, value)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]; pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos); v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]; pos += 4
        else:
            raise ValueError('wire type %d' % wt)
        yield fno, wt, v


def _shape(buf):
    dims = []
    for fno, _, v in _proto_fields(buf):
        if fno == 2:  # TensorShapeProto.dim
            size = 0
            for f2, _, v2 in _proto_fields(v):
                if f2 == 1:
                    size = v2
            dims.append(size)
    return tuple(dims)


class BundleEntry:
    __slots__ = ('name', 'dtype', 'shape', 'shard_id', 'offset', 'size', 'crc32c')

    def __repr__(self):
        return 'BundleEntry(%s, %s, shard=%d, off=%d, size=%d)' % (self.name, self.shape, self.shard_id,
                                                                  self.offset, self.size)


def read_index(index_path):
    """Returns (num_shards, {name: BundleEntry}) for a `variables.index` file."""
    with open(index_path, 'rb') as f:
        buf = f.read()
    if struct.unpack_from('<Q', buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError('%s is not a table file' % index_path)
    pos = len(buf) - 48
    _, pos = _varint(buf, pos); _, pos = _varint(buf, pos)       # metaindex handle
    ioff, pos = _varint(buf, pos); isize, pos = _varint(buf, pos)  # index handle
    num_shards, entries = 1, {}
    for _, handle in _block_entries(buf, ioff, isize):
        boff, p = _varint(handle, 0); bsize, p = _varint(handle, p)
        for key, val in _block_entries(buf, boff, bsize):
            if key == b'':
                for fno, _, v in _proto_fields(val):
                    if fno == 1:
                        num_shards = v
                continue
            e = BundleEntry()
            e.name, e.dtype, e.shape = key.decode(), 0, ()
            e.shard_id = e.offset = e.size = e.crc32c = 0
            for fno, _, v in _proto_fields(val):
                if fno == 1: e.dtype = v
                elif fno == 2: e.shape = _shape(v)
                elif fno == 3: e.shard_id = v
                elif fno == 4: e.offset = v
                elif fno == 5: e.size = v
                elif fno == 6: e.crc32c = v
            entries[e.name] = e
    return num_shards, entries


_CRC_TABLE = None


def crc32c(data):
    """CRC-32C (Castagnoli), table driven."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = np.array(t, np.uint32)
    crc = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    """TF's crc32c::Mask: rotate right 15 and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def load_tensors(prefix, names=None, verify_crc=False):
    """Reads tensors of `<prefix>.index` / `<prefix>.data-*`; returns {name: ndarray}.

    names: iterable of entry names (default: every float entry that is not an optimizer slot).
    """
    num_shards, entries = read_index(prefix + '.index')
    if names is None:
        names = [n for n, e in entries.items() if e.dtype == 1 and '.OPTIMIZER_SLOT' not in n]
    out, files = {}, {}
    for n in names:
        e = entries[n]
        path = '%s.data-%05d-of-%05d' % (prefix, e.shard_id, num_shards)
        if path not in files:
            if not os.path.exists(path):
                raise FileNotFoundError('checkpoint shard %s is missing' % path)
            files[path] = open(path, 'rb')
        f = files[path]; f.seek(e.offset); raw = f.read(e.size)
        if len(raw) != e.size:
            raise IOError('short read of %s' % n)
        if verify_crc and masked_crc32c(raw) != e.crc32c:
            raise IOError('crc32c mismatch for %s' % n)
        out[n] = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e.dtype]).newbyteorder('<')).reshape(e.shape).copy()
    for f in files.values():
        f.close()
    return out


def keras_weights(prefix, verify_crc=False):
    """{'layer_with_weights-<i>/<var>': ndarray} for the model variables of a Keras checkpoint."""
    suffix = '/.ATTRIBUTES/VARIABLE_VALUE'
    _, entries = read_index(prefix + '.index')
    names = [n for n in entries if n.startswith('layer_with_weights-') and n.endswith(suffix)
             and '.OPTIMIZER_SLOT' not in n]
    raw = load_tensors(prefix, names, verify_crc)
    return {n[:-len(suffix)]: v for n, v in raw.items()}


def keras_weight_shapes(index_path):
    """{'layer_with_weights-<i>/<var>': shape} from an index file alone."""
    suffix = '/.ATTRIBUTES/VARIABLE_VALUE'
    _, entries = read_index(index_path)
    return {n[:-len(suffix)]: e.shape for n, e in entries.items()
            if n.startswith('layer_with_weights-') and n.endswith(suffix) and '.OPTIMIZER_SLOT' not in n}


# ---- writer ---------------------------------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_varint(fno, v):
    return _put_varint((fno << 3) | 0) + _put_varint(v)


def _pb_bytes(fno, b):
    return _put_varint((fno << 3) | 2) + _put_varint(len(b)) + b


def _table_block(entries):
    """One uncompressed table block (every entry is a restart point) + its trailer (type byte, masked crc32c)."""
    body, restarts = bytearray(), []
    for key, val in entries:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(val)) + key + val
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack('<I', r)
    body += struct.pack('<I', len(restarts))
    trailer = b'\x00'
    return bytes(body), trailer + struct.pack('<I', masked_crc32c(bytes(body) + trailer))


def save_tensors(prefix, tensors):
    """Writes {name: float32 ndarray} as a one-shard TF tensor bundle: `<prefix>.index` (LevelDB-format table: header entry ''
    + one BundleEntryProto per tensor, sorted by name) and `<prefix>.data-00000-of-00001` (raw little-endian tensors), with the
    per-tensor masked crc32c TensorFlow verifies on load.  The inverse of load_tensors (deepxi/model.py:2377-2383 SaveWeights)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda n: n.encode())
    entries = [(b'', _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))]          # BundleHeaderProto{num_shards=1, version{producer=1}}
    offset = 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for n in names:
            a = np.ascontiguousarray(np.asarray(tensors[n], dtype='<f4'))
            raw = a.tobytes()
            f.write(raw)
            shape = b''.join(_pb_bytes(2, _pb_varint(1, int(d))) for d in a.shape)
            val = _pb_varint(1, 1) + _pb_bytes(2, shape)                           # dtype DT_FLOAT, TensorShapeProto
            if offset:
                val += _pb_varint(4, offset)
            val += _pb_varint(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack('<I', masked_crc32c(raw))
            entries.append((n.encode(), val))
            offset += len(raw)
    out = bytearray()
    data, trailer = _table_block(entries)
    data_handle = _put_varint(0) + _put_varint(len(data))
    out += data + trailer
    meta, mtrailer = _table_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out += meta + mtrailer
    index, itrailer = _table_block([(entries[-1][0] + b'\xff', data_handle)])    # separator >= the last key of the data block
    index_handle = _put_varint(len(out)) + _put_varint(len(index))
    out += index + itrailer
    footer = meta_handle + index_handle
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', _MAGIC)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))


def save_keras_weights(prefix, weights):
    """{'layer_with_weights-<i>/<var>': ndarray} -> a bundle with the checkpoint's own keys ('<name>/.ATTRIBUTES/VARIABLE_VALUE') that
    keras_weights() / load_tensors() of this module read back (crc-verified).  It is a valid tensor bundle for name-keyed readers
    (tf.train.load_checkpoint), but it carries NO '_CHECKPOINTABLE_OBJECT_GRAPH' entry, so Keras' object-based model.load_weights
    (deepxi/model.py:279) cannot restore from it: the writer exists to round-trip this repo's weights, not to feed TensorFlow."""
    save_tensors(prefix, {n + '/.ATTRIBUTES/VARIABLE_VALUE': v for n, v in weights.items()})
