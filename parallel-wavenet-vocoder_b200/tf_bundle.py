"""Reader (and a minimal writer, for tests) of TensorFlow "tensor bundle" checkpoints (V2 format).

The reference restores its weights with `tf.train.Saver(...).restore(sess, ckpt)` from files that
tensorpack's `ModelSaver` wrote into `hp.logdir` (reference generate.py:55-66, train.py:51):
`model-<step>.index`, `model-<step>.data-00000-of-00001` and a text file `checkpoint` naming the
latest prefix. TensorFlow is not available here, so the format is parsed natively:

* `<prefix>.index` is an SSTable in TensorFlow's `table::Table` format (LevelDB's): data blocks of
  prefix-compressed (key, value) entries + restart array, each followed by a 5-byte trailer
  (compression type, masked CRC32C), an index block mapping separator keys to data-block handles,
  and a 48-byte footer (metaindex handle, index handle, magic 0xdb4775248b80fb57).
  Key "" holds a `BundleHeaderProto`; every other key is a tensor name and its value a
  `BundleEntryProto` {1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size, 6: crc32c}.
* `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at (offset, size).

Only what the generation path needs is implemented: uncompressed blocks (what TF's BundleWriter
emits), full (unsliced) tensors, float32/float64/int32/int64 dtypes.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_BYTES = 48
BLOCK_TRAILER_BYTES = 5
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}      # tensorflow.DataType
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ----------------------------------------------------------------------------- primitives
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(value):
    out = bytearray()
    while True:
        b = value & 0x7F
        value >>= 7
        if value:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


_CRC_TABLE = None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), bytewise table; used to verify small blocks / tensors in tests."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            table.append(c)
        _CRC_TABLE = table
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    crc = crc32c(data)
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _parse_proto(buf):
    """Minimal protobuf wire parser -> {field: [values]} (varint / 64-bit / bytes / 32-bit)."""
    fields, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            value, pos = _varint(buf, pos)
        elif wire == 1:
            value = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = _varint(buf, pos)
            value = bytes(buf[pos:pos + n])
            pos += n
        elif wire == 5:
            value = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wire)
        fields.setdefault(field, []).append(value)
    return fields


def _read_block(data, offset, size, verify):
    block = data[offset:offset + size]
    trailer = data[offset + size:offset + size + BLOCK_TRAILER_BYTES]
    if len(block) != size or len(trailer) != BLOCK_TRAILER_BYTES:
        raise ValueError('truncated table block')
    if trailer[0] != 0:
        raise NotImplementedError('compressed table blocks (type %d) are not supported' % trailer[0])
    if verify:
        want = struct.unpack('<I', trailer[1:5])[0]
        if masked_crc32c(block + trailer[0:1]) != want:
            raise ValueError('table block checksum mismatch')
    return block


def _block_entries(block):
    """Decode the prefix-compressed entries of a table block -> [(key, value)]."""
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key, out = 0, b'', []
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        value_len, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + value_len])))
        pos += value_len
    return out


# ----------------------------------------------------------------------------- reader
class BundleReader:
    def __init__(self, prefix, verify_index=True):
        self.prefix = prefix
        with open(prefix + '.index', 'rb') as fh:
            data = fh.read()
        if len(data) < FOOTER_BYTES or struct.unpack_from('<Q', data, len(data) - 8)[0] != TABLE_MAGIC:
            raise ValueError('%s.index is not a TensorFlow table file (bad magic)' % prefix)
        footer = data[len(data) - FOOTER_BYTES:]
        pos = 0
        _, pos = _varint(footer, pos)         # metaindex offset
        _, pos = _varint(footer, pos)         # metaindex size
        index_off, pos = _varint(footer, pos)
        index_size, pos = _varint(footer, pos)
        self.entries = {}
        self.num_shards = 1
        for _, handle in _block_entries(_read_block(data, index_off, index_size, verify_index)):
            off, p = _varint(handle, 0)
            size, _ = _varint(handle, p)
            for key, value in _block_entries(_read_block(data, off, size, verify_index)):
                if key == b'':
                    header = _parse_proto(value)
                    self.num_shards = header.get(1, [1])[0]
                    if header.get(2, [0])[0] != 0:
                        raise NotImplementedError('big-endian bundles are not supported')
                    continue
                f = _parse_proto(value)
                shape = []
                if 2 in f:
                    for dim in _parse_proto(f[2][0]).get(2, []):
                        shape.append(_parse_proto(dim).get(1, [0])[0])
                if 7 in f:
                    raise NotImplementedError('sliced (partitioned) variable %r is not supported' % key.decode())
                self.entries[key.decode()] = dict(dtype=f.get(1, [0])[0], shape=tuple(shape), shard=f.get(3, [0])[0],
                                                  offset=f.get(4, [0])[0], size=f.get(5, [0])[0], crc=f.get(6, [None])[0])

    def keys(self):
        return list(self.entries)

    def shape(self, name):
        return self.entries[name]['shape']

    def tensor(self, name, verify=False):
        e = self.entries[name]
        if e['dtype'] not in DTYPES:
            raise NotImplementedError('%s: dtype enum %d is not supported' % (name, e['dtype']))
        path = '%s.data-%05d-of-%05d' % (self.prefix, e['shard'], self.num_shards)
        with open(path, 'rb') as fh:
            fh.seek(e['offset'])
            raw = fh.read(e['size'])
        if len(raw) != e['size']:
            raise ValueError('%s: data file %s is truncated' % (name, path))
        if verify and e['crc'] is not None and masked_crc32c(raw) != e['crc']:
            raise ValueError('%s: tensor checksum mismatch' % name)
        return np.frombuffer(raw, dtype=DTYPES[e['dtype']]).reshape(e['shape']).copy()


def latest_checkpoint(logdir):
    """`tf.train.latest_checkpoint`: the prefix named by `<logdir>/checkpoint`, else the newest *.index."""
    state = os.path.join(logdir, 'checkpoint')
    if os.path.exists(state):
        with open(state) as fh:
            for line in fh:
                if line.startswith('model_checkpoint_path:'):
                    name = line.split(':', 1)[1].strip().strip('"')
                    prefix = name if os.path.isabs(name) else os.path.join(logdir, name)
                    if os.path.exists(prefix + '.index'):
                        return prefix
    cands = [os.path.join(logdir, f[:-6]) for f in os.listdir(logdir) if f.endswith('.index')] if os.path.isdir(logdir) else []
    return max(cands, key=lambda p: os.path.getmtime(p + '.index')) if cands else None


def load_variables(prefix, names, use_ema=False, dtype=np.float32):
    """name -> array for the requested variable names. With `use_ema` every variable is read from its
    `<name>/ExponentialMovingAverage` shadow -- the map reference generate.py:58-63 hands to tf.train.Saver, whose
    restore FAILS when a shadow is absent; so does this (a checkpoint without shadows would otherwise yield non-EMA
    audio without any notice)."""
    reader = BundleReader(prefix)
    suffix = '/ExponentialMovingAverage'
    out, missing = {}, []
    for name in names:
        key = name + suffix if use_ema else name
        if key not in reader.entries:
            missing.append(key)
            continue
        out[name] = reader.tensor(key).astype(dtype, copy=False)
    if missing:
        hint = ' (train.use_ema is set but the checkpoint has no EMA shadows)' if use_ema and all(m.endswith(suffix) for m in missing) else ''
        raise KeyError('%d variables missing from %s%s, first: %s' % (len(missing), prefix, hint, missing[0]))
    return out


# ----------------------------------------------------------------------------- writer (tests / export)
def _proto_field(field, wire, payload):
    return _put_varint((field << 3) | wire) + payload


def _table_block(items, restart_interval=16):
    buf, restarts, last = bytearray(), [], b''
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack('<I', r)
    buf += struct.pack('<I', len(restarts))
    return bytes(buf)


def write_bundle(prefix, tensors, block_entries=64):
    """Write `tensors` (name -> ndarray) as a one-shard bundle readable by TensorFlow and by BundleReader."""
    names = sorted(tensors)
    data = bytearray()
    items = [(b'', _proto_field(1, 0, _put_varint(1)) + _proto_field(3, 2, _put_varint(2) + _proto_field(1, 0, _put_varint(1))))]
    for name in names:
        arr = np.ascontiguousarray(tensors[name])
        raw = arr.tobytes()
        shape = b''.join(_proto_field(2, 2, _put_varint(len(d)) + d) for d in
                         (_proto_field(1, 0, _put_varint(int(s))) for s in arr.shape))
        entry = (_proto_field(1, 0, _put_varint(DTYPE_IDS[arr.dtype])) + _proto_field(2, 2, _put_varint(len(shape)) + shape)
                 + _proto_field(4, 0, _put_varint(len(data))) + _proto_field(5, 0, _put_varint(len(raw)))
                 + _proto_field(6, 5, struct.pack('<I', masked_crc32c(raw))))
        items.append((name.encode(), entry))
        data += raw
    with open(prefix + '.data-00000-of-00001', 'wb') as fh:
        fh.write(data)
    out, index_items = bytearray(), []
    for i in range(0, len(items), block_entries):
        chunk = items[i:i + block_entries]
        block = _table_block(chunk)
        index_items.append((chunk[-1][0], _put_varint(len(out)) + _put_varint(len(block))))
        out += block + b'\x00' + struct.pack('<I', masked_crc32c(block + b'\x00'))
    meta = _table_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out += meta + b'\x00' + struct.pack('<I', masked_crc32c(meta + b'\x00'))
    index = _table_block(index_items, restart_interval=1)
    index_handle = _put_varint(len(out)) + _put_varint(len(index))
    out += index + b'\x00' + struct.pack('<I', masked_crc32c(index + b'\x00'))
    footer = meta_handle + index_handle
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    with open(prefix + '.index', 'wb') as fh:
        fh.write(out)
    with open(os.path.join(os.path.dirname(prefix) or '.', 'checkpoint'), 'w') as fh:
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (os.path.basename(prefix), os.path.basename(prefix)))
