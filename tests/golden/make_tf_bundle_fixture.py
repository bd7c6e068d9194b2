#!/usr/bin/env python
"""Writes tests/golden/tf_fixture/model-7.{index,data-00000-of-00001} + `checkpoint`: a small TensorFlow V2 tensor-bundle
checkpoint produced by a SECOND, independent encoder -- not parallel-wavenet-vocoder_b200/tf_bundle.py's writer.

Why: the reference ships no checkpoint and TensorFlow cannot be installed here, so tf_bundle.py's reader was only ever
tested against its own writer (a shared misreading of the format would go unnoticed). This encoder is written straight
from the format descriptions (LevelDB table_format.md; tensorflow/core/util/tensor_bundle/tensor_bundle.h and
protobuf/tensor_bundle.proto, tensor_shape.proto) and deliberately makes the choices the other writer does not:
restart interval 1 (no prefix compression at all) in the first data block, interval 4 in the others, two to five
entries per data block, a NON-empty metaindex block, a bit-serial CRC-32C, protobuf fields emitted by hand including
fields the reader must skip (BundleHeaderProto.version as a nested message, TensorShapeProto.Dim.name). It is still
not TensorFlow: parity with a TF-written file remains unpinned and DESIGN.md says so.

    python tests/golden/make_tf_bundle_fixture.py        # regenerates the committed fixture (deterministic)
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'tf_fixture')


def varint(n):
    b = b''
    while n >= 0x80:
        b += bytes([(n & 0x7F) | 0x80])
        n >>= 7
    return b + bytes([n])


def crc32c_bitwise(data):
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def pb_varint(field, value):
    return varint(field << 3 | 0) + varint(value)


def pb_bytes(field, payload):
    return varint(field << 3 | 2) + varint(len(payload)) + payload


def pb_fixed32(field, value):
    return varint(field << 3 | 5) + struct.pack('<I', value)


def shape_proto(shape):
    out = b''
    for i, d in enumerate(shape):
        dim = pb_varint(1, d) + (pb_bytes(2, b'd%d' % i) if i == 0 else b'')     # Dim{size, name}: name must be skipped
        out += pb_bytes(2, dim)
    return out


def block(entries, restart_interval):
    body, restarts, prev = b'', [], b''
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(body))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    for r in restarts:
        body += struct.pack('<I', r)
    return body + struct.pack('<I', len(restarts))


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(20261017)
    tensors = {
        'iaf_vocoder/cond/dense': rng.uniform(-1, 1, (1, 80, 80)).astype(np.float32),
        'iaf_vocoder/cond/dense/ExponentialMovingAverage': rng.uniform(-1, 1, (1, 80, 80)).astype(np.float32),
        'iaf_vocoder/iaf0/scalar/causal_layer/filter': rng.uniform(-1, 1, (2, 1, 64)).astype(np.float32),
        'iaf_vocoder/iaf0/scalar/causal_layer/filter/ExponentialMovingAverage': rng.uniform(-1, 1, (2, 1, 64)).astype(np.float32),
        'iaf_vocoder/iaf0/scalar/dilated_stack/layer0/filter_bias': rng.normal(0, 0.1, (64,)).astype(np.float32),
        'iaf_vocoder/iaf0/scalar/dilated_stack/layer0/filter': rng.uniform(-1, 1, (2, 64, 64)).astype(np.float32),
        'global_step': np.array(7, dtype=np.int64),
        'learning_rate': np.array(2e-4, dtype=np.float32),
        'beta1_power': np.array(0.9 ** 7, dtype=np.float64),
        'EMA/step_counter': np.array([1, 2, 3], dtype=np.int32),
    }
    dtype_enum = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}
    data, entries = b'', []
    header = pb_varint(1, 1) + pb_varint(2, 0) + pb_bytes(3, pb_varint(1, 1))      # num_shards, LITTLE endian, version{producer: 1}
    entries.append((b'', header))
    for name in sorted(tensors):                    # table keys are sorted bytewise
        arr = tensors[name]
        raw = arr.tobytes()
        entry = pb_varint(1, dtype_enum[arr.dtype]) + pb_bytes(2, shape_proto(arr.shape)) + pb_varint(3, 0) + pb_varint(4, len(data)) + \
            pb_varint(5, len(raw)) + pb_fixed32(6, mask(crc32c_bitwise(raw)))
        entries.append((name.encode(), entry))
        data += raw
    with open(os.path.join(OUT, 'model-7.data-00000-of-00001'), 'wb') as fh:
        fh.write(data)

    table, index_entries = b'', []

    def emit(payload):
        nonlocal table
        off = len(table)
        table += payload + b'\x00' + struct.pack('<I', mask(crc32c_bitwise(payload + b'\x00')))
        return varint(off) + varint(len(payload))

    cuts = [0, 2, 7, len(entries)]                  # 2, 5 and the remaining entries per data block
    for bi in range(len(cuts) - 1):
        chunk = entries[cuts[bi]:cuts[bi + 1]]
        handle = emit(block(chunk, 1 if bi == 0 else 4))
        index_entries.append((chunk[-1][0], handle))        # separator: the block's last key
    meta = emit(block([(b'filter.none', varint(0) + varint(0))], 16))
    index = emit(block(index_entries, 1))
    footer = meta + index
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    with open(os.path.join(OUT, 'model-7.index'), 'wb') as fh:
        fh.write(table + footer)
    with open(os.path.join(OUT, 'checkpoint'), 'w') as fh:
        fh.write('model_checkpoint_path: "model-7"\nall_model_checkpoint_paths: "model-7"\n')
    np.savez(os.path.join(OUT, 'expected.npz'), **{k.replace('/', '|'): v for k, v in tensors.items()})
    print('wrote', OUT, len(table) + len(footer), 'index bytes,', len(data), 'data bytes')


if __name__ == '__main__':
    main()
