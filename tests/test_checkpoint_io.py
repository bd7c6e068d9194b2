"""Weights in: TensorFlow tensor-bundle checkpoints (parsed natively) and the .npz container."""
import os

import numpy as np
import pytest

from conftest import pkg, small_case


def _weights_with_ema(hp, seed=0):
    W = pkg('weights')
    live = W.init_weights(hp, seed=seed, bias_std=0.1)
    ema = {k + W.EMA_SUFFIX: (v * 0.5).astype(np.float32) for k, v in live.items()}
    extra = {'global_step': np.array(123, dtype=np.int64), 'learning_rate': np.array(2e-4, dtype=np.float32),
             'beta1_power': np.array(0.9, dtype=np.float32)}
    return live, ema, extra


def test_bundle_roundtrip_and_ema_mapping(hp, tmp_path):
    small_case(hp)
    B = pkg('tf_bundle')
    live, ema, extra = _weights_with_ema(hp)
    prefix = str(tmp_path / 'model-1000')
    B.write_bundle(prefix, {**live, **ema, **extra}, block_entries=7)    # several data blocks
    r = B.BundleReader(prefix)
    assert set(r.keys()) == set(live) | set(ema) | set(extra)
    name = 'iaf_vocoder/iaf0/scalar/dilated_stack/layer1/filter'
    assert r.shape(name) == (2, 64, 64)
    assert np.array_equal(r.tensor(name, verify=True), live[name])
    assert r.tensor('global_step') == 123
    names = list(pkg('weights').variable_shapes(hp).keys())
    got = B.load_variables(prefix, names, use_ema=False)
    assert all(np.array_equal(got[k], live[k]) for k in names)
    got = B.load_variables(prefix, names, use_ema=True)              # reference generate.py:58-63
    assert all(np.array_equal(got[k], ema[k + '/ExponentialMovingAverage']) for k in names)
    assert B.latest_checkpoint(str(tmp_path)) == prefix


def test_bundle_detects_corruption(hp, tmp_path):
    small_case(hp)
    B = pkg('tf_bundle')
    live, _, _ = _weights_with_ema(hp)
    prefix = str(tmp_path / 'model-1')
    B.write_bundle(prefix, live)
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[10] ^= 0xFF
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    r = B.BundleReader(prefix)
    first = sorted(live)[0]
    with pytest.raises(ValueError):
        r.tensor(first, verify=True)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[20] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError):
        B.BundleReader(prefix)
    with pytest.raises(ValueError):
        open(prefix + '.index', 'wb').write(b'not a table' * 10)
        B.BundleReader(prefix)


def test_crc32c_known_answers():
    B = pkg('tf_bundle')
    assert B.crc32c(b'123456789') == 0xE3069283            # the CRC-32C check value
    assert B.crc32c(b'\x00' * 32) == 0x8A9136AA             # RFC 3720 B.4


def test_io_finds_and_loads_checkpoints(hp, tmp_path):
    small_case(hp)
    io, B, W = pkg('io'), pkg('tf_bundle'), pkg('weights')
    live, ema, extra = _weights_with_ema(hp)
    logdir = str(tmp_path)
    assert io.find_checkpoint(logdir) is None                    # -> generate.py prints 'No checkpoint found'
    B.write_bundle(os.path.join(logdir, 'model-5'), {**live, **ema, **extra})
    path = io.find_checkpoint(logdir)
    assert path == os.path.join(logdir, 'model-5.index')
    assert io.find_checkpoint(logdir, 'model-5') == path          # named, as `generate.py case model-5`
    got = io.load_checkpoint(path, use_ema=True)
    W.check_weights(hp, got)
    assert np.array_equal(got['iaf_vocoder/cond/dense'], ema['iaf_vocoder/cond/dense/ExponentialMovingAverage'])
    # .npz container with the same naming
    W.save_npz(os.path.join(logdir, 'weights.npz'), {**live, **ema})
    os.utime(os.path.join(logdir, 'weights.npz'), (2e9, 2e9))
    path = io.find_checkpoint(logdir)
    assert path.endswith('weights.npz')
    got = io.load_checkpoint(path, use_ema=False)
    assert np.array_equal(got['iaf_vocoder/cond/dense'], live['iaf_vocoder/cond/dense'])
    with pytest.raises(FileNotFoundError):
        io.find_checkpoint(logdir, 'model-6')


def test_use_ema_without_shadows_is_an_error(hp, tmp_path):
    """tf.train.Saver(var_list={average_name: v}).restore fails on a checkpoint without shadows (reference
    generate.py:58-63); falling back to the live variables silently would produce non-EMA audio."""
    small_case(hp)
    B = pkg('tf_bundle')
    W = pkg('weights')
    live, _, _ = _weights_with_ema(hp)
    prefix = str(tmp_path / 'model-5')
    B.write_bundle(prefix, live)
    names = list(W.variable_shapes(hp).keys())
    with pytest.raises(KeyError, match='no EMA shadows'):
        B.load_variables(prefix, names, use_ema=True)
    assert all(np.array_equal(B.load_variables(prefix, names)[k], live[k]) for k in names)
    W.save_npz(str(tmp_path / 'w.npz'), live)
    with pytest.raises(KeyError):
        W.load_npz(str(tmp_path / 'w.npz'), use_ema=True)


def test_bundle_from_an_independent_encoder(hp):
    """tests/golden/tf_fixture: a bundle written by a second encoder built from the format descriptions (no prefix
    compression in the first block, several data blocks, a non-empty metaindex block, protobuf fields the reader must skip,
    float64 / int32 / int64 entries) -- read back tensor for tensor, checksums verified, EMA map applied."""
    import os
    from conftest import ROOT
    B = pkg('tf_bundle')
    fx = os.path.join(ROOT, 'tests', 'golden', 'tf_fixture')
    want = {k.replace('|', '/'): v for k, v in np.load(os.path.join(fx, 'expected.npz')).items()}
    assert B.latest_checkpoint(fx) == os.path.join(fx, 'model-7')
    r = B.BundleReader(os.path.join(fx, 'model-7'))
    assert sorted(r.keys()) == sorted(want)
    for name, arr in want.items():
        got = r.tensor(name, verify=True)
        assert got.dtype == arr.dtype and got.shape == arr.shape and np.array_equal(got, arr), name
    names = ['iaf_vocoder/cond/dense', 'iaf_vocoder/iaf0/scalar/causal_layer/filter']
    ema = B.load_variables(os.path.join(fx, 'model-7'), names, use_ema=True)
    assert all(np.array_equal(ema[n], want[n + '/ExponentialMovingAverage']) for n in names)
    live = B.load_variables(os.path.join(fx, 'model-7'), names)
    assert all(np.array_equal(live[n], want[n]) for n in names)
