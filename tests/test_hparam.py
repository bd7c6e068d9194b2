"""Config surface: same files, keys and merge semantics as the reference's hparam.py."""
import pytest

from conftest import pkg


def test_default_values_match_reference(hp):
    assert hp.signal.sr == 16000 and hp.signal.hop_length == 80 and hp.signal.n_mels == 80
    m = hp.model
    assert (m.filter_width, m.residual_channels, m.dilation_channels, m.skip_channels, m.condition_channels) == (2, 64, 64, 128, 80)
    assert m.n_iaf == 4 and [len(d) for d in m.dilations] == [10, 10, 10, 30]
    assert m.dilations[3][10:20] == [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]
    assert m.use_biases is True and m.use_skip_connection is False
    assert m.normalize == '' and m.normalize_cond == '' and m.normalize_wavenet == ''
    assert m.cond_upsample_method == 'repeat'
    assert hp.generate.length == 64000 and hp.generate.batch_size == 3
    assert hp.train.use_ema is True and hp.train.ema_decay == 0.998
    assert hp.logdir == hp.logdir_path + '/default' and hp.case == 'default'


def test_case_overrides_merge_recursively(hp):
    hp.set_hparam_yaml('ema/len4000')
    assert hp.train.batch_size == 8 and hp.train.num_gpu == 8
    assert hp.train.lr == 0.0002                       # untouched defaults survive inside the section
    assert hp.logdir.endswith('/ema/len4000')
    hp.set_hparam_yaml('test/tran')
    assert hp.model.cond_upsample_method == 'transposed_conv' and hp.model.n_iaf == 4
    hp.set_hparam_yaml('bench/c3')
    assert hp.signal.sr == 24000 and hp.generate.length == 96000 and hp.engine.precision == 'bf16'
    hp.set_hparam_yaml('bench/c2')
    assert hp.engine.precision == 'f16x3' and hp.engine.seed == 0


def test_unknown_case_falls_back_to_defaults(hp):
    hp.set_hparam_yaml('no/such/case')                 # reference hparam.py:59: `if case in user_hp else default_hp`
    assert hp.generate.length == 64000 and hp.case == 'no/such/case'


def test_overlay_semantics():
    H = pkg('hparam')
    user = {'a': {'x': 1}, 'b': 5}
    merged = H.overlay(user, {'a': {'x': 0, 'y': 2}, 'b': {'z': 1}, 'c': 3})
    assert merged == {'a': {'x': 1, 'y': 2}, 'b': 5, 'c': 3}


def test_attribute_access(hp):
    assert hp['model']['n_iaf'] == hp.model.n_iaf
    with pytest.raises(AttributeError):
        hp.model.nonexistent


def test_root_level_dropin_modules_resolve(hp):
    """`from hparam import hparam as hp` / `from models import IAFVocoder` as in reference generate.py:12-13."""
    import importlib
    h = importlib.import_module('hparam')
    assert h.hparam is hp
