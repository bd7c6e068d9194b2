"""generate.py keeps the reference's `fire.Fire(generate)` command line (reference generate.py:16-21,77-78) without
depending on fire, and the entry point fails loudly -- no CPU fallback -- when there is no CUDA device."""
import importlib.util
import os

import pytest
import torch

from conftest import ROOT, pkg


def _load():
    spec = importlib.util.spec_from_file_location('generate_entry', os.path.join(ROOT, 'generate.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fire_compatible_arguments():
    cli = _load()._cli
    assert cli([]) == {}
    assert cli(['test/tran']) == {'case': 'test/tran'}
    assert cli(['ema/lj', 'model-1200']) == {'case': 'ema/lj', 'ckpt': 'model-1200'}
    assert cli(['ema/lj', 'model-1200', 'True']) == {'case': 'ema/lj', 'ckpt': 'model-1200', 'debug': True}
    assert cli(['--case=default', '--ckpt', 'model-7', '--debug']) == {'case': 'default', 'ckpt': 'model-7', 'debug': True}
    assert cli(['default', '--ckpt=None']) == {'case': 'default', 'ckpt': None}
    assert cli(['--debug=False']) == {'debug': False}
    with pytest.raises(SystemExit):
        cli(['--no-such-flag=1'])
    with pytest.raises(SystemExit):
        cli(['--ckpt'])


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour WITHOUT a CUDA device')
def test_generate_raises_without_a_gpu(tmp_path, monkeypatch):
    """The synthetic bench case needs no files; without a device the forward pass must raise (the product path never
    routes through the oracle or any CPU implementation)."""
    monkeypatch.chdir(ROOT)
    mod = _load()
    with pytest.raises(Exception) as err:
        mod.generate('bench/c1')
    assert not isinstance(err.value, (ImportError, FileNotFoundError, KeyError)), repr(err.value)


def test_generate_flow_and_sinks_with_a_stand_in_model(tmp_path, monkeypatch):
    """Everything of generate() around the forward pass -- hparams case, synthetic dataset, 'no checkpoint' branch,
    audio summaries, .npy, optional PCM16 wav files -- with the model object replaced by a stand-in that returns a
    known waveform (the forward pass itself is the GPU suite's business; this is NOT a CPU implementation of it)."""
    import numpy as np
    import models as models_shim
    monkeypatch.chdir(ROOT)
    mod = _load()
    hp = pkg('hparam').hparam

    class StandIn:
        def __init__(self, batch_size, length):
            self.n, self.t = batch_size, length

        def __call__(self, wav, melspec, is_training=False, noise=None):
            assert wav is None and melspec.shape == (self.n, 1 + self.t // 80, 80) and noise.shape == (self.n, self.t)
            t = np.arange(self.t, dtype=np.float32) / 16000.0
            return torch.from_numpy(np.stack([0.5 * np.sin(2 * np.pi * 220.0 * (i + 1) * t) for i in range(self.n)])[..., None])

    monkeypatch.setattr(models_shim, 'IAFVocoder', StandIn)
    logdir = str(tmp_path / 'logdir')

    def with_overrides(self, case):              # the case's hparams, redirected to a scratch logdir, wav files on
        return self.set_hparam_dict({'logdir_path': logdir, 'data_path': 'synthetic', 'engine': {'write_wav': True},
                                     'generate': {'batch_size': 2, 'length': 1600}}, case=case)
    monkeypatch.setattr(type(hp), 'set_hparam_yaml', with_overrides)
    pred = mod.generate('bench/c1')
    assert pred.shape == (2, 1600, 1)
    out = hp.logdir
    assert np.array_equal(np.load(os.path.join(out, 'pred_wav.npy')), pred)
    from scipy.io import wavfile
    for i in range(2):
        sr, raw = wavfile.read(os.path.join(out, 'pred_%d.wav' % i))
        assert sr == 16000 and raw.dtype == np.int16 and len(raw) == 1600
        assert np.abs(raw / 32768.0 - pred[i, :, 0]).max() <= 0.5 / 32768 + 1e-7
    assert any(f.startswith('events.out.tfevents') for f in os.listdir(out))
