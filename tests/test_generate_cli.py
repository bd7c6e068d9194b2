"""generate.py keeps the reference's `fire.Fire(generate)` command line (reference generate.py:16-21,77-78) without
depending on fire, and the entry point fails loudly -- no CPU fallback -- when there is no CUDA device."""
import importlib.util
import os

import pytest
import torch

from conftest import ROOT


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
