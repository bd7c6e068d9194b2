import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = 'parallel-wavenet-vocoder_b200'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a box without a CUDA device."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def pkg(mod=None):
    return importlib.import_module(PKG + ('.' + mod if mod else ''))


@pytest.fixture()
def hp():
    """The config singleton reset to default.yaml for every test."""
    h = pkg('hparam').hparam
    h.set_hparam_yaml('default')
    return h


def small_case(hp, channels=64, dilations=((1, 2, 4, 512), (1, 8, 64)), n=2, t=1600, precision='fp32'):
    """A small graph; `precision` defaults to the exact fp32 kernels (tests opt in to the tensor-core modes)."""
    hp.set_hparam_dict({
        'model': {'n_iaf': len(dilations), 'dilations': [list(d) for d in dilations],
                  'residual_channels': channels, 'dilation_channels': channels, 'skip_channels': 2 * channels},
        'generate': {'batch_size': n, 'length': t},
        'engine': {'precision': precision},
    }, case='test/small')
    return hp


def load_golden(hp, name):
    """A committed fixture made by tests/golden/make_golden_from_reference.py: configures `hp`,
    regenerates the weights from the stored recipe and verifies their checksum."""
    import numpy as np
    g = np.load(os.path.join(ROOT, 'tests', 'golden', name), allow_pickle=False)
    n_layers = [int(v) for v in g['n_layers']]
    dil = [[int(v) for v in row[:n]] for row, n in zip(g['dilations'], n_layers)]
    model = {'n_iaf': int(g['n_iaf']), 'dilations': dil}
    model['use_skip_connection'] = (name == 'ref_skip.npz')     # generated with model.use_skip_connection=True (modules.py:147)
    for key in ('filter_width', 'residual_channels', 'dilation_channels', 'skip_channels', 'use_skip_connection'):
        if key in g.files:              # the free shape parameters (ref_shapes.npz)
            model[key] = bool(g[key]) if key == 'use_skip_connection' else int(g[key])
    if 'cond_upsample_method' in g.files:
        model['cond_upsample_method'] = str(g['cond_upsample_method'])
    if 'normalize' in g.files:          # every normaliser call site (reference modules.py:263-284)
        model.update({'normalize': str(g['normalize']), 'normalize_cond': str(g['normalize']), 'normalize_wavenet': str(g['normalize'])})
    hp.set_hparam_dict({'model': model}, case='golden/' + name)
    seed, bias_std, gain = g['weight_recipe']
    weights = pkg('weights').init_weights(hp, seed=int(seed), bias_std=float(bias_std), gain=float(gain))
    flat = np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in weights.values()])
    chk = np.array([flat.sum(), np.abs(flat).sum(), flat[::9973].sum()])
    assert np.allclose(chk, g['weight_checksum'], rtol=0, atol=1e-9), 'regenerated weights differ from the fixture recipe'
    return weights, g['noise'], g['mel'], g['wav'], dil
