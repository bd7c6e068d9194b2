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


def pkg(mod=None):
    return importlib.import_module(PKG + ('.' + mod if mod else ''))


@pytest.fixture()
def hp():
    """The config singleton reset to default.yaml for every test."""
    h = pkg('hparam').hparam
    h.set_hparam_yaml('default')
    return h


def small_case(hp, channels=64, dilations=((1, 2, 4, 512), (1, 8, 64)), n=2, t=1600, precision='fp32'):
    hp.set_hparam_dict({
        'model': {'n_iaf': len(dilations), 'dilations': [list(d) for d in dilations],
                  'residual_channels': channels, 'dilation_channels': channels, 'skip_channels': 2 * channels},
        'generate': {'batch_size': n, 'length': t},
        'engine': {'precision': precision},
    }, case='test/small')
    return hp
