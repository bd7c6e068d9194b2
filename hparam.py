"""Drop-in for the reference's `hparam` module: `from hparam import hparam as hp`
(reference generate.py:12, models.py:12). The implementation lives in the package."""
import importlib as _importlib

_impl = _importlib.import_module('parallel-wavenet-vocoder_b200.hparam')
Hparam = _impl.Hparam
hparam = _impl.hparam
