"""Drop-in for the reference's `models` module on the generation path:
`from models import IAFVocoder` (reference generate.py:13). Backed by hand-written sm_100a
kernels behind the C-ABI of include/pwv.h; see parallel-wavenet-vocoder_b200/vocoder.py."""
import importlib as _importlib

_impl = _importlib.import_module('parallel-wavenet-vocoder_b200.vocoder')
IAFVocoder = _impl.IAFVocoder
