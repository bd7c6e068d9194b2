def stft(*a, **k):     # imported at module scope by reference modules.py:8; only losses call it
    raise NotImplementedError('tf_shim: stft is training-only')
