"""ctypes binding of the C-ABI in include/pwv.h (libpwv_b200.so, built in-tree by
`__graft_entry__.build()` / `csrc/Makefile`).

There is deliberately no fallback: if the library is missing or fails to load, importing the
product path raises, and every compute entry point fails without a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PWV_LIB') or os.path.join(_HERE, 'libpwv_b200.so')   # PWV_LIB: A/B builds

PWV_MAX_FLOWS = 8
PWV_MAX_LAYERS = 64
PWV_MAX_UPSAMPLE = 4
UPSAMPLE = {'repeat': 0, 'transposed_conv': 1}      # any other method: 2 = no conditioning (reference models.py:134-135)
PREC = {'fp32': 0, 'f16x3': 1, 'bf16': 2}

EXPORTS = (
    'pwv_version', 'pwv_last_error', 'pwv_device_count', 'pwv_model_create', 'pwv_model_destroy',
    'pwv_model_num_variables', 'pwv_model_variable', 'pwv_model_load_weight', 'pwv_model_finalize',
    'pwv_workspace_bytes', 'pwv_forward', 'pwv_forward_host', 'pwv_last_launch_count',
    'pwv_set_profiling', 'pwv_profile_read', 'pwv_debug_set_trace', 'pwv_debug_set',
    'pwv_melspec_create', 'pwv_melspec_destroy', 'pwv_melspec_forward',
)


class PwvHparams(ctypes.Structure):
    _fields_ = [
        ('n_iaf', ctypes.c_int32), ('filter_width', ctypes.c_int32),
        ('residual_channels', ctypes.c_int32), ('dilation_channels', ctypes.c_int32),
        ('skip_channels', ctypes.c_int32), ('condition_channels', ctypes.c_int32),
        ('n_mels', ctypes.c_int32), ('hop_length', ctypes.c_int32),
        ('use_biases', ctypes.c_int32), ('use_skip_connection', ctypes.c_int32),
        ('precision', ctypes.c_int32),
        ('n_layers', ctypes.c_int32 * PWV_MAX_FLOWS),
        ('dilations', (ctypes.c_int32 * PWV_MAX_LAYERS) * PWV_MAX_FLOWS),
        ('cond_upsample', ctypes.c_int32), ('n_upsample', ctypes.c_int32),
        ('upsample_strides', ctypes.c_int32 * PWV_MAX_UPSAMPLE),
        ('normalize', ctypes.c_int32), ('normalize_cond', ctypes.c_int32), ('normalize_wavenet', ctypes.c_int32),
    ]


class PwvMelConfig(ctypes.Structure):
    _fields_ = [('n_fft', ctypes.c_int32), ('win_length', ctypes.c_int32), ('hop_length', ctypes.c_int32), ('n_mels', ctypes.c_int32),
                ('min_db', ctypes.c_float), ('max_db', ctypes.c_float), ('normalise', ctypes.c_int32)]


class PwvTaps(ctypes.Structure):
    _fields_ = [
        ('flow_out', ctypes.c_void_p),
        ('layer_flow', ctypes.c_int32), ('layer_body', ctypes.c_int32), ('layer_index', ctypes.c_int32),
        ('layer_out', ctypes.c_void_p),
        ('scale_shift', ctypes.c_void_p),
    ]


class PwvError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'pwv error {code}: {message}')
        self.code = code


_lib = None


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `make -C {os.path.join(_HERE, "csrc")}`). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    lib.pwv_version.restype = c.c_int
    lib.pwv_last_error.restype = c.c_char_p
    lib.pwv_device_count.restype = c.c_int
    lib.pwv_model_create.argtypes = [c.POINTER(PwvHparams), c.POINTER(c.c_void_p)]
    lib.pwv_model_destroy.argtypes = [c.c_void_p]
    lib.pwv_model_num_variables.argtypes = [c.c_void_p]
    lib.pwv_model_variable.argtypes = [c.c_void_p, c.c_int, c.POINTER(c.c_char_p), c.POINTER(c.c_int64), c.POINTER(c.c_int)]
    lib.pwv_model_load_weight.argtypes = [c.c_void_p, c.c_char_p, c.c_void_p, c.POINTER(c.c_int64), c.c_int]
    lib.pwv_model_finalize.argtypes = [c.c_void_p]
    lib.pwv_workspace_bytes.argtypes = [c.c_void_p, c.c_int, c.c_int, c.POINTER(c.c_size_t)]
    lib.pwv_forward.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t,
                                c.c_int, c.c_int, c.c_void_p, c.POINTER(PwvTaps)]
    lib.pwv_forward_host.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    lib.pwv_last_launch_count.argtypes = [c.c_void_p]
    lib.pwv_set_profiling.argtypes = [c.c_void_p, c.c_int]
    lib.pwv_debug_set_trace.argtypes = [c.c_void_p, c.c_void_p, c.c_int]
    lib.pwv_debug_set.argtypes = [c.c_void_p, c.c_char_p, c.c_int]
    lib.pwv_melspec_create.argtypes = [c.POINTER(PwvMelConfig), c.c_void_p, c.POINTER(c.c_void_p)]
    lib.pwv_melspec_destroy.argtypes = [c.c_void_p]
    lib.pwv_melspec_forward.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    lib.pwv_profile_read.argtypes = [c.c_void_p, c.POINTER(c.c_double), c.POINTER(c.c_int), c.POINTER(c.c_double)]
    for name in EXPORTS:
        if name not in ('pwv_last_error',):
            getattr(lib, name).restype = c.c_int
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise PwvError(rc, load().pwv_last_error().decode('utf-8', 'replace'))
    return rc


def make_hparams(dims, precision='fp32'):
    """`dims` is weights.model_dims(hp)."""
    if precision not in PREC:
        raise ValueError(f'engine.precision must be one of {sorted(PREC)}, got {precision!r}')
    h = PwvHparams()
    h.n_iaf = dims['n_iaf']
    h.filter_width = dims['k']
    h.residual_channels = dims['R']
    h.dilation_channels = dims['D']
    h.skip_channels = dims['S']
    h.condition_channels = dims['Cc']
    h.n_mels = dims['n_mels']
    h.hop_length = dims['hop']
    h.use_biases = int(dims['use_biases'])
    h.use_skip_connection = int(dims['use_skip'])
    h.precision = PREC[precision]
    method = dims.get('cond_upsample', 'repeat')
    h.cond_upsample = UPSAMPLE.get(method, 2)
    strides = list(dims.get('upsample_strides', ())) if method == 'transposed_conv' else []
    if len(strides) > PWV_MAX_UPSAMPLE:
        raise ValueError(f'{len(strides)} upsample stages > {PWV_MAX_UPSAMPLE}')
    h.n_upsample = len(strides)
    for i, st in enumerate(strides):
        h.upsample_strides[i] = int(st)
    for field, key in (('normalize', 'norm_flow'), ('normalize_cond', 'norm_cond'), ('normalize_wavenet', 'norm_wavenet')):
        method = dims.get(key, '') or ''
        if method not in ('', 'in'):
            raise NotImplementedError(f"normaliser {method!r}: '' and 'in' (reference modules.py:274-284) are implemented; "
                                      f"'bn' is tf.layers.batch_normalization")
        setattr(h, field, 1 if method == 'in' else 0)
    if dims['n_iaf'] > PWV_MAX_FLOWS:
        raise ValueError(f'n_iaf {dims["n_iaf"]} > {PWV_MAX_FLOWS}')
    for i, dil in enumerate(dims['dilations']):
        if len(dil) > PWV_MAX_LAYERS:
            raise ValueError(f'flow {i}: {len(dil)} layers > {PWV_MAX_LAYERS}')
        h.n_layers[i] = len(dil)
        for j, d in enumerate(dil):
            h.dilations[i][j] = int(d)
    return h
