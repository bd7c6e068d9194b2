"""The C-ABI library loads, exports every symbol include/pwv.h declares, validates hparams and
variable names on the host, and FAILS LOUDLY (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, pkg, small_case


def _declared_symbols():
    with open(os.path.join(ROOT, 'include', 'pwv.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pwv_[a-z_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    L = pkg('_lib')
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(L.EXPORTS) == declared
    assert lib.pwv_version() == 201


def test_hparams_struct_matches_header():
    L = pkg('_lib')
    # 11 scalars + 8 + 8*64 int32, then cond_upsample, n_upsample, upsample_strides[4], the three normaliser switches
    assert ctypes.sizeof(L.PwvHparams) == 4 * (11 + 8 + 8 * 64 + 2 + 4 + 3)


def _create(hp, precision='fp32'):
    L = pkg('_lib')
    lib = L.load()
    h = ctypes.c_void_p()
    hparams = L.make_hparams(pkg('weights').model_dims(hp), precision)
    rc = lib.pwv_model_create(ctypes.byref(hparams), ctypes.byref(h))
    return lib, h, rc


def test_model_create_validates(hp):
    L = pkg('_lib')
    lib, h, rc = _create(hp)
    assert rc == 0 and h.value
    assert lib.pwv_model_num_variables(h) == 1241
    name = ctypes.c_char_p()
    shape = (ctypes.c_int64 * 4)()
    ndim = ctypes.c_int()
    assert lib.pwv_model_variable(h, 0, ctypes.byref(name), shape, ctypes.byref(ndim)) == 0
    assert name.value == b'iaf_vocoder/cond/dense' and list(shape)[:3] == [1, 80, 80] and ndim.value == 3
    names = []
    for i in range(1241):
        lib.pwv_model_variable(h, i, ctypes.byref(name), shape, ctypes.byref(ndim))
        names.append(name.value.decode())
    assert names == list(pkg('weights').variable_shapes(hp).keys())
    lib.pwv_model_destroy(h)

    hp.model.filter_width = 3                           # any shape on the general fp32 chain, none of it on tensor cores
    _, h3, rc = _create(hp)
    assert rc == 0
    lib.pwv_model_variable(h3, 1, ctypes.byref(name), shape, ctypes.byref(ndim))
    assert name.value.endswith(b'causal_layer/filter') and list(shape)[:3] == [3, 1, 64]
    lib.pwv_model_destroy(h3)
    _, _, rc = _create(hp, 'f16x3')
    assert rc == -1 and b'general fp32 path' in lib.pwv_last_error()
    hp.model.filter_width = 0
    _, _, rc = _create(hp)
    assert rc == -1 and b'filter_width' in lib.pwv_last_error()
    hp.model.filter_width = 2
    hp.model.use_skip_connection = True                 # accepted by the fp32 and the tensor-core (plane) paths
    for prec in ('fp32', 'f16x3'):
        _, h2, rc = _create(hp, prec)
        assert rc == 0
        assert lib.pwv_debug_set(h2, b'path', 0) == 0 and lib.pwv_debug_set(h2, b'no_such_switch', 1) == -1
        lib.pwv_model_destroy(h2)
    hp.model.use_skip_connection = False
    hp.model.residual_channels = 48                     # R != D, S != 2R: general chain; bf16 / f16x3 refuse
    _, h4, rc = _create(hp)
    assert rc == 0
    lib.pwv_model_destroy(h4)
    _, _, rc = _create(hp, 'bf16')
    assert rc == -1
    hp.model.residual_channels = 0
    _, _, rc = _create(hp)
    assert rc == -1 and b'positive' in lib.pwv_last_error()
    with pytest.raises(ValueError):
        L.make_hparams(pkg('weights').model_dims(hp), 'fp16')


def test_load_weight_checks_names_and_shapes(hp):
    small_case(hp)
    lib, h, rc = _create(hp)
    assert rc == 0
    w = np.zeros((1, 80, 80), np.float32)
    shp = (ctypes.c_int64 * 3)(1, 80, 80)
    assert lib.pwv_model_load_weight(h, b'iaf_vocoder/cond/dense', w.ctypes.data_as(ctypes.c_void_p), shp, 3) == 0
    assert lib.pwv_model_load_weight(h, b'iaf_vocoder/cond/nope', w.ctypes.data_as(ctypes.c_void_p), shp, 3) == -5
    bad = (ctypes.c_int64 * 3)(1, 80, 81)
    assert lib.pwv_model_load_weight(h, b'iaf_vocoder/cond/dense', w.ctypes.data_as(ctypes.c_void_p), bad, 3) == -5
    # finalize before everything is loaded is a state error naming the first missing variable
    assert lib.pwv_model_finalize(h) == -2 and b'was not loaded' in lib.pwv_last_error()
    lib.pwv_model_destroy(h)


def test_workspace_query_and_shape_errors(hp):
    lib, h, rc = _create(hp)
    out = ctypes.c_size_t()
    assert lib.pwv_workspace_bytes(h, 8, 16000, ctypes.byref(out)) == 0
    # two ping/pong activation buffers of [2][N][T][64] fp32 dominate
    assert out.value >= 2 * 2 * 8 * 16000 * 64 * 4
    assert lib.pwv_workspace_bytes(h, 8, 16001, ctypes.byref(out)) == -1
    assert b'hop_length' in lib.pwv_last_error()
    lib.pwv_model_destroy(h)


def test_no_cpu_fallback(hp):
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    small_case(hp)
    L, V, W = pkg('_lib'), pkg('vocoder'), pkg('weights')
    assert L.load().pwv_device_count() < 0
    with pytest.raises(L.PwvError) as e:
        V.PwvModel(W.model_dims(hp), W.init_weights(hp, seed=0))
    assert e.value.code == -3
    with pytest.raises(Exception):
        V.IAFVocoder(batch_size=1, length=1600)


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg_dir = os.path.join(ROOT, 'parallel-wavenet-vocoder_b200')
    offenders = []
    for base in [pkg_dir] + [os.path.join(ROOT, f) for f in ('generate.py', 'generate_multi.py', 'models.py', 'hparam.py')]:
        paths = [base] if os.path.isfile(base) else [os.path.join(d, f) for d, _, fs in os.walk(base) for f in fs if f.endswith(('.py', '.cu', '.cuh'))]
        for p in paths:
            if os.path.exists(p) and re.search(r'^\s*(from|import)\s+oracle\b|oracle[./]iaf_oracle', open(p).read(), flags=re.M):
                offenders.append(p)
    assert not offenders, offenders


def test_plain_c_consumer_builds_and_runs(tmp_path):
    """include/pwv.h compiles as C (gcc -std=c99 -pedantic) and a C program drives the library."""
    import subprocess
    L = pkg('_lib')
    L.load()
    exe = str(tmp_path / 'c_abi_smoke')
    src = os.path.join(ROOT, 'tests', 'c_abi_smoke.c')
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-o', exe, src, L.LIB_PATH,
                    '-Wl,-rpath,' + os.path.dirname(L.LIB_PATH)], check=True)
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
    assert 'c_abi_smoke ok' in out.stdout


def test_unconditional_graph_variable_list(hp):
    """cond_upsample_method outside 'repeat' / 'transposed_conv' leaves the graph unconditional (reference
    models.py:134-135, modules.py:216-222): no cond/* variable, no gc_filter / gc_gate -- in the weight container and
    in the C-ABI's variable list alike; normalize_cond cannot apply."""
    small_case(hp)
    hp.model.cond_upsample_method = 'none'
    W, L = pkg('weights'), pkg('_lib')
    names = list(W.variable_shapes(hp).keys())
    assert not any('/cond/' in n or '/gc_' in n for n in names)
    lib, h, rc = _create(hp)
    assert rc == 0
    name, shape, ndim = ctypes.c_char_p(), (ctypes.c_int64 * 4)(), ctypes.c_int()
    got = []
    for i in range(lib.pwv_model_num_variables(h)):
        lib.pwv_model_variable(h, i, ctypes.byref(name), shape, ctypes.byref(ndim))
        got.append(name.value.decode())
    assert got == names
    lib.pwv_model_destroy(h)
    hp.model.normalize_cond = 'in'
    with pytest.raises(ValueError):
        W.variable_shapes(hp)
    with pytest.raises(ValueError):
        pkg('vocoder')._assert_supported(hp)


def test_batch_split_arithmetic_of_the_host_side():
    """PwvModel._utterances_per_pass: how many utterances one pwv_forward call takes when the workspace is bounded
    (host logic only; the GPU suite checks that the split result is bit-identical)."""
    V, L = pkg('vocoder'), pkg('_lib')
    m = V.PwvModel.__new__(V.PwvModel)           # no library handle: the method only asks workspace_bytes
    m._ws, m.max_workspace_bytes = None, 1000
    m.workspace_bytes = lambda k, t: 100 + 300 * k
    assert m._utterances_per_pass(5, 10, 'cpu') == 3        # 100 + 900 fits exactly, 1300 does not
    assert m._utterances_per_pass(2, 10, 'cpu') == 2
    m.max_workspace_bytes = 399
    with pytest.raises(MemoryError, match='one utterance'):
        m._utterances_per_pass(5, 10, 'cpu')

    def needs_more_than_the_device(k, t):        # what pwv_workspace_bytes reports past the device's size
        if k > 2:
            raise L.PwvError(-4, 'workspace exceeds the device')
        return 400 * k
    m.workspace_bytes, m.max_workspace_bytes = needs_more_than_the_device, 10_000
    assert m._utterances_per_pass(64, 10, 'cpu') == 2
    m.close = lambda: None
