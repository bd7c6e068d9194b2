"""CPU tests of the oracle: against the golden vectors produced by the reference's own graph code,
against the literal time_to_batch replay, and the structural properties of the path."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden, pkg, small_case
from oracle import iaf_oracle as O


@pytest.mark.parametrize('name', ['ref_small.npz', 'ref_flows.npz', 'ref_skip.npz', 'ref_tran.npz', 'ref_norm.npz', 'ref_norm_tran.npz', 'ref_nocond.npz', 'ref_shapes.npz'])
def test_oracle_reproduces_reference_golden(hp, name):
    weights, noise, mel, wav, dil = load_golden(hp, name)
    skip = bool(hp.model.use_skip_connection)        # ref_skip / ref_shapes were generated with model.use_skip_connection=True
    if name == 'ref_small.npz':          # keep the CPU suite quick: 1 of the 2 utterances
        noise, mel, wav = noise[:1], mel[:1], wav[:1]
    got = O.iaf_vocoder_forward(noise, mel, weights, dil, 80, use_skip_connection=skip, dtype=np.float64)
    assert np.abs(got - wav).max() < 1e-10
    got32 = O.iaf_vocoder_forward(noise, mel, weights, dil, 80, use_skip_connection=skip, dtype=np.float32)
    assert np.abs(got32 - wav).max() < 1e-4 * max(1.0, np.abs(wav).max())


def test_variable_list_matches_reference_graph(hp):
    """Names and creation order of the 1241 variables the reference's default graph creates."""
    with open(os.path.join(ROOT, 'tests', 'golden', 'ref_varlist.txt')) as fh:
        ref = fh.read().split()
    W = pkg('weights')
    assert list(W.variable_shapes(hp).keys()) == ref
    assert W.count_parameters(hp) == 4848392


@pytest.mark.parametrize('n,t,d,k', [(3, 37, 1, 2), (3, 37, 2, 2), (2, 100, 8, 2), (3, 64, 16, 2), (2, 50, 64, 2),
                                     (1, 1000, 512, 2), (2, 7, 3, 2), (2, 40, 4, 3)])
def test_direct_conv_equals_literal_time_to_batch(n, t, d, k):
    rng = np.random.RandomState(n * 1000 + t + d)
    x, w = rng.randn(n, t, 5), rng.randn(k, 5, 4)
    assert np.abs(O.causal_conv_literal(x, w, d) - O.causal_conv(x, w, d)).max() < 1e-12


def test_upsample_repeat_and_crop():
    rng = np.random.RandomState(0)
    for hop, t in [(80, 4000), (80, 16000), (80, 96000), (5, 35)]:
        t_mel = 1 + t // hop
        mel = rng.randn(2, t_mel, 3)
        w = rng.randn(1, 3, 4)
        cond = O.upsample_cond_repeat(mel, w, hop)
        assert cond.shape == (2, t, 4)
        proj = np.maximum(mel @ w[0], 0)
        s = np.arange(t)
        assert np.array_equal(cond, proj[:, (s + hop // 2) // hop, :])


def test_causality_and_batch_independence(hp):
    small_case(hp, dilations=((1, 2, 4, 8), (1, 16)), n=3, t=800)
    weights = pkg('weights').init_weights(hp, seed=1, bias_std=0.1)
    noise, mel = O.synthetic_inputs(3, 800, 80, 80)
    dil = hp.model.dilations
    y = O.iaf_vocoder_forward(noise, mel, weights, dil, 80)
    assert np.array_equal(O.iaf_vocoder_forward(noise[1:2], mel[1:2], weights, dil, 80)[0], y[1])
    t0 = 500
    noise2, mel2 = noise.copy(), mel.copy()
    noise2[:, t0:] += 1.0
    mel2[:, (t0 + 40) // 80 + 1:, :] *= -1.0
    y2 = O.iaf_vocoder_forward(noise2, mel2, weights, dil, 80)
    assert np.array_equal(y2[:, :t0], y[:, :t0]) and not np.array_equal(y2[:, t0:], y[:, t0:])


def test_length_must_be_multiple_of_hop(hp):
    small_case(hp)
    weights = pkg('weights').init_weights(hp, seed=1)
    noise = np.zeros((1, 120), np.float32)
    mel = np.zeros((1, 2, 80), np.float32)
    with pytest.raises(ValueError):
        O.iaf_vocoder_forward(noise, mel, weights, hp.model.dilations, 80)


def test_instance_norm_variables_on_the_product_path(hp):
    """With every normaliser set to 'in' the weight container lists the reference's beta/gamma variables at each call
    site (order pinned by the fixture generator against the reference's own graph code), the oracle normalises per
    utterance and channel over time, and the C-ABI declares the same variable list ('bn' is refused, not ignored)."""
    small_case(hp, dilations=((1, 2), (4,)), n=2, t=160, precision='fp32')
    hp.model.normalize = hp.model.normalize_cond = hp.model.normalize_wavenet = 'in'
    W = pkg('weights')
    names = list(W.variable_shapes(hp).keys())
    assert 'iaf_vocoder/cond/normalize/normalize/beta' in names and 'iaf_vocoder/normalize1/gamma' in names
    assert names.index('iaf_vocoder/iaf0/scalar/dilated_stack/layer0/normalize_filter/beta') == \
        names.index('iaf_vocoder/iaf0/scalar/dilated_stack/layer0/gate_bias') + 1
    weights = W.init_weights(hp, seed=3, bias_std=0.1)
    noise, mel = O.synthetic_inputs(2, 160, 80, 80)
    out = O.iaf_vocoder_forward(noise, mel, weights, [[1, 2], [4]], 80, dtype=np.float64)
    # the last flow's output went through normalize1: per utterance mean = beta, std = |gamma| (up to the 1e-8 epsilon)
    beta, gamma = float(weights['iaf_vocoder/normalize1/beta'][0]), float(weights['iaf_vocoder/normalize1/gamma'][0])
    assert np.allclose(out.mean(axis=1), beta, atol=1e-9) and np.allclose(out.std(axis=1), abs(gamma), rtol=1e-6)
    pkg('vocoder')._assert_supported(hp)              # 'in' is on the B200 path (un-fused fp32 kernels, csrc/pwv_norm.cuh) ...
    assert pkg('vocoder').resolve_precision(W.model_dims(hp), 'fp32') == 'fp32'
    # ... and the C-ABI lists the same variables in the same order
    import ctypes
    L = pkg('_lib')
    lib = L.load()
    h = ctypes.c_void_p()
    hparams = L.make_hparams(W.model_dims(hp), 'fp32')
    assert lib.pwv_model_create(ctypes.byref(hparams), ctypes.byref(h)) == 0
    name, shape, ndim = ctypes.c_char_p(), (ctypes.c_int64 * 4)(), ctypes.c_int()
    got = []
    for i in range(lib.pwv_model_num_variables(h)):
        lib.pwv_model_variable(h, i, ctypes.byref(name), shape, ctypes.byref(ndim))
        got.append(name.value.decode())
    assert got == names
    lib.pwv_model_destroy(h)
    hparams = L.make_hparams(W.model_dims(hp), 'f16x3')           # a statistic over the whole time axis: fp32 kernels only
    assert lib.pwv_model_create(ctypes.byref(hparams), ctypes.byref(h)) == -1 and b'normalisers' in lib.pwv_last_error()
    hp.model.normalize = 'bn'                         # 'bn' (tf.layers.batch_normalization) stays unimplemented
    with pytest.raises(NotImplementedError):
        pkg('vocoder')._assert_supported(hp)
    with pytest.raises(NotImplementedError):
        W.variable_shapes(hp)
