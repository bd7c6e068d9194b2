"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN `models.py` / `modules.py`.

Runs only where /root/reference exists (the build container). TensorFlow 1.x is not installable, so
the reference modules are imported under `oracle/tf_shim` (a numpy stand-in for the TF ops they
call); hparams come from the reference's own `hparams/default.yaml` through the reference's own
`hparam.py` (with `yaml.load_all` given the safe loader PyYAML >= 6 requires). What the fixtures
pin: graph wiring, variable names/shapes/creation order, op order, crop arithmetic -- everything
in the reference's Python. What they do not pin: TensorFlow's kernels (restated by the shim).

    python tests/golden/make_golden_from_reference.py

writes tests/golden/ref_shapes.npz (filter_width 3, R/D/S = 24/40/56, skip sum), tests/golden/ref_norm.npz / ref_norm_tran.npz (all normalisers 'in'), tests/golden/ref_tran.npz (cond_upsample_method 'transposed_conv'), tests/golden/ref_small.npz (default hparams, N=2, T=1600, fp64 arithmetic on float32-valued
inputs and weights, non-zero biases), tests/golden/ref_flows.npz (a 2-flow graph with per-flow
outputs) and tests/golden/ref_varlist.txt (the graph's variable names in creation order).
"""
import importlib
import os
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
PKG = 'parallel-wavenet-vocoder_b200'


def import_reference():
    """Import the reference's hparam/models modules under the TF stand-in."""
    # search order: the reference first (its hparam.py / models.py / modules.py must win over this
    # repo's root-level drop-ins of the same names), then the TF stand-in, then this repo
    for path in (ROOT, os.path.join(ROOT, 'oracle', 'tf_shim'), REF):
        if path in sys.path:
            sys.path.remove(path)
        sys.path.insert(0, path)
    import tensorflow as tf                      # the shim
    real_load_all = yaml.load_all
    yaml.load_all = lambda stream, Loader=None: real_load_all(stream, Loader=yaml.SafeLoader)
    cwd = os.getcwd()
    os.chdir(REF)                                # reference hparam.py:56 opens 'hparams/default.yaml'
    try:
        # our repo root also has hparam.py / models.py shims: make sure the reference's win
        for name in ('hparam', 'models', 'modules'):
            sys.modules.pop(name, None)
        ref_hparam = importlib.import_module('hparam')
        assert os.path.dirname(os.path.abspath(ref_hparam.__file__)) == REF, ref_hparam.__file__
        ref_hparam.hparam.set_hparam_yaml('default')
        ref_models = importlib.import_module('models')
        assert os.path.dirname(os.path.abspath(ref_models.__file__)) == REF, ref_models.__file__
    finally:
        os.chdir(cwd)
        yaml.load_all = real_load_all
    return tf, ref_hparam.hparam, ref_models


def run_reference(tf, ref_models, weights, noise, mel):
    """Execute the reference graph code in float64 on the given (float32-valued) tensors."""
    tf.set_variables({k: np.asarray(v, dtype=np.float64) for k, v in weights.items()})
    tf.set_logistic_sample(np.asarray(noise, dtype=np.float64))
    n, t = noise.shape
    model = ref_models.IAFVocoder(batch_size=n, length=t)
    wav = model(None, np.asarray(mel, dtype=np.float64), is_training=False)   # reference models.py:23
    return np.asarray(wav)[:, :, 0], tf.created_variables()


def pack(noise, mel, wav, weights, dilations, recipe):
    """Weights are NOT stored (19 MB): they are regenerated from `recipe` = (seed, bias_std, gain) with
    weights.init_weights (numpy's frozen legacy generator); a checksum guards against drift."""
    flat = np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in weights.values()])
    return {'noise': noise, 'mel': mel, 'wav': wav, 'n_iaf': np.int64(len(dilations)),
            'dilations': np.array([list(d) + [0] * (64 - len(d)) for d in dilations], dtype=np.int64),
            'n_layers': np.array([len(d) for d in dilations], dtype=np.int64),
            'weight_recipe': np.array(recipe, dtype=np.float64),
            'weight_checksum': np.array([flat.sum(), np.abs(flat).sum(), flat[::9973].sum()])}


def main():
    tf, ref_hp, ref_models = import_reference()
    my_hp = importlib.import_module(PKG + '.hparam').hparam
    W = importlib.import_module(PKG + '.weights')
    from oracle import iaf_oracle as O
    hop, n_mels = ref_hp.signal.hop_length, ref_hp.signal.n_mels

    # ---- fixture 1: the reference's default graph
    my_hp.set_hparam_yaml('default')
    n, t = 2, 1600
    weights = W.init_weights(my_hp, seed=42, bias_std=0.1, dtype=np.float32)
    noise, mel = O.synthetic_inputs(n, t, hop, n_mels, mel_seed=101, noise_seed=102, dtype=np.float32)
    wav, created = run_reference(tf, ref_models, weights, noise, mel)
    assert created == list(W.variable_shapes(my_hp).keys()), 'variable list / creation order differs'
    ours = O.iaf_vocoder_forward(noise, mel, weights, ref_hp.model.dilations, hop, dtype=np.float64)
    err = np.abs(ours - wav).max()
    print('default graph: reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (err, np.abs(wav).max()))
    assert err < 1e-12
    np.savez_compressed(os.path.join(HERE, 'ref_small.npz'), **pack(noise, mel, wav, weights, ref_hp.model.dilations, (42, 0.1, 1.0)))
    with open(os.path.join(HERE, 'ref_varlist.txt'), 'w') as fh:
        fh.write('\n'.join(created) + '\n')

    # ---- fixture 2: a short 2-flow graph, larger kernels, T not a multiple of 64, d >= T tap
    dil = [[1, 2, 4, 512], [3, 1, 256]]
    ref_hp.model.dilations = dil
    ref_hp.model.n_iaf = 2
    my_hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': dil}}, case='golden/flows')
    n, t = 3, 400
    weights = W.init_weights(my_hp, seed=43, bias_std=0.2, gain=2.0, dtype=np.float32)
    noise, mel = O.synthetic_inputs(n, t, hop, n_mels, mel_seed=103, noise_seed=104, dtype=np.float32)
    wav, created = run_reference(tf, ref_models, weights, noise, mel)
    assert created == list(W.variable_shapes(my_hp).keys())
    taps = {}
    ours = O.iaf_vocoder_forward(noise, mel, weights, dil, hop, dtype=np.float64, taps=taps)
    err = np.abs(ours - wav).max()
    print('2-flow graph : reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (err, np.abs(wav).max()))
    assert err < 1e-12
    np.savez_compressed(os.path.join(HERE, 'ref_flows.npz'), **pack(noise, mel, wav, weights, dil, (43, 0.2, 2.0)))

    # ---- fixture 3: the same 2-flow graph with use_skip_connection=True (reference modules.py:147)
    ref_hp.model.use_skip_connection = True
    weights = W.init_weights(my_hp, seed=44, bias_std=0.1, dtype=np.float32)
    wav, created = run_reference(tf, ref_models, weights, noise, mel)
    ours = O.iaf_vocoder_forward(noise, mel, weights, dil, hop, use_skip_connection=True, dtype=np.float64)
    err = np.abs(ours - wav).max()
    print('skip-sum graph: reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (err, np.abs(wav).max()))
    assert err < 1e-12
    np.savez_compressed(os.path.join(HERE, 'ref_skip.npz'), **pack(noise, mel, wav, weights, dil, (44, 0.1, 1.0)))
    ref_hp.model.use_skip_connection = False

    # ---- fixture 4: the same 2-flow graph with cond_upsample_method='transposed_conv' (reference models.py:109-124,
    #      the reference's own hparams case test/tran)
    ref_hp.model.cond_upsample_method = 'transposed_conv'
    my_hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': dil, 'cond_upsample_method': 'transposed_conv'}}, case='golden/tran')
    weights = W.init_weights(my_hp, seed=45, bias_std=0.1, dtype=np.float32)
    wav, created = run_reference(tf, ref_models, weights, noise, mel)
    assert created == list(W.variable_shapes(my_hp).keys()), 'variable list / creation order differs (transposed_conv)'
    ours = O.iaf_vocoder_forward(noise, mel, weights, dil, hop, dtype=np.float64)
    err = np.abs(ours - wav).max()
    print('transposed_conv graph: reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (err, np.abs(wav).max()))
    assert err < 1e-12
    d = pack(noise, mel, wav, weights, dil, (45, 0.1, 1.0))
    d['cond_upsample_method'] = np.array('transposed_conv')
    np.savez_compressed(os.path.join(HERE, 'ref_tran.npz'), **d)
    ref_hp.model.cond_upsample_method = 'repeat'

    # ---- fixtures 5, 6: every normaliser call site switched to 'in' (reference modules.py:263-284: instance
    #      normalisation over time, executed from the reference's own instance_normalization), with 'repeat' and with
    #      'transposed_conv' conditioning (where the reference normalises a 4-D tensor over its size-1 axis, so each
    #      stage's output collapses to beta -- replayed literally by the oracle). The B200 path rejects these options;
    #      the fixtures pin the oracle for them.
    norm = {'normalize': 'in', 'normalize_cond': 'in', 'normalize_wavenet': 'in'}
    for method, fname, seed in (('repeat', 'ref_norm.npz', 46), ('transposed_conv', 'ref_norm_tran.npz', 47)):
        for key, val in norm.items():
            setattr(ref_hp.model, key, val)
        ref_hp.model.cond_upsample_method = method
        my_hp.set_hparam_dict({'model': dict(norm, n_iaf=2, dilations=dil, cond_upsample_method=method)}, case='golden/' + fname)
        weights = W.init_weights(my_hp, seed=seed, bias_std=0.1, dtype=np.float32)
        wav, created = run_reference(tf, ref_models, weights, noise, mel)
        assert created == list(W.variable_shapes(my_hp).keys()), 'variable list / creation order differs (normalisers, %s)' % method
        ours = O.iaf_vocoder_forward(noise, mel, weights, dil, hop, dtype=np.float64)
        err = np.abs(ours - wav).max()
        print('instance-norm graph (%s): reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (method, err, np.abs(wav).max()))
        assert err < 1e-11
        d = pack(noise, mel, wav, weights, dil, (seed, 0.1, 1.0))
        d['cond_upsample_method'] = np.array(method)
        d['normalize'] = np.array('in')
        np.savez_compressed(os.path.join(HERE, fname), **d)
        for key in norm:
            setattr(ref_hp.model, key, '')
    ref_hp.model.cond_upsample_method = 'repeat'

    # ---- fixture 7: any other cond_upsample_method leaves the graph UNCONDITIONAL (reference models.py:134-135: cond = None,
    #      so no cond/* variable and no gc_filter / gc_gate in any layer, modules.py:216-222)
    ref_hp.model.cond_upsample_method = 'none'
    my_hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': dil, 'cond_upsample_method': 'none'}}, case='golden/nocond')
    weights = W.init_weights(my_hp, seed=48, bias_std=0.1, dtype=np.float32)
    wav, created = run_reference(tf, ref_models, weights, noise, mel)
    assert created == list(W.variable_shapes(my_hp).keys()), 'variable list / creation order differs (unconditional)'
    assert not any('/cond/' in v or '/gc_' in v for v in created)
    ours = O.iaf_vocoder_forward(noise, mel, weights, dil, hop, dtype=np.float64)
    err = np.abs(ours - wav).max()
    print('unconditional graph: reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (err, np.abs(wav).max()))
    assert err < 1e-12
    d = pack(noise, mel, wav, weights, dil, (48, 0.1, 1.0))
    d['cond_upsample_method'] = np.array('none')
    np.savez_compressed(os.path.join(HERE, 'ref_nocond.npz'), **d)
    ref_hp.model.cond_upsample_method = 'repeat'

    # ---- fixture 8: the free shape parameters (reference modules.py:210-244, hparams/default.yaml:22-26): filter_width 3,
    #      residual / dilation / skip channels all different and none a multiple of 16, skip sum on
    shapes = {'filter_width': 3, 'residual_channels': 24, 'dilation_channels': 40, 'skip_channels': 56, 'use_skip_connection': True}
    for key, value in shapes.items():
        setattr(ref_hp.model, key, value)
    my_hp.set_hparam_dict({'model': dict(shapes, n_iaf=2, dilations=dil)}, case='golden/shapes')
    weights = W.init_weights(my_hp, seed=49, bias_std=0.1, dtype=np.float32)
    wav, created = run_reference(tf, ref_models, weights, noise, mel)
    assert created == list(W.variable_shapes(my_hp).keys()), 'variable list / creation order differs (free shapes)'
    assert weights['iaf_vocoder/iaf0/scalar/dilated_stack/layer0/filter'].shape == (3, 24, 40)
    ours = O.iaf_vocoder_forward(noise, mel, weights, dil, hop, use_skip_connection=True, dtype=np.float64)
    err = np.abs(ours - wav).max()
    print('free-shape graph: reference code (under shim) vs oracle: max|delta| = %.3e, |wav|max = %.3f' % (err, np.abs(wav).max()))
    assert err < 1e-12
    d = pack(noise, mel, wav, weights, dil, (49, 0.1, 1.0))
    for key, value in shapes.items():
        d[key] = np.int64(value)
    np.savez_compressed(os.path.join(HERE, 'ref_shapes.npz'), **d)
    ref_hp.model.filter_width, ref_hp.model.residual_channels, ref_hp.model.dilation_channels = 2, 64, 64
    ref_hp.model.skip_channels, ref_hp.model.use_skip_connection = 128, False
    print('wrote ref_small.npz, ref_flows.npz, ref_varlist.txt (%d variables in the default graph)' % len(W.variable_shapes(my_hp.set_hparam_yaml('default'))))


if __name__ == '__main__':
    main()
