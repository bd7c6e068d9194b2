"""The weight contract of the generation path: TF variable names, shapes and layouts.

Every tensor is keyed by the exact variable name the reference's graph creates and kept in
the reference's layout (conv kernels are `[k, Cin, Cout]`, reference modules.py:179,210-211,
217-221,239-244,152-163; models.py:128). Scopes: `iaf_vocoder/cond/dense` (models.py:24-25,128),
`iaf_vocoder/iaf{i}/{scalar|shifter}/...` (models.py:35,48,63 -- the scaler's scope really is
spelled 'scalar'), `causal_layer/filter` (modules.py:133,179), `dilated_stack/layer{j}/...`
(modules.py:137-139,210-248), `postprocessing/...` (modules.py:145,152-164).
"""
from collections import OrderedDict

import numpy as np

BODIES = ('scalar', 'shifter')      # body 0 = scaler (scope 'scalar'), body 1 = shifter
ROOT = 'iaf_vocoder'
EMA_SUFFIX = '/ExponentialMovingAverage'   # tf.train.ExponentialMovingAverage.average_name(v)
UPSAMPLE_STRIDES = (4, 4, 5)               # the constant IAFVocoder.__call__ passes (reference models.py:26)


def model_dims(hp):
    m = hp.model
    return dict(k=int(m.filter_width), R=int(m.residual_channels), D=int(m.dilation_channels),
                S=int(m.skip_channels), Cc=int(m.condition_channels), n_mels=int(hp.signal.n_mels),
                hop=int(hp.signal.hop_length), n_iaf=int(m.n_iaf),
                dilations=[list(map(int, d)) for d in m.dilations[:int(m.n_iaf)]],
                use_biases=bool(m.use_biases), use_skip=bool(m.use_skip_connection),
                cond_upsample=str(m.get('cond_upsample_method', 'repeat') or 'repeat'),
                upsample_strides=list(UPSAMPLE_STRIDES),
                # normalisers (reference modules.py:263-284): '' off; the weight container and the oracle know 'in'
                # (instance normalisation over time); the B200 path itself rejects every non-empty value (vocoder.py)
                norm_flow=str(m.get('normalize', '') or ''), norm_cond=str(m.get('normalize_cond', '') or ''),
                norm_wavenet=str(m.get('normalize_wavenet', '') or ''))


def variable_shapes(hp):
    """OrderedDict name -> shape, in the order the reference's graph creates the variables."""
    d = model_dims(hp)
    k, R, D, S, Cc = d['k'], d['R'], d['D'], d['S'], d['Cc']
    shapes = OrderedDict()
    for key in ('norm_flow', 'norm_cond', 'norm_wavenet'):
        if d[key] not in ('', 'in'):
            raise NotImplementedError(f"normaliser {d[key]!r}: only '' and 'in' have a variable layout here "
                                      f"(reference modules.py:263-284; 'bn' creates tf.layers variables)")

    def norm(scope, channels, on):      # instance_normalization's variables, beta before gamma (modules.py:279-280)
        if on:
            shapes[f'{scope}/beta'] = (channels,)
            shapes[f'{scope}/gamma'] = (channels,)
    nc, nw, nf = d['norm_cond'] == 'in', d['norm_wavenet'] == 'in', d['norm_flow'] == 'in'
    if d['cond_upsample'] == 'transposed_conv':     # reference models.py:109-124: [1, stride, Cc (out), Cin]
        cin = d['n_mels']
        for i, stride in enumerate(d['upsample_strides']):
            shapes[f'{ROOT}/cond/transposed_conv_{i}_weights'] = (1, stride, Cc, cin)
            norm(f'{ROOT}/cond/normalize_transposed_conv_{i}', Cc, nc)       # models.py:121-122
            cin = Cc
    elif d['cond_upsample'] == 'repeat':
        shapes[f'{ROOT}/cond/dense'] = (1, d['n_mels'], Cc)
    conditioned = d['cond_upsample'] in ('repeat', 'transposed_conv')        # anything else: cond = None (models.py:134-135)
    if not conditioned and nc:
        raise ValueError("normalize_cond with an unconditional graph: the reference fails too (normalize(None), models.py:27-29)")
    norm(f'{ROOT}/cond/normalize/normalize', Cc, nc)                         # models.py:27-29
    for i in range(d['n_iaf']):
        for body in BODIES:
            p = f'{ROOT}/iaf{i}/{body}'
            shapes[f'{p}/causal_layer/filter'] = (k, 1, R)
            norm(f'{p}/causal_layer/normalize', R, nw)                       # modules.py:181-182
            for j, _ in enumerate(d['dilations'][i]):
                q = f'{p}/dilated_stack/layer{j}'
                shapes[f'{q}/filter'] = (k, R, D)
                shapes[f'{q}/gate'] = (k, R, D)
                if conditioned:                                              # modules.py:216-222: only with a condition
                    shapes[f'{q}/gc_filter'] = (1, Cc, D)
                    shapes[f'{q}/gc_gate'] = (1, Cc, D)
                if d['use_biases']:
                    shapes[f'{q}/filter_bias'] = (D,)
                    shapes[f'{q}/gate_bias'] = (D,)
                norm(f'{q}/normalize_filter', D, nw)                         # modules.py:230-234
                norm(f'{q}/normalize_gate', D, nw)
                shapes[f'{q}/dense'] = (1, D, R)
                shapes[f'{q}/skip'] = (1, D, S)
                if d['use_biases']:
                    shapes[f'{q}/dense_bias'] = (R,)
                    shapes[f'{q}/skip_bias'] = (S,)
                norm(f'{q}/normalize_skip_output', S, nw)                    # modules.py:253-257
                norm(f'{q}/normalize_dense_output', R, nw)
            q = f'{p}/postprocessing'
            norm(f'{q}/normalize_postprocess1', S, nw)                       # modules.py:149-151
            shapes[f'{q}/postprocess1'] = (1, S, S)
            if d['use_biases']:
                shapes[f'{q}/postprocess1_bias'] = (S,)
            norm(f'{q}/normalize_postprocess2', S, nw)                       # modules.py:158-160
            shapes[f'{q}/postprocess2'] = (1, S, 1)
            if d['use_biases']:
                shapes[f'{q}/postprocess2_bias'] = (1,)
        norm(f'{ROOT}/normalize{i}', 1, nf)                                  # models.py:70
    return shapes


def count_parameters(hp):
    return int(sum(int(np.prod(s)) for s in variable_shapes(hp).values()))


def init_weights(hp, seed=0, bias_std=0.0, gain=1.0, dtype=np.float32):
    """Random weights shaped like a fresh reference graph.

    Conv kernels follow `tf.get_variable`'s default initializer (Glorot uniform: limit =
    sqrt(6 / (fan_in + fan_out)), fans = k*Cin and k*Cout); biases are zero as in the reference
    (`tf.zeros_initializer`, modules.py:155,164,225-226,247-248) unless `bias_std` > 0, which the
    parity tests use so that the bias paths carry signal. `gain` scales the kernels (stress tests).
    Drawn in float64 from numpy's frozen legacy generator so the same seed gives the same
    tensors everywhere, then cast once.
    """
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in variable_shapes(hp).items():
        if len(shape) == 1:
            w = rng.normal(0.0, bias_std, size=shape) if bias_std > 0 else np.zeros(shape)
            if name.endswith('/gamma'):        # instance-norm scale: tf.ones_initializer (modules.py:280), jittered like the biases
                w = w + 1.0
        elif len(shape) == 4:            # conv2d_transpose filter [1, width, Cout, Cin]
            _, width, cout, cin = shape
            limit = np.sqrt(6.0 / (width * cin + width * cout))
            w = rng.uniform(-limit, limit, size=shape) * gain
        else:
            k, cin, cout = shape
            limit = np.sqrt(6.0 / (k * cin + k * cout))
            w = rng.uniform(-limit, limit, size=shape) * gain
        out[name] = np.ascontiguousarray(w.astype(dtype))
    return out


def save_npz(path, weights):
    np.savez(path, **{k.replace('/', '|'): v for k, v in weights.items()})


def load_npz(path, use_ema=False):
    """Load a name->array container. With `use_ema`, `<name>/ExponentialMovingAverage` entries
    take precedence over `<name>` (the mapping reference generate.py:58-63 builds)."""
    raw = {k.replace('|', '/'): v for k, v in np.load(path).items()}
    out = OrderedDict()
    for name, value in raw.items():
        if name.endswith(EMA_SUFFIX):
            continue
        out[name] = value
    if use_ema:
        shadows = [name for name in raw if name.endswith(EMA_SUFFIX)]
        if out and not shadows:
            raise KeyError(f'{path}: train.use_ema is set but the container has no {EMA_SUFFIX} entries')
        for name in shadows:
            out[name[:-len(EMA_SUFFIX)]] = raw[name]
    return out


def check_weights(hp, weights):
    shapes = variable_shapes(hp)
    missing = [n for n in shapes if n not in weights]
    if missing:
        raise KeyError(f'{len(missing)} variables missing, first: {missing[0]}')
    for name, shape in shapes.items():
        if tuple(weights[name].shape) != tuple(shape):
            raise ValueError(f'{name}: shape {tuple(weights[name].shape)} != {shape}')
