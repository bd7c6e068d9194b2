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
                upsample_strides=list(UPSAMPLE_STRIDES))


def variable_shapes(hp):
    """OrderedDict name -> shape, in the order the reference's graph creates the variables."""
    d = model_dims(hp)
    k, R, D, S, Cc = d['k'], d['R'], d['D'], d['S'], d['Cc']
    shapes = OrderedDict()
    if d['cond_upsample'] == 'transposed_conv':     # reference models.py:109-124: [1, stride, Cc (out), Cin]
        cin = d['n_mels']
        for i, stride in enumerate(d['upsample_strides']):
            shapes[f'{ROOT}/cond/transposed_conv_{i}_weights'] = (1, stride, Cc, cin)
            cin = Cc
    else:
        shapes[f'{ROOT}/cond/dense'] = (1, d['n_mels'], Cc)
    for i in range(d['n_iaf']):
        for body in BODIES:
            p = f'{ROOT}/iaf{i}/{body}'
            shapes[f'{p}/causal_layer/filter'] = (k, 1, R)
            for j, _ in enumerate(d['dilations'][i]):
                q = f'{p}/dilated_stack/layer{j}'
                shapes[f'{q}/filter'] = (k, R, D)
                shapes[f'{q}/gate'] = (k, R, D)
                shapes[f'{q}/gc_filter'] = (1, Cc, D)
                shapes[f'{q}/gc_gate'] = (1, Cc, D)
                if d['use_biases']:
                    shapes[f'{q}/filter_bias'] = (D,)
                    shapes[f'{q}/gate_bias'] = (D,)
                shapes[f'{q}/dense'] = (1, D, R)
                shapes[f'{q}/skip'] = (1, D, S)
                if d['use_biases']:
                    shapes[f'{q}/dense_bias'] = (R,)
                    shapes[f'{q}/skip_bias'] = (S,)
            q = f'{p}/postprocessing'
            shapes[f'{q}/postprocess1'] = (1, S, S)
            if d['use_biases']:
                shapes[f'{q}/postprocess1_bias'] = (S,)
            shapes[f'{q}/postprocess2'] = (1, S, 1)
            if d['use_biases']:
                shapes[f'{q}/postprocess2_bias'] = (1,)
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
        for name, value in raw.items():
            if name.endswith(EMA_SUFFIX):
                out[name[:-len(EMA_SUFFIX)]] = value
    return out


def check_weights(hp, weights):
    shapes = variable_shapes(hp)
    missing = [n for n in shapes if n not in weights]
    if missing:
        raise KeyError(f'{len(missing)} variables missing, first: {missing[0]}')
    for name, shape in shapes.items():
        if tuple(weights[name].shape) != tuple(shape):
            raise ValueError(f'{name}: shape {tuple(weights[name].shape)} != {shape}')
