"""CPU ORACLE of the IAF-vocoder generation path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement, op for op and in the reference's order of operations, of the forward pass
`generate.py` runs (andabi/parallel-wavenet-vocoder @ 6c2fa069). Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module; the product
path (`parallel-wavenet-vocoder_b200/`) never does and fails loudly without its CUDA library.

PARITY PINNING. The reference ships no tests, golden vectors, checkpoints or audio, and its
arithmetic lives in TensorFlow 1.x (`tensorflow >= 1.4`, unpinned, requirements.txt:1), which is
not installable here. The oracle is therefore pinned two ways, both weaker than running TF:
  (1) `tests/golden/make_golden_from_reference.py` imports the reference's OWN `modules.py` /
      `models.py` under a numpy stand-in for the dozen TF ops they call (`oracle/tf_shim`) and
      stores inputs + outputs as fixtures; the oracle must reproduce them (tests/test_oracle.py).
      This pins the model wiring (what is connected to what, in which order, under which names)
      to the reference's source, but NOT TensorFlow's kernels -- the shim restates those.
  (2) `causal_conv_literal` below replays the reference's time_to_batch / pad / VALID-conv /
      batch_to_time / slice sequence literally and is checked against the direct x[t-d] form.
So: parity is pinned to the reference's graph code, unpinned with respect to TF's own kernels.

All functions take/return numpy arrays; `dtype` float64 is ground truth, float32 stands in for
"TF on CPU in fp32". Layouts are the reference's: activations (N, T, C), kernels [k, Cin, Cout].
"""
import numpy as np

ROOT = 'iaf_vocoder'
UPSAMPLE_STRIDES = (4, 4, 5)     # the constant reference models.py:26 passes to _upsample_cond


# ----------------------------------------------------------------------------- array backends
class NumpyOps:
    """Default backend: numpy (float64 ground truth / float32)."""
    name = 'numpy'
    tanh = staticmethod(np.tanh)
    exp = staticmethod(np.exp)
    zeros_like = staticmethod(np.zeros_like)

    @staticmethod
    def relu(x):
        return np.maximum(x, 0)

    @staticmethod
    def asarray(x, dtype):
        return np.asarray(x, dtype=dtype)

    @staticmethod
    def tile_last(x, reps):
        return np.tile(x, [1, 1, reps])

    @staticmethod
    def to_numpy(x):
        return x

    @staticmethod
    def moments_time(x):
        return x.mean(axis=1, keepdims=True), x.var(axis=1, keepdims=True)


class TorchOps:
    """Same restatement evaluated by torch's CPU kernels (MKL matmul, threaded elementwise): the
    kernel family TF-CPU would use. Only bench.py's CPU-baseline legs use it, for a fair timing."""
    name = 'torch-cpu'

    def __init__(self):
        import torch
        self.t = torch
        self.tanh, self.exp, self.zeros_like, self.relu = torch.tanh, torch.exp, torch.zeros_like, torch.relu

    def asarray(self, x, dtype):
        return self.t.as_tensor(np.asarray(x, dtype=dtype))

    def tile_last(self, x, reps):
        return x.repeat(1, 1, reps)

    @staticmethod
    def to_numpy(x):
        return x.numpy()

    @staticmethod
    def moments_time(x):
        return x.mean(dim=1, keepdim=True), x.var(dim=1, keepdim=True, unbiased=False)


_OPS = NumpyOps()


# ----------------------------------------------------------------------------- causal_conv
def conv1d_valid(x, w):
    """tf.nn.conv1d(x, w, stride=1, padding='VALID'): cross-correlation
    y[n, t, co] = sum_j sum_ci x[n, t + j, ci] * w[j, ci, co]   (reference modules.py:33,42)."""
    k = w.shape[0]
    t_out = x.shape[1] - k + 1
    y = x[:, 0:t_out, :] @ w[0]
    for j in range(1, k):
        y = y + x[:, j:j + t_out, :] @ w[j]
    return y


def causal_conv_literal(x, w, dilation):
    """Literal replay of reference modules.py:11-43 (time_to_batch -> left pad k-1 -> VALID conv ->
    batch_to_time -> slice). Used only to validate `causal_conv`."""
    n, t, c = x.shape
    k = w.shape[0]
    if dilation > 1:
        pad_elements = dilation - 1 - (t + dilation - 1) % dilation          # modules.py:14
        padded = np.pad(x, [(0, 0), (0, pad_elements), (0, 0)])              # modules.py:15
        reshaped = padded.reshape(-1, dilation, c)                           # modules.py:16
        transposed = reshaped.transpose(1, 0, 2)                             # modules.py:17
        tb = transposed.reshape(n * dilation, -1, c)                         # modules.py:18
        tb = np.pad(tb, [(0, 0), (k - 1, 0), (0, 0)])                        # modules.py:32
        conv = conv1d_valid(tb, w)                                           # modules.py:33
        co = conv.shape[2]
        prepared = conv.reshape(dilation, -1, co)                            # modules.py:22
        transposed = prepared.transpose(1, 0, 2)                             # modules.py:23
        restored = transposed.reshape(conv.shape[0] // dilation, -1, co)     # modules.py:24-25
        return restored[:, :t, :]                                            # modules.py:37-39
    padded = np.pad(x, [(0, 0), (k - 1, 0), (0, 0)])                         # modules.py:41
    return conv1d_valid(padded, w)                                           # modules.py:42


def causal_conv(x, w, dilation):
    """Direct form of reference modules.py:11-43:
    y[n, t] = sum_j x[n, t - (k-1-j)*dilation] @ w[j], taps before t=0 read zeros."""
    n, t, _ = x.shape
    k = w.shape[0]
    y = None
    for j in range(k):
        shift = (k - 1 - j) * dilation
        if shift == 0:
            term = x @ w[j]
        else:
            xs = _OPS.zeros_like(x)
            if shift < t:
                xs[:, shift:, :] = x[:, :t - shift, :]
            term = xs @ w[j]
        y = term if y is None else y + term
    return y


def conv1x1(x, w):
    """tf.nn.conv1d(..., stride=1, padding='SAME') with a width-1 kernel = per-timestep matmul."""
    assert w.shape[0] == 1
    return x @ w[0]


# ----------------------------------------------------------------------------- normalisers
def instance_normalization(x, gamma, beta, epsilon=1e-8):
    """Reference modules.py:274-284: per utterance and channel, mean / (population) variance over the TIME axis,
    (x - mean) / (variance + 1e-8) ** .5, then gamma * . + beta."""
    mean, variance = _OPS.moments_time(x)                                    # :278
    return gamma * ((x - mean) / ((variance + epsilon) ** .5)) + beta        # :282-283


def normalize(x, W, scope):
    """Reference modules.py:263-270 at the call site whose variable scope is `scope`. Which method the graph was
    built with is read off the weight container: method 'in' creates `<scope>/beta`, `<scope>/gamma`; method ''
    creates nothing and is the identity. ('bn' is not restated: tf.layers.batch_normalization lives in TensorFlow.)"""
    if scope + '/gamma' in W:
        return instance_normalization(x, W[scope + '/gamma'], W[scope + '/beta'])
    return x


# ----------------------------------------------------------------------------- WaveNet body
def _tanh(x):
    return _OPS.tanh(x)


def _sigmoid(x):
    return 1.0 / (1.0 + _OPS.exp(-x))


def dilation_layer(cur, cond, W, prefix, dilation, use_biases):
    """Reference modules.py:185-259. Returns (skip_output, dense_output, z)."""
    f = causal_conv(cur, W[prefix + '/filter'], dilation)                    # :213
    g = causal_conv(cur, W[prefix + '/gate'], dilation)                      # :214
    if cond is not None:
        f = f + conv1x1(cond, W[prefix + '/gc_filter'])                      # :216-219
        g = g + conv1x1(cond, W[prefix + '/gc_gate'])                        # :220-222
    if use_biases:
        f = f + W[prefix + '/filter_bias']                                   # :227
        g = g + W[prefix + '/gate_bias']                                     # :228
    f = normalize(f, W, prefix + '/normalize_filter')                        # :230-232
    g = normalize(g, W, prefix + '/normalize_gate')                          # :233-234
    z = _tanh(f) * _sigmoid(g)                                               # :236
    transformed = conv1x1(z, W[prefix + '/dense'])                           # :239-240
    skip = conv1x1(z, W[prefix + '/skip'])                                   # :243-244
    if use_biases:
        transformed = transformed + W[prefix + '/dense_bias']                # :249
        skip = skip + W[prefix + '/skip_bias']                               # :250
    dense_output = cur + transformed                                         # :251
    skip = normalize(skip, W, prefix + '/normalize_skip_output')             # :253-255
    dense_output = normalize(dense_output, W, prefix + '/normalize_dense_output')    # :256-257
    return skip, dense_output, z                                             # :259


def wavenet(x, cond, W, prefix, dilations, use_biases=True, use_skip_connection=False, taps=None):
    """Reference modules.py:129-166, WaveNet.__call__. x (N,T,1), cond (N,T,Cc) -> (N,T,1)."""
    cur = causal_conv(x, W[prefix + '/causal_layer/filter'], 1)              # :133-134,174-183 (no bias)
    cur = normalize(cur, W, prefix + '/causal_layer/normalize')              # :181-182
    if taps is not None:
        taps[prefix + '/causal_layer'] = cur
    outputs = []
    for j, d in enumerate(dilations):                                        # :138-142
        skip, cur, _ = dilation_layer(cur, cond, W, f'{prefix}/dilated_stack/layer{j}', d, use_biases)
        outputs.append(skip)
        if taps is not None:
            taps[f'{prefix}/dilated_stack/layer{j}'] = cur
    if use_skip_connection:                                                  # :147 (python sum: left to right from 0)
        total = 0
        for o in outputs:
            total = total + o
    else:
        total = outputs[-1]
    h = _OPS.relu(total)                                                     # :148
    h = normalize(h, W, prefix + '/postprocessing/normalize_postprocess1')   # :149-151
    h = conv1x1(h, W[prefix + '/postprocessing/postprocess1'])               # :152-153
    if use_biases:
        h = h + W[prefix + '/postprocessing/postprocess1_bias']              # :155-156
    h = _OPS.relu(h)                                                         # :157
    h = normalize(h, W, prefix + '/postprocessing/normalize_postprocess2')   # :158-160
    y = conv1x1(h, W[prefix + '/postprocessing/postprocess2'])               # :161-162
    if use_biases:
        y = y + W[prefix + '/postprocessing/postprocess2_bias']              # :164-165
    return y


# ----------------------------------------------------------------------------- IAFVocoder
def upsample_cond_repeat(mel, w_dense, hop):
    """Reference models.py:127-133: 1x1 conv (no bias) -> relu -> tile+reshape (== repeat each
    frame `hop` times along time) -> crop hop//2 at the front and -(-hop//2) ... i.e. Python's
    `-hop // 2` (floor division of the negated value) at the back."""
    n, t_mel, _ = mel.shape
    cc = w_dense.shape[2]
    cond = _OPS.relu(conv1x1(mel, w_dense))                                  # :129-130
    cond = _OPS.tile_last(cond, hop).reshape(-1, t_mel * hop, cc)            # :131-132
    return cond[:, hop // 2: -hop // 2, :]                                   # :133


def upsample_cond_transposed(mel, W, strides, hop):
    """Reference models.py:109-124 (normalize_cond ''): per stage i, tf.nn.conv2d_transpose of the (n, 1, len, Cin)
    sequence with filter w_i [1, stride, Cout, Cin], strides [1, 1, stride, 1], SAME padding, then relu. The filter
    is exactly as wide as the stride, so output position l*stride + j has the single source l:
        out[n, l*stride + j, co] = sum_ci in[n, l, ci] * w_i[0, j, co, ci]
    (conv2d_transpose is the gradient of conv2d w.r.t. its input: no kernel flip). Crop hop//2 at both ends."""
    cond = mel
    for i, stride in enumerate(strides):
        w = W[f'{ROOT}/cond/transposed_conv_{i}_weights']                    # :113-115
        n, length, cin = cond.shape
        assert w.shape[0] == 1 and w.shape[1] == stride and w.shape[3] == cin, (w.shape, stride, cin)
        cout = w.shape[2]
        out = cond.reshape(n * length, cin) @ w[0].reshape(stride * cout, cin).T       # [n*len, stride*cout]
        cond = _OPS.relu(out.reshape(n, length * stride, cout))                        # :118-120
        # :121-122 -- the reference normalises the 4-D tensor (n, 1, length, C) and instance_normalization's
        # "time axis" is axis 1 (modules.py:277), here the dummy height of size 1: mean = x, variance = 0, so the
        # stage's output collapses to beta. Replayed literally (a quirk of the reference, not of this restatement).
        cond = normalize(cond.reshape(n, 1, length * stride, cout), W,
                         f'{ROOT}/cond/normalize_transposed_conv_{i}').reshape(n, length * stride, cout)
    return cond[:, hop // 2: -hop // 2, :]                                   # :124


def logistic_noise(shape, seed, dtype=np.float64):
    """Sample of Logistic(0,1) as tf.contrib.distributions does it: log(u) - log1p(-u), u~U(0,1)
    (reference models.py:32-33). Deterministic stand-in with u clipped to [1e-7, 1-1e-7]."""
    u = np.random.RandomState(seed).uniform(1e-7, 1.0 - 1e-7, size=shape)
    return (np.log(u) - np.log1p(-u)).astype(dtype)


def iaf_vocoder_forward(noise, mel, W, dilations, hop, use_biases=True, use_skip_connection=False,
                        dtype=np.float64, taps=None, ops=None):
    """Reference models.py:23-78 (is_training=False; conditioning upsampled by 'repeat' or 'transposed_conv' and the
    normalisers '' or 'in', all read off the variables present in W).

    noise (N,T) or (N,T,1): the logistic sample the reference draws in-graph (models.py:32-33) --
    an INPUT here so that both sides see the same numbers. mel (N, 1+T//hop, n_mels).
    W: name -> array in TF layout. dilations: list (per flow) of lists. Returns (N, T) `dtype`.
    """
    global _OPS
    prev_ops, _OPS = _OPS, (ops or NumpyOps())
    try:
        return _OPS.to_numpy(_forward(noise, mel, W, dilations, hop, use_biases, use_skip_connection, dtype, taps))
    finally:
        _OPS = prev_ops


def _forward(noise, mel, W, dilations, hop, use_biases, use_skip_connection, dtype, taps):
    W = {k: _OPS.asarray(v, dtype) for k, v in W.items()}
    mel = _OPS.asarray(mel, dtype)
    x = _OPS.asarray(noise, dtype).reshape(noise.shape[0], noise.shape[1], 1)
    n, t, _ = x.shape
    if f'{ROOT}/cond/dense' in W:
        cond = upsample_cond_repeat(mel, W[f'{ROOT}/cond/dense'], hop)       # models.py:26,127-133
    elif f'{ROOT}/cond/transposed_conv_0_weights' in W:
        cond = upsample_cond_transposed(mel, W, UPSAMPLE_STRIDES, hop)       # models.py:26,109-124
    else:
        cond = None                                                          # models.py:134-135: any other method, no conditioning
    if cond is not None:
        cond = normalize(cond, W, f'{ROOT}/cond/normalize/normalize')        # models.py:27-29 (after the crop)
    if cond is not None and cond.shape[1] != t:
        raise ValueError(f'cond length {cond.shape[1]} != {t}: length must be a multiple of hop '
                         f'and mel must have 1 + length//hop frames')
    for i, dil in enumerate(dilations):                                      # models.py:34
        p = f'{ROOT}/iaf{i}'
        scale = wavenet(x, cond, W, p + '/scalar', dil, use_biases, use_skip_connection, taps)
        shift = wavenet(x, cond, W, p + '/shifter', dil, use_biases, use_skip_connection, taps)
        if taps is not None:                                                 # the two WaveNet outputs (modules.py:56-57)
            taps[p + '/scalar'] = _OPS.to_numpy(scale[:, :, 0]).copy()
            taps[p + '/shifter'] = _OPS.to_numpy(shift[:, :, 0]).copy()
        x = x * scale + shift                                                # modules.py:57-59
        x = normalize(x, W, f'{ROOT}/normalize{i}')                          # models.py:70
        if taps is not None:
            taps[p] = _OPS.to_numpy(x[:, :, 0]).copy()
    return x[:, :, 0]


def synthetic_inputs(n, t, hop, n_mels, mel_seed=1234, noise_seed=1235, dtype=np.float32):
    """SURVEY 8(d) synthetic inputs: mel ~ U(-1,1) (the range reference audio.py:278-286 maps to),
    logistic noise."""
    mel = np.random.RandomState(mel_seed).uniform(-1.0, 1.0, size=(n, 1 + t // hop, n_mels))
    noise = logistic_noise((n, t), noise_seed)
    return noise.astype(dtype), mel.astype(dtype)
