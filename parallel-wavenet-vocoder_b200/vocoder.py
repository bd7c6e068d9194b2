"""Host-side mirror of the reference's model object for the generation path.

`IAFVocoder(batch_size, length)(wav, melspec, is_training=False, noise=None)` keeps the reference's
call shape (reference models.py:18-23,78): it reads the global `hparam`, ignores `wav` exactly as
the reference's forward does (models.py:23 never reads it), and returns the predicted waveform
`(N, length, 1)` float32. Where the reference builds a TF graph that `sess.run` later executes
(generate.py:38,68), this object owns a `pwv_model` (C-ABI, include/pwv.h) and runs the
hand-written sm_100a kernels directly. torch tensors are used as device buffers only.

The logistic noise the reference samples in-graph (models.py:32-33) is drawn here with torch on
the device unless the caller passes `noise` (parity tests and the bench do).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import weights as W
from .hparam import hparam as hp


def _assert_supported(hp):
    m = hp.model
    for key in ('normalize', 'normalize_cond', 'normalize_wavenet'):
        if m.get(key) and m[key] != 'in':
            raise NotImplementedError(f"model.{key}={m[key]!r}: '' and 'in' (instance normalisation over time, reference "
                                      f"modules.py:274-284) are on the B200 path; 'bn' (tf.layers.batch_normalization) is not")
    # cond_upsample_method outside 'repeat' / 'transposed_conv': the reference conditions on nothing (models.py:134-135:
    # cond = None, no cond/* and gc_* variables) -- so does this path; normalize_cond would then normalise None there
    if m.cond_upsample_method not in ('repeat', 'transposed_conv') and m.get('normalize_cond'):
        raise ValueError(f"model.cond_upsample_method={m.cond_upsample_method!r} leaves the graph unconditional; "
                         f"normalize_cond={m['normalize_cond']!r} cannot apply (the reference fails there too, models.py:27-29)")
    strides = list(W.UPSAMPLE_STRIDES)   # the reference asserts this even for 'repeat' (models.py:26,106)
    if int(np.prod(strides)) != int(hp.signal.hop_length):
        raise AssertionError(f'prod({strides}) != hop_length {hp.signal.hop_length} (reference models.py:106)')


def tensor_core_covers(dims):
    """Graphs the tcgen05 kernels implement: R = D = 64, S = 128 (csrc/pwv_tc2.cuh: the whole layer in one kernel) and
    R = D in {128, 256}, S = 2R without skip connections (csrc/pwv_tc3.cuh: streamed-K gate and dense passes)."""
    if dims['R'] != dims['D'] or dims['S'] != 2 * dims['R'] or dims['k'] != 2:
        return False                # free parameters in the reference (modules.py:210-244): the general fp32 chain (csrc/pwv_gen.cuh)
    if dims.get('norm_flow') or dims.get('norm_cond') or dims.get('norm_wavenet'):
        return False                # a statistic over the whole time axis sits between the stages: un-fused fp32 kernels
    return dims['R'] == 64 or (dims['R'] in (128, 256) and not dims['use_skip'])


def resolve_precision(dims, precision):
    """'auto' -> 'f16x3' (tcgen05, fp32-level parity) when the tensor-core kernels cover the graph, else the exact
    fp32 FFMA kernels -- said out loud, never silently (a user asking for 'auto' at 128 channels is told what runs)."""
    if precision in (None, '', 'auto'):
        if tensor_core_covers(dims):
            return 'f16x3'
        import warnings
        warnings.warn(f"engine.precision 'auto': filter_width {dims['k']}, residual/dilation/skip channels {dims['R']}/{dims['D']}/{dims['S']} "
                      f"(use_skip_connection={bool(dims['use_skip'])}, normalisers {dims.get('norm_flow')!r}/{dims.get('norm_cond')!r}/"
                      f"{dims.get('norm_wavenet')!r}) are outside the tensor-core kernels' coverage; "
                      f"running the exact fp32 FFMA kernels", RuntimeWarning, stacklevel=2)
        return 'fp32'
    return precision


class PwvModel:
    """Thin RAII wrapper over a finalized `pwv_model`."""

    def __init__(self, dims, weights, precision='auto', debug=None):
        """`debug`: {switch: int} passed to pwv_debug_set (tests and measurement tools only; include/pwv.h)."""
        self.lib = _lib.load()
        self.dims = dims
        precision = resolve_precision(dims, precision)
        self.precision = precision
        self._h = ctypes.c_void_p()
        hparams = _lib.make_hparams(dims, precision)
        _lib.check(self.lib.pwv_model_create(ctypes.byref(hparams), ctypes.byref(self._h)))
        for key, value in (debug or {}).items():
            _lib.check(self.lib.pwv_debug_set(self._h, key.encode(), int(value)))
        n = _lib.check(self.lib.pwv_model_num_variables(self._h))
        name = ctypes.c_char_p()
        shape = (ctypes.c_int64 * 4)()
        ndim = ctypes.c_int()
        for i in range(n):
            _lib.check(self.lib.pwv_model_variable(self._h, i, ctypes.byref(name), shape, ctypes.byref(ndim)))
            key = name.value.decode()
            if key not in weights:
                raise KeyError(f'weight container has no variable {key!r}')
            arr = np.ascontiguousarray(weights[key], dtype=np.float32)
            shp = (ctypes.c_int64 * max(arr.ndim, 1))(*arr.shape)
            _lib.check(self.lib.pwv_model_load_weight(self._h, name.value, arr.ctypes.data_as(ctypes.c_void_p),
                                                      shp, arr.ndim))
        _lib.check(self.lib.pwv_model_finalize(self._h))
        self._ws = None
        self.max_workspace_bytes = None     # None: bounded by the device's free memory (see _utterances_per_pass)

    def close(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self.lib.pwv_model_destroy(h)
            h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: ctypes may already be torn down
            pass

    def workspace_bytes(self, n, t):
        out = ctypes.c_size_t()
        _lib.check(self.lib.pwv_workspace_bytes(self._h, n, t, ctypes.byref(out)))
        return out.value

    def _workspace(self, n, t, device):
        need = self.workspace_bytes(n, t)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = None                       # (free the old block before asking for the larger one)
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _workspace_limit(self, device):
        """Bytes a workspace may take: `max_workspace_bytes` if set, else 90 % of what the device could give now
        (free memory + the workspace this model already holds)."""
        if self.max_workspace_bytes:
            return int(self.max_workspace_bytes)
        free, _ = torch.cuda.mem_get_info(device)
        held = self._ws.numel() if self._ws is not None and self._ws.device == device else 0
        return int(0.9 * (free + held))

    def _utterances_per_pass(self, n, t, device):
        """How many utterances one pwv_forward call may take. The path is independent per utterance (bit for bit:
        tests/test_gpu_parity.py), so a batch whose workspace would not fit -- a full-rate conditioning
        ('transposed_conv', normalize_cond) materialises 2 x layers x N x T x 2D floats -- is run in several passes."""
        def need(k):
            try:
                return self.workspace_bytes(k, t)
            except _lib.PwvError as e:
                if e.code != -4:              # PWV_ENOMEM: larger than the whole device
                    raise
                return float('inf')
        if self._ws is not None and self._ws.device == device and need(n) <= self._ws.numel():
            return n
        limit = self._workspace_limit(device)
        if need(n) <= limit:
            return n
        one = need(1)
        if one > limit:
            raise MemoryError(f'one utterance of {t} samples needs a {one}-byte workspace, {limit} are available')
        k = max(1, min(n, int(limit // one)))       # (a lower estimate: the fixed part is counted once per utterance)
        while k > 1 and need(k) > limit:
            k -= 1
        while k < n and need(k + 1) <= limit:
            k += 1
        return k

    def forward(self, noise, mel, out=None, taps=None):
        """noise (N,T) f32 cuda, mel (N,1+T//hop,n_mels) f32 cuda -> wav (N,T) f32 cuda.
        Asynchronous on torch's current stream. `taps`: dict with optional keys 'flow_out',
        'scale_shift' (True -> allocated) and 'layer' = (flow, body, index)."""
        assert noise.is_cuda and mel.is_cuda and noise.dtype == torch.float32 and mel.dtype == torch.float32
        noise = noise.contiguous()
        mel = mel.contiguous()
        n, t = noise.shape
        assert mel.shape == (n, 1 + t // self.dims['hop'], self.dims['n_mels']), mel.shape
        if out is None:
            out = torch.empty((n, t), dtype=torch.float32, device=noise.device)
        k = n if taps else self._utterances_per_pass(n, t, noise.device)
        if k < n:
            for i in range(0, n, k):
                self.forward(noise[i:i + k], mel[i:i + k], out=out[i:i + k])
            return out
        ws = self._workspace(n, t, noise.device)
        tp = None
        captured = {}
        if taps:
            tp = _lib.PwvTaps()
            tp.layer_flow = tp.layer_body = tp.layer_index = -1
            if taps.get('flow_out'):
                captured['flow_out'] = torch.empty((self.dims['n_iaf'], n, t), dtype=torch.float32, device=noise.device)
                tp.flow_out = captured['flow_out'].data_ptr()
            if taps.get('scale_shift'):
                captured['scale_shift'] = torch.empty((self.dims['n_iaf'], 2, n, t), dtype=torch.float32, device=noise.device)
                tp.scale_shift = captured['scale_shift'].data_ptr()
            if taps.get('layer') is not None:
                tp.layer_flow, tp.layer_body, tp.layer_index = taps['layer']
                captured['layer_out'] = torch.empty((n, t, self.dims['R']), dtype=torch.float32, device=noise.device)
                tp.layer_out = captured['layer_out'].data_ptr()
        stream = torch.cuda.current_stream(noise.device).cuda_stream
        with torch.cuda.device(noise.device):
            _lib.check(self.lib.pwv_forward(self._h, noise.data_ptr(), mel.data_ptr(), out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), n, t, ctypes.c_void_p(stream),
                                            ctypes.byref(tp) if tp is not None else None))
        return (out, captured) if taps else out

    def forward_host(self, noise, mel, out=None):
        """Host buffers in, host buffer out (numpy float32 or pinned CPU torch tensors); the H2D /
        D2H copies and the synchronise happen inside the C call (pwv_forward_host)."""
        noise_a = _as_host(noise)
        mel_a = _as_host(mel)
        n, t = noise_a.shape
        if out is None:
            out = np.empty((n, t), dtype=np.float32)
        out_a = _as_host(out)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.pwv_forward_host(self._h, _host_ptr(noise_a), _host_ptr(mel_a), _host_ptr(out_a),
                                             n, t, ctypes.c_void_p(stream)))
        return out

    def last_launch_count(self):
        return _lib.check(self.lib.pwv_last_launch_count(self._h))

    def set_profiling(self, enable):
        """0 / False: off; 1 / True: every gated-layer launch timed in isolation; 2: each flow's chain of
        gated-layer launches timed as launched in production (include/pwv.h)."""
        _lib.check(self.lib.pwv_set_profiling(self._h, 2 if enable == 2 else int(bool(enable))))

    def profile_read(self):
        """-> (summed ms of the gated-layer launches, their count, ms of the whole forward)."""
        layer_ms, n, fwd = ctypes.c_double(), ctypes.c_int(), ctypes.c_double()
        _lib.check(self.lib.pwv_profile_read(self._h, ctypes.byref(layer_ms), ctypes.byref(n), ctypes.byref(fwd)))
        return layer_ms.value, n.value, fwd.value


def _as_host(a):
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()
        return a
    a = np.asarray(a)
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a


def _host_ptr(a):
    return ctypes.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)


def sample_logistic(n, t, device, seed=None):
    """Logistic(0,1) sample, log(u) - log1p(-u) (what reference models.py:32-33 draws in-graph)."""
    gen = torch.Generator(device=device)
    if seed is not None:
        gen.manual_seed(int(seed))
    u = torch.rand((n, t), generator=gen, device=device, dtype=torch.float32).clamp_(1e-7, 1.0 - 1e-7)
    return torch.log(u) - torch.log1p(-u)


class IAFVocoder:
    """Drop-in for the reference's `IAFVocoder` on the generation path (reference models.py:16-78)."""

    def __init__(self, batch_size, length, weights=None, device='cuda'):
        _assert_supported(hp)
        self.batch_size = int(batch_size)
        self.length = int(length)
        self.t_mel = 1 + self.length // int(hp.signal.hop_length)      # reference models.py:20
        if self.length % int(hp.signal.hop_length) != 0:
            raise ValueError(f'length {length} must be a multiple of hop_length {hp.signal.hop_length} '
                             f'(the reference fails with a shape error otherwise, modules.py:218)')
        self.device = torch.device(device)
        self.dims = W.model_dims(hp)
        engine = hp.get('engine', {}) or {}
        self.precision = resolve_precision(self.dims, engine.get('precision', 'auto'))
        if weights is None:      # a fresh reference graph: Glorot kernels, zero biases (generate.py:56)
            weights = W.init_weights(hp, seed=int(engine.get('seed', 0)))
        W.check_weights(hp, weights)
        self.weights = weights
        with torch.cuda.device(self.device):
            self.model = PwvModel(self.dims, weights, self.precision)

    def load_weights(self, weights):
        W.check_weights(hp, weights)
        self.weights = weights
        with torch.cuda.device(self.device):
            self.model = PwvModel(self.dims, weights, self.precision)

    def __call__(self, wav, melspec, is_training=False, name='iaf_vocoder', noise=None, noise_seed=None):
        if is_training:
            raise NotImplementedError('the B200 path implements generation (is_training=False) only')
        mel = torch.as_tensor(melspec, dtype=torch.float32).to(self.device, non_blocking=True)
        n = mel.shape[0]
        if mel.shape[1:] != (self.t_mel, self.dims['n_mels']):
            raise ValueError(f'melspec shape {tuple(mel.shape)} != (N, {self.t_mel}, {self.dims["n_mels"]})')
        if noise is None:
            noise = sample_logistic(n, self.length, self.device, noise_seed)
        else:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device, non_blocking=True).reshape(n, self.length)
        out = self.model.forward(noise, mel)
        return out.unsqueeze(-1)                                       # (N, length, 1) like models.py:78
