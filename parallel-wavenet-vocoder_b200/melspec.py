"""Mel front end of the generation path: wav file -> (wav chunk, normalised mel-dB spectrogram).

Restates what the reference's `Dataset._get_wav_and_melspec` does for `is_training=False`
(reference data_load.py:37-56) with the librosa-0.5.x calls it makes (reference audio.py:14-35,
102-141,232-243,254-286,327-356) written out in torch / numpy / scipy -- librosa, soundfile and
resampy are not available here:

  read_wav      librosa.load(path, sr=sr, mono=True)           -> scipy.io.wavfile + polyphase resample
  trim_wav      librosa.effects.trim(wav)  (top_db=60, frame 2048, hop 512, ref=max)
  first chunk   wav[0:length], zero-padded to `length`         (data_load.py:45-50)
  wav2melspec_db  |stft(n_fft, hop, win, hann, centred, reflect)| -> slaney mel basis ->
                amplitude_to_db (amin 1e-5, top_db 80) -> clip((db-min)/(max-min),0,1)*2-1

PARITY NOTE: this row has no oracle beyond the restated formulas (the reference's dependency is
absent): the mel basis is cross-checked against torchaudio's slaney filterbank in the CPU tests;
resampling uses scipy's polyphase filter instead of resampy's kaiser_best, so files that are not
already at `signal.sr` differ from the reference at the -80 dB level.
"""
import numpy as np
import torch


def read_wav(path, sr):
    """float32 mono waveform in [-1, 1] at `sr` Hz."""
    from scipy.io import wavfile
    file_sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        wav = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        wav = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        wav = (data.astype(np.float32) - 128.0) / 128.0
    else:
        wav = data.astype(np.float32)
    if wav.ndim == 2:
        wav = wav.mean(axis=1)
    if file_sr != sr:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(file_sr), int(sr))
        wav = resample_poly(wav, int(sr) // g, int(file_sr) // g).astype(np.float32)
    return np.ascontiguousarray(wav, dtype=np.float32)


def _frame_rms_power(wav, frame_length, hop_length):
    """librosa.feature.rmse(y, frame_length, hop_length)**2: centred (reflect-padded) frames."""
    pad = frame_length // 2
    y = np.pad(wav, pad, mode='reflect') if len(wav) > pad else np.pad(wav, pad, mode='constant')
    n_frames = 1 + (len(y) - frame_length) // hop_length
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n_frames)[:, None]
    return np.mean(np.abs(y[idx]) ** 2, axis=1)


def trim_wav(wav, top_db=60, frame_length=2048, hop_length=512):
    """librosa.effects.trim: drop leading / trailing frames more than `top_db` below the loudest."""
    if len(wav) == 0:
        return wav
    mse = _frame_rms_power(wav, frame_length, hop_length)
    ref = max(float(mse.max()), 1e-10)
    db = 10.0 * np.log10(np.maximum(1e-10, mse)) - 10.0 * np.log10(ref)
    nonsilent = np.flatnonzero(db > -top_db)
    if nonsilent.size == 0:
        return wav[:0]
    start = int(nonsilent[0]) * hop_length
    end = min(len(wav), (int(nonsilent[-1]) + 1) * hop_length)
    return wav[start:end]


def fix_length(wav, length):
    if len(wav) >= length:
        return wav[:length]
    return np.pad(wav, (0, length - len(wav)))


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels) with its defaults (slaney scale, area-normalised
    triangles): (n_mels, 1 + n_fft // 2) float32."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)


def wav2melspec_db(wav, sr, n_fft, win_length, hop_length, n_mels, max_db=None, min_db=None, device='cpu'):
    """(T,) or (N, T) waveform -> (t_mel, n_mels) / (N, t_mel, n_mels) float32, t_mel = 1 + T // hop."""
    x = torch.as_tensor(wav, dtype=torch.float32, device=device)
    squeeze = x.dim() == 1
    if squeeze:
        x = x[None]
    window = torch.hann_window(win_length, periodic=True, dtype=torch.float32, device=x.device)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=True,
                      pad_mode='reflect', return_complex=True)
    mag = spec.abs()                                                         # (N, 1+n_fft/2, t_mel)
    basis = torch.from_numpy(mel_basis(sr, n_fft, n_mels)).to(x.device)
    mel = torch.matmul(basis, mag)                                           # (N, n_mels, t_mel)
    amin, top_db = 1e-5, 80.0                                                # librosa.amplitude_to_db defaults
    db = 20.0 * torch.log10(torch.clamp(mel, min=amin))
    db = torch.maximum(db, db.amax(dim=(1, 2), keepdim=True) - top_db)
    if max_db and min_db:
        db = (torch.clamp((db - min_db) / (max_db - min_db), 0, 1) - 0.5) * 2
    out = db.transpose(1, 2).contiguous()
    return out[0] if squeeze else out


class MelFrontEnd:
    """wav2melspec_db on the GPU through the C-ABI (`pwv_melspec_*`, csrc/pwv_mel.cuh): device waveform in, device
    mel out, no host round trip. The CPU function above is its restated checker (tests/test_gpu_melspec.py)."""

    def __init__(self, sr, n_fft, win_length, hop_length, n_mels, max_db=None, min_db=None):
        import ctypes
        from . import _lib
        self._lib_mod, self.lib = _lib, _lib.load()
        self.hop, self.n_mels = int(hop_length), int(n_mels)
        cfg = _lib.PwvMelConfig()
        cfg.n_fft, cfg.win_length, cfg.hop_length, cfg.n_mels = int(n_fft), int(win_length), int(hop_length), int(n_mels)
        cfg.normalise = int(bool(max_db and min_db))
        cfg.min_db, cfg.max_db = float(min_db or 0.0), float(max_db or 0.0)
        basis = np.ascontiguousarray(mel_basis(sr, n_fft, n_mels), dtype=np.float32)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.pwv_melspec_create(ctypes.byref(cfg), basis.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._h)))

    def __call__(self, wav):
        """(N, T) float32 CUDA tensor -> (N, 1 + T // hop, n_mels) float32 CUDA tensor (asynchronous on the current stream)."""
        import ctypes
        assert wav.is_cuda and wav.dtype == torch.float32 and wav.dim() == 2
        wav = wav.contiguous()
        n, t = wav.shape
        out = torch.empty((n, 1 + t // self.hop, self.n_mels), dtype=torch.float32, device=wav.device)
        stream = torch.cuda.current_stream(wav.device).cuda_stream
        with torch.cuda.device(wav.device):
            self._lib_mod.check(self.lib.pwv_melspec_forward(self._h, wav.data_ptr(), out.data_ptr(), n, t, ctypes.c_void_p(stream)))
        return out

    def close(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            self.lib.pwv_melspec_destroy(h)
            h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def wav_and_melspec(path, hp_signal, length, device='cpu'):
    """reference data_load.py:37-56 for generation: -> (wav (length, 1) f32, melspec (1+length//hop, n_mels) f32)."""
    wav = read_wav(path, sr=int(hp_signal.sr))
    wav = trim_wav(wav)
    wav = fix_length(wav[:length], length)
    if str(device).startswith('cuda'):
        front = MelFrontEnd(int(hp_signal.sr), int(hp_signal.n_fft), int(hp_signal.win_length), int(hp_signal.hop_length),
                            int(hp_signal.n_mels), max_db=hp_signal.max_db, min_db=hp_signal.min_db)
        mel = front(torch.from_numpy(wav)[None].to(device))[0]
    else:
        mel = wav2melspec_db(wav, sr=int(hp_signal.sr), n_fft=int(hp_signal.n_fft), win_length=int(hp_signal.win_length),
                             hop_length=int(hp_signal.hop_length), n_mels=int(hp_signal.n_mels),
                             max_db=hp_signal.max_db, min_db=hp_signal.min_db, device=device)
    return wav[:, None].astype(np.float32), mel.cpu().numpy().astype(np.float32)
