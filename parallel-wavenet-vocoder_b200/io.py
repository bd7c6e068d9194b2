"""Inputs and outputs either side of the forward pass of `generate.py`.

* `GenerationData`: the generation split of the dataset (reference data_load.py:18-25: glob, the
  part after `train.dataset_ratio`). `data_path: synthetic` yields the bench inputs instead
  (mel ~ U(-1,1) as reference audio.py:278-286 normalises to, logistic noise). Real wav files go
  through melspec.py (the reference's librosa front end restated in torch/scipy).
* `find_checkpoint` / `load_checkpoint`: weights from `hp.logdir`: TensorFlow tensor-bundle
  checkpoints (`model-*.index` + `.data-*`, parsed natively by tf_bundle.py) or a `.npz` keyed by TF
  variable names (weights.py).
* `write_audio_summaries`: `audio/pred`, `audio/gt` into a TensorBoard event file in `hp.logdir`
  (reference generate.py:41-45,71-73) when tensorboard is importable, plus `.npy` copies.
* `write_wav`: PCM16 wav files (reference audio.py:19-20 `write_wav`, soundfile's 'PCM_16'); `engine.write_wav: true`
  makes `generate.py` leave `pred_<i>.wav` beside the summaries.
"""
import glob
import os

import numpy as np
import torch

from .hparam import hparam as hp


def synthetic_batch(n, t, hop, n_mels, mel_seed=1234, noise_seed=1235):
    """SURVEY 8(d) synthetic inputs: mel ~ U(-1, 1) (the range reference audio.py:278-286 normalises to) and the
    logistic sample log(u) - log1p(-u) the reference draws in-graph (models.py:32-33). -> (noise (n,t), mel (n,1+t//hop,n_mels)) f32"""
    mel = np.random.RandomState(mel_seed).uniform(-1.0, 1.0, size=(n, 1 + t // hop, n_mels))
    u = np.random.RandomState(noise_seed).uniform(1e-7, 1.0 - 1e-7, size=(n, t))
    return (np.log(u) - np.log1p(-u)).astype(np.float32), mel.astype(np.float32)


class GenerationData:
    def __init__(self, data_path, batch_size, length):
        self.batch_size = int(batch_size)
        self.length = int(length)
        self.synthetic = (data_path == 'synthetic')
        if self.synthetic:
            self.wav_files = []
            return
        files = sorted(glob.glob(data_path))
        if len(files) > 1:                                # reference data_load.py:22-23
            split = int(len(files) * hp.train.dataset_ratio)
            files = files[split:]
        self.wav_files = files

    def next_batch(self):
        """-> (gt_wav (N,T,1) or None, melspec (N,t_mel,n_mels) f32, noise (N,T) f32 or None)"""
        hop, n_mels = int(hp.signal.hop_length), int(hp.signal.n_mels)
        n, t = self.batch_size, self.length
        if self.synthetic:
            engine = hp.get('engine', {}) or {}
            noise, mel = synthetic_batch(n, t, hop, n_mels, 1234, int(engine.get('noise_seed', 1235)))
            return None, mel, noise
        from . import melspec
        if not self.wav_files:
            raise FileNotFoundError(f'no wav files match data_path {hp.data_path!r}')
        wavs, mels = [], []
        for i in range(n):                       # the reference batches consecutive (shuffled) files; here: in order
            wav, mel = melspec.wav_and_melspec(self.wav_files[i % len(self.wav_files)], hp.signal, t,
                                               device='cuda' if torch.cuda.is_available() else 'cpu')
            wavs.append(wav)
            mels.append(mel)
        return np.stack(wavs), np.stack(mels), None


def find_checkpoint(logdir, ckpt=None):
    """`<logdir>/<ckpt>` when named (reference generate.py:55), else the latest checkpoint in `logdir`:
    a TensorFlow bundle prefix (via the `checkpoint` state file, as tf.train.latest_checkpoint does)
    or this build's `.npz` container, whichever is newer. Returns a path or None."""
    from . import tf_bundle
    if ckpt:
        path = os.path.join(logdir, ckpt)
        if os.path.exists(path + '.index'):
            return path + '.index'
        for cand in (path, path + '.npz'):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(path)
    cands = sorted(glob.glob(os.path.join(logdir, '*.npz')), key=os.path.getmtime)
    prefix = tf_bundle.latest_checkpoint(logdir) if os.path.isdir(logdir) else None
    if prefix and (not cands or os.path.getmtime(prefix + '.index') >= os.path.getmtime(cands[-1])):
        return prefix + '.index'
    return cands[-1] if cands else None


def load_checkpoint(path, use_ema=False):
    """name -> float32 array for every variable of the current hparams' graph. `.npz` containers are
    keyed by TF variable names; `<prefix>.index` is a TensorFlow tensor bundle (tf_bundle.py). With
    `use_ema` the `<name>/ExponentialMovingAverage` shadows win (reference generate.py:58-63)."""
    from . import tf_bundle
    from . import weights as W
    if path.endswith('.npz'):
        return W.load_npz(path, use_ema=use_ema)
    if path.endswith('.index'):
        return tf_bundle.load_variables(path[:-len('.index')], list(W.variable_shapes(hp).keys()), use_ema=use_ema)
    raise ValueError(f'{path}: unknown checkpoint format (expected .npz or a TensorFlow bundle .index)')


def write_audio_summaries(logdir, sr, pred, gt=None):
    os.makedirs(logdir, exist_ok=True)
    pred = np.asarray(pred, dtype=np.float32)
    np.save(os.path.join(logdir, 'pred_wav.npy'), pred)
    try:
        from torch.utils.tensorboard import SummaryWriter
    except Exception:           # tensorboard not installed: the .npy copy is the sink
        return
    writer = SummaryWriter(logdir)
    for i in range(min(pred.shape[0], 3)):               # tf.summary.audio default max_outputs=3
        writer.add_audio('audio/pred/%d' % i, torch.from_numpy(pred[i].reshape(1, -1)).clamp(-1, 1), 0, sample_rate=int(sr))
        if gt is not None:
            writer.add_audio('audio/gt/%d' % i, torch.as_tensor(gt[i]).reshape(1, -1).clamp(-1, 1), 0, sample_rate=int(sr))
    writer.close()


def write_wav(wav, sr, path):
    """Mono PCM16 wav (reference audio.py:19-20: `sf.write(path, wav, sr, format='wav', subtype='PCM_16')`): float
    samples in [-1, 1) scaled by 32768 and clipped to the int16 range, as libsndfile's default conversion does."""
    import wave
    data = np.asarray(wav, dtype=np.float64).reshape(-1)
    pcm = np.clip(np.rint(data * 32768.0), -32768, 32767).astype('<i2')
    with wave.open(path, 'wb') as fh:
        fh.setnchannels(1)
        fh.setsampwidth(2)
        fh.setframerate(int(sr))
        fh.writeframes(pcm.tobytes())
