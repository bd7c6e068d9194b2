"""`generate.py` end to end on the GPU -- the north_star's drop-in entry point (reference generate.py:16-75):
hparams case -> synthetic generation batch -> model -> EMA shadows restored from a TensorFlow tensor-bundle
checkpoint in `hp.logdir` -> one forward pass -> `audio/pred` summaries + `pred_wav.npy`."""
import importlib.util
import io
import os
import shutil
import wave

import numpy as np
import pytest

from conftest import ROOT, pkg
from oracle import iaf_oracle as O

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _load():
    spec = importlib.util.spec_from_file_location('generate_entry', os.path.join(ROOT, 'generate.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_generate_restores_ema_checkpoint_and_writes_summaries(hp, monkeypatch, capsys):
    monkeypatch.chdir(ROOT)
    hp.set_hparam_yaml('parity/small')
    logdir = hp.logdir
    shutil.rmtree(logdir, ignore_errors=True)
    os.makedirs(logdir)
    W, B = pkg('weights'), pkg('tf_bundle')
    # the checkpoint a tensorpack run leaves behind: live variables, their EMA shadows, optimiser slots
    live = W.init_weights(hp, seed=21, bias_std=0.1)
    shadow = W.init_weights(hp, seed=22, bias_std=0.1)
    bundle = dict(live)
    bundle.update({k + W.EMA_SUFFIX: v for k, v in shadow.items()})
    bundle['global_step'] = np.array(1200, dtype=np.int64)
    B.write_bundle(os.path.join(logdir, 'model-1200'), bundle, block_entries=16)
    with open(os.path.join(logdir, 'checkpoint'), 'w') as fh:
        fh.write('model_checkpoint_path: "model-1200"\nall_model_checkpoint_paths: "model-1200"\n')

    pred = _load().generate('parity/small')
    printed = capsys.readouterr().out
    assert 'Successfully loaded checkpoint' in printed and 'Done.' in printed
    n, t = int(hp.generate.batch_size), int(hp.generate.length)
    assert pred.shape == (n, t, 1)

    # against the oracle on the SHADOW weights (train.use_ema: reference generate.py:58-63), same synthetic batch
    _, mel, noise = pkg('io').GenerationData('synthetic', n, t).next_batch()
    d = W.model_dims(hp)
    ref = O.iaf_vocoder_forward(noise, mel, shadow, d['dilations'], d['hop'], dtype=np.float64)
    assert np.abs(pred[:, :, 0] - ref).max() <= 1e-4
    ref_live = O.iaf_vocoder_forward(noise, mel, live, d['dilations'], d['hop'], dtype=np.float64)
    assert np.abs(pred[:, :, 0] - ref_live).max() > 1e-3          # the shadows, not the live variables, were used

    # sinks: pred_wav.npy and the TensorBoard event file with audio/pred
    saved = np.load(os.path.join(logdir, 'pred_wav.npy'))
    assert np.array_equal(saved, pred)
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    ea = EventAccumulator(logdir, size_guidance={'audio': 0})
    ea.Reload()
    tags = ea.Tags()['audio']
    assert 'audio/pred/0' in tags and 'audio/pred/1' in tags
    ev = ea.Audio('audio/pred/0')[0]
    assert int(ev.sample_rate) == int(hp.signal.sr)
    with wave.open(io.BytesIO(ev.encoded_audio_string)) as wf:
        assert wf.getframerate() == int(hp.signal.sr) and wf.getnframes() == t
        pcm = np.frombuffer(wf.readframes(t), dtype=np.int16).astype(np.float64) / 32767.0
    assert np.abs(pcm - np.clip(pred[0, :, 0], -1, 1)).max() <= 2.0 / 32767.0
