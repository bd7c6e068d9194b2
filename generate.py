#!/usr/bin/env python
"""`python generate.py [case] [ckpt] [--debug]` -- the reference's generation entry point
(reference generate.py:16-78) on the B200 path.

Same flow: load the case's hparams, build the dataset's generation split, build the model, restore
weights from `hp.logdir` (EMA shadows when `train.use_ema`; random init with a notice when no
checkpoint exists, as the reference does), run ONE forward pass, write `audio/pred` and `audio/gt`
summaries into `hp.logdir`, print `Done.`. What changed: the forward pass is the sm_100a library
instead of a TF session, and `data_path: synthetic` (this build's bench/parity cases) feeds
U(-1,1) mel + logistic noise instead of reading wav files.
"""
import importlib
import os
import sys

import numpy as np

_PKG = 'parallel-wavenet-vocoder_b200'


def generate(case='default', ckpt=None, debug=False):
    """
    :param case: experiment case name
    :param ckpt: checkpoint (file name inside hp.logdir) to load the model from
    :param debug: the reference attaches tfdbg; here: check every flow output for NaN/Inf
    """
    from hparam import hparam as hp
    from models import IAFVocoder
    io = importlib.import_module(_PKG + '.io')
    hp.set_hparam_yaml(case)

    # dataset (generation split) -- reference generate.py:27-35
    dataset = io.GenerationData(hp.data_path, hp.generate.batch_size, hp.generate.length)
    print('dataset size is {}'.format(len(dataset.wav_files)))
    gt_wav, melspec, noise = dataset.next_batch()

    # model + weights -- reference generate.py:31,55-66
    model = IAFVocoder(batch_size=hp.generate.batch_size, length=hp.generate.length)
    ckpt_path = io.find_checkpoint(hp.logdir, ckpt)
    if ckpt_path:
        model.load_weights(io.load_checkpoint(ckpt_path, use_ema=hp.train.use_ema))
        print('Successfully loaded checkpoint {}'.format(ckpt_path))
    else:
        print('No checkpoint found at {}.'.format(hp.logdir))

    # feed forward (the reference's single sess.run, generate.py:68)
    pred_wav = model(gt_wav, melspec, is_training=False, noise=noise)
    pred_wav = pred_wav.cpu().numpy()
    if debug and not np.isfinite(pred_wav).all():
        raise FloatingPointError('non-finite samples in the predicted waveform')

    # summaries -- reference generate.py:41-45,71-73
    io.write_audio_summaries(hp.logdir, hp.signal.sr, pred=pred_wav, gt=gt_wav)
    if (hp.get('engine', {}) or {}).get('write_wav'):          # PCM16 files like reference audio.py:19-20 (off by default)
        for i, w in enumerate(np.asarray(pred_wav).reshape(len(pred_wav), -1)):
            io.write_wav(w, hp.signal.sr, os.path.join(hp.logdir, 'pred_%d.wav' % i))
    print('Done.')
    return pred_wav


def _cli(argv):
    """`fire.Fire(generate)`-compatible argument handling (fire is not a dependency here):
    positionals in order (case, ckpt, debug), or --case=X / --case X, --ckpt=..., --debug."""
    kwargs, positional = {}, []
    it = iter(argv)
    for a in it:
        if a.startswith('--'):
            key, eq, val = a[2:].partition('=')
            key = key.replace('-', '_')
            if key not in ('case', 'ckpt', 'debug'):
                raise SystemExit('unknown flag --' + key)
            if key == 'debug' and not eq:
                val = 'True'
            elif not eq:
                val = next(it, None)
                if val is None:
                    raise SystemExit('flag --%s needs a value' % key)
            kwargs[key] = val
        else:
            positional.append(a)
    for key, val in zip(('case', 'ckpt', 'debug'), positional):
        kwargs.setdefault(key, val)
    if 'debug' in kwargs:
        kwargs['debug'] = str(kwargs['debug']).lower() in ('1', 'true', 'yes')
    if kwargs.get('ckpt') in ('None', ''):
        kwargs['ckpt'] = None
    return kwargs


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    generate(**_cli(sys.argv[1:]))
