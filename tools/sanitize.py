#!/usr/bin/env python
"""Small forward passes of every arithmetic path for compute-sanitizer (memcheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize.py > profiles/..._memcheck.txt

Graphs: a 2-flow graph with a d >= T tap and a ragged last tile ('repeat' conditioning, all three precisions, the
plane kernels k_layer_h with and without the tile-flag handshake, the round-1 kernels behind debug path 0) and the
same graph with cond_upsample_method='transposed_conv' and with use_skip_connection; then the wide tensor-core
kernels, the 'in' normaliser kernels, the general-shape chain and the mel front end. Prints max|delta| vs the oracle
(the checker; not the thing checked)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam
W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder')
from oracle import iaf_oracle as O

n, t = 3, 1040
for method, skip in (('repeat', False), ('transposed_conv', False), ('repeat', True)):
    hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': [[1, 2, 512], [4, 1]], 'cond_upsample_method': method, 'use_skip_connection': skip},
                        'generate': {'batch_size': n, 'length': t}}, case='sanitize/' + method)
    weights = W.init_weights(hp, seed=3, bias_std=0.1)
    d = W.model_dims(hp)
    noise, mel = O.synthetic_inputs(n, t, 80, 80)
    ref = O.iaf_vocoder_forward(noise, mel, weights, d['dilations'], d['hop'], True, skip, dtype=np.float64)
    runs = [('f16x3', {}), ('f16x3', {'tile_flags': 0}), ('bf16', {}), ('fp32', {})]
    if not skip:
        runs += [('f16x3', {'path': 0, 'seg': 100}), ('f16x3', {'path': 0, 'flow': 0})]
    for precision, debug in runs:
        model = V.PwvModel(d, weights, precision, debug=debug)
        out = model.forward(torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda())
        torch.cuda.synchronize()
        print(method, 'skip' if skip else '', precision, debug, 'max|delta|', float(np.abs(out.cpu().numpy() - ref).max()), flush=True)
        del model

# the wide tensor-core kernels (k_wide_h, 128 channels), the 'in' normaliser kernels, the general-shape chain, the mel front end
for name, model_hp, precision in (
        ('wide128', {'residual_channels': 128, 'dilation_channels': 128, 'skip_channels': 256}, 'f16x3'),
        ('norm', {'normalize': 'in', 'normalize_cond': 'in', 'normalize_wavenet': 'in', 'use_skip_connection': True}, 'fp32'),
        ('shapes', {'filter_width': 3, 'residual_channels': 24, 'dilation_channels': 40, 'skip_channels': 56, 'use_skip_connection': True}, 'fp32')):
    hp.set_hparam_yaml('default')
    hp.set_hparam_dict({'model': dict(model_hp, n_iaf=2, dilations=[[1, 2, 512], [4, 1]]), 'generate': {'batch_size': n, 'length': t}}, case='sanitize/' + name)
    weights = W.init_weights(hp, seed=4, bias_std=0.1)
    d = W.model_dims(hp)
    noise, mel = O.synthetic_inputs(n, t, 80, 80)
    ref = O.iaf_vocoder_forward(noise, mel, weights, d['dilations'], d['hop'], True, bool(d['use_skip']), dtype=np.float64)
    model = V.PwvModel(d, weights, precision)
    out = model.forward(torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda())
    torch.cuda.synchronize()
    print(name, precision, 'max|delta|', float(np.abs(out.cpu().numpy() - ref).max()), 'max|ref|', float(np.abs(ref).max()), flush=True)
    del model
M = importlib.import_module(P + '.melspec')
hp.set_hparam_yaml('default')
wav = np.random.RandomState(0).randn(2, 4000).astype(np.float32) * 0.1
sg = hp.signal
front = M.MelFrontEnd(sg.sr, sg.n_fft, sg.win_length, sg.hop_length, sg.n_mels, sg.max_db, sg.min_db)
mel_gpu = front(torch.from_numpy(wav).cuda())
torch.cuda.synchronize()
print('mel front end', tuple(mel_gpu.shape), 'finite', bool(torch.isfinite(mel_gpu).all()), flush=True)
