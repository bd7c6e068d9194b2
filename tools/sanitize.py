#!/usr/bin/env python
"""Small forward passes of every arithmetic path for compute-sanitizer (memcheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize.py > profiles/..._memcheck.txt

Graphs: a 2-flow graph with a d >= T tap and a ragged last tile ('repeat' conditioning, all three precisions, the
plane kernels k_layer_h with and without the tile-flag handshake, the round-1 kernels behind debug path 0) and the
same graph with cond_upsample_method='transposed_conv' and with use_skip_connection. Prints max|delta| vs the oracle
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
