#!/usr/bin/env python
"""Small forward passes of every arithmetic path for compute-sanitizer (memcheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize.py > profiles/..._memcheck.txt

Graphs: a 2-flow graph with a d >= T tap and a ragged last tile ('repeat' conditioning, all three precisions,
k_flow_tc as one launch per flow and as one launch per layer) and the same graph with
cond_upsample_method='transposed_conv'. Prints max|delta| vs the oracle (the checker; not the thing checked)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam
W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder')
from oracle import iaf_oracle as O

n, t = 3, 1040
for method in ('repeat', 'transposed_conv'):
    hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': [[1, 2, 512], [4, 1]], 'cond_upsample_method': method},
                        'generate': {'batch_size': n, 'length': t}}, case='sanitize/' + method)
    weights = W.init_weights(hp, seed=3, bias_std=0.1)
    d = W.model_dims(hp)
    noise, mel = O.synthetic_inputs(n, t, 80, 80)
    ref = O.iaf_vocoder_forward(noise, mel, weights, d['dilations'], d['hop'], dtype=np.float64)
    for precision, quiet in (('f16x3', '100'), ('f16x3', '1'), ('bf16', '100'), ('fp32', '0')):    # second field: PWV_TC_SEG
        os.environ['PWV_TC_SEG'] = quiet
        model = V.PwvModel(d, weights, precision)
        out = model.forward(torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda())
        torch.cuda.synchronize()
        print(method, precision, 'seg=' + quiet, 'max|delta|', float(np.abs(out.cpu().numpy() - ref).max()), flush=True)
        del model
