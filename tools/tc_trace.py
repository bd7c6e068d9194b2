#!/usr/bin/env python
"""Print the phase timeline (clock64 deltas) of CTA 0 of the last tensor-core layer launch."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam
W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder'); L = importlib.import_module(P + '._lib')
from oracle import iaf_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3'
# one flow, layers d=1,2 so the LAST mode-0 launch is layer index 0? -> use 3 layers; the trace keeps the last mode-0 or mode-1 launch
hp.set_hparam_dict({'model': {'n_iaf': 1, 'dilations': [[1, 64, 2]]}, 'generate': {'batch_size': 8, 'length': 16000}}, case='trace')
weights = W.init_weights(hp, seed=0)
m = V.PwvModel(W.model_dims(hp), weights, prec)
noise, mel = O.synthetic_inputs(8, 16000, 80, 80)
noise, mel = torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda()
buf = torch.zeros(4 * 16 * 16, dtype=torch.int64, device='cuda')
for it in range(3):
    m.forward(noise, mel)
torch.cuda.synchronize()
# trace only the middle layer: hack -- set trace, run, the last writer is the mode-1 layer; so use 2-layer model instead
hp.set_hparam_dict({'model': {'n_iaf': 1, 'dilations': [[64, 2]]}, 'generate': {'batch_size': 8, 'length': 16000}}, case='trace')
weights = W.init_weights(hp, seed=0)
m = V.PwvModel(W.model_dims(hp), weights, prec)
for it in range(2):
    m.forward(noise, mel)
L.check(m.lib.pwv_debug_set_trace(m._h, ctypes.c_void_p(buf.data_ptr())))
m.forward(noise, mel)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(4, 16, 16)
print('NOTE: the buffer holds the union of the mode-0 layer (events 0-9) and the mode-1 layer (overwrites 0-5, 8-9)')
base = t[t > 0].min()
names = ['enter', 'x_landed', 'x_prepped', 'y_landed', 'a_ready', 'd1_ready', 'z_ready', 'd2_ready', 'out_ready']
for role in (0, 1):
    print(f'--- worker slot {role} (cycles since first stamp; delta from previous event)')
    for j in range(8):
        row = t[role, j]
        if row[0] == 0: break
        s = []
        prev = row[0]
        for k, nm in enumerate(names):
            if row[k] == 0: continue
            s.append(f'{nm}={row[k]-base}(+{row[k]-prev})')
            prev = row[k]
        print(f' tile {j}: ' + ' '.join(s))
print('--- MMA thread: [slot][phase] ready-seen / issued')
for j in range(8):
    row = t[2, j]
    if not row.any(): break
    print(f' j={j}: ' + ' '.join(f's{s}p{ph}:{row[s*8+ph*2]-base}/{row[s*8+ph*2+1]-base}' for s in (0, 1) for ph in (0, 1) if row[s*8+ph*2]))
print('--- producer: [slot] x_refilled / out_stored / y_refilled')
for j in range(8):
    row = t[3, j]
    if not row.any(): break
    print(f' j={j}: ' + ' '.join(f's{s}:{row[s*8]-base}/{row[s*8+1]-base}/{row[s*8+2]-base}' for s in (0, 1)))
