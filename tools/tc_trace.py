#!/usr/bin/env python
"""Print the phase timeline (clock64 deltas) of CTA 0 of one tensor-core gated-layer launch.
usage: tc_trace.py [precision] [layer_index] [KEY=INT ...]   (default f16x3, layer 2 of flow 0, d=4; the c2 batch)
KEY=INT are pwv_debug_set switches (path=0 traces the round-1 kernels)."""
import ctypes, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam
W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder'); L = importlib.import_module(P + '._lib')
IO = importlib.import_module(P + '.io')
args = [a for a in sys.argv[1:] if '=' not in a]
debug = {a.split('=')[0]: int(a.split('=')[1]) for a in sys.argv[1:] if '=' in a}
prec = args[0] if len(args) > 0 else 'f16x3'
launch = int(args[1]) if len(args) > 1 else 2
hp.set_hparam_yaml('bench/c2')
weights = W.init_weights(hp, seed=0)
m = V.PwvModel(W.model_dims(hp), weights, prec, debug=debug)
noise, mel = IO.synthetic_batch(8, 16000, 80, 80)
noise, mel = torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda()
buf = torch.zeros(4 * 16 * 16, dtype=torch.int64, device='cuda')
for it in range(2):
    m.forward(noise, mel)
L.check(m.lib.pwv_debug_set_trace(m._h, ctypes.c_void_p(buf.data_ptr()), launch))
m.forward(noise, mel)
torch.cuda.synchronize()
L.check(m.lib.pwv_debug_set_trace(m._h, None, -1))
t = buf.cpu().numpy().reshape(4, 16, 16)
base = t[t > 0].min()
if debug.get('path', 1) == 1:      # k_layer_h (pwv_tc2.cuh)
    names = ['enter', '-', '-', '-', 'first_copy_done', 'd1_seen', 'z_ready', 'd2_seen', 'out_staged', 'd2_x_loaded', 'next_copied', '-', '-', '-']
else:
    names = ['enter', 'x_landed', 'x_prepped', 'y_landed', 'a_ready', 'd1_ready', 'z_ready', 'd2_ready', 'out_ready', 'e2_loaded', 'e2_done', 'copied', 'e1_ld0', 'e1_ld1']
print(f'precision {prec}, switches {debug}, gated layer {launch}; cycles since the first stamp (delta from the previous event)')
for role in (0, 1):
    print(f'--- worker slot {role}')
    for j in range(8):
        row = t[role, j]
        if not row.any(): break
        s, prev = [], row[row > 0].min()
        for k in sorted(range(len(names)), key=lambda k: row[k]):
            if row[k] == 0: continue
            s.append(f'{names[k]}={row[k]-base}(+{row[k]-prev})')
            prev = row[k]
        print(f' tile {j}: ' + ' '.join(s))
print('--- MMA thread: [slot][phase] ready-seen / issued')
for j in range(8):
    row = t[2, j]
    if not row.any(): break
    print(f' j={j}: ' + ' '.join(f's{s}p{ph}:{row[s*8+ph*2]-base}/{row[s*8+ph*2+1]-base}' for s in (0, 1) for ph in (0, 1) if row[s*8+ph*2]))
print('--- producers: [slot] x_refilled / - / stored+y_refilled')
for j in range(8):
    row = t[3, j]
    if not row.any(): break
    print(f' j={j}: ' + ' '.join(f's{s}:{row[s*8]-base}/{row[s*8+1]-base}/{row[s*8+2]-base}' for s in (0, 1)))
del m
