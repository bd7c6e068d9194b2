#!/usr/bin/env python
"""Which launch form of the gated layers is faster at which job size? One process, default hparams, T = 16000:
for every batch size N and arithmetic mode, the c2-style timed loop (L2 flushed before every step, CUDA events) with
  pdl    : one k_layer_tc launch per layer (640 threads, polled mbarriers), programmatic dependent launch, no tile flags
  layers : one k_flow_tc launch per layer (PWV_TC_SEG=1; 512-thread form unless PWV_TC_QUIET=0)
  flow   : one persistent k_flow_tc launch per flow, tiles chained by flags (PWV_TC_SEG=100)
usage: sweep_modes.py [N,N,...] [T]   -> one JSON line per (N, precision, mode) on stdout"""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam
W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder')
from oracle import iaf_oracle as O

ns = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '1,2,4,8,16,32').split(',')]
t = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
hp.set_hparam_yaml('bench/c2')
weights = W.init_weights(hp, seed=0)
dims = W.model_dims(hp)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
MODES = {'pdl': {'PWV_NO_TILE_FLAGS': '1'}, 'layers': {'PWV_TC_SEG': '1'}, 'flow': {'PWV_TC_SEG': '100'}}
for n in ns:
    noise, mel = O.synthetic_inputs(n, t, 80, 80)
    noise, mel = torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda()
    out = torch.empty((n, t), dtype=torch.float32, device='cuda')
    for prec in ('f16x3', 'bf16'):
        res = {}
        for mode, env in MODES.items():
            for k in ('PWV_NO_TILE_FLAGS', 'PWV_TC_FLOW', 'PWV_TC_SEG'):
                os.environ.pop(k, None)
            os.environ.update(env)
            m = V.PwvModel(dims, weights, prec)
            steps = 10 if n * t <= 2_000_000 else 4
            for _ in range(3):
                m.forward(noise, mel, out=out)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            for a, b in ev:
                flush.fill_(1)
                a.record(); m.forward(noise, mel, out=out); b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
            res[mode] = ms
            del m
        print(json.dumps({'N': n, 'T': t, 'precision': prec, 'tiles_per_cta_layer': round(n * ((t + 127) // 128) / 74, 1),
                          'ms_pdl': round(res['pdl'], 4), 'ms_layers': round(res['layers'], 4), 'ms_flow': round(res['flow'], 4),
                          'best': min(res, key=res.get), 'samples_per_s_best': round(n * t / min(res.values()) * 1e3)}), flush=True)
