#!/usr/bin/env python
"""BASELINE config c5: single-flow dilation-depth sweep (10/20/30 layers) x channel sweep (64/128/256).

For every (L, C) point: one flow (scaler + shifter bodies) with dilations = first L of [1..512]x3,
R = D = C, S = 2C, N*T >= 4M samples so the activations (>= 1 GB per buffer) cannot sit in L2. Reports
the gated-layer kernel's device time per launch (CUDA events inside the library), algorithmic GB/s
(2 bodies x N*T x 2*C*4 bytes per layer launch) against the measured HBM peak, and fp32-equivalent
TFLOP/s. Every channel count runs on the tcgen05 path (f16x3: k_layer_h at C = 64, the streamed-K k_wide_h gate + dense
passes at C = 128 / 256 -- there "layer_launch_us" is the time of a layer's two passes) and on the fp32 FFMA path.

    python tools/sweep_c5.py [--quick] > profiles/r2_c5_sweep.jsonl
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
P = 'parallel-wavenet-vocoder_b200'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true', help='N*T = 1M instead of 4M')
    args = ap.parse_args()
    hp = importlib.import_module(P + '.hparam').hparam
    W = importlib.import_module(P + '.weights')
    V = importlib.import_module(P + '.vocoder')
    peak = 6650.0
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk))['hbm_gbs'])
    n, t = (16, 64000) if args.quick else (64, 64000)
    base = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 3
    out = []
    for c in (64, 128, 256):
        for layers in (10, 20, 30):
            for prec in ('f16x3', 'fp32'):
                hp.set_hparam_dict({'model': {'n_iaf': 1, 'dilations': [base[:layers]], 'residual_channels': c,
                                              'dilation_channels': c, 'skip_channels': 2 * c},
                                    'generate': {'batch_size': n, 'length': t}}, case='c5')
                dims = W.model_dims(hp)
                model = V.PwvModel(dims, W.init_weights(hp, seed=0), prec)
                g = torch.Generator(device='cuda').manual_seed(1)
                noise = torch.randn((n, t), device='cuda', generator=g)
                mel = torch.rand((n, 1 + t // 80, 80), device='cuda', generator=g) * 2 - 1
                outbuf = torch.empty((n, t), device='cuda')
                model.forward(noise, mel, out=outbuf)
                model.set_profiling(True)
                times = []
                for _ in range(2):
                    model.forward(noise, mel, out=outbuf)
                    layer_ms, launches, fwd_ms = model.profile_read()
                    times.append((layer_ms, launches, fwd_ms))
                model.set_profiling(False)
                layer_ms, launches, fwd_ms = min(times)
                bytes_per_launch = 2 * n * t * 2 * c * 4
                mac_per_launch = 2 * n * t * (2 * c * 2 * c + c * c)
                us = layer_ms * 1e3 / launches
                rec = {'C': c, 'L': layers, 'precision': prec, 'N': n, 'T': t, 'layer_launch_us': us, 'launches': launches,
                       'algorithmic_GBps': bytes_per_launch / (us * 1e-6) / 1e9, 'hbm_peak_GBps': peak,
                       'frac_of_hbm_peak': bytes_per_launch / (us * 1e-6) / 1e9 / peak,
                       'tflops_fp32_equiv': 2 * mac_per_launch / (us * 1e-6) / 1e12,
                       'samples_per_s_forward': n * t / (fwd_ms * 1e-3), 'finite': bool(torch.isfinite(outbuf).all())}
                out.append(rec)
                print(json.dumps(rec), flush=True)
                del model
                torch.cuda.empty_cache()
    return out


if __name__ == '__main__':
    main()
