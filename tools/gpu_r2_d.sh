#!/bin/bash
# Round 2, GPU call D: wide-channel kernels (k_wide_h), stress-case diagnosis, k_layer_h with the K-splits off
mkdir -p gpurun_out
timeout -k 5 90 python - > gpurun_out/d_tiny.log 2>&1 <<'PY'
import __graft_entry__ as g
g.smoke()
PY
echo "tiny rc=$?"; tail -2 gpurun_out/d_tiny.log
echo "== stress-case diagnosis"
timeout -k 5 120 python - > gpurun_out/d_stress.log 2>&1 <<'PY'
import importlib, sys, numpy as np, torch
sys.path.insert(0, 'tests')
from conftest import pkg, small_case
from oracle import iaf_oracle as O
hp = pkg('hparam').hparam
hp.set_hparam_yaml('default')
small_case(hp, dilations=((1, 2, 4, 8, 16, 32, 64, 128, 256, 512),), t=2400, precision='f16x3')
W = pkg('weights'); V = pkg('vocoder')
for gain in (1.0, 2.0, 3.0):
    weights = W.init_weights(hp, seed=5, gain=gain)
    noise, mel = O.synthetic_inputs(2, 2400, 80, 80)
    d = W.model_dims(hp)
    ref = O.iaf_vocoder_forward(noise, mel, weights, d['dilations'], d['hop'], dtype=np.float64)
    for prec, dbg in (('fp32', {}), ('f16x3', {'path': 0, 'variant': 0}), ('f16x3', {'path': 1, 'variant': 0}), ('f16x3', {'path': 1, 'variant': 1})):
        out = V.PwvModel(d, weights, prec, debug=dbg).forward(torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda()).cpu().numpy()
        err = np.abs(out - ref)
        print('gain', gain, prec, dbg, 'max|ref| %.2f  max err %.3e  rel-to-max %.3e  rms %.3e' % (np.abs(ref).max(), err.max(), err.max() / np.abs(ref).max(), np.sqrt((err ** 2).mean())), flush=True)
PY
echo "stress rc=$?"; cat gpurun_out/d_stress.log | tail -14
echo "== wide channels"
timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -x -q -k "wide_channels or (small_against_oracle and (128 or 256))" > gpurun_out/d_t_wide.log 2>&1; WIDE=$?; echo "wide rc=$WIDE"
tail -8 gpurun_out/d_t_wide.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 100 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/d_bench_$name.json 2> gpurun_out/d_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/d_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run v3
run v3_again
run v3_bf16_c2 --precision bf16
if [ $WIDE -eq 0 ]; then
  timeout -k 5 400 python tools/sweep_c5.py --quick > gpurun_out/d_c5_quick.jsonl 2> gpurun_out/d_c5_quick.err; echo "c5 quick rc=$?"
  python - <<'PY'
import json
for line in open('gpurun_out/d_c5_quick.jsonl'):
    r = json.loads(line)
    print('   C=%d L=%d %-5s layer %.1f us  frac %.3f  %.0f TF' % (r['C'], r['L'], r['precision'], r['layer_launch_us'], r['frac_of_hbm_peak'], r['tflops_fp32_equiv']))
PY
fi
echo "== full GPU suite"
timeout -k 5 700 python -m pytest tests -m gpu -q > gpurun_out/d_t2.log 2>&1; echo "t2 rc=$?"
tail -8 gpurun_out/d_t2.log
