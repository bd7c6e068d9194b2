#!/bin/bash
# Round 2, GPU call F: k_layer_h v4 (no staging barrier, E1 accumulator prefetch), bf16 gate A/B, normaliser tests
mkdir -p gpurun_out
timeout -k 5 90 python - > gpurun_out/f_tiny.log 2>&1 <<'PY'
import __graft_entry__ as g
g.smoke()
PY
echo "tiny rc=$?"; tail -1 gpurun_out/f_tiny.log
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -k "small_against_oracle or default_hparams or edge_shapes or golden or stress or instance_normalisers or bf16_mode or properties_at_full" > gpurun_out/f_t1.log 2>&1; echo "t1 rc=$?"
tail -4 gpurun_out/f_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/f_bench_$name.json 2> gpurun_out/f_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/f_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run v4
run v4_scalar --debug variant=0
run v4_again
run old --debug path=0
run v4_bf16_c2 --precision bf16
run v4_bf16_c2_ex2 --precision bf16 --debug variant=2
run v4_c3_bf16 --workload c3 --steps 5
run v4_c3_bf16_ex2 --workload c3 --steps 5 --debug variant=2
run v4_c4shard --workload c4 --steps 5
timeout -k 5 60 python tools/tc_trace.py f16x3 2 > gpurun_out/f_trace_f16x3_l2.txt 2>&1; echo "trace rc=$?"
timeout -k 5 60 python tools/tc_trace.py bf16 2 > gpurun_out/f_trace_bf16_l2.txt 2>&1
timeout -k 5 60 python tools/tc_trace.py bf16 2 variant=2 > gpurun_out/f_trace_bf16_ex2_l2.txt 2>&1
echo "== bf16 drift of the ex2 gate"
timeout -k 5 120 python - > gpurun_out/f_bf16_drift.log 2>&1 <<'PY'
import importlib, sys, numpy as np, torch
sys.path.insert(0, 'tests')
from conftest import pkg
from oracle import iaf_oracle as O
hp = pkg('hparam').hparam
hp.set_hparam_yaml('default')
W = pkg('weights'); V = pkg('vocoder')
weights = W.init_weights(hp, seed=0, bias_std=0.1)
noise, mel = O.synthetic_inputs(2, 4000, 80, 80)
d = W.model_dims(hp)
ref = O.iaf_vocoder_forward(noise, mel, weights, d['dilations'], d['hop'], dtype=np.float64)
for dbg in ({}, {'variant': 2}, {'path': 0}):
    out = V.PwvModel(d, weights, 'bf16', debug=dbg).forward(torch.from_numpy(noise).cuda(), torch.from_numpy(mel).cuda()).cpu().numpy()
    e = np.abs(out - ref)
    print('bf16', dbg, 'max %.3e rms %.3e' % (e.max(), np.sqrt((e ** 2).mean())), flush=True)
PY
cat gpurun_out/f_bf16_drift.log | tail -4
