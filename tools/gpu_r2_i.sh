#!/bin/bash
# Round 2, GPU call I: boxes -> TMEM by tcgen05.cp (switch cp=1) against the workers' copy. Everything after the first
# tiny cp forward is skipped if that does not come back.
mkdir -p gpurun_out
timeout -k 5 60 python - > gpurun_out/i_tiny.log 2>&1 <<'PY'
import importlib, numpy as np, torch
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam; W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder'); IO = importlib.import_module(P + '.io')
hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': [[1, 2, 4, 512], [1, 8]]}, 'generate': {'batch_size': 3, 'length': 4000}}, case='t')
d = W.model_dims(hp); w = W.init_weights(hp, seed=1, bias_std=0.1)
n, m = IO.synthetic_batch(3, 4000, 80, 80)
for prec in ('f16x3', 'bf16'):
    a = V.PwvModel(d, w, prec, debug={'cp': 1}).forward(torch.from_numpy(n).cuda(), torch.from_numpy(m).cuda())
    b = V.PwvModel(d, w, prec).forward(torch.from_numpy(n).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize(); print(prec, 'cp == copy:', torch.equal(a, b), float((a - b).abs().max()))
PY
OK=$?; echo "tiny cp rc=$OK"; cat gpurun_out/i_tiny.log | tail -3
[ $OK -eq 0 ] || exit 1
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -x -q -k "layer_h_switches" > gpurun_out/i_t1.log 2>&1; echo "t1 rc=$?"
tail -4 gpurun_out/i_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 100 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/i_bench_$name.json 2> gpurun_out/i_bench_$name.err
  rc=$?; echo "bench $name rc=$rc"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/i_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
  return $rc
}
run cp --debug cp=1 || exit 1
run base
run cp2 --debug cp=1
run bf16_cp --precision bf16 --debug cp=1
run bf16_base --precision bf16
run c3_bf16_cp --workload c3 --steps 5 --debug cp=1
run c4shard_cp --workload c4 --steps 5 --debug cp=1
timeout -k 5 60 python tools/tc_trace.py f16x3 2 cp=1 > gpurun_out/i_trace_cp_f16x3_l2.txt 2>&1; echo "trace rc=$?"
timeout -k 5 60 python tools/tc_trace.py bf16 2 cp=1 > gpurun_out/i_trace_cp_bf16_l2.txt 2>&1
