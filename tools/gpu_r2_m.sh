#!/bin/bash
# Round 2, GPU call M: f16x3 with z in the dead gate-accumulator columns (early copy of the next tile) against the old order
mkdir -p gpurun_out
timeout -k 5 90 python - > gpurun_out/m_tiny.log 2>&1 <<'PY'
import importlib, numpy as np, torch
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam; W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder'); IO = importlib.import_module(P + '.io')
for skip in (False, True):
    hp.set_hparam_dict({'model': {'n_iaf': 2, 'dilations': [[1, 2, 4, 512], [1, 8]], 'use_skip_connection': skip}, 'generate': {'batch_size': 3, 'length': 4000}}, case='t%d' % skip)
    d = W.model_dims(hp); w = W.init_weights(hp, seed=1, bias_std=0.1)
    for n, t in ((3, 4000), (1, 320), (8, 16000)):
        nz, ml = IO.synthetic_batch(n, t, 80, 80)
        a = V.PwvModel(d, w, 'f16x3').forward(torch.from_numpy(nz).cuda(), torch.from_numpy(ml).cuda())
        b = V.PwvModel(d, w, 'f16x3', debug={'z_in_d': 0}).forward(torch.from_numpy(nz).cuda(), torch.from_numpy(ml).cuda())
        c = V.PwvModel(d, w, 'f16x3', debug={'split2': 1}).forward(torch.from_numpy(nz).cuda(), torch.from_numpy(ml).cuda())
        torch.cuda.synchronize(); print(skip, n, t, 'z_in_d == old:', torch.equal(a, b), float((a - b).abs().max()), ' split2 delta', float((a - c).abs().max()), flush=True)
PY
OK=$?; echo "tiny rc=$OK"; tail -6 gpurun_out/m_tiny.log
[ $OK -eq 0 ] || exit 1
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -x -q -k "switches or variants or golden or taps or stress or skip" > gpurun_out/m_t1.log 2>&1; echo "t1 rc=$?"
tail -4 gpurun_out/m_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 100 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/m_bench_$name.json 2> gpurun_out/m_bench_$name.err
  rc=$?; echo "bench $name rc=$rc"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/m_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
  return $rc
}
run c2_zd || exit 1
run c2_old --debug z_in_d=0
run c2_zd2
run c2_old2 --debug z_in_d=0
run c4shard_zd --workload c4 --steps 5
run c4shard_old --workload c4 --steps 5 --debug z_in_d=0
run c1_zd --workload c1
run c1_old --workload c1 --debug z_in_d=0
timeout -k 5 60 python tools/tc_trace.py f16x3 2 > gpurun_out/m_trace_f16x3_l2.txt 2>&1; echo "trace rc=$?"
