#!/bin/bash
# Round 2, GPU call B: plane kernel with the corrected register split. Every step has a short timeout; the new-kernel
# steps are skipped if the first tiny forward does not come back.
mkdir -p gpurun_out
echo "== tiny forward on the new path"
timeout -k 5 90 python - > gpurun_out/b_tiny.log 2>&1 <<'PY'
import __graft_entry__ as g
g.smoke()
PY
NEW_OK=$?; echo "tiny rc=$NEW_OK"; tail -3 gpurun_out/b_tiny.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 100 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/b_bench_$name.json 2> gpurun_out/b_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/b_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
if [ $NEW_OK -eq 0 ]; then
  echo "== quick parity of the new path"
  timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -x -q -k "small_against_oracle or default_hparams or edge_shapes or golden or stress" > gpurun_out/b_t1.log 2>&1; echo "t1 rc=$?"
  tail -12 gpurun_out/b_t1.log
  run new
  run new_pk --debug variant=1
  run new_noflags --debug tile_flags=0
  run new_bf16_c2 --precision bf16
  timeout -k 5 60 python tools/tc_trace.py f16x3 2 > gpurun_out/b_trace_new_f16x3_l2.txt 2>&1; echo "trace rc=$?"
  timeout -k 5 60 python tools/tc_trace.py bf16 2 > gpurun_out/b_trace_new_bf16_l2.txt 2>&1
  run new_c3_bf16 --workload c3 --steps 5
  run new_c1 --workload c1
fi
echo "== round-1 kernels with the register re-partition"
timeout -k 5 60 python - > gpurun_out/b_tiny_rg.log 2>&1 <<'PY'
import importlib, numpy as np, torch
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam; W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder'); IO = importlib.import_module(P + '.io')
hp.set_hparam_dict({'model': {'n_iaf': 1, 'dilations': [[1, 2, 4]]}, 'generate': {'batch_size': 2, 'length': 1600}}, case='t')
d = W.model_dims(hp); w = W.init_weights(hp, seed=1)
n, m = IO.synthetic_batch(2, 1600, 80, 80)
a = V.PwvModel(d, w, 'f16x3', debug={'path': 0, 'variant': 2}).forward(torch.from_numpy(n).cuda(), torch.from_numpy(m).cuda())
b = V.PwvModel(d, w, 'f16x3', debug={'path': 0, 'variant': 0}).forward(torch.from_numpy(n).cuda(), torch.from_numpy(m).cuda())
torch.cuda.synchronize(); print('rg == plain:', torch.equal(a, b))
PY
RG_OK=$?; echo "tiny rg rc=$RG_OK"; tail -2 gpurun_out/b_tiny_rg.log
if [ $RG_OK -eq 0 ]; then
  run old_rg --debug path=0 --debug variant=2
  timeout -k 5 60 python tools/tc_trace.py f16x3 2 path=0 variant=2 trace_flow=1 > gpurun_out/b_trace_old_rg_f16x3_l2.txt 2>&1
fi
run old --debug path=0
echo "== full GPU suite"
timeout -k 5 900 python -m pytest tests -m gpu -q -x > gpurun_out/b_t2.log 2>&1; echo "t2 rc=$?"
tail -15 gpurun_out/b_t2.log
if [ $NEW_OK -eq 0 ]; then
  echo "== ncu"
  timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/b_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:k_layer_h -s 40 -c 2 -o gpurun_out/b_prof_layer_h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/b_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls gpurun_out | grep "^b_" | tr '\n' ' '
