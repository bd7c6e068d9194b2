#!/bin/bash
# Round 2, GPU call I: boxes -> TMEM by tcgen05.cp (switch cp=1) against the workers' copy
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -x -q -k "layer_h_switches" > gpurun_out/i_t1.log 2>&1; OK=$?; echo "t1 rc=$OK"
tail -6 gpurun_out/i_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/i_bench_$name.json 2> gpurun_out/i_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/i_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run base
run cp --debug cp=1
run base2
run cp2 --debug cp=1
run bf16_base --precision bf16
run bf16_cp --precision bf16 --debug cp=1
run c3_bf16_cp --workload c3 --steps 5 --debug cp=1
run c3_bf16_base --workload c3 --steps 5
run c4shard_cp --workload c4 --steps 5 --debug cp=1
timeout -k 5 60 python tools/tc_trace.py f16x3 2 cp=1 > gpurun_out/i_trace_cp_f16x3_l2.txt 2>&1; echo "trace rc=$?"
timeout -k 5 60 python tools/tc_trace.py bf16 2 cp=1 > gpurun_out/i_trace_cp_bf16_l2.txt 2>&1
