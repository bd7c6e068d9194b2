#!/bin/bash
# Round 2, GPU call G: k_layer_h v5 (one mbarrier arrival per warp, suspend-hint waits)
mkdir -p gpurun_out
timeout -k 5 90 python - > gpurun_out/g_tiny.log 2>&1 <<'PY'
import __graft_entry__ as g
g.smoke()
PY
OK=$?; echo "tiny rc=$OK"; tail -1 gpurun_out/g_tiny.log
[ $OK -eq 0 ] || exit 1
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -k "small_against_oracle or default_hparams or edge_shapes or golden or stress or bf16_mode or properties_at_full or variants_bit or use_skip or transposed" > gpurun_out/g_t1.log 2>&1; echo "t1 rc=$?"
tail -4 gpurun_out/g_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/g_bench_$name.json 2> gpurun_out/g_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/g_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run v5
run old --debug path=0
run v5_again
run v5_bf16_c2 --precision bf16
run v5_c3_bf16 --workload c3 --steps 5
run v5_c4shard --workload c4 --steps 5
run v5_c1 --workload c1
timeout -k 5 60 python tools/tc_trace.py f16x3 2 > gpurun_out/g_trace_f16x3_l2.txt 2>&1; echo "trace rc=$?"
timeout -k 5 60 python tools/tc_trace.py bf16 2 > gpurun_out/g_trace_bf16_l2.txt 2>&1
