#!/bin/bash
# Round 2, GPU call C: k_layer_h v2 (GEMM1 / GEMM2 K-splits, completion counters, probes one tile ahead)
mkdir -p gpurun_out
timeout -k 5 90 python - > gpurun_out/c_tiny.log 2>&1 <<'PY'
import __graft_entry__ as g
g.smoke()
PY
NEW_OK=$?; echo "tiny rc=$NEW_OK"; tail -2 gpurun_out/c_tiny.log
[ $NEW_OK -eq 0 ] || exit 1
timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -x -q -k "small_against_oracle or default_hparams or edge_shapes or golden or stress or use_skip or properties_at_full" > gpurun_out/c_t1.log 2>&1; echo "t1 rc=$?"
tail -5 gpurun_out/c_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 100 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/c_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run v2
run v2_s10 --debug split2=0
run v2_s01 --debug split1=0
run v2_s00 --debug split1=0 --debug split2=0
run v2_scalar --debug variant=0
run v2_noflags --debug tile_flags=0
run v2_again
run v2_bf16_c2 --precision bf16
run v2_c3_bf16 --workload c3 --steps 5
run v2_c3_f16x3 --workload c3 --steps 5 --precision f16x3
run v2_c1 --workload c1
timeout -k 5 60 python tools/tc_trace.py f16x3 2 > gpurun_out/c_trace_f16x3_l2.txt 2>&1; echo "trace rc=$?"
timeout -k 5 60 python tools/tc_trace.py f16x3 7 > gpurun_out/c_trace_f16x3_l7.txt 2>&1
timeout -k 5 60 python tools/tc_trace.py bf16 2 > gpurun_out/c_trace_bf16_l2.txt 2>&1
echo "== full GPU suite"
timeout -k 5 600 python -m pytest tests -m gpu -q -x > gpurun_out/c_t2.log 2>&1; echo "t2 rc=$?"
tail -6 gpurun_out/c_t2.log
echo "== default bench line (cpu baseline + e2e)"
timeout -k 5 200 python bench.py > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err; echo "default rc=$?"
echo "== ncu"
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/c_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:k_layer_h -s 40 -c 2 -o gpurun_out/c_prof_layer_h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/c_ncu_full.log 2>&1; echo "ncu full rc=$?"
