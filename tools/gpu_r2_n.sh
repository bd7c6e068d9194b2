#!/bin/bash
# Round 2, GPU call N: round-end evidence again on the final kernel source (whole suite, default line, ncu captures, workloads)
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -q -m gpu > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/n_tests.log
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/n_smoke.log
S=$(date +%s); timeout -k 5 600 python bench.py > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err; echo "default bench rc=$? in $(( $(date +%s) - S )) s"
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/n_bench_reference.json 2> gpurun_out/n_bench_reference.err; echo "reference arm rc=$?"
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/n_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:k_layer_h -s 40 -c 2 -o gpurun_out/n_prof_layer_h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/n_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:k_layer_h -s 40 -c 2 -o gpurun_out/n_prof_layer_h_bf16 -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/n_ncu_full_bf16.log 2>&1; echo "ncu full bf16 rc=$?"
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 150 python bench.py --steps 10 --no-cpu-baseline --sustain-s 1.5 "$@" > gpurun_out/n_bench_$name.json 2> gpurun_out/n_bench_$name.err
  rc=$?; echo "bench $name rc=$rc"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/n_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   %s value %.3e  e2e %.3e  ms/step %.3f  us/layer %.2f  frac %.3f  sustained ms %.3f @ %s MHz  clocks %s' % (d['engine']['precision'], d['value'], d['e2e']['value'], d['ms_per_step'], r['us_per_layer'], r['frac'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run c2_f16x3
run c2_bf16 --precision bf16
run c1 --workload c1
run c3_bf16 --workload c3 --precision bf16 --steps 5
run c3_f16x3 --workload c3 --precision f16x3 --steps 5
run c4shard --workload c4 --steps 5
python - <<'PY'
import json
d = json.load(open('gpurun_out/n_bench_default.json'))
print('default line: value %.3e e2e %.3e cpu %.3e frac %.3f traffic %s clocks %s launches %s' % (d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['gpu_launches']))
PY
timeout -k 5 400 python tools/sweep_c5.py --quick > gpurun_out/n_c5_quick.jsonl 2> gpurun_out/n_c5_quick.err; echo "c5 quick rc=$?"; tail -3 gpurun_out/n_c5_quick.jsonl | cut -c1-300
