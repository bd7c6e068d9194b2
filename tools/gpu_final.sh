#!/bin/bash
# Round-end validation: full GPU parity suite, the default bench line, and the other workloads under the size-based launch policy.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_final.log 2>&1; RC=$?; echo "pytest -m gpu rc=$RC"; tail -4 gpurun_out/t_gpu_final.log
[ $RC -eq 124 ] || [ $RC -eq 137 ] && exit 1
timeout -k 5 100 python bench.py > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_final_c2.json
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d['roofline']
    print('  ms/step', round(d['ms_per_step'], 4), 'value %.3e' % d['value'], 'us/layer', round(r['avg_launch_us'], 2), 'frac', round(r['frac'], 4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'launches/step', d['gpu_launches'] // d['steps'])
except Exception as e:
    print('  no line:', e)
PY
}
for cfg in "c1 f16x3 20" "c3 bf16 4" "c3 f16x3 4"; do
  set -- $cfg
  timeout -k 5 60 python bench.py --workload $1 --precision $2 --steps $3 --no-cpu-baseline --no-e2e > gpurun_out/bench_final_$1_$2.json 2> gpurun_out/bench_final_$1_$2.err
  echo "bench $1 $2 rc=$?"; line gpurun_out/bench_final_$1_$2.json
done
