#!/bin/bash
# A/B of the k_flow_tc hand-off style (PWV_TC_QUIET) and the per-layer tile rotation (PWV_TC_ROTATE).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -q -x -k "flow_kernel" > gpurun_out/t_flow2.log 2>&1; RC=$?; echo "flow tests rc=$RC"; tail -15 gpurun_out/t_flow2.log
if [ $RC -ne 0 ]; then echo "flow tests failed: benches skipped"; exit 1; fi
for prec in f16x3 bf16; do
  for cfg in "0 0" "1 0" "0 1" "1 1"; do
    set -- $cfg
    export PWV_TC_QUIET=$1 PWV_TC_ROTATE=$2
    tag=${prec}_q$1_r$2
    timeout -k 5 120 python bench.py --steps 10 --precision $prec --no-cpu-baseline --no-e2e > gpurun_out/ab2_$tag.json 2> gpurun_out/ab2_$tag.err
    echo "bench $tag rc=$?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab2_$tag.json'))
    r = d['roofline']
    print('  ms/step', round(d['ms_per_step'], 4), 'chain us', round(r['avg_launch_us'], 2), 'frac', round(r['frac'], 4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'launches', d['gpu_launches'])
except Exception as e:
    print('  no line:', e)
PY
  done
done
export PWV_TC_QUIET=1 PWV_TC_ROTATE=0
PWV_TRACE_FLOW=1 timeout -k 5 100 python tools/tc_trace.py f16x3 2 > gpurun_out/trace_flow_quiet_f16x3_l2.txt 2>&1
head -12 gpurun_out/trace_flow_quiet_f16x3_l2.txt | cut -c1-330
