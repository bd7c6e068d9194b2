#!/bin/bash
# A/B: gate phases of the two slots mutually exclusive (PWV_TC_E1LOCK); then compute-sanitizer on the small graphs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -q -x -k "flow_kernel" > gpurun_out/t_flow3.log 2>&1; RC=$?; echo "flow tests rc=$RC"; tail -3 gpurun_out/t_flow3.log
if [ $RC -ne 0 ]; then exit 1; fi
for prec in f16x3 bf16; do
  for lk in 0 1; do
    export PWV_TC_E1LOCK=$lk
    tag=${prec}_e1lock$lk
    timeout -k 5 120 python bench.py --steps 10 --precision $prec --no-cpu-baseline --no-e2e > gpurun_out/ab3_$tag.json 2> gpurun_out/ab3_$tag.err
    echo "bench $tag rc=$?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab3_$tag.json'))
    r = d['roofline']
    print('  ms/step', round(d['ms_per_step'], 4), 'chain us', round(r['avg_launch_us'], 2), 'frac', round(r['frac'], 4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('  no line:', e)
PY
  done
done
export PWV_TC_E1LOCK=1
PWV_TRACE_FLOW=1 timeout -k 5 100 python tools/tc_trace.py f16x3 2 > gpurun_out/trace_flow_e1lock_f16x3_l2.txt 2>&1
head -8 gpurun_out/trace_flow_e1lock_f16x3_l2.txt | cut -c1-330
unset PWV_TC_E1LOCK
timeout -k 5 240 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck_r1b.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_r1b.txt
timeout -k 5 240 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/sanitizer_synccheck_r1b.txt 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/sanitizer_synccheck_r1b.txt
