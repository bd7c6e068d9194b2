#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -q -x -k "flow_kernel" > gpurun_out/t_flow.log 2>&1; RC=$?; echo "flow test rc=$RC"; tail -5 gpurun_out/t_flow.log
[ $RC -eq 124 ] || [ $RC -eq 137 ] && exit 1
for prec in f16x3 bf16; do
  for st in 0 1; do
    export PWV_TC_STAGGER=$st
    timeout -k 5 150 python bench.py --steps 10 --precision $prec --no-cpu-baseline --no-e2e > gpurun_out/ab_${prec}_stagger_$st.json 2> gpurun_out/ab_${prec}_stagger_$st.err
    echo "bench $prec stagger=$st rc=$?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_${prec}_stagger_$st.json'))
    r = d['roofline']
    print('  ms/step', round(d['ms_per_step'], 4), 'chain us', round(r['avg_launch_us'], 2), 'frac', round(r['frac'], 4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'launches', d['gpu_launches'])
except Exception as e:
    print('  no line:', e)
PY
  done
done
unset PWV_TC_STAGGER
PWV_TC_FLOW=0 timeout -k 5 150 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/ab_f16x3_flow0.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ab_f16x3_flow0.json')); print('per-layer launches: ms/step', d['ms_per_step'])"
PWV_TRACE_FLOW=1 timeout -k 5 100 python tools/tc_trace.py f16x3 2 > gpurun_out/trace_flow_f16x3_l2.txt 2>&1
head -10 gpurun_out/trace_flow_f16x3_l2.txt
