#!/bin/bash
# A/B of the gated-layer kernel variants on one B200 (run under gpurun): bit-identity tests, bench lines, traces.
mkdir -p gpurun_out
T=tests/test_gpu_parity.py::test_layer_kernel_variants_bit_identical
timeout -k 5 300 python -m pytest "$T[3-f16x3]" "$T[3-bf16]" -q -x > gpurun_out/t_var3.log 2>&1; RC3=$?; echo "var3 rc=$RC3"; tail -3 gpurun_out/t_var3.log
VARS="0"; [ $RC3 -ne 124 ] && [ $RC3 -ne 137 ] && VARS="0 3"
for prec in f16x3 bf16; do
  for v in $VARS; do
    timeout -k 5 150 python bench.py --steps 10 --tc-variant $v --precision $prec --no-cpu-baseline --no-e2e > gpurun_out/ab_${prec}_v$v.json 2> gpurun_out/ab_${prec}_v$v.err
    echo "bench $prec v$v rc=$?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_${prec}_v$v.json'))
    r = d['roofline']
    print('  ms/step', round(d['ms_per_step'], 4), 'chain us', round(r['avg_launch_us'], 2), 'iso us', round(r['isolated_launch_us'], 2), 'frac', round(r['frac'], 4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('  no line:', e)
PY
  done
done
for v in $VARS; do
  PWV_TC_VARIANT=$v timeout -k 5 120 python tools/tc_trace.py f16x3 2 > gpurun_out/trace_f16x3_v$v.txt 2>&1
done
[ -f gpurun_out/trace_f16x3_v3.txt ] && head -24 gpurun_out/trace_f16x3_v3.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
