#!/bin/bash
# A/B: HOIST form of the epilogues; then the other BASELINE workloads with the round's final kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PWV_TC_HOIST=1 timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -q -x -k "flow_kernel" > gpurun_out/t_flow4.log 2>&1; RC=$?; echo "flow tests (HOIST) rc=$RC"; tail -3 gpurun_out/t_flow4.log
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d['roofline']
    print('  ms/step', round(d['ms_per_step'], 4), 'value %.3e' % d['value'], 'chain us', round(r['avg_launch_us'], 2), 'frac', round(r['frac'], 4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('  no line:', e)
PY
}
if [ $RC -eq 0 ]; then
for prec in f16x3 bf16; do
  for h in 0 1 0 1; do
    export PWV_TC_HOIST=$h
    tag=${prec}_hoist$h
    timeout -k 5 120 python bench.py --steps 10 --precision $prec --no-cpu-baseline --no-e2e > gpurun_out/ab4_$tag.json 2> gpurun_out/ab4_$tag.err
    echo "bench $tag rc=$?"; line gpurun_out/ab4_$tag.json
  done
done
fi
unset PWV_TC_HOIST
for wl in c1 c3; do
  for prec in f16x3 bf16; do
    timeout -k 5 200 python bench.py --workload $wl --steps 10 --precision $prec --no-cpu-baseline > gpurun_out/final_${wl}_$prec.json 2> gpurun_out/final_${wl}_$prec.err
    echo "bench $wl $prec rc=$?"; line gpurun_out/final_${wl}_$prec.json
  done
done
