#!/bin/bash
# A/B at the c1 and c3 sizes: per-layer launches (PDL chain / tile flags) vs the flow kernel (polled / named barriers, rotation).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
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
run() {  # tag workload precision steps
  timeout -k 5 200 python bench.py --workload $2 --steps $4 --precision $3 --no-cpu-baseline --no-e2e > gpurun_out/ab5_$1.json 2> gpurun_out/ab5_$1.err
  echo "bench $1 rc=$?"; line gpurun_out/ab5_$1.json
}
for wl in c3 c1; do
  st=5; [ $wl = c1 ] && st=20
  for prec in bf16 f16x3; do
    PWV_NO_TILE_FLAGS=1 run ${wl}_${prec}_layers_pdl $wl $prec $st
    PWV_TC_FLOW=0 run ${wl}_${prec}_layers_flags $wl $prec $st
    PWV_TC_QUIET=0 PWV_TC_ROTATE=0 run ${wl}_${prec}_flow_q0r0 $wl $prec $st
    PWV_TC_QUIET=1 PWV_TC_ROTATE=0 run ${wl}_${prec}_flow_q1r0 $wl $prec $st
    PWV_TC_QUIET=1 PWV_TC_ROTATE=1 run ${wl}_${prec}_flow_q1r1 $wl $prec $st
  done
done
