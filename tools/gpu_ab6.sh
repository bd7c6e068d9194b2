#!/bin/bash
# Timing experiment (results undefined for the unsafe modes): what do the gpu-scope fences / the flags cost per tile?
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
  timeout -k 5 200 python bench.py --workload $2 --steps $4 --precision $3 --no-cpu-baseline --no-e2e > gpurun_out/ab6_$1.json 2> gpurun_out/ab6_$1.err
  echo "bench $1 rc=$?"; line gpurun_out/ab6_$1.json
}
for wl in c2 c3; do
  st=10; [ $wl = c3 ] && st=5
  for prec in f16x3 bf16; do
    for u in 0 1 2; do
      PWV_TC_DEBUG_UNSAFE=$u run ${wl}_${prec}_unsafe$u $wl $prec $st
    done
  done
done
