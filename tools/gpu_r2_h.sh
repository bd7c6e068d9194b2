#!/bin/bash
# Round 2, GPU call H: tcgen05.cp probe, unconditional-graph fixture on the GPU, c1 with / without tile flags
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cp_probe tools/cp_probe.cu > gpurun_out/h_cp_build.log 2>&1 && timeout -k 5 60 /tmp/cp_probe > gpurun_out/h_cp_probe.txt 2>&1
echo "cp probe rc=$?"; cat gpurun_out/h_cp_probe.txt
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or instance_normalisers or small_against_oracle" > gpurun_out/h_t1.log 2>&1; echo "t1 rc=$?"
tail -4 gpurun_out/h_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/h_bench_$name.json 2> gpurun_out/h_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/h_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f @ %s MHz  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['sustained']['clocks'].get('sm_mhz'), d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run c1 --workload c1
run c1_old --workload c1 --debug path=0
run c2
