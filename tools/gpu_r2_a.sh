#!/bin/bash
# Round 2, GPU call A: is the plane kernel (k_layer_h) correct, and how does it compare with the round-1 kernels?
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== quick parity of the new path"
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -k "small_against_oracle or default_hparams or edge_shapes or golden or stress" > gpurun_out/a_t1.log 2>&1; echo "t1 rc=$?"
tail -12 gpurun_out/a_t1.log
run() {  # name, extra args...
  name=$1; shift
  timeout -k 5 150 python bench.py --steps 10 --no-cpu-baseline --no-e2e --sustain-s 1 "$@" > gpurun_out/a_bench_$name.json 2> gpurun_out/a_bench_$name.err
  echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/a_bench_%s.json' % sys.argv[1]))
    r = d['roofline']
    print('   ms/step %.3f  us/layer %.2f  frac %.3f  iso_us %.2f  sustained ms %.3f  clocks %s' % (d['ms_per_step'], r['us_per_layer'], r['frac'], r['isolated_launch_us'], d['sustained']['ms_per_step'], d['clocks']['sm_mhz']))
except Exception as e:
    print('   no line:', e)
PY
}
run new
run new_pk --debug variant=1
run new_noflags --debug tile_flags=0
run new_nopdl --debug pdl=0
run old --debug path=0
run old_rg --debug path=0 --debug variant=2
run new_bf16_c2 --precision bf16
run old_bf16_c2 --precision bf16 --debug path=0
echo "== traces"
timeout -k 5 100 python tools/tc_trace.py f16x3 2 > gpurun_out/a_trace_new_f16x3_l2.txt 2>&1; echo "trace rc=$?"
timeout -k 5 100 python tools/tc_trace.py bf16 2 > gpurun_out/a_trace_new_bf16_l2.txt 2>&1
timeout -k 5 100 python tools/tc_trace.py f16x3 2 path=0 variant=2 flow=0 > gpurun_out/a_trace_old_rg_f16x3_l2.txt 2>&1
echo "== full GPU suite"
timeout -k 5 1500 python -m pytest tests -m gpu -q -x > gpurun_out/a_t2.log 2>&1; echo "t2 rc=$?"
tail -15 gpurun_out/a_t2.log
echo "== c3"
run new_c3_bf16 --workload c3 --steps 5
run old_c3_bf16 --workload c3 --steps 5 --debug path=0
echo "== ncu"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/a_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:k_layer_h -s 40 -c 2 -o gpurun_out/a_prof_layer_h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0 > gpurun_out/a_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -30
