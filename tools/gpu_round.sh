#!/bin/bash
# One gpurun call: full GPU parity suite, default bench line, ncu launch list + full capture of the flow kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_all.log 2>&1; RC=$?; echo "pytest -m gpu rc=$RC"; tail -6 gpurun_out/t_gpu_all.log
[ $RC -eq 124 ] || [ $RC -eq 137 ] && exit 1
timeout -k 5 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_flow.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 240 ncu --set full --clock-control none --import-source on -k regex:k_flow_tc -s 4 -c 1 -o gpurun_out/prof_flow_f16x3 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
