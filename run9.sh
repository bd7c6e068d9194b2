set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "f16x3 or bf16" 2>&1 | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 --precision f16x3 --no-cpu-baseline > gpurun_out/bench_f16x3_v5.json 2> gpurun_out/bench_f16x3_v5.err; cat gpurun_out/bench_f16x3_v5.json; tail -5 gpurun_out/bench_f16x3_v5.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline --no-e2e > gpurun_out/bench_bf16_v5.json 2>&1; cat gpurun_out/bench_bf16_v5.json
timeout 120 python tools/tc_trace.py f16x3 5 2>&1 | tee gpurun_out/tc_trace_v5_l5.txt | tail -40
timeout 120 python tools/tc_trace.py f16x3 11 2>&1 | tee gpurun_out/tc_trace_v5_l11.txt | tail -40
