set -x
timeout 120 ./tools/tc_probe 2>&1 | tee gpurun_out/tc_probe.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "f16x3 and small" 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -q -k "f16x3 or bf16" 2>&1 | tail -25
timeout 300 python bench.py --steps 10 --warmup 3 --precision f16x3 --no-cpu-baseline > gpurun_out/bench_f16x3.json 2> gpurun_out/bench_f16x3.err; cat gpurun_out/bench_f16x3.json; tail -5 gpurun_out/bench_f16x3.err
