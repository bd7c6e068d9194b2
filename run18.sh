set -x
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/sweep_c5.py --quick > gpurun_out/c5_sweep_quick.jsonl 2> gpurun_out/c5.err; tail -3 gpurun_out/c5.err; wc -l gpurun_out/c5_sweep_quick.jsonl
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-300
python generate.py bench/c1 2>&1 | tail -5
ls -la gpurun_out/logdir/bench/c1/ | head
