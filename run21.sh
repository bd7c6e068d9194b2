set -x
python bench.py > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final.err; cut -c1-250 gpurun_out/bench_final_c2.json
python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/bench_final_c2_bf16.json 2>> gpurun_out/bench_final.err; cut -c1-250 gpurun_out/bench_final_c2_bf16.json
python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/bench_final_c3_bf16.json 2>> gpurun_out/bench_final.err; cut -c1-250 gpurun_out/bench_final_c3_bf16.json
python bench.py --steps 5 --warmup 3 --workload c3 --precision f16x3 --no-cpu-baseline --no-e2e > gpurun_out/bench_final_c3_f16x3.json 2>> gpurun_out/bench_final.err; cut -c1-250 gpurun_out/bench_final_c3_f16x3.json
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/bench_final_c1.json 2>> gpurun_out/bench_final.err; cut -c1-250 gpurun_out/bench_final_c1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_layer_tc -s 130 -c 2 -o gpurun_out/prof_layer_tc_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none -k regex:k_post_tc -s 3 -c 1 -o gpurun_out/prof_post_tc_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1
python tools/sweep_c5.py > gpurun_out/c5_sweep.jsonl 2> gpurun_out/c5.err; wc -l gpurun_out/c5_sweep.jsonl; tail -2 gpurun_out/c5.err
tail -3 gpurun_out/bench_final.err
