set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "c3_size" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b2.out 2> gpurun_out/b2.err; wc -l gpurun_out/b2.out; cut -c1-120 gpurun_out/b2.out; grep -c "NCCL version" gpurun_out/b2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/b2ref.out 2>> gpurun_out/b2.err; wc -l gpurun_out/b2ref.out; cut -c1-120 gpurun_out/b2ref.out
