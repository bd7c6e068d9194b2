#!/bin/bash
# Round 2, last 8-GPU call: bench.py exactly as the driver launches it (c4 strong-scaled, default e2e form) on the final kernels, and the reference arm under torchrun
N=${1:-8}
mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/mg${N}d_c4.json 2> gpurun_out/mg${N}d_c4.err
echo "bench rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/mg${N}d_c4.json'))
print('   value %.3e  ms/step %.3f  e2e %.3e  sustained %.3e  frac %.3f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['sustained']['value'], d['roofline']['frac'], d['clocks']))
PY
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/mg${N}d_ref.json 2> gpurun_out/mg${N}d_ref.err; echo "ref arm rc=$?"
head -c 400 gpurun_out/mg${N}d_ref.json; echo
