#!/bin/bash
# Round 2, third 8-GPU call: the forward legs of the copy probe, then bench.py with the device size cached at finalize
N=${1:-8}
mkdir -p gpurun_out
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/hostshard_probe.py > gpurun_out/mg${N}c_probe.txt 2> gpurun_out/mg${N}c_probe.err
echo "probe rc=$?"; cat gpurun_out/mg${N}c_probe.txt; tail -3 gpurun_out/mg${N}c_probe.err
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg${N}c_c4_hostshard.json 2> gpurun_out/mg${N}c_c4_hostshard.err
echo "bench rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/mg${N}c_c4_hostshard.json'))
print('   value %.3e  ms/step %.3f  e2e %.3e  sustained %.3e  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['sustained']['value'], d['roofline']['frac']))
PY
