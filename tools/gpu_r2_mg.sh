#!/bin/bash
# Round 2, multi-GPU call: bench.py under torchrun exactly as the driver launches it (c4 sharded over the GPUs), both e2e forms
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() {  # name, args
  name=$1; shift
  timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/mg${N}_$name.json 2> gpurun_out/mg${N}_$name.err
  echo "bench $name rc=$?"
  python - "$N" "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/mg%s_%s.json' % (sys.argv[1], sys.argv[2])))
    print('   value %.3e  ms/step %.3f  e2e %.3e (%s)  sustained %.3e  frac %.3f  workload %s' % (d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else float('nan'),
          (d['e2e']['api'][:40] if d.get('e2e') else ''), d['sustained']['value'] if d.get('sustained') else float('nan'), d['roofline']['frac'], d['config']['workload'][:30]))
except Exception as e:
    print('   no line:', e)
    print(open('gpurun_out/mg%s_%s.err' % (sys.argv[1], sys.argv[2])).read()[-1500:])
PY
}
run c4_hostshard
run c4_nccl --e2e-mode nccl
run c2_hostshard --workload c2
if [ "$N" = "2" ]; then
  NCCL_DEBUG=INFO timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --sustain-s 0 > gpurun_out/mg${N}_ncclinfo.json 2> gpurun_out/mg${N}_ncclinfo.err; echo "nccl info rc=$?"
  head -c 600 gpurun_out/mg${N}_ncclinfo.json; echo; grep -c "NCCL INFO" gpurun_out/mg${N}_ncclinfo.err
  timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/mg${N}_ref.json 2> gpurun_out/mg${N}_ref.err; echo "ref arm rc=$?"
  head -c 900 gpurun_out/mg${N}_ref.json; echo
fi
