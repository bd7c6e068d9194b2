#!/bin/bash
# Round 2, second 8-GPU call: where the hostshard e2e overhead goes (copy probe with / without the NUMA binding), then
# bench.py as the driver launches it with both e2e forms.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/mg${N}b_topo.txt 2>&1
for mode in "" "--no-bind"; do
  timeout -k 5 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/hostshard_probe.py $mode > gpurun_out/mg${N}b_probe${mode}.txt 2> gpurun_out/mg${N}b_probe${mode}.err
  echo "probe $mode rc=$?"; cat gpurun_out/mg${N}b_probe${mode}.txt
done
run() {  # name, args
  name=$1; shift
  timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/mg${N}b_$name.json 2> gpurun_out/mg${N}b_$name.err
  echo "bench $name rc=$?"
  python - "$N" "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/mg%sb_%s.json' % (sys.argv[1], sys.argv[2])))
    print('   value %.3e  ms/step %.3f  e2e %.3e (%s)  sustained %.3e  frac %.3f  bound cpus %s' % (d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else float('nan'),
          (d['e2e']['api'][:40] if d.get('e2e') else ''), d['sustained']['value'] if d.get('sustained') else float('nan'), d['roofline']['frac'], d['engine'].get('cpus_bound_to_gpu_socket')))
except Exception as e:
    print('   no line:', e)
    print(open('gpurun_out/mg%sb_%s.err' % (sys.argv[1], sys.argv[2])).read()[-1500:])
PY
}
run c4_hostshard
run c4_nccl --e2e-mode nccl
