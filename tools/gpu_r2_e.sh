#!/bin/bash
# Round 2, GPU call E: 'in' normalisers, full GPU suite, full-size c5 sweep, ncu of the wide kernels
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -x -q -k "instance_normalisers or golden" > gpurun_out/e_t_norm.log 2>&1; echo "norm rc=$?"
tail -12 gpurun_out/e_t_norm.log
echo "== full GPU suite"
timeout -k 5 800 python -m pytest tests -m gpu -q > gpurun_out/e_t2.log 2>&1; echo "t2 rc=$?"
tail -8 gpurun_out/e_t2.log
echo "== c5 sweep, N*T = 4M"
timeout -k 5 900 python tools/sweep_c5.py > gpurun_out/e_c5.jsonl 2> gpurun_out/e_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/e_c5.jsonl'):
    r = json.loads(line)
    print('   C=%d L=%d %-5s layer %.1f us  frac %.3f  %.0f TF  fwd %.3g samples/s' % (r['C'], r['L'], r['precision'], r['layer_launch_us'], r['frac_of_hbm_peak'], r['tflops_fp32_equiv'], r['samples_per_s_forward']))
PY
echo "== ncu of the wide passes (C = 128, quick size)"
cat > /tmp/wide_once.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
P = 'parallel-wavenet-vocoder_b200'
hp = importlib.import_module(P + '.hparam').hparam; W = importlib.import_module(P + '.weights'); V = importlib.import_module(P + '.vocoder'); IO = importlib.import_module(P + '.io')
c = int(sys.argv[1])
hp.set_hparam_dict({'model': {'n_iaf': 1, 'dilations': [[1, 2, 4, 8, 16, 32, 64, 128, 256, 512]], 'residual_channels': c, 'dilation_channels': c, 'skip_channels': 2 * c},
                    'generate': {'batch_size': 16, 'length': 64000}}, case='c5')
d = W.model_dims(hp)
m = V.PwvModel(d, W.init_weights(hp, seed=0), 'f16x3')
n, mel = IO.synthetic_batch(16, 64000, 80, 80)
n, mel = torch.from_numpy(n).cuda(), torch.from_numpy(mel).cuda()
for _ in range(2):
    m.forward(n, mel)
torch.cuda.synchronize()
PY
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:k_wide_h -s 22 -c 2 -o gpurun_out/e_prof_wide128 -f python /tmp/wide_once.py 128 > gpurun_out/e_ncu_wide.log 2>&1; echo "ncu wide rc=$?"
