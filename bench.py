#!/usr/bin/env python
"""bench.py -- throughput of the IAF-vocoder generation path (audio samples/s).

    python bench.py --gpus N --steps K --warmup W            # this build (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

One "step" = one forward pass (noise + mel -> waveform, 4 IAF flows) over one batch.
Workload: N=1 -> BASELINE.json configs[1] ("c2": batch = 8 utterances x 1 s @ 16 kHz, default hparams,
fp32-parity arithmetic); N>1 -> configs[3] ("c4": 256 utterances x 1 s sharded over the N GPUs, 256/N
per GPU: the config the whole-box metric is quoted on; strong scaling, utterances are independent, no
data-path collective), launched by torchrun with one rank per GPU. `--workload` overrides.

Prints ONE JSON line (rank 0). `value` = whole-job samples/s with inputs resident in HBM, timed
with CUDA events per step (L2 flushed between steps), max over ranks. `sustained` = the same forward
back to back for >= 1.5 s (the power-capped regime a 60 ms timed region never sees). `e2e` = the same
metric through the C-ABI host-buffer call `pwv_forward_host` (pinned host inputs, H2D + kernels + D2H +
sync inside the timed region); at N>1 every rank runs it on its own shard of ONE shared pinned host batch.
`roofline` = the gated-layer kernel's algorithmic bytes / its measured launch time against
MEASURED_PEAKS.json. `cpu_baseline` = the numpy oracle (a port of the reference's TF-CPU path; TF itself
is not installable) timed on this box's host cores. Both arms print the same `config`.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = 'parallel-wavenet-vocoder_b200'
FALLBACK_HBM_GBS = 6650.0      # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
FALLBACK_BF16_TFLOPS = 1590.0

WORKLOADS = {
    # name: (case in hparams.yaml, description)
    'c1': ('bench/c1', 'c1: batch=1 utt x 1 s (16000 samples), 16 kHz, default hparams'),
    'c2': ('bench/c2', 'c2: batch=8 utt x 1 s (16000 samples) per GPU, 16 kHz, 4 IAF flows, default hparams'),
    'c3': ('bench/c3', 'c3: batch=64 utt x 4 s (96000 samples) per GPU, 24 kHz, default hparams'),
    'c4': ('bench/c4', 'c4: batch=256 utt x 1 s (16000 samples) total, sharded over the GPUs (256/N per GPU), 16 kHz, 4 IAF flows, default hparams'),
}


def default_workload(world):
    return 'c2' if world == 1 else 'c4'


def workload_config(name, hp, world):
    """The `config` object both arms print: what is computed, nothing about how."""
    n_total, t = int(hp.generate.batch_size), int(hp.generate.length)
    strong = name == 'c4'
    return {'workload': WORKLOADS[name][1], 'utterances_per_step_whole_job': n_total if strong else n_total * world,
            'length': t, 'sr': int(hp.signal.sr), 'n_gpus': world, 'hparams': 'hparams/default.yaml',
            'l2': 'GPU arm: flushed (256 MB write) before every timed step',
            'timing': 'GPU arm: CUDA events per step, max over ranks; CPU arm: perf_counter per step'}


def pkg(mod):
    return importlib.import_module(PKG + '.' + mod)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=float(p['hbm_gbs']), bf16=float(p['bf16_tflops']),
                    bf16_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), source='measured')
    return dict(hbm=FALLBACK_HBM_GBS, bf16=FALLBACK_BF16_TFLOPS, bf16_sustained=1400.0, source='fallback')


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            if not (t_begin - 0.05 <= ts <= t_end + 0.15):
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples in the timed region'], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': float(max(power))}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on host cores
# ------------------------------------------------------------------------------------------------
_THREADS = {}


def calibrate_threads(fn):
    """All host cores are offered; these small-matrix ops scale badly past a socket's worth of
    threads, so the fastest count among {all, 1/2, 1/4, ... >= 8} on a short probe is used (a CPU
    arm that is slower with more threads would flatter the GPU side)."""
    import torch
    if 'best' in _THREADS:
        return _THREADS['best']
    cores = host_cores()
    cands, c = [], cores
    while c >= 8:
        cands.append(c)
        c //= 2
    cands = cands or [cores]
    best, best_t = cands[0], float('inf')
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    _THREADS['best'] = best
    return best


def time_cpu_oracle(hp, n, t, repeats=1, warm=False):
    """The numpy oracle (float32, BLAS on all host cores) on an (n, t) batch -> samples/s, seconds."""
    from oracle import iaf_oracle as O
    W = pkg('weights')
    d = W.model_dims(hp)
    weights = W.init_weights(hp, seed=0)
    noise, mel = O.synthetic_inputs(n, t, d['hop'], d['n_mels'])
    import torch
    ops = O.TorchOps()                       # torch-CPU kernels (MKL + threaded elementwise)
    threads = calibrate_threads(lambda: O.iaf_vocoder_forward(noise[:1, :d['hop'] * 50], mel[:1, :51], weights, d['dilations'],
                                                             d['hop'], dtype=np.float32, ops=ops))
    torch.set_num_threads(threads)
    time_cpu_oracle.threads = threads
    if warm:
        O.iaf_vocoder_forward(noise[:1, :d['hop'] * 10], mel[:1, :11], weights, d['dilations'], d['hop'], dtype=np.float32, ops=ops)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.iaf_vocoder_forward(noise, mel, weights, d['dilations'], d['hop'], d['use_biases'], False, dtype=np.float32, ops=ops)
        times.append(time.perf_counter() - t0)
    return n * t / float(np.median(times)), times


def run_reference_arm(args, hp, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path. TensorFlow 1.x cannot be
    installed (no wheel for this Python, no network), so this is the oracle port, on all host
    cores. A step = the GPU arm's batch when that is the c1/c2 batch; for the bigger workloads a bounded
    sample of it (8 utterances of the workload's length, at most 16000 samples each) so that the run ends
    within minutes."""
    if rank != 0:
        return
    n_total, t = int(hp.generate.batch_size), int(hp.generate.length)
    n_s, t_s = min(n_total, 8), min(t, 16000)
    for _ in range(max(args.warmup, 0) and 1):      # one warm-up pass is enough for BLAS thread start
        time_cpu_oracle(hp, 1, 1600)
    _, times = time_cpu_oracle(hp, n_s, t_s, repeats=args.steps)
    total = float(sum(times))
    value = n_s * t_s * len(times) / total
    cores = _THREADS.get('best', host_cores())
    whole = (n_s, t_s) == (n_total, t)
    line = {
        'impl': 'reference', 'metric': 'audio_samples_per_sec', 'value': value, 'unit': 'samples/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'strong' if args.workload == 'c4' else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.workload, hp, world),
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': cores, 'cores_available': host_cores(), 'kind': 'port',
                         'sample': (f'{len(times)} x the whole step batch ({n_s} utt x {t_s} samples)' if whole else
                                    f'{len(times)} x a bounded sample of the step batch ({n_s} of {n_total} utt x {t_s} of {t} samples)') +
                                   ', default hparams, fp32; oracle port of the reference TF-CPU forward (oracle/iaf_oracle.py on torch-CPU kernels); TF 1.x not installable'},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    return line


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class StdoutGuard:
    """stdout carries exactly ONE JSON line: anything native libraries write to fd 1 meanwhile (NCCL's
    version banner, for one) is diverted to stderr until the line is printed."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


def src_sha16(rel):
    import hashlib
    with open(os.path.join(ROOT, rel), 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()[:16]


def measured_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None
    when that capture was taken from a different revision of the kernel source (a stale figure is worse than none)."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None, 'no ncu capture committed'
    with open(path) as fh:
        tj = json.load(fh)
    ent = tj.get(kernel_key)
    if not ent:
        return None, f'no ncu capture of {kernel_key}'
    if ent.get('src_sha16') != src_sha16(ent.get('src', PKG + '/csrc/pwv_tc2.cuh')):
        return None, f'ncu capture of {kernel_key} predates the current kernel source'
    return ent.get('dram_bytes_per_launch'), ent.get('capture')


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS), help='default: c2 on 1 GPU, c4 on N > 1')
    ap.add_argument('--precision', default=None, help="override engine.precision: fp32 | f16x3 | bf16")
    ap.add_argument('--debug', action='append', default=[], metavar='KEY=INT', help='pwv_debug_set switch (A/B runs), repeatable')
    ap.add_argument('--e2e-mode', default='hostshard', choices=['hostshard', 'nccl'],
                    help='N > 1: shared pinned host batch, every rank copies its own shard (default) | rank-0 H2D + NCCL scatter/gather')
    ap.add_argument('--sustain-s', type=float, default=1.5, help='seconds of back-to-back forwards for the sustained leg (0: off)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    debug = {}
    for kv in args.debug:
        k, _, v = kv.partition('=')
        debug[k] = int(v)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.workload is None:
        args.workload = default_workload(world)
    hp = pkg('hparam').hparam
    hp.set_hparam_yaml(WORKLOADS[args.workload][0])

    if args.impl == 'reference':
        line = run_reference_arm(args, hp, rank, world)
        if line is not None:
            guard.emit(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa_cpus = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        numa_cpus = pkg('dist').bind_to_gpu_numa(local_rank)   # the rank's host buffers and copy threads next to its GPU's PCIe root
        dist.init_process_group('nccl', device_id=dev)      # (NCCL's log lines go wherever NCCL_DEBUG sends them; fd 1 is guarded)
    if world != args.gpus and rank == 0:
        print(f'bench.py: note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}', file=sys.stderr)

    V, W, IO = pkg('vocoder'), pkg('weights'), pkg('io')
    precision = V.resolve_precision(W.model_dims(hp), args.precision or hp.engine.precision)
    dims = W.model_dims(hp)
    n_total, t = int(hp.generate.batch_size), int(hp.generate.length)
    strong = args.workload == 'c4'
    if strong and n_total % world:
        raise SystemExit(f'bench.py: c4 shards {n_total} utterances evenly; --gpus {world} does not divide it')
    n = n_total // world if strong else n_total             # c4 is strong-scaled, the others weak
    n_job = n * world
    weights = W.init_weights(hp, seed=0)
    model = V.PwvModel(dims, weights, precision, debug=debug)
    noise_h, mel_h = IO.synthetic_batch(n, t, dims['hop'], dims['n_mels'], mel_seed=1234 + rank, noise_seed=1235 + rank)
    noise = torch.from_numpy(noise_h).to(dev)
    mel = torch.from_numpy(mel_h).to(dev)
    out = torch.empty((n, t), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(ev0, ev1):
        flush.fill_(1)                          # evict L2 (not timed)
        ev0.record()
        model.forward(noise, mel, out=out)
        ev1.record()

    for _ in range(args.warmup):
        one_step(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = time.time()
    for ev0, ev1 in events:
        one_step(ev0, ev1)
    barrier()
    t_end = time.time()
    step_ms = [a.elapsed_time(b) for a, b in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    samples_per_step = n_job * t
    value = samples_per_step * args.steps / (total_ms * 1e-3)
    launches_per_forward = model.last_launch_count()
    launches = launches_per_forward * args.steps

    # ---- sustained leg: the same forward back to back for >= sustain_s seconds (one event pair around the loop; the
    #      working set of a forward, >= 180 MB of activations streamed ~60 times, turns L2 over by itself)
    sustained = None
    t_sus0 = t_sus1 = None
    if args.sustain_s > 0:
        iters = max(args.steps, int(np.ceil(args.sustain_s * 1e3 / (total_ms / args.steps))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_sus0 = time.time()
        e0.record()
        for _ in range(iters):
            model.forward(noise, mel, out=out)
        e1.record()
        barrier()
        t_sus1 = time.time()
        sus_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
        sus_ms = float(sus_ms.item())
        sustained = {'value': samples_per_step * iters / (sus_ms * 1e-3), 'unit': 'samples/s', 'iterations': iters, 'seconds': sus_ms * 1e-3,
                     'ms_per_step': sus_ms / iters, 'note': 'back-to-back forwards, no L2 flush in between, one CUDA-event pair around the loop'}
    clocks = None
    if rank == 0:
        clocks = sampler.stop(t_begin, t_end)           # samples of the K timed steps
        if sustained is not None:
            sustained['clocks'] = window_clocks(sampler.lines, t_sus0, t_sus1)
            sustained['gpu_launches'] = launches_per_forward * sustained['iterations']     # (`gpu_launches` of the line: the K timed steps only)

    # ---- roofline of the dominant kernel (gated dilated layer), separate profiled steps
    pk = peaks()

    def profiled(mode):
        model.set_profiling(mode)
        lms, ln, fms = [], 0, []
        for _ in range(3):
            flush.fill_(1)
            model.forward(noise, mel, out=out)
            lm, ln, fm = model.profile_read()
            lms.append(lm); fms.append(fm)
        model.set_profiling(False)
        return lms, ln, fms
    # mode 2: one event pair around each flow's chain of gated-layer launches, launched as in the timed steps
    # (programmatic dependent launch on) -> the kernel's average launch duration inside the step;
    # mode 1: every launch bracketed (serialised) -> the isolated launch duration, what ncu's list shows
    layer_ms, layer_n, fwd_ms = profiled(2)
    iso_ms, iso_n, _ = profiled(1)
    # One launch of the gated-layer kernel runs one layer of both bodies. Algorithmic bytes per launch:
    # each body reads its 64-channel input once and writes its output once, sizeof(act) bytes per element:
    # 4 in fp32 / f16x3 (two fp16 planes hi + lo), 2 in bf16 on the plane path (one bf16 plane).
    planes_path = precision != 'fp32' and debug.get('path', 1) == 1
    act_bytes = 2 if (precision == 'bf16' and planes_path) else 4
    n_gated = sum(len(d) for d in dims['dilations'])                     # gated layers per forward (x 2 bodies each)
    bytes_per_layer = 2 * n * t * (2 * dims['R'] * act_bytes)
    layer_s = float(np.median(layer_ms)) * 1e-3                          # device time of all gated-layer launches
    achieved = n_gated * bytes_per_layer / layer_s / 1e9
    mac_per_layer = 2 * n * t * (2 * dims['R'] * 2 * dims['D'] + dims['D'] * dims['R'])
    general = dims['k'] != 2 or dims['R'] != dims['D'] or dims['S'] != 2 * dims['R'] or dims['R'] not in (64, 128, 256)
    if precision == 'fp32' and general:
        kernel, kkey = 'k_gen_gemm chain: un-fused general-shape gated layer, fp32 FFMA (csrc/pwv_gen.cuh)', 'k_gen:fp32'
    elif precision == 'fp32':
        kernel, kkey = 'k_layer_simt: gated dilated layer, fp32 FFMA, both bodies per launch', 'k_layer_simt:fp32'
    elif planes_path:
        kernel = ('k_layer_h: gated dilated layer on tcgen05, activations as 16-bit planes in HBM, both bodies per launch; '
                  'one launch per layer chained by programmatic dependent launch')
        kkey = 'k_layer_h:' + precision
    else:
        kernel, kkey = 'k_layer_tc / k_flow_tc (round-1 fp32-row kernels, debug path 0)', 'k_layer_tc:' + precision
    traffic, traffic_src = measured_traffic(kkey)
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': achieved / pk['hbm'],
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': pk['source'], 'kernel': kernel,
                'bytes_per_launch': n_gated * bytes_per_layer / max(layer_n, 1), 'avg_launch_us': layer_s * 1e6 / max(layer_n, 1),
                'isolated_launch_us': float(np.median(iso_ms)) * 1e3 / max(iso_n, 1),
                'frac_isolated': n_gated * bytes_per_layer / (float(np.median(iso_ms)) * 1e-3) / 1e9 / pk['hbm'],
                'launches_per_step': layer_n, 'gated_layers_per_step': n_gated, 'us_per_layer': layer_s * 1e6 / n_gated,
                'share_of_step': float(np.median(layer_ms) / np.median(fwd_ms)),
                'tflops_fp32_equiv': 2 * n_gated * mac_per_layer / layer_s / 1e12,
                'act_bytes_per_element': act_bytes,
                'note': f'algorithmic bytes = {2 * dims["R"] * act_bytes} B per sample per body-layer (SURVEY 8d: sizeof(act) x 2 x R); avg_launch_us = CUDA events around '
                        'each flow\'s chain of gated-layer launches as launched in the timed steps / launches; isolated_launch_us = every launch bracketed '
                        '(serialised, what ncu lists)'}

    # ---- e2e: host buffers in, host buffer out, copies inside the timed region, through the C-ABI call
    #      pwv_forward_host (H2D + kernels + D2H + sync inside the call).
    #      N = 1: pinned host buffers of the process.
    #      N > 1 (hostshard): ONE pinned host batch in shared memory holds the whole job; every rank runs
    #             pwv_forward_host on its own block of utterances (its own PCIe link), rank 0 owns the complete
    #             output after the barrier (parallel-wavenet-vocoder_b200/dist.py).
    #      N > 1 (nccl): rank-0 H2D -> NCCL scatter -> pwv_forward per rank -> NCCL gather -> rank-0 D2H (round-1 form).
    e2e = None
    if not args.no_e2e:
        steps_e2e = max(3, min(args.steps, 10))
        t_mel = 1 + t // dims['hop']
        if world == 1:
            pin_n = torch.from_numpy(noise_h).pin_memory()
            pin_m = torch.from_numpy(mel_h).pin_memory()
            pin_o = torch.empty((n, t), dtype=torch.float32).pin_memory()
            for _ in range(2):
                model.forward_host(pin_n, pin_m, pin_o)
            e2e_s = 0.0
            barrier()
            for _ in range(steps_e2e):
                flush.fill_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                model.forward_host(pin_n, pin_m, pin_o)      # H2D + kernels + D2H + sync inside
                e2e_s += time.perf_counter() - t0
            api = 'pwv_forward_host (pinned host buffers)'
        elif args.e2e_mode == 'hostshard':
            D = pkg('dist')
            batch = D.SharedHostBatch(n_job, t, t_mel, dims['n_mels'])
            noise_all, mel_all = IO.synthetic_batch(n_job, t, dims['hop'], dims['n_mels'])    # every rank: the same job batch ...
            lo, hi = batch.bounds[rank]
            batch.fill_shard(noise_all[lo:hi], mel_all[lo:hi])                                # ... of which it loads its own block
            del noise_all, mel_all
            for _ in range(2):
                D.hostshard_forward(model.forward_host, batch)
            e2e_s = 0.0
            for _ in range(steps_e2e):
                flush.fill_(1)
                barrier()
                t0 = time.perf_counter()
                D.hostshard_forward(model.forward_host, batch)    # per rank: H2D + kernels + D2H + sync; then the barrier
                e2e_s += time.perf_counter() - t0
            api = ('one pinned host batch in shared memory (pinned: %s); every rank: pwv_forward_host on its own shard '
                   '(H2D + kernels + D2H + sync), barrier; rank 0 owns the whole output' % batch.pinned)
            batch.close()
        else:
            D = pkg('dist')
            if rank == 0:
                noise_all, mel_all = IO.synthetic_batch(n_job, t, dims['hop'], dims['n_mels'])
                pin_n = torch.from_numpy(noise_all).pin_memory()
                pin_m = torch.from_numpy(mel_all).pin_memory()
                pin_o = torch.empty((n_job, t), dtype=torch.float32).pin_memory()

            def one_e2e():
                nz = pin_n.to(dev, non_blocking=True) if rank == 0 else None
                ml = pin_m.to(dev, non_blocking=True) if rank == 0 else None
                full = D.sharded_forward(lambda a, b: model.forward(a, b), nz, ml, n_job, t, t_mel, dims['n_mels'], dev)
                if rank == 0:
                    pin_o.copy_(full, non_blocking=True)
                torch.cuda.synchronize()
            for _ in range(2):
                one_e2e()
            e2e_s = 0.0
            for _ in range(steps_e2e):
                flush.fill_(1)
                barrier()
                t0 = time.perf_counter()
                one_e2e()
                e2e_s += time.perf_counter() - t0
            api = 'rank-0 pinned host batch -> H2D -> NCCL scatter -> pwv_forward per rank -> NCCL gather -> D2H'
        h2d, d2h = int(n_job * t * 4 + n_job * t_mel * dims['n_mels'] * 4), int(n_job * t * 4)
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {'value': samples_per_step * steps_e2e / float(tt.item()), 'unit': 'samples/s',
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': steps_e2e, 'api': api}

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sn, st = min(n, 8), min(t, 16000)
        cps, times = time_cpu_oracle(hp, sn, st, repeats=1, warm=True)
        cpu = {'value': cps, 'unit': 'samples/s', 'cores': _THREADS.get('best', host_cores()), 'cores_available': host_cores(), 'kind': 'port',
               'sample': f'1 pass over {sn} utt x {st} samples (the {args.workload} batch), oracle port on torch-CPU kernels, fp32, {times[0]:.1f} s'}

    if rank == 0:
        line = {
            'metric': 'audio_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'f16x3': 'f32 (fp16 hi/lo 3-term tensor-core split, fp32 accumulate; activations stored as fp16 hi + lo planes)',
                      'bf16': 'bf16 (fp32 accumulate)'}[precision],
            'data': 'synthetic',
            'config': workload_config(args.workload, hp, world),
            'engine': {'precision': precision, 'per_gpu_batch': n, 'activation_layout': 'planes' if planes_path else 'fp32 rows', 'debug': debug,
                       'cpus_bound_to_gpu_socket': numa_cpus},
            'roofline': roofline, 'sustained': sustained, 'cpu_baseline': cpu, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def window_clocks(lines, t_begin, t_end):
    sm, power, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for ts, line in lines:
        parts = [p.strip() for p in line.split(',')]
        if len(parts) < 7 or not (t_begin <= ts <= t_end + 0.1):
            continue
        try:
            sm.append(float(parts[0])); power.append(float(parts[2]))
        except ValueError:
            continue
        for name, flag in zip(names, parts[3:7]):
            if flag.lower().startswith('active'):
                reasons.add(name)
    if not sm:
        return {'sm_mhz': None, 'samples': 0, 'reasons': []}
    return {'sm_mhz': float(np.median(sm)), 'sm_mhz_min': float(min(sm)), 'samples': len(sm), 'power_w_max': float(max(power)), 'reasons': sorted(reasons)}


if __name__ == '__main__':
    main()
