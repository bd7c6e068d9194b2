"""Where does the multi-GPU e2e overhead go? Under torchrun: every rank copies ITS block of a SharedHostBatch
(c4: 256 utt x 16000) to its GPU and back, all ranks at once, and prints the per-rank copy times next to the same
copies from process-private pinned memory, and the cost of the barrier. Evidence for DESIGN section 6.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/hostshard_probe.py"""
import importlib
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
D = importlib.import_module('parallel-wavenet-vocoder_b200.dist')


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    cpus = D.bind_to_gpu_numa(local) if '--no-bind' not in sys.argv else None
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n, t, t_mel, n_mels = 256, 16000, 201, 80
    batch = D.SharedHostBatch(n, t, t_mel, n_mels)
    nz, ml, wv = batch.shard()
    d_nz, d_ml, d_wv = (torch.empty(x.shape, device='cuda') for x in (nz, ml, wv))
    p_nz, p_ml, p_wv = (torch.empty(x.shape).pin_memory() for x in (nz, ml, wv))

    def copies(a, b, c):
        d_nz.copy_(a, non_blocking=True)
        d_ml.copy_(b, non_blocking=True)
        c.copy_(d_wv, non_blocking=True)
        torch.cuda.synchronize()

    out = {}
    for name, bufs in (('shared', (nz, ml, wv)), ('private', (p_nz, p_ml, p_wv))):
        for _ in range(3):
            copies(*bufs)
        ts = []
        for _ in range(10):
            dist.barrier()
            t0 = time.perf_counter()
            copies(*bufs)
            ts.append(time.perf_counter() - t0)
        out[name] = sorted(ts)[len(ts) // 2] * 1e3
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.barrier()
        ts.append(time.perf_counter() - t0)
    out['barrier'] = sorted(ts)[len(ts) // 2] * 1e3
    # the driver query pwv_forward_host made per call until this probe was written, all ranks at once
    ts = []
    for _ in range(10):
        dist.barrier()
        t0 = time.perf_counter()
        torch.cuda.mem_get_info()
        ts.append(time.perf_counter() - t0)
    out['memgetinfo'] = sorted(ts)[len(ts) // 2] * 1e3
    # the whole path on this rank's shard: pwv_forward_host from the shared batch / from private pinned buffers /
    # pwv_forward on device-resident inputs, each step released by a barrier as in bench.py's e2e leg
    hp = importlib.import_module('parallel-wavenet-vocoder_b200.hparam').hparam
    W = importlib.import_module('parallel-wavenet-vocoder_b200.weights')
    V = importlib.import_module('parallel-wavenet-vocoder_b200.vocoder')
    hp.set_hparam_yaml('bench/c4')
    dims = W.model_dims(hp)
    model = V.PwvModel(dims, W.init_weights(hp, seed=0), 'f16x3')
    d_nz.normal_()
    d_ml.uniform_(-1, 1)
    nz.copy_(d_nz)
    ml.copy_(d_ml)
    p_nz.copy_(d_nz)
    p_ml.copy_(d_ml)
    legs = (('fwd_host_shared', lambda: model.forward_host(nz, ml, wv)), ('fwd_host_private', lambda: model.forward_host(p_nz, p_ml, p_wv)),
            ('fwd_device', lambda: (model.forward(d_nz, d_ml), torch.cuda.synchronize())))
    for name, fn in legs:
        for _ in range(2):
            fn()
        ts, tb = [], []
        for _ in range(8):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fn()
            t1 = time.perf_counter()
            dist.barrier()
            ts.append(t1 - t0)
            tb.append(time.perf_counter() - t0)
        out[name] = sorted(ts)[len(ts) // 2] * 1e3
        out[name + '+barrier'] = sorted(tb)[len(tb) // 2] * 1e3
    mb = (nz.numel() + ml.numel() + wv.numel()) * 4 / 1e6
    res = [None] * world
    dist.all_gather_object(res, (rank, out, cpus))
    if rank == 0:
        for r, o, c in res:
            print('rank %d: %.1f MB per step; shared %.3f ms (%.1f GB/s)  private pinned %.3f ms (%.1f GB/s)  barrier %.3f ms  cpus bound %s'
                  % (r, mb, o['shared'], mb / o['shared'], o['private'], mb / o['private'], o['barrier'], c))
            print('        cudaMemGetInfo %.3f ms | forward_host shared %.3f (+barrier %.3f)  private %.3f (+barrier %.3f)  forward on device %.3f (+barrier %.3f) ms'
                  % (o['memgetinfo'], o['fwd_host_shared'], o['fwd_host_shared+barrier'], o['fwd_host_private'], o['fwd_host_private+barrier'],
                     o['fwd_device'], o['fwd_device+barrier']))
    batch.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
