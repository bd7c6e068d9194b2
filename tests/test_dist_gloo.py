"""World-size-2/3 `gloo` runs (CPU) of the utterance scatter / gather plumbing used for N > 1."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_forward(noise, mel):
    # any per-utterance function: batch rows must not interact
    return noise * 2.0 + mel.mean(dim=(1, 2), keepdim=False)[:, None]


def _worker(rank, world, port, n_total, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    D = pkg('dist')
    t, t_mel, n_mels = 160, 3, 80
    g = torch.Generator().manual_seed(7)
    noise = torch.randn((n_total, t), generator=g)
    mel = torch.randn((n_total, t_mel, n_mels), generator=g)
    full = D.sharded_forward(_fake_forward, noise if rank == 0 else None, mel if rank == 0 else None,
                             n_total, t, t_mel, n_mels, torch.device('cpu'))
    if rank == 0:
        np.save(out_path, full.numpy())
        expect = _fake_forward(noise, mel).numpy()
        assert np.array_equal(full.numpy(), expect)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,n_total', [(2, 8), (2, 5), (3, 2)])
def test_scatter_forward_gather(tmp_path, world, n_total):
    out = str(tmp_path / 'full.npy')
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    assert np.load(out).shape == (n_total, 160)


def test_shard_bounds():
    D = pkg('dist')
    assert D.shard_bounds(256, 8) == [(32 * r, 32 * (r + 1)) for r in range(8)]
    assert D.shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert D.shard_bounds(2, 3) == [(0, 1), (1, 2), (2, 2)]
    b = D.shard_bounds(1000, 7)
    assert b[0][0] == 0 and b[-1][1] == 1000 and all(x[1] == y[0] for x, y in zip(b, b[1:]))


def _host_forward(noise, mel, wav):
    # stand-in for pwv_forward_host: host views in, host view out
    wav.copy_(_fake_forward(noise, mel))


def _worker_hostshard(rank, world, port, n_total, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    D = pkg('dist')
    t, t_mel, n_mels = 160, 3, 80
    batch = D.SharedHostBatch(n_total, t, t_mel, n_mels, name='pwv_test_%d' % port, pin=False)
    g = torch.Generator().manual_seed(11)
    noise = torch.randn((n_total, t), generator=g)
    mel = torch.randn((n_total, t_mel, n_mels), generator=g)
    lo, hi = batch.bounds[rank]
    if n_total % 2:                                      # both ways of loading the batch
        batch.fill(noise if rank == 0 else None, mel if rank == 0 else None)
    else:
        batch.fill_shard(noise[lo:hi], mel[lo:hi])
    nz, ml, wv = batch.shard()
    assert nz.shape == (hi - lo, t) and ml.shape == (hi - lo, t_mel, n_mels) and wv.shape == (hi - lo, t)
    assert nz.is_contiguous() and ml.is_contiguous() and wv.is_contiguous()
    assert torch.equal(nz, noise[lo:hi])                 # every rank sees what rank 0 wrote
    for _ in range(2):                                   # reusable step after step
        full = D.hostshard_forward(_host_forward, batch)
    if rank == 0:
        assert torch.equal(full, _fake_forward(noise, mel))
        np.save(out_path, full.numpy())
        assert not os.path.exists(batch.path)            # the name is unlinked once every rank has mapped it
    else:
        assert full is None
    batch.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,n_total', [(2, 8), (2, 5), (3, 2)])
def test_shared_host_batch(tmp_path, world, n_total):
    """The e2e form of bench.py at N > 1: one host batch in shared memory, every rank works on its own block."""
    out = str(tmp_path / 'full.npy')
    mp.spawn(_worker_hostshard, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    assert np.load(out).shape == (n_total, 160)
