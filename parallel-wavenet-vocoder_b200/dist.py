"""Utterance sharding across the GPUs of one box.

The path has no exchange step: every utterance is independent, weights (19.4 MB) are
replicated. One process per GPU (torchrun). Two ways of moving the inputs / outputs (~12 bytes per
audio sample in total) between the job's host batch and the ranks:

* `SharedHostBatch` (default of bench.py's e2e leg): the whole job's `(noise, mel, wav)` lives in ONE
  host buffer in POSIX shared memory that every rank maps and pins (cudaHostRegister). Each rank copies
  its own contiguous block of utterances host->device over its OWN PCIe link, runs the forward and
  copies its block of `wav` back into the shared buffer -- 8 links in parallel, nothing funnelled through
  rank 0 (round 1's scatter/gather form lost 37 % at 8 GPUs to exactly that funnel: one H2D of the whole
  batch, 21 point-to-point ops, one D2H). The only collective left is the barrier that tells rank 0 the
  output is complete.
* `scatter_inputs` / `gather_outputs`: rank `src` holds the batch on ITS device and NCCL moves the
  shards over NVLink (grouped `send/recv`, one `batch_isend_irecv` per direction). Kept for callers whose
  batch is already device-resident on one rank; works with `gloo` too (CPU tests).

Uneven splits are supported (the first `N % world` ranks take one more utterance).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """[(start, stop)] per rank: contiguous blocks, sizes differ by at most one."""
    base, extra = divmod(int(n), int(world))
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def _exchange(ops):
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()


def scatter_inputs(noise, mel, n_total, t, t_mel, n_mels, device, src=0, group=None):
    """`noise`/`mel` are the full tensors on rank `src` (ignored elsewhere). Returns this rank's
    shard `(noise_shard, mel_shard)` on `device`."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(n_total, world)
    lo, hi = bounds[rank]
    if rank == src:
        noise = noise.to(device).contiguous()
        mel = mel.to(device).contiguous()
        ops = []
        for r, (a, b) in enumerate(bounds):
            if r == src or b == a:
                continue
            ops.append(dist.P2POp(dist.isend, noise[a:b], r, group))
            ops.append(dist.P2POp(dist.isend, mel[a:b], r, group))
        _exchange(ops)
        return noise[lo:hi], mel[lo:hi]
    noise_s = torch.empty((hi - lo, t), dtype=torch.float32, device=device)
    mel_s = torch.empty((hi - lo, t_mel, n_mels), dtype=torch.float32, device=device)
    if hi > lo:
        _exchange([dist.P2POp(dist.irecv, noise_s, src, group), dist.P2POp(dist.irecv, mel_s, src, group)])
    return noise_s, mel_s


def gather_outputs(wav_shard, n_total, dst=0, group=None):
    """Inverse of scatter: rank `dst` returns the full `(n_total, T)` tensor, the others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(n_total, world)
    t = wav_shard.shape[1]
    if rank == dst:
        full = torch.empty((n_total, t), dtype=wav_shard.dtype, device=wav_shard.device)
        lo, hi = bounds[rank]
        full[lo:hi] = wav_shard
        ops = [dist.P2POp(dist.irecv, full[a:b], r, group) for r, (a, b) in enumerate(bounds) if r != dst and b > a]
        _exchange(ops)
        return full
    if wav_shard.shape[0]:
        _exchange([dist.P2POp(dist.isend, wav_shard.contiguous(), dst, group)])
    return None


def sharded_forward(forward, noise, mel, n_total, t, t_mel, n_mels, device, src=0, group=None):
    """scatter -> `forward(noise_shard, mel_shard) -> wav_shard` on every rank -> gather on `src`."""
    noise_s, mel_s = scatter_inputs(noise, mel, n_total, t, t_mel, n_mels, device, src, group)
    if noise_s.shape[0]:
        wav_s = forward(noise_s, mel_s)
    else:
        wav_s = torch.empty((0, t), dtype=torch.float32, device=device)
    return gather_outputs(wav_s, n_total, src, group)


class SharedHostBatch:
    """The job's host batch in POSIX shared memory, mapped by every rank of the node.

    Layout (float32): noise [N][T] | mel [N][t_mel][n_mels] | wav [N][T]. Rank `src` creates and fills it,
    the others map it after a barrier; with `pin=True` every rank page-locks its mapping so that the
    copies to and from its GPU are real asynchronous DMA transfers. `shard()` returns this rank's views.
    """

    def __init__(self, n_total, t, t_mel, n_mels, name=None, src=0, group=None, pin=True, directory='/dev/shm'):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_total, self.t, self.t_mel, self.n_mels = int(n_total), int(t), int(t_mel), int(n_mels)
        self.src, self.group = src, group
        sizes = [self.n_total * self.t, self.n_total * self.t_mel * self.n_mels, self.n_total * self.t]
        self._offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(self._offsets[-1])
        if name is None:
            name = 'pwv_batch_%s_%s' % (os.environ.get('MASTER_PORT', '0'), os.environ.get('TORCHELASTIC_RUN_ID', 'x'))
        self.path = os.path.join(directory, name)
        if self.rank == src:
            with open(self.path, 'wb') as fh:
                fh.truncate(total * 4)
        dist.barrier(group)
        self.buf = torch.from_file(self.path, shared=True, size=total, dtype=torch.float32)
        self.bounds = shard_bounds(self.n_total, self.world)
        # First touch decides which NUMA node a page of the (so far unbacked) file lands on: every rank touches ITS block
        # of each region before anybody pins the buffer, so a rank's DMA stays on the socket the rank runs on (measured at
        # 8 ranks with every page on rank 0's node: 2.9 ms per step of extra copy time on the far-socket GPUs).
        for view in self.shard():
            view.zero_()
        dist.barrier(group)
        self.pinned = False
        if pin and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), total * 4, 0)
            self.pinned = int(rc) == 0
        dist.barrier(group)
        if self.rank == src:                       # every rank has the file open: unlink the name now
            os.unlink(self.path)

    def _view(self, i, shape):
        a, b = int(self._offsets[i]), int(self._offsets[i + 1])
        return self.buf[a:b].view(*shape)

    @property
    def noise(self):
        return self._view(0, (self.n_total, self.t))

    @property
    def mel(self):
        return self._view(1, (self.n_total, self.t_mel, self.n_mels))

    @property
    def wav(self):
        return self._view(2, (self.n_total, self.t))

    def fill(self, noise, mel):
        """Rank `src` writes the job's inputs; everyone returns after they are visible to all ranks."""
        if self.rank == self.src:
            self.noise.copy_(torch.as_tensor(noise))
            self.mel.copy_(torch.as_tensor(mel))
        dist.barrier(self.group)

    def fill_shard(self, noise_shard, mel_shard):
        """Every rank writes its own block (a data loader per rank); returns after all blocks are visible."""
        nz, ml, _ = self.shard()
        if nz.shape[0]:
            nz.copy_(torch.as_tensor(noise_shard))
            ml.copy_(torch.as_tensor(mel_shard))
        dist.barrier(self.group)

    def shard(self):
        """-> (noise, mel, wav) host views of this rank's utterances (contiguous, pinned if `pinned`)."""
        lo, hi = self.bounds[self.rank]
        return self.noise[lo:hi], self.mel[lo:hi], self.wav[lo:hi]

    def close(self):
        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            self.pinned = False


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs NVML reports as local to GPU `device_index` (its socket), so that host buffers it
    touches and the threads that drive its copies sit next to its PCIe root. Best effort: returns the CPU count, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def hostshard_forward(forward_host, batch):
    """Every rank runs `forward_host(noise_view, mel_view, wav_view)` (host buffers in, host buffer out -- the
    C-ABI call pwv_forward_host) on its own shard of the shared host batch; after the barrier rank `src` holds
    the whole job's output in `batch.wav`."""
    noise, mel, wav = batch.shard()
    if noise.shape[0]:
        forward_host(noise, mel, wav)
    dist.barrier(batch.group)
    return batch.wav if batch.rank == batch.src else None
