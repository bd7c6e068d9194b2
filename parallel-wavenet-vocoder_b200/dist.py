"""Utterance sharding across the GPUs of one box.

The path has no exchange step: every utterance is independent, weights (19.4 MB) are
replicated. One process per GPU (torchrun); NCCL moves only inputs and outputs:
rank `src` scatters `(noise[N,T], mel[N,t_mel,n_mels])` in contiguous blocks of utterances and
gathers `wav[N,T]` (~12 bytes per audio sample in total). Uneven splits are supported (the first
`N % world` ranks take one more utterance). Works with any backend (`nccl` on GPUs, `gloo` in the
CPU tests) because it only uses point-to-point `send/recv` via `batch_isend_irecv`.
"""
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
