"""Multi-GPU reading: contiguous frame ranges per rank, optional all-gather.

Frames are independent, so a stream shards with no collective on the data
path: rank ``r`` of ``W`` decodes frames
``[r*N/W, (r+1)*N/W)`` on its own GPU, reading only its own byte range of the
file.  Only when the caller wants the whole decoded stream on every device is
a collective used: one NCCL all-gather of the decoded shards over NVLink
(``gather=True``).  Gathering *decoded* float32 moves 16x the packed bytes for
2-bit data; prefer leaving the shards where they are.

One process per GPU (``torchrun``); ``torch.distributed`` is plumbing only.
"""
import os

import torch

__all__ = ['shard_bounds', 'shard_samples', 'read_sharded']


def shard_bounds(nitem, rank, world):
    """[start, stop) of ``nitem`` items for ``rank``: contiguous, covering,
    sizes differing by at most one."""
    if not 0 <= rank < world:
        raise ValueError('rank {} outside world of size {}'.format(rank,
                                                                   world))
    return rank * nitem // world, (rank + 1) * nitem // world


def shard_samples(fh, rank, world):
    """Sample range [start, stop) of stream reader ``fh`` owned by ``rank``:
    whole frames, except that the last rank takes the tail (e.g. GUPPI's
    final overlap)."""
    f0, f1 = shard_bounds(fh._nframe, rank, world)
    spf = fh.samples_per_frame
    start = f0 * spf
    stop = fh.shape[0] if rank == world - 1 else f1 * spf
    return start, stop


def _dist_env(rank, world):
    if rank is None or world is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return (int(os.environ.get('RANK', 0)),
                int(os.environ.get('WORLD_SIZE', 1)))
    return rank, world


def read_sharded(fh, rank=None, world=None, gather=False, group=None):
    """Decode this rank's contiguous share of ``fh``.

    Returns ``(data, (start, stop))``: the decoded samples of the shard (on
    the reader's device when it was opened with ``device=``) and the sample
    range they cover.  With ``gather=True`` every rank instead receives the
    whole stream: shards are padded to a common length, all-gathered
    (NCCL over NVLink for CUDA tensors, gloo for host tensors) and trimmed.
    """
    rank, world = _dist_env(rank, world)
    start, stop = shard_samples(fh, rank, world)
    fh.seek(start)
    data = fh.read(stop - start)
    if not gather or world == 1:
        return data, (start, stop)
    import torch.distributed as dist
    tensor = data if isinstance(data, torch.Tensor) else torch.from_numpy(data)
    bounds = [shard_samples(fh, r, world) for r in range(world)]
    longest = max(b - a for a, b in bounds)
    padded = torch.zeros((longest,) + tuple(tensor.shape[1:]),
                         dtype=tensor.dtype, device=tensor.device)
    padded[:tensor.shape[0]] = tensor
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded, group=group)
    whole = torch.cat([p[:b - a] for p, (a, b) in zip(pieces, bounds)])
    if not isinstance(data, torch.Tensor):
        whole = whole.numpy()
    return whole, (0, fh.shape[0])
