"""Stream sharding across GPUs: streams are independent, so rank r of `world` owns one
contiguous block of stream indices and runs the same kernels on it; nothing crosses GPUs on the data
path.  Only the tiny per-stream result vectors are gathered (to the host, or with one all_gather
when a torch.distributed group exists)."""


def shard_range(nstreams, world, rank):
    """Contiguous block [start, start+count) of rank `rank`; blocks differ by at most one stream."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(nstreams), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def gather_results(local, nstreams, group=None):
    """all_gather ragged per-stream result tensors (first dim = streams of this rank) into stream
    order on every rank.  `local` is a torch tensor; without an initialised process group it is
    returned unchanged (single GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_range(nstreams, world, r)[1] for r in range(world)]
    width = max(counts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
