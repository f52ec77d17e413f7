"""One-process-per-GPU plumbing for the path (torch.distributed; NCCL on the
GPU box, gloo in the CPU tests).  The path shards by sample and needs no
data-path collective; what crosses ranks is the tiny parameter-gradient /
logged-loss vector that DDP's bucket carries in training."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [start, stop) slice of n samples for `rank`."""
    return n * rank // world, n * (rank + 1) // world


def average_across_ranks(tensors):
    """All-reduce (mean) a list of small tensors in one flat bucket; returns new
    tensors.  Equal per-rank batches + per-rank mean losses => the result equals
    the single-process global-batch value (train.py:197-199 are means)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [t.clone() for t in tensors]
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    out, o = [], 0
    for t in tensors:
        out.append(flat[o:o + t.numel()].view_as(t).to(t.dtype))
        o += t.numel()
    return out


def wrap_ddp(model, device_id):
    """DistributedDataParallel over NCCL with one bucket (the whole model is
    13 MB of float32 gradients)."""
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[device_id], bucket_cap_mb=32,
                                                     gradient_as_bucket_view=True)
