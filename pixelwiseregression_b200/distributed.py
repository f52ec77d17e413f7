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


def proportional_shards(rates, total, quantum=64, lo=None, hi=None):
    """Split `total` samples over the ranks in proportion to `rates` (samples per second each rank sustained on
    its own host link), in multiples of `quantum`, every share clamped to [lo, hi].  Pure and deterministic:
    every rank computes the same list from the all-gathered rates.

    Why: the e2e feed is PCIe-bound, a lock-step job runs at the pace of its slowest link, and on an 8-GPU box
    the links are not equal (GPUs behind a shared upstream port halve each other; profiles/r2_pcie_probe_n8.json:
    20-42 GB/s per GPU).  Equal shards then waste the fast links; shares proportional to the measured rates make
    every rank's transfer take the same time.  The loss stays the global-batch mean through `n_mean` =
    global B*J and SUM-reduced gradients (ops.fused_decoder_loss)."""
    n = len(rates)
    if n == 0 or total % quantum != 0:
        raise ValueError("total must be a multiple of quantum and rates non-empty")
    units = total // quantum
    lo_u = 0 if lo is None else -(-int(lo) // quantum)
    hi_u = units if hi is None else int(hi) // quantum
    if lo_u * n > units or hi_u * n < units or lo_u > hi_u:
        raise ValueError("bounds [%r, %r] cannot hold %d samples on %d ranks" % (lo, hi, total, n))
    r = [max(float(x), 0.0) for x in rates]
    if not all(x == x and x != float("inf") for x in r) or sum(r) <= 0.0:
        r = [1.0] * n
    # water-filling over the clamped shares: ranks pinned to a bound leave the rest to the others
    share = [0] * n
    free = list(range(n))
    left = units
    ideal = [0.0] * n
    while free:
        s = sum(r[i] for i in free)
        for i in free:
            ideal[i] = left * (r[i] / s) if s > 0 else left / len(free)
        pinned = [i for i in free if ideal[i] < lo_u or ideal[i] > hi_u]
        if not pinned:
            break
        for i in pinned:
            share[i] = lo_u if ideal[i] < lo_u else hi_u
            left -= share[i]
            free.remove(i)
    for i in free:
        share[i] = min(max(int(ideal[i]), lo_u), hi_u)
    # hand out / take back whole quanta by largest fractional part (ties: lowest rank), inside the bounds
    rest = units - sum(share)
    order = sorted(free, key=lambda i: (-(ideal[i] - int(ideal[i])), i))
    k = 0
    while rest != 0 and order:
        i = order[k % len(order)]
        if rest > 0 and share[i] < hi_u:
            share[i] += 1
            rest -= 1
        elif rest < 0 and share[i] > lo_u:
            share[i] -= 1
            rest += 1
        k += 1
        if k > 4 * units + 4 * n:
            break
    if rest != 0:                  # the free ranks could not absorb it: spread over everybody
        for i in sorted(range(n), key=lambda i: (-r[i], i)) * (abs(rest) + 1):
            if rest == 0:
                break
            if rest > 0 and share[i] < hi_u:
                share[i] += 1
                rest -= 1
            elif rest < 0 and share[i] > lo_u:
                share[i] -= 1
                rest += 1
    assert sum(share) == units, (share, units)
    return [s * quantum for s in share]


def gather_rates(rate, device=None):
    """All-gather one float per rank (this rank's measured samples/s) -> list over ranks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(rate)]
    t = torch.tensor([float(rate)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]
