"""CPU, world_size 2 over gloo: batch sharding of the path.  Each rank runs the
decoder + loss + backward on its half of the batch (oracle arithmetic on CPU, the
checker allowed in tests) and the rank-averaged dL/dw and loss terms must equal
the single-process result on the whole batch — the property DDP relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import decoder_oracle as do
from pixelwiseregression_b200 import distributed as pd
from pixelwiseregression_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _case(B, J, seed):
    d = synth.make_decoder_inputs(B, J, seed)
    rng = np.random.default_rng(seed + 1)
    tg = (rng.uniform(0, 0.05, (B, J, 64, 64)).astype(np.float32),
          (rng.standard_normal((B, J, 64, 64)) * d["mask"]).astype(np.float32),
          rng.uniform(-0.5, 0.5, (B, J, 3)).astype(np.float32))
    return d, tg


def _loss_and_gw(d, tg, lo, hi, alpha):
    t = lambda a: torch.from_numpy(a[lo:hi]).double()
    w = torch.from_numpy(d["w"]).double().requires_grad_(True)
    p, Dm, uvd = do.decoder_forward(t(d["z"]), w, t(d["D"]), t(d["label"]), t(d["mask"]))
    terms = do.stage_losses(p, Dm, uvd, t(tg[0]), t(tg[1]), t(tg[2]))
    do.combine_losses(terms, alpha).backward()
    return torch.stack([x.detach() for x in terms]), w.grad.detach()


def _worker(rank, world, port, B, J, alpha, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, tg = _case(B, J, 3)
        lo, hi = pd.shard_range(B, rank, world)
        terms, gw = _loss_and_gw(d, tg, lo, hi, alpha)
        terms, gw = pd.average_across_ranks([terms, gw])
        if rank == 0:
            torch.save({"terms": terms, "gw": gw}, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 128, 4096):
        for world in (1, 2, 3, 8):
            parts = [pd.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_average_across_ranks_without_process_group_is_identity():
    a, b = torch.arange(3.0), torch.ones(2, 2)
    x, y = pd.average_across_ranks([a, b])
    assert torch.equal(x, a) and torch.equal(y, b)


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_two_rank_average_equals_global_batch(tmp_path, alpha):
    B, J, world = 4, 3, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, J, alpha, out), nprocs=world, join=True)
    got = torch.load(out)
    d, tg = _case(B, J, 3)
    terms, gw = _loss_and_gw(d, tg, 0, B, alpha)
    assert torch.allclose(got["terms"].double(), terms, rtol=1e-6, atol=1e-9)
    assert torch.allclose(got["gw"].double(), gw, rtol=1e-5, atol=1e-9)


# ---- bandwidth-proportional shards of the global batch (the e2e feed at N > 1) ----

def test_proportional_shards_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for _ in range(50):
            rates = rng.uniform(5.0, 60.0, world)
            total, q = world * 4096, 64
            shards = pd.proportional_shards(list(rates), total, q, 2048, 6144)
            assert sum(shards) == total and all(s % q == 0 for s in shards)
            assert all(2048 <= s <= 6144 for s in shards)
            ideal = total * rates / rates.sum()
            free = [i for i in range(world) if 2048 < shards[i] < 6144]
            if len(free) == world:                                   # nobody clamped: within one quantum of the ideal share
                assert all(abs(shards[i] - ideal[i]) <= q for i in free)
            order = np.argsort(rates)
            assert all(shards[a] <= shards[b] + q for a, b in zip(order, order[1:]))     # monotone up to rounding


def test_proportional_shards_edge_cases():
    assert pd.proportional_shards([1.0, 1.0], 8192) == [4096, 4096]
    assert pd.proportional_shards([1.0, 100.0], 8192, 64, 2048, 6144) == [2048, 6144]           # clamped both ways
    assert pd.proportional_shards([0.0, 0.0], 8192) == [4096, 4096]                             # no information: equal
    assert pd.proportional_shards([float("nan"), 3.0], 8192) == [4096, 4096]
    assert pd.proportional_shards([7.0], 4096) == [4096]
    # the eight links of profiles/r2_pcie_probe_n8.json (window kernel, GB/s): the slow one gets the smallest share
    shards = pd.proportional_shards([29.4, 29.5, 29.5, 25.8, 31.8, 33.5, 40.2, 30.5], 8 * 4096, 64, 2048, 6144)
    assert sum(shards) == 8 * 4096 and min(shards) == shards[3] and max(shards) == shards[6]
    with pytest.raises(ValueError):
        pd.proportional_shards([1.0, 1.0], 8192 + 1)
    with pytest.raises(ValueError):
        pd.proportional_shards([1.0, 1.0], 8192, 64, 5000, 6144)


def _shard_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # uneven shards + n_mean: per-rank partial sums over the GLOBAL B*J, SUM-reduced == the global-batch mean
        B, J, alpha = 6, 3, 0.5
        rates = pd.gather_rates([10.0, 20.0][rank])
        shards = pd.proportional_shards(rates, B, 1, 1, B)
        lo = sum(shards[:rank])
        d, tg = _case(B, J, 5)
        terms, gw = _loss_and_gw(d, tg, lo, lo + shards[rank], alpha)
        scale = shards[rank] / B                         # local mean -> this rank's part of the global mean
        flat = torch.cat([terms * scale, gw.reshape(-1) * scale])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if rank == 0:
            torch.save({"rates": rates, "shards": shards, "flat": flat}, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_with_uneven_shards_sum_to_the_global_batch(tmp_path):
    out = str(tmp_path / "s.pt")
    mp.spawn(_shard_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["rates"] == [10.0, 20.0] and got["shards"] == [2, 4]
    d, tg = _case(6, 3, 5)
    terms, gw = _loss_and_gw(d, tg, 0, 6, 0.5)
    ref = torch.cat([terms, gw.reshape(-1)])
    assert torch.allclose(got["flat"], ref, rtol=1e-6, atol=1e-9)
