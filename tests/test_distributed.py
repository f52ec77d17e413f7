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
