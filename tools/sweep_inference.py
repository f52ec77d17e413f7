#!/usr/bin/env python
"""Large-batch inference sweep of the SFR + decoder path (BASELINE configs[4]: HAND17 shape,
J = 21, batch 256 ... 16384 per GPU).

    python tools/sweep_inference.py                       # one GPU
    torchrun --nproc-per-node 8 tools/sweep_inference.py  # every rank sweeps its own shard (weak)

One pass = what test.py:88-113 does around the backbone: test-only SFR of the frames
(`pwr_sfr_crop`), decoder forward of the last stage without storing the heat maps
(`pwr_decoder_fwd`, H = NULL), `pwr_recover_uvd` to pixel / mm coordinates.  The backbone is
cuDNN and out of scope, so its conv outputs are random logits resident in HBM.

Every batch size is timed twice: launched call by call from Python, and replayed as one CUDA
graph (the library never synchronises or allocates, so the whole pass captures).  Small batches are
bound by the host issuing 4 launches through ctypes; the graph removes that.

Prints one JSON line per batch size and a markdown table at the end (rank 0).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelwiseregression_b200 import _lib, ops, roofline, sfr, synth  # noqa: E402


def sweep(shape, batches, steps=20, warmup=5, world=1, rank=0, local=0, peak=None, echo=False):
    """Rows of the sweep (one dict per batch size); the process group must exist when world > 1."""
    dev = torch.device("cuda", local)
    if not peak:
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except (OSError, KeyError, ValueError):
            peak = 6650.0
    J = shape.joints
    intr = (shape.fx, shape.fy, shape.halfu, shape.halfv)
    w = torch.ones(J, 1, device=dev)
    rows = []
    for B in batches:
        d = synth.make_frames_device(shape, B, seed=rank, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(100 + rank)
        z = torch.randn(B, J, 64, 64, device=dev, generator=g)
        D = torch.randn(B, J, 64, 64, device=dev, generator=g)
        com = None if shape.com_from_frame else d["com"]
        arena = sfr.SfrArena()

        def one_pass():
            t = sfr.build_sfr(d["frames"], com, d["cube"], fx=shape.fx, fy=shape.fy, frame_f64=shape.frame_f64,
                              test_only=True, arena=arena)
            _, uvd, _, _ = ops.decoder_forward_raw(z, w, D, t.label_img, t.mask, store_heat=False, want_stats=False)
            return ops.recover_uvd(uvd, t.box_size, t.com, t.cube_size, intrinsics=intr)

        def timed(fn):
            for _ in range(warmup):
                fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(e) / steps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return ms.item()

        ms_calls = timed(one_pass)
        # per entry point: CUDA events around each launch on the launching stream
        _lib.PROFILE = []
        for _ in range(steps):
            one_pass()
        torch.cuda.synchronize()
        per_kernel = {}
        for what, s, e in _lib.PROFILE:
            per_kernel.setdefault(what, []).append(s.elapsed_time(e))
        _lib.PROFILE = None
        per_kernel = {k: sum(v) / len(v) for k, v in per_kernel.items()}
        # the same pass as one CUDA graph
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_pass()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = one_pass()
        ms_graph = timed(graph.replay)
        out = [o.clone() for o in out]
        ref = one_pass()
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(out, ref))
        bytes_per_sample = roofline.sfr_crop_bytes() + roofline.decoder_fwd_bytes(J, store_heat=False) + 36 * J
        row = dict(shape=shape.name, joints=J, n_gpus=world, batch_per_gpu=B, ms_calls=ms_calls, ms_graph=ms_graph,
                   samples_per_s_calls=B * world / ms_calls * 1e3, samples_per_s_graph=B * world / ms_graph * 1e3,
                   algorithmic_bytes_per_sample=bytes_per_sample,
                   roofline_frac_calls=B * bytes_per_sample / (ms_calls * 1e-3) / 1e9 / peak,
                   roofline_frac_graph=B * bytes_per_sample / (ms_graph * 1e-3) / 1e9 / peak,
                   kernels_ms=per_kernel, graph_equals_calls=bool(same), inputs_mb=(d["frames"].nbytes + z.nbytes + D.nbytes) / 1e6)
        rows.append(row)
        if echo and rank == 0:
            print(json.dumps(row), flush=True)
        del d, z, D, graph, out, ref, arena
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="HAND17")
    ap.add_argument("--batches", default="256,512,1024,2048,4096,8192,16384")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--peak-gbs", type=float, default=0.0, help="0 = MEASURED_PEAKS.json hbm_gbs")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = sweep(synth.SHAPES[args.shape], [int(b) for b in args.batches.split(",")], args.steps, args.warmup, world,
                 rank, local, args.peak_gbs, echo=True)
    if rank == 0:
        print("\n| batch/GPU | ms (calls) | ms (graph) | samples/s (graph, %d GPU) | %% HBM roofline (calls) | %% HBM roofline (graph) |" % world)
        print("|---|---|---|---|---|---|")
        for r in rows:
            print("| %d | %.3f | %.3f | %.3g | %.1f | %.1f |" % (r["batch_per_gpu"], r["ms_calls"], r["ms_graph"],
                                                                r["samples_per_s_graph"], 100 * r["roofline_frac_calls"],
                                                                100 * r["roofline_frac_graph"]))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
