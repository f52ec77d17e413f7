#!/usr/bin/env python
"""Diagnostics of the double-buffered host feed: per-step wall time and the fetch kernel's own duration
(CUDA events on the copy stream) with / without the compute stream busy.
    python tools/ab_e2e.py [--compute none|sfr|full] [--steps 12] [--depth 2]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelwiseregression_b200 import _lib, feed, ops, sfr, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--compute", default="full")
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--depth", type=int, default=2)
ap.add_argument("--batch", type=int, default=4096)
args = ap.parse_args()
shape = synth.NYU
B, J = args.batch, shape.joints
d = synth.make_frames_device(shape, B, seed=0, device="cuda")
raw = d["frames"].round().clamp_(0, 65535).to(torch.int32).to(torch.uint16)
host_frames = torch.empty(raw.shape, dtype=raw.dtype, pin_memory=True).copy_(raw)
host = {k: d[k].cpu().numpy() for k in ("com", "cube", "uvd")}
z = torch.randn(B, J, 64, 64, device="cuda").requires_grad_(True)
D = torch.randn(B, J, 64, 64, device="cuda").requires_grad_(True)
w = (torch.rand(J, 1, device="cuda") + 0.5).requires_grad_(True)
del d, raw
hf = feed.HostFeed(shape, B, frame_format="nyu_gb16", prefilter=(40.0, shape.halfu, shape.halfv), depth=args.depth)
out_host = torch.empty(4 + B * J * 3, pin_memory=True)
done = torch.cuda.Event()


def consume(t):
    if args.compute == "none":
        torch.cuda.current_stream().wait_event(hf.slots[t % len(hf.slots)].ready)
        hf.slots[t % len(hf.slots)].consumed.record()
    else:
        batch = hf.build(t)
        if args.compute == "full":
            total, terms, uvd_out = ops.fused_decoder_loss(z, w, D, batch.label_img, batch.mask, batch.heatmaps, batch.depthmaps,
                                                           batch.uvd, store_heat=True)[:3]
            z.grad = D.grad = w.grad = None
            total.backward()
            out_host[:1].copy_(total.detach().reshape(1), non_blocking=True)
            out_host[4:].copy_(uvd_out.reshape(-1), non_blocking=True)
    done.record()


_lib.PROFILE = []
t = hf.submit(host_frames, host["com"], host["cube"], host["uvd"])
for _ in range(3):
    nxt = hf.submit(host_frames, host["com"], host["cube"], host["uvd"])
    consume(t)
    done.synchronize()
    t = nxt
torch.cuda.synchronize()
_lib.PROFILE = []
t0 = time.perf_counter()
for _ in range(args.steps):
    nxt = hf.submit(host_frames, host["com"], host["cube"], host["uvd"])
    consume(t)
    done.synchronize()
    t = nxt
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / args.steps * 1e3
ev = {}
for what, s, e in _lib.PROFILE:
    ev.setdefault(what, []).append(s.elapsed_time(e))
_lib.PROFILE = None
print("compute=%s depth=%d: %.3f ms/step (%.0f samples/s); kernel ms: %s" % (
    args.compute, args.depth, dt, B / dt * 1e3, {k: round(sum(v) / len(v), 3) for k, v in ev.items()}))
