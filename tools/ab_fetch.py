#!/usr/bin/env python
"""Standalone timing of pwr_sfr_fetch from pinned host memory (B = 4096 NYU raw frames).
    python tools/ab_fetch.py [--prefilter 0|1] [--batch 4096]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelwiseregression_b200 import sfr, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--prefilter", type=int, default=1)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--device-frames", action="store_true")
args = ap.parse_args()
shape = synth.NYU
d = synth.make_frames_device(shape, args.batch, seed=0, device="cuda")
raw = d["frames"].round().clamp_(0, 65535).to(torch.int32).to(torch.uint16)
host = raw if args.device_frames else torch.empty(raw.shape, dtype=raw.dtype, pin_memory=True).copy_(raw)
kw = dict(fx=shape.fx, fy=shape.fy, frame_format="nyu_gb16", prefilter=(40.0, shape.halfu, shape.halfv) if args.prefilter else None)
win_hw = sfr.window_size(d["com"], d["cube"], shape.fx, shape.fy, shape.height, shape.width, "nyu_gb16")
win_hw = (win_hw[0], min(win_hw[1] + 128, shape.width))          # room for the wider alignments of the A/B builds
fw = sfr.fetch_windows(host, d["com"], d["cube"], win_hw=win_hw, **kw)
assert int(fw.status) == 0
torch.cuda.synchronize()
fetched = int(fw.fetched_bytes)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(args.iters):
    sfr.fetch_windows(host, d["com"], d["cube"], win_hw=win_hw, out=fw, **kw)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / args.iters
ext = fw.extent.cpu()
print("lib=%s prefilter=%d B=%d win=%s: %.3f ms per fetch, %d B (%.1f KB/sample, mean region %.0f x %.0f px) -> %.1f GB/s, %.0f samples/s" % (
    os.environ.get("PWR_LIB_PATH", "default"), args.prefilter, args.batch, win_hw, ms, fetched, fetched / args.batch / 1e3,
    ext[:, 2].float().mean(), ext[:, 3].float().mean(), fetched / ms / 1e6, args.batch / ms * 1e3))
