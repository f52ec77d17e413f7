#!/usr/bin/env python
"""A/B timing of the SFR builder alone (B = 4096 NYU): PWR_LIB_PATH selects the library variant.
    python tools/ab_sfr.py [--arena 0|1] [--format f32|nyu_gb16] [--targets dense|sparse] [--test-only]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelwiseregression_b200 import sfr, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--arena", type=int, default=1)
ap.add_argument("--format", default="f32")
ap.add_argument("--targets", default="dense")
ap.add_argument("--test-only", action="store_true")
ap.add_argument("--shape", default="NYU")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
shape = synth.SHAPES[args.shape]
d = synth.make_frames_device(shape, args.batch, seed=0, device="cuda")
frames = d["frames"]
kw = dict(fx=shape.fx, fy=shape.fy)
if args.format != "f32":
    frames = frames.round().clamp_(0, 65535).to(torch.int32).to(torch.uint16)
    kw.update(frame_format=args.format, prefilter=(40.0, shape.halfu, shape.halfv))
arena = sfr.SfrArena() if args.arena else None
uvd = None if args.test_only else d["uvd"]


def run():
    return sfr.build_sfr(frames, d["com"], d["cube"], uvd, test_only=args.test_only, arena=arena,
                         **(kw if args.test_only else dict(kw, targets=args.targets)))


for _ in range(5):
    run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(args.iters):
    run()
e.record()
torch.cuda.synchronize()
print("lib=%s arena=%d format=%s targets=%s test_only=%s shape=%s: %.4f ms per build" % (
    os.environ.get("PWR_LIB_PATH", "default"), args.arena, args.format, args.targets, args.test_only, args.shape,
    s.elapsed_time(e) / args.iters))
