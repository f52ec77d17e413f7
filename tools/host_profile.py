#!/usr/bin/env python
"""Where does the HOST time of one step go?  cProfile over the public-API step at a small batch
(kernels are microseconds there, so the wall clock is the host issuing them).
    python tools/host_profile.py [--batch 128] [--iters 300]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelwiseregression_b200 import ops, sfr, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=300)
args = ap.parse_args()
shape = synth.NYU
B, J = args.batch, shape.joints
d = synth.make_frames_device(shape, B, seed=0, device="cuda")
z = torch.randn(B, J, 64, 64, device="cuda").requires_grad_(True)
D = torch.randn(B, J, 64, 64, device="cuda").requires_grad_(True)
w = (torch.rand(J, 1, device="cuda") + 0.5).requires_grad_(True)
arena = sfr.SfrArena()


def step():
    b = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, arena=arena)
    total, terms, uvd = ops.fused_decoder_loss(z, w, D, b.label_img, b.mask, b.heatmaps, b.depthmaps, b.uvd,
                                               store_heat=False)
    z.grad = D.grad = w.grad = None
    total.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.iters):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue %.1f us/step, with final sync %.1f us/step" % ((t1 - t0) / args.iters * 1e6, (t2 - t0) / args.iters * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(args.iters):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
