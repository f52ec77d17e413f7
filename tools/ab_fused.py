#!/usr/bin/env python
"""A/B of the one-pass last stage without dense target maps (compact targets / uvd-only loss), B = 4096 NYU:
the lean three-CTA/SM kernel against decoder_fused_kernel (option fused_no_lean), outputs compared bit for bit,
then both timed.  PWR_LIB_PATH selects a library variant (build flags -DPWR_FL_*).
    python tools/ab_fused.py [--batch 4096] [--iters 30]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelwiseregression_b200 import _lib, ops, roofline, sfr, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--method", default="softmax")
args = ap.parse_args()
shape = synth.NYU
B, J = args.batch, shape.joints
d = synth.make_frames_device(shape, B, seed=0, device="cuda")
batch = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, targets="sparse")
del d
g = torch.Generator(device="cuda").manual_seed(1000)
z = torch.randn(B, J, 64, 64, device="cuda", generator=g)
D = torch.randn(B, J, 64, 64, device="cuda", generator=g)
w = (torch.rand(J, 1, device="cuda", generator=g) + 0.5) if args.method == "softmax" else None
sparse = ops.SparseTargets(batch.taps, batch.uvd)
L, m = batch.label_img, batch.mask
CASES = {"compact targets, alpha 0.5": dict(alpha=0.5),
         "compact targets, alpha 1": dict(alpha=1.0),
         "compact targets, no heat store": dict(alpha=1.0, store_heat=False),
         "uvd term only (no logged terms)": dict(alpha=1.0, want_loss=False),
         "forward + loss only": dict(alpha=1.0, want_grads=False)}


def run(kw):
    return ops.decoder_fused_raw(z, w, D, L, m, sparse, args.method, **kw)


def timed(kw):
    for _ in range(5):
        run(kw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.iters):
        run(kw)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / args.iters


print("lib=%s B=%d J=%d method=%s" % (os.environ.get("PWR_LIB_PATH", "default"), B, J, args.method))
for name, kw in CASES.items():
    lean = run(kw)
    with _lib.option("fused_no_lean", 1):
        ref = run(kw)
    torch.cuda.synchronize()
    same = all((x is None and y is None) or torch.equal(x, y) for x, y in zip(lean, ref))
    worst = max([float((x.float() - y.float()).abs().max() / (y.float().abs().max() + 1e-30))
                 for x, y in zip(lean, ref) if x is not None] + [0.0])
    t_lean = timed(kw)
    with _lib.option("fused_no_lean", 1):
        t_ref = timed(kw)
    maps_out = (1 if kw.get("store_heat", True) else 0) + (2 if kw.get("want_grads", True) else 0)
    nbytes = B * ((2 * J + 2 + maps_out * J) * 16384 + 80 * J)
    print("%-34s lean %.4f ms (%.0f GB/s)  two-CTA kernel %.4f ms  bit-identical=%s worst rel diff %.2e" % (
        name, t_lean, nbytes / t_lean / 1e6, t_ref, same, worst))
