#!/usr/bin/env python
"""One launch of every hot kernel of the path per iteration, in a fixed order, for `ncu --set full`
(tools/capture_profiles.sh): B = 4096 NYU, J = 14.

    python tools/prof_step.py [--iters 3]      # ncu: -k regex:'sfr_|decoder_' --launch-skip 2*N --launch-count N

Order (N = 16 matching kernels per iteration; sfr_prep_kernel precedes every sfr_build_kernel):
  SFR build  float32 frames, dense targets | raw uint16 (NYU G/B) frames + hand rectangle, dense | float32, compact targets
  fetch      plan + copy of the raw frames' crop windows (device-resident source: the PCIe side is in bench.py)
  one-pass   last stage, dense targets | compact targets | forward + loss only (inner stage, stats saved)
  forward    with the heat-map store (direct kernel) | without (pipelined kernel)
  backward   + loss, last stage (targets in the slots) | inner stage alpha = 1 (upstream maps in the slots) |
             inner stage alpha = 0.5 (six-slot stage: targets + upstream maps)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelwiseregression_b200 import ops, sfr, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--batch", type=int, default=4096)
args = ap.parse_args()
shape = synth.NYU
B, J = args.batch, shape.joints
d = synth.make_frames_device(shape, B, seed=0, device="cuda")
raw = d["frames"].round().clamp_(0, 65535).to(torch.int32).to(torch.uint16)
g = torch.Generator(device="cuda").manual_seed(1000)
z = torch.randn(B, J, 64, 64, device="cuda", generator=g)
D = torch.randn(B, J, 64, 64, device="cuda", generator=g)
w = torch.rand(J, 1, device="cuda", generator=g) + 0.5
gH = torch.randn(B, J, 64, 64, device="cuda", generator=g) * 1e-4
gD = torch.randn(B, J, 64, 64, device="cuda", generator=g) * 1e-4
one = torch.ones((), device="cuda")
kw = dict(fx=shape.fx, fy=shape.fy)
raw_kw = dict(kw, frame_format="nyu_gb16", prefilter=(40.0, shape.halfu, shape.halfv))
a1, a2, a3 = sfr.SfrArena(), sfr.SfrArena(), sfr.SfrArena()
win_hw = sfr.window_size(d["com"], d["cube"], shape.fx, shape.fy, shape.height, shape.width, "nyu_gb16")
fw = None
for it in range(args.iters):
    bt = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], arena=a1, targets="both", **kw)
    sfr.build_sfr(raw, d["com"], d["cube"], d["uvd"], arena=a2, **raw_kw)
    sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], arena=a3, targets="sparse", **kw)
    fw = sfr.fetch_windows(raw, d["com"], d["cube"], win_hw=win_hw, out=fw, **raw_kw)
    dense = (bt.heatmaps, bt.depthmaps, bt.uvd)
    sparse = ops.SparseTargets(bt.taps, bt.uvd)
    L, m = bt.label_img, bt.mask
    ops.decoder_fused_raw(z, w, D, L, m, dense, "softmax", 1.0)
    ops.decoder_fused_raw(z, w, D, L, m, sparse, "softmax", 1.0)
    H, uvd, stats, lp = ops.decoder_forward_raw(z, w, D, L, m, targets=dense)            # one-pass kernel, forward only
    ops.decoder_forward_raw(z, w, D, L, m)                                              # direct forward, H stored
    ops.decoder_forward_raw(z, w, D, L, m, store_heat=False, want_stats=False)          # pipelined forward
    ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, targets=dense, alpha=1.0, want_loss=True)
    ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, None, gH, gD, targets=dense, alpha=1.0, loss_scale_dev=one)
    ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, None, gH, gD, targets=dense, alpha=0.5, loss_scale_dev=one)
torch.cuda.synchronize()
print("ok")
