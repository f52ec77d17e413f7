#!/usr/bin/env python
"""End-to-end training step on synthetic NYU/MSRA/HAND17-shaped data (BASELINE configs 3/4).

    python examples/train_synthetic.py --decoder fused --batch 128 --steps 20
    torchrun --nproc-per-node 8 examples/train_synthetic.py --decoder fused

One step = on-GPU SFR target build from raw frames -> PixelwiseRegression forward (cuDNN hourglass
backbone + decoder) -> loss of train.py:194-205 -> backward -> AdamW step, exactly the loop body of
train.py:158-208 with the DataLoader replaced by `sfr.build_sfr`.  `--decoder` selects how the decoder
and the loss run:
    eager   the reference's own lines (model.py:83-95,123-130,151; train.py:197-205) as PyTorch ops
    dropin  pixelwiseregression_b200.model (fused decoder kernels), loss lines unchanged
    fused   model.forward_loss (decoder + loss + backward fused)
Under torchrun the model is wrapped in DDP (NCCL); every rank builds its own batch.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelwiseregression_b200 import distributed as pd  # noqa: E402
from pixelwiseregression_b200 import model as M  # noqa: E402
from pixelwiseregression_b200 import sfr, synth  # noqa: E402


def eager_forward(net, img, label_img, mask):
    """PixelwiseRegression.forward with the decoder written as the reference writes it."""
    F = torch.nn.functional
    f = net.conv(img)
    results = []
    for stage in net.stages:
        f, z, d_raw = stage.features_and_logits(f)
        plane = stage.plane_regression
        B, J, H, W = z.shape
        heat = F.softmax(plane.w * z.view(B, J, -1), dim=2).view(B, J, H, W)
        u = torch.sum(plane.filter[0].view(1, 1, H, W) * heat, dim=(2, 3)).unsqueeze(-1)
        v = torch.sum(plane.filter[1].view(1, 1, H, W) * heat, dim=(2, 3)).unsqueeze(-1)
        rec = d_raw + label_img
        mh = heat * mask
        d = (torch.sum(mh * (mask * rec), dim=(2, 3)) / (torch.sum(mh, dim=(2, 3)) + 1e-14)).unsqueeze(-1)
        uvd = torch.cat([u, v, d], dim=2)
        results.append((heat, d_raw, uvd))
        f = torch.cat([heat, d_raw, label_img], dim=1)
    return results


def train_py_loss(results, uvd, heatmaps, depthmaps, alpha, lambda_h, lambda_d):
    loss = 0
    for _heatmaps, _depthmaps, _uvd in results:
        heatmap_loss = lambda_h * torch.mean(torch.sum((_heatmaps - heatmaps) ** 2, dim=(2, 3)))
        depthmap_loss = lambda_d * torch.mean(torch.sum((_depthmaps - depthmaps) ** 2, dim=(2, 3)))
        uvd_loss = torch.mean(torch.sum((_uvd - uvd) ** 2, dim=2))
        loss = loss + alpha * uvd_loss + (1 - alpha) * (heatmap_loss + depthmap_loss)
    return loss


class _LossModule(torch.nn.Module):
    """Lets DDP see the fused criterion as the wrapped module's forward."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, *a):
        return self.net.forward_loss(*a)[0]


def run_training(shape, batch, steps, warmup, decoder="fused", features=128, level=4, stages=2, alpha=1.0,
                 world=1, rank=0, local=0, on_timed_start=None):
    """One training configuration, timed on the device (CUDA events, max over ranks).  The process group must
    already exist when world > 1.  Returns dict(ms_per_step, samples_per_s, loss, peak_mem_gb)."""
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    net = M.PixelwiseRegression(shape.joints, stage=stages, features=features, level=level,
                                norm_method="instance").to(dev)
    # the fused criterion is exposed to DDP as a module whose forward returns the loss
    wrapped = _LossModule(net) if decoder == "fused" else net
    model = pd.wrap_ddp(wrapped, local) if world > 1 else wrapped
    optim = torch.optim.AdamW(net.parameters(), lr=1e-3)
    d = synth.make_frames_device(shape, batch, seed=rank, device=dev)
    lambda_h, lambda_d = 1.0, 0.01
    arena = sfr.SfrArena()

    def step():
        b = sfr.build_sfr(d["frames"], None if shape.com_from_frame else d["com"], d["cube"], d["uvd"],
                          fx=shape.fx, fy=shape.fy, frame_f64=shape.frame_f64, arena=arena)
        # the reference raises on samples it cannot build (datasets.py:323-327, 362-365, 385-390), so they never
        # reach the loss; here they are flagged and dropped before the model (a no-op when all are valid)
        b = sfr.select_valid(b)
        optim.zero_grad(set_to_none=True)
        if decoder == "fused":
            loss = model(b.img, b.label_img, b.mask, b.uvd, b.heatmaps, b.depthmaps, alpha, lambda_h, lambda_d)
        else:
            if decoder == "eager":
                results = eager_forward(net, b.img, b.label_img, b.mask)
            else:
                results = model(b.img, b.label_img, b.mask)
            loss = train_py_loss(results, b.uvd, b.heatmaps, b.depthmaps, alpha, lambda_h, lambda_d)
        loss.backward()
        optim.step()
        return loss

    torch.cuda.reset_peak_memory_stats(dev)
    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if on_timed_start is not None:
        on_timed_start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"ms_per_step": ms.item(), "samples_per_s": batch * world / ms.item() * 1e3, "loss": loss.item(),
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    del model, wrapped, net, optim, d, arena
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decoder", default="fused", choices=["eager", "dropin", "fused"])
    ap.add_argument("--shape", default="NYU")
    ap.add_argument("--batch", type=int, default=128, help="per GPU")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--level", type=int, default=4)
    ap.add_argument("--stages", type=int, default=2)
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--json", default="", help="also write the result as one JSON line to this file (rank 0)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = synth.SHAPES[args.shape]
    r = run_training(shape, args.batch, args.steps, args.warmup, args.decoder, args.features, args.level, args.stages,
                     args.alpha, world, rank, local)
    if rank == 0:
        print("decoder=%s shape=%s gpus=%d batch/gpu=%d: %.2f ms/step, %.0f samples/s, loss %.5f, peak mem %.1f GB" % (
            args.decoder, shape.name, world, args.batch, r["ms_per_step"], r["samples_per_s"], r["loss"],
            r["peak_mem_gb"]))
        if args.json:
            import json
            with open(args.json, "w") as f:
                f.write(json.dumps({
                    "metric": "end-to-end training samples/s", "value": r["samples_per_s"],
                    "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": r["ms_per_step"], "scaling": "weak", "dtype": "f32", "data": "synthetic",
                    "config": {"workload": "configs[2]: %s-shape end-to-end training step (on-GPU SFR build, cuDNN "
                                           "hourglass backbone, fused decoder + loss, AdamW), batch %d/GPU, DDP over NCCL"
                                           % (shape.name, args.batch),
                               "decoder": args.decoder, "features": args.features, "stages": args.stages,
                               "level": args.level, "joints": shape.joints},
                    "loss": r["loss"], "peak_mem_gb": r["peak_mem_gb"]}) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
