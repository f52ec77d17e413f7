"""Deterministic synthetic depth frames / joints / logits.

No datasets or checkpoints exist on the GPU box, so every test and benchmark
runs on synthetic inputs shaped like the reference's datasets.  The constants
are the reference's per-dataset intrinsics and cube sizes:
NYU  datasets.py:693-696, MSRA datasets.py:406-409, HAND17 datasets.py:862-865,
ICVL datasets.py:521-524.
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class DatasetShape:
    name: str
    fx: float
    fy: float
    halfu: float
    halfv: float
    height: int
    width: int
    cube: float
    joints: int
    z_range: tuple
    frame_f64: bool = False      # reference holds the frame as float64 (MSRA)
    com_from_frame: bool = False  # load_from_text returns com=None (MSRA)


NYU = DatasetShape("NYU", 588.037, 587.075, 320, 240, 480, 640, 150, 14, (500.0, 1000.0))
HAND17 = DatasetShape("HAND17", 475.065948, 475.065857, 315.944855, 245.287079, 480, 640, 150, 21,
                      (500.0, 1000.0))
ICVL = DatasetShape("ICVL", 241.42, 241.42, 160, 120, 240, 320, 125, 16, (300.0, 550.0))
MSRA = DatasetShape("MSRA", 241.42, 241.42, 160, 120, 240, 320, 125, 21, (300.0, 550.0),
                    frame_f64=True, com_from_frame=True)

SHAPES = {s.name: s for s in (NYU, HAND17, ICVL, MSRA)}


def make_frames(shape: DatasetShape, batch: int, seed: int = 0, mixed_cube: bool = True):
    """Host (NumPy) generator used by the parity tests.

    Returns dict(frames [B,Hf,Wf] float32, com [B,3] float64, cube [B] float64,
    uvd [B,J,3] float64).  A disc of radius 0.7*cube/z*fx px around the CoM is
    filled with z + U(-100,100) mm; 5 % of the disc is clutter beyond the cube
    and 5 % exact zeros, so the depth window and the mask are both exercised.
    CoMs are non-integer and far enough off-centre that crop boxes straddle the
    frame edge for part of the batch (zero-padding path)."""
    rng = np.random.default_rng(seed)
    hf, wf, J = shape.height, shape.width, shape.joints
    frames = np.zeros((batch, hf, wf), dtype=np.float32)
    com = np.empty((batch, 3), dtype=np.float64)
    com[:, 0] = rng.uniform(0.25, 0.75, batch) * wf
    com[:, 1] = rng.uniform(0.25, 0.75, batch) * hf
    com[:, 2] = rng.uniform(shape.z_range[0], shape.z_range[1], batch)
    cube = np.full(batch, float(shape.cube))
    if mixed_cube and shape.name == "NYU":
        # datasets.py:818-819: a second test subject uses int(cube*5/6)
        cube[rng.uniform(size=batch) < 0.3] = float(int(shape.cube * 5 / 6))
    uvd = np.empty((batch, J, 3), dtype=np.float64)
    yy, xx = np.mgrid[0:hf, 0:wf]
    for b in range(batch):
        u0, v0, z0 = com[b]
        rad = 0.7 * cube[b] / z0 * shape.fx
        disc = (xx - u0) ** 2 + (yy - v0) ** 2 < rad * rad
        n = int(disc.sum())
        vals = z0 + rng.uniform(-100.0, 100.0, n)
        sel = rng.uniform(size=n)
        vals[sel < 0.05] = z0 + cube[b] + rng.uniform(1.0, 300.0, int((sel < 0.05).sum()))
        vals[sel > 0.95] = 0.0
        frames[b][disc] = vals.astype(np.float32)
        box = max(int(cube[b] / z0 * shape.fx + cube[b] / z0 * shape.fy), 2)
        shift = box // 2
        uvd[b, :, 0] = u0 + rng.uniform(-0.6, 0.6, J) * shift
        uvd[b, :, 1] = v0 + rng.uniform(-0.6, 0.6, J) * shift
        uvd[b, :, 2] = z0 + rng.uniform(-100.0, 100.0, J)
    if shape.com_from_frame:
        # the CoM is recomputed from the frame by the SFR builder; keep the
        # generator's value only as a hint for joint placement
        pass
    return dict(frames=frames, com=com, cube=cube, uvd=uvd)


def make_decoder_inputs(batch: int, joints: int, seed: int = 0, hw: int = 64):
    """Decoder micro-benchmark logits: z, D ~ N(0,1) float32,
    w ~ U(0.5,1.5), label in [-1,1] with a binary mask."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((batch, joints, hw, hw)).astype(np.float32)
    D = rng.standard_normal((batch, joints, hw, hw)).astype(np.float32)
    w = rng.uniform(0.5, 1.5, (joints, 1)).astype(np.float32)
    mask = (rng.uniform(size=(batch, 1, hw, hw)) < 0.4).astype(np.float32)
    label = (rng.uniform(-1.0, 1.0, (batch, 1, hw, hw)).astype(np.float32)) * mask
    return dict(z=z, D=D, w=w, label=label, mask=mask)


def make_frames_device(shape: DatasetShape, batch: int, seed: int = 0, device="cuda", chunk: int = 256,
                       mixed_cube: bool = True):
    """Device-side generator with the same recipe as make_frames (different random
    stream), for batches too large to build in NumPy (the B=4096 micro-benchmark
    holds 5 GB of NYU frames).  Returns CUDA tensors: frames [B,Hf,Wf] float32,
    com [B,3] float64, cube [B] float64, uvd [B,J,3] float64."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    hf, wf, J = shape.height, shape.width, shape.joints
    f64 = dict(device=device, dtype=torch.float64)
    rnd = lambda *s: torch.rand(*s, generator=g, **f64)
    com = torch.stack([(0.25 + 0.5 * rnd(batch)) * wf, (0.25 + 0.5 * rnd(batch)) * hf,
                       shape.z_range[0] + (shape.z_range[1] - shape.z_range[0]) * rnd(batch)], dim=1)
    cube = torch.full((batch,), float(shape.cube), **f64)
    if mixed_cube and shape.name == "NYU":
        cube = torch.where(rnd(batch) < 0.3, torch.full_like(cube, float(int(shape.cube * 5 / 6))), cube)
    box = torch.clamp((cube / com[:, 2] * shape.fx + cube / com[:, 2] * shape.fy).floor(), min=2)
    shift = (box / 2).floor()
    uvd = torch.stack([com[:, 0:1] + (rnd(batch, J) * 1.2 - 0.6) * shift[:, None],
                       com[:, 1:2] + (rnd(batch, J) * 1.2 - 0.6) * shift[:, None],
                       com[:, 2:3] + (rnd(batch, J) * 200.0 - 100.0)], dim=2)
    frames = torch.empty(batch, hf, wf, device=device, dtype=torch.float32)
    yy = torch.arange(hf, device=device, dtype=torch.float32).view(1, hf, 1)
    xx = torch.arange(wf, device=device, dtype=torch.float32).view(1, 1, wf)
    for s in range(0, batch, chunk):
        e = min(s + chunk, batch)
        c = com[s:e].float()
        cb = cube[s:e].float()
        rad = 0.7 * cb / c[:, 2] * shape.fx
        disc = (xx - c[:, 0].view(-1, 1, 1)) ** 2 + (yy - c[:, 1].view(-1, 1, 1)) ** 2 < (rad * rad).view(-1, 1, 1)
        val = c[:, 2].view(-1, 1, 1) + (torch.rand(e - s, hf, wf, generator=g, device=device) * 200.0 - 100.0)
        sel = torch.rand(e - s, hf, wf, generator=g, device=device)
        clutter = c[:, 2].view(-1, 1, 1) + cb.view(-1, 1, 1) + 1.0 + 299.0 * torch.rand(
            e - s, hf, wf, generator=g, device=device)
        val = torch.where(sel < 0.05, clutter, val)
        val = torch.where(sel > 0.95, torch.zeros_like(val), val)
        frames[s:e] = torch.where(disc, val, torch.zeros_like(val))
    return dict(frames=frames, com=com, cube=cube, uvd=uvd)
