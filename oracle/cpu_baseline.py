"""CPU timing of the oracle port (used ONLY by bench.py's `cpu_baseline` leg and
by `bench.py --impl reference`): the reference's SFR + decoder + loss path
restated in NumPy/PyTorch, run on the box's host cores.

SFR runs one sample per task across a process pool (the reference runs
process_single_data in `cpu_count` DataLoader workers, train.py:99); the
decoder + loss forward/backward (model.py:79-132, train.py:197-207) runs as
PyTorch CPU ops with all intra-op threads.  OpenCV is used for resize/blur when
importable (as the reference does), else the NumPy restatement.
"""
import os
import time

import numpy as np
import torch

from . import decoder_oracle as do
from . import sfr_oracle as so

_POOL_STATE = {}


def _backend():
    try:
        import cv2  # noqa: F401
        cv2.setNumThreads(1)
        return "cv2"
    except Exception:
        return "numpy"


def _sfr_one(b):
    s = _POOL_STATE
    return so.process_sample(s["frames"][b], s["uvd"][b], s["com"][b], s["cube"][b], s["fx"], s["fy"],
                             backend=s["backend"])


_POOL = {}


def _get_pool(workers):
    """Persistent fork pool (created after the inputs are in _POOL_STATE so the
    children inherit them copy-on-write, like DataLoader workers inherit the dataset)."""
    import multiprocessing as mp
    key = (workers, id(_POOL_STATE.get("frames")))
    if _POOL.get("key") != key:
        if _POOL.get("pool") is not None:
            _POOL["pool"].terminate()
        _POOL["pool"] = mp.get_context("fork").Pool(workers)
        _POOL["key"] = key
    return _POOL["pool"]


def run_sfr(frames, uvd, com, cube, fx, fy, workers):
    """Returns the stacked SFR batch (dict of arrays)."""
    if _POOL_STATE.get("frames") is not frames:
        _POOL_STATE.update(frames=frames, uvd=uvd, com=com, cube=cube, fx=fx, fy=fy, backend=_backend())
    n = len(frames)
    if workers <= 1:
        outs = [_sfr_one(b) for b in range(n)]
    else:
        outs = _get_pool(workers).map(_sfr_one, range(n), chunksize=max(1, n // (workers * 4)))
    return {k: np.stack([o[k] for o in outs]) for k in so.FIELDS}


def shutdown():
    if _POOL.get("pool") is not None:
        _POOL["pool"].terminate()
        _POOL.clear()
    _POOL_STATE.clear()


def run_decoder(z, w, D, batch, alpha=1.0, lambda_h=1.0, lambda_d=0.01):
    """Decoder forward + train.py loss + autograd backward on CPU tensors."""
    z = z.clone().requires_grad_(True)
    D = D.clone().requires_grad_(True)
    w = w.clone().requires_grad_(True)
    p, Dm, uvd = do.decoder_forward(z, w, D, torch.from_numpy(batch["label_img"]), torch.from_numpy(batch["mask"]))
    terms = do.stage_losses(p, Dm, uvd, torch.from_numpy(batch["heatmaps"]), torch.from_numpy(batch["dmap"]),
                            torch.from_numpy(batch["uvd"]), lambda_h, lambda_d)
    loss = do.combine_losses(terms, alpha)
    loss.backward()
    return loss.item()


def prepare(shape, n, seed=0):
    """Synthetic inputs of one bounded sample (generated outside any timed region)."""
    from pixelwiseregression_b200 import synth
    d = synth.make_frames(shape, n, seed)
    g = torch.Generator().manual_seed(seed)
    d["z"] = torch.randn(n, shape.joints, 64, 64, generator=g)
    d["D"] = torch.randn(n, shape.joints, 64, 64, generator=g)
    d["w"] = torch.rand(shape.joints, 1, generator=g) + 0.5
    return d


def run_path(shape, d, workers):
    """One pass of the whole path over the prepared sample; returns (total, sfr, decoder) seconds."""
    t0 = time.perf_counter()
    batch = run_sfr(d["frames"], d["uvd"], d["com"], d["cube"], shape.fx, shape.fy, workers)
    t1 = time.perf_counter()
    run_decoder(d["z"], d["w"], d["D"], batch)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1


def describe(shape, n, workers):
    return ("%d %s-shape samples per pass: SFR build (%s, %d processes) + decoder fwd + loss + bwd "
            "(torch CPU, %d threads)" % (n, shape.name, _backend(), workers, workers))


def time_path(shape, n, seed=0, workers=None, repeats=1):
    """Time `repeats` passes of the whole path over `n` synthetic samples (inputs prepared
    and the worker pool warmed outside the timed region).
    Returns dict(samples_per_s, seconds=[...], cores, sample)."""
    workers = workers or os.cpu_count() or 1
    torch.set_num_threads(workers)
    d = prepare(shape, n, seed)
    run_sfr(d["frames"][:workers], d["uvd"][:workers], d["com"][:workers], d["cube"][:workers], shape.fx, shape.fy, 1)
    _POOL_STATE.clear()
    run_path(shape, {k: (v[:min(n, 2 * workers)] if k != "w" else v) for k, v in d.items()}, workers)   # warm pool
    _POOL_STATE.clear()
    secs = [run_path(shape, d, workers) for _ in range(repeats)]
    best = min(s[0] for s in secs)
    shutdown()
    return dict(samples_per_s=n / best, seconds=secs, cores=workers, sample=describe(shape, n, workers))
