"""Batched SFR target builder on the GPU (pwr_sfr_* of include/pwr.h).

Mirrors HandDataset.process_single_data, non-augmented branch
(datasets.py:301-403), for a whole batch at once: the returned tensors have
exactly the shapes and dtypes `default_collate` produces from the reference's
9-tuple (train) / 6-tuple (test_only), plus `valid [B] uint8`, which replaces
the reference's exception path (datasets.py:323-327, 362-365, 385-390).
"""
from collections import namedtuple

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

SFRBatch = namedtuple("SFRBatch", ["img", "label_img", "mask", "box_size", "cube_size", "com", "uvd", "heatmaps",
                                   "depthmaps", "valid", "taps"], defaults=(None,))
SFRTestBatch = namedtuple("SFRTestBatch", ["img", "label_img", "mask", "box_size", "cube_size", "com", "valid"])


def _f64(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=device)


def center_of_mass(frames):
    """datasets.py:208-211 for a batch: [B,Hf,Wf] float32 CUDA -> com [B,3] float64
    (mean column, mean row, mean depth of the pixels > 0)."""
    require_cuda(frames)
    if frames.dtype != torch.float32 or frames.dim() != 3:
        raise _lib.PwrError("frames must be a [B, Hf, Wf] float32 tensor")
    frames = frames.contiguous()
    B, Hf, Wf = frames.shape
    com = torch.empty(B, 3, device=frames.device, dtype=torch.float64)
    with torch.cuda.device(frames.device):
        rc = _lib.load().pwr_sfr_com(ptr(frames), Hf, Wf, ptr(com), B, stream_ptr(frames.device))
    check(rc, "pwr_sfr_com")
    return com


_FRAME_DTYPES = {"f32": torch.float32, "nyu_gb16": torch.uint16, "u16": torch.uint16}


def draw_augmentation(batch, generator=None, rotation=True, scale=True, shift=True):
    """The random draws of the reference's augmented branch for `batch` samples, as a float64
    [B,4] array (scale, shift_u, shift_v, angle_deg): scale = 0.8 + U*0.4 (datasets.py:230),
    shifts = -5 + U*10 (:236-237), angle = U*60 - 30.  Upstream quirks kept: the rotation is
    redrawn inside random_rotated and applied whenever ANY augmentation is on (utils.py:72),
    and the "mm" shift lands on the pixel coordinates of the centre (uvd2xyz is a no-op on a
    1-D array, :238-241).  `generator`: a numpy Generator (default: fresh, unseeded)."""
    rng = generator if generator is not None else np.random.default_rng()
    aug = np.empty((batch, 4), dtype=np.float64)
    aug[:, 0] = 0.8 + rng.random(batch) * 0.4 if scale else 1.0
    aug[:, 1] = -5 + rng.random(batch) * 10 if shift else 0.0
    aug[:, 2] = -5 + rng.random(batch) * 10 if shift else 0.0
    aug[:, 3] = rng.random(batch) * 60 - 30
    if not (rotation or scale or shift):
        return None
    return aug


def _aug_device_params(aug, B, device):
    """[B,4] (scale, shift_u, shift_v, angle_deg) -> the [B,8] float64 block of include/pwr.h.
    The trigonometry is done here in float64 NumPy (the reference's libm), not on the GPU."""
    aug = np.asarray(aug.cpu() if isinstance(aug, torch.Tensor) else aug, dtype=np.float64)
    if aug.shape != (B, 4):
        raise _lib.PwrError("augment must be [B, 4] = (scale, shift_u, shift_v, angle_deg)")
    a_m = aug[:, 3] * (np.pi / 180)          # cv2.getRotationMatrix2D: angle *= CV_PI/180
    a_j = aug[:, 3] / 180.0 * np.pi          # utils.py:77
    out = np.zeros((B, 8), dtype=np.float64)
    out[:, 0:3] = aug[:, 0:3]
    out[:, 3], out[:, 4] = np.cos(a_m), np.sin(a_m)
    out[:, 5], out[:, 6] = np.cos(a_j), np.sin(a_j)
    return torch.from_numpy(out).to(device)


def build_sfr(frames, com, cube, uvd=None, *, fx, fy, frame_f64=False, test_only=False, frame_format="f32",
              prefilter=None, augment=None, targets="dense"):
    """frames [B,Hf,Wf] CUDA depth frames: float32 mm (`frame_format="f32"`, what
    process_single_data receives), or raw sensor samples decoded on the fly exactly as
    the reference's loaders do: "nyu_gb16" = uint16 G<<8|B of the NYU PNG
    (datasets.py:810), "u16" = 16-bit grey PNG (ICVL :632, HAND17 :940).
    `prefilter=(margin, halfu, halfv)` applies the hand rectangle of load_from_text
    (NYU/HAND17 margin 40, ICVL 30; datasets.py:841-853) inside the crop taps.
    `targets`: "dense" renders heatmaps / depthmaps [B,J,64,64] like the reference; "sparse" returns
    them as `taps` [B,J,64] uint8 (one 64-byte pwr_joint_taps per joint) for the loss kernels to
    evaluate on the fly (ops.fused_decoder_loss / model.forward_loss accept it in place of the
    dense heat maps), which removes the write and re-read of 2J maps per sample; "both" does both.
    `augment` = [B,4] (scale, shift_u, shift_v, angle_deg), e.g. from `draw_augmentation`: the
    reference's augmented branch (datasets.py:216-299, train.py's default flags) with these
    draws; samples whose augmented branch would raise fall back to the plain branch.
    com [B,3] float64 hand
    centre (u, v, z) or None to use the centre-of-mass fallback; cube [B] (or a
    scalar) half cube size; uvd [B,J,3] float64 joint annotations (train mode).
    `frame_f64=True` reproduces datasets whose frames the reference holds as
    float64 (MSRA, datasets.py:516).

    Returns SFRBatch (train) or SFRTestBatch (test_only), all float32 except
    `valid` (uint8)."""
    require_cuda(frames)
    if frame_format not in _FRAME_DTYPES:
        raise _lib.PwrError("unknown frame_format %r" % (frame_format,))
    if frames.dtype != _FRAME_DTYPES[frame_format] or frames.dim() != 3:
        raise _lib.PwrError("frames must be a [B, Hf, Wf] %s tensor for frame_format=%r" % (
            _FRAME_DTYPES[frame_format], frame_format))
    if frame_f64 and frame_format != "f32":
        raise _lib.PwrError("float64 frame semantics (MSRA) exist for decoded float32 frames only")
    lib = _lib.load()
    frames = frames.contiguous()
    dev = frames.device
    B, Hf, Wf = frames.shape
    fmt = _lib.FRAME_FORMATS[frame_format]
    pf = (-1.0, 0.0, 0.0) if prefilter is None else (float(prefilter[0]), 2.0 * prefilter[1], 2.0 * prefilter[2])
    if com is None:
        if frame_format != "f32":
            raise _lib.PwrError("the centre-of-mass fallback (MSRA) takes decoded float32 frames")
        com = center_of_mass(frames)
    com = _f64(com, dev)
    if not isinstance(cube, torch.Tensor) and np.ndim(cube) == 0:
        cube = np.full(B, float(cube))
    cube = _f64(cube, dev)
    if tuple(com.shape) != (B, 3) or tuple(cube.shape) != (B,):
        raise _lib.PwrError("com must be [B,3] and cube [B]")
    f32 = dict(device=dev, dtype=torch.float32)
    img = torch.empty(B, 1, 128, 128, **f32)
    label_img = torch.empty(B, 1, 64, 64, **f32)
    mask = torch.empty(B, 1, 64, 64, **f32)
    box_size = torch.empty(B, **f32)
    cube_size = torch.empty(B, **f32)
    com_out = torch.empty(B, 3, **f32)
    valid = torch.empty(B, device=dev, dtype=torch.uint8)
    s = stream_ptr(dev)
    J = 0 if (test_only or uvd is None) else int(uvd.shape[1])
    ws_bytes = int(lib.pwr_sfr_workspace_bytes(B, J))
    workspace = torch.empty(max(ws_bytes, 16), device=dev, dtype=torch.uint8)     # scratch, no init needed
    if test_only:
        if augment is not None:
            raise _lib.PwrError("you can not transform the test data")     # datasets.py:64-65
        with torch.cuda.device(dev), _lib.timed("pwr_sfr_crop"):
            rc = lib.pwr_sfr_crop(ptr(frames), fmt, Hf, Wf, ptr(com), ptr(cube), float(fx), float(fy), int(frame_f64),
                                  pf[0], pf[1], pf[2], ptr(img), ptr(label_img), ptr(mask), ptr(box_size), ptr(cube_size), ptr(com_out),
                                  ptr(valid), ptr(workspace), ws_bytes, B, s)
        check(rc, "pwr_sfr_crop")
        return SFRTestBatch(img, label_img, mask, box_size, cube_size, com_out, valid)
    if uvd is None:
        raise _lib.PwrError("train-mode SFR needs joint annotations (uvd)")
    uvd = _f64(uvd, dev)
    if uvd.dim() != 3 or uvd.shape[0] != B or uvd.shape[2] != 3:
        raise _lib.PwrError("uvd must be [B, J, 3]")
    J = uvd.shape[1]
    if targets not in ("dense", "sparse", "both"):
        raise _lib.PwrError("targets must be 'dense', 'sparse' or 'both'")
    uvd_norm = torch.empty(B, J, 3, **f32)
    heatmaps = torch.empty(B, J, 64, 64, **f32) if targets != "sparse" else None
    dmap = torch.empty(B, J, 64, 64, **f32) if targets != "sparse" else None
    taps = torch.empty(B, J, 64, device=dev, dtype=torch.uint8) if targets != "dense" else None
    aug_dev = _aug_device_params(augment, B, dev) if augment is not None else None
    with torch.cuda.device(dev), _lib.timed("pwr_sfr_build"):
        rc = lib.pwr_sfr_build(ptr(frames), fmt, Hf, Wf, ptr(com), ptr(cube), ptr(uvd), ptr(aug_dev), float(fx), float(fy),
                               int(frame_f64), pf[0], pf[1], pf[2], ptr(img), ptr(label_img), ptr(mask), ptr(box_size), ptr(cube_size),
                               ptr(com_out), ptr(uvd_norm), ptr(heatmaps), ptr(dmap), ptr(taps), ptr(valid),
                               ptr(workspace), ws_bytes, B, J, s)
    check(rc, "pwr_sfr_build")
    return SFRBatch(img, label_img, mask, box_size, cube_size, com_out, uvd_norm, heatmaps, dmap, valid, taps)
