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

FrameWindows = namedtuple("FrameWindows", ["windows", "extent", "frame_hw", "fetched_bytes", "status"])
FrameWindows.__doc__ = """What `fetch_windows` leaves on the device: `windows` [B, win_h, win_w] (the frames' own
element type) holding, per sample, only the region of the frame the builder reads; `extent` [B,4] int32 = (row0,
col0, rows, cols) of that region in frame coordinates; `frame_hw` = (Hf, Wf) of the original frames;
`fetched_bytes` (uint64 device scalar: bytes pulled from the source) and `status` (int32 device scalar, non-zero
if a region did not fit).  Pass it to `build_sfr` in place of the frames."""


class SfrArena:
    """Output tensors + scratch of `build_sfr`, allocated once and re-used by every call that passes
    `arena=` (same batch, joints, mode): no allocator traffic on the launch path, fixed addresses for CUDA-graph
    capture.  The returned SFRBatch aliases the arena: it is overwritten by the next call."""

    def __init__(self):
        self.key = None
        self.t = {}

    def get(self, key, make):
        if self.key != key:
            self.key, self.t = key, make()
        return self.t


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
    with _lib.launch(frames.device):
        rc = _lib.load().pwr_sfr_com(ptr(frames), Hf, Wf, ptr(com), B, stream_ptr(frames.device))
    check(rc, "pwr_sfr_com")
    return com


def load_bb(raw, boxes):
    """HAND17 `process_mode='bb'` loader on the GPU (datasets.py:974-996): raw [B,Hf,Wf] uint16 CUDA frames,
    boxes [B,4] = (ustart, vstart, du, dv) -> float32 frames with everything outside the box and everything
    deeper than mean + 100 mm (two-pass mean) zeroed.  Continue as the reference does for this mode
    (datasets.py:203-214): `build_sfr(frames, None, cube, test_only=True, frame_f64=True, fx=.., fy=..)`."""
    require_cuda(raw)
    if raw.dtype != torch.uint16 or raw.dim() != 3:
        raise _lib.PwrError("raw must be a [B, Hf, Wf] uint16 tensor")
    raw = raw.contiguous()
    B, Hf, Wf = raw.shape
    boxes = _f64(boxes, raw.device)
    if tuple(boxes.shape) != (B, 4):
        raise _lib.PwrError("boxes must be [B, 4] = (ustart, vstart, du, dv)")
    out = torch.empty(B, Hf, Wf, device=raw.device, dtype=torch.float32)
    with _lib.launch(raw.device):
        rc = _lib.load().pwr_sfr_bb_filter(ptr(raw), Hf, Wf, ptr(boxes), ptr(out), B, stream_ptr(raw.device))
    check(rc, "pwr_sfr_bb_filter")
    return out


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


def window_size(com, cube, fx, fy, Hf, Wf, frame_format="f32", augment=False):
    """(win_h, win_w) large enough for every sample of the batch: the largest crop box side
    max(int(cube/z*fx + cube/z*fy), 2) rounded to even (datasets.py:306-309, utils.py:169), clipped to the frame,
    plus 16-byte alignment slack per row (and the +-5 px centre shift of the augmented branch)."""
    com = np.asarray(com.cpu() if isinstance(com, torch.Tensor) else com, dtype=np.float64)
    cube = np.broadcast_to(np.asarray(cube.cpu() if isinstance(cube, torch.Tensor) else cube, dtype=np.float64),
                           (len(com),))
    with np.errstate(all="ignore"):
        side = cube / com[:, 2] * fx + cube / com[:, 2] * fy
    side = side[np.isfinite(side) & (np.abs(side) < 1e6)]
    box = max(int(side.max()), 2) if side.size else 2
    box = 2 * (box // 2) + (12 if augment else 0)
    per16 = 4 if frame_format == "f32" else 8
    win_w = min(-(-(box + per16) // per16) * per16, -(-Wf // per16) * per16)
    return max(min(box, Hf), 1), max(win_w, per16)


def fetch_windows(frames, com, cube, *, fx, fy, frame_format="f32", prefilter=None, augment=None, win_hw=None,
                  out=None):
    """Host -> device feed (pwr_sfr_fetch): `frames` [B,Hf,Wf] in PINNED host memory (torch `pin_memory()`;
    device-addressable under UVA) or on the device; `com` [B,3], `cube` [B] float64 (device tensors, or host arrays
    that are copied first: they are a few KB).  Copies, per sample, only the crop box AND hand rectangle the builder will read into a
    compact window buffer and returns FrameWindows for `build_sfr(frames=...)`.  Enqueued on the current stream.
    `win_hw`: window size; default from `window_size` (needs host copies of com / cube -> pass it explicitly
    when they only exist on the device).  `out`: a previous FrameWindows of the same shape to overwrite."""
    if not isinstance(frames, torch.Tensor) or frames.dim() != 3 or frames.dtype != _FRAME_DTYPES.get(frame_format):
        raise _lib.PwrError("frames must be a [B, Hf, Wf] %s tensor for frame_format=%r" % (
            _FRAME_DTYPES.get(frame_format), frame_format))
    if not frames.is_cuda and not frames.is_pinned():
        raise _lib.PwrError("fetch_windows reads the frames from the GPU: host frames must be pinned (pin_memory())")
    if not frames.is_contiguous():
        raise _lib.PwrError("frames must be contiguous")
    if not torch.cuda.is_available():
        raise _lib.PwrError("fetch_windows needs a CUDA device (there is no CPU fallback)")
    dev = frames.device if frames.is_cuda else (com.device if isinstance(com, torch.Tensor) and com.is_cuda
                                                else torch.device("cuda", torch.cuda.current_device()))
    com, cube = _f64(com, dev), _f64(cube, dev)           # a few KB: host arrays are copied, device tensors pass through
    B, Hf, Wf = frames.shape
    if cube.dim() == 0:
        cube = cube.expand(B).contiguous()
    if tuple(com.shape) != (B, 3) or tuple(cube.shape) != (B,):
        raise _lib.PwrError("com must be [B,3] and cube [B]")
    lib = _lib.load()
    aug_dev = None
    if augment is not None:
        aug_dev = augment if (isinstance(augment, torch.Tensor) and augment.is_cuda and augment.shape == (B, 8)) \
            else _aug_device_params(augment, B, dev)
    if win_hw is None:
        win_hw = window_size(com, cube, fx, fy, Hf, Wf, frame_format, augment is not None)
    win_h, win_w = int(win_hw[0]), int(win_hw[1])
    if out is not None and tuple(out.windows.shape) == (B, win_h, win_w) and out.windows.dtype == frames.dtype:
        fw = out
    else:
        fw = FrameWindows(torch.empty(B, win_h, win_w, device=dev, dtype=frames.dtype),
                          torch.empty(B, 4, device=dev, dtype=torch.int32), (Hf, Wf),
                          torch.zeros(1, device=dev, dtype=torch.int64), torch.zeros(1, device=dev, dtype=torch.int32))
    pf = (-1.0, 0.0, 0.0) if prefilter is None else (float(prefilter[0]), 2.0 * prefilter[1], 2.0 * prefilter[2])
    com = com.to(torch.float64).contiguous()
    cube = cube.to(torch.float64).contiguous()
    with _lib.launch(dev, "pwr_sfr_fetch"):
        rc = lib.pwr_sfr_fetch(ptr(frames), _lib.FRAME_FORMATS[frame_format], Hf, Wf, ptr(com), ptr(cube), ptr(aug_dev),
                               float(fx), float(fy), pf[0], pf[1], pf[2], ptr(fw.windows), win_h, win_w, ptr(fw.extent),
                               ptr(fw.fetched_bytes), ptr(fw.status), B, stream_ptr(dev))
    check(rc, "pwr_sfr_fetch")
    return fw


def select_valid(batch):
    """The reference RAISES on a sample with an empty crop, an out-of-range heat-map index, NaN or sum(mask) < 10
    (datasets.py:323-327, 362-365, 385-390), so such samples never reach the loss; here they come back with
    valid == 0 and all-zero targets.  This drops them (device-side boolean indexing: one sync), which is what the
    reference's DataLoader effectively does.  Returns the batch unchanged when every sample is valid."""
    keep = batch.valid.bool()
    if bool(keep.all()):
        return batch
    return type(batch)(*[None if t is None else t[keep] for t in batch])


# the one heat-map configuration the kernels implement (train.py:28-30 defaults; constants in sfr.cu / decoder.cu)
KERNEL_SIZE, SIGMOID, LABEL_SIZE, IMAGE_SIZE = 7, 1.5, 64, 128


def build_sfr(frames, com, cube, uvd=None, *, fx, fy, frame_f64=False, test_only=False, frame_format="f32",
              prefilter=None, augment=None, targets="dense", arena=None, kernel_size=KERNEL_SIZE, sigmoid=SIGMOID,
              label_size=LABEL_SIZE, image_size=IMAGE_SIZE):
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
    `frames` may also be the FrameWindows of `fetch_windows` (window mode: only the part of each frame
    the builder reads was brought to the device); results are bit-identical to whole frames.
    `arena`: an SfrArena to re-use outputs and scratch across calls (no allocations on the launch path).
    `kernel_size` / `sigmoid` / `label_size` / `image_size` exist to refuse configurations of the reference's
    flags (train.py:28-30, datasets.py:47-49) other than the 7 / 1.5 / 64 / 128 the kernels implement.

    Returns SFRBatch (train) or SFRTestBatch (test_only), all float32 except
    `valid` (uint8)."""
    if (kernel_size, float(sigmoid), label_size, image_size) != (KERNEL_SIZE, SIGMOID, LABEL_SIZE, IMAGE_SIZE):
        raise _lib.PwrError("the SFR kernels implement kernel_size=7, sigmoid=1.5, label_size=64, image_size=128 only "
                            "(got %r, %r, %r, %r); other settings would silently build different targets" % (
                                kernel_size, sigmoid, label_size, image_size))
    win = None
    if isinstance(frames, FrameWindows):
        win, frames = frames, frames.windows
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
    win_extent, win_h, win_w = None, 0, 0
    if win is not None:
        if com is None:
            raise _lib.PwrError("window mode needs the hand centre (the centre-of-mass fallback reads whole frames)")
        win_extent, win_h, win_w = win.extent, Hf, Wf
        Hf, Wf = win.frame_hw
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
    J = 0 if (test_only or uvd is None) else int(uvd.shape[1])
    if targets not in ("dense", "sparse", "both"):
        raise _lib.PwrError("targets must be 'dense', 'sparse' or 'both'")
    ws_bytes = int(lib.pwr_sfr_workspace_bytes(B, J))

    def allocate():
        t = dict(img=torch.empty(B, 1, 128, 128, **f32), label_img=torch.empty(B, 1, 64, 64, **f32),
                 mask=torch.empty(B, 1, 64, 64, **f32), box_size=torch.empty(B, **f32),
                 cube_size=torch.empty(B, **f32), com_out=torch.empty(B, 3, **f32),
                 valid=torch.empty(B, device=dev, dtype=torch.uint8),
                 workspace=torch.empty(max(ws_bytes, 16), device=dev, dtype=torch.uint8))    # scratch, no init needed
        if J:
            t["uvd_norm"] = torch.empty(B, J, 3, **f32)
            t["heatmaps"] = torch.empty(B, J, 64, 64, **f32) if targets != "sparse" else None
            t["dmap"] = torch.empty(B, J, 64, 64, **f32) if targets != "sparse" else None
            t["taps"] = torch.empty(B, J, 64, device=dev, dtype=torch.uint8) if targets != "dense" else None
        return t

    t = allocate() if arena is None else arena.get((str(dev), B, J, targets), allocate)
    img, label_img, mask, box_size, cube_size = t["img"], t["label_img"], t["mask"], t["box_size"], t["cube_size"]
    com_out, valid, workspace = t["com_out"], t["valid"], t["workspace"]
    s = stream_ptr(dev)
    if test_only:
        if augment is not None:
            raise _lib.PwrError("you can not transform the test data")     # datasets.py:64-65
        with _lib.launch(dev, "pwr_sfr_crop"):
            rc = lib.pwr_sfr_crop(ptr(frames), fmt, Hf, Wf, ptr(com), ptr(cube), float(fx), float(fy), int(frame_f64),
                                  pf[0], pf[1], pf[2], ptr(img), ptr(label_img), ptr(mask), ptr(box_size), ptr(cube_size), ptr(com_out),
                                  ptr(valid), ptr(workspace), ws_bytes, ptr(win_extent), win_h, win_w, B, s)
        check(rc, "pwr_sfr_crop")
        return SFRTestBatch(img, label_img, mask, box_size, cube_size, com_out, valid)
    if uvd is None:
        raise _lib.PwrError("train-mode SFR needs joint annotations (uvd)")
    uvd = _f64(uvd, dev)
    if uvd.dim() != 3 or uvd.shape[0] != B or uvd.shape[2] != 3:
        raise _lib.PwrError("uvd must be [B, J, 3]")
    uvd_norm, heatmaps, dmap, taps = t["uvd_norm"], t["heatmaps"], t["dmap"], t["taps"]
    aug_dev = None
    if augment is not None:
        aug_dev = augment if (isinstance(augment, torch.Tensor) and augment.is_cuda and tuple(augment.shape) == (B, 8)) \
            else _aug_device_params(augment, B, dev)
    with _lib.launch(dev, "pwr_sfr_build"):
        rc = lib.pwr_sfr_build(ptr(frames), fmt, Hf, Wf, ptr(com), ptr(cube), ptr(uvd), ptr(aug_dev), float(fx), float(fy),
                               int(frame_f64), pf[0], pf[1], pf[2], ptr(img), ptr(label_img), ptr(mask), ptr(box_size), ptr(cube_size),
                               ptr(com_out), ptr(uvd_norm), ptr(heatmaps), ptr(dmap), ptr(taps), ptr(valid),
                               ptr(workspace), ws_bytes, ptr(win_extent), win_h, win_w, B, J, s)
    check(rc, "pwr_sfr_build")
    return SFRBatch(img, label_img, mask, box_size, cube_size, com_out, uvd_norm, heatmaps, dmap, valid, taps)
