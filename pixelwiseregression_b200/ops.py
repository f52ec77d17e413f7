"""Host-side operators over libpwr_b200.so: raw launches and autograd glue.

PyTorch is plumbing here (device memory, streams, autograd graph); every
function below enqueues hand-written sm_100a kernels on the current stream
through the C ABI of include/pwr.h.  There is no eager / CPU fallback.

Reference lines replaced: model.py:79-97, 123-132, 151 (decoder forward),
what autograd derives from them plus train.py:197-207 (backward + loss),
utils.py:332-337 + datasets.py:100-111 (recover_uvd / uvd2xyz).
"""
from collections import namedtuple

import torch

from . import _lib
from ._lib import METHODS, as_f32, check, ptr, require_cuda, stream_ptr

HW = 64
MAP = HW * HW


def _check_maps(z, label_img, mask):
    if z.dim() != 4 or z.shape[2] != HW or z.shape[3] != HW:
        raise _lib.PwrError("decoder maps must be [B, J, 64, 64], got %s" % (tuple(z.shape),))
    B = z.shape[0]
    for name, t in (("label_img", label_img), ("mask", mask)):
        if t is not None and tuple(t.shape) != (B, 1, HW, HW):
            raise _lib.PwrError("%s must be [B, 1, 64, 64], got %s" % (name, tuple(t.shape)))


SparseTargets = namedtuple("SparseTargets", ["taps", "uvd"])
SparseTargets.__doc__ = """Compact targets of a batch: `taps` [B,J,64] uint8 (pwr_joint_taps records written by
sfr.build_sfr(targets="sparse")) from which the loss kernels evaluate the heat-map and depth-map targets on
the fly, and the normalised joint coordinates `uvd` [B,J,3]."""


def _unpack_targets(targets, B, J):
    """(heat_gt, dmap_gt, uvd_gt) dense maps, or SparseTargets -> (heat, dmap, uvd, taps) for the C ABI."""
    if isinstance(targets, SparseTargets):
        taps = targets.taps
        require_cuda(taps)
        if taps.dtype != torch.uint8 or tuple(taps.shape) != (B, J, 64) or not taps.is_contiguous():
            raise _lib.PwrError("sparse targets must be a contiguous [B, J, 64] uint8 tensor of pwr_joint_taps")
        return None, None, as_f32(targets.uvd), taps
    heat_gt, dmap_gt, uvd_gt = (as_f32(t) for t in targets)
    return heat_gt, dmap_gt, uvd_gt, None


def _conv_maps(z, D, method):
    """z and D as contiguous tensors of ONE kernel-supported element type: float32, or the
    float16 / bfloat16 conv outputs of autocast, read by the kernels as they are (the arithmetic
    is float32 either way, which is what autocast does for these ops)."""
    dt = z.dtype
    if dt not in _lib.MAP_DTYPES or method == "given" or (D is not None and D.dtype != dt):
        dt = torch.float32
    z = z.to(dt).contiguous()
    D = D.to(dt).contiguous() if D is not None else None
    return z, D, _lib.MAP_DTYPES[dt]


def decoder_forward_raw(z, w, D, label_img, mask, method="softmax", store_heat=True, want_stats=True,
                        targets=None):
    """One launch of pwr_decoder_fwd.  Returns (H or None, uvd, stats or None,
    loss_partial or None).  z, D: float32 / float16 / bfloat16 CUDA tensors (same dtype);
    everything else float32; H, uvd are float32."""
    require_cuda(z, w, D, label_img, mask)
    lib = _lib.load()
    z, D, map_dtype = _conv_maps(z, D, method)
    label_img = as_f32(label_img)
    mask = as_f32(mask)
    _check_maps(z, label_img if D is not None else None, mask if D is not None else None)
    B, J = z.shape[0], z.shape[1]
    wv = as_f32(w).reshape(-1) if w is not None else None
    H = torch.empty(z.shape, device=z.device, dtype=torch.float32) if store_heat else None
    uvd = torch.empty(B, J, 3, device=z.device, dtype=torch.float32)
    stats = torch.empty(B, J, 4, device=z.device, dtype=torch.float32) if want_stats else None
    heat_gt = dmap_gt = uvd_gt = taps = loss_partial = None
    if targets is not None:
        heat_gt, dmap_gt, uvd_gt, taps = _unpack_targets(targets, B, J)
        loss_partial = torch.empty(B, J, 3, device=z.device, dtype=torch.float32)
    with _lib.launch(z.device, "pwr_decoder_fwd"):
        rc = lib.pwr_decoder_fwd(ptr(z), ptr(wv), ptr(D), ptr(label_img), ptr(mask), ptr(heat_gt), ptr(dmap_gt),
                                 ptr(uvd_gt), ptr(taps), ptr(H), ptr(uvd), ptr(stats), ptr(loss_partial), B, J,
                                 METHODS[method], map_dtype, stream_ptr(z.device))
    check(rc, "pwr_decoder_fwd")
    return H, uvd, stats, loss_partial


def decoder_backward_raw(z, w, D, label_img, mask, stats, uvd, g_uvd=None, gH_up=None, gD_up=None,
                         method="softmax", targets=None, alpha=1.0, lambda_h=1.0, lambda_d=0.01,
                         loss_scale=1.0, loss_scale_dev=None, n_mean=0, want_loss=False, want_gz=True,
                         want_gD=True):
    """One launch of pwr_decoder_bwd (targets None) or pwr_decoder_bwd_loss.
    Returns (gz, gD, gw_partial [B,J] or None, loss_partial [B,J,3] or None); gz, gD (and
    gD_up) carry the element type of z, D."""
    require_cuda(z, w, D, label_img, mask, stats, uvd, g_uvd, gH_up, gD_up)
    lib = _lib.load()
    z, D, map_dtype = _conv_maps(z, D, method)
    B, J = z.shape[0], z.shape[1]
    wv = as_f32(w).reshape(-1) if w is not None else None
    gz = torch.empty_like(z) if want_gz else None
    gD = torch.empty_like(z) if (want_gD and D is not None) else None
    gw_partial = torch.empty(B, J, device=z.device, dtype=torch.float32) if method == "softmax" else None
    g_uvd = as_f32(g_uvd)
    gH_up = as_f32(gH_up)
    gD_up = gD_up.to(z.dtype).contiguous() if gD_up is not None else None
    # every converted tensor is bound to a local that outlives the launch: a temporary inside the
    # argument list would be freed (and its block possibly re-used by the next temporary) before the kernel runs
    label_img, mask, stats, uvd = as_f32(label_img), as_f32(mask), as_f32(stats), as_f32(uvd)
    scale_dev = as_f32(loss_scale_dev)
    s = stream_ptr(z.device)
    with _lib.launch(z.device, "pwr_decoder_bwd" if targets is None else "pwr_decoder_bwd_loss"):
        if targets is None:
            rc = lib.pwr_decoder_bwd(ptr(z), ptr(wv), ptr(D), ptr(label_img), ptr(mask),
                                     ptr(stats), ptr(uvd), ptr(g_uvd), ptr(gH_up), ptr(gD_up), ptr(gz), ptr(gD),
                                     ptr(gw_partial), B, J, METHODS[method], map_dtype, s)
            check(rc, "pwr_decoder_bwd")
            return gz, gD, gw_partial, None
        heat_gt, dmap_gt, uvd_gt, taps = _unpack_targets(targets, B, J)
        loss_partial = torch.empty(B, J, 3, device=z.device, dtype=torch.float32) if want_loss else None
        rc = lib.pwr_decoder_bwd_loss(ptr(z), ptr(wv), ptr(D), ptr(label_img), ptr(mask),
                                      ptr(stats), ptr(uvd), ptr(g_uvd), ptr(gH_up), ptr(gD_up), ptr(heat_gt),
                                      ptr(dmap_gt), ptr(uvd_gt), ptr(taps), float(alpha), float(lambda_h),
                                      float(lambda_d),
                                      float(loss_scale), ptr(scale_dev), int(n_mean), ptr(gz),
                                      ptr(gD), ptr(gw_partial),
                                      ptr(loss_partial), B, J, METHODS[method], map_dtype, s)
    check(rc, "pwr_decoder_bwd_loss")
    return gz, gD, gw_partial, loss_partial


def decoder_fused_raw(z, w, D, label_img, mask, targets, method="softmax", alpha=1.0, lambda_h=1.0, lambda_d=0.01,
                      loss_scale=1.0, loss_scale_dev=None, n_mean=0, store_heat=True, want_grads=True,
                      want_loss=True):
    """One launch of pwr_decoder_fwd_bwd_loss: last-stage forward + stage loss + backward with z, D and
    the targets read once.  Returns (H or None, uvd, gz, gD, gw_partial or None, loss_partial [B,J,3] or None).
    `want_loss=False` skips the logged loss terms; with alpha = 1 the target maps are then not read at all."""
    require_cuda(z, w, D, label_img, mask)
    lib = _lib.load()
    z, D, map_dtype = _conv_maps(z, D, method)
    label_img, mask = as_f32(label_img), as_f32(mask)
    _check_maps(z, label_img, mask)
    B, J = z.shape[0], z.shape[1]
    wv = as_f32(w).reshape(-1) if w is not None else None
    heat_gt, dmap_gt, uvd_gt, taps = _unpack_targets(targets, B, J)
    f32 = dict(device=z.device, dtype=torch.float32)
    H = torch.empty(z.shape, **f32) if store_heat else None
    uvd = torch.empty(B, J, 3, **f32)
    gz = torch.empty_like(z) if want_grads else None
    gD = torch.empty_like(z) if want_grads else None
    gw_partial = torch.empty(B, J, **f32) if (want_grads and method == "softmax") else None
    loss_partial = torch.empty(B, J, 3, **f32) if want_loss else None
    loss_scale_dev = as_f32(loss_scale_dev)
    with _lib.launch(z.device, "pwr_decoder_fwd_bwd_loss"):
        rc = lib.pwr_decoder_fwd_bwd_loss(ptr(z), ptr(wv), ptr(D), ptr(label_img), ptr(mask), ptr(heat_gt), ptr(dmap_gt),
                                          ptr(uvd_gt), ptr(taps), float(alpha), float(lambda_h), float(lambda_d),
                                          float(loss_scale), ptr(loss_scale_dev), int(n_mean), ptr(H), ptr(uvd),
                                          ptr(gz), ptr(gD), ptr(gw_partial), ptr(loss_partial), B, J, METHODS[method],
                                          map_dtype, stream_ptr(z.device))
    check(rc, "pwr_decoder_fwd_bwd_loss")
    return H, uvd, gz, gD, gw_partial, loss_partial


def reduce_partials(partial):
    """[B, J] or [B, J, C] per-(sample, joint) partials -> [J] / [J, C] batch sums
    (deterministic tree, pwr_reduce_partials)."""
    require_cuda(partial)
    partial = as_f32(partial)
    B, J = partial.shape[0], partial.shape[1]
    C = partial.shape[2] if partial.dim() == 3 else 1
    out = torch.empty((J, C) if partial.dim() == 3 else (J,), device=partial.device, dtype=torch.float32)
    with _lib.launch(partial.device):
        rc = _lib.load().pwr_reduce_partials(ptr(partial), ptr(out), B, J, C,
                                             stream_ptr(partial.device))
    check(rc, "pwr_reduce_partials")
    return out


def scale_inplace_(x, scale, x2=None, small=None):
    """x *= scale (and x2, and the small float32 vector `small`), with `scale` a 0-dim CUDA tensor: no host sync,
    one launch for all three, an early exit inside the kernel when scale == 1."""
    require_cuda(x, scale, x2, small)
    scale = as_f32(scale)
    if x2 is not None and (x2.dtype != x.dtype or x2.numel() != x.numel()):
        raise _lib.PwrError("scale_inplace_: x and x2 must have the same element type and size")
    if small is not None and (small.dtype != torch.float32 or not small.is_contiguous() or small.numel() > 256):
        raise _lib.PwrError("scale_inplace_: `small` must be a contiguous float32 tensor of at most 256 elements")
    with _lib.launch(x.device):
        rc = _lib.load().pwr_scale_inplace(ptr(x), ptr(x2), ptr(small), 0 if small is None else small.numel(), ptr(scale),
                                           x.numel(), _lib.MAP_DTYPES[x.dtype], stream_ptr(x.device))
    check(rc, "pwr_scale_inplace")
    return x


def stage_loss(loss_partial, lambda_h, lambda_d, alpha, n_mean=0, gw_partial=None):
    """train.py:197-205 from per-(b,j) sums of squares in one launch (pwr_stage_loss):
    returns a [4] tensor (heatmap_loss, depthmap_loss, uvd_loss, combined loss); with `gw_partial` [B,J] also
    dL/dw [J] summed over the batch by the same launch -> (out4, gw)."""
    require_cuda(loss_partial, gw_partial)
    loss_partial = as_f32(loss_partial)
    B, J = loss_partial.shape[0], loss_partial.shape[1]
    out = torch.empty(4 + (J if gw_partial is not None else 0), device=loss_partial.device, dtype=torch.float32)
    gw = out[4:] if gw_partial is not None else None
    gw_partial = as_f32(gw_partial)
    with _lib.launch(loss_partial.device):
        rc = _lib.load().pwr_stage_loss(ptr(loss_partial), B, J, float(lambda_h), float(lambda_d), float(alpha),
                                        int(n_mean), ptr(out), ptr(gw_partial), ptr(gw), stream_ptr(loss_partial.device))
    check(rc, "pwr_stage_loss")
    return (out[:4], gw) if gw_partial is not None else out


def stage_loss_from_partials(loss_partial, lambda_h, lambda_d, n_mean=0):
    """[3] tensor (heatmap_loss, depthmap_loss, uvd_loss), train.py:197-199."""
    return stage_loss(loss_partial, lambda_h, lambda_d, 1.0, n_mean)[:3]


def _make_targets(heat_gt, dmap_gt, uvd_gt):
    """The autograd functions receive plain tensors: a uint8 `heat_gt` IS the sparse taps tensor."""
    if heat_gt is None:
        return None
    if heat_gt.dtype == torch.uint8:
        return SparseTargets(heat_gt, uvd_gt)
    return (heat_gt, dmap_gt, uvd_gt)


class DecoderFunction(torch.autograd.Function):
    """Differentiable fused decoder: (z, w, D, label_img, mask) -> (heatmaps,
    depthmaps, uvd), model.py:147-151 without the two conv stacks.  `depthmaps`
    is D itself (model.py:132 returns the conv output unchanged); routing it
    through this node lets the backward kernel fold the dense upstream gradient
    on the depth maps (next stage's conv, loss) into its single pass.

    With `targets` the stage loss of train.py:197-205 rides along: its value
    comes out of the forward kernel (4th/5th outputs: combined loss, [3] terms)
    and its gradient is added inside the backward kernel, scaled by the
    upstream gradient on the combined loss (read on the device)."""

    @staticmethod
    def forward(ctx, z, w, D, label_img, mask, method, heat_gt=None, dmap_gt=None, uvd_gt=None, alpha=1.0,
                lambda_h=1.0, lambda_d=0.01, n_mean=0):
        ctx.set_materialize_grads(False)
        targets = _make_targets(heat_gt, dmap_gt, uvd_gt)
        H, uvd, stats, loss_partial = decoder_forward_raw(z, w, D, label_img, mask, method, targets=targets)
        ctx.method = method
        ctx.in_dtypes = (z.dtype, D.dtype)
        ctx.loss_cfg = (alpha, lambda_h, lambda_d, n_mean) if targets is not None else None
        ctx.save_for_backward(z, w, D, label_img, mask, stats, uvd, heat_gt, dmap_gt, uvd_gt)
        # heat maps and coordinates are float32 even for half-precision logits, as under autocast
        outs = (H, D.view_as(D), uvd)
        if targets is None:
            return outs
        out4 = stage_loss(loss_partial, lambda_h, lambda_d, alpha, n_mean)
        terms, total = out4[:3], out4[3]
        ctx.mark_non_differentiable(terms)
        return outs + (total, terms)

    @staticmethod
    def backward(ctx, gH, gD_up, g_uvd, g_total=None, g_terms=None):
        z, w, D, label_img, mask, stats, uvd, heat_gt, dmap_gt, uvd_gt = ctx.saved_tensors
        if ctx.loss_cfg is not None and g_total is not None:
            alpha, lambda_h, lambda_d, n_mean = ctx.loss_cfg
            gz, gD, gw_partial, _ = decoder_backward_raw(
                z, w, D, label_img, mask, stats, uvd, g_uvd, gH, gD_up, ctx.method,
                targets=_make_targets(heat_gt, dmap_gt, uvd_gt), alpha=alpha, lambda_h=lambda_h, lambda_d=lambda_d,
                loss_scale_dev=g_total, n_mean=n_mean)
        else:
            gz, gD, gw_partial, _ = decoder_backward_raw(z, w, D, label_img, mask, stats, uvd, g_uvd, gH, gD_up,
                                                         ctx.method)
        gw = None
        if w is not None and ctx.needs_input_grad[1]:
            gw = reduce_partials(gw_partial).view_as(w).to(w.dtype)
        return (gz.to(ctx.in_dtypes[0]), gw, gD.to(ctx.in_dtypes[1])) + (None,) * 10


def fused_decoder(z, w, D, label_img, mask, method="softmax"):
    """Autograd-connected fused decoder.  Returns (heatmaps, depthmaps, uvd)."""
    if torch.is_grad_enabled() and (z.requires_grad or D.requires_grad or (w is not None and w.requires_grad)):
        return DecoderFunction.apply(z, w, D, label_img, mask, method)
    H, uvd, _, _ = decoder_forward_raw(z, w, D, label_img, mask, method, want_stats=False)
    return H, D, uvd


def fused_decoder_with_loss(z, w, D, label_img, mask, heat_gt, dmap_gt, uvd_gt, method="softmax", alpha=1.0,
                            lambda_h=1.0, lambda_d=0.01, n_mean=0):
    """Inner-stage variant: returns (heatmaps, depthmaps, uvd, stage_loss, terms[3])
    with heatmaps/depthmaps/uvd/stage_loss all differentiable.  `n_mean`: see fused_decoder_loss."""
    return DecoderFunction.apply(z, w, D, label_img, mask, method, heat_gt, dmap_gt, uvd_gt, float(alpha),
                                 float(lambda_h), float(lambda_d), int(n_mean))


class PlaneFunction(torch.autograd.Function):
    """PlaneRegression.forward called on its own (model.py:79-97): no depth branch."""

    @staticmethod
    def forward(ctx, z, w, method):
        ctx.set_materialize_grads(False)
        H, uvd, stats, _ = decoder_forward_raw(z, w, None, None, None, method)
        ctx.method = method
        ctx.save_for_backward(z, w, stats, uvd)
        return H, uvd[:, :, :2].contiguous()

    @staticmethod
    def backward(ctx, gH, g_uv):
        z, w, stats, uvd = ctx.saved_tensors
        g_uvd = None
        if g_uv is not None:
            g_uvd = torch.zeros_like(uvd)
            g_uvd[:, :, :2] = g_uv
        gz, _, gw_partial, _ = decoder_backward_raw(z, w, None, None, None, stats, uvd, g_uvd, gH, None, ctx.method)
        gw = None
        if w is not None and ctx.needs_input_grad[1]:
            gw = reduce_partials(gw_partial).view_as(w).to(w.dtype)
        return gz.to(z.dtype), gw, None


class DepthFunction(torch.autograd.Function):
    """DepthRegression.forward called on its own with caller-supplied heat maps
    (model.py:123-132): the heat maps take the place of the logits
    (PWR_METHOD_GIVEN), gradients flow to both the heat maps and D."""

    @staticmethod
    def forward(ctx, D, heatmaps, label_img, mask):
        ctx.set_materialize_grads(False)
        _, uvd, stats, _ = decoder_forward_raw(heatmaps, None, D, label_img, mask, "given", store_heat=False)
        ctx.save_for_backward(D, heatmaps, label_img, mask, stats, uvd)
        return uvd[:, :, 2:].contiguous().to(D.dtype)

    @staticmethod
    def backward(ctx, g_d):
        D, heatmaps, label_img, mask, stats, uvd = ctx.saved_tensors
        g_uvd = torch.zeros_like(uvd)
        if g_d is not None:
            g_uvd[:, :, 2:] = g_d
        gH, gD, _, _ = decoder_backward_raw(heatmaps, None, D, label_img, mask, stats, uvd, g_uvd, None, None,
                                            "given")
        return gD.to(D.dtype), gH.to(heatmaps.dtype), None, None


# Loss scale built into the eagerly computed float16 gradients of the last stage (see DecoderLossFunction.forward)
EAGER_FP16_SCALE = 65536.0

# False sends ops.fused_decoder_loss through the two-kernel route (pwr_decoder_fwd, then pwr_decoder_bwd_loss);
# bench.py flips it to time both, tests to compare them.
ONE_PASS_LAST_STAGE = True


class DecoderLossFunction(torch.autograd.Function):
    """Last-stage decoder fused with the stage loss, train.py:197-207.

    forward runs the forward kernel and then, when gradients are needed, the
    fused backward+loss kernel straight away (unit upstream): the loss value
    and the gradients come out of one pass over (z, D, heat_gt, dmap_gt).  The
    autograd backward only rescales the stored gradients by the upstream scalar
    (a no-op launch when it is 1, e.g. without a GradScaler).  Only `total` is
    differentiable; the returned heat maps / uvd / loss terms are detached
    (use DecoderFunction when later layers consume the maps)."""

    @staticmethod
    def forward(ctx, z, w, D, label_img, mask, heat_gt, dmap_gt, uvd_gt, method, alpha, lambda_h, lambda_d,
                store_heat, n_mean):
        ctx.set_materialize_grads(False)       # no zero-filled gradient tensors for the detached outputs
        need_grad = any(ctx.needs_input_grad[:3])
        targets = _make_targets(heat_gt, dmap_gt, uvd_gt)
        # float16 conv outputs (autocast + GradScaler, train.py:170-189): the eager gradients are stored in
        # float16, and at unit loss scale their typical size 2/(B*J) * p * err (1e-6 .. 1e-10) is subnormal or
        # flushes to zero.  The reference scales the loss BEFORE backward, so its float16 gradients are 2^16
        # larger; do the same here: the kernel produces them pre-scaled by EAGER_FP16_SCALE and backward()
        # multiplies by g_total / EAGER_FP16_SCALE (exactly 1 for GradScaler's initial scale of 2^16).
        pre = EAGER_FP16_SCALE if (need_grad and z.dtype == torch.float16 and method != "given") else 1.0
        if need_grad and D is not None and method != "given" and ONE_PASS_LAST_STAGE:
            # forward + loss + backward in one visit of (z, D, targets)
            H, uvd, gz, gD, gw_partial, loss_partial = decoder_fused_raw(
                z, w, D, label_img, mask, targets, method, alpha, lambda_h, lambda_d, loss_scale=pre,
                n_mean=n_mean, store_heat=store_heat)
        else:
            H, uvd, stats, _ = decoder_forward_raw(z, w, D, label_img, mask, method, store_heat=store_heat)
            gz, gD, gw_partial, loss_partial = decoder_backward_raw(
                z, w, D, label_img, mask, stats, uvd, None, None, None, method, targets, alpha, lambda_h, lambda_d,
                loss_scale=pre, n_mean=n_mean, want_loss=True, want_gz=need_grad, want_gD=need_grad)
        if need_grad and w is not None and gw_partial is not None:
            out4, gw = stage_loss(loss_partial, lambda_h, lambda_d, alpha, n_mean, gw_partial)    # loss + dL/dw: one launch
            gw = gw.view_as(w)
        else:
            out4, gw = stage_loss(loss_partial, lambda_h, lambda_d, alpha, n_mean), None
        terms, total = out4[:3], out4[3]
        # the eager gradients live on the node until its (single) backward consumes them
        ctx.grads = (gz, gD, gw)
        ctx.pre = pre
        outs = (total, terms, uvd) + ((H,) if store_heat else ())
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g_total, *unused):
        if ctx.grads is None:
            raise _lib.PwrError("fused_decoder_loss: backward through the same node twice (the eager gradients are "
                                "rescaled in place and handed over once; retain_graph is not supported here)")
        gz, gD, gw = ctx.grads
        ctx.grads = None
        if gz is None:
            raise _lib.PwrError("fused_decoder_loss: forward ran without gradient tracking")
        if g_total is None:
            return (None,) * 14
        if ctx.pre != 1.0:
            g_total = g_total.float() * (1.0 / ctx.pre)
        scale_inplace_(gz, g_total, gD, gw)                 # one launch; a no-op when the upstream gradient is 1
        return (gz, gw, gD) + (None,) * 11


def fused_decoder_loss(z, w, D, label_img, mask, heat_gt, dmap_gt, uvd_gt, method="softmax", alpha=1.0,
                       lambda_h=1.0, lambda_d=0.01, store_heat=True, n_mean=0):
    """Returns (total_loss, loss_terms[3] = (heatmap, depthmap, uvd), uvd[, heatmaps]).
    `heat_gt` may be the uint8 taps tensor of sfr.build_sfr(targets="sparse") (then `dmap_gt` is
    ignored / None): the targets are evaluated inside the loss kernel.
    `n_mean` = the B*J the means of train.py:197-199 run over; 0 = this call's own B*J (what DDP's
    gradient AVERAGING needs); the global B*J when per-rank losses / gradients are SUMMED across ranks
    or micro-batches."""
    return DecoderLossFunction.apply(z, w, D, label_img, mask, heat_gt, dmap_gt, uvd_gt, method, float(alpha),
                                     float(lambda_h), float(lambda_d), bool(store_heat), int(n_mean))


def recover_uvd(uvd_norm, box_size, com, cube_size, intrinsics=None):
    """utils.py:332-337 on the GPU (no .cpu() round trip, test.py:106-113).
    Returns uvd in pixels/mm; with intrinsics=(fx, fy, halfu, halfv) also xyz
    (datasets.py:100-111)."""
    require_cuda(uvd_norm, box_size, com, cube_size)
    B, J = uvd_norm.shape[0], uvd_norm.shape[1]
    uvd_norm, box_size, cube_size, com = as_f32(uvd_norm), as_f32(box_size), as_f32(cube_size), as_f32(com)
    uvd_px = torch.empty_like(uvd_norm)
    xyz = torch.empty_like(uvd_norm) if intrinsics is not None else None
    fx, fy, hu, hv = intrinsics if intrinsics is not None else (1.0, 1.0, 0.0, 0.0)
    with _lib.launch(uvd_norm.device):
        rc = _lib.load().pwr_recover_uvd(ptr(uvd_norm), ptr(box_size), ptr(cube_size),
                                         ptr(com), fx, fy, hu, hv, ptr(uvd_px), ptr(xyz), B, J,
                                         stream_ptr(uvd_norm.device))
    check(rc, "pwr_recover_uvd")
    return (uvd_px, xyz) if intrinsics is not None else uvd_px


def joint_error(uvd_pred, uvd_true, box_size, com, cube_size, intrinsics):
    """Mean joint error per sample in mm (train.py:254-276, test.py:106-113) on the GPU:
    both normalised uvd tensors go through recover_uvd and uvd2xyz inside one kernel.
    intrinsics = (fx, fy, halfu, halfv).  Returns [B] float32."""
    require_cuda(uvd_pred, uvd_true, box_size, com, cube_size)
    B, J = uvd_pred.shape[0], uvd_pred.shape[1]
    err = torch.empty(B, device=uvd_pred.device, dtype=torch.float32)
    fx, fy, hu, hv = intrinsics
    uvd_pred, uvd_true = as_f32(uvd_pred), as_f32(uvd_true)
    box_size, cube_size, com = as_f32(box_size), as_f32(cube_size), as_f32(com)
    with _lib.launch(uvd_pred.device):
        rc = _lib.load().pwr_joint_error(ptr(uvd_pred), ptr(uvd_true), ptr(box_size),
                                         ptr(cube_size), ptr(com), fx, fy, hu, hv, ptr(err), B, J,
                                         stream_ptr(uvd_pred.device))
    check(rc, "pwr_joint_error")
    return err
