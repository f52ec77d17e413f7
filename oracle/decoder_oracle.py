"""CPU oracle for the differentiable decoder, its loss and its backward
(TEST INFRASTRUCTURE, not product).

A plain-PyTorch restatement (any dtype; float64 for gradient checks) of
  model.py:79-97    PlaneRegression.forward, post-conv part  -> plane_decode
  model.py:123-132  DepthRegression.forward, post-conv part  -> depth_decode
  model.py:144-151  PredictionBlock glue (cat -> uvd)        -> decoder_forward
  train.py:197-205  per-stage losses and their combination   -> stage_losses
  utils.py:24-35    generate_com_filter                      -> com_filter
and the closed-form gradient of that composition (decoder_backward), which is
what the fused CUDA backward implements; tests check it against autograd of
decoder_forward in float64.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

Parity pinning: the reference ships no tests or golden vectors; this oracle is
pinned against the UNMODIFIED reference modules (PlaneRegression /
DepthRegression with their conv stacks replaced by Identity so that the input
IS the logit map) run in the build container -> tests/golden/decoder_*.npz
(oracle/make_golden.py).
"""
import torch


def com_filter(size=64, dtype=torch.float32, device="cpu"):
    """utils.py:24-35 -> [2, size, size]: U[i,j]=(j-size//2)/(size-1),
    V[i,j]=(i-size//2)/(size-1), evaluated in float64 then cast (model.py:68)."""
    idx = torch.arange(size, dtype=torch.float64, device=device)
    coord = (idx - size // 2) / (size - 1)
    U = coord.view(1, size).expand(size, size)
    V = coord.view(size, 1).expand(size, size)
    return torch.stack([U, V]).to(dtype).contiguous()


def plane_decode(z, w, method="softmax"):
    """model.py:79-97.  z [B,J,H,W] logits -> (heatmaps [B,J,H,W], uv [B,J,2])."""
    B, J, H, W = z.shape
    if method == "softmax":
        p = torch.softmax(w * z.reshape(B, J, -1), dim=2).view(B, J, H, W)
    else:
        p = torch.relu(z) + 1e-14
        p = p / p.sum(dim=(2, 3), keepdim=True)
    filt = com_filter(H, z.dtype, z.device)
    u = torch.sum(filt[0].view(1, 1, H, W) * p, dim=(2, 3)).unsqueeze(-1)
    v = torch.sum(filt[1].view(1, 1, H, W) * p, dim=(2, 3)).unsqueeze(-1)
    return p, torch.cat([u, v], dim=2)


def depth_decode(D, p, label_img, mask):
    """model.py:123-132.  Returns depth coordinate [B,J,1]."""
    rec = D + label_img
    mr = mask * rec
    mh = p * mask
    d = torch.sum(mh * mr, dim=(2, 3)) / (torch.sum(mh, dim=(2, 3)) + 1e-14)
    return d.unsqueeze(-1)


def decoder_forward(z, w, D, label_img, mask, method="softmax"):
    """model.py:147-151 with both conv stacks removed: returns
    (heatmaps, depthmaps(=D), uvd [B,J,3])."""
    p, uv = plane_decode(z, w, method)
    d = depth_decode(D, p, label_img, mask)
    return p, D, torch.cat([uv, d], dim=2)


def stage_losses(p, D, uvd, heat_gt, dmap_gt, uvd_gt, lambda_h=1.0, lambda_d=0.01):
    """train.py:197-199 -> (heatmap_loss, depthmap_loss, uvd_loss)."""
    lh = lambda_h * torch.mean(torch.sum((p - heat_gt) ** 2, dim=(2, 3)))
    ld = lambda_d * torch.mean(torch.sum((D - dmap_gt) ** 2, dim=(2, 3)))
    lu = torch.mean(torch.sum((uvd - uvd_gt) ** 2, dim=2))
    return lh, ld, lu


def combine_losses(losses, alpha=1.0):
    """train.py:202-205."""
    lh, ld, lu = losses
    return alpha * lu + (1 - alpha) * (lh + ld)


def decoder_backward(z, w, D, label_img, mask, g_uvd, gH_up=None, gD_up=None, method="softmax",
                     targets=None, alpha=1.0, lambda_h=1.0, lambda_d=0.01):
    """Closed-form backward of decoder_forward (+ optionally the stage loss).

    g_uvd  [B,J,3]   upstream gradient on the decoded coordinates
    gH_up  [B,J,H,W] upstream gradient on the heatmaps (next stage's conv), or None
    gD_up  [B,J,H,W] upstream gradient on the depth maps, or None
    targets (heat_gt, dmap_gt, uvd_gt): add d(combined stage loss)/d(.) with
            unit upstream, exactly as train.py:197-207 would through autograd.
    Returns (gz, gD, gw [J,1] or None)."""
    B, J, H, W = z.shape
    N = B * J
    p, _, uvd = decoder_forward(z, w, D, label_img, mask, method)
    filt = com_filter(H, z.dtype, z.device)
    U = filt[0].view(1, 1, H, W)
    V = filt[1].view(1, 1, H, W)
    g_uvd = g_uvd.clone()
    if targets is not None:
        heat_gt, dmap_gt, uvd_gt = targets
        g_uvd = g_uvd + 2 * alpha * (uvd - uvd_gt) / N
    gu = g_uvd[:, :, 0].view(B, J, 1, 1)
    gv = g_uvd[:, :, 1].view(B, J, 1, 1)
    gd = g_uvd[:, :, 2].view(B, J, 1, 1)
    den = (torch.sum(p * mask, dim=(2, 3)) + 1e-14).view(B, J, 1, 1)
    d = uvd[:, :, 2].view(B, J, 1, 1)
    rec = D + label_img
    gp = gu * U + gv * V + gd * mask * (mask * rec - d) / den
    gD = gd * p * mask * mask / den
    if targets is not None:
        gp = gp + 2 * (1 - alpha) * lambda_h * (p - heat_gt) / N
        gD = gD + 2 * (1 - alpha) * lambda_d * (D - dmap_gt) / N
    if gH_up is not None:
        gp = gp + gH_up
    if gD_up is not None:
        gD = gD + gD_up
    if method == "softmax":
        gy = p * (gp - torch.sum(gp * p, dim=(2, 3), keepdim=True))
        gz = w.view(1, J, 1, 1) * gy
        gw = torch.sum(gy * z, dim=(0, 2, 3)).view(J, 1)
    else:
        # p = r / S, r = relu(z) + 1e-14, S = sum r
        r = torch.relu(z) + 1e-14
        S = r.sum(dim=(2, 3), keepdim=True)
        gr = (gp - torch.sum(gp * p, dim=(2, 3), keepdim=True)) / S
        gz = gr * (z > 0).to(z.dtype)
        gw = None
    return gz, gD, gw
