"""CPU: the decoder/loss/backward oracle against the reference-generated golden
vectors and against autograd in float64."""
import numpy as np
import pytest
import torch

from oracle import decoder_oracle as do
from helpers import GRAD_RTOL, assert_close, load_golden

GOLDEN_SETS = ["decoder_softmax_a1", "decoder_softmax_a05_up", "decoder_sum_a05_up"]


def tensors(g, dtype):
    t = lambda k: torch.from_numpy(g[k]).to(dtype)
    return dict(z=t("z"), D=t("D"), w=t("w"), label=t("label"), mask=t("mask"), heat_gt=t("heat_gt"),
                dmap_gt=t("dmap_gt"), uvd_gt=t("uvd_gt"))


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_forward_and_loss_match_reference_golden(name):
    g = load_golden(name)
    x = tensors(g, torch.float32)
    method = str(g["method"])
    p, D, uvd = do.decoder_forward(x["z"], x["w"], x["D"], x["label"], x["mask"], method)
    assert_close("heat", p.numpy(), g["ref_heat"])
    assert_close("uvd", uvd.numpy(), g["ref_uvd"])
    lh, ld, lu = do.stage_losses(p, D, uvd, x["heat_gt"], x["dmap_gt"], x["uvd_gt"],
                                 float(g["lambda_h"]), float(g["lambda_d"]))
    total = do.combine_losses((lh, ld, lu), float(g["alpha"]))
    assert_close("losses", [lh.item(), ld.item(), lu.item(), total.item()], g["ref_losses"])


@pytest.mark.parametrize("name", GOLDEN_SETS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_closed_form_backward_matches_reference_golden(name, dtype):
    g = load_golden(name)
    x = tensors(g, dtype)
    method = str(g["method"])
    B, J = x["z"].shape[:2]
    up = "gH_up" in g
    g_uvd = torch.from_numpy(g["g_uvd_up"]).to(dtype) if up else torch.zeros(B, J, 3, dtype=dtype)
    gz, gD, gw = do.decoder_backward(
        x["z"], x["w"], x["D"], x["label"], x["mask"], g_uvd,
        torch.from_numpy(g["gH_up"]).to(dtype) if up else None,
        torch.from_numpy(g["gD_up"]).to(dtype) if up else None, method,
        targets=(x["heat_gt"], x["dmap_gt"], x["uvd_gt"]), alpha=float(g["alpha"]),
        lambda_h=float(g["lambda_h"]), lambda_d=float(g["lambda_d"]))
    assert_close("gz", gz.numpy(), g["ref_gz"], GRAD_RTOL)
    assert_close("gD", gD.numpy(), g["ref_gD"], GRAD_RTOL)
    if method == "softmax":
        assert_close("gw", gw.numpy(), g["ref_gw"], GRAD_RTOL)


@pytest.mark.parametrize("method", ["softmax", "sum"])
@pytest.mark.parametrize("alpha", [1.0, 0.3])
def test_closed_form_backward_equals_autograd_fp64(method, alpha):
    torch.manual_seed(0)
    B, J = 2, 3
    dt = torch.float64
    z = torch.randn(B, J, 64, 64, dtype=dt, requires_grad=True)
    D = torch.randn(B, J, 64, 64, dtype=dt, requires_grad=True)
    w = torch.rand(J, 1, dtype=dt).add(0.5).requires_grad_(True)
    mask = (torch.rand(B, 1, 64, 64) < 0.4).to(dt)
    label = torch.rand(B, 1, 64, 64, dtype=dt) * mask
    heat_gt = torch.rand(B, J, 64, 64, dtype=dt) * 0.01
    dmap_gt = torch.randn(B, J, 64, 64, dtype=dt)
    uvd_gt = torch.rand(B, J, 3, dtype=dt) - 0.5
    gH = torch.randn(B, J, 64, 64, dtype=dt) * 1e-3
    gDu = torch.randn(B, J, 64, 64, dtype=dt) * 1e-3
    gu = torch.randn(B, J, 3, dtype=dt)
    p, Dm, uvd = do.decoder_forward(z, w, D, label, mask, method)
    loss = do.combine_losses(do.stage_losses(p, Dm, uvd, heat_gt, dmap_gt, uvd_gt, 1.0, 0.01), alpha)
    (loss + (p * gH).sum() + (Dm * gDu).sum() + (uvd * gu).sum()).backward()
    gz, gD, gw = do.decoder_backward(z.detach(), w.detach(), D.detach(), label, mask, gu, gH, gDu, method,
                                     targets=(heat_gt, dmap_gt, uvd_gt), alpha=alpha)
    assert torch.allclose(gz, z.grad, rtol=1e-10, atol=1e-14)
    assert torch.allclose(gD, D.grad, rtol=1e-10, atol=1e-14)
    if method == "softmax":
        assert torch.allclose(gw, w.grad, rtol=1e-9, atol=1e-13)


def test_com_filter_matches_reference_definition():
    f = do.com_filter(64, torch.float64)
    assert f.shape == (2, 64, 64)
    assert f[0, 5, 40].item() == (40 - 32) / 63 and f[1, 5, 40].item() == (5 - 32) / 63


def test_reference_model_weights_load_and_oracle_decoder_reproduces_reference_outputs():
    """BASELINE configs[0] on the CPU: the unmodified reference model's state_dict loads into the drop-in
    module tree (strict), and its backbone (plain PyTorch convs) + the oracle decoder reproduce the
    reference's eval outputs on the reference's own SFR crops."""
    import torch
    from helpers import load_model_golden
    g, net = load_model_golden()
    img, label, mask = (torch.from_numpy(g[n]) for n in ("img", "label_img", "mask"))
    with torch.no_grad():
        f = net.conv(img)
        results = []
        for stage in net.stages:
            f, z, d_raw = stage.features_and_logits(f)
            plane = stage.plane_regression
            H, D, uvd = do.decoder_forward(z, plane.temperature, d_raw, label, mask, plane.method)
            results.append((H, D, uvd))
            f = torch.cat([H, D, label], dim=1)
    for i, (H, D, uvd) in enumerate(results):
        assert_close("uvd stage %d" % i, uvd.numpy(), g["ref_uvd_%d" % i])
    assert_close("heat", results[-1][0][:1].numpy(), g["ref_heat_last"])
    assert_close("dmap", results[-1][1][:1].numpy(), g["ref_dmap_last"])
