"""GPU parity: fused decoder forward / backward / backward+loss (through the C
ABI) against the reference-generated golden vectors, the float64 oracle on seeded
inputs, and size-independent properties at B=4096."""
import numpy as np
import pytest
import torch

from oracle import decoder_oracle as do
from pixelwiseregression_b200 import _lib, ops, synth
from helpers import GRAD_RTOL, assert_close, load_golden

pytestmark = pytest.mark.gpu

GOLDEN_SETS = ["decoder_softmax_a1", "decoder_softmax_a05_up", "decoder_sum_a05_up"]
DEV = "cuda:0"


def cu(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dtype)


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_forward_matches_reference_golden(name):
    g = load_golden(name)
    method = str(g["method"])
    w = cu(g["w"]) if method == "softmax" else None
    targets = (cu(g["heat_gt"]), cu(g["dmap_gt"]), cu(g["uvd_gt"]))
    H, uvd, stats, lp = ops.decoder_forward_raw(cu(g["z"]), w, cu(g["D"]), cu(g["label"]), cu(g["mask"]), method,
                                                targets=targets)
    assert_close("heat", H.cpu().numpy(), g["ref_heat"])
    assert_close("uvd", uvd.cpu().numpy(), g["ref_uvd"])
    terms = ops.stage_loss_from_partials(lp, float(g["lambda_h"]), float(g["lambda_d"])).cpu().numpy()
    assert_close("loss terms", terms, g["ref_losses"][:3])


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_backward_loss_matches_reference_golden(name):
    g = load_golden(name)
    method = str(g["method"])
    w = cu(g["w"]) if method == "softmax" else None
    z, D, L, m = cu(g["z"]), cu(g["D"]), cu(g["label"]), cu(g["mask"])
    targets = (cu(g["heat_gt"]), cu(g["dmap_gt"]), cu(g["uvd_gt"]))
    up = "gH_up" in g
    H, uvd, stats, _ = ops.decoder_forward_raw(z, w, D, L, m, method)
    gz, gD, gwp, lp = ops.decoder_backward_raw(
        z, w, D, L, m, stats, uvd, cu(g["g_uvd_up"]) if up else None, cu(g["gH_up"]) if up else None,
        cu(g["gD_up"]) if up else None, method, targets, float(g["alpha"]), float(g["lambda_h"]),
        float(g["lambda_d"]), want_loss=True)
    assert_close("gz", gz.cpu().numpy(), g["ref_gz"], GRAD_RTOL)
    assert_close("gD", gD.cpu().numpy(), g["ref_gD"], GRAD_RTOL)
    if method == "softmax":
        assert_close("gw", ops.reduce_partials(gwp).view(-1, 1).cpu().numpy(), g["ref_gw"], GRAD_RTOL)
    terms = ops.stage_loss_from_partials(lp, float(g["lambda_h"]), float(g["lambda_d"])).cpu().numpy()
    assert_close("loss terms", terms, g["ref_losses"][:3])


def oracle_case(B, J, method, alpha, upstream, seed):
    rng = np.random.default_rng(seed)
    d = synth.make_decoder_inputs(B, J, seed)
    d["z"] = (d["z"] * 3).astype(np.float32)            # sharper softmax than N(0,1)
    heat_gt = rng.uniform(0, 0.05, (B, J, 64, 64)).astype(np.float32)
    dmap_gt = (rng.standard_normal((B, J, 64, 64)) * d["mask"]).astype(np.float32)
    uvd_gt = rng.uniform(-0.5, 0.5, (B, J, 3)).astype(np.float32)
    ups = None
    if upstream:
        ups = ((rng.standard_normal((B, J, 3)) * 1e-2).astype(np.float32),
               (rng.standard_normal((B, J, 64, 64)) * 1e-3).astype(np.float32),
               (rng.standard_normal((B, J, 64, 64)) * 1e-3).astype(np.float32))
    return d, (heat_gt, dmap_gt, uvd_gt), ups


@pytest.mark.parametrize("method", ["softmax", "sum"])
@pytest.mark.parametrize("J", [14, 21])
@pytest.mark.parametrize("alpha,upstream", [(1.0, False), (0.5, True), (1.0, True)])
@pytest.mark.parametrize("direct", [False, True])
def test_forward_backward_match_fp64_oracle(method, J, alpha, upstream, direct):
    """direct=False: the pipelined backward kernels (targets + dense upstream gradients at once = the six-slot
    stage); direct=True: the one-CTA-per-item kernel forced through the dispatch option."""
    B = 37 if not direct else 6            # 37*J items: several per CTA of the persistent kernels, ragged ranges
    d, tg, ups = oracle_case(B, J, method, alpha, upstream, seed=J + int(alpha * 10))
    dd = torch.float64
    t64 = lambda a: torch.from_numpy(a).to(dd)
    w64 = t64(d["w"])
    p_ref, _, uvd_ref = do.decoder_forward(t64(d["z"]), w64, t64(d["D"]), t64(d["label"]), t64(d["mask"]), method)
    g_uvd = t64(ups[0]) if ups else torch.zeros(B, J, 3, dtype=dd)
    gz_ref, gD_ref, gw_ref = do.decoder_backward(
        t64(d["z"]), w64, t64(d["D"]), t64(d["label"]), t64(d["mask"]), g_uvd, t64(ups[1]) if ups else None,
        t64(ups[2]) if ups else None, method, targets=tuple(t64(a) for a in tg), alpha=alpha)
    losses_ref = do.stage_losses(p_ref, t64(d["D"]), uvd_ref, *[t64(a) for a in tg], 1.0, 0.01)

    w = cu(d["w"]) if method == "softmax" else None
    z, D, L, m = cu(d["z"]), cu(d["D"]), cu(d["label"]), cu(d["mask"])
    H, uvd, stats, _ = ops.decoder_forward_raw(z, w, D, L, m, method)
    assert_close("heat", H.cpu().numpy(), p_ref.numpy())
    assert_close("uvd", uvd.cpu().numpy(), uvd_ref.numpy())
    with _lib.option("bwd_direct", int(direct)):
        gz, gD, gwp, lp = ops.decoder_backward_raw(
            z, w, D, L, m, stats, uvd, cu(ups[0]) if ups else None, cu(ups[1]) if ups else None,
            cu(ups[2]) if ups else None, method, tuple(cu(a) for a in tg), alpha, 1.0, 0.01, want_loss=True)
    assert_close("gz", gz.cpu().numpy(), gz_ref.numpy(), GRAD_RTOL)
    assert_close("gD", gD.cpu().numpy(), gD_ref.numpy(), GRAD_RTOL)
    if method == "softmax":
        assert_close("gw", ops.reduce_partials(gwp).view(-1, 1).cpu().numpy(), gw_ref.numpy(), GRAD_RTOL)
    terms = ops.stage_loss_from_partials(lp, 1.0, 0.01).cpu().numpy()
    assert_close("losses", terms, [x.item() for x in losses_ref])


def test_autograd_function_matches_oracle_autograd():
    """DecoderFunction inside an autograd graph (dense upstream on H and D, like
    the inner stage of the model) == autograd through the float64 oracle."""
    B, J = 4, 14
    d, tg, ups = oracle_case(B, J, "softmax", 0.5, True, seed=5)
    def run(dtype, dev, fused):
        z = torch.from_numpy(d["z"]).to(dev, dtype).requires_grad_(True)
        D = torch.from_numpy(d["D"]).to(dev, dtype).requires_grad_(True)
        w = torch.from_numpy(d["w"]).to(dev, dtype).requires_grad_(True)
        L, m = torch.from_numpy(d["label"]).to(dev, dtype), torch.from_numpy(d["mask"]).to(dev, dtype)
        if fused:
            H, Dm, uvd = ops.fused_decoder(z, w, D, L, m, "softmax")
        else:
            H, Dm, uvd = do.decoder_forward(z, w, D, L, m, "softmax")
        tt = [torch.from_numpy(a).to(dev, dtype) for a in tg]
        loss = do.combine_losses(do.stage_losses(H, Dm, uvd, *tt, 1.0, 0.01), 0.5)
        extra = (H * torch.from_numpy(ups[1]).to(dev, dtype)).sum() + (Dm ** 2 * torch.from_numpy(ups[2]).to(dev, dtype)).sum()
        (loss + extra).backward()
        return [t.grad.detach().cpu().double().numpy() for t in (z, D, w)] + [loss.item()]
    ref = run(torch.float64, "cpu", False)
    got = run(torch.float32, DEV, True)
    for name, a, b in zip(("gz", "gD", "gw"), got, ref):
        assert_close(name, a, b, GRAD_RTOL)
    assert abs(got[3] - ref[3]) <= 1e-5 * abs(ref[3])


@pytest.mark.parametrize("alpha", [1.0, 0.4])
def test_eager_loss_function_matches_oracle(alpha):
    B, J = 4, 14
    d, tg, _ = oracle_case(B, J, "softmax", alpha, False, seed=9)
    dd = torch.float64
    z64 = torch.from_numpy(d["z"]).to(dd).requires_grad_(True)
    D64 = torch.from_numpy(d["D"]).to(dd).requires_grad_(True)
    w64 = torch.from_numpy(d["w"]).to(dd).requires_grad_(True)
    p, Dm, uvd_ref = do.decoder_forward(z64, w64, D64, torch.from_numpy(d["label"]).to(dd),
                                        torch.from_numpy(d["mask"]).to(dd))
    terms_ref = do.stage_losses(p, Dm, uvd_ref, *[torch.from_numpy(a).to(dd) for a in tg], 1.0, 0.01)
    (3.0 * do.combine_losses(terms_ref, alpha)).backward()

    z = cu(d["z"]).requires_grad_(True)
    D = cu(d["D"]).requires_grad_(True)
    w = cu(d["w"]).requires_grad_(True)
    total, terms, uvd, H = ops.fused_decoder_loss(z, w, D, cu(d["label"]), cu(d["mask"]), *[cu(a) for a in tg],
                                                  method="softmax", alpha=alpha)
    (3.0 * total).backward()          # non-unit upstream exercises pwr_scale_inplace
    assert_close("terms", terms.cpu().numpy(), [t.item() for t in terms_ref])
    assert_close("heat", H.cpu().numpy(), p.detach().numpy())
    assert_close("gz", z.grad.cpu().numpy(), z64.grad.numpy(), GRAD_RTOL)
    assert_close("gD", D.grad.cpu().numpy(), D64.grad.numpy(), GRAD_RTOL)
    assert_close("gw", w.grad.cpu().numpy(), w64.grad.numpy(), GRAD_RTOL)


def test_inner_stage_fused_loss_matches_oracle():
    B, J = 3, 14
    d, tg, ups = oracle_case(B, J, "softmax", 0.5, True, seed=13)
    dd = torch.float64
    z64 = torch.from_numpy(d["z"]).to(dd).requires_grad_(True)
    D64 = torch.from_numpy(d["D"]).to(dd).requires_grad_(True)
    w64 = torch.from_numpy(d["w"]).to(dd).requires_grad_(True)
    p, Dm, uvd_ref = do.decoder_forward(z64, w64, D64, torch.from_numpy(d["label"]).to(dd),
                                        torch.from_numpy(d["mask"]).to(dd))
    loss_ref = do.combine_losses(do.stage_losses(p, Dm, uvd_ref, *[torch.from_numpy(a).to(dd) for a in tg], 1.0, 0.01), 0.5)
    (2.0 * loss_ref + (p * torch.from_numpy(ups[1]).to(dd)).sum() + (Dm * torch.from_numpy(ups[2]).to(dd)).sum()).backward()
    z = cu(d["z"]).requires_grad_(True)
    D = cu(d["D"]).requires_grad_(True)
    w = cu(d["w"]).requires_grad_(True)
    H, Dout, uvd, stage_loss, terms = ops.fused_decoder_with_loss(z, w, D, cu(d["label"]), cu(d["mask"]),
                                                                  *[cu(a) for a in tg], alpha=0.5)
    (2.0 * stage_loss + (H * cu(ups[1])).sum() + (Dout * cu(ups[2])).sum()).backward()
    assert abs(stage_loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    assert_close("gz", z.grad.cpu().numpy(), z64.grad.numpy(), GRAD_RTOL)
    assert_close("gD", D.grad.cpu().numpy(), D64.grad.numpy(), GRAD_RTOL)
    assert_close("gw", w.grad.cpu().numpy(), w64.grad.numpy(), GRAD_RTOL)


def test_standalone_plane_and_depth_functions():
    B, J = 3, 5
    d, _, _ = oracle_case(B, J, "softmax", 1.0, False, seed=3)
    dd = torch.float64
    z64 = torch.from_numpy(d["z"]).to(dd).requires_grad_(True)
    w64 = torch.from_numpy(d["w"]).to(dd).requires_grad_(True)
    p, uv = do.plane_decode(z64, w64)
    hin64 = torch.rand(B, J, 64, 64, dtype=dd).requires_grad_(True)
    D64 = torch.from_numpy(d["D"]).to(dd).requires_grad_(True)
    dep = do.depth_decode(D64, hin64, torch.from_numpy(d["label"]).to(dd), torch.from_numpy(d["mask"]).to(dd))
    ((uv ** 2).sum() + (p ** 2).sum() + (dep ** 2).sum()).backward()
    z = cu(d["z"]).requires_grad_(True)
    w = cu(d["w"]).requires_grad_(True)
    H, uv_g = ops.PlaneFunction.apply(z, w, "softmax")
    hin = hin64.detach().float().to(DEV).requires_grad_(True)
    D = cu(d["D"]).requires_grad_(True)
    dep_g = ops.DepthFunction.apply(D, hin, cu(d["label"]), cu(d["mask"]))
    ((uv_g ** 2).sum() + (H ** 2).sum() + (dep_g ** 2).sum()).backward()
    assert_close("uv", uv_g.detach().cpu().numpy(), uv.detach().numpy())
    assert_close("dep", dep_g.detach().cpu().numpy(), dep.detach().numpy())
    assert_close("gz", z.grad.cpu().numpy(), z64.grad.numpy(), GRAD_RTOL)
    assert_close("gw", w.grad.cpu().numpy(), w64.grad.numpy(), GRAD_RTOL)
    assert_close("gheat", hin.grad.cpu().numpy(), hin64.grad.numpy(), GRAD_RTOL)
    assert_close("gD", D.grad.cpu().numpy(), D64.grad.numpy(), GRAD_RTOL)


def test_edge_inputs():
    """All-masked sample (den = 1e-14), negative temperature, huge logits, B=0."""
    B, J = 2, 3
    z = torch.randn(B, J, 64, 64, device=DEV) * 50
    D = torch.randn(B, J, 64, 64, device=DEV)
    w = torch.tensor([[1.0], [-0.7], [2.5]], device=DEV)
    L = torch.rand(B, 1, 64, 64, device=DEV)
    m = torch.zeros(B, 1, 64, 64, device=DEV)
    m[1, :, 10:30, 10:30] = 1
    H, uvd, stats, _ = ops.decoder_forward_raw(z, w, D, L, m)
    p_ref, _, uvd_ref = do.decoder_forward(z.double().cpu(), w.double().cpu(), D.double().cpu(), L.double().cpu(),
                                           m.double().cpu())
    assert torch.isfinite(H).all() and torch.isfinite(uvd).all()
    assert float(uvd[0, :, 2].abs().max()) == 0.0
    assert_close("heat", H.cpu().numpy(), p_ref.numpy())
    assert_close("uvd", uvd.cpu().numpy(), uvd_ref.numpy())
    e = torch.empty(0, J, 64, 64, device=DEV)
    H0, uvd0, _, _ = ops.decoder_forward_raw(e, w, e, torch.empty(0, 1, 64, 64, device=DEV),
                                             torch.empty(0, 1, 64, 64, device=DEV))
    assert H0.shape == (0, J, 64, 64) and uvd0.shape == (0, J, 3)


def test_full_batch_properties():
    """B=4096, J=14 (benchmark size): properties that need no oracle."""
    B, J = 4096, 14
    g = torch.Generator(device=DEV).manual_seed(0)
    z = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    w = torch.rand(J, 1, device=DEV, generator=g) + 0.5
    m = (torch.rand(B, 1, 64, 64, device=DEV, generator=g) < 0.4).float()
    L = torch.rand(B, 1, 64, 64, device=DEV, generator=g) * m
    H, uvd, stats, _ = ops.decoder_forward_raw(z, w, D, L, m)
    assert float((H.sum(dim=(2, 3)) - 1).abs().max()) < 1e-5                 # softmax normalisation
    assert float(uvd[:, :, :2].abs().max()) <= 0.5 + 1e-6                    # inside the U/V range
    # shift invariance of the softmax: adding a per-map constant to z changes nothing
    H2, uvd2, _, _ = ops.decoder_forward_raw(z + 3.0, w, D, L, m)
    assert float((H2 - H).abs().max()) < 1e-6 * float(H.max()) + 1e-9
    # d is a convex combination of the masked reconstruction
    rec = (D + L) * m
    assert bool((uvd[:, :, 2] <= rec.amax(dim=(2, 3)) + 1e-4).all() and (uvd[:, :, 2] >= rec.amin(dim=(2, 3)) - 1e-4).all())
    # backward is linear in the upstream gradients
    g1 = torch.randn(B, J, 3, device=DEV, generator=g)
    g2 = torch.randn(B, J, 3, device=DEV, generator=g)
    a = ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, g1)
    b = ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, g2)
    c = ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, g1 + g2)
    for x, y, s in zip(a[:3], b[:3], c[:3]):
        assert float((x + y - s).abs().max()) <= 1e-5 * float(s.abs().max()) + 1e-9
    # softmax gradient sums to zero over each map (p sums to one)
    assert float(a[0].sum(dim=(2, 3)).abs().max()) < 1e-4 * float(a[0].abs().sum(dim=(2, 3)).max())
    # deterministic
    a2 = ops.decoder_backward_raw(z, w, D, L, m, stats, uvd, g1)
    assert torch.equal(a[0], a2[0]) and torch.equal(a[1], a2[1]) and torch.equal(a[2], a2[2])


def test_cpu_tensors_are_rejected():
    from pixelwiseregression_b200._lib import PwrError
    z = torch.zeros(1, 2, 64, 64)
    with pytest.raises(PwrError):
        ops.decoder_forward_raw(z, torch.ones(2, 1), z, torch.zeros(1, 1, 64, 64), torch.zeros(1, 1, 64, 64))


def test_recover_uvd_and_joint_error_match_oracle():
    """utils.py:332-337, datasets.py:100-111, train.py:254-276 on the GPU vs the NumPy restatement."""
    from oracle import sfr_oracle as so
    rng = np.random.default_rng(4)
    B, J = 257, 21
    shape = synth.HAND17
    pred = rng.uniform(-0.5, 0.5, (B, J, 3)).astype(np.float32)
    true = (pred + rng.normal(0, 0.02, (B, J, 3))).astype(np.float32)
    box = rng.integers(100, 352, B).astype(np.float32)
    cube = np.full(B, 150, np.float32)
    com = np.stack([rng.integers(100, 500, B), rng.integers(100, 400, B), rng.uniform(500, 1000, B)], 1).astype(np.float32)
    intr = (shape.fx, shape.fy, shape.halfu, shape.halfv)
    uvd_px, xyz = ops.recover_uvd(cu(pred), cu(box), cu(com), cu(cube), intrinsics=intr)
    ref_px = so.recover_uvd(pred, box, com, cube)
    assert (uvd_px.cpu().numpy() == ref_px).all()                         # same float32 operation order
    assert_close("xyz", xyz.cpu().numpy(), so.uvd2xyz(ref_px, *intr))
    err = ops.joint_error(cu(pred), cu(true), cu(box), cu(com), cu(cube), intr)
    assert_close("joint error", err.cpu().numpy(), so.joint_error(pred, true, box, com, cube, *intr))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("method", ["softmax", "sum"])
def test_half_precision_conv_outputs_equal_upcast_path(dtype, method):
    """fp16 / bf16 logits (autocast, train.py:170-172) read directly by the kernels give
    exactly what the float32 kernels give on the up-cast values (autocast's own semantics), with the
    gradients rounded once to the conv dtype."""
    B, J = 5, 14
    d, tg, ups = oracle_case(B, J, method, 0.5, True, seed=21)
    z16, D16 = cu(d["z"]).to(dtype), cu(d["D"]).to(dtype)
    w = cu(d["w"]) if method == "softmax" else None
    L, m = cu(d["label"]), cu(d["mask"])
    targets = tuple(cu(a) for a in tg)
    gH_up, gD_up16 = cu(ups[1]), cu(ups[2]).to(dtype)
    H16, uvd16, st16, lp16 = ops.decoder_forward_raw(z16, w, D16, L, m, method, targets=targets)
    H32, uvd32, st32, lp32 = ops.decoder_forward_raw(z16.float(), w, D16.float(), L, m, method, targets=targets)
    assert H16.dtype == torch.float32 and torch.equal(H16, H32) and torch.equal(uvd16, uvd32) and torch.equal(lp16, lp32)
    for kw in (dict(targets=targets, alpha=0.5, want_loss=True),                 # pipelined, targets in the slots
               dict(gH_up=gH_up, gD_up=gD_up16),                                 # pipelined, upstream maps in the slots
               dict(gH_up=gH_up, gD_up=gD_up16, targets=targets, alpha=0.5)):    # direct-load kernel (both pairs)
        kw32 = dict(kw)
        if "gD_up" in kw32:
            kw32["gD_up"] = kw32["gD_up"].float()
        a = ops.decoder_backward_raw(z16, w, D16, L, m, st16, uvd16, cu(ups[0]), method=method, **kw)
        b = ops.decoder_backward_raw(z16.float(), w, D16.float(), L, m, st32, uvd32, cu(ups[0]), method=method, **kw32)
        assert a[0].dtype == dtype and a[1].dtype == dtype
        assert torch.equal(a[0], b[0].to(dtype)) and torch.equal(a[1], b[1].to(dtype))
        if method == "softmax":
            assert torch.equal(a[2], b[2])
        if a[3] is not None:
            assert torch.equal(a[3], b[3])


def test_autocast_training_step_through_the_fused_decoder():
    """fp16 logits inside an autograd graph with a GradScaler-like upstream scale."""
    B, J = 3, 14
    d, tg, _ = oracle_case(B, J, "softmax", 0.5, False, seed=22)
    z = cu(d["z"]).half().requires_grad_(True)
    D = cu(d["D"]).half().requires_grad_(True)
    w = cu(d["w"]).requires_grad_(True)
    total, terms, uvd, H = ops.fused_decoder_loss(z, w, D, cu(d["label"]), cu(d["mask"]), *[cu(a) for a in tg], alpha=0.5)
    (total * 1024.0).backward()
    assert z.grad.dtype == torch.float16 and D.grad.dtype == torch.float16 and H.dtype == torch.float32
    zf = z.detach().float().requires_grad_(True)
    Df = D.detach().float().requires_grad_(True)
    wf = w.detach().clone().requires_grad_(True)
    total2, *_ = ops.fused_decoder_loss(zf, wf, Df, cu(d["label"]), cu(d["mask"]), *[cu(a) for a in tg], alpha=0.5)
    (total2 * 1024.0).backward()
    assert torch.equal(total, total2)
    assert_close("gz", z.grad.float().cpu().numpy(), zf.grad.cpu().numpy(), 2e-3)     # two fp16 roundings (eager, then scale)
    assert_close("gD", D.grad.float().cpu().numpy(), Df.grad.cpu().numpy(), 2e-3)
    assert_close("gw", w.grad.cpu().numpy(), wf.grad.cpu().numpy(), 1e-6)
    # drop-in autograd node with half inputs
    z2 = cu(d["z"]).half().requires_grad_(True)
    D2 = cu(d["D"]).half().requires_grad_(True)
    H2, Dm, uvd2 = ops.fused_decoder(z2, w, D2, cu(d["label"]), cu(d["mask"]))
    assert H2.dtype == torch.float32 and Dm.dtype == torch.float16 and uvd2.dtype == torch.float32
    ((H2 ** 2).sum() + (Dm.float() ** 2).sum() * 1e-3 + (uvd2 ** 2).sum()).backward()
    assert z2.grad.dtype == torch.float16 and D2.grad.dtype == torch.float16


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_sparse_targets_equal_dense_targets(alpha):
    """sfr.build_sfr(targets="both"): the loss kernels evaluating heat-map / depth-map targets on the
    fly from the 64-byte taps give the same loss terms and gradients as reading the dense maps."""
    from pixelwiseregression_b200 import sfr
    shape = synth.NYU
    B, J = 12, shape.joints
    d = synth.make_frames(shape, B, seed=51)
    uvd_np = d["uvd"].copy()
    uvd_np[0, 0, :2] = d["com"][0, :2] + np.array([-0.97, 0.3]) * 100      # a joint on the map border (reflection)
    batch = sfr.build_sfr(torch.from_numpy(d["frames"]).cuda(), d["com"], d["cube"], uvd_np, fx=shape.fx, fy=shape.fy,
                          targets="both")
    assert batch.taps.shape == (B, J, 64) and batch.taps.dtype == torch.uint8
    g = torch.Generator(device=DEV).manual_seed(3)
    z = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 3
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    w = torch.rand(J, 1, device=DEV, generator=g) + 0.5
    gH = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 1e-3
    gDu = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 1e-3
    dense = (batch.heatmaps, batch.depthmaps, batch.uvd)
    sparse = ops.SparseTargets(batch.taps, batch.uvd)
    # forward with loss (inner stage)
    _, uvd_d, st, lp_d = ops.decoder_forward_raw(z, w, D, batch.label_img, batch.mask, targets=dense)
    _, uvd_s, _, lp_s = ops.decoder_forward_raw(z, w, D, batch.label_img, batch.mask, targets=sparse)
    assert torch.equal(uvd_d, uvd_s)
    assert_close("fwd loss partials", lp_s.cpu().numpy(), lp_d.cpu().numpy(), 1e-5)
    # backward + loss: pipelined (no upstream), pipelined with upstream maps (only possible with sparse
    # targets), and the direct kernel
    for up in (dict(), dict(gH_up=gH, gD_up=gDu)):
        a = ops.decoder_backward_raw(z, w, D, batch.label_img, batch.mask, st, uvd_d, targets=dense, alpha=alpha,
                                     want_loss=True, **up)
        b = ops.decoder_backward_raw(z, w, D, batch.label_img, batch.mask, st, uvd_d, targets=sparse, alpha=alpha,
                                     want_loss=True, **up)
        assert_close("gz", b[0].cpu().numpy(), a[0].cpu().numpy(), 1e-5)
        assert_close("gD", b[1].cpu().numpy(), a[1].cpu().numpy(), 1e-5)
        assert_close("gw", b[2].cpu().numpy(), a[2].cpu().numpy(), 1e-5)
        assert_close("loss partials", b[3].cpu().numpy(), a[3].cpu().numpy(), 1e-5)
    # through the public fused criterion
    zz = z.clone().requires_grad_(True)
    total_d, terms_d, *_ = ops.fused_decoder_loss(zz, w, D, batch.label_img, batch.mask, batch.heatmaps,
                                                  batch.depthmaps, batch.uvd, alpha=alpha)
    total_s, terms_s, *_ = ops.fused_decoder_loss(zz, w, D, batch.label_img, batch.mask, batch.taps, None, batch.uvd,
                                                  alpha=alpha)
    assert_close("terms", terms_s.cpu().numpy(), terms_d.cpu().numpy(), 1e-5)
    assert abs(total_s.item() - total_d.item()) <= 1e-5 * abs(total_d.item())


@pytest.mark.parametrize("method,dtype", [("softmax", torch.float32), ("sum", torch.float32),
                                          ("softmax", torch.float16), ("sum", torch.bfloat16)])
@pytest.mark.parametrize("B,J", [(1, 3), (7, 14), (41, 21), (300, 5)])
def test_pipelined_forward_equals_direct_forward(method, dtype, B, J, monkeypatch):
    """The persistent TMA-pipelined forward (no fused loss) and the one-CTA-per-item forward compute the
    same heat maps bit for bit (same exp argument, same 1/sum up to the summation order) and the same
    coordinates within float32 summation-order noise; ragged item ranges per CTA, sample boundaries
    inside a range (L/m buffer switch) and negative temperatures included."""
    rng = np.random.default_rng(B * 100 + J)
    d = synth.make_decoder_inputs(B, J, seed=B + J)
    z, D = cu(d["z"] * 3).to(dtype), cu(d["D"]).to(dtype)
    L, m = cu(d["label"]), cu(d["mask"])
    w = None
    if method == "softmax":
        wv = rng.uniform(0.5, 1.5, (J, 1)).astype(np.float32)
        wv[::3] *= -1.0                                     # extremum = min for these joints
        w = cu(wv)
    for store_heat in (True, False):
        with _lib.option("fwd_direct", 1):
            Hd, uvd_d, st_d, _ = ops.decoder_forward_raw(z, w, D, L, m, method, store_heat=store_heat)
        with _lib.option("fwd_pipe", 1):                    # default: pipelined only without the heat-map store
            Hp, uvd_p, st_p, _ = ops.decoder_forward_raw(z, w, D, L, m, method, store_heat=store_heat)
        torch.cuda.synchronize()
        assert torch.equal(st_d[..., 0], st_p[..., 0])      # extremum: exact
        assert_close("1/sum", st_p[..., 1].cpu().numpy(), st_d[..., 1].cpu().numpy(), 1e-6)
        assert_close("uvd", uvd_p.cpu().numpy(), uvd_d.cpu().numpy(), 2e-6)
        assert_close("stats", st_p.cpu().numpy(), st_d.cpu().numpy(), 2e-6)
        if store_heat:
            assert_close("heat", Hp.cpu().numpy(), Hd.cpu().numpy(), 1e-6)
        else:
            assert Hp is None and Hd is None
    # against the float64 oracle as well (the pipelined kernel is the default path)
    t64 = lambda a: a.detach().cpu().to(torch.float64)
    p_ref, _, uvd_ref = do.decoder_forward(t64(z), t64(w) if w is not None else None, t64(D), t64(L), t64(m), method)
    with _lib.option("fwd_pipe", 1):
        Hp, uvd_p, _, _ = ops.decoder_forward_raw(z, w, D, L, m, method)
    _, uvd_n, _, _ = ops.decoder_forward_raw(z, w, D, L, m, method, store_heat=False, want_stats=False)   # default route
    assert_close("heat vs oracle", Hp.cpu().numpy(), p_ref.numpy())
    assert_close("uvd vs oracle", uvd_p.cpu().numpy(), uvd_ref.numpy())
    assert torch.equal(uvd_n, uvd_p)


@pytest.mark.parametrize("method,dtype", [("softmax", torch.float32), ("sum", torch.float32),
                                          ("softmax", torch.bfloat16), ("sum", torch.float16)])
@pytest.mark.parametrize("B,J", [(1, 1), (9, 14), (301, 5)])
def test_lean_backward_equals_pipelined_backward(method, dtype, B, J, monkeypatch):
    """Configurations without dense target / upstream maps (compact targets, uvd-only loss, plain
    backward) run in the lean multi-CTA-per-SM kernel; PWR_BWD_LEAN=0 sends them through the
    one-CTA-per-SM pipelined kernel.  Same arithmetic per pixel, different summation order."""
    from pixelwiseregression_b200 import sfr
    shape = synth.NYU
    rng = np.random.default_rng(B + 7 * J)
    d = synth.make_frames(shape, B, seed=B + J)
    uvd_j = d["uvd"][:, :J] if J <= shape.joints else None
    batch = sfr.build_sfr(torch.from_numpy(d["frames"]).cuda(), d["com"], d["cube"], uvd_j, fx=shape.fx, fy=shape.fy,
                          targets="sparse")
    g = torch.Generator(device=DEV).manual_seed(B * J)
    z = (torch.randn(B, J, 64, 64, device=DEV, generator=g) * 3).to(dtype)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g).to(dtype)
    w = None
    if method == "softmax":
        wv = rng.uniform(0.5, 1.5, (J, 1)).astype(np.float32)
        wv[::4] *= -1.0
        w = cu(wv)
    L, m = batch.label_img, batch.mask
    _, uvd, st, _ = ops.decoder_forward_raw(z, w, D, L, m, method)
    g_uvd = torch.randn(B, J, 3, device=DEV, generator=g) * 1e-2
    sparse = ops.SparseTargets(batch.taps, batch.uvd)
    cases = [dict(targets=sparse, alpha=0.5, want_loss=True, g_uvd=g_uvd),     # compact targets, all three terms
             dict(targets=sparse, alpha=1.0, want_loss=False),                 # uvd term only, no map reads at all
             dict(g_uvd=g_uvd),                                                # plain backward
             dict(g_uvd=g_uvd, want_gD=False)]
    for kw in cases:
        with _lib.option("bwd_no_lean", 1):
            a = ops.decoder_backward_raw(z, w, D, L, m, st, uvd, method=method, **kw)
        b = ops.decoder_backward_raw(z, w, D, L, m, st, uvd, method=method, **kw)
        torch.cuda.synchronize()
        for name, x, y in zip(("gz", "gD", "gw_partial", "loss_partial"), b, a):
            assert (x is None) == (y is None), name
            if x is not None:
                tol = 1e-5 if dtype == torch.float32 else 1e-2          # half outputs: one rounding of a ~1e-6-different value
                assert_close(name, x.float().cpu().numpy(), y.float().cpu().numpy(), tol)


@pytest.mark.parametrize("method,dtype", [("softmax", torch.float32), ("sum", torch.float32),
                                          ("softmax", torch.float16), ("sum", torch.bfloat16)])
@pytest.mark.parametrize("B,J,alpha", [(1, 1, 0.5), (10, 14, 1.0), (10, 14, 0.3), (151, 5, 0.5)])
def test_one_pass_last_stage_equals_forward_then_backward(method, dtype, B, J, alpha, monkeypatch):
    """pwr_decoder_fwd_bwd_loss (forward + loss + backward in one visit of z, D and the targets) against
    pwr_decoder_fwd followed by pwr_decoder_bwd_loss, with dense and with compact targets, through the raw
    entry point and through ops.fused_decoder_loss; and against the float64 oracle."""
    from pixelwiseregression_b200 import sfr
    shape = synth.NYU
    rng = np.random.default_rng(3 * B + J)
    d = synth.make_frames(shape, B, seed=B + 2 * J)
    batch = sfr.build_sfr(torch.from_numpy(d["frames"]).cuda(), d["com"], d["cube"], d["uvd"][:, :J], fx=shape.fx,
                          fy=shape.fy, targets="both")
    g = torch.Generator(device=DEV).manual_seed(B * J + 1)
    z = (torch.randn(B, J, 64, 64, device=DEV, generator=g) * 3).to(dtype)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g).to(dtype)
    w = None
    if method == "softmax":
        wv = rng.uniform(0.5, 1.5, (J, 1)).astype(np.float32)
        wv[::5] *= -1.0
        w = cu(wv)
    L, m = batch.label_img, batch.mask
    dense = (batch.heatmaps, batch.depthmaps, batch.uvd)
    sparse = ops.SparseTargets(batch.taps, batch.uvd)
    tol = 2e-5 if dtype == torch.float32 else 1e-2            # half gradients: one rounding of a ~1e-6-different value
    for targets in (dense, sparse):
        for store_heat in (True, False):
            H1, uvd1, gz1, gD1, gw1, lp1 = ops.decoder_fused_raw(z, w, D, L, m, targets, method, alpha,
                                                                store_heat=store_heat)
            H2, uvd2, st2, _ = ops.decoder_forward_raw(z, w, D, L, m, method, store_heat=store_heat)
            gz2, gD2, gw2, lp2 = ops.decoder_backward_raw(z, w, D, L, m, st2, uvd2, method=method, targets=targets,
                                                          alpha=alpha, want_loss=True)
            torch.cuda.synchronize()
            assert (H1 is None) == (not store_heat)
            if store_heat:
                assert_close("H", H1.cpu().numpy(), H2.cpu().numpy(), 2e-6)
            assert_close("uvd", uvd1.cpu().numpy(), uvd2.cpu().numpy(), 2e-6)
            assert_close("loss partials", lp1.cpu().numpy(), lp2.cpu().numpy(), 1e-5)
            assert_close("gz", gz1.float().cpu().numpy(), gz2.float().cpu().numpy(), tol)
            assert_close("gD", gD1.float().cpu().numpy(), gD2.float().cpu().numpy(), tol)
            if method == "softmax":
                assert_close("gw", ops.reduce_partials(gw1).cpu().numpy(), ops.reduce_partials(gw2).cpu().numpy(), 2e-5)
    # public criterion, both routes
    outs = []
    for one_pass in (True, False):
        monkeypatch.setattr(ops, "ONE_PASS_LAST_STAGE", one_pass)
        zz, DD = z.clone().requires_grad_(True), D.clone().requires_grad_(True)
        ww = w.clone().requires_grad_(True) if w is not None else None
        total, terms, uvd_o, H_o = ops.fused_decoder_loss(zz, ww, DD, L, m, batch.heatmaps, batch.depthmaps, batch.uvd,
                                                          method, alpha)
        total.backward()
        outs.append((total.detach(), terms, uvd_o, zz.grad, DD.grad, ww.grad if ww is not None else None))
    for a, b in zip(*outs):
        if a is not None:
            assert_close("criterion", a.float().cpu().numpy(), b.float().cpu().numpy(), tol)
    # float64 oracle
    t64 = lambda a: a.detach().cpu().to(torch.float64)
    w64 = t64(w) if w is not None else None
    p_ref, _, uvd_ref = do.decoder_forward(t64(z), w64, t64(D), t64(L), t64(m), method)
    gz_ref, gD_ref, gw_ref = do.decoder_backward(t64(z), w64, t64(D), t64(L), t64(m), torch.zeros(B, J, 3, dtype=torch.float64),
                                                 None, None, method, targets=tuple(t64(a) for a in dense), alpha=alpha)
    H1, uvd1, gz1, gD1, gw1, lp1 = ops.decoder_fused_raw(z, w, D, L, m, dense, method, alpha)
    assert_close("H vs oracle", H1.cpu().numpy(), p_ref.numpy())
    assert_close("uvd vs oracle", uvd1.cpu().numpy(), uvd_ref.numpy())
    if dtype == torch.float32:
        assert_close("gz vs oracle", gz1.cpu().numpy(), gz_ref.numpy(), GRAD_RTOL)
        assert_close("gD vs oracle", gD1.cpu().numpy(), gD_ref.numpy(), GRAD_RTOL)
        if method == "softmax":
            assert_close("gw vs oracle", ops.reduce_partials(gw1).view(-1, 1).cpu().numpy(), gw_ref.numpy(), GRAD_RTOL)
