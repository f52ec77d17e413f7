"""GPU parity pins added in round 2 ("parity evidence holes"):

* the benchmark configuration itself (B = 4096 NYU) against the oracle on randomly drawn samples;
* the `n_mean` argument (global-batch mean under data parallelism) of the three loss entry points;
* every instantiation of the one-pass last-stage kernel with several items per CTA (the shape the
  compute-sanitizer runs of tools/sanitize.sh use);
* non-contiguous / float64 arguments of the host wrappers (no aliasing temporaries);
* float16 last stage under a GradScaler-sized loss scale with realistically small gradients;
* two NCCL ranks through the CUDA kernels (skipped below 2 GPUs).
"""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as do
from oracle import sfr_oracle as so
from pixelwiseregression_b200 import ops, sfr, synth
from helpers import GRAD_RTOL, SFR_FIELDS, assert_close, assert_sfr_matches

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_benchmark_batch_sampled_against_oracle():
    """bench.py's configuration (B = 4096 NYU frames, J = 14, N(0,1) logits): build on the GPU, pull 32
    random samples to the host, and compare the SFR outputs and the one-pass loss / gz / gD / dL/dw of
    exactly those samples with the oracle (masks / box exact, maps 1e-5, gradients 1e-4)."""
    shape = synth.NYU
    B, J, S = 4096, shape.joints, 32
    d = synth.make_frames_device(shape, B, seed=0, device=DEV)
    batch = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy)
    g = torch.Generator(device=DEV).manual_seed(1000)
    z = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    w = torch.rand(J, 1, device=DEV, generator=g) + 0.5
    alpha = 0.5
    H, uvd, gz, gD, gwp, lp = ops.decoder_fused_raw(z, w, D, batch.label_img, batch.mask,
                                                    (batch.heatmaps, batch.depthmaps, batch.uvd), "softmax", alpha)
    total = ops.stage_loss(lp, 1.0, 0.01, alpha)
    torch.cuda.synchronize()
    assert bool(batch.valid.all())
    idx = torch.from_numpy(np.sort(np.random.default_rng(7).choice(B, S, replace=False))).to(DEV)
    pick = lambda t: t.index_select(0, idx).cpu()
    # ---- SFR on the sampled frames
    ref = so.process_batch(pick(d["frames"]).numpy(), pick(d["uvd"]).numpy(), pick(d["com"]).numpy(),
                           pick(d["cube"]).numpy(), shape.fx, shape.fy)
    got = {k: pick(v).numpy() for k, v in batch._asdict().items() if v is not None}
    got["dmap"] = got.pop("depthmaps")
    assert_sfr_matches(got, ref, SFR_FIELDS, ref["valid"], prefix="B4096:")
    for n in ("img", "label_img"):
        assert (got[n] == ref[n]).all(), n
    # ---- decoder + loss + backward on the same samples; the means run over the FULL batch (N = B*J)
    t64 = lambda t: pick(t).double()
    z64, D64, w64 = t64(z), t64(D), w.double().cpu()
    L64, m64 = t64(batch.label_img), t64(batch.mask)
    tg = (t64(batch.heatmaps), t64(batch.depthmaps), t64(batch.uvd))
    p_ref, _, uvd_ref = do.decoder_forward(z64, w64, D64, L64, m64)
    gz_ref, gD_ref, gw_ref = do.decoder_backward(z64, w64, D64, L64, m64, torch.zeros(S, J, 3, dtype=torch.float64),
                                                 None, None, "softmax", targets=tg, alpha=alpha)
    scale = S / B                    # the oracle's means ran over S*J items, the kernel's over B*J
    assert_close("H", pick(H).numpy(), p_ref.numpy())
    assert_close("uvd", pick(uvd).numpy(), uvd_ref.numpy())
    assert_close("gz", pick(gz).numpy(), (gz_ref * scale).numpy(), GRAD_RTOL)
    assert_close("gD", pick(gD).numpy(), (gD_ref * scale).numpy(), GRAD_RTOL)
    assert_close("gw (sampled items)", pick(gwp).double().sum(0).view(J, 1).numpy(), (gw_ref * scale).numpy(), GRAD_RTOL)
    lp_ref = torch.stack([((p_ref - tg[0]) ** 2).sum((2, 3)), ((D64 - tg[1]) ** 2).sum((2, 3)),
                          ((uvd_ref - tg[2]) ** 2).sum(2)], dim=2)
    assert_close("loss partials", pick(lp).numpy(), lp_ref.numpy())
    # the [4] loss vector is what pwr_stage_loss makes of ALL partials: check it against a float64 sum of them
    lp_all = lp.double().cpu()
    means = lp_all.sum((0, 1)) / (B * J)
    expect = torch.tensor([means[0], 0.01 * means[1], means[2],
                           alpha * means[2] + (1 - alpha) * (means[0] + 0.01 * means[1])])
    assert_close("stage loss", total.cpu().numpy(), expect.numpy())


@pytest.mark.parametrize("targets_kind", ["dense", "sparse"])
def test_n_mean_two_half_batches_sum_to_the_full_batch(targets_kind):
    """train.py:197-199 are means over B*J.  Under data parallelism with SUMMED gradients (or micro-batching)
    every shard must divide by the global B*J: two half batches run with n_mean = B*J must add up to the
    full batch, for the one-pass kernel, the backward+loss kernel and pwr_stage_loss."""
    shape = synth.NYU
    B, J, alpha = 12, 5, 0.4
    d = synth.make_frames(shape, B, seed=11)
    batch = sfr.build_sfr(torch.from_numpy(d["frames"]).to(DEV), d["com"], d["cube"], d["uvd"][:, :J], fx=shape.fx,
                          fy=shape.fy, targets="both")
    g = torch.Generator(device=DEV).manual_seed(5)
    z = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 2
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    w = torch.rand(J, 1, device=DEV, generator=g) + 0.5

    def tg(lo, hi):
        if targets_kind == "sparse":
            return ops.SparseTargets(batch.taps[lo:hi].contiguous(), batch.uvd[lo:hi].contiguous())
        return (batch.heatmaps[lo:hi], batch.depthmaps[lo:hi], batch.uvd[lo:hi])

    def fused(lo, hi, n_mean):
        H, uvd, gz, gD, gwp, lp = ops.decoder_fused_raw(z[lo:hi], w, D[lo:hi], batch.label_img[lo:hi], batch.mask[lo:hi],
                                                        tg(lo, hi), "softmax", alpha, n_mean=n_mean)
        return gz, gD, ops.reduce_partials(gwp), ops.stage_loss(lp, 1.0, 0.01, alpha, n_mean)

    def two_kernel(lo, hi, n_mean):
        _, uvd, st, _ = ops.decoder_forward_raw(z[lo:hi], w, D[lo:hi], batch.label_img[lo:hi], batch.mask[lo:hi])
        gz, gD, gwp, lp = ops.decoder_backward_raw(z[lo:hi], w, D[lo:hi], batch.label_img[lo:hi], batch.mask[lo:hi], st,
                                                   uvd, targets=tg(lo, hi), alpha=alpha, n_mean=n_mean, want_loss=True)
        return gz, gD, ops.reduce_partials(gwp), ops.stage_loss(lp, 1.0, 0.01, alpha, n_mean)

    for route in (fused, two_kernel):
        full = route(0, B, 0)
        a, b = route(0, B // 2, B * J), route(B // 2, B, B * J)
        torch.cuda.synchronize()
        assert_close("gz", torch.cat([a[0], b[0]]).cpu().numpy(), full[0].cpu().numpy(), 1e-6)
        assert_close("gD", torch.cat([a[1], b[1]]).cpu().numpy(), full[1].cpu().numpy(), 1e-6)
        assert_close("gw", (a[2] + b[2]).cpu().numpy(), full[2].cpu().numpy(), 1e-5)
        assert_close("loss", (a[3] + b[3]).cpu().numpy(), full[3].cpu().numpy(), 1e-6)
        # and n_mean = 0 on a half batch means "this call's own B*J": twice the contribution
        h = route(0, B // 2, 0)
        assert_close("own mean", h[0].cpu().numpy(), (2 * a[0]).cpu().numpy(), 1e-6)
    # public criterion
    zz, DD, ww = (t.clone().requires_grad_(True) for t in (z[:B // 2], D[:B // 2], w))
    tot, terms, _, _ = ops.fused_decoder_loss(zz, ww, DD, batch.label_img[:B // 2], batch.mask[:B // 2],
                                              batch.heatmaps[:B // 2], batch.depthmaps[:B // 2], batch.uvd[:B // 2],
                                              alpha=alpha, n_mean=B * J)
    tot.backward()
    ref = fused(0, B // 2, B * J)
    assert_close("criterion gz", zz.grad.cpu().numpy(), ref[0].cpu().numpy(), 1e-6)
    assert_close("criterion loss", tot.detach().cpu().numpy(), ref[3][3].cpu().numpy(), 1e-6)


FUSED_INSTANCES = [(m, t, dt) for m in ("softmax", "sum") for t in ("dense", "sparse")
                   for dt in (torch.float32, torch.float16, torch.bfloat16)]


@pytest.mark.parametrize("method,targets_kind,dtype", FUSED_INSTANCES)
def test_one_pass_kernel_every_instantiation_many_items_per_cta(method, targets_kind, dtype):
    """All 12 instantiations of decoder_fused_kernel with ~7 items per CTA (grid = 296), so that every slot of
    the six-slot ring is re-used several times and sample boundaries fall inside a CTA's range; compared with
    the forward kernel followed by the backward+loss kernel.  tools/sanitize.sh runs this test under
    compute-sanitizer racecheck / memcheck / initcheck."""
    shape = synth.NYU
    B, J, alpha = 300, 7, 0.5
    d = synth.make_frames_device(shape, B, seed=21, device=DEV)
    batch = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"][:, :J].contiguous(), fx=shape.fx, fy=shape.fy,
                          targets="both")
    g = torch.Generator(device=DEV).manual_seed(9)
    z = (torch.randn(B, J, 64, 64, device=DEV, generator=g) * 2).to(dtype)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g).to(dtype)
    w = (torch.rand(J, 1, device=DEV, generator=g) + 0.5) if method == "softmax" else None
    targets = ((batch.heatmaps, batch.depthmaps, batch.uvd) if targets_kind == "dense"
               else ops.SparseTargets(batch.taps, batch.uvd))
    L, m = batch.label_img, batch.mask
    # float16 gradients of a mean over B*J are subnormal at unit loss scale (one float16 step = 6e-8 is 5 % of the
    # largest of them): run both routes with GradScaler's 2^16, as ops.fused_decoder_loss does
    ls = 65536.0 if dtype == torch.float16 else 1.0
    H1, uvd1, gz1, gD1, gw1, lp1 = ops.decoder_fused_raw(z, w, D, L, m, targets, method, alpha, loss_scale=ls)
    H2, uvd2, st2, _ = ops.decoder_forward_raw(z, w, D, L, m, method)
    gz2, gD2, gw2, lp2 = ops.decoder_backward_raw(z, w, D, L, m, st2, uvd2, method=method, targets=targets, alpha=alpha,
                                                  loss_scale=ls, want_loss=True)
    torch.cuda.synchronize()
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert_close("H", H1.cpu().numpy(), H2.cpu().numpy(), 2e-6)
    assert_close("uvd", uvd1.cpu().numpy(), uvd2.cpu().numpy(), 2e-6)
    assert_close("loss partials", lp1.cpu().numpy(), lp2.cpu().numpy(), 1e-5)
    assert_close("gz", gz1.float().cpu().numpy(), gz2.float().cpu().numpy(), tol)
    assert_close("gD", gD1.float().cpu().numpy(), gD2.float().cpu().numpy(), tol)
    if method == "softmax":
        assert_close("gw", ops.reduce_partials(gw1).cpu().numpy(), ops.reduce_partials(gw2).cpu().numpy(), 2e-5)
    # run-to-run determinism (a race would show up here as well)
    again = ops.decoder_fused_raw(z, w, D, L, m, targets, method, alpha, loss_scale=ls)
    for x, y in zip((H1, uvd1, gz1, gD1, gw1, lp1), again):
        assert (x is None and y is None) or torch.equal(x, y)


def test_wrappers_take_noncontiguous_and_float64_arguments():
    """converted copies of label / mask / box / cube must stay alive (and distinct) until the launch."""
    B, J = 6, 4
    g = torch.Generator(device=DEV).manual_seed(2)
    z = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    w = torch.rand(J, 1, device=DEV, generator=g) + 0.5
    Lm = torch.rand(B, 2, 64, 64, device=DEV, generator=g)
    Lm[:, 1] = (Lm[:, 1] < 0.4).float()
    L_nc, m_nc = Lm[:, 0:1], Lm[:, 1:2]                      # non-contiguous views of one buffer
    assert not L_nc.is_contiguous() and not m_nc.is_contiguous()
    L_c, m_c = L_nc.contiguous(), m_nc.contiguous()
    _, uvd, st, _ = ops.decoder_forward_raw(z, w, D, L_c, m_c)
    g_uvd = torch.randn(B, J, 3, device=DEV, generator=g)
    ref = ops.decoder_backward_raw(z, w, D, L_c, m_c, st, uvd, g_uvd)
    got = ops.decoder_backward_raw(z, w, D, L_nc.double(), m_nc.double(), st, uvd, g_uvd)
    tg = (torch.rand(B, J, 64, 64, device=DEV, generator=g) * 0.01, torch.randn(B, J, 64, 64, device=DEV, generator=g),
          torch.rand(B, J, 3, device=DEV, generator=g))
    ref_l = ops.decoder_backward_raw(z, w, D, L_c, m_c, st, uvd, targets=tg, alpha=0.5, want_loss=True)
    got_l = ops.decoder_backward_raw(z, w, D, L_nc, m_nc, st, uvd, targets=tg, alpha=0.5, want_loss=True)
    torch.cuda.synchronize()
    for a, b in list(zip(ref, got)) + list(zip(ref_l, got_l)):
        assert (a is None and b is None) or torch.equal(a, b)
    # recover_uvd / joint_error with float64 box / cube / com
    uvd_n = torch.rand(B, J, 3, device=DEV, generator=g) - 0.5
    uvd_t = torch.rand(B, J, 3, device=DEV, generator=g) - 0.5
    box = torch.rand(B, device=DEV, generator=g) * 100 + 150
    cube = torch.full((B,), 150.0, device=DEV)
    com = torch.rand(B, 3, device=DEV, generator=g) * 300 + 200
    intr = (588.0, 587.0, 320.0, 240.0)
    a = ops.recover_uvd(uvd_n, box, com, cube, intr)
    b = ops.recover_uvd(uvd_n.double(), box.double(), com.double(), cube.double(), intr)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    ea = ops.joint_error(uvd_n, uvd_t, box, com, cube, intr)
    eb = ops.joint_error(uvd_n.double(), uvd_t.double(), box.double(), com.double(), cube.double(), intr)
    assert torch.equal(ea, eb)


def test_fp16_last_stage_keeps_small_gradients_under_a_loss_scale():
    """with float16 conv outputs the eagerly computed last-stage gradients are 2/(B*J) * p * err,
    typically 1e-6 .. 1e-10: below float16's normal range at unit loss scale.  With GradScaler's scale (2^16)
    applied the way the reference applies it (before backward), they survive; so must ours."""
    B, J = 64, 14
    scale = 65536.0
    g = torch.Generator(device=DEV).manual_seed(3)
    z32 = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 0.3            # flat maps: p ~ 1/4096
    D32 = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 0.1
    w = (torch.rand(J, 1, device=DEV, generator=g) + 0.5).requires_grad_(True)
    m = (torch.rand(B, 1, 64, 64, device=DEV, generator=g) < 0.4).float()
    L = torch.rand(B, 1, 64, 64, device=DEV, generator=g) * m
    heat = torch.softmax(torch.randn(B, J, 4096, device=DEV, generator=g), 2).view(B, J, 64, 64)
    dm = torch.randn(B, J, 64, 64, device=DEV, generator=g) * m
    uv = torch.rand(B, J, 3, device=DEV, generator=g) - 0.5
    z = z32.half().requires_grad_(True)
    D = D32.half().requires_grad_(True)
    total, _, _ = ops.fused_decoder_loss(z, w, D, L, m, heat, dm, uv, alpha=1.0, store_heat=False)
    (total * scale).backward()
    # reference semantics: float32 autograd on the same (float16-rounded) values, scaled loss, gradients cast to half
    zr = z.detach().float().requires_grad_(True)
    Dr = D.detach().float().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    p, Dm, uvd = do.decoder_forward(zr, wr, Dr, L, m)
    loss = do.combine_losses(do.stage_losses(p, Dm, uvd, heat, dm, uv), 1.0)
    (loss * scale).backward()
    ref_gz = zr.grad.half().float()
    got_gz = z.grad.float()
    assert float(ref_gz.abs().max()) > 0
    nz = ref_gz != 0
    # what the reference keeps, we keep: same non-zero pattern up to float16 rounding of borderline values
    kept = float((got_gz[nz] != 0).float().mean())
    assert kept > 0.999, kept
    assert_close("scaled gz", got_gz.cpu().numpy(), zr.grad.cpu().numpy(), 2e-3)
    assert_close("scaled gD", D.grad.float().cpu().numpy(), Dr.grad.cpu().numpy(), 2e-3)
    assert_close("scaled gw", w.grad.cpu().numpy(), wr.grad.cpu().numpy(), 1e-3)
    # magnitude check: unscaled, most of these gradients would not be representable in float16
    unscaled = zr.grad / scale
    assert float((unscaled.abs() < 6e-8).float().mean()) > 0.3


# --------------------------------------------------------------------------- #
# two ranks over NCCL through the CUDA kernels
# --------------------------------------------------------------------------- #
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        shape = synth.NYU
        B, J, alpha = 8, shape.joints, 0.5
        d = synth.make_frames(shape, B, seed=4)
        from pixelwiseregression_b200 import distributed as pd
        lo, hi = pd.shard_range(B, rank, world)
        batch = sfr.build_sfr(torch.from_numpy(d["frames"][lo:hi]).to(dev), d["com"][lo:hi], d["cube"][lo:hi],
                              d["uvd"][lo:hi], fx=shape.fx, fy=shape.fy)
        g = torch.Generator(device="cpu").manual_seed(6)
        z = torch.randn(B, J, 64, 64, generator=g)[lo:hi].to(dev).requires_grad_(True)
        D = torch.randn(B, J, 64, 64, generator=g)[lo:hi].to(dev).requires_grad_(True)
        w = (torch.rand(J, 1, generator=g) + 0.5).to(dev).requires_grad_(True)
        # SUM-reduced gradients need the global mean inside the kernel: n_mean = B*J
        total, terms, _ = ops.fused_decoder_loss(z, w, D, batch.label_img, batch.mask, batch.heatmaps, batch.depthmaps,
                                                 batch.uvd, alpha=alpha, store_heat=False, n_mean=B * J)
        total.backward()
        vec = torch.cat([w.grad.reshape(-1), total.detach().reshape(1)])
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        gz_all = [torch.empty_like(z.grad) for _ in range(world)]
        dist.all_gather(gz_all, z.grad)
        if rank == 0:
            torch.save({"vec": vec.cpu(), "gz": torch.cat(gz_all).cpu()}, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_nccl_ranks_with_n_mean_equal_one_rank_global_batch(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r0.pt")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    shape = synth.NYU
    B, J, alpha = 8, shape.joints, 0.5
    d = synth.make_frames(shape, B, seed=4)
    batch = sfr.build_sfr(torch.from_numpy(d["frames"]).to(DEV), d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy)
    g = torch.Generator(device="cpu").manual_seed(6)
    z = torch.randn(B, J, 64, 64, generator=g).to(DEV).requires_grad_(True)
    D = torch.randn(B, J, 64, 64, generator=g).to(DEV).requires_grad_(True)
    w = (torch.rand(J, 1, generator=g) + 0.5).to(DEV).requires_grad_(True)
    total, _, _ = ops.fused_decoder_loss(z, w, D, batch.label_img, batch.mask, batch.heatmaps, batch.depthmaps,
                                         batch.uvd, alpha=alpha, store_heat=False)
    total.backward()
    ref = torch.cat([w.grad.reshape(-1), total.detach().reshape(1)]).cpu()
    assert_close("dL/dw + loss over 2 ranks", got["vec"].numpy(), ref.numpy(), 1e-5)
    assert_close("gz over 2 ranks", got["gz"].numpy(), z.grad.cpu().numpy(), 1e-6)


@pytest.mark.parametrize("method,dtype", [("softmax", torch.float32), ("sum", torch.float32), ("softmax", torch.bfloat16)])
@pytest.mark.parametrize("targets_kind", ["dense", "sparse"])
def test_inner_stage_forward_loss_through_the_one_pass_kernel_equals_the_direct_kernel(method, dtype, targets_kind):
    """pwr_decoder_fwd with targets (an inner stage: heat maps stored, loss value at forward time, stats saved for the
    backward) runs in the one-pass kernel without gradient outputs; the one-CTA-per-item kernel (dispatch option)
    must give the same heat maps, coordinates, stats and loss partials, and the saved stats must drive the backward."""
    from pixelwiseregression_b200 import _lib
    shape = synth.NYU
    B, J = 61, 9                                            # 549 items over 296 CTAs: ragged ranges, slot re-use
    d = synth.make_frames_device(shape, B, seed=33, device=DEV)
    batch = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"][:, :J].contiguous(), fx=shape.fx, fy=shape.fy,
                          targets="both")
    g = torch.Generator(device=DEV).manual_seed(4)
    z = (torch.randn(B, J, 64, 64, device=DEV, generator=g) * 2).to(dtype)
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g).to(dtype)
    w = None
    if method == "softmax":
        w = torch.rand(J, 1, device=DEV, generator=g) + 0.5
        w[::4] *= -1
    targets = ((batch.heatmaps, batch.depthmaps, batch.uvd) if targets_kind == "dense"
               else ops.SparseTargets(batch.taps, batch.uvd))
    L, m = batch.label_img, batch.mask
    with _lib.option("fwd_direct", 1):
        Hd, uvd_d, st_d, lp_d = ops.decoder_forward_raw(z, w, D, L, m, method, targets=targets)
    Hp, uvd_p, st_p, lp_p = ops.decoder_forward_raw(z, w, D, L, m, method, targets=targets)
    torch.cuda.synchronize()
    assert torch.equal(st_d[..., 0], st_p[..., 0])          # extremum: exact
    assert_close("heat", Hp.cpu().numpy(), Hd.cpu().numpy(), 2e-6)
    assert_close("uvd", uvd_p.cpu().numpy(), uvd_d.cpu().numpy(), 2e-6)
    assert_close("stats", st_p.cpu().numpy(), st_d.cpu().numpy(), 2e-6)
    assert_close("loss partials", lp_p.cpu().numpy(), lp_d.cpu().numpy(), 1e-5)
    gH = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 1e-3
    a = ops.decoder_backward_raw(z, w, D, L, m, st_p, uvd_p, None, gH, None, method, targets, 0.5)
    b = ops.decoder_backward_raw(z, w, D, L, m, st_d, uvd_d, None, gH, None, method, targets, 0.5)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert_close("gz", a[0].float().cpu().numpy(), b[0].float().cpu().numpy(), tol)
    assert_close("gD", a[1].float().cpu().numpy(), b[1].float().cpu().numpy(), tol)


@pytest.mark.parametrize("method", ["softmax", "sum"])
@pytest.mark.parametrize("B,J", [(1, 1), (2, 3), (300, 14)])
def test_lean_one_pass_kernel_equals_the_two_cta_one_pass_kernel(method, B, J):
    """Without a dense target map to visit (compact targets, or the uvd term alone) and float32 logits,
    pwr_decoder_fwd_bwd_loss runs decoder_fused_lean_kernel: three CTAs per SM, gp / p parked in the ring slots of
    z / D instead of in registers.  Same arithmetic and summation order as decoder_fused_kernel (option
    `fused_no_lean`) up to the compiler's choice of fused multiply-adds (measured: identical or 1 ulp apart), and
    bitwise reproducible from run to run; B = 300, J = 14 gives ~9.5 items per CTA at grid 444
    (each of the two slot pairs re-used four times, sample boundaries inside a CTA's range); (1, 1) and (2, 3)
    leave most of the ring unused.  tools/sanitize.sh runs this test under racecheck / memcheck / initcheck."""
    from pixelwiseregression_b200 import _lib
    shape = synth.NYU
    d = synth.make_frames_device(shape, B, seed=33 + B, device=DEV)
    batch = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"][:, :J].contiguous(), fx=shape.fx, fy=shape.fy,
                          targets="both")
    g = torch.Generator(device=DEV).manual_seed(B + J)
    z = torch.randn(B, J, 64, 64, device=DEV, generator=g) * 2
    D = torch.randn(B, J, 64, 64, device=DEV, generator=g)
    w = None
    if method == "softmax":
        w = torch.rand(J, 1, device=DEV, generator=g) + 0.5
        w[::3] *= -1.0                                  # negative temperatures take the minimum as the extremum
    L, m = batch.label_img, batch.mask
    compact = ops.SparseTargets(batch.taps, batch.uvd)
    dense = (batch.heatmaps, batch.depthmaps, batch.uvd)
    cases = [(compact, dict(alpha=0.5)), (compact, dict(alpha=1.0, store_heat=False)),
             (compact, dict(alpha=0.3, want_grads=False)),            # forward + loss only (gz = gD = NULL)
             (dense, dict(alpha=1.0, want_loss=False)),               # uvd term only: the target maps are not visited
             (compact, dict(alpha=1.0, want_loss=False, store_heat=False))]
    for targets, kw in cases:
        lean = ops.decoder_fused_raw(z, w, D, L, m, targets, method, **kw)
        again = ops.decoder_fused_raw(z, w, D, L, m, targets, method, **kw)
        with _lib.option("fused_no_lean", 1):
            ref = ops.decoder_fused_raw(z, w, D, L, m, targets, method, **kw)
        torch.cuda.synchronize()
        for name, x, y, x2 in zip(("H", "uvd", "gz", "gD", "gw_partial", "loss_partial"), lean, ref, again):
            assert (x is None) == (y is None), name
            if x is not None:
                assert_close("%s %s" % (name, kw), x.cpu().numpy(), y.cpu().numpy(), 2e-6)
                assert torch.equal(x, x2), ("run-to-run", name, kw)
    # and against the float64 oracle (uvd-only loss, alpha = 1)
    H, uvd, gz, gD, gwp, _ = ops.decoder_fused_raw(z, w, D, L, m, compact, method, alpha=1.0)
    t64 = lambda a: a.detach().cpu().to(torch.float64)
    w64 = t64(w) if w is not None else None
    p_ref, _, uvd_ref = do.decoder_forward(t64(z), w64, t64(D), t64(L), t64(m), method)
    gz_ref, gD_ref, gw_ref = do.decoder_backward(t64(z), w64, t64(D), t64(L), t64(m), torch.zeros(B, J, 3, dtype=torch.float64),
                                                 None, None, method, targets=tuple(t64(a) for a in dense), alpha=1.0)
    assert_close("H vs oracle", H.cpu().numpy(), p_ref.numpy())
    assert_close("uvd vs oracle", uvd.cpu().numpy(), uvd_ref.numpy())
    assert_close("gz vs oracle", gz.cpu().numpy(), gz_ref.numpy(), GRAD_RTOL)
    assert_close("gD vs oracle", gD.cpu().numpy(), gD_ref.numpy(), GRAD_RTOL)
    if method == "softmax":
        assert_close("gw vs oracle", ops.reduce_partials(gwp).view(-1, 1).cpu().numpy(), gw_ref.numpy(), GRAD_RTOL)
