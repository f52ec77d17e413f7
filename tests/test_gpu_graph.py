"""The library never synchronises, allocates or frees, so whole passes of the path capture
into a CUDA graph.  Replays must reproduce the call-by-call results bit for bit, also after the
inputs were overwritten in place (the graph holds pointers, not values)."""
import numpy as np
import pytest
import torch

from pixelwiseregression_b200 import ops, sfr, synth

pytestmark = pytest.mark.gpu


def _device_inputs(shape, batch, seed):
    d = synth.make_frames(shape, batch, seed)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    J = shape.joints
    return dict(frames=torch.from_numpy(d["frames"]).to(dev), com=torch.from_numpy(d["com"]).to(dev),
                cube=torch.from_numpy(d["cube"]).to(dev), uvd=torch.from_numpy(d["uvd"]).to(dev),
                z=torch.randn(batch, J, 64, 64, device=dev, generator=g),
                D=torch.randn(batch, J, 64, 64, device=dev, generator=g),
                w=torch.rand(J, 1, device=dev, generator=g) + 0.5)


def _capture(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    return graph, out


def test_inference_pass_replays_bit_exact():
    shape = synth.HAND17
    a, b = _device_inputs(shape, 16, 0), _device_inputs(shape, 16, 1)
    intr = (shape.fx, shape.fy, shape.halfu, shape.halfv)

    def one_pass(x):
        t = sfr.build_sfr(x["frames"], x["com"], x["cube"], fx=shape.fx, fy=shape.fy, test_only=True)
        _, uvd, _, _ = ops.decoder_forward_raw(x["z"], x["w"], x["D"], t.label_img, t.mask, store_heat=False,
                                               want_stats=False)
        return (t.img, t.valid) + tuple(ops.recover_uvd(uvd, t.box_size, t.com, t.cube_size, intrinsics=intr))

    graph, out = _capture(lambda: one_pass(a))
    for inputs in (a, b):
        for k in a:
            a[k].copy_(inputs[k])          # second round: new values behind the captured pointers
        graph.replay()
        ref = one_pass(a)
        torch.cuda.synchronize()
        for got, want in zip(out, ref):
            assert torch.equal(got, want)


def test_training_step_replays_bit_exact():
    """SFR build -> decoder forward -> fused backward + loss (the benchmark's step) as one graph."""
    shape = synth.NYU
    a, b = _device_inputs(shape, 12, 2), _device_inputs(shape, 12, 3)

    def step(x):
        t = sfr.build_sfr(x["frames"], x["com"], x["cube"], x["uvd"], fx=shape.fx, fy=shape.fy)
        H, uvd, stats, _ = ops.decoder_forward_raw(x["z"], x["w"], x["D"], t.label_img, t.mask)
        gz, gD, gw_partial, loss_partial = ops.decoder_backward_raw(
            x["z"], x["w"], x["D"], t.label_img, t.mask, stats, uvd, targets=(t.heatmaps, t.depthmaps, t.uvd),
            alpha=0.5, want_loss=True)
        gw = ops.reduce_partials(gw_partial)
        loss = ops.stage_loss(loss_partial, 1.0, 0.01, 0.5)
        return gz, gD, gw, loss, uvd, t.heatmaps

    graph, out = _capture(lambda: step(a))
    for inputs in (a, b):
        for k in a:
            a[k].copy_(inputs[k])
        graph.replay()
        ref = step(a)
        torch.cuda.synchronize()
        for got, want in zip(out, ref):
            assert torch.equal(got, want)
        assert np.isfinite(out[3].cpu().numpy()).all()
