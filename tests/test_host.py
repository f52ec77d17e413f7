"""CPU: host-side logic that needs no GPU — drop-in module layout, synthetic
generators, refusal of CPU tensors (no silent fallback)."""
import json
import os

import numpy as np
import pytest
import torch

from pixelwiseregression_b200 import model, ops, sfr, synth
from pixelwiseregression_b200._lib import PwrError
from helpers import GOLDEN


def test_state_dict_layout_equals_reference():
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    for name, s in spec.items():
        m = getattr(model, s.get("class", "PixelwiseRegression"))(**s["kwargs"])
        got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert got == s["keys"], name
    assert len(spec["nyu_instance"]["keys"]) == 344
    assert spec["fullregression_nyu"]["class"] == "FullRegression"     # the ablation modules are re-exported too


def test_constructor_signatures_match_reference():
    import inspect
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]
    assert sig(model.PixelwiseRegression) == ["joints", "stage", "label_size", "features", "level", "kernel_size",
                                              "norm_method", "heatmap_method"]
    assert sig(model.PredictionBlock) == ["in_dim", "joints", "label_size", "features", "level", "kernel_size", "norm",
                                          "heatmap_method"]
    assert sig(model.PlaneRegression) == ["features", "joints", "label_size", "kernel_size", "norm", "inplace",
                                          "normalization_method"]
    assert sig(model.DepthRegression) == ["features", "joints", "kernel_size", "norm", "inplace"]
    assert list(inspect.signature(model.PixelwiseRegression.forward).parameters)[1:] == ["img", "label_img", "mask"]
    assert list(inspect.signature(model.DepthRegression.forward).parameters)[1:] == ["f", "heatmaps", "label_img", "mask"]


def test_com_filter_definition():
    f = model.com_filter(64, 64)
    assert f.dtype == torch.float32 and f.shape == (2, 64, 64)
    assert f[0, 3, 40].item() == np.float32((40 - 32) / 63) and f[1, 3, 40].item() == np.float32((3 - 32) / 63)


def test_cpu_tensors_raise_instead_of_falling_back():
    m = model.PixelwiseRegression(14, stage=1, features=32, level=1, norm_method="instance")
    with pytest.raises(PwrError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 128, 128), torch.zeros(1, 1, 64, 64), torch.zeros(1, 1, 64, 64))
    with pytest.raises(PwrError):
        sfr.build_sfr(torch.zeros(1, 480, 640), np.zeros((1, 3)), 150.0, np.zeros((1, 14, 3)), fx=1.0, fy=1.0)
    with pytest.raises(PwrError):
        ops.reduce_partials(torch.zeros(4, 3))


def test_synthetic_generators_are_deterministic():
    a = synth.make_frames(synth.MSRA, 2, 5)
    b = synth.make_frames(synth.MSRA, 2, 5)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert a["frames"].shape == (2, 240, 320) and a["uvd"].shape == (2, 21, 3)
    d = synth.make_frames_device(synth.NYU, 3, 1, device="cpu")
    assert d["frames"].shape == (3, 480, 640) and d["uvd"].dtype == torch.float64
    frac = float((d["frames"] > 0).float().mean())
    assert 0.01 < frac < 0.5


def test_window_size_covers_the_oracle_crop_geometry():
    """sfr.window_size (host side of the PCIe feed) must be at least the in-frame extent of every crop box the
    oracle derives (datasets.py:306-311, utils.py:167-173), plus 16-byte alignment slack per row."""
    from oracle import sfr_oracle as so
    for shape, fmt in ((synth.NYU, "nyu_gb16"), (synth.HAND17, "u16"), (synth.ICVL, "f32"), (synth.MSRA, "f32")):
        d = synth.make_frames(shape, 64, seed=9)
        d["com"][0, :2] = (1.5, 2.5)
        d["com"][1, 2] = 120.0                       # a box larger than the frame
        win_h, win_w = sfr.window_size(d["com"], d["cube"], shape.fx, shape.fy, shape.height, shape.width, fmt)
        per16 = 4 if fmt == "f32" else 8
        assert win_w % per16 == 0 and win_h <= shape.height
        for b in range(64):
            box = so.crop_box(d["com"][b, 2], d["cube"][b], shape.fx, shape.fy)
            r0, c0, shift, nrows, ncols, rs, cs = so.crop_geometry(d["com"][b], box, shape.height, shape.width)
            fr0, fc0 = rs - shift, cs - shift
            rows = max(0, min(fr0 + nrows, shape.height) - max(fr0, 0))
            ca, cb = max(fc0, 0), min(fc0 + ncols, shape.width)
            cols = max(0, -(-cb // per16) * per16 - ca // per16 * per16)
            assert rows <= win_h and cols <= win_w, (shape.name, b, rows, cols, win_h, win_w)


def test_host_feed_and_fetch_refuse_to_run_without_cuda():
    from pixelwiseregression_b200 import feed
    if not torch.cuda.is_available():
        with pytest.raises(PwrError, match="no CPU fallback"):
            feed.HostFeed(synth.NYU, 4)
    with pytest.raises(PwrError):
        sfr.fetch_windows(torch.zeros(1, 480, 640), torch.zeros(1, 3), torch.zeros(1), fx=1.0, fy=1.0)
    with pytest.raises(PwrError, match="kernel_size=7"):
        sfr.build_sfr(torch.zeros(1, 480, 640), np.zeros((1, 3)), 150.0, np.zeros((1, 14, 3)), fx=1.0, fy=1.0, kernel_size=9)


def test_both_bench_arms_describe_the_same_workload():
    """bench.py's two arms print `config` from one function with the same arguments, so the driver's same_config
    check compares like with like (arm-specific remarks live outside it)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from pixelwiseregression_b200 import roofline, synth
    a = bench.workload_config(synth.NYU, 4096, "f32", False, 1.0, 1.0, 0.01)
    b = bench.workload_config(synth.NYU, 4096, "f32", False, 1.0, 1.0, 0.01)
    assert a == b and a["algorithmic_bytes_per_sample"] == roofline.step_one_pass_bytes(14) == 2261404
    assert "larger than L2" in a["l2"] and a["batch_per_gpu"] == 4096 and a["joints"] == 14
