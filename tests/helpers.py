"""Shared test helpers: golden loading and the tolerance rules of SURVEY §8(d)."""
import os

import numpy as np

from pixelwiseregression_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SFR_FIELDS = ("img", "label_img", "mask", "box_size", "cube_size", "com", "uvd", "heatmaps", "dmap")
# north_star: crop ints and masks bit-exact; maps and coordinates 1e-5 relative
EXACT_FIELDS = ("mask", "box_size", "cube_size", "com")
RTOL = 1e-5
GRAD_RTOL = 1e-4


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_shape(g):
    return synth.SHAPES[str(g["shape_name"])]


def assert_close(name, got, ref, rtol=RTOL):
    """|got-ref| <= rtol*|ref| + rtol*max|ref| (per tensor), as in SURVEY §8(d)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    if ref.size == 0:
        return
    atol = rtol * float(np.abs(ref).max())
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + atol
    bad = err > bound
    assert not bad.any(), "%s: %d/%d outside tolerance, max err %.3e (max|ref| %.3e)" % (
        name, int(bad.sum()), ref.size, float(err.max()), float(np.abs(ref).max()))


def assert_sfr_matches(got, ref, names, valid_ref, prefix=""):
    """Compare two SFR result dicts on the samples the reference accepts."""
    valid_ref = np.asarray(valid_ref).astype(bool)
    assert (np.asarray(got["valid"]).astype(bool) == valid_ref).all(), (
        prefix, got["valid"].tolist(), valid_ref.astype(int).tolist())
    sel = valid_ref
    for n in names:
        a, b = np.asarray(got[n])[sel], np.asarray(ref[n])[sel]
        if n in EXACT_FIELDS:
            assert (a == b).all(), "%s%s not bit-exact" % (prefix, n)
        else:
            assert_close(prefix + n, a, b)
    if "heatmaps" in names:
        # the Dmap support is (heatmap > 0) & mask: discontinuous, must be exact
        a, b = np.asarray(got["heatmaps"])[sel], np.asarray(ref["heatmaps"])[sel]
        assert ((a > 0) == (b > 0)).all(), prefix + "heat-map support differs"
        a, b = np.asarray(got["dmap"])[sel], np.asarray(ref["dmap"])[sel]
        assert ((a != 0) == (b != 0)).all(), prefix + "Dmap support differs"


def load_model_golden(name="model_nyu_eval"):
    """Reference PixelwiseRegression weights + inputs + eval outputs (oracle/make_golden.py:golden_model).
    Returns (golden dict, drop-in model with the reference's weights loaded, strict)."""
    import torch
    from pixelwiseregression_b200 import model as M
    g = load_golden(name)
    net = M.PixelwiseRegression(int(g["joints"]), stage=int(g["stages"]), features=int(g["features"]),
                                level=int(g["level"]), norm_method="instance")
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd_")}
    net.load_state_dict(sd, strict=True)          # the reference's own keys and shapes, nothing missing or extra
    net.eval()
    return g, net
