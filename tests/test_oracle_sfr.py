"""CPU: the SFR oracle against the reference-generated golden vectors, against
OpenCV itself, and (build container only) against the live reference."""
import numpy as np
import pytest

from oracle import sfr_oracle as so
from oracle import ref_shim
from pixelwiseregression_b200 import synth
from helpers import SFR_FIELDS, assert_close, assert_sfr_matches, golden_shape, load_golden

GOLDEN_SETS = ["sfr_nyu", "sfr_nyu_test_only", "sfr_hand17", "sfr_msra", "sfr_icvl", "sfr_edge"]


def run_oracle_on_golden(g, backend="numpy"):
    shape = golden_shape(g)
    frames = g["frames"].astype(np.float64) if shape.frame_f64 else g["frames"]
    com = None if shape.com_from_frame else g["com"]
    return so.process_batch(frames, g["uvd"], com, g["cube"], shape.fx, shape.fy,
                            test_only=bool(g["test_only"]), backend=backend)


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    got = run_oracle_on_golden(g)
    names = SFR_FIELDS[:6] if bool(g["test_only"]) else SFR_FIELDS
    ref = {n: g["ref_" + n] for n in names}
    assert_sfr_matches(got, ref, names, g["ref_valid"], prefix=name + ":")


RAW_SETS = ["sfr_nyu_raw", "sfr_nyu_raw_val", "sfr_hand17_raw", "sfr_icvl_raw"]


def oracle_on_raw_golden(g, backend="numpy"):
    """Raw sensor samples -> decode -> load_from_text prefilter -> SFR, all oracle code."""
    shape = golden_shape(g)
    fmt, margin = str(g["frame_format"]), float(g["margin"])
    frames = []
    for b in range(len(g["raw"])):
        raw = g["raw"][b]
        if fmt == "nyu_gb16":
            rgb = np.zeros(raw.shape + (3,), np.uint8)
            rgb[..., 1] = raw >> 8
            rgb[..., 2] = raw & 255
            img = so.decode_nyu(rgb)
        else:
            img = so.decode_u16(raw)
        cube = int(g["cube"][b])
        frames.append(so.prefilter(img, g["com"][b], cube, shape.fx, shape.fy, shape.halfu, shape.halfv, int(margin)))
    return so.process_batch(frames, g["uvd"], g["com"], g["cube"], shape.fx, shape.fy, backend=backend)


@pytest.mark.parametrize("name", RAW_SETS)
def test_oracle_raw_frames_match_reference_load_from_text(name):
    g = load_golden(name)
    got = oracle_on_raw_golden(g)
    ref = {n: g["ref_" + n] for n in SFR_FIELDS}
    assert_sfr_matches(got, ref, SFR_FIELDS, g["ref_valid"], prefix=name + ":")
    assert g["ref_valid"].all()


def test_u16_png_decode_is_the_identity_and_nyu_decode_is_not():
    u = np.arange(65536, dtype=np.uint32).astype(np.uint16).reshape(256, 256)
    assert (so.decode_u16(u) == u.astype(np.float32)).all()
    rgb = np.zeros((1, 1, 3), np.uint8)
    rgb[..., 1], rgb[..., 2] = 2, 238
    assert so.decode_nyu(rgb)[0, 0] == np.float32(750.00006)       # float32 rounding of the reference's decode line


AUG_SETS = ["sfr_nyu_aug", "sfr_nyu_aug_fallback", "sfr_msra_aug"]


def oracle_on_aug_golden(g, backend="numpy"):
    shape = golden_shape(g)
    frames = g["frames"].astype(np.float64) if shape.frame_f64 else g["frames"]
    com = None if shape.com_from_frame else g["com"]
    return so.process_batch(frames, g["uvd"], com, g["cube"], shape.fx, shape.fy, backend=backend, aug=g["aug"])


@pytest.mark.parametrize("name", AUG_SETS)
def test_oracle_augmented_branch_matches_reference(name):
    """datasets.py:216-299 with train.py's default flags; the reference's random draws were
    recorded (oracle/make_golden.py) and are replayed here."""
    g = load_golden(name)
    got = oracle_on_aug_golden(g)
    ref = {n: g["ref_" + n] for n in SFR_FIELDS}
    assert_sfr_matches(got, ref, SFR_FIELDS, g["ref_valid"], prefix=name + ":")
    if name == "sfr_nyu_aug_fallback":
        # at least one sample raised inside the augmented branch and silently fell back
        plain = so.process_batch(g["frames"], g["uvd"], g["com"], g["cube"], golden_shape(g).fx, golden_shape(g).fy)
        same = [np.abs(plain["img"][b] - g["ref_img"][b]).max() < 1e-6 for b in range(len(g["frames"]))]
        assert any(same) and not all(same), same


def test_warp_affine_restatement_is_bit_exact_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        for _ in range(6):
            src = (rng.uniform(-100, 100, (128, 128)) * (rng.uniform(size=(128, 128)) < 0.7)).astype(dt)
            angle, scale = rng.uniform(-30, 30), rng.uniform(0.8, 1.2)
            M = cv2.getRotationMatrix2D((64, 64), angle, scale)
            assert (M == so.rotation_matrix(angle, scale)).all()
            assert (so.warp_affine(src, M) == cv2.warpAffine(src, M, (128, 128))).all()


def test_edge_golden_covers_reject_and_accept():
    g = load_golden("sfr_edge")
    v = g["ref_valid"]
    assert v.sum() >= 8 and (v == 0).sum() >= 4


def test_resize_restatement_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for n in (28, 128, 150, 176, 234, 353, 880):
        for dt in (np.float32, np.float64):
            src = (rng.uniform(-100, 100, (n, n)) * (rng.uniform(size=(n, n)) < 0.7)).astype(dt)
            ref = cv2.resize(src, (128, 128))
            got = so.resize_bilinear(src)
            assert got.dtype == ref.dtype
            assert_close("resize %d %s" % (n, dt.__name__), got, ref, rtol=2e-7 if dt is np.float32 else 1e-14)
            assert ((got != 0) == (ref != 0)).all()
            half = cv2.resize(ref, (64, 64))
            assert_close("half", so.resize_half(ref), half, rtol=2e-7 if dt is np.float32 else 1e-14)
            assert ((so.resize_half(ref) != 0) == (half != 0)).all()
    # non-square source (truncated crop)
    src = rng.uniform(-50, 50, (200, 131)).astype(np.float32)
    assert_close("nonsquare", so.resize_bilinear(src), cv2.resize(src, (128, 128)), rtol=2e-7)


def test_blur_restatement_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    k = cv2.getGaussianKernel(7, 1.5).ravel()
    assert (k == so.gaussian_kernel7()).all()
    x = np.arange(7) - 3.0
    closed = np.exp(-x * x / (2 * 1.5 * 1.5))
    assert np.abs(closed / closed.sum() - k).max() < 1e-16
    for _ in range(50):
        u, v = rng.uniform(-1, 62.99, 2)
        h = so.splat4(u, v)
        ref = cv2.GaussianBlur(h, (7, 7), 1.5)
        got = so.gaussian_blur7(h)
        assert np.abs(got - ref).max() < 1e-15
        assert ((got > 0) == (ref > 0)).all()


def test_splat_is_centre_of_mass_preserving():
    rng = np.random.default_rng(2)
    for _ in range(100):
        u, v = rng.uniform(0, 62.99, 2)
        h = so.splat4(u, v)
        yy, xx = np.mgrid[0:64, 0:64]
        assert abs(h.sum() - 1) < 1e-12
        assert abs((h * xx).sum() - u) < 1e-10 and abs((h * yy).sum() - v) < 1e-10
    with pytest.raises(IndexError):
        so.splat4(63.0, 10.0)


def test_recover_uvd_inverts_normalisation():
    g = load_golden("sfr_nyu")
    rec = so.recover_uvd(g["ref_uvd"], g["ref_box_size"], g["ref_com"], g["ref_cube_size"])
    assert np.abs(rec - g["uvd"]).max() < 2e-3  # float32 pixels / millimetres


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("shape_name,batch,seed", [("NYU", 6, 100), ("MSRA", 3, 101), ("HAND17", 3, 102)])
def test_oracle_matches_live_reference(shape_name, batch, seed):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from make_golden import run_reference_sfr
    _, _, datasets = ref_shim.load()
    shape = synth.SHAPES[shape_name]
    d = synth.make_frames(shape, batch, seed)
    frames = d["frames"].astype(np.float64) if shape.frame_f64 else d["frames"]
    ref = run_reference_sfr(datasets, shape, frames, d["uvd"], d["com"], d["cube"])
    got = so.process_batch(frames, d["uvd"], None if shape.com_from_frame else d["com"], d["cube"],
                           shape.fx, shape.fy)
    assert_sfr_matches(got, {n: ref["ref_" + n] for n in SFR_FIELDS}, SFR_FIELDS, ref["ref_valid"])


def test_hand17_bb_mode_matches_reference_golden():
    """HAND17 `process_mode='bb'` (datasets.py:199-206, 974-996): the oracle's bounding-box loader equals the
    reference's load_from_text_bb bit for bit (a float64 frame), and the test-only SFR with the CoM / cube
    fallback equals the reference's bb branch of process_single_data."""
    g = load_golden("sfr_hand17_bb")
    shape = golden_shape(g)
    frames = np.stack([so.load_bb(g["raw"][b], *g["boxes"][b]) for b in range(len(g["raw"]))])
    assert frames.dtype == np.float64 and np.array_equal(frames, g["ref_frames"])
    assert (frames[0] > 0).sum() < (so.decode_u16(g["raw"][0]) > 0).sum()          # the filter removed the clutter
    cube = np.full(len(frames), float(shape.cube))
    got = so.process_batch(frames, None, None, cube, shape.fx, shape.fy, test_only=True)
    ref = {n: g["ref_" + n] for n in SFR_FIELDS[:6]}
    assert_sfr_matches(got, ref, SFR_FIELDS[:6], np.ones(len(frames), np.uint8), prefix="bb:")
