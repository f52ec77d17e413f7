"""GPU parity: the CUDA SFR builder (through the C ABI) against the reference-
generated golden vectors, against the NumPy oracle on seeded inputs, and through
size-independent properties at the benchmark's full batch size."""
import numpy as np
import pytest
import torch

from oracle import sfr_oracle as so
from pixelwiseregression_b200 import sfr, synth
from helpers import SFR_FIELDS, assert_close, assert_sfr_matches, golden_shape, load_golden

pytestmark = pytest.mark.gpu

GOLDEN_SETS = ["sfr_nyu", "sfr_nyu_test_only", "sfr_hand17", "sfr_msra", "sfr_icvl", "sfr_edge"]


def run_gpu(frames, uvd, com, cube, shape, test_only=False):
    dev = torch.device("cuda:0")
    out = sfr.build_sfr(torch.from_numpy(frames).to(dev), None if com is None else com, cube,
                        None if test_only else uvd, fx=shape.fx, fy=shape.fy, frame_f64=shape.frame_f64,
                        test_only=test_only)
    torch.cuda.synchronize()
    d = {k: v.cpu().numpy() for k, v in out._asdict().items() if v is not None}
    if "depthmaps" in d:
        d["dmap"] = d.pop("depthmaps")
    return d


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_gpu_matches_reference_golden(name):
    g = load_golden(name)
    shape = golden_shape(g)
    test_only = bool(g["test_only"])
    got = run_gpu(g["frames"], g["uvd"], None if shape.com_from_frame else g["com"], g["cube"], shape, test_only)
    names = SFR_FIELDS[:6] if test_only else SFR_FIELDS
    ref = {n: g["ref_" + n] for n in names}
    assert_sfr_matches(got, ref, names, g["ref_valid"], prefix=name + ":")


@pytest.mark.parametrize("shape_name,batch,seed", [("NYU", 24, 0), ("HAND17", 8, 1), ("MSRA", 8, 2), ("ICVL", 8, 3)])
def test_gpu_matches_oracle_seeded(shape_name, batch, seed):
    shape = synth.SHAPES[shape_name]
    d = synth.make_frames(shape, batch, seed)
    frames64 = d["frames"].astype(np.float64) if shape.frame_f64 else d["frames"]
    com = None if shape.com_from_frame else d["com"]
    ref = so.process_batch(frames64, d["uvd"], com, d["cube"], shape.fx, shape.fy)
    got = run_gpu(d["frames"], d["uvd"], com, d["cube"], shape)
    assert_sfr_matches(got, ref, SFR_FIELDS, ref["valid"], prefix=shape_name + ":")
    # same operation order as the oracle -> the image path is bit-identical, not just close
    for n in ("img", "label_img"):
        assert (got[n] == ref[n]).all(), "%s differs bitwise from the oracle" % n
    assert ref["valid"].all()


RAW_SETS = ["sfr_nyu_raw", "sfr_nyu_raw_val", "sfr_hand17_raw", "sfr_icvl_raw"]


def gpu_on_raw(g):
    shape = golden_shape(g)
    out = sfr.build_sfr(torch.from_numpy(g["raw"]).cuda(), g["com"], g["cube"], g["uvd"], fx=shape.fx, fy=shape.fy,
                        frame_format=str(g["frame_format"]), prefilter=(float(g["margin"]), shape.halfu, shape.halfv))
    torch.cuda.synchronize()
    d = {k: v.cpu().numpy() for k, v in out._asdict().items() if v is not None}
    d["dmap"] = d.pop("depthmaps")
    return d


@pytest.mark.parametrize("name", RAW_SETS)
def test_gpu_raw_frames_match_reference_load_from_text(name):
    """Raw sensor samples + in-kernel PNG decode + hand rectangle == the reference's own
    load_from_text followed by process_single_data."""
    g = load_golden(name)
    got = gpu_on_raw(g)
    ref = {n: g["ref_" + n] for n in SFR_FIELDS}
    assert_sfr_matches(got, ref, SFR_FIELDS, g["ref_valid"], prefix=name + ":")


@pytest.mark.parametrize("name", RAW_SETS)
def test_gpu_raw_frames_bitwise_equal_oracle(name):
    from test_oracle_sfr import oracle_on_raw_golden
    g = load_golden(name)
    got, ref = gpu_on_raw(g), oracle_on_raw_golden(g)
    for n in ("img", "label_img", "mask"):
        assert (got[n] == ref[n]).all(), n


def test_prefilter_rectangle_python_slice_semantics():
    """A CoM near the left edge makes `right` small / `left` clamp; one far outside makes the
    Python slice wrap: compare against the oracle's literal restatement."""
    shape = synth.NYU
    d = synth.make_frames(shape, 4, 31, mixed_cube=False)
    com = d["com"].copy()
    com[0, 0] = 20.3
    com[1, 0] = -150.0          # int(com_u + du) < 0 -> negative `right` wraps around the frame
    com[2, 1] = 470.9
    frames = [so.prefilter(d["frames"][b], com[b], 150, shape.fx, shape.fy, shape.halfu, shape.halfv, 40) for b in range(4)]
    ref = so.process_batch(frames, d["uvd"], com, d["cube"], shape.fx, shape.fy, test_only=True)
    out = sfr.build_sfr(torch.from_numpy(d["frames"]).cuda(), com, d["cube"], fx=shape.fx, fy=shape.fy, test_only=True,
                        prefilter=(40, shape.halfu, shape.halfv))
    got = {k: v.cpu().numpy() for k, v in out._asdict().items() if v is not None}
    assert_sfr_matches(got, ref, SFR_FIELDS[:6], ref["valid"])
    assert (got["img"] == ref["img"]).all()


AUG_SETS = ["sfr_nyu_aug", "sfr_nyu_aug_fallback", "sfr_msra_aug"]


def gpu_on_aug(g):
    shape = golden_shape(g)
    out = sfr.build_sfr(torch.from_numpy(g["frames"]).cuda(), None if shape.com_from_frame else g["com"], g["cube"],
                        g["uvd"], fx=shape.fx, fy=shape.fy, frame_f64=shape.frame_f64, augment=g["aug"])
    torch.cuda.synchronize()
    d = {k: v.cpu().numpy() for k, v in out._asdict().items() if v is not None}
    d["dmap"] = d.pop("depthmaps")
    return d


@pytest.mark.parametrize("name", AUG_SETS)
def test_gpu_augmented_branch_matches_reference(name):
    """The reference's augmented branch (shift, cv2.warpAffine rotation + scale, joint rotation)
    replayed on the GPU with the reference's own recorded random draws, including samples that
    fall back to the plain branch."""
    g = load_golden(name)
    got = gpu_on_aug(g)
    ref = {n: g["ref_" + n] for n in SFR_FIELDS}
    assert_sfr_matches(got, ref, SFR_FIELDS, g["ref_valid"], prefix=name + ":")


@pytest.mark.parametrize("name", AUG_SETS)
def test_gpu_augmented_bitwise_equal_oracle(name):
    from test_oracle_sfr import oracle_on_aug_golden
    g = load_golden(name)
    got, ref = gpu_on_aug(g), oracle_on_aug_golden(g)
    for n in ("img", "label_img", "mask", "box_size", "com"):
        assert (got[n] == ref[n]).all(), n
    assert_sfr_matches(got, ref, SFR_FIELDS, ref["valid"])


def test_gpu_augmented_seeded_batch_matches_oracle():
    shape = synth.NYU
    B = 32
    d = synth.make_frames(shape, B, 41)
    aug = sfr.draw_augmentation(B, np.random.default_rng(5))
    ref = so.process_batch(d["frames"], d["uvd"], d["com"], d["cube"], shape.fx, shape.fy, aug=aug)
    out = sfr.build_sfr(torch.from_numpy(d["frames"]).cuda(), d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy,
                        augment=aug)
    got = {k: v.cpu().numpy() for k, v in out._asdict().items() if v is not None}
    got["dmap"] = got.pop("depthmaps")
    assert_sfr_matches(got, ref, SFR_FIELDS, ref["valid"])
    assert (got["img"] == ref["img"]).all()
    # augmentation actually changes the sample
    plain = so.process_batch(d["frames"], d["uvd"], d["com"], d["cube"], shape.fx, shape.fy)
    assert np.abs(plain["img"] - ref["img"]).max() > 1e-3
    with pytest.raises(Exception):
        sfr.build_sfr(torch.from_numpy(d["frames"]).cuda(), d["com"], d["cube"], fx=shape.fx, fy=shape.fy,
                      test_only=True, augment=aug)


def test_gpu_test_only_matches_oracle():
    shape = synth.NYU
    d = synth.make_frames(shape, 8, 7)
    ref = so.process_batch(d["frames"], None, d["com"], d["cube"], shape.fx, shape.fy, test_only=True)
    got = run_gpu(d["frames"], None, d["com"], d["cube"], shape, test_only=True)
    assert_sfr_matches(got, ref, SFR_FIELDS[:6], ref["valid"])
    assert (got["img"] == ref["img"]).all()


def test_center_of_mass_kernel_matches_oracle():
    shape = synth.MSRA
    d = synth.make_frames(shape, 6, 11)
    got = sfr.center_of_mass(torch.from_numpy(d["frames"]).cuda()).cpu().numpy()
    for b in range(6):
        ref = so.com_from_frame(d["frames"][b].astype(np.float64))
        assert (got[b, :2] == ref[:2]).all()                      # exact integer sums / count
        assert abs(got[b, 2] - ref[2]) <= 1e-12 * abs(ref[2])


def test_empty_batch_and_bad_inputs():
    from pixelwiseregression_b200._lib import PwrError
    shape = synth.NYU
    out = sfr.build_sfr(torch.zeros(0, 480, 640, device="cuda"), np.zeros((0, 3)), np.zeros(0), np.zeros((0, 14, 3)),
                        fx=shape.fx, fy=shape.fy)
    assert out.img.shape == (0, 1, 128, 128) and out.heatmaps.shape == (0, 14, 64, 64)
    with pytest.raises(PwrError):
        sfr.build_sfr(torch.zeros(1, 480, 640), np.zeros((1, 3)), 150.0, np.zeros((1, 14, 3)), fx=1.0, fy=1.0)
    with pytest.raises(PwrError):
        sfr.build_sfr(torch.zeros(1, 480, 640, device="cuda", dtype=torch.float64), np.zeros((1, 3)), 150.0,
                      np.zeros((1, 14, 3)), fx=1.0, fy=1.0)


def test_nan_and_degenerate_samples_are_flagged_not_fatal():
    shape = synth.NYU
    d = synth.make_frames(shape, 4, 21)
    com = d["com"].copy()
    com[1, 2] = 0.0            # cube / 0 -> inf -> int() raises in the reference
    com[2, 0] = np.nan
    uvd = d["uvd"].copy()
    uvd[3, 5, 0] = np.nan
    got = run_gpu(d["frames"], uvd, com, d["cube"], shape)
    assert got["valid"].tolist() == [1, 0, 0, 0]


def test_full_batch_properties():
    """B=4096 NYU (the benchmark configuration): properties that need no oracle."""
    shape = synth.NYU
    B = 4096
    d = synth.make_frames_device(shape, B, seed=3)
    out = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy)
    torch.cuda.synchronize()
    assert bool(out.valid.all())
    # mask == (label != 0), label == 2x2 mean of img up to one rounding
    assert torch.equal(out.mask, (out.label_img != 0).float())
    pooled = torch.nn.functional.avg_pool2d(out.img, 2)
    assert float((pooled - out.label_img).abs().max()) <= 2e-7 * float(out.img.abs().max()) + 1e-9
    assert float(out.img.abs().max()) < 1.0           # |depth - com_z| < cube
    # heat maps: unit mass and centre of mass at the joint (all joints are interior by construction)
    heat = out.heatmaps.double()
    assert float((heat.sum(dim=(2, 3)) - 1).abs().max()) < 1e-6
    xs = torch.arange(64, device="cuda", dtype=torch.float64)
    ku = (heat.sum(dim=2) * xs).sum(dim=2)
    kv = (heat.sum(dim=3) * xs).sum(dim=2)
    expect_u = out.uvd[:, :, 0].double() * 63 + 32     # uvd_norm = k-space / 63 shifted by 32
    expect_v = out.uvd[:, :, 1].double() * 63 + 32
    assert float((ku - expect_u).abs().max()) < 1e-4 and float((kv - expect_v).abs().max()) < 1e-4
    # Dmap support = (heat > 0) & mask, values = uvd_d - label
    supp = (out.heatmaps > 0) & (out.mask > 0)
    assert torch.equal(out.depthmaps != 0, supp & (out.depthmaps != 0))
    expect = (out.uvd[:, :, 2].view(B, -1, 1, 1) - out.label_img) * supp
    assert float((out.depthmaps - expect).abs().max()) < 1e-6
    # idempotence / determinism: a second launch is bitwise identical
    again = sfr.build_sfr(d["frames"], d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy)
    for a, b in zip(out, again):
        assert (a is None and b is None) or torch.equal(a, b)


def test_gpu_hand17_bb_frames_match_reference_golden():
    """Frames of the HAND17 bounding-box loader (float64 upstream, integer-valued, so exact in float32) through
    the GPU builder with the float64 image path and the centre-of-mass fallback on 480x640 frames, against the
    reference's `process_mode='bb'` branch.  (The bounding-box filter itself is host-side, see oracle.load_bb.)"""
    g = load_golden("sfr_hand17_bb")
    shape = golden_shape(g)
    frames32 = g["ref_frames"].astype(np.float32)
    assert np.array_equal(frames32.astype(np.float64), g["ref_frames"])
    out = sfr.build_sfr(torch.from_numpy(frames32).cuda(), None, float(shape.cube), fx=shape.fx, fy=shape.fy,
                        frame_f64=True, test_only=True)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out._asdict().items()}
    ref = {n: g["ref_" + n] for n in SFR_FIELDS[:6]}
    assert_sfr_matches(got, ref, SFR_FIELDS[:6], np.ones(len(frames32), np.uint8), prefix="bb:")


def test_gpu_hand17_bb_loader_matches_reference_golden():
    """The bounding-box loader itself on the GPU (pwr_sfr_bb_filter): box mask + two-pass "mean + 100 mm" background
    removal from the raw uint16 frame == the reference's load_from_text_bb, bit for bit; and the whole `bb` branch
    (loader -> CoM fallback -> test-only SFR with float64 frame semantics) == the reference's process_single_data."""
    g = load_golden("sfr_hand17_bb")
    shape = golden_shape(g)
    raw = torch.from_numpy(g["raw"]).cuda()
    frames = sfr.load_bb(raw, g["boxes"])
    torch.cuda.synchronize()
    assert np.array_equal(frames.cpu().numpy().astype(np.float64), g["ref_frames"])
    out = sfr.build_sfr(frames, None, float(shape.cube), fx=shape.fx, fy=shape.fy, frame_f64=True, test_only=True)
    got = {k: v.cpu().numpy() for k, v in out._asdict().items()}
    ref = {n: g["ref_" + n] for n in SFR_FIELDS[:6]}
    assert_sfr_matches(got, ref, SFR_FIELDS[:6], np.ones(len(g["raw"]), np.uint8), prefix="bb loader:")
    # Python slice semantics of the box (negative start wraps, oversize stop clips) against the oracle
    boxes = np.array([[-30.5, 100.2, 200.0, 150.0], [500.0, 400.0, 400.0, 300.0], [10.0, 10.0, 0.5, 50.0]])
    raw3 = torch.from_numpy(g["raw"][:1]).cuda().expand(3, -1, -1).contiguous()
    got3 = sfr.load_bb(raw3, boxes).cpu().numpy()
    import warnings
    for b in range(3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_b = so.load_bb(g["raw"][0], *boxes[b])
        assert np.array_equal(got3[b].astype(np.float64), ref_b), b


@pytest.mark.parametrize("cube", [150.0, 420.0])
@pytest.mark.parametrize("fmt", ["u16", "nyu_gb16"])
def test_raw_frame_window_table_equals_per_tap_arithmetic(cube, fmt):
    """Raw 16-bit frames are windowed through a per-sample lookup table (decode + window + centring of every raw
    value that can fall inside the window, filled with the per-tap functions); cube = 420 exceeds the table and
    takes the per-tap arithmetic.  Both must equal the build from the DECODED float32 frames bit for bit."""
    shape = synth.NYU
    d = synth.make_frames(shape, 12, seed=13, mixed_cube=False)
    raw = np.clip(np.rint(d["frames"]), 0, 65535).astype(np.uint16)
    if fmt == "u16":
        decoded = np.stack([so.decode_u16(r) for r in raw])
    else:
        # oracle.decode_nyu takes one frame of RGB PNG samples: G in channel 1, B in channel 2
        rgb = np.zeros(raw.shape + (3,), np.uint8)
        rgb[..., 1], rgb[..., 2] = raw >> 8, raw & 255
        decoded = np.stack([so.decode_nyu(f) for f in rgb])
    cubes = np.full(12, cube)
    kw = dict(fx=shape.fx, fy=shape.fy, prefilter=(40.0, shape.halfu, shape.halfv))
    a = sfr.build_sfr(torch.from_numpy(raw).cuda(), d["com"], cubes, d["uvd"], frame_format=fmt, targets="both", **kw)
    b = sfr.build_sfr(torch.from_numpy(np.ascontiguousarray(decoded, dtype=np.float32)).cuda(), d["com"], cubes, d["uvd"],
                      targets="both", **kw)
    torch.cuda.synchronize()
    for name, x, y in zip(a._fields, a, b):
        assert torch.equal(x, y), "%s differs between the raw-frame and the decoded-frame build (cube %g)" % (name, cube)
    assert bool(a.valid.any())


@pytest.mark.parametrize("shape_name,fmt", [("NYU", "f32"), ("NYU", "nyu_gb16"), ("HAND17", "u16"), ("MSRA", "f32")])
def test_staged_source_rows_equal_the_direct_gather(shape_name, fmt):
    """The builder gathers its taps straight from HBM (default) or stages the source rows of a band in shared memory
    with bulk-TMA copies (dispatch option `sfr_staged`; inside that kernel, bands whose rows do not fit the stage
    still gather): bit-identical outputs, including boxes over the frame corners, a box larger than the frame (no band
    fits: gather inside the staged kernel) and a tiny far-away hand (a handful of source rows)."""
    from pixelwiseregression_b200 import _lib
    shape = synth.SHAPES[shape_name]
    B = 20
    d = synth.make_frames(shape, B, seed=23)
    d["com"][0, :2] = (2.2, 3.7)
    d["com"][1, :2] = (shape.width - 1.5, shape.height - 2.5)
    d["com"][2, 2] = 140.0                             # box far larger than the frame
    d["com"][3, 2] = 4000.0                            # box of ~40 px: upsampling, few source rows
    frames = d["frames"] if fmt == "f32" else np.clip(np.rint(d["frames"]), 0, 65535).astype(np.uint16)
    kw = dict(fx=shape.fx, fy=shape.fy, frame_f64=shape.frame_f64)
    if fmt != "f32":
        kw.update(frame_format=fmt, prefilter=(40.0, shape.halfu, shape.halfv), frame_f64=False)
    com = None if shape.com_from_frame else d["com"]
    fr = torch.from_numpy(frames).cuda()
    for mode in (dict(targets="both"), dict(targets="sparse"), dict(test_only=True)):
        uvd = None if mode.get("test_only") else d["uvd"]
        gather = sfr.build_sfr(fr, com, d["cube"], uvd, **kw, **mode)
        with _lib.option("sfr_staged", 1):
            staged = sfr.build_sfr(fr, com, d["cube"], uvd, **kw, **mode)
        torch.cuda.synchronize()
        for name, x, y in zip(staged._fields, staged, gather):
            assert (x is None and y is None) or torch.equal(x, y), "%s differs (staged vs gather, %s)" % (name, mode)
    # window mode (fetched crop windows) through the staged kernel as well
    if com is not None:
        comd, cubed = torch.from_numpy(d["com"]).cuda(), torch.from_numpy(d["cube"]).cuda()
        wkw = {k: v for k, v in kw.items() if k != "frame_f64"}
        fw = sfr.fetch_windows(fr, comd, cubed, win_hw=(shape.height, shape.width), **wkw)
        with _lib.option("sfr_staged", 1):
            a = sfr.build_sfr(fw, comd, cubed, d["uvd"], **kw)
        b = sfr.build_sfr(fr, comd, cubed, d["uvd"], **kw)
        for name, x, y in zip(a._fields, a, b):
            assert (x is None and y is None) or torch.equal(x, y), "%s differs (window mode, staged)" % name
