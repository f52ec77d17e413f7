"""GPU parity of the host -> device feed (pwr_sfr_fetch + window mode of pwr_sfr_build / pwr_sfr_crop):
building from the fetched windows must be BIT-IDENTICAL to building from whole frames - which the other
test files pin to the reference - for every frame format, with the prefilter, with augmentation, on the
reference's edge cases, and through the double-buffered HostFeed."""
import numpy as np
import pytest
import torch

from pixelwiseregression_b200 import _lib, feed, sfr, synth
from helpers import golden_shape, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def same(a, b, what=""):
    for name, x, y in zip(a._fields, a, b):
        assert (x is None) == (y is None), (what, name)
        if x is not None:
            assert torch.equal(x, y), "%s %s differs between window mode and whole frames" % (what, name)


def raw_of(frames):
    return np.clip(np.rint(frames), 0, 65535).astype(np.uint16)


CASES = [("NYU", "f32", False), ("NYU", "nyu_gb16", True), ("HAND17", "u16", True), ("ICVL", "u16", True),
         ("ICVL", "f32", False)]


@pytest.mark.parametrize("shape_name,fmt,prefilter", CASES)
@pytest.mark.parametrize("where", ["pinned", "device"])
def test_window_mode_is_bit_identical_to_whole_frames(shape_name, fmt, prefilter, where):
    shape = synth.SHAPES[shape_name]
    B = 24
    d = synth.make_frames(shape, B, seed=17)
    d["com"][0, :2] = (3.2, 2.7)                       # boxes hanging over each frame corner
    d["com"][1, :2] = (shape.width - 2.5, shape.height - 1.5)
    d["com"][2, 2] = 230.0                             # a box larger than the frame
    frames = torch.from_numpy(d["frames"] if fmt == "f32" else raw_of(d["frames"]))
    kw = dict(fx=shape.fx, fy=shape.fy, frame_format=fmt,
              prefilter=(40.0, shape.halfu, shape.halfv) if prefilter else None)
    com = torch.from_numpy(d["com"]).to(DEV)
    cube = torch.from_numpy(d["cube"]).to(DEV)
    src = frames.pin_memory() if where == "pinned" else frames.to(DEV)
    fw = sfr.fetch_windows(src, com, cube, **kw)                # 16-byte loads / stores (default)
    torch.cuda.synchronize()
    ext0 = fw.extent.cpu().numpy()
    for bi in range(B):                                          # every window holds exactly its region of the frame
        r, c = int(ext0[bi, 2]), int(ext0[bi, 3])
        if r and c:
            r0, c0 = int(ext0[bi, 0]), int(ext0[bi, 1])
            assert torch.equal(fw.windows[bi, :r, :c].cpu(), frames[bi, r0:r0 + r, c0:c0 + c])
    for targets in ("dense", "both"):
        whole = sfr.build_sfr(frames.to(DEV), com, cube, d["uvd"], targets=targets, **kw)
        win = sfr.build_sfr(fw, com, cube, d["uvd"], targets=targets, **kw)
        torch.cuda.synchronize()
        same(win, whole, "%s/%s" % (shape_name, fmt))
    whole = sfr.build_sfr(frames.to(DEV), com, cube, test_only=True, **kw)
    win = sfr.build_sfr(fw, com, cube, test_only=True, **kw)
    same(win, whole, "test-only")
    assert int(fw.status) == 0
    # only the needed part crossed: far less than the frames, at least the crops' worth
    elem = 4 if fmt == "f32" else 2
    fetched = int(fw.fetched_bytes)
    assert 0 < fetched < 0.6 * frames.numel() * elem
    ext = fw.extent.cpu().numpy()
    assert fetched == int((ext[:, 2].astype(np.int64) * ext[:, 3]).sum()) * elem
    assert (ext[:, 1] * elem % 16 == 0).all() and (ext[:, 3] * elem % 16 == 0).all()
    assert (ext[:, 0] >= 0).all() and (ext[:, 0] + ext[:, 2] <= shape.height).all()
    assert (ext[:, 1] >= 0).all() and (ext[:, 1] + ext[:, 3] <= shape.width).all()


def test_window_mode_on_the_reference_edge_cases():
    """sfr_edge.npz: corner / truncated / empty crops, joints on the borders - window mode must flag and build
    exactly what whole frames do (which test_gpu_sfr pins to the reference)."""
    g = load_golden("sfr_edge")
    shape = golden_shape(g)
    frames = torch.from_numpy(g["frames"])
    com, cube = torch.from_numpy(g["com"]).to(DEV), torch.from_numpy(np.asarray(g["cube"], np.float64)).to(DEV)
    kw = dict(fx=shape.fx, fy=shape.fy)
    win_hw = (shape.height, shape.width)               # degenerate CoMs: size for the worst case
    fw = sfr.fetch_windows(frames.pin_memory(), com, cube, win_hw=win_hw, **kw)
    whole = sfr.build_sfr(frames.to(DEV), com, cube, g["uvd"], **kw)
    win = sfr.build_sfr(fw, com, cube, g["uvd"], **kw)
    torch.cuda.synchronize()
    same(win, whole, "edge")
    assert whole.valid.cpu().numpy().tolist() == g["ref_valid"].astype(int).tolist()


@pytest.mark.parametrize("name", ["sfr_nyu_aug", "sfr_nyu_aug_fallback"])
def test_window_mode_with_augmentation(name):
    """The augmented branch crops around the shifted centre and falls back to the plain branch when it raises:
    the fetch takes the union of both regions."""
    g = load_golden(name)
    shape = golden_shape(g)
    frames = torch.from_numpy(g["frames"])
    com, cube = torch.from_numpy(g["com"]).to(DEV), torch.from_numpy(np.asarray(g["cube"], np.float64)).to(DEV)
    aug = g["aug"]
    kw = dict(fx=shape.fx, fy=shape.fy)
    fw = sfr.fetch_windows(frames.pin_memory(), com, cube, augment=aug, **kw)
    whole = sfr.build_sfr(frames.to(DEV), com, cube, g["uvd"], augment=aug, **kw)
    win = sfr.build_sfr(fw, com, cube, g["uvd"], augment=aug, **kw)
    torch.cuda.synchronize()
    same(win, whole, name)
    assert int(fw.status) == 0


def test_undersized_windows_are_reported_not_silent():
    shape = synth.NYU
    d = synth.make_frames(shape, 4, seed=2)
    frames = torch.from_numpy(d["frames"]).pin_memory()
    com, cube = torch.from_numpy(d["com"]).to(DEV), torch.from_numpy(d["cube"]).to(DEV)
    fw = sfr.fetch_windows(frames, com, cube, fx=shape.fx, fy=shape.fy, win_hw=(64, 64))
    assert int(fw.status) != 0
    with pytest.raises(_lib.PwrError):
        sfr.fetch_windows(torch.from_numpy(d["frames"]), com, cube, fx=shape.fx, fy=shape.fy)     # pageable host memory
    with pytest.raises(_lib.PwrError):
        sfr.build_sfr(fw, None, cube, d["uvd"], fx=shape.fx, fy=shape.fy)                          # CoM fallback needs whole frames


def test_host_feed_double_buffering_equals_direct_builds():
    """Four different batches through a depth-2 HostFeed (submit k+1 before build k, slots re-used twice), raw NYU
    frames with the prefilter: every batch equals the direct whole-frame build; the byte counter is per batch."""
    shape = synth.NYU
    B, n = 16, 4
    kw = dict(frame_format="nyu_gb16", prefilter=(40.0, shape.halfu, shape.halfv))
    data = []
    for k in range(n):
        d = synth.make_frames(shape, B, seed=40 + k)
        d["raw"] = torch.from_numpy(raw_of(d["frames"])).pin_memory()
        data.append(d)
    hf = feed.HostFeed(shape, B, targets="both", **kw)
    t = hf.submit(data[0]["raw"], data[0]["com"], data[0]["cube"], data[0]["uvd"])
    got, fetched = [], []
    for k in range(n):
        nxt = hf.submit(data[k + 1]["raw"], data[k + 1]["com"], data[k + 1]["cube"], data[k + 1]["uvd"]) if k + 1 < n else None
        batch = hf.build(t)
        got.append(type(batch)(*[None if x is None else x.clone() for x in batch]))     # the arena is re-used
        fetched.append(hf.fetched_bytes(t))
        t = nxt
    torch.cuda.synchronize()
    for k in range(n):
        d = data[k]
        ref = sfr.build_sfr(d["raw"].to(DEV), d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, targets="both", **kw)
        same(got[k], ref, "batch %d" % k)
        assert bool(ref.valid.all())
    assert len(set(fetched)) == n and all(0 < f < 0.3 * B * 480 * 640 * 2 for f in fetched)
    with pytest.raises(_lib.PwrError):
        hf.build(0)                                     # long gone


def test_arena_reuse_is_allocation_free_and_identical():
    shape = synth.HAND17
    d = synth.make_frames(shape, 8, seed=5)
    frames = torch.from_numpy(d["frames"]).to(DEV)
    arena = sfr.SfrArena()
    a = sfr.build_sfr(frames, d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, arena=arena)
    ptrs = [t.data_ptr() for t in a if t is not None]
    keep = [None if t is None else t.clone() for t in a]
    b = sfr.build_sfr(frames, d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, arena=arena)
    assert ptrs == [t.data_ptr() for t in b if t is not None]
    ref = sfr.build_sfr(frames, d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy)
    for x, y, z in zip(keep, b, ref):
        assert (x is None and y is None) or (torch.equal(x, y) and torch.equal(y, z))


def test_unsupported_heatmap_configuration_is_refused():
    shape = synth.NYU
    d = synth.make_frames(shape, 2, seed=1)
    frames = torch.from_numpy(d["frames"]).to(DEV)
    for bad in (dict(kernel_size=5), dict(sigmoid=2.0), dict(label_size=32), dict(image_size=96)):
        with pytest.raises(_lib.PwrError, match="kernel_size=7"):
            sfr.build_sfr(frames, d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, **bad)


def test_select_valid_drops_what_the_reference_raises_on():
    shape = synth.NYU
    d = synth.make_frames(shape, 5, seed=3)
    com = d["com"].copy()
    com[1, 2] = 0.0
    com[3, 0] = np.nan
    batch = sfr.build_sfr(torch.from_numpy(d["frames"]).to(DEV), com, d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy)
    assert batch.valid.cpu().tolist() == [1, 0, 1, 0, 1]
    kept = sfr.select_valid(batch)
    assert kept.img.shape[0] == 3 and kept.heatmaps.shape[0] == 3 and bool(kept.valid.all())
    assert torch.equal(kept.img[1], batch.img[2])
    assert sfr.select_valid(kept) is kept


def test_host_feed_test_only_and_augmented_modes():
    """HostFeed in the two other modes of process_single_data: the 6-tuple of test frames (HAND17, raw 16-bit grey
    samples) and augmented training batches (per-submit draws; the fetch covers the shifted crop and its fallback)."""
    shape = synth.HAND17
    B = 10
    d = synth.make_frames(shape, B, seed=61)
    raw = torch.from_numpy(raw_of(d["frames"])).pin_memory()
    kw = dict(frame_format="u16", prefilter=(40.0, shape.halfu, shape.halfv))
    hf = feed.HostFeed(shape, B, test_only=True, **kw)
    got = hf.build(hf.submit(raw, d["com"], d["cube"]))
    ref = sfr.build_sfr(raw.to(DEV), d["com"], d["cube"], fx=shape.fx, fy=shape.fy, test_only=True, **kw)
    torch.cuda.synchronize()
    same(got, ref, "test-only feed")
    assert isinstance(got, sfr.SFRTestBatch)
    # augmented NYU batches, float32 frames: two submits with different draws through the same slots
    shape = synth.NYU
    d = synth.make_frames(shape, B, seed=62)
    frames = torch.from_numpy(d["frames"]).pin_memory()
    hf = feed.HostFeed(shape, B, augment=True, depth=1)
    for seed in (1, 2):
        aug = sfr.draw_augmentation(B, np.random.default_rng(seed))
        got = hf.build(hf.submit(frames, d["com"], d["cube"], d["uvd"], augment=aug))
        ref = sfr.build_sfr(frames.to(DEV), d["com"], d["cube"], d["uvd"], fx=shape.fx, fy=shape.fy, augment=aug)
        torch.cuda.synchronize()
        same(got, ref, "augmented feed, draw %d" % seed)
        hf.fetched_bytes(0 if seed == 1 else 1)            # raises if a region did not fit its window


def test_fetch_timing_and_uneven_shards_keep_the_global_mean():
    """HostFeed(timing=True).fetch_ms feeds distributed.proportional_shards; two feeds of unequal batch with
    n_mean = global B*J must add up to one feed of the whole batch (what the N > 1 e2e leg of bench.py relies on)."""
    from pixelwiseregression_b200 import distributed as pd, ops
    shape = synth.NYU
    B, J = 12, shape.joints
    d = synth.make_frames(shape, B, seed=23)
    raw = torch.from_numpy(raw_of(d["frames"])).pin_memory()
    pf = (40.0, shape.halfu, shape.halfv)
    g = torch.Generator(device="cpu").manual_seed(5)
    z = torch.randn(B, J, 64, 64, generator=g).to(DEV)
    D = torch.randn(B, J, 64, 64, generator=g).to(DEV)
    w = (torch.rand(J, 1, generator=g) + 0.5).to(DEV)

    def run(lo, hi, n_mean):
        hf = feed.HostFeed(shape, hi - lo, frame_format="nyu_gb16", prefilter=pf, timing=True)
        t = hf.submit(raw[lo:hi], d["com"][lo:hi], d["cube"][lo:hi], d["uvd"][lo:hi])
        batch = hf.build(t)
        ms = hf.fetch_ms(t)
        assert ms > 0.0
        zz, DD, ww = (x.clone().requires_grad_(True) for x in (z[lo:hi], D[lo:hi], w))
        total, terms, _ = ops.fused_decoder_loss(zz, ww, DD, batch.label_img, batch.mask, batch.heatmaps, batch.depthmaps,
                                                 batch.uvd, alpha=0.5, n_mean=n_mean)[:3]
        total.backward()
        return total.detach(), terms.detach(), ww.grad, zz.grad, (hi - lo) / ms

    whole = run(0, B, 0)
    shards = pd.proportional_shards([1.0, 2.0], B, 1, 1, B)
    assert shards == [4, 8]
    a, b = run(0, shards[0], B * J), run(shards[0], B, B * J)
    for i, name in ((0, "loss"), (1, "terms"), (2, "dL/dw")):
        s = a[i] + b[i]
        assert torch.allclose(s, whole[i], rtol=2e-5, atol=1e-8), name
    assert torch.allclose(torch.cat([a[3], b[3]]), whole[3], rtol=1e-5, atol=1e-10)
    with pytest.raises(_lib.PwrError):
        hf = feed.HostFeed(shape, B, frame_format="nyu_gb16", prefilter=pf)
        hf.fetch_ms(hf.submit(raw, d["com"], d["cube"], d["uvd"]))
