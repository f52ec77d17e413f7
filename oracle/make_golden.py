"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Runs the reference's own HandDataset.process_single_data (datasets.py:185-403)
and its PlaneRegression / DepthRegression forward (model.py:76-132) + the
train.py:197-207 loss/backward on seeded synthetic inputs and stores inputs and
outputs as small compressed fixtures.  The reference cannot travel to the GPU
box; these vectors can.  Versions the vectors were produced with are stored in
each file (`versions`).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from pixelwiseregression_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIELDS = ("img", "label_img", "mask", "box_size", "cube_size", "com", "uvd", "heatmaps", "dmap")


def versions():
    import cv2
    import scipy
    return np.array("numpy %s; opencv %s; scipy %s; torch %s" % (
        np.__version__, cv2.__version__, scipy.__version__, torch.__version__))


def run_reference_sfr(datasets, shape, frames, uvd, com, cube, test_only=False):
    """frames in the dtype the reference would hold them in."""
    B = len(frames)
    samples = [dict(frame=frames[b], uvd=uvd[b], com=com[b],
                    cube=int(cube[b]) if float(cube[b]) == int(cube[b]) else float(cube[b]))
               for b in range(B)]
    ds = ref_shim.make_synthetic_dataset(datasets, samples, shape.fx, shape.fy, shape.halfu, shape.halfv,
                                         int(shape.cube), shape.joints, test_only=test_only,
                                         msra_com=shape.com_from_frame)
    J = shape.joints
    zero = dict(img=np.zeros((1, 128, 128), np.float32), label_img=np.zeros((1, 64, 64), np.float32),
                mask=np.zeros((1, 64, 64), np.float32), box_size=np.float32(0), cube_size=np.float32(0),
                com=np.zeros(3, np.float32), uvd=np.zeros((J, 3), np.float32),
                heatmaps=np.zeros((J, 64, 64), np.float32), dmap=np.zeros((J, 64, 64), np.float32))
    names = FIELDS[:6] if test_only else FIELDS
    outs, valid = [], []
    for b in range(B):
        try:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                tup = ds[b]
            outs.append({n: t.numpy() for n, t in zip(names, tup)})
            valid.append(1)
        except Exception:
            outs.append({n: zero[n] for n in names})
            valid.append(0)
    res = {"ref_" + n: np.stack([o[n] for o in outs]) for n in names}
    res["ref_valid"] = np.array(valid, dtype=np.uint8)
    return res


def golden_sfr(datasets, name, shape, batch, seed, test_only=False):
    d = synth.make_frames(shape, batch, seed)
    frames = d["frames"].astype(np.float64) if shape.frame_f64 else d["frames"]
    res = run_reference_sfr(datasets, shape, frames, d["uvd"], d["com"], d["cube"], test_only)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), frames=d["frames"], uvd=d["uvd"], com=d["com"],
                        cube=d["cube"], shape_name=np.array(shape.name), test_only=np.array(test_only),
                        versions=versions(), **res)
    print(name, "valid", res["ref_valid"].tolist())


def edge_cases(shape):
    """Hand-placed NYU-shaped samples for the reference's corner behaviour."""
    rng = np.random.default_rng(1234)
    base = synth.make_frames(shape, 1, 77, mixed_cube=False)
    frames, uvds, coms, cubes, notes = [], [], [], [], []

    def add(frame, uvd, com, cube, note):
        frames.append(frame.copy())
        uvds.append(uvd.copy())
        coms.append(np.array(com, dtype=np.float64))
        cubes.append(float(cube))
        notes.append(note)

    f0, u0, c0 = base["frames"][0], base["uvd"][0], base["com"][0]
    box = max(int(150 / c0[2] * shape.fx + 150 / c0[2] * shape.fy), 2)
    bs = 2 * (box // 2)

    def joints_at(label_coords):
        """uvd whose heat-map pixel coordinates land near `label_coords`."""
        u = u0.copy()
        for j, (ku, kv) in enumerate(label_coords):
            u[j, 0] = (ku - 32) / 63.0 * (bs - 1) + int(c0[0])
            u[j, 1] = (kv - 32) / 63.0 * (bs - 1) + int(c0[1])
        return u

    border = [(0.0, 0.0), (0.5, 0.5), (2.3, 1.2), (61.7, 2.3), (62.999, 62.999), (1.0, 61.0),
              (30.0, 30.0), (30.5, 30.0), (30.0, 30.5), (30.25, 30.75), (30.75, 30.25), (62.0, 31.5),
              (3.0, 3.0), (59.99, 60.01)]
    add(f0, joints_at(border), c0, 150, "joints on the border: REFLECT_101 mass, integer coords")
    wrap = list(border)
    wrap[0] = (-0.5, 10.0)
    wrap[1] = (10.0, -0.25)
    wrap[2] = (-1.5, -1.5)
    add(f0, joints_at(wrap), c0, 150, "negative heat-map indices wrap NumPy-style")
    oor = list(border)
    oor[3] = (63.2, 20.0)
    add(f0, joints_at(oor), c0, 150, "low_u+1 == 64 -> Out of range -> rejected")
    oor2 = list(border)
    oor2[5] = (20.0, 63.0)
    add(f0, joints_at(oor2), c0, 150, "low_v+1 == 64 exactly -> rejected")
    far = list(border)
    far[7] = (-70.0, 20.0)
    add(f0, joints_at(far), c0, 150, "index < -64 -> IndexError -> rejected")

    # CoM in a frame corner: the crop box straddles two frame edges
    hf, wf = shape.height, shape.width
    yy, xx = np.mgrid[0:hf, 0:wf]
    for (cu, cv, cz, note) in [(12.7, 9.2, 640.5, "top-left corner"),
                               (wf - 3.4, hf - 5.6, 555.5, "bottom-right corner"),
                               (wf - 0.2, 200.3, 700.0, "int(com_u) == Wf-1"),
                               (float(wf), float(hf), 612.0, "int(com) == frame size (slice still full)"),
                               (wf + 40.6, 240.0, 800.0, "CoM beyond the frame: truncated (non-square) crop"),
                               (300.0, hf + 25.3, 650.0, "CoM below the frame: truncated rows"),
                               (-3.5, 100.0, 700.0, "negative CoM: empty crop -> rejected"),
                               (100.5, -0.5, 700.0, "com_v in (-1,0): int() truncates to 0")]:
        fr = np.zeros((hf, wf), np.float32)
        rad = 0.7 * 150 / cz * shape.fx
        disc = (xx - cu) ** 2 + (yy - cv) ** 2 < rad * rad
        fr[disc] = (cz + rng.uniform(-100, 100, int(disc.sum()))).astype(np.float32)
        b = max(int(150 / cz * shape.fx + 150 / cz * shape.fy), 2) // 2
        uv = np.stack([cu + rng.uniform(-0.3, 0.3, shape.joints) * b,
                       cv + rng.uniform(-0.3, 0.3, shape.joints) * b,
                       cz + rng.uniform(-100, 100, shape.joints)], 1)
        add(fr, uv, (cu, cv, cz), 150, note)

    # fewer than 10 valid label pixels -> rejected; exactly-at-window pixels
    fr = np.zeros((hf, wf), np.float32)
    fr[238:241, 318:321] = 750.0 + 20.0
    add(fr, u0 * 0 + np.array([320.0, 240.0, 760.0]), (320.4, 240.6, 750.0), 150, "sum(mask) < 10 -> rejected")
    fr = f0.copy()
    pos = np.argwhere(fr > 0)
    fr[tuple(pos[10])] = np.float32(c0[2])            # exactly com_z -> becomes 0 after centring
    fr[tuple(pos[11])] = np.float32(c0[2] + 150.0)    # on the strict window edge
    fr[tuple(pos[12])] = np.float32(c0[2] - 150.0)
    add(fr, u0, c0, 150, "pixels == com_z and on the strict window bounds")
    # very near hand -> huge box (> frame), very far -> tiny box
    fr = np.zeros((hf, wf), np.float32)
    fr[100:400, 150:500] = (200.0 + rng.uniform(-50, 50, (300, 350))).astype(np.float32)
    add(fr, u0 * 0 + np.array([320.0, 240.0, 200.0]) + rng.uniform(-30, 30, (shape.joints, 3)),
        (320.5, 240.5, 200.0), 150, "box larger than the frame (2s = 880)")
    fr = np.zeros((hf, wf), np.float32)
    fr[230:250, 310:330] = (6000.0 + rng.uniform(-50, 50, (20, 20))).astype(np.float32)
    add(fr, u0 * 0 + np.array([320.0, 240.0, 6000.0]) + rng.uniform(-5, 5, (shape.joints, 3)),
        (320.5, 240.5, 6000.0), 150, "tiny box (2s = 28): up-sampling resize")
    return (np.stack(frames), np.stack(uvds), np.stack(coms), np.array(cubes), np.array(notes))


def golden_edge(datasets):
    shape = synth.NYU
    frames, uvd, com, cube, notes = edge_cases(shape)
    res = run_reference_sfr(datasets, shape, frames, uvd, com, cube)
    np.savez_compressed(os.path.join(GOLDEN, "sfr_edge.npz"), frames=frames, uvd=uvd, com=com, cube=cube,
                        shape_name=np.array(shape.name), test_only=np.array(False), notes=notes,
                        versions=versions(), **res)
    for n, v in zip(notes, res["ref_valid"]):
        print("  edge: valid=%d  %s" % (v, n))


def golden_raw(datasets, name, shape, frame_format, margin, batch, seed, val_mode=False):
    """Raw sensor frames through the reference's OWN load_from_text (PNG decode line, hand
    rectangle, depth window; NYU datasets.py:797-859, HAND17 :925-972, ICVL :626-690) and then
    its process_single_data.  plt.imread is the only stand-in: matplotlib is not installed, so
    it is replaced by its documented conversion (oracle.sfr_oracle.imread_float) over in-memory
    PNG sample arrays."""
    from oracle import sfr_oracle as so
    d = synth.make_frames(shape, batch, seed, mixed_cube=False)
    raw = np.clip(np.rint(d["frames"]), 0, 65535).astype(np.uint16)          # sensor counts (mm)
    store = {}
    cls = {"NYU": datasets.NYUDataset, "HAND17": datasets.HAND17Dataset, "ICVL": datasets.ICVLDataset}[shape.name]
    ds = object.__new__(cls)
    ds.fx, ds.fy, ds.halfu, ds.halfv = shape.fx, shape.fy, shape.halfu, shape.halfv
    ds.cube_size, ds.path = int(shape.cube), "/synthetic"
    ds.dataset = "val" if val_mode else "train"
    first = 2500 if val_mode else 0          # NYU val index > 2440 -> int(cube * 5 / 6)
    centers = np.zeros((first + batch, 3))
    centers[first:] = d["com"]
    ds.train_centers = ds.test_centers = centers
    ds.train_lookup = {}
    old_imread = getattr(datasets.plt, "imread", None)
    datasets.plt.imread = lambda path: so.imread_float(store[path])
    images, uvds, coms, cubes = [], [], [], []
    try:
        for b in range(batch):
            idx = first + b
            if shape.name == "NYU":
                path = "/synthetic/train/depth_1_%07d.png" % (idx + 1)
                rgb = np.zeros(raw[b].shape + (3,), np.uint8)
                rgb[..., 1] = raw[b] >> 8
                rgb[..., 2] = raw[b] & 255
                store[path] = rgb
                joints = d["uvd"][b]
            elif shape.name == "HAND17":
                rel = "image_D%08d.png" % (idx + 1)
                path = rel
                store[os.path.join(ds.path, "training", "images", rel)] = raw[b]
                uvd = d["uvd"][b]                                  # text holds xyz; the loader maps it to uvd
                joints = uvd.copy()
                joints[:, 0] = (uvd[:, 0] - ds.halfu) / ds.fx * uvd[:, 2]
                joints[:, 1] = (uvd[:, 1] - ds.halfv) / ds.fy * uvd[:, 2]
            else:
                path = "/synthetic/Training/Depth/seq/image_%04d.png" % idx
                store[path] = raw[b]
                ds.train_lookup["/".join(path.split("/")[-2:])] = idx
                joints = d["uvd"][b]
            text = path + " " + " ".join(repr(float(x)) for x in joints.reshape(-1))
            image, joint_uvd, com, cube = ds.load_from_text(text)
            images.append(image)
            uvds.append(np.asarray(joint_uvd, dtype=np.float64))
            coms.append(np.asarray(com, dtype=np.float64))
            cubes.append(float(ds.cube_size if cube is None else cube))
    finally:
        datasets.plt.imread = old_imread
    res = run_reference_sfr(datasets, shape, images, uvds, coms, cubes)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), raw=raw, uvd=np.stack(uvds), com=np.stack(coms),
                        cube=np.array(cubes), shape_name=np.array(shape.name), frame_format=np.array(frame_format),
                        margin=np.float64(margin), test_only=np.array(False), versions=versions(), **res)
    print(name, "valid", res["ref_valid"].tolist(), "cubes", sorted(set(cubes)))


def golden_bb(datasets, name, batch, seed):
    """HAND17 `process_mode='bb'` (datasets.py:199-206, 974-996): the reference's own load_from_text_bb
    (plt.imread replaced by its documented conversion, as in golden_raw) and the bb branch of its
    process_single_data (test-only; CoM and cube from the fallback)."""
    from oracle import sfr_oracle as so
    shape = synth.HAND17
    d = synth.make_frames(shape, batch, seed, mixed_cube=False)
    raw = np.clip(np.rint(d["frames"]), 0, 65535).astype(np.uint16)
    rng = np.random.default_rng(seed)
    # clutter behind the hand inside the box (what the two-pass mean + 100 mm rule removes) and a far wall
    for b in range(batch):
        wall = rng.uniform(size=raw[b].shape) < 0.02
        raw[b][wall & (raw[b] == 0)] = np.uint16(d["com"][b, 2] + 400)
    boxes = []
    for b in range(batch):
        u, v, z = d["com"][b]
        half = 0.9 * shape.cube / z * shape.fx
        boxes.append([u - half + rng.uniform(-3, 3), v - half + rng.uniform(-3, 3), 2 * half + 0.5, 2 * half + 1.25])
    boxes = np.array(boxes)
    store = {}
    old_imread = getattr(datasets.plt, "imread", None)
    datasets.plt.imread = lambda path: so.imread_float(store[path])
    try:
        ds = object.__new__(datasets.HAND17Dataset)      # file-reading constructor bypassed: only I/O is replaced
        ds.fx, ds.fy, ds.halfu, ds.halfv = shape.fx, shape.fy, shape.halfu, shape.halfv
        ds.path, ds.cube_size, ds.process_mode, ds.test_only = "/synthetic", int(shape.cube), "bb", True
        ds.image_size, ds.label_size, ds.kernel_size, ds.sigmoid, ds.joint_number = 128, 64, 7, 1.5, shape.joints
        ds.using_rotation = ds.using_scale = ds.using_shift = ds.using_flip = False
        frames64, outs = [], []
        for b in range(batch):
            rel = "image_D%08d.png" % (b + 1)
            store[os.path.join(ds.path, "frame", "images", rel)] = raw[b]
            text = rel + " " + " ".join(repr(float(x)) for x in boxes[b])
            frames64.append(ds.load_from_text_bb(text))
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                tup = ds.process_single_data(text)
            outs.append([t.numpy() for t in tup])
    finally:
        datasets.plt.imread = old_imread
    names = FIELDS[:6]
    res = {"ref_" + n: np.stack([o[i] for o in outs]) for i, n in enumerate(names)}
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), raw=raw, boxes=boxes, ref_frames=np.stack(frames64),
                        shape_name=np.array(shape.name), versions=versions(), **res)
    print(name, "box sizes", res["ref_box_size"].tolist(), "com z", res["ref_com"][:, 2].tolist())


def golden_augmented(datasets, name, shape, batch, seed, spread=0.6):
    """The reference's augmented branch (datasets.py:216-299, train.py's default flags) with
    `random.random` replaced by a recorded sequence, so the draws can be replayed on the GPU."""
    import random
    d = synth.make_frames(shape, batch, seed)
    if spread != 0.6:                                   # push joints towards the crop border
        rng = np.random.default_rng(seed + 99)
        for b in range(batch):
            z0 = d["com"][b, 2]
            sh = max(int(d["cube"][b] / z0 * shape.fx + d["cube"][b] / z0 * shape.fy), 2) // 2
            d["uvd"][b, :, 0] = d["com"][b, 0] + rng.uniform(-spread, spread, shape.joints) * sh
            d["uvd"][b, :, 1] = d["com"][b, 1] + rng.uniform(-spread, spread, shape.joints) * sh
            d["uvd"][b, 0, :2] = d["com"][b, :2] + 0.9 * sh      # a corner joint: leaves the map once rotated
    frames = d["frames"].astype(np.float64) if shape.frame_f64 else d["frames"]
    samples = [dict(frame=frames[b], uvd=d["uvd"][b], com=d["com"][b], cube=int(d["cube"][b])) for b in range(batch)]
    ds = ref_shim.make_synthetic_dataset(datasets, samples, shape.fx, shape.fy, shape.halfu, shape.halfv,
                                         int(shape.cube), shape.joints, msra_com=shape.com_from_frame, augment=True)
    draw_rng = np.random.default_rng(seed + 7)
    real_random = random.random
    outs, augs, valid = [], [], []
    try:
        for b in range(batch):
            log = []

            def fake():
                v = float(draw_rng.uniform())
                log.append(v)
                return v
            random.random = fake
            try:
                import contextlib
                import io
                with contextlib.redirect_stdout(io.StringIO()):
                    tup = ds[b]
                outs.append({n: t.numpy() for n, t in zip(FIELDS, tup)})
                valid.append(1)
            except Exception:
                outs.append(None)
                valid.append(0)
            random.random = real_random
            # draws: angle (discarded, datasets.py:225), scale :230, shift_x :236, shift_y :237, angle (utils.py:72)
            assert len(log) == 5, log
            augs.append([0.8 + log[1] * 0.4, -5 + log[2] * 10, -5 + log[3] * 10, log[4] * 60 - 30])
    finally:
        random.random = real_random
    J = shape.joints
    zero = dict(img=np.zeros((1, 128, 128), np.float32), label_img=np.zeros((1, 64, 64), np.float32),
                mask=np.zeros((1, 64, 64), np.float32), box_size=np.float32(0), cube_size=np.float32(0),
                com=np.zeros(3, np.float32), uvd=np.zeros((J, 3), np.float32),
                heatmaps=np.zeros((J, 64, 64), np.float32), dmap=np.zeros((J, 64, 64), np.float32))
    res = {"ref_" + n: np.stack([(o or zero)[n] for o in outs]) for n in FIELDS}
    res["ref_valid"] = np.array(valid, dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), frames=d["frames"], uvd=d["uvd"], com=d["com"],
                        cube=d["cube"], aug=np.array(augs), shape_name=np.array(shape.name),
                        test_only=np.array(False), versions=versions(), **res)
    print(name, "valid", valid, "aug[0]", np.round(augs[0], 3).tolist())


def golden_decoder(model_mod, name, method, B, J, seed, alpha, with_upstream):
    """Reference PlaneRegression/DepthRegression with `.conv` swapped for
    Identity (so the module input is the logit map itself), reference loss
    lines, autograd backward."""
    torch.manual_seed(seed)
    d = synth.make_decoder_inputs(B, J, seed)
    z = torch.from_numpy(d["z"]).requires_grad_(True)
    D = torch.from_numpy(d["D"]).requires_grad_(True)
    label = torch.from_numpy(d["label"])
    mask = torch.from_numpy(d["mask"])
    plane = model_mod.PlaneRegression(8, J, 64, normalization_method=method)
    depth = model_mod.DepthRegression(8, J)
    plane.conv = torch.nn.Identity()
    depth.conv = torch.nn.Identity()
    if method == "softmax":
        with torch.no_grad():
            plane.w.copy_(torch.from_numpy(d["w"]))
    rng = np.random.default_rng(seed + 1)
    heat_gt = torch.from_numpy(rng.uniform(0, 0.05, (B, J, 64, 64)).astype(np.float32))
    dmap_gt = torch.from_numpy((rng.standard_normal((B, J, 64, 64)) * d["mask"]).astype(np.float32))
    uvd_gt = torch.from_numpy(rng.uniform(-0.5, 0.5, (B, J, 3)).astype(np.float32))
    lambda_h, lambda_d = 1.0, 0.01

    zz = z.clone()  # relu(inplace=True) in the 'sum' branch must not hit the leaf
    heat, uv = plane(zz)
    dm, dep = depth(D, heat, label, mask)
    uvd = torch.cat([uv, dep], dim=2)
    # train.py:197-205
    heatmap_loss = lambda_h * torch.mean(torch.sum((heat - heat_gt) ** 2, dim=(2, 3)))
    depthmap_loss = lambda_d * torch.mean(torch.sum((dm - dmap_gt) ** 2, dim=(2, 3)))
    uvd_loss = torch.mean(torch.sum((uvd - uvd_gt) ** 2, dim=2))
    loss = alpha * uvd_loss + (1 - alpha) * (heatmap_loss + depthmap_loss)
    out = dict(z=d["z"], D=d["D"], w=d["w"], label=d["label"], mask=d["mask"], heat_gt=heat_gt.numpy(),
               dmap_gt=dmap_gt.numpy(), uvd_gt=uvd_gt.numpy(), alpha=np.float64(alpha),
               lambda_h=np.float64(lambda_h), lambda_d=np.float64(lambda_d), method=np.array(method))
    if with_upstream:
        gH_up = torch.from_numpy((rng.standard_normal((B, J, 64, 64)) * 1e-3).astype(np.float32))
        gD_up = torch.from_numpy((rng.standard_normal((B, J, 64, 64)) * 1e-3).astype(np.float32))
        g_uvd_up = torch.from_numpy((rng.standard_normal((B, J, 3)) * 1e-2).astype(np.float32))
        total = loss + (heat * gH_up).sum() + (dm * gD_up).sum() + (uvd * g_uvd_up).sum()
        out.update(gH_up=gH_up.numpy(), gD_up=gD_up.numpy(), g_uvd_up=g_uvd_up.numpy())
    else:
        total = loss
    total.backward()
    out.update(ref_heat=heat.detach().numpy(), ref_uvd=uvd.detach().numpy(),
               ref_losses=np.array([heatmap_loss.item(), depthmap_loss.item(), uvd_loss.item(), loss.item()]),
               ref_gz=z.grad.numpy(), ref_gD=D.grad.numpy(), versions=versions())
    if method == "softmax":
        out["ref_gw"] = plane.w.grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "losses", out["ref_losses"])


def golden_model(model_mod, datasets, name, shape, B, seed, features=32, level=2, stages=2):
    """BASELINE configs[0] (reference PixelwiseRegression, eval forward on the CPU, crops from the
    reference's own SFR builder) at a width that keeps the fixture small: the reference's weights travel in
    the file, so the drop-in model must load them (state_dict compatibility) and reproduce the outputs."""
    torch.manual_seed(seed)
    d = synth.make_frames(shape, B, seed)
    ref = run_reference_sfr(datasets, shape, d["frames"], d["uvd"], d["com"], d["cube"], test_only=True)
    assert ref["ref_valid"].all()
    net = model_mod.PixelwiseRegression(shape.joints, stage=stages, features=features, level=level, norm_method="instance")
    with torch.no_grad():
        for st in net.stages:                      # trained temperatures are not all 1 (and can be negative)
            st.plane_regression.w.uniform_(0.5, 1.5)
        net.stages[-1].plane_regression.w[0] = -0.8
    net.eval()
    img, label, mask = (torch.from_numpy(ref["ref_" + n]) for n in ("img", "label_img", "mask"))
    with torch.no_grad():
        results = net(img, label, mask)
    out = {"sd_" + k: v.numpy() for k, v in net.state_dict().items()}
    out.update(img=img.numpy(), label_img=label.numpy(), mask=mask.numpy(), joints=np.int64(shape.joints),
               features=np.int64(features), level=np.int64(level), stages=np.int64(stages), versions=versions())
    for i, (heat, dm, uvd) in enumerate(results):
        out["ref_uvd_%d" % i] = uvd.numpy()
    out["ref_heat_last"] = results[-1][0][:1].numpy()       # one sample of maps is enough to pin the layout
    out["ref_dmap_last"] = results[-1][1][:1].numpy()
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    n_par = sum(p.numel() for p in net.parameters())
    print(name, "params", n_par, "uvd[0,0]", results[-1][2][0, 0].tolist())


STATE_DICT_SPECS = {
    # name -> (reference class, constructor kwargs): the module trees whose state_dict layout the drop-in
    # model.py must reproduce key for key (released checkpoints load through utils.py:309-314)
    "nyu_instance": ("PixelwiseRegression", dict(joints=14, stage=2, features=128, level=4, norm_method="instance")),
    "msra_batch_sum": ("PixelwiseRegression", dict(joints=21, stage=2, features=64, level=2, norm_method="batch",
                                                   heatmap_method="sum")),
    "fullregression_nyu": ("FullRegression", dict(joints=14, stage=2, features=64, level=2, norm_method="instance")),
}


def golden_state_dict_keys(model_mod):
    """tests/golden/state_dict_keys.json: [key, shape] lists of the reference's own modules."""
    import json
    spec = {}
    for name, (cls, kwargs) in STATE_DICT_SPECS.items():
        net = getattr(model_mod, cls)(**kwargs)
        spec[name] = {"class": cls, "kwargs": kwargs,
                      "keys": [[k, list(v.shape)] for k, v in net.state_dict().items()]}
        print("state_dict", name, len(spec[name]["keys"]), "entries")
    with open(os.path.join(GOLDEN, "state_dict_keys.json"), "w") as f:
        json.dump(spec, f)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    model_mod, _, datasets = ref_shim.load()
    golden_state_dict_keys(model_mod)
    if sys.argv[1:] == ["--state-dict-only"]:
        return
    golden_sfr(datasets, "sfr_nyu", synth.NYU, 4, 0)
    golden_sfr(datasets, "sfr_nyu_test_only", synth.NYU, 2, 5, test_only=True)
    golden_sfr(datasets, "sfr_hand17", synth.HAND17, 2, 1)
    golden_sfr(datasets, "sfr_msra", synth.MSRA, 3, 2)
    golden_sfr(datasets, "sfr_icvl", synth.ICVL, 2, 3)
    golden_edge(datasets)
    golden_augmented(datasets, "sfr_nyu_aug", synth.NYU, 4, 30)
    golden_augmented(datasets, "sfr_nyu_aug_fallback", synth.NYU, 6, 31, spread=0.5)
    golden_augmented(datasets, "sfr_msra_aug", synth.MSRA, 2, 32)
    golden_raw(datasets, "sfr_nyu_raw", synth.NYU, "nyu_gb16", 40, 3, 20)
    golden_raw(datasets, "sfr_nyu_raw_val", synth.NYU, "nyu_gb16", 40, 2, 21, val_mode=True)
    golden_raw(datasets, "sfr_hand17_raw", synth.HAND17, "u16", 40, 2, 22)
    golden_raw(datasets, "sfr_icvl_raw", synth.ICVL, "u16", 30, 2, 23)
    golden_decoder(model_mod, "decoder_softmax_a1", "softmax", 2, 3, 10, 1.0, False)
    golden_decoder(model_mod, "decoder_softmax_a05_up", "softmax", 2, 3, 11, 0.5, True)
    golden_decoder(model_mod, "decoder_sum_a05_up", "sum", 2, 3, 12, 0.5, True)
    golden_model(model_mod, datasets, "model_nyu_eval", synth.NYU, 6, 40)
    golden_bb(datasets, "sfr_hand17_bb", 3, 50)


if __name__ == "__main__":
    main()
