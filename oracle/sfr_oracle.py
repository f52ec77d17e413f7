"""CPU oracle for the SFR target builder (TEST INFRASTRUCTURE, not product).

A NumPy restatement of the reference's non-augmented SFR path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path (pixelwiseregression_b200/) never
does and fails loudly when the CUDA library is missing.

Reference lines followed (relative to the reference tree):
  datasets.py:208-211   CoM fallback (MSRA)               -> com_from_frame
  datasets.py:306-309   crop box                          -> crop_box
  utils.py:167-173      center_crop                       -> center_crop
  datasets.py:312-319   depth window, centring, int CoM   -> window_and_centre
  datasets.py:323       cv2.resize -> 128x128             -> resize_bilinear
  datasets.py:330-332   cv2.resize -> 64x64, mask         -> resize_half / mask
  datasets.py:350-358   uvd centring + heatmap coords     -> joint_coords
  utils.py:37-62        generate_heatmap (4-tap splat)    -> splat4
  utils.py:64-65        cv2.GaussianBlur 7x7 sigma 1.5    -> gaussian_blur7
  datasets.py:369-383   Dmap + normalisation              -> process_sample
  datasets.py:385-390   reject gate                       -> process_sample
  utils.py:332-337      recover_uvd                       -> recover_uvd

Third-party arithmetic that is NOT in the reference tree is restated from its
published algorithm: OpenCV (env.yml: `opencv`, unpinned; pinned here to the
container's 4.13.0) `cv2.resize` INTER_LINEAR and `cv2.GaussianBlur`; SciPy
`center_of_mass`.  Parity pinning: the reference has no tests or golden
vectors, so this oracle is pinned against outputs of the UNMODIFIED reference
run in the build container (oracle/make_golden.py -> tests/golden/*.npz) and
against cv2 directly when it is importable (tests/test_oracle_sfr.py).

Dtype rules (NumPy-2 / NEP-50 promotion, probed on the reference):
  * NYU/ICVL/HAND17 frames are float32 -> crop, window, resize, label are
    float32; `crop[crop>0] -= com_z` is evaluated in float64 and rounded back
    to float32; the window compare is done in float64.
  * MSRA frames are float64 -> the whole image path is float64 until the final
    `.float()`.
  * Joint coordinates, heatmaps and Dmap are float64 until `.float()`.
"""
import math
import numpy as np

IMAGE_SIZE = 128
LABEL_SIZE = 64
KSIZE = 7
SIGMA = 1.5


# --------------------------------------------------------------------------- #
# scalar / index helpers
# --------------------------------------------------------------------------- #
def com_from_frame(frame):
    """datasets.py:208-211.  mean depth of positive pixels and the centre of
    mass of the binary mask (scipy center_of_mass of `img > 0`)."""
    pos = frame > 0
    cnt = int(pos.sum())
    mean = np.mean(frame[pos])
    rows, cols = np.nonzero(pos)
    # scipy: sum(input * grid) / sum(input) with input = bool mask -> exact ints
    com_row = np.float64(rows.sum()) / np.float64(cnt)
    com_col = np.float64(cols.sum()) / np.float64(cnt)
    return np.array([com_col, com_row, np.float64(mean)], dtype=np.float64)


def crop_box(com_z, cube, fx, fy):
    """datasets.py:306-309."""
    du = np.float64(cube) / np.float64(com_z) * np.float64(fx)
    dv = np.float64(cube) / np.float64(com_z) * np.float64(fy)
    return max(int(du + dv), 2)


def crop_geometry(com, box, hf, wf):
    """utils.py:167-173 index arithmetic.  Returns (r0, c0, shift, nrows,
    ncols): the crop covers frame rows [r0-shift, r0-shift+nrows) and columns
    [c0-shift, c0-shift+ncols) with zeros outside the frame.  nrows/ncols are
    2*shift unless the padded-array slice truncates (CoM beyond the frame) and
    0 when the Python slice is empty (negative CoM)."""
    r0 = int(com[1])
    c0 = int(com[0])
    shift = box // 2

    def extent(start, full):
        n = full + 2 * shift
        s = slice(start, start + 2 * shift).indices(n)
        return max(0, s[1] - s[0]), s[0]

    nrows, rs = extent(r0, hf)
    ncols, cs = extent(c0, wf)
    return r0, c0, shift, nrows, ncols, rs, cs


def center_crop(frame, com, box):
    """utils.py:167-173 without materialising the padded frame."""
    hf, wf = frame.shape
    r0, c0, shift, nrows, ncols, rs, cs = crop_geometry(com, box, hf, wf)
    out = np.zeros((nrows, ncols), dtype=frame.dtype)
    # padded index p maps to frame index p - shift
    fr0, fc0 = rs - shift, cs - shift
    ra, rb = max(fr0, 0), min(fr0 + nrows, hf)
    ca, cb = max(fc0, 0), min(fc0 + ncols, wf)
    if rb > ra and cb > ca:
        out[ra - fr0:rb - fr0, ca - fc0:cb - fc0] = frame[ra:rb, ca:cb]
    return out


def window_and_centre(crop, com_z, cube):
    """datasets.py:312,315.  Strict float64 window, then subtract com_z from
    positive pixels (float64 arithmetic, rounded to the crop dtype)."""
    z = np.float64(com_z)
    lo = z - cube
    hi = z + cube
    keep = np.logical_and(crop.astype(np.float64) > lo, crop.astype(np.float64) < hi)
    crop = crop * keep.astype(crop.dtype)
    pos = crop > 0
    crop[pos] = (crop[pos].astype(np.float64) - z).astype(crop.dtype)
    return crop


# --------------------------------------------------------------------------- #
# raw sensor frames: PNG channel decode and the load_from_text prefilter
# --------------------------------------------------------------------------- #
def imread_float(png):
    """matplotlib.pyplot.imread on a PNG (matplotlib is not in the reference tree nor
    installed here; restated from matplotlib.image._pil_png_to_float_array): 8-bit
    channels -> float32(v / 255), 16-bit grey -> float32(v / 65535), correctly
    rounded float32 divisions."""
    png = np.asarray(png)
    if png.dtype == np.uint8:
        return np.divide(png, 2 ** 8 - 1, dtype=np.float32)
    if png.dtype == np.uint16:
        return np.divide(png, 2 ** 16 - 1, dtype=np.float32)
    raise TypeError("PNG sample type %s" % png.dtype)


def decode_nyu(rgb_u8):
    """datasets.py:809-810: depth = (G * 256 + B) * 255 on the float32 channels."""
    img = imread_float(rgb_u8)
    return (img[:, :, 1] * 256 + img[:, :, 2]) * 255


def decode_u16(grey_u16):
    """datasets.py:632 (ICVL), :940/:950 (HAND17): plt.imread(path) * 65535."""
    return imread_float(grey_u16) * 65535


def load_bb(grey_u16, ustart, vstart, du, dv):
    """HAND17 `process_mode='bb'` loader, datasets.py:974-996 (test frames that come with a bounding
    box instead of joint annotations): keep the box, then drop everything deeper than 100 mm behind
    the mean depth of what is left, in two passes.  `MM = np.zeros(image.shape)` is float64, so the
    frame the reference carries on is float64 (like MSRA); CoM and cube then come from the fallback of
    process_single_data (datasets.py:203-214)."""
    image = decode_u16(grey_u16)
    mm = np.zeros(image.shape)
    mm[int(vstart):int(vstart + dv), int(ustart):int(ustart + du)] = 1
    image = image * mm
    mean = np.mean(image[image > 0])
    first = image.copy()
    first[first > mean + 100] = 0
    mean = np.mean(first[first > 0])
    image[image > mean + 100] = 0
    return image


def prefilter(image, com, cube, fx, fy, halfu, halfv, margin):
    """The hand rectangle + depth window of load_from_text (NYU datasets.py:841-857 with
    margin 40, HAND17 :956-972 margin 40, ICVL :666-681 margin 30).  Python slice
    semantics (a negative `right` wraps) are the reference's."""
    du = (cube - margin) / com[2] * fx
    dv = (cube - margin) / com[2] * fy
    left = max(int(com[0] - du), 0)
    top = max(int(com[1] - dv), 0)
    right = int(min(int(com[0] + du), halfu * 2))
    buttom = int(min(int(com[1] + dv), halfv * 2))
    mm = np.zeros_like(image)
    mm[top:buttom, left:right] = 1
    image = image * mm
    keep = np.logical_and(image < com[2] + cube, image > com[2] - cube)
    return image * keep


# --------------------------------------------------------------------------- #
# OpenCV restatements
# --------------------------------------------------------------------------- #
def _linear_taps(dst, src):
    """cv::resize INTER_LINEAR coefficient table (imgproc/resize.cpp): scale =
    1/(dst/src) in double, fx = (float)((d+0.5)*scale-0.5), floor, clamp."""
    inv_scale = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv_scale
    idx = np.empty(dst, dtype=np.int64)
    a1 = np.empty(dst, dtype=np.float32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= src - 1:
            s, f = src - 1, np.float32(0)
        idx[d] = s
        a1[d] = f
    a0 = (np.float32(1.0) - a1).astype(np.float32)
    return idx, a0, a1


def resize_bilinear(src, dh=IMAGE_SIZE, dw=IMAGE_SIZE):
    """cv2.resize(src, (dw, dh)) INTER_LINEAR for CV_32F / CV_64F: horizontal
    2-tap pass then vertical 2-tap pass, coefficients held as float, products
    and sums in the working type (float for CV_32F, double for CV_64F), no
    antialiasing when shrinking."""
    sh, sw = src.shape
    wt = src.dtype.type
    xi, xa0, xa1 = _linear_taps(dw, sw)
    yi, ya0, ya1 = _linear_taps(dh, sh)
    xi1 = np.minimum(xi + 1, sw - 1)
    yi1 = np.minimum(yi + 1, sh - 1)
    xa0 = xa0.astype(wt)
    xa1 = xa1.astype(wt)
    ya0 = ya0.astype(wt)[:, None]
    ya1 = ya1.astype(wt)[:, None]
    top = src[yi][:, xi] * xa0 + src[yi][:, xi1] * xa1
    bot = src[yi1][:, xi] * xa0 + src[yi1][:, xi1] * xa1
    return (top * ya0 + bot * ya1).astype(wt)


def resize_half(img):
    """cv2.resize 128->64 INTER_LINEAR: OpenCV reroutes an exact 2x shrink to
    the INTER_AREA fast path = mean of each 2x2 block."""
    wt = img.dtype.type
    s = (img[0::2, 0::2] + img[0::2, 1::2]) + (img[1::2, 0::2] + img[1::2, 1::2])
    return (s * wt(0.25)).astype(wt)


# cv::getGaussianKernel(7, 1.5, CV_64F) as produced by OpenCV 4.13.0 (softdouble
# arithmetic, platform independent): exp(-x^2/(2 sigma^2)) / sum, x = -3..3.
_GAUSS7_HALF = (float.fromhex("0x1.2c18a51a3e5e7p-5"), float.fromhex("0x1.c7ce552574441p-4"),
                float.fromhex("0x1.bbe4f897eb627p-3"), float.fromhex("0x1.152db38ecae3ep-2"))


def gaussian_kernel7():
    """cv::getGaussianKernel(7, 1.5, CV_64F); the closed form agrees to 1 ulp."""
    h = _GAUSS7_HALF
    return np.array([h[0], h[1], h[2], h[3], h[2], h[1], h[0]], dtype=np.float64)


def _reflect101(i, n):
    if i < 0:
        return -i
    if i >= n:
        return 2 * n - 2 - i
    return i


def gaussian_blur7(h):
    """cv2.GaussianBlur(h, (7,7), 1.5) on a float64 plane: separable, row pass
    then column pass, BORDER_REFLECT_101."""
    n = h.shape[0]
    g = gaussian_kernel7()
    idx = np.array([[_reflect101(i + k - 3, n) for k in range(KSIZE)] for i in range(n)])
    rows = np.zeros_like(h)
    for k in range(KSIZE):
        rows += g[k] * h[:, idx[:, k]]
    out = np.zeros_like(h)
    for k in range(KSIZE):
        out += g[k] * rows[idx[:, k], :]
    return out


# --------------------------------------------------------------------------- #
# joints
# --------------------------------------------------------------------------- #
def splat4(u, v, size=LABEL_SIZE):
    """utils.py:37-62.  Centre-of-mass preserving 4-tap splat.  Negative
    indices wrap NumPy-style; an index >= size raises (the reference turns that
    into a rejected sample, datasets.py:362-365)."""
    h = np.zeros((size, size))
    low_u = int(np.floor(u))
    low_v = int(np.floor(v))
    du = u - low_u
    dv = v - low_v
    min_d = max(du + dv - 1, 0)
    max_d = min(du, dv)
    d = (max_d + min_d) / 2
    b = du - d
    c = dv - d
    a = 1 + d - du - dv
    h[low_v, low_u] = a
    h[low_v, low_u + 1] = b
    h[low_v + 1, low_u] = c
    h[low_v + 1, low_u + 1] = d
    return h


def joint_coords(uvd, com_int, box):
    """datasets.py:350-358.  Returns (centred+resized uvd [J,3] float64,
    heatmap pixel coords [J,2] float64)."""
    c = uvd.astype(np.float64) - com_int
    r = c.copy()
    r[:, :2] = r[:, :2] / (box - 1) * (IMAGE_SIZE - 1)
    k = r.copy()
    k[:, :2] = k[:, :2] / (IMAGE_SIZE - 1) * (LABEL_SIZE - 1) + np.array([LABEL_SIZE // 2, LABEL_SIZE // 2])
    return r, k


# --------------------------------------------------------------------------- #
# one sample, end to end
# --------------------------------------------------------------------------- #
# --------------------------------------------------------------------------- #
# augmentation (datasets.py:216-299, utils.py:67-82)
# --------------------------------------------------------------------------- #
def rotation_matrix(angle_deg, scale, center=(IMAGE_SIZE // 2, IMAGE_SIZE // 2)):
    """cv2.getRotationMatrix2D (utils.py:74): `angle *= CV_PI/180` folds the constant first."""
    a = angle_deg * (np.pi / 180)
    alpha = math.cos(a) * scale
    beta = math.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]],
                     [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]])


def invert_affine(M):
    """cv::invertAffineTransform, as cv::warpAffine applies it to a forward matrix."""
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22, A12, A21 = M[1, 1] * D, M[0, 0] * D, -M[0, 1] * D, -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]])


def warp_affine(src, M):
    """cv2.warpAffine(src, M, src.shape) with the defaults the reference uses (utils.py:75):
    INTER_LINEAR, BORDER_CONSTANT 0.  OpenCV evaluates the inverse map in 10-bit fixed
    point, quantises the source position to 1/32 pixel and blends the four neighbours with
    float table weights (1-fy/32)(1-fx/32)...; this restatement is bit-exact against OpenCV
    4.13.0 for float32 and float64 images (tests/test_oracle_sfr.py)."""
    size = src.shape[0]
    iM = invert_affine(M)
    ab_bits, inter_bits = 10, 5
    ab_scale = 1 << ab_bits
    round_delta = ab_scale // 32 // 2
    xs = np.arange(size)
    adelta = np.rint(iM[0, 0] * xs * ab_scale).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * ab_scale).astype(np.int64)
    X0 = np.rint((iM[0, 1] * xs + iM[0, 2]) * ab_scale).astype(np.int64) + round_delta      # per row y
    Y0 = np.rint((iM[1, 1] * xs + iM[1, 2]) * ab_scale).astype(np.int64) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (ab_bits - inter_bits)
    Y = (Y0[:, None] + bdelta[None, :]) >> (ab_bits - inter_bits)
    sx, sy, fx, fy = X >> inter_bits, Y >> inter_bits, X & 31, Y & 31
    tab = np.stack([1 - np.arange(32, dtype=np.float32) / np.float32(32), np.arange(32, dtype=np.float32) / np.float32(32)], 1)
    padded = np.zeros((size + 2, size + 2), src.dtype)
    padded[1:-1, 1:-1] = src

    def S(r, c):
        ok = (r >= -1) & (r <= size) & (c >= -1) & (c <= size)
        return np.where(ok, padded[np.clip(r + 1, 0, size + 1), np.clip(c + 1, 0, size + 1)], src.dtype.type(0))
    w00 = tab[fy, 0] * tab[fx, 0]
    w01 = tab[fy, 0] * tab[fx, 1]
    w10 = tab[fy, 1] * tab[fx, 0]
    w11 = tab[fy, 1] * tab[fx, 1]
    out = S(sy, sx) * w00 + S(sy, sx + 1) * w01 + S(sy + 1, sx) * w10 + S(sy + 1, sx + 1) * w11
    return out.astype(src.dtype)


def rotate_joints(r, angle_deg, scale):
    """utils.py:77-80: uvd[:, :2] @ Rot.T * scale (written out without FMA)."""
    a = angle_deg / 180.0 * np.pi
    c, s = np.cos(a), np.sin(a)
    u, v = r[:, 0].copy(), r[:, 1].copy()
    r = r.copy()
    r[:, 0] = (u * c + v * s) * scale
    r[:, 1] = (u * -s + v * c) * scale
    return r


def _process(frame, uvd, com, cube, fx, fy, test_only, backend, aug):
    """One pass of process_single_data; `aug` None = the non-augmented branch
    (datasets.py:301-365), else (scale, shift_u, shift_v, angle_deg) = the augmented branch
    (:216-299).  Returns (out dict, raised) where `raised` mirrors a Python exception."""
    if backend == "cv2":
        import cv2
        do_resize = lambda s: cv2.resize(s, (IMAGE_SIZE, IMAGE_SIZE))
        do_half = lambda s: cv2.resize(s, (LABEL_SIZE, LABEL_SIZE))
        do_blur = lambda s: cv2.GaussianBlur(s, (KSIZE, KSIZE), SIGMA)
        do_warp = lambda s, M: cv2.warpAffine(s, M, (IMAGE_SIZE, IMAGE_SIZE))
    else:
        do_resize, do_half, do_blur, do_warp = resize_bilinear, resize_half, gaussian_blur7, warp_affine

    J = 0 if uvd is None else uvd.shape[0]
    f32 = np.float32
    out = dict(valid=False,
               img=np.zeros((1, IMAGE_SIZE, IMAGE_SIZE), f32),
               label_img=np.zeros((1, LABEL_SIZE, LABEL_SIZE), f32),
               mask=np.zeros((1, LABEL_SIZE, LABEL_SIZE), f32),
               box_size=f32(0), cube_size=f32(cube), com=np.zeros(3, f32),
               uvd=np.zeros((J, 3), f32),
               heatmaps=np.zeros((J, LABEL_SIZE, LABEL_SIZE), f32),
               dmap=np.zeros((J, LABEL_SIZE, LABEL_SIZE), f32))
    if com is None:
        com = com_from_frame(frame)
    com = np.array(com, dtype=np.float64)
    # integral cube sizes are Python ints in the reference (weak scalars)
    cube_w = int(cube) if float(cube) == int(cube) else float(cube)
    if aug is not None:
        # datasets.py:234-241: uvd2xyz / xyz2uvd do nothing to a 1-D array, so the "mm" shift is
        # applied to the pixel coordinates of the centre as they are
        scale, shift_u, shift_v, angle = (float(a) for a in aug)
        com[0] += shift_u
        com[1] += shift_v

    box = crop_box(com[2], cube_w, fx, fy)
    crop = center_crop(frame, com, box)
    if crop.shape[0] == 0 or crop.shape[1] == 0:
        return out, True                            # cv2.resize raises
    crop = window_and_centre(crop.copy(), com[2], cube_w)
    com[0] = int(com[0])
    com[1] = int(com[1])
    box = crop.shape[0]

    img = do_resize(crop)
    wt = img.dtype.type
    if aug is not None:
        r, _ = joint_coords(uvd, com, box)
        img = do_warp(img, rotation_matrix(angle, scale))                  # utils.py:74-75
        r = rotate_joints(r, angle, scale)                                 # utils.py:77-80
        img = (img * wt(scale)).astype(wt)                                 # datasets.py:284
        r[:, 2] *= scale                                                   # datasets.py:285
        k = r.copy()
        k[:, :2] = k[:, :2] / (IMAGE_SIZE - 1) * (LABEL_SIZE - 1) + np.array([LABEL_SIZE // 2, LABEL_SIZE // 2])
    label = do_half(img)
    mask = (label != 0).astype(np.float64)

    out["img"] = (img / wt(cube_w)).astype(f32)[None]
    out["label_img"] = (label / wt(cube_w)).astype(f32)[None]
    out["mask"] = mask.astype(f32)[None]
    out["box_size"] = f32(box)
    out["com"] = com.astype(f32)
    bad = bool(np.isnan(out["img"]).any() or np.isnan(out["label_img"]).any())
    if test_only:
        # datasets.py:334-348 returns before the reject gate
        out["valid"] = True
        return out, False

    if aug is None:
        r, k = joint_coords(uvd, com, box)
    heat = np.zeros((LABEL_SIZE, LABEL_SIZE, J))
    try:
        for j in range(J):
            heat[:, :, j] = do_blur(splat4(k[j, 0], k[j, 1]))
    except (IndexError, ValueError, OverflowError):
        return out, True                            # "heatmap error"
    dmap = np.zeros_like(heat)
    for j in range(J):
        heatmask = (heat[:, :, j] > 0).astype(np.float64) * mask
        dmap[:, :, j] = (r[j, 2] - label) * heatmask
    dmap = dmap / cube_w
    nuvd = r.copy()
    nuvd[:, :2] = nuvd[:, :2] / (IMAGE_SIZE - 1)
    nuvd[:, 2] = nuvd[:, 2] / cube_w

    out["uvd"] = nuvd.astype(f32)
    out["heatmaps"] = np.ascontiguousarray(heat.transpose(2, 0, 1)).astype(f32)
    out["dmap"] = np.ascontiguousarray(dmap.transpose(2, 0, 1)).astype(f32)
    bad = bad or bool(np.isnan(nuvd).any() or np.isnan(heat).any() or np.isnan(dmap).any())
    out["valid"] = (not bad) and float(mask.sum()) >= 10
    return out, False


def process_sample(frame, uvd, com, cube, fx, fy, test_only=False, backend="numpy", aug=None):
    """HandDataset.process_single_data (datasets.py:185-403).  `frame` float32
    (NYU/ICVL/HAND17 semantics) or float64 (MSRA semantics); `com` None -> CoM fallback.
    `aug` = (scale, shift_u, shift_v, angle_deg): the augmented branch with these draws; as
    in the reference (`try:` :216 / `except:` :301) anything that raises inside it silently
    falls back to the non-augmented branch.  Returns a dict with float32 arrays named like
    the reference tuple plus `valid` (False where the reference raises)."""
    if aug is not None:
        out, raised = _process(frame, uvd, com, cube, fx, fy, test_only, backend, aug)
        if not raised:
            return out
    return _process(frame, uvd, com, cube, fx, fy, test_only, backend, None)[0]


FIELDS = ("img", "label_img", "mask", "box_size", "cube_size", "com", "uvd", "heatmaps", "dmap")


def process_batch(frames, uvd, com, cube, fx, fy, test_only=False, backend="numpy", aug=None):
    """Stack process_sample over a batch (the DataLoader's default_collate)."""
    B = len(frames)
    outs = [process_sample(frames[b], None if uvd is None else uvd[b],
                           None if com is None else com[b], cube[b], fx, fy,
                           test_only=test_only, backend=backend, aug=None if aug is None else aug[b])
            for b in range(B)]
    res = {k: np.stack([o[k] for o in outs]) for k in FIELDS}
    res["valid"] = np.array([o["valid"] for o in outs], dtype=np.uint8)
    return res


def recover_uvd(uvd, box_size, com, cube):
    """utils.py:332-337 (inverse of the uvd normalisation), float32."""
    out = uvd.astype(np.float32).copy()
    out[:, :, :2] = out[:, :, :2] * (box_size.astype(np.float32) - 1).reshape(-1, 1, 1)
    out[:, :, 2] = out[:, :, 2] * cube.astype(np.float32)[:, None]
    return out + com.astype(np.float32)[:, None, :]


def uvd2xyz(uvd, fx, fy, halfu, halfv):
    """datasets.py:100-111 on a float32 [B,J,3] array (NumPy float32 arithmetic)."""
    x = uvd.copy()
    x[:, :, 0] = (x[:, :, 0] - halfu) / fx * x[:, :, 2]
    x[:, :, 1] = (x[:, :, 1] - halfv) / fy * x[:, :, 2]
    return x


def joint_error(uvd_pred, uvd_true, box_size, com, cube, fx, fy, halfu, halfv):
    """train.py:254-276: per-sample mean joint error (mm) from normalised uvd."""
    p = uvd2xyz(recover_uvd(uvd_pred, box_size, com, cube), fx, fy, halfu, halfv)
    t = uvd2xyz(recover_uvd(uvd_true, box_size, com, cube), fx, fy, halfu, halfv)
    return np.mean(np.sqrt(np.sum((p - t) ** 2, axis=2)), axis=1)
