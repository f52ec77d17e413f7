// SFR (Spatial-Form Representation) target builder for sm_100a.
//
// Turns raw depth frames + joint annotations into the reference's training
// tuple (datasets.py:301-403): crop / window / centre / resize the depth frame,
// build the label image and mask, render per-joint Gaussian heat maps and depth
// maps.  In the reference this is ~40-60 ms of NumPy/OpenCV per sample inside
// DataLoader workers.
//
// Two kernels per call, sharing a caller-provided workspace:
//   sfr_prep_kernel   one warp per sample: lane 0 derives the crop geometry,
//                     lanes 0..J-1 the splat taps of one joint each, all in
//                     float64 with the reference's operation order; writes the
//                     scalar outputs (box, cube, int CoM, normalised uvd) and
//                     zeroes the sample's gate counter.  Dependent float64
//                     division chains are latency-, not throughput-limited, so
//                     they are done ONCE per sample here instead of stalling
//                     every CTA of the main kernel.
//   sfr_build_kernel  kBands (8 with dense maps, 2 without) CTAs per sample, one per horizontal band of the
//                     label image.  Each CTA (a) streams zeros over its band
//                     of all 2J maps with 128-bit stores while the prepared
//                     geometry is fetched, (b) resamples its 2*R x 128 image
//                     rows straight from the frame in HBM (window + centring
//                     per tap), writes img / label / mask and keeps the
//                     un-normalised label band in shared memory, (c) patches
//                     the <= 8x8 float64 footprint of every joint that touches
//                     the band.  The per-sample reject gate (sum(mask) < 10,
//                     NaN; datasets.py:385-390) is one packed atomicAdd per
//                     CTA on the sample's counter; the CTA that arrives last
//                     writes `valid`.
//   sfr_aug_kernel    the augmented branch (datasets.py:216-299): one CTA per
//                     sample, resized crop staged in shared memory, rotation +
//                     scale through cv2.warpAffine's fixed-point bilinear map.
// Frames may be decoded float32 or raw PNG samples (NYU G/B, 16-bit grey)
// decoded per tap; the hand rectangle of load_from_text tightens the tap bounds.
//
// Arithmetic contract (bit-level where the result is discontinuous):
//   * crop box, int CoM, slice extents: float64 with explicit _rn intrinsics
//     (no FMA contraction) and Python slice semantics -> bit-exact integers.
//   * image path in the frame's reference dtype (float32, or float64 for MSRA):
//     cv::resize INTER_LINEAR coefficient recipe, horizontal then vertical,
//     un-fused multiplies/adds; 2x2 mean; mask = (label != 0).
//   * joints, splat taps, Gaussian (cv::getGaussianKernel(7,1.5) constants,
//     BORDER_REFLECT_101) and Dmap in float64, rounded to float32 at the store.
#include "common.cuh"

namespace pwr {

int get_option(int option);      // dispatch overrides, defined in decoder.cu (pwr_set_option)

// CTAs (horizontal bands of the label image) per sample.  With dense maps a CTA's band of all 2J maps
// is zero-filled while the taps are fetched, and 8 bands measured best (2/4/8/16: 0.60/0.59/0.54/0.55 ms
// at B=4096 NYU); without them (test-only SFR, compact targets) nothing overlaps that prologue and
// fewer, longer CTAs win (8/4/2: 0.289/0.258/0.250 ms HAND17 crop).
#ifndef PWR_SFR_BANDS
#define PWR_SFR_BANDS 8
#endif
#ifndef PWR_SFR_BANDS_LEAN
#define PWR_SFR_BANDS_LEAN 2
#endif
constexpr int kBandsDense = PWR_SFR_BANDS;
constexpr int kBandsLean = PWR_SFR_BANDS_LEAN;
#ifndef PWR_SFR_BANDS_STAGED_U16
#define PWR_SFR_BANDS_STAGED_U16 8
#endif
#ifndef PWR_SFR_BANDS_STAGED_F32
#define PWR_SFR_BANDS_STAGED_F32 16
#endif
constexpr int kBandsStagedU16 = PWR_SFR_BANDS_STAGED_U16;
constexpr int kBandsStagedF32 = PWR_SFR_BANDS_STAGED_F32;
constexpr int kBandsMax = 16;

// cv::getGaussianKernel(7, 1.5, CV_64F), OpenCV 4.13.0 (softdouble, exact bits)
__constant__ double kGauss7[7] = {0x1.2c18a51a3e5e7p-5, 0x1.c7ce552574441p-4, 0x1.bbe4f897eb627p-3,
                                  0x1.152db38ecae3ep-2, 0x1.bbe4f897eb627p-3, 0x1.c7ce552574441p-4,
                                  0x1.2c18a51a3e5e7p-5};

struct SampleGeom {
    double z, cube;
    double scale_y, scale_x;   // 1 / (128 / n), as cv::resize computes it
    float lo_dn, hi_up;        // float-exact equivalents of the float64 window bounds
    double lo, hi;
    int ok;                    // crop non-empty and all scalars finite
    int fr0, fc0;              // frame row / column of crop(0,0) (may be negative)
    int nrows, ncols;          // crop extent (== 2*shift unless truncated)
    int r0, c0;                // int(com_v), int(com_u)
    int pr0, pr1, pc0, pc1;    // frame rows [pr0,pr1) x cols [pc0,pc1) that may be non-zero: the whole frame,
                               // or the hand rectangle of load_from_text when the prefilter is on; in window
                               // mode additionally clipped to the part of the frame that was fetched
    int org_r, org_c;          // frame row / column of element (0,0) of this sample's pixel buffer: (0,0) for
                               // whole frames, the window origin written by pwr_sfr_fetch in window mode
    int pad;                   // explicit tail padding (the struct is copied word by word; keeps initcheck clean)
};
static_assert(sizeof(SampleGeom) == 112, "no implicit padding");

// Part of frame b that pwr_sfr_fetch copied into windows[b]: rows [row0, row0+rows) x cols [col0, col0+cols).
struct WinExtent { int row0, col0, rows, cols; };

struct JointParam {
    double tap[4];             // a, b, c, d of utils.py:48-58
    double cd;                 // centred depth  uvd_z - com_z
    int tx0, tx1, ty0, ty1;    // wrapped heat-map indices of the four taps
    int ok;
    int pad[3];                // explicit tail padding, see SampleGeom
};
static_assert(sizeof(JointParam) == 72, "no implicit padding");

struct TapX { int s0, s1; float a0, a1; };

// Augmentation of one sample (datasets.py:216-299): rotation + scale about the centre of the
// 128x128 image through cv2.warpAffine's inverse fixed-point map, intensity * scale.
struct WarpParam {
    double iM[6];              // inverse affine map, cv::invertAffineTransform of getRotationMatrix2D
    double scale;
    int active;                // 0: the augmented branch raised -> plain branch (datasets.py:301)
    int pad;
};

// mbarrier / bulk-TMA helpers (same forms as decoder.cu)
__device__ __forceinline__ uint32_t sfr_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sfr_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sfr_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sfr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sfr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sfr_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "SFR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra SFR_DONE;\n"
        "bra SFR_WAIT;\n"
        "SFR_DONE:\n"
        "}" ::"r"(sfr_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sfr_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sfr_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sfr_smem_u32(bar)) : "memory");
}

// Python slice(start, stop).indices(n) for step 1
__device__ __forceinline__ void py_slice(long long start, long long stop, long long n, int& first, int& count) {
    if (start < 0) { start += n; if (start < 0) start = 0; } else if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; } else if (stop > n) stop = n;
    first = static_cast<int>(start);
    count = stop > start ? static_cast<int>(stop - start) : 0;
}

// int(x) for a float64 that may be far outside the int range (Python ints are unbounded; every use
// below is followed by a clamp to the frame, so saturating first gives the same result)
__device__ __forceinline__ long long py_int(double x) {
    if (x > 1.0e12) x = 1.0e12;
    if (x < -1.0e12) x = -1.0e12;
    return static_cast<long long>(x);
}

__device__ void sample_geometry(SampleGeom& g, const double* com, double cube, double fx, double fy, int Hf, int Wf,
                                double pf_margin, double pf_umax, double pf_vmax) {
    const double cu = com[0], cv = com[1], z = com[2];
    g.z = z; g.cube = cube; g.ok = 0;
    g.org_r = 0; g.org_c = 0; g.pad = 0;
    g.pr0 = 0; g.pr1 = Hf; g.pc0 = 0; g.pc1 = Wf;
    g.fr0 = g.fc0 = 0; g.nrows = g.ncols = 0; g.r0 = g.c0 = 0;
    g.scale_x = g.scale_y = 1.0;
    // datasets.py:306-309
    const double du = __dmul_rn(__ddiv_rn(cube, z), fx);
    const double dv = __dmul_rn(__ddiv_rn(cube, z), fy);
    const double sum = __dadd_rn(du, dv);
    g.lo = __dsub_rn(z, cube);
    g.hi = __dadd_rn(z, cube);
    g.lo_dn = __double2float_rd(g.lo);
    g.hi_up = __double2float_ru(g.hi);
    if (!isfinite(sum) || !isfinite(cu) || !isfinite(cv) || !isfinite(z) || !isfinite(cube)) return;  // int(nan) raises
    if (fabs(sum) > 1.0e6 || fabs(cu) > 1.0e9 || fabs(cv) > 1.0e9) return;
    int box = static_cast<int>(sum);          // int() truncates toward zero
    if (box < 2) box = 2;
    const int shift = box / 2;
    g.r0 = static_cast<int>(cv);
    g.c0 = static_cast<int>(cu);
    // utils.py:167-173: slice of the zero-padded frame
    int rs, cs;
    py_slice(g.r0, static_cast<long long>(g.r0) + 2 * shift, static_cast<long long>(Hf) + 2 * shift, rs, g.nrows);
    py_slice(g.c0, static_cast<long long>(g.c0) + 2 * shift, static_cast<long long>(Wf) + 2 * shift, cs, g.ncols);
    g.fr0 = rs - shift;
    g.fc0 = cs - shift;
    if (pf_margin >= 0.0) {
        // load_from_text (NYU datasets.py:841-853, HAND17 :956-968, ICVL :666-678): zero everything
        // outside MM[top:buttom, left:right]; Python slice semantics included
        const double pdu = __dmul_rn(__ddiv_rn(__dsub_rn(cube, pf_margin), z), fx);
        const double pdv = __dmul_rn(__ddiv_rn(__dsub_rn(cube, pf_margin), z), fy);
        long long left = py_int(__dsub_rn(cu, pdu)), right = py_int(__dadd_rn(cu, pdu));
        long long top = py_int(__dsub_rn(cv, pdv)), buttom = py_int(__dadd_rn(cv, pdv));
        if (left < 0) left = 0;
        if (top < 0) top = 0;
        right = py_int(fmin(static_cast<double>(right), pf_umax));
        buttom = py_int(fmin(static_cast<double>(buttom), pf_vmax));
        int first, count;
        py_slice(top, buttom, Hf, first, count);
        g.pr0 = first; g.pr1 = first + count;
        py_slice(left, right, Wf, first, count);
        g.pc0 = first; g.pc1 = first + count;
    }
    if (g.nrows == 0 || g.ncols == 0) return;  // cv2.resize raises on an empty crop
    g.scale_y = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(kImage), static_cast<double>(g.nrows)));
    g.scale_x = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(kImage), static_cast<double>(g.ncols)));
    g.ok = 1;
}

// Window mode: the pixel buffer of the sample holds only `w` of the frame.  Every tap of the crop lies inside
// the box, and pwr_sfr_fetch copied (box AND non-zero rectangle) rounded outwards to 16 bytes, so clipping the
// non-zero rectangle to the fetched extent changes no tap's value and guarantees no read outside the buffer.
__device__ __forceinline__ void clip_to_window(SampleGeom& g, const WinExtent& w) {
    g.org_r = w.row0; g.org_c = w.col0;
    g.pr0 = max(g.pr0, w.row0); g.pr1 = min(g.pr1, w.row0 + w.rows);
    g.pc0 = max(g.pc0, w.col0); g.pc1 = min(g.pc1, w.col0 + w.cols);
}

// cv::resize INTER_LINEAR tap for destination index d, source extent n
__device__ __forceinline__ TapX linear_tap(int d, int n, double scale) {
    float f = __double2float_rn(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
    int s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n - 1) { s = n - 1; f = 0.f; }
    TapX t;
    t.s0 = s;
    t.s1 = min(s + 1, n - 1);
    t.a1 = f;
    t.a0 = __fsub_rn(1.f, f);
    return t;
}

// utils.py:37-62 + datasets.py:350-358 for one joint
// `rot` = {cos, sin, scale} of the augmentation (utils.py:77-80, datasets.py:285) or nullptr.
__device__ void joint_param(JointParam& p, float* uvd_norm_out, const double* uvd, const SampleGeom& g,
                            const double* rot = nullptr) {
    const double cu = __dsub_rn(uvd[0], static_cast<double>(g.c0));
    const double cv = __dsub_rn(uvd[1], static_cast<double>(g.r0));
    double cd = __dsub_rn(uvd[2], g.z);
    const double bm1 = static_cast<double>(g.nrows - 1);          // box_size = crop.shape[0]
    double ru = __dmul_rn(__ddiv_rn(cu, bm1), 127.0);
    double rv = __dmul_rn(__ddiv_rn(cv, bm1), 127.0);
    if (rot != nullptr) {
        // uvd[:, :2] @ Rot.T with Rot = [[c, s], [-s, c]], then * scale; depth * scale
        const double c = rot[0], sn = rot[1], sc = rot[2];
        const double u2 = __dmul_rn(__dadd_rn(__dmul_rn(ru, c), __dmul_rn(rv, sn)), sc);
        const double v2 = __dmul_rn(__dadd_rn(__dmul_rn(ru, -sn), __dmul_rn(rv, c)), sc);
        ru = u2; rv = v2;
        cd = __dmul_rn(cd, sc);
    }
    const double ku = __dadd_rn(__dmul_rn(__ddiv_rn(ru, 127.0), 63.0), 32.0);
    const double kv = __dadd_rn(__dmul_rn(__ddiv_rn(rv, 127.0), 63.0), 32.0);
    p.cd = cd;
    p.ok = 0; p.pad[0] = p.pad[1] = p.pad[2] = 0;
    p.tx0 = p.tx1 = p.ty0 = p.ty1 = 0;
    p.tap[0] = p.tap[1] = p.tap[2] = p.tap[3] = 0.0;
    const double nu = __ddiv_rn(ru, 127.0), nv = __ddiv_rn(rv, 127.0), nd = __ddiv_rn(cd, g.cube);
    if (uvd_norm_out != nullptr) {
        uvd_norm_out[0] = __double2float_rn(nu);
        uvd_norm_out[1] = __double2float_rn(nv);
        uvd_norm_out[2] = __double2float_rn(nd);
    }
    if (!isfinite(ku) || !isfinite(kv) || isnan(nd)) return;       // int(floor(nan/inf)) raises / NaN gate
    const double flu = floor(ku), flv = floor(kv);
    if (flu < -64.0 || flu > 62.0 || flv < -64.0 || flv > 62.0) return;   // IndexError -> "Out of range"
    const int lu = static_cast<int>(flu), lv = static_cast<int>(flv);
    const double du = __dsub_rn(ku, flu), dv = __dsub_rn(kv, flv);
    const double t = __dsub_rn(__dadd_rn(du, dv), 1.0);
    const double min_d = (0.0 > t) ? 0.0 : t;                      // max(du + dv - 1, 0)
    const double max_d = (dv < du) ? dv : du;                      // min(du, dv)
    const double d = __ddiv_rn(__dadd_rn(max_d, min_d), 2.0);
    p.tap[1] = __dsub_rn(du, d);
    p.tap[2] = __dsub_rn(dv, d);
    p.tap[0] = __dsub_rn(__dsub_rn(__dadd_rn(1.0, d), du), dv);
    p.tap[3] = d;
    p.tx0 = lu < 0 ? lu + kLabel : lu;                             // negative indices wrap NumPy-style
    p.tx1 = lu + 1 < 0 ? lu + 1 + kLabel : lu + 1;
    p.ty0 = lv < 0 ? lv + kLabel : lv;
    p.ty1 = lv + 1 < 0 ? lv + 1 + kLabel : lv + 1;
    p.ok = 1;
}

// weight with which a unit impulse at index t reaches output index x through the
// 7-tap Gaussian with BORDER_REFLECT_101 on a 64-long axis
__device__ __forceinline__ double gauss_reach(int x, int t) {
    double wgt = 0.0;
    const int k0 = t - x + 3;                   // direct
    if (k0 >= 0 && k0 <= 6) wgt += kGauss7[k0];
    const int k1 = 3 - x - t;                   // reflected at the low border: -(x+k-3) == t
    if (t >= 1 && k1 >= 0 && k1 <= 6) wgt += kGauss7[k1];
    const int k2 = 2 * kLabel + 1 - x - t;      // reflected at the high border: 126-(x+k-3) == t
    if (t <= kLabel - 2 && k2 >= 0 && k2 <= 6) wgt += kGauss7[k2];
    return wgt;
}

// ---- working-type arithmetic (float for float32 frames, double for MSRA) ----
template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float rcp(float b) { return __frcp_rn(b); }
    // a / b, correctly rounded, from rb = RN(1/b): q = RN(a*rb) is within 1 ulp, the FMA residual is
    // exact, and one correction step lands on RN(a/b) (Markstein).  |a| < cube-ish and b = cube, so no
    // overflow / denormal corner can occur; NaN propagates like the true division.
    static __device__ __forceinline__ float div_by(float a, float b, float rb) {
        const float q = __fmul_rn(a, rb);
        const float r = __fmaf_rn(-q, b, a);
        return __fmaf_rn(r, rb, q);
    }
    // window + centring of one frame pixel, datasets.py:312,315
    static __device__ __forceinline__ float window(float v, const SampleGeom& g) {
        const float keep = (v > g.lo_dn && v < g.hi_up) ? 1.f : 0.f;   // == float64 compare, see SampleGeom
        v = __fmul_rn(v, keep);
        if (v > 0.f) v = __double2float_rn(__dsub_rn(static_cast<double>(v), g.z));
        return v;
    }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double rcp(double b) { return b; }
    static __device__ __forceinline__ double div_by(double a, double b, double) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double window(double v, const SampleGeom& g) {
        const double keep = (v > g.lo && v < g.hi) ? 1.0 : 0.0;
        v = __dmul_rn(v, keep);
        if (v > 0.0) v = __dsub_rn(v, g.z);
        return v;
    }
};

struct SfrArgs {
    const void* frames; int Hf, Wf;
    const double* com; const double* cube; const double* uvd;
    double fx, fy;
    float* img; float* label_img; float* mask;
    float* box_size; float* cube_size; float* com_out;
    float* uvd_norm; float* heatmaps; float* dmap;
    uint8_t* valid;
    int B, J;
    SampleGeom* prep_geom;      // workspace: [B]
    JointParam* prep_joints;    // workspace: [B*J]
    int* prep_flags;            // workspace: [B] joints_bad
    unsigned int* gate;         // workspace: [B] packed (bands arrived << 24 | NaN << 16 | mask count)
    double pf_margin, pf_umax, pf_vmax;   // load_from_text prefilter: margin < 0 = off; 2*halfu, 2*halfv
    pwr_joint_taps* taps_out;   // [B,J] compact ("sparse") targets, or NULL
    const double* aug;          // [B,8] (scale, shift_u, shift_v, cos a, sin a, cos a', sin a', -) or NULL
    WarpParam* prep_warp;       // workspace: [B] (augmentation only)
    const WinExtent* win;       // [B] window mode (frames = [B, win_h, win_w] windows written by pwr_sfr_fetch) or NULL
    int pitch;                  // elements per row of a sample's pixel buffer: Wf, or win_w
    long long frame_stride;     // elements per sample: Hf*Wf, or win_h*win_w
};

// Raw sensor formats.  The float32 value the reference would hold is reproduced bit
// for bit: plt.imread turns PNG samples into float32(v / 255) resp. float32(v / 65535).
//   FMT_F32   frames already decoded (what process_single_data receives)
//   FMT_GB16  NYU: uint16 = G << 8 | B of the PNG;  depth = ((G/255)*256 + B/255)*255, datasets.py:810
//   FMT_U16   16-bit grey PNG (ICVL :632, HAND17 :940): (v/65535)*65535 == v exactly for all 65536 values
enum { FMT_F32 = 0, FMT_GB16 = 1, FMT_U16 = 2 };

// One pixel as it sits in memory (float bits for FMT_F32, the 16-bit sample otherwise) and its decode: split so
// that a block of taps can issue all its loads before any decode arithmetic (loads in flight are what the
// gather lives on).
template <int FMT, bool SMEM = false>
__device__ __forceinline__ unsigned int load_raw(const void* __restrict__ frame, int idx) {
    if (SMEM) {     // staged source rows: ordinary (shared-memory) loads
        if (FMT == FMT_F32) return static_cast<const unsigned int*>(frame)[idx];
        return static_cast<const unsigned short*>(frame)[idx];
    }
    if (FMT == FMT_F32) return __float_as_uint(__ldg(static_cast<const float*>(frame) + idx));
    return __ldg(static_cast<const unsigned short*>(frame) + idx);
}
template <int FMT>
__device__ __forceinline__ float decode_px(unsigned int v) {
    if (FMT == FMT_F32) return __uint_as_float(v);
    if (FMT == FMT_U16) return static_cast<float>(v);
    // x / 255 correctly rounded through the reciprocal (exact for x = 0..255, checked exhaustively)
    const float rc = 0x1.010102p-8f;                     // RN(1/255)
    const float gq = static_cast<float>(v >> 8), bq = static_cast<float>(v & 255u);
    float g = __fmul_rn(gq, rc); g = __fmaf_rn(__fmaf_rn(-g, 255.f, gq), rc, g);
    float bl = __fmul_rn(bq, rc); bl = __fmaf_rn(__fmaf_rn(-bl, 255.f, bq), rc, bl);
    return __fmul_rn(__fadd_rn(__fmul_rn(g, 256.f), bl), 255.f);
}
template <int FMT>
__device__ __forceinline__ float load_px(const void* __restrict__ frame, int idx) {
    return decode_px<FMT>(load_raw<FMT>(frame, idx));
}

// Bilinear taps of one 2x2 image block (one label pixel), gathered from the frame with the
// depth window + centring applied per tap.  INTERIOR: every source row / column of this band
// lies inside the frame, so no bounds predicates are needed.
// Raw 16-bit frames: inside the depth window a sample takes at most 2*cube + 1 distinct raw values, so decode
// (datasets.py:810 / :940) + window + centring (:312-315) of a tap collapse into one shared-memory lookup,
// tab[raw - base] (0 outside the table = outside the window).  Each CTA fills the table for its sample with the
// very functions the per-tap path uses, so the values are bit-identical; what goes away is the per-tap
// float32 -> float64 -> float32 round trip (XU pipe: 27 % busy in the float32-frame kernel, ncu r1) and the
// PNG decode arithmetic (which made the raw-frame build SLOWER than the float32 one in r1: 0.59 vs 0.54 ms).
constexpr int kWinTab = 768;                       // covers cube <= 380 mm; larger cubes use the per-tap arithmetic
struct WinTab { const float* tab; int base; int n; };
// tab[n] is a 0.f sentinel: every raw value outside the table clamps onto it, so the lookup is one unconditional
// shared-memory load (r2: the predicated form made ptxas rebuild the shared-window address per tap - S2R + MOV +
// VIADD + LEA under the predicate, 64 of the 252 instructions of a 2x2 block).
__device__ __forceinline__ float tab_lookup(unsigned int raw, const WinTab& t) {
    const unsigned int i = min(raw - static_cast<unsigned int>(t.base), static_cast<unsigned int>(t.n));
    return t.tab[i];
}

template <typename T, int FMT, bool INTERIOR, bool TAB = false, bool SMEM = false>
__device__ __forceinline__ void resample_block(T (&px)[2][2], const void* __restrict__ frame, const SampleGeom& g,
                                               const TapX* ytap2, const TapX* xtap2, int pitch,
                                               const WinTab& wt = WinTab{nullptr, 0, 0}) {
    // Phase A: all 16 tap loads of the block are issued before any of them is consumed (the kernel lives
    // on loads in flight: with 8 + 8 interleaved with arithmetic it measured 8 % slower, r2).
    unsigned int v[2][2][4];                       // raw samples; 0 decodes to 0.f in every format
    TapX ty[2], tx[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) { ty[d] = ytap2[d]; tx[d] = xtap2[d]; }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int fr_a = g.fr0 + ty[dy].s0, fr_b = g.fr0 + ty[dy].s1;
        const bool ra = INTERIOR || (fr_a >= g.pr0 && fr_a < g.pr1), rb = INTERIOR || (fr_b >= g.pr0 && fr_b < g.pr1);
        // element offsets inside the sample's pixel buffer (Hf*Wf < 2^31, checked on the host)
        const int row_a = (fr_a - g.org_r) * pitch - g.org_c, row_b = (fr_b - g.org_r) * pitch - g.org_c;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int fc_a = g.fc0 + tx[dx].s0, fc_b = g.fc0 + tx[dx].s1;
            const bool ca = INTERIOR || (fc_a >= g.pc0 && fc_a < g.pc1), cb = INTERIOR || (fc_b >= g.pc0 && fc_b < g.pc1);
            v[dy][dx][0] = (ra && ca) ? load_raw<FMT, SMEM>(frame, row_a + fc_a) : 0u;
            v[dy][dx][1] = (ra && cb) ? load_raw<FMT, SMEM>(frame, row_a + fc_b) : 0u;
            v[dy][dx][2] = (rb && ca) ? load_raw<FMT, SMEM>(frame, row_b + fc_a) : 0u;
            v[dy][dx][3] = (rb && cb) ? load_raw<FMT, SMEM>(frame, row_b + fc_b) : 0u;
        }
    }
    // Phase B: window + centring per tap, horizontal then vertical lerp (cv::resize's order, un-fused)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            T w00, w01, w10, w11;
            if (TAB) {
                w00 = static_cast<T>(tab_lookup(v[dy][dx][0], wt)); w01 = static_cast<T>(tab_lookup(v[dy][dx][1], wt));
                w10 = static_cast<T>(tab_lookup(v[dy][dx][2], wt)); w11 = static_cast<T>(tab_lookup(v[dy][dx][3], wt));
            } else {
                w00 = Arith<T>::window(static_cast<T>(decode_px<FMT>(v[dy][dx][0])), g);
                w01 = Arith<T>::window(static_cast<T>(decode_px<FMT>(v[dy][dx][1])), g);
                w10 = Arith<T>::window(static_cast<T>(decode_px<FMT>(v[dy][dx][2])), g);
                w11 = Arith<T>::window(static_cast<T>(decode_px<FMT>(v[dy][dx][3])), g);
            }
            const T xa0 = static_cast<T>(tx[dx].a0), xa1 = static_cast<T>(tx[dx].a1);
            const T top = Arith<T>::add(Arith<T>::mul(w00, xa0), Arith<T>::mul(w01, xa1));
            const T bot = Arith<T>::add(Arith<T>::mul(w10, xa0), Arith<T>::mul(w11, xa1));
            px[dy][dx] = Arith<T>::add(Arith<T>::mul(top, static_cast<T>(ty[dy].a0)), Arith<T>::mul(bot, static_cast<T>(ty[dy].a1)));
        }
    }
}

// Phase 2, pass B: the float64 footprint of the listed joints inside label rows [y_lo, y_hi):
// <= 11 candidate rows x 11 candidate columns per joint (offsets -3..3 around tap 0, then 0..3
// around tap 1), most rejected by integer tests.  `label_s` holds rows y_lo.. of the un-normalised label.
template <typename T>
__device__ __forceinline__ void patch_footprints(const SfrArgs& a, const SampleGeom& g, const JointParam* joints,
                                                 const int* list, int list_n, const T* label_s, int b, int y_lo,
                                                 int y_hi, int tid, int nthreads) {
    constexpr int kCand = 11;
    const int total = list_n * kCand * kCand;
    for (int c = tid; c < total; c += nthreads) {
        const int jl = c / (kCand * kCand);
        const int r = c - jl * (kCand * kCand);
        const int sy = r / kCand, sx = r - sy * kCand;
        const int j = list[jl];
        const JointParam& jp = joints[j];
        const int y = sy < 7 ? jp.ty0 + sy - 3 : jp.ty1 + sy - 7;
        const int x = sx < 7 ? jp.tx0 + sx - 3 : jp.tx1 + sx - 7;
        if (y < y_lo || y >= y_hi || x < 0 || x >= kLabel) continue;
        if (sy >= 7 && abs(y - jp.ty0) <= 3) continue;     // already listed around tap 0
        if (sx >= 7 && abs(x - jp.tx0) <= 3) continue;
        const double wy0 = gauss_reach(y, jp.ty0), wy1 = gauss_reach(y, jp.ty1);
        const double wx0 = gauss_reach(x, jp.tx0), wx1 = gauss_reach(x, jp.tx1);
        // separable order of cv2.GaussianBlur: rows first, then columns
        const double r0 = __dadd_rn(__dmul_rn(jp.tap[0], wx0), __dmul_rn(jp.tap[1], wx1));
        const double r1 = __dadd_rn(__dmul_rn(jp.tap[2], wx0), __dmul_rn(jp.tap[3], wx1));
        const double h = __dadd_rn(__dmul_rn(wy0, r0), __dmul_rn(wy1, r1));
        const T lab = label_s[(y - y_lo) * kLabel + x];
        const size_t o = (static_cast<size_t>(b) * a.J + j) * kMap + y * kLabel + x;
        a.heatmaps[o] = __double2float_rn(h);
        // datasets.py:372-374,380: (d_j - label) * [heat > 0] * mask / cube
        if (h > 0.0 && lab != T(0))
            a.dmap[o] = __double2float_rn(__ddiv_rn(__dsub_rn(jp.cd, static_cast<double>(lab)), g.cube));
    }
}

// ---------------------------------------------------------------------------
// prep: per-sample geometry + per-joint taps, once per sample
// ---------------------------------------------------------------------------
constexpr int kPrepThreads = 128;                 // 4 samples (warps) per CTA

__device__ void write_scalar_outputs(const SfrArgs& a, int b, const SampleGeom& g) {
    a.gate[b] = 0u;
    a.box_size[b] = static_cast<float>(g.nrows);     // datasets.py:319
    a.cube_size[b] = static_cast<float>(g.cube);
    a.com_out[3 * b + 0] = static_cast<float>(g.c0);
    a.com_out[3 * b + 1] = static_cast<float>(g.r0);
    a.com_out[3 * b + 2] = static_cast<float>(g.z);
}

// all joints of sample b with geometry g (one lane per joint); returns true if any joint raised
__device__ bool prep_joints(const SfrArgs& a, int b, int lane, const SampleGeom& g, const double* rot) {
    int bad = 0;
    for (int j = lane; j < a.J; j += 32) {
        float* un = a.uvd_norm + (static_cast<size_t>(b) * a.J + j) * 3;
        JointParam jp;
        if (g.ok) {
            joint_param(jp, un, a.uvd + (static_cast<size_t>(b) * a.J + j) * 3, g, rot);
        } else {
            jp.ok = 0; jp.cd = 0.0; jp.tx0 = jp.tx1 = jp.ty0 = jp.ty1 = 0; jp.pad[0] = jp.pad[1] = jp.pad[2] = 0;
            jp.tap[0] = jp.tap[1] = jp.tap[2] = jp.tap[3] = 0.0;
            un[0] = 0.f; un[1] = 0.f; un[2] = 0.f;
        }
        bad |= jp.ok ? 0 : 1;
        a.prep_joints[static_cast<size_t>(b) * a.J + j] = jp;
        if (a.taps_out != nullptr) {
            pwr_joint_taps t;
            t.tap[0] = jp.tap[0]; t.tap[1] = jp.tap[1]; t.tap[2] = jp.tap[2]; t.tap[3] = jp.tap[3];
            t.cd = jp.cd;
            t.cd_norm = g.ok ? __ddiv_rn(jp.cd, g.cube) : 0.0;
            t.tx0 = static_cast<int16_t>(jp.tx0); t.tx1 = static_cast<int16_t>(jp.tx1);
            t.ty0 = static_cast<int16_t>(jp.ty0); t.ty1 = static_cast<int16_t>(jp.ty1);
            t.ok = jp.ok; t.pad = 0;
            a.taps_out[static_cast<size_t>(b) * a.J + j] = t;
        }
    }
    return __any_sync(0xffffffffu, bad) != 0;
}

template <bool TRAIN, bool AUG>
__global__ void __launch_bounds__(kPrepThreads)
sfr_prep_kernel(SfrArgs a) {
    const int b = blockIdx.x * (kPrepThreads / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= a.B) return;
    SampleGeom* gp = a.prep_geom + b;
    if (AUG) {
        // augmented branch first (datasets.py:216-299); whatever raises in it falls back to the
        // plain branch (:301), exactly like the reference's try / except
        const double* au = a.aug + 8 * static_cast<size_t>(b);
        if (lane == 0) {
            SampleGeom g;
            const double com2[3] = {__dadd_rn(a.com[3 * b + 0], au[1]), __dadd_rn(a.com[3 * b + 1], au[2]), a.com[3 * b + 2]};
            sample_geometry(g, com2, a.cube[b], a.fx, a.fy, a.Hf, a.Wf, a.pf_margin, a.pf_umax, a.pf_vmax);
            if (a.win != nullptr) clip_to_window(g, a.win[b]);
            *gp = g;
        }
        __syncwarp();
        const double rot[3] = {au[5], au[6], au[0]};
        bool raised = !gp->ok;
        if (!raised) raised = prep_joints(a, b, lane, *gp, rot);
        raised = __any_sync(0xffffffffu, raised) != 0;
        if (lane == 0) {
            WarpParam wp;
            wp.active = raised ? 0 : 1;
            wp.scale = au[0];
            // cv2.getRotationMatrix2D((64, 64), angle, scale) and cv::invertAffineTransform
            const double alpha = __dmul_rn(au[3], au[0]), beta = __dmul_rn(au[4], au[0]), cx = 64.0, cy = 64.0;
            const double m02 = __dsub_rn(__dmul_rn(__dsub_rn(1.0, alpha), cx), __dmul_rn(beta, cy));
            const double m12 = __dadd_rn(__dmul_rn(beta, cx), __dmul_rn(__dsub_rn(1.0, alpha), cy));
            double D = __dsub_rn(__dmul_rn(alpha, alpha), __dmul_rn(beta, -beta));
            D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
            const double A11 = __dmul_rn(alpha, D), A22 = __dmul_rn(alpha, D);
            const double A12 = __dmul_rn(-beta, D), A21 = __dmul_rn(beta, D);       // -M01*D, -M10*D
            wp.iM[0] = A11; wp.iM[1] = A12;
            wp.iM[2] = __dsub_rn(__dmul_rn(-A11, m02), __dmul_rn(A12, m12));
            wp.iM[3] = A21; wp.iM[4] = A22;
            wp.iM[5] = __dsub_rn(__dmul_rn(-A21, m02), __dmul_rn(A22, m12));
            wp.pad = 0;
            a.prep_warp[b] = wp;
        }
        if (!raised) {
            if (lane == 0) { write_scalar_outputs(a, b, *gp); a.prep_flags[b] = 0; }
            return;
        }
        __syncwarp();
    }
    if (lane == 0) {
        SampleGeom g;
        sample_geometry(g, a.com + 3 * b, a.cube[b], a.fx, a.fy, a.Hf, a.Wf, a.pf_margin, a.pf_umax, a.pf_vmax);
        if (a.win != nullptr) clip_to_window(g, a.win[b]);
        *gp = g;
        write_scalar_outputs(a, b, g);
    }
    __syncwarp();
    if (TRAIN) {
        const bool bad = prep_joints(a, b, lane, *gp, nullptr);
        if (lane == 0) a.prep_flags[b] = bad ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
// STAGED (dispatch option PWR_OPT_SFR_STAGED, off by default): the source rows of the band (the crop box rows its taps
// touch, clipped to the non-zero rectangle, columns rounded outwards to 16 bytes) are brought into shared memory with
// one 1-D bulk-TMA copy per row, all in flight at once and independent of registers / occupancy; the resample then
// gathers from shared memory.  A band whose rows do not fit kStageBytes (large crop boxes) keeps the direct gather.
// Bit-identical to the gather (tests/test_gpu_sfr.py) and measured SLOWER on B200 at B = 4096 (r2, ms per build,
// gather -> staged): NYU float32 dense 0.546 -> 0.669, compact 0.244 -> 0.421; raw uint16 dense 0.511 -> 0.574,
// compact 0.243 -> 0.355; HAND17 test-only 0.222 -> 0.362 (float32), 0.220 -> 0.304 (uint16).  Why: the gather kernels
// are not waiting for bytes - the dense one moves its real DRAM traffic at 92 % of the copy peak, the compact ones
// issue 2.3 instructions per clock per SM - so staging only adds what it costs: 36 KB per CTA (4 CTAs per SM instead
// of 5-6), a full-CTA wait on the row copies, and bands short enough to fit the stage (8 / 16 per sample), each of
// which repeats the per-sample prologue (x-tap table in float64, window table).
constexpr int kStageBytes = 36 * 1024;

template <typename T, int FMT, bool TRAIN, int kBands, bool STAGED = false>
__global__ void __launch_bounds__(kThreads)
sfr_build_kernel(SfrArgs a) {
    constexpr int kBandRows = kLabel / kBands;                    // label rows per CTA
    constexpr int kLabelIters = kBandRows * kLabel / kThreads;
    static_assert(kLabel % kBands == 0 && kImage + 2 * kBandRows <= kThreads, "tap tables are built by one thread each");
    __shared__ SampleGeom geom;
    __shared__ JointParam joints[TRAIN ? PWR_MAX_JOINTS : 1];
    __shared__ TapX xtap[kImage];
    __shared__ TapX ytap[2 * kBandRows];
    __shared__ T label_s[kBandRows * kLabel];
    __shared__ int band_list[TRAIN ? PWR_MAX_JOINTS : 1];   // joints whose footprint touches this band
    __shared__ int band_list_n;
    __shared__ int band_flags[2];                           // [0] mask count, [1] NaN seen (this CTA)
    constexpr bool kTab = (FMT != FMT_F32) && sizeof(T) == 4;      // raw 16-bit frames: window lookup table
    __shared__ float wtab[kTab ? kWinTab + 1 : 1];          // + the 0.f sentinel at [n]
    extern __shared__ __align__(128) unsigned char stage_raw[];    // STAGED: kStageBytes of source rows
    __shared__ __align__(8) uint64_t stage_bar;

    const int band = blockIdx.x % kBands;
    const int b = blockIdx.x / kBands;
    const int tid = threadIdx.x;
    const int y_lo = band * kBandRows, y_hi = y_lo + kBandRows;
    static_assert(sizeof(SampleGeom) % 4 == 0 && sizeof(JointParam) % 4 == 0, "word copies");

    // ---- prologue: fetch the prepared geometry / joint taps (L2-resident, written by the prep
    // kernel) into shared memory with coalesced word loads ...
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.prep_geom + b);
        if (tid < static_cast<int>(sizeof(SampleGeom) / 4)) reinterpret_cast<uint32_t*>(&geom)[tid] = __ldcg(src + tid);
        if (TRAIN) {
            const uint32_t* js = reinterpret_cast<const uint32_t*>(a.prep_joints + static_cast<size_t>(b) * a.J);
            const int words = a.J * static_cast<int>(sizeof(JointParam) / 4);
            for (int i = tid; i < words; i += kThreads) reinterpret_cast<uint32_t*>(joints)[i] = __ldcg(js + i);
        }
        if (tid == 0) {
            band_list_n = 0; band_flags[0] = 0; band_flags[1] = 0;
            if (STAGED) {
                sfr_mbar_init(&stage_bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
    }
    // ... while pass A of phase 2 streams the zeros: a heat map is zero outside the <= 8x8
    // footprint of the blurred 4-tap splat, and so is the depth map, so every map band is
    // zero-filled with 128-bit stores here and patched after phase 1.
    const bool dense = TRAIN && a.heatmaps != nullptr;      // dense maps wanted (else: compact taps only)
    if (dense) {
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t band_off = static_cast<size_t>(b) * a.J * kMap + y_lo * kLabel + (tid & 31) * 4;
        for (int j = tid >> 5; j < a.J; j += kWarps) {
            float* hp = a.heatmaps + band_off + static_cast<size_t>(j) * kMap;
            float* dp = a.dmap + band_off + static_cast<size_t>(j) * kMap;
#pragma unroll
            for (int i = 0; i < kBandRows * kLabel / 128; ++i) {
                st_stream(hp + i * 128, zero4);
                st_stream(dp + i * 128, zero4);
            }
        }
    }
    __syncthreads();
    const SampleGeom g = geom;
    constexpr int kElem = (FMT == FMT_F32) ? 4 : 2;
    const unsigned char* frame_bytes = static_cast<const unsigned char*>(a.frames) +
                                       static_cast<size_t>(b) * a.frame_stride * kElem;
    // ---- STAGED: plan + issue the row copies first, so they fly while the tap tables are built
    bool staged = false;
    int st_rlo = 0, st_c0 = 0, st_pitch = 0;       // frame row of staged row 0; buffer column of staged column 0; elements per staged row
    if (STAGED && g.ok) {
        constexpr int per16 = 16 / kElem;
        const TapX t0 = linear_tap(band * 2 * kBandRows, g.nrows, g.scale_y);
        const TapX t1 = linear_tap(band * 2 * kBandRows + 2 * kBandRows - 1, g.nrows, g.scale_y);
        const int rlo = max(g.fr0 + t0.s0, g.pr0), rhi = min(g.fr0 + t1.s1 + 1, g.pr1);
        const int c_lo = max(g.fc0, g.pc0) - g.org_c, c_hi = min(g.fc0 + g.ncols, g.pc1) - g.org_c;     // buffer columns
        const int c_lo_al = c_lo / per16 * per16, c_hi_al = min((c_hi + per16 - 1) / per16 * per16, a.pitch);
        const int n_rows = rhi - rlo, seg = c_hi_al - c_lo_al;
        // staged rows start on 128-byte boundaries: no two bulk copies ever write the same shared-memory line
        const int stride = (seg * kElem + 127) / 128 * 128;
        if (n_rows > 0 && seg > 0 && c_lo >= 0 && n_rows * stride <= kStageBytes) {
            staged = true; st_rlo = rlo; st_c0 = c_lo_al; st_pitch = stride / kElem;
            if (tid < 32) {
                const uint32_t seg_bytes = static_cast<uint32_t>(seg) * kElem;
                if (tid == 0) sfr_mbar_expect_tx(&stage_bar, seg_bytes * static_cast<uint32_t>(n_rows));
                __syncwarp();
                for (int r = tid; r < n_rows; r += 32)
                    sfr_bulk_g2s(stage_raw + static_cast<size_t>(r) * stride,
                                 frame_bytes + (static_cast<size_t>(rlo + r - g.org_r) * a.pitch + c_lo_al) * kElem,
                                 seg_bytes, &stage_bar);
            }
        }
    }
    WinTab wt = {wtab, 0, 0};
    if (kTab && g.ok) {
        // raw values that can decode into (lo, hi): [floor(lo) - 2, ceil(hi) + 2] (the decode is within 0.02 of the
        // raw value); every other raw value windows to 0
        const float lo_f = floorf(g.lo_dn), hi_f = ceilf(g.hi_up);
        if (lo_f > -1.0e9f && hi_f < 1.0e9f) {
            const int b0 = max(static_cast<int>(lo_f) - 2, 0), n = static_cast<int>(hi_f) + 3 - b0;
            if (n > 0 && n <= kWinTab) { wt.base = b0; wt.n = n; }
        }
        for (int i = tid; i < wt.n; i += kThreads)
            wtab[i] = Arith<float>::window(decode_px<FMT>(static_cast<unsigned int>(wt.base + i)), g);
        if (tid == 0) wtab[wt.n] = 0.f;
    }
    if (g.ok) {
        if (tid < kImage) xtap[tid] = linear_tap(tid, g.ncols, g.scale_x);
        else if (tid < kImage + 2 * kBandRows)
            ytap[tid - kImage] = linear_tap(band * 2 * kBandRows + (tid - kImage), g.nrows, g.scale_y);
        if (dense && tid < a.J) {
            const JointParam& jp = joints[tid];                     // rows ty0-3..ty0+3 or ty1..ty1+3 hit the band?
            if (jp.ok && ((jp.ty0 + 3 >= y_lo && jp.ty0 - 3 < y_hi) || (jp.ty1 + 3 >= y_lo && jp.ty1 < y_hi)))
                band_list[atomicAdd(&band_list_n, 1)] = tid;
        }
    }
    __syncthreads();

    // ---- phase 1: image band, label band, mask band -------------------------
    float* img_b = a.img + static_cast<size_t>(b) * kImage * kImage;
    float* lab_b = a.label_img + static_cast<size_t>(b) * kMap;
    float* msk_b = a.mask + static_cast<size_t>(b) * kMap;
    const void* frame = frame_bytes;
    const T cube_t = static_cast<T>(g.cube);
    const T cube_r = Arith<T>::rcp(cube_t);
    const bool interior = g.ok && g.fc0 >= g.pc0 && g.fc0 + g.ncols <= g.pc1 && g.fr0 + ytap[0].s0 >= g.pr0 &&
                          g.fr0 + ytap[2 * kBandRows - 1].s1 < g.pr1;
    // staged rows: the same tap arithmetic on a "pixel buffer" that is the shared-memory stage (origin = first
    // staged row / column, pitch = staged row length); the bounds predicates are the frame's, unchanged
    SampleGeom gs = g;
    if (STAGED && staged) {
        gs.org_r = st_rlo; gs.org_c = g.org_c + st_c0;
        sfr_mbar_wait(&stage_bar, 0);
    }
    int my_count = 0, my_nan = 0;
#pragma unroll 2
    for (int it = 0; it < kLabelIters; ++it) {
        const int lrow = it * (kThreads / kLabel) + (tid >> 6);     // label row inside the band
        const int lx = tid & (kLabel - 1);
        const int gly = y_lo + lrow;                               // label row in the sample
        T lab = T(0);
        float2 o0 = make_float2(0.f, 0.f), o1 = o0;
        if (g.ok) {
            T px[2][2];
            if (STAGED && staged) {
                if (kTab && wt.n > 0) {
                    if (interior) resample_block<T, FMT, true, kTab, true>(px, stage_raw, gs, &ytap[2 * lrow], &xtap[2 * lx], st_pitch, wt);
                    else          resample_block<T, FMT, false, kTab, true>(px, stage_raw, gs, &ytap[2 * lrow], &xtap[2 * lx], st_pitch, wt);
                } else {
                    if (interior) resample_block<T, FMT, true, false, true>(px, stage_raw, gs, &ytap[2 * lrow], &xtap[2 * lx], st_pitch);
                    else          resample_block<T, FMT, false, false, true>(px, stage_raw, gs, &ytap[2 * lrow], &xtap[2 * lx], st_pitch);
                }
            } else if (kTab && wt.n > 0) {
                if (interior) resample_block<T, FMT, true, kTab>(px, frame, g, &ytap[2 * lrow], &xtap[2 * lx], a.pitch, wt);
                else          resample_block<T, FMT, false, kTab>(px, frame, g, &ytap[2 * lrow], &xtap[2 * lx], a.pitch, wt);
            } else {
                if (interior) resample_block<T, FMT, true>(px, frame, g, &ytap[2 * lrow], &xtap[2 * lx], a.pitch);
                else          resample_block<T, FMT, false>(px, frame, g, &ytap[2 * lrow], &xtap[2 * lx], a.pitch);
            }
            // 2x2 mean (cv::resize reroutes an exact 2x INTER_LINEAR shrink to the area path)
            lab = Arith<T>::mul(Arith<T>::add(Arith<T>::add(px[0][0], px[0][1]), Arith<T>::add(px[1][0], px[1][1])),
                                T(0.25));
            o0 = make_float2(static_cast<float>(Arith<T>::div_by(px[0][0], cube_t, cube_r)),
                             static_cast<float>(Arith<T>::div_by(px[0][1], cube_t, cube_r)));
            o1 = make_float2(static_cast<float>(Arith<T>::div_by(px[1][0], cube_t, cube_r)),
                             static_cast<float>(Arith<T>::div_by(px[1][1], cube_t, cube_r)));
        }
        const float labn = static_cast<float>(Arith<T>::div_by(lab, cube_t, cube_r));
        const bool hand = g.ok && (lab != T(0));
        *reinterpret_cast<float2*>(img_b + (2 * gly) * kImage + 2 * lx) = o0;
        *reinterpret_cast<float2*>(img_b + (2 * gly + 1) * kImage + 2 * lx) = o1;
        lab_b[gly * kLabel + lx] = g.ok ? labn : 0.f;
        msk_b[gly * kLabel + lx] = hand ? 1.f : 0.f;
        label_s[lrow * kLabel + lx] = lab;
        my_count += hand ? 1 : 0;
        my_nan |= (isnan(o0.x) || isnan(o0.y) || isnan(o1.x) || isnan(o1.y) || isnan(labn)) ? 1 : 0;
    }
    {
        // block totals of my_count / my_nan
        int c = my_count, n = my_nan;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); n |= __shfl_xor_sync(0xffffffffu, n, o); }
        if ((tid & 31) == 0) { atomicAdd(&band_flags[0], c); atomicOr(&band_flags[1], n); }
    }
    __syncthreads();                                      // label_s and band_flags complete

    // ---- per-sample reject gate: one packed atomicAdd per band; the band that arrives last
    // sees the totals of all kBands bands and writes `valid` (datasets.py:362-365, 385-390)
    if (tid == 0) {
        const unsigned int mine = (1u << 24) | (band_flags[1] ? (1u << 16) : 0u) | static_cast<unsigned int>(band_flags[0]);
        const unsigned int total = atomicAdd(a.gate + b, mine) + mine;
        if ((total >> 24) == static_cast<unsigned int>(kBands)) {
            const int count = static_cast<int>(total & 0xffffu);        // <= 4096
            const bool nan_seen = ((total >> 16) & 0xffu) != 0;
            uint8_t v = g.ok ? 1 : 0;
            if (TRAIN) v = (g.ok && !a.prep_flags[b] && !nan_seen && count >= 10) ? 1 : 0;
            a.valid[b] = v;
        }
    }

    // ---- phase 2, pass B: the <= 8x8 footprint of every joint that touches this band
    // (pass A zero-filled the maps before the barriers above, so CTA-scope order holds)
    if (dense && g.ok)
        patch_footprints<T>(a, g, joints, band_list, band_list_n, label_s, b, y_lo, y_hi, tid, kThreads);
}

// ---------------------------------------------------------------------------
// augmented samples: one CTA per sample (the rotation mixes all rows of the image)
// ---------------------------------------------------------------------------
constexpr int kAugThreads = 512;

template <typename T>
struct AugSmem {
    T img[kImage * kImage];          // un-normalised resized crop before the warp (datasets.py:271)
    T label[kMap];                   // un-normalised label of the augmented image
    TapX xtap[kImage], ytap[kImage];
    int adelta[kImage], bdelta[kImage], x0[kImage], y0[kImage];   // cv::warpAffine fixed-point tables
    SampleGeom geom;
    WarpParam warp;
    JointParam joints[PWR_MAX_JOINTS];
    int list[PWR_MAX_JOINTS];
    int list_n;
    int count, nan_seen;
};

// One pixel of cv2.warpAffine(img, M, (128,128)) * scale (utils.py:75, datasets.py:284): 10-bit
// fixed-point inverse map, source position quantised to 1/32 px, float table weights, BORDER_CONSTANT 0.
template <typename T>
__device__ __forceinline__ T warp_pixel(const AugSmem<T>& sm, int x, int y, T scale_t) {
    const int X = (sm.x0[y] + sm.adelta[x]) >> 5, Y = (sm.y0[y] + sm.bdelta[x]) >> 5;
    const int sx = X >> 5, sy = Y >> 5;
    const float fx1 = static_cast<float>(X & 31) * 0.03125f, fy1 = static_cast<float>(Y & 31) * 0.03125f;
    const float fx0 = __fsub_rn(1.f, fx1), fy0 = __fsub_rn(1.f, fy1);
    const bool r0 = sy >= 0 && sy < kImage, r1 = sy + 1 >= 0 && sy + 1 < kImage;
    const bool c0 = sx >= 0 && sx < kImage, c1 = sx + 1 >= 0 && sx + 1 < kImage;
    const T s00 = (r0 && c0) ? sm.img[sy * kImage + sx] : T(0);
    const T s01 = (r0 && c1) ? sm.img[sy * kImage + sx + 1] : T(0);
    const T s10 = (r1 && c0) ? sm.img[(sy + 1) * kImage + sx] : T(0);
    const T s11 = (r1 && c1) ? sm.img[(sy + 1) * kImage + sx + 1] : T(0);
    T v = Arith<T>::mul(s00, static_cast<T>(__fmul_rn(fy0, fx0)));
    v = Arith<T>::add(v, Arith<T>::mul(s01, static_cast<T>(__fmul_rn(fy0, fx1))));
    v = Arith<T>::add(v, Arith<T>::mul(s10, static_cast<T>(__fmul_rn(fy1, fx0))));
    v = Arith<T>::add(v, Arith<T>::mul(s11, static_cast<T>(__fmul_rn(fy1, fx1))));
    return Arith<T>::mul(v, scale_t);
}

template <typename T, int FMT>
__global__ void __launch_bounds__(kAugThreads)
sfr_aug_kernel(SfrArgs a) {
    extern __shared__ __align__(16) unsigned char aug_smem_raw[];
    AugSmem<T>& sm = *reinterpret_cast<AugSmem<T>*>(aug_smem_raw);
    const int b = blockIdx.x;
    const int tid = threadIdx.x;

    {   // prepared geometry / warp / joint taps -> shared memory
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.prep_geom + b);
        if (tid < static_cast<int>(sizeof(SampleGeom) / 4)) reinterpret_cast<uint32_t*>(&sm.geom)[tid] = __ldcg(src + tid);
        const uint32_t* ws = reinterpret_cast<const uint32_t*>(a.prep_warp + b);
        if (tid >= 64 && tid < 64 + static_cast<int>(sizeof(WarpParam) / 4))
            reinterpret_cast<uint32_t*>(&sm.warp)[tid - 64] = __ldcg(ws + tid - 64);
        const uint32_t* js = reinterpret_cast<const uint32_t*>(a.prep_joints + static_cast<size_t>(b) * a.J);
        const int words = a.J * static_cast<int>(sizeof(JointParam) / 4);
        for (int i = tid; i < words; i += kAugThreads) reinterpret_cast<uint32_t*>(sm.joints)[i] = __ldcg(js + i);
        if (tid == 0) { sm.list_n = 0; sm.count = 0; sm.nan_seen = 0; }
    }
    const bool dense = a.heatmaps != nullptr;
    if (dense) {   // pass A of phase 2: zero every map of the sample
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float* hp = a.heatmaps + static_cast<size_t>(b) * a.J * kMap;
        float* dp = a.dmap + static_cast<size_t>(b) * a.J * kMap;
        for (int i = tid; i < a.J * (kMap / 4); i += kAugThreads) {
            st_stream(hp + i * 4, zero4);
            st_stream(dp + i * 4, zero4);
        }
    }
    __syncthreads();
    const SampleGeom g = sm.geom;
    const bool warp_on = sm.warp.active != 0;
    if (g.ok) {
        if (tid < kImage) {
            sm.xtap[tid] = linear_tap(tid, g.ncols, g.scale_x);
            // cv::warpAffine: adelta[x] = saturate_cast<int>(M[0]*x*1024), X0(y) = saturate_cast<int>((M[1]*y + M[2])*1024) + 16
            const double* m = sm.warp.iM;
            const double t = static_cast<double>(tid);
            sm.adelta[tid] = __double2int_rn(__dmul_rn(__dmul_rn(m[0], t), 1024.0));
            sm.bdelta[tid] = __double2int_rn(__dmul_rn(__dmul_rn(m[3], t), 1024.0));
            sm.x0[tid] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], t), m[2]), 1024.0)) + 16;
            sm.y0[tid] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], t), m[5]), 1024.0)) + 16;
        } else if (tid < 2 * kImage) {
            sm.ytap[tid - kImage] = linear_tap(tid - kImage, g.nrows, g.scale_y);
        } else if (tid < 2 * kImage + a.J) {
            const int j = tid - 2 * kImage;
            if (sm.joints[j].ok) sm.list[atomicAdd(&sm.list_n, 1)] = j;
        }
    }
    __syncthreads();

    // ---- phase 1a: the resized crop (datasets.py:271), one pixel per thread and iteration
    const void* frame = static_cast<const unsigned char*>(a.frames) +
                        static_cast<size_t>(b) * a.frame_stride * (FMT == FMT_F32 ? 4 : 2);
    if (g.ok) {
        for (int p = tid; p < kImage * kImage; p += kAugThreads) {
            const int y = p >> 7, x = p & (kImage - 1);
            const TapX ty = sm.ytap[y], tx = sm.xtap[x];
            const int fr_a = g.fr0 + ty.s0, fr_b = g.fr0 + ty.s1, fc_a = g.fc0 + tx.s0, fc_b = g.fc0 + tx.s1;
            const bool ra = fr_a >= g.pr0 && fr_a < g.pr1, rb = fr_b >= g.pr0 && fr_b < g.pr1;
            const bool ca = fc_a >= g.pc0 && fc_a < g.pc1, cb = fc_b >= g.pc0 && fc_b < g.pc1;
            const int row_a = (fr_a - g.org_r) * a.pitch - g.org_c, row_b = (fr_b - g.org_r) * a.pitch - g.org_c;
            const float v00 = (ra && ca) ? load_px<FMT>(frame, row_a + fc_a) : 0.f;
            const float v01 = (ra && cb) ? load_px<FMT>(frame, row_a + fc_b) : 0.f;
            const float v10 = (rb && ca) ? load_px<FMT>(frame, row_b + fc_a) : 0.f;
            const float v11 = (rb && cb) ? load_px<FMT>(frame, row_b + fc_b) : 0.f;
            const T w00 = Arith<T>::window(static_cast<T>(v00), g), w01 = Arith<T>::window(static_cast<T>(v01), g);
            const T w10 = Arith<T>::window(static_cast<T>(v10), g), w11 = Arith<T>::window(static_cast<T>(v11), g);
            const T xa0 = static_cast<T>(tx.a0), xa1 = static_cast<T>(tx.a1);
            const T top = Arith<T>::add(Arith<T>::mul(w00, xa0), Arith<T>::mul(w01, xa1));
            const T bot = Arith<T>::add(Arith<T>::mul(w10, xa0), Arith<T>::mul(w11, xa1));
            sm.img[p] = Arith<T>::add(Arith<T>::mul(top, static_cast<T>(ty.a0)), Arith<T>::mul(bot, static_cast<T>(ty.a1)));
        }
    }
    __syncthreads();

    // ---- phase 1b: rotate + scale (utils.py:74-75, datasets.py:284), 2x2 mean, outputs
    float* img_b = a.img + static_cast<size_t>(b) * kImage * kImage;
    float* lab_b = a.label_img + static_cast<size_t>(b) * kMap;
    float* msk_b = a.mask + static_cast<size_t>(b) * kMap;
    const T cube_t = static_cast<T>(g.cube);
    const T cube_r = Arith<T>::rcp(cube_t);
    const T scale_t = static_cast<T>(sm.warp.scale);
    int my_count = 0, my_nan = 0;
    for (int p = tid; p < kMap; p += kAugThreads) {
        const int ly = p >> 6, lx = p & (kLabel - 1);
        T lab = T(0);
        float2 o0 = make_float2(0.f, 0.f), o1 = o0;
        if (g.ok) {
            T px[2][2];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx)
                    px[dy][dx] = warp_on ? warp_pixel<T>(sm, 2 * lx + dx, 2 * ly + dy, scale_t)
                                         : sm.img[(2 * ly + dy) * kImage + 2 * lx + dx];
            lab = Arith<T>::mul(Arith<T>::add(Arith<T>::add(px[0][0], px[0][1]), Arith<T>::add(px[1][0], px[1][1])),
                                T(0.25));
            o0 = make_float2(static_cast<float>(Arith<T>::div_by(px[0][0], cube_t, cube_r)),
                             static_cast<float>(Arith<T>::div_by(px[0][1], cube_t, cube_r)));
            o1 = make_float2(static_cast<float>(Arith<T>::div_by(px[1][0], cube_t, cube_r)),
                             static_cast<float>(Arith<T>::div_by(px[1][1], cube_t, cube_r)));
        }
        const float labn = static_cast<float>(Arith<T>::div_by(lab, cube_t, cube_r));
        const bool hand = g.ok && (lab != T(0));
        *reinterpret_cast<float2*>(img_b + (2 * ly) * kImage + 2 * lx) = o0;
        *reinterpret_cast<float2*>(img_b + (2 * ly + 1) * kImage + 2 * lx) = o1;
        lab_b[p] = g.ok ? labn : 0.f;
        msk_b[p] = hand ? 1.f : 0.f;
        sm.label[p] = lab;
        my_count += hand ? 1 : 0;
        my_nan |= (isnan(o0.x) || isnan(o0.y) || isnan(o1.x) || isnan(o1.y) || isnan(labn)) ? 1 : 0;
    }
    {
        int c = my_count, n = my_nan;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); n |= __shfl_xor_sync(0xffffffffu, n, o); }
        if ((tid & 31) == 0) { atomicAdd(&sm.count, c); atomicOr(&sm.nan_seen, n); }
    }
    __syncthreads();
    if (tid == 0)   // reject gate, datasets.py:362-365, 385-390 (one CTA owns the whole sample)
        a.valid[b] = (g.ok && !a.prep_flags[b] && !sm.nan_seen && sm.count >= 10) ? 1 : 0;

    // ---- phase 2, pass B
    if (dense && g.ok) patch_footprints<T>(a, g, sm.joints, sm.list, sm.list_n, sm.label, b, 0, kLabel, tid, kAugThreads);
}

// ---------------------------------------------------------------------------
// host -> device feed: fetch only what the builder will read (pwr_sfr_fetch)
// ---------------------------------------------------------------------------
// train.py:161-166 ships every tensor of the batch with .to(device); for the depth frame that is the whole
// 480x640 image although the builder reads only the crop box (176-352 px at NYU depths) inside the hand
// rectangle.  Here the frames stay in pinned (device-mapped) host memory and this kernel pulls, per sample,
// exactly that region over PCIe with coalesced 16-byte reads - no host-side repacking, no per-sample memcpy
// call - into a compact [B, win_h, win_w] buffer the builder then reads in "window mode".  A plan kernel
// re-derives the crop geometry with the builder's own code (so the two agree bit for bit); a small persistent
// copy kernel moves the bytes.  The same kernels work on device-resident frames (a plain gather in HBM).
#ifndef PWR_FETCH_GRID
#define PWR_FETCH_GRID 32
#endif
constexpr int kFetchGrid = PWR_FETCH_GRID;      // CTAs of the copy kernel in total (NOT per SM), see below
constexpr int kFetchThreads = 256;
constexpr int kFetchGroups = 4;              // CTAs per sample
constexpr int kFetchUnroll = 4;              // 16-byte loads in flight per thread

struct FetchArgs {
    const void* frames; int elem; int Hf, Wf;
    const double* com; const double* cube; const double* aug;
    double fx, fy, pf_margin, pf_umax, pf_vmax;
    void* windows; int win_h, win_w;
    WinExtent* extent; unsigned long long* fetched_bytes; int* status;
    int B;
};

// rows / cols of the frame the builder can touch for this geometry: box AND non-zero rectangle
__device__ __forceinline__ void needed_region(const SampleGeom& g, int& r0, int& r1, int& c0, int& c1) {
    if (!g.ok) { r0 = r1 = c0 = c1 = 0; return; }
    r0 = max(g.fr0, g.pr0); r1 = min(g.fr0 + g.nrows, g.pr1);
    c0 = max(g.fc0, g.pc0); c1 = min(g.fc0 + g.ncols, g.pc1);
    if (r1 <= r0 || c1 <= c0) r0 = r1 = c0 = c1 = 0;
}

// plan: one thread per sample derives the region with the builder's own geometry code
__global__ void __launch_bounds__(128)
sfr_fetch_plan_kernel(FetchArgs a) {
    const int b = blockIdx.x * 128 + threadIdx.x;
    if (b >= a.B) return;
    const int per16 = 16 / a.elem;                       // elements per 16-byte chunk
    SampleGeom g;
    int r0, r1, c0, c1;
    sample_geometry(g, a.com + 3 * b, a.cube[b], a.fx, a.fy, a.Hf, a.Wf, a.pf_margin, a.pf_umax, a.pf_vmax);
    needed_region(g, r0, r1, c0, c1);
    if (a.aug != nullptr) {
        // the augmented branch crops around the shifted centre (datasets.py:236-246) and falls back to
        // the plain branch when it raises: fetch the union of both regions
        const double* au = a.aug + 8 * static_cast<size_t>(b);
        const double com2[3] = {__dadd_rn(a.com[3 * b + 0], au[1]), __dadd_rn(a.com[3 * b + 1], au[2]), a.com[3 * b + 2]};
        int q0, q1, p0, p1;
        sample_geometry(g, com2, a.cube[b], a.fx, a.fy, a.Hf, a.Wf, a.pf_margin, a.pf_umax, a.pf_vmax);
        needed_region(g, q0, q1, p0, p1);
        if (q1 > q0) {
            if (r1 > r0) { r0 = min(r0, q0); r1 = max(r1, q1); c0 = min(c0, p0); c1 = max(c1, p1); }
            else { r0 = q0; r1 = q1; c0 = p0; c1 = p1; }
        }
    }
    // outwards to 16 bytes (Wf * elem % 16 == 0).  Wider alignment was measured and lost: 64 / 128 bytes raise the
    // PCIe rate from 44 to 49 / 51 GB/s but move 13 / 30 % more bytes (tools/ab_fetch.py, r2).
    c0 = c0 / per16 * per16;
    c1 = min((c1 + per16 - 1) / per16 * per16, a.Wf);
    int rows = r1 - r0, cols = c1 - c0;
    if (rows > a.win_h || cols > a.win_w) {               // the caller sized the windows too small
        if (a.status != nullptr) atomicOr(a.status, 1);
        rows = min(rows, a.win_h); cols = min(cols, a.win_w / per16 * per16);
    }
    WinExtent ext = {r0, c0, rows, cols};
    a.extent[b] = ext;
    if (a.fetched_bytes != nullptr) atomicAdd(a.fetched_bytes, static_cast<unsigned long long>(rows) * cols * a.elem);
}

// copy: a SMALL persistent grid (kFetchGrid = 32 CTAs in total) walks the (sample, row group) items.  The kernel is
// PCIe-bound - 32 x 256 threads x 4 x 16 B = 0.5 MB in flight is more than the link needs - and it runs next to the
// compute stream's kernels for most of a step.  Its loads are system-memory reads with microseconds of latency: on
// every SM that hosts a fetch CTA they fill the load/store unit's miss queues, and a concurrent kernel that gathers
// with ordinary loads (the SFR build) crawls on that SM.  Measured (r2, B = 4096 NYU raw frames, tools/ab_e2e.py): one
// CTA per item (16 384 CTAs) or even ONE CTA on each of the 148 SMs delayed the concurrent SFR build from 0.51 to
// 6.1 ms (the whole transfer) -> 7.6 ms per step; 32 CTAs confine the damage to 32 SMs -> SFR build 0.63 ms, step
// 6.07 ms = the transfer itself (grid 8 / 16 / 32 / 64: 7.27 / 6.17 / 6.07 / 6.08 ms per step).  The bulk-TMA
// decoder kernels never were affected (their loads do not go through the LSU).  A bulk-TMA form of this copy (rows
// host -> shared -> HBM with cp.async.bulk both ways, one row per lane in flight; tools/pcie_probe.cu keeps it) was
// built and measured: the same 40-45 GB/s - the PCIe rate follows the fragment size, not who issues the reads - and
// 6.27 ms per step; it was removed again rather than shipped as a slower option.
__global__ void __launch_bounds__(kFetchThreads)
sfr_fetch_copy_kernel(FetchArgs a) {
    const int per16 = 16 / a.elem;
    const long long items = static_cast<long long>(a.B) * kFetchGroups;
    for (long long w = blockIdx.x; w < items; w += gridDim.x) {
        const int b = static_cast<int>(w / kFetchGroups), grp = static_cast<int>(w % kFetchGroups);
        const WinExtent e = a.extent[b];
        const int cpr = e.cols / per16;                   // 16-byte chunks per row
        const int total = e.rows * cpr;
        const int per = (total + kFetchGroups - 1) / kFetchGroups;
        const int lo = grp * per, hi = min(total, lo + per);
        const unsigned char* src = static_cast<const unsigned char*>(a.frames) +
                                   (static_cast<size_t>(b) * a.Hf * a.Wf + static_cast<size_t>(e.row0) * a.Wf + e.col0) * a.elem;
        unsigned char* dst = static_cast<unsigned char*>(a.windows) + static_cast<size_t>(b) * a.win_h * a.win_w * a.elem;
        const size_t src_pitch = static_cast<size_t>(a.Wf) * a.elem, dst_pitch = static_cast<size_t>(a.win_w) * a.elem;
        for (int i = lo + threadIdx.x; i < hi; i += kFetchThreads * kFetchUnroll) {
            uint4 v[kFetchUnroll];
            int rr[kFetchUnroll], cc[kFetchUnroll];
#pragma unroll
            for (int u = 0; u < kFetchUnroll; ++u) {
                const int k = i + u * kFetchThreads;
                rr[u] = k / cpr; cc[u] = k - rr[u] * cpr;
                if (k < hi) v[u] = __ldcs(reinterpret_cast<const uint4*>(src + rr[u] * src_pitch + cc[u] * 16));
            }
#pragma unroll
            for (int u = 0; u < kFetchUnroll; ++u) {
                const int k = i + u * kFetchThreads;
                if (k < hi) *reinterpret_cast<uint4*>(dst + rr[u] * dst_pitch + cc[u] * 16) = v[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// centre-of-mass fallback, datasets.py:208-211
// ---------------------------------------------------------------------------
constexpr int kComThreads = 512;

__global__ void __launch_bounds__(kComThreads)
sfr_com_kernel(const float* __restrict__ frames, int Hf, int Wf, double* __restrict__ com) {
    __shared__ double s_z[kComThreads / 32];
    __shared__ long long s_r[kComThreads / 32], s_c[kComThreads / 32], s_n[kComThreads / 32];
    const int b = blockIdx.x;
    const float* frame = frames + static_cast<size_t>(b) * Hf * Wf;
    double sz = 0.0;
    long long sr = 0, sc = 0, sn = 0;
    const int n = Hf * Wf;
    for (int i = threadIdx.x; i < n; i += kComThreads) {
        const float v = __ldg(frame + i);
        if (v > 0.f) {
            sz += static_cast<double>(v);
            sr += i / Wf;
            sc += i % Wf;
            sn += 1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sc += __shfl_xor_sync(0xffffffffu, sc, o);
        sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_z[warp] = sz; s_r[warp] = sr; s_c[warp] = sc; s_n[warp] = sn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        sz = 0.0; sr = sc = sn = 0;
        for (int wv = 0; wv < kComThreads / 32; ++wv) { sz += s_z[wv]; sr += s_r[wv]; sc += s_c[wv]; sn += s_n[wv]; }
        const double cnt = static_cast<double>(sn);
        com[3 * b + 0] = __ddiv_rn(static_cast<double>(sc), cnt);
        com[3 * b + 1] = __ddiv_rn(static_cast<double>(sr), cnt);
        com[3 * b + 2] = __ddiv_rn(sz, cnt);
    }
}

// ---------------------------------------------------------------------------
// HAND17 bounding-box loader, datasets.py:974-996 (process_mode='bb', test frames)
// ---------------------------------------------------------------------------
// Keep the annotated box, then drop everything deeper than 100 mm behind the mean depth of what is left, in
// two passes (the reference's two np.mean over the frame).  Depths are integer sensor counts, so the sums are
// exact in 64-bit integers whatever the order, and the means are one correctly rounded float64 division
// each: bit-identical to NumPy.  One CTA per sample; the box (<= ~300 x 300 uint16) is re-read from L1/L2.
constexpr int kBbThreads = 512;

__device__ __forceinline__ void bb_block_sum(unsigned long long& sum, unsigned long long& cnt, unsigned long long* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                                   // scratch may still be read from the previous pass
    if (lane == 0) { scratch[2 * warp] = sum; scratch[2 * warp + 1] = cnt; }
    __syncthreads();
    sum = 0; cnt = 0;
    for (int wv = 0; wv < kBbThreads / 32; ++wv) { sum += scratch[2 * wv]; cnt += scratch[2 * wv + 1]; }
}

__global__ void __launch_bounds__(kBbThreads)
sfr_bb_kernel(const unsigned short* __restrict__ raw, const double* __restrict__ boxes, float* __restrict__ out,
              int Hf, int Wf) {
    __shared__ unsigned long long scratch[2 * kBbThreads / 32];
    const int b = blockIdx.x;
    const unsigned short* frame = raw + static_cast<size_t>(b) * Hf * Wf;
    float* dst = out + static_cast<size_t>(b) * Hf * Wf;
    const double ustart = boxes[4 * b + 0], vstart = boxes[4 * b + 1], du = boxes[4 * b + 2], dv = boxes[4 * b + 3];
    // MM[int(vstart):int(vstart+dv), int(ustart):int(ustart+du)] = 1, Python slice semantics
    int r0, nr, c0, nc;
    py_slice(py_int(vstart), py_int(__dadd_rn(vstart, dv)), Hf, r0, nr);
    py_slice(py_int(ustart), py_int(__dadd_rn(ustart, du)), Wf, c0, nc);
    const int npx = nr * nc;
    // pass 1: mean of the positive pixels inside the box
    unsigned long long sum = 0, cnt = 0;
    for (int i = threadIdx.x; i < npx; i += kBbThreads) {
        const unsigned int v = __ldg(frame + (r0 + i / nc) * Wf + c0 + i % nc);
        if (v > 0u) { sum += v; cnt += 1; }
    }
    bb_block_sum(sum, cnt, scratch);
    const double cut1 = __dadd_rn(__ddiv_rn(static_cast<double>(sum), static_cast<double>(cnt)), 100.0);
    // pass 2: mean of what survives `> mean + 100 -> 0`
    sum = 0; cnt = 0;
    for (int i = threadIdx.x; i < npx; i += kBbThreads) {
        const unsigned int v = __ldg(frame + (r0 + i / nc) * Wf + c0 + i % nc);
        if (v > 0u && !(static_cast<double>(v) > cut1)) { sum += v; cnt += 1; }
    }
    bb_block_sum(sum, cnt, scratch);
    const double cut2 = __dadd_rn(__ddiv_rn(static_cast<double>(sum), static_cast<double>(cnt)), 100.0);
    // pass 3: the filtered frame (zero outside the box)
    const int n = Hf * Wf;
    for (int i = threadIdx.x; i < n; i += kBbThreads) {
        const int r = i / Wf, c = i - r * Wf;
        float o = 0.f;
        if (r >= r0 && r < r0 + nr && c >= c0 && c < c0 + nc) {
            const unsigned int v = __ldg(frame + i);
            if (!(static_cast<double>(v) > cut2)) o = static_cast<float>(v);
        }
        dst[i] = o;
    }
}

// workspace layout: [B] SampleGeom | [B*J] JointParam | [B] int joints_bad | [B] u32 gate | [B] WarpParam
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static size_t workspace_bytes(int B, int J) {
    return align_up(sizeof(SampleGeom) * static_cast<size_t>(B), 256) +
           align_up(sizeof(JointParam) * static_cast<size_t>(B) * (J > 0 ? J : 0), 256) +
           align_up(sizeof(int) * static_cast<size_t>(B), 256) + align_up(sizeof(unsigned int) * static_cast<size_t>(B), 256) +
           align_up(sizeof(WarpParam) * static_cast<size_t>(B), 256);
}
static void carve_workspace(SfrArgs& a, void* ws) {
    unsigned char* p = static_cast<unsigned char*>(ws);
    a.prep_geom = reinterpret_cast<SampleGeom*>(p);
    p += align_up(sizeof(SampleGeom) * static_cast<size_t>(a.B), 256);
    a.prep_joints = reinterpret_cast<JointParam*>(p);
    p += align_up(sizeof(JointParam) * static_cast<size_t>(a.B) * a.J, 256);
    a.prep_flags = reinterpret_cast<int*>(p);
    p += align_up(sizeof(int) * static_cast<size_t>(a.B), 256);
    a.gate = reinterpret_cast<unsigned int*>(p);
    p += align_up(sizeof(unsigned int) * static_cast<size_t>(a.B), 256);
    a.prep_warp = reinterpret_cast<WarpParam*>(p);
}

template <bool TRAIN>
static int launch_sfr(const SfrArgs& a, int frame_f64, int fmt, cudaStream_t stream) {
    if (fmt != FMT_F32 && fmt != FMT_GB16 && fmt != FMT_U16) return PWR_E_METHOD;
    if (frame_f64 && fmt != FMT_F32) return PWR_E_METHOD;       // float64 semantics exist for decoded frames only
    const unsigned prep_grid = (a.B + kPrepThreads / 32 - 1) / (kPrepThreads / 32);
    if (TRAIN && a.aug != nullptr) {
        const int dev = current_device();
        // augmented batch: one CTA per sample, the whole resized image staged in shared memory
        sfr_prep_kernel<true, true><<<prep_grid, kPrepThreads, 0, stream>>>(a);
        if (int rc = launch_status()) return rc;
#define PWR_LAUNCH_AUG(T, F)                                                                                  \
    do {                                                                                                      \
        PWR_ENSURE_DYN_SMEM(static_cast<int>(sizeof(AugSmem<T>)), dev, sfr_aug_kernel<T, F>);                 \
        sfr_aug_kernel<T, F><<<a.B, kAugThreads, sizeof(AugSmem<T>), stream>>>(a);                            \
    } while (0)
        if (frame_f64)            PWR_LAUNCH_AUG(double, FMT_F32);
        else if (fmt == FMT_F32)  PWR_LAUNCH_AUG(float, FMT_F32);
        else if (fmt == FMT_GB16) PWR_LAUNCH_AUG(float, FMT_GB16);
        else                      PWR_LAUNCH_AUG(float, FMT_U16);
#undef PWR_LAUNCH_AUG
        return launch_status();
    }
    sfr_prep_kernel<TRAIN, false><<<prep_grid, kPrepThreads, 0, stream>>>(a);
    if (int rc = launch_status()) return rc;
#define PWR_LAUNCH_SFR(NB)                                                                                         \
    do {                                                                                                           \
        const unsigned grid = static_cast<unsigned>(a.B) * NB;                                                     \
        if (frame_f64)            sfr_build_kernel<double, FMT_F32, TRAIN, NB><<<grid, kThreads, 0, stream>>>(a);  \
        else if (fmt == FMT_F32)  sfr_build_kernel<float, FMT_F32, TRAIN, NB><<<grid, kThreads, 0, stream>>>(a);   \
        else if (fmt == FMT_GB16) sfr_build_kernel<float, FMT_GB16, TRAIN, NB><<<grid, kThreads, 0, stream>>>(a);  \
        else                      sfr_build_kernel<float, FMT_U16, TRAIN, NB><<<grid, kThreads, 0, stream>>>(a);   \
    } while (0)
    // staged (bulk-TMA) variant: source rows of a band in kStageBytes of dynamic shared memory
#define PWR_LAUNCH_SFR_STAGED(T, F, NB)                                                                            \
    do {                                                                                                           \
        PWR_ENSURE_DYN_SMEM(kStageBytes, dev, sfr_build_kernel<T, F, TRAIN, NB, true>);                            \
        sfr_build_kernel<T, F, TRAIN, NB, true><<<static_cast<unsigned>(a.B) * NB, kThreads, kStageBytes, stream>>>(a); \
    } while (0)
    const int elem = fmt == FMT_F32 ? 4 : 2;
    const bool can_stage = get_option(PWR_OPT_SFR_STAGED) != 0 && (static_cast<long long>(a.pitch) * elem) % 16 == 0 &&
                           !misaligned(a.frames);
    const bool dense_maps = TRAIN && a.heatmaps != nullptr;
    if (can_stage) {
        const int dev = current_device();
        // 2-byte samples: 8 bands (<= 46 source rows of <= 368 px); 4-byte samples: 16 bands (half the rows)
        if (frame_f64)            PWR_LAUNCH_SFR_STAGED(double, FMT_F32, kBandsStagedF32);
        else if (fmt == FMT_F32)  PWR_LAUNCH_SFR_STAGED(float, FMT_F32, kBandsStagedF32);
        else if (fmt == FMT_GB16) PWR_LAUNCH_SFR_STAGED(float, FMT_GB16, kBandsStagedU16);
        else                      PWR_LAUNCH_SFR_STAGED(float, FMT_U16, kBandsStagedU16);
        return launch_status();
    }
    if (dense_maps) {
        PWR_LAUNCH_SFR(kBandsDense);
        return launch_status();
    }
    // Without dense maps, 2 long bands per sample win at large batches (no zero-fill to overlap the prologue);
    // a small batch (config 5 sweeps from 256) cannot fill 148 SMs x 6 CTAs with 2 CTAs per sample, so it takes
    // the 8-band grid: 4x the CTAs, each a quarter as long.
    if (static_cast<long long>(a.B) * kBandsLean < static_cast<long long>(sm_count(current_device())) * 12)
        PWR_LAUNCH_SFR(kBandsDense);
    else
        PWR_LAUNCH_SFR(kBandsLean);
#undef PWR_LAUNCH_SFR_STAGED
#undef PWR_LAUNCH_SFR
    return launch_status();
}

static int check_frames(int Hf, int Wf, int B) {
    if (B < 0 || Hf < 1 || Wf < 1 || Hf > 16384 || Wf > 16384) return PWR_E_SHAPE;
    if (static_cast<long long>(B) * kBandsMax > 0x7fffffffLL) return PWR_E_SHAPE;
    return 0;
}

// whole frames, or the windows of pwr_sfr_fetch
static int set_pixel_source(SfrArgs& a, const int* win_extent, int win_h, int win_w) {
    if (win_extent == nullptr) {
        a.win = nullptr; a.pitch = a.Wf; a.frame_stride = static_cast<long long>(a.Hf) * a.Wf;
        return 0;
    }
    if (win_h < 1 || win_w < 1 || win_h > 16384 || win_w > 16384) return PWR_E_SHAPE;
    if (misaligned(win_extent)) return PWR_E_ALIGN;
    a.win = reinterpret_cast<const WinExtent*>(win_extent);
    a.pitch = win_w; a.frame_stride = static_cast<long long>(win_h) * win_w;
    return 0;
}

}  // namespace pwr

using namespace pwr;

extern "C" size_t pwr_sfr_workspace_bytes(int B, int J) {
    if (B < 0 || J < 0 || J > PWR_MAX_JOINTS) return 0;
    return workspace_bytes(B, J);
}

extern "C" int pwr_sfr_com(const float* frames, int Hf, int Wf, double* com, int B, void* stream) {
    if (int rc = check_frames(Hf, Wf, B)) return rc;
    if (B == 0) return 0;
    if (frames == nullptr || com == nullptr) return PWR_E_NULL;
    sfr_com_kernel<<<B, kComThreads, 0, static_cast<cudaStream_t>(stream)>>>(frames, Hf, Wf, com);
    return launch_status();
}

extern "C" int pwr_sfr_crop(const void* frames, int frame_format, int Hf, int Wf, const double* com,
                            const double* cube, double fx, double fy, int frame_f64, double prefilter_margin,
                            double prefilter_umax, double prefilter_vmax, float* img, float* label_img, float* mask, float* box_size,
                            float* cube_size, float* com_out, uint8_t* valid, void* workspace, size_t workspace_size,
                            const int* win_extent, int win_h, int win_w, int B, void* stream) {
    if (int rc = check_frames(Hf, Wf, B)) return rc;
    if (B == 0) return 0;
    if (frames == nullptr || com == nullptr || cube == nullptr || box_size == nullptr || cube_size == nullptr ||
        com_out == nullptr || valid == nullptr)
        return PWR_E_NULL;
    PWR_REQUIRE_PTR(img); PWR_REQUIRE_PTR(label_img); PWR_REQUIRE_PTR(mask); PWR_REQUIRE_PTR(workspace);
    if (workspace_size < workspace_bytes(B, 0)) return PWR_E_SHAPE;
    SfrArgs a = {frames, Hf, Wf, com, cube, nullptr, fx, fy, img, label_img, mask, box_size, cube_size, com_out,
                 nullptr, nullptr, nullptr, valid, B, 0, nullptr, nullptr, nullptr, nullptr,
                 prefilter_margin, prefilter_umax, prefilter_vmax, nullptr, nullptr, nullptr, nullptr, 0, 0};
    if (int rc = set_pixel_source(a, win_extent, win_h, win_w)) return rc;
    carve_workspace(a, workspace);
    return launch_sfr<false>(a, frame_f64, frame_format, static_cast<cudaStream_t>(stream));
}

extern "C" int pwr_sfr_build(const void* frames, int frame_format, int Hf, int Wf, const double* com,
                             const double* cube, const double* uvd, const double* aug, double fx, double fy,
                             int frame_f64, double prefilter_margin, double prefilter_umax, double prefilter_vmax,
                             float* img,
                             float* label_img,
                             float* mask, float* box_size, float* cube_size, float* com_out, float* uvd_norm,
                             float* heatmaps, float* dmap, pwr_joint_taps* joint_taps, uint8_t* valid,
                             void* workspace, size_t workspace_size,
                             const int* win_extent, int win_h, int win_w,
                             int B, int J, void* stream) {
    if (int rc = check_frames(Hf, Wf, B)) return rc;
    if (J < 1 || J > PWR_MAX_JOINTS) return PWR_E_SHAPE;
    if (B == 0) return 0;
    if (frames == nullptr || com == nullptr || cube == nullptr || uvd == nullptr || box_size == nullptr ||
        cube_size == nullptr || com_out == nullptr || uvd_norm == nullptr || valid == nullptr)
        return PWR_E_NULL;
    PWR_REQUIRE_PTR(img); PWR_REQUIRE_PTR(label_img); PWR_REQUIRE_PTR(mask);
    PWR_OPTIONAL_PTR(heatmaps); PWR_OPTIONAL_PTR(dmap); PWR_OPTIONAL_PTR(joint_taps); PWR_REQUIRE_PTR(workspace);
    if ((heatmaps == nullptr) != (dmap == nullptr)) return PWR_E_NULL;          // both dense maps or neither
    if (heatmaps == nullptr && joint_taps == nullptr) return PWR_E_NULL;        // some form of targets is required
    if (workspace_size < workspace_bytes(B, J)) return PWR_E_SHAPE;
    SfrArgs a = {frames, Hf, Wf, com, cube, uvd, fx, fy, img, label_img, mask, box_size, cube_size, com_out,
                 uvd_norm, heatmaps, dmap, valid, B, J, nullptr, nullptr, nullptr, nullptr,
                 prefilter_margin, prefilter_umax, prefilter_vmax, joint_taps, aug, nullptr, nullptr, 0, 0};
    if (int rc = set_pixel_source(a, win_extent, win_h, win_w)) return rc;
    carve_workspace(a, workspace);
    return launch_sfr<true>(a, frame_f64, frame_format, static_cast<cudaStream_t>(stream));
}

extern "C" int pwr_sfr_fetch(const void* frames, int frame_format, int Hf, int Wf, const double* com,
                             const double* cube, const double* aug, double fx, double fy, double prefilter_margin,
                             double prefilter_umax, double prefilter_vmax, void* windows, int win_h, int win_w,
                             int* win_extent, unsigned long long* fetched_bytes, int* status, int B, void* stream) {
    if (int rc = check_frames(Hf, Wf, B)) return rc;
    if (frame_format != FMT_F32 && frame_format != FMT_GB16 && frame_format != FMT_U16) return PWR_E_METHOD;
    const int elem = frame_format == FMT_F32 ? 4 : 2;
    if ((Wf * elem) % 16 != 0 || win_h < 1 || win_w < 1 || (win_w * elem) % 16 != 0 || win_h > 16384 || win_w > 16384)
        return PWR_E_SHAPE;                                   // rows of both buffers must keep 16-byte alignment
    if (B == 0) return 0;
    if (com == nullptr || cube == nullptr) return PWR_E_NULL;
    PWR_REQUIRE_PTR(frames); PWR_REQUIRE_PTR(windows); PWR_REQUIRE_PTR(win_extent);
    if (static_cast<long long>(B) * kFetchGroups > 0x7fffffffLL) return PWR_E_SHAPE;
    FetchArgs a = {frames, elem, Hf, Wf, com, cube, aug, fx, fy, prefilter_margin, prefilter_umax, prefilter_vmax,
                   windows, win_h, win_w, reinterpret_cast<WinExtent*>(win_extent), fetched_bytes, status, B};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    sfr_fetch_plan_kernel<<<(B + 127) / 128, 128, 0, s>>>(a);
    if (int rc = launch_status()) return rc;
    const long long items = static_cast<long long>(B) * kFetchGroups;
    const long long cap = kFetchGrid;
    sfr_fetch_copy_kernel<<<static_cast<unsigned>(items < cap ? items : cap), kFetchThreads, 0, s>>>(a);
    return launch_status();
}

extern "C" int pwr_sfr_bb_filter(const uint16_t* raw, int Hf, int Wf, const double* boxes, float* frames_out, int B,
                                 void* stream) {
    if (int rc = check_frames(Hf, Wf, B)) return rc;
    if (B == 0) return 0;
    if (raw == nullptr || boxes == nullptr || frames_out == nullptr) return PWR_E_NULL;
    sfr_bb_kernel<<<B, kBbThreads, 0, static_cast<cudaStream_t>(stream)>>>(raw, boxes, frames_out, Hf, Wf);
    return launch_status();
}
