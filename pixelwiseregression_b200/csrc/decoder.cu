// Fused differentiable decoder for sm_100a: forward, backward, backward+loss, and the one-pass last stage.
//
// Replaces the ~16 ATen launches per stage of the reference
//   model.py:79-97   PlaneRegression.forward (post-conv): softmax / relu-sum
//                    normalisation and the soft-argmax sums over U and V
//   model.py:123-132 DepthRegression.forward (post-conv): masked, heat-map
//                    weighted depth
//   model.py:151     cat -> uvd
//   train.py:197-207 stage loss and everything autograd derives from the above
// by ONE forward and ONE backward kernel - or, for the last stage, one kernel for
// both.  Work unit ("item") = one (sample, joint) pair = one 64x64 logit map + one
// 64x64 depth map, 16 pixels (four 128-bit accesses) per thread per map at 256
// threads, so every map crosses HBM exactly once.  No tensor cores: nothing here is
// a contraction; the bound is HBM bandwidth.
//
// Kernels (which one an entry point launches is decided in the extern "C" functions
// at the end of the file; DESIGN.md section 4 has the measurements behind each):
//   decoder_fwd_kernel       one CTA per item; forward with the heat-map store, with
//                            or without the loss riding along (inner stages)
//   decoder_fwd_pipe_kernel  persistent, 3 CTAs/SM, bulk-TMA ring; forward WITHOUT the
//                            heat-map store (inference, last stage)
//   decoder_bwd_pipe_kernel  persistent, 1 CTA/SM, 2 x 64 KB stages; backward(+loss)
//                            with dense targets or dense upstream gradients
//   decoder_bwd_lean_kernel  persistent, 2 CTAs/SM; backward(+loss) with no dense map
//                            beyond z, D (compact targets, uvd-only loss)
//   decoder_bwd_kernel       one CTA per item; only the plane-only call (no depth branch) lands here
//   decoder_fused_kernel     persistent, 2 CTAs/SM, six-slot FIFO ring; last stage
//                            forward + loss + backward in one pass
//
// Pixel mapping: chunk c = tid + 256*i (i = 0..3) covers pixels 4c..4c+3,
// i.e. column x = 4*(tid & 15) + k and row y = (tid >> 4) + 16*i.  The
// coordinate filter of utils.py:24-35 (U = (x-32)/63, V = (y-32)/63) is
// synthesised from the indices: sums are taken over the integer offsets
// (x-32), (y-32) and scaled by 1/63 once.
#include <atomic>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace pwr {

// Element type of the conv outputs z, D and of their gradients: float32, or the
// float16 / bfloat16 tensors autocast produces under --mixed_precision (train.py:170-172).  The
// arithmetic is float32 either way (what autocast does for softmax / sum / mul with a float32
// operand); only the HBM traffic of z, D, gz, gD halves.
template <typename TZ> struct MapIO;
template <> struct MapIO<float> {
    static __device__ __forceinline__ float4 ld(const void* base, size_t e) {
        return __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(base) + e));
    }
    static __device__ __forceinline__ void st(void* base, size_t e, float4 v) {
        __stcs(reinterpret_cast<float4*>(static_cast<float*>(base) + e), v);
    }
    static __device__ __forceinline__ float4 smem(const void* slot, int chunk) {
        return static_cast<const float4*>(slot)[chunk];
    }
};
template <> struct MapIO<__half> {
    static __device__ __forceinline__ float4 cvt(uint2 r) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float4 ld(const void* base, size_t e) {
        return cvt(__ldcs(reinterpret_cast<const uint2*>(static_cast<const __half*>(base) + e)));
    }
    static __device__ __forceinline__ void st(void* base, size_t e, float4 v) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned int*>(&a);
        r.y = *reinterpret_cast<const unsigned int*>(&b);
        __stcs(reinterpret_cast<uint2*>(static_cast<__half*>(base) + e), r);
    }
    static __device__ __forceinline__ float4 smem(const void* slot, int chunk) {
        return cvt(static_cast<const uint2*>(slot)[chunk]);
    }
};
template <> struct MapIO<__nv_bfloat16> {
    static __device__ __forceinline__ float4 cvt(uint2 r) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float4 ld(const void* base, size_t e) {
        return cvt(__ldcs(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(base) + e)));
    }
    static __device__ __forceinline__ void st(void* base, size_t e, float4 v) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned int*>(&a);
        r.y = *reinterpret_cast<const unsigned int*>(&b);
        __stcs(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(base) + e), r);
    }
    static __device__ __forceinline__ float4 smem(const void* slot, int chunk) {
        return cvt(static_cast<const uint2*>(slot)[chunk]);
    }
};

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kEps = 1e-14f;             // model.py:89,128

struct PixelCoords {
    float xs;          // (x - 32) of the first of the thread's four columns
    float ys0;         // (y - 32) of the thread's first row; row i adds 16*i
};
__device__ __forceinline__ PixelCoords pixel_coords() {
    PixelCoords pc;
    pc.xs = static_cast<float>(static_cast<int>((threadIdx.x & 15) * 4) - 32);
    pc.ys0 = static_cast<float>(static_cast<int>(threadIdx.x >> 4) - 32);
    return pc;
}

__device__ __forceinline__ float comp(const float4& v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ void set_comp(float4& v, int k, float x) {
    if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}

// Extremum of the logits in the direction of the temperature sign, so that
// c*(z - ext) <= 0 for every pixel even when a trained w is negative.
__device__ __forceinline__ float block_extremum(const float4 (&zv)[kVec], bool want_max, float* scratch) {
    float e = want_max ? -INFINITY : INFINITY;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        if (want_max) e = fmaxf(fmaxf(fmaxf(e, zv[i].x), fmaxf(zv[i].y, zv[i].z)), zv[i].w);
        else          e = fminf(fminf(fminf(e, zv[i].x), fminf(zv[i].y, zv[i].z)), zv[i].w);
    }
    e = want_max ? warp_max(e) : warp_min(e);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = e;
    __syncthreads();
    float r = scratch[0];
#pragma unroll
    for (int wv = 1; wv < kWarps; ++wv) r = want_max ? fmaxf(r, scratch[wv]) : fminf(r, scratch[wv]);
    __syncthreads();
    return r;
}

// MUFU.EX2: 2 ulp, results below 2^-126 flush to zero (irrelevant after the
// 1/sum normalisation: such pixels are < 1e-38 of the map's mass).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Un-normalised heat value: 2^(c*z - off) for softmax, relu(z)+1e-14 for sum,
// z itself when the caller hands in an already normalised heat map.
template <int METHOD>
__device__ __forceinline__ float heat_raw(float z, float c, float off) {
    if (METHOD == PWR_METHOD_SOFTMAX) return ex2_approx(fmaf(z, c, -off));
    if (METHOD == PWR_METHOD_GIVEN) return z;
    return fmaxf(z, 0.f) + kEps;
}

// ---------------------------------------------------------------------------
// sparse targets: heat-map / depth-map targets evaluated on the fly from pwr_joint_taps
// ---------------------------------------------------------------------------
// cv::getGaussianKernel(7, 1.5) rounded to float32 (the dense targets are float32 roundings of the
// float64 blur; evaluating in float32 here differs by < 4e-7 relative, far inside the 1e-5 bound)
__constant__ float kGauss7f[7] = {0x1.2c18a51a3e5e7p-5, 0x1.c7ce552574441p-4, 0x1.bbe4f897eb627p-3,
                                  0x1.152db38ecae3ep-2, 0x1.bbe4f897eb627p-3, 0x1.c7ce552574441p-4,
                                  0x1.2c18a51a3e5e7p-5};

// weight with which a unit impulse at index t reaches output x through the 7-tap Gaussian with
// BORDER_REFLECT_101 on a 64-long axis (same decomposition as gauss_reach in sfr.cu)
__device__ __forceinline__ float gauss_reach_f(int x, int t) {
    float wgt = 0.f;
    const int k0 = t - x + 3;
    if (k0 >= 0 && k0 <= 6) wgt += kGauss7f[k0];
    const int k1 = 3 - x - t;
    if (t >= 1 && k1 >= 0 && k1 <= 6) wgt += kGauss7f[k1];
    const int k2 = 2 * kLabel + 1 - x - t;
    if (t <= kLabel - 2 && k2 >= 0 && k2 <= 6) wgt += kGauss7f[k2];
    return wgt;
}

// Footprint table: the blurred 4-tap splat is non-zero on <= 8x8 pixels.  Before a CTA walks the
// 4096 pixels of a map, 121 threads evaluate the heat target on the 11 x 11 candidate grid
// (rows ty0-3..ty0+3 then ty1..ty1+3, same for columns; entries that fall outside the map or
// repeat an earlier row/column stay 0 and are never looked up) into shared memory; the hot loop
// then only does an integer range test per 4-pixel chunk and, inside the footprint, a table lookup.
constexpr int kCand = 11;
constexpr int kFootprint = kCand * kCand;

struct TapsIdx { int tx0, tx1, ty0, ty1, ok; float cdn; };

// `raw` = the 16 words of a pwr_joint_taps record (global or shared memory)
__device__ __forceinline__ TapsIdx taps_index(const uint32_t* raw) {
    const pwr_joint_taps* p = reinterpret_cast<const pwr_joint_taps*>(raw);
    TapsIdx t;
    t.tx0 = p->tx0; t.tx1 = p->tx1; t.ty0 = p->ty0; t.ty1 = p->ty1; t.ok = p->ok;
    t.cdn = static_cast<float>(p->cd_norm);
    return t;
}

// one candidate of the footprint table (thread `c` of the first kFootprint threads)
__device__ __forceinline__ float footprint_entry(const uint32_t* raw, int c) {
    const pwr_joint_taps* p = reinterpret_cast<const pwr_joint_taps*>(raw);
    const int sy = c / kCand, sx = c - sy * kCand;
    const int ty0 = p->ty0, ty1 = p->ty1, tx0 = p->tx0, tx1 = p->tx1;
    const int y = sy < 7 ? ty0 + sy - 3 : ty1 + sy - 7;
    const int x = sx < 7 ? tx0 + sx - 3 : tx1 + sx - 7;
    if (!p->ok || y < 0 || y >= kLabel || x < 0 || x >= kLabel) return 0.f;
    if ((sy >= 7 && abs(y - ty0) <= 3) || (sx >= 7 && abs(x - tx0) <= 3)) return 0.f;
    const float a = static_cast<float>(p->tap[0]), b = static_cast<float>(p->tap[1]);
    const float cc = static_cast<float>(p->tap[2]), d = static_cast<float>(p->tap[3]);
    const float wy0 = gauss_reach_f(y, ty0), wy1 = gauss_reach_f(y, ty1);
    const float wx0 = gauss_reach_f(x, tx0), wx1 = gauss_reach_f(x, tx1);
    return wy0 * (a * wx0 + b * wx1) + wy1 * (cc * wx0 + d * wx1);
}

// targets of the 4 pixels (row y, columns x0..x0+3): heat from the table, Dmap =
// (cd/cube - label_img) * [heat > 0] * mask (datasets.py:372-374, 380); zero outside the footprint
__device__ __forceinline__ void sparse_lookup(const TapsIdx& t, const float* fp, int y, int x0, const float4& l4,
                                              const float4& m4, float4& hg, float4& dg) {
    hg = make_float4(0.f, 0.f, 0.f, 0.f);
    dg = hg;
    int sy = -1;
    if (abs(y - t.ty0) <= 3) sy = y - t.ty0 + 3;
    else if (y - t.ty1 >= 0 && y - t.ty1 <= 3) sy = 7 + y - t.ty1;
    if (!t.ok || sy < 0) return;
    if (!((x0 + 3 >= t.tx0 - 3 && x0 <= t.tx0 + 3) || (x0 + 3 >= t.tx1 && x0 <= t.tx1 + 3))) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x0 + k;
        int sx = -1;
        if (abs(x - t.tx0) <= 3) sx = x - t.tx0 + 3;
        else if (x - t.tx1 >= 0 && x - t.tx1 <= 3) sx = 7 + x - t.tx1;
        const float h = sx >= 0 ? fp[sy * kCand + sx] : 0.f;
        set_comp(hg, k, h);
        set_comp(dg, k, (h > 0.f && comp(m4, k) != 0.f) ? t.cdn - comp(l4, k) : 0.f);
    }
}

// LOSS template parameter of the kernels below
enum { LOSS_NONE = 0, LOSS_DENSE = 1, LOSS_SPARSE = 2 };   // no loss | dense target maps | pwr_joint_taps

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int METHOD, int LOSS, typename TZ>
__global__ void __launch_bounds__(kThreads)
decoder_fwd_kernel(const void* __restrict__ z, const float* __restrict__ w, const void* __restrict__ D,
                   const float* __restrict__ L, const float* __restrict__ m,
                   const float* __restrict__ heat_gt, const float* __restrict__ dmap_gt,
                   const float* __restrict__ uvd_gt, const pwr_joint_taps* __restrict__ taps, float* __restrict__ H,
                   float* __restrict__ uvd, float* __restrict__ stats, float* __restrict__ loss_partial, int J) {
    __shared__ float scratch[kWarps * 5];
    const int bj = blockIdx.x;
    const int b = bj / J;
    const int j = bj - b * J;
    const size_t off = static_cast<size_t>(bj) * kMap + threadIdx.x * 4;
    const size_t offb = static_cast<size_t>(b) * kMap + threadIdx.x * 4;
    const bool depth = (D != nullptr);          // PlaneRegression alone: no depth branch

    float4 zv[kVec], dv[kVec], lv[kVec], mv[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) zv[i] = MapIO<TZ>::ld(z, off + i * (kThreads * 4));
    if (depth) {
#pragma unroll
        for (int i = 0; i < kVec; ++i) mv[i] = ld_keep(m + offb + i * (kThreads * 4));
#pragma unroll
        for (int i = 0; i < kVec; ++i) lv[i] = ld_keep(L + offb + i * (kThreads * 4));
#pragma unroll
        for (int i = 0; i < kVec; ++i) dv[i] = MapIO<TZ>::ld(D, off + i * (kThreads * 4));
    } else {
#pragma unroll
        for (int i = 0; i < kVec; ++i) mv[i] = lv[i] = dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    float c = 0.f, shift = 0.f, zext = 0.f;
    if (METHOD == PWR_METHOD_SOFTMAX) {
        c = w[j] * kLog2e;
        zext = block_extremum(zv, c >= 0.f, scratch);
        shift = zext * c;                          // the backward recomputes exactly this product
    }

    const PixelCoords pc = pixel_coords();
    constexpr bool sparse = (LOSS == LOSS_SPARSE);
    __shared__ float fp[sparse ? kFootprint : 1];
    TapsIdx tp;
    if (sparse) {
        const uint32_t* raw = reinterpret_cast<const uint32_t*>(taps + bj);
        tp = taps_index(raw);
        if (threadIdx.x < kFootprint) fp[threadIdx.x] = footprint_entry(raw, threadIdx.x);
        __syncthreads();
    }
    float4 hgs[sparse ? kVec : 1];              // sparse heat targets, kept for the second loop
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum e, e*(x-32), e*(y-32), e*m, e*m*m*(D+L)
    float ld2 = 0.f;                            // sum (D - Dgt)^2
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const float ys = pc.ys0 + 16.f * i;
        float rowsum = 0.f;
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LOSS) {
            if (sparse) sparse_lookup(tp, fp, static_cast<int>(threadIdx.x >> 4) + 16 * i, static_cast<int>(threadIdx.x & 15) * 4,
                                      lv[i], mv[i], hgs[sparse ? i : 0], dg);
            else dg = ld_stream(dmap_gt + off + i * (kThreads * 4));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float e = heat_raw<METHOD>(comp(zv[i], k), c, shift);
            set_comp(zv[i], k, e);
            const float mk = comp(mv[i], k);
            const float em = e * mk;
            rowsum += e;
            acc[1] = fmaf(e, pc.xs + static_cast<float>(k), acc[1]);
            acc[3] += em;
            acc[4] = fmaf(em, mk * (comp(dv[i], k) + comp(lv[i], k)), acc[4]);
            if (LOSS) { const float ed = comp(dv[i], k) - comp(dg, k); ld2 = fmaf(ed, ed, ld2); }
        }
        acc[0] += rowsum;
        acc[2] = fmaf(rowsum, ys, acc[2]);
    }
    block_sum<5>(acc, scratch);

    const float inv_s = (METHOD == PWR_METHOD_GIVEN) ? 1.f : 1.f / acc[0];
    float lh2[2] = {0.f, ld2};                  // sum (p - Hgt)^2, sum (D - Dgt)^2
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        float4 h = zv[i];
        h.x *= inv_s; h.y *= inv_s; h.z *= inv_s; h.w *= inv_s;
        if (LOSS) {
            const float4 hg = sparse ? hgs[sparse ? i : 0] : ld_stream(heat_gt + off + i * (kThreads * 4));
            const float e0 = h.x - hg.x, e1 = h.y - hg.y, e2 = h.z - hg.z, e3 = h.w - hg.w;
            lh2[0] += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
        }
        if (H != nullptr) st_stream(H + off + i * (kThreads * 4), h);
    }
    if (LOSS) block_sum<2>(lh2, scratch);
    if (threadIdx.x == 0) {
        const float den = fmaf(acc[3], inv_s, kEps);
        const float d = (acc[4] * inv_s) / den;
        const float u = acc[1] * inv_s / 63.f, v = acc[2] * inv_s / 63.f;
        uvd[bj * 3 + 0] = u;
        uvd[bj * 3 + 1] = v;
        uvd[bj * 3 + 2] = d;
        if (stats != nullptr)
            reinterpret_cast<float4*>(stats)[bj] = make_float4(zext, inv_s, den, d);
        if (LOSS) {
            const float eu = u - uvd_gt[bj * 3 + 0], ev = v - uvd_gt[bj * 3 + 1], ed = d - uvd_gt[bj * 3 + 2];
            loss_partial[bj * 3 + 0] = lh2[0];
            loss_partial[bj * 3 + 1] = lh2[1];
            loss_partial[bj * 3 + 2] = eu * eu + ev * ev + ed * ed;
        }
    }
}

// ---------------------------------------------------------------------------
// backward (optionally fused with the stage loss)
// ---------------------------------------------------------------------------
struct LossCoef {
    float cu;   // loss_scale * 2*alpha / N
    float ch;   // loss_scale * 2*(1-alpha)*lambda_h / N
    float cd;   // loss_scale * 2*(1-alpha)*lambda_d / N
    const float* scale_dev;   // optional device scalar multiplying all three (upstream d/d(loss), e.g. a GradScaler)
};

// LOSS = false: plain backward.  LOSS = true: adds d(loss)/d(.) of train.py:197-205;
// the target maps are only read when they matter (non-zero map weights or
// loss_partial requested), so the default alpha = 1 costs no extra traffic
// unless the caller wants the logged loss values.
template <int METHOD, int LOSS, typename TZ>
__global__ void __launch_bounds__(kThreads)
decoder_bwd_kernel(const void* __restrict__ z, const float* __restrict__ w, const void* __restrict__ D,
                   const float* __restrict__ L, const float* __restrict__ m, const float* __restrict__ stats,
                   const float* __restrict__ uvd, const float* __restrict__ g_uvd,
                   const float* __restrict__ gH_up, const void* __restrict__ gD_up,
                   const float* __restrict__ heat_gt, const float* __restrict__ dmap_gt,
                   const float* __restrict__ uvd_gt, const pwr_joint_taps* __restrict__ taps, LossCoef coef,
                   void* __restrict__ gz, void* __restrict__ gD, float* __restrict__ gw_partial,
                   float* __restrict__ loss_partial, int J) {
    __shared__ float scratch[kWarps * 5];
    if (LOSS && coef.scale_dev != nullptr) {
        const float up = *coef.scale_dev;
        coef.cu *= up; coef.ch *= up; coef.cd *= up;
    }
    const int bj = blockIdx.x;
    const int b = bj / J;
    const int j = bj - b * J;
    const size_t off = static_cast<size_t>(bj) * kMap + threadIdx.x * 4;
    const size_t offb = static_cast<size_t>(b) * kMap + threadIdx.x * 4;
    const bool depth = (D != nullptr);
    const bool map_loss = LOSS && (loss_partial != nullptr || coef.ch != 0.f || coef.cd != 0.f);
    const bool sparse = (LOSS == LOSS_SPARSE) && map_loss;
    __shared__ float fp[LOSS == LOSS_SPARSE ? kFootprint : 1];
    TapsIdx tp;
    if (LOSS == LOSS_SPARSE) {
        const uint32_t* raw = reinterpret_cast<const uint32_t*>(taps + bj);
        tp = taps_index(raw);
        if (threadIdx.x < kFootprint) fp[threadIdx.x] = footprint_entry(raw, threadIdx.x);
        __syncthreads();
    }

    float4 zv[kVec], pv[kVec], gv[kVec];   // logits, heat p, dL/dp
#pragma unroll
    for (int i = 0; i < kVec; ++i) zv[i] = MapIO<TZ>::ld(z, off + i * (kThreads * 4));

    const float4 st = reinterpret_cast<const float4*>(stats)[bj];   // (z extremum, 1/sum, den, d)
    const float wj = (METHOD == PWR_METHOD_SOFTMAX) ? w[j] : 1.f;
    const float c = wj * kLog2e;
    const float shift = st.x * c;
    float gu = 0.f, gvv = 0.f, gd = 0.f, lu = 0.f;
    if (g_uvd != nullptr) { gu = g_uvd[bj * 3 + 0]; gvv = g_uvd[bj * 3 + 1]; gd = g_uvd[bj * 3 + 2]; }
    if (LOSS) {
        const float eu = uvd[bj * 3 + 0] - uvd_gt[bj * 3 + 0];
        const float ev = uvd[bj * 3 + 1] - uvd_gt[bj * 3 + 1];
        const float ed = uvd[bj * 3 + 2] - uvd_gt[bj * 3 + 2];
        gu = fmaf(coef.cu, eu, gu); gvv = fmaf(coef.cu, ev, gvv); gd = fmaf(coef.cu, ed, gd);
        lu = eu * eu + ev * ev + ed * ed;
    }
    const float gu63 = gu * (1.f / 63.f), gv63 = gvv * (1.f / 63.f);
    const float gdd = depth ? __fdividef(gd, st.z) : 0.f;        // g_d / den
    const float dcoord = st.w, zref = st.x;
    const PixelCoords pc = pixel_coords();
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // acc: sum gp*p, sum (p-Hgt)^2, sum (D-Dgt)^2, T1 = sum p*gp*(z-zref), T2 = sum p*(z-zref):
    // dL/dw = sum_px p*(gp - S1)*z = T1 - S1*T2 (sum_px p*(gp - S1) = 0 frees the reference point),
    // so one block reduction serves everything and z is dead after this loop
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned int zpos = 0;               // PWR_METHOD_SUM: bit (4*i+k) = z > 0
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const size_t o = off + i * (kThreads * 4);
        const size_t ob = offb + i * (kThreads * 4);
        float4 d4 = zero4, l4 = zero4, m4 = zero4, hg = zero4, dg = zero4, uh = zero4, ud = zero4;
        if (depth) { d4 = MapIO<TZ>::ld(D, o); l4 = ld_keep(L + ob); m4 = ld_keep(m + ob); }
        if (LOSS == LOSS_SPARSE) {
            if (sparse) sparse_lookup(tp, fp, static_cast<int>(threadIdx.x >> 4) + 16 * i, static_cast<int>(threadIdx.x & 15) * 4, l4, m4, hg, dg);
        } else if (map_loss) { hg = ld_stream(heat_gt + o); if (depth) dg = ld_stream(dmap_gt + o); }
        if (gH_up != nullptr) uh = ld_stream(gH_up + o);
        if (gD_up != nullptr) ud = MapIO<TZ>::ld(gD_up, o);
        const float gyrow = gv63 * (pc.ys0 + 16.f * i);
        float4 gd4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float zk = comp(zv[i], k);
            const float p = heat_raw<METHOD>(zk, c, shift) * st.y;
            const float mk = comp(m4, k), dk = comp(d4, k);
            const float rec = mk * (dk + comp(l4, k));
            float gp = fmaf(gu63, pc.xs + static_cast<float>(k), gyrow);
            gp = fmaf(gdd * mk, rec - dcoord, gp);
            float gdk = gdd * p * mk * mk;
            if (LOSS) {
                const float eh = p - comp(hg, k);
                const float ed = dk - comp(dg, k);
                gp = fmaf(coef.ch, eh, gp);
                gdk = fmaf(coef.cd, ed, gdk);
                acc[1] = fmaf(eh, eh, acc[1]);
                acc[2] = fmaf(ed, ed, acc[2]);
            }
            gp += comp(uh, k);
            gdk += comp(ud, k);
            const float pg = gp * p;
            acc[0] += pg;
            if (METHOD == PWR_METHOD_SOFTMAX) {
                acc[3] = fmaf(pg, zk - zref, acc[3]);
                acc[4] = fmaf(p, zk - zref, acc[4]);
            }
            if (METHOD == PWR_METHOD_SUM && zk > 0.f) zpos |= 1u << (4 * i + k);
            set_comp(pv[i], k, p);
            set_comp(gv[i], k, gp);
            set_comp(gd4, k, gdk);
        }
        if (gD != nullptr) MapIO<TZ>::st(gD, o, gd4);
    }
    if (METHOD != PWR_METHOD_GIVEN || LOSS) block_sum<5>(acc, scratch);

    const float s1 = acc[0];
    if (gz != nullptr) {
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            float4 g4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g;
                if (METHOD == PWR_METHOD_SOFTMAX) {
                    g = wj * (comp(pv[i], k) * (comp(gv[i], k) - s1));                 // w * dL/d(w z)
                } else if (METHOD == PWR_METHOD_SUM) {
                    g = ((zpos >> (4 * i + k)) & 1u) ? (comp(gv[i], k) - s1) * st.y : 0.f;   // through relu and 1/sum
                } else {
                    g = comp(gv[i], k);                                        // heat map given: dL/dp itself
                }
                set_comp(g4, k, g);
            }
            MapIO<TZ>::st(gz, off + i * (kThreads * 4), g4);
        }
    }
    if (METHOD == PWR_METHOD_SOFTMAX && gw_partial != nullptr && threadIdx.x == 0)
        gw_partial[bj] = acc[3] - s1 * acc[4];
    if (LOSS && loss_partial != nullptr && threadIdx.x == 0) {
        loss_partial[bj * 3 + 0] = acc[1];
        loss_partial[bj * 3 + 1] = acc[2];
        loss_partial[bj * 3 + 2] = lu;
    }
}

// ---------------------------------------------------------------------------
// backward, persistent + TMA-pipelined variant (the hot configurations)
// ---------------------------------------------------------------------------
// The direct-load kernel above keeps three 16-float arrays live across a block
// reduction (104 registers -> 2 CTAs/SM) and only has a few 128-bit loads in
// flight per thread, so it is latency- rather than bandwidth-bound (ncu r1:
// 55 % DRAM, 45 % of stall samples on the first use of a loaded value).  Here
// one CTA per SM stays resident and walks a contiguous range of (b,j) items;
// a single elected thread streams the next item's maps into a shared-memory
// ring with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx) while
// all 512 threads compute the current item out of shared memory.  Bytes in
// flight no longer depend on registers or occupancy: a whole 64 KB stage per
// SM is always outstanding.  L and m are fetched once per sample, not once per
// (sample, joint).
//
// A stage has four 16 KB slots: z, D and two optional maps whose meaning the
// host picks: (heat_gt, dmap_gt) for the fused loss, or (gH_up, gD_up) for
// dense upstream gradients.  Configurations that need both pairs at once (an inner
// stage trained with alpha < 1 on dense targets) run the six-slot instantiation (BOTH):
// 2 x 96 KB of stages, with L, m read through L2 instead of a shared-memory pair.
#ifndef PWR_PIPE_THREADS
#define PWR_PIPE_THREADS 512
#endif
constexpr int kPipeThreads = PWR_PIPE_THREADS;
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeVec = kMap / 4 / kPipeThreads;     // float4 chunks per thread per map = 2
constexpr int kPipeStages = 2;
constexpr int kSlotBytes = kMap * 4;                  // 16 KB
constexpr int kPipeSmemBytes = kPipeStages * 4 * kSlotBytes + 2 * 2 * kSlotBytes + 64;
// six-slot variant (dense targets AND dense upstream gradients at once: an inner stage with alpha < 1):
// z, D, heat_gt, dmap_gt, gH_up, gD_up per stage; L, m are then read through L2 instead of a shared-memory pair
constexpr int kPipeSmemBytesBoth = kPipeStages * 6 * kSlotBytes + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Block-wide sums for the persistent kernel.  The caller alternates between two scratch buffers,
// so the usual trailing barrier (scratch reuse) is not needed: the next writer of this buffer is two
// items away and every item ends with a __syncthreads.
template <int N>
__device__ __forceinline__ void pipe_block_sum(float (&v)[N], float* scratch) {
    static_assert(kPipeWarps == 8 || kPipeWarps == 16 || kPipeWarps == 32, "second level is a shuffle tree over the warps");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * kPipeWarps + warp] = v[i];
    }
    __syncthreads();
    // every warp reduces the 16 partials of each value itself: one shared load per lane and four
    // shuffles instead of 16 loads + adds per thread
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float s = scratch[i * kPipeWarps + (lane & (kPipeWarps - 1))];
#pragma unroll
        for (int o = kPipeWarps / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        v[i] = s;
    }
}

// ---------------------------------------------------------------------------
// forward, persistent + TMA-pipelined variant (no fused loss: inference and the plain training forward)
// ---------------------------------------------------------------------------
// The one-CTA-per-(b,j) forward above is a chain load -> extremum -> sums -> store with three CTAs
// per SM; when the heat maps are not stored (inference, last stage) nothing overlaps the load
// latency and it reaches 79 % of the copy peak (ncu r1, HAND17 B=4096).  Here a few persistent CTAs
// per SM each walk a contiguous range of items; the z / D maps of a CTA's next item(s) are always in
// flight (ring of 32 KB stages filled by 1-D bulk TMA copies), the CTAs of an SM interleave their
// reduction phases, and the five block sums cost 8 shuffles per warp instead of 25 (a forward
// item moves only 32 KB, so the shuffle pipe - one warp instruction per clock per SM - is a
// first-order cost: a lock-step 512-thread CTA with tree reductions was measured 40 % SLOWER than
// the direct kernel).  L and m stay in registers across the J consecutive items of a sample (with
// per-item cached loads instead the L1 thrashed - 0.5 % hit rate - and 30 % of the stall samples
// sat on their first use).  Measured at HAND17 B=4096 without the heat-map store: direct 0.576 ms,
// this kernel 0.452 ms (256 threads, 2 stages, 3 CTAs/SM) / 0.488 (3 stages, 2 CTAs/SM) / 0.528
// (128 threads, 1 stage, 6 CTAs/SM, L/m not register-resident).
#ifndef PWR_FWD_PIPE_THREADS
#define PWR_FWD_PIPE_THREADS 256
#endif
#ifndef PWR_FWD_PIPE_STAGES
#define PWR_FWD_PIPE_STAGES 2
#endif
constexpr int kFwdThreads = PWR_FWD_PIPE_THREADS;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kFwdVec = kMap / 4 / kFwdThreads;       // float4 chunks per thread per map
constexpr int kFwdStages = PWR_FWD_PIPE_STAGES;
constexpr int kFwdSmemBytes = kFwdStages * 2 * kSlotBytes + 64;
constexpr int kFwdCtasPerSm = (227 * 1024 / (kFwdSmemBytes + 1024)) < (2048 / kFwdThreads)
                                  ? (227 * 1024 / (kFwdSmemBytes + 1024)) : (2048 / kFwdThreads);
static_assert(kFwdWarps == 4 || kFwdWarps == 8 || kFwdWarps == 16, "second level reads whole float4s of partials");

// Warp totals of five values in 8 shuffles: at every butterfly step a lane keeps only the values its
// half of the pair is responsible for, so the number of live values halves as the partner distance
// does.  Totals end up in lane 0 (v0), 4 (v1), 8 (v2), 16 (v3), 20 (v4); fixed order => deterministic.
__device__ __forceinline__ float warp_sum5_scattered(const float (&v)[5]) {
    const int lane = threadIdx.x & 31;
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
    // xor 16: lower half keeps v0 v1 v2, upper half keeps v3 v4
    float r = __shfl_xor_sync(0xffffffffu, b16 ? v[0] : v[3], 16);
    const float x0 = (b16 ? v[3] : v[0]) + r;
    r = __shfl_xor_sync(0xffffffffu, b16 ? v[1] : v[4], 16);
    const float x1 = (b16 ? v[4] : v[1]) + r;
    r = __shfl_xor_sync(0xffffffffu, b16 ? v[2] : 0.f, 16);
    const float x2 = b16 ? 0.f : v[2] + r;
    // xor 8: bit-3-clear lanes keep x0 x1, bit-3-set lanes keep x2
    r = __shfl_xor_sync(0xffffffffu, b8 ? x0 : x2, 8);
    const float p = (b8 ? x2 : x0) + r;
    r = __shfl_xor_sync(0xffffffffu, b8 ? x1 : 0.f, 8);
    const float q = x1 + r;                                   // meaningful on bit-3-clear lanes only
    // xor 4: bit-3-clear lanes split (p | q); bit-3-set lanes keep reducing p
    r = __shfl_xor_sync(0xffffffffu, b8 ? p : (b4 ? p : q), 4);
    float wv = (b8 ? p : (b4 ? q : p)) + r;
    wv += __shfl_xor_sync(0xffffffffu, wv, 2);
    wv += __shfl_xor_sync(0xffffffffu, wv, 1);
    return wv;
}
// scratch layout [5][NW]; lanes 0, 4, 8, 16, 20 hold value 0..4
template <int NW = kFwdWarps>
__device__ __forceinline__ void store_scattered5(float wv, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx = -1;
    if (lane == 0) idx = 0; else if (lane == 4) idx = 1; else if (lane == 8) idx = 2;
    else if (lane == 16) idx = 3; else if (lane == 20) idx = 4;
    if (idx >= 0) scratch[idx * NW + warp] = wv;
}
template <int NW = kFwdWarps>
__device__ __forceinline__ float sum_partials(const float* row) {      // NW floats, fixed order
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NW / 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(row)[i];
        s += (t.x + t.y) + (t.z + t.w);
    }
    return s;
}

struct FwdPipeArgs {
    const void* z; const float* w; const void* D; const float* L; const float* m;
    float* H; float* uvd; float* stats;
    int J; int items;
};

template <int METHOD, typename TZ>
__global__ void __launch_bounds__(kFwdThreads, kFwdCtasPerSm)
decoder_fwd_pipe_kernel(FwdPipeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);                               // [stages][2][4096]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + kFwdStages * 2 * kMap);     // [stages]
    __shared__ __align__(16) float scr_e[2][kFwdWarps];       // double-buffered: one barrier per reduction
    __shared__ __align__(16) float scr_s[2][5 * kFwdWarps];

    const int tid = threadIdx.x;
    const long long first = static_cast<long long>(a.items) * blockIdx.x / gridDim.x;
    const long long last = static_cast<long long>(a.items) * (blockIdx.x + 1) / gridDim.x;
    if (first >= last) return;
    if (tid == 0) {
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    constexpr uint32_t kZBytes = kMap * sizeof(TZ);
    // producer (thread 0): stream item `it` into the stage of its index
    auto issue = [&](long long it, int kk) {
        float* st = stage_base + (kk % kFwdStages) * 2 * kMap;
        uint64_t* bar = &full[kk % kFwdStages];
        const size_t off = static_cast<size_t>(it) * kMap;
        mbar_expect_tx(bar, 2 * kZBytes);
        bulk_g2s(st, static_cast<const TZ*>(a.z) + off, kZBytes, bar);
        bulk_g2s(st + kMap, static_cast<const TZ*>(a.D) + off, kZBytes, bar);
    };
    if (tid == 0) {
        for (int i = 0; i < kFwdStages && first + i < last; ++i) issue(first + i, i);
    }

    int b_cur = static_cast<int>(first / a.J);
    int j_cur = static_cast<int>(first - static_cast<long long>(b_cur) * a.J);
    const float xs = static_cast<float>(static_cast<int>((tid & 15) * 4) - 32);
    const float ys0 = static_cast<float>(static_cast<int>(tid >> 4) - 32);     // row of chunk i: + (kFwdThreads/16)*i
    constexpr bool kPreloadLM = kFwdVec <= 4;
    float4 lv[kPreloadLM ? kFwdVec : 1], mv[kPreloadLM ? kFwdVec : 1];
    int b_loaded = -1;
    float c_next = (METHOD == PWR_METHOD_SOFTMAX) ? a.w[j_cur] * kLog2e : 0.f;     // fetched one item ahead
    int k = 0;
    for (long long it = first; it < last; ++it, ++k) {
        const int s = k % kFwdStages;
        const uint32_t parity = (k / kFwdStages) & 1;
        const float* sz = stage_base + s * 2 * kMap;
        const float* sD = sz + kMap;
        const float c = c_next;
        // label and mask of the sample stay in registers across its J items (with 32 pixels per
        // thread the 64 registers are not there: cached loads inside pass 2 instead)
        const size_t offb = static_cast<size_t>(b_cur) * kMap + tid * 4;
        if (kPreloadLM && b_cur != b_loaded) {
            b_loaded = b_cur;
#pragma unroll
            for (int i = 0; i < (kPreloadLM ? kFwdVec : 1); ++i) mv[i] = ld_keep(a.m + offb + i * (kFwdThreads * 4));
#pragma unroll
            for (int i = 0; i < (kPreloadLM ? kFwdVec : 1); ++i) lv[i] = ld_keep(a.L + offb + i * (kFwdThreads * 4));
        }

        mbar_wait(&full[s], parity);

        float4 zv[kFwdVec];
#pragma unroll
        for (int i = 0; i < kFwdVec; ++i) zv[i] = MapIO<TZ>::smem(sz, tid + i * kFwdThreads);
        float shift = 0.f, zext = 0.f;
        if (METHOD == PWR_METHOD_SOFTMAX) {
            const bool want_max = c >= 0.f;
            float e = want_max ? -INFINITY : INFINITY;
#pragma unroll
            for (int i = 0; i < kFwdVec; ++i) {
                if (want_max) e = fmaxf(fmaxf(fmaxf(e, zv[i].x), fmaxf(zv[i].y, zv[i].z)), zv[i].w);
                else          e = fminf(fminf(fminf(e, zv[i].x), fminf(zv[i].y, zv[i].z)), zv[i].w);
            }
            e = want_max ? warp_max(e) : warp_min(e);
            float* se = scr_e[k & 1];
            if ((tid & 31) == 0) se[tid >> 5] = e;
            __syncthreads();
            zext = se[0];
#pragma unroll
            for (int wv = 1; wv < kFwdWarps; ++wv) zext = want_max ? fmaxf(zext, se[wv]) : fminf(zext, se[wv]);
            shift = zext * c;                      // the backward recomputes exactly this product
        }
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum e, e*(x-32), e*(y-32), e*m, e*m*m*(D+L)
#pragma unroll
        for (int i = 0; i < kFwdVec; ++i) {
            const float4 d4 = MapIO<TZ>::smem(sD, tid + i * kFwdThreads);
            const float4 m4 = kPreloadLM ? mv[kPreloadLM ? i : 0] : ld_keep(a.m + offb + i * (kFwdThreads * 4));
            const float4 l4 = kPreloadLM ? lv[kPreloadLM ? i : 0] : ld_keep(a.L + offb + i * (kFwdThreads * 4));
            const float ys = ys0 + static_cast<float>(kFwdThreads / 16) * i;
            float rowsum = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float e = heat_raw<METHOD>(comp(zv[i], kk), c, shift);
                set_comp(zv[i], kk, e);
                const float mk = comp(m4, kk);
                const float em = e * mk;
                rowsum += e;
                acc[1] = fmaf(e, xs + static_cast<float>(kk), acc[1]);
                acc[3] += em;
                acc[4] = fmaf(em, mk * (comp(d4, kk) + comp(l4, kk)), acc[4]);
            }
            acc[0] += rowsum;
            acc[2] = fmaf(rowsum, ys, acc[2]);
        }
        float* ss = scr_s[k & 1];
        store_scattered5(warp_sum5_scattered(acc), ss);
        __syncthreads();
        // every thread is past its shared-memory reads of this item: its stage can take item k + stages
        if (tid == 0 && it + kFwdStages < last) issue(it + kFwdStages, k + kFwdStages);
        if (++j_cur == a.J) { j_cur = 0; ++b_cur; }
        if (METHOD == PWR_METHOD_SOFTMAX && it + 1 < last) c_next = a.w[j_cur] * kLog2e;

        const float inv_s = (METHOD == PWR_METHOD_GIVEN) ? 1.f : 1.f / sum_partials(ss);
        if (a.H != nullptr) {
            float* Hp = a.H + static_cast<size_t>(it) * kMap;
#pragma unroll
            for (int i = 0; i < kFwdVec; ++i) {
                float4 h = zv[i];
                h.x *= inv_s; h.y *= inv_s; h.z *= inv_s; h.w *= inv_s;
                st_stream(Hp + (tid + i * kFwdThreads) * 4, h);
            }
        }
        if (tid == 0) {
            const float su = sum_partials(ss + kFwdWarps), sv = sum_partials(ss + 2 * kFwdWarps);
            const float sm = sum_partials(ss + 3 * kFwdWarps), sd = sum_partials(ss + 4 * kFwdWarps);
            const float den = fmaf(sm, inv_s, kEps);
            const float d = (sd * inv_s) / den;
            float* o = a.uvd + static_cast<size_t>(it) * 3;
            o[0] = su * inv_s / 63.f;
            o[1] = sv * inv_s / 63.f;
            o[2] = d;
            if (a.stats != nullptr) reinterpret_cast<float4*>(a.stats)[it] = make_float4(zext, inv_s, den, d);
        }
    }
}

struct PipeArgs {
    const void* z; const float* w; const void* D; const float* L; const float* m;
    const float* stats; const float* uvd; const float* g_uvd;
    const float* slot2; const void* slot3;     // (heat_gt, dmap_gt) or (gH_up, gD_up); either may be NULL
    const float* slot4; const void* slot5;     // BOTH variant only: (gH_up, gD_up) next to the targets in slot2/3
    const float* uvd_gt;
    const pwr_joint_taps* taps;                // sparse targets (then slot2/slot3 carry no targets)
    LossCoef coef;
    void* gz; void* gD; float* gw_partial; float* loss_partial;
    int J; int items;
    int slots_are_targets;                     // 1: slot2/3 = heat_gt/dmap_gt, 0: = gH_up/gD_up
};

template <int METHOD, int LOSS, typename TZ, bool BOTH = false>
__global__ void __launch_bounds__(kPipeThreads, 1)
decoder_bwd_pipe_kernel(PipeArgs a) {
    static_assert(!BOTH || LOSS == LOSS_DENSE, "six slots = dense targets + dense upstream gradients");
    constexpr int kSlots = BOTH ? 6 : 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);                               // [stages][kSlots][4096]
    float* lm_base = stage_base + kPipeStages * kSlots * kMap;                            // [2][2][4096] (not BOTH)
    uint64_t* full = reinterpret_cast<uint64_t*>(lm_base + (BOTH ? 0 : 2 * 2 * kMap));    // [stages]
    __shared__ float scratch[2][kPipeWarps * 5];      // double-buffered: one barrier per reduction
    // Per-item scalars (stats, upstream / predicted / target uvd, the 64-byte taps record) are
    // prefetched one item ahead by the first 32 threads: the global loads are issued at the top of
    // iteration k and parked in shared memory at its end, so no item starts with an exposed
    // L2 / HBM round trip.  Word layout: 0-3 stats, 4-6 g_uvd, 7-9 uvd, 10-12 uvd_gt, 16-31 taps.
    __shared__ __align__(16) uint32_t scal[2][32];
    __shared__ float fp[LOSS == LOSS_SPARSE ? kFootprint : 1];

    const int tid = threadIdx.x;
    const long long first = static_cast<long long>(a.items) * blockIdx.x / gridDim.x;
    const long long last = static_cast<long long>(a.items) * (blockIdx.x + 1) / gridDim.x;
    if (first >= last) return;

    if (LOSS && a.coef.scale_dev != nullptr) {
        const float up = *a.coef.scale_dev;
        a.coef.cu *= up; a.coef.ch *= up; a.coef.cd *= up;
    }
    if (tid == 0) {
        for (int s = 0; s < kPipeStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const bool has2 = a.slot2 != nullptr, has3 = a.slot3 != nullptr;
    const bool has4 = BOTH && a.slot4 != nullptr, has5 = BOTH && a.slot5 != nullptr;
    const bool tg = BOTH || a.slots_are_targets != 0;
    constexpr uint32_t kZBytes = kMap * sizeof(TZ);                  // z, D (and gD_up) in the conv dtype
    const uint32_t slot3_bytes = tg ? kSlotBytes : kZBytes;          // dmap_gt is float32, gD_up is TZ
    const uint32_t stage_tx = 2 * kZBytes + (has2 ? kSlotBytes : 0) + (has3 ? slot3_bytes : 0) +
                              (has4 ? kSlotBytes : 0) + (has5 ? kZBytes : 0);

    // producer (thread 0): stream item `it` into stage `s`; fetch L, m when the sample changes
    auto issue = [&](long long it, int b, int s, int prev_b, int lm_buf) {
        const size_t off = static_cast<size_t>(it) * kMap;
        float* st = stage_base + s * kSlots * kMap;
        const bool new_lm = !BOTH && (b != prev_b);
        mbar_expect_tx(&full[s], stage_tx + (new_lm ? 2 * kSlotBytes : 0));
        bulk_g2s(st, static_cast<const TZ*>(a.z) + off, kZBytes, &full[s]);
        bulk_g2s(st + kMap, static_cast<const TZ*>(a.D) + off, kZBytes, &full[s]);
        if (has2) bulk_g2s(st + 2 * kMap, a.slot2 + off, kSlotBytes, &full[s]);
        if (has3) {
            if (tg) bulk_g2s(st + 3 * kMap, static_cast<const float*>(a.slot3) + off, kSlotBytes, &full[s]);
            else    bulk_g2s(st + 3 * kMap, static_cast<const TZ*>(a.slot3) + off, kZBytes, &full[s]);
        }
        if (has4) bulk_g2s(st + 4 * kMap, a.slot4 + off, kSlotBytes, &full[s]);
        if (has5) bulk_g2s(st + 5 * kMap, static_cast<const TZ*>(a.slot5) + off, kZBytes, &full[s]);
        if (new_lm) {
            float* lm = lm_base + lm_buf * 2 * kMap;
            bulk_g2s(lm, a.L + static_cast<size_t>(b) * kMap, kSlotBytes, &full[s]);
            bulk_g2s(lm + kMap, a.m + static_cast<size_t>(b) * kMap, kSlotBytes, &full[s]);
        }
    };

    auto scalar_word = [&](long long it, int wd) -> uint32_t {
        const size_t bj = static_cast<size_t>(it);
        if (wd < 4) return __float_as_uint(a.stats[bj * 4 + wd]);
        if (wd < 7) return a.g_uvd != nullptr ? __float_as_uint(a.g_uvd[bj * 3 + wd - 4]) : 0u;
        if (wd < 10) return (LOSS && a.uvd != nullptr) ? __float_as_uint(a.uvd[bj * 3 + wd - 7]) : 0u;
        if (wd < 13) return LOSS ? __float_as_uint(a.uvd_gt[bj * 3 + wd - 10]) : 0u;
        if (wd >= 16 && LOSS == LOSS_SPARSE) return reinterpret_cast<const uint32_t*>(a.taps + bj)[wd - 16];
        return 0u;
    };

    // lm_cur: buffer holding the current item's L/m; toggles whenever the sample changes
    int lm_cur = 0;
    int b_cur = static_cast<int>(first / a.J);
    int j_cur = static_cast<int>(first - static_cast<long long>(b_cur) * a.J);
    if (tid == 0) issue(first, b_cur, 0, -1, 0);
    if (tid < 32) scal[0][tid] = scalar_word(first, tid);
    __syncthreads();

    const float xs = static_cast<float>(static_cast<int>((tid & 15) * 4) - 32);
    const float ys0 = static_cast<float>(static_cast<int>(tid >> 4) - 32);     // row of chunk i: + (kPipeThreads/16)*i

    int k = 0;
    for (long long it = first; it < last; ++it, ++k) {
        const int s = k % kPipeStages;
        const uint32_t parity = (k / kPipeStages) & 1;
        // prefetch the next item into the other stage (its readers finished before the barrier
        // that ended the previous iteration)
        const long long nxt = it + 1;
        const int j_next = (j_cur + 1 == a.J) ? 0 : j_cur + 1;
        const int b_next = (nxt < last && j_next == 0) ? b_cur + 1 : b_cur;
        const int lm_next = (b_next != b_cur) ? (lm_cur ^ 1) : lm_cur;
        if (tid == 0 && nxt < last) issue(nxt, b_next, (k + 1) % kPipeStages, b_cur, lm_next);
        const uint32_t next_word = (tid < 32 && nxt < last) ? scalar_word(nxt, tid) : 0u;   // lands during the compute

        const int bj = static_cast<int>(it);
        const int j = j_cur;
        const float* sc = reinterpret_cast<const float*>(scal[k & 1]);
        const float4 st = *reinterpret_cast<const float4*>(sc);           // (z extremum, 1/sum, den, d)
        const float wj = (METHOD == PWR_METHOD_SOFTMAX) ? a.w[j] : 1.f;
        const float c = wj * kLog2e;
        const float shift = st.x * c, zref = st.x;
        float gu = sc[4], gvv = sc[5], gd = sc[6], lu = 0.f;
        if (LOSS) {
            const float eu = sc[7] - sc[10], ev = sc[8] - sc[11], ed = sc[9] - sc[12];
            gu = fmaf(a.coef.cu, eu, gu); gvv = fmaf(a.coef.cu, ev, gvv); gd = fmaf(a.coef.cu, ed, gd);
            lu = eu * eu + ev * ev + ed * ed;
        }
        const float gu63 = gu * (1.f / 63.f), gv63 = gvv * (1.f / 63.f);
        const float gdd = __fdividef(gd, st.z);
        const float dcoord = st.w;
        constexpr bool sparse = (LOSS == LOSS_SPARSE);
        TapsIdx tp;
        if (sparse) {
            // footprint table of this item from the prefetched taps record (fp's previous readers
            // finished before the barrier that ended the previous iteration)
            tp = taps_index(scal[k & 1] + 16);
            if (tid < kFootprint) fp[tid] = footprint_entry(scal[k & 1] + 16, tid);
            __syncthreads();
        }

        const float* sz = stage_base + s * kSlots * kMap;     // slot bases (each slot is 16 KiB apart)
        const float* sD = sz + kMap;
        const float4* s2 = reinterpret_cast<const float4*>(sz + 2 * kMap);
        const float* s3 = sz + 3 * kMap;
        const float4* sL = reinterpret_cast<const float4*>(lm_base + lm_cur * 2 * kMap);
        const float4* sM = sL + kMap / 4;

        mbar_wait(&full[s], parity);

        // acc: sum gp*p, sum (p-Hgt)^2, sum (D-Dgt)^2, and for dL/dw = sum_px gy*z with gy = p*(gp - S1):
        // T1 = sum p*gp*(z - zref), T2 = sum p*(z - zref)  =>  sum gy*z = T1 - S1*T2 (sum gy = 0 makes the
        // reference point free; zref = the extremum, where p peaks, keeps both sums small), so dL/dw
        // needs no second pass over z and no second block reduction.
        float4 pv[kPipeVec], gv[kPipeVec];
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < kPipeVec; ++i) {
            const int cidx = tid + i * kPipeThreads;
            const float4 z4 = MapIO<TZ>::smem(sz, cidx), d4 = MapIO<TZ>::smem(sD, cidx);
            // BOTH: no shared-memory pair for L, m (the six slots fill it); the J items of a sample share them in L2
            const float4 l4 = BOTH ? ld_keep(a.L + static_cast<size_t>(b_cur) * kMap + cidx * 4) : sL[cidx];
            const float4 m4 = BOTH ? ld_keep(a.m + static_cast<size_t>(b_cur) * kMap + cidx * 4) : sM[cidx];
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            // slot 2 / 3 hold either dense targets (tg) or dense upstream gradients
            const float4 q2 = has2 ? s2[cidx] : zero4;
            const float4 q3 = has3 ? (tg ? MapIO<float>::smem(s3, cidx) : MapIO<TZ>::smem(s3, cidx)) : zero4;
            float4 t2 = tg ? q2 : zero4, t3 = tg ? q3 : zero4;          // targets
            float4 u2 = tg ? zero4 : q2, u3 = tg ? zero4 : q3;          // upstream gradients
            if (BOTH) {
                if (has4) u2 = reinterpret_cast<const float4*>(sz + 4 * kMap)[cidx];
                if (has5) u3 = MapIO<TZ>::smem(sz + 5 * kMap, cidx);
            }
            if (sparse) sparse_lookup(tp, fp, cidx >> 4, (cidx & 15) * 4, l4, m4, t2, t3);
            const bool have_h = sparse || (tg && has2), have_d = sparse || (tg && has3);
            const float gyrow = gv63 * (ys0 + static_cast<float>(kPipeThreads / 16) * i);
            float4 gd4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float p = heat_raw<METHOD>(comp(z4, kk), c, shift) * st.y;
                const float mk = comp(m4, kk), dk = comp(d4, kk);
                const float rec = mk * (dk + comp(l4, kk));
                float gp = fmaf(gu63, xs + static_cast<float>(kk), gyrow);
                gp = fmaf(gdd * mk, rec - dcoord, gp);
                float gdk = gdd * p * mk * mk;
                if (LOSS) {
                    const float eh = have_h ? p - comp(t2, kk) : 0.f;
                    const float ed = have_d ? dk - comp(t3, kk) : 0.f;
                    gp = fmaf(a.coef.ch, eh, gp);
                    gdk = fmaf(a.coef.cd, ed, gdk);
                    acc[1] = fmaf(eh, eh, acc[1]);
                    acc[2] = fmaf(ed, ed, acc[2]);
                }
                gp += comp(u2, kk);
                gdk += comp(u3, kk);
                const float pg = gp * p;
                acc[0] += pg;
                if (METHOD == PWR_METHOD_SOFTMAX) {
                    const float dz = comp(z4, kk) - zref;
                    acc[3] = fmaf(pg, dz, acc[3]);
                    acc[4] = fmaf(p, dz, acc[4]);
                }
                set_comp(pv[i], kk, p);
                set_comp(gv[i], kk, gp);
                set_comp(gd4, kk, gdk);
            }
            if (a.gD != nullptr) MapIO<TZ>::st(a.gD, static_cast<size_t>(bj) * kMap + cidx * 4, gd4);
        }
        if (METHOD != PWR_METHOD_GIVEN || LOSS) pipe_block_sum<5>(acc, scratch[k & 1]);

        const float s1 = acc[0];
        if (a.gz != nullptr) {
#pragma unroll
            for (int i = 0; i < kPipeVec; ++i) {
                const int cidx = tid + i * kPipeThreads;
                float4 g4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float g;
                    if (METHOD == PWR_METHOD_SOFTMAX) {
                        g = wj * (comp(pv[i], kk) * (comp(gv[i], kk) - s1));       // w * dL/d(w z)
                    } else if (METHOD == PWR_METHOD_SUM) {
                        // p > eps/S  <=>  z > 0: through relu and 1/sum
                        g = comp(MapIO<TZ>::smem(sz, cidx), kk) > 0.f ? (comp(gv[i], kk) - s1) * st.y : 0.f;
                    } else {
                        g = comp(gv[i], kk);
                    }
                    set_comp(g4, kk, g);
                }
                MapIO<TZ>::st(a.gz, static_cast<size_t>(bj) * kMap + cidx * 4, g4);
            }
        }
        if (METHOD == PWR_METHOD_SOFTMAX && a.gw_partial != nullptr && tid == 0)
            a.gw_partial[bj] = acc[3] - s1 * acc[4];
        if (LOSS && a.loss_partial != nullptr && tid == 0) {
            a.loss_partial[bj * 3 + 0] = acc[1];
            a.loss_partial[bj * 3 + 1] = acc[2];
            a.loss_partial[bj * 3 + 2] = lu;
        }
        if (tid < 32 && nxt < last) scal[(k + 1) & 1][tid] = next_word;
        // every thread is done with stage s and (if the sample changes) with the old L/m buffer
        __syncthreads();
        b_cur = b_next;
        j_cur = j_next;
        lm_cur = lm_next;
    }
}

// ---------------------------------------------------------------------------
// small helpers: batch reduction of per-(b,j) partials, scaling, recover_uvd
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
reduce_partials_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int J, int C) {
    __shared__ float scratch[kWarps];
    const int jc = blockIdx.x;               // j*C + c
    const int j = jc / C, c = jc - j * C;
    float v[1] = {0.f};
    for (int b = threadIdx.x; b < B; b += kThreads) v[0] += in[(static_cast<size_t>(b) * J + j) * C + c];
    block_sum<1>(v, scratch);
    if (threadIdx.x == 0) out[jc] = v[0];
}

// train.py:197-205 from the per-(b,j) sums of squares: one CTA, fixed order.
constexpr int kLossThreads = 1024;
__global__ void __launch_bounds__(kLossThreads)
stage_loss_kernel(const float* __restrict__ loss_partial, int n_items, float scale_h, float scale_d, float scale_u,
                  float alpha, float* __restrict__ out, const float* __restrict__ gw_partial, float* __restrict__ gw_out,
                  int B, int J) {
    __shared__ float scratch[(kLossThreads / 32) * 3];
    if (blockIdx.x > 0) {
        // CTAs 1..J: dL/dw[j] = sum over the batch of gw_partial[b, j] (the launch pwr_reduce_partials would be)
        const int j = blockIdx.x - 1;
        float g = 0.f;
        for (int b = threadIdx.x; b < B; b += kLossThreads) g += gw_partial[static_cast<size_t>(b) * J + j];
        g = warp_sum(g);
        if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = g;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int wv = 0; wv < kLossThreads / 32; ++wv) t += scratch[wv];
            gw_out[j] = t;
        }
        return;
    }
    float v[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < n_items; i += kLossThreads) {
        v[0] += loss_partial[i * 3 + 0];
        v[1] += loss_partial[i * 3 + 1];
        v[2] += loss_partial[i * 3 + 2];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) { scratch[warp * 3 + 0] = v[0]; scratch[warp * 3 + 1] = v[1]; scratch[warp * 3 + 2] = v[2]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s[3] = {0.f, 0.f, 0.f};
        for (int wv = 0; wv < kLossThreads / 32; ++wv) { s[0] += scratch[wv * 3]; s[1] += scratch[wv * 3 + 1]; s[2] += scratch[wv * 3 + 2]; }
        const float lh = s[0] * scale_h, ld = s[1] * scale_d, lu = s[2] * scale_u;
        out[0] = lh; out[1] = ld; out[2] = lu;
        out[3] = alpha * lu + (1.f - alpha) * (lh + ld);
    }
}

template <typename TZ>
__global__ void __launch_bounds__(kThreads)
scale_inplace_kernel(void* __restrict__ x, void* __restrict__ x2, const float* __restrict__ scale, long long n4,
                     float* __restrict__ small, int n_small) {
    const float s = *scale;
    if (s == 1.0f) return;
    if (small != nullptr && blockIdx.x == 0 && threadIdx.x < n_small) small[threadIdx.x] *= s;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < n4; i += stride) {
        float4 v = MapIO<TZ>::ld(x, static_cast<size_t>(i) * 4);
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        MapIO<TZ>::st(x, static_cast<size_t>(i) * 4, v);
        if (x2 != nullptr) {
            float4 u = MapIO<TZ>::ld(x2, static_cast<size_t>(i) * 4);
            u.x *= s; u.y *= s; u.z *= s; u.w *= s;
            MapIO<TZ>::st(x2, static_cast<size_t>(i) * 4, u);
        }
    }
}

// utils.py:332-337 (float32 torch arithmetic) + datasets.py:100-111
__global__ void recover_uvd_kernel(const float* __restrict__ uvd_norm, const float* __restrict__ box,
                                   const float* __restrict__ cube, const float* __restrict__ com,
                                   float fx, float fy, float halfu, float halfv,
                                   float* __restrict__ uvd_px, float* __restrict__ xyz, int B, int J) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * J) return;
    const int b = i / J;
    const float s = __fsub_rn(box[b], 1.f);
    const float u = __fadd_rn(__fmul_rn(uvd_norm[i * 3 + 0], s), com[b * 3 + 0]);
    const float v = __fadd_rn(__fmul_rn(uvd_norm[i * 3 + 1], s), com[b * 3 + 1]);
    const float d = __fadd_rn(__fmul_rn(uvd_norm[i * 3 + 2], cube[b]), com[b * 3 + 2]);
    if (uvd_px != nullptr) { uvd_px[i * 3 + 0] = u; uvd_px[i * 3 + 1] = v; uvd_px[i * 3 + 2] = d; }
    if (xyz != nullptr) {
        xyz[i * 3 + 0] = __fmul_rn(__fdiv_rn(__fsub_rn(u, halfu), fx), d);
        xyz[i * 3 + 1] = __fmul_rn(__fdiv_rn(__fsub_rn(v, halfv), fy), d);
        xyz[i * 3 + 2] = d;
    }
}

// train.py:254-276 / test.py:106-113: mean over joints of |xyz_pred - xyz_true| per sample, both
// sides through recover_uvd (float32 torch order) and uvd2xyz (float32 NumPy order)
__global__ void joint_error_kernel(const float* __restrict__ uvd_pred, const float* __restrict__ uvd_true,
                                   const float* __restrict__ box, const float* __restrict__ cube,
                                   const float* __restrict__ com, float fx, float fy, float halfu, float halfv,
                                   float* __restrict__ err, int B, int J) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float s = __fsub_rn(box[b], 1.f), cb = cube[b];
    const float cu = com[b * 3 + 0], cv = com[b * 3 + 1], cz = com[b * 3 + 2];
    float acc = 0.f;
    for (int j = 0; j < J; ++j) {
        float xyz[2][3];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float* p = (k == 0 ? uvd_pred : uvd_true) + (static_cast<size_t>(b) * J + j) * 3;
            const float u = __fadd_rn(__fmul_rn(p[0], s), cu);
            const float v = __fadd_rn(__fmul_rn(p[1], s), cv);
            const float d = __fadd_rn(__fmul_rn(p[2], cb), cz);
            xyz[k][0] = __fmul_rn(__fdiv_rn(__fsub_rn(u, halfu), fx), d);
            xyz[k][1] = __fmul_rn(__fdiv_rn(__fsub_rn(v, halfv), fy), d);
            xyz[k][2] = d;
        }
        const float dx = __fsub_rn(xyz[0][0], xyz[1][0]), dy = __fsub_rn(xyz[0][1], xyz[1][1]);
        const float dz = __fsub_rn(xyz[0][2], xyz[1][2]);
        acc = __fadd_rn(acc, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
    }
    err[b] = __fdiv_rn(acc, static_cast<float>(J));
}

static int check_bj(int B, int J) {
    if (B < 0 || J < 1 || J > PWR_MAX_JOINTS) return PWR_E_SHAPE;
    if (static_cast<long long>(B) * J > 0x7fffffffLL / 4) return PWR_E_SHAPE;
    return 0;
}
// ---------------------------------------------------------------------------
// last stage in ONE pass: forward + stage loss + backward (pwr_decoder_fwd_bwd_loss)
// ---------------------------------------------------------------------------
// train.py:192-207 for the last stage: nothing consumes its heat maps but the loss, so forward and
// backward can share one visit of (z, D, heat_gt, dmap_gt): the logits are read once instead of
// twice and `stats` never travels.  Per sample (7J + 2) maps move instead of (3J + 2) + (6J + 2)
// (H stored; 6J + 2 without).
// An item is a chain of three block reductions (extremum, forward sums, backward sums); one lock-step
// 512-thread CTA per SM spent 6 750 cycles per item on it and LOST to the two-kernel route (1.36 vs
// 1.32 ms, r1).  Hence two independent 256-thread CTAs per SM, each with its own ring of six 16 KB
// slots (96 KB): the 2 or 4 maps of an item (z, D[, heat_gt, dmap_gt]) are bulk-TMA loads in one
// FIFO, up to six in flight or resident, so while an item is in its backward pass the logits of the
// next one (and with compact targets of the next two) are already landing.  L, m live in registers
// across the J items of a sample; block sums cost 8 shuffles per warp.
#ifndef PWR_FUSED_THREADS
#define PWR_FUSED_THREADS 256
#endif
constexpr int kFusedThreads = PWR_FUSED_THREADS;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedVec = kMap / 4 / kFusedThreads;
constexpr int kFusedSlots = 6;
constexpr int kFusedCtasPerSm = 2;
constexpr int kFusedSmemBytes = kFusedSlots * kSlotBytes + 64;
static_assert(kFusedWarps % 4 == 0, "per-warp partials are read as float4");
static_assert(kFusedVec * 4 <= 32, "one z > 0 bit per pixel of a thread");

struct FusedArgs {
    const void* z; const float* w; const void* D; const float* L; const float* m;
    const float* heat_gt; const float* dmap_gt; const float* uvd_gt; const pwr_joint_taps* taps;
    LossCoef coef;
    float* H; float* uvd; void* gz; void* gD; float* gw_partial; float* loss_partial;
    float* stats;              // [items,4] for a later backward (inner stage: forward + loss only), or NULL
    int J; int items;
};

template <int METHOD, int LOSS, typename TZ>
__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm)
decoder_fused_kernel(FusedArgs a) {
    static_assert(LOSS != LOSS_NONE && METHOD != PWR_METHOD_GIVEN, "last stage with a loss");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slots = reinterpret_cast<float*>(smem_raw);                                    // [6][4096]
    uint64_t* full = reinterpret_cast<uint64_t*>(slots + kFusedSlots * kMap);             // [6]
    __shared__ __align__(16) float scr_e[2][kFusedWarps];
    __shared__ __align__(16) float scr_f[2][5 * kFusedWarps];     // forward sums
    __shared__ __align__(16) float scr_b[2][5 * kFusedWarps];     // backward sums
    __shared__ __align__(16) uint32_t scal[2][32];                // 10-12 uvd_gt, 16-31 taps (one item ahead)
    __shared__ float fp[LOSS == LOSS_SPARSE ? kFootprint : 1];
    constexpr bool sparse = (LOSS == LOSS_SPARSE);

    const int tid = threadIdx.x;
    const long long first = static_cast<long long>(a.items) * blockIdx.x / gridDim.x;
    const long long last = static_cast<long long>(a.items) * (blockIdx.x + 1) / gridDim.x;
    if (first >= last) return;
    if (a.coef.scale_dev != nullptr) {
        const float up = *a.coef.scale_dev;
        a.coef.cu *= up; a.coef.ch *= up; a.coef.cd *= up;
    }
    if (tid == 0) {
        for (int s = 0; s < kFusedSlots; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool maps = !sparse && a.heat_gt != nullptr;        // dense target maps are visited (else: uvd term only)
    const int lpi = maps ? 4 : 2;                             // loads per item: z, D[, heat_gt, dmap_gt]
    constexpr uint32_t kZBytes = kMap * sizeof(TZ);
    auto scalar_word = [&](long long it, int wd) -> uint32_t {
        const size_t bj = static_cast<size_t>(it);
        if (wd >= 10 && wd < 13) return __float_as_uint(a.uvd_gt[bj * 3 + wd - 10]);
        if (wd >= 16 && sparse) return reinterpret_cast<const uint32_t*>(a.taps + bj)[wd - 16];
        return 0u;
    };
    if (tid < 32) scal[0][tid] = scalar_word(first, tid);
    __syncthreads();

    // producer (thread 0): one FIFO of loads, load n -> slot n % 6
    long long p_it = first;
    int p_q = 0, n_issued = 0;
    auto fill = [&](int limit) {
        while (n_issued < limit && p_it < last) {
            const int sl = n_issued % kFusedSlots;
            float* dst = slots + sl * kMap;
            const size_t off = static_cast<size_t>(p_it) * kMap;
            if (p_q < 2) {
                mbar_expect_tx(&full[sl], kZBytes);
                bulk_g2s(dst, static_cast<const TZ*>(p_q == 0 ? a.z : a.D) + off, kZBytes, &full[sl]);
            } else {
                mbar_expect_tx(&full[sl], kSlotBytes);
                bulk_g2s(dst, (p_q == 2 ? a.heat_gt : a.dmap_gt) + off, kSlotBytes, &full[sl]);
            }
            ++n_issued;
            if (++p_q == lpi) { p_q = 0; ++p_it; }
        }
    };
    if (tid == 0) fill(kFusedSlots);

    int b_cur = static_cast<int>(first / a.J);
    int j_cur = static_cast<int>(first - static_cast<long long>(b_cur) * a.J);
    const float xs = static_cast<float>(static_cast<int>((tid & 15) * 4) - 32);
    const float ys0 = static_cast<float>(static_cast<int>(tid >> 4) - 32);     // row of chunk i: + (kFusedThreads/16)*i
    float4 lv[kFusedVec], mv[kFusedVec];
    int b_loaded = -1;
    float w_next = (METHOD == PWR_METHOD_SOFTMAX) ? a.w[j_cur] : 1.f;
    int k = 0;
    for (long long it = first; it < last; ++it, ++k) {
        const float wj = w_next;
        const float c = wj * kLog2e;
        if (b_cur != b_loaded) {                      // label and mask stay in registers for the J items of a sample
            b_loaded = b_cur;
            const size_t offb = static_cast<size_t>(b_cur) * kMap + tid * 4;
#pragma unroll
            for (int i = 0; i < kFusedVec; ++i) mv[i] = ld_keep(a.m + offb + i * (kFusedThreads * 4));
#pragma unroll
            for (int i = 0; i < kFusedVec; ++i) lv[i] = ld_keep(a.L + offb + i * (kFusedThreads * 4));
        }
        const uint32_t next_word = (tid < 32 && it + 1 < last) ? scalar_word(it + 1, tid) : 0u;
        const float* sc = reinterpret_cast<const float*>(scal[k & 1]);
        TapsIdx tp;
        if (sparse) {
            tp = taps_index(scal[k & 1] + 16);
            if (tid < kFootprint) fp[tid] = footprint_entry(scal[k & 1] + 16, tid);
        }
        // slots and barrier phases of this item's loads
        const int n0 = k * lpi;
        const int sl_z = n0 % kFusedSlots, sl_d = (n0 + 1) % kFusedSlots;
        const float* sz = slots + sl_z * kMap;
        const float* sD = slots + sl_d * kMap;

        // ---- forward: extremum, un-normalised heat, five sums (decoder_fwd_kernel's arithmetic) ----
        mbar_wait(&full[sl_z], (n0 / kFusedSlots) & 1);
        float4 zv[kFusedVec], pv[kFusedVec];
#pragma unroll
        for (int i = 0; i < kFusedVec; ++i) zv[i] = MapIO<TZ>::smem(sz, tid + i * kFusedThreads);
        float shift = 0.f, zext = 0.f;
        if (METHOD == PWR_METHOD_SOFTMAX) {
            const bool want_max = c >= 0.f;
            float e = want_max ? -INFINITY : INFINITY;
#pragma unroll
            for (int i = 0; i < kFusedVec; ++i) {
                if (want_max) e = fmaxf(fmaxf(fmaxf(e, zv[i].x), fmaxf(zv[i].y, zv[i].z)), zv[i].w);
                else          e = fminf(fminf(fminf(e, zv[i].x), fminf(zv[i].y, zv[i].z)), zv[i].w);
            }
            e = want_max ? warp_max(e) : warp_min(e);
            float* se = scr_e[k & 1];
            if ((tid & 31) == 0) se[tid >> 5] = e;
            __syncthreads();                                             // (also publishes fp)
            zext = se[0];
#pragma unroll
            for (int wv = 1; wv < kFusedWarps; ++wv) zext = want_max ? fmaxf(zext, se[wv]) : fminf(zext, se[wv]);
            shift = zext * c;
        } else if (sparse) {
            __syncthreads();                                             // fp visible
        }
        mbar_wait(&full[sl_d], ((n0 + 1) / kFusedSlots) & 1);
        float accf[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum e, e*(x-32), e*(y-32), e*m, e*m*m*(D+L)
#pragma unroll
        for (int i = 0; i < kFusedVec; ++i) {
            const float4 d4 = MapIO<TZ>::smem(sD, tid + i * kFusedThreads);
            const float ys = ys0 + static_cast<float>(kFusedThreads / 16) * i;
            float rowsum = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float e = heat_raw<METHOD>(comp(zv[i], kk), c, shift);
                set_comp(pv[i], kk, e);
                const float mk = comp(mv[i], kk);
                const float em = e * mk;
                rowsum += e;
                accf[1] = fmaf(e, xs + static_cast<float>(kk), accf[1]);
                accf[3] += em;
                accf[4] = fmaf(em, mk * (comp(d4, kk) + comp(lv[i], kk)), accf[4]);
            }
            accf[0] += rowsum;
            accf[2] = fmaf(rowsum, ys, accf[2]);
        }
        float* sf = scr_f[k & 1];
        store_scattered5<kFusedWarps>(warp_sum5_scattered(accf), sf);
        __syncthreads();
        const float inv_s = 1.f / sum_partials<kFusedWarps>(sf);
        const float den = fmaf(sum_partials<kFusedWarps>(sf + 3 * kFusedWarps), inv_s, kEps);
        const float dcoord = (sum_partials<kFusedWarps>(sf + 4 * kFusedWarps) * inv_s) / den;
        const float u = sum_partials<kFusedWarps>(sf + kFusedWarps) * inv_s / 63.f;
        const float v = sum_partials<kFusedWarps>(sf + 2 * kFusedWarps) * inv_s / 63.f;

        // ---- loss on the coordinates: the upstream of the backward ----
        const float eu = u - sc[10], ev = v - sc[11], ed = dcoord - sc[12];
        const float gu = a.coef.cu * eu, gvv = a.coef.cu * ev, gd = a.coef.cu * ed;
        const float lu = eu * eu + ev * ev + ed * ed;
        const float gu63 = gu * (1.f / 63.f), gv63 = gvv * (1.f / 63.f);
        const float gdd = __fdividef(gd, den);

        // ---- backward pass (decoder_bwd_pipe_kernel's arithmetic), heat maps stored on the way ----
        const float4* s2 = reinterpret_cast<const float4*>(slots + ((n0 + 2) % kFusedSlots) * kMap);
        const float4* s3 = reinterpret_cast<const float4*>(slots + ((n0 + 3) % kFusedSlots) * kMap);
        if (maps) {
            mbar_wait(&full[(n0 + 2) % kFusedSlots], ((n0 + 2) / kFusedSlots) & 1);
            mbar_wait(&full[(n0 + 3) % kFusedSlots], ((n0 + 3) / kFusedSlots) & 1);
        }
        float4 gv[kFusedVec];
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum gp*p, sum (p-Hgt)^2, sum (D-Dgt)^2, T1, T2
        unsigned int zpos = 0;
#pragma unroll
        for (int i = 0; i < kFusedVec; ++i) {
            const int cidx = tid + i * kFusedThreads;
            const float4 d4 = MapIO<TZ>::smem(sD, cidx), l4 = lv[i], m4 = mv[i];
            float4 t2 = make_float4(0.f, 0.f, 0.f, 0.f), t3 = t2;
            if (sparse) sparse_lookup(tp, fp, cidx >> 4, (cidx & 15) * 4, l4, m4, t2, t3);
            else if (maps) { t2 = s2[cidx]; t3 = s3[cidx]; }
            const bool have_t = sparse || maps;
            const float gyrow = gv63 * (ys0 + static_cast<float>(kFusedThreads / 16) * i);
            float4 gd4, p4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float zk = comp(zv[i], kk);
                const float p = comp(pv[i], kk) * inv_s;
                const float mk = comp(m4, kk), dk = comp(d4, kk);
                const float rec = mk * (dk + comp(l4, kk));
                float gp = fmaf(gu63, xs + static_cast<float>(kk), gyrow);
                gp = fmaf(gdd * mk, rec - dcoord, gp);
                float gdk = gdd * p * mk * mk;
                if (have_t) {
                    const float eh = p - comp(t2, kk), edm = dk - comp(t3, kk);
                    gp = fmaf(a.coef.ch, eh, gp);
                    gdk = fmaf(a.coef.cd, edm, gdk);
                    acc[1] = fmaf(eh, eh, acc[1]);
                    acc[2] = fmaf(edm, edm, acc[2]);
                }
                const float pg = gp * p;
                acc[0] += pg;
                if (METHOD == PWR_METHOD_SOFTMAX) {
                    const float dz = zk - zext;
                    acc[3] = fmaf(pg, dz, acc[3]);
                    acc[4] = fmaf(p, dz, acc[4]);
                }
                if (METHOD == PWR_METHOD_SUM && zk > 0.f) zpos |= 1u << (4 * i + kk);
                set_comp(p4, kk, p);
                set_comp(gv[i], kk, gp);
                set_comp(gd4, kk, gdk);
            }
            pv[i] = p4;
            if (a.H != nullptr) st_stream(a.H + static_cast<size_t>(it) * kMap + cidx * 4, p4);
            if (a.gD != nullptr) MapIO<TZ>::st(a.gD, static_cast<size_t>(it) * kMap + cidx * 4, gd4);
        }
        // next item's scalars: parked before the barrier below, read after it
        if (tid < 32 && it + 1 < last) scal[(k + 1) & 1][tid] = next_word;
        float* sb = scr_b[k & 1];
        store_scattered5<kFusedWarps>(warp_sum5_scattered(acc), sb);
        __syncthreads();
        // every thread is past its shared-memory reads of this item: its slots go back to the FIFO
        if (tid == 0) fill((k + 1) * lpi + kFusedSlots);
        if (++j_cur == a.J) { j_cur = 0; ++b_cur; }
        if (METHOD == PWR_METHOD_SOFTMAX && it + 1 < last) w_next = a.w[j_cur];

        const float s1 = sum_partials<kFusedWarps>(sb);
        if (a.gz != nullptr) {
#pragma unroll
            for (int i = 0; i < kFusedVec; ++i) {
                float4 g4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float g;
                    if (METHOD == PWR_METHOD_SOFTMAX) g = wj * (comp(pv[i], kk) * (comp(gv[i], kk) - s1));
                    else g = ((zpos >> (4 * i + kk)) & 1u) ? (comp(gv[i], kk) - s1) * inv_s : 0.f;
                    set_comp(g4, kk, g);
                }
                MapIO<TZ>::st(a.gz, static_cast<size_t>(it) * kMap + (tid + i * kFusedThreads) * 4, g4);
            }
        }
        if (tid == 0) {
            float* o = a.uvd + static_cast<size_t>(it) * 3;
            o[0] = u; o[1] = v; o[2] = dcoord;
            if (a.stats != nullptr) reinterpret_cast<float4*>(a.stats)[it] = make_float4(zext, inv_s, den, dcoord);
            if (METHOD == PWR_METHOD_SOFTMAX && a.gw_partial != nullptr)
                a.gw_partial[it] = sum_partials<kFusedWarps>(sb + 3 * kFusedWarps) - s1 * sum_partials<kFusedWarps>(sb + 4 * kFusedWarps);
            if (a.loss_partial != nullptr) {
                a.loss_partial[it * 3 + 0] = sum_partials<kFusedWarps>(sb + kFusedWarps);
                a.loss_partial[it * 3 + 1] = sum_partials<kFusedWarps>(sb + 2 * kFusedWarps);
                a.loss_partial[it * 3 + 2] = lu;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// one-pass last stage, lean variant: no dense target map is visited (compact targets, or uvd-only loss)
// ---------------------------------------------------------------------------
// With compact targets an item moves only 32 KB in and up to 48 KB out, and decoder_fused_kernel stops being
// bandwidth-bound: ncu r2 shows 632 M warp instructions issued at 2.1 per clock per SM from 16 resident warps,
// barrier + short-scoreboard stalls on top, DRAM at 56 % (1.05 ms where the bytes need 0.74).  Its 118-128
// registers (z, e, gp, L, m of 16 pixels per thread = 80 of them) pin it at two CTAs per SM.  Here the per-pixel
// state between the phases lives in the ring instead of in registers: the forward pass only accumulates the
// sums; the backward pass re-reads z, D from the item's slots, recomputes e (one more ex2 per pixel on an idle
// XU pipe) and writes gp over z and p over D IN PLACE (every thread rewrites exactly the 16-byte chunks it has
// just read); the dL/dz pass after the block sums reads them back.  L, m stay register-resident across the J
// items of a sample.  That fits 80 registers -> three CTAs (24 warps) per SM, each with a ring of two (z, D)
// pairs.  The slots of item k are therefore busy until its last pass; they are handed back to the producer
// after the first block barrier of item k + 1 (every thread has executed fence.proxy.async after its last
// generic access to them).  float32 maps only: half-precision logits are 8 KB per slot and cannot take the
// float32 gp / p of the same pixels in place; those keep decoder_fused_kernel.
#ifndef PWR_FL_THREADS
#define PWR_FL_THREADS 256
#endif
#ifndef PWR_FL_CTAS
#define PWR_FL_CTAS 3
#endif
#ifndef PWR_FL_PAIRS
#define PWR_FL_PAIRS 2
#endif
#ifndef PWR_FL_FOLD
#define PWR_FL_FOLD 0          // 1: extremum of item k + 1 rides on the backward barrier of item k (needs >= 3 pairs)
#endif
constexpr int kFLThreads = PWR_FL_THREADS;
constexpr int kFLWarps = kFLThreads / 32;
constexpr int kFLVec = kMap / 4 / kFLThreads;
constexpr int kFLPairs = PWR_FL_PAIRS;
constexpr int kFLCtasPerSm = PWR_FL_CTAS;
constexpr bool kFLFold = PWR_FL_FOLD != 0;
constexpr int kFLSmemBytes = kFLPairs * 2 * kSlotBytes + 64;
static_assert(kFLWarps % 4 == 0 && kFLThreads >= 128, "per-warp partials are read as float4; 121 footprint threads");
static_assert(kFLVec * 4 <= 32 && kFLVec >= 1, "one z > 0 bit per pixel of a thread");
static_assert(kFLCtasPerSm * (kFLSmemBytes + 3072) <= 227 * 1024, "shared memory of the resident CTAs");
static_assert(!kFLFold || kFLPairs >= 3, "the folded extremum reads item k + 1 while item k + 2 is in flight");

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int METHOD, int LOSS>
__global__ void __launch_bounds__(kFLThreads, kFLCtasPerSm)
decoder_fused_lean_kernel(FusedArgs a) {
    static_assert(LOSS != LOSS_NONE && METHOD != PWR_METHOD_GIVEN, "last stage with a loss");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slots = reinterpret_cast<float*>(smem_raw);                                    // [pairs][2][4096]
    uint64_t* full = reinterpret_cast<uint64_t*>(slots + kFLPairs * 2 * kMap);            // [pairs][2]
    __shared__ __align__(16) float scr_e[2][kFLWarps];
    __shared__ __align__(16) float scr_f[2][5 * kFLWarps];        // forward sums
    __shared__ __align__(16) float scr_b[2][5 * kFLWarps];        // backward sums
    __shared__ __align__(16) uint32_t scal[2][32];                // 10-12 uvd_gt, 16-31 taps (one item ahead)
    __shared__ float fp[LOSS == LOSS_SPARSE ? kFootprint : 1];
    constexpr bool sparse = (LOSS == LOSS_SPARSE);

    const int tid = threadIdx.x;
    const long long first = static_cast<long long>(a.items) * blockIdx.x / gridDim.x;
    const long long last = static_cast<long long>(a.items) * (blockIdx.x + 1) / gridDim.x;
    if (first >= last) return;
    if (a.coef.scale_dev != nullptr) {
        const float up = *a.coef.scale_dev;
        a.coef.cu *= up; a.coef.ch *= up; a.coef.cd *= up;
    }
    if (tid == 0) {
        for (int s = 0; s < kFLPairs * 2; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto scalar_word = [&](long long it, int wd) -> uint32_t {
        const size_t bj = static_cast<size_t>(it);
        if (wd >= 10 && wd < 13) return __float_as_uint(a.uvd_gt[bj * 3 + wd - 10]);
        if (wd >= 16 && sparse) return reinterpret_cast<const uint32_t*>(a.taps + bj)[wd - 16];
        return 0u;
    };
    if (tid < 32) scal[0][tid] = scalar_word(first, tid);
    __syncthreads();

    // producer (thread 0): item n of this CTA -> pair n % kFLPairs, one barrier per map
    auto issue = [&](long long it, int n) {
        if (it >= last) return;
        const int pr = n % kFLPairs;
        float* dst = slots + pr * 2 * kMap;
        const size_t off = static_cast<size_t>(it) * kMap;
        mbar_expect_tx(&full[2 * pr], kSlotBytes);
        bulk_g2s(dst, static_cast<const float*>(a.z) + off, kSlotBytes, &full[2 * pr]);
        mbar_expect_tx(&full[2 * pr + 1], kSlotBytes);
        bulk_g2s(dst + kMap, static_cast<const float*>(a.D) + off, kSlotBytes, &full[2 * pr + 1]);
    };
    if (tid == 0) {
        for (int n = 0; n < kFLPairs; ++n) issue(first + n, n);
    }

    int b_cur = static_cast<int>(first / a.J);
    int j_cur = static_cast<int>(first - static_cast<long long>(b_cur) * a.J);
    const float xs = static_cast<float>(static_cast<int>((tid & 15) * 4) - 32);
    const float ys0 = static_cast<float>(static_cast<int>(tid >> 4) - 32);     // row of chunk i: + (kFLThreads/16)*i
    float4 lv[kFLVec], mv[kFLVec];
    int b_loaded = -1;
    float w_next = (METHOD == PWR_METHOD_SOFTMAX) ? a.w[j_cur] : 1.f;

    // block extremum of the logits in slot `sz` in the direction of sign(c): per-warp results -> se[warp]
    auto warp_extremum = [&](const float* sz, bool want_max, float* se) {
        float e = want_max ? -INFINITY : INFINITY;
#pragma unroll
        for (int i = 0; i < kFLVec; ++i) {
            const float4 z4 = reinterpret_cast<const float4*>(sz)[tid + i * kFLThreads];
            if (want_max) e = fmaxf(fmaxf(fmaxf(e, z4.x), fmaxf(z4.y, z4.z)), z4.w);
            else          e = fminf(fminf(fminf(e, z4.x), fminf(z4.y, z4.z)), z4.w);
        }
        e = want_max ? warp_max(e) : warp_min(e);
        if ((tid & 31) == 0) se[tid >> 5] = e;
    };
    if (kFLFold && METHOD == PWR_METHOD_SOFTMAX) {       // first item: nobody computed its extremum yet
        mbar_wait(&full[0], 0);
        warp_extremum(slots, w_next * kLog2e >= 0.f, scr_e[0]);
        __syncthreads();
    }

    int k = 0;
    for (long long it = first; it < last; ++it, ++k) {
        const float wj = w_next;
        const float c = wj * kLog2e;
        if (b_cur != b_loaded) {                      // label and mask stay in registers for the J items of a sample
            b_loaded = b_cur;
            const size_t offb = static_cast<size_t>(b_cur) * kMap + tid * 4;
#pragma unroll
            for (int i = 0; i < kFLVec; ++i) mv[i] = ld_keep(a.m + offb + i * (kFLThreads * 4));
#pragma unroll
            for (int i = 0; i < kFLVec; ++i) lv[i] = ld_keep(a.L + offb + i * (kFLThreads * 4));
        }
        int j_next = j_cur + 1, b_next = b_cur;
        if (j_next == a.J) { j_next = 0; ++b_next; }
        if (METHOD == PWR_METHOD_SOFTMAX && it + 1 < last) w_next = a.w[j_next];
        const uint32_t next_word = (tid < 32 && it + 1 < last) ? scalar_word(it + 1, tid) : 0u;
        const float* sc = reinterpret_cast<const float*>(scal[k & 1]);
        TapsIdx tp;
        if (sparse) {
            tp = taps_index(scal[k & 1] + 16);
            if (tid < kFootprint) fp[tid] = footprint_entry(scal[k & 1] + 16, tid);   // published by the barriers below
        }
        const int pr = k % kFLPairs;
        const uint32_t ph = (k / kFLPairs) & 1;
        float* sz = slots + pr * 2 * kMap;
        float* sD = sz + kMap;
        bool handed_back = false;                     // the slots of item k - 1 go back after this item's first barrier
        auto hand_back = [&]() {
            if (!handed_back && tid == 0 && k > 0) issue(it - 1 + kFLPairs, k - 1 + kFLPairs);
            handed_back = true;
        };

        // ---- forward: extremum, five sums (decoder_fwd_kernel's arithmetic); nothing per pixel is kept ----
        mbar_wait(&full[2 * pr], ph);
        float shift = 0.f, zext = 0.f;
        if (METHOD == PWR_METHOD_SOFTMAX) {
            const bool want_max = c >= 0.f;
            float* se = scr_e[k & 1];
            if (!kFLFold) {
                warp_extremum(sz, want_max, se);
                __syncthreads();
                hand_back();
            }
            zext = se[0];
#pragma unroll
            for (int wv = 1; wv < kFLWarps; ++wv) zext = want_max ? fmaxf(zext, se[wv]) : fminf(zext, se[wv]);
            shift = zext * c;
        }
        mbar_wait(&full[2 * pr + 1], ph);
        float accf[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum e, e*(x-32), e*(y-32), e*m, e*m*m*(D+L)
#pragma unroll
        for (int i = 0; i < kFLVec; ++i) {
            const float4 z4 = reinterpret_cast<const float4*>(sz)[tid + i * kFLThreads];
            const float4 d4 = reinterpret_cast<const float4*>(sD)[tid + i * kFLThreads];
            const float ys = ys0 + static_cast<float>(kFLThreads / 16) * i;
            float rowsum = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float e = heat_raw<METHOD>(comp(z4, kk), c, shift);
                const float mk = comp(mv[i], kk);
                const float em = e * mk;
                rowsum += e;
                accf[1] = fmaf(e, xs + static_cast<float>(kk), accf[1]);
                accf[3] += em;
                accf[4] = fmaf(em, mk * (comp(d4, kk) + comp(lv[i], kk)), accf[4]);
            }
            accf[0] += rowsum;
            accf[2] = fmaf(rowsum, ys, accf[2]);
        }
        float* sf = scr_f[k & 1];
        store_scattered5<kFLWarps>(warp_sum5_scattered(accf), sf);
        __syncthreads();
        hand_back();
        const float inv_s = 1.f / sum_partials<kFLWarps>(sf);
        const float den = fmaf(sum_partials<kFLWarps>(sf + 3 * kFLWarps), inv_s, kEps);
        const float dcoord = (sum_partials<kFLWarps>(sf + 4 * kFLWarps) * inv_s) / den;
        const float u = sum_partials<kFLWarps>(sf + kFLWarps) * inv_s / 63.f;
        const float v = sum_partials<kFLWarps>(sf + 2 * kFLWarps) * inv_s / 63.f;

        // ---- loss on the coordinates: the upstream of the backward ----
        const float eu = u - sc[10], ev = v - sc[11], ed = dcoord - sc[12];
        const float gu = a.coef.cu * eu, gvv = a.coef.cu * ev, gd = a.coef.cu * ed;
        const float lu = eu * eu + ev * ev + ed * ed;
        const float gu63 = gu * (1.f / 63.f), gv63 = gvv * (1.f / 63.f);
        const float gdd = __fdividef(gd, den);

        // ---- backward pass (decoder_fused_kernel's arithmetic): gp over z, p over D, heat maps stored on the way ----
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // sum gp*p, sum (p-Hgt)^2, sum (D-Dgt)^2, T1, T2
        const bool keep = a.gz != nullptr;
#pragma unroll
        for (int i = 0; i < kFLVec; ++i) {
            const int cidx = tid + i * kFLThreads;
            const float4 z4 = reinterpret_cast<const float4*>(sz)[cidx];
            const float4 d4 = reinterpret_cast<const float4*>(sD)[cidx];
            const float4 l4 = lv[i], m4 = mv[i];
            float4 t2 = make_float4(0.f, 0.f, 0.f, 0.f), t3 = t2;
            if (sparse) sparse_lookup(tp, fp, cidx >> 4, (cidx & 15) * 4, l4, m4, t2, t3);
            const float gyrow = gv63 * (ys0 + static_cast<float>(kFLThreads / 16) * i);
            float4 gd4, p4, gp4, q4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float zk = comp(z4, kk);
                const float p = heat_raw<METHOD>(zk, c, shift) * inv_s;
                const float mk = comp(m4, kk), dk = comp(d4, kk);
                const float rec = mk * (dk + comp(l4, kk));
                float gp = fmaf(gu63, xs + static_cast<float>(kk), gyrow);
                gp = fmaf(gdd * mk, rec - dcoord, gp);
                float gdk = gdd * p * mk * mk;
                if (sparse) {
                    const float eh = p - comp(t2, kk), edm = dk - comp(t3, kk);
                    gp = fmaf(a.coef.ch, eh, gp);
                    gdk = fmaf(a.coef.cd, edm, gdk);
                    acc[1] = fmaf(eh, eh, acc[1]);
                    acc[2] = fmaf(edm, edm, acc[2]);
                }
                const float pg = gp * p;
                acc[0] += pg;
                if (METHOD == PWR_METHOD_SOFTMAX) {
                    const float dz = zk - zext;
                    acc[3] = fmaf(pg, dz, acc[3]);
                    acc[4] = fmaf(p, dz, acc[4]);
                }
                set_comp(p4, kk, p);
                set_comp(gp4, kk, gp);
                set_comp(gd4, kk, gdk);
                // what the dL/dz pass multiplies (gp - s1) with: p (softmax) or [z > 0] / sum (relu-sum)
                set_comp(q4, kk, METHOD == PWR_METHOD_SOFTMAX ? p : (zk > 0.f ? inv_s : 0.f));
            }
            if (a.H != nullptr) st_stream(a.H + static_cast<size_t>(it) * kMap + cidx * 4, p4);
            if (a.gD != nullptr) st_stream(static_cast<float*>(a.gD) + static_cast<size_t>(it) * kMap + cidx * 4, gd4);
            if (keep) {
                reinterpret_cast<float4*>(sz)[cidx] = gp4;
                reinterpret_cast<float4*>(sD)[cidx] = q4;
            }
        }
        // next item's scalars: parked before the barrier below, read after it
        if (tid < 32 && it + 1 < last) scal[(k + 1) & 1][tid] = next_word;
        if (kFLFold && METHOD == PWR_METHOD_SOFTMAX && it + 1 < last) {
            const int prn = (k + 1) % kFLPairs;
            mbar_wait(&full[2 * prn], ((k + 1) / kFLPairs) & 1);
            warp_extremum(slots + prn * 2 * kMap, w_next * kLog2e >= 0.f, scr_e[(k + 1) & 1]);
        }
        float* sb = scr_b[k & 1];
        store_scattered5<kFLWarps>(warp_sum5_scattered(acc), sb);
        __syncthreads();
        j_cur = j_next; b_cur = b_next;

        const float s1 = sum_partials<kFLWarps>(sb);
        if (keep) {
#pragma unroll
            for (int i = 0; i < kFLVec; ++i) {
                const int cidx = tid + i * kFLThreads;
                const float4 gp4 = reinterpret_cast<const float4*>(sz)[cidx];
                const float4 q4 = reinterpret_cast<const float4*>(sD)[cidx];
                float4 g4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float g;
                    if (METHOD == PWR_METHOD_SOFTMAX) g = wj * (comp(q4, kk) * (comp(gp4, kk) - s1));
                    else g = comp(q4, kk) != 0.f ? (comp(gp4, kk) - s1) * comp(q4, kk) : 0.f;
                    set_comp(g4, kk, g);
                }
                st_stream(static_cast<float*>(a.gz) + static_cast<size_t>(it) * kMap + cidx * 4, g4);
            }
            fence_proxy_async_smem();      // the next accesses to these slots are the producer's bulk copies
        }
        if (tid == 0) {
            float* o = a.uvd + static_cast<size_t>(it) * 3;
            o[0] = u; o[1] = v; o[2] = dcoord;
            if (a.stats != nullptr) reinterpret_cast<float4*>(a.stats)[it] = make_float4(zext, inv_s, den, dcoord);
            if (METHOD == PWR_METHOD_SOFTMAX && a.gw_partial != nullptr)
                a.gw_partial[it] = sum_partials<kFLWarps>(sb + 3 * kFLWarps) - s1 * sum_partials<kFLWarps>(sb + 4 * kFLWarps);
            if (a.loss_partial != nullptr) {
                a.loss_partial[it * 3 + 0] = sum_partials<kFLWarps>(sb + kFLWarps);
                a.loss_partial[it * 3 + 1] = sum_partials<kFLWarps>(sb + 2 * kFLWarps);
                a.loss_partial[it * 3 + 2] = lu;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// backward, lean pipelined variant: no dense target / upstream-gradient maps
// ---------------------------------------------------------------------------
// With compact (pwr_joint_taps) targets, or with no map terms at all, an item moves only 32 KB in
// and 32 KB out, and the one-CTA-per-SM kernel above stops being bandwidth-bound: its 512 threads walk
// wait -> footprint table -> pass -> block sums -> store in lock step (1.01 ms vs 0.60 ms of traffic at
// B=4096 J=14, ncu r1).  Same remedy as the pipelined forward: several independent CTAs per SM so the
// phases of different items overlap, label / mask of the sample kept in registers across its J items,
// and the five block sums in 8 shuffles per warp.
#ifndef PWR_LEAN_THREADS
#define PWR_LEAN_THREADS 256
#endif
#ifndef PWR_LEAN_STAGES
#define PWR_LEAN_STAGES 3
#endif
#ifndef PWR_LEAN_CTAS
#define PWR_LEAN_CTAS 2
#endif
constexpr int kLeanThreads = PWR_LEAN_THREADS;
constexpr int kLeanWarps = kLeanThreads / 32;
constexpr int kLeanVec = kMap / 4 / kLeanThreads;
constexpr int kLeanStages = PWR_LEAN_STAGES;
constexpr int kLeanCtasPerSm = PWR_LEAN_CTAS;
constexpr int kLeanSmemBytes = kLeanStages * 2 * kSlotBytes + 64;
static_assert(kLeanThreads == kFwdThreads, "the scattered block sums are laid out for kFwdWarps warps");
static_assert(kLeanVec * 4 <= 32, "PWR_METHOD_SUM keeps one z > 0 bit per pixel of a thread");
static_assert(kLeanCtasPerSm * (kLeanSmemBytes + 2048) <= 227 * 1024, "shared memory of the resident CTAs");

template <int METHOD, int LOSS, typename TZ>
__global__ void __launch_bounds__(kLeanThreads, kLeanCtasPerSm)
decoder_bwd_lean_kernel(PipeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_base = reinterpret_cast<float*>(smem_raw);                               // [stages][2][4096]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + kLeanStages * 2 * kMap);    // [stages]
    __shared__ __align__(16) float scr[2][5 * kLeanWarps];    // double-buffered: one barrier per reduction
    __shared__ __align__(16) uint32_t scal[2][32];            // per-item scalars, fetched one item ahead
    __shared__ float fp[LOSS == LOSS_SPARSE ? kFootprint : 1];

    const int tid = threadIdx.x;
    const long long first = static_cast<long long>(a.items) * blockIdx.x / gridDim.x;
    const long long last = static_cast<long long>(a.items) * (blockIdx.x + 1) / gridDim.x;
    if (first >= last) return;
    if (LOSS && a.coef.scale_dev != nullptr) {
        const float up = *a.coef.scale_dev;
        a.coef.cu *= up; a.coef.ch *= up; a.coef.cd *= up;
    }
    if (tid == 0) {
        for (int s = 0; s < kLeanStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t kZBytes = kMap * sizeof(TZ);
    auto issue = [&](long long it, int kk) {
        float* st = stage_base + (kk % kLeanStages) * 2 * kMap;
        uint64_t* bar = &full[kk % kLeanStages];
        const size_t off = static_cast<size_t>(it) * kMap;
        mbar_expect_tx(bar, 2 * kZBytes);
        bulk_g2s(st, static_cast<const TZ*>(a.z) + off, kZBytes, bar);
        bulk_g2s(st + kMap, static_cast<const TZ*>(a.D) + off, kZBytes, bar);
    };
    // word layout of scal: 0-3 stats, 4-6 g_uvd, 7-9 uvd, 10-12 uvd_gt, 16-31 taps
    auto scalar_word = [&](long long it, int wd) -> uint32_t {
        const size_t bj = static_cast<size_t>(it);
        if (wd < 4) return __float_as_uint(a.stats[bj * 4 + wd]);
        if (wd < 7) return a.g_uvd != nullptr ? __float_as_uint(a.g_uvd[bj * 3 + wd - 4]) : 0u;
        if (wd < 10) return (LOSS && a.uvd != nullptr) ? __float_as_uint(a.uvd[bj * 3 + wd - 7]) : 0u;
        if (wd < 13) return LOSS ? __float_as_uint(a.uvd_gt[bj * 3 + wd - 10]) : 0u;
        if (wd >= 16 && LOSS == LOSS_SPARSE) return reinterpret_cast<const uint32_t*>(a.taps + bj)[wd - 16];
        return 0u;
    };
    if (tid < 32) scal[0][tid] = scalar_word(first, tid);
    __syncthreads();                                   // barriers initialised, first scalars visible
    if (tid == 0) {
        for (int i = 0; i < kLeanStages && first + i < last; ++i) issue(first + i, i);
    }

    int b_cur = static_cast<int>(first / a.J);
    int j_cur = static_cast<int>(first - static_cast<long long>(b_cur) * a.J);
    const float xs = static_cast<float>(static_cast<int>((tid & 15) * 4) - 32);
    const float ys0 = static_cast<float>(static_cast<int>(tid >> 4) - 32);     // row of chunk i: + (kLeanThreads/16)*i
    float4 lv[kLeanVec], mv[kLeanVec];
    int b_loaded = -1;
    float w_next = (METHOD == PWR_METHOD_SOFTMAX) ? a.w[j_cur] : 1.f;
    constexpr bool sparse = (LOSS == LOSS_SPARSE);
    int k = 0;
    for (long long it = first; it < last; ++it, ++k) {
        const int s = k % kLeanStages;
        const uint32_t parity = (k / kLeanStages) & 1;
        const float* sz = stage_base + s * 2 * kMap;
        const float* sD = sz + kMap;
        const float wj = w_next;
        if (b_cur != b_loaded) {                      // label and mask stay in registers for the J items of a sample
            b_loaded = b_cur;
            const size_t offb = static_cast<size_t>(b_cur) * kMap + tid * 4;
#pragma unroll
            for (int i = 0; i < kLeanVec; ++i) mv[i] = ld_keep(a.m + offb + i * (kLeanThreads * 4));
#pragma unroll
            for (int i = 0; i < kLeanVec; ++i) lv[i] = ld_keep(a.L + offb + i * (kLeanThreads * 4));
        }
        const uint32_t next_word = (tid < 32 && it + 1 < last) ? scalar_word(it + 1, tid) : 0u;   // lands during the pass

        const float* sc = reinterpret_cast<const float*>(scal[k & 1]);
        const float4 st = *reinterpret_cast<const float4*>(sc);           // (z extremum, 1/sum, den, d)
        const float c = wj * kLog2e;
        const float shift = st.x * c, zref = st.x;
        float gu = sc[4], gvv = sc[5], gd = sc[6], lu = 0.f;
        if (LOSS) {
            const float eu = sc[7] - sc[10], ev = sc[8] - sc[11], ed = sc[9] - sc[12];
            gu = fmaf(a.coef.cu, eu, gu); gvv = fmaf(a.coef.cu, ev, gvv); gd = fmaf(a.coef.cu, ed, gd);
            lu = eu * eu + ev * ev + ed * ed;
        }
        const float gu63 = gu * (1.f / 63.f), gv63 = gvv * (1.f / 63.f);
        const float gdd = __fdividef(gd, st.z);
        const float dcoord = st.w;
        TapsIdx tp;
        if (sparse) {
            // footprint table of this item (its previous readers finished before the barrier that
            // ended the previous item's sums)
            tp = taps_index(scal[k & 1] + 16);
            if (tid < kFootprint) fp[tid] = footprint_entry(scal[k & 1] + 16, tid);
            __syncthreads();
        }

        mbar_wait(&full[s], parity);

        // acc: sum gp*p, sum (p-Hgt)^2, sum (D-Dgt)^2, T1 = sum p*gp*(z - zref), T2 = sum p*(z - zref)
        float4 pv[kLeanVec], gv[kLeanVec];
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        unsigned int zpos = 0;                          // PWR_METHOD_SUM: bit (4*i+kk) = z > 0
#pragma unroll
        for (int i = 0; i < kLeanVec; ++i) {
            const int cidx = tid + i * kLeanThreads;
            const float4 z4 = MapIO<TZ>::smem(sz, cidx), d4 = MapIO<TZ>::smem(sD, cidx);
            const float4 l4 = lv[i], m4 = mv[i];
            float4 t2 = make_float4(0.f, 0.f, 0.f, 0.f), t3 = t2;
            if (sparse) sparse_lookup(tp, fp, cidx >> 4, (cidx & 15) * 4, l4, m4, t2, t3);
            const float gyrow = gv63 * (ys0 + static_cast<float>(kLeanThreads / 16) * i);
            float4 gd4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float zk = comp(z4, kk);
                const float p = heat_raw<METHOD>(zk, c, shift) * st.y;
                const float mk = comp(m4, kk), dk = comp(d4, kk);
                const float rec = mk * (dk + comp(l4, kk));
                float gp = fmaf(gu63, xs + static_cast<float>(kk), gyrow);
                gp = fmaf(gdd * mk, rec - dcoord, gp);
                float gdk = gdd * p * mk * mk;
                if (sparse) {
                    const float eh = p - comp(t2, kk), ed = dk - comp(t3, kk);
                    gp = fmaf(a.coef.ch, eh, gp);
                    gdk = fmaf(a.coef.cd, ed, gdk);
                    acc[1] = fmaf(eh, eh, acc[1]);
                    acc[2] = fmaf(ed, ed, acc[2]);
                }
                const float pg = gp * p;
                acc[0] += pg;
                if (METHOD == PWR_METHOD_SOFTMAX) {
                    const float dz = zk - zref;
                    acc[3] = fmaf(pg, dz, acc[3]);
                    acc[4] = fmaf(p, dz, acc[4]);
                }
                if (METHOD == PWR_METHOD_SUM && zk > 0.f) zpos |= 1u << (4 * i + kk);
                set_comp(pv[i], kk, p);
                set_comp(gv[i], kk, gp);
                set_comp(gd4, kk, gdk);
            }
            if (a.gD != nullptr) MapIO<TZ>::st(a.gD, static_cast<size_t>(it) * kMap + cidx * 4, gd4);
        }
        // next item's scalars: parked before the barrier below, read after it
        if (tid < 32 && it + 1 < last) scal[(k + 1) & 1][tid] = next_word;
        float* ss = scr[k & 1];
        store_scattered5(warp_sum5_scattered(acc), ss);
        __syncthreads();
        // every thread is past its shared-memory reads of this item: its stage can take item k + stages
        if (tid == 0 && it + kLeanStages < last) issue(it + kLeanStages, k + kLeanStages);
        if (++j_cur == a.J) { j_cur = 0; ++b_cur; }
        if (METHOD == PWR_METHOD_SOFTMAX && it + 1 < last) w_next = a.w[j_cur];

        const float s1 = (METHOD != PWR_METHOD_GIVEN || LOSS) ? sum_partials(ss) : 0.f;
        if (a.gz != nullptr) {
#pragma unroll
            for (int i = 0; i < kLeanVec; ++i) {
                float4 g4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float g;
                    if (METHOD == PWR_METHOD_SOFTMAX) g = wj * (comp(pv[i], kk) * (comp(gv[i], kk) - s1));   // w * dL/d(w z)
                    else if (METHOD == PWR_METHOD_SUM) g = ((zpos >> (4 * i + kk)) & 1u) ? (comp(gv[i], kk) - s1) * st.y : 0.f;
                    else g = comp(gv[i], kk);
                    set_comp(g4, kk, g);
                }
                MapIO<TZ>::st(a.gz, static_cast<size_t>(it) * kMap + (tid + i * kLeanThreads) * 4, g4);
            }
        }
        if (tid == 0) {
            if (METHOD == PWR_METHOD_SOFTMAX && a.gw_partial != nullptr)
                a.gw_partial[it] = sum_partials(ss + 3 * kLeanWarps) - s1 * sum_partials(ss + 4 * kLeanWarps);
            if (LOSS && a.loss_partial != nullptr) {
                a.loss_partial[it * 3 + 0] = sum_partials(ss + kLeanWarps);
                a.loss_partial[it * 3 + 1] = sum_partials(ss + 2 * kLeanWarps);
                a.loss_partial[it * 3 + 2] = lu;
            }
        }
    }
}

// Dispatch overrides (A/B measurements, tests of both variants of a kernel): process-wide atomics set
// through pwr_set_option.  Nothing on the launch path reads the environment; the Python binding
// translates PWR_BWD_DIRECT / PWR_FWD_DIRECT / PWR_FWD_PIPE / PWR_BWD_LEAN once, when it loads the library.
static std::atomic<int> g_options[PWR_OPT_COUNT] = {};
static bool force_direct_bwd() { return g_options[PWR_OPT_BWD_DIRECT].load(std::memory_order_relaxed) != 0; }
static bool force_direct_fwd() { return g_options[PWR_OPT_FWD_DIRECT].load(std::memory_order_relaxed) != 0; }
static bool force_pipe_fwd() { return g_options[PWR_OPT_FWD_PIPE].load(std::memory_order_relaxed) != 0; }
static bool no_lean_bwd() { return g_options[PWR_OPT_BWD_NO_LEAN].load(std::memory_order_relaxed) != 0; }
static bool no_lean_fused() { return g_options[PWR_OPT_FUSED_NO_LEAN].load(std::memory_order_relaxed) != 0; }
static bool bad_method(int method) {
    return method != PWR_METHOD_SOFTMAX && method != PWR_METHOD_SUM && method != PWR_METHOD_GIVEN;
}

}  // namespace pwr

using namespace pwr;

// Dispatch over (method, loss, conv dtype).  PWR_METHOD_GIVEN (caller-supplied heat maps) exists in
// float32 only.
static bool bad_dtype(int method, int map_dtype) {
    if (map_dtype != PWR_DTYPE_F32 && map_dtype != PWR_DTYPE_F16 && map_dtype != PWR_DTYPE_BF16) return true;
    return method == PWR_METHOD_GIVEN && map_dtype != PWR_DTYPE_F32;
}
#define PWR_DISPATCH_LOSS(LAUNCH, M, TZ)                                                              \
    do {                                                                                              \
        if (loss_mode == LOSS_NONE) LAUNCH(M, LOSS_NONE, TZ);                                         \
        else if (loss_mode == LOSS_DENSE) LAUNCH(M, LOSS_DENSE, TZ);                                  \
        else LAUNCH(M, LOSS_SPARSE, TZ);                                                              \
    } while (0)
#define PWR_DISPATCH(LAUNCH)                                                                          \
    do {                                                                                              \
        if (method == PWR_METHOD_GIVEN) PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_GIVEN, float);           \
        else if (method == PWR_METHOD_SOFTMAX) {                                                      \
            if (map_dtype == PWR_DTYPE_F32)      PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_SOFTMAX, float);          \
            else if (map_dtype == PWR_DTYPE_F16) PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_SOFTMAX, __half);         \
            else                                 PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_SOFTMAX, __nv_bfloat16);  \
        } else {                                                                                      \
            if (map_dtype == PWR_DTYPE_F32)      PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_SUM, float);              \
            else if (map_dtype == PWR_DTYPE_F16) PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_SUM, __half);             \
            else                                 PWR_DISPATCH_LOSS(LAUNCH, PWR_METHOD_SUM, __nv_bfloat16);      \
        }                                                                                             \
    } while (0)

// One launcher for the one-pass kernel: the last stage (forward + loss + backward), and - with no gradient
// outputs and `stats` saved - the forward of an inner stage whose loss value is needed at forward time.
static int launch_fused(const FusedArgs& a, int method, int map_dtype, cudaStream_t s) {
    const bool sparse = a.taps != nullptr;
    const int dev = current_device(), sms = sm_count(dev);
    // No dense target map to visit (compact targets, or the uvd term alone) and float32 logits: the lean
    // variant (three CTAs per SM, per-pixel state in the ring).  Measured B=4096 NYU in DESIGN.md section 4.
    if ((sparse || a.heat_gt == nullptr) && map_dtype == PWR_DTYPE_F32 && !no_lean_fused()) {
        const int lgrid = a.items < sms * kFLCtasPerSm ? a.items : sms * kFLCtasPerSm;
#define PWR_LAUNCH_FL(M, LS)                                                                                 \
    do {                                                                                                     \
        PWR_ENSURE_DYN_SMEM(kFLSmemBytes, dev, decoder_fused_lean_kernel<M, LS>);                            \
        decoder_fused_lean_kernel<M, LS><<<lgrid, kFLThreads, kFLSmemBytes, s>>>(a);                         \
    } while (0)
        if (method == PWR_METHOD_SOFTMAX) { if (sparse) PWR_LAUNCH_FL(PWR_METHOD_SOFTMAX, LOSS_SPARSE); else PWR_LAUNCH_FL(PWR_METHOD_SOFTMAX, LOSS_DENSE); }
        else                              { if (sparse) PWR_LAUNCH_FL(PWR_METHOD_SUM, LOSS_SPARSE); else PWR_LAUNCH_FL(PWR_METHOD_SUM, LOSS_DENSE); }
#undef PWR_LAUNCH_FL
        return launch_status();
    }
    const int grid = a.items < sms * kFusedCtasPerSm ? a.items : sms * kFusedCtasPerSm;
#define PWR_LAUNCH_FUSED(M, LS, TZ)                                                                          \
    do {                                                                                                     \
        PWR_ENSURE_DYN_SMEM(kFusedSmemBytes, dev, decoder_fused_kernel<M, LS, TZ>);                          \
        decoder_fused_kernel<M, LS, TZ><<<grid, kFusedThreads, kFusedSmemBytes, s>>>(a);                     \
    } while (0)
#define PWR_FUSED_TZ(M, LS)                                                                                  \
    do {                                                                                                     \
        if (map_dtype == PWR_DTYPE_F32)      PWR_LAUNCH_FUSED(M, LS, float);                                 \
        else if (map_dtype == PWR_DTYPE_F16) PWR_LAUNCH_FUSED(M, LS, __half);                                \
        else                                 PWR_LAUNCH_FUSED(M, LS, __nv_bfloat16);                         \
    } while (0)
    if (method == PWR_METHOD_SOFTMAX) { if (sparse) PWR_FUSED_TZ(PWR_METHOD_SOFTMAX, LOSS_SPARSE); else PWR_FUSED_TZ(PWR_METHOD_SOFTMAX, LOSS_DENSE); }
    else                              { if (sparse) PWR_FUSED_TZ(PWR_METHOD_SUM, LOSS_SPARSE); else PWR_FUSED_TZ(PWR_METHOD_SUM, LOSS_DENSE); }
#undef PWR_FUSED_TZ
#undef PWR_LAUNCH_FUSED
    return launch_status();
}

extern "C" int pwr_decoder_fwd(const void* z, const float* w, const void* D, const float* L, const float* m,
                               const float* heat_gt, const float* dmap_gt, const float* uvd_gt,
                               const pwr_joint_taps* taps, float* H, float* uvd, float* stats, float* loss_partial,
                               int B, int J, int method, int map_dtype, void* stream) {
    if (bad_method(method) || bad_dtype(method, map_dtype)) return PWR_E_METHOD;
    if (int rc = check_bj(B, J)) return rc;
    if (B == 0) return 0;
    PWR_REQUIRE_PTR(z);
    if (uvd == nullptr) return PWR_E_NULL;
    PWR_OPTIONAL_PTR(D); PWR_OPTIONAL_PTR(H); PWR_OPTIONAL_PTR(stats);
    if (D != nullptr) { PWR_REQUIRE_PTR(L); PWR_REQUIRE_PTR(m); }
    if (method == PWR_METHOD_SOFTMAX && w == nullptr) return PWR_E_NULL;
    const bool loss = loss_partial != nullptr;
    if (loss) {
        if (taps == nullptr) { PWR_REQUIRE_PTR(heat_gt); PWR_REQUIRE_PTR(dmap_gt); }
        else PWR_REQUIRE_PTR(taps);
        if (uvd_gt == nullptr || D == nullptr) return PWR_E_NULL;
    }
    const int loss_mode = !loss ? LOSS_NONE : (taps != nullptr ? LOSS_SPARSE : LOSS_DENSE);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Forward with the stage loss riding along (an inner stage, model.py:205-208 + train.py:197-199): the
    // one-pass kernel without gradient outputs - z, D and the target maps arrive through its bulk-TMA ring
    // instead of 24 direct loads per thread (measured B=4096 NYU: direct kernel 0.99 ms = 75 % of the copy
    // peak on 5J + 2 maps in, J out).  `stats` is saved for the backward kernel.
    if (loss && method != PWR_METHOD_GIVEN && !force_direct_fwd()) {
        FusedArgs a;
        a.coef.cu = a.coef.ch = a.coef.cd = 0.f; a.coef.scale_dev = nullptr;
        a.z = z; a.w = w; a.D = D; a.L = L; a.m = m;
        a.heat_gt = taps == nullptr ? heat_gt : nullptr; a.dmap_gt = taps == nullptr ? dmap_gt : nullptr;
        a.uvd_gt = uvd_gt; a.taps = taps;
        a.H = H; a.uvd = uvd; a.gz = nullptr; a.gD = nullptr; a.gw_partial = nullptr; a.loss_partial = loss_partial;
        a.stats = stats; a.J = J; a.items = B * J;
        return launch_fused(a, method, map_dtype, s);
    }
    // No fused loss and the depth branch present -> persistent TMA-pipelined forward.
    // Measured (B200, f32): without the heat-map store 0.452 vs 0.576 ms (HAND17 B=4096, 6.5 TB/s); with it
    // the direct kernel below is already at 98 % of the copy peak and stays the default.
    if (loss_mode == LOSS_NONE && D != nullptr && (H == nullptr || force_pipe_fwd()) && !force_direct_fwd()) {
        FwdPipeArgs a;
        a.z = z; a.w = w; a.D = D; a.L = L; a.m = m; a.H = H; a.uvd = uvd; a.stats = stats; a.J = J; a.items = B * J;
        const int dev = current_device(), sms = sm_count(dev);
        const int grid = a.items < sms * kFwdCtasPerSm ? a.items : sms * kFwdCtasPerSm;
#define PWR_LAUNCH_FWD_PIPE(M, TZ)                                                                            \
    do {                                                                                                      \
        PWR_ENSURE_DYN_SMEM(kFwdSmemBytes, dev, decoder_fwd_pipe_kernel<M, TZ>);    \
                                                                  \
        decoder_fwd_pipe_kernel<M, TZ><<<grid, kFwdThreads, kFwdSmemBytes, s>>>(a);                            \
    } while (0)
        if (method == PWR_METHOD_GIVEN) PWR_LAUNCH_FWD_PIPE(PWR_METHOD_GIVEN, float);
        else if (method == PWR_METHOD_SOFTMAX) {
            if (map_dtype == PWR_DTYPE_F32)      PWR_LAUNCH_FWD_PIPE(PWR_METHOD_SOFTMAX, float);
            else if (map_dtype == PWR_DTYPE_F16) PWR_LAUNCH_FWD_PIPE(PWR_METHOD_SOFTMAX, __half);
            else                                 PWR_LAUNCH_FWD_PIPE(PWR_METHOD_SOFTMAX, __nv_bfloat16);
        } else {
            if (map_dtype == PWR_DTYPE_F32)      PWR_LAUNCH_FWD_PIPE(PWR_METHOD_SUM, float);
            else if (map_dtype == PWR_DTYPE_F16) PWR_LAUNCH_FWD_PIPE(PWR_METHOD_SUM, __half);
            else                                 PWR_LAUNCH_FWD_PIPE(PWR_METHOD_SUM, __nv_bfloat16);
        }
#undef PWR_LAUNCH_FWD_PIPE
        return launch_status();
    }
#define PWR_LAUNCH_FWD(M, LS, TZ)                                                                            \
    decoder_fwd_kernel<M, LS, TZ><<<B * J, kThreads, 0, s>>>(z, w, D, L, m, heat_gt, dmap_gt, uvd_gt, taps, H, \
                                                             uvd, stats, loss_partial, J)
    PWR_DISPATCH(PWR_LAUNCH_FWD);
#undef PWR_LAUNCH_FWD
    return launch_status();
}

static int launch_bwd(bool loss, const void* z, const float* w, const void* D, const float* L, const float* m,
                      const float* stats, const float* uvd, const float* g_uvd, const float* gH_up,
                      const void* gD_up, const float* heat_gt, const float* dmap_gt, const float* uvd_gt,
                      const pwr_joint_taps* taps, LossCoef coef, void* gz, void* gD, float* gw_partial,
                      float* loss_partial, int B, int J, int method, int map_dtype, void* stream) {
    if (bad_method(method) || bad_dtype(method, map_dtype)) return PWR_E_METHOD;
    if (int rc = check_bj(B, J)) return rc;
    if (B == 0) return 0;
    PWR_REQUIRE_PTR(z); PWR_REQUIRE_PTR(stats);
    PWR_OPTIONAL_PTR(D); PWR_OPTIONAL_PTR(gz); PWR_OPTIONAL_PTR(gD); PWR_OPTIONAL_PTR(gH_up); PWR_OPTIONAL_PTR(gD_up);
    if (D != nullptr) { PWR_REQUIRE_PTR(L); PWR_REQUIRE_PTR(m); }
    if (D == nullptr && (gD != nullptr || gD_up != nullptr)) return PWR_E_NULL;
    if (method == PWR_METHOD_SOFTMAX && w == nullptr) return PWR_E_NULL;
    if (loss) {
        if (taps == nullptr) { PWR_REQUIRE_PTR(heat_gt); PWR_REQUIRE_PTR(dmap_gt); }
        else PWR_REQUIRE_PTR(taps);
        if (uvd == nullptr || uvd_gt == nullptr || D == nullptr) return PWR_E_NULL;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Hot configurations -> persistent TMA-pipelined kernel: depth branch present, and at most one
    // pair of extra DENSE maps (targets XOR upstream gradients); sparse targets need no slot.
    const bool map_terms = loss && (loss_partial != nullptr || coef.ch != 0.f || coef.cd != 0.f);
    const bool need_targets = map_terms && taps == nullptr;
    if (!map_terms) taps = nullptr;
    const int loss_mode = !loss ? LOSS_NONE : (taps != nullptr ? LOSS_SPARSE : LOSS_DENSE);
    const bool need_up = gH_up != nullptr || gD_up != nullptr;
    if (D != nullptr && !force_direct_bwd()) {
        PipeArgs a;
        const bool lean = !need_targets && !need_up && !no_lean_bwd();      // no dense map beyond z, D
        const bool both = need_targets && need_up;                          // six-slot stage
        a.z = z; a.w = w; a.D = D; a.L = L; a.m = m; a.stats = stats; a.uvd = uvd; a.g_uvd = g_uvd;
        a.slot2 = need_targets ? heat_gt : gH_up;
        a.slot3 = need_targets ? static_cast<const void*>(dmap_gt) : gD_up;
        a.slot4 = both ? gH_up : nullptr;
        a.slot5 = both ? gD_up : nullptr;
        a.uvd_gt = uvd_gt; a.taps = taps; a.coef = coef; a.gz = gz; a.gD = gD; a.gw_partial = gw_partial;
        a.loss_partial = loss_partial; a.J = J; a.items = B * J;
        a.slots_are_targets = need_targets ? 1 : 0;
        const int dev = current_device(), sms = sm_count(dev);
        if (lean) {
            const int lean_grid = a.items < sms * kLeanCtasPerSm ? a.items : sms * kLeanCtasPerSm;
#define PWR_LAUNCH_LEAN(M, LS, TZ)                                                                             \
    do {                                                                                                       \
        PWR_ENSURE_DYN_SMEM(kLeanSmemBytes, dev, decoder_bwd_lean_kernel<M, LS, TZ>);    \
                                                                  \
        decoder_bwd_lean_kernel<M, LS, TZ><<<lean_grid, kLeanThreads, kLeanSmemBytes, s>>>(a);                 \
    } while (0)
            PWR_DISPATCH(PWR_LAUNCH_LEAN);
#undef PWR_LAUNCH_LEAN
            return launch_status();
        }
        const int grid = a.items < sms ? a.items : sms;
        if (both) {
            // dense targets + dense upstream gradients (inner stage, alpha < 1): loss_mode is LOSS_DENSE here
#define PWR_LAUNCH_PIPE6(M, TZ)                                                                                \
    do {                                                                                                       \
        PWR_ENSURE_DYN_SMEM(kPipeSmemBytesBoth, dev, decoder_bwd_pipe_kernel<M, LOSS_DENSE, TZ, true>);        \
        decoder_bwd_pipe_kernel<M, LOSS_DENSE, TZ, true><<<grid, kPipeThreads, kPipeSmemBytesBoth, s>>>(a);    \
    } while (0)
            if (method == PWR_METHOD_GIVEN) PWR_LAUNCH_PIPE6(PWR_METHOD_GIVEN, float);
            else if (method == PWR_METHOD_SOFTMAX) {
                if (map_dtype == PWR_DTYPE_F32)      PWR_LAUNCH_PIPE6(PWR_METHOD_SOFTMAX, float);
                else if (map_dtype == PWR_DTYPE_F16) PWR_LAUNCH_PIPE6(PWR_METHOD_SOFTMAX, __half);
                else                                 PWR_LAUNCH_PIPE6(PWR_METHOD_SOFTMAX, __nv_bfloat16);
            } else {
                if (map_dtype == PWR_DTYPE_F32)      PWR_LAUNCH_PIPE6(PWR_METHOD_SUM, float);
                else if (map_dtype == PWR_DTYPE_F16) PWR_LAUNCH_PIPE6(PWR_METHOD_SUM, __half);
                else                                 PWR_LAUNCH_PIPE6(PWR_METHOD_SUM, __nv_bfloat16);
            }
#undef PWR_LAUNCH_PIPE6
            return launch_status();
        }
#define PWR_LAUNCH_PIPE(M, LS, TZ)                                                                             \
    do {                                                                                                       \
        PWR_ENSURE_DYN_SMEM(kPipeSmemBytes, dev, decoder_bwd_pipe_kernel<M, LS, TZ>);    \
                                                                  \
        decoder_bwd_pipe_kernel<M, LS, TZ><<<grid, kPipeThreads, kPipeSmemBytes, s>>>(a);                      \
    } while (0)
        PWR_DISPATCH(PWR_LAUNCH_PIPE);
#undef PWR_LAUNCH_PIPE
        return launch_status();
    }
#define PWR_LAUNCH_BWD(M, LS, TZ)                                                                              \
    decoder_bwd_kernel<M, LS, TZ><<<B * J, kThreads, 0, s>>>(z, w, D, L, m, stats, uvd, g_uvd, gH_up, gD_up,    \
                                                             heat_gt, dmap_gt, uvd_gt, taps, coef, gz, gD,      \
                                                             gw_partial, loss_partial, J)
    PWR_DISPATCH(PWR_LAUNCH_BWD);
#undef PWR_LAUNCH_BWD
    return launch_status();
}

extern "C" int pwr_decoder_bwd(const void* z, const float* w, const void* D, const float* L, const float* m,
                               const float* stats, const float* uvd, const float* g_uvd, const float* gH_up,
                               const void* gD_up, void* gz, void* gD, float* gw_partial, int B, int J,
                               int method, int map_dtype, void* stream) {
    LossCoef coef = {0.f, 0.f, 0.f, nullptr};
    return launch_bwd(false, z, w, D, L, m, stats, uvd, g_uvd, gH_up, gD_up, nullptr, nullptr, nullptr, nullptr, coef,
                      gz, gD, gw_partial, nullptr, B, J, method, map_dtype, stream);
}

extern "C" int pwr_decoder_bwd_loss(const void* z, const float* w, const void* D, const float* L, const float* m,
                                    const float* stats, const float* uvd, const float* g_uvd, const float* gH_up,
                                    const void* gD_up, const float* heat_gt, const float* dmap_gt,
                                    const float* uvd_gt, const pwr_joint_taps* taps, float alpha, float lambda_h,
                                    float lambda_d,
                                    float loss_scale, const float* loss_scale_dev, int n_mean, void* gz,
                                    void* gD, float* gw_partial, float* loss_partial, int B, int J, int method,
                                    int map_dtype, void* stream) {
    const double n = n_mean > 0 ? static_cast<double>(n_mean) : static_cast<double>(B) * J;
    LossCoef coef;
    coef.cu = static_cast<float>(loss_scale * 2.0 * alpha / n);
    coef.ch = static_cast<float>(loss_scale * 2.0 * (1.0 - alpha) * lambda_h / n);
    coef.cd = static_cast<float>(loss_scale * 2.0 * (1.0 - alpha) * lambda_d / n);
    coef.scale_dev = loss_scale_dev;
    return launch_bwd(true, z, w, D, L, m, stats, uvd, g_uvd, gH_up, gD_up, heat_gt, dmap_gt, uvd_gt, taps, coef, gz,
                      gD, gw_partial, loss_partial, B, J, method, map_dtype, stream);
}

extern "C" int pwr_decoder_fwd_bwd_loss(const void* z, const float* w, const void* D, const float* L, const float* m,
                                        const float* heat_gt, const float* dmap_gt, const float* uvd_gt,
                                        const pwr_joint_taps* taps, float alpha, float lambda_h, float lambda_d,
                                        float loss_scale, const float* loss_scale_dev, int n_mean, float* H,
                                        float* uvd, void* gz, void* gD, float* gw_partial, float* loss_partial,
                                        int B, int J, int method, int map_dtype, void* stream) {
    if (method != PWR_METHOD_SOFTMAX && method != PWR_METHOD_SUM) return PWR_E_METHOD;
    if (bad_dtype(method, map_dtype)) return PWR_E_METHOD;
    if (int rc = check_bj(B, J)) return rc;
    if (B == 0) return 0;
    PWR_REQUIRE_PTR(z); PWR_REQUIRE_PTR(D); PWR_REQUIRE_PTR(L); PWR_REQUIRE_PTR(m);
    PWR_OPTIONAL_PTR(H); PWR_OPTIONAL_PTR(gz); PWR_OPTIONAL_PTR(gD);
    if (uvd == nullptr || uvd_gt == nullptr) return PWR_E_NULL;
    if (method == PWR_METHOD_SOFTMAX && w == nullptr) return PWR_E_NULL;
    if (taps == nullptr) { PWR_REQUIRE_PTR(heat_gt); PWR_REQUIRE_PTR(dmap_gt); }
    else PWR_REQUIRE_PTR(taps);
    const double n = n_mean > 0 ? static_cast<double>(n_mean) : static_cast<double>(B) * J;
    FusedArgs a;
    a.coef.cu = static_cast<float>(loss_scale * 2.0 * alpha / n);
    a.coef.ch = static_cast<float>(loss_scale * 2.0 * (1.0 - alpha) * lambda_h / n);
    a.coef.cd = static_cast<float>(loss_scale * 2.0 * (1.0 - alpha) * lambda_d / n);
    a.coef.scale_dev = loss_scale_dev;
    // the target maps are only visited when their loss terms matter (logged values or non-zero weights)
    const bool map_terms = loss_partial != nullptr || a.coef.ch != 0.f || a.coef.cd != 0.f;
    a.z = z; a.w = w; a.D = D; a.L = L; a.m = m;
    a.heat_gt = (map_terms && taps == nullptr) ? heat_gt : nullptr;
    a.dmap_gt = (map_terms && taps == nullptr) ? dmap_gt : nullptr;
    a.uvd_gt = uvd_gt; a.taps = map_terms ? taps : nullptr;
    a.H = H; a.uvd = uvd; a.gz = gz; a.gD = gD; a.gw_partial = gw_partial; a.loss_partial = loss_partial;
    a.J = J; a.items = B * J;
    a.stats = nullptr;
    return launch_fused(a, method, map_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int pwr_reduce_partials(const float* in, float* out, int B, int J, int C, void* stream) {
    if (out == nullptr || (in == nullptr && B != 0)) return PWR_E_NULL;
    if (B < 0 || J < 1 || C < 1 || J * C > 65535) return PWR_E_SHAPE;
    reduce_partials_kernel<<<J * C, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(in, out, B, J, C);
    return launch_status();
}

extern "C" int pwr_stage_loss(const float* loss_partial, int B, int J, float lambda_h, float lambda_d, float alpha,
                              int n_mean, float* out4, const float* gw_partial, float* gw_out, void* stream) {
    if (int rc = check_bj(B, J)) return rc;
    if (out4 == nullptr || (loss_partial == nullptr && B != 0)) return PWR_E_NULL;
    if ((gw_partial == nullptr) != (gw_out == nullptr)) return PWR_E_NULL;
    const double n = n_mean > 0 ? static_cast<double>(n_mean) : static_cast<double>(B) * J;
    stage_loss_kernel<<<1 + (gw_out != nullptr ? J : 0), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        loss_partial, B * J, static_cast<float>(lambda_h / n), static_cast<float>(lambda_d / n),
        static_cast<float>(1.0 / n), alpha, out4, gw_partial, gw_out, B, J);
    return launch_status();
}

extern "C" int pwr_scale_inplace(void* x, void* x2, float* small, int n_small, const float* scale_dev, long long n,
                                 int map_dtype, void* stream) {
    if (n < 0 || (n & 3) != 0 || n_small < 0 || n_small > kThreads) return PWR_E_SHAPE;   // whole maps only
    if (n == 0 && n_small == 0) return 0;
    PWR_REQUIRE_PTR(x); PWR_OPTIONAL_PTR(x2);
    if (scale_dev == nullptr || (n_small > 0 && small == nullptr)) return PWR_E_NULL;
    const long long n4 = n / 4;
    long long blocks = (n4 + kThreads - 1) / kThreads;
    if (blocks < 1) blocks = 1;
    const long long cap = static_cast<long long>(sm_count(current_device())) * 16;
    if (blocks > cap) blocks = cap;
    const unsigned g = static_cast<unsigned>(blocks);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* sm = n_small > 0 ? small : nullptr;
    if (map_dtype == PWR_DTYPE_F32)       scale_inplace_kernel<float><<<g, kThreads, 0, s>>>(x, x2, scale_dev, n4, sm, n_small);
    else if (map_dtype == PWR_DTYPE_F16)  scale_inplace_kernel<__half><<<g, kThreads, 0, s>>>(x, x2, scale_dev, n4, sm, n_small);
    else if (map_dtype == PWR_DTYPE_BF16) scale_inplace_kernel<__nv_bfloat16><<<g, kThreads, 0, s>>>(x, x2, scale_dev, n4, sm, n_small);
    else return PWR_E_METHOD;
    return launch_status();
}

extern "C" int pwr_recover_uvd(const float* uvd_norm, const float* box_size, const float* cube_size,
                               const float* com, double fx, double fy, double halfu, double halfv, float* uvd_px,
                               float* xyz, int B, int J, void* stream) {
    if (uvd_norm == nullptr || box_size == nullptr || cube_size == nullptr || com == nullptr) return PWR_E_NULL;
    if (B < 0 || J < 1) return PWR_E_SHAPE;
    if (B == 0) return 0;
    const int n = B * J;
    recover_uvd_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        uvd_norm, box_size, cube_size, com, static_cast<float>(fx), static_cast<float>(fy),
        static_cast<float>(halfu), static_cast<float>(halfv), uvd_px, xyz, B, J);
    return launch_status();
}

extern "C" int pwr_joint_error(const float* uvd_pred, const float* uvd_true, const float* box_size,
                               const float* cube_size, const float* com, double fx, double fy, double halfu,
                               double halfv, float* err, int B, int J, void* stream) {
    if (B < 0 || J < 1) return PWR_E_SHAPE;
    if (B == 0) return 0;
    if (uvd_pred == nullptr || uvd_true == nullptr || box_size == nullptr || cube_size == nullptr ||
        com == nullptr || err == nullptr)
        return PWR_E_NULL;
    joint_error_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        uvd_pred, uvd_true, box_size, cube_size, com, static_cast<float>(fx), static_cast<float>(fy),
        static_cast<float>(halfu), static_cast<float>(halfv), err, B, J);
    return launch_status();
}

extern "C" int pwr_version(void) { return PWR_VERSION; }

extern "C" int pwr_set_option(int option, int value) {
    if (option < 0 || option >= PWR_OPT_COUNT) return PWR_E_METHOD;
    return g_options[option].exchange(value, std::memory_order_relaxed);
}
namespace pwr { int get_option(int option) { return g_options[option].load(std::memory_order_relaxed); } }

extern "C" const char* pwr_error_string(int rc) {
    switch (rc) {
        case 0: return "ok";
        case PWR_E_NULL: return "required pointer is NULL";
        case PWR_E_SHAPE: return "shape out of range";
        case PWR_E_ALIGN: return "pointer not 16-byte aligned";
        case PWR_E_METHOD: return "unknown heat-map normalisation";
        default: return rc > 0 ? cudaGetErrorString(static_cast<cudaError_t>(rc)) : "unknown error";
    }
}
