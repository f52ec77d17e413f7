// PCIe host->device probe behind the design of the e2e feed:
// how fast can ONE GPU pull (a) whole raw frames, (b) only the crop windows of the
// frames, from pinned host memory - with the copy engine (cudaMemcpyAsync, one call or
// one call per window, cudaMemcpy2DAsync, cudaMemcpyBatchAsync) or with a kernel that
// reads the mapped host memory directly (128-bit loads, coalesced row segments)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pcie_probe tools/pcie_probe.cu
//   build/pcie_probe [B]            (prints one JSON object)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Win { int r0, c0, rows, cols; };      // window of frame b: rows [r0, r0+rows), cols [c0, c0+cols), c0 % 8 == 0, cols % 8 == 0

// whole-buffer copy through the SMs
__global__ void copy_all(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

// one CTA per (sample, row group): gathers the window rows into a compact [B, WR, WC] u16 buffer
template <int UNROLL>
__global__ void fetch_windows(const uint16_t* __restrict__ frames, int Hf, int Wf, const Win* __restrict__ wins,
                              uint16_t* __restrict__ out, int WR, int WC, int groups) {
    const int b = blockIdx.x / groups, g = blockIdx.x % groups;
    const Win w = wins[b];
    const int chunks_per_row = w.cols / 8;                 // 16-byte chunks
    const int total = w.rows * chunks_per_row;
    const uint16_t* src = frames + static_cast<size_t>(b) * Hf * Wf;
    uint16_t* dst = out + static_cast<size_t>(b) * WR * WC;
    const int per = (total + groups - 1) / groups;
    const int lo = g * per, hi = min(total, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x * UNROLL) {
        uint4 v[UNROLL];
        int rr[UNROLL], cc[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int k = i + u * blockDim.x;
            rr[u] = k / chunks_per_row; cc[u] = k - rr[u] * chunks_per_row;
            if (k < hi) v[u] = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(w.r0 + rr[u]) * Wf + w.c0 + cc[u] * 8);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int k = i + u * blockDim.x;
            if (k < hi) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(rr[u]) * WC + cc[u] * 8) = v[u];
        }
    }
}


// ---- bulk-TMA variant: one warp per CTA, every lane moves one window row at a time host -> shared -> HBM
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void __launch_bounds__(32)
fetch_windows_tma(const uint16_t* __restrict__ frames, int Hf, int Wf, const Win* __restrict__ wins,
                  uint16_t* __restrict__ out, int WR, int WC, int B) {
    extern __shared__ __align__(128) unsigned char ring[];          // [32][slot_bytes]
    __shared__ __align__(8) uint64_t bars[32];
    const int lane = threadIdx.x;
    const int slot_bytes = WC * 2;
    unsigned char* slot = ring + static_cast<size_t>(lane) * slot_bytes;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[lane])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const Win w = wins[b];
        const uint32_t bytes = static_cast<uint32_t>(w.cols) * 2;
        const uint16_t* src = frames + static_cast<size_t>(b) * Hf * Wf + static_cast<size_t>(w.r0) * Wf + w.c0;
        uint16_t* dst = out + static_cast<size_t>(b) * WR * WC;
        for (int r = lane; r < w.rows; r += 32) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[lane])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(slot)), "l"(src + static_cast<size_t>(r) * Wf), "r"(bytes), "r"(smem_u32(&bars[lane])) : "memory");
            asm volatile(
                "{\n.reg .pred P1;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D1;\nbra W1;\nD1:\n}" ::"r"(
                    smem_u32(&bars[lane])), "r"(parity) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + static_cast<size_t>(r) * WC),
                         "r"(smem_u32(slot)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            parity ^= 1;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float time_ms(cudaStream_t s, int reps, void (*fn)(void*), void* ctx) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    fn(ctx);
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(a, s));
    for (int i = 0; i < reps; ++i) fn(ctx);
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

struct Ctx {
    cudaStream_t s;
    uint16_t* host; uint16_t* dev; uint16_t* out; Win* wins_d; std::vector<Win> wins;
    int B, Hf, Wf, WR, WC, groups, threads;
    size_t bytes;
};

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 2048;
    const int Hf = 480, Wf = 640, WR = 352, WC = 368;
    Ctx c;
    c.B = B; c.Hf = Hf; c.Wf = Wf; c.WR = WR; c.WC = WC;
    c.bytes = static_cast<size_t>(B) * Hf * Wf * 2;
    CK(cudaSetDevice(0));
    CK(cudaStreamCreate(&c.s));
    CK(cudaHostAlloc(&c.host, c.bytes, cudaHostAllocDefault));      // what torch's pin_memory does
    for (size_t i = 0; i < c.bytes / 2; i += 997) c.host[i] = static_cast<uint16_t>(i);
    CK(cudaMalloc(&c.dev, c.bytes));
    CK(cudaMalloc(&c.out, static_cast<size_t>(B) * WR * WC * 2));
    // NYU-like windows: z ~ U(500,1000), box = 176000/z, centre U(0.25,0.75) of the frame, clipped
    srand(1);
    size_t win_bytes = 0, row_bytes = 0;
    for (int b = 0; b < B; ++b) {
        const double z = 500.0 + 500.0 * (rand() / (double)RAND_MAX);
        const int box = static_cast<int>(176000.0 / z) / 2 * 2;
        const int cu = static_cast<int>((0.25 + 0.5 * (rand() / (double)RAND_MAX)) * Wf);
        const int cv = static_cast<int>((0.25 + 0.5 * (rand() / (double)RAND_MAX)) * Hf);
        int r0 = cv - box / 2, r1 = cv + box / 2, c0 = cu - box / 2, c1 = cu + box / 2;
        if (r0 < 0) r0 = 0; if (r1 > Hf) r1 = Hf; if (c0 < 0) c0 = 0; if (c1 > Wf) c1 = Wf;
        c0 &= ~7; c1 = (c1 + 7) & ~7;
        Win w = {r0, c0, r1 - r0, c1 - c0};
        c.wins.push_back(w);
        win_bytes += static_cast<size_t>(w.rows) * w.cols * 2;
        row_bytes += static_cast<size_t>(w.rows) * Wf * 2;
    }
    CK(cudaMalloc(&c.wins_d, B * sizeof(Win)));
    CK(cudaMemcpy(c.wins_d, c.wins.data(), B * sizeof(Win), cudaMemcpyHostToDevice));

    printf("{\"B\": %d, \"frame_bytes\": %zu, \"window_bytes\": %zu, \"row_window_bytes\": %zu", B, c.bytes, win_bytes, row_bytes);

    float ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
        CK(cudaMemcpyAsync(c.dev, c.host, c.bytes, cudaMemcpyHostToDevice, c.s)); }, &c);
    printf(", \"memcpy_full\": {\"ms\": %.3f, \"gbs\": %.2f}", ms, c.bytes / ms / 1e6);

    for (int blocks : {148, 296, 592, 1184}) {
        c.groups = blocks;
        ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
            copy_all<<<c.groups, 512, 0, c.s>>>(reinterpret_cast<const uint4*>(c.host), reinterpret_cast<uint4*>(c.dev), c.bytes / 16); }, &c);
        CK(cudaGetLastError());
        printf(", \"kernel_full_%d\": {\"ms\": %.3f, \"gbs\": %.2f}", blocks, ms, c.bytes / ms / 1e6);
    }

    // per-sample contiguous row windows with one cudaMemcpyAsync each
    ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
        for (int b = 0; b < c.B; ++b) {
            const Win& w = c.wins[b];
            CK(cudaMemcpyAsync(c.dev + static_cast<size_t>(b) * c.Hf * c.Wf + static_cast<size_t>(w.r0) * c.Wf,
                               c.host + static_cast<size_t>(b) * c.Hf * c.Wf + static_cast<size_t>(w.r0) * c.Wf,
                               static_cast<size_t>(w.rows) * c.Wf * 2, cudaMemcpyHostToDevice, c.s));
        } }, &c);
    printf(", \"memcpy_rows_per_sample\": {\"ms\": %.3f, \"gbs_useful\": %.2f, \"samples_per_s\": %.0f}", ms, row_bytes / ms / 1e6, B / ms * 1e3);


    // per-sample contiguous row windows, ONE cudaMemcpyBatchAsync call
    {
        static std::vector<void*> dsts, srcs; static std::vector<size_t> sizes;
        for (int b = 0; b < B; ++b) {
            const Win& w = c.wins[b];
            dsts.push_back(c.dev + static_cast<size_t>(b) * Hf * Wf + static_cast<size_t>(w.r0) * Wf);
            srcs.push_back(c.host + static_cast<size_t>(b) * Hf * Wf + static_cast<size_t>(w.r0) * Wf);
            sizes.push_back(static_cast<size_t>(w.rows) * Wf * 2);
        }
        struct BCtx { Ctx* c; void** d; void** s; size_t* z; } bc = {&c, dsts.data(), srcs.data(), sizes.data()};
        cudaMemcpyAttributes at = {};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        static cudaMemcpyAttributes at_s; at_s = at;
        size_t idx0 = 0; static size_t idx_s; idx_s = idx0;
        size_t fail = 0;
        cudaError_t e = cudaMemcpyBatchAsync(bc.d, bc.s, bc.z, B, &at_s, &idx_s, 1, &fail, c.s);
        if (e == cudaSuccess) {
            CK(cudaStreamSynchronize(c.s));
            cudaEvent_t a, b2; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b2));
            CK(cudaEventRecord(a, c.s));
            for (int r = 0; r < 3; ++r) CK(cudaMemcpyBatchAsync(bc.d, bc.s, bc.z, B, &at_s, &idx_s, 1, &fail, c.s));
            CK(cudaEventRecord(b2, c.s)); CK(cudaEventSynchronize(b2));
            CK(cudaEventElapsedTime(&ms, a, b2)); ms /= 3;
            printf(", \"memcpy_batch_rows\": {\"ms\": %.3f, \"gbs_useful\": %.2f, \"samples_per_s\": %.0f}", ms, row_bytes / ms / 1e6, B / ms * 1e3);
        } else {
            printf(", \"memcpy_batch_rows\": {\"error\": \"%s\"}", cudaGetErrorString(e));
            cudaGetLastError();
        }
    }

    // per-sample 2-D windows with cudaMemcpy2DAsync
    ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
        for (int b = 0; b < c.B; ++b) {
            const Win& w = c.wins[b];
            CK(cudaMemcpy2DAsync(c.out + static_cast<size_t>(b) * c.WR * c.WC, c.WC * 2,
                                 c.host + static_cast<size_t>(b) * c.Hf * c.Wf + static_cast<size_t>(w.r0) * c.Wf + w.c0, c.Wf * 2,
                                 static_cast<size_t>(w.cols) * 2, w.rows, cudaMemcpyHostToDevice, c.s));
        } }, &c);
    printf(", \"memcpy2d_per_sample\": {\"ms\": %.3f, \"gbs_useful\": %.2f, \"samples_per_s\": %.0f}", ms, win_bytes / ms / 1e6, B / ms * 1e3);

    // zero-copy gather kernel
    for (int groups : {1, 2, 4, 8}) {
        for (int threads : {256, 512}) {
            c.groups = groups; c.threads = threads;
            ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
                fetch_windows<4><<<c.B * c.groups, c.threads, 0, c.s>>>(c.host, c.Hf, c.Wf, c.wins_d, c.out, c.WR, c.WC, c.groups); }, &c);
            CK(cudaGetLastError());
            printf(", \"kernel_windows_g%d_t%d\": {\"ms\": %.3f, \"gbs_useful\": %.2f, \"samples_per_s\": %.0f}", groups, threads, ms,
                   win_bytes / ms / 1e6, B / ms * 1e3);
        }
    }

    // bulk-TMA gather: rows host -> shared -> HBM, one row per lane in flight
    for (int grid : {32, 64, 148, 592}) {
        c.groups = grid;
        CK(cudaFuncSetAttribute(fetch_windows_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * c.WC * 2));
        ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
            fetch_windows_tma<<<c.groups, 32, 32 * c.WC * 2, c.s>>>(c.host, c.Hf, c.Wf, c.wins_d, c.out, c.WR, c.WC, c.B); }, &c);
        CK(cudaGetLastError());
        printf(", \"tma_windows_grid%d\": {\"ms\": %.3f, \"gbs_useful\": %.2f, \"samples_per_s\": %.0f}", grid, ms,
               win_bytes / ms / 1e6, B / ms * 1e3);
    }
    // LDG gather with a small persistent-like grid for comparison (groups = 1, few CTAs is not expressible here: B*groups CTAs)
    // same kernel on device-resident frames (upper bound without PCIe)
    c.groups = 4;
    ms = time_ms(c.s, 3, [](void* p) { Ctx& c = *static_cast<Ctx*>(p);
        fetch_windows<4><<<c.B * c.groups, 256, 0, c.s>>>(c.dev, c.Hf, c.Wf, c.wins_d, c.out, c.WR, c.WC, c.groups); }, &c);
    printf(", \"kernel_windows_from_hbm\": {\"ms\": %.3f, \"gbs_useful\": %.2f}", ms, win_bytes / ms / 1e6);
    printf("}\n");
    return 0;
}
