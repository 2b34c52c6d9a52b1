// Micro-benchmark: what would the forward gain from (a) L1-resident coarse levels (all CTAs of an SM read rows of
// the same 32..192 KB window, found through %smid) and (b) reading the x0/x1 corner rows of a sample as ONE
// 128-byte request at a 64-byte-aligned address (half of them straddle two 128-byte lines)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/l1_gather_bench tools/l1_gather_bench.cu
// Prints one JSON object per line (rows are 64 bytes = one bf16 Dh-32 channel row).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\":\"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

// MODE 0: 4 lanes x 16 B = one 64-byte row per quad, ld.global.nc (what the forward does today)
// MODE 1: same, plain ld.global (L1-allocating default policy)
// MODE 2: 8 lanes x 16 B = rows r, r+1 (128 contiguous bytes at a 64-byte-aligned address)
// MODE 3: 4 lanes x 32 B (LDG.256) = rows r, r+1
// MODE 4: 2 lanes x 32 B (LDG.256) = one 64-byte row
// MODE 5: 4 lanes x 32 B (LDG.256) = row r and row r + 8 (channel-last: the next pixel of the same head, +512 B)
// PM: channel-last addressing of the model -- the 8 heads of a pixel are 8 consecutive 64-byte rows, a group of
// lanes works on ONE head (group index mod 8) and picks random pixels, as an item of the forward does
template <int MODE, bool PM = false>
__global__ void gather_kernel(const uint4* __restrict__ buf, uint32_t window_rows, uint32_t n_windows, bool per_sm,
                              int iters, float* __restrict__ sink) {
    constexpr int G = MODE == 2 ? 8 : MODE == 4 ? 2 : 4;      // lanes per group (MODE 5: 4 lanes, two rows)
    const int lane = threadIdx.x % G;
    const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t w = per_sm ? smid() % n_windows : hash32(blockIdx.x * 2654435761u) % n_windows;
    const uint4* base = buf + (uint64_t)w * window_rows * 4;
    float acc = 0.f;
    uint32_t s = hash32(grp + 12345u);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        uint32_t row = hash32(s) % (window_rows - 16);
        if (PM) row = (row & ~7u) | (grp & 7u);
        uint4 v;
        if (MODE == 1) v = base[row * 4 + lane];
        else if (MODE >= 3) {
            uint4 u;
            const uint4* p = MODE == 5 ? base + (row + (lane >> 1) * 8) * 4 + (lane & 1) * 2 : base + row * 4 + lane * 2;
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w), "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                         : "l"(p));
            v.x ^= u.x; v.w ^= u.w;
        }
        else v = __ldg(base + row * 4 + lane);
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    // resident threads per SM (dynamic shared memory as the limiter): the forward keeps 1024 threads x 128 bytes
    // of loads in flight per SM, this benchmark 2048 x (4 unrolled loads) unless limited
    {
        float* sink0; CK(cudaMalloc(&sink0, 256));
        const size_t bytes0 = 64ull << 20;
        uint4* buf0; CK(cudaMalloc(&buf0, bytes0)); CK(cudaMemset(buf0, 1, bytes0));
        CK(cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        CK(cudaFuncSetAttribute(gather_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        CK(cudaFuncSetAttribute(gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        const int iters0 = 512, threads0 = 256, grid0 = sms * 16;
        const uint32_t wrows = 4096 * 1024 / 64, nwin = (uint32_t)(bytes0 / (4096 * 1024));
        // 8 CTAs x 256 threads resident per SM in every row; only the shared-memory carve-out (and with it the
        // L1 size) changes -- and, second loop, the resident threads at a small carve-out
        for (int total_kb : {0, 32, 64, 100, 132, 164, 200}) {
            const size_t smem = (size_t)total_kb * 1024 / 8;
            const float a = time_ms([&] { gather_kernel<0><<<grid0, threads0, smem>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float b = time_ms([&] { gather_kernel<3><<<grid0, threads0, smem>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float c = time_ms([&] { gather_kernel<4><<<grid0, threads0, smem>>>(buf0, wrows, nwin, false, iters0, sink0); });
            printf("{\"resident_threads_per_sm\":2048,\"smem_per_sm_KB\":%d,\"ldg128_64B_Grows_s\":%.1f,\"ldg256_pair128B_Grows_s\":%.1f,\"ldg256_64B_Grows_s\":%.1f}\n",
                   total_kb, (double)grid0 * threads0 / 4 * iters0 / a / 1e6, (double)grid0 * threads0 / 4 * iters0 * 2 / b / 1e6,
                   (double)grid0 * threads0 / 2 * iters0 / c / 1e6);
        }
        {
            const float a = time_ms([&] { gather_kernel<0, true><<<grid0, threads0, 100 * 1024 / 8>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float c = time_ms([&] { gather_kernel<4, true><<<grid0, threads0, 100 * 1024 / 8>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float d = time_ms([&] { gather_kernel<5, true><<<grid0, threads0, 100 * 1024 / 8>>>(buf0, wrows, nwin, false, iters0, sink0); });
            printf("{\"addressing\":\"channel-last, 4 lanes x 32 B = the rows of pixel x and x + 1 of one head (+512 B)\",\"ldg256_2rows_Grows_s\":%.1f}\n",
                   (double)grid0 * threads0 / 4 * iters0 * 2 / d / 1e6);
            const size_t smem4 = (size_t)(228 * 1024 / 5) + 1024;        // 4 resident CTAs = 1024 threads, as the forward
            CK(cudaFuncSetAttribute(gather_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
            CK(cudaFuncSetAttribute(gather_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
            const float a4 = time_ms([&] { gather_kernel<0, true><<<grid0, threads0, smem4>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float c4 = time_ms([&] { gather_kernel<4, true><<<grid0, threads0, smem4>>>(buf0, wrows, nwin, false, iters0, sink0); });
            printf("{\"addressing\":\"channel-last (pixel, head) rows, one head per lane group\",\"smem_per_sm_KB\":100,"
                   "\"ldg128_64B_Grows_s\":%.1f,\"ldg256_64B_Grows_s\":%.1f,\"ldg128_64B_1024thr_186KB_Grows_s\":%.1f,\"ldg256_64B_1024thr_186KB_Grows_s\":%.1f}\n",
                   (double)grid0 * threads0 / 4 * iters0 / a / 1e6, (double)grid0 * threads0 / 2 * iters0 / c / 1e6,
                   (double)grid0 * threads0 / 4 * iters0 / a4 / 1e6, (double)grid0 * threads0 / 2 * iters0 / c4 / 1e6);
        }
        for (int ctas : {6, 4, 3, 2}) {
            // 228 KB / (ctas + 1) < smem per CTA <= 228 KB / ctas limits residency; use the smallest such size
            const size_t smem = (size_t)(228 * 1024 / (ctas + 1)) + 1024;
            const float a = time_ms([&] { gather_kernel<0><<<grid0, threads0, smem>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float b = time_ms([&] { gather_kernel<3><<<grid0, threads0, smem>>>(buf0, wrows, nwin, false, iters0, sink0); });
            const float c = time_ms([&] { gather_kernel<4><<<grid0, threads0, smem>>>(buf0, wrows, nwin, false, iters0, sink0); });
            printf("{\"resident_threads_per_sm\":%d,\"smem_per_sm_KB\":%d,\"ldg128_64B_Grows_s\":%.1f,\"ldg256_pair128B_Grows_s\":%.1f,\"ldg256_64B_Grows_s\":%.1f}\n",
                   ctas * 256, (int)(smem * ctas / 1024), (double)grid0 * threads0 / 4 * iters0 / a / 1e6, (double)grid0 * threads0 / 4 * iters0 * 2 / b / 1e6,
                   (double)grid0 * threads0 / 2 * iters0 / c / 1e6);
        }
        CK(cudaFree(buf0));
    }
    float* sink; CK(cudaMalloc(&sink, 256));
    const size_t bytes = 64ull << 20;                 // L2 resident
    uint4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
    const int iters = 512, threads = 256, grid = sms * 16;
    for (int carve : {0}) {                       // shared-memory carve-out in percent (rest is L1)
        for (int kb : {64, 128, 192, 400, 4096}) {
            const uint32_t wrows = kb * 1024 / 64, nwin = (uint32_t)(bytes / (kb * 1024));
            for (int per_sm = 0; per_sm < 2; ++per_sm) {
                float ms[5];
                CK(cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                CK(cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                CK(cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                ms[0] = time_ms([&] { gather_kernel<0><<<grid, threads>>>(buf, wrows, nwin, per_sm, iters, sink); });
                ms[1] = time_ms([&] { gather_kernel<1><<<grid, threads>>>(buf, wrows, nwin, per_sm, iters, sink); });
                ms[2] = time_ms([&] { gather_kernel<2><<<grid, threads>>>(buf, wrows, nwin, per_sm, iters, sink); });
                ms[3] = time_ms([&] { gather_kernel<3><<<grid, threads>>>(buf, wrows, nwin, per_sm, iters, sink); });
                ms[4] = time_ms([&] { gather_kernel<4><<<grid, threads>>>(buf, wrows, nwin, per_sm, iters, sink); });
                const double rows4 = (double)grid * threads / 4 * iters, rows8 = (double)grid * threads / 8 * iters * 2;
                printf("{\"window_KB\":%d,\"window_per\":\"%s\",\"smem_carveout_pct\":%d,"
                       "\"ldg_nc_64B_Grows_s\":%.1f,\"ld_64B_Grows_s\":%.1f,\"ldg_nc_pair128B_Grows_s\":%.1f,"
                       "\"ldg256_pair128B_Grows_s\":%.1f,\"ldg256_64B_Grows_s\":%.1f}\n",
                       kb, per_sm ? "SM" : "CTA", carve, rows4 / ms[0] / 1e6, rows4 / ms[1] / 1e6, rows8 / ms[2] / 1e6,
                       (double)grid * threads / 4 * iters * 2 / ms[3] / 1e6, (double)grid * threads / 2 * iters / ms[4] / 1e6);
            }
        }
    }
    return 0;
}
