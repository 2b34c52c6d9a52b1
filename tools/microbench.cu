// Micro-benchmarks that size the design of the sampler kernels on B200:
// random row gathers (L2 / L1 / shared), vector reductions into global memory,
// shared-memory float atomics, warp-exclusive shared accumulation, memset.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/microbench tools/microbench.cu
// Prints one JSON object per line.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\":\"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// ---- gather: G lanes x 16 B per row, rows chosen by hash inside a per-CTA window ----------------
// window_rows: rows addressable by one CTA (models "one (n,h) head" vs "whole buffer")
template <int G>
__global__ void gather_kernel(const uint4* __restrict__ buf, uint64_t total_rows, uint32_t window_rows,
                              int iters, float* __restrict__ sink) {
    const int lane = threadIdx.x % G;
    const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint64_t win0 = ((uint64_t)hash32(blockIdx.x * 2654435761u) % (total_rows / window_rows)) * window_rows;
    float acc = 0.f;
    uint32_t s = hash32(grp + 12345u);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const uint64_t row = win0 + (hash32(s) % window_rows);
        const uint4 v = __ldg(buf + row * G + lane);
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
    }
    if (acc == 123.456f) sink[0] = acc;
}

// ---- shared-memory gather ---------------------------------------------------------------------
template <int G>
__global__ void lds_gather_kernel(int rows, int iters, float* __restrict__ sink) {
    extern __shared__ uint4 sm[];
    for (int i = threadIdx.x; i < rows * G; i += blockDim.x) sm[i] = make_uint4(i, i, i, i);
    __syncthreads();
    const int lane = threadIdx.x % G;
    uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x / G + 999u);
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const uint4 v = sm[(hash32(s) % rows) * G + lane];
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
    }
    if (acc == 123.456f) sink[0] = acc;
}

// ---- vector reductions into global memory: 8 lanes x red.v4.f32 = one 128-byte row -------------
__global__ void red_v4_kernel(float* __restrict__ buf, uint64_t total_rows, uint32_t window_rows, int iters) {
    const int lane = threadIdx.x % 8;
    const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / 8;
    const uint64_t win0 = ((uint64_t)hash32(blockIdx.x * 2654435761u) % (total_rows / window_rows)) * window_rows;
    uint32_t s = hash32(grp + 777u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        float* p = buf + (win0 + (hash32(s) % window_rows)) * 32 + lane * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
}
// 16 lanes x red.v2.f32
__global__ void red_v2_kernel(float* __restrict__ buf, uint64_t total_rows, uint32_t window_rows, int iters) {
    const int lane = threadIdx.x % 16;
    const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / 16;
    const uint64_t win0 = ((uint64_t)hash32(blockIdx.x * 2654435761u) % (total_rows / window_rows)) * window_rows;
    uint32_t s = hash32(grp + 777u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        float* p = buf + (win0 + (hash32(s) % window_rows)) * 32 + lane * 2;
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(1.f), "f"(2.f) : "memory");
    }
}
// 32 lanes x scalar red
__global__ void red_s_kernel(float* __restrict__ buf, uint64_t total_rows, uint32_t window_rows, int iters) {
    const int lane = threadIdx.x % 32;
    const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const uint64_t win0 = ((uint64_t)hash32(blockIdx.x * 2654435761u) % (total_rows / window_rows)) * window_rows;
    uint32_t s = hash32(grp + 777u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(buf + (win0 + (hash32(s) % window_rows)) * 32 + lane, 1.0f);
    }
}
// bf16x2 packed: 4 lanes x red.v4.bf16x2 = 64-byte bf16 row of 32 channels
__global__ void red_bf16_kernel(uint32_t* __restrict__ buf, uint64_t total_rows, uint32_t window_rows, int iters) {
    const int lane = threadIdx.x % 4;
    const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / 4;
    const uint64_t win0 = ((uint64_t)hash32(blockIdx.x * 2654435761u) % (total_rows / window_rows)) * window_rows;
    uint32_t s = hash32(grp + 777u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        uint32_t* p = buf + (win0 + (hash32(s) % window_rows)) * 16 + lane * 4;
        asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u) : "memory");
    }
}

// ---- shared-memory float atomics: 32 lanes = one 128-byte row -----------------------------------
__global__ void atoms_kernel(int rows, int iters, float* __restrict__ sink) {
    extern __shared__ float smf[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) smf[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x % 32;
    uint32_t s = hash32(blockIdx.x * blockDim.x / 32 + threadIdx.x / 32 + 4242u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&smf[(hash32(s) % rows) * 32 + lane], 1.0f);
    }
    __syncthreads();
    if (smf[threadIdx.x] == 123.456f) sink[0] = 1.f;
}
// 8 lanes x 4 floats: each lane does 4 scalar shared atomics (row of 32 floats per 8 lanes)
__global__ void atoms8_kernel(int rows, int iters, float* __restrict__ sink) {
    extern __shared__ float smf[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) smf[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x % 8;
    uint32_t s = hash32((blockIdx.x * blockDim.x + threadIdx.x) / 8 + 4242u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        float* p = &smf[(hash32(s) % rows) * 32];
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(p + k * 8 + lane, 1.0f);   // conflict-free within the 8 lanes
    }
    __syncthreads();
    if (smf[threadIdx.x] == 123.456f) sink[0] = 1.f;
}
// warp-exclusive accumulation without atomics (lane = channel, rows partitioned per warp)
__global__ void excl_kernel(int rows, int iters, float* __restrict__ sink) {
    extern __shared__ float smf[];
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) smf[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, nwarps = blockDim.x / 32;
    const int rows_per_warp = rows / nwarps;
    uint32_t s = hash32(blockIdx.x * nwarps + warp + 31u);
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        float* p = &smf[(warp * rows_per_warp + hash32(s) % rows_per_warp) * 32 + lane];
        *p += 1.0f;
    }
    __syncthreads();
    if (smf[threadIdx.x] == 123.456f) sink[0] = 1.f;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();                                  // warm-up
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
    printf("{\"device\":\"%s\",\"sms\":%d,\"l2_mb\":%.1f}\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0);
    const int sms = prop.multiProcessorCount;
    float* sink; CK(cudaMalloc(&sink, 256));

    // ---------------- gathers ----------------
    const size_t big_bytes = 512ull << 20;               // 512 MiB, >> L2
    uint4* big; CK(cudaMalloc(&big, big_bytes)); CK(cudaMemset(big, 1, big_bytes));
    struct GCase { const char* name; int G; size_t footprint; uint32_t window_rows; };
    const GCase gcases[] = {
        {"gather64B_random_512MB", 4, big_bytes, 0},
        {"gather64B_window_400KB(level0 of one head, head-major)", 4, big_bytes, 6400},
        {"gather64B_window_100KB(level1 of one head)", 4, big_bytes, 1600},
        {"gather64B_footprint_64MB_random(L2 resident)", 4, 64ull << 20, 0},
        {"gather128B_random_512MB", 8, big_bytes, 0},
        {"gather128B_window_800KB", 8, big_bytes, 6400},
        {"gather128B_footprint_64MB_random(L2 resident)", 8, 64ull << 20, 0},
        {"gather32B_random_512MB", 2, big_bytes, 0},
    };
    for (const GCase& c : gcases) {
        const int G = c.G, iters = 256, threads = 256;
        const uint64_t rows = c.footprint / (16 * G);
        const uint32_t win = c.window_rows ? c.window_rows : (uint32_t)rows;
        const int grid = sms * 16;
        const double total_rows = (double)grid * threads / G * iters;
        float ms = 0;
        if (G == 4) ms = time_ms([&] { gather_kernel<4><<<grid, threads>>>(big, rows, win, iters, sink); });
        if (G == 8) ms = time_ms([&] { gather_kernel<8><<<grid, threads>>>(big, rows, win, iters, sink); });
        if (G == 2) ms = time_ms([&] { gather_kernel<2><<<grid, threads>>>(big, rows, win, iters, sink); });
        printf("{\"bench\":\"%s\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps\":%.1f}\n", c.name, ms,
               total_rows / ms / 1e6, total_rows * 16 * G / ms / 1e6);
    }

    // ---------------- shared gathers ----------------
    CK(cudaFuncSetAttribute(lds_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(lds_gather_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int G : {4, 8}) {
        const int rows = 128 * 1024 / (16 * G), iters = 2048, threads = 1024, grid = sms;
        const double total_rows = (double)grid * threads / G * iters;
        float ms = G == 4 ? time_ms([&] { lds_gather_kernel<4><<<grid, threads, 128 * 1024>>>(rows, iters, sink); })
                          : time_ms([&] { lds_gather_kernel<8><<<grid, threads, 128 * 1024>>>(rows, iters, sink); });
        printf("{\"bench\":\"lds_gather_%dB_rows_128KB\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps\":%.1f}\n", 16 * G, ms,
               total_rows / ms / 1e6, total_rows * 16 * G / ms / 1e6);
    }

    // ---------------- global reductions ----------------
    float* acc = reinterpret_cast<float*>(big);
    struct RCase { const char* name; size_t footprint; uint32_t window_rows; };
    const RCase rcases[] = {
        {"random_512MB", big_bytes, 0},
        {"footprint_64MB(L2 resident)", 64ull << 20, 0},
        {"window_800KB(level0 of one head, fp32)", big_bytes, 6400},
        {"window_800KB_in_64MB", 64ull << 20, 6400},
    };
    for (const RCase& c : rcases) {
        const int iters = 128, threads = 256, grid = sms * 16;
        const uint64_t rows = c.footprint / 128;
        const uint32_t win = c.window_rows ? c.window_rows : (uint32_t)rows;
        float ms = time_ms([&] { red_v4_kernel<<<grid, threads>>>(acc, rows, win, iters); });
        double total_rows = (double)grid * threads / 8 * iters;
        printf("{\"bench\":\"red_v4_f32_128Brow_%s\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", c.name, ms,
               total_rows / ms / 1e6, total_rows * 128 / ms / 1e6);
        ms = time_ms([&] { red_v2_kernel<<<grid, threads>>>(acc, rows, win, iters); });
        total_rows = (double)grid * threads / 16 * iters;
        printf("{\"bench\":\"red_v2_f32_128Brow_%s\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", c.name, ms,
               total_rows / ms / 1e6, total_rows * 128 / ms / 1e6);
        ms = time_ms([&] { red_s_kernel<<<grid, threads>>>(acc, rows, win, iters); });
        total_rows = (double)grid * threads / 32 * iters;
        printf("{\"bench\":\"red_scalar_f32_128Brow_%s\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", c.name, ms,
               total_rows / ms / 1e6, total_rows * 128 / ms / 1e6);
        const uint64_t rows16 = c.footprint / 64;
        const uint32_t win16 = c.window_rows ? c.window_rows : (uint32_t)rows16;
        ms = time_ms([&] { red_bf16_kernel<<<grid, threads>>>(reinterpret_cast<uint32_t*>(big), rows16, win16, iters); });
        total_rows = (double)grid * threads / 4 * iters;
        printf("{\"bench\":\"red_v4_bf16x2_64Brow_%s\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", c.name, ms,
               total_rows / ms / 1e6, total_rows * 64 / ms / 1e6);
    }

    // ---------------- shared atomics / exclusive accumulation ----------------
    CK(cudaFuncSetAttribute(atoms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(atoms8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(excl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
        const int rows = 1024, iters = 1024, threads = 1024, grid = sms;      // 128 KB of fp32 rows
        float ms = time_ms([&] { atoms_kernel<<<grid, threads, rows * 128>>>(rows, iters, sink); });
        double total_rows = (double)grid * threads / 32 * iters;
        printf("{\"bench\":\"smem_atomicAdd_f32_row32lanes\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", ms,
               total_rows / ms / 1e6, total_rows * 128 / ms / 1e6);
        ms = time_ms([&] { atoms8_kernel<<<grid, threads, rows * 128>>>(rows, iters, sink); });
        total_rows = (double)grid * threads / 8 * iters;
        printf("{\"bench\":\"smem_atomicAdd_f32_row8lanes_x4\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", ms,
               total_rows / ms / 1e6, total_rows * 128 / ms / 1e6);
        ms = time_ms([&] { excl_kernel<<<grid, threads, rows * 128>>>(rows, iters, sink); });
        total_rows = (double)grid * threads / 32 * iters;
        printf("{\"bench\":\"smem_exclusive_rmw_row32lanes\",\"ms\":%.4f,\"Grows_per_s\":%.3f,\"GBps_payload\":%.1f}\n", ms,
               total_rows / ms / 1e6, total_rows * 128 / ms / 1e6);
    }

    // ---------------- memset / copy ----------------
    {
        float ms = time_ms([&] { CK(cudaMemsetAsync(big, 0, big_bytes)); });
        printf("{\"bench\":\"memset_512MB\",\"ms\":%.4f,\"GBps\":%.1f}\n", ms, big_bytes / ms / 1e6);
        uint4* dst; CK(cudaMalloc(&dst, big_bytes));
        ms = time_ms([&] { CK(cudaMemcpyAsync(dst, big, big_bytes, cudaMemcpyDeviceToDevice)); });
        printf("{\"bench\":\"memcpy_d2d_512MB\",\"ms\":%.4f,\"GBps_rw\":%.1f}\n", ms, 2.0 * big_bytes / ms / 1e6);
    }
    return 0;
}
