// Microbenchmark: can TMA's tile::gather4 (sm_100) fetch random 64-byte rows faster than LDG.128 through L1?
// The forward kernel is bound by the SM's per-line request rate on the load/store path (about 0.6 lines per
// cycle and SM, 178 G rows/s for L2-resident 64-byte rows; profiles/r01_microbench.jsonl).  gather4 brings four
// rows of a 2-D tensor (here: the four bilinear corners of a sample in the (N*S, C) view of the pyramid, 64-byte
// box = one head's channels) into shared memory with one instruction, bypassing the L1 tag stage.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/tma_gather_bench tools/tma_gather_bench.cu -lcuda
//   tools/build/tma_gather_bench [box_rows]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kWarps = 8;            // warps per CTA
constexpr int kBatch = 8;            // gather4 per stage and warp  (32 rows = 2 KB)
constexpr int kStages = 2;
constexpr int kRowBytes = 64;        // box width: 32 bf16

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                        uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        :: "r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// every warp: lane 0 issues kBatch gather4 per stage, all lanes consume (one 16-byte LDS per row quarter)
__global__ void __launch_bounds__(kWarps * 32)
tma_gather_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int width_rows, int iters, float* sink,
                  unsigned long long* timeout_flag) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t stage_bytes = kBatch * 4 * kRowBytes;
    const uint32_t my = sbase + warp * (kStages * stage_bytes);
    const uint32_t bars = sbase + kWarps * kStages * stage_bytes + warp * kStages * 8;
    if (lane == 0)
        for (int s = 0; s < kStages; ++s) mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t seed = (blockIdx.x * kWarps + warp) * 7919u + 13u;
    auto issue = [&](int stage, int it) {
        if (lane == 0) {
            mbar_expect_tx(bars + stage * 8, stage_bytes);
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t r = hash(seed + it * kBatch + b) % (uint32_t)(rows - width_rows - 2);
                const int col = (int)((hash(seed ^ (it * 131 + b)) & 7u) * 32u);       // one of 8 heads
                gather4(my + stage * stage_bytes + b * 4 * kRowBytes, &tmap, col, (int)r, (int)r + 1,
                        (int)r + width_rows, (int)r + width_rows + 1, bars + stage * 8);
            }
        }
    };
    float acc = 0.0f;
    issue(0, 0);
    for (int it = 0; it < iters; ++it) {
        const int stage = it & 1;
        if (it + 1 < iters) issue(stage ^ 1, it + 1);
        const uint32_t parity = (it >> 1) & 1;
        long long spins = 0;
        while (!mbar_try_wait(bars + stage * 8, parity)) {
            if (++spins > 20000000LL) { if (lane == 0) *timeout_flag = 1; return; }   // never hang the box
        }
        // consume: lane l reads 16 bytes of row l
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(my + stage * stage_bytes + lane * kRowBytes + (lane & 3) * 16));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
        __syncwarp();
    }
    if (acc == 123.456f) sink[0] = acc;
}

// the same access pattern through LDG.128 (4 lanes per 64-byte row, 8 rows per warp instruction)
__global__ void __launch_bounds__(kWarps * 32)
ldg_gather_kernel(const char* base, int rows, int width_rows, int iters, float* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t seed = (blockIdx.x * kWarps + warp) * 7919u + 13u;
    float acc = 0.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < kBatch; b += 2) {
            // lanes 0-15: sample b (4 corners x 4 lanes), lanes 16-31: sample b+1
            const int bb = b + (lane >> 4), corner = (lane >> 2) & 3, q = lane & 3;
            const uint32_t r = hash(seed + it * kBatch + bb) % (uint32_t)(rows - width_rows - 2);
            const int col = (int)((hash(seed ^ (it * 131 + bb)) & 7u) * 32u);
            const size_t row = (size_t)r + (corner & 1) + (corner >> 1) * width_rows;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + row * 512 + col * 2 + q * 16));
            acc += __uint_as_float(v.x) + __uint_as_float(v.w);
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const int rows = 131072, cols = 256;                  // 64 MB of bf16: L2 resident, like r01's microbenchmark
    void* d = nullptr;
    CK(cudaMalloc(&d, (size_t)rows * cols * 2));
    CK(cudaMemset(d, 0, (size_t)rows * cols * 2));
    float* sink; CK(cudaMalloc(&sink, 4));
    unsigned long long* flag; CK(cudaMalloc(&flag, 8)); CK(cudaMemset(flag, 0, 8));

    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((EncodeFn)fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"error\":\"cuTensorMapEncodeTiled %d\"}\n", (int)r); return 1; }

    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 2000;
    const size_t smem = kWarps * kStages * kBatch * 4 * kRowBytes + kWarps * kStages * 8 + 64;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
        const int grid = sms * ctas_per_sm;
        const double total_rows = (double)grid * kWarps * iters * kBatch * 4;
        float ms;
        tma_gather_kernel<<<grid, kWarps * 32, smem>>>(tmap, rows, 80, 50, sink, flag);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        tma_gather_kernel<<<grid, kWarps * 32, smem>>>(tmap, rows, 80, iters, sink, flag);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long f = 0; CK(cudaMemcpy(&f, flag, 8, cudaMemcpyDeviceToHost));
        printf("{\"bench\":\"tma_gather4_64B_rows\",\"box_rows\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"Grows_per_s\":%.1f,\"timed_out\":%llu}\n",
               box_rows, ctas_per_sm, ms, total_rows / ms / 1e6, f);
        if (f) break;
        ldg_gather_kernel<<<grid, kWarps * 32>>>((const char*)d, rows, 80, 50, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        ldg_gather_kernel<<<grid, kWarps * 32>>>((const char*)d, rows, 80, iters, sink);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"bench\":\"ldg128_64B_rows\",\"ctas_per_sm\":%d,\"ms\":%.4f,\"Grows_per_s\":%.1f}\n", ctas_per_sm, ms,
               total_rows / ms / 1e6);
    }
    return 0;
}
