// Gate epilogue: what follows the sampler in every decoder layer.
//
// Reference: Gate.forward, /root/reference/src/models/detrpose/transformer.py:231-235 --
//     gates = sigmoid(Linear(cat[x1, x2]));  g1, g2 = gates.chunk(2, -1);  y = LayerNorm(g1*x1 + g2*x2)
// The Linear stays a library GEMM (host side); everything after it -- sigmoid, chunk, blend, LayerNorm
// with affine -- is ONE pass here: per row it reads the 2C pre-activations and the two C-wide inputs once
// and writes C outputs (+ mean / rstd for the backward), instead of the reference's six elementwise /
// normalisation kernels.  The backward recomputes the blend from the same operands, so nothing but
// (mean, rstd) is kept alive between forward and backward.
//
// HBM-bound elementwise work: one warp per row, each lane owns EPL = C/32 consecutive channels and moves
// them with 16-byte vectors (8-byte for bf16 rows of 128 or 384 channels), all loads of a row issued before the first use; the
// row statistics are two shuffle reductions (two-pass variance from registers).  The backward walks rows
// with a grid-stride loop so that the gamma / beta gradients accumulate in registers and leave as one
// shared-memory reduction + C atomics per CTA.
#include "msda_common.cuh"
#include "msda_kernels.cuh"

namespace msda {
namespace {

constexpr int kGateWarps = 8;
constexpr unsigned kFull = 0xffffffffu;

template <bool BF, int EPL>
__device__ __forceinline__ void load_row(const void* base, int64_t elem_off, float* out) {
    if constexpr (BF && EPL % 8 == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem_off);
#pragma unroll
        for (int i = 0; i < EPL / 8; ++i) {
            const uint4 r = __ldg(p + i);
            out[8 * i + 0] = __uint_as_float(r.x << 16);
            out[8 * i + 1] = __uint_as_float(r.x & 0xffff0000u);
            out[8 * i + 2] = __uint_as_float(r.y << 16);
            out[8 * i + 3] = __uint_as_float(r.y & 0xffff0000u);
            out[8 * i + 4] = __uint_as_float(r.z << 16);
            out[8 * i + 5] = __uint_as_float(r.z & 0xffff0000u);
            out[8 * i + 6] = __uint_as_float(r.w << 16);
            out[8 * i + 7] = __uint_as_float(r.w & 0xffff0000u);
        }
    } else if constexpr (BF) {
        const uint2* p = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(base) + elem_off);
#pragma unroll
        for (int i = 0; i < EPL / 4; ++i) {
            const uint2 r = __ldg(p + i);
            out[4 * i + 0] = __uint_as_float(r.x << 16);
            out[4 * i + 1] = __uint_as_float(r.x & 0xffff0000u);
            out[4 * i + 2] = __uint_as_float(r.y << 16);
            out[4 * i + 3] = __uint_as_float(r.y & 0xffff0000u);
        }
    } else {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
#pragma unroll
        for (int i = 0; i < EPL / 4; ++i) {
            const float4 r = __ldg(p + i);
            out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
        }
    }
}

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&h);
}

template <bool BF, int EPL>
__device__ __forceinline__ void store_row(void* base, int64_t elem_off, const float* v) {
    if constexpr (BF && EPL % 8 == 0) {
        uint4* p = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem_off);
#pragma unroll
        for (int i = 0; i < EPL / 8; ++i)
            p[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                              pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    } else if constexpr (BF) {
        uint2* p = reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(base) + elem_off);
#pragma unroll
        for (int i = 0; i < EPL / 4; ++i)
            p[i] = make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
    } else {
        float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + elem_off);
#pragma unroll
        for (int i = 0; i < EPL / 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
    return x;
}

// ex2.approx + rcp.approx (about 2 ulp each): the IEEE division alone was a third of the kernel's instructions
// and made the bf16 variant issue-bound
__device__ __forceinline__ float sigmoidf(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <bool PBF, bool XBF, int EPL>
__global__ void __launch_bounds__(kGateWarps * 32)
gate_fwd_kernel(const void* __restrict__ pre, const void* __restrict__ x1, const void* __restrict__ x2,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float eps,
                void* __restrict__ y, float* __restrict__ stats, const int64_t rows) {
    constexpr int C = EPL * 32;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kGateWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int c0 = lane * EPL;

    float p1[EPL], p2[EPL], a[EPL], b[EPL];
    load_row<PBF, EPL>(pre, row * (2 * C) + c0, p1);
    load_row<PBF, EPL>(pre, row * (2 * C) + C + c0, p2);
    load_row<XBF, EPL>(x1, row * C + c0, a);
    load_row<XBF, EPL>(x2, row * C + c0, b);

    float z[EPL], s = 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        z[i] = sigmoidf(p1[i]) * a[i] + sigmoidf(p2[i]) * b[i];
        s += z[i];
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        z[i] -= mean;
        q += z[i] * z[i];
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);

    float g[EPL], bt[EPL];
    load_row<false, EPL>(gamma, c0, g);
    load_row<false, EPL>(beta, c0, bt);
#pragma unroll
    for (int i = 0; i < EPL; ++i) z[i] = z[i] * rstd * g[i] + bt[i];
    store_row<XBF, EPL>(y, row * C + c0, z);
    if (stats != nullptr && lane == 0) reinterpret_cast<float2*>(stats)[row] = make_float2(mean, rstd);
}

template <bool PBF, bool XBF, int EPL>
__global__ void __launch_bounds__(kGateWarps * 32)
gate_bwd_kernel(const void* __restrict__ pre, const void* __restrict__ x1, const void* __restrict__ x2,
                const float* __restrict__ gamma, const float* __restrict__ stats, const void* __restrict__ gy,
                void* __restrict__ gpre, void* __restrict__ gx1, void* __restrict__ gx2,
                float* __restrict__ ggamma, float* __restrict__ gbeta, const int64_t rows) {
    constexpr int C = EPL * 32;
    __shared__ float red[2][kGateWarps][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = lane * EPL;

    float g[EPL], dgam[EPL], dbet[EPL];
    load_row<false, EPL>(gamma, c0, g);
#pragma unroll
    for (int i = 0; i < EPL; ++i) dgam[i] = dbet[i] = 0.0f;

    for (int64_t row = (int64_t)blockIdx.x * kGateWarps + warp; row < rows; row += (int64_t)gridDim.x * kGateWarps) {
        float p1[EPL], p2[EPL], a[EPL], b[EPL], dy[EPL];
        load_row<PBF, EPL>(pre, row * (2 * C) + c0, p1);
        load_row<PBF, EPL>(pre, row * (2 * C) + C + c0, p2);
        load_row<XBF, EPL>(x1, row * C + c0, a);
        load_row<XBF, EPL>(x2, row * C + c0, b);
        load_row<XBF, EPL>(gy, row * C + c0, dy);
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + row);

        float s1[EPL], s2[EPL], zh[EPL], dzh[EPL];
        float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            s1[i] = sigmoidf(p1[i]);
            s2[i] = sigmoidf(p2[i]);
            zh[i] = (s1[i] * a[i] + s2[i] * b[i] - st.x) * st.y;      // normalised blend
            dzh[i] = dy[i] * g[i];
            m1 += dzh[i];
            m2 += dzh[i] * zh[i];
            dgam[i] += dy[i] * zh[i];
            dbet[i] += dy[i];
        }
        m1 = warp_sum(m1) * (1.0f / C);
        m2 = warp_sum(m2) * (1.0f / C);
        float o1[EPL], o2[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const float dz = st.y * (dzh[i] - m1 - zh[i] * m2);
            o1[i] = dz * s1[i];                                       // d x1
            o2[i] = dz * s2[i];                                       // d x2
            p1[i] = o1[i] * a[i] * (1.0f - s1[i]);                    // d pre[:, :C]  (dz * x1 * s1 * (1 - s1))
            p2[i] = o2[i] * b[i] * (1.0f - s2[i]);
        }
        store_row<XBF, EPL>(gx1, row * C + c0, o1);
        store_row<XBF, EPL>(gx2, row * C + c0, o2);
        store_row<PBF, EPL>(gpre, row * (2 * C) + c0, p1);
        store_row<PBF, EPL>(gpre, row * (2 * C) + C + c0, p2);
    }

    // gamma / beta gradients: registers -> shared (one row per warp) -> one atomic per channel and CTA
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        red[0][warp][c0 + i] = dgam[i];
        red[1][warp][c0 + i] = dbet[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += kGateWarps * 32) {
        const int which = c / C, ch = c - which * C;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kGateWarps; ++w) s += red[which][w][ch];
        atomicAdd((which ? gbeta : ggamma) + ch, s);
    }
}

template <bool PBF, bool XBF, int EPL>
cudaError_t launch_fwd(const void* pre, const void* x1, const void* x2, const float* gamma, const float* beta,
                       float eps, void* y, float* stats, int64_t rows, cudaStream_t st) {
    const int64_t blocks = (rows + kGateWarps - 1) / kGateWarps;
    gate_fwd_kernel<PBF, XBF, EPL><<<(unsigned)blocks, kGateWarps * 32, 0, st>>>(pre, x1, x2, gamma, beta, eps, y,
                                                                               stats, rows);
    return cudaGetLastError();
}

template <bool PBF, bool XBF, int EPL>
cudaError_t launch_bwd(const void* pre, const void* x1, const void* x2, const float* gamma, const float* stats,
                       const void* gy, void* gpre, void* gx1, void* gx2, float* ggamma, float* gbeta, int64_t rows,
                       int sm_count, cudaStream_t st) {
    constexpr int C = EPL * 32;
    cudaError_t e = cudaMemsetAsync(ggamma, 0, C * sizeof(float), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(gbeta, 0, C * sizeof(float), st);
    if (e != cudaSuccess) return e;
    const int64_t need = (rows + kGateWarps - 1) / kGateWarps;
    const int64_t cap = (int64_t)sm_count * 8;                        // a few resident CTAs per SM, grid-stride rows
    gate_bwd_kernel<PBF, XBF, EPL><<<(unsigned)(need < cap ? need : cap), kGateWarps * 32, 0, st>>>(
        pre, x1, x2, gamma, stats, gy, gpre, gx1, gx2, ggamma, gbeta, rows);
    return cudaGetLastError();
}

}  // namespace

bool gate_supported(int C) { return C == 128 || C == 256 || C == 384 || C == 512; }

#define MSDA_GATE_DISPATCH(FN, ...)                                                            \
    switch (C / 32) {                                                                          \
        case 4:  return pre_bf16 ? (x_bf16 ? FN<true, true, 4>(__VA_ARGS__) : FN<true, false, 4>(__VA_ARGS__))     \
                                 : (x_bf16 ? FN<false, true, 4>(__VA_ARGS__) : FN<false, false, 4>(__VA_ARGS__)); \
        case 8:  return pre_bf16 ? (x_bf16 ? FN<true, true, 8>(__VA_ARGS__) : FN<true, false, 8>(__VA_ARGS__))     \
                                 : (x_bf16 ? FN<false, true, 8>(__VA_ARGS__) : FN<false, false, 8>(__VA_ARGS__)); \
        case 12: return pre_bf16 ? (x_bf16 ? FN<true, true, 12>(__VA_ARGS__) : FN<true, false, 12>(__VA_ARGS__))   \
                                 : (x_bf16 ? FN<false, true, 12>(__VA_ARGS__) : FN<false, false, 12>(__VA_ARGS__)); \
        case 16: return pre_bf16 ? (x_bf16 ? FN<true, true, 16>(__VA_ARGS__) : FN<true, false, 16>(__VA_ARGS__))   \
                                 : (x_bf16 ? FN<false, true, 16>(__VA_ARGS__) : FN<false, false, 16>(__VA_ARGS__)); \
        default: return cudaErrorInvalidValue;                                                 \
    }

cudaError_t gate_forward(const void* pre, bool pre_bf16, const void* x1, const void* x2, bool x_bf16,
                         const float* gamma, const float* beta, float eps, void* y, float* stats, int64_t rows,
                         int C, cudaStream_t st) {
    MSDA_GATE_DISPATCH(launch_fwd, pre, x1, x2, gamma, beta, eps, y, stats, rows, st)
}

cudaError_t gate_backward(const void* pre, bool pre_bf16, const void* x1, const void* x2, bool x_bf16,
                          const float* gamma, const float* stats, const void* gy, void* gpre, void* gx1, void* gx2,
                          float* ggamma, float* gbeta, int64_t rows, int C, int sm_count, cudaStream_t st) {
    MSDA_GATE_DISPATCH(launch_bwd, pre, x1, x2, gamma, stats, gy, gpre, gx1, gx2, ggamma, gbeta, rows, sm_count, st)
}

}  // namespace msda
