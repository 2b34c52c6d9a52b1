// Backward kernels: one pass emits grad_value (vector reductions into an fp32
// accumulation pyramid), grad_locations and grad_attention.
// Replaces autograd through ms_deform_attn_core_pytorch
// (/root/reference/src/models/detrpose/ms_deform_attn.py:145-193), i.e. ATen
// grid_sampler_2d_backward per level + the backward of cat / mul / sum.
//
// Per sample only four dot products d_k = <grad_out, value[corner k]> are
// needed for the two small gradients:
//   grad_attn = sum_k w_k d_k
//   grad_x    = A * W_l * ((d_ne - d_nw) * wy0 + (d_se - d_sw) * wy1)
//   grad_y    = A * H_l * ((d_sw - d_nw) * wx0 + (d_se - d_ne) * wx1)
// (dropped corners have d_k = 0 and receive no grad_value, GridSampler.h:238-243;
//  the factors W_l, H_l are ATen's W/2, H/2 times the 2 of "2*loc - 1", :161).
#include "msda_kernels.cuh"

namespace msda {

// load E consecutive channels of a grad_out row as fp32
template <int E, bool GBF>
__device__ __forceinline__ void load_go(const char* row, int c0, float* f) {
    if constexpr (GBF) {
        if constexpr (E == 8) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + c0 * 2));
            unpack<true>(v, f);
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(row + c0 * 2));
            f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; e += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + (c0 + e) * 4));
            f[e] = v.x; f[e + 1] = v.y; f[e + 2] = v.z; f[e + 3] = v.w;
        }
    }
}

// Variant 0 ("flat"): same item/lane mapping as the flat forward.
template <int G, int K, bool VBF, bool GBF>
__global__ void __launch_bounds__(kBwdThreads)
bwd_flat_kernel(const Problem pb, const char* __restrict__ value,
                const float* __restrict__ loc, const float* __restrict__ attn,
                const char* __restrict__ grad_out, float* __restrict__ grad_value,
                float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
    constexpr int E = Vec<VBF>::kElems;
    constexpr int ES = VBF ? 2 : 4;
    constexpr int CH = K * E;

    const int lane = threadIdx.x % G;
    const int64_t item = ((int64_t)blockIdx.x * kBwdThreads + threadIdx.x) / G;
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    if (item >= items) return;          // G divides 32: whole groups leave together
    const int h = (int)(item % pb.H);
    const int n = (int)(item / ((int64_t)pb.H * pb.Lq));

    const int LP = pb.L * pb.P;
    const float* locp = loc + item * LP * 2;
    const float* attp = attn + item * LP;
    const char* vbase = value + ((int64_t)n * pb.vs_n + (int64_t)h * pb.vs_h + lane * E) * ES;
    const int64_t row_bytes = pb.vs_s * ES;
    // grad_value is (N, S, H, Dh) fp32 contiguous
    const int64_t gv_row = (int64_t)pb.H * pb.Dh;
    float* gvbase = grad_value ? grad_value + ((int64_t)n * pb.S * gv_row + (int64_t)h * pb.Dh + lane * E)
                               : nullptr;

    float go[CH];
    {
        const char* grow = grad_out + item * pb.Dh * (GBF ? 2 : 4);
#pragma unroll
        for (int k = 0; k < K; ++k) load_go<E, GBF>(grow, (k * G + lane) * E, go + k * E);
    }

    // mask of the lanes of this group inside the warp (for the shuffles)
    const unsigned gmask = (G == 32) ? 0xffffffffu
                                     : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));

    for (int l = 0; l < pb.L; ++l) {
        const int Hl = pb.geom.h[l], Wl = pb.geom.w[l];
        const char* lbase = vbase + (int64_t)pb.geom.start[l] * row_bytes;
        float* glbase = gvbase ? gvbase + (int64_t)pb.geom.start[l] * gv_row : nullptr;
        for (int p = 0; p < pb.P; ++p) {
            const int sidx = l * pb.P + p;
            const float2 xy = __ldg(reinterpret_cast<const float2*>(locp) + sidx);
            const float a = __ldg(attp + sidx);
            const Sample s = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
            const int xc0 = min(max(s.x0, 0), Wl - 1), xc1 = min(max(s.x0 + 1, 0), Wl - 1);
            const int yc0 = min(max(s.y0, 0), Hl - 1), yc1 = min(max(s.y0 + 1, 0), Hl - 1);
            const int pix[4] = {yc0 * Wl + xc0, yc0 * Wl + xc1, yc1 * Wl + xc0, yc1 * Wl + xc1};
            const float w[4] = {s.w_nw, s.w_ne, s.w_sw, s.w_se};
            const bool ok[4] = {s.vx0 && s.vy0, s.vx1 && s.vy0, s.vx0 && s.vy1, s.vx1 && s.vy1};
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint4 raw[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    raw[c] = ok[c] ? ldg_nc_v4(lbase + (int64_t)pix[c] * row_bytes + k * G * 16)
                                   : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float f[E];
                    unpack<VBF>(raw[c], f);
#pragma unroll
                    for (int e = 0; e < E; ++e) d[c] = fmaf(f[e], go[k * E + e], d[c]);
                    const float wa = w[c] * a;
                    if (glbase != nullptr && ok[c] && wa != 0.0f) {
                        float* dst = glbase + (int64_t)pix[c] * gv_row + k * G * E;
#pragma unroll
                        for (int e = 0; e < E; e += 4)
                            red_add_v4(dst + e, wa * go[k * E + e], wa * go[k * E + e + 1],
                                       wa * go[k * E + e + 2], wa * go[k * E + e + 3]);
                    }
                }
            }
            if (grad_loc != nullptr) {
                float ga = w[0] * d[0] + w[1] * d[1] + w[2] * d[2] + w[3] * d[3];
                float gx = (d[1] - d[0]) * s.wy0 + (d[3] - d[2]) * s.wy1;
                float gy = (d[2] - d[0]) * s.wx0 + (d[3] - d[1]) * s.wx1;
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) {
                    ga += __shfl_xor_sync(gmask, ga, off);
                    gx += __shfl_xor_sync(gmask, gx, off);
                    gy += __shfl_xor_sync(gmask, gy, off);
                }
                if (lane == 0) {
                    grad_attn[item * LP + sidx] = ga;
                    reinterpret_cast<float2*>(grad_loc)[item * LP + sidx] =
                        make_float2(a * (float)Wl * gx, a * (float)Hl * gy);
                }
            }
        }
    }
}

template <int G, int K, bool VBF>
static cudaError_t launch_flat(const Problem& pb, const void* value, const float* loc, const float* attn,
                               const void* go, bool go_bf16, float* gv, float* gl, float* ga,
                               cudaStream_t st) {
    const int64_t threads = (int64_t)pb.N * pb.Lq * pb.H * G;
    const unsigned grid = (unsigned)((threads + kBwdThreads - 1) / kBwdThreads);
    if (go_bf16)
        bwd_flat_kernel<G, K, VBF, true><<<grid, kBwdThreads, 0, st>>>(
            pb, (const char*)value, loc, attn, (const char*)go, gv, gl, ga);
    else
        bwd_flat_kernel<G, K, VBF, false><<<grid, kBwdThreads, 0, st>>>(
            pb, (const char*)value, loc, attn, (const char*)go, gv, gl, ga);
    return cudaGetLastError();
}

cudaError_t backward_flat(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                          const float* attn, const void* grad_out, bool go_bf16, float* grad_value,
                          float* grad_loc, float* grad_attn, cudaStream_t st) {
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;
#define MSDA_BWD_CASE(NV, G, K)                                                                  \
    case NV:                                                                                     \
        return value_bf16                                                                        \
                   ? launch_flat<G, K, true>(pb, value, loc, attn, grad_out, go_bf16, grad_value, \
                                             grad_loc, grad_attn, st)                            \
                   : launch_flat<G, K, false>(pb, value, loc, attn, grad_out, go_bf16, grad_value, \
                                              grad_loc, grad_attn, st);
    switch (nv) {
        MSDA_BWD_CASE(1, 1, 1)
        MSDA_BWD_CASE(2, 2, 1)
        MSDA_BWD_CASE(3, 1, 3)
        MSDA_BWD_CASE(4, 4, 1)
        MSDA_BWD_CASE(6, 2, 3)
        MSDA_BWD_CASE(8, 8, 1)
        MSDA_BWD_CASE(12, 4, 3)
        MSDA_BWD_CASE(16, 8, 2)
        MSDA_BWD_CASE(24, 8, 3)
        MSDA_BWD_CASE(32, 8, 4)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_BWD_CASE
}

}  // namespace msda
