// Forward kernels: bilinear gather + attention-weighted reduction over levels
// and points, fused in registers.
// Replaces ms_deform_attn_core_pytorch forward,
// /root/reference/src/models/detrpose/ms_deform_attn.py:145-193.
#include "msda_kernels.cuh"

namespace msda {

// ---------------------------------------------------------------------------
// Variant 0 ("flat"): one group of G lanes per (n, q, h) item, each lane owns
// K 16-byte vectors of the head's channel row (K*G*16 bytes = Dh * sizeof(T)).
// Items are ordered h-fastest so that locations / attention / output of one
// warp are contiguous in memory; every corner row is read with G coalesced
// 16-byte loads through the read-only path (L1-allocating: coarse levels get
// re-used across queries).
// ---------------------------------------------------------------------------
template <int G, int K, bool VBF, bool OBF>
__global__ void __launch_bounds__(kFwdThreads)
fwd_flat_kernel(const Problem pb, const char* __restrict__ value,
                const float* __restrict__ loc, const float* __restrict__ attn,
                char* __restrict__ out) {
    constexpr int E = Vec<VBF>::kElems;          // channels per 16-byte vector
    constexpr int ES = VBF ? 2 : 4;              // bytes per value element
    constexpr int CH = K * E;                    // channels owned by this lane

    const int lane = threadIdx.x % G;
    const int64_t item = ((int64_t)blockIdx.x * kFwdThreads + threadIdx.x) / G;
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    if (item >= items) return;
    const int h = (int)(item % pb.H);
    const int n = (int)(item / ((int64_t)pb.H * pb.Lq));

    const int LP = pb.L * pb.P;
    const float* locp = loc + item * LP * 2;
    const float* attp = attn + item * LP;
    const char* vbase = value + ((int64_t)n * pb.vs_n + (int64_t)h * pb.vs_h + lane * E) * ES;
    const int64_t row_bytes = pb.vs_s * ES;

    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.0f;

    for (int l = 0; l < pb.L; ++l) {
        const int Hl = pb.geom.h[l], Wl = pb.geom.w[l];
        const char* lbase = vbase + (int64_t)pb.geom.start[l] * row_bytes;
        for (int p = 0; p < pb.P; ++p) {
            const int sidx = l * pb.P + p;
            const float2 xy = __ldg(reinterpret_cast<const float2*>(locp) + sidx);
            const float a = __ldg(attp + sidx);
            const Sample s = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
            const int xc0 = min(max(s.x0, 0), Wl - 1), xc1 = min(max(s.x0 + 1, 0), Wl - 1);
            const int yc0 = min(max(s.y0, 0), Hl - 1), yc1 = min(max(s.y0 + 1, 0), Hl - 1);
            const char* r0 = lbase + (int64_t)(yc0 * Wl) * row_bytes;
            const char* r1 = lbase + (int64_t)(yc1 * Wl) * row_bytes;
            const float w[4] = {s.w_nw * a, s.w_ne * a, s.w_sw * a, s.w_se * a};
            const bool ok[4] = {s.vx0 && s.vy0, s.vx1 && s.vy0, s.vx0 && s.vy1, s.vx1 && s.vy1};
            const char* cp[4] = {r0 + xc0 * row_bytes, r0 + xc1 * row_bytes,
                                 r1 + xc0 * row_bytes, r1 + xc1 * row_bytes};
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint4 raw[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    raw[c] = ok[c] ? ldg_nc_v4(cp[c] + k * G * 16) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float f[E];
                    unpack<VBF>(raw[c], f);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[k * E + e] = fmaf(f[e], w[c], acc[k * E + e]);
                }
            }
        }
    }

    // out[n, q, h*Dh + channel]; item = (n*Lq + q)*H + h, so the row offset is item*Dh
    constexpr int OS = OBF ? 2 : 4;
    char* obase = out + (item * pb.Dh + lane * E) * OS;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        char* o = obase + k * G * E * OS;
        if constexpr (OBF) {
            if constexpr (E == 8) {
                uint4 v;
                v.x = pack_bf16x2(acc[k * E + 0], acc[k * E + 1]);
                v.y = pack_bf16x2(acc[k * E + 2], acc[k * E + 3]);
                v.z = pack_bf16x2(acc[k * E + 4], acc[k * E + 5]);
                v.w = pack_bf16x2(acc[k * E + 6], acc[k * E + 7]);
                *reinterpret_cast<uint4*>(o) = v;
            } else {
                uint2 v;
                v.x = pack_bf16x2(acc[k * E + 0], acc[k * E + 1]);
                v.y = pack_bf16x2(acc[k * E + 2], acc[k * E + 3]);
                *reinterpret_cast<uint2*>(o) = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; e += 4)
                *reinterpret_cast<float4*>(o + e * 4) =
                    make_float4(acc[k * E + e], acc[k * E + e + 1], acc[k * E + e + 2], acc[k * E + e + 3]);
        }
    }
}

template <int G, int K, bool VBF>
static cudaError_t launch_flat(const Problem& pb, const void* value, const float* loc, const float* attn,
                               void* out, bool out_bf16, cudaStream_t st) {
    const int64_t threads = (int64_t)pb.N * pb.Lq * pb.H * G;
    const unsigned grid = (unsigned)((threads + kFwdThreads - 1) / kFwdThreads);
    if (out_bf16)
        fwd_flat_kernel<G, K, VBF, true><<<grid, kFwdThreads, 0, st>>>(
            pb, (const char*)value, loc, attn, (char*)out);
    else
        fwd_flat_kernel<G, K, VBF, false><<<grid, kFwdThreads, 0, st>>>(
            pb, (const char*)value, loc, attn, (char*)out);
    return cudaGetLastError();
}

cudaError_t forward_flat(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                         const float* attn, void* out, bool out_bf16, cudaStream_t st) {
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;     // 16-byte vectors per channel row
#define MSDA_FWD_CASE(NV, G, K)                                                          \
    case NV:                                                                             \
        return value_bf16 ? launch_flat<G, K, true>(pb, value, loc, attn, out, out_bf16, st)  \
                          : launch_flat<G, K, false>(pb, value, loc, attn, out, out_bf16, st);
    switch (nv) {
        MSDA_FWD_CASE(1, 1, 1)
        MSDA_FWD_CASE(2, 2, 1)
        MSDA_FWD_CASE(3, 1, 3)
        MSDA_FWD_CASE(4, 4, 1)
        MSDA_FWD_CASE(6, 2, 3)
        MSDA_FWD_CASE(8, 8, 1)
        MSDA_FWD_CASE(12, 4, 3)
        MSDA_FWD_CASE(16, 8, 2)
        MSDA_FWD_CASE(24, 8, 3)
        MSDA_FWD_CASE(32, 8, 4)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_FWD_CASE
}

}  // namespace msda
