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
    // 32-bit division whenever the item index fits (a 64-bit divide is a ~100-instruction subroutine per thread)
    int h, n;
    if (items <= 0x7fffffffLL) {
        const unsigned it = (unsigned)item;
        h = (int)(it % (unsigned)pb.H);
        n = (int)(it / ((unsigned)pb.H * (unsigned)pb.Lq));
    } else {
        h = (int)(item % pb.H);
        n = (int)(item / ((int64_t)pb.H * pb.Lq));
    }

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

// ---------------------------------------------------------------------------
// Variant 1 ("lean"): same item/lane mapping, but the per-sample arithmetic is
// done ONCE per sample instead of once per lane: phase 1 walks the CTA's
// samples with fully coalesced loads of locations / attention (the CTA's items
// are consecutive, so its samples are one contiguous range) and leaves
// {4 corner weights * attention, 4 row byte offsets} in shared memory; phase 2
// is the gather + FFMA2 accumulation, reading two 16-byte parameter vectors
// per sample.  The instruction count per lane-sample drops ~2.5x, which is
// what bounds this kernel (ncu: issue-bound, DRAM traffic == algorithmic).
// ---------------------------------------------------------------------------
struct __align__(16) SampleParams {
    float w[4];          // nw, ne, sw, se weight times attention, 0 for dropped corners
    uint32_t off[4];     // byte offset of the (clamped) corner row from the (n, h) base; off[0] == ~0u: no valid corner
};

// Fused prologue: with fz.ref != nullptr the kernel takes the raw Linear outputs
// instead of locations / attention -- `loc` points at the sampling offsets (N, Lq, H, L, P, 2), `attn` at the
// attention logits (N, Lq, H, L*P) -- and does ms_deform_attn.py:392-393 (softmax over L*P) and :412-416
// (ref + offsets / (W_l, H_l), divide and add separately rounded) in phase 1, with the arithmetic of
// `locations_kernel`, so locations and weights are bit-identical to the two-kernel path and never touch HBM.
// attn_out (optional) keeps the softmaxed weights for the backward.
struct FusedPrologue {
    const float* ref;       // (N, Lq, ref_levels, 2) reference points; nullptr: not fused
    int ref_levels;         // 1 or L
    float* attn_out;        // (N, Lq, H, L, P) or nullptr
};

template <int G, int K, bool VBF, bool OBF, int MINB>
__global__ void __launch_bounds__(kFwdThreads, MINB)
fwd_lean_kernel(const Problem pb, const char* __restrict__ value,
                const float* __restrict__ loc, const float* __restrict__ attn,
                char* __restrict__ out, const FusedPrologue fz) {
    constexpr int E = Vec<VBF>::kElems;
    constexpr int E2 = E / 2;
    constexpr int ES = VBF ? 2 : 4;
    constexpr int IPC = kFwdThreads / G;             // items per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int LP = pb.L * pb.P;
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    const int64_t item0 = (int64_t)blockIdx.x * IPC;
    const int nitems = (int)min((int64_t)IPC, items - item0);
    const uint32_t row_bytes = (uint32_t)(pb.vs_s * ES);

    // ---- phase 1: one thread per sample ----
    // An item's parameters take LP*32 + 16 bytes: the 16-byte pad staggers the items of a warp
    // over the banks (LP*32 alone is a multiple of 128 for the model shapes: 8-way conflicts).
    const int item_stride = LP * 32 + 16;
    const bool fused = fz.ref != nullptr;
    float2* stat_s = reinterpret_cast<float2*>(smem_raw + IPC * item_stride);     // {max, sum} per item (fused)
    if (fused) {
        // softmax statistics, one thread per item (same loops as locations_kernel: same rounding)
        if (tid < nitems) {
            const float* lg = attn + (item0 + tid) * LP;
            float m = -INFINITY;
            for (int i = 0; i < LP; ++i) m = fmaxf(m, __ldg(lg + i));
            float sum = 0.0f;
            for (int i = 0; i < LP; ++i) sum += expf(__ldg(lg + i) - m);
            stat_s[tid] = make_float2(m, sum);
        }
        __syncthreads();
    }
    {
        const int nsamp = nitems * LP;
        const float2* lsrc = reinterpret_cast<const float2*>(loc) + item0 * LP;
        const float* asrc = attn + item0 * LP;
        int il = tid / LP, sl = tid - il * LP;       // item inside the CTA, sample inside the item
        const int dil = kFwdThreads / LP, dsl = kFwdThreads - dil * LP;
        // the loads of up to three samples are issued before the first dependent instruction: one exposed
        // DRAM latency per batch instead of one per sample (a thread has 3 samples for G = 4, 6 for G = 2)
        constexpr int PF = 3;
        for (int s0 = tid; s0 < nsamp; s0 += PF * kFwdThreads) {
            float2 xy_[PF];
            float a_[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int s = s0 + u * kFwdThreads;
                xy_[u] = s < nsamp ? __ldg(lsrc + s) : make_float2(0.0f, 0.0f);
                a_[u] = s < nsamp ? __ldg(asrc + s) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (s0 + u * kFwdThreads >= nsamp) break;
                float2 xy = xy_[u];
                float a = a_[u];
                const int l = sl / pb.P;
                const int Hl = pb.geom.h[l], Wl = pb.geom.w[l];
                if (fused) {
                    const float2 st = stat_s[il];
                    a = expf(a - st.x) / st.y;
                    const int64_t nq = (item0 + il) / pb.H;
                    const float2 r = __ldg(reinterpret_cast<const float2*>(fz.ref) + nq * fz.ref_levels +
                                           (fz.ref_levels == 1 ? 0 : l));
                    xy = make_float2(__fadd_rn(r.x, __fdiv_rn(xy.x, (float)Wl)), __fadd_rn(r.y, __fdiv_rn(xy.y, (float)Hl)));
                    if (fz.attn_out != nullptr) fz.attn_out[item0 * LP + s0 + u * kFwdThreads] = a;
                }
                const Sample sm = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
                const int xc0 = min(max(sm.x0, 0), Wl - 1), xc1 = min(max(sm.x0 + 1, 0), Wl - 1);
                const int yc0 = min(max(sm.y0, 0), Hl - 1), yc1 = min(max(sm.y0 + 1, 0), Hl - 1);
                const uint32_t r0 = (uint32_t)(pb.geom.start[l] + yc0 * Wl), r1 = (uint32_t)(pb.geom.start[l] + yc1 * Wl);
                unsigned char* dst = smem_raw + il * item_stride + sl * 32;
                reinterpret_cast<float4*>(dst)[0] = make_float4(sm.w_nw * a, sm.w_ne * a, sm.w_sw * a, sm.w_se * a);
                // A corner outside the map keeps an exact zero weight and points at the clamped pixel, which is
                // a VALID corner of the same sample whenever the sample touches the map at all (so the result,
                // Inf/NaN propagation included, is what per-corner dropping gives); a sample with no valid
                // corner is flagged and skipped as a whole.  One predicate per sample instead of one per corner.
                const bool any = (sm.vx0 | sm.vx1) & (sm.vy0 | sm.vy1);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(any ? (r0 + xc0) * row_bytes : 0xffffffffu,
                                                              (r0 + xc1) * row_bytes, (r1 + xc0) * row_bytes,
                                                              (r1 + xc1) * row_bytes);
                il += dil; sl += dsl;
                if (sl >= LP) { sl -= LP; ++il; }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: G lanes per item gather and accumulate ----
    const int lane = tid % G;
    const int il = tid / G;
    // rows that are not a power-of-two number of 16-byte vectors (Dh = 48 bf16: 6) use the next power of
    // two of lanes with the surplus lanes idle: one request per corner row instead of three
    if (il >= nitems || lane * 16 >= pb.Dh * ES) return;
    const int64_t item = item0 + il;
    // 32-bit division whenever the item index fits (a 64-bit divide is a ~100-instruction subroutine per thread)
    int h, n;
    if (items <= 0x7fffffffLL) {
        const unsigned it = (unsigned)item;
        h = (int)(it % (unsigned)pb.H);
        n = (int)(it / ((unsigned)pb.H * (unsigned)pb.Lq));
    } else {
        h = (int)(item % pb.H);
        n = (int)(item / ((int64_t)pb.H * pb.Lq));
    }
    const char* vbase = value + ((int64_t)n * pb.vs_n + (int64_t)h * pb.vs_h + lane * E) * ES;

    float2 acc[K * E2];
#pragma unroll
    for (int c = 0; c < K * E2; ++c) acc[c] = make_float2(0.0f, 0.0f);

    const unsigned char* ip = smem_raw + il * item_stride;
    // B samples per step: all of their 4*K*B row loads are issued before the first use, so that every
    // thread keeps 8 (K == 1) independent 16-byte loads in flight -- the kernel is latency-bound otherwise
    constexpr int B = K == 1 ? 2 : 1;
    for (int s = 0; s < LP; s += B) {
        float wv[B][4];
        uint4 raw[B][K][4];
        bool live[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const bool in = s + b < LP;                                  // odd tail: re-reads the previous entry
            const unsigned char* e = ip + (in ? s + b : s) * 32;
            const float4 w = reinterpret_cast<const float4*>(e)[0];
            const uint4 o = reinterpret_cast<const uint4*>(e)[1];
            wv[b][0] = w.x; wv[b][1] = w.y; wv[b][2] = w.z; wv[b][3] = w.w;
            live[b] = in && o.x != 0xffffffffu;                          // sample with at least one valid corner
            const uint32_t ov[4] = {o.x, o.y, o.z, o.w};
            if (live[b]) {
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) raw[b][k][c] = ldg_nc_v4(vbase + ov[c] + k * G * 16);
            }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (!live[b]) continue;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float2 f[E2];
                    const uint4 r = raw[b][k][c];
                    if constexpr (VBF) {
                        f[0] = make_float2(bf16_lo(r.x), bf16_hi(r.x));
                        f[1] = make_float2(bf16_lo(r.y), bf16_hi(r.y));
                        f[2] = make_float2(bf16_lo(r.z), bf16_hi(r.z));
                        f[3] = make_float2(bf16_lo(r.w), bf16_hi(r.w));
                    } else {
                        f[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y));
                        f[1] = make_float2(__uint_as_float(r.z), __uint_as_float(r.w));
                    }
                    const float2 ww = make_float2(wv[b][c], wv[b][c]);
#pragma unroll
                    for (int e2 = 0; e2 < E2; ++e2) acc[k * E2 + e2] = __ffma2_rn(f[e2], ww, acc[k * E2 + e2]);
                }
        }
    }

    constexpr int OS = OBF ? 2 : 4;
    char* obase = out + (item * pb.Dh + lane * E) * OS;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        char* o = obase + k * G * E * OS;
        const float2* a2 = acc + k * E2;
        if constexpr (OBF) {
            if constexpr (E == 8) {
                uint4 v;
                v.x = pack_bf16x2(a2[0].x, a2[0].y); v.y = pack_bf16x2(a2[1].x, a2[1].y);
                v.z = pack_bf16x2(a2[2].x, a2[2].y); v.w = pack_bf16x2(a2[3].x, a2[3].y);
                *reinterpret_cast<uint4*>(o) = v;
            } else {
                uint2 v;
                v.x = pack_bf16x2(a2[0].x, a2[0].y); v.y = pack_bf16x2(a2[1].x, a2[1].y);
                *reinterpret_cast<uint2*>(o) = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < E2; e += 2)
                *reinterpret_cast<float4*>(o + e * 8) = make_float4(a2[e].x, a2[e].y, a2[e + 1].x, a2[e + 1].y);
        }
    }
}

static size_t lean_smem(const Problem& pb, int ipc) {
    return (size_t)ipc * (pb.L * pb.P * sizeof(SampleParams) + 16 + sizeof(float2));
}

template <int G, int K, bool VBF>
static cudaError_t launch_lean(const Problem& pb, const void* value, const float* loc, const float* attn,
                               void* out, bool out_bf16, int min_blocks, const FusedPrologue& fz, cudaStream_t st) {
    constexpr int IPC = kFwdThreads / G;
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    const unsigned grid = (unsigned)((items + IPC - 1) / IPC);
    const size_t smem = lean_smem(pb, IPC);
    auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, kFwdThreads, smem, st>>>(pb, (const char*)value, loc, attn, (char*)out, fz);
        return cudaGetLastError();
    };
    // min_blocks: occupancy target (registers per thread are capped accordingly); K == 1 only
    if (K == 1 && min_blocks == 6)
        return out_bf16 ? launch(fwd_lean_kernel<G, K, VBF, true, 6>) : launch(fwd_lean_kernel<G, K, VBF, false, 6>);
    if (K == 1 && min_blocks == 5)
        return out_bf16 ? launch(fwd_lean_kernel<G, K, VBF, true, 5>) : launch(fwd_lean_kernel<G, K, VBF, false, 5>);
    return out_bf16 ? launch(fwd_lean_kernel<G, K, VBF, true, 4>) : launch(fwd_lean_kernel<G, K, VBF, false, 4>);
}

// The lean kernel needs 32-bit row offsets and its parameter table in shared memory.
bool forward_lean_supported(const Problem& pb, bool value_bf16) {
    const int es = value_bf16 ? 2 : 4;
    const int nv = pb.Dh * es / 16;
    if (!(nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8 || nv == 12 || nv == 16)) return false;
    if ((int64_t)pb.S * pb.vs_s * es >= (int64_t)0x7fffffff) return false;
    const int g = nv == 3 ? 1 : nv == 6 ? 8 : nv == 12 ? 4 : nv == 16 ? 8 : nv;
    return lean_smem(pb, kFwdThreads / g) <= 96 * 1024;
}

cudaError_t forward_lean(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                         const float* attn, void* out, bool out_bf16, int min_blocks, cudaStream_t st,
                         const float* ref, int ref_levels, float* attn_out) {
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;
    const FusedPrologue fz{ref, ref_levels, attn_out};
#define MSDA_LEAN_CASE(NV, G, K)                                                          \
    case NV:                                                                              \
        return value_bf16 ? launch_lean<G, K, true>(pb, value, loc, attn, out, out_bf16, min_blocks, fz, st)  \
                          : launch_lean<G, K, false>(pb, value, loc, attn, out, out_bf16, min_blocks, fz, st);
    switch (nv) {
        MSDA_LEAN_CASE(1, 1, 1)
        MSDA_LEAN_CASE(2, 2, 1)
        MSDA_LEAN_CASE(3, 1, 3)
        MSDA_LEAN_CASE(4, 4, 1)
        MSDA_LEAN_CASE(6, 8, 1)
        MSDA_LEAN_CASE(8, 8, 1)
        MSDA_LEAN_CASE(12, 4, 3)
        MSDA_LEAN_CASE(16, 8, 2)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_LEAN_CASE
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
