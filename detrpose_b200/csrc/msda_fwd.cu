// Forward kernels: bilinear gather + attention-weighted reduction over levels
// and points, fused in registers.
// Replaces ms_deform_attn_core_pytorch forward,
// /root/reference/src/models/detrpose/ms_deform_attn.py:145-193.
#include <atomic>

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

// acc[0 .. E/2) += row * w for one 16-byte piece of a channel row (float2 accumulators, FFMA2)
template <bool VBF>
__device__ __forceinline__ void fma_row(const uint4& r, const float w, float2* acc) {
    constexpr int P2 = Vec<VBF>::kElems / 2;
    const float2 ww = make_float2(w, w);
    float2 f[P2];
    if constexpr (VBF) {
        f[0] = make_float2(bf16_lo(r.x), bf16_hi(r.x));
        f[1] = make_float2(bf16_lo(r.y), bf16_hi(r.y));
        f[2] = make_float2(bf16_lo(r.z), bf16_hi(r.z));
        f[3] = make_float2(bf16_lo(r.w), bf16_hi(r.w));
    } else {
        f[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y));
        f[1] = make_float2(__uint_as_float(r.z), __uint_as_float(r.w));
    }
#pragma unroll
    for (int e2 = 0; e2 < P2; ++e2) acc[e2] = __ffma2_rn(f[e2], ww, acc[e2]);
}

// N2 float2 (2 * N2 consecutive channels) -> out, with the widest stores the count allows
template <bool OBF, int N2>
__device__ __forceinline__ void store_out(char* o, const float2* a) {
    if constexpr (OBF) {
        if constexpr (N2 % 4 == 0) {
#pragma unroll
            for (int i = 0; i < N2; i += 4)
                *reinterpret_cast<uint4*>(o + i * 4) =
                    make_uint4(pack_bf16x2(a[i].x, a[i].y), pack_bf16x2(a[i + 1].x, a[i + 1].y),
                               pack_bf16x2(a[i + 2].x, a[i + 2].y), pack_bf16x2(a[i + 3].x, a[i + 3].y));
        } else if constexpr (N2 % 2 == 0) {
#pragma unroll
            for (int i = 0; i < N2; i += 2)
                *reinterpret_cast<uint2*>(o + i * 4) =
                    make_uint2(pack_bf16x2(a[i].x, a[i].y), pack_bf16x2(a[i + 1].x, a[i + 1].y));
        } else {
#pragma unroll
            for (int i = 0; i < N2; ++i) *reinterpret_cast<uint32_t*>(o + i * 4) = pack_bf16x2(a[i].x, a[i].y);
        }
    } else {
        if constexpr (N2 % 2 == 0) {
#pragma unroll
            for (int i = 0; i < N2; i += 2)
                *reinterpret_cast<float4*>(o + i * 8) = make_float4(a[i].x, a[i].y, a[i + 1].x, a[i + 1].y);
        } else {
#pragma unroll
            for (int i = 0; i < N2; ++i) *reinterpret_cast<float2*>(o + i * 8) = a[i];
        }
    }
}

// PAIR (K == 1, rows of at most 32 bytes): 2*G lanes per item, the lower G take the two x0 corners of a sample
// and the upper G the two x1 corners; the two halves are summed by one shuffle per accumulator at the end and
// each half stores half of the item's output vector.  With one or two lanes per item a warp carries 32 or 16
// items and every thread waits on its own chain of loads: DETRPose-N (Dh 16, bf16) runs 188 -> 118 us at batch 64
// this way (97 us on a head-major pyramid, where the x0 / x1 rows are one contiguous request).
// Streaming L2 prefetch of the pyramid.  The gather reads every pyramid byte about six times, but the FIRST touch
// of a row is a random 64-byte DRAM read (35 % of the sectors miss L2: a batch of pyramids is larger than L2).
// Each CTA therefore asks L2, with one bulk-prefetch instruction, for the contiguous slice of the pyramid that
// the CTAs `ahead` positions later gather from: DRAM is read sequentially and a little ahead of time.  Worth 3 %
// (140 -> 135 us); looking further ahead (one or more waves of CTAs) is slower than no prefetch at all.
// bytes == 0: off (pyramid not one dense block).
struct L2Prefetch {
    int64_t bytes;          // whole pyramid, all images
    uint32_t per_cta;       // slice per CTA (multiple of 16)
    uint32_t ahead;         // in CTAs
};

template <int G, int K, bool VBF, bool OBF, bool PAIR>
__global__ void __launch_bounds__(kFwdThreads, 4)
fwd_lean_kernel(const Problem pb, const char* __restrict__ value,
                const float* __restrict__ loc, const float* __restrict__ attn,
                char* __restrict__ out, const FusedPrologue fz, const L2Prefetch pf) {
    constexpr int W = 16;                            // bytes per lane and corner row: one LDG.128
    constexpr int E = Vec<VBF>::kElems;              // channels per lane vector
    constexpr int E2 = E / 2;
    constexpr int ES = VBF ? 2 : 4;
    static_assert(!PAIR || (K == 1 && G * W <= 64), "PAIR: one vector per lane, rows of at most 64 bytes");
    constexpr int GG = PAIR ? 2 * G : G;             // lanes per item
    constexpr int IPC = kFwdThreads / GG;            // items per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    if (pf.bytes != 0 && tid == 0) {
        // the first `ahead` CTAs also fetch the slices nobody is ahead of
        for (int64_t c = blockIdx.x < pf.ahead ? blockIdx.x : (int64_t)blockIdx.x + pf.ahead;
             c <= (int64_t)blockIdx.x + pf.ahead; c += pf.ahead) {
            const int64_t off = c * pf.per_cta;
            if (off < pf.bytes) {
                const uint32_t len = (uint32_t)min((int64_t)pf.per_cta, pf.bytes - off);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(value + off), "r"(len) : "memory");
            }
        }
    }
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
        // item inside the CTA, sample inside the item (tid, kFwdThreads < 2^32 / LP: fastdiv is exact)
        int il = (int)fastdiv((uint32_t)tid, pb.magic_lp), sl = tid - il * LP;
        const int dil = (int)fastdiv((uint32_t)kFwdThreads, pb.magic_lp), dsl = kFwdThreads - dil * LP;
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
                const int l = (int)fastdiv((uint32_t)sl, pb.magic_p);
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
                // PAIR keeps the two corners of one x side next to each other: {nw, sw, ne, se}
                reinterpret_cast<float4*>(dst)[0] = PAIR ? make_float4(sm.w_nw * a, sm.w_sw * a, sm.w_ne * a, sm.w_se * a)
                                                         : make_float4(sm.w_nw * a, sm.w_ne * a, sm.w_sw * a, sm.w_se * a);
                // A corner outside the map keeps an exact zero weight and points at the clamped pixel, which is
                // a VALID corner of the same sample whenever the sample touches the map at all (so the result,
                // Inf/NaN propagation included, is what per-corner dropping gives); a sample with no valid
                // corner is flagged and skipped as a whole.  One predicate per sample instead of one per corner.
                const bool any = (sm.vx0 | sm.vx1) & (sm.vy0 | sm.vy1);
                if constexpr (PAIR)
                    reinterpret_cast<uint4*>(dst)[1] = make_uint4(any ? (r0 + xc0) * row_bytes : 0xffffffffu,
                                                                  (r1 + xc0) * row_bytes,
                                                                  any ? (r0 + xc1) * row_bytes : 0xffffffffu,
                                                                  (r1 + xc1) * row_bytes);
                else
                    reinterpret_cast<uint4*>(dst)[1] = make_uint4(any ? (r0 + xc0) * row_bytes : 0xffffffffu,
                                                                  (r0 + xc1) * row_bytes, (r1 + xc0) * row_bytes,
                                                                  (r1 + xc1) * row_bytes);
                il += dil; sl += dsl;
                if (sl >= LP) { sl -= LP; ++il; }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: G lanes per item (PAIR: per x side of an item) gather and accumulate ----
    const int lane = tid % G;
    const bool active = tid / GG < nitems;
    const int il = active ? tid / GG : 0;            // surplus threads shadow item 0 (no early exit: PAIR shuffles)
    const int half = PAIR ? (tid / G) & 1 : 0;
    // rows that are not a power-of-two number of W-byte vectors (Dh = 48 bf16: 96 bytes) use the next power of
    // two of lanes with the surplus lanes idle: one request per corner row instead of three
    if (!PAIR && (!active || lane * W >= pb.Dh * ES)) return;
    const int64_t item = item0 + il;
    // (n, h) of the item: the CTA's first item is divided once (CTA-uniform, 64-bit safe), the item inside the
    // CTA adds at most IPC to its head index: one IMAD.HI per thread instead of two run-time divisions
    int h, n;
    if (items <= 0x7fffffffLL) {
        const uint32_t nq0 = (uint32_t)item0 / (uint32_t)pb.H;                 // CTA-uniform
        const uint32_t n0 = nq0 / (uint32_t)pb.Lq, q0 = nq0 - n0 * (uint32_t)pb.Lq;
        const uint32_t hh = ((uint32_t)item0 - nq0 * (uint32_t)pb.H) + (uint32_t)il;
        const uint32_t dq = fastdiv(hh, pb.magic_h);
        h = (int)(hh - dq * (uint32_t)pb.H);
        const uint32_t qq = q0 + dq;                                           // dq <= IPC
        n = (int)(pb.Lq >= IPC ? n0 + (qq >= (uint32_t)pb.Lq) : n0 + qq / (uint32_t)pb.Lq);
    } else {
        h = (int)(item % pb.H);
        n = (int)(item / ((int64_t)pb.H * pb.Lq));
    }
    const char* vbase = value + ((int64_t)n * pb.vs_n + (int64_t)h * pb.vs_h + lane * E) * ES;

    float2 acc[K * E2];
#pragma unroll
    for (int c = 0; c < K * E2; ++c) acc[c] = make_float2(0.0f, 0.0f);

    const unsigned char* ip = smem_raw + il * item_stride;
    constexpr int OS = OBF ? 2 : 4;
    if constexpr (PAIR) {
        // four samples per step: 128 bytes of independent loads in flight per thread, as in the unpaired loop
        constexpr int B = 4;
        const unsigned char* hp = ip + half * 8;
        for (int s = 0; s < LP; s += B) {
            float2 wv[B];
            uint4 raw[B][2];
            bool live[B];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const bool in = s + b < LP;
                const unsigned char* e = hp + (in ? s + b : s) * 32;
                wv[b] = *reinterpret_cast<const float2*>(e);
                const uint2 o = *reinterpret_cast<const uint2*>(e + 16);
                live[b] = in && o.x != 0xffffffffu;
                if (live[b]) {
                    raw[b][0] = ldg_nc_v4(vbase + o.x);
                    raw[b][1] = ldg_nc_v4(vbase + o.y);
                }
            }
#pragma unroll
            for (int b = 0; b < B; ++b) {
                if (!live[b]) continue;
                fma_row<VBF>(raw[b][0], wv[b].x, acc);
                fma_row<VBF>(raw[b][1], wv[b].y, acc);
            }
        }
        // x0 side + x1 side; afterwards each side stores its half of the lane's E channels
#pragma unroll
        for (int e2 = 0; e2 < E2; ++e2) {
            acc[e2].x += __shfl_xor_sync(0xffffffffu, acc[e2].x, G);
            acc[e2].y += __shfl_xor_sync(0xffffffffu, acc[e2].y, G);
        }
        if (!active) return;
        constexpr int HE2 = E2 / 2;
        float2 mine[HE2];
#pragma unroll
        for (int e2 = 0; e2 < HE2; ++e2) mine[e2] = half ? acc[HE2 + e2] : acc[e2];
        store_out<OBF, HE2>(out + (item * pb.Dh + lane * E + half * (E / 2)) * OS, mine);
        return;
    } else {
        // B samples per step: all of their 4*K*B row loads are issued before the first use, so that every
        // thread keeps 128 bytes of independent loads in flight -- the kernel is latency-bound otherwise
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
                        for (int c = 0; c < 4; ++c) raw[b][k][c] = ldg_nc_v4(vbase + ov[c] + k * G * W);
                }
            }
#pragma unroll
            for (int b = 0; b < B; ++b) {
                if (!live[b]) continue;
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) fma_row<VBF>(raw[b][k][c], wv[b][c], acc + k * E2);
            }
        }
        char* obase = out + (item * pb.Dh + lane * E) * OS;
#pragma unroll
        for (int k = 0; k < K; ++k) store_out<OBF, E2>(obase + k * G * E * OS, acc + k * E2);
    }
}

// lanes per item of the lean kernel for a channel row of nv 16-byte vectors
struct LeanShape { int g, k; };
static LeanShape lean_shape(int nv) {
    switch (nv) {
        case 1: return {1, 1};
        case 2: return {2, 1};
        case 3: return {1, 3};
        case 4: return {4, 1};
        case 6: return {8, 1};
        case 8: return {8, 1};
        case 12: return {4, 3};
        case 16: return {8, 2};
        default: return {0, 0};
    }
}

// SM count and L2 size of the current device (queried once per device, then served from a small table)
static void device_limits(int& sms, int& l2) {
    static std::atomic<int> table[16][2];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return;
    int a = table[dev][0].load(std::memory_order_relaxed), b = table[dev][1].load(std::memory_order_relaxed);
    if (a == 0 || b == 0) {
        if (cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&b, cudaDevAttrL2CacheSize, dev) != cudaSuccess || a <= 0 || b <= 0) return;
        table[dev][0].store(a, std::memory_order_relaxed);
        table[dev][1].store(b, std::memory_order_relaxed);
    }
    sms = a; l2 = b;
}

static size_t lean_smem(const Problem& pb, int ipc) {
    return (size_t)ipc * (pb.L * pb.P * sizeof(SampleParams) + 16 + sizeof(float2));
}

template <int G, int K, bool VBF, bool PAIR>
static cudaError_t launch_lean(const Problem& pb, const void* value, const float* loc, const float* attn,
                               void* out, bool out_bf16, const FusedPrologue& fz, bool l2_prefetch, cudaStream_t st) {
    constexpr int IPC = kFwdThreads / (PAIR ? 2 * G : G);
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    const unsigned grid = (unsigned)((items + IPC - 1) / IPC);
    const size_t smem = lean_smem(pb, IPC);
    // L2 prefetch: only when the pyramid is one dense channel-last block (N, S, H, Dh), a quarter wave ahead
    L2Prefetch pf{0, 0, 0};
    {
        const int es = VBF ? 2 : 4;
        const int64_t img = (int64_t)pb.S * pb.H * pb.Dh;
        const bool pm = pb.vs_h == pb.Dh && pb.vs_s == (int64_t)pb.H * pb.Dh;
        if (l2_prefetch && pm && pb.vs_n == img && grid > 0) {
            int sms = 148, l2 = 126 << 20;
            device_limits(sms, l2);
            pf.bytes = img * pb.N * es;
            pf.per_cta = (uint32_t)(((pf.bytes + grid - 1) / grid + 15) / 16 * 16);
            pf.ahead = (uint32_t)max(1, sms / 4);
            // few queries per image: the slices get large and most of their rows are touched once or never;
            // the prefetch then costs more than the misses it saves (Len_q 300, 4 levels: 88 -> 98 us)
            if (pf.per_cta > 64 * 1024) pf.bytes = 0;
            // a batch of pyramids well inside L2 is likely resident already (the hint then only costs: +2-6 % on
            // back-to-back launches at batch 8); from half of L2 on it pays even when it fits (training batch,
            // 69 MB, cold L2: 64 -> 60 us)
            if (pf.bytes <= (int64_t)l2 / 2) pf.bytes = 0;
        }
    }
    auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, kFwdThreads, smem, st>>>(pb, (const char*)value, loc, attn, (char*)out, fz, pf);
        return cudaGetLastError();
    };
    return out_bf16 ? launch(fwd_lean_kernel<G, K, VBF, true, PAIR>) : launch(fwd_lean_kernel<G, K, VBF, false, PAIR>);
}

// The lean kernel needs 32-bit row offsets and its parameter table in shared memory.
bool forward_lean_supported(const Problem& pb, bool value_bf16) {
    const int es = value_bf16 ? 2 : 4;
    const int nv = pb.Dh * es / 16;
    const int g = lean_shape(nv).g;
    if (g == 0) return false;
    if ((int64_t)pb.S * pb.vs_s * es >= (int64_t)0x7fffffff) return false;
    return lean_smem(pb, kFwdThreads / g) <= 96 * 1024;
}

cudaError_t forward_lean(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                         const float* attn, void* out, bool out_bf16, cudaStream_t st,
                         const float* ref, int ref_levels, float* attn_out, int pair_mode, int l2_prefetch) {
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;
    const FusedPrologue fz{ref, ref_levels, attn_out};
    const bool l2pf = l2_prefetch != 0;
    // rows of 16 or 32 bytes: one or two lanes per item leave a warp with too few rows in flight per instruction
    // stream; the paired form (two lanes groups per item) is the default there
    const bool pair = nv <= 2 && pair_mode != 0;
#define MSDA_LEAN(G, K, PAIR)                                                                                 \
    return value_bf16 ? launch_lean<G, K, true, PAIR>(pb, value, loc, attn, out, out_bf16, fz, l2pf, st) \
                      : launch_lean<G, K, false, PAIR>(pb, value, loc, attn, out, out_bf16, fz, l2pf, st)
    if (pair) {       // 2 * G lanes per item
        switch (nv) {
            case 1: MSDA_LEAN(1, 1, true);
            case 2: MSDA_LEAN(2, 1, true);
            default: break;
        }
    }
    switch (nv) {
        case 1: MSDA_LEAN(1, 1, false);
        case 2: MSDA_LEAN(2, 1, false);
        case 3: MSDA_LEAN(1, 3, false);
        case 4: MSDA_LEAN(4, 1, false);
        case 6: MSDA_LEAN(8, 1, false);
        case 8: MSDA_LEAN(8, 1, false);
        case 12: MSDA_LEAN(4, 3, false);
        case 16: MSDA_LEAN(8, 2, false);
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_LEAN
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
