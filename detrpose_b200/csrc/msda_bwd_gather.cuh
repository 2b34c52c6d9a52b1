// Backward, gather form ("pixel-owner"): no global atomics, no zero-fill.
//
// One CTA owns one (image n, head h, level l).  Per chunk of queries it
//   P0  stages the head's grad_out rows in shared memory,
//   P1  counts the level's samples per base pixel (x0, y0)          [native u32 shared atomics]
//   P2  exclusive-scans the counts,
//   P3  scatters one 16-byte record {A, wx1, wy1, id} per sample into bin order,
//   P4  walks the level's pixels: the G lanes that own a pixel read its value row ONCE, then
//       visit the (at most four) bins whose samples touch it, accumulating
//       grad_value[pixel] += A*w*grad_out[q] in registers and storing the corner dot product
//       <grad_out[q], value[pixel]> for that (sample, corner) in shared memory,
//   P5  turns the four dots of every sample into grad_attention / grad_locations.
//
// Replaces autograd through ms_deform_attn_core_pytorch
// (/root/reference/src/models/detrpose/ms_deform_attn.py:145-193): ATen
// grid_sampler_2d_backward's per-corner atomics into grad_input and its grad_grid,
// plus the backward of cat / mul / sum.  Dropped corners (GridSampler.h:205,238-243) are
// never visited because only in-range pixels gather.
//
// Bins: base pixels x0 in [-1, W-1], y0 in [-1, H-1] -> (W+1)*(H+1) bins, row-major, so that
// the two bins feeding a pixel from one row, (x-1, y) and (x, y), are adjacent and their
// records form one contiguous range.  Counters are packed two u16 per u32 (a chunk holds
// fewer than 65536 samples) to keep the DETRPose shape (1080 queries x 4 points, 32 bf16
// channels, 81x81 bins) inside 227 KB with a single chunk.
// The kernel and its launch templates live in this header so that the instantiations can be spread over several
// translation units (csrc/msda_bwd_gather_case*.cu) and compile in parallel; msda_bwd_gather.cu holds the plan and
// the dispatch.
#pragma once

#include "msda_kernels.cuh"

namespace msda {

// Diagnostic: device buffer that receives clock64() at the phase boundaries of every CTA (first query chunk only):
// slot blockIdx.x * 8 + phase.  Set by msda_b200_debug_phase_buffer, passed to the kernel as an argument.
unsigned long long* phase_buffer();

namespace gather_detail {

constexpr int kMaxSmem = 227 * 1024;

struct GatherPlan {
    int q_chunk;        // queries per chunk
    int n_chunks;
    int bins_words;     // u32 words of packed counters (largest level)
    size_t smem_bytes;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

// shared-memory carve-up for a chunk of `qc` queries
struct SmemLayout {
    int off_g, off_rec, off_dots, off_bins, off_scan, off_order, total;
    __host__ __device__ SmemLayout(int qc, int P, int row_bytes, int bins_words) {
        off_g = 0;
        off_rec = align16((qc + 1) * row_bytes);      // + one all-zero row read by idle visit slots
        off_dots = off_rec + qc * P * 16;
        off_bins = off_dots + qc * P * 16;
        off_scan = off_bins + align16(bins_words * 4);
        off_order = off_scan + 64 * 4;              // per-warp visiting order of a 64-pixel tile (u16)
        total = off_order + 32 * 64 * 2;
    }
};

__device__ __forceinline__ unsigned half_of(unsigned packed, int b) { return (packed >> ((b & 1) * 16)) & 0xffffu; }

}  // namespace gather_detail
using namespace gather_detail;

// fp32 pair helpers: packed FFMA2 (sm_100) halves the FMA issue slots of the visit loop
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) { return __ffma2_rn(a, b, c); }

template <bool BF>
__device__ __forceinline__ void unpack2(const uint4& v, float2* f) {     // 16 bytes -> E/2 fp32 pairs
    if constexpr (BF) {
        f[0] = make_float2(bf16_lo(v.x), bf16_hi(v.x));
        f[1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
        f[2] = make_float2(bf16_lo(v.z), bf16_hi(v.z));
        f[3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
    } else {
        f[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
        f[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
    }
}

// Explicit shared-space accesses on 32-bit addresses: keeps generic->shared conversions out of the
// hot loops.  volatile keeps their order relative to the barriers; ptxas still schedules them freely.
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts_f4(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned lds_u16(uint32_t a) {
    unsigned short r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ unsigned lds_u8(uint32_t a) {
    unsigned r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts_u8(uint32_t a, unsigned v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory");
}
// 256-bit global accesses (sm_100): the 4 lanes that own a pixel write its 128-byte fp32 gradient
// row as ONE request instead of two half-line requests (the SM's request rate bounds this phase)
__device__ __forceinline__ void stg_v8(float* p, const float2 a, const float2 b, const float2 c, const float2 d) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y), "f"(d.x), "f"(d.y) : "memory");
}
__device__ __forceinline__ void ldg_v8(const float* p, float2& a, float2& b, float2& c, float2& d) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y), "=f"(c.x), "=f"(c.y), "=f"(d.x), "=f"(d.y) : "l"(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(PENDING) : "memory"); }

// ONE: the whole query range fits one chunk and grad_value is overwritten -- the common case; the
// read-modify-write paths and the chunk bookkeeping compile away (less register pressure).
template <int G, int K, bool VBF, int THREADS, bool SMALL, bool ONE, bool FUSED = false>
__global__ void __launch_bounds__(THREADS, 1)
bwd_gather_kernel(const Problem pb, const char* __restrict__ value, const float* __restrict__ loc,
                  const float* __restrict__ attn, const char* __restrict__ grad_out,
                  float* __restrict__ grad_value, float* __restrict__ grad_loc,
                  float* __restrict__ grad_attn, const int accumulate, const int q_chunk,
                  const int bins_words_max, const float* __restrict__ ref, const int ref_levels,
                  unsigned long long* __restrict__ phase_buf) {
    constexpr int E = Vec<VBF>::kElems;
    constexpr int E2 = E / 2;
    constexpr int ES = VBF ? 2 : 4;
    constexpr int VPR = G * K;                       // 16-byte vectors per channel row
    constexpr int UPW = 32 / G;                      // pixel units a warp handles per iteration
    constexpr int NWARPS = THREADS / 32;
    constexpr int NB = 4;                            // visits per batch (dots are transpose-reduced per batch)
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem[];

    const int tid = threadIdx.x;
    const int lane = tid % G;
    const int lane32 = tid & 31, warp = tid >> 5, gsub = lane32 / G;
    // CTA -> (image, head, level).  The levels of one (image, head) sit next to each other so that they
    // share grad_out / locations / attention in L2; only in the tail of the grid (the last kTail pairs)
    // the long finest-level CTAs are issued first, so that the last CTAs to start are the short ones.
    int l, nh;
    {
        constexpr int kTail = 148;
        const int NH = pb.N * pb.H, b = blockIdx.x;
        const int head_ctas = max(NH - kTail, 0) * pb.L;
        if (b < head_ctas || pb.L == 1) { nh = b / pb.L; l = b - nh * pb.L; }
        else {
            const int r = b - head_ctas, t = NH - head_ctas / pb.L;       // t pairs in the tail
            if (r < t) { nh = head_ctas / pb.L + r; l = 0; }
            else { const int r2 = r - t; nh = head_ctas / pb.L + r2 / (pb.L - 1); l = 1 + r2 % (pb.L - 1); }
        }
    }
    const int h = nh % pb.H;
    const int n = nh / pb.H;
    const int Hl = pb.geom.h[l], Wl = pb.geom.w[l];
    const int BW = Wl + 1;                           // bins per row: x0 in [-1, W-1]
    const int nbins = BW * (Hl + 1);
    // counters are indexed by bin+1: after the scatter half k holds end(bin k-1), half 0 stays 0,
    // so begin(bin b) = ends16[b] and end(bin b) = ends16[b + 1]
    const int nwords = (nbins + 3) / 2;
    const int P = pb.P, LP = pb.L * pb.P;
    const int row_bytes = pb.Dh * ES;

    const SmemLayout lay(q_chunk, P, row_bytes, bins_words_max);
    float4* rec_s = reinterpret_cast<float4*>(smem + lay.off_rec);
    float* dots_s = reinterpret_cast<float*>(smem + lay.off_dots);
    // the unsorted records live in the dots region (a_dots) until P4
    unsigned* bins = reinterpret_cast<unsigned*>(smem + lay.off_bins);
    unsigned* scan_s = reinterpret_cast<unsigned*>(smem + lay.off_scan);
    int* work_s = reinterpret_cast<int*>(smem + lay.off_order + 32 * 32);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t a_g = sbase + lay.off_g + lane * 16;                        // this lane's vector of g row 0
    uint32_t a_rec = sbase + lay.off_rec;
    uint32_t a_dots = sbase + lay.off_dots;
    const uint32_t a_ends = sbase + lay.off_bins;
    // keep the three bases the visit loop uses in registers: left alone, the compiler rebuilds them from the
    // shared-window base and the kernel parameters (S2UR / ULEA / LDC) inside the loop
    uint32_t a_order = sbase + lay.off_order + warp * 32;                // lane holding the pixel of rank r
    asm volatile("" : "+r"(a_g), "+r"(a_rec), "+r"(a_dots), "+r"(a_order));

    const char* vlevel = value + ((int64_t)n * pb.vs_n + (int64_t)pb.geom.start[l] * pb.vs_s +
                                  (int64_t)h * pb.vs_h + lane * E) * ES;
    const int64_t vrow = pb.vs_s * ES;
    const int64_t gv_row = (int64_t)pb.H * pb.Dh;
    // byte / element offsets inside one level fit 32 bits (make_plan checks): cheaper address arithmetic
    const uint32_t vrow32 = (uint32_t)vrow, gv_row32 = (uint32_t)gv_row;
    float* gvlevel = grad_value
        ? grad_value + ((int64_t)n * pb.S + pb.geom.start[l]) * gv_row + (int64_t)h * pb.Dh + lane * E
        : nullptr;
    // 256-bit row stores (one request per 128-byte row) measured slower than two 128-bit stores here
    // (602 vs 568 us at batch 64): kept behind this switch
    constexpr bool wide_store = false;
    const int npix = Hl * Wl;
    const float inv_w = 1.0f / (float)Wl;
    const float fW = (float)Wl, fH = (float)Hl;
    // (q, p) of sample tid and the step to sample tid + THREADS
    const int q_t0 = tid / P, p_t0 = tid - q_t0 * P;
    const int dq = THREADS / P, dp = THREADS - dq * P;

    unsigned long long* const pbuf = phase_buf;
#define MSDA_STAMP(i) do { if (pbuf != nullptr && tid == 0 && chunk == 0) pbuf[(size_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)

    for (int q0 = 0, chunk = 0; q0 < pb.Lq; q0 += q_chunk, ++chunk) {
        const int qc = min(q_chunk, pb.Lq - q0);
        const int nsamp = qc * P;
        MSDA_STAMP(0);
        // sample (q, p) of this head and level sits at float2/float index s0 + q * sstride + p
        const int64_t s0 = (((int64_t)n * pb.Lq + q0) * pb.H + h) * LP + l * P;
        const int64_t sstride = (int64_t)pb.H * LP;

        // ---- P0: clear the counters; stage (asynchronously) this head's grad_out rows and the level's
        //      locations / attention, so that no phase below waits on a dependent global load ----
        // staging area for locations (8 B/sample) and attention (4 B/sample): the sorted-record region,
        // which is free until P3 and again after P4
        // (locations at a_rec, attention weights behind them: read in P1 through the shared-window addresses)
        auto stage_samples = [&]() {
            if ((P & 3) == 0) {
                // a query's P samples are contiguous: 16-byte copies (2 per 4 locations, 1 per 4 weights)
                const int vq = P / 4;                                // 16-byte attention vectors per query
                // locations: the two 16-byte halves of a 32-byte piece go to neighbouring lanes, so one warp
                // instruction touches 16 lines instead of 32 (the SM's cost is per line, not per byte)
                for (int j = tid; j < 2 * qc * vq; j += THREADS) {
                    const int i = j >> 1, half = j & 1;
                    const int q = i / vq, v = i - q * vq;
                    cp_async16(a_rec + (q * P + v * 4) * 8 + half * 16,
                               reinterpret_cast<const float2*>(loc) + s0 + q * sstride + v * 4 + half * 2);
                }
                for (int i = tid; i < qc * vq; i += THREADS) {
                    const int q = i / vq, v = i - q * vq;
                    cp_async16(a_rec + nsamp * 8 + (q * P + v * 4) * 4, attn + s0 + q * sstride + v * 4);
                }
            } else {
                int q = q_t0, p = p_t0;
                for (int i = tid; i < nsamp; i += THREADS) {
                    const int64_t sidx = s0 + q * sstride + p;
                    cp_async8(a_rec + i * 8, reinterpret_cast<const float2*>(loc) + sidx);
                    cp_async4(a_rec + nsamp * 8 + i * 4, attn + sidx);
                    q += dq; p += dp;
                    if (p >= P) { p -= P; ++q; }
                }
            }
        };
        {
            for (int i = tid; i < nwords; i += THREADS) bins[i] = 0u;
            if (tid == 0) work_s[0] = 0;
            if (tid < VPR) reinterpret_cast<uint4*>(smem + lay.off_g)[qc * VPR + tid] = make_uint4(0, 0, 0, 0);
            stage_samples();
            cp_async_commit();                               // group A: locations + attention
            const char* gsrc = grad_out + (((int64_t)n * pb.Lq + q0) * pb.H * pb.Dh + (int64_t)h * pb.Dh) * ES;
            const int64_t gstride = (int64_t)pb.H * pb.Dh * ES;
            for (int i = tid; i < qc * VPR; i += THREADS) {
                const int q = i / VPR, v = i % VPR;
                cp_async16(sbase + lay.off_g + i * 16, gsrc + q * gstride + v * 16);
            }
            cp_async_commit();                               // group B: grad_out rows, needed from P4 on
            cp_async_wait<1>();
        }
        __syncthreads();
        MSDA_STAMP(1);

        // ---- P1: one pass over the samples: build the (unsorted) record, count per base pixel ----
        // record = {A*wy0, A*wy1, wx1, (bin+1) | corner validity << 20}.  Samples that are not binned
        // (no valid corner, or |A| < 1e-30 so that wy cannot be recovered from A*wy later: their
        // contribution to grad_value / grad_locations is below 1e-30 |grad_out| and is dropped) are finished here.
        {
            const uint32_t a_att = a_rec + nsamp * 8;          // shared-window addresses: nothing to rebuild per sample
            auto sample = [&](const int i, const int q) {
                float2 xy = lds_f2(a_rec + i * 8);
                const float a = lds_f32(a_att + i * 4);
                if constexpr (FUSED) {
                    // fused prologue (row f1): the staged pairs are sampling offsets; location = ref + offset /
                    // (W_l, H_l) with the divide and the add rounded separately (ms_deform_attn.py:414-416)
                    const float2 r = __ldg(reinterpret_cast<const float2*>(ref) +
                                           ((int64_t)n * pb.Lq + q0 + q) * ref_levels + (ref_levels == 1 ? 0 : l));
                    xy = make_float2(__fadd_rn(r.x, __fdiv_rn(xy.x, fW)), __fadd_rn(r.y, __fdiv_rn(xy.y, fH)));
                }
                const Sample s = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
                const bool inside = s.x0 >= -1 && s.x0 < Wl && s.y0 >= -1 && s.y0 < Hl;
                float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);     // unbinned: .w == 0
                if (inside && !(fabsf(a) < 1e-30f)) {        // NaN weights stay on the main path (NaN out)
                    const int b = (s.y0 + 1) * BW + (s.x0 + 1) + 1;
                    atomicAdd(&bins[b >> 1], 1u << ((b & 1) * 16));
                    const int packed = b | ((int)s.vx0 << 20) | ((int)s.vx1 << 21) | ((int)s.vy0 << 22) | ((int)s.vy1 << 23);
                    t = make_float4(a * s.wy0, a * s.wy1, s.wx1, __int_as_float(packed));
                } else if (SMALL) {
                    // rare: grad_locations is zero (A == 0 or all corners dropped); grad_attention needs the
                    // four corner dots, gathered directly
                    float ga = 0.0f;
                    if (inside) {
                        // the staged copy of grad_out may still be in flight: read the row from global
                        const char* grow = reinterpret_cast<const char*>(grad_out) +
                            ((((int64_t)n * pb.Lq + q0 + q) * pb.H + h) * pb.Dh) * ES;
                        const char* vl = value + ((int64_t)n * pb.vs_n + (int64_t)pb.geom.start[l] * pb.vs_s +
                                                  (int64_t)h * pb.vs_h) * ES;
                        // (the corner products are only needed here: the common path never computes them)
                        const float wk[4] = {(s.vx0 & s.vy0) ? __fmul_rn(s.wx0, s.wy0) : 0.0f,
                                             (s.vx1 & s.vy0) ? __fmul_rn(s.wx1, s.wy0) : 0.0f,
                                             (s.vx0 & s.vy1) ? __fmul_rn(s.wx0, s.wy1) : 0.0f,
                                             (s.vx1 & s.vy1) ? __fmul_rn(s.wx1, s.wy1) : 0.0f};
                        const int px[4] = {s.x0, s.x0 + 1, s.x0, s.x0 + 1};
                        const int py[4] = {s.y0, s.y0, s.y0 + 1, s.y0 + 1};
                        for (int k = 0; k < 4; ++k) {
                            if (wk[k] == 0.0f || px[k] < 0 || px[k] >= Wl || py[k] < 0 || py[k] >= Hl) continue;
                            const char* vr = vl + (int64_t)(py[k] * Wl + px[k]) * vrow;
                            float d = 0.0f;
                            for (int c = 0; c < pb.Dh; ++c) {
                                const float vv = VBF ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(vr)[c])
                                                     : reinterpret_cast<const float*>(vr)[c];
                                const float gg = VBF ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(grow)[c])
                                                     : reinterpret_cast<const float*>(grow)[c];
                                d = fmaf(vv, gg, d);
                            }
                            ga = fmaf(wk[k], d, ga);
                        }
                    }
                    t.x = ga;                                // final {grad_attn, grad_loc = 0} of this sample
                }
                sts_f4(a_dots + i * 16, t);
            };
            int q = q_t0, p = p_t0;
            for (int i = tid; i < nsamp; i += THREADS) {
                sample(i, q);
                q += dq; p += dp;
                if (p >= P) { p -= P; ++q; }
            }
        }
        __syncthreads();
        MSDA_STAMP(2);

        // ---- P2: exclusive scan of the packed counters (in place: count -> start) ----
        {
            const int per = (nwords + THREADS - 1) / THREADS;
            const int w0 = min(tid * per, nwords), w1 = min(w0 + per, nwords);
            unsigned local = 0;
            for (int j = w0; j < w1; ++j) { const unsigned c = bins[j]; local += (c & 0xffffu) + (c >> 16); }
            unsigned incl = local;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned t = __shfl_up_sync(FULL, incl, off);
                if (lane32 >= off) incl += t;
            }
            if (lane32 == 31) scan_s[warp] = incl;
            __syncthreads();
            if (tid < 32) {
                const unsigned wsum = tid < NWARPS ? scan_s[tid] : 0u;
                unsigned winc = wsum;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned t = __shfl_up_sync(FULL, winc, off);
                    if (tid >= off) winc += t;
                }
                scan_s[32 + tid] = winc - wsum;          // exclusive warp offsets
            }
            __syncthreads();
            unsigned running = scan_s[32 + warp] + incl - local;
            for (int j = w0; j < w1; ++j) {
                const unsigned c = bins[j];
                const unsigned lo = running; running += c & 0xffffu;
                const unsigned hi = running; running += c >> 16;
                bins[j] = lo | (hi << 16);
            }
        }
        __syncthreads();
        MSDA_STAMP(3);

        // ---- P3: move the records into bin order (start -> end, in place) ----
        // sorted record = {A*wy0, A*wy1, wx1, corner validity << 28 | query << 16 | sample slot}
        {
            int q = q_t0, p = p_t0;
            for (int i = tid; i < nsamp; i += THREADS) {
                float4 r = lds_f4(a_dots + i * 16);
                const int packed = __float_as_int(r.w);
                if (packed != 0) {
                    const int b = packed & 0xfffff;
                    const unsigned old = atomicAdd(&bins[b >> 1], 1u << ((b & 1) * 16));
                    r.w = __uint_as_float(((unsigned)(packed >> 20) << 28) | ((unsigned)q << 16) | (unsigned)i);
                    rec_s[half_of(old, b)] = r;
                }
                q += dq; p += dp;
                if (p >= P) { p -= P; ++q; }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        MSDA_STAMP(4);

        // ---- P4: pixel owners gather ----
        {
            const bool rmw = !ONE && (accumulate || chunk > 0);
            // dense coarse levels: `split` groups share one pixel and take every split-th record
            int split = 1, split_log2 = 0;
            while (split < UPW && (npix * split < (THREADS / G) * 2 || nsamp * 4 > 32 * npix * split)) {
                split <<= 1;
                ++split_log2;
            }

            // A unit = (pixel, part).  Packed bin bounds: eu = e0_up | em_up << 16, ed likewise for the
            // lower bin row, nt = n_up | total << 16 (records in the upper row / in both rows).
            struct Unit { int pix; unsigned eu, ed, nt; bool valid; };

            auto bounds = [&](const int pix, unsigned& eu, unsigned& ed, unsigned& nt) {
                int y = (int)(((float)pix + 0.5f) * inv_w);
                int x = pix - y * Wl;
                if (x < 0) { --y; x += Wl; } else if (x >= Wl) { ++y; x -= Wl; }   // float rounding guard
                const uint32_t a_dn = a_ends + (y * BW + x) * 2;   // bin (x0 = x-1, y0 = y-1); next is (x0 = x)
                const uint32_t a_up = a_dn + BW * 2;               // bin (x0 = x-1, y0 = y)
                const unsigned e0u = lds_u16(a_up), emu = lds_u16(a_up + 2), e1u = lds_u16(a_up + 4);
                const unsigned e0d = lds_u16(a_dn), emd = lds_u16(a_dn + 2), e1d = lds_u16(a_dn + 4);
                eu = e0u | (emu << 16);
                ed = e0d | (emd << 16);
                nt = (e1u - e0u) | ((e1u - e0u + e1d - e0d) << 16);
            };
            // the value row of the unit's pixel (needed for the dots only), issued ahead of its use
            auto load_v = [&](const Unit& u, const int part, uint4* raw) {
                const bool need = SMALL && u.valid && (int)(u.nt >> 16) > part;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    raw[k] = need ? ldg_nc_v4(vlevel + (uint32_t)u.pix * vrow32 + k * G * 16) : make_uint4(0, 0, 0, 0);
            };

            auto run = [&](const Unit& u, const int part, const uint4* raw) {
                const int n_up = u.nt & 0xffffu, total = u.nt >> 16;
                const int e0_up = u.eu & 0xffffu, em_up = u.eu >> 16;
                const int em_dn = u.ed >> 16;
                const int delta_dn = (int)(u.ed & 0xffffu) - n_up;
                const int mine = total > part ? (total - part + split - 1) >> split_log2 : 0;
                const int trips = __reduce_max_sync(FULL, mine);
                if (trips == 0 && (rmw || gvlevel == nullptr)) return;

                float2 v[K * E2], acc[K * E2];
#pragma unroll
                for (int k = 0; k < K; ++k) unpack2<VBF>(raw[k], v + k * E2);
#pragma unroll
                for (int c = 0; c < K * E2; ++c) acc[c] = make_float2(0.0f, 0.0f);

                int j = part;
                if constexpr (G == 4) {
                    // Cooperative decode: in a batch of 4 visits lane i of the group decodes visit i (record,
                    // weight, g-row address, dot slot); weight and address are then broadcast inside the
                    // group, and after the transpose-reduce lane i holds the dot of the visit it decoded.
                    // (the broadcast address is the byte offset of the g row; every lane adds its own a_g)
                    for (int t = 0; t < trips; t += 4, j += 4 * split) {
                        const int jm = j + lane * split;
                        float w_m = 0.0f;
                        uint32_t ga_m = qc * (VPR * 16), slot_m = 0u;             // idle slot: zero row, zero weight
                        if (jm < total) {
                            const bool up = jm < n_up;
                            const int e = jm + (up ? e0_up : delta_dn);
                            const bool isx1 = e < (up ? em_up : em_dn);     // sample sits one pixel to the left
                            const float4 rec = lds_f4(a_rec + e * 16);
                            w_m = (up ? rec.x : rec.y) * (isx1 ? rec.z : 1.0f - rec.z);
                            const unsigned id = __float_as_uint(rec.w);
                            slot_m = a_dots + (id & 0xffffu) * 16u + (up ? 0u : 8u) + (isx1 ? 4u : 0u);
                            ga_m = ((id >> 16) & 0xfffu) * (VPR * 16);
                        }
                        float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        // the visits of a batch in one fall-through switch on the (warp-uniform) number that is
                        // left: one compare chain per batch instead of a test and a branch per visit
                        auto visit = [&](const int u) {
                            const float w = __shfl_sync(FULL, w_m, u, 4);
                            const uint32_t ga = __shfl_sync(FULL, ga_m, u, 4);
                            float2 d2 = make_float2(0.0f, 0.0f);
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                float2 g[E2];
                                unpack2<VBF>(lds_u4(a_g + ga + k * G * 16), g);
#pragma unroll
                                for (int c = 0; c < E2; ++c) {
                                    if (SMALL) d2 = fma2(v[k * E2 + c], g[c], d2);
                                    acc[k * E2 + c] = fma2(g[c], make_float2(w, w), acc[k * E2 + c]);
                                }
                            }
                            d[u] = d2.x + d2.y;
                        };
                        switch (min(trips - t, 4)) {
                            case 4: visit(3);
                            case 3: visit(2);
                            case 2: visit(1);
                            default: visit(0);
                        }
                        if (SMALL) {
                            const bool hi2 = lane & 2, hi1 = lane & 1;
                            float k0 = hi2 ? d[2] : d[0], k1 = hi2 ? d[3] : d[1];
                            k0 += __shfl_xor_sync(FULL, hi2 ? d[0] : d[2], 2);
                            k1 += __shfl_xor_sync(FULL, hi2 ? d[1] : d[3], 2);
                            float keep = hi1 ? k1 : k0;
                            keep += __shfl_xor_sync(FULL, hi1 ? k0 : k1, 1);
                            if (jm < total) sts_f32(slot_m, keep);
                        }
                    }
                } else {
                    for (int t = 0; t < trips; t += NB) {
                        float d[NB];
                        uint32_t slot[NB];
#pragma unroll
                        for (int b = 0; b < NB; ++b, j += split) {
                            d[b] = 0.0f;
                            slot[b] = 0u;
                            if (j < total) {
                                const bool up = j < n_up;
                                const int e = j + (up ? e0_up : delta_dn);
                                const bool isx1 = e < (up ? em_up : em_dn);     // sample sits one pixel to the left
                                const float4 rec = lds_f4(a_rec + e * 16);
                                const float w = (up ? rec.x : rec.y) * (isx1 ? rec.z : 1.0f - rec.z);
                                const unsigned id = __float_as_uint(rec.w);
                                slot[b] = a_dots + (id & 0xffffu) * 16u + (up ? 0u : 8u) + (isx1 ? 4u : 0u);
                                const uint32_t grow = a_g + ((id >> 16) & 0xfffu) * (VPR * 16);
                                float2 d2 = make_float2(0.0f, 0.0f);
#pragma unroll
                                for (int k = 0; k < K; ++k) {
                                    float2 g[E2];
                                    unpack2<VBF>(lds_u4(grow + k * G * 16), g);
#pragma unroll
                                    for (int c = 0; c < E2; ++c) {
                                        if (SMALL) d2 = fma2(v[k * E2 + c], g[c], d2);
                                        acc[k * E2 + c] = fma2(g[c], make_float2(w, w), acc[k * E2 + c]);
                                    }
                                }
                                d[b] = d2.x + d2.y;
                            }
                        }
                        if (SMALL) {
#pragma unroll
                            for (int b = 0; b < NB; ++b) {
                                float x = d[b];
#pragma unroll
                                for (int off = G / 2; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
                                if (lane == 0 && slot[b] != 0u) sts_f32(slot[b], x);
                            }
                        }
                    }
                }
                // combine the parts of one pixel (adjacent groups of the same warp)
                for (int off = G; off < G * split; off <<= 1) {
#pragma unroll
                    for (int c = 0; c < K * E2; ++c) {
                        acc[c].x += __shfl_xor_sync(FULL, acc[c].x, off);
                        acc[c].y += __shfl_xor_sync(FULL, acc[c].y, off);
                    }
                }
                if (gvlevel != nullptr && u.valid && part == 0 && !(rmw && total == 0)) {
                    float* dst = gvlevel + (uint32_t)u.pix * gv_row32;
                    if (E == 8 && wide_store) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            float2* a2 = acc + k * E2;
                            if (rmw) {
                                float2 t0, t1, t2, t3;
                                ldg_v8(dst + k * G * E, t0, t1, t2, t3);
                                a2[0].x += t0.x; a2[0].y += t0.y; a2[1].x += t1.x; a2[1].y += t1.y;
                                a2[2].x += t2.x; a2[2].y += t2.y; a2[3].x += t3.x; a2[3].y += t3.y;
                            }
                            stg_v8(dst + k * G * E, a2[0], a2[1], a2[2], a2[3]);
                        }
                    } else
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int c = 0; c < E2; c += 2) {
                            float4* p4 = reinterpret_cast<float4*>(dst + k * G * E + c * 2);
                            float4 o = make_float4(acc[k * E2 + c].x, acc[k * E2 + c].y, acc[k * E2 + c + 1].x,
                                                   acc[k * E2 + c + 1].y);
                            if (rmw) { const float4 t = *p4; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
                            *p4 = o;
                        }
                }
            };

            if (split == 1) {
                // sparse level: warps fetch tiles of 32 pixels; each lane reads the bounds of one pixel,
                // the pixels are ranked by record count and handed to the groups in that order, so the
                // UPW pixels processed together carry similar work
                // 32 pixels per tile (one per lane), 16 when the level has fewer tiles than warps.  (Tile sizes that
                // make the tile count a multiple of the warp count -- 25 pixels for a 40 x 40 level -- were measured:
                // the wait at the barrier that ends P4 shrinks, but tiles that are not a multiple of the 8 pixels a
                // warp processes together waste more slots than that saves: 513 vs 469 us.)
                const int tpx = npix >= NWARPS * 32 ? 32 : 16;
                const int ntiles = (npix + tpx - 1) / tpx;
                for (;;) {
                    int tile = 0;
                    if (lane32 == 0) tile = atomicAdd(&work_s[0], 1);
                    tile = __shfl_sync(FULL, tile, 0);
                    if (tile >= ntiles) break;
                    const int mypix = tile * tpx + lane32;
                    unsigned eu = 0, ed = 0, nt = 0;
                    int cnt = -1;
                    if (lane32 < tpx && mypix < npix) { bounds(mypix, eu, ed, nt); cnt = min(15, (int)(nt >> 16)); }
                    const int maxc = __reduce_max_sync(FULL, cnt);
                    const unsigned lt = (1u << lane32) - 1u;
                    int base = 0;
                    for (int c = maxc; c >= 0; --c) {
                        const unsigned m = __ballot_sync(FULL, cnt == c);
                        if (cnt == c) sts_u8(a_order + base + __popc(m & lt), (unsigned)lane32);
                        base += __popc(m);
                    }
                    __syncwarp();
                    const int nvalid = base;
                    auto fetch = [&](const int u0) {
                        Unit u;
                        const int r = u0 + gsub;
                        u.valid = r < nvalid;
                        const int src = u.valid ? (int)lds_u8(a_order + r) : 0;
                        u.eu = __shfl_sync(FULL, eu, src);
                        u.ed = __shfl_sync(FULL, ed, src);
                        u.nt = __shfl_sync(FULL, nt, src);
                        if (!u.valid) u.nt = 0u;
                        u.pix = tile * tpx + src;
                        return u;
                    };
                    Unit nxt = fetch(0);
                    uint4 raw_nxt[K];
                    load_v(nxt, 0, raw_nxt);
                    for (int u0 = 0; u0 < nvalid; u0 += UPW) {
                        const Unit cur = nxt;
                        uint4 raw[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) raw[k] = raw_nxt[k];
                        if (u0 + UPW < nvalid) { nxt = fetch(u0 + UPW); load_v(nxt, 0, raw_nxt); }
                        run(cur, 0, raw);
                    }
                    __syncwarp();
                }
            } else {
                const int units = npix * split;
                const int nblocks = (units + UPW - 1) / UPW;
                for (;;) {
                    int blk = 0;
                    if (lane32 == 0) blk = atomicAdd(&work_s[0], 1);
                    blk = __shfl_sync(FULL, blk, 0);
                    if (blk >= nblocks) break;
                    const int ui = blk * UPW + gsub;
                    Unit u;
                    u.valid = ui < units;
                    u.pix = u.valid ? ui >> split_log2 : 0;
                    u.eu = u.ed = u.nt = 0u;
                    if (u.valid) bounds(u.pix, u.eu, u.ed, u.nt);
                    uint4 raw[K];
                    load_v(u, ui & (split - 1), raw);
                    run(u, ui & (split - 1), raw);
                }
            }
        }
        __syncthreads();
        MSDA_STAMP(5);

        // ---- P5: per-sample gradients from the four corner dots, walking the sorted records ----
        // (A, wy) are recovered from A*wy0 and A*wy1 (wy0 + wy1 == 1); nothing is re-read from global.
        // Results replace the dots in the sample's slot: {grad_attn, grad_x, grad_y, -}; unbinned samples
        // already hold theirs (P1).  A second sweep emits them per query with 16-byte stores: the SM's
        // write path is bound by the number of store transactions, not by bytes.
        if (SMALL) {
            const int nrec = (int)lds_u16(a_ends + nbins * 2);          // end of the last bin
            const float sW = FUSED ? 1.0f : fW, sH = FUSED ? 1.0f : fH;
            for (int e = tid; e < nrec; e += THREADS) {
                const float4 rec = rec_s[e];
                const unsigned id = __float_as_uint(rec.w);
                const int slot = id & 0xffffu;
                const float4 dd = *reinterpret_cast<const float4*>(dots_s + slot * 4);
                const float a = rec.x + rec.y;
                const float wy1 = __fdividef(rec.y, a), wy0 = 1.0f - wy1;
                const float wx1 = rec.z, wx0 = 1.0f - wx1;
                const bool vx0 = id & (1u << 28), vx1 = id & (1u << 29), vy0 = id & (1u << 30), vy1 = id & (1u << 31);
                const float d0 = (vx0 && vy0) ? dd.x : 0.0f;     // nw
                const float d1 = (vx1 && vy0) ? dd.y : 0.0f;     // ne
                const float d2 = (vx0 && vy1) ? dd.z : 0.0f;     // sw
                const float d3 = (vx1 && vy1) ? dd.w : 0.0f;     // se
                const float ga = wx0 * wy0 * d0 + wx1 * wy0 * d1 + wx0 * wy1 * d2 + wx1 * wy1 * d3;
                const float gx = (d1 - d0) * wy0 + (d3 - d2) * wy1;
                const float gy = (d2 - d0) * wx0 + (d3 - d1) * wx1;
                // d(pixel coordinate)/d(location) = (W_l, H_l); with the fused prologue the gradient is taken
                // w.r.t. the offsets, location = ref + offset / (W_l, H_l), and the factors cancel
                *reinterpret_cast<float4*>(dots_s + slot * 4) = make_float4(ga, a * sW * gx, a * sH * gy, 0.0f);
            }
            __syncthreads();
            if (pbuf != nullptr && tid == 0 && chunk == 0) pbuf[(size_t)blockIdx.x * 8 + 7] = (unsigned long long)l;
            if ((P & 3) == 0) {
                const int vq = P / 4;
                for (int i = tid; i < qc * vq; i += THREADS) {
                    const int q = i / vq, v = i - q * vq;
                    const float4* r = reinterpret_cast<const float4*>(dots_s) + q * P + v * 4;
                    const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3];
                    const int64_t sidx = s0 + q * sstride + v * 4;
                    *reinterpret_cast<float4*>(grad_attn + sidx) = make_float4(r0.x, r1.x, r2.x, r3.x);
                    // one 32-byte sector per (query, 4 samples): a single 256-bit store
                    stg_v8(reinterpret_cast<float*>(reinterpret_cast<float2*>(grad_loc) + sidx),
                           make_float2(r0.y, r0.z), make_float2(r1.y, r1.z), make_float2(r2.y, r2.z),
                           make_float2(r3.y, r3.z));
                }
            } else {
                int q = q_t0, p = p_t0;
                for (int i = tid; i < nsamp; i += THREADS) {
                    const float4 r = reinterpret_cast<const float4*>(dots_s)[i];
                    const int64_t sidx = s0 + q * sstride + p;
                    grad_attn[sidx] = r.x;
                    reinterpret_cast<float2*>(grad_loc)[sidx] = make_float2(r.y, r.z);
                    q += dq; p += dp;
                    if (p >= P) { p -= P; ++q; }
                }
            }
        }
        __syncthreads();
        MSDA_STAMP(6);
    }
#undef MSDA_STAMP
}

// reference points of the fused-prologue call in flight on this thread (set by backward_gather around the
// dispatch below; the dispatch macros stay as they are)

template <int G, int K, bool VBF, int THREADS, bool SMALL, bool ONE, bool FUSED>
static cudaError_t launch_gather_impl(const Problem& pb, const GatherPlan& plan, const void* value, const float* loc,
                                      const float* attn, const void* go, float* gv, float* gl, float* ga,
                                      int accumulate, cudaStream_t st, const float* ref, int ref_levels) {
    auto kern = bwd_gather_kernel<G, K, VBF, THREADS, SMALL, ONE, FUSED>;
    static thread_local int configured_for = -1;      // per-thread cache of the attribute call (per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_for != dev) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return e;
        configured_for = dev;
    }
    const unsigned grid = (unsigned)(pb.N * pb.H * pb.L);
    kern<<<grid, THREADS, plan.smem_bytes, st>>>(pb, (const char*)value, loc, attn, (const char*)go, gv, gl, ga,
                                                  accumulate, plan.q_chunk, plan.bins_words, ref, ref_levels, phase_buffer());
    return cudaGetLastError();
}

template <int G, int K, bool VBF, int THREADS, bool FUSED>
static cudaError_t launch_gather_f(const Problem& pb, const GatherPlan& plan, const void* value, const float* loc,
                                   const float* attn, const void* go, float* gv, float* gl, float* ga,
                                   int accumulate, cudaStream_t st, const float* ref, const int rl) {
    const bool one = plan.n_chunks == 1 && !accumulate;
    if (gl != nullptr)
        return one ? launch_gather_impl<G, K, VBF, THREADS, true, true, FUSED>(pb, plan, value, loc, attn, go, gv, gl, ga, accumulate, st, ref, rl)
                   : launch_gather_impl<G, K, VBF, THREADS, true, false, FUSED>(pb, plan, value, loc, attn, go, gv, gl, ga, accumulate, st, ref, rl);
    return launch_gather_impl<G, K, VBF, THREADS, false, false, FUSED>(pb, plan, value, loc, attn, go, gv, gl, ga, accumulate, st, ref, rl);
}

// One (lanes per pixel, vectors per lane, threads) combination, both value dtypes, with or without the fused
// prologue (ref != nullptr): explicitly instantiated in its own translation unit.
struct GatherArgs {
    const Problem* pb; const GatherPlan* plan; const void* value; const float* loc; const float* attn;
    const void* go; float* gv; float* gl; float* ga; int accumulate; cudaStream_t st; const float* ref; int ref_levels;
};

template <int G, int K, int THREADS>
cudaError_t gather_case(bool vbf, const GatherArgs& a) {
    // the fused-prologue form (reference points given) is a compile-time variant: the default kernel carries
    // none of its code or registers
#define MSDA_GC(VBF_, FUSED_)                                                                                   \
    launch_gather_f<G, K, VBF_, THREADS, FUSED_>(*a.pb, *a.plan, a.value, a.loc, a.attn, a.go, a.gv, a.gl, a.ga, \
                                                 a.accumulate, a.st, a.ref, a.ref_levels)
    if (a.ref != nullptr) return vbf ? MSDA_GC(true, true) : MSDA_GC(false, true);
    return vbf ? MSDA_GC(true, false) : MSDA_GC(false, false);
#undef MSDA_GC
}

}  // namespace msda
