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
#include "msda_kernels.cuh"

namespace msda {

namespace {

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
        off_rec = align16(qc * row_bytes);
        off_dots = off_rec + qc * P * 16;
        off_bins = off_dots + qc * P * 16;
        off_scan = off_bins + align16(bins_words * 4);
        off_order = off_scan + 64 * 4;              // per-warp visiting order of a 64-pixel tile (u16)
        total = off_order + 32 * 64 * 2;
    }
};

__device__ __forceinline__ unsigned half_of(unsigned packed, int b) { return (packed >> ((b & 1) * 16)) & 0xffffu; }

}  // namespace

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

template <int G, int K, bool VBF, int THREADS, bool SMALL>
__global__ void __launch_bounds__(THREADS, 1)
bwd_gather_kernel(const Problem pb, const char* __restrict__ value, const float* __restrict__ loc,
                  const float* __restrict__ attn, const char* __restrict__ grad_out,
                  float* __restrict__ grad_value, float* __restrict__ grad_loc,
                  float* __restrict__ grad_attn, const int accumulate, const int q_chunk,
                  const int bins_words_max) {
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
    const int l = blockIdx.x % pb.L;
    const int h = (blockIdx.x / pb.L) % pb.H;
    const int n = blockIdx.x / (pb.L * pb.H);
    const int Hl = pb.geom.h[l], Wl = pb.geom.w[l];
    const int BW = Wl + 1;                           // bins per row: x0 in [-1, W-1]
    const int nbins = BW * (Hl + 1);
    // counters are indexed by bin+1: after the scatter half k holds end(bin k-1), half 0 stays 0,
    // so begin(bin b) = ends16[b] and end(bin b) = ends16[b + 1]
    const int nwords = (nbins + 3) / 2;
    const int P = pb.P, LP = pb.L * pb.P;
    const int row_bytes = pb.Dh * ES;

    const SmemLayout lay(q_chunk, P, row_bytes, bins_words_max);
    uint4* g_s = reinterpret_cast<uint4*>(smem + lay.off_g);
    float4* rec_s = reinterpret_cast<float4*>(smem + lay.off_rec);
    float* dots_s = reinterpret_cast<float*>(smem + lay.off_dots);
    float4* tmp_s = reinterpret_cast<float4*>(smem + lay.off_dots);      // unsorted records live here until P4
    unsigned* bins = reinterpret_cast<unsigned*>(smem + lay.off_bins);
    const unsigned short* ends16 = reinterpret_cast<const unsigned short*>(bins);
    unsigned* scan_s = reinterpret_cast<unsigned*>(smem + lay.off_scan);
    unsigned char* order_s = smem + lay.off_order + warp * 32;           // lane holding the pixel of rank r
    int* work_s = reinterpret_cast<int*>(smem + lay.off_order + 32 * 32);

    const char* vlevel = value + ((int64_t)n * pb.vs_n + (int64_t)pb.geom.start[l] * pb.vs_s +
                                  (int64_t)h * pb.vs_h + lane * E) * ES;
    const int64_t vrow = pb.vs_s * ES;
    const int64_t gv_row = (int64_t)pb.H * pb.Dh;
    float* gvlevel = grad_value
        ? grad_value + ((int64_t)n * pb.S + pb.geom.start[l]) * gv_row + (int64_t)h * pb.Dh + lane * E
        : nullptr;
    const int npix = Hl * Wl;
    const float inv_w = 1.0f / (float)Wl;
    const float fW = (float)Wl, fH = (float)Hl;

    for (int q0 = 0, chunk = 0; q0 < pb.Lq; q0 += q_chunk, ++chunk) {
        const int qc = min(q_chunk, pb.Lq - q0);
        const int nsamp = qc * P;
        // sample (q, p) of this head and level sits at float2/float index s0 + q * sstride + p
        const int64_t s0 = (((int64_t)n * pb.Lq + q0) * pb.H + h) * LP + l * P;
        const int64_t sstride = (int64_t)pb.H * LP;

        // ---- P0: stage grad_out rows of this head, clear the counters ----
        {
            const char* gsrc = grad_out + (((int64_t)n * pb.Lq + q0) * pb.H * pb.Dh + (int64_t)h * pb.Dh) * ES;
            const int64_t gstride = (int64_t)pb.H * pb.Dh * ES;
            for (int i = tid; i < qc * VPR; i += THREADS) {
                const int q = i / VPR, v = i % VPR;
                g_s[i] = __ldg(reinterpret_cast<const uint4*>(gsrc + q * gstride + v * 16));
            }
            for (int i = tid; i < nwords; i += THREADS) bins[i] = 0u;
            if (tid == 0) work_s[0] = 0;
        }
        __syncthreads();

        // ---- P1: one pass over the samples: build the (unsorted) record, count per base pixel ----
        // record = {A*wy0, A*wy1, wx1, bin+1 (0: sample has no valid corner)}
        {
            int q = tid / P, p = tid - q * P;
            const int dq = THREADS / P, dp = THREADS - dq * P;
            for (int i = tid; i < nsamp; i += THREADS) {
                const int64_t sidx = s0 + q * sstride + p;
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + sidx);
                const float a = __ldg(attn + sidx);
                const Sample s = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
                int b = 0;
                if (s.x0 >= -1 && s.x0 < Wl && s.y0 >= -1 && s.y0 < Hl) {
                    b = (s.y0 + 1) * BW + (s.x0 + 1) + 1;
                    atomicAdd(&bins[b >> 1], 1u << ((b & 1) * 16));
                }
                tmp_s[i] = make_float4(a * s.wy0, a * s.wy1, s.wx1, __int_as_float(b));
                q += dq; p += dp;
                if (p >= P) { p -= P; ++q; }
            }
        }
        __syncthreads();

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

        // ---- P3: move the records into bin order (start -> end, in place) ----
        // sorted record = {A*wy0, A*wy1, wx1, (first g vector index << 16) | sample slot}
        {
            int q = tid / P;
            const int dq = THREADS / P, dp = THREADS - dq * P;
            int p = tid - q * P;
            for (int i = tid; i < nsamp; i += THREADS) {
                float4 r = tmp_s[i];
                const int b = __float_as_int(r.w);
                if (b != 0) {
                    const unsigned old = atomicAdd(&bins[b >> 1], 1u << ((b & 1) * 16));
                    r.w = __uint_as_float(((unsigned)(q * VPR) << 16) | (unsigned)i);
                    rec_s[half_of(old, b)] = r;
                }
                q += dq; p += dp;
                if (p >= P) { p -= P; ++q; }
            }
        }
        __syncthreads();

        // ---- P4: pixel owners gather ----
        {
            const bool rmw = accumulate || chunk > 0;
            // dense coarse levels: `split` groups share one pixel and take every split-th record
            int split = 1;
            while (split < UPW && (npix * split < (THREADS / G) * 2 || nsamp * 4 > 16 * npix * split)) split <<= 1;

            // One unit = (pixel, part).  eu = e0_up | em_up << 16, ed likewise for the lower bin row,
            // nt = n_up | total << 16 (records in the upper row / in both rows).
            auto process = [&](const int pix, const int part, const bool valid, const unsigned eu,
                               const unsigned ed, const unsigned nt) {
                const int n_up = nt & 0xffffu, total = nt >> 16;
                const int e0_up = eu & 0xffffu, em_up = eu >> 16;
                const int em_dn = ed >> 16;
                const int delta_dn = (int)(ed & 0xffffu) - n_up;
                const int mine = total > part ? (total - part + split - 1) / split : 0;
                const int trips = __reduce_max_sync(FULL, mine);
                if (trips == 0 && (rmw || gvlevel == nullptr)) return;

                float2 v[K * E2], acc[K * E2];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint4 raw = (SMALL && mine > 0) ? ldg_nc_v4(vlevel + (int64_t)pix * vrow + k * G * 16)
                                                          : make_uint4(0, 0, 0, 0);
                    unpack2<VBF>(raw, v + k * E2);
                }
#pragma unroll
                for (int c = 0; c < K * E2; ++c) acc[c] = make_float2(0.0f, 0.0f);

                int j = part;
                for (int t = 0; t < trips; t += NB) {
                    float d[NB];
                    unsigned slot[NB];
#pragma unroll
                    for (int u = 0; u < NB; ++u, j += split) {
                        d[u] = 0.0f;
                        slot[u] = 0xffffffffu;
                        if (j < total) {
                            const bool up = j < n_up;
                            const int e = j + (up ? e0_up : delta_dn);
                            const bool isx1 = e < (up ? em_up : em_dn);     // sample sits one pixel to the left
                            const float4 rec = rec_s[e];
                            const float w = (up ? rec.x : rec.y) * (isx1 ? rec.z : 1.0f - rec.z);
                            const unsigned id = __float_as_uint(rec.w);
                            slot[u] = (id & 0xffffu) * 4u + (up ? 0u : 2u) + (isx1 ? 1u : 0u);
                            const uint4* grow = g_s + (id >> 16) + lane;
                            float2 d2 = make_float2(0.0f, 0.0f);
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                float2 g[E2];
                                unpack2<VBF>(grow[k * G], g);
#pragma unroll
                                for (int c = 0; c < E2; ++c) {
                                    if (SMALL) d2 = fma2(v[k * E2 + c], g[c], d2);
                                    acc[k * E2 + c] = fma2(g[c], make_float2(w, w), acc[k * E2 + c]);
                                }
                            }
                            d[u] = d2.x + d2.y;
                        }
                    }
                    if (SMALL) {
                        if constexpr (G == 4) {
                            // 4 visits x 4 lanes: transpose-reduce, lane i ends with the dot of visit i
                            const bool hi2 = lane & 2, hi1 = lane & 1;
                            float k0 = hi2 ? d[2] : d[0], k1 = hi2 ? d[3] : d[1];
                            k0 += __shfl_xor_sync(FULL, hi2 ? d[0] : d[2], 2);
                            k1 += __shfl_xor_sync(FULL, hi2 ? d[1] : d[3], 2);
                            float keep = hi1 ? k1 : k0;
                            keep += __shfl_xor_sync(FULL, hi1 ? k0 : k1, 1);
                            const unsigned sl = hi2 ? (hi1 ? slot[3] : slot[2]) : (hi1 ? slot[1] : slot[0]);
                            if (sl != 0xffffffffu) dots_s[sl] = keep;
                        } else {
#pragma unroll
                            for (int u = 0; u < NB; ++u) {
                                float x = d[u];
#pragma unroll
                                for (int off = G / 2; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
                                if (lane == 0 && slot[u] != 0xffffffffu) dots_s[slot[u]] = x;
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
                if (gvlevel != nullptr && valid && part == 0 && !(rmw && total == 0)) {
                    float* dst = gvlevel + (int64_t)pix * gv_row;
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

            // bin bounds of one pixel, packed (see process)
            auto bounds = [&](const int pix, unsigned& eu, unsigned& ed, unsigned& nt) {
                int y = (int)(((float)pix + 0.5f) * inv_w);
                int x = pix - y * Wl;
                if (x < 0) { --y; x += Wl; } else if (x >= Wl) { ++y; x -= Wl; }   // float rounding guard
                const int b_dn = y * BW + x;               // bin (x0 = x-1, y0 = y-1); +1 is (x0 = x)
                const int b_up = b_dn + BW;                // bin (x0 = x-1, y0 = y)
                const unsigned e0u = ends16[b_up], emu = ends16[b_up + 1], e1u = ends16[b_up + 2];
                const unsigned e0d = ends16[b_dn], emd = ends16[b_dn + 1], e1d = ends16[b_dn + 2];
                eu = e0u | (emu << 16);
                ed = e0d | (emd << 16);
                nt = (e1u - e0u) | ((e1u - e0u + e1d - e0d) << 16);
            };

            if (split == 1) {
                // sparse level: warps fetch tiles of 32 pixels; each lane reads the bounds of one pixel,
                // the pixels are ranked by record count and handed to the groups in that order, so the
                // UPW pixels processed together carry similar work
                const int ntiles = (npix + 31) / 32;
                for (;;) {
                    int tile = 0;
                    if (lane32 == 0) tile = atomicAdd(&work_s[0], 1);
                    tile = __shfl_sync(FULL, tile, 0);
                    if (tile >= ntiles) break;
                    const int mypix = tile * 32 + lane32;
                    unsigned eu = 0, ed = 0, nt = 0;
                    int cnt = -1;
                    if (mypix < npix) { bounds(mypix, eu, ed, nt); cnt = min(15, (int)(nt >> 16)); }
                    const int maxc = __reduce_max_sync(FULL, cnt);
                    const unsigned lt = (1u << lane32) - 1u;
                    int base = 0;
                    for (int c = maxc; c >= 0; --c) {
                        const unsigned m = __ballot_sync(FULL, cnt == c);
                        if (cnt == c) order_s[base + __popc(m & lt)] = (unsigned char)lane32;
                        base += __popc(m);
                    }
                    __syncwarp();
                    const int nvalid = base;
                    for (int u0 = 0; u0 < nvalid; u0 += UPW) {
                        const int u = u0 + gsub;
                        const bool valid = u < nvalid;
                        const int src = valid ? (int)order_s[u] : 0;
                        const unsigned seu = __shfl_sync(FULL, eu, src);
                        const unsigned sed = __shfl_sync(FULL, ed, src);
                        const unsigned snt = __shfl_sync(FULL, nt, src);
                        process(tile * 32 + src, 0, valid, seu, sed, valid ? snt : 0u);
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
                    const int u = blk * UPW + gsub;
                    const bool valid = u < units;
                    const int pix = valid ? u / split : 0;
                    unsigned eu = 0, ed = 0, nt = 0;
                    if (valid) bounds(pix, eu, ed, nt);
                    process(pix, u & (split - 1), valid, eu, ed, nt);
                }
            }
        }
        __syncthreads();

        // ---- P5: per-sample gradients from the four corner dots ----
        if (SMALL) {
            int q = tid / P, p = tid - q * P;
            const int dq = THREADS / P, dp = THREADS - dq * P;
            for (int i = tid; i < nsamp; i += THREADS) {
                const int64_t sidx = s0 + q * sstride + p;
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + sidx);
                const float a = __ldg(attn + sidx);
                const Sample s = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
                const float4 dd = *reinterpret_cast<const float4*>(dots_s + i * 4);
                const float d0 = (s.vx0 && s.vy0) ? dd.x : 0.0f;     // nw
                const float d1 = (s.vx1 && s.vy0) ? dd.y : 0.0f;     // ne
                const float d2 = (s.vx0 && s.vy1) ? dd.z : 0.0f;     // sw
                const float d3 = (s.vx1 && s.vy1) ? dd.w : 0.0f;     // se
                const float ga = s.w_nw * d0 + s.w_ne * d1 + s.w_sw * d2 + s.w_se * d3;
                const float gx = (d1 - d0) * s.wy0 + (d3 - d2) * s.wy1;
                const float gy = (d2 - d0) * s.wx0 + (d3 - d1) * s.wx1;
                grad_attn[sidx] = ga;
                reinterpret_cast<float2*>(grad_loc)[sidx] = make_float2(a * fW * gx, a * fH * gy);
                q += dq; p += dp;
                if (p >= P) { p -= P; ++q; }
            }
        }
        __syncthreads();
    }
}

template <int G, int K, bool VBF, int THREADS, bool SMALL>
static cudaError_t launch_gather_impl(const Problem& pb, const GatherPlan& plan, const void* value, const float* loc,
                                      const float* attn, const void* go, float* gv, float* gl, float* ga,
                                      int accumulate, cudaStream_t st) {
    auto kern = bwd_gather_kernel<G, K, VBF, THREADS, SMALL>;
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
                                                  accumulate, plan.q_chunk, plan.bins_words);
    return cudaGetLastError();
}

template <int G, int K, bool VBF, int THREADS>
static cudaError_t launch_gather(const Problem& pb, const GatherPlan& plan, const void* value, const float* loc,
                                 const float* attn, const void* go, float* gv, float* gl, float* ga,
                                 int accumulate, cudaStream_t st) {
    return gl != nullptr
        ? launch_gather_impl<G, K, VBF, THREADS, true>(pb, plan, value, loc, attn, go, gv, gl, ga, accumulate, st)
        : launch_gather_impl<G, K, VBF, THREADS, false>(pb, plan, value, loc, attn, go, gv, gl, ga, accumulate, st);
}

// Returns false when the shape does not fit the gather kernel (caller falls back to the flat one).
static bool make_plan(const Problem& pb, int row_bytes, GatherPlan& plan) {
    int max_bins = 0;
    for (int l = 0; l < pb.L; ++l) max_bins = max(max_bins, (pb.geom.w[l] + 1) * (pb.geom.h[l] + 1));
    plan.bins_words = (max_bins + 3) / 2;
    const int fixed = align16(plan.bins_words * 4) + 64 * 4 + 32 * 64 * 2 + 16;
    const int per_query = row_bytes + pb.P * 32;
    int qmax = (kMaxSmem - fixed) / per_query;
    qmax = min(qmax, 65535 / pb.P);                  // u16 counters and ids
    if (qmax < 32) return false;
    plan.n_chunks = (pb.Lq + qmax - 1) / qmax;
    plan.q_chunk = (pb.Lq + plan.n_chunks - 1) / plan.n_chunks;
    plan.smem_bytes = SmemLayout(plan.q_chunk, pb.P, row_bytes, plan.bins_words).total;
    return plan.smem_bytes <= (size_t)kMaxSmem;
}

bool backward_gather_supported(const Problem& pb, bool value_bf16) {
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;
    if (!(nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6 || nv == 8 || nv == 12 || nv == 16)) return false;
    if ((int64_t)pb.N * pb.H * pb.L > 0x7fffffffLL) return false;
    GatherPlan plan;
    return make_plan(pb, pb.Dh * (value_bf16 ? 2 : 4), plan);
}

cudaError_t backward_gather(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                            const float* attn, const void* grad_out, float* grad_value, float* grad_loc,
                            float* grad_attn, int accumulate, cudaStream_t st) {
    GatherPlan plan;
    if (!make_plan(pb, pb.Dh * (value_bf16 ? 2 : 4), plan)) return cudaErrorInvalidValue;
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;
#define MSDA_GATHER_CASE(NV, G, K, T)                                                                       \
    case NV:                                                                                                \
        return value_bf16 ? launch_gather<G, K, true, T>(pb, plan, value, loc, attn, grad_out, grad_value,   \
                                                         grad_loc, grad_attn, accumulate, st)               \
                          : launch_gather<G, K, false, T>(pb, plan, value, loc, attn, grad_out, grad_value,  \
                                                          grad_loc, grad_attn, accumulate, st);
    switch (nv) {
        MSDA_GATHER_CASE(1, 1, 1, 1024)
        MSDA_GATHER_CASE(2, 2, 1, 1024)
        MSDA_GATHER_CASE(3, 1, 3, 512)
        MSDA_GATHER_CASE(4, 4, 1, 1024)
        MSDA_GATHER_CASE(6, 2, 3, 512)
        MSDA_GATHER_CASE(8, 8, 1, 1024)
        MSDA_GATHER_CASE(12, 4, 3, 512)
        MSDA_GATHER_CASE(16, 8, 2, 512)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_GATHER_CASE
}

}  // namespace msda
