// Forward, staged variant: one CTA owns one (image n, head h) and a slice of its queries.  The
// coarse pyramid levels of that head (every level from `first_staged` on: 40x40 + 20x20 = 128 KB
// of bf16 rows for the DETRPose shapes) are brought into shared memory by TMA
// (cp.async.bulk.tensor on a 4-D tensor map of the strided value pyramid, completion on an
// mbarrier) while the CTA computes its first batch of sample parameters; 2/3 of the corner rows
// then come from shared memory and only the finest level is gathered through L2.
// Same arithmetic as fwd_lean_kernel (msda_fwd.cu); replaces ms_deform_attn_core_pytorch forward,
// /root/reference/src/models/detrpose/ms_deform_attn.py:145-193.
#include <cuda.h>

#include "msda_kernels.cuh"

namespace msda {

namespace {

constexpr int kBoxRows = 64;                 // pixel rows per TMA box
constexpr int kMaxSmemStaged = 227 * 1024;

struct StagePlan {
    int first_staged;                        // levels >= first_staged live in shared memory
    int parts;                               // query slices per (n, h)
    int q_per_part;
    int32_t smem_off[MSDA_MAX_LEVELS];       // byte offset of a staged level inside the staging area
    int stage_bytes;                         // staging area (whole boxes)
    int n_boxes;
};

__device__ __forceinline__ uint4 lds_u4s(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

}  // namespace

template <int G, int K, bool VBF, bool OBF, int kStagedThreads, int kMinBlocks>
__global__ void __launch_bounds__(kStagedThreads, kMinBlocks)
fwd_staged_kernel(const __grid_constant__ CUtensorMap tmap, const Problem pb, const StagePlan plan,
                  const char* __restrict__ value, const float* __restrict__ loc,
                  const float* __restrict__ attn, char* __restrict__ out) {
    constexpr int E = Vec<VBF>::kElems;
    constexpr int E2 = E / 2;
    constexpr int ES = VBF ? 2 : 4;
    constexpr int IPW = 32 / G;                      // items (queries of this head) a warp handles at a time
    constexpr int NWARPS = kStagedThreads / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int lane32 = tid & 31, wid = tid >> 5;
    const int LP = pb.L * pb.P;
    const int part = blockIdx.x % plan.parts;
    const int h = (blockIdx.x / plan.parts) % pb.H;
    const int n = blockIdx.x / (plan.parts * pb.H);
    const int q_begin = part * plan.q_per_part;
    const int q_end = min(pb.Lq, q_begin + plan.q_per_part);
    const uint32_t row_bytes = (uint32_t)(pb.vs_s * ES);          // global row pitch
    const uint32_t srow_bytes = (uint32_t)(pb.Dh * ES);           // staged row pitch (dense)

    const int item_stride = LP * 32 + 16;
    unsigned char* params = smem_raw + plan.stage_bytes + wid * (IPW * item_stride);   // this warp's table
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bar = sbase + plan.stage_bytes + NWARPS * IPW * item_stride;
    int* work = reinterpret_cast<int*>(smem_raw + plan.stage_bytes + NWARPS * IPW * item_stride + 8);

    // ---- start the TMA staging of the coarse levels of (n, h) ----
    if (tid == 0) {
        *work = 0;
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)plan.n_boxes * kBoxRows * srow_bytes);
        for (int l = plan.first_staged; l < pb.L; ++l) {
            const int rows = pb.geom.h[l] * pb.geom.w[l];
            for (int r = 0; r < rows; r += kBoxRows)
                tma_load_4d(sbase + plan.smem_off[l] + r * srow_bytes, &tmap, 0, h, pb.geom.start[l] + r, n, bar);
        }
    }
    __syncthreads();                                  // barrier and work counter initialised

    const int lane = tid % G;
    const int il = lane32 / G;
    const char* vbase = value + ((int64_t)n * pb.vs_n + (int64_t)h * pb.vs_h + lane * E) * ES;
    const uint32_t a_stage = sbase + lane * E * ES;
    bool staged_ready = false;

    // Warps work independently (no CTA-wide barrier after this point): each fetches groups of IPW queries,
    // builds the sample parameters of its group in its own table and gathers, so that the parameter
    // loads, the L2 gathers of the finest level and the shared-memory gathers of different warps overlap.
    for (;;) {
        int grp = 0;
        if (lane32 == 0) grp = atomicAdd(work, 1);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        const int q0 = q_begin + grp * IPW;
        if (q0 >= q_end) break;
        const int nitems = min(IPW, q_end - q0);
        // ---- phase 1: the lanes of the warp share the samples of its group ----
        {
            const int nsamp = nitems * LP;
            for (int s = lane32; s < nsamp; s += 32) {
                const int i_l = s / LP, sl = s - i_l * LP;
                const int64_t sidx = (((int64_t)n * pb.Lq + q0 + i_l) * pb.H + h) * LP + sl;
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + sidx);
                const float a = __ldg(attn + sidx);
                const int l = sl / pb.P;
                const int Hl = pb.geom.h[l], Wl = pb.geom.w[l];
                const Sample sm = make_sample(xy.x, xy.y, Hl, Wl, pb.coord_mode);
                const int xc0 = min(max(sm.x0, 0), Wl - 1), xc1 = min(max(sm.x0 + 1, 0), Wl - 1);
                const int yc0 = min(max(sm.y0, 0), Hl - 1), yc1 = min(max(sm.y0 + 1, 0), Hl - 1);
                uint32_t base, pitch;
                if (l >= plan.first_staged) { base = (uint32_t)plan.smem_off[l]; pitch = srow_bytes; }
                else { base = (uint32_t)pb.geom.start[l] * row_bytes; pitch = row_bytes; }
                const uint32_t r0 = base + (uint32_t)(yc0 * Wl) * pitch, r1 = base + (uint32_t)(yc1 * Wl) * pitch;
                unsigned char* dst = params + i_l * item_stride + sl * 32;
                reinterpret_cast<float4*>(dst)[0] = make_float4(sm.w_nw * a, sm.w_ne * a, sm.w_sw * a, sm.w_se * a);
                // Corners outside the map keep an exact zero weight and point at the clamped pixel, which is a
                // valid corner of the same sample whenever the sample touches the map at all: no predicated
                // loads in phase 2.  (A sample entirely outside multiplies a border row by zero.)
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(r0 + xc0 * pitch, r0 + xc1 * pitch,
                                                              r1 + xc0 * pitch, r1 + xc1 * pitch);
            }
        }
        __syncwarp();
        if (!staged_ready) { mbar_wait(bar, 0); staged_ready = true; }

        // ---- phase 2: G lanes per item gather (shared memory for staged levels) and accumulate ----
        if (il < nitems) {
            float2 acc[K * E2];
#pragma unroll
            for (int c = 0; c < K * E2; ++c) acc[c] = make_float2(0.0f, 0.0f);
            const unsigned char* ip = params + il * item_stride;
            // one pair of samples: 8 row loads issued before the first use, then FFMA2 accumulation
            auto accumulate = [&](const uint4 (&raw)[2][K][4], const float (&wv)[2][4]) {
#pragma unroll
                for (int b = 0; b < 2; ++b)
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
                            for (int e = 0; e < E2; ++e) acc[k * E2 + e] = __ffma2_rn(f[e], ww, acc[k * E2 + e]);
                        }
            };
            auto pair = [&](const int s, const bool two, auto load_row) {
                float wv[2][4];
                uint4 raw[2][K][4];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    // an odd tail repeats the previous sample with zero weights
                    const unsigned char* e = ip + (s + ((b == 0 || two) ? b : 0)) * 32;
                    const float4 w = reinterpret_cast<const float4*>(e)[0];
                    const uint4 o = reinterpret_cast<const uint4*>(e)[1];
                    const bool live = b == 0 || two;
                    wv[b][0] = live ? w.x : 0.0f; wv[b][1] = live ? w.y : 0.0f;
                    wv[b][2] = live ? w.z : 0.0f; wv[b][3] = live ? w.w : 0.0f;
                    const uint32_t ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int c = 0; c < 4; ++c) raw[b][k][c] = load_row(ov[c] + k * G * 16);
                }
                accumulate(raw, wv);
            };
            for (int l = 0; l < pb.L; ++l) {
                if (l >= plan.first_staged) {                     // warp-uniform: rows of this level are staged
                    for (int p = 0; p < pb.P; p += 2)
                        pair(l * pb.P + p, p + 1 < pb.P, [&](const uint32_t o) { return lds_u4s(a_stage + o); });
                } else {
                    for (int p = 0; p < pb.P; p += 2)
                        pair(l * pb.P + p, p + 1 < pb.P, [&](const uint32_t o) { return ldg_nc_v4(vbase + o); });
                }
            }
            constexpr int OS = OBF ? 2 : 4;
            const int64_t item = ((int64_t)n * pb.Lq + q0 + il) * pb.H + h;
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
        __syncwarp();                                 // the table is overwritten by the warp's next group
    }
    if (!staged_ready) mbar_wait(bar, 0);             // never leave with the bulk copies in flight
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// `small`: 256-thread CTAs that stage only what fits beside three co-resident CTAs (the coarsest
// level for the DETRPose shapes) instead of one 768-thread CTA per SM staging everything it can.
bool make_stage_plan(const Problem& pb, bool value_bf16, int G, bool small, StagePlan& plan) {
    const int es = value_bf16 ? 2 : 4;
    const int srow = pb.Dh * es;
    const int threads = small ? 256 : 768;
    const int ipc = threads / G;
    const int params_bytes = ipc * (pb.L * pb.P * 32 + 16) + 16;      // one table of 32 / G items per warp
    const int budget = (small ? kMaxSmemStaged / 4 : kMaxSmemStaged) - params_bytes - 256;
    // stage the longest suffix of levels that fits
    int first = pb.L, bytes = 0;
    for (int l = pb.L - 1; l >= 0; --l) {
        const int rows = pb.geom.h[l] * pb.geom.w[l];
        const int b = ((rows + kBoxRows - 1) / kBoxRows) * kBoxRows * srow;
        if (bytes + b > budget) break;
        bytes += b;
        first = l;
    }
    if (first >= pb.L) return false;
    plan.first_staged = first;
    int off = 0, boxes = 0;
    for (int l = 0; l < MSDA_MAX_LEVELS; ++l) plan.smem_off[l] = 0;
    for (int l = first; l < pb.L; ++l) {
        const int rows = pb.geom.h[l] * pb.geom.w[l];
        const int nb = (rows + kBoxRows - 1) / kBoxRows;
        plan.smem_off[l] = off;
        off += nb * kBoxRows * srow;
        boxes += nb;
    }
    plan.stage_bytes = off;                      // multiple of kBoxRows * srow (>= 1 KB): keeps 128-byte alignment
    plan.n_boxes = boxes;
    if (small) {             // one pass of queries per CTA
        plan.parts = (pb.Lq + ipc - 1) / ipc;
        plan.q_per_part = ipc;
    } else {                 // enough CTAs for ~4 waves of 148 SMs, but at least one full pass per CTA
        int parts = 1;
        while ((int64_t)pb.N * pb.H * parts < 592 && pb.Lq / (parts * 2) >= ipc) parts *= 2;
        plan.parts = parts;
        plan.q_per_part = (pb.Lq + parts - 1) / parts;
    }
    return true;
}

template <int G, int K, bool VBF, int THREADS, int MINB>
cudaError_t launch_staged_t(const Problem& pb, const StagePlan& plan, const CUtensorMap& tmap, const void* value,
                            const float* loc, const float* attn, void* out, bool out_bf16, cudaStream_t st) {
    constexpr int IPC = THREADS / G;
    const size_t smem = (size_t)plan.stage_bytes + (size_t)IPC * (pb.L * pb.P * 32 + 16) + 16;
    const unsigned grid = (unsigned)(pb.N * pb.H * plan.parts);
    auto launch = [&](auto kern) -> cudaError_t {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemStaged);
        if (e != cudaSuccess) return e;
        kern<<<grid, THREADS, smem, st>>>(tmap, pb, plan, (const char*)value, loc, attn, (char*)out);
        return cudaGetLastError();
    };
    return out_bf16 ? launch(fwd_staged_kernel<G, K, VBF, true, THREADS, MINB>)
                    : launch(fwd_staged_kernel<G, K, VBF, false, THREADS, MINB>);
}

template <int G, int K, bool VBF>
cudaError_t launch_staged(const Problem& pb, const StagePlan& plan, bool small, const CUtensorMap& tmap,
                          const void* value, const float* loc, const float* attn, void* out, bool out_bf16,
                          cudaStream_t st) {
    return small ? launch_staged_t<G, K, VBF, 256, 4>(pb, plan, tmap, value, loc, attn, out, out_bf16, st)
                 : launch_staged_t<G, K, VBF, 768, 1>(pb, plan, tmap, value, loc, attn, out, out_bf16, st);
}

int lanes_for(int nv) { return nv == 3 ? 1 : nv == 6 ? 2 : nv == 12 ? 4 : nv == 16 ? 8 : nv; }

}  // namespace

bool forward_staged_supported(const Problem& pb, bool value_bf16, bool small) {
    const int es = value_bf16 ? 2 : 4;
    const int nv = pb.Dh * es / 16;
    if (!(nv == 2 || nv == 4 || nv == 6 || nv == 8)) return false;
    if (get_encode_fn() == nullptr) return false;
    if ((int64_t)pb.S * pb.vs_s * es >= (int64_t)0x7fffffff) return false;
    if (pb.Dh > 256 || pb.S < 1) return false;
    StagePlan plan;
    return make_stage_plan(pb, value_bf16, lanes_for(nv), small, plan);
}

cudaError_t forward_staged(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                           const float* attn, void* out, bool out_bf16, bool small, cudaStream_t st) {
    const int es = value_bf16 ? 2 : 4;
    const int nv = pb.Dh * es / 16;
    StagePlan plan;
    if (!make_stage_plan(pb, value_bf16, lanes_for(nv), small, plan)) return cudaErrorInvalidValue;
    // 4-D tensor map of the strided pyramid: (channel, head, pixel, image)
    CUtensorMap tmap;
    const cuuint64_t dims[4] = {(cuuint64_t)pb.Dh, (cuuint64_t)pb.H, (cuuint64_t)pb.S, (cuuint64_t)pb.N};
    // a single image may arrive with an arbitrary (even zero) image stride: any valid value will do
    const cuuint64_t img_stride = pb.N > 1 ? (cuuint64_t)pb.vs_n * es : (cuuint64_t)pb.S * pb.vs_s * es;
    const cuuint64_t strides[3] = {(cuuint64_t)pb.vs_h * es, (cuuint64_t)pb.vs_s * es, img_stride};
    const cuuint32_t box[4] = {(cuuint32_t)pb.Dh, 1u, (cuuint32_t)kBoxRows, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = get_encode_fn()(&tmap, value_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                       4, const_cast<void*>(value), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
#define MSDA_STAGED_CASE(NV, G, K)                                                                        \
    case NV:                                                                                              \
        return value_bf16 ? launch_staged<G, K, true>(pb, plan, small, tmap, value, loc, attn, out, out_bf16, st)  \
                          : launch_staged<G, K, false>(pb, plan, small, tmap, value, loc, attn, out, out_bf16, st);
    switch (nv) {
        MSDA_STAGED_CASE(2, 2, 1)
        MSDA_STAGED_CASE(4, 4, 1)
        MSDA_STAGED_CASE(6, 2, 3)
        MSDA_STAGED_CASE(8, 8, 1)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_STAGED_CASE
}

}  // namespace msda
