// Auxiliary kernels around the sampler: integer index export, the
// location/softmax prologue, and the layout repack between the reference's
// per-level strided views and the kernel-native channel-last pyramid.
#include "msda_kernels.cuh"

namespace msda {

// ---------------------------------------------------------------------------
// Index export (parity instrument): (y0, x0) per sample, exactly as the
// forward/backward kernels derive them (same make_sample()).
// ---------------------------------------------------------------------------
__global__ void sample_indices_kernel(const Problem pb, const float* __restrict__ loc,
                                      int32_t* __restrict__ idx_out, int32_t* __restrict__ level_start_out) {
    const int64_t total = (int64_t)pb.N * pb.Lq * pb.H * pb.L * pb.P;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (level_start_out != nullptr && i < pb.L) level_start_out[i] = pb.geom.start[i];
    if (i >= total) return;
    const int l = (int)((i / pb.P) % pb.L);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + i);
    const Sample s = make_sample(xy.x, xy.y, pb.geom.h[l], pb.geom.w[l], pb.coord_mode);
    reinterpret_cast<int2*>(idx_out)[i] = make_int2(s.y0, s.x0);
}

cudaError_t sample_indices(const Problem& pb, const float* loc, int32_t* idx_out,
                           int32_t* level_start_out, cudaStream_t st) {
    const int64_t total = (int64_t)pb.N * pb.Lq * pb.H * pb.L * pb.P;
    const unsigned grid = (unsigned)((total + 255) / 256);
    sample_indices_kernel<<<grid, 256, 0, st>>>(pb, loc, idx_out, level_start_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Prologue: attention = softmax over L*P; locations = ref + offsets / (W_l, H_l)
// ms_deform_attn.py:392-393, :412-416.  One thread per (n, q, h).
// The divide and the add are separately rounded IEEE ops, like the reference's
// elementwise torch ops, so the resulting locations are bit-identical.
// ---------------------------------------------------------------------------
__global__ void locations_kernel(const Problem pb, const float* __restrict__ offsets,
                                 const float* __restrict__ logits, const float* __restrict__ ref,
                                 int ref_levels, float* __restrict__ loc, float* __restrict__ attn) {
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= items) return;
    const int LP = pb.L * pb.P;
    const int64_t nq = item / pb.H;

    const float* lg = logits + item * LP;
    float m = -INFINITY;
    for (int i = 0; i < LP; ++i) m = fmaxf(m, __ldg(lg + i));
    float sum = 0.0f;
    for (int i = 0; i < LP; ++i) sum += expf(__ldg(lg + i) - m);
    float* ao = attn + item * LP;
    for (int i = 0; i < LP; ++i) ao[i] = expf(__ldg(lg + i) - m) / sum;

    const float2* off = reinterpret_cast<const float2*>(offsets) + item * LP;
    float2* lo = reinterpret_cast<float2*>(loc) + item * LP;
    for (int l = 0; l < pb.L; ++l) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(ref) + nq * ref_levels + (ref_levels == 1 ? 0 : l));
        const float wl = (float)pb.geom.w[l], hl = (float)pb.geom.h[l];
        for (int p = 0; p < pb.P; ++p) {
            const float2 o = __ldg(off + l * pb.P + p);
            lo[l * pb.P + p] = make_float2(__fadd_rn(r.x, __fdiv_rn(o.x, wl)),
                                           __fadd_rn(r.y, __fdiv_rn(o.y, hl)));
        }
    }
}

cudaError_t locations(const Problem& pb, const float* offsets, const float* logits, const float* ref,
                      int ref_levels, float* loc, float* attn, cudaStream_t st) {
    const int64_t items = (int64_t)pb.N * pb.Lq * pb.H;
    const unsigned grid = (unsigned)((items + 127) / 128);
    locations_kernel<<<grid, 128, 0, st>>>(pb, offsets, logits, ref, ref_levels, loc, attn);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Layout repack.  Source: value[l] (N*H, Dh, H_l*W_l) with arbitrary strides
// (transformer.py:1285-1286 hands over spatial-innermost views for N > 1).
// Destination: channel-last pyramid (N, S, H, Dh).  Tiled transpose through
// shared memory: 32 spatial positions x Dh channels per block.
// ---------------------------------------------------------------------------
constexpr int kTileS = 32;

struct TileMap {
    int32_t first_tile[MSDA_MAX_LEVELS + 1];   // running sum of ceil(HW_l / kTileS)
};

template <bool BF> __device__ __forceinline__ float load_elem(const void* p, int64_t i) {
    if constexpr (BF) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    else return reinterpret_cast<const float*>(p)[i];
}
template <bool BF> __device__ __forceinline__ void store_elem(void* p, int64_t i, float v) {
    if constexpr (BF) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(p)[i] = v;
}

template <bool SBF, bool DBF>
__global__ void __launch_bounds__(256)
repack_kernel(const Problem pb, const LevelViews src, const TileMap tm, void* __restrict__ dst) {
    extern __shared__ float tile[];               // [Dh][kTileS + 1]
    const int nh = blockIdx.y;
    const int n = nh / pb.H, h = nh % pb.H;
    int l = 0;
    while (l + 1 < pb.L && (int)blockIdx.x >= tm.first_tile[l + 1]) ++l;
    const int s0 = ((int)blockIdx.x - tm.first_tile[l]) * kTileS;
    const int hw = pb.geom.h[l] * pb.geom.w[l];
    const int tx = threadIdx.x % kTileS, ty = threadIdx.x / kTileS;      // 32 x 8

    if (s0 + tx < hw) {
        const int64_t sbase = (int64_t)nh * src.s_nh[l] + (int64_t)(s0 + tx) * src.s_s[l];
        for (int c = ty; c < pb.Dh; c += 8)
            tile[c * (kTileS + 1) + tx] = load_elem<SBF>(src.ptr[l], sbase + (int64_t)c * src.s_c[l]);
    }
    __syncthreads();
    // write: channel fastest
    const int64_t drow = (int64_t)pb.H * pb.Dh;
    for (int i = threadIdx.x; i < kTileS * pb.Dh; i += 256) {
        const int s = i / pb.Dh, c = i % pb.Dh;
        if (s0 + s < hw) {
            const int64_t d = ((int64_t)n * pb.S + pb.geom.start[l] + s0 + s) * drow + (int64_t)h * pb.Dh + c;
            store_elem<DBF>(dst, d, tile[c * (kTileS + 1) + s]);
        }
    }
}

template <bool DBF>
__global__ void __launch_bounds__(256)
unpack_grad_kernel(const Problem pb, const float* __restrict__ gv, const LevelViews dst, const TileMap tm) {
    extern __shared__ float tile[];               // [Dh][kTileS + 1]
    const int nh = blockIdx.y;
    const int n = nh / pb.H, h = nh % pb.H;
    int l = 0;
    while (l + 1 < pb.L && (int)blockIdx.x >= tm.first_tile[l + 1]) ++l;
    const int s0 = ((int)blockIdx.x - tm.first_tile[l]) * kTileS;
    const int hw = pb.geom.h[l] * pb.geom.w[l];
    const int64_t srow = (int64_t)pb.H * pb.Dh;
    for (int i = threadIdx.x; i < kTileS * pb.Dh; i += 256) {
        const int s = i / pb.Dh, c = i % pb.Dh;
        if (s0 + s < hw)
            tile[c * (kTileS + 1) + s] =
                gv[((int64_t)n * pb.S + pb.geom.start[l] + s0 + s) * srow + (int64_t)h * pb.Dh + c];
    }
    __syncthreads();
    const int tx = threadIdx.x % kTileS, ty = threadIdx.x / kTileS;
    if (s0 + tx < hw) {
        const int64_t dbase = (int64_t)nh * dst.s_nh[l] + (int64_t)(s0 + tx) * dst.s_s[l];
        for (int c = ty; c < pb.Dh; c += 8)
            store_elem<DBF>(const_cast<void*>(dst.ptr[l]), dbase + (int64_t)c * dst.s_c[l],
                            tile[c * (kTileS + 1) + tx]);
    }
}

// ---------------------------------------------------------------------------
// Vectorised repack / un-repack for the layout the reference actually hands over for N > 1
// (spatial stride 1, 16-byte aligned rows): 64 positions x Dh channels per block, 16-byte global
// accesses on both sides, fp32 staging tile in shared memory with a 65-word pitch (conflict-free
// on the transposed side).  Anything else takes the scalar kernels above.
// ---------------------------------------------------------------------------
constexpr int kVecTileS = 64;

struct TileMap64 { int32_t first_tile[MSDA_MAX_LEVELS + 1]; };

template <bool BF> struct ElemVec { static constexpr int n = BF ? 8 : 4; };

template <bool BF>
__device__ __forceinline__ void vec_to_floats(const uint4& v, float* f) {
    if constexpr (BF) {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    } else {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
}
template <bool BF>
__device__ __forceinline__ uint4 floats_to_vec(const float* f) {
    if constexpr (BF)
        return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    else
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}

// U sub-tiles of 64 positions per block: a thread issues the loads of all of them before the first conversion,
// so that 16 * U bytes per thread are in flight (one 16-byte load per thread and barrier left the copy bound by
// latency: 3.5 TB/s at U = 1, tools/bench_repack.py).
template <bool SBF, bool DBF, int U>
__global__ void __launch_bounds__(256)
repack_vec_kernel(const Problem pb, const LevelViews src, const TileMap64 tm, char* __restrict__ dst) {
    extern __shared__ float tile[];                   // [U][Dh][65]
    constexpr int VS = ElemVec<SBF>::n, VD = ElemVec<DBF>::n, PITCH = kVecTileS + 1;
    const int nh = blockIdx.y, n = nh / pb.H, h = nh % pb.H;
    int l = 0;
    while (l + 1 < pb.L && (int)blockIdx.x >= tm.first_tile[l + 1]) ++l;
    const int s00 = ((int)blockIdx.x - tm.first_tile[l]) * (kVecTileS * U);
    const int hw = pb.geom.h[l] * pb.geom.w[l];
    const char* sp = reinterpret_cast<const char*>(src.ptr[l]);
    constexpr int SES = SBF ? 2 : 4, DES = DBF ? 2 : 4;
    const int sub = pb.Dh * PITCH;                    // floats per sub-tile
    // load: channel rows, VS positions per 16-byte vector
    for (int idx = threadIdx.x; idx < pb.Dh * (kVecTileS / VS); idx += 256) {
        const int c = idx / (kVecTileS / VS), v = idx % (kVecTileS / VS);
        const int64_t e0 = (int64_t)nh * src.s_nh[l] + (int64_t)c * src.s_c[l];
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s00 + u * kVecTileS + v * VS;
            if (s + VS <= hw) raw[u] = __ldg(reinterpret_cast<const uint4*>(sp + (e0 + s) * SES));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s00 + u * kVecTileS + v * VS;
            float f[VS];
            if (s + VS <= hw) {
                vec_to_floats<SBF>(raw[u], f);
            } else {
#pragma unroll
                for (int j = 0; j < VS; ++j) f[j] = (s + j < hw) ? load_elem<SBF>(sp, e0 + s + j) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < VS; ++j) tile[u * sub + c * PITCH + v * VS + j] = f[j];
        }
    }
    __syncthreads();
    // store: position rows, VD channels per 16-byte vector
    const int64_t drow = (int64_t)pb.H * pb.Dh;
    for (int idx = threadIdx.x; idx < kVecTileS * (pb.Dh / VD); idx += 256) {
        const int p = idx / (pb.Dh / VD), cv = idx % (pb.Dh / VD);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s00 + u * kVecTileS + p;
            if (s >= hw) continue;
            float f[VD];
#pragma unroll
            for (int j = 0; j < VD; ++j) f[j] = tile[u * sub + (cv * VD + j) * PITCH + p];
            const int64_t d = ((int64_t)n * pb.S + pb.geom.start[l] + s) * drow + (int64_t)h * pb.Dh + cv * VD;
            *reinterpret_cast<uint4*>(dst + d * DES) = floats_to_vec<DBF>(f);
        }
    }
}

template <bool DBF, int U>
__global__ void __launch_bounds__(256)
unpack_grad_vec_kernel(const Problem pb, const float* __restrict__ gv, const LevelViews dst, const TileMap64 tm) {
    extern __shared__ float tile[];                   // [U][Dh][65]
    constexpr int VD = ElemVec<DBF>::n, PITCH = kVecTileS + 1, DES = DBF ? 2 : 4;
    const int nh = blockIdx.y, n = nh / pb.H, h = nh % pb.H;
    int l = 0;
    while (l + 1 < pb.L && (int)blockIdx.x >= tm.first_tile[l + 1]) ++l;
    const int s00 = ((int)blockIdx.x - tm.first_tile[l]) * (kVecTileS * U);
    const int hw = pb.geom.h[l] * pb.geom.w[l];
    const int64_t srow = (int64_t)pb.H * pb.Dh;
    const int sub = pb.Dh * PITCH;
    for (int idx = threadIdx.x; idx < kVecTileS * (pb.Dh / 4); idx += 256) {
        const int p = idx / (pb.Dh / 4), cv = idx % (pb.Dh / 4);
        float4 f[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s00 + u * kVecTileS + p;
            f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < hw)
                f[u] = __ldg(reinterpret_cast<const float4*>(
                    gv + ((int64_t)n * pb.S + pb.geom.start[l] + s) * srow + (int64_t)h * pb.Dh + cv * 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float* t = tile + u * sub + p;
            t[(cv * 4 + 0) * PITCH] = f[u].x; t[(cv * 4 + 1) * PITCH] = f[u].y;
            t[(cv * 4 + 2) * PITCH] = f[u].z; t[(cv * 4 + 3) * PITCH] = f[u].w;
        }
    }
    __syncthreads();
    char* dp = reinterpret_cast<char*>(const_cast<void*>(dst.ptr[l]));
    for (int idx = threadIdx.x; idx < pb.Dh * (kVecTileS / VD); idx += 256) {
        const int c = idx / (kVecTileS / VD), v = idx % (kVecTileS / VD);
        const int64_t e0 = (int64_t)nh * dst.s_nh[l] + (int64_t)c * dst.s_c[l];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s00 + u * kVecTileS + v * VD;
            if (s >= hw) continue;
            float f[VD];
#pragma unroll
            for (int j = 0; j < VD; ++j) f[j] = tile[u * sub + c * PITCH + v * VD + j];
            if (s + VD <= hw) {
                *reinterpret_cast<uint4*>(dp + (e0 + s) * DES) = floats_to_vec<DBF>(f);
            } else {
#pragma unroll
                for (int j = 0; j < VD; ++j) if (s + j < hw) store_elem<DBF>(dp, e0 + s + j, f[j]);
            }
        }
    }
}

// the vector kernels need unit spatial stride and 16-byte aligned channel rows on the strided side
static bool views_vectorisable(const Problem& pb, const LevelViews& v, bool bf16) {
    const int es = bf16 ? 2 : 4;
    if (pb.Dh % 8 != 0) return false;
    for (int l = 0; l < pb.L; ++l) {
        if (v.s_s[l] != 1) return false;
        if ((reinterpret_cast<uintptr_t>(v.ptr[l]) & 15u) || (v.s_nh[l] * es) % 16 || (v.s_c[l] * es) % 16) return false;
    }
    return true;
}

// sub-tiles per block: as many as keep the staging tile at or below 33 KB (Dh 32: 4, Dh 64: 2, Dh 128: 1)
static int vec_sub_tiles(const Problem& pb) { return pb.Dh <= 32 ? 4 : pb.Dh <= 64 ? 2 : 1; }

static TileMap64 make_tile_map64(const Problem& pb, int u) {
    TileMap64 tm;
    int acc = 0;
    for (int l = 0; l < pb.L; ++l) {
        tm.first_tile[l] = acc;
        acc += (pb.geom.h[l] * pb.geom.w[l] + kVecTileS * u - 1) / (kVecTileS * u);
    }
    for (int l = pb.L; l <= MSDA_MAX_LEVELS; ++l) tm.first_tile[l] = acc;
    return tm;
}

static TileMap make_tile_map(const Problem& pb) {
    TileMap tm;
    int acc = 0;
    for (int l = 0; l < pb.L; ++l) {
        tm.first_tile[l] = acc;
        acc += (pb.geom.h[l] * pb.geom.w[l] + kTileS - 1) / kTileS;
    }
    for (int l = pb.L; l <= MSDA_MAX_LEVELS; ++l) tm.first_tile[l] = acc;
    return tm;
}

cudaError_t repack(const Problem& pb, const LevelViews& src, bool src_bf16, void* dst, bool dst_bf16,
                   cudaStream_t st) {
    if (views_vectorisable(pb, src, src_bf16) && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int u = vec_sub_tiles(pb);
        const TileMap64 tm64 = make_tile_map64(pb, u);
        const dim3 grid(tm64.first_tile[pb.L], pb.N * pb.H);
        const size_t smem = (size_t)u * pb.Dh * (kVecTileS + 1) * sizeof(float);
#define MSDA_REPACK(U)                                                                                      \
        if (src_bf16) {                                                                                     \
            if (dst_bf16) repack_vec_kernel<true, true, U><<<grid, 256, smem, st>>>(pb, src, tm64, (char*)dst);   \
            else repack_vec_kernel<true, false, U><<<grid, 256, smem, st>>>(pb, src, tm64, (char*)dst);     \
        } else {                                                                                            \
            if (dst_bf16) repack_vec_kernel<false, true, U><<<grid, 256, smem, st>>>(pb, src, tm64, (char*)dst);  \
            else repack_vec_kernel<false, false, U><<<grid, 256, smem, st>>>(pb, src, tm64, (char*)dst);    \
        }
        if (u == 4) { MSDA_REPACK(4) } else if (u == 2) { MSDA_REPACK(2) } else { MSDA_REPACK(1) }
#undef MSDA_REPACK
        return cudaGetLastError();
    }
    const TileMap tm = make_tile_map(pb);
    const dim3 grid(tm.first_tile[pb.L], pb.N * pb.H);
    const size_t smem = (size_t)pb.Dh * (kTileS + 1) * sizeof(float);
    if (src_bf16) {
        if (dst_bf16) repack_kernel<true, true><<<grid, 256, smem, st>>>(pb, src, tm, dst);
        else repack_kernel<true, false><<<grid, 256, smem, st>>>(pb, src, tm, dst);
    } else {
        if (dst_bf16) repack_kernel<false, true><<<grid, 256, smem, st>>>(pb, src, tm, dst);
        else repack_kernel<false, false><<<grid, 256, smem, st>>>(pb, src, tm, dst);
    }
    return cudaGetLastError();
}

cudaError_t unpack_grad(const Problem& pb, const float* grad_value, const LevelViews& dst, bool dst_bf16,
                        cudaStream_t st) {
    if (views_vectorisable(pb, dst, dst_bf16) && (reinterpret_cast<uintptr_t>(grad_value) & 15u) == 0) {
        const int u = vec_sub_tiles(pb);
        const TileMap64 tm64 = make_tile_map64(pb, u);
        const dim3 grid(tm64.first_tile[pb.L], pb.N * pb.H);
        const size_t smem = (size_t)u * pb.Dh * (kVecTileS + 1) * sizeof(float);
#define MSDA_UNPACK(U)                                                                                      \
        if (dst_bf16) unpack_grad_vec_kernel<true, U><<<grid, 256, smem, st>>>(pb, grad_value, dst, tm64);  \
        else unpack_grad_vec_kernel<false, U><<<grid, 256, smem, st>>>(pb, grad_value, dst, tm64);
        if (u == 4) { MSDA_UNPACK(4) } else if (u == 2) { MSDA_UNPACK(2) } else { MSDA_UNPACK(1) }
#undef MSDA_UNPACK
        return cudaGetLastError();
    }
    const TileMap tm = make_tile_map(pb);
    const dim3 grid(tm.first_tile[pb.L], pb.N * pb.H);
    const size_t smem = (size_t)pb.Dh * (kTileS + 1) * sizeof(float);
    if (dst_bf16) unpack_grad_kernel<true><<<grid, 256, smem, st>>>(pb, grad_value, dst, tm);
    else unpack_grad_kernel<false><<<grid, 256, smem, st>>>(pb, grad_value, dst, tm);
    return cudaGetLastError();
}

}  // namespace msda
