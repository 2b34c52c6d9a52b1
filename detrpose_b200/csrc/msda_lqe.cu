// LQE sampler: the keypoint-quality statistics of the pose decoder.
//
// Reference: LQE.forward, /root/reference/src/models/detrpose/transformer.py:274-283 --
//     v    = grid_sample(feat, 2*poses - 1, bilinear, zeros, align_corners=False).permute(0, 2, 3, 1)   (:278-280)
//     top  = v.topk(k, dim=-1)[0]                                                                       (:282)
//     stat = cat([top, top.mean(-1, keepdim=True)], -1)                                                 (:284)
// i.e. per keypoint: sample all C channels of the finest feature map, keep the k largest (descending) and
// their mean.  The reference materialises the (B, C, L, 17) sampled tensor, permutes it and runs a
// sort-based top-k over it; here one warp owns one keypoint: lanes split the channels, the four corner
// loads per channel reuse the sampler's own coordinate / zero-padding code (make_sample), the top-k is k
// rounds of a shuffle arg-max over values that never leave registers, and only k+1 floats (+ k channel
// indices for the backward) are written.  `feat` is addressed through its strides, so both NCHW (the
// encoder's layout: the two x-neighbours share a sector) and channels-last (lanes coalesce) are read in
// place.
//
// Backward: the gradient of stat reaches only the k selected channels; 4k lanes own one (channel, corner)
// each: a scalar fp32 atomic into grad_feat and a shuffle reduction for the pose gradient.
#include <cfloat>

#include "msda_common.cuh"
#include "msda_kernels.cuh"

namespace msda {
namespace {

constexpr int kLqeWarps = 8;
constexpr unsigned kFull = 0xffffffffu;

template <bool BF>
__device__ __forceinline__ float load_elem(const void* base, int64_t off) {
    if constexpr (BF) {
        const unsigned short r = __ldg(static_cast<const unsigned short*>(base) + off);
        return __uint_as_float((unsigned)r << 16);
    } else {
        return __ldg(static_cast<const float*>(base) + off);
    }
}

// torch.topk's order (transformer.py:282): NaN counts as the largest value, ties go to the lower channel;
// a candidate with channel kNone ("this lane has nothing left") loses against everything
constexpr int kNone = 0x7fffffff;
__device__ __forceinline__ bool lqe_better(float a, int ca, float b, int cb) {
    if (cb == kNone) return ca != kNone;
    if (ca == kNone) return false;
    const bool an = a != a, bn = b != b;
    if (an != bn) return an;
    if (!an && a != b) return a > b;
    return ca < cb;
}

struct LqeGeom {
    int64_t sb, sc, sy, sx;     // feat strides in elements
    int B, C, Hf, Wf, P, K, coord_mode;
};

template <bool BF, int NJ>
__global__ void __launch_bounds__(kLqeWarps * 32)
lqe_fwd_kernel(const void* __restrict__ feat, const float* __restrict__ poses, float* __restrict__ stat,
               int32_t* __restrict__ topk_idx, const LqeGeom gm) {
    const int lane = threadIdx.x & 31;
    const int64_t pt = (int64_t)blockIdx.x * kLqeWarps + (threadIdx.x >> 5);
    if (pt >= (int64_t)gm.B * gm.P) return;
    const int b = (int)(pt / gm.P);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(poses) + pt);
    const Sample s = make_sample(xy.x, xy.y, gm.Hf, gm.Wf, gm.coord_mode);

    const int64_t o_nw = (int64_t)b * gm.sb + (int64_t)s.y0 * gm.sy + (int64_t)s.x0 * gm.sx;
    const bool v_nw = s.vx0 & s.vy0, v_ne = s.vx1 & s.vy0, v_sw = s.vx0 & s.vy1, v_se = s.vx1 & s.vy1;

    float vals[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int64_t o = o_nw + (int64_t)(j * 32 + lane) * gm.sc;
        const float a = v_nw ? load_elem<BF>(feat, o) : 0.0f;
        const float c = v_ne ? load_elem<BF>(feat, o + gm.sx) : 0.0f;
        const float d = v_sw ? load_elem<BF>(feat, o + gm.sy) : 0.0f;
        const float e = v_se ? load_elem<BF>(feat, o + gm.sy + gm.sx) : 0.0f;
        vals[j] = fmaf(e, s.w_se, fmaf(d, s.w_sw, fmaf(c, s.w_ne, a * s.w_nw)));
    }

    float sum = 0.0f, mine = 0.0f;
    int mine_c = 0;
    unsigned taken = 0u;                      // channels of this lane already selected (bit j)
    for (int r = 0; r < gm.K; ++r) {
        // best of this lane among the channels not taken yet, then arg-max over the warp.  Non-finite
        // samples are ordinary candidates (NaN first, as torch.topk orders them): the statistics then carry
        // the NaN / Inf like the reference's do, and every stored index is a real channel.
        float bv = 0.0f;
        int bc = kNone;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (!((taken >> j) & 1u) && lqe_better(vals[j], j * 32 + lane, bv, bc)) { bv = vals[j]; bc = j * 32 + lane; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(kFull, bv, off);
            const int oc = __shfl_xor_sync(kFull, bc, off);
            if (lqe_better(ov, oc, bv, bc)) { bv = ov; bc = oc; }
        }
        if ((bc & 31) == lane && bc != kNone) taken |= 1u << (bc >> 5);
        sum += bv;
        if (lane == r) { mine = bv; mine_c = bc; }
    }
    float* out = stat + pt * (gm.K + 1);
    if (lane < gm.K) {
        out[lane] = mine;
        if (topk_idx != nullptr) topk_idx[pt * gm.K + lane] = mine_c;
    } else if (lane == gm.K) {
        out[gm.K] = sum / (float)gm.K;
    }
}

template <bool BF>
__global__ void __launch_bounds__(kLqeWarps * 32)
lqe_bwd_kernel(const void* __restrict__ feat, const float* __restrict__ poses, const int32_t* __restrict__ topk_idx,
               const float* __restrict__ grad_stat, float* __restrict__ grad_feat, float* __restrict__ grad_poses,
               const LqeGeom gm) {
    const int lane = threadIdx.x & 31;
    const int64_t pt = (int64_t)blockIdx.x * kLqeWarps + (threadIdx.x >> 5);
    if (pt >= (int64_t)gm.B * gm.P) return;
    const int b = (int)(pt / gm.P);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(poses) + pt);
    const Sample s = make_sample(xy.x, xy.y, gm.Hf, gm.Wf, gm.coord_mode);

    // lane = 4 * j + corner (corner: 0 nw, 1 ne, 2 sw, 3 se); K <= 8
    const int j = lane >> 2, corner = lane & 3;
    const bool right = corner & 1, low = corner & 2;
    float gx = 0.0f, gy = 0.0f;
    if (j < gm.K) {
        const float g = __ldg(grad_stat + pt * (gm.K + 1) + j) + __ldg(grad_stat + pt * (gm.K + 1) + gm.K) / (float)gm.K;
        const int c = __ldg(topk_idx + pt * gm.K + j);
        // an index outside [0, C) cannot come from the forward kernel; never turn one into an address
        const bool valid = (right ? s.vx1 : s.vx0) & (low ? s.vy1 : s.vy0) & (c >= 0) & (c < gm.C);
        if (valid) {
            const int64_t o = (int64_t)b * gm.sb + (int64_t)c * gm.sc + (int64_t)(s.y0 + (low ? 1 : 0)) * gm.sy +
                              (int64_t)(s.x0 + (right ? 1 : 0)) * gm.sx;
            const float wx = right ? s.wx1 : s.wx0, wy = low ? s.wy1 : s.wy0;
            if (grad_feat != nullptr) atomicAdd(grad_feat + o, g * wx * wy);
            const float v = load_elem<BF>(feat, o) * g;
            gx = (right ? v : -v) * wy;            // d(weight)/dx = +-wy, d(weight)/dy = +-wx
            gy = (low ? v : -v) * wx;
        }
    }
    if (grad_poses != nullptr) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            gx += __shfl_xor_sync(kFull, gx, off);
            gy += __shfl_xor_sync(kFull, gy, off);
        }
        // pixel coordinate = pose * size - 0.5: the chain rule carries the map size (GridSampler.h unnormalise)
        if (lane == 0)
            reinterpret_cast<float2*>(grad_poses)[pt] = make_float2(gx * (float)gm.Wf, gy * (float)gm.Hf);
    }
}

}  // namespace

bool lqe_supported(int C, int K) { return (C == 128 || C == 256 || C == 384 || C == 512) && K >= 1 && K <= 8; }

static LqeGeom make_geom(const int64_t* strides, int B, int C, int Hf, int Wf, int P, int K, int coord_mode) {
    LqeGeom g;
    g.sb = strides[0]; g.sc = strides[1]; g.sy = strides[2]; g.sx = strides[3];
    g.B = B; g.C = C; g.Hf = Hf; g.Wf = Wf; g.P = P; g.K = K; g.coord_mode = coord_mode;
    return g;
}

cudaError_t lqe_forward(const void* feat, bool feat_bf16, const int64_t* strides, const float* poses, float* stat,
                        int32_t* topk_idx, int B, int C, int Hf, int Wf, int P, int K, int coord_mode,
                        cudaStream_t st) {
    const LqeGeom g = make_geom(strides, B, C, Hf, Wf, P, K, coord_mode);
    const unsigned blocks = (unsigned)(((int64_t)B * P + kLqeWarps - 1) / kLqeWarps);
#define MSDA_LQE_FWD(NJ)                                                                                   \
    if (feat_bf16) lqe_fwd_kernel<true, NJ><<<blocks, kLqeWarps * 32, 0, st>>>(feat, poses, stat, topk_idx, g); \
    else lqe_fwd_kernel<false, NJ><<<blocks, kLqeWarps * 32, 0, st>>>(feat, poses, stat, topk_idx, g);          \
    break;
    switch (C / 32) {
        case 4: MSDA_LQE_FWD(4)
        case 8: MSDA_LQE_FWD(8)
        case 12: MSDA_LQE_FWD(12)
        case 16: MSDA_LQE_FWD(16)
        default: return cudaErrorInvalidValue;
    }
#undef MSDA_LQE_FWD
    return cudaGetLastError();
}

cudaError_t lqe_backward(const void* feat, bool feat_bf16, const int64_t* strides, const float* poses,
                         const int32_t* topk_idx, const float* grad_stat, float* grad_feat, float* grad_poses,
                         int B, int C, int Hf, int Wf, int P, int K, int coord_mode, cudaStream_t st) {
    const LqeGeom g = make_geom(strides, B, C, Hf, Wf, P, K, coord_mode);
    const unsigned blocks = (unsigned)(((int64_t)B * P + kLqeWarps - 1) / kLqeWarps);
    if (feat_bf16) lqe_bwd_kernel<true><<<blocks, kLqeWarps * 32, 0, st>>>(feat, poses, topk_idx, grad_stat, grad_feat, grad_poses, g);
    else lqe_bwd_kernel<false><<<blocks, kLqeWarps * 32, 0, st>>>(feat, poses, topk_idx, grad_stat, grad_feat, grad_poses, g);
    return cudaGetLastError();
}

}  // namespace msda
