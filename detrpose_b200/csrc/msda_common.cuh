// Shared device helpers for the sm_100a multi-scale deformable attention kernels.
//
// Arithmetic contract (what the parity tests pin):
//   pixel coordinate  : ms_deform_attn.py:161 (g = 2*loc - 1) followed by ATen
//                       grid_sampler_unnormalize, GridSampler.h:33-35
//                       (((g + 1) * size - 1) / 2), one IEEE rounding per op
//                       (MSDA_COORD_UNFUSED) or with the multiply-subtract fused
//                       (MSDA_COORD_FMA);
//   corner weights    : nw=(x1-x)(y1-y) ne=(x-x0)(y1-y) sw=(x1-x)(y-y0) se=(x-x0)(y-y0)
//   zero padding      : each corner dropped individually when outside [0,W)x[0,H)
//                       (GridSampler.h:205), forward and backward alike (:238-243).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace msda {

struct LevelGeom {
    int32_t h[MSDA_MAX_LEVELS];
    int32_t w[MSDA_MAX_LEVELS];
    int32_t start[MSDA_MAX_LEVELS];   // exclusive running sum of h*w
};

struct Problem {
    int32_t N, Lq, H, Dh, L, P;
    int32_t S;                        // sum_l h*w
    int64_t vs_n, vs_s, vs_h;         // value strides in elements (channel stride 1)
    int32_t coord_mode;
    uint32_t magic_lp, magic_p, magic_h;   // ceil(2^32 / d) for d = L*P, P, H: see fastdiv
    LevelGeom geom;
};

// n / d for n * d < 2^32 with magic = ceil(2^32 / d) (host: div_magic): one IMAD.HI instead of the ~20
// instructions of a 32-bit division by a run-time value
__host__ __device__ inline uint32_t div_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ULL + d - 1) / d); }
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

__device__ __forceinline__ float pixel_coord(float loc, float size, int coord_mode) {
    // every op explicitly rounded so that nvcc cannot contract across them
    const float g = __fsub_rn(__fmul_rn(2.0f, loc), 1.0f);
    const float t = __fadd_rn(g, 1.0f);
    const float u = (coord_mode == MSDA_COORD_FMA) ? __fmaf_rn(t, size, -1.0f)
                                                  : __fsub_rn(__fmul_rn(t, size), 1.0f);
    return __fmul_rn(u, 0.5f);
}

// One bilinear sample: integer top-left corner, the four corner weights
// (already zeroed for corners outside the map) and clamped corner coordinates
// that are always safe to dereference.
struct Sample {
    int x0, y0;                // floor of the pixel coordinate (may be -1 .. size)
    float wx0, wx1;            // x1 - x, x - x0
    float wy0, wy1;            // y1 - y, y - y0
    float w_nw, w_ne, w_sw, w_se;
    bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ Sample make_sample(float lx, float ly, int H, int W, int coord_mode) {
    Sample s;
    float x = pixel_coord(lx, (float)W, coord_mode);
    float y = pixel_coord(ly, (float)H, coord_mode);
    // keep the int conversion defined for absurd / non-finite locations: anything
    // beyond one pixel outside the map has no valid corner anyway
    x = fminf(fmaxf(x, -2.0f), (float)W + 1.0f);
    y = fminf(fmaxf(y, -2.0f), (float)H + 1.0f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
    const float wx0 = __fsub_rn(x1f, x), wx1 = __fsub_rn(x, x0f);
    const float wy0 = __fsub_rn(y1f, y), wy1 = __fsub_rn(y, y0f);
    s.x0 = (int)x0f;
    s.y0 = (int)y0f;
    s.wx0 = wx0; s.wx1 = wx1;
    s.wy0 = wy0; s.wy1 = wy1;
    s.vx0 = (s.x0 >= 0) & (s.x0 < W);
    s.vx1 = (s.x0 + 1 >= 0) & (s.x0 + 1 < W);
    s.vy0 = (s.y0 >= 0) & (s.y0 < H);
    s.vy1 = (s.y0 + 1 >= 0) & (s.y0 + 1 < H);
    s.w_nw = (s.vx0 & s.vy0) ? __fmul_rn(wx0, wy0) : 0.0f;
    s.w_ne = (s.vx1 & s.vy0) ? __fmul_rn(wx1, wy0) : 0.0f;
    s.w_sw = (s.vx0 & s.vy1) ? __fmul_rn(wx0, wy1) : 0.0f;
    s.w_se = (s.vx1 & s.vy1) ? __fmul_rn(wx1, wy1) : 0.0f;
    return s;
}

// ---- 16-byte vector access -------------------------------------------------

template <bool BF16> struct Vec;          // one 16-byte vector of value elements
template <> struct Vec<false> { static constexpr int kElems = 4; };
template <> struct Vec<true>  { static constexpr int kElems = 8; };

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// unpack one 16-byte vector into fp32 lanes
template <bool BF16>
__device__ __forceinline__ void unpack(const uint4& v, float* f) {
    if constexpr (BF16) {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
        f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
        f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    } else {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);   // .x = lo (low half)
    return *reinterpret_cast<const uint32_t*>(&p);
}

// vector reduction into global memory, no return value (sm_90+)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace msda
