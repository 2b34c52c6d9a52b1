// C ABI (include/msda_b200.h): argument checking, problem description, kernel
// selection.  No torch types, no global mutable state besides the per-thread
// error string and the (atomic) kernel-variant knobs.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "msda_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

std::atomic<int> g_fwd_variant{-1};   // -1: automatic
std::atomic<int> g_bwd_variant{-1};

// Validates sizes and fills the kernel-side problem description.
int make_problem(msda::Problem& pb, int N, int Lq, int H, int Dh, int L, int P,
                 const int32_t* spatial_shapes, int coord_mode) {
    if (spatial_shapes == nullptr) return fail(MSDA_ERR_NULL, "spatial_shapes is NULL");
    if (N <= 0 || Lq <= 0 || H <= 0 || L <= 0 || P <= 0)
        return fail(MSDA_ERR_SHAPE, "non-positive size N=%d Lq=%d H=%d L=%d P=%d", N, Lq, H, L, P);
    if (L > MSDA_MAX_LEVELS || P > MSDA_MAX_POINTS)
        return fail(MSDA_ERR_LEVELS, "L=%d (max %d) or P=%d (max %d) too large", L, MSDA_MAX_LEVELS, P,
                    MSDA_MAX_POINTS);
    if (Dh < 8 || Dh > 128 || Dh % 8 != 0)
        return fail(MSDA_ERR_DHEAD, "Dh=%d unsupported (multiple of 8 in 8..128)", Dh);
    if (coord_mode != MSDA_COORD_UNFUSED && coord_mode != MSDA_COORD_FMA)
        return fail(MSDA_ERR_SHAPE, "unknown coord_mode %d", coord_mode);
    pb.N = N; pb.Lq = Lq; pb.H = H; pb.Dh = Dh; pb.L = L; pb.P = P;
    pb.coord_mode = coord_mode;
    pb.magic_lp = msda::div_magic((uint32_t)(L * P));
    pb.magic_p = msda::div_magic((uint32_t)P);
    pb.magic_h = msda::div_magic((uint32_t)H);
    int64_t acc = 0;
    for (int l = 0; l < MSDA_MAX_LEVELS; ++l) {
        if (l < L) {
            const int32_t h = spatial_shapes[2 * l], w = spatial_shapes[2 * l + 1];
            if (h <= 0 || w <= 0 || h > 16384 || w > 16384)
                return fail(MSDA_ERR_SHAPE, "level %d has shape %dx%d", l, h, w);
            pb.geom.h[l] = h; pb.geom.w[l] = w; pb.geom.start[l] = (int32_t)acc;
            acc += (int64_t)h * w;
            if (acc > (1 << 30)) return fail(MSDA_ERR_SHAPE, "pyramid too large");
        } else {
            pb.geom.h[l] = 1; pb.geom.w[l] = 1; pb.geom.start[l] = (int32_t)acc;
        }
    }
    pb.S = (int32_t)acc;
    pb.vs_n = (int64_t)pb.S * H * Dh; pb.vs_s = (int64_t)H * Dh; pb.vs_h = Dh;
    return MSDA_OK;
}

int set_value_strides(msda::Problem& pb, const void* value, int value_dtype, const int64_t* strides) {
    if (value == nullptr || strides == nullptr) return fail(MSDA_ERR_NULL, "value / value_strides is NULL");
    if (value_dtype != MSDA_F32 && value_dtype != MSDA_BF16)
        return fail(MSDA_ERR_DTYPE, "unknown value dtype %d", value_dtype);
    const int es = value_dtype == MSDA_BF16 ? 2 : 4;
    pb.vs_n = strides[0]; pb.vs_s = strides[1]; pb.vs_h = strides[2];
    if (!aligned16(value) || (pb.vs_n * es) % 16 || (pb.vs_s * es) % 16 || (pb.vs_h * es) % 16)
        return fail(MSDA_ERR_ALIGN, "value rows must be 16-byte aligned (ptr %p, strides %lld %lld %lld, %d B elems)",
                    value, (long long)pb.vs_n, (long long)pb.vs_s, (long long)pb.vs_h, es);
    if (pb.vs_s < 0 || pb.vs_h < 0 || pb.vs_n < 0) return fail(MSDA_ERR_SHAPE, "negative value stride");
    return MSDA_OK;
}

int fill_views(msda::LevelViews& v, const void* const* ptrs, const int64_t* strides, int L) {
    if (ptrs == nullptr || strides == nullptr) return fail(MSDA_ERR_NULL, "level_ptrs / level_strides is NULL");
    for (int l = 0; l < MSDA_MAX_LEVELS; ++l) {
        if (l < L) {
            if (ptrs[l] == nullptr) return fail(MSDA_ERR_NULL, "level_ptrs[%d] is NULL", l);
            v.ptr[l] = ptrs[l];
            v.s_nh[l] = strides[3 * l]; v.s_c[l] = strides[3 * l + 1]; v.s_s[l] = strides[3 * l + 2];
        } else {
            v.ptr[l] = nullptr; v.s_nh[l] = v.s_c[l] = v.s_s[l] = 0;
        }
    }
    return MSDA_OK;
}

}  // namespace

extern "C" {

MSDA_API int msda_b200_abi_version(void) { return MSDA_B200_ABI_VERSION; }

MSDA_API const char* msda_b200_last_error(void) { return g_err; }

MSDA_API int msda_b200_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(MSDA_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e)); }
    int v = 0;
    if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
    if (cc_major) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); *cc_major = v; }
    if (cc_minor) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); *cc_minor = v; }
    return MSDA_OK;
}

MSDA_API int msda_b200_debug_phase_buffer(void* device_buffer) {
    const cudaError_t e = msda::set_phase_buffer(static_cast<unsigned long long*>(device_buffer));
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_debug_phase_buffer");
}

MSDA_API int msda_b200_set_variant(int fwd_variant, int bwd_variant) {
    g_fwd_variant.store(fwd_variant);
    g_bwd_variant.store(bwd_variant);
    return MSDA_OK;
}

MSDA_API int msda_b200_forward(const void* value, int value_dtype, const int64_t* value_strides,
                      const int32_t* spatial_shapes, const float* locations, const float* attention,
                      void* out, int out_dtype, int N, int Lq, int H, int Dh, int L, int P,
                      int coord_mode, void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, Lq, H, Dh, L, P, spatial_shapes, coord_mode);
    if (rc) return rc;
    if ((rc = set_value_strides(pb, value, value_dtype, value_strides))) return rc;
    if (!locations || !attention || !out) return fail(MSDA_ERR_NULL, "locations / attention / out is NULL");
    if (out_dtype != MSDA_F32 && out_dtype != MSDA_BF16) return fail(MSDA_ERR_DTYPE, "unknown out dtype %d", out_dtype);
    if (!aligned16(out) || !aligned16(locations) || !aligned16(attention))
        return fail(MSDA_ERR_ALIGN, "locations / attention / out must be 16-byte aligned");
    // variant 1 (default when the shape fits): lean kernel; variant 0: flat kernel
    // 100 + v: variant v without the streaming L2 prefetch of the pyramid (150: the automatic choice without it)
    const int raw_variant = g_fwd_variant.load();
    const int l2_prefetch = raw_variant >= 100 ? 0 : 1;
    const int variant = raw_variant >= 100 ? (raw_variant == 150 ? -1 : raw_variant - 100) : raw_variant;
    const bool vbf = value_dtype == MSDA_BF16;
    const bool can_lean = msda::forward_lean_supported(pb, vbf);
    if (variant == 1 && !can_lean) return fail(MSDA_ERR_SHAPE, "lean forward does not support this shape");
    cudaError_t e;
    if (variant == 2 || variant == 3) {   // TMA-staged coarse levels (opt-in): 2 = one 768-thread CTA per SM
        const bool small = variant == 3;  // staging all it can, 3 = 256-thread CTAs staging the coarsest level(s)
        if (!msda::forward_staged_supported(pb, vbf, small))
            return fail(MSDA_ERR_SHAPE, "staged forward does not support this shape");
        e = msda::forward_staged(pb, value, vbf, locations, attention, out, out_dtype == MSDA_BF16, small,
                                 (cudaStream_t)stream);
    } else {
        // variant 7: the one-lane-group form also for rows of at most 32 bytes (default there: two lane groups)
        e = (can_lean && variant != 0)
            ? msda::forward_lean(pb, value, vbf, locations, attention, out, out_dtype == MSDA_BF16,
                                 (cudaStream_t)stream, nullptr, 1, nullptr, variant == 7 ? 0 : -1, l2_prefetch)
            : msda::forward_flat(pb, value, vbf, locations, attention, out, out_dtype == MSDA_BF16, (cudaStream_t)stream);
    }
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_forward launch");
}

MSDA_API int msda_b200_backward(const void* value, int value_dtype, const int64_t* value_strides,
                       const int32_t* spatial_shapes, const float* locations, const float* attention,
                       const void* grad_out, int grad_out_dtype, float* grad_value, int accumulate,
                       float* grad_locations, float* grad_attention, int N, int Lq, int H, int Dh, int L, int P,
                       int coord_mode, void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, Lq, H, Dh, L, P, spatial_shapes, coord_mode);
    if (rc) return rc;
    if ((rc = set_value_strides(pb, value, value_dtype, value_strides))) return rc;
    if (!locations || !attention || !grad_out) return fail(MSDA_ERR_NULL, "locations / attention / grad_out is NULL");
    if ((grad_locations == nullptr) != (grad_attention == nullptr))
        return fail(MSDA_ERR_NULL, "grad_locations and grad_attention must be given together");
    if (grad_out_dtype != MSDA_F32 && grad_out_dtype != MSDA_BF16)
        return fail(MSDA_ERR_DTYPE, "unknown grad_out dtype %d", grad_out_dtype);
    if (!aligned16(grad_out) || !aligned16(locations) || !aligned16(attention) ||
        (grad_value && !aligned16(grad_value)) || (grad_attention && !aligned16(grad_attention)))
        return fail(MSDA_ERR_ALIGN, "backward buffers must be 16-byte aligned");
    // the gather kernel writes grad_locations with one 256-bit store per (query, 4 points)
    if (grad_locations && (reinterpret_cast<uintptr_t>(grad_locations) & 31u))
        return fail(MSDA_ERR_ALIGN, "grad_locations must be 32-byte aligned");
    if (!grad_value && !grad_locations) return MSDA_OK;
    const cudaStream_t st = (cudaStream_t)stream;
    const bool vbf = value_dtype == MSDA_BF16;
    // variant 1 (default when the shape fits): gather form, no atomics (2: same with 512 threads x 128
    // registers); variant 0: flat + vector reductions
    const int variant = g_bwd_variant.load();
    const bool can_gather = grad_out_dtype == value_dtype && msda::backward_gather_supported(pb, vbf);
    if ((variant == 1 || variant == 2) && !can_gather)
        return fail(MSDA_ERR_SHAPE, "gather-form backward does not support this shape/dtype combination");
    cudaError_t e;
    if (can_gather && variant != 0) {
        e = msda::backward_gather(pb, value, vbf, locations, attention, grad_out, grad_value, grad_locations,
                                  grad_attention, accumulate, variant == 2 ? 512 : 1024, st);
    } else {
        if (grad_value && !accumulate) {
            e = cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)N * pb.S * H * Dh, st);
            if (e != cudaSuccess) return cuda_fail(e, "msda_b200_backward memset");
        }
        e = msda::backward_flat(pb, value, vbf, locations, attention, grad_out, grad_out_dtype == MSDA_BF16,
                                grad_value, grad_locations, grad_attention, st);
    }
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_backward launch");
}

// ---- fused prologue (row f1) ----
static bool fused_ok(const msda::Problem& pb, bool vbf) {
    return msda::forward_lean_supported(pb, vbf) && msda::backward_gather_supported(pb, vbf);
}

MSDA_API int msda_b200_fused_supported(int value_dtype, const int64_t* value_strides, const int32_t* spatial_shapes,
                              int N, int Lq, int H, int Dh, int L, int P) {
    msda::Problem pb;
    if (make_problem(pb, N, Lq, H, Dh, L, P, spatial_shapes, MSDA_COORD_FMA)) return 0;
    if (value_dtype != MSDA_F32 && value_dtype != MSDA_BF16) return 0;
    if (value_strides) { pb.vs_n = value_strides[0]; pb.vs_s = value_strides[1]; pb.vs_h = value_strides[2]; }
    return fused_ok(pb, value_dtype == MSDA_BF16) ? 1 : 0;
}

MSDA_API int msda_b200_forward_fused(const void* value, int value_dtype, const int64_t* value_strides,
                            const int32_t* spatial_shapes, const float* offsets, const float* logits,
                            const float* ref_points, int ref_levels, void* out, int out_dtype,
                            float* attention_out, int N, int Lq, int H, int Dh, int L, int P,
                            int coord_mode, void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, Lq, H, Dh, L, P, spatial_shapes, coord_mode);
    if (rc) return rc;
    if ((rc = set_value_strides(pb, value, value_dtype, value_strides))) return rc;
    if (!offsets || !logits || !ref_points || !out) return fail(MSDA_ERR_NULL, "offsets / logits / ref_points / out is NULL");
    if (ref_levels != 1 && ref_levels != L) return fail(MSDA_ERR_SHAPE, "ref_levels=%d must be 1 or L=%d", ref_levels, L);
    if (out_dtype != MSDA_F32 && out_dtype != MSDA_BF16) return fail(MSDA_ERR_DTYPE, "unknown out dtype %d", out_dtype);
    if (!aligned16(out) || !aligned16(offsets) || !aligned16(logits) || (attention_out && !aligned16(attention_out)) ||
        (reinterpret_cast<uintptr_t>(ref_points) & 7u))
        return fail(MSDA_ERR_ALIGN, "fused forward: offsets / logits / out / attention_out must be 16-byte, ref_points 8-byte aligned");
    const bool vbf = value_dtype == MSDA_BF16;
    if (!msda::forward_lean_supported(pb, vbf)) return fail(MSDA_ERR_SHAPE, "fused forward does not support this shape");
    const cudaError_t e = msda::forward_lean(pb, value, vbf, offsets, logits, out, out_dtype == MSDA_BF16,
                                             (cudaStream_t)stream, ref_points, ref_levels, attention_out);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_forward_fused launch");
}

MSDA_API int msda_b200_backward_fused(const void* value, int value_dtype, const int64_t* value_strides,
                             const int32_t* spatial_shapes, const float* offsets, const float* ref_points,
                             int ref_levels, const float* attention, const void* grad_out, int grad_out_dtype,
                             float* grad_value, int accumulate, float* grad_offsets, float* grad_attention,
                             int N, int Lq, int H, int Dh, int L, int P, int coord_mode, void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, Lq, H, Dh, L, P, spatial_shapes, coord_mode);
    if (rc) return rc;
    if ((rc = set_value_strides(pb, value, value_dtype, value_strides))) return rc;
    if (!offsets || !ref_points || !attention || !grad_out)
        return fail(MSDA_ERR_NULL, "offsets / ref_points / attention / grad_out is NULL");
    if (ref_levels != 1 && ref_levels != L) return fail(MSDA_ERR_SHAPE, "ref_levels=%d must be 1 or L=%d", ref_levels, L);
    if ((grad_offsets == nullptr) != (grad_attention == nullptr))
        return fail(MSDA_ERR_NULL, "grad_offsets and grad_attention must be given together");
    if (grad_out_dtype != value_dtype)
        return fail(MSDA_ERR_DTYPE, "fused backward: grad_out dtype %d must equal the value dtype %d", grad_out_dtype, value_dtype);
    if (!aligned16(grad_out) || !aligned16(offsets) || !aligned16(attention) || (grad_value && !aligned16(grad_value)) ||
        (grad_attention && !aligned16(grad_attention)) || (reinterpret_cast<uintptr_t>(ref_points) & 7u))
        return fail(MSDA_ERR_ALIGN, "fused backward buffers must be 16-byte aligned (ref_points 8-byte)");
    if (grad_offsets && (reinterpret_cast<uintptr_t>(grad_offsets) & 31u))
        return fail(MSDA_ERR_ALIGN, "grad_offsets must be 32-byte aligned");
    if (!grad_value && !grad_offsets) return MSDA_OK;
    const bool vbf = value_dtype == MSDA_BF16;
    if (!msda::backward_gather_supported(pb, vbf)) return fail(MSDA_ERR_SHAPE, "fused backward does not support this shape");
    const int variant = g_bwd_variant.load();
    const cudaError_t e = msda::backward_gather(pb, value, vbf, offsets, attention, grad_out, grad_value, grad_offsets,
                                                grad_attention, accumulate,
                                                variant == 2 ? 512 : 1024, (cudaStream_t)stream, ref_points, ref_levels);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_backward_fused launch");
}

MSDA_API int msda_b200_softmax_backward(const float* attention, const float* grad_attention, float* grad_logits,
                               int64_t rows, int cols, void* stream) {
    if (!attention || !grad_attention || !grad_logits) return fail(MSDA_ERR_NULL, "softmax backward: NULL argument");
    if (rows <= 0 || cols <= 0 || cols > MSDA_MAX_LEVELS * MSDA_MAX_POINTS)
        return fail(MSDA_ERR_SHAPE, "softmax backward: rows=%lld cols=%d", (long long)rows, cols);
    if (!aligned16(attention) || !aligned16(grad_attention) || !aligned16(grad_logits))
        return fail(MSDA_ERR_ALIGN, "softmax backward: buffers must be 16-byte aligned");
    const cudaError_t e = msda::softmax_backward(attention, grad_attention, grad_logits, rows, cols, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_softmax_backward launch");
}

MSDA_API int msda_b200_sample_indices(const int32_t* spatial_shapes, const float* locations, int32_t* idx_out,
                             int32_t* level_start_out, int N, int Lq, int H, int L, int P, int coord_mode,
                             void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, Lq, H, 8, L, P, spatial_shapes, coord_mode);
    if (rc) return rc;
    if (!locations || !idx_out) return fail(MSDA_ERR_NULL, "locations / idx_out is NULL");
    const cudaError_t e = msda::sample_indices(pb, locations, idx_out, level_start_out, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_sample_indices launch");
}

MSDA_API int msda_b200_locations(const float* offsets, const float* logits, const float* ref_points, int ref_levels,
                        const int32_t* spatial_shapes, float* locations, float* attention, int N, int Lq,
                        int H, int L, int P, void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, Lq, H, 8, L, P, spatial_shapes, MSDA_COORD_UNFUSED);
    if (rc) return rc;
    if (!offsets || !logits || !ref_points || !locations || !attention)
        return fail(MSDA_ERR_NULL, "a prologue buffer is NULL");
    if (ref_levels != 1 && ref_levels != L)
        return fail(MSDA_ERR_SHAPE, "ref_levels=%d must be 1 or L=%d", ref_levels, L);
    const cudaError_t e = msda::locations(pb, offsets, logits, ref_points, ref_levels, locations, attention,
                                          (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_locations launch");
}

MSDA_API int msda_b200_repack(const void* const* level_ptrs, const int64_t* level_strides, int src_dtype,
                     const int32_t* spatial_shapes, void* dst, int dst_dtype, int N, int H, int Dh, int L,
                     void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, 1, H, Dh, L, 1, spatial_shapes, MSDA_COORD_UNFUSED);
    if (rc) return rc;
    msda::LevelViews v;
    if ((rc = fill_views(v, level_ptrs, level_strides, L))) return rc;
    if (!dst) return fail(MSDA_ERR_NULL, "dst is NULL");
    if ((src_dtype != MSDA_F32 && src_dtype != MSDA_BF16) || (dst_dtype != MSDA_F32 && dst_dtype != MSDA_BF16))
        return fail(MSDA_ERR_DTYPE, "unknown dtype %d / %d", src_dtype, dst_dtype);
    if ((int64_t)N * H > 65535) return fail(MSDA_ERR_SHAPE, "N*H=%lld exceeds 65535", (long long)N * H);
    const cudaError_t e = msda::repack(pb, v, src_dtype == MSDA_BF16, dst, dst_dtype == MSDA_BF16, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_repack launch");
}

MSDA_API int msda_b200_unpack_grad(const float* grad_value, const int32_t* spatial_shapes, void* const* level_ptrs,
                          const int64_t* level_strides, int dst_dtype, int N, int H, int Dh, int L,
                          void* stream) {
    msda::Problem pb;
    int rc = make_problem(pb, N, 1, H, Dh, L, 1, spatial_shapes, MSDA_COORD_UNFUSED);
    if (rc) return rc;
    msda::LevelViews v;
    if ((rc = fill_views(v, (const void* const*)level_ptrs, level_strides, L))) return rc;
    if (!grad_value) return fail(MSDA_ERR_NULL, "grad_value is NULL");
    if (dst_dtype != MSDA_F32 && dst_dtype != MSDA_BF16) return fail(MSDA_ERR_DTYPE, "unknown dtype %d", dst_dtype);
    if ((int64_t)N * H > 65535) return fail(MSDA_ERR_SHAPE, "N*H=%lld exceeds 65535", (long long)N * H);
    const cudaError_t e = msda::unpack_grad(pb, grad_value, v, dst_dtype == MSDA_BF16, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "msda_b200_unpack_grad launch");
}

namespace {
int check_gate(const void* pre, int pre_dtype, const void* x1, const void* x2, int x_dtype, int64_t rows, int C) {
    if (pre == nullptr || x1 == nullptr || x2 == nullptr) return fail(MSDA_ERR_NULL, "gate: NULL input");
    if ((pre_dtype != MSDA_F32 && pre_dtype != MSDA_BF16) || (x_dtype != MSDA_F32 && x_dtype != MSDA_BF16))
        return fail(MSDA_ERR_DTYPE, "gate: unknown dtype %d / %d", pre_dtype, x_dtype);
    if (rows <= 0) return fail(MSDA_ERR_SHAPE, "gate: rows=%lld", (long long)rows);
    if (!msda::gate_supported(C)) return fail(MSDA_ERR_SHAPE, "gate: C=%d unsupported (128, 256, 384, 512)", C);
    if (!aligned16(pre) || !aligned16(x1) || !aligned16(x2)) return fail(MSDA_ERR_ALIGN, "gate: unaligned input");
    return MSDA_OK;
}
}  // namespace

MSDA_API int msda_b200_gate_forward(const void* pre, int pre_dtype, const void* x1, const void* x2, int x_dtype,
                                    const float* gamma, const float* beta, float eps, void* y, float* stats,
                                    int64_t rows, int C, void* stream) {
    if (int rc = check_gate(pre, pre_dtype, x1, x2, x_dtype, rows, C)) return rc;
    if (gamma == nullptr || beta == nullptr || y == nullptr) return fail(MSDA_ERR_NULL, "gate: NULL gamma/beta/y");
    if (!aligned16(gamma) || !aligned16(beta) || !aligned16(y) || (reinterpret_cast<uintptr_t>(stats) & 7u))
        return fail(MSDA_ERR_ALIGN, "gate: unaligned gamma/beta/y/stats");
    const cudaError_t e = msda::gate_forward(pre, pre_dtype == MSDA_BF16, x1, x2, x_dtype == MSDA_BF16, gamma, beta,
                                             eps, y, stats, rows, C, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "gate forward launch");
}

MSDA_API int msda_b200_gate_backward(const void* pre, int pre_dtype, const void* x1, const void* x2, int x_dtype,
                                     const float* gamma, const float* stats, const void* grad_y, void* grad_pre,
                                     void* grad_x1, void* grad_x2, float* grad_gamma, float* grad_beta,
                                     int64_t rows, int C, void* stream) {
    if (int rc = check_gate(pre, pre_dtype, x1, x2, x_dtype, rows, C)) return rc;
    if (gamma == nullptr || stats == nullptr || grad_y == nullptr || grad_pre == nullptr || grad_x1 == nullptr ||
        grad_x2 == nullptr || grad_gamma == nullptr || grad_beta == nullptr)
        return fail(MSDA_ERR_NULL, "gate backward: NULL argument");
    if (!aligned16(gamma) || !aligned16(grad_y) || !aligned16(grad_pre) || !aligned16(grad_x1) ||
        !aligned16(grad_x2) || (reinterpret_cast<uintptr_t>(stats) & 7u))
        return fail(MSDA_ERR_ALIGN, "gate backward: unaligned argument");
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "gate backward: device query");
    e = msda::gate_backward(pre, pre_dtype == MSDA_BF16, x1, x2, x_dtype == MSDA_BF16, gamma, stats, grad_y,
                            grad_pre, grad_x1, grad_x2, grad_gamma, grad_beta, rows, C, sms, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "gate backward launch");
}


namespace {
int check_lqe(const void* feat, int feat_dtype, const int64_t* strides, const float* poses, int B, int C, int Hf,
              int Wf, int P, int K, int coord_mode) {
    if (feat == nullptr || strides == nullptr || poses == nullptr) return fail(MSDA_ERR_NULL, "lqe: NULL input");
    if (feat_dtype != MSDA_F32 && feat_dtype != MSDA_BF16) return fail(MSDA_ERR_DTYPE, "lqe: unknown dtype %d", feat_dtype);
    if (B <= 0 || P <= 0 || Hf <= 0 || Wf <= 0 || Hf > 16384 || Wf > 16384)
        return fail(MSDA_ERR_SHAPE, "lqe: bad size B=%d P=%d map %dx%d", B, P, Hf, Wf);
    if (!msda::lqe_supported(C, K))
        return fail(MSDA_ERR_SHAPE, "lqe: C=%d (128, 256, 384, 512) or K=%d (1..8) unsupported", C, K);
    if (coord_mode != MSDA_COORD_UNFUSED && coord_mode != MSDA_COORD_FMA)
        return fail(MSDA_ERR_SHAPE, "lqe: unknown coord_mode %d", coord_mode);
    for (int i = 0; i < 4; ++i)
        if (strides[i] < 0) return fail(MSDA_ERR_SHAPE, "lqe: negative feat stride");
    if (reinterpret_cast<uintptr_t>(poses) & 7u) return fail(MSDA_ERR_ALIGN, "lqe: poses must be 8-byte aligned");
    return MSDA_OK;
}
}  // namespace

MSDA_API int msda_b200_lqe_forward(const void* feat, int feat_dtype, const int64_t* feat_strides, const float* poses,
                                   float* stat, int32_t* topk_idx, int B, int C, int Hf, int Wf, int P, int K,
                                   int coord_mode, void* stream) {
    if (int rc = check_lqe(feat, feat_dtype, feat_strides, poses, B, C, Hf, Wf, P, K, coord_mode)) return rc;
    if (stat == nullptr) return fail(MSDA_ERR_NULL, "lqe: stat is NULL");
    const cudaError_t e = msda::lqe_forward(feat, feat_dtype == MSDA_BF16, feat_strides, poses, stat, topk_idx, B, C,
                                            Hf, Wf, P, K, coord_mode, (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "lqe forward launch");
}

MSDA_API int msda_b200_lqe_backward(const void* feat, int feat_dtype, const int64_t* feat_strides, const float* poses,
                                    const int32_t* topk_idx, const float* grad_stat, float* grad_feat,
                                    float* grad_poses, int B, int C, int Hf, int Wf, int P, int K, int coord_mode,
                                    void* stream) {
    if (int rc = check_lqe(feat, feat_dtype, feat_strides, poses, B, C, Hf, Wf, P, K, coord_mode)) return rc;
    if (topk_idx == nullptr || grad_stat == nullptr) return fail(MSDA_ERR_NULL, "lqe backward: NULL topk_idx / grad_stat");
    if (reinterpret_cast<uintptr_t>(grad_poses) & 7u) return fail(MSDA_ERR_ALIGN, "lqe: grad_poses must be 8-byte aligned");
    const cudaError_t e = msda::lqe_backward(feat, feat_dtype == MSDA_BF16, feat_strides, poses, topk_idx, grad_stat,
                                             grad_feat, grad_poses, B, C, Hf, Wf, P, K, coord_mode,
                                             (cudaStream_t)stream);
    return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "lqe backward launch");
}

}  // extern "C"
