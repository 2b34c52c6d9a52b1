// Backward, gather form: plan, dispatch and the softmax backward of the fused prologue.  The kernel itself is in
// msda_bwd_gather.cuh; its instantiations are compiled in msda_bwd_gather_case*.cu.
#include <atomic>

#include "msda_bwd_gather.cuh"

namespace msda {

static std::atomic<unsigned long long*> g_phase_buf_host{nullptr};
unsigned long long* phase_buffer() { return g_phase_buf_host.load(); }
cudaError_t set_phase_buffer(unsigned long long* buf) {
    g_phase_buf_host.store(buf);
    return cudaSuccess;
}

extern template cudaError_t gather_case<1, 1, 1024>(bool, const GatherArgs&);
extern template cudaError_t gather_case<2, 1, 1024>(bool, const GatherArgs&);
extern template cudaError_t gather_case<1, 3, 512>(bool, const GatherArgs&);
extern template cudaError_t gather_case<4, 1, 1024>(bool, const GatherArgs&);
extern template cudaError_t gather_case<2, 3, 512>(bool, const GatherArgs&);
extern template cudaError_t gather_case<4, 2, 1024>(bool, const GatherArgs&);
extern template cudaError_t gather_case<4, 3, 512>(bool, const GatherArgs&);
extern template cudaError_t gather_case<8, 2, 512>(bool, const GatherArgs&);
extern template cudaError_t gather_case<2, 1, 512>(bool, const GatherArgs&);
extern template cudaError_t gather_case<4, 1, 512>(bool, const GatherArgs&);
extern template cudaError_t gather_case<8, 1, 512>(bool, const GatherArgs&);

// Returns false when the shape does not fit the gather kernel (caller falls back to the flat one).
static bool make_plan(const Problem& pb, int row_bytes, GatherPlan& plan) {
    int max_bins = 0;
    for (int l = 0; l < pb.L; ++l) max_bins = max(max_bins, (pb.geom.w[l] + 1) * (pb.geom.h[l] + 1));
    plan.bins_words = (max_bins + 3) / 2;
    // the kernel addresses a level's value rows / gradient rows with 32-bit offsets from the level base
    const int es = row_bytes / pb.Dh;
    for (int l = 0; l < pb.L; ++l) {
        const int64_t npix = (int64_t)pb.geom.w[l] * pb.geom.h[l];
        if (npix * pb.vs_s * es >= (int64_t)0xffffffffLL || npix * pb.H * pb.Dh >= (int64_t)0x7fffffffLL) return false;
    }
    const int fixed = align16(plan.bins_words * 4) + 64 * 4 + 32 * 64 * 2 + 16 + row_bytes + 16;
    const int per_query = row_bytes + pb.P * 32;
    int qmax = (kMaxSmem - fixed) / per_query;
    qmax = min(qmax, min(65535 / pb.P, 4095));       // u16 counters / sample slots, 12-bit query ids
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
                            float* grad_attn, int accumulate, int threads_pref, cudaStream_t st,
                            const float* ref, int ref_levels) {
    GatherPlan plan;
    if (!make_plan(pb, pb.Dh * (value_bf16 ? 2 : 4), plan)) return cudaErrorInvalidValue;
    const int nv = pb.Dh * (value_bf16 ? 2 : 4) / 16;
    const GatherArgs a{&pb, &plan, value, loc, attn, grad_out, grad_value, grad_loc, grad_attn, accumulate, st, ref, ref_levels};
    if (threads_pref == 512) {          // 512 threads x 128 registers instead of 1024 x 64
        switch (nv) {
            case 2: return gather_case<2, 1, 512>(value_bf16, a);
            case 4: return gather_case<4, 1, 512>(value_bf16, a);
            case 8: return gather_case<8, 1, 512>(value_bf16, a);
            default: break;
        }
    }
    switch (nv) {
        case 1: return gather_case<1, 1, 1024>(value_bf16, a);
        case 2: return gather_case<2, 1, 1024>(value_bf16, a);
        case 3: return gather_case<1, 3, 512>(value_bf16, a);
        case 4: return gather_case<4, 1, 1024>(value_bf16, a);
        case 6: return gather_case<2, 3, 512>(value_bf16, a);
        case 8: return gather_case<4, 2, 1024>(value_bf16, a);
        case 12: return gather_case<4, 3, 512>(value_bf16, a);
        case 16: return gather_case<8, 2, 512>(value_bf16, a);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------
// Softmax backward of the fused prologue: grad_logits = a * (g - sum_j a_j g_j) over the L*P weights of one
// (n, q, h) (ms_deform_attn.py:393).  The sum spans all levels, i.e. several CTAs of the gather kernel, so
// it is a pass of its own: one thread per row, rows are consecutive 16-byte-aligned runs of `cols` floats.
// ---------------------------------------------------------------------------
template <int COLS>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ attn, const float* __restrict__ grad_attn,
                   float* __restrict__ grad_logits, const int64_t rows, const int cols_rt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int cols = COLS > 0 ? COLS : cols_rt;
    const float* a = attn + r * cols;
    const float* g = grad_attn + r * cols;
    float* o = grad_logits + r * cols;
    if constexpr (COLS > 0 && COLS % 4 == 0) {
        float4 av[COLS / 4], gv[COLS / 4];
#pragma unroll
        for (int i = 0; i < COLS / 4; ++i) {
            av[i] = __ldg(reinterpret_cast<const float4*>(a) + i);
            gv[i] = __ldg(reinterpret_cast<const float4*>(g) + i);
        }
        float dot = 0.0f;
#pragma unroll
        for (int i = 0; i < COLS / 4; ++i)
            dot += av[i].x * gv[i].x + av[i].y * gv[i].y + av[i].z * gv[i].z + av[i].w * gv[i].w;
#pragma unroll
        for (int i = 0; i < COLS / 4; ++i)
            reinterpret_cast<float4*>(o)[i] = make_float4(av[i].x * (gv[i].x - dot), av[i].y * (gv[i].y - dot),
                                                          av[i].z * (gv[i].z - dot), av[i].w * (gv[i].w - dot));
    } else {
        float dot = 0.0f;
        for (int i = 0; i < cols; ++i) dot += __ldg(a + i) * __ldg(g + i);
        for (int i = 0; i < cols; ++i) o[i] = __ldg(a + i) * (__ldg(g + i) - dot);
    }
}

cudaError_t softmax_backward(const float* attn, const float* grad_attn, float* grad_logits, int64_t rows,
                             int cols, cudaStream_t st) {
    const unsigned grid = (unsigned)((rows + 255) / 256);
    switch (cols) {
        case 8: softmax_bwd_kernel<8><<<grid, 256, 0, st>>>(attn, grad_attn, grad_logits, rows, cols); break;
        case 12: softmax_bwd_kernel<12><<<grid, 256, 0, st>>>(attn, grad_attn, grad_logits, rows, cols); break;
        case 16: softmax_bwd_kernel<16><<<grid, 256, 0, st>>>(attn, grad_attn, grad_logits, rows, cols); break;
        default: softmax_bwd_kernel<0><<<grid, 256, 0, st>>>(attn, grad_attn, grad_logits, rows, cols); break;
    }
    return cudaGetLastError();
}

}  // namespace msda