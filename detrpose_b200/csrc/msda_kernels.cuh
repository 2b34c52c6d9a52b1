// Internal launcher declarations shared by the C-ABI translation unit.
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kFwdThreads = 256;
constexpr int kBwdThreads = 256;

// msda_fwd.cu
cudaError_t forward_flat(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                         const float* attn, void* out, bool out_bf16, cudaStream_t st);

bool forward_lean_supported(const Problem& pb, bool value_bf16);
// ref != nullptr: fused prologue -- loc = sampling offsets, attn = attention logits (see msda_fwd.cu)
cudaError_t forward_lean(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                         const float* attn, void* out, bool out_bf16, cudaStream_t st,
                         const float* ref = nullptr, int ref_levels = 1, float* attn_out = nullptr,
                         int pair_mode = -1,        // two lane groups per item (rows <= 32 bytes): -1 auto, 0 off
                         int l2_prefetch = 1);      // streaming L2 prefetch of the pyramid ahead of the gathers

// msda_fwd_staged.cu
bool forward_staged_supported(const Problem& pb, bool value_bf16, bool small);
cudaError_t forward_staged(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                           const float* attn, void* out, bool out_bf16, bool small, cudaStream_t st);

// msda_bwd.cu
cudaError_t backward_flat(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                          const float* attn, const void* grad_out, bool go_bf16, float* grad_value,
                          float* grad_loc, float* grad_attn, cudaStream_t st);

// msda_bwd_gather.cu
bool backward_gather_supported(const Problem& pb, bool value_bf16);
// ref != nullptr: fused prologue -- loc = sampling offsets, grad_loc receives the gradient w.r.t. the offsets
cudaError_t backward_gather(const Problem& pb, const void* value, bool value_bf16, const float* loc,
                            const float* attn, const void* grad_out, float* grad_value, float* grad_loc,
                            float* grad_attn, int accumulate, int threads_pref, cudaStream_t st,
                            const float* ref = nullptr, int ref_levels = 1);
// grad_logits = a * (g - sum_j a_j g_j) per row of `cols` softmaxed weights (ms_deform_attn.py:393 backward)
cudaError_t softmax_backward(const float* attn, const float* grad_attn, float* grad_logits, int64_t rows,
                             int cols, cudaStream_t st);

cudaError_t set_phase_buffer(unsigned long long* buf);

// msda_aux.cu
cudaError_t sample_indices(const Problem& pb, const float* loc, int32_t* idx_out,
                           int32_t* level_start_out, cudaStream_t st);
cudaError_t locations(const Problem& pb, const float* offsets, const float* logits, const float* ref,
                      int ref_levels, float* loc, float* attn, cudaStream_t st);
struct LevelViews {
    const void* ptr[MSDA_MAX_LEVELS];
    int64_t s_nh[MSDA_MAX_LEVELS], s_c[MSDA_MAX_LEVELS], s_s[MSDA_MAX_LEVELS];
};
cudaError_t repack(const Problem& pb, const LevelViews& src, bool src_bf16, void* dst, bool dst_bf16,
                   cudaStream_t st);
cudaError_t unpack_grad(const Problem& pb, const float* grad_value, const LevelViews& dst, bool dst_bf16,
                        cudaStream_t st);

// msda_gate.cu
bool gate_supported(int C);
cudaError_t gate_forward(const void* pre, bool pre_bf16, const void* x1, const void* x2, bool x_bf16,
                         const float* gamma, const float* beta, float eps, void* y, float* stats, int64_t rows,
                         int C, cudaStream_t st);
cudaError_t gate_backward(const void* pre, bool pre_bf16, const void* x1, const void* x2, bool x_bf16,
                          const float* gamma, const float* stats, const void* gy, void* gpre, void* gx1, void* gx2,
                          float* ggamma, float* gbeta, int64_t rows, int C, int sm_count, cudaStream_t st);

// msda_lqe.cu
bool lqe_supported(int C, int K);
cudaError_t lqe_forward(const void* feat, bool feat_bf16, const int64_t* strides, const float* poses, float* stat,
                        int32_t* topk_idx, int B, int C, int Hf, int Wf, int P, int K, int coord_mode,
                        cudaStream_t st);
cudaError_t lqe_backward(const void* feat, bool feat_bf16, const int64_t* strides, const float* poses,
                         const int32_t* topk_idx, const float* grad_stat, float* grad_feat, float* grad_poses,
                         int B, int C, int Hf, int Wf, int P, int K, int coord_mode, cudaStream_t st);

}  // namespace msda
