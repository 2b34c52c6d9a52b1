/*
 * msda_b200.h -- C ABI of the B200 (sm_100a) multi-scale deformable attention
 * sampling core.  This is the drop-in boundary for DETRPose's hot path.
 *
 * The reference has no FFI of its own (it is pure PyTorch); the entry points
 * below are what a binding for this path would bind, one per reference
 * function they replace:
 *
 *   msda_b200_forward   <- ms_deform_attn_core_pytorch, forward
 *                          /root/reference/src/models/detrpose/ms_deform_attn.py:145-193
 *                          (L x F.grid_sample :178 + cat :184 + mul/sum :192)
 *   msda_b200_backward  <- autograd of the same function (ATen
 *                          grid_sampler_2d_backward + mul/sum/cat backward)
 *   msda_b200_locations <- the location/softmax prologue of MSDeformAttn.forward
 *                          ms_deform_attn.py:392-393 (softmax over L*P) and
 *                          :412-416 (ref + offsets / [W_l, H_l])
 *   msda_b200_repack / msda_b200_unpack_grad
 *                       <- the value construction of the caller,
 *                          /root/reference/src/models/detrpose/transformer.py:1285-1286
 *                          (permute + flatten + split), inverted: strided
 *                          per-level views -> one channel-last pyramid
 *   msda_b200_sample_indices
 *                       <- the integer corner indices ATen derives inside
 *                          grid_sampler_2d (GridSampler.h:27-35 + floor);
 *                          exported so index parity can be tested bit-exactly
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types; every *device* pointer is
 *     owned by the caller (PyTorch allocates and frees), the library keeps no
 *     state between calls apart from a per-thread error string;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never synchronises, is re-entrant and thread-safe (autograd calls
 *     backward from worker threads);
 *   - returns 0 on success, a negative MSDA_ERR_* for argument errors, a
 *     positive cudaError_t for CUDA failures; never throws or aborts.
 *     msda_b200_last_error() returns a per-thread description;
 *   - "host" arrays are read during the call and may be freed on return.
 *
 * Data layout in HBM (kernel-native)
 *   value pyramid : channel-last, element (n, s, h, c) at
 *                   base + n*stride_n + s*stride_s + h*stride_h + c  (c contiguous),
 *                   s = level_start[l] + y*W_l + x.  The reference's `memory`
 *                   tensor (N, S, C) is exactly this with strides (S*C, C, Dh).
 *   locations     : (N, Lq, H, L, P, 2) fp32 contiguous, last axis (x, y) in [0,1] units
 *   attention     : (N, Lq, H, L, P)    fp32 contiguous
 *   output        : (N, Lq, H*Dh) contiguous, fp32 or bf16
 *   grad_value    : fp32 buffer (N, S, H, Dh) contiguous; the backward either
 *                   overwrites it or adds into it (`accumulate`), so decoder
 *                   layers that share `value` can accumulate in place.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define MSDA_API __attribute__((visibility("default")))
#else
#define MSDA_API
#endif

/* element types */
#define MSDA_F32  0
#define MSDA_BF16 1

/* argument-check error codes (negative); CUDA errors are returned as positive cudaError_t */
#define MSDA_OK              0
#define MSDA_ERR_NULL       -1   /* a required pointer is NULL */
#define MSDA_ERR_SHAPE      -2   /* a size is <= 0 or exceeds the supported range */
#define MSDA_ERR_DTYPE      -3   /* unknown element type code */
#define MSDA_ERR_DHEAD      -4   /* Dh not supported (must be a multiple of 8, 8..128) */
#define MSDA_ERR_ALIGN      -5   /* a base pointer / stride breaks the 16-byte row alignment */
#define MSDA_ERR_LEVELS     -6   /* L or P exceeds MSDA_MAX_LEVELS / MSDA_MAX_POINTS */
#define MSDA_ERR_NO_DEVICE  -7   /* no sm_100 device / kernel image not loadable */

#define MSDA_MAX_LEVELS 8
#define MSDA_MAX_POINTS 16

/* Rounding chain used to turn a normalised location into a pixel coordinate.
 * UNFUSED reproduces the reference's elementwise fp32 ops one rounding per op
 * (2*loc-1, +1, *size, -1, /2); FMA contracts "(g+1)*size - 1" into one
 * fused multiply-add the way nvcc compiles ATen's CUDA grid sampler. */
#define MSDA_COORD_UNFUSED 0
#define MSDA_COORD_FMA     1

MSDA_API int msda_b200_abi_version(void);
MSDA_API const char* msda_b200_last_error(void);

/* Number of SMs / name of the device the library will launch on (current device). */
MSDA_API int msda_b200_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Tuning knob for benchmarks: pin the forward / backward kernel variant
 * (-1 = automatic selection, the default).  Process-wide, atomic.
 * forward : 0 flat, 1 lean (the automatic choice when the shape fits), 2 / 3 TMA-staged coarse levels (one big
 *           CTA per SM / small CTAs), 7 lean with one lane group per item also for rows of at most 32 bytes;
 *           100 + v = variant v without the streaming L2 prefetch of the pyramid (150: automatic without it)
 * backward: 0 flat (vector reductions), 1 gather form (automatic when the shape fits), 2 gather form with
 *           512 threads x 128 registers
 * A pinned variant that does not support a shape makes the call fail (MSDA_ERR_SHAPE), it never falls back. */
MSDA_API int msda_b200_set_variant(int fwd_variant, int bwd_variant);

/* Diagnostic: device buffer of 8 uint64 per backward CTA (N*H*L CTAs) that receives clock64()
 * at the phase boundaries of the gather-form backward; NULL (default) disables it. */
MSDA_API int msda_b200_debug_phase_buffer(void* device_buffer);

/*
 * Forward: out[n,q,h*Dh+c] = sum_{l,p} attn[n,q,h,l,p] * bilinear_zero_pad(value_l[n,h,c], loc[n,q,h,l,p])
 *
 * value          device, channel-last pyramid (see layout above), value_dtype
 * value_strides  host[3]: (stride_n, stride_s, stride_h) in ELEMENTS; channel stride is 1
 * spatial_shapes host[2*L]: (H_l, W_l); level starts are the running sum of H_l*W_l
 */
MSDA_API int msda_b200_forward(const void* value, int value_dtype, const int64_t* value_strides,
                      const int32_t* spatial_shapes,
                      const float* locations, const float* attention,
                      void* out, int out_dtype,
                      int N, int Lq, int H, int Dh, int L, int P,
                      int coord_mode, void* stream);

/*
 * Backward of the above for all three inputs.
 *
 * grad_out        device (N, Lq, H*Dh) contiguous, grad_out_dtype
 * grad_value      device fp32 (N, S, H, Dh) contiguous (may be NULL: skip)
 * accumulate      0: grad_value is OVERWRITTEN (every element written, the caller does not zero-fill);
 *                 1: the gradient is ADDED to the buffer (decoder layers sharing one `value`)
 * grad_locations  device fp32, shape of locations, overwritten (may be NULL together with grad_attention);
 *                 base pointer 32-BYTE aligned (written with 256-bit stores), else MSDA_ERR_ALIGN
 * grad_attention  device fp32, shape of attention, overwritten; 16-byte aligned like every other buffer
 */
MSDA_API int msda_b200_backward(const void* value, int value_dtype, const int64_t* value_strides,
                       const int32_t* spatial_shapes,
                       const float* locations, const float* attention,
                       const void* grad_out, int grad_out_dtype,
                       float* grad_value, int accumulate, float* grad_locations, float* grad_attention,
                       int N, int Lq, int H, int Dh, int L, int P,
                       int coord_mode, void* stream);

/*
 * Integer corner indices: idx[(n,q,h,l,p)] = (y0, x0) int32, the floor of the
 * pixel coordinate, exactly as the forward/backward kernels compute it.
 * level_start_out: optional DEVICE int32[L], receives the level offsets.
 */
MSDA_API int msda_b200_sample_indices(const int32_t* spatial_shapes, const float* locations,
                             int32_t* idx_out, int32_t* level_start_out,
                             int N, int Lq, int H, int L, int P,
                             int coord_mode, void* stream);

/*
 * Location/softmax prologue (ms_deform_attn.py:392-393, :412-416):
 *   attention = softmax_{L*P}(logits);  locations = ref + offsets / (W_l, H_l)
 *
 * offsets    device fp32 (N, Lq, H, L, P, 2)   (output of the offsets Linear)
 * logits     device fp32 (N, Lq, H, L*P)       (output of the attention Linear)
 * ref_points device fp32 (N, Lq, ref_levels, 2), ref_levels is 1 (broadcast) or L
 */
MSDA_API int msda_b200_locations(const float* offsets, const float* logits, const float* ref_points,
                        int ref_levels, const int32_t* spatial_shapes,
                        float* locations, float* attention,
                        int N, int Lq, int H, int L, int P, void* stream);

/*
 * Sampler with the prologue fused in.  Replaces, in one launch per direction,
 * ms_deform_attn.py:392-393 (softmax over L*P), :412-416 (ref + offsets / (W_l, H_l), 2-D reference points)
 * and the sampling core :145-193.  Locations and weights are computed with the arithmetic of
 * msda_b200_locations (divide and add separately rounded: corner indices stay bit-exact) and never reach HBM.
 *
 * offsets        device fp32 (N, Lq, H, L, P, 2)  (output of the sampling_offsets Linear), 16-byte aligned
 * logits         device fp32 (N, Lq, H, L*P)      (output of the attention_weights Linear), 16-byte aligned
 * ref_points     device fp32 (N, Lq, ref_levels, 2), ref_levels 1 (broadcast over levels) or L, 8-byte aligned
 * attention_out  device fp32 (N, Lq, H, L, P) or NULL: the softmaxed weights, kept for the backward
 *
 * Shapes the fused kernels do not cover return MSDA_ERR_SHAPE (query with msda_b200_fused_supported; the
 * caller then runs msda_b200_locations + msda_b200_forward / _backward).
 */
MSDA_API int msda_b200_fused_supported(int value_dtype, const int64_t* value_strides,
                              const int32_t* spatial_shapes, int N, int Lq, int H, int Dh, int L, int P);

MSDA_API int msda_b200_forward_fused(const void* value, int value_dtype, const int64_t* value_strides,
                            const int32_t* spatial_shapes,
                            const float* offsets, const float* logits, const float* ref_points, int ref_levels,
                            void* out, int out_dtype, float* attention_out,
                            int N, int Lq, int H, int Dh, int L, int P,
                            int coord_mode, void* stream);

/*
 * Backward of the fused sampler.  `attention` is attention_out of the forward.  grad_offsets (shape of
 * offsets, 32-byte aligned) is the gradient w.r.t. the sampling offsets -- the (W_l, H_l) of the pixel
 * mapping and of `offsets / (W_l, H_l)` cancel; grad_attention is the gradient w.r.t. the softmaxed weights:
 * msda_b200_softmax_backward turns it into the gradient w.r.t. the logits (the sum over L*P spans several
 * CTAs of the backward, hence a pass of its own).  The gradient w.r.t. ref_points is the sum of
 * grad_offsets * (W_l, H_l) over heads, (levels,) points; DETRPose detaches them (transformer.py:1246).
 */
MSDA_API int msda_b200_backward_fused(const void* value, int value_dtype, const int64_t* value_strides,
                             const int32_t* spatial_shapes,
                             const float* offsets, const float* ref_points, int ref_levels,
                             const float* attention,
                             const void* grad_out, int grad_out_dtype,
                             float* grad_value, int accumulate, float* grad_offsets, float* grad_attention,
                             int N, int Lq, int H, int Dh, int L, int P,
                             int coord_mode, void* stream);

/* grad_logits[r, i] = attention[r, i] * (grad_attention[r, i] - sum_j attention[r, j] * grad_attention[r, j])
 * for `rows` = N*Lq*H rows of `cols` = L*P contiguous fp32 values (all three buffers 16-byte aligned;
 * grad_logits may alias grad_attention). */
MSDA_API int msda_b200_softmax_backward(const float* attention, const float* grad_attention, float* grad_logits,
                               int64_t rows, int cols, void* stream);

/*
 * Repack the reference's per-level strided views into one channel-last pyramid.
 *
 * level_ptrs     host[L] of device pointers: element (nh=0, c=0, s=0) of value[l]
 * level_strides  host[3*L]: strides of value[l] (N*H, Dh, H_l*W_l) in elements
 * dst            device (N, S, H, Dh) contiguous, dst_dtype (conversion allowed)
 */
MSDA_API int msda_b200_repack(const void* const* level_ptrs, const int64_t* level_strides, int src_dtype,
                     const int32_t* spatial_shapes, void* dst, int dst_dtype,
                     int N, int H, int Dh, int L, void* stream);

/*
 * Inverse of repack for the gradient: fp32 channel-last (N, S, H, Dh) ->
 * per-level (N*H, Dh, H_l*W_l) buffers with the given strides and dtype.
 */
MSDA_API int msda_b200_unpack_grad(const float* grad_value, const int32_t* spatial_shapes,
                          void* const* level_ptrs, const int64_t* level_strides, int dst_dtype,
                          int N, int H, int Dh, int L, void* stream);

/*
 * Gate epilogue -- replaces everything after the Linear in
 * Gate.forward, transformer.py:233-235 (sigmoid, chunk, blend, LayerNorm with affine):
 *   y = LayerNorm(sigmoid(pre[:, :C]) * x1 + sigmoid(pre[:, C:]) * x2) * gamma + beta
 *
 * pre     device (rows, 2*C) contiguous, pre_dtype: output of the 2C->2C Linear, BEFORE the sigmoid
 * x1, x2  device (rows, C) contiguous, x_dtype (the layer input and the sampler output, :429)
 * gamma, beta  device fp32 [C] (LayerNorm weight / bias); eps as in nn.LayerNorm
 * y       device (rows, C) contiguous, x_dtype
 * stats   device fp32 (rows, 2) receiving (mean, rstd) per row for the backward, or NULL (inference)
 * C in {128, 256, 384, 512}; all pointers 16-byte aligned.
 */
MSDA_API int msda_b200_gate_forward(const void* pre, int pre_dtype, const void* x1, const void* x2, int x_dtype,
                           const float* gamma, const float* beta, float eps,
                           void* y, float* stats, int64_t rows, int C, void* stream);

/*
 * Backward of the gate epilogue (the reference gets it from autograd over the same ops).
 * grad_y (rows, C) x_dtype;  outputs, all overwritten: grad_pre (rows, 2*C) pre_dtype,
 * grad_x1 / grad_x2 (rows, C) x_dtype, grad_gamma / grad_beta fp32 [C] (zero-filled by the
 * library, then summed over rows with fp32 atomics: summation order is not fixed).
 */
MSDA_API int msda_b200_gate_backward(const void* pre, int pre_dtype, const void* x1, const void* x2, int x_dtype,
                            const float* gamma, const float* stats, const void* grad_y,
                            void* grad_pre, void* grad_x1, void* grad_x2,
                            float* grad_gamma, float* grad_beta, int64_t rows, int C, void* stream);

/*
 * LQE sampler -- replaces transformer.py:278-284 (grid_sample of the finest
 * feature map at the predicted keypoints, permute, top-k over channels, mean, cat):
 *   stat[b, p, 0..K-1] = the K largest of { bilinear(feat[b, c], poses[b, p]) : c < C }, descending
 *   stat[b, p, K]      = their mean
 *
 * feat          device (B, C, Hf, Wf), feat_dtype, addressed through feat_strides (host[4], elements):
 *               NCHW and channels-last tensors are both read in place
 * poses         device fp32 (B, P, 2), normalised (x, y) as in the sampler (pixel = pose * size - 0.5,
 *               zero padding per corner); P = queries * keypoints
 * stat          device fp32 (B, P, K + 1)
 * topk_idx      device int32 (B, P, K): channel of each kept value (lowest channel on ties), needed by
 *               the backward; may be NULL
 * C in {128, 256, 384, 512}, 1 <= K <= 8.
 */
MSDA_API int msda_b200_lqe_forward(const void* feat, int feat_dtype, const int64_t* feat_strides,
                          const float* poses, float* stat, int32_t* topk_idx,
                          int B, int C, int Hf, int Wf, int P, int K, int coord_mode, void* stream);

/*
 * Backward of the LQE sampler.  grad_stat fp32 (B, P, K+1).
 * grad_feat   device fp32 with feat's shape AND strides; the gradient is ADDED (the caller zero-fills
 *             or accumulates; scalar fp32 atomics, 4*K per keypoint); may be NULL
 * grad_poses  device fp32 (B, P, 2), overwritten; may be NULL
 */
MSDA_API int msda_b200_lqe_backward(const void* feat, int feat_dtype, const int64_t* feat_strides,
                           const float* poses, const int32_t* topk_idx, const float* grad_stat,
                           float* grad_feat, float* grad_poses,
                           int B, int C, int Hf, int Wf, int P, int K, int coord_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
