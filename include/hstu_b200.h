/*
 * hstu_b200.h — C ABI of the B200-native HSTU jagged attention (librecsys_b200.so).
 *
 * Drop-in boundary: what the reference's Python (`hstu.hstu_attn_varlen_func`,
 * third_party/FBGEMM/fbgemm_gpu/experimental/hstu/hstu/cuda_hstu_attention.py:677-775 -> HstuAttnVarlenFunc :248-670)
 * calls on sm_100: `hstu_varlen_fwd_100` / `hstu_varlen_bwd_100`
 * (src/hstu_blackwell/hstu_ops_gpu.py:85-252, :257-512).  Re-cut as plain C: raw DEVICE pointers, element strides,
 * a cudaStream_t passed as void*, no allocation inside, 0 / negative error code, nothing synchronises.
 *
 * Tensors: q, k, v are bf16 (T, H, D) with unit last stride and arbitrary token/head strides that are multiples of 8
 * elements (views of the fused (T, 4*H*D) uvqk buffer are consumed in place, examples/hstu/ops/fused_hstu_op.py:494-501);
 * out / dq / dk / dv are bf16 (T, H, D) contiguous.  cu_seqlens int32 (B+1) — cu_seqlens_q == cu_seqlens_k (self attention,
 * the training case).  num_contexts / num_targets int32 (B), nullable.  window (-1,0) = causal, (-1,-1) = full, else local.
 * Math: O_i = 1/scaling_seqlen * sum_{j in mask(i)} silu(alpha * q_i.k_j) v_j   (mask: src/hstu_blackwell/mask.py:61-127).
 */
#ifndef HSTU_B200_H_
#define HSTU_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HSTU_ERR_ARG (-1100)
#define HSTU_ERR_UNSUPPORTED (-1101)
#define HSTU_ERR_WORKSPACE (-1102)

/* Workspace: the persistent kernels pull tiles from a global counter that lives in caller memory (hstu_workspace_bytes() bytes of
 * device memory per call, contents irrelevant; must stay alive until the call's kernels have run).  This is the "workspace-size query"
 * of SURVEY 8(b); there is no dQ accumulation workspace (see hstu_bwd_sm100). */
int64_t hstu_workspace_bytes(void);

/* strides[6] = {q_token, q_head, k_token, k_head, v_token, v_head} in elements */
int hstu_fwd_sm100(const void* q, const void* k, const void* v, void* out, const int32_t* cu_seqlens, const int32_t* num_contexts,
                   const int32_t* num_targets, int batch, int heads, int head_dim, int total_tokens, int max_seqlen, int scaling_seqlen,
                   int target_group_size, int window_left, int window_right, float alpha, const int64_t* strides, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* strides[8] = q,k,v as above + {do_token, do_head}.  dq/dk/dv contiguous (T,H,D) bf16.  No gradient workspace: dK/dV come from a
 * KV-stationary kernel and dQ from a Q-stationary kernel, so nothing is accumulated through global memory (the reference
 * zero-fills and reduce-adds a dense fp32 [B,H,max_seqlen,D] dQ workspace every call, hstu_ops_gpu.py:373-382). */
int hstu_bwd_sm100(const void* dout, const void* q, const void* k, const void* v, void* dq, void* dk, void* dv, const int32_t* cu_seqlens,
                   const int32_t* num_contexts, const int32_t* num_targets, int batch, int heads, int head_dim, int total_tokens,
                   int max_seqlen, int scaling_seqlen, int target_group_size, int window_left, int window_right, float alpha,
                   const int64_t* strides, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------------------
 * HSTU layer glue (csrc/hstu_glue.cu): the row-wise stages between the GEMMs and the attention kernel of FusedHSTULayerFunction
 * (examples/hstu/ops/fused_hstu_op.py:196-251, :421-483 and its backward).  Replace the reference's Triton kernels:
 *   hstu_layer_norm_fwd / _bwd        <- triton_weighted_layer_norm_fwd / _bwd      (ops/triton_ops/triton_layer_norm.py:313, :386)
 *   hstu_ln_mul_dropout_fwd / _bwd    <- triton_layer_norm_mul_dropout_fwd / _bwd   (ops/triton_ops/triton_norm_mul_dropout.py:361, :426; concat_ux=False)
 *   hstu_silu_fwd / _bwd              <- triton_silu_fwd / _bwd                     (ops/triton_ops/triton_silu.py:91, :108)
 * Conventions: raw DEVICE pointers, 16-byte aligned; rows = tokens; D = normalised width, multiple of 8, <= 8192 (HSTU_ERR_UNSUPPORTED
 * above; <= 1024 a warp owns a row, above that a CTA does — the row is register resident either way); strides in ELEMENTS, multiples of 8 (strided views of the fused uvqk buffer are read in place);
 * dtype 0 = fp32, 1 = fp16, 2 = bf16 for every tensor argument of a call (weight / bias included), statistics (mean, rstd) and the
 * weight / bias gradients are fp32; math is fp32; nothing allocates or synchronises.  weight / bias may be NULL (plain normalisation).
 * Backward calls need hstu_glue_workspace_bytes(D) bytes of device scratch (per-CTA dw / db partials, summed in a fixed order).
 * Dropout (training != 0 and ratio > 0): element (row, col) is kept iff 16 bits of Philox4x32-10(seed; row, col / 8) >= round(ratio * 65536),
 * kept values are scaled by 1 / (1 - round(ratio * 65536) / 65536); forward and backward regenerate the mask from (seed, row, col). */
int64_t hstu_glue_workspace_bytes(int D);
/* y = (x - mean) * rstd * weight + bias;  mean / rstd [rows] are outputs */
int hstu_layer_norm_fwd(const void* x, int64_t x_stride, const void* weight, const void* bias, void* y, int64_t y_stride, float* mean, float* rstd,
                        int64_t rows, int D, float eps, int dtype, void* stream);
/* dx = LN backward of dy (+ dx_add, nullable: the residual branch's gradient);  dweight / dbias [D] fp32, nullable */
int hstu_layer_norm_bwd(const void* dy, int64_t dy_stride, const void* x, int64_t x_stride, const void* weight, const float* mean, const float* rstd,
                        const void* dx_add, int64_t dx_add_stride, void* dx, int64_t dx_stride, float* dweight, float* dbias, void* workspace,
                        int64_t workspace_bytes, int64_t rows, int D, int dtype, void* stream);
/* y = dropout(LN(x) * u) */
int hstu_ln_mul_dropout_fwd(const void* x, int64_t x_stride, const void* u, int64_t u_stride, const void* weight, const void* bias, void* y,
                            int64_t y_stride, float* mean, float* rstd, int64_t rows, int D, float eps, float dropout_ratio, uint64_t seed,
                            int training, int dtype, void* stream);
/* dx, du, dweight, dbias of the above in one pass; y_out (nullable) = the forward output recomputed (compute_y of the reference) */
int hstu_ln_mul_dropout_bwd(const void* dy, int64_t dy_stride, const void* x, int64_t x_stride, const void* u, int64_t u_stride, const void* weight,
                            const void* bias, const float* mean, const float* rstd, void* dx, int64_t dx_stride, void* du, int64_t du_stride,
                            void* y_out, int64_t y_stride, float* dweight, float* dbias, void* workspace, int64_t workspace_bytes, int64_t rows,
                            int D, float dropout_ratio, uint64_t seed, int training, int dtype, void* stream);
/* y = x * sigmoid(x) over n contiguous elements (n multiple of 8) */
int hstu_silu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream);
/* dx[rows, W] = dy * silu'(x) with dy given as 1..4 column segments (HOST arrays: base pointer, row stride, width; W = sum of widths,
 * each a multiple of 8) — du / dv / dq / dk are read where their producers left them; x, dx contiguous [rows, W] */
int hstu_silu_bwd(int num_segments, const void* const* seg_ptr, const int64_t* seg_stride, const int32_t* seg_width, const void* x, void* dx,
                  int64_t rows, int dtype, void* stream);
/* the same with the column sums of dx on the way: dbias[W] (fp32) = sum over rows of dx — the bias gradient of the GEMM in front of the SiLU
 * (`dy = torch.sum(dz, dim=0)` in triton_addmm_silu_bwd, triton_addmm.py:293-294: a separate pass over dz there).  Needs
 * hstu_silu_bwd_bias_workspace_bytes(W) bytes of scratch; that query returns 0 when W / 8 cannot be made to divide the launch (then call
 * hstu_silu_bwd and sum separately; HSTU_ERR_UNSUPPORTED from this entry point). */
int64_t hstu_silu_bwd_bias_workspace_bytes(int W);
int hstu_silu_bwd_bias(int num_segments, const void* const* seg_ptr, const int64_t* seg_stride, const int32_t* seg_width, const void* x, void* dx,
                       float* dbias, void* workspace, int64_t workspace_bytes, int64_t rows, int dtype, void* stream);
/* test aid: keep[rows, D] (uint8) = the dropout mask the two ln_mul_dropout calls apply for (ratio, seed) */
int hstu_dropout_mask(int64_t rows, int D, float dropout_ratio, uint64_t seed, uint8_t* keep, void* stream);

/* development aid: one-CTA tcgen05 GEMM that pins the descriptor conventions (csrc/sm100_probe.cu) */
int sm100_probe_gemm(const void* A, const void* B, float* C, int variant, const uint32_t* overrides, void* stream);
/* development aid: install (or clear with NULL) a cycle-accounting buffer of 128 int32.  While installed, hstu_fwd_sm100 runs its
   instrumented instantiation and adds per-role wait / work cycle counts of every CTA to the buffer (DEVICE memory, zeroed by the
   caller); hstu_bwd_sm100 stores the counts of CTA (0,0,0) (device or host-mapped memory).  tools/hstu_cycles.py prints them.
   Never installed on the product path. */
int hstu_set_debug_buffer(int* buffer);

#ifdef __cplusplus
}
#endif
#endif
