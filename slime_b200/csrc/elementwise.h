// Host-side launchers of the HBM-bound (non-GEMM) kernels of the SliME prefill path.
#pragma once
#include "common.cuh"

// y = LayerNorm(x) * w + b.  Source row of output row r is
//   (r / in_group) * in_group_stride + in_offset + r % in_group      (in_group <= 0: identity)
// so the CLS-dropping slice of the ViT output can be folded into the read.
int slime_launch_layernorm(const bf16* x, int x_ld, const bf16* w, const bf16* b, bf16* y, int y_ld,
                           int rows, int D, float eps, int in_group, int in_group_stride,
                           int in_offset, cudaStream_t stream);

// HF LlamaRMSNorm: y = w * bf16(x * rsqrt(mean(x^2) + eps)).  src_rows (optional) gathers input rows.
int slime_launch_rmsnorm(const bf16* x, int x_ld, const bf16* w, bf16* y, int y_ld, int rows, int D,
                         float eps, const int* src_rows, cudaStream_t stream);

// Norm folding (gemm.h GemmParams::row_scale / sumsq_out):
//   row_rstd        : rstd[r] = rsqrt(mean(x[r, :]^2) + eps) straight from the rows (first layer: no GEMM produced them)
//   sumsq_to_rstd   : rstd[r] = rsqrt(sum(partials[r, 0..parts)) / D + eps) from the partials a residual GEMM's epilogue wrote
int slime_launch_row_rstd(const bf16* x, int x_ld, float* rstd, int rows, int D, float eps, cudaStream_t stream);
int slime_launch_sumsq_to_rstd(const float* partials, int parts, float* rstd, int rows, int D, float eps,
                               cudaStream_t stream);

// pixels [Nc,3,336,336] -> patches [Nc*576, Kpad] (k = c*196 + ky*14 + kx, zero padded to Kpad)
int slime_launch_im2col(const bf16* pixels, bf16* patches, int Nc, int image, int patch, int Kpad,
                        cudaStream_t stream);

// CLIP embeddings + pre_layrnorm: h[c,0]=LN(cls+pos[0]); h[c,1+t]=LN(patch[c,t]+pos[1+t])
int slime_launch_clip_embed_ln(const bf16* patch_out, const bf16* cls, const bf16* pos, const bf16* w,
                               const bf16* b, bf16* h, int Nc, int tokens, int D, float eps,
                               cudaStream_t stream);

// dst[r, :] = src[(r / group) * group_stride + offset + r % group, :]   (row slice / gather copy)
// vision-tower output with the global crops of all images first, the local crops behind them (elementwise.cu)
int slime_launch_vit_split_rows(const bf16* src, bf16* dst, int crops, int D, int P, int TK, int crop0, int per_image,
                                int images, cudaStream_t stream);
int slime_launch_copy_rows(const bf16* src, int src_ld, bf16* dst, int dst_ld, int rows, int D,
                           int group, int group_stride, int offset, cudaStream_t stream);
// dst[dst_rows[r], :] = src[r, :]  (row scatter; negative target drops the row)
int slime_launch_scatter_rows(const bf16* src, int src_ld, bf16* dst, int dst_ld, int rows, int D,
                              const int* dst_rows, cudaStream_t stream);

// y = a + b (row-periodic b: row % period), used to pre-compute constant query/pos sums at load time
int slime_launch_add_rows(const bf16* a, const bf16* b, bf16* y, int rows, int D, int period,
                          cudaStream_t stream);

// RoPE (rotate-half) in place on the q and k heads of a packed qkv buffer [rows, ld]
int slime_launch_rope(bf16* qkv, int ld, int rows, int n_q_heads, int n_k_heads, int head_dim,
                      const int* pos_ids, const float* cos_sin_table, int max_pos, cudaStream_t stream);
int slime_launch_rope_table(float* table, int max_pos, int head_dim, float theta, cudaStream_t stream);

// Gated mix of the two global experts:  gates = softmax(x W_g) renormalised by (sum + 1e-6)
int slime_launch_gate_mix(const bf16* x, const bf16* w_gate, const bf16* e0, const bf16* e1, bf16* out,
                          int rows, int Dm, int H, cudaStream_t stream);
