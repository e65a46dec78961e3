// Host-side interface of the tcgen05 bf16 GEMM (gemm_sm100.cu, gemm2_sm100.cu).
#pragma once
#include "common.cuh"

enum GemmEpilogue : int {
  GEMM_EPI_NONE = 0,        // out = acc (+bias) (+residual)
  GEMM_EPI_QUICK_GELU = 1,  // out = qgelu(acc + bias)          x*sigmoid(1.702x)  (CLIP MLP)
  GEMM_EPI_GELU_ERF = 2,    // out = gelu(acc + bias)           exact erf GELU     (mm_projector)
  GEMM_EPI_SWIGLU = 3,      // out[:, j] = silu(acc[:, 2j]) * acc[:, 2j+1]          (Llama MLP)
  GEMM_EPI_ROPE = 4,        // rotary embedding on the fp32 accumulators of the packed Llama QKV projection: columns
                            // < rope_cols hold the (i, i + hd/2) feature pairs of every head ADJACENT (weight rows
                            // interleaved at load, slime_b200/weights.py), rotated by the row's position
};

struct GemmParams {
  int M, N, K;           // C[M,N] = A[M,K] * W[N,K]^T
  const bf16* bias;      // [N] or nullptr
  const bf16* residual;  // [*, res_ld] or nullptr; added after the activation
  int res_ld;
  int res_period;      // >0: residual row = row % res_period (row-periodic table), else row
  const int* row_map;  // optional: output row = row_map[row]; negative = drop the row
  bf16* out;           // bf16 output (may alias residual) or nullptr
  float* out_f32;      // optional fp32 output (used for logits) or nullptr
  int out_ld;          // leading dimension of the output (elements)
  // GEMM_EPI_ROPE only (HF llama/modeling_llama.py:152-176 apply_rotary_pos_emb, rotate_half)
  const int* rope_pos = nullptr;       // [M] position id of every row
  const float2* rope_table = nullptr;  // [rope_max_pos][rope_half] (cos, sin) fp32
  int rope_half = 0;                   // head_dim / 2
  int rope_cols = 0;                   // columns [0, rope_cols) are rotated (q and k heads), the rest (v) is not
  int rope_max_pos = 0;
  // ---- norm folding (prefill): the RMSNorm in front of a projection is folded into it - gamma is multiplied into the
  // weight columns at load, the GEMM runs on the UN-normalised residual stream and the epilogue scales every output row
  // by its 1/rms (HF llama/modeling_llama.py:62-67).  The 1/rms values come from per-row partial sums of squares that
  // the PREVIOUS residual GEMM's epilogue writes (one partial per 64 output columns, of the bf16-rounded values) and a
  // tiny finishing kernel turns into rstd[row]: the separate RMSNorm pass over the stream (read + write) disappears.
  const float* row_scale = nullptr;  // [M] multiplied into the accumulators first (before RoPE / bias / activation)
  float* sumsq_out = nullptr;        // [M][sumsq_parts] partial sums of squares of this GEMM's bf16 output rows (EPI_NONE)
  int sumsq_parts = 0;               // = N / 64
  int store_hint = 0;  // L2 policy of the epilogue's output stores: 0 default, 1 evict_first (the output streams through L2
                       // once and would otherwise push the resident A panel out); filled in by slime_launch_gemm
  int epi_mode = 0;    // HBM access pattern of the epilogue (gemm_epilogue.cuh): 0 direct, 1 staged through shared
                       // memory (coalesced stores and residual loads); filled in by slime_launch_gemm
  // ---- decode-step problems (M <= 32 rows; gemm_skinny.cu) ----
  float* splitk_ws = nullptr;   // optional fp32 scratch for split-K partial sums [splits, M, N]
  size_t splitk_ws_floats = 0;  // its capacity; too small / absent: the problem runs unsplit
  int force_splits = 0;         // tests: > 0 forces the weight-streaming kernel with this many k-splits
  // In-kernel split-K finish (no finishing launch): [N / 8] ints, ZERO on entry and zero again on exit.  The warp that
  // delivers the last partial of an 8-column tile (atomic ticket) adds all partials in split order - so the result does
  // not depend on which warp came last - and runs the epilogue.
  int* tile_counters = nullptr;
  // RMSNorm of the A rows on their way into shared memory (weight-streaming path only; the decode step's
  // input_layernorm / post_attention_layernorm / final norm, HF modeling_llama.py:62-67):
  //   staged row = a_norm_w * elem(A[m] * rsqrt(mean(A[m]^2) + a_norm_eps)), the mean over all K columns.
  const bf16* a_norm_w = nullptr;
  float a_norm_eps = 0.f;
  // optional RMSNorm of the output rows (EPI_NONE, bf16 out): norm_out = norm_w * bf16(out * rsqrt(mean(out^2) + eps)).
  // Fused into the split-K finishing kernel on the weight-streaming path, a separate rmsnorm launch otherwise.
  const bf16* norm_w = nullptr;
  bf16* norm_out = nullptr;
  int norm_ld = 0;
  float norm_eps = 0.f;
  // optional KV-cache append of the decode step's packed QKV projection: output columns [kv_q_cols, kv_q_cols + kv_dim)
  // of row m also go to kv_k[(m * kv_cache_len + kv_lens[m]) * kv_dim + ...], the next kv_dim columns to kv_v (dropped
  // when the slot is outside the cache).  Done by the epilogue on the weight-streaming path, by a kv_append launch after
  // the GEMM otherwise.
  // optional: weights of a LATER projection to pull into L2 while this one's split-K finishing kernel runs
  const void* l2_prefetch = nullptr;
  size_t l2_prefetch_bytes = 0;
  bf16* kv_k = nullptr;
  bf16* kv_v = nullptr;
  const int* kv_lens = nullptr;
  int kv_cache_len = 0, kv_dim = 0, kv_q_cols = 0;
};

// 0 = 1-CTA kernel only, 1 = always the 2-CTA kernel, 2 = 2-CTA for problems that fill the GPU.
// (the SLIME_GEMM_2CTA environment variable overrides this at run time)
#ifndef SLIME_GEMM_2CTA_DEFAULT
#define SLIME_GEMM_2CTA_DEFAULT 2
#endif

// Returns 0 on success; negative SLIME_E* otherwise (message via slime_set_error).
int slime_launch_gemm(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi,
                      int num_sms, cudaStream_t stream);
// cta_group::2 variant (256 x 256 cluster tiles); arguments already validated by slime_launch_gemm.
int slime_launch_gemm_2cta(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi,
                           int num_sms, cudaStream_t stream);

// Weight-streaming kernel for M <= 32 rows (gemm_skinny.cu; the decode step).  slime_launch_gemm routes to it when
// slime_gemm_skinny_applies(); SLIME_GEMM_SKINNY=0 / slime_gemm_set_skinny_mode(0) keep such problems on the tcgen05 path.
bool slime_gemm_skinny_enabled();
bool slime_gemm_skinny_applies(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi, int num_sms);
int slime_launch_gemm_skinny(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi,
                             int num_sms, cudaStream_t stream);

// Epilogue access pattern used by slime_launch_gemm: 0 direct, 1 staged, 2 chosen by shape (SLIME_GEMM_EPI_MODE overrides).
#ifndef SLIME_GEMM_EPI_MODE_DEFAULT
#define SLIME_GEMM_EPI_MODE_DEFAULT 2
#endif

// m-tiles per rasterisation group for a problem with reduction length K (see gemm_sm100.cu)
int slime_gemm_group_m(int K, int tile_rows);

// Cached 2-D TMA descriptor over a row-major bf16 matrix [rows, cols] with leading dimension ld:
// box = {64 columns (one 128-byte swizzled row), box_rows rows}, SWIZZLE_128B, zero fill out of bounds.
int slime_get_tmap(const bf16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
