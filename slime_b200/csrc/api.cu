// C-ABI of libslime_b200 (include/slime_b200.h): context, weight registry, and the per-stage
// orchestration of the SliME prefill path.  Every stage is a fixed sequence of launches of the
// kernels in gemm_sm100.cu / attention_tc2.cu / elementwise.cu / router.cu / splice.cu on the
// caller's stream; scratch comes from the caller's workspace through a bump arena whose dry-run
// twin implements the *_workspace_bytes() queries, so the two can never disagree.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "attention.h"
#include "elementwise.h"
#include "errors.h"
#include "gemm.h"
#include "router.h"
#include "splice.h"

namespace {

struct Tensor {
  const bf16* p = nullptr;
  int64_t rows = 0, cols = 0;
};

struct Arena {
  char* base;
  size_t cap;
  size_t off = 0;
  bool dry;
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c), dry(b == nullptr) {}
  template <typename T>
  T* get(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    const size_t bytes = count * sizeof(T);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
  bool ok() const { return dry || off <= cap; }
};

#define ARENA_CHECK(a, what)                                                                  \
  do {                                                                                        \
    if (!(a).ok()) {                                                                          \
      slime_set_error("%s: workspace too small (%zu bytes needed, %zu given)", what, (a).off, \
                      (a).cap);                                                               \
      return SLIME_EWORKSPACE;                                                                \
    }                                                                                         \
  } while (0)

struct VitLayer {
  const bf16 *ln1_w, *ln1_b, *qkv_w, *qkv_b, *o_w, *o_b, *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};
struct Resampler {
  int nq;
  const bf16 *query, *pos_q, *pos_k, *ln_q_w, *ln_q_b, *ln_kv_w, *ln_kv_b, *ln_post_w, *ln_post_b,
      *in_w, *in_b, *out_w, *out_b;
  bf16 *derived_q, *derived_kvbias;
};
struct LlmLayer {
  const bf16 *in_norm_w, *qkv_w, *o_w, *post_norm_w, *gate_up_w, *down_w;
};
// TextGuidedRouterAttention (mm_resampler_type == "qformer"): cross_attn in/out projections, three LayerNorms, prob_proj
struct QformerRouter {
  const bf16 *ln_q_w, *ln_q_b, *ln_kv_w, *ln_kv_b, *ln_post_w, *ln_post_b, *in_w, *in_b, *out_w, *out_b, *fc1_w, *fc1_b,
      *fc2_w, *fc2_b;
};

}  // namespace

struct slime_ctx {
  slime_model_desc d;
  int device = 0;
  int num_sms = 148;
  std::unordered_map<std::string, Tensor> w;
  float* rope_table = nullptr;  // [max_pos, head_dim/2] (cos, sin) fp32, owned
  int* err_flag = nullptr;      // device int, owned
  bool finalized = false;
  bf16* kv_cache = nullptr;  // caller-owned [layers][2][kv_cache_batch][kv_cache_len][kv_heads*head_dim], or nullptr
  int kv_cache_batch = 0, kv_cache_len = 0;
  bool has_vit = false, has_rs[2] = {false, false}, has_proj = false, has_llm = false, has_router = false;
  QformerRouter qf = {};
  std::mutex mu;
  // resolved at finalize
  int vit_tokens = 0, vit_patches = 0, vit_kpad = 0;
  const bf16 *vit_patch_w = nullptr, *vit_cls = nullptr, *vit_pos = nullptr, *vit_pre_w = nullptr,
             *vit_pre_b = nullptr;
  std::vector<VitLayer> vit;
  Resampler rs[2];
  const bf16 *proj_fc1_w = nullptr, *proj_fc1_b = nullptr, *proj_fc2_w = nullptr, *proj_fc2_b = nullptr,
             *proj_w_gate = nullptr;
  const bf16 *llm_embed = nullptr, *llm_norm_w = nullptr, *llm_lm_head = nullptr;
  std::vector<LlmLayer> llm;
};

namespace {

#ifndef SLIME_DECODE_PREFETCH_DEFAULT
#define SLIME_DECODE_PREFETCH_DEFAULT 0  // measured: every mask is slower than none (profiles/r01_decode_bench.txt)
#endif
constexpr int VIT_CHUNK_CROPS = 128;  // crops per pass through the ViT (bounds the workspace to ~1.6 GB)

int find_weight(slime_ctx* c, const std::string& name, int64_t rows, int64_t cols, const bf16** out) {
  auto it = c->w.find(name);
  if (it == c->w.end()) {
    slime_set_error("weight '%s' was never registered", name.c_str());
    return SLIME_ESTATE;
  }
  if ((rows >= 0 && it->second.rows != rows) || (cols >= 0 && it->second.cols != cols)) {
    slime_set_error("weight '%s' has shape [%lld,%lld], expected [%lld,%lld]", name.c_str(),
                    static_cast<long long>(it->second.rows), static_cast<long long>(it->second.cols),
                    static_cast<long long>(rows), static_cast<long long>(cols));
    return SLIME_EINVAL;
  }
  if ((reinterpret_cast<uintptr_t>(it->second.p) & 15) != 0) {
    slime_set_error("weight '%s' is not 16-byte aligned", name.c_str());
    return SLIME_EINVAL;
  }
  *out = it->second.p;
  return SLIME_OK;
}

// Optional arguments of the decode-step GEMMs (M <= 32 rows, gemm_skinny.cu): split-K scratch and the RMSNorm that
// follows the projection (fused into the split-K finishing kernel when possible).
struct GemmExtra {
  float* splitk_ws = nullptr;
  size_t splitk_ws_floats = 0;
  const void* l2_prefetch = nullptr;  // weights of a later projection, pulled into L2 by this one's finishing kernel
  size_t l2_prefetch_bytes = 0;
  bf16* kv_k = nullptr;  // qkv_rope only: append this step's K / V rows to the cache (slot kv_lens[row] of sequence row)
  bf16* kv_v = nullptr;
  const int* kv_lens = nullptr;
  int kv_cache_len = 0;
  const bf16* norm_w = nullptr;
  bf16* norm_out = nullptr;
  int norm_ld = 0;
  float norm_eps = 0.f;
  // prefill norm folding (gemm.h): per-row 1/rms applied by the epilogue; partial sums of squares of the output rows
  const float* row_scale = nullptr;
  float* sumsq_out = nullptr;
  // fused decode chain (gemm.h): in-kernel split-K finish, RMSNorm of the A rows while they are staged
  int* tile_counters = nullptr;
  const bf16* a_norm_w = nullptr;
  float a_norm_eps = 0.f;
};

int gemm(slime_ctx* c, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K,
         const bf16* bias, const bf16* residual, int res_ld, int res_period, const int* row_map, int epi,
         bf16* out, float* out_f32, int out_ld, cudaStream_t s, const GemmExtra* ex = nullptr) {
  GemmParams p;
  if (ex != nullptr) {
    p.splitk_ws = ex->splitk_ws;
    p.splitk_ws_floats = ex->splitk_ws_floats;
    p.l2_prefetch = ex->l2_prefetch;
    p.l2_prefetch_bytes = ex->l2_prefetch_bytes;
    p.norm_w = ex->norm_w;
    p.norm_out = ex->norm_out;
    p.norm_ld = ex->norm_ld;
    p.norm_eps = ex->norm_eps;
    p.row_scale = ex->row_scale;
    p.sumsq_out = ex->sumsq_out;
    p.sumsq_parts = ex->sumsq_out != nullptr ? N / 64 : 0;
    p.tile_counters = ex->tile_counters;
    p.a_norm_w = ex->a_norm_w;
    p.a_norm_eps = ex->a_norm_eps;
  }
  p.M = M; p.N = N; p.K = K;
  p.bias = bias;
  p.residual = residual;
  p.res_ld = res_ld;
  p.res_period = res_period;
  p.row_map = row_map;
  p.out = out;
  p.out_f32 = out_f32;
  p.out_ld = out_ld;
  return slime_launch_gemm(A, lda, W, ldw, p, epi, c->num_sms, s);
}

// Packed Llama QKV projection + rotary embedding of the q / k heads (HF llama/modeling_llama.py:152-176, 262-288).
// With SLIME_FLAG_ROPE_INTERLEAVED the q / k weight rows of every head were interleaved at load so a feature pair
// (i, i + hd/2) sits in adjacent accumulator columns and the rotation happens in the GEMM epilogue on the fp32
// accumulators (q and k stay in that permuted feature order: q.k is invariant under a common permutation, and
// the KV cache / decode step use the same order).  Without the flag: plain GEMM, then rope_kernel in place.
int qkv_rope(slime_ctx* c, const bf16* x, const bf16* qkv_w, int rows, const int* pos, bf16* qkv, cudaStream_t s,
             const GemmExtra* ex = nullptr) {
  const slime_model_desc& d = c->d;
  const int H = d.hidden, hd = d.head_dim, QKV = (d.heads + 2 * d.kv_heads) * hd;
  if ((d.flags & SLIME_FLAG_ROPE_INTERLEAVED) == 0) {
    SLIME_PROPAGATE(gemm(c, x, H, qkv_w, H, rows, QKV, H, nullptr, nullptr, 0, 0, nullptr, GEMM_EPI_NONE, qkv, nullptr,
                         QKV, s, ex));
    SLIME_PROPAGATE(slime_launch_rope(qkv, QKV, rows, d.heads, d.kv_heads, hd, pos, c->rope_table, d.max_pos, s));
    if (ex != nullptr && ex->kv_k != nullptr)
      return slime_launch_kv_append(qkv + d.heads * hd, qkv + (d.heads + d.kv_heads) * hd, QKV, ex->kv_k, ex->kv_v,
                                    d.kv_heads * hd, ex->kv_lens, rows, ex->kv_cache_len, s);
    return SLIME_OK;
  }
  GemmParams p;
  if (ex != nullptr) {
    p.splitk_ws = ex->splitk_ws;
    p.splitk_ws_floats = ex->splitk_ws_floats;
    p.l2_prefetch = ex->l2_prefetch;
    p.l2_prefetch_bytes = ex->l2_prefetch_bytes;
    p.kv_k = ex->kv_k;
    p.kv_v = ex->kv_v;
    p.kv_lens = ex->kv_lens;
    p.kv_cache_len = ex->kv_cache_len;
    p.kv_dim = d.kv_heads * hd;
    p.kv_q_cols = d.heads * hd;
    p.row_scale = ex->row_scale;
    p.tile_counters = ex->tile_counters;
    p.a_norm_w = ex->a_norm_w;
    p.a_norm_eps = ex->a_norm_eps;
  }
  p.M = rows; p.N = QKV; p.K = H;
  p.bias = nullptr;
  p.residual = nullptr;
  p.res_ld = 0;
  p.res_period = 0;
  p.row_map = nullptr;
  p.out = qkv;
  p.out_f32 = nullptr;
  p.out_ld = QKV;
  p.rope_pos = pos;
  p.rope_table = reinterpret_cast<const float2*>(c->rope_table);
  p.rope_half = hd / 2;
  p.rope_cols = (d.heads + d.kv_heads) * hd;
  p.rope_max_pos = d.max_pos;
  return slime_launch_gemm(x, H, qkv_w, H, p, GEMM_EPI_ROPE, c->num_sms, s);
}

// ------------------------------------------------------------------------------------------
// stage bodies (dry == true: only account for workspace)
// ------------------------------------------------------------------------------------------
int vit_body(slime_ctx* c, Arena& a, const bf16* pixels, int n_crops, bf16* feats, cudaStream_t s, int per_image = 0) {
  const slime_model_desc& d = c->d;
  const int D = d.vit_hidden, I = d.vit_mlp, P = c->vit_patches, TK = c->vit_tokens, KP = c->vit_kpad;
  const int chunk = n_crops < VIT_CHUNK_CROPS ? n_crops : VIT_CHUNK_CROPS;
  const size_t rows_max = static_cast<size_t>(chunk) * TK;
  bf16* patches = a.get<bf16>(static_cast<size_t>(chunk) * P * KP);
  bf16* patch_out = a.get<bf16>(static_cast<size_t>(chunk) * P * D);
  bf16* h = a.get<bf16>(rows_max * D);
  bf16* t = a.get<bf16>(rows_max * D);
  bf16* qkv = a.get<bf16>(rows_max * 3 * D);
  bf16* att = a.get<bf16>(rows_max * D);
  bf16* u = a.get<bf16>(rows_max * I);
  ARENA_CHECK(a, "vision_tower");
  if (a.dry) return SLIME_OK;

  const int hd = D / d.vit_heads;
  const size_t px_per_crop = static_cast<size_t>(3) * d.vit_image * d.vit_image;
  for (int c0 = 0; c0 < n_crops; c0 += chunk) {
    const int nc = (n_crops - c0) < chunk ? (n_crops - c0) : chunk;
    const int rows = nc * TK;
    SLIME_PROPAGATE(slime_launch_im2col(pixels + c0 * px_per_crop, patches, nc, d.vit_image, d.vit_patch, KP, s));
    SLIME_PROPAGATE(gemm(c, patches, KP, c->vit_patch_w, KP, nc * P, D, KP, nullptr, nullptr, 0, 0, nullptr,
                         GEMM_EPI_NONE, patch_out, nullptr, D, s));
    SLIME_PROPAGATE(slime_launch_clip_embed_ln(patch_out, c->vit_cls, c->vit_pos, c->vit_pre_w, c->vit_pre_b, h,
                                               nc, TK, D, d.vit_ln_eps, s));
    for (int l = 0; l < d.vit_layers_used; ++l) {
      const VitLayer& L = c->vit[l];
      SLIME_PROPAGATE(slime_launch_layernorm(h, D, L.ln1_w, L.ln1_b, t, D, rows, D, d.vit_ln_eps, 0, 0, 0, s));
      SLIME_PROPAGATE(gemm(c, t, D, L.qkv_w, D, rows, 3 * D, D, L.qkv_b, nullptr, 0, 0, nullptr, GEMM_EPI_NONE,
                           qkv, nullptr, 3 * D, s));
      AttnParams ap;
      ap.q = qkv; ap.k = qkv + D; ap.v = qkv + 2 * D; ap.o = att;
      ap.q_ld = ap.k_ld = ap.v_ld = 3 * D; ap.o_ld = D;
      ap.cu_q = ap.cu_k = nullptr;
      ap.seqlen_q = ap.seqlen_k = TK;
      ap.q_batch_rows = ap.k_batch_rows = ap.o_batch_rows = TK;
      ap.batch = nc; ap.num_heads = d.vit_heads; ap.num_kv_heads = d.vit_heads; ap.head_dim = hd;
      ap.scale = 1.0f / sqrtf(static_cast<float>(hd)); ap.causal = 0;
      SLIME_PROPAGATE(slime_launch_attention(ap, s));
      SLIME_PROPAGATE(gemm(c, att, D, L.o_w, D, rows, D, D, L.o_b, h, D, 0, nullptr, GEMM_EPI_NONE, h, nullptr, D, s));
      SLIME_PROPAGATE(slime_launch_layernorm(h, D, L.ln2_w, L.ln2_b, t, D, rows, D, d.vit_ln_eps, 0, 0, 0, s));
      SLIME_PROPAGATE(gemm(c, t, D, L.fc1_w, D, rows, I, D, L.fc1_b, nullptr, 0, 0, nullptr, GEMM_EPI_QUICK_GELU,
                           u, nullptr, I, s));
      SLIME_PROPAGATE(gemm(c, u, I, L.fc2_w, I, rows, D, I, L.fc2_b, h, D, 0, nullptr, GEMM_EPI_NONE, h, nullptr, D, s));
    }
    // feature_select 'patch': drop the CLS row of every crop
    if (per_image > 0) {  // global crops of all images first, local crops behind them
      SLIME_PROPAGATE(slime_launch_vit_split_rows(h, feats, nc, D, P, TK, c0, per_image, n_crops / per_image, s));
    } else {
      SLIME_PROPAGATE(slime_launch_copy_rows(h, D, feats + static_cast<size_t>(c0) * P * D, D, nc * P, D, P, TK, 1, s));
    }
  }
  return SLIME_OK;
}

int resampler_body(slime_ctx* c, Arena& a, int which, const bf16* x, int n, bf16* out, cudaStream_t s) {
  const slime_model_desc& d = c->d;
  const Resampler& R = c->rs[which];
  const int D = d.vit_hidden, NK = c->vit_patches, nq = R.nq;
  const size_t kv_rows = static_cast<size_t>(n) * NK, q_rows = static_cast<size_t>(n) * nq;
  bf16* kvn = a.get<bf16>(kv_rows * D);
  bf16* kv = a.get<bf16>(kv_rows * 2 * D);
  bf16* att = a.get<bf16>(q_rows * D);
  bf16* o = a.get<bf16>(q_rows * D);
  ARENA_CHECK(a, "resampler");
  if (a.dry || n <= 0) return SLIME_OK;
  const int heads = D / 128;
  SLIME_PROPAGATE(slime_launch_layernorm(x, D, R.ln_kv_w, R.ln_kv_b, kvn, D, static_cast<int>(kv_rows), D,
                                         d.rs_ln_eps, 0, 0, 0, s));
  // [K | V] = ln_kv(x) [Wk;Wv]^T + [bk;bv] + [pos_k Wk^T | 0]   (the position term is row-periodic)
  SLIME_PROPAGATE(gemm(c, kvn, D, R.in_w + static_cast<size_t>(D) * D, D, static_cast<int>(kv_rows), 2 * D, D,
                       R.in_b + D, R.derived_kvbias, 2 * D, NK, nullptr, GEMM_EPI_NONE, kv, nullptr, 2 * D, s));
  AttnParams ap;
  ap.q = R.derived_q; ap.k = kv; ap.v = kv + D; ap.o = att;
  ap.q_ld = D; ap.k_ld = ap.v_ld = 2 * D; ap.o_ld = D;
  ap.cu_q = ap.cu_k = nullptr;
  ap.seqlen_q = nq; ap.seqlen_k = NK;
  ap.q_batch_rows = 0; ap.k_batch_rows = NK; ap.o_batch_rows = nq;
  ap.batch = n; ap.num_heads = heads; ap.num_kv_heads = heads; ap.head_dim = 128;
  ap.scale = 1.0f / sqrtf(128.0f); ap.causal = 0;
  SLIME_PROPAGATE(slime_launch_attention(ap, s));
  SLIME_PROPAGATE(gemm(c, att, D, R.out_w, D, static_cast<int>(q_rows), D, D, R.out_b, nullptr, 0, 0, nullptr,
                       GEMM_EPI_NONE, o, nullptr, D, s));
  SLIME_PROPAGATE(slime_launch_layernorm(o, D, R.ln_post_w, R.ln_post_b, out, D, static_cast<int>(q_rows), D,
                                         d.rs_ln_eps, 0, 0, 0, s));
  return SLIME_OK;
}

int projector_body(slime_ctx* c, Arena& a, const bf16* x, int rows, const int* row_map, bf16* out,
                   cudaStream_t s) {
  const int D = c->d.vit_hidden, H = c->d.hidden;
  bf16* u = a.get<bf16>(static_cast<size_t>(rows) * H);
  ARENA_CHECK(a, "projector");
  if (a.dry || rows <= 0) return SLIME_OK;
  SLIME_PROPAGATE(gemm(c, x, D, c->proj_fc1_w, D, rows, H, D, c->proj_fc1_b, nullptr, 0, 0, nullptr,
                       GEMM_EPI_GELU_ERF, u, nullptr, H, s));
  SLIME_PROPAGATE(gemm(c, u, H, c->proj_fc2_w, H, rows, H, H, c->proj_fc2_b, nullptr, 0, 0, row_map,
                       GEMM_EPI_NONE, out, nullptr, H, s));
  return SLIME_OK;
}

int gated_body(slime_ctx* c, Arena& a, const bf16* x, int n, bf16* out, cudaStream_t s) {
  const slime_model_desc& d = c->d;
  const int D = d.vit_hidden, H = d.hidden, NQ = c->rs[1].nq;
  const int rows = n * NQ;
  if (d.mm_learnable_gated == 0) return projector_body(c, a, x, rows, nullptr, out, s);
  bf16* r = a.get<bf16>(static_cast<size_t>(rows) * D);
  if (d.mm_learnable_gated == 1) {
    SLIME_PROPAGATE(resampler_body(c, a, 1, x, n, r, s));
    return projector_body(c, a, r, rows, nullptr, out, s);
  }
  bf16* e0 = a.get<bf16>(static_cast<size_t>(rows) * H);
  bf16* e1 = a.get<bf16>(static_cast<size_t>(rows) * H);
  SLIME_PROPAGATE(resampler_body(c, a, 1, x, n, r, s));
  SLIME_PROPAGATE(projector_body(c, a, x, rows, nullptr, e0, s));
  SLIME_PROPAGATE(projector_body(c, a, r, rows, nullptr, e1, s));
  ARENA_CHECK(a, "gated_projector");
  if (a.dry || n <= 0) return SLIME_OK;
  SLIME_PROPAGATE(slime_launch_gate_mix(x, c->proj_w_gate, e0, e1, out, rows, D, H, s));
  return SLIME_OK;
}

// ids != nullptr: prompt given as token ids (rows gathered from the embedding table, placeholders skipped);
// ids == nullptr: prompt given as a dense [B, T, H] embedding tensor `text` (the reference's module-level API).
// 'qformer' router: probs1 = softmax(prob_proj(ln_post(cross_attn(ln_q(local), ln_kv(text), ln_kv(text)))) / temp)
// (reference multimodal_resampler/builder.py:148-160), then the sampler's own softmax + top-p (:258-273).
int router_qformer_body(slime_ctx* c, Arena& a, const bf16* local, int n_per, const int* n_valid, const long long* ids,
                        const bf16* text, const unsigned char* mask, int B, int T, float* probs_out, int* sel_idx,
                        int* sel_count, cudaStream_t s) {
  const int H = c->d.hidden, Dh = H / 4, heads = H / 128;
  const size_t rows = static_cast<size_t>(B) * (n_per > 0 ? n_per : 1), trows = static_cast<size_t>(B) * T;
  int* dst_row = a.get<int>(trows);
  int* cu_k = a.get<int>(B + 1);
  bf16* tpack = a.get<bf16>(trows * H);       // kept prompt rows, packed
  bf16* tn = a.get<bf16>(trows * H);          // ln_kv(text)
  bf16* kv = a.get<bf16>(trows * 2 * H);      // K | V
  bf16* xn = a.get<bf16>(rows * H);           // ln_q(local), later the attention output
  bf16* q = a.get<bf16>(rows * H);            // Q, later out_proj + ln_post
  bf16* hid = a.get<bf16>(rows * Dh);         // first prob_proj layer (pre-ReLU)
  float* logit = a.get<float>(rows);
  ARENA_CHECK(a, "router (qformer)");
  if (a.dry || B <= 0) return SLIME_OK;
  if (n_per <= 0) {
    SLIME_CHECK_CUDA(cudaMemsetAsync(sel_count, 0, sizeof(int) * B, s));
    return SLIME_OK;
  }
  SLIME_REQUIRE(c->has_router, "router: mm_resampler_type='qformer' but the weight group 'router' is not registered");
  const QformerRouter& R = c->qf;
  const int n_rows = B * n_per, t_rows = B * T;
  SLIME_CHECK_CUDA(cudaMemsetAsync(tpack, 0, trows * H * sizeof(bf16), s));  // rows past the packed end stay finite
  SLIME_PROPAGATE(slime_launch_qf_pack_text(ids, mask, ids != nullptr ? c->llm_embed : text, B, T, H, c->d.image_token,
                                            c->d.vocab, dst_row, cu_k, tpack, s));
  constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default (builder.py:108 norm_layer=nn.LayerNorm)
  SLIME_PROPAGATE(slime_launch_layernorm(tpack, H, R.ln_kv_w, R.ln_kv_b, tn, H, t_rows, H, LN_EPS, 0, 0, 0, s));
  SLIME_PROPAGATE(slime_launch_layernorm(local, H, R.ln_q_w, R.ln_q_b, xn, H, n_rows, H, LN_EPS, 0, 0, 0, s));
  // packed in_proj: rows [0, H) = q, [H, 2H) = k, [2H, 3H) = v (torch.nn.MultiheadAttention)
  SLIME_PROPAGATE(gemm(c, xn, H, R.in_w, H, n_rows, H, H, R.in_b, nullptr, 0, 0, nullptr, GEMM_EPI_NONE, q, nullptr, H, s));
  SLIME_PROPAGATE(gemm(c, tn, H, R.in_w + static_cast<size_t>(H) * H, H, t_rows, 2 * H, H, R.in_b + H, nullptr, 0, 0, nullptr,
                       GEMM_EPI_NONE, kv, nullptr, 2 * H, s));
  AttnParams ap;
  ap.q = q; ap.k = kv; ap.v = kv + H; ap.o = xn;
  ap.q_ld = H; ap.k_ld = ap.v_ld = 2 * H; ap.o_ld = H;
  ap.cu_q = nullptr; ap.cu_k = cu_k;           // every sample's n_per queries against its own kept prompt tokens
  ap.seqlen_q = n_per; ap.seqlen_k = T;
  ap.q_batch_rows = n_per; ap.k_batch_rows = 0; ap.o_batch_rows = n_per;
  ap.batch = B; ap.num_heads = heads; ap.num_kv_heads = heads; ap.head_dim = 128;
  ap.scale = 1.0f / sqrtf(128.0f); ap.causal = 0;
  ap.total_q_rows = n_rows; ap.total_k_rows = t_rows;
  SLIME_PROPAGATE(slime_launch_attention(ap, s));
  SLIME_PROPAGATE(gemm(c, xn, H, R.out_w, H, n_rows, H, H, R.out_b, nullptr, 0, 0, nullptr, GEMM_EPI_NONE, q, nullptr, H, s));
  SLIME_PROPAGATE(slime_launch_layernorm(q, H, R.ln_post_w, R.ln_post_b, xn, H, n_rows, H, LN_EPS, 0, 0, 0, s));
  SLIME_PROPAGATE(gemm(c, xn, H, R.fc1_w, H, n_rows, Dh, H, R.fc1_b, nullptr, 0, 0, nullptr, GEMM_EPI_NONE, hid, nullptr, Dh, s));
  SLIME_PROPAGATE(slime_launch_qf_logits(hid, R.fc2_w, R.fc2_b, logit, n_rows, Dh, B, n_per, n_valid, c->d.temp, s));
  // the sampler soft-maxes the router's (already soft-maxed) output once more before the top-p rule
  SLIME_PROPAGATE(slime_launch_router_select(logit, B, n_per, n_valid, c->d.temp, c->d.top_p, 0, probs_out, sel_idx,
                                             sel_count, s));
  return SLIME_OK;
}

int router_body(slime_ctx* c, Arena& a, const bf16* local, int n_per, const int* n_valid, const long long* ids,
                const bf16* text, const unsigned char* mask, int B, int T, float* probs_out, int* sel_idx,
                int* sel_count, cudaStream_t s) {
  if ((c->d.flags & SLIME_FLAG_ROUTER_QFORMER) != 0)
    return router_qformer_body(c, a, local, n_per, n_valid, ids, text, mask, B, T, probs_out, sel_idx, sel_count, s);
  const int H = c->d.hidden;
  float* inv_norm = a.get<float>(static_cast<size_t>(B) * T);
  float* tvec = a.get<float>(static_cast<size_t>(B) * H);
  float* score = a.get<float>(static_cast<size_t>(B) * (n_per > 0 ? n_per : 1));
  ARENA_CHECK(a, "router");
  if (a.dry || B <= 0) return SLIME_OK;
  if (n_per <= 0) {
    SLIME_CHECK_CUDA(cudaMemsetAsync(sel_count, 0, sizeof(int) * B, s));
    return SLIME_OK;
  }
  SLIME_PROPAGATE(slime_launch_text_dir(ids, mask, ids != nullptr ? c->llm_embed : text, inv_norm, tvec, B, T, H,
                                        c->d.image_token, c->d.vocab, s));
  SLIME_PROPAGATE(slime_launch_router_score(local, tvec, score, B * n_per, n_per, H, s));
  SLIME_PROPAGATE(slime_launch_router_select(score, B, n_per, n_valid, c->d.temp, c->d.top_p, 0, probs_out,
                                             sel_idx, sel_count, s));
  return SLIME_OK;
}

int decoder_body(slime_ctx* c, Arena& a, const bf16* embeds, const int* cu, const int* pos_ids, int B,
                 int total, int max_seqlen, float* logits_last, bf16* logits_all, bf16* hidden_out,
                 cudaStream_t s) {
  const slime_model_desc& d = c->d;
  const int H = d.hidden, I = d.mlp, hd = d.head_dim;
  const int QKV = (d.heads + 2 * d.kv_heads) * hd, QD = d.heads * hd, KD = d.kv_heads * hd;
  bf16* h = a.get<bf16>(static_cast<size_t>(total) * H);
  bf16* t = a.get<bf16>(static_cast<size_t>(total) * H);
  bf16* qkv = a.get<bf16>(static_cast<size_t>(total) * QKV);
  bf16* att = a.get<bf16>(static_cast<size_t>(total) * QD);
  bf16* act = a.get<bf16>(static_cast<size_t>(total) * I);
  int* last_rows = a.get<int>(B);
  bf16* last_h = a.get<bf16>(static_cast<size_t>(B) * H);
  int* cache_rows = a.get<int>(total);  // packed row -> slot of the per-sequence KV cache (when one is attached)
  float* rstd = a.get<float>(total);    // norm folding: 1/rms of every row of the residual stream
  float* sumsq = a.get<float>(static_cast<size_t>(total) * (H / 64 + 1));  // ... and the partials a residual GEMM writes
  ARENA_CHECK(a, "decoder");
  if (a.dry || total <= 0) return SLIME_OK;

  // The residual stream starts in the caller's `embeds` (read only) and moves to h with the first o-projection
  // (residual = embeds, out = h): no copy of the spliced rows.
  const bf16* hin = embeds;
  // SLIME_FLAG_NORM_FOLDED: the two RMSNorms of a layer live inside the projections that follow them (gamma in the weight
  // columns, 1/rms in the epilogue; gemm.h) - the GEMMs read the residual stream h directly, `t` is not used
  const bool folded = (d.flags & SLIME_FLAG_NORM_FOLDED) != 0 && H % 64 == 0;
  const int parts = H / 64;
  if (folded) SLIME_PROPAGATE(slime_launch_row_rstd(hin, H, rstd, total, H, d.rms_eps, s));
  for (int l = 0; l < d.layers; ++l) {
    const LlmLayer& L = c->llm[l];
    GemmExtra exs, exq;  // exs: residual GEMM that also writes the partials; exq: projection scaled by 1/rms
    exs.sumsq_out = folded ? sumsq : nullptr;
    exq.row_scale = folded ? rstd : nullptr;
    if (!folded) SLIME_PROPAGATE(slime_launch_rmsnorm(hin, H, L.in_norm_w, t, H, total, H, d.rms_eps, nullptr, s));
    SLIME_PROPAGATE(qkv_rope(c, folded ? hin : t, L.qkv_w, total, pos_ids, qkv, s, folded ? &exq : nullptr));
    if (c->kv_cache != nullptr) {
      // keep K (post-RoPE) and V of every real token for the decode steps that follow the prefill
      if (l == 0)
        SLIME_PROPAGATE(slime_launch_cache_rows(cu, pos_ids, B, total, c->kv_cache_batch, c->kv_cache_len, cache_rows, s));
      const size_t plane = static_cast<size_t>(c->kv_cache_batch) * c->kv_cache_len * KD;
      bf16* kc = c->kv_cache + (static_cast<size_t>(l) * 2 + 0) * plane;
      bf16* vc = c->kv_cache + (static_cast<size_t>(l) * 2 + 1) * plane;
      SLIME_PROPAGATE(slime_launch_scatter_rows(qkv + QD, QKV, kc, KD, total, KD, cache_rows, s));
      SLIME_PROPAGATE(slime_launch_scatter_rows(qkv + QD + KD, QKV, vc, KD, total, KD, cache_rows, s));
    }
    AttnParams ap;
    ap.q = qkv; ap.k = qkv + QD; ap.v = qkv + QD + KD; ap.o = att;
    ap.q_ld = ap.k_ld = ap.v_ld = QKV; ap.o_ld = QD;
    ap.cu_q = cu; ap.cu_k = cu;
    ap.seqlen_q = ap.seqlen_k = max_seqlen;
    ap.q_batch_rows = ap.k_batch_rows = ap.o_batch_rows = 0;
    ap.batch = B; ap.num_heads = d.heads; ap.num_kv_heads = d.kv_heads; ap.head_dim = hd;
    ap.scale = 1.0f / sqrtf(static_cast<float>(hd)); ap.causal = 1;
    ap.total_q_rows = ap.total_k_rows = total;
    SLIME_PROPAGATE(slime_launch_attention(ap, s));
    SLIME_PROPAGATE(gemm(c, att, QD, L.o_w, QD, total, H, QD, nullptr, hin, H, 0, nullptr, GEMM_EPI_NONE, h, nullptr, H, s,
                         folded ? &exs : nullptr));
    hin = h;
    if (folded) {
      SLIME_PROPAGATE(slime_launch_sumsq_to_rstd(sumsq, parts, rstd, total, H, d.rms_eps, s));
    } else {
      SLIME_PROPAGATE(slime_launch_rmsnorm(h, H, L.post_norm_w, t, H, total, H, d.rms_eps, nullptr, s));
    }
    SLIME_PROPAGATE(gemm(c, folded ? h : t, H, L.gate_up_w, H, total, 2 * I, H, nullptr, nullptr, 0, 0, nullptr,
                         GEMM_EPI_SWIGLU, act, nullptr, I, s, folded ? &exq : nullptr));
    SLIME_PROPAGATE(gemm(c, act, I, L.down_w, I, total, H, I, nullptr, h, H, 0, nullptr, GEMM_EPI_NONE, h, nullptr, H, s,
                         folded && l + 1 < d.layers ? &exs : nullptr));
    if (folded && l + 1 < d.layers) SLIME_PROPAGATE(slime_launch_sumsq_to_rstd(sumsq, parts, rstd, total, H, d.rms_eps, s));
  }
  if (logits_last != nullptr) {
    SLIME_PROPAGATE(slime_launch_last_rows(cu, B, last_rows, s));
    SLIME_PROPAGATE(slime_launch_rmsnorm(hin, H, c->llm_norm_w, last_h, H, B, H, d.rms_eps, last_rows, s));
    SLIME_PROPAGATE(gemm(c, last_h, H, c->llm_lm_head, H, B, d.vocab, H, nullptr, nullptr, 0, 0, nullptr,
                         GEMM_EPI_NONE, nullptr, logits_last, d.vocab, s));
  }
  if (logits_all != nullptr || hidden_out != nullptr) {
    bf16* hn = hidden_out != nullptr ? hidden_out : t;
    SLIME_PROPAGATE(slime_launch_rmsnorm(hin, H, c->llm_norm_w, hn, H, total, H, d.rms_eps, nullptr, s));
    if (logits_all != nullptr) {
      SLIME_PROPAGATE(gemm(c, hn, H, c->llm_lm_head, H, total, d.vocab, H, nullptr, nullptr, 0, 0, nullptr,
                           GEMM_EPI_NONE, logits_all, nullptr, d.vocab, s));
    }
  }
  return SLIME_OK;
}

int g_decode_pf = -1;  // -1 unset (SLIME_DECODE_PREFETCH or the default), else the mask
int decode_prefetch_mask() {
  if (g_decode_pf < 0) {
    const char* e = getenv("SLIME_DECODE_PREFETCH");
    g_decode_pf = (e != nullptr && e[0] >= '0' && e[0] <= '9') ? atoi(e) : SLIME_DECODE_PREFETCH_DEFAULT;
  }
  return g_decode_pf;
}

// One decode step for B sequences: x [B, H] = embeddings of the tokens to append; lens[b] = tokens already cached.
// HBM-bound (every weight byte is read once per step): the projections run on the weight-streaming kernel of
// gemm_skinny.cu for B <= 32, attention on the split-KV kernel of decode_attn.cu; 7-8 launches per layer -
//   qkv (+RoPE, K/V append) [+ split-K finish] | attention [+ split merge] | o-proj + finish(residual, RMSNorm) |
//   gate/up (SwiGLU) | down-proj + finish(residual, next layer's RMSNorm).
// Fused chain (opt-in: SLIME_DECODE_FUSED=7 / slime_set_decode_fused(7)): 5 launches per layer.
// The finishing kernels disappear - split-K partials are summed by the warp that delivers a tile's last partial and the
// kv splits of the attention by the last CTA of a (sequence, kv head), both by atomic ticket and in a fixed order - and
// the RMSNorms move into the activation staging of the projection that consumes them (every CTA normalises the <= 32
// rows it stages; the residual stream h is the only activation buffer between layers):
//   qkv (RMSNorm in, RoPE, K/V append) | attention | o-proj (+residual) | gate/up (RMSNorm in, SwiGLU) | down (+residual).
// Bit mask (A/B): 1 = split-K sums finished inside the projections, 2 = kv splits merged inside the attention kernel,
// 4 = RMSNorm in the consumer's staging (else: fused into the o-/down-projection's finishing kernel).  7 = all.
#ifndef SLIME_DECODE_FUSED_DEFAULT
#define SLIME_DECODE_FUSED_DEFAULT 0  // measured: every fusion is slower than its PDL finishing launch (profiles/r02_decode_experiments.txt)
#endif
int g_decode_fused = -1;
int decode_fused_mask() {
  if (g_decode_fused < 0) {
    const char* e = getenv("SLIME_DECODE_FUSED");
    g_decode_fused = (e != nullptr && e[0] >= '0' && e[0] <= '7') ? e[0] - '0' : SLIME_DECODE_FUSED_DEFAULT;
  }
  return g_decode_fused;
}

int decode_body(slime_ctx* c, Arena& a, const bf16* x_in, const int* lens, int B, float* logits, cudaStream_t s) {
  const slime_model_desc& d = c->d;
  const int H = d.hidden, I = d.mlp, hd = d.head_dim;
  const int QKV = (d.heads + 2 * d.kv_heads) * hd, QD = d.heads * hd, KD = d.kv_heads * hd;
  bf16* h = a.get<bf16>(static_cast<size_t>(B) * H);
  bf16* t = a.get<bf16>(static_cast<size_t>(B) * H);
  bf16* qkv = a.get<bf16>(static_cast<size_t>(B) * QKV);
  bf16* att = a.get<bf16>(static_cast<size_t>(B) * QD);
  bf16* act = a.get<bf16>(static_cast<size_t>(B) * I);
  // split-K scratch of the projections: up to 8 splits of the [B, <= max(QKV, H)] outputs, 4 of the wide ones (more
  // than 16 sequences stage at most 2048 k per split: hidden size 5120 needs 4 splits of gate/up and the LM head;
  // tests/test_skinny_plan_cpu.py mirrors the plan and checks this bound for every model size and batch)
  const size_t wide = static_cast<size_t>(2 * I > d.vocab ? 2 * I : d.vocab);
  const size_t narrow = static_cast<size_t>(QKV > H ? QKV : H);
  const size_t sk_floats = static_cast<size_t>(B) * (8 * narrow > 4 * wide ? 8 * narrow : 4 * wide);
  float* skws = a.get<float>(sk_floats);
  // kv splits of the attention: sized for the attached cache (the dry run has none: assume the worst case)
  const int cache_len = c->kv_cache != nullptr ? c->kv_cache_len : d.max_pos;
  const int asplits = slime_decode_attention_splits(B, d.heads, d.kv_heads, hd, cache_len, c->num_sms);
  const int asplits_max = slime_decode_attention_splits(B, d.heads, d.kv_heads, hd, d.max_pos, c->num_sms);
  float* aws = a.get<float>(slime_decode_attention_ws_floats(B, d.heads, asplits_max > asplits ? asplits_max : asplits));
  // tickets of the fused chain: one int per 8-column tile of the widest projection + one per (sequence, kv head)
  const size_t n_tile_cnt = wide / 8 + 1;
  const size_t n_cnt = n_tile_cnt + static_cast<size_t>(B > 0 ? B : 1) * d.kv_heads;
  int* counters = a.get<int>(n_cnt);
  ARENA_CHECK(a, "decode");
  if (a.dry || B <= 0) return SLIME_OK;
  const int fm = (B <= 32 && asplits >= 1 && H % 32 == 0 && I % 32 == 0 && slime_gemm_skinny_enabled()) ? decode_fused_mask() : 0;
  SLIME_CHECK_CUDA(cudaMemcpyAsync(h, x_in, static_cast<size_t>(B) * H * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
  if (fm & 3) SLIME_CHECK_CUDA(cudaMemsetAsync(counters, 0, n_cnt * sizeof(int), s));
  const size_t plane = static_cast<size_t>(c->kv_cache_batch) * c->kv_cache_len * KD;
  GemmExtra ex;
  ex.splitk_ws = skws;
  ex.splitk_ws_floats = sk_floats;
  const int pf = decode_prefetch_mask();
  int* const tile_cnt = (fm & 1) ? counters : nullptr;
  int* const attn_cnt = (fm & 2) ? counters + n_tile_cnt : nullptr;
  if (fm & 4) {  // the residual stream h is the only activation buffer between layers
    ex.tile_counters = tile_cnt;
    GemmExtra exa = ex;  // projections that read the residual stream through an RMSNorm
    exa.a_norm_eps = d.rms_eps;
    for (int l = 0; l < d.layers; ++l) {
      const LlmLayer& L = c->llm[l];
      bf16* kc = c->kv_cache + (static_cast<size_t>(l) * 2 + 0) * plane;
      bf16* vc = c->kv_cache + (static_cast<size_t>(l) * 2 + 1) * plane;
      GemmExtra exq = exa;
      exq.a_norm_w = L.in_norm_w;
      exq.kv_k = kc;
      exq.kv_v = vc;
      exq.kv_lens = lens;
      exq.kv_cache_len = c->kv_cache_len;
      SLIME_PROPAGATE(qkv_rope(c, h, L.qkv_w, B, lens, qkv, s, &exq));
      SLIME_PROPAGATE(slime_launch_decode_attention(qkv, QKV, kc, vc, c->kv_cache_len, lens, B, d.heads, d.kv_heads, hd,
                                                    1.0f / sqrtf(static_cast<float>(hd)), att, QD, asplits, aws, nullptr, 0, s,
                                                    attn_cnt));
      SLIME_PROPAGATE(gemm(c, att, QD, L.o_w, QD, B, H, QD, nullptr, h, H, 0, nullptr, GEMM_EPI_NONE, h, nullptr, H, s, &ex));
      exa.a_norm_w = L.post_norm_w;
      SLIME_PROPAGATE(gemm(c, h, H, L.gate_up_w, H, B, 2 * I, H, nullptr, nullptr, 0, 0, nullptr, GEMM_EPI_SWIGLU, act,
                           nullptr, I, s, &exa));
      SLIME_PROPAGATE(gemm(c, act, I, L.down_w, I, B, H, I, nullptr, h, H, 0, nullptr, GEMM_EPI_NONE, h, nullptr, H, s, &ex));
    }
    exa.a_norm_w = c->llm_norm_w;
    return gemm(c, h, H, c->llm_lm_head, H, B, d.vocab, H, nullptr, nullptr, 0, 0, nullptr, GEMM_EPI_NONE, nullptr, logits,
                d.vocab, s, &exa);
  }
  SLIME_PROPAGATE(slime_launch_rmsnorm(h, H, c->llm[0].in_norm_w, t, H, B, H, d.rms_eps, nullptr, s));
  for (int l = 0; l < d.layers; ++l) {
    const LlmLayer& L = c->llm[l];
    bf16* kc = c->kv_cache + (static_cast<size_t>(l) * 2 + 0) * plane;
    bf16* vc = c->kv_cache + (static_cast<size_t>(l) * 2 + 1) * plane;
    // L2 prefetch duties (bit mask, slime_set_decode_prefetch; OFF by default - on B200 every variant measured
    // slower than none, 4.26 -> 5.7..7.9 ms per step): the latency-bound kernels between the projections would keep
    // HBM busy with weights that are needed next -
    //   1: the QKV finishing kernel pulls the o-projection (used after the attention),
    //   2: the down-projection's finishing kernel pulls the next layer's QKV weights (or the head of lm_head),
    //   4: the attention kernel pulls the first PF_PART bytes of gate/up, 8: the o-projection's finishing kernel too.
    constexpr size_t PF_PART = 32u << 20;
    const size_t gu_bytes = static_cast<size_t>(2) * I * H * sizeof(bf16);
    GemmExtra exq = ex;  // QKV projection + RoPE; K / V of the new token go straight into the cache
    exq.kv_k = kc;
    exq.kv_v = vc;
    exq.kv_lens = lens;
    exq.kv_cache_len = c->kv_cache_len;
    exq.tile_counters = tile_cnt;
    if (pf & 1) {
      exq.l2_prefetch = L.o_w;
      exq.l2_prefetch_bytes = static_cast<size_t>(H) * QD * sizeof(bf16);
    }
    SLIME_PROPAGATE(qkv_rope(c, t, L.qkv_w, B, lens, qkv, s, &exq));
    SLIME_PROPAGATE(slime_launch_decode_attention(qkv, QKV, kc, vc, c->kv_cache_len, lens, B, d.heads, d.kv_heads, hd,
                                                  1.0f / sqrtf(static_cast<float>(hd)), att, QD, asplits, aws,
                                                  (pf & 4) ? L.gate_up_w : nullptr,
                                                  (pf & 4) ? (gu_bytes < PF_PART ? gu_bytes : PF_PART) : 0, s, attn_cnt));
    GemmExtra exn = ex;  // projection + residual, then the RMSNorm that feeds the next GEMM
    exn.norm_out = t;
    exn.norm_ld = H;
    exn.norm_eps = d.rms_eps;
    exn.norm_w = L.post_norm_w;
    if ((pf & 8) && gu_bytes > PF_PART) {
      exn.l2_prefetch = reinterpret_cast<const char*>(L.gate_up_w) + ((pf & 4) ? PF_PART : 0);
      exn.l2_prefetch_bytes = (gu_bytes - ((pf & 4) ? PF_PART : 0)) < PF_PART ? gu_bytes - ((pf & 4) ? PF_PART : 0) : PF_PART;
    }
    SLIME_PROPAGATE(gemm(c, att, QD, L.o_w, QD, B, H, QD, nullptr, h, H, 0, nullptr, GEMM_EPI_NONE, h, nullptr, H, s, &exn));
    SLIME_PROPAGATE(gemm(c, t, H, L.gate_up_w, H, B, 2 * I, H, nullptr, nullptr, 0, 0, nullptr, GEMM_EPI_SWIGLU, act,
                         nullptr, I, s, &ex));
    exn.norm_w = (l + 1 < d.layers) ? c->llm[l + 1].in_norm_w : c->llm_norm_w;
    exn.l2_prefetch = nullptr;
    exn.l2_prefetch_bytes = 0;
    if (pf & 2) {
      const size_t head_bytes = static_cast<size_t>(d.vocab) * H * sizeof(bf16);
      exn.l2_prefetch = (l + 1 < d.layers) ? c->llm[l + 1].qkv_w : c->llm_lm_head;
      exn.l2_prefetch_bytes = (l + 1 < d.layers) ? static_cast<size_t>(QKV) * H * sizeof(bf16)
                                                 : (head_bytes < 2 * PF_PART ? head_bytes : 2 * PF_PART);
    }
    SLIME_PROPAGATE(gemm(c, act, I, L.down_w, I, B, H, I, nullptr, h, H, 0, nullptr, GEMM_EPI_NONE, h, nullptr, H, s, &exn));
  }
  SLIME_PROPAGATE(gemm(c, t, H, c->llm_lm_head, H, B, d.vocab, H, nullptr, nullptr, 0, 0, nullptr, GEMM_EPI_NONE, nullptr,
                       logits, d.vocab, s, &ex));
  return SLIME_OK;
}

int finalize_body(slime_ctx* c, Arena& a, cudaStream_t s) {
  const int D = c->d.vit_hidden;
  const int NK = c->vit_patches;
  const int nq_max = c->d.rs_local_queries > c->d.rs_global_queries ? c->d.rs_local_queries : c->d.rs_global_queries;
  bf16* t1 = a.get<bf16>(static_cast<size_t>(nq_max) * D);
  bf16* t2 = a.get<bf16>(static_cast<size_t>(nq_max) * D);
  ARENA_CHECK(a, "finalize_weights");
  if (a.dry) return SLIME_OK;
  for (int which = 0; which < 2; ++which) {
    if (!c->has_rs[which]) continue;
    Resampler& R = c->rs[which];
    // Q = (ln_q(query) + pos_q) Wq^T + bq   -- input independent (sampler.py:159-161)
    SLIME_PROPAGATE(slime_launch_layernorm(R.query, D, R.ln_q_w, R.ln_q_b, t1, D, R.nq, D, c->d.rs_ln_eps, 0, 0, 0, s));
    SLIME_PROPAGATE(slime_launch_add_rows(t1, R.pos_q, t2, R.nq, D, 0, s));
    SLIME_PROPAGATE(gemm(c, t2, D, R.in_w, D, R.nq, D, D, R.in_b, nullptr, 0, 0, nullptr, GEMM_EPI_NONE,
                         R.derived_q, nullptr, D, s));
    // [pos_k Wk^T | 0]: the key-side position term of (ln_kv(x) + pos_k) Wk^T   (sampler.py:162)
    SLIME_CHECK_CUDA(cudaMemsetAsync(R.derived_kvbias, 0, static_cast<size_t>(NK) * 2 * D * sizeof(bf16), s));
    SLIME_PROPAGATE(gemm(c, R.pos_k, D, R.in_w + static_cast<size_t>(D) * D, D, NK, D, D, nullptr, nullptr, 0, 0,
                         nullptr, GEMM_EPI_NONE, R.derived_kvbias, nullptr, 2 * D, s));
  }
  return SLIME_OK;
}

enum Group { G_VIT = 1, G_RS_LOCAL = 2, G_RS_GLOBAL = 4, G_PROJ = 8, G_LLM = 16 };

int check_ready(slime_ctx* c, int groups) {
  if (c == nullptr) {
    slime_set_error("null context");
    return SLIME_EINVAL;
  }
  if (!c->finalized) {
    slime_set_error("weights not finalized: call slime_ctx_finalize_weights first");
    return SLIME_ESTATE;
  }
  const int have = (c->has_vit ? G_VIT : 0) | (c->has_rs[0] ? G_RS_LOCAL : 0) | (c->has_rs[1] ? G_RS_GLOBAL : 0) |
                   (c->has_proj ? G_PROJ : 0) | (c->has_llm ? G_LLM : 0);
  if ((have & groups) != groups) {
    slime_set_error("stage needs weight groups 0x%x but only 0x%x are registered (1 vit, 2 rs_local, 4 rs_global, "
                    "8 proj, 16 llm)", groups, have);
    return SLIME_ESTATE;
  }
  return SLIME_OK;
}

}  // namespace

// ==========================================================================================
// extern "C"
// ==========================================================================================
extern "C" {

int slime_version(void) { return SLIME_ABI_VERSION; }
int slime_elem_dtype(void) { return SLIME_ELEM_DTYPE; }
const char* slime_last_error(void) { return slime_get_error(); }

int slime_ctx_create(slime_ctx** out, int device, const slime_model_desc* desc) {
  SLIME_REQUIRE(out != nullptr && desc != nullptr, "ctx_create: null argument");
  cudaDeviceProp prop;
  SLIME_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    slime_set_error("device %d is sm_%d%d; slime_b200 only runs on sm_100 (B200) and has no fallback",
                    device, prop.major, prop.minor);
    return SLIME_EARCH;
  }
  const slime_model_desc& d = *desc;
  SLIME_REQUIRE(d.vit_hidden % 128 == 0 && d.vit_hidden % d.vit_heads == 0, "bad vit_hidden %d", d.vit_hidden);
  const int vhd = d.vit_hidden / d.vit_heads;
  SLIME_REQUIRE(vhd == 64 || vhd == 128, "ViT head_dim %d unsupported (64 or 128)", vhd);
  SLIME_REQUIRE(d.head_dim == 64 || d.head_dim == 128, "decoder head_dim %d unsupported", d.head_dim);
  SLIME_REQUIRE(d.vit_image % d.vit_patch == 0, "image %d not a multiple of patch %d", d.vit_image, d.vit_patch);
  SLIME_REQUIRE(d.hidden % 8 == 0 && d.mlp % 8 == 0 && d.vocab % 8 == 0, "hidden/mlp/vocab must be multiples of 8");
  SLIME_REQUIRE(d.kv_heads > 0 && d.heads % d.kv_heads == 0, "bad heads %d / kv_heads %d", d.heads, d.kv_heads);
  SLIME_REQUIRE(d.max_pos > 0 && d.temp > 0.f, "bad max_pos/temp");
  SLIME_CHECK_CUDA(cudaSetDevice(device));
  slime_ctx* c = new slime_ctx();
  c->d = d;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->vit_patches = (d.vit_image / d.vit_patch) * (d.vit_image / d.vit_patch);
  c->vit_tokens = c->vit_patches + 1;
  c->vit_kpad = ((3 * d.vit_patch * d.vit_patch + 63) / 64) * 64;
  cudaError_t e = cudaMalloc(&c->rope_table, static_cast<size_t>(d.max_pos) * (d.head_dim / 2) * 2 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->err_flag, sizeof(int));
  if (e != cudaSuccess) {
    slime_set_error("ctx_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    delete c;
    return SLIME_ECUDA;
  }
  cudaMemset(c->err_flag, 0, sizeof(int));
  int rc = slime_launch_rope_table(c->rope_table, d.max_pos, d.head_dim, d.rope_theta, nullptr);
  if (rc != SLIME_OK) {
    delete c;
    return rc;
  }
  SLIME_CHECK_CUDA(cudaDeviceSynchronize());
  *out = c;
  return SLIME_OK;
}

void slime_ctx_destroy(slime_ctx* ctx) {
  if (ctx == nullptr) return;
  if (ctx->rope_table) cudaFree(ctx->rope_table);
  if (ctx->err_flag) cudaFree(ctx->err_flag);
  delete ctx;
}

int slime_ctx_set_weight(slime_ctx* ctx, const char* name, const void* dev_ptr, int64_t rows, int64_t cols) {
  SLIME_REQUIRE(ctx != nullptr && name != nullptr && dev_ptr != nullptr, "set_weight: null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  Tensor t;
  t.p = static_cast<const bf16*>(dev_ptr);
  t.rows = rows;
  t.cols = cols;
  ctx->w[name] = t;
  ctx->finalized = false;
  return SLIME_OK;
}

size_t slime_finalize_workspace_bytes(const slime_ctx* ctx) {
  if (ctx == nullptr) return 0;
  const int nq = ctx->d.rs_global_queries > ctx->d.rs_local_queries ? ctx->d.rs_global_queries : ctx->d.rs_local_queries;
  return 2 * (static_cast<size_t>(nq) * ctx->d.vit_hidden * sizeof(bf16) + 256) + 256;
}

int slime_ctx_finalize_weights(slime_ctx* c, void* ws, size_t ws_bytes, void* stream) {
  SLIME_REQUIRE(c != nullptr, "finalize: null context");
  std::lock_guard<std::mutex> lk(c->mu);
  const slime_model_desc& d = c->d;
  const int D = d.vit_hidden, I = d.vit_mlp, H = d.hidden;
#define W(name, r, cc, dst) SLIME_PROPAGATE(find_weight(c, name, r, cc, &(dst)))
  c->has_vit = c->w.count("vit.patch_w") > 0;
  c->has_rs[0] = c->w.count("rs_local.query") > 0;
  c->has_rs[1] = c->w.count("rs_global.query") > 0;
  c->has_proj = c->w.count("proj.fc1_w") > 0;
  c->has_llm = c->w.count("llm.embed") > 0;
  SLIME_REQUIRE(c->has_vit || c->has_rs[0] || c->has_rs[1] || c->has_proj || c->has_llm || c->w.count("router.in_proj_w") > 0,
                "finalize: no weight group registered");
  if (c->has_vit) {
  W("vit.patch_w", D, c->vit_kpad, c->vit_patch_w);
  W("vit.cls", 1, D, c->vit_cls);
  W("vit.pos", c->vit_tokens, D, c->vit_pos);
  W("vit.pre_ln_w", 1, D, c->vit_pre_w);
  W("vit.pre_ln_b", 1, D, c->vit_pre_b);
  c->vit.resize(d.vit_layers_used);
  for (int l = 0; l < d.vit_layers_used; ++l) {
    const std::string p = "vit.layers." + std::to_string(l) + ".";
    VitLayer& L = c->vit[l];
    W(p + "ln1_w", 1, D, L.ln1_w);
    W(p + "ln1_b", 1, D, L.ln1_b);
    W(p + "qkv_w", 3 * D, D, L.qkv_w);
    W(p + "qkv_b", 1, 3 * D, L.qkv_b);
    W(p + "o_w", D, D, L.o_w);
    W(p + "o_b", 1, D, L.o_b);
    W(p + "ln2_w", 1, D, L.ln2_w);
    W(p + "ln2_b", 1, D, L.ln2_b);
    W(p + "fc1_w", I, D, L.fc1_w);
    W(p + "fc1_b", 1, I, L.fc1_b);
    W(p + "fc2_w", D, I, L.fc2_w);
    W(p + "fc2_b", 1, D, L.fc2_b);
  }
  }  // has_vit
  for (int which = 0; which < 2; ++which) {
    if (!c->has_rs[which]) continue;
    const std::string p = which == 0 ? "rs_local." : "rs_global.";
    Resampler& R = c->rs[which];
    R.nq = which == 0 ? d.rs_local_queries : d.rs_global_queries;
    const bf16 *dq = nullptr, *dk = nullptr;
    W(p + "query", R.nq, D, R.query);
    W(p + "pos_q", R.nq, D, R.pos_q);
    W(p + "pos_k", c->vit_patches, D, R.pos_k);
    W(p + "ln_q_w", 1, D, R.ln_q_w);
    W(p + "ln_q_b", 1, D, R.ln_q_b);
    W(p + "ln_kv_w", 1, D, R.ln_kv_w);
    W(p + "ln_kv_b", 1, D, R.ln_kv_b);
    W(p + "ln_post_w", 1, D, R.ln_post_w);
    W(p + "ln_post_b", 1, D, R.ln_post_b);
    W(p + "in_proj_w", 3 * D, D, R.in_w);
    W(p + "in_proj_b", 1, 3 * D, R.in_b);
    W(p + "out_w", D, D, R.out_w);
    W(p + "out_b", 1, D, R.out_b);
    W(p + "derived_q", R.nq, D, dq);
    W(p + "derived_kvbias", c->vit_patches, 2 * D, dk);
    R.derived_q = const_cast<bf16*>(dq);
    R.derived_kvbias = const_cast<bf16*>(dk);
  }
  if (c->has_proj) {
  W("proj.fc1_w", H, D, c->proj_fc1_w);
  W("proj.fc1_b", 1, H, c->proj_fc1_b);
  W("proj.fc2_w", H, H, c->proj_fc2_w);
  W("proj.fc2_b", 1, H, c->proj_fc2_b);
  W("proj.w_gate", D, 2, c->proj_w_gate);
  }  // has_proj
  c->has_router = c->w.count("router.in_proj_w") > 0;
  if (c->has_router) {
  SLIME_REQUIRE(H % 128 == 0 && (H / 4) % 8 == 0, "qformer router needs hidden_size %% 128 == 0 (heads = H/128)");
  W("router.ln_q_w", 1, H, c->qf.ln_q_w);
  W("router.ln_q_b", 1, H, c->qf.ln_q_b);
  W("router.ln_kv_w", 1, H, c->qf.ln_kv_w);
  W("router.ln_kv_b", 1, H, c->qf.ln_kv_b);
  W("router.ln_post_w", 1, H, c->qf.ln_post_w);
  W("router.ln_post_b", 1, H, c->qf.ln_post_b);
  W("router.in_proj_w", 3 * H, H, c->qf.in_w);
  W("router.in_proj_b", 1, 3 * H, c->qf.in_b);
  W("router.out_w", H, H, c->qf.out_w);
  W("router.out_b", 1, H, c->qf.out_b);
  W("router.fc1_w", H / 4, H, c->qf.fc1_w);
  W("router.fc1_b", 1, H / 4, c->qf.fc1_b);
  W("router.fc2_w", 1, H / 4, c->qf.fc2_w);
  W("router.fc2_b", 1, 1, c->qf.fc2_b);
  }  // has_router
  const int QKV = (d.heads + 2 * d.kv_heads) * d.head_dim;
  if (c->has_llm) {
  W("llm.embed", d.vocab, H, c->llm_embed);
  W("llm.norm_w", 1, H, c->llm_norm_w);
  W("llm.lm_head", d.vocab, H, c->llm_lm_head);
  c->llm.resize(d.layers);
  for (int l = 0; l < d.layers; ++l) {
    const std::string p = "llm.layers." + std::to_string(l) + ".";
    LlmLayer& L = c->llm[l];
    W(p + "in_norm_w", 1, H, L.in_norm_w);
    W(p + "qkv_w", QKV, H, L.qkv_w);
    W(p + "o_w", H, d.heads * d.head_dim, L.o_w);
    W(p + "post_norm_w", 1, H, L.post_norm_w);
    W(p + "gate_up_w", 2 * d.mlp, H, L.gate_up_w);
    W(p + "down_w", H, d.mlp, L.down_w);
  }
  }  // has_llm
#undef W
  SLIME_REQUIRE(D % 128 == 0, "Resampler needs mm_hidden_size %% 128 == 0 (heads = D/128)");
  Arena a(ws, ws_bytes);
  SLIME_REQUIRE(ws != nullptr, "finalize: null workspace");
  SLIME_PROPAGATE(finalize_body(c, a, static_cast<cudaStream_t>(stream)));
  c->finalized = true;
  return SLIME_OK;
}

// ---- vision tower ----
size_t slime_vision_tower_workspace_bytes(const slime_ctx* ctx, int n_crops) {
  Arena a(nullptr, 0);
  vit_body(const_cast<slime_ctx*>(ctx), a, nullptr, n_crops > 0 ? n_crops : 1, nullptr, nullptr);
  return a.off + 256;
}
int slime_vision_tower_fwd(slime_ctx* ctx, const void* pixels, int n_crops, void* feats, void* ws,
                           size_t ws_bytes, void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_VIT));
  SLIME_REQUIRE(pixels && feats && ws, "vision_tower: null pointer");
  if (n_crops <= 0) return SLIME_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return vit_body(ctx, a, static_cast<const bf16*>(pixels), n_crops, static_cast<bf16*>(feats),
                  static_cast<cudaStream_t>(stream));
}

int slime_vision_tower_fwd_split(slime_ctx* ctx, const void* pixels, int n_images, int crops_per_image, void* feats,
                                 void* ws, size_t ws_bytes, void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_VIT));
  SLIME_REQUIRE(pixels && feats && ws, "vision_tower: null pointer");
  SLIME_REQUIRE(crops_per_image >= 1, "vision_tower: crops_per_image must be >= 1 (%d given)", crops_per_image);
  if (n_images <= 0) return SLIME_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return vit_body(ctx, a, static_cast<const bf16*>(pixels), n_images * crops_per_image, static_cast<bf16*>(feats),
                  static_cast<cudaStream_t>(stream), crops_per_image);
}

// ---- resampler ----
size_t slime_resampler_workspace_bytes(const slime_ctx* ctx, int which, int n) {
  Arena a(nullptr, 0);
  resampler_body(const_cast<slime_ctx*>(ctx), a, which, nullptr, n > 0 ? n : 1, nullptr, nullptr);
  return a.off + 256;
}
int slime_resampler_fwd(slime_ctx* ctx, int which, const void* x, int n, void* out, void* ws, size_t ws_bytes,
                        void* stream) {
  SLIME_REQUIRE(which == 0 || which == 1, "resampler: which must be 0 (local) or 1 (global)");
  SLIME_PROPAGATE(check_ready(ctx, which == 0 ? G_RS_LOCAL : G_RS_GLOBAL));
  if (n <= 0) return SLIME_OK;
  SLIME_REQUIRE(x && out && ws, "resampler: null pointer");
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return resampler_body(ctx, a, which, static_cast<const bf16*>(x), n, static_cast<bf16*>(out),
                        static_cast<cudaStream_t>(stream));
}

// ---- projector ----
size_t slime_projector_workspace_bytes(const slime_ctx* ctx, int rows) {
  Arena a(nullptr, 0);
  projector_body(const_cast<slime_ctx*>(ctx), a, nullptr, rows > 0 ? rows : 1, nullptr, nullptr, nullptr);
  return a.off + 256;
}
int slime_projector_fwd(slime_ctx* ctx, const void* x, int rows, const int32_t* row_map, void* out, void* ws,
                        size_t ws_bytes, void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_PROJ));
  if (rows <= 0) return SLIME_OK;
  SLIME_REQUIRE(x && out && ws, "projector: null pointer");
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return projector_body(ctx, a, static_cast<const bf16*>(x), rows, row_map, static_cast<bf16*>(out),
                        static_cast<cudaStream_t>(stream));
}
size_t slime_gated_projector_workspace_bytes(const slime_ctx* ctx, int n) {
  Arena a(nullptr, 0);
  gated_body(const_cast<slime_ctx*>(ctx), a, nullptr, n > 0 ? n : 1, nullptr, nullptr);
  return a.off + 256;
}
int slime_gated_projector_fwd(slime_ctx* ctx, const void* x, int n, void* out, void* ws, size_t ws_bytes,
                              void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_PROJ | (ctx != nullptr && ctx->d.mm_learnable_gated == 0 ? 0 : G_RS_GLOBAL)));
  if (n <= 0) return SLIME_OK;
  SLIME_REQUIRE(x && out && ws, "gated_projector: null pointer");
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return gated_body(ctx, a, static_cast<const bf16*>(x), n, static_cast<bf16*>(out),
                    static_cast<cudaStream_t>(stream));
}

// ---- router ----
size_t slime_router_workspace_bytes(const slime_ctx* ctx, int batch, int n_per, int prompt_len) {
  Arena a(nullptr, 0);
  router_body(const_cast<slime_ctx*>(ctx), a, nullptr, n_per, nullptr, nullptr, nullptr, nullptr, batch > 0 ? batch : 1,
              prompt_len > 0 ? prompt_len : 1, nullptr, nullptr, nullptr, nullptr);
  return a.off + 256;
}
int slime_router_fwd(slime_ctx* ctx, const void* local, int n_per, const int32_t* n_valid, const int64_t* ids,
                     const uint8_t* mask, int batch, int prompt_len, float* probs_out, int32_t* sel_idx,
                     int32_t* sel_count, void* ws, size_t ws_bytes, void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_LLM));
  SLIME_REQUIRE(ids && sel_idx && sel_count && ws, "router: null pointer");
  SLIME_REQUIRE(n_per <= 0 || local != nullptr, "router: null local features");
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return router_body(ctx, a, static_cast<const bf16*>(local), n_per, n_valid,
                     reinterpret_cast<const long long*>(ids), nullptr, mask, batch, prompt_len, probs_out, sel_idx,
                     sel_count, static_cast<cudaStream_t>(stream));
}
int slime_router_fwd_embeds(slime_ctx* ctx, const void* local, int n_per, const int32_t* n_valid,
                            const void* text_embeds, const uint8_t* mask, int batch, int prompt_len, float* probs_out,
                            int32_t* sel_idx, int32_t* sel_count, void* ws, size_t ws_bytes, void* stream) {
  SLIME_REQUIRE(ctx != nullptr && text_embeds && sel_idx && sel_count && ws, "router_embeds: null pointer");
  SLIME_REQUIRE(n_per <= 0 || local != nullptr, "router_embeds: null local features");
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return router_body(ctx, a, static_cast<const bf16*>(local), n_per, n_valid, nullptr,
                     static_cast<const bf16*>(text_embeds), mask, batch, prompt_len, probs_out, sel_idx, sel_count,
                     static_cast<cudaStream_t>(stream));
}
int slime_router_select(slime_ctx* ctx, const float* probs, int batch, int n_per, const int32_t* n_valid,
                        int32_t* sel_idx, int32_t* sel_count, void* stream) {
  SLIME_REQUIRE(ctx != nullptr && probs && sel_idx && sel_count, "router_select: null pointer");
  return slime_launch_router_select(probs, batch, n_per, n_valid, ctx->d.temp, ctx->d.top_p, 1, nullptr,
                                    sel_idx, sel_count, static_cast<cudaStream_t>(stream));
}

// ---- splice ----
size_t slime_splice_plan_ints(int batch, int prompt_len) {
  return static_cast<size_t>(batch) * prompt_len + static_cast<size_t>(batch) * SLIME_PLAN_STRIDE + batch + 1 + 64;
}
// plan_buf layout: [valid_pos B*T][plan B*STRIDE][cu B+1]
static inline int* plan_valid(int32_t* buf) { return buf; }
static inline int* plan_plan(int32_t* buf, int B, int T) { return buf + static_cast<size_t>(B) * T; }
static inline int* plan_cu(int32_t* buf, int B, int T) {
  return buf + static_cast<size_t>(B) * T + static_cast<size_t>(B) * SLIME_PLAN_STRIDE;
}

int slime_splice_plan(slime_ctx* ctx, const int64_t* ids, const uint8_t* mask, int batch, int prompt_len,
                      int n_global, int has_sep, const int32_t* sel_count, int32_t* plan_buf, int32_t* host_cu,
                      void* stream) {
  SLIME_REQUIRE(ctx && ids && plan_buf && host_cu, "splice_plan: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lk(ctx->mu);
  SLIME_CHECK_CUDA(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s));
  SLIME_PROPAGATE(slime_launch_splice_plan(reinterpret_cast<const long long*>(ids), mask, batch, prompt_len,
                                           ctx->d.image_token, n_global, has_sep, sel_count, ctx->d.max_len,
                                           plan_valid(plan_buf), plan_plan(plan_buf, batch, prompt_len),
                                           plan_cu(plan_buf, batch, prompt_len), ctx->err_flag, s));
  int err = 0;
  SLIME_CHECK_CUDA(cudaMemcpyAsync(host_cu, plan_cu(plan_buf, batch, prompt_len), sizeof(int) * (batch + 1),
                                   cudaMemcpyDeviceToHost, s));
  SLIME_CHECK_CUDA(cudaMemcpyAsync(&err, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  SLIME_CHECK_CUDA(cudaStreamSynchronize(s));
  if (err != 0) {
    slime_set_error("splice: a prompt holds more than one image placeholder; the SliME path pairs exactly "
                    "one image with each sample (llava_arch.py:222-255)");
    return SLIME_EINVAL;
  }
  return SLIME_OK;
}

int slime_splice_plan_async(slime_ctx* ctx, const int64_t* ids, const uint8_t* mask, int batch, int prompt_len,
                            int n_global, int has_sep, const int32_t* sel_count, int32_t* plan_buf, void* stream) {
  SLIME_REQUIRE(ctx && ids && plan_buf, "splice_plan_async: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lk(ctx->mu);
  SLIME_CHECK_CUDA(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s));
  return slime_launch_splice_plan(reinterpret_cast<const long long*>(ids), mask, batch, prompt_len, ctx->d.image_token,
                                  n_global, has_sep, sel_count, ctx->d.max_len, plan_valid(plan_buf),
                                  plan_plan(plan_buf, batch, prompt_len), plan_cu(plan_buf, batch, prompt_len),
                                  ctx->err_flag, s);
}

int slime_splice_check(slime_ctx* ctx, void* stream) {
  SLIME_REQUIRE(ctx != nullptr, "splice_check: null context");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int err = 0;
  SLIME_CHECK_CUDA(cudaMemcpyAsync(&err, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  SLIME_CHECK_CUDA(cudaStreamSynchronize(s));
  if (err != 0) {
    slime_set_error("splice: a prompt holds more than one image placeholder; the SliME path pairs exactly "
                    "one image with each sample (llava_arch.py:222-255)");
    return SLIME_EINVAL;
  }
  return SLIME_OK;
}

int slime_splice_gather(slime_ctx* ctx, const int64_t* ids, int batch, int prompt_len, const int32_t* plan_buf,
                        const void* global_feats, int n_global, int64_t global_sample_rows,
                        const void* local_feats, int64_t local_sample_rows, const int32_t* sel_idx,
                        int sel_stride, int has_sep, void* out_embeds, int32_t* pos_ids, int total_rows,
                        void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_LLM));
  SLIME_REQUIRE(ids && plan_buf && out_embeds, "splice_gather: null pointer");
  int32_t* pb = const_cast<int32_t*>(plan_buf);
  return slime_launch_splice_gather(reinterpret_cast<const long long*>(ids), prompt_len, plan_valid(pb),
                                    plan_plan(pb, batch, prompt_len), plan_cu(pb, batch, prompt_len), batch,
                                    ctx->llm_embed, ctx->d.hidden, ctx->d.sep_token, has_sep,
                                    static_cast<const bf16*>(global_feats), n_global, global_sample_rows,
                                    static_cast<const bf16*>(local_feats), local_sample_rows, sel_idx, sel_stride,
                                    static_cast<bf16*>(out_embeds), pos_ids, total_rows,
                                    static_cast<cudaStream_t>(stream));
}

int slime_splice_pad(slime_ctx* ctx, const int32_t* plan_buf, const void* packed_embeds, const int64_t* labels_in,
                     int batch, int prompt_len, int lmax, void* out_embeds, uint8_t* out_mask, int64_t* out_pos,
                     int64_t* out_labels, void* stream) {
  SLIME_REQUIRE(ctx && plan_buf, "splice_pad: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t* pb = const_cast<int32_t*>(plan_buf);
  const int left = (ctx->d.flags & SLIME_FLAG_LEFT_PAD) ? 1 : 0;
  if (out_mask || out_pos || out_labels) {
    SLIME_PROPAGATE(slime_launch_splice_pad_meta(plan_plan(pb, batch, prompt_len), plan_valid(pb),
                                                 reinterpret_cast<const long long*>(labels_in), prompt_len, batch,
                                                 lmax, left, -100, out_mask, reinterpret_cast<long long*>(out_pos),
                                                 reinterpret_cast<long long*>(out_labels), s));
  }
  if (out_embeds != nullptr) {
    SLIME_REQUIRE(packed_embeds != nullptr, "splice_pad: packed embeds missing");
    SLIME_PROPAGATE(slime_launch_splice_pad_embeds(static_cast<const bf16*>(packed_embeds),
                                                   plan_cu(pb, batch, prompt_len), batch, lmax, ctx->d.hidden, left,
                                                   static_cast<bf16*>(out_embeds), s));
  }
  return SLIME_OK;
}

// ---- decoder ----
size_t slime_decoder_workspace_bytes(const slime_ctx* ctx, int total_rows, int batch) {
  Arena a(nullptr, 0);
  decoder_body(const_cast<slime_ctx*>(ctx), a, nullptr, nullptr, nullptr, batch > 0 ? batch : 1,
               total_rows > 0 ? total_rows : 1, 0, nullptr, nullptr, nullptr, nullptr);
  return a.off + 256;
}
int slime_decoder_prefill_fwd(slime_ctx* ctx, const void* embeds, const int32_t* cu_seqlens, const int32_t* pos_ids,
                              int batch, int total_rows, int max_seqlen, float* logits_last, void* logits_all,
                              void* hidden_out, void* ws, size_t ws_bytes, void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_LLM));
  SLIME_REQUIRE(embeds && cu_seqlens && pos_ids && ws, "decoder: null pointer");
  SLIME_REQUIRE(max_seqlen <= ctx->d.max_pos, "decoder: sequence length %d exceeds max_pos %d", max_seqlen,
                ctx->d.max_pos);
  SLIME_REQUIRE(ctx->kv_cache == nullptr || batch <= ctx->kv_cache_batch,
                "decoder: batch %d exceeds the attached KV cache's %d sequences", batch, ctx->kv_cache_batch);
  if (total_rows <= 0 || batch <= 0) return SLIME_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return decoder_body(ctx, a, static_cast<const bf16*>(embeds), cu_seqlens, pos_ids, batch, total_rows, max_seqlen,
                      logits_last, static_cast<bf16*>(logits_all), static_cast<bf16*>(hidden_out),
                      static_cast<cudaStream_t>(stream));
}

// ---- KV cache + decode step ----
size_t slime_kv_cache_bytes(const slime_ctx* ctx, int batch, int cache_len) {
  if (ctx == nullptr || batch <= 0 || cache_len <= 0) return 0;
  return static_cast<size_t>(ctx->d.layers) * 2 * batch * cache_len * ctx->d.kv_heads * ctx->d.head_dim * sizeof(bf16);
}
int slime_decoder_set_kv_cache(slime_ctx* ctx, void* cache, int batch, int cache_len) {
  SLIME_REQUIRE(ctx != nullptr, "set_kv_cache: null context");
  SLIME_REQUIRE(cache == nullptr || (batch > 0 && cache_len > 0 && cache_len <= ctx->d.max_pos),
                "set_kv_cache: bad geometry batch=%d cache_len=%d (max_pos %d)", batch, cache_len, ctx->d.max_pos);
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->kv_cache = static_cast<bf16*>(cache);
  ctx->kv_cache_batch = cache != nullptr ? batch : 0;
  ctx->kv_cache_len = cache != nullptr ? cache_len : 0;
  return SLIME_OK;
}
size_t slime_decoder_decode_workspace_bytes(const slime_ctx* ctx, int batch) {
  Arena a(nullptr, 0);
  decode_body(const_cast<slime_ctx*>(ctx), a, nullptr, nullptr, batch > 0 ? batch : 1, nullptr, nullptr);
  return a.off + 256;
}
int slime_decoder_decode_fwd(slime_ctx* ctx, const void* x, const int32_t* lens, int batch, float* logits, void* ws,
                             size_t ws_bytes, void* stream) {
  SLIME_PROPAGATE(check_ready(ctx, G_LLM));
  SLIME_REQUIRE(x && lens && logits && ws, "decode: null pointer");
  SLIME_REQUIRE(ctx->kv_cache != nullptr, "decode: no KV cache attached (slime_decoder_set_kv_cache)");
  SLIME_REQUIRE(batch <= ctx->kv_cache_batch, "decode: batch %d exceeds the KV cache's %d sequences", batch,
                ctx->kv_cache_batch);
  if (batch <= 0) return SLIME_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  Arena a(ws, ws_bytes);
  return decode_body(ctx, a, static_cast<const bf16*>(x), lens, batch, logits, static_cast<cudaStream_t>(stream));
}

// ---- single ops ----
int slime_op_gemm(const void* a, int lda, const void* w, int ldw, int m, int n, int k, const void* bias,
                  const void* residual, int res_ld, int res_period, const int32_t* row_map, int epilogue, void* out,
                  float* out_f32, int out_ld, void* stream) {
  GemmParams p;
  p.M = m; p.N = n; p.K = k;
  p.bias = static_cast<const bf16*>(bias);
  p.residual = static_cast<const bf16*>(residual);
  p.res_ld = res_ld;
  p.res_period = res_period;
  p.row_map = row_map;
  p.out = static_cast<bf16*>(out);
  p.out_f32 = out_f32;
  p.out_ld = out_ld;
  int dev = 0, sms = 148;
  SLIME_CHECK_CUDA(cudaGetDevice(&dev));
  SLIME_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return slime_launch_gemm(static_cast<const bf16*>(a), lda, static_cast<const bf16*>(w), ldw, p, epilogue, sms,
                           static_cast<cudaStream_t>(stream));
}

int slime_set_decode_prefetch(int mask) {
  g_decode_pf = mask;
  return SLIME_OK;
}

int slime_set_decode_fused(int mask) {
  g_decode_fused = (mask < 0 || mask > 7) ? -1 : mask;
  return SLIME_OK;
}

int slime_op_gemm_skinny(const void* a, int lda, const void* w, int ldw, int m, int n, int k, const void* bias,
                         const void* residual, int res_ld, int epilogue, void* out, float* out_f32, int out_ld,
                         int splits, float* ws, size_t ws_floats, const void* norm_w, void* norm_out, float norm_eps,
                         const int32_t* rope_pos, const float* rope_table, int rope_half, int rope_cols, int rope_max_pos,
                         void* stream) {
  return slime_op_gemm_skinny_fused(a, lda, w, ldw, m, n, k, bias, residual, res_ld, epilogue, out, out_f32, out_ld, splits,
                                    ws, ws_floats, norm_w, norm_out, norm_eps, rope_pos, rope_table, rope_half, rope_cols,
                                    rope_max_pos, nullptr, nullptr, 0.f, stream);
}

int slime_op_gemm_skinny_fused(const void* a, int lda, const void* w, int ldw, int m, int n, int k, const void* bias,
                               const void* residual, int res_ld, int epilogue, void* out, float* out_f32, int out_ld,
                               int splits, float* ws, size_t ws_floats, const void* norm_w, void* norm_out, float norm_eps,
                               const int32_t* rope_pos, const float* rope_table, int rope_half, int rope_cols,
                               int rope_max_pos, int32_t* tile_counters, const void* a_norm_w, float a_norm_eps,
                               void* stream) {
  GemmParams p;
  p.tile_counters = tile_counters;
  p.a_norm_w = static_cast<const bf16*>(a_norm_w);
  p.a_norm_eps = a_norm_eps;
  p.M = m; p.N = n; p.K = k;
  p.bias = static_cast<const bf16*>(bias);
  p.residual = static_cast<const bf16*>(residual);
  p.res_ld = res_ld;
  p.res_period = 0;
  p.row_map = nullptr;
  p.out = static_cast<bf16*>(out);
  p.out_f32 = out_f32;
  p.out_ld = out_ld;
  p.force_splits = splits;
  p.splitk_ws = ws;
  p.splitk_ws_floats = ws_floats;
  p.norm_w = static_cast<const bf16*>(norm_w);
  p.norm_out = static_cast<bf16*>(norm_out);
  p.norm_ld = out_ld;
  p.norm_eps = norm_eps;
  p.rope_pos = rope_pos;
  p.rope_table = reinterpret_cast<const float2*>(rope_table);
  p.rope_half = rope_half;
  p.rope_cols = rope_cols;
  p.rope_max_pos = rope_max_pos;
  SLIME_REQUIRE(splits >= 1, "gemm_skinny: splits must be >= 1");
  int dev = 0, sms = 148;
  SLIME_CHECK_CUDA(cudaGetDevice(&dev));
  SLIME_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return slime_launch_gemm_skinny(static_cast<const bf16*>(a), lda, static_cast<const bf16*>(w), ldw, p, epilogue, sms,
                                  static_cast<cudaStream_t>(stream));
}

int slime_op_decode_attention(const void* q, int q_ld, const void* kcache, const void* vcache, int cache_len,
                              const int32_t* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                              void* out, int out_ld, int splits, float* ws, void* stream) {
  return slime_op_decode_attention_fused(q, q_ld, kcache, vcache, cache_len, lens, batch, heads, kv_heads, head_dim, scale,
                                         out, out_ld, splits, ws, nullptr, stream);
}

int slime_op_decode_attention_fused(const void* q, int q_ld, const void* kcache, const void* vcache, int cache_len,
                                    const int32_t* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                                    void* out, int out_ld, int splits, float* ws, int32_t* merge_counters, void* stream) {
  return slime_launch_decode_attention(static_cast<const bf16*>(q), q_ld, static_cast<const bf16*>(kcache),
                                       static_cast<const bf16*>(vcache), cache_len, lens, batch, heads, kv_heads,
                                       head_dim, scale, static_cast<bf16*>(out), out_ld, splits, ws, nullptr, 0,
                                       static_cast<cudaStream_t>(stream), merge_counters);
}

int slime_op_attention(const void* q, const void* k, const void* v, void* o, int q_ld, int k_ld, int v_ld, int o_ld,
                       const int32_t* cu_q, const int32_t* cu_k, int seqlen_q, int seqlen_k, int64_t q_batch_rows,
                       int64_t k_batch_rows, int64_t o_batch_rows, int batch, int heads, int kv_heads, int head_dim,
                       float scale, int causal, int64_t total_q_rows, int64_t total_k_rows, int impl, void* stream) {
  AttnParams ap;
  ap.q = static_cast<const bf16*>(q);
  ap.k = static_cast<const bf16*>(k);
  ap.v = static_cast<const bf16*>(v);
  ap.o = static_cast<bf16*>(o);
  ap.q_ld = q_ld; ap.k_ld = k_ld; ap.v_ld = v_ld; ap.o_ld = o_ld;
  ap.cu_q = cu_q; ap.cu_k = cu_k;
  ap.seqlen_q = seqlen_q; ap.seqlen_k = seqlen_k;
  ap.q_batch_rows = q_batch_rows; ap.k_batch_rows = k_batch_rows; ap.o_batch_rows = o_batch_rows;
  ap.batch = batch; ap.num_heads = heads; ap.num_kv_heads = kv_heads; ap.head_dim = head_dim;
  ap.scale = scale; ap.causal = causal;
  ap.total_q_rows = total_q_rows; ap.total_k_rows = total_k_rows;
  (void)impl;  // (kept in the signature: there is one prefill attention kernel now)
  return slime_launch_attention(ap, static_cast<cudaStream_t>(stream));
}

int slime_op_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int dim, float eps,
                       void* stream) {
  return slime_launch_layernorm(static_cast<const bf16*>(x), dim, static_cast<const bf16*>(w),
                                static_cast<const bf16*>(b), static_cast<bf16*>(y), dim, rows, dim, eps, 0, 0, 0,
                                static_cast<cudaStream_t>(stream));
}
int slime_op_rmsnorm(const void* x, const void* w, void* y, int rows, int dim, float eps, void* stream) {
  return slime_launch_rmsnorm(static_cast<const bf16*>(x), dim, static_cast<const bf16*>(w), static_cast<bf16*>(y),
                              dim, rows, dim, eps, nullptr, static_cast<cudaStream_t>(stream));
}
int slime_op_rope(slime_ctx* ctx, void* qkv, int ld, int rows, const int32_t* pos_ids, void* stream) {
  SLIME_REQUIRE(ctx != nullptr, "rope: null context");
  return slime_launch_rope(static_cast<bf16*>(qkv), ld, rows, ctx->d.heads, ctx->d.kv_heads, ctx->d.head_dim, pos_ids,
                           ctx->rope_table, ctx->d.max_pos, static_cast<cudaStream_t>(stream));
}
int slime_op_qkv_rope(slime_ctx* ctx, const void* x, const void* qkv_w, int rows, const int32_t* pos_ids, void* qkv,
                      void* stream) {
  SLIME_REQUIRE(ctx != nullptr && x != nullptr && qkv_w != nullptr && qkv != nullptr && pos_ids != nullptr,
                "qkv_rope: null argument");
  if (rows <= 0) return SLIME_OK;
  return qkv_rope(ctx, static_cast<const bf16*>(x), static_cast<const bf16*>(qkv_w), rows, pos_ids,
                  static_cast<bf16*>(qkv), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
