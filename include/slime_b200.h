/* slime_b200 C-ABI: the drop-in boundary of the B200-native SliME prefill path.
 *
 * The reference (yfzhang114/SliME) has no FFI of its own - its seam is the Python module API of
 * llava/model.  The Python shims in slime_b200/ keep that API and forward
 * every stage to the entry points below through ctypes.  Each entry point names the reference
 * function it replaces (paths relative to the reference repo; "HF:" = transformers).
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer owned by the caller (PyTorch); bf16 unless stated;
 *     row-major, 16-byte aligned.  The library never frees or retains activation pointers; weight
 *     pointers registered with slime_ctx_set_weight are borrowed until slime_ctx_destroy.
 *   - `ws`/`ws_bytes` is caller-provided scratch; size it with the matching *_workspace_bytes().
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); only
 *     slime_splice_plan synchronises (it must hand sequence lengths to the host).
 *   - return value: 0 on success, negative SLIME_E* on failure; slime_last_error() gives the text.
 *   - sm_100 only: slime_ctx_create fails with SLIME_EARCH on any other device.  No CPU fallback.
 */
#ifndef SLIME_B200_H_
#define SLIME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIME_ABI_VERSION 1

#define SLIME_OK 0
#define SLIME_EINVAL (-1)
#define SLIME_ECUDA (-2)
#define SLIME_EARCH (-3)
#define SLIME_EWORKSPACE (-4)
#define SLIME_ESTATE (-5)

#define SLIME_FLAG_LEFT_PAD 1u         /* config.tokenizer_padding_side == "left" */
#define SLIME_FLAG_USE_GLOBAL_ONLY 2u  /* config.use_global_only */
#define SLIME_FLAG_USE_LOCAL_ONLY 4u   /* config.use_local_only */
/* llm.layers.N.qkv_w was registered with the q / k rows of every head interleaved - row 2i = q_proj row i, row
 * 2i+1 = q_proj row i + head_dim/2 (same for k_proj) - so the rotary embedding (HF llama/modeling_llama.py:152-176)
 * is applied in the epilogue of the QKV GEMM instead of by a separate pass over q and k. */
#define SLIME_FLAG_ROPE_INTERLEAVED 8u
/* config.mm_resampler_type == "qformer": slime_router_fwd / slime_router_fwd_embeds score the local tokens with the
 * cross-attention router (reference multimodal_resampler/builder.py:94-162 TextGuidedRouterAttention; weight group
 * "router.*") instead of the cosine router.  Like the reference, its softmax output is soft-maxed once more by the
 * sampler (builder.py:160 and :258) before the top-p rule. */
#define SLIME_FLAG_ROUTER_QFORMER 16u
#define SLIME_FLAG_NORM_FOLDED 32u /* the Llama RMSNorm weights are already multiplied into the columns of the qkv / gate-up
                                     weights (slime_b200/weights.py): the prefill applies 1/rms in those GEMMs' epilogues
                                     instead of running separate RMSNorm passes; the decode step normalises without gamma */

typedef struct slime_ctx slime_ctx;

/* Model description = the config attributes the reference reads on the hot path. */
typedef struct slime_model_desc {
  /* CLIP ViT (HF:models/clip/modeling_clip.py; mm_vision_tower) */
  int32_t vit_hidden;       /* 1024 */
  int32_t vit_layers_used;  /* layers executed = num_hidden_layers + 1 + mm_vision_select_layer (23) */
  int32_t vit_heads;        /* 16 */
  int32_t vit_mlp;          /* 4096 */
  int32_t vit_image;        /* 336 */
  int32_t vit_patch;        /* 14 */
  float vit_ln_eps;         /* 1e-5 */
  /* Resampler / projector (multimodal_resampler/sampler.py, multimodal_projector/builder.py) */
  int32_t rs_local_queries;   /* mm_resampler_dim = 144 */
  int32_t rs_global_queries;  /* GatedBlock.target_sequence_length = 576 */
  float rs_ln_eps;            /* 1e-6 */
  int32_t mm_learnable_gated; /* -1: gated mix, 0/1: that expert only */
  /* Llama decoder (HF:models/llama/modeling_llama.py) */
  int32_t hidden;
  int32_t layers;
  int32_t heads;
  int32_t kv_heads;
  int32_t head_dim;
  int32_t mlp;
  int32_t vocab;
  float rope_theta;
  float rms_eps;
  int32_t max_pos; /* RoPE table length (>= longest spliced sequence) */
  /* router / splice */
  float top_p;           /* mm_resampler_topp */
  float temp;            /* mm_resampler_temp */
  int64_t image_token;   /* IMAGE_TOKEN_INDEX = -200 (llava/constants.py:9) */
  int64_t sep_token;     /* config.seperator */
  int32_t max_len;       /* tokenizer_model_max_length, 0 = no truncation */
  uint32_t flags;        /* SLIME_FLAG_* */
} slime_model_desc;

int slime_version(void);
/* 16-bit element type this build computes in: 0 = bfloat16 (libslime_b200.so), 2 = IEEE half (libslime_b200_fp16.so,
 * the same sources compiled with -DSLIME_FP16).  Every "bf16" tensor in this header is of that type. */
int slime_elem_dtype(void);
const char* slime_last_error(void);

int slime_ctx_create(slime_ctx** out, int device, const slime_model_desc* desc);
void slime_ctx_destroy(slime_ctx* ctx);

/* Register one weight tensor by canonical name (see slime_b200/weights.py for the mapping from the
 * reference state-dict keys).  rows/cols are validated against the model description. */
int slime_ctx_set_weight(slime_ctx* ctx, const char* name, const void* dev_ptr, int64_t rows, int64_t cols);
/* Validates that every weight is present and pre-computes the input-independent tensors of the two
 * Resamplers (projected queries, projected key position table) into the registered "*.derived_*"
 * buffers.  Must be called once after the weights are registered (and again if they change). */
size_t slime_finalize_workspace_bytes(const slime_ctx* ctx);
int slime_ctx_finalize_weights(slime_ctx* ctx, void* ws, size_t ws_bytes, void* stream);

/* ---- stage 1: CLIPVisionTower.forward + feature_select  (multimodal_encoder/clip_encoder.py:36-58)
 * pixels [Nc,3,S,S] -> feats [Nc, (S/P)^2, vit_hidden] = hidden_states[select_layer] without CLS. */
size_t slime_vision_tower_workspace_bytes(const slime_ctx* ctx, int n_crops);
int slime_vision_tower_fwd(slime_ctx* ctx, const void* pixels, int n_crops, void* feats, void* ws,
                           size_t ws_bytes, void* stream);
/* Same for n_images images of crops_per_image crops each (crop 0 of an image = its global view; pixels
 * [n_images * crops_per_image, 3, S, S]), with the output grouped for the adapter stages: feats[0 .. n_images) = the global
 * crops, feats[n_images ..) = the local crops in image order - the two slices the reference cuts per sample
 * (llava/model/llava_arch.py:212-225) are contiguous and need no gather.  Workspace as slime_vision_tower_fwd. */
int slime_vision_tower_fwd_split(slime_ctx* ctx, const void* pixels, int n_images, int crops_per_image, void* feats,
                                 void* ws, size_t ws_bytes, void* stream);

/* ---- stage 2: Resampler.forward  (multimodal_resampler/sampler.py:140-170)
 * which = 0: sampler.post_qformer (local compression, 576 -> 144 tokens per crop)
 * which = 1: mm_projector.attn    (global 576-query resampler)
 * x [n, 576, vit_hidden] -> out [n, nq, vit_hidden]. */
size_t slime_resampler_workspace_bytes(const slime_ctx* ctx, int which, int n);
int slime_resampler_fwd(slime_ctx* ctx, int which, const void* x, int n, void* out, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- stage 3a: GatedBlock.projection (multimodal_projector/builder.py:53-57,180-181)
 * x [rows, vit_hidden] -> out[row_map ? row_map[r] : r, :hidden]; row_map folds the spatial merge
 * (llava_arch.py:240-244) into the store. */
size_t slime_projector_workspace_bytes(const slime_ctx* ctx, int rows);
int slime_projector_fwd(slime_ctx* ctx, const void* x, int rows, const int32_t* row_map, void* out,
                        void* ws, size_t ws_bytes, void* stream);
/* ---- stage 3b: GatedBlock.forward on the global crop (multimodal_projector/builder.py:179-209)
 * x [n*576, vit_hidden] -> out [n*576, hidden]. */
size_t slime_gated_projector_workspace_bytes(const slime_ctx* ctx, int n);
int slime_gated_projector_fwd(slime_ctx* ctx, const void* x, int n, void* out, void* ws, size_t ws_bytes,
                              void* stream);

/* ---- stage 4: TextGuidedSampler.forward (multimodal_resampler/builder.py:248-281) with the cosine
 * selector (:189-201) and get_pure_text_embedding (llava_arch.py:162-210) folded in.
 * local [B, n_per, hidden] (sample b uses its first n_valid[b] rows; n_valid may be NULL),
 * ids [B,T] int64, mask [B,T] uint8 (may be NULL = all ones)
 * -> sel_idx [B, n_per] int32 (ascending kept indices), sel_count [B] int32,
 *    probs_out [B, n_per] fp32 (optional, the softmax the selection was made from). */
size_t slime_router_workspace_bytes(const slime_ctx* ctx, int batch, int n_per, int prompt_len);
int slime_router_fwd(slime_ctx* ctx, const void* local, int n_per, const int32_t* n_valid,
                     const int64_t* ids, const uint8_t* mask, int batch, int prompt_len, float* probs_out,
                     int32_t* sel_idx, int32_t* sel_count, void* ws, size_t ws_bytes, void* stream);
/* Same, with the prompt given as a dense [B, T, hidden] embedding tensor (the signature of the reference's
 * TextGuidedSampler.forward(local_f, text_embedding, attn_mask), multimodal_resampler/builder.py:248). */
int slime_router_fwd_embeds(slime_ctx* ctx, const void* local, int n_per, const int32_t* n_valid,
                            const void* text_embeds, const uint8_t* mask, int batch, int prompt_len,
                            float* probs_out, int32_t* sel_idx, int32_t* sel_count, void* ws, size_t ws_bytes,
                            void* stream);
/* Selection only, from given probabilities (bit-exact index parity test entry). */
int slime_router_select(slime_ctx* ctx, const float* probs, int batch, int n_per, const int32_t* n_valid,
                        int32_t* sel_idx, int32_t* sel_count, void* stream);

/* ---- stage 5: the splice (llava_arch.py:249-255,361-459).
 * plan: computes per-sample lengths and cu_seqlens, copies cu_seqlens[B+1] to host_cu (SYNCHRONISES).
 *   plan_buf: device int32 scratch of slime_splice_plan_ints(B,T) ints, reused by gather/pad.
 * gather: writes the packed rows out_embeds [cu[B], hidden] and pos_ids [cu[B]] (int32). */
size_t slime_splice_plan_ints(int batch, int prompt_len);
int slime_splice_plan(slime_ctx* ctx, const int64_t* ids, const uint8_t* mask, int batch, int prompt_len,
                      int n_global, int has_sep, const int32_t* sel_count, int32_t* plan_buf,
                      int32_t* host_cu, void* stream);
/* No-host-sync variant (the syncs it replaces are llava_arch.py:170,378): only enqueues, the spliced lengths
 * stay on the device (cu_seqlens inside plan_buf).  The caller sizes every later buffer by the UPPER BOUND
 * batch * ((prompt_len - 1) + n_global + has_sep + local rows per sample) and passes that bound as `total_rows` to
 * slime_splice_gather (rows past the real total are zero-filled, pos_ids 0) and to slime_decoder_prefill_fwd (zero rows
 * stay zero through every layer and belong to no sequence).  With no host round trip the whole prefill can be captured
 * in a CUDA graph.  slime_splice_check reads back the "more than one image placeholder" flag of the last plan
 * (synchronises the stream; optional). */
int slime_splice_plan_async(slime_ctx* ctx, const int64_t* ids, const uint8_t* mask, int batch, int prompt_len,
                            int n_global, int has_sep, const int32_t* sel_count, int32_t* plan_buf, void* stream);
int slime_splice_check(slime_ctx* ctx, void* stream);
int slime_splice_gather(slime_ctx* ctx, const int64_t* ids, int batch, int prompt_len,
                        const int32_t* plan_buf, const void* global_feats, int n_global,
                        int64_t global_sample_rows, const void* local_feats, int64_t local_sample_rows,
                        const int32_t* sel_idx, int sel_stride, int has_sep, void* out_embeds,
                        int32_t* pos_ids, int total_rows, void* stream);
/* Padded views exactly as prepare_inputs_labels_for_multimodal returns them (any output may be NULL). */
int slime_splice_pad(slime_ctx* ctx, const int32_t* plan_buf, const void* packed_embeds,
                     const int64_t* labels_in, int batch, int prompt_len, int lmax, void* out_embeds,
                     uint8_t* out_mask, int64_t* out_pos, int64_t* out_labels, void* stream);

/* ---- stage 6: LlamaForCausalLM.forward(inputs_embeds=...) prefill (HF llama/modeling_llama.py:355-507)
 * embeds [total, hidden] packed rows, cu_seqlens [B+1] int32 (device), pos_ids [total] int32
 * -> logits_last [B, vocab] fp32 (last real token of every sequence; may be NULL)
 *    logits_all  [total, vocab] bf16 (may be NULL)
 *    hidden_out  [total, hidden] final-norm'ed hidden states (may be NULL) */
size_t slime_decoder_workspace_bytes(const slime_ctx* ctx, int total_rows, int batch);
int slime_decoder_prefill_fwd(slime_ctx* ctx, const void* embeds, const int32_t* cu_seqlens,
                              const int32_t* pos_ids, int batch, int total_rows, int max_seqlen,
                              float* logits_last, void* logits_all, void* hidden_out, void* ws,
                              size_t ws_bytes, void* stream);

/* ---- KV cache + decode step: the step right after the prefill in generate() (llava_llama.py:139, early-out
 * llava_arch.py:279).  The cache is caller-owned bf16 [layers][2][batch][cache_len][kv_heads*head_dim].
 * While a cache is attached, slime_decoder_prefill_fwd also stores K (post-RoPE) and V of every real token into it.
 * decode: x [batch, hidden] = embeddings of the tokens to append, lens [batch] int32 (device) = tokens already cached
 * -> logits [batch, vocab] fp32; the caller increments lens afterwards. */
size_t slime_kv_cache_bytes(const slime_ctx* ctx, int batch, int cache_len);
int slime_decoder_set_kv_cache(slime_ctx* ctx, void* cache, int batch, int cache_len);
size_t slime_decoder_decode_workspace_bytes(const slime_ctx* ctx, int batch);
int slime_decoder_decode_fwd(slime_ctx* ctx, const void* x, const int32_t* lens, int batch, float* logits, void* ws,
                             size_t ws_bytes, void* stream);

/* ---- image pre-processing: process_images / process_anyres_image + CLIPImageProcessor.preprocess
 * (llava/mm_utils.py:99-153,177-210,231-259), bit-exact with Pillow's bicubic `Image.resize`.
 * One job = one resize of one image placed on a canvas that is then cut into crop x crop tiles:
 *   source  : RGB bytes [src_h, src_w, 3] at src + src_offset, seen through a padded "virtual" source of
 *             virt_w x virt_h with the image at (virt_x, virt_y) and `fill` elsewhere (expand2square; no padding:
 *             virt = src, offsets 0)
 *   resize  : virtual source -> out_w x out_h (antialiased bicubic, 8-bit intermediate, as PIL)
 *   canvas  : canvas_w x canvas_h (multiples of crop), black, resized image pasted at (paste_x, paste_y) - negative
 *             offsets centre-crop; its tiles, row-major, are crops first_crop, first_crop+1, ... of `out`
 *   out     : [n_crops, 3, crop, crop] in out_dtype (0 bf16, 1 fp32, 2 fp16), value = lut[c][byte]
 * `jobs` and `lut` ([3][256] floats: the processor's rescale + normalise of every byte value) are HOST arrays;
 * src / out / ws are device pointers.  Needs no slime_ctx (no weights).
 * Alignment: src 4 bytes, every src_offset a multiple of 4 and the buffer readable up to the next multiple of 4 past
 * each image (the kernels fetch the interleaved bytes as aligned 32-bit words); out 16 bytes; ws 256 bytes;
 * crop a multiple of 4. */
typedef struct slime_resize_job {
  int64_t src_offset;
  int32_t src_w, src_h;
  int32_t virt_w, virt_h, virt_x, virt_y;
  int32_t out_w, out_h;
  int32_t canvas_w, canvas_h;
  int32_t paste_x, paste_y;
  int32_t first_crop;
  uint8_t fill[4];
} slime_resize_job;
size_t slime_preprocess_workspace_bytes(const slime_resize_job* jobs, int n_jobs);
int slime_preprocess_fwd(const uint8_t* src, const slime_resize_job* jobs, int n_jobs, int crop, const float* lut,
                         void* out, int out_dtype, void* ws, size_t ws_bytes, void* stream);

/* ---- single-op entry points (unit parity tests of the kernels through the same ABI) ---- */
int slime_op_gemm(const void* a, int lda, const void* w, int ldw, int m, int n, int k, const void* bias,
                  const void* residual, int res_ld, int res_period, const int32_t* row_map, int epilogue,
                  void* out, float* out_f32, int out_ld, void* stream);
/* Weight-streaming GEMM of the decode step (M <= 32 rows, K % 32 == 0; csrc/gemm_skinny.cu) with an explicit number
 * of k-splits: splits > 1 (or norm_w != NULL) needs ws with splits * m * n floats.  norm_w / norm_out: RMSNorm of the
 * output rows fused into the split-K finishing kernel (epilogue 0, bf16 out).  rope_*: epilogue 4 only. */
int slime_op_gemm_skinny(const void* a, int lda, const void* w, int ldw, int m, int n, int k, const void* bias,
                         const void* residual, int res_ld, int epilogue, void* out, float* out_f32, int out_ld,
                         int splits, float* ws, size_t ws_floats, const void* norm_w, void* norm_out, float norm_eps,
                         const int32_t* rope_pos, const float* rope_table, int rope_half, int rope_cols, int rope_max_pos,
                         void* stream);
/* Same with the fused decode chain's options: tile_counters ([n / 8] ints, zero on entry, zero again on exit) finishes
 * the split-K sum inside the kernel; a_norm_w ([k]) applies RMSNorm(a_norm_eps) to the rows of a while they are staged. */
int slime_op_gemm_skinny_fused(const void* a, int lda, const void* w, int ldw, int m, int n, int k, const void* bias,
                               const void* residual, int res_ld, int epilogue, void* out, float* out_f32, int out_ld,
                               int splits, float* ws, size_t ws_floats, const void* norm_w, void* norm_out, float norm_eps,
                               const int32_t* rope_pos, const float* rope_table, int rope_half, int rope_cols,
                               int rope_max_pos, int32_t* tile_counters, const void* a_norm_w, float a_norm_eps,
                               void* stream);
/* Single-query attention over a KV cache [batch, cache_len, kv_heads * head_dim]: sequence b attends its first
 * lens[b] + 1 positions.  splits >= 1: split-KV kernel (ws = batch * heads * splits * (head_dim + 2) floats when
 * splits > 1); splits == 0: one CTA per (q head, sequence). */
int slime_op_decode_attention(const void* q, int q_ld, const void* kcache, const void* vcache, int cache_len,
                              const int32_t* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                              void* out, int out_ld, int splits, float* ws, void* stream);
/* merge_counters ([batch * kv_heads] ints, zero on entry and exit): the kv splits are merged inside the kernel. */
int slime_op_decode_attention_fused(const void* q, int q_ld, const void* kcache, const void* vcache, int cache_len,
                                    const int32_t* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                                    void* out, int out_ld, int splits, float* ws, int32_t* merge_counters, void* stream);
int slime_op_attention(const void* q, const void* k, const void* v, void* o, int q_ld, int k_ld, int v_ld,
                       int o_ld, const int32_t* cu_q, const int32_t* cu_k, int seqlen_q, int seqlen_k,
                       int64_t q_batch_rows, int64_t k_batch_rows, int64_t o_batch_rows, int batch,
                       int heads, int kv_heads, int head_dim, float scale, int causal, int64_t total_q_rows,
                       int64_t total_k_rows, int impl /* 0 default, 1 mma.sync kernel, 2 tcgen05 kernel */,
                       void* stream);
int slime_op_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int dim, float eps,
                       void* stream);
int slime_op_rmsnorm(const void* x, const void* w, void* y, int rows, int dim, float eps, void* stream);
int slime_op_rope(slime_ctx* ctx, void* qkv, int ld, int rows, const int32_t* pos_ids, void* stream);
/* Packed Llama QKV projection + RoPE of the q / k heads exactly as the decoder stage runs it: x [rows, hidden] times
 * qkv_w [(heads + 2 kv_heads) * head_dim, hidden]^T -> qkv [rows, (heads + 2 kv_heads) * head_dim].  With
 * SLIME_FLAG_ROPE_INTERLEAVED in the ctx's desc, qkv_w must hold the interleaved q / k rows and the q / k columns of the
 * result come out in that interleaved feature order (rotation fused into the GEMM epilogue); without it the plain
 * layout is used (GEMM, then the in-place rotation pass). */
int slime_op_qkv_rope(slime_ctx* ctx, const void* x, const void* qkv_w, int rows, const int32_t* pos_ids, void* qkv,
                      void* stream);

/* ---- accounting: kernels launched by the library since load; optional CUDA-event profiling of the
 * library's own launches (class 0 = tcgen05 GEMM [work = FLOPs], 1 = attention, 2 = other) ---- */
/* GEMM kernel selection: 0 = 1-CTA kernel only, 1 = force the 2-CTA (cta_group::2) kernel, 2 = 2-CTA for problems
 * that fill the GPU, -1 = back to the default (SLIME_GEMM_2CTA environment variable / build default). */
int slime_gemm_set_2cta_mode(int mode);
/* GEMM epilogue HBM access pattern: 0 = every thread stores its own row, 1 = chunks transposed through shared memory
 * (stores and residual loads of 8 rows x 64 contiguous bytes per instruction).  Never called: SLIME_GEMM_EPI_MODE
 * environment variable, else the build default. */
int slime_gemm_set_epi_mode(int mode);
/* Decode-step kernel selection (A/B measurements): weight-streaming GEMM for M <= 32 rows (1, default) or the tcgen05
 * tile kernels (0); split-KV decode attention (1, default) or one CTA per (q head, sequence) (0); -1 = back to the
 * SLIME_GEMM_SKINNY / SLIME_DECODE_ATTN environment variables. */
int slime_gemm_set_skinny_mode(int mode);
int slime_decode_attention_set_mode(int mode);
/* Programmatic dependent launch of the decode-step kernels (their weight prefetch / prologue overlaps the previous
 * kernel's tail): 1 on (default), 0 ordinary stream-ordered launches, -1 = back to the SLIME_PDL environment variable. */
int slime_set_pdl_mode(int mode);
/* L2 prefetch duties of the decode step's latency-bound kernels (bit mask; csrc/api.cu decode_body): 1 = the QKV
 * finishing kernel pulls the o-projection's weights, 2 = the down-projection's finishing kernel pulls the next layer's
 * QKV weights, 4 / 8 = the attention kernel / the o-projection's finishing kernel pull 32 MB of gate/up each;
 * -1 = back to SLIME_DECODE_PREFETCH / the build default (0: measured slower than no prefetch in every variant). */
int slime_set_decode_prefetch(int mask);
/* Column width of the cta_group::2 GEMM's cluster tiles: 0 (default) = 256 unless 192-column tiles finish the problem in
 * fewer width-weighted waves (small problems only, e.g. the N = 4096 projections of a batch-1 prefill), 256 / 192 = forced
 * (tests, A/B), -1 = back to the default / SLIME_GEMM2_BN.  Results do not depend on it (same k order per element). */
int slime_gemm_set_tile_n(int bn);
/* Programmatic dependent launch for the prefill chain (tcgen05 GEMMs, attention, norm kernels): 1 (default) = the next kernel's
 * prologue overlaps the tail of the previous one (each such kernel waits for its predecessor before its first global
 * access), 0 = ordinary launches, -1 = back to the default / SLIME_PREFILL_PDL. */
int slime_set_prefill_pdl(int mode);
/* Decode-step launch chain, bit mask: 1 = split-K partial sums finished inside the projection kernels, 2 = attention kv
 * splits merged inside the attention kernel (both by atomic ticket, fixed summation order), 4 = RMSNorm applied while the
 * consuming projection stages its rows instead of in the o-/down-projection's finishing kernel.  7 = 5 launches per layer,
 * 0 (the default: faster under programmatic dependent launch, profiles/r02_decode_experiments.txt) = 9, one finishing
 * launch per split reduction; -1 = back to the default / SLIME_DECODE_FUSED. */
int slime_set_decode_fused(int mask);
/* ---- multi-GPU: the path's ONE collective - samples are independent, every rank runs the whole prefill on its
 * shard of the batch, and the last-token logits are all-gathered over NCCL (NVLink 5 / NVSwitch).  The reference has no
 * collective (N independent processes, outputs concatenated: scripts/llama/eval/gqa.sh:20-43).  NCCL is bound at run time
 * (dlopen libnccl.so.2).  Bootstrap: rank 0 -> slime_comm_unique_id -> the 128 bytes travel by any side channel -> every rank
 * slime_comm_init on its own device.  slime_allgather_logits only enqueues on `stream` (use a side stream: the next
 * prefill does not depend on it): out [world * rows_local, vocab] fp32 in rank order. */
typedef struct slime_comm slime_comm;
int slime_comm_unique_id(void* out_128_bytes);
int slime_comm_init(slime_comm** out, const void* id_128_bytes, int rank, int world);
int slime_comm_nccl_version(void);
int slime_allgather_logits(slime_comm* comm, const float* local_logits, float* out, int rows_local, int vocab, void* stream);
void slime_comm_destroy(slime_comm* comm);
/* debug: CTA 0 of the tcgen05 attention kernel stamps clock64() of its first 64 tiles into buf [64][16] (NULL = off) */
int slime_attention_set_trace(long long* buf);
/* two-query-tile attention kernel (csrc/attention_tc2.cu): how many of every 8 score-column pairs are exponentiated by a
 * polynomial on the FMA pipe instead of MUFU.EX2 (0, 2, 3 or 4; -1 = back to SLIME_ATTN_POLY / the build default). */
int slime_attention_set_poly(int pairs_of_8);
long long slime_launch_count(void);
int slime_profile_enable(int on);
int slime_profile_collect(double* ms3, double* work3, long long* launches3);

#ifdef __cplusplus
}
#endif
#endif /* SLIME_B200_H_ */
