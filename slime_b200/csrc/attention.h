// Host-side interface of the fused attention kernels (attention_tc2.cu, decode_attn.cu).
#pragma once
#include "common.cuh"

// attention_tc2.cu: how many of every 8 score-column pairs are exponentiated by a polynomial on the FMA pipe instead of
// MUFU.EX2 (0, 2, 3 or 4), by head dim - the measured optimum on B200 (profiles/r02_attention_experiments.txt)
#ifndef SLIME_ATTN_POLY_HD128
#define SLIME_ATTN_POLY_HD128 3
#endif
#ifndef SLIME_ATTN_POLY_HD64
#define SLIME_ATTN_POLY_HD64 2
#endif

struct AttnParams {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  int q_ld, k_ld, v_ld, o_ld;  // row strides (elements); heads are contiguous slices of HD inside a row
  const int* cu_q;             // [B+1] packed (varlen) row offsets, or nullptr for fixed-length batches
  const int* cu_k;             // [B+1] or nullptr
  int seqlen_q, seqlen_k;      // fixed lengths (cu_* == nullptr) or upper bounds (varlen)
  long long q_batch_rows;      // rows between consecutive batches when cu_q == nullptr (0 = shared Q)
  long long k_batch_rows;      // same for K/V when cu_k == nullptr
  long long o_batch_rows;      // same for O when cu_q == nullptr
  int batch;
  int num_heads, num_kv_heads;
  int head_dim;  // 64 or 128
  float scale;   // softmax scale (1/sqrt(head_dim))
  int causal;    // 1: query i attends keys <= i + (seqlen_k - seqlen_q)
  long long total_q_rows = 0;  // rows of the q matrix (needed for the TMA map when cu_q != nullptr; 0 = derive)
  long long total_k_rows = 0;  // rows of the k / v matrices (same)
  long long* trace = nullptr;  // debug: CTA 0 writes clock64() stamps of its first 64 tiles here ([64][16])
};

int slime_launch_attention(const AttnParams& p, cudaStream_t stream);
// tcgen05 / TMEM implementation (attention_tc2.cu)
int slime_launch_attention_tc2(const AttnParams& p, int num_sms, cudaStream_t stream);

// ---- decode step (decode_attn.cu) ----
// q [batch, heads*head_dim] (row stride q_ld) against the cache k/v [batch, cache_len, kv_heads*head_dim];
// sequence b attends its first lens[b] + 1 cached positions.  splits >= 1 selects the split-KV kernel (one CTA per
// kv head x sequence x split; ws = slime_decode_attention_ws_floats() floats when splits > 1); splits == 0, or a
// shape the split kernel does not cover, runs the one-CTA-per-(q head, sequence) kernel.
int slime_launch_decode_attention(const bf16* q, int q_ld, const bf16* kcache, const bf16* vcache, int cache_len,
                                  const int* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                                  bf16* out, int out_ld, int splits, float* ws, const void* pf_ptr, size_t pf_bytes,
                                  cudaStream_t stream, int* merge_counters = nullptr);
// merge_counters: [batch * kv_heads] ints, zero on entry and on exit - the mma.sync split-KV kernel then merges the kv
// splits itself (last CTA of a (sequence, kv head) by atomic ticket, fixed split order) and no merge kernel is launched.
// (pf_ptr, pf_bytes): optional immutable region (a later projection's weights) the split-KV kernel pulls into L2
// kv splits that fill the GPU for this problem (0: the split kernel does not apply / is switched off)
int slime_decode_attention_splits(int batch, int heads, int kv_heads, int head_dim, int cache_len, int num_sms);
size_t slime_decode_attention_ws_floats(int batch, int heads, int splits);
// rows[i] = sample(i) * cache_len + pos_ids[i] for the packed prefill rows (-1 when pos is outside [0, cache_len) or the
// sample outside the cache's batch)
int slime_launch_cache_rows(const int* cu, const int* pos_ids, int B, int total, int cache_batch, int cache_len, int* rows,
                            cudaStream_t stream);
// cache[b, lens[b]] = this step's K / V rows (k, v: [B, KD] slices with row stride ld), both planes in one launch
int slime_launch_kv_append(const bf16* k, const bf16* v, int ld, bf16* kcache, bf16* vcache, int KD, const int* lens,
                           int B, int cache_len, cudaStream_t stream);
// rows[b] = b * cache_len + lens[b]
int slime_launch_append_rows(const int* lens, int B, int cache_len, int* rows, cudaStream_t stream);
