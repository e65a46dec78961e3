// Host-side interface of the fused attention kernels (attention_fa.cu).
#pragma once
#include "common.cuh"

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
  int impl = 0;                // 0 = default (tcgen05), 1 = mma.sync flash kernel, 2 = tcgen05
};

int slime_launch_attention(const AttnParams& p, cudaStream_t stream);
// tcgen05 / TMEM implementation (attention_tc.cu)
int slime_launch_attention_tc(const AttnParams& p, int num_sms, cudaStream_t stream);
