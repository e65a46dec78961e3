// Host-side launchers of the text-guided router kernels (router.cu).
#pragma once
#include "common.cuh"

// tvec[b,:] = sum over kept prompt tokens of E[id] / max(|E[id]|, 1e-8); inv_norm_ws is [B*T] scratch.
int slime_launch_text_dir(const long long* ids, const unsigned char* mask, const bf16* embed,
                          float* inv_norm_ws, float* tvec, int B, int T, int H, long long image_token,
                          int vocab, cudaStream_t stream);
// score[i] = <x_i, tvec[i / rows_per_sample]> / max(|x_i|, 1e-8)
int slime_launch_router_score(const bf16* x, const float* tvec, float* score, int rows,
                              int rows_per_sample, int H, cudaStream_t stream);
// softmax(score / temp) (skipped when from_probs) -> top-p selection; see router.cu for the exact rule.
int slime_launch_router_select(const float* in, int B, int n_per, const int* n_valid, float temp,
                               float top_p, int from_probs, float* probs_out, int* sel_idx,
                               int* sel_count, cudaStream_t stream);

// ---- 'qformer' router (TextGuidedRouterAttention) ----
// Packs the kept prompt tokens (mask != 0 and a real token id) of every sample into contiguous rows of `packed`
// (E[ids] rows when ids != nullptr, else rows of the dense [B*T, H] tensor `src`): dst_row [B*T] (scratch), cu [B+1].
int slime_launch_qf_pack_text(const long long* ids, const unsigned char* mask, const bf16* src, int B, int T, int H,
                              long long image_token, int vocab, int* dst_row, int* cu, bf16* packed,
                              cudaStream_t stream);
// logit[i] = <relu(h[i, :Dh]), w2> + b2, then in place per sample: softmax(logit / temp) over its first n_valid rows
int slime_launch_qf_logits(const bf16* h, const bf16* w2, const bf16* b2, float* logit, int rows, int Dh, int B,
                           int n_per, const int* n_valid, float temp, cudaStream_t stream);
