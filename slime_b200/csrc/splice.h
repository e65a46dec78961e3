// Host-side launchers of the token-splice kernels (splice.cu).
#pragma once
#include "common.cuh"

#define SLIME_PLAN_STRIDE 8  // ints per sample in the plan buffer: n_text, img_pos, n_img, img_len, L

int slime_launch_splice_plan(const long long* ids, const unsigned char* mask, int B, int T,
                             long long image_token, int n_global, int has_sep, const int* sel_count,
                             int max_len, int* valid_pos, int* plan, int* cu_seqlens, int* err_flag,
                             cudaStream_t stream);
int slime_launch_splice_gather(const long long* ids, int T, const int* valid_pos, const int* plan,
                               const int* cu_seqlens, int B, const bf16* embed, int H, long long sep_token,
                               int has_sep, const bf16* glob, int n_global, long long glob_sample_rows,
                               const bf16* local, long long local_sample_rows, const int* sel_idx,
                               int sel_stride, bf16* out, int* pos_ids, int total_rows,
                               cudaStream_t stream);
int slime_launch_splice_pad_meta(const int* plan, const int* valid_pos, const long long* labels_in, int T,
                                 int B, int Lmax, int left_pad, long long ignore_index,
                                 unsigned char* out_mask, long long* out_pos, long long* out_labels,
                                 cudaStream_t stream);
int slime_launch_splice_pad_embeds(const bf16* packed, const int* cu_seqlens, int B, int Lmax, int H,
                                   int left_pad, bf16* out, cudaStream_t stream);
int slime_launch_last_rows(const int* cu_seqlens, int B, int* rows, cudaStream_t stream);
