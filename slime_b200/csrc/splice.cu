// Token splice (reference llava/model/llava_arch.py:361-459 prepare_inputs_labels_for_multimodal and
// :249-255 encode_images concat):  per sample
//     [text before <image>] + [576 global] + [separator] + [K selected local, ascending] + [text after]
// Integer/index work - bit-exact with the reference by construction.  Instead of the reference's
// right-padded [B, Lmax, H] tensor the decoder consumes the PACKED rows [sum L_i, H] + cu_seqlens;
// the padded views (inputs_embeds / attention_mask / position_ids / labels) are produced on demand by
// splice_pad_* for API parity.
//
//   splice_plan   : strips masked prompt slots, locates the image placeholder, computes L_i and the
//                   exclusive prefix cu_seqlens (one CTA; B is small)
//   splice_gather : one warp per output row: embedding-table gather for text rows, row copies for the
//                   global / separator / selected-local rows; also writes per-row position ids
#include "errors.h"
#include "splice.h"

namespace {

// plan[b] layout (ints): 0 n_text, 1 img_pos (text tokens before the image), 2 n_img, 3 img_len, 4 L
constexpr int PLAN_STRIDE = 8;

__global__ void __launch_bounds__(1024) splice_plan_kernel(
    const long long* __restrict__ ids, const unsigned char* __restrict__ mask, int B, int T,
    long long image_token, int n_global, int has_sep, const int* __restrict__ sel_count, int max_len,
    int* __restrict__ valid_pos, int* __restrict__ plan, int* __restrict__ cu_seqlens,
    int* __restrict__ err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += 32) {
    const long long* idr = ids + static_cast<long long>(b) * T;
    const unsigned char* mr = mask != nullptr ? mask + static_cast<long long>(b) * T : nullptr;
    int* vp = valid_pos + static_cast<long long>(b) * T;
    int n_text = 0, n_img = 0, img_pos = -1;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const bool valid = t < T && (mr == nullptr || mr[t] != 0);
      const bool is_img = valid && idr[t] == image_token;
      const bool is_text = valid && !is_img;
      const unsigned tm = __ballot_sync(0xffffffffu, is_text);
      const unsigned im = __ballot_sync(0xffffffffu, is_img);
      if (is_text) vp[n_text + __popc(tm & ((1u << lane) - 1))] = t;
      if (im != 0 && img_pos < 0) {
        const int first = __ffs(im) - 1;
        img_pos = n_text + __popc(tm & ((1u << first) - 1));
      }
      n_text += __popc(tm);
      n_img += __popc(im);
    }
    if (lane == 0) {
      int img_len = 0;
      if (n_img > 0) img_len = n_global + (has_sep ? 1 : 0) + (sel_count != nullptr ? sel_count[b] : 0);
      int L = n_text + img_len;
      if (max_len > 0 && L > max_len) L = max_len;
      int* pl = plan + b * PLAN_STRIDE;
      pl[0] = n_text;
      pl[1] = img_pos < 0 ? n_text : img_pos;
      pl[2] = n_img;
      pl[3] = img_len;
      pl[4] = L;
      if (n_img > 1) atomicExch(err_flag, 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    cu_seqlens[0] = 0;
    for (int b = 0; b < B; ++b) {
      acc += plan[b * PLAN_STRIDE + 4];
      cu_seqlens[b + 1] = acc;
    }
  }
}

SLIME_DEVINL int find_sample(const int* cu, int B, int row) {
  int lo = 0, hi = B;  // cu[lo] <= row < cu[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= row)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128) splice_gather_kernel(
    const long long* __restrict__ ids, int T, const int* __restrict__ valid_pos,
    const int* __restrict__ plan, const int* __restrict__ cu_seqlens, int B,
    const bf16* __restrict__ embed, int H, long long sep_token, int has_sep,
    const bf16* __restrict__ glob, int n_global, long long glob_sample_rows,
    const bf16* __restrict__ local, long long local_sample_rows, const int* __restrict__ sel_idx,
    int sel_stride, bf16* __restrict__ out, int* __restrict__ pos_ids, int total_rows) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total_rows) return;
  if (row >= cu_seqlens[B]) {
    // no-host-sync mode: the caller sized the buffers by an upper bound; rows past the real total are zero (they belong
    // to no sequence and stay zero through the decoder)
    bf16* z = out + static_cast<long long>(row) * H;
    for (int c = lane; c < (H >> 3); c += 32) *reinterpret_cast<uint4*>(z + c * 8) = make_uint4(0, 0, 0, 0);
    if (lane == 0 && pos_ids != nullptr) pos_ids[row] = 0;
    return;
  }
  const int b = find_sample(cu_seqlens, B, row);
  const int j = row - cu_seqlens[b];
  const int* pl = plan + b * PLAN_STRIDE;
  const int img_pos = pl[1], img_len = pl[3];
  const bf16* src;
  if (j < img_pos) {
    src = embed + ids[static_cast<long long>(b) * T + valid_pos[static_cast<long long>(b) * T + j]] * H;
  } else if (j < img_pos + img_len) {
    const int q = j - img_pos;
    if (q < n_global) {
      src = glob + (b * glob_sample_rows + q) * H;
    } else if (has_sep && q == n_global) {
      src = embed + sep_token * H;
    } else {
      const int s = q - n_global - (has_sep ? 1 : 0);
      src = local + (b * local_sample_rows + sel_idx[static_cast<long long>(b) * sel_stride + s]) * H;
    }
  } else {
    const int tj = j - img_len;  // index into the compacted text tokens
    src = embed + ids[static_cast<long long>(b) * T + valid_pos[static_cast<long long>(b) * T + tj]] * H;
  }
  bf16* dst = out + static_cast<long long>(row) * H;
  for (int c = lane; c < (H >> 3); c += 32) {
    *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(src + c * 8);
  }
  if (lane == 0 && pos_ids != nullptr) pos_ids[row] = j;
}

// padded [B, Lmax] metadata exactly as the reference returns it (llava_arch.py:415-459)
__global__ void splice_pad_meta_kernel(const int* __restrict__ plan, const int* __restrict__ valid_pos,
                                       const long long* __restrict__ labels_in, int T, int B, int Lmax,
                                       int left_pad, long long ignore_index,
                                       unsigned char* __restrict__ out_mask,
                                       long long* __restrict__ out_pos, long long* __restrict__ out_labels) {
  const long long total = static_cast<long long>(B) * Lmax;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / Lmax);
    const int s = static_cast<int>(i % Lmax);
    const int* pl = plan + b * PLAN_STRIDE;
    const int L = pl[4], img_pos = pl[1], img_len = pl[3];
    const int j = left_pad ? s - (Lmax - L) : s;
    const bool real = j >= 0 && j < L;
    if (out_mask != nullptr) out_mask[i] = real ? 1 : 0;
    if (out_pos != nullptr) out_pos[i] = real ? j : 0;
    if (out_labels != nullptr) {
      long long lab = ignore_index;
      if (real && labels_in != nullptr) {
        if (j < img_pos) {
          lab = labels_in[static_cast<long long>(b) * T + valid_pos[static_cast<long long>(b) * T + j]];
        } else if (j >= img_pos + img_len) {
          lab = labels_in[static_cast<long long>(b) * T + valid_pos[static_cast<long long>(b) * T + (j - img_len)]];
        }
      }
      out_labels[i] = lab;
    }
  }
}

// packed rows -> zero-padded [B, Lmax, H]
__global__ void __launch_bounds__(128) splice_pad_embeds_kernel(const bf16* __restrict__ packed,
                                                                const int* __restrict__ cu_seqlens, int B,
                                                                int Lmax, int H, int left_pad,
                                                                bf16* __restrict__ out) {
  const long long prow = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (prow >= static_cast<long long>(B) * Lmax) return;
  const int b = static_cast<int>(prow / Lmax);
  const int s = static_cast<int>(prow % Lmax);
  const int L = cu_seqlens[b + 1] - cu_seqlens[b];
  const int j = left_pad ? s - (Lmax - L) : s;
  bf16* dst = out + prow * H;
  if (j >= 0 && j < L) {
    const bf16* src = packed + static_cast<long long>(cu_seqlens[b] + j) * H;
    for (int c = lane; c < (H >> 3); c += 32)
      *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(src + c * 8);
  } else {
    for (int c = lane; c < (H >> 3); c += 32) *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0, 0, 0, 0);
  }
}

// rows[b] = cu_seqlens[b+1] - 1  (last real token of each sequence)
__global__ void last_rows_kernel(const int* __restrict__ cu_seqlens, int B, int* __restrict__ rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) rows[b] = max(cu_seqlens[b + 1] - 1, cu_seqlens[b]);
}

}  // namespace

int slime_launch_splice_plan(const long long* ids, const unsigned char* mask, int B, int T,
                             long long image_token, int n_global, int has_sep, const int* sel_count,
                             int max_len, int* valid_pos, int* plan, int* cu_seqlens, int* err_flag,
                             cudaStream_t stream) {
  SLIME_REQUIRE(B > 0 && T > 0, "splice: empty batch");
  splice_plan_kernel<<<1, 1024, 0, stream>>>(ids, mask, B, T, image_token, n_global, has_sep, sel_count,
                                             max_len, valid_pos, plan, cu_seqlens, err_flag);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_splice_gather(const long long* ids, int T, const int* valid_pos, const int* plan,
                               const int* cu_seqlens, int B, const bf16* embed, int H, long long sep_token,
                               int has_sep, const bf16* glob, int n_global, long long glob_sample_rows,
                               const bf16* local, long long local_sample_rows, const int* sel_idx,
                               int sel_stride, bf16* out, int* pos_ids, int total_rows,
                               cudaStream_t stream) {
  SLIME_REQUIRE(H % 8 == 0, "splice: hidden size %d must be a multiple of 8", H);
  if (total_rows <= 0) return SLIME_OK;
  splice_gather_kernel<<<(total_rows + 3) / 4, 128, 0, stream>>>(
      ids, T, valid_pos, plan, cu_seqlens, B, embed, H, sep_token, has_sep, glob, n_global,
      glob_sample_rows, local, local_sample_rows, sel_idx, sel_stride, out, pos_ids, total_rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_splice_pad_meta(const int* plan, const int* valid_pos, const long long* labels_in, int T,
                                 int B, int Lmax, int left_pad, long long ignore_index,
                                 unsigned char* out_mask, long long* out_pos, long long* out_labels,
                                 cudaStream_t stream) {
  if (B <= 0 || Lmax <= 0) return SLIME_OK;
  const long long total = static_cast<long long>(B) * Lmax;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  splice_pad_meta_kernel<<<grid, 256, 0, stream>>>(plan, valid_pos, labels_in, T, B, Lmax, left_pad,
                                                   ignore_index, out_mask, out_pos, out_labels);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_splice_pad_embeds(const bf16* packed, const int* cu_seqlens, int B, int Lmax, int H,
                                   int left_pad, bf16* out, cudaStream_t stream) {
  if (B <= 0 || Lmax <= 0) return SLIME_OK;
  const long long rows = static_cast<long long>(B) * Lmax;
  splice_pad_embeds_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, stream>>>(packed, cu_seqlens, B,
                                                                                      Lmax, H, left_pad, out);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_last_rows(const int* cu_seqlens, int B, int* rows, cudaStream_t stream) {
  if (B <= 0) return SLIME_OK;
  last_rows_kernel<<<(B + 127) / 128, 128, 0, stream>>>(cu_seqlens, B, rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}
