// Text-guided router (reference llava/model/multimodal_resampler/builder.py:177-201 cosine selector,
// :248-281 top-p selection) and the text-side reduction it needs.
//
// The reference materialises cos(x_i, e_t) for every (local token, text token) pair - an N x T x H
// broadcast - and then sums over t.  Algebraically
//     score_i = sum_t m_t * <x_i, e_t> / (max(|x_i|,eps) * max(|e_t|,eps))
//             = <x_i, tvec> / max(|x_i|, eps),      tvec = sum_t m_t * e_t / max(|e_t|, eps)
// so the router is one reduction over the prompt (text_inv_norm + text_dir) and one GEMV-like pass
// over the local tokens (router_score), followed by a per-sample single-CTA softmax / stable
// descending sort / prefix-sum / compaction (router_select) that reproduces the reference's
// selection rule exactly:  count = #(cumsum(sorted p) <= top_p);  keep the first count+1 sorted
// entries (all of them if count == N);  return their indices in ascending order.
#include "errors.h"
#include "router.h"

namespace {

constexpr float COS_EPS = 1e-8f;  // torch.nn.functional.cosine_similarity default eps

// inv_norm[b,t] = keep(b,t) / max(|E[id]|, eps);  keep = attention_mask != 0 and id is a real token
__global__ void text_inv_norm_kernel(const long long* __restrict__ ids, const unsigned char* __restrict__ mask,
                                     const bf16* __restrict__ embed, float* __restrict__ inv_norm, int B,
                                     int T, int H, long long image_token, int vocab) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T) return;
  // ids == nullptr: `embed` is a dense [B*T, H] tensor and row t is its own embedding
  const long long id = ids != nullptr ? ids[row] : row;
  const bool keep = (mask == nullptr || mask[row] != 0) &&
                    (ids == nullptr || (id != image_token && id >= 0 && id < vocab));
  float r = 0.f;
  if (keep) {
    const bf16* e = embed + id * H;
    float sq = 0.f;
    for (int c = lane; c < (H >> 3); c += 32) {
      const uint4 u = *reinterpret_cast<const uint4*>(e + c * 8);
      float2 t;
      t = unpack_bf16x2(u.x); sq += t.x * t.x + t.y * t.y;
      t = unpack_bf16x2(u.y); sq += t.x * t.x + t.y * t.y;
      t = unpack_bf16x2(u.z); sq += t.x * t.x + t.y * t.y;
      t = unpack_bf16x2(u.w); sq += t.x * t.x + t.y * t.y;
    }
    sq = warp_sum(sq);
    r = 1.0f / fmaxf(sqrtf(sq), COS_EPS);
  }
  if (lane == 0) inv_norm[row] = r;
}

// tvec[b, col] = sum_t inv_norm[b,t] * E[ids[b,t], col]     (thread per column pair, loop over t)
__global__ void text_dir_kernel(const long long* __restrict__ ids, const float* __restrict__ inv_norm,
                                const bf16* __restrict__ embed, float* __restrict__ tvec, int T, int H) {
  const int b = blockIdx.y;
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= H) return;
  float a0 = 0.f, a1 = 0.f;
  const long long* idr = ids + static_cast<long long>(b) * T;
  const float* nr = inv_norm + static_cast<long long>(b) * T;
  for (int t = 0; t < T; ++t) {
    const float w = nr[t];
    if (w != 0.f) {
      const long long er = ids != nullptr ? idr[t] : static_cast<long long>(b) * T + t;
      const float2 e = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(embed + er * H + col));
      a0 += w * e.x;
      a1 += w * e.y;
    }
  }
  tvec[static_cast<long long>(b) * H + col] = a0;
  tvec[static_cast<long long>(b) * H + col + 1] = a1;
}

// score[i] = <x_i, tvec[b]> / max(|x_i|, eps)   (warp per local token; rows_per_sample tokens each)
__global__ void router_score_kernel(const bf16* __restrict__ x, const float* __restrict__ tvec,
                                    float* __restrict__ score, int rows, int rows_per_sample, int H) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / rows_per_sample;
  const bf16* xr = x + static_cast<long long>(row) * H;
  const float* tv = tvec + static_cast<long long>(b) * H;
  float dot = 0.f, sq = 0.f;
  for (int c = lane; c < (H >> 3); c += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + c * 8);
    const float4 t0 = *reinterpret_cast<const float4*>(tv + c * 8);
    const float4 t1 = *reinterpret_cast<const float4*>(tv + c * 8 + 4);
    float2 v;
    v = unpack_bf16x2(u.x); dot += v.x * t0.x + v.y * t0.y; sq += v.x * v.x + v.y * v.y;
    v = unpack_bf16x2(u.y); dot += v.x * t0.z + v.y * t0.w; sq += v.x * v.x + v.y * v.y;
    v = unpack_bf16x2(u.z); dot += v.x * t1.x + v.y * t1.y; sq += v.x * v.x + v.y * v.y;
    v = unpack_bf16x2(u.w); dot += v.x * t1.z + v.y * t1.w; sq += v.x * v.x + v.y * v.y;
  }
  dot = warp_sum(dot);
  sq = warp_sum(sq);
  if (lane == 0) score[row] = dot / fmaxf(sqrtf(sq), COS_EPS);
}

// ------------------------------------------------------------------------------------------
// per-sample selection: softmax (optional) -> stable descending sort -> cumsum -> top-p -> compaction
// ------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX = 4096;  // max local tokens per sample (17 crops -> 16*144 = 2304)

SLIME_DEVINL bool before(float pa, int ia, float pb, int ib) {
  // descending by probability, ascending index among equal probabilities (stable order)
  return pa > pb || (pa == pb && ia < ib);
}

template <typename T>
SLIME_DEVINL T block_reduce(T v, T* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (other > v ? other : v) : v + other;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < (SEL_THREADS / 32) ? red[lane] : (is_max ? red[0] : T(0));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T other = __shfl_xor_sync(0xffffffffu, w, o);
      w = is_max ? (other > w ? other : w) : w + other;
    }
    if (lane == 0) red[0] = w;
  }
  __syncthreads();
  const T out = red[0];
  __syncthreads();
  return out;
}

// in: score or prob [B, n_stride]; sample b uses its first n_valid[b] entries (all n_stride when
// n_valid == nullptr).  from_probs: the input already holds probabilities, skip the softmax.
__global__ void __launch_bounds__(SEL_THREADS) router_select_kernel(
    const float* __restrict__ in, int n_stride, const int* __restrict__ n_valid, float temp, float top_p,
    int from_probs, float* __restrict__ probs_out, int* __restrict__ sel_idx, int* __restrict__ sel_count) {
  __shared__ float key[SEL_MAX];
  __shared__ int idx[SEL_MAX];
  __shared__ double scan[SEL_MAX / 4];  // per-thread partial sums (SEL_THREADS entries)
  __shared__ float redf[32];
  __shared__ int redi[32];
  __shared__ int s_count;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int n_per = n_valid != nullptr ? min(n_valid[b], n_stride) : n_stride;
  const float* src = in + static_cast<long long>(b) * n_stride;
  int* out_idx = sel_idx + static_cast<long long>(b) * n_stride;

  if (n_per <= 0) {
    if (tid == 0) sel_count[b] = 0;
    return;
  }

  // ---- probabilities ----
  if (!from_probs) {
    float mx = -INFINITY;
    for (int i = tid; i < n_per; i += SEL_THREADS) mx = fmaxf(mx, src[i] / temp);
    mx = block_reduce<float>(mx, redf, true);
    float sum = 0.f;
    for (int i = tid; i < n_per; i += SEL_THREADS) {
      const float e = expf(src[i] / temp - mx);
      key[i] = e;
      sum += e;
    }
    sum = block_reduce<float>(sum, redf, false);
    for (int i = tid; i < n_per; i += SEL_THREADS) key[i] = key[i] / sum;
  } else {
    for (int i = tid; i < n_per; i += SEL_THREADS) key[i] = src[i];
  }
  __syncthreads();
  if (probs_out != nullptr) {
    for (int i = tid; i < n_per; i += SEL_THREADS) probs_out[static_cast<long long>(b) * n_stride + i] = key[i];
  }
  int npow = 1;
  while (npow < n_per) npow <<= 1;
  for (int i = tid; i < npow; i += SEL_THREADS) {
    idx[i] = i;
    if (i >= n_per) key[i] = -1.0f;  // padding sorts last (probabilities are >= 0)
  }
  __syncthreads();

  // ---- bitonic sort, order: descending probability, ascending index on ties ----
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow; i += SEL_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const float pa = key[i], pb = key[l];
          const int ia = idx[i], ib = idx[l];
          const bool up = (i & k) == 0;  // this subsequence should be in "before" order
          const bool swap = up ? before(pb, ib, pa, ia) : before(pa, ia, pb, ib);
          if (swap) {
            key[i] = pb; key[l] = pa;
            idx[i] = ib; idx[l] = ia;
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- inclusive prefix sum in double (torch CPU cumsum accumulates float in double), then
  //      count entries whose float(cumsum) <= top_p ----
  const int per = (n_per + SEL_THREADS - 1) / SEL_THREADS;  // <= 4
  const int lo = tid * per;
  double local = 0.0;
  for (int i = 0; i < per; ++i) {
    const int p = lo + i;
    if (p < n_per) local += static_cast<double>(key[p]);
  }
  scan[tid] = local;
  __syncthreads();
  // Hillis-Steele inclusive scan over SEL_THREADS partials
  for (int off = 1; off < SEL_THREADS; off <<= 1) {
    double add = 0.0;
    if (tid >= off) add = scan[tid - off];
    __syncthreads();
    scan[tid] += add;
    __syncthreads();
  }
  double run = tid > 0 ? scan[tid - 1] : 0.0;
  int cnt = 0;
  for (int i = 0; i < per; ++i) {
    const int p = lo + i;
    if (p < n_per) {
      run += static_cast<double>(key[p]);
      if (static_cast<float>(run) <= top_p) cnt += 1;
    }
  }
  cnt = block_reduce<int>(cnt, redi, false);
  if (tid == 0) s_count = cnt < n_per ? cnt + 1 : n_per;
  __syncthreads();
  const int keep = s_count;

  // ---- mark the kept original indices, compact them in ascending order ----
  __syncthreads();
  for (int i = tid; i < n_per; i += SEL_THREADS) key[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < keep; i += SEL_THREADS) key[idx[i]] = 1.f;
  __syncthreads();
  int mine = 0;
  for (int i = 0; i < per; ++i) {
    const int p = lo + i;
    if (p < n_per && key[p] != 0.f) mine += 1;
  }
  int* iscan = reinterpret_cast<int*>(scan);
  iscan[tid] = mine;
  __syncthreads();
  for (int off = 1; off < SEL_THREADS; off <<= 1) {
    int add = 0;
    if (tid >= off) add = iscan[tid - off];
    __syncthreads();
    iscan[tid] += add;
    __syncthreads();
  }
  int pos = tid > 0 ? iscan[tid - 1] : 0;
  for (int i = 0; i < per; ++i) {
    const int p = lo + i;
    if (p < n_per && key[p] != 0.f) out_idx[pos++] = p;
  }
  if (tid == 0) sel_count[b] = keep;
}

// ------------------------------------------------------------------------------------------
// 'qformer' router (reference multimodal_resampler/builder.py:94-162, TextGuidedRouterAttention): the local tokens
// cross-attend the prompt, an MLP turns every attended token into one logit, softmax over the tokens.  The dense
// parts (LayerNorms, in/out projections, 128-wide heads, first MLP layer) run on the library's GEMM / attention /
// LayerNorm kernels (api.cu router_qformer_body); what is specific to this router lives here:
//   * packing the kept prompt tokens of every sample into contiguous rows (the attention kernel takes key ranges,
//     key_padding_mask = ~mask in the reference) - the cross-attention has no positional term, so only the SET of
//     kept tokens matters and the placeholder-removal / zero-padding of get_pure_text_embedding reduces to "keep";
//   * logit_i = <relu(h_i), w2> + b2 and the per-sample softmax.
// ------------------------------------------------------------------------------------------
// One CTA; warp w handles samples w, w + nwarps, ...: count the kept tokens, then (after the prefix sum over the samples)
// assign every kept token its packed row.  dst_row[b*T + t] = packed row or -1;  cu[b] = first packed row of sample b.
__global__ void __launch_bounds__(1024) qf_text_rows_kernel(const long long* __restrict__ ids,
                                                            const unsigned char* __restrict__ mask, int B, int T,
                                                            long long image_token, int vocab, int* __restrict__ dst_row,
                                                            int* __restrict__ cu) {
  extern __shared__ int qf_cnt[];  // [B + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  auto keep = [&](int b, int t) {
    const long long i = static_cast<long long>(b) * T + t;
    if (mask != nullptr && mask[i] == 0) return false;
    if (ids == nullptr) return true;
    const long long id = ids[i];
    return id != image_token && id >= 0 && id < vocab;
  };
  for (int b = warp; b < B; b += nw) {
    int n = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      n += __popc(__ballot_sync(0xffffffffu, t < T && keep(b, t)));
    }
    if (lane == 0) qf_cnt[b + 1] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    qf_cnt[0] = 0;
    for (int b = 0; b < B; ++b) qf_cnt[b + 1] += qf_cnt[b];
    for (int b = 0; b <= B; ++b) cu[b] = qf_cnt[b];
  }
  __syncthreads();
  for (int b = warp; b < B; b += nw) {
    int base = qf_cnt[b];
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const bool k = t < T && keep(b, t);
      const unsigned bal = __ballot_sync(0xffffffffu, k);
      if (t < T) dst_row[static_cast<long long>(b) * T + t] = k ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
      base += __popc(bal);
    }
  }
}

// packed[dst_row[i]] = E[ids[i]] (ids != nullptr) or text[i]; 16 bytes per thread
__global__ void qf_gather_text_kernel(const long long* __restrict__ ids, const bf16* __restrict__ src,
                                      const int* __restrict__ dst_row, bf16* __restrict__ packed, int rows, int H) {
  const int chunks = H >> 3;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(rows) * chunks) return;
  const int r = static_cast<int>(i / chunks), ch = static_cast<int>(i % chunks);
  const int d = dst_row[r];
  if (d < 0) return;
  const long long srow = ids != nullptr ? ids[r] : r;
  *reinterpret_cast<uint4*>(packed + static_cast<long long>(d) * H + ch * 8) =
      *reinterpret_cast<const uint4*>(src + srow * H + ch * 8);
}

// logit[i] = <relu(h[i, :]), w2> + b2   (warp per row; prob_proj = Linear, ReLU, Linear)
__global__ void qf_logit_kernel(const bf16* __restrict__ h, const bf16* __restrict__ w2, const bf16* __restrict__ b2,
                                float* __restrict__ logit, int rows, int Dh) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int c = lane; c < Dh; c += 32)
    acc += fmaxf(elem_to_float(h[static_cast<long long>(row) * Dh + c]), 0.f) * elem_to_float(w2[c]);
  acc = warp_sum(acc);
  if (lane == 0) logit[row] = acc + elem_to_float(b2[0]);
}

// in place: x[b, :n] = softmax(x[b, :n] / temp)   (one CTA per sample; n = n_valid[b] or n_per)
__global__ void __launch_bounds__(256) qf_softmax_kernel(float* __restrict__ x, int n_per, const int* __restrict__ n_valid,
                                                         float temp) {
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = n_valid != nullptr ? min(n_valid[b], n_per) : n_per;
  float* xb = x + static_cast<long long>(b) * n_per;
  const float inv_t = 1.0f / temp;
  float m = -INFINITY;
  for (int i = tid; i < n; i += 256) m = fmaxf(m, xb[i] * inv_t);
  m = warp_max(m);
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < n; i += 256) sum += expf(xb[i] * inv_t - m);
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  for (int i = tid; i < n; i += 256) xb[i] = expf(xb[i] * inv_t - m) / sum;
}

}  // namespace

int slime_launch_qf_pack_text(const long long* ids, const unsigned char* mask, const bf16* src, int B, int T, int H,
                              long long image_token, int vocab, int* dst_row, int* cu, bf16* packed,
                              cudaStream_t stream) {
  SLIME_REQUIRE(H % 8 == 0 && B <= 4096, "qformer router: bad H=%d / B=%d", H, B);
  if (B <= 0 || T <= 0) return SLIME_OK;
  qf_text_rows_kernel<<<1, 1024, (B + 1) * sizeof(int), stream>>>(ids, mask, B, T, image_token, vocab, dst_row, cu);
  SLIME_AFTER_LAUNCH();
  const long long total = static_cast<long long>(B) * T * (H / 8);
  qf_gather_text_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(ids, src, dst_row, packed, B * T, H);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_qf_logits(const bf16* h, const bf16* w2, const bf16* b2, float* logit, int rows, int Dh, int B,
                           int n_per, const int* n_valid, float temp, cudaStream_t stream) {
  SLIME_REQUIRE(temp > 0.f, "qformer router: temperature must be positive");
  if (rows <= 0) return SLIME_OK;
  qf_logit_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(h, w2, b2, logit, rows, Dh);
  SLIME_AFTER_LAUNCH();
  qf_softmax_kernel<<<B, 256, 0, stream>>>(logit, n_per, n_valid, temp);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_text_dir(const long long* ids, const unsigned char* mask, const bf16* embed,
                          float* inv_norm_ws, float* tvec, int B, int T, int H, long long image_token,
                          int vocab, cudaStream_t stream) {
  SLIME_REQUIRE(H % 8 == 0, "router: hidden size %d must be a multiple of 8", H);
  if (B <= 0 || T <= 0) return SLIME_OK;
  text_inv_norm_kernel<<<(B * T + 3) / 4, 128, 0, stream>>>(ids, mask, embed, inv_norm_ws, B, T, H,
                                                            image_token, vocab);
  SLIME_AFTER_LAUNCH();
  dim3 grid((H / 2 + 127) / 128, B);
  text_dir_kernel<<<grid, 128, 0, stream>>>(ids, inv_norm_ws, embed, tvec, T, H);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_router_score(const bf16* x, const float* tvec, float* score, int rows,
                              int rows_per_sample, int H, cudaStream_t stream) {
  SLIME_REQUIRE(H % 8 == 0, "router: hidden size %d must be a multiple of 8", H);
  if (rows <= 0) return SLIME_OK;
  router_score_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(x, tvec, score, rows, rows_per_sample, H);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_router_select(const float* in, int B, int n_per, const int* n_valid, float temp,
                               float top_p, int from_probs, float* probs_out, int* sel_idx,
                               int* sel_count, cudaStream_t stream) {
  SLIME_REQUIRE(n_per <= SEL_MAX, "router: %d local tokens per sample exceeds the supported %d", n_per,
                SEL_MAX);
  SLIME_REQUIRE(temp > 0.f, "router: temperature must be positive");
  if (B <= 0) return SLIME_OK;
  router_select_kernel<<<B, SEL_THREADS, 0, stream>>>(in, n_per, n_valid, temp, top_p, from_probs, probs_out,
                                                      sel_idx, sel_count);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}
