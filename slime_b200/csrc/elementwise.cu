// HBM-bound kernels of the SliME prefill path: normalisation, embedding assembly, RoPE, row
// gathers, the gated expert mix.  All of them are one pass over their rows with 16-byte vector
// loads/stores, fp32 arithmetic, bf16 storage; none of them is shaped into a GEMM.
#include "elementwise.h"
#include "errors.h"

namespace {

constexpr int NT = 128;

struct Vec8 {
  float v[8];
};
SLIME_DEVINL Vec8 load8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  float2 t;
  t = unpack_bf16x2(u.x); r.v[0] = t.x; r.v[1] = t.y;
  t = unpack_bf16x2(u.y); r.v[2] = t.x; r.v[3] = t.y;
  t = unpack_bf16x2(u.z); r.v[4] = t.x; r.v[5] = t.y;
  t = unpack_bf16x2(u.w); r.v[6] = t.x; r.v[7] = t.y;
  return r;
}
SLIME_DEVINL void store8(bf16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
SLIME_DEVINL float round_bf16(float x) { return elem_to_float(float_to_elem(x)); }  // round to the element type

// Sum over the TPR threads that share a row.  TPR == 32: one warp per row (4 rows per CTA);
// TPR == 128: the whole CTA works on one row.
template <int TPR>
SLIME_DEVINL float row_sum(float v, float* red) {
  v = warp_sum(v);
  if (TPR == 32) return v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

SLIME_DEVINL long long src_row_of(int r, int group, int group_stride, int offset) {
  if (group <= 0) return r;
  return static_cast<long long>(r / group) * group_stride + offset + (r % group);
}

// ------------------------------------------------------------------------------------------
// LayerNorm
// ------------------------------------------------------------------------------------------
template <int TPR, int MAXV>
__global__ void __launch_bounds__(NT) layernorm_kernel(const bf16* __restrict__ x, int x_ld,
                                                       const bf16* __restrict__ w,
                                                       const bf16* __restrict__ b, bf16* __restrict__ y,
                                                       int y_ld, int rows, int D, float eps, int group,
                                                       int group_stride, int offset) {
  __shared__ float red[4];
  pdl_trigger();
  pdl_wait();  // (no-ops unless launched with the programmatic attribute: slime_launch_prefill)
  constexpr int RPB = NT / TPR;
  const int r = blockIdx.x * RPB + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const bool active = r < rows;
  const int nchunks = D >> 3;
  const bf16* xr = x + src_row_of(active ? r : 0, group, group_stride, offset) * x_ld;
  float vals[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = t + i * TPR;
    if (active && c < nchunks) {
      const Vec8 v = load8(xr + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vals[i][j] = v.v[j];
        s += v.v[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[i][j] = 0.f;
    }
  }
  const float mean = row_sum<TPR>(s, red) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = t + i * TPR;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = vals[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float var = row_sum<TPR>(sq, red) / D;
  const float rstd = rsqrtf(var + eps);
  if (!active) return;
  bf16* yr = y + static_cast<long long>(r) * y_ld;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = t + i * TPR;
    if (c < nchunks) {
      const Vec8 wv = load8(w + c * 8);
      const Vec8 bv = load8(b + c * 8);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (vals[i][j] - mean) * rstd * wv.v[j] + bv.v[j];
      store8(yr + c * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// RMSNorm (HF LlamaRMSNorm semantics: normalise in fp32, round to bf16, then multiply by weight)
// ------------------------------------------------------------------------------------------
template <int TPR, int MAXV>
__global__ void __launch_bounds__(NT) rmsnorm_kernel(const bf16* __restrict__ x, int x_ld,
                                                     const bf16* __restrict__ w, bf16* __restrict__ y,
                                                     int y_ld, int rows, int D, float eps,
                                                     const int* __restrict__ src_rows) {
  __shared__ float red[4];
  pdl_trigger();
  pdl_wait();
  constexpr int RPB = NT / TPR;
  const int r = blockIdx.x * RPB + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const bool active = r < rows;
  const int nchunks = D >> 3;
  long long sr = active ? r : 0;
  if (active && src_rows != nullptr) sr = src_rows[r];
  const bf16* xr = x + sr * x_ld;
  float vals[MAXV][8];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = t + i * TPR;
    if (active && c < nchunks) {
      const Vec8 v = load8(xr + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vals[i][j] = v.v[j];
        sq += v.v[j] * v.v[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[i][j] = 0.f;
    }
  }
  const float rstd = rsqrtf(row_sum<TPR>(sq, red) / D + eps);
  if (!active) return;
  bf16* yr = y + static_cast<long long>(r) * y_ld;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = t + i * TPR;
    if (c < nchunks) {
      const Vec8 wv = load8(w + c * 8);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = wv.v[j] * round_bf16(vals[i][j] * rstd);
      store8(yr + c * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// CLIP patch gather (im2col) and embedding assembly + pre-LayerNorm
// ------------------------------------------------------------------------------------------
__global__ void im2col_kernel(const bf16* __restrict__ px, bf16* __restrict__ out, int Nc, int image,
                              int patch, int Kpad) {
  const int grid = image / patch;
  const int K = 3 * patch * patch;
  const long long total = static_cast<long long>(Nc) * grid * grid * Kpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kpad);
    const long long prow = i / Kpad;
    bf16 v = float_to_elem(0.f);
    if (k < K) {
      const int c = k / (patch * patch);
      const int ky = (k / patch) % patch;
      const int kx = k % patch;
      const int pidx = static_cast<int>(prow % (grid * grid));
      const long long crop = prow / (grid * grid);
      const int py = pidx / grid, pxx = pidx % grid;
      v = px[((crop * 3 + c) * image + (py * patch + ky)) * image + (pxx * patch + kx)];
    }
    out[i] = v;
  }
}

template <int MAXV>
__global__ void __launch_bounds__(NT) clip_embed_ln_kernel(const bf16* __restrict__ patch_out,
                                                           const bf16* __restrict__ cls,
                                                           const bf16* __restrict__ pos,
                                                           const bf16* __restrict__ w,
                                                           const bf16* __restrict__ b,
                                                           bf16* __restrict__ h, int Nc, int tokens,
                                                           int D, float eps) {
  // one warp per token row; `tokens` = 1 + patches per crop
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= Nc * tokens) return;
  const int crop = r / tokens, t = r % tokens;
  const bf16* src = (t == 0) ? cls : patch_out + (static_cast<long long>(crop) * (tokens - 1) + (t - 1)) * D;
  const bf16* pr = pos + static_cast<long long>(t) * D;
  const int nchunks = D >> 3;
  float vals[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nchunks) {
      const Vec8 a = load8(src + c * 8);
      const Vec8 p = load8(pr + c * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vals[i][j] = round_bf16(a.v[j] + p.v[j]);  // HF adds in bf16 before pre_layrnorm
        s += vals[i][j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[i][j] = 0.f;
    }
  }
  const float mean = warp_sum(s) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = vals[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
  bf16* hr = h + static_cast<long long>(r) * D;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nchunks) {
      const Vec8 wv = load8(w + c * 8);
      const Vec8 bv = load8(b + c * 8);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (vals[i][j] - mean) * rstd * wv.v[j] + bv.v[j];
      store8(hr + c * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// row copies
// ------------------------------------------------------------------------------------------
__global__ void copy_rows_kernel(const bf16* __restrict__ src, int src_ld, bf16* __restrict__ dst,
                                 int dst_ld, int rows, int D, int group, int group_stride, int offset,
                                 const int* __restrict__ dst_rows) {
  const int nchunks = D >> 3;
  const long long total = static_cast<long long>(rows) * nchunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / nchunks);
    const int c = static_cast<int>(i % nchunks);
    long long dr = r;
    if (dst_rows != nullptr) {
      dr = dst_rows[r];
      if (dr < 0) continue;
    }
    const uint4 v = *reinterpret_cast<const uint4*>(src + src_row_of(r, group, group_stride, offset) * src_ld + c * 8);
    *reinterpret_cast<uint4*>(dst + dr * dst_ld + c * 8) = v;
  }
}

// Output rows of the vision tower for a batch of images with `per_image` crops each (crop 0 = the global view):
// the CLS row of every crop is dropped and crop ci = img * per_image + j lands in slot img (j == 0: all global crops first,
// [images, P, D]) or images + img * (per_image - 1) + j - 1 (the local crops behind them) - the adapter stages read both
// groups as contiguous tensors without a gather (reference llava_arch.py:212-225 slices them per sample).
__global__ void vit_split_rows_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int rows, int D, int P, int TK,
                                      int crop0, int per_image, int images) {
  const int nchunks = D >> 3;
  const long long total = static_cast<long long>(rows) * nchunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / nchunks);
    const int c = static_cast<int>(i % nchunks);
    const int crop = r / P, pr = r - crop * P;
    const int ci = crop0 + crop, img = ci / per_image, j = ci - img * per_image;
    const long long slot = j == 0 ? img : static_cast<long long>(images) + static_cast<long long>(img) * (per_image - 1) + j - 1;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (static_cast<long long>(crop) * TK + 1 + pr) * D + c * 8);
    *reinterpret_cast<uint4*>(dst + (slot * P + pr) * D + c * 8) = v;
  }
}

__global__ void add_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                bf16* __restrict__ y, int rows, int D, int period) {
  const int nchunks = D >> 3;
  const long long total = static_cast<long long>(rows) * nchunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / nchunks);
    const int c = static_cast<int>(i % nchunks);
    const int br = period > 0 ? r % period : r;
    const Vec8 av = load8(a + static_cast<long long>(r) * D + c * 8);
    const Vec8 bv = load8(b + static_cast<long long>(br) * D + c * 8);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = av.v[j] + bv.v[j];
    store8(y + static_cast<long long>(r) * D + c * 8, o);
  }
}

// ------------------------------------------------------------------------------------------
// RoPE
// ------------------------------------------------------------------------------------------
// table[pos][i] = (cos, sin) of pos * theta^(-2i/head_dim), i < head_dim/2, computed in fp32 like
// HF LlamaRotaryEmbedding (llama/modeling_llama.py:114-137) but kept in fp32 (HF rounds to bf16).
__global__ void rope_table_kernel(float2* __restrict__ table, int max_pos, int half, float theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * half) return;
  const int pos = i / half, f = i % half;
  const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * f) / static_cast<float>(2 * half));
  const float ang = static_cast<float>(pos) * inv_freq;
  float s, c;
  sincosf(ang, &s, &c);
  table[i] = make_float2(c, s);
}

// One thread handles 8 consecutive feature pairs (i, i + half) of one head of one token.
__global__ void rope_kernel(bf16* __restrict__ qkv, int ld, int rows, int n_heads_total, int head_dim,
                            const int* __restrict__ pos_ids, const float2* __restrict__ table,
                            int max_pos) {
  const int half = head_dim >> 1;
  const int cpb = half >> 3;  // 8-wide chunks per half head
  const long long total = static_cast<long long>(rows) * n_heads_total * cpb;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpb);
    const int hd = static_cast<int>((i / cpb) % n_heads_total);
    const int r = static_cast<int>(i / (static_cast<long long>(cpb) * n_heads_total));
    int pos = pos_ids[r];
    pos = min(max(pos, 0), max_pos - 1);
    bf16* base = qkv + static_cast<long long>(r) * ld + hd * head_dim + c * 8;
    const Vec8 lo = load8(base);
    const Vec8 hi = load8(base + half);
    const float2* cs = table + static_cast<long long>(pos) * half + c * 8;
    float olo[8], ohi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 t = cs[j];
      olo[j] = lo.v[j] * t.x - hi.v[j] * t.y;
      ohi[j] = hi.v[j] * t.x + lo.v[j] * t.y;
    }
    store8(base, olo);
    store8(base + half, ohi);
  }
}

// ------------------------------------------------------------------------------------------
// gated mix of the two global experts (reference multimodal_projector/builder.py:137-171,203-206)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) gate_mix_kernel(const bf16* __restrict__ x,
                                                      const bf16* __restrict__ w_gate,
                                                      const bf16* __restrict__ e0,
                                                      const bf16* __restrict__ e1, bf16* __restrict__ out,
                                                      int rows, int Dm, int H) {
  __shared__ float red0[4], red1[4];
  const int r = blockIdx.x;
  if (r >= rows) return;
  const bf16* xr = x + static_cast<long long>(r) * Dm;
  float l0 = 0.f, l1 = 0.f;
  for (int k = threadIdx.x * 2; k < Dm; k += NT * 2) {
    // w_gate is [Dm, 2] row-major: (k,0),(k,1),(k+1,0),(k+1,1) are 8 contiguous bytes
    const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xr + k));
    const uint2 wv = *reinterpret_cast<const uint2*>(w_gate + 2 * k);
    const float2 w0 = unpack_bf16x2(wv.x), w1 = unpack_bf16x2(wv.y);
    l0 += xv.x * w0.x + xv.y * w1.x;
    l1 += xv.x * w0.y + xv.y * w1.y;
  }
  l0 = warp_sum(l0);
  l1 = warp_sum(l1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red0[warp] = l0;
    red1[warp] = l1;
  }
  __syncthreads();
  l0 = red0[0] + red0[1] + red0[2] + red0[3];
  l1 = red1[0] + red1[1] + red1[2] + red1[3];
  const float m = fmaxf(l0, l1);
  const float p0 = expf(l0 - m), p1 = expf(l1 - m);
  const float inv = 1.0f / (p0 + p1);
  const float s0 = p0 * inv, s1 = p1 * inv;
  const float g0 = s0 / (s0 + s1 + 1e-6f), g1 = s1 / (s0 + s1 + 1e-6f);
  const bf16* a = e0 + static_cast<long long>(r) * H;
  const bf16* b = e1 + static_cast<long long>(r) * H;
  bf16* o = out + static_cast<long long>(r) * H;
  for (int c = threadIdx.x; c < (H >> 3); c += NT) {
    const Vec8 av = load8(a + c * 8);
    const Vec8 bv = load8(b + c * 8);
    float ov[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ov[j] = g0 * av.v[j] + g1 * bv.v[j];
    store8(o + c * 8, ov);
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// one warp per row: rstd = rsqrt(mean(x^2) + eps) (the squares are taken of the 16-bit values, summed in fp32)
__global__ void __launch_bounds__(256) row_rstd_kernel(const bf16* __restrict__ x, int x_ld, float* __restrict__ rstd,
                                                       int rows, int D, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + static_cast<size_t>(row) * x_ld;
  float acc = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    acc += (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y) + (cc.x * cc.x + cc.y * cc.y) + (d.x * d.x + d.y * d.y);
  }
  acc = warp_sum(acc);
  if (lane == 0) rstd[row] = rsqrtf(acc / static_cast<float>(D) + eps);
}

// one warp per row: sums the row's partials in a fixed order (lane-strided, then the shuffle tree)
__global__ void __launch_bounds__(256) sumsq_to_rstd_kernel(const float* __restrict__ partials, int parts,
                                                            float* __restrict__ rstd, int rows, int D, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* pr = partials + static_cast<size_t>(row) * parts;
  float acc = 0.f;
  for (int i = lane; i < parts; i += 32) acc += pr[i];
  acc = warp_sum(acc);
  if (lane == 0) rstd[row] = rsqrtf(acc / static_cast<float>(D) + eps);
}

}  // namespace

int slime_launch_layernorm(const bf16* x, int x_ld, const bf16* w, const bf16* b, bf16* y, int y_ld,
                           int rows, int D, float eps, int in_group, int in_group_stride,
                           int in_offset, cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && D <= 8192, "layernorm: bad D=%d", D);
  if (rows <= 0) return SLIME_OK;
  cudaError_t le;
  if (D <= 1024) {
    le = slime_launch_prefill(layernorm_kernel<32, 4>, dim3((rows + 3) / 4), dim3(NT), 0, stream, x, x_ld, w, b, y, y_ld, rows,
                              D, eps, in_group, in_group_stride, in_offset);
  } else if (D <= 4096) {
    le = slime_launch_prefill(layernorm_kernel<128, 4>, dim3(rows), dim3(NT), 0, stream, x, x_ld, w, b, y, y_ld, rows, D, eps,
                              in_group, in_group_stride, in_offset);
  } else {
    le = slime_launch_prefill(layernorm_kernel<128, 8>, dim3(rows), dim3(NT), 0, stream, x, x_ld, w, b, y, y_ld, rows, D, eps,
                              in_group, in_group_stride, in_offset);
  }
  SLIME_CHECK_CUDA(le);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_row_rstd(const bf16* x, int x_ld, float* rstd, int rows, int D, float eps, cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && x_ld % 8 == 0, "row_rstd: bad D=%d", D);
  if (rows <= 0) return SLIME_OK;
  SLIME_CHECK_CUDA(slime_launch_prefill(row_rstd_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, x, x_ld, rstd, rows, D, eps));
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_sumsq_to_rstd(const float* partials, int parts, float* rstd, int rows, int D, float eps,
                               cudaStream_t stream) {
  SLIME_REQUIRE(parts > 0 && parts <= 1024, "sumsq_to_rstd: bad parts=%d", parts);
  if (rows <= 0) return SLIME_OK;
  SLIME_CHECK_CUDA(slime_launch_prefill(sumsq_to_rstd_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, partials, parts, rstd,
                                        rows, D, eps));
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_rmsnorm(const bf16* x, int x_ld, const bf16* w, bf16* y, int y_ld, int rows, int D,
                         float eps, const int* src_rows, cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && D <= 8192, "rmsnorm: bad D=%d", D);
  if (rows <= 0) return SLIME_OK;
  cudaError_t le;
  if (D <= 1024) {
    le = slime_launch_prefill(rmsnorm_kernel<32, 4>, dim3((rows + 3) / 4), dim3(NT), 0, stream, x, x_ld, w, y, y_ld, rows, D, eps,
                              src_rows);
  } else if (D <= 4096) {
    le = slime_launch_prefill(rmsnorm_kernel<128, 4>, dim3(rows), dim3(NT), 0, stream, x, x_ld, w, y, y_ld, rows, D, eps, src_rows);
  } else {
    le = slime_launch_prefill(rmsnorm_kernel<128, 8>, dim3(rows), dim3(NT), 0, stream, x, x_ld, w, y, y_ld, rows, D, eps, src_rows);
  }
  SLIME_CHECK_CUDA(le);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_im2col(const bf16* pixels, bf16* patches, int Nc, int image, int patch, int Kpad,
                        cudaStream_t stream) {
  SLIME_REQUIRE(image % patch == 0 && Kpad >= 3 * patch * patch && Kpad % 8 == 0, "im2col: bad geometry");
  if (Nc <= 0) return SLIME_OK;
  const long long total = static_cast<long long>(Nc) * (image / patch) * (image / patch) * Kpad;
  im2col_kernel<<<grid_for(total, 256), 256, 0, stream>>>(pixels, patches, Nc, image, patch, Kpad);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_clip_embed_ln(const bf16* patch_out, const bf16* cls, const bf16* pos, const bf16* w,
                               const bf16* b, bf16* h, int Nc, int tokens, int D, float eps,
                               cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && D <= 2048, "clip_embed_ln: bad D=%d", D);
  if (Nc <= 0) return SLIME_OK;
  const int rows = Nc * tokens;
  if (D <= 1024) {
    clip_embed_ln_kernel<4><<<(rows + 3) / 4, NT, 0, stream>>>(patch_out, cls, pos, w, b, h, Nc, tokens, D, eps);
  } else {
    clip_embed_ln_kernel<8><<<(rows + 3) / 4, NT, 0, stream>>>(patch_out, cls, pos, w, b, h, Nc, tokens, D, eps);
  }
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_copy_rows(const bf16* src, int src_ld, bf16* dst, int dst_ld, int rows, int D,
                           int group, int group_stride, int offset, cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0, "copy_rows: bad D=%d", D);
  if (rows <= 0) return SLIME_OK;
  copy_rows_kernel<<<grid_for(static_cast<long long>(rows) * (D / 8), 256), 256, 0, stream>>>(
      src, src_ld, dst, dst_ld, rows, D, group, group_stride, offset, nullptr);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_vit_split_rows(const bf16* src, bf16* dst, int crops, int D, int P, int TK, int crop0, int per_image,
                                int images, cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && per_image >= 1 && images >= 1, "vit_split_rows: bad D=%d per_image=%d", D, per_image);
  if (crops <= 0) return SLIME_OK;
  const long long rows = static_cast<long long>(crops) * P;
  vit_split_rows_kernel<<<grid_for(rows * (D / 8), 256), 256, 0, stream>>>(src, dst, static_cast<int>(rows), D, P, TK, crop0,
                                                                         per_image, images);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_scatter_rows(const bf16* src, int src_ld, bf16* dst, int dst_ld, int rows, int D,
                              const int* dst_rows, cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0, "scatter_rows: bad D=%d", D);
  if (rows <= 0) return SLIME_OK;
  copy_rows_kernel<<<grid_for(static_cast<long long>(rows) * (D / 8), 256), 256, 0, stream>>>(
      src, src_ld, dst, dst_ld, rows, D, 0, 0, 0, dst_rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_add_rows(const bf16* a, const bf16* b, bf16* y, int rows, int D, int period,
                          cudaStream_t stream) {
  SLIME_REQUIRE(D % 8 == 0, "add_rows: bad D=%d", D);
  if (rows <= 0) return SLIME_OK;
  add_rows_kernel<<<grid_for(static_cast<long long>(rows) * (D / 8), 256), 256, 0, stream>>>(a, b, y, rows,
                                                                                           D, period);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_rope_table(float* table, int max_pos, int head_dim, float theta, cudaStream_t stream) {
  const int half = head_dim / 2;
  const int total = max_pos * half;
  rope_table_kernel<<<(total + 255) / 256, 256, 0, stream>>>(reinterpret_cast<float2*>(table), max_pos, half,
                                                             theta);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_rope(bf16* qkv, int ld, int rows, int n_q_heads, int n_k_heads, int head_dim,
                      const int* pos_ids, const float* cos_sin_table, int max_pos, cudaStream_t stream) {
  SLIME_REQUIRE(head_dim % 16 == 0 && ld % 8 == 0, "rope: bad head_dim=%d", head_dim);
  if (rows <= 0) return SLIME_OK;
  const int heads = n_q_heads + n_k_heads;  // q heads then k heads are contiguous in the packed row
  const long long total = static_cast<long long>(rows) * heads * (head_dim / 16);
  rope_kernel<<<grid_for(total, 256), 256, 0, stream>>>(qkv, ld, rows, heads, head_dim, pos_ids,
                                                        reinterpret_cast<const float2*>(cos_sin_table),
                                                        max_pos);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_gate_mix(const bf16* x, const bf16* w_gate, const bf16* e0, const bf16* e1, bf16* out,
                          int rows, int Dm, int H, cudaStream_t stream) {
  SLIME_REQUIRE(Dm % 2 == 0 && H % 8 == 0, "gate_mix: bad dims Dm=%d H=%d", Dm, H);
  if (rows <= 0) return SLIME_OK;
  gate_mix_kernel<<<rows, NT, 0, stream>>>(x, w_gate, e0, e1, out, rows, Dm, H);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}
