// Single-query (decode step) attention over the KV cache written by the prefill - the step that follows the
// prefill in generate() (reference llava_llama.py:139 -> HF generation loop).
// HBM-bound: every cached K/V row of the sequence is read once per q head group; one CTA per (q head, sequence),
// 4 warps stride over the cached positions with an online softmax each and are merged through shared memory.
#include "attention.h"
#include "errors.h"

namespace {

constexpr int NT = 128;

template <int HD>
__global__ void __launch_bounds__(NT) decode_attn_kernel(const bf16* __restrict__ q, int q_ld,
                                                         const bf16* __restrict__ kcache,
                                                         const bf16* __restrict__ vcache, int cache_len,
                                                         const int* __restrict__ lens, int heads, int kv_heads,
                                                         float scale_log2, bf16* __restrict__ out, int out_ld) {
  constexpr int EPL = HD / 32;  // elements per lane
  __shared__ float s_m[4], s_l[4];
  __shared__ float s_acc[4][HD];
  const int head = blockIdx.x, b = blockIdx.y;
  const int kvh = head / (heads / kv_heads);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lens[b] + 1;  // cached tokens + the one appended in this step
  const int KD = kv_heads * HD;

  float qv[EPL];
  {
    const bf16* qp = q + static_cast<long long>(b) * q_ld + head * HD + lane * EPL;
#pragma unroll
    for (int i = 0; i < EPL; ++i) qv[i] = elem_to_float(qp[i]);
  }
  float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;

  const long long base = (static_cast<long long>(b) * cache_len) * KD + kvh * HD + lane * EPL;
  for (int pos = warp; pos < len; pos += 4) {
    const bf16* kp = kcache + base + static_cast<long long>(pos) * KD;
    const bf16* vp = vcache + base + static_cast<long long>(pos) * KD;
    float kv[EPL], vv[EPL];
    if (EPL == 4) {
      const uint2 ku = *reinterpret_cast<const uint2*>(kp);
      const uint2 vu = *reinterpret_cast<const uint2*>(vp);
      float2 t;
      t = unpack_bf16x2(ku.x); kv[0] = t.x; kv[1] = t.y;
      t = unpack_bf16x2(ku.y); kv[2] = t.x; kv[3] = t.y;
      t = unpack_bf16x2(vu.x); vv[0] = t.x; vv[1] = t.y;
      t = unpack_bf16x2(vu.y); vv[2] = t.x; vv[3] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        kv[i] = elem_to_float(kp[i]);
        vv[i] = elem_to_float(vp[i]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += qv[i] * kv[i];
    s = warp_sum(s) * scale_log2;
    const float m_new = fmaxf(m, s);
    const float alpha = exp2f(m - m_new);  // m == -inf -> 0
    const float pexp = exp2f(s - m_new);
    l = l * alpha + pexp;
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = acc[i] * alpha + pexp * vv[i];
    m = m_new;
  }
  if (lane == 0) {
    s_m[warp] = m;
    s_l[warp] = l;
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) s_acc[warp][lane * EPL + i] = acc[i];
  __syncthreads();
  if (warp == 0) {
    const float mt = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float lt = 0.f, o[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) o[i] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float a = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - mt);
      lt += s_l[w] * a;
#pragma unroll
      for (int i = 0; i < EPL; ++i) o[i] += s_acc[w][lane * EPL + i] * a;
    }
    const float inv = lt > 0.f ? 1.0f / lt : 0.f;
    bf16* op = out + static_cast<long long>(b) * out_ld + head * HD + lane * EPL;
#pragma unroll
    for (int i = 0; i < EPL; ++i) op[i] = float_to_elem(o[i] * inv);
  }
}

// cache_rows[i] = sample(i) * cache_len + pos_ids[i]   (packed prefill row -> slot in the per-sequence cache)
__global__ void cache_rows_kernel(const int* __restrict__ cu, const int* __restrict__ pos_ids, int B, int total,
                                  int cache_len, int* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= i) lo = mid; else hi = mid;
  }
  const int pos = pos_ids[i];
  rows[i] = pos < cache_len ? lo * cache_len + pos : -1;
}

// rows[b] = b * cache_len + lens[b]   (slot of the token appended in this decode step)
__global__ void append_rows_kernel(const int* __restrict__ lens, int B, int cache_len, int* __restrict__ rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) rows[b] = lens[b] < cache_len ? b * cache_len + lens[b] : -1;
}

}  // namespace

int slime_launch_decode_attention(const bf16* q, int q_ld, const bf16* kcache, const bf16* vcache, int cache_len,
                                  const int* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                                  bf16* out, int out_ld, cudaStream_t stream) {
  SLIME_REQUIRE(head_dim == 64 || head_dim == 128, "decode attention: head_dim %d unsupported", head_dim);
  if (batch <= 0) return SLIME_OK;
  dim3 grid(heads, batch);
  const float sl2 = scale * 1.4426950408889634f;
  slime_prof_begin(1, 0.0, stream);
  if (head_dim == 128) {
    decode_attn_kernel<128><<<grid, NT, 0, stream>>>(q, q_ld, kcache, vcache, cache_len, lens, heads, kv_heads, sl2, out, out_ld);
  } else {
    decode_attn_kernel<64><<<grid, NT, 0, stream>>>(q, q_ld, kcache, vcache, cache_len, lens, heads, kv_heads, sl2, out, out_ld);
  }
  slime_prof_end(stream);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_cache_rows(const int* cu, const int* pos_ids, int B, int total, int cache_len, int* rows,
                            cudaStream_t stream) {
  if (total <= 0) return SLIME_OK;
  cache_rows_kernel<<<(total + 255) / 256, 256, 0, stream>>>(cu, pos_ids, B, total, cache_len, rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_append_rows(const int* lens, int B, int cache_len, int* rows, cudaStream_t stream) {
  if (B <= 0) return SLIME_OK;
  append_rows_kernel<<<(B + 127) / 128, 128, 0, stream>>>(lens, B, cache_len, rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}
