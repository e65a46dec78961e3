// Fused (flash-style) attention forward for the SliME prefill path, bf16 in / fp32 softmax / bf16 out.
//
// One kernel template serves the three attention shapes on the path:
//   * CLIP ViT self-attention: 16 heads, head_dim 64, S = 577, non-causal, fixed length
//     (HF clip/modeling_clip.py:318-331),
//   * Resampler cross-attention: 8 heads, head_dim 128, 144 or 576 learned queries shared by every
//     crop against 576 keys (reference llava/model/multimodal_resampler/sampler.py:160-164),
//   * Llama decoder self-attention: head_dim 128, causal, GQA, packed variable-length sequences
//     (HF llama/modeling_llama.py:269-286); packing replaces the reference's right-padding + mask.
//
// Round-1 implementation: 64x64 tiles, 4 warps, K/V double-buffered with cp.async, S = QK^T and
// O += PV on mma.sync.m16n8k16 with the online-softmax running max/sum in registers.  (Attention is
// 3 % of the decoder FLOPs and 9 % of the ViT FLOPs; the tcgen05/TMEM version is the planned upgrade.)
#include <cstdlib>

#include "attention.h"
#include "errors.h"

namespace {

constexpr int BM = 64;
constexpr int BN = 64;
constexpr int NTHREADS = 128;

template <int HD>
SLIME_DEVINL uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * HD * 2 + ((chunk ^ (row & 7)) << 4));
}

SLIME_DEVINL void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
SLIME_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
SLIME_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
SLIME_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
SLIME_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
SLIME_DEVINL void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." SLIME_MMA_SYNC_TYPE "." SLIME_MMA_SYNC_TYPE ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Copies a [64 x HD] bf16 tile (rows >= rows_valid are zero-filled) into swizzled shared memory.
template <int HD>
SLIME_DEVINL void load_tile(uint32_t smem_base, const bf16* g, long long ld, int rows_valid, int tid) {
  constexpr int CH = HD / 8;
#pragma unroll
  for (int i = tid; i < 64 * CH; i += NTHREADS) {
    const int row = i / CH;
    const int chunk = i % CH;
    const bool ok = row < rows_valid;
    const bf16* src = ok ? (g + row * ld + chunk * 8) : g;
    cp_async16(smem_base + swz<HD>(row, chunk), src, ok ? 16 : 0);
  }
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NTHREADS) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int TILE_BYTES = 64 * HD * 2;
  const uint32_t sQ = smem_u32(smem_attn);
  const uint32_t sK = sQ + TILE_BYTES;
  const uint32_t sV = sK + 2 * TILE_BYTES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int b = blockIdx.z;
  const int head = blockIdx.y;
  const int kv_head = head / (p.num_heads / p.num_kv_heads);
  const int m_blk = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;

  long long q_row0, k_row0, o_row0;
  int len_q, len_k;
  if (p.cu_q != nullptr) {
    q_row0 = p.cu_q[b];
    len_q = p.cu_q[b + 1] - p.cu_q[b];
    o_row0 = q_row0;
  } else {
    q_row0 = b * p.q_batch_rows;
    o_row0 = b * p.o_batch_rows;
    len_q = p.seqlen_q;
  }
  if (p.cu_k != nullptr) {
    k_row0 = p.cu_k[b];
    len_k = p.cu_k[b + 1] - p.cu_k[b];
  } else {
    k_row0 = b * p.k_batch_rows;
    len_k = p.seqlen_k;
  }
  const int m0 = m_blk * BM;
  if (m0 >= len_q) return;
  const int causal_off = len_k - len_q;

  int n_tiles = (len_k + BN - 1) / BN;
  if (CAUSAL) {
    const int last_col = min(len_k, m0 + BM + causal_off);  // exclusive
    n_tiles = max(0, (last_col + BN - 1) / BN);
  }

  const bf16* qg = p.q + (q_row0 + m0) * p.q_ld + head * HD;
  const bf16* kg = p.k + k_row0 * p.k_ld + kv_head * HD;
  const bf16* vg = p.v + k_row0 * p.v_ld + kv_head * HD;

  load_tile<HD>(sQ, qg, p.q_ld, len_q - m0, tid);
  if (n_tiles > 0) {
    load_tile<HD>(sK, kg, p.k_ld, len_k, tid);
    load_tile<HD>(sV, vg, p.v_ld, len_k, tid);
  }
  cp_async_commit();

  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  }
  float row_max[2] = {-INFINITY, -INFINITY};
  float row_sum[2] = {0.f, 0.f};
  uint32_t q_frag[HD / 16][4];
  const float scale_log2 = p.scale * 1.4426950408889634f;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_tiles) {
      const int nb = buf ^ 1;
      load_tile<HD>(sK + nb * TILE_BYTES, kg + static_cast<long long>(j + 1) * BN * p.k_ld, p.k_ld,
                    len_k - (j + 1) * BN, tid);
      load_tile<HD>(sV + nb * TILE_BYTES, vg + static_cast<long long>(j + 1) * BN * p.v_ld, p.v_ld,
                    len_k - (j + 1) * BN, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        ldsm_x4(sQ + swz<HD>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), q_frag[ks][0],
                q_frag[ks][1], q_frag[ks][2], q_frag[ks][3]);
      }
    }

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t sKb = sK + buf * TILE_BYTES;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(sKb + swz<HD>(np * 16 + ((lane >> 4) << 3) + (lane & 7), ks * 2 + ((lane >> 3) & 1)),
                r0, r1, r2, r3);
        mma_bf16(s[2 * np], q_frag[ks], r0, r1);
        mma_bf16(s[2 * np + 1], q_frag[ks], r2, r3);
      }
    }

    // ---- masking (sequence tail, causal diagonal) ----
    const int col_base = j * BN;
    const bool need_mask = (col_base + BN > len_k) || (CAUSAL && (col_base + BN - 1 > m0 + causal_off));
    if (need_mask) {
      const int row_lo = m0 + warp * 16 + (lane >> 2);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = col_base + nb * 8 + ((lane & 3) << 1) + (e & 1);
          const int row = row_lo + ((e >> 1) << 3);
          bool masked = col >= len_k;
          if (CAUSAL) masked = masked || (col > row + causal_off);
          if (masked) s[nb][e] = -INFINITY;
        }
      }
    }

    // ---- online softmax ----
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = row_max[r];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * r], s[nb][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_scaled = (mx == -INFINITY) ? 0.f : mx * scale_log2;
      alpha[r] = (row_max[r] == -INFINITY) ? 0.f : exp2f(row_max[r] * scale_log2 - m_scaled);
      row_max[r] = mx;
      float sum = 0.f;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float p0 = exp2f(s[nb][2 * r] * scale_log2 - m_scaled);
        const float p1 = exp2f(s[nb][2 * r + 1] * scale_log2 - m_scaled);
        s[nb][2 * r] = p0;
        s[nb][2 * r + 1] = p1;
        sum += p0 + p1;
      }
      row_sum[r] = row_sum[r] * alpha[r] + sum;
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o_acc[i][0] *= alpha[0];
      o_acc[i][1] *= alpha[0];
      o_acc[i][2] *= alpha[1];
      o_acc[i][3] *= alpha[1];
    }

    // ---- O += P V ----
    const uint32_t sVb = sV + buf * TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(sVb + swz<HD>(kk * 16 + (((lane >> 3) & 1) << 3) + (lane & 7), dp * 2 + (lane >> 4)),
                  r0, r1, r2, r3);
        mma_bf16(o_acc[2 * dp], pa, r0, r1);
        mma_bf16(o_acc[2 * dp + 1], pa, r2, r3);
      }
    }
    __syncthreads();
  }
  if (n_tiles == 0) {
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- finalise: O /= l, stage through this warp's own Q rows, coalesced 16-byte stores ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = row_sum[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    row_sum[r] = (l > 0.f) ? 1.0f / l : 0.f;
  }
  uint8_t* sQ_gen = smem_attn;
#pragma unroll
  for (int nb = 0; nb < HD / 8; ++nb) {
    const int r_lo = warp * 16 + (lane >> 2);
    const int r_hi = r_lo + 8;
    const int byte_in_chunk = (lane & 3) * 4;
    *reinterpret_cast<uint32_t*>(sQ_gen + swz<HD>(r_lo, nb) + byte_in_chunk) =
        pack_bf16x2(o_acc[nb][0] * row_sum[0], o_acc[nb][1] * row_sum[0]);
    *reinterpret_cast<uint32_t*>(sQ_gen + swz<HD>(r_hi, nb) + byte_in_chunk) =
        pack_bf16x2(o_acc[nb][2] * row_sum[1], o_acc[nb][3] * row_sum[1]);
  }
  __syncwarp();
  bf16* og = p.o + (o_row0 + m0) * p.o_ld + head * HD;
  constexpr int CH = HD / 8;
#pragma unroll
  for (int i = lane; i < 16 * CH; i += 32) {
    const int row = warp * 16 + i / CH;
    const int chunk = i % CH;
    if (m0 + row < len_q) {
      const uint4 val = *reinterpret_cast<const uint4*>(sQ_gen + swz<HD>(row, chunk));
      *reinterpret_cast<uint4*>(og + static_cast<long long>(row) * p.o_ld + chunk * 8) = val;
    }
  }
}

template <int HD, bool CAUSAL>
int launch(const AttnParams& p, cudaStream_t stream) {
  constexpr int SMEM = 5 * 64 * HD * 2;
  auto kern = attn_fwd_kernel<HD, CAUSAL>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  dim3 grid((p.seqlen_q + BM - 1) / BM, p.num_heads, p.batch);
  // FLOPs are only known exactly for fixed-length batches (4*Sq*Sk*hd per head, halved when causal)
  double flops = 0.0;
  if (p.cu_q == nullptr)
    flops = 4.0 * p.seqlen_q * static_cast<double>(p.seqlen_k) * HD * p.num_heads * p.batch * (CAUSAL ? 0.5 : 1.0);
  slime_prof_begin(1, flops, stream);
  kern<<<grid, NTHREADS, SMEM, stream>>>(p);
  slime_prof_end(stream);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

}  // namespace

static long long* g_attn_trace = nullptr;
extern "C" int slime_attention_set_trace(long long* buf) {
  g_attn_trace = buf;
  return SLIME_OK;
}

int slime_launch_attention(const AttnParams& p_in, cudaStream_t stream) {
  AttnParams p = p_in;
  p.trace = g_attn_trace;
  SLIME_REQUIRE(p.q && p.k && p.v && p.o, "attention: null tensor");
  SLIME_REQUIRE(p.head_dim == 64 || p.head_dim == 128, "attention: head_dim %d unsupported", p.head_dim);
  SLIME_REQUIRE(p.num_kv_heads > 0 && p.num_heads % p.num_kv_heads == 0, "attention: bad GQA heads %d/%d",
                p.num_heads, p.num_kv_heads);
  SLIME_REQUIRE(p.q_ld % 8 == 0 && p.k_ld % 8 == 0 && p.v_ld % 8 == 0 && p.o_ld % 8 == 0,
                "attention: row strides must be multiples of 8 elements");
  SLIME_REQUIRE(((reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.k) |
                  reinterpret_cast<uintptr_t>(p.v) | reinterpret_cast<uintptr_t>(p.o)) & 15) == 0,
                "attention: tensors must be 16-byte aligned");
  if (p.batch <= 0 || p.seqlen_q <= 0) return SLIME_OK;
  // implementation choice: tcgen05/TMEM kernel by default; SLIME_ATTN_IMPL=fa2 (or impl == 1) selects the
  // mma.sync kernel kept for A/B measurements
  static int env_impl = -1;
  static int num_sms = 0;
  if (env_impl < 0) {
    const char* e = getenv("SLIME_ATTN_IMPL");
    env_impl = (e == nullptr) ? SLIME_ATTN_DEFAULT_IMPL : ((e[0] == 'f' || e[0] == '1') ? 1 : ((e[0] == 'p' || e[0] == '3') ? 3 : 2));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int impl = p.impl != 0 ? p.impl : env_impl;
  if (impl == 2) return slime_launch_attention_tc(p, num_sms, stream);
  if (p.head_dim == 64) {
    return p.causal ? launch<64, true>(p, stream) : launch<64, false>(p, stream);
  }
  return p.causal ? launch<128, true>(p, stream) : launch<128, false>(p, stream);
}
