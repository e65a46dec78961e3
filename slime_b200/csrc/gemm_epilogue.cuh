// Epilogue of the tcgen05 GEMM (included by gemm_sm100.cu inside its anonymous namespace).
//
// 8 epilogue warps: warp w may only touch TMEM lane quadrant w % 4, so two warps share each
// quadrant and split the accumulator columns in halves.  Each thread owns one output row and walks
// its columns in chunks of 32; the loop is software-pipelined - the tcgen05.ld of chunk c+1 and the
// global loads of its residual values are issued BEFORE chunk c is processed - because with only
// 8 resident warps per SM nothing else hides the TMEM / global-load latency (the first ncu capture
// of the K = 1024 ViT GEMMs showed the tensor pipe 18-46 % active, waiting on the epilogue).
#pragma once

// sigmoid(y) = 0.5 + 0.5 * tanh(y / 2): ONE MUFU op (tanh.approx) instead of ex2 + rcp.  The SFUs retire only
// 16 results per clock per SM, and the ncu capture of the K = 1024 ViT fc1 GEMM showed its quick-GELU epilogue
// (2 MUFU per element) exactly as long as the main loop (tensor pipe 70 % active).
SLIME_DEVINL float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SLIME_DEVINL float fast_sigmoid(float y) { return fmaf(0.5f, fast_tanh(0.5f * y), 0.5f); }
SLIME_DEVINL float act_quick_gelu(float x) { return x * fast_sigmoid(1.702f * x); }
SLIME_DEVINL float act_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
SLIME_DEVINL float act_silu(float x) { return x * fast_sigmoid(x); }

struct EpiRow {
  bool store_ok;
  int out_row;
  int res_row;
};

template <int EPI>
SLIME_DEVINL void epi_issue_residual(const GemmParams& p, const EpiRow& er, int col0, uint4 (&res)[4]) {
  if constexpr (EPI != GEMM_EPI_SWIGLU) {
    if (p.residual != nullptr && er.store_ok) {
      const bf16* rp = p.residual + static_cast<size_t>(er.res_row) * p.res_ld + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col0 + g * 8 < p.N) res[g] = *reinterpret_cast<const uint4*>(rp + g * 8);
      }
    }
  }
}

template <int EPI>
SLIME_DEVINL void epi_process_chunk(const GemmParams& p, const EpiRow& er, int col0, const uint32_t (&r)[32],
                                    const uint4 (&res)[4]) {
  if (!er.store_ok || col0 >= p.N) return;
  if constexpr (EPI == GEMM_EPI_SWIGLU) {
    // columns are (gate_j, up_j) interleaved -> 16 outputs per 32 accumulator columns
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int col = col0 + g * 16;
      if (col >= p.N) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gate = __uint_as_float(r[g * 16 + 2 * j]);
        const float up = __uint_as_float(r[g * 16 + 2 * j + 1]);
        o[j] = act_silu(gate) * up;
      }
      uint4 pk;
      pk.x = pack_bf16x2(o[0], o[1]);
      pk.y = pack_bf16x2(o[2], o[3]);
      pk.z = pack_bf16x2(o[4], o[5]);
      pk.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(er.out_row) * p.out_ld + (col >> 1)) = pk;
    }
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = col0 + g * 8;
      if (col >= p.N) break;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
      if (p.bias != nullptr) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
        const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z),
                     b3 = unpack_bf16x2(b.w);
        v[0] += b0.x; v[1] += b0.y; v[2] += b1.x; v[3] += b1.y;
        v[4] += b2.x; v[5] += b2.y; v[6] += b3.x; v[7] += b3.y;
      }
      if constexpr (EPI == GEMM_EPI_QUICK_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_quick_gelu(v[j]);
      } else if constexpr (EPI == GEMM_EPI_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_gelu_erf(v[j]);
      }
      if (p.residual != nullptr) {
        const uint4 q = res[g];
        const float2 q0 = unpack_bf16x2(q.x), q1 = unpack_bf16x2(q.y), q2 = unpack_bf16x2(q.z),
                     q3 = unpack_bf16x2(q.w);
        v[0] += q0.x; v[1] += q0.y; v[2] += q1.x; v[3] += q1.y;
        v[4] += q2.x; v[5] += q2.y; v[6] += q3.x; v[7] += q3.y;
      }
      if (p.out_f32 != nullptr) {
        float4* dst = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(er.out_row) * p.out_ld + col);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        uint4 pk;
        pk.x = pack_bf16x2(v[0], v[1]);
        pk.y = pack_bf16x2(v[2], v[3]);
        pk.z = pack_bf16x2(v[4], v[5]);
        pk.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(er.out_row) * p.out_ld + col) = pk;
      }
    }
  }
}

// One accumulator tile (128 x BLOCK_N fp32 in TMEM at column tmem_acc) -> HBM.
//   quad : TMEM lane quadrant of this warp;  half : which half of the columns this warp covers
template <int BLOCK_N, int EPI>
SLIME_DEVINL void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int m0, int n0, int quad, int half, int lane) {
  constexpr int NCH = BLOCK_N / 64;  // 32-column chunks per warp (half of the tile's columns)
  EpiRow er;
  const int row = m0 + quad * 32 + lane;
  const bool row_ok = row < p.M;
  er.out_row = row;
  if (row_ok && p.row_map != nullptr) er.out_row = p.row_map[row];
  er.store_ok = row_ok && er.out_row >= 0;
  er.res_row = (p.res_period > 0) ? row % p.res_period : row;

  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + half * (BLOCK_N / 2);
  const int col_begin = n0 + half * (BLOCK_N / 2);

  uint32_t acc[2][32];
  uint4 res[2][4];
  tmem_ld_32x32b_x32(taddr, acc[0]);
  epi_issue_residual<EPI>(p, er, col_begin, res[0]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (i + 1 < NCH) {
      tmem_ld_32x32b_x32(taddr + (i + 1) * 32, acc[(i + 1) & 1]);
      epi_issue_residual<EPI>(p, er, col_begin + (i + 1) * 32, res[(i + 1) & 1]);
    }
    epi_process_chunk<EPI>(p, er, col_begin + i * 32, acc[i & 1], res[i & 1]);
    if (i + 1 < NCH) tmem_ld_wait();
  }
}
