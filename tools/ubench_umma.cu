// Micro-benchmark: how many cycles does one tcgen05.mma of the shapes the attention kernel issues really cost?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slime_b200/csrc -I include tools/ubench_umma.cu -o tools/bin/ubench_umma
//
// One CTA per SM (148), one issuing thread per CTA, operands resident in shared memory / TMEM (contents are
// irrelevant for timing).  For every case: `iters` groups of 8 MMAs (K = 128 = 8 x 16) are issued back to back into
// the same accumulator, one commit at the end.  Reported per MMA: cycles until the LAST ISSUE returned (issue cost /
// queue back-pressure) and cycles until the commit barrier flipped (execution).  The attention kernel
// (slime_b200/csrc/attention_tc2.cu) issues SS 128x128x16 (S = Q K^T) and TS 128xHDx16 with an MN-major B (O += P V).
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"

namespace {

constexpr int SLAB = 128 * 128;  // 128 rows x 64 bf16 (one 128-byte swizzled slab)

enum Mode { SS_K = 0, TS_MN = 1, SS_MN = 2 };

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) ubench(int iters, long long* out, int with_ld) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                // 2 slabs (K = 128)
  uint8_t* sB = smem + 2 * SLAB;     // up to 256 rows x 128 K: 4 slabs
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * SLAB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_holder);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_holder;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = MODE == SS_K ? make_idesc_bf16_major(128, N, 0, 0) : make_idesc_bf16_major(128, N, 0, 1);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (MODE == SS_K) {
          const uint64_t da = make_umma_desc_sw128(smem_u32(sA + (kk >> 2) * SLAB)) + 2 * (kk & 3);
          const uint64_t db = make_umma_desc_sw128(smem_u32(sB + (kk >> 2) * (N / 128 > 1 ? 2 * SLAB : SLAB))) + 2 * (kk & 3);
          umma_bf16_ss(tmem, da, db, idesc, 1u);
        } else if (MODE == TS_MN) {
          const uint64_t dv = make_umma_desc_mn_sw128(smem_u32(sB), SLAB) + static_cast<uint64_t>(kk * (2048 >> 4));
          umma_bf16_ts(tmem, tmem + 256 + kk * 8, dv, idesc, 1u);
        } else {
          const uint64_t da = make_umma_desc_sw128(smem_u32(sA + (kk >> 2) * SLAB)) + 2 * (kk & 3);
          const uint64_t dv = make_umma_desc_mn_sw128(smem_u32(sB), SLAB) + static_cast<uint64_t>(kk * (2048 >> 4));
          umma_bf16_ss(tmem, da, dv, idesc, 1u);
        }
      }
    }
    t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    t2 = clock64();
  } else if (with_ld && warp >= 2) {
    // optional TMEM read traffic next to the MMAs (what the softmax warps do): 2 warps re-read 64 columns in a loop
    uint32_t r[32];
    uint32_t acc = 0;
    for (int it = 0; it < iters * 2; ++it) {
      tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 384 + (it & 1) * 32, r);
      tmem_ld_wait();
      acc += r[it & 31];
    }
    if (acc == 0x12345678u) out[3] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0 && lane == 0 && blockIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// The attention kernel's MMA stream (hd 128): per kv tile g, S(g+1) = Q K^T (8 SS MMAs into S buffer (g+1) % NBUF) then
// O += P(g) V (8 TS MMAs, A = P(g) read from TMEM).  P_ALIAS: P(g) lives in the first 64 columns of S buffer g % NBUF
// (what the hd-128 path of attention_tc2.cu does) - the S MMA issued right after PV(g-1) then OVERWRITES the columns PV(g-1) reads when
// NBUF == 2.  P_ALIAS == 0: P in its own columns.  No barriers, no softmax: pure tensor-pipe stream.
template <int NBUF, int P_ALIAS>
__global__ void __launch_bounds__(128, 1) ustream(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;             // 2 slabs
  uint8_t* sK = smem + 2 * SLAB;  // 2 slabs
  uint8_t* sV = smem + 4 * SLAB;  // 2 slabs
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * SLAB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_holder);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_holder;
  long long t0 = 0, t2 = 0;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc_s = make_idesc_bf16_major(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16_major(128, 128, 0, 1);
    // columns: S buffers at 0, 128, (256); O at NBUF * 128 (hd 128); separate P at 448 (64 columns, NBUF == 2 only)
    const uint32_t o_col = NBUF * 128;
    t0 = clock64();
    for (int g = 0; g < iters; ++g) {
      const uint32_t s_next = ((g + 1) % NBUF) * 128;
      const uint32_t p_cur = P_ALIAS ? (g % NBUF) * 128 : 448;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da = make_umma_desc_sw128(smem_u32(sQ + (kk >> 2) * SLAB)) + 2 * (kk & 3);
        const uint64_t db = make_umma_desc_sw128(smem_u32(sK + (kk >> 2) * SLAB)) + 2 * (kk & 3);
        umma_bf16_ss(tmem + s_next, da, db, idesc_s, kk != 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t dv = make_umma_desc_mn_sw128(smem_u32(sV), SLAB) + static_cast<uint64_t>(kk * (2048 >> 4));
        umma_bf16_ts(tmem + o_col, tmem + p_cur + kk * 8, dv, idesc_pv, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    t2 = clock64();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0 && lane == 0 && blockIdx.x == 0) out[0] = t2 - t0;
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int NBUF, int P_ALIAS>
void run_stream(const char* name, int iters, long long* d_out) {
  auto k = ustream<NBUF, P_ALIAS>;
  const int smem = 1024 + 6 * SLAB;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[1];
  for (int rep = 0; rep < 3; ++rep) {
    k<<<148, 128, smem>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-72s %7.1f cycles per kv tile (16 MMAs, floor 1024)\n", name, static_cast<double>(h[0]) / iters);
}

template <int N, int MODE>
void run(const char* name, int iters, long long* d_out, int with_ld) {
  auto k = ubench<N, MODE>;
  const int smem = 1024 + 6 * SLAB;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2];
  for (int rep = 0; rep < 3; ++rep) {
    k<<<148, 128, smem>>>(iters, d_out, with_ld);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double n = 8.0 * iters;
  printf("%-44s %s issue %7.1f cyc/MMA   done %7.1f cyc/MMA   (floor %d)\n", name, with_ld ? "+tmem ld" : "        ", h[0] / n,
         h[1] / n, 128 * N / 256);
}

}  // namespace

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  const int iters = 512;
  for (int with_ld = 0; with_ld < 2; ++with_ld) {
    run<128, SS_K>("SS  128x128x16  (S = Q K^T, hd 128)", iters, d_out, with_ld);
    run<256, SS_K>("SS  128x256x16  (GEMM 1-CTA tile)", iters, d_out, with_ld);
    run<64, SS_K>("SS  128x64x16", iters, d_out, with_ld);
    run<128, TS_MN>("TS  128x128x16  B MN-major (O += P V, hd 128)", iters, d_out, with_ld);
    run<64, TS_MN>("TS  128x64x16   B MN-major (O += P V, hd 64)", iters, d_out, with_ld);
    run<128, SS_MN>("SS  128x128x16  B MN-major", iters, d_out, with_ld);
  }
  run_stream<2, 1>("stream: S(g+1) then PV(g), 2 S buffers, P aliases its S buffer (kernel today)", 2048, d_out);
  run_stream<3, 1>("stream: 3 S buffers, P aliases its S buffer", 2048, d_out);
  run_stream<2, 0>("stream: 2 S buffers, P in its own TMEM columns", 2048, d_out);
  return 0;
}
