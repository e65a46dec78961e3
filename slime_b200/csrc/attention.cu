// Entry of the fused prefill attention (attention.h): argument checks, then the tcgen05 kernel with two query tiles
// per CTA sharing one K/V stream (attention_tc2.cu).  (The round-1 kernel - one query tile per CTA - lost on every shape
// of the path and was removed: profiles/r02_attention_experiments.txt.)
#include <cstdlib>

#include "attention.h"
#include "errors.h"

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
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return slime_launch_attention_tc2(p, num_sms, stream);
}
