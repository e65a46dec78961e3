// The path's one collective behind the C-ABI (include/slime_b200.h): the all-gather of the
// [B_local, V] fp32 last-token logits over NCCL (NVLink 5 / NVSwitch).  The reference has no collective at all - its
// multi-GPU story is N independent processes whose outputs are concatenated (scripts/llama/eval/gqa.sh:20-43).
//
// NCCL is bound at RUN TIME (dlopen of libnccl.so.2): the library has no link-time dependency on it, a single-GPU user
// never loads it, and inside a PyTorch process the already-loaded NCCL (torch's bundled build) is the one that is found.
// Bootstrap: rank 0 calls slime_comm_unique_id, ships the 128 bytes to the other ranks by any side channel
// (slime_b200/parallel.py uses the torch.distributed store), every rank calls slime_comm_init.
#include <dlfcn.h>

#include <cstring>

#include "errors.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;     // ncclSuccess == 0
constexpr int kNcclFloat32 = 7;  // ncclDataType_t: ncclFloat32 (nccl.h)

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle != nullptr) return SLIME_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    slime_set_error("comm: cannot load libnccl.so.2 (%s)", dlerror());
    return SLIME_ESTATE;
  }
#define BIND(field, sym)                                                       \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, sym));      \
  if (g_nccl.field == nullptr) {                                               \
    slime_set_error("comm: libnccl has no symbol %s", sym);                    \
    return SLIME_ESTATE;                                                       \
  }
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(AllGather, "ncclAllGather");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(GetErrorString, "ncclGetErrorString");
  BIND(GetVersion, "ncclGetVersion");
#undef BIND
  g_nccl.handle = h;
  return SLIME_OK;
}

#define SLIME_CHECK_NCCL(expr)                                                                   \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != 0) {                                                                               \
      slime_set_error("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return SLIME_ECUDA;                                                                        \
    }                                                                                            \
  } while (0)

}  // namespace

struct slime_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

extern "C" {

int slime_comm_unique_id(void* out_128_bytes) {
  SLIME_REQUIRE(out_128_bytes != nullptr, "comm_unique_id: null buffer");
  SLIME_PROPAGATE(load_nccl());
  ncclUniqueId id;
  SLIME_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(out_128_bytes, id.internal, sizeof(id.internal));
  return SLIME_OK;
}

int slime_comm_init(slime_comm** out, const void* id_128_bytes, int rank, int world) {
  SLIME_REQUIRE(out != nullptr && id_128_bytes != nullptr, "comm_init: null argument");
  SLIME_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank %d of %d", rank, world);
  SLIME_PROPAGATE(load_nccl());
  ncclUniqueId id;
  std::memcpy(id.internal, id_128_bytes, sizeof(id.internal));
  slime_comm* c = new slime_comm();
  c->rank = rank;
  c->world = world;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);  // uses the calling thread's current CUDA device
  if (r != 0) {
    slime_set_error("comm_init: ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return SLIME_ECUDA;
  }
  *out = c;
  return SLIME_OK;
}

int slime_comm_nccl_version(void) {
  if (load_nccl() != SLIME_OK) return -1;
  int v = 0;
  return g_nccl.GetVersion(&v) == 0 ? v : -1;
}

// out [world * rows_local, vocab] fp32 <- every rank's local [rows_local, vocab] block, in rank order; enqueued on `stream`.
int slime_allgather_logits(slime_comm* comm, const float* local_logits, float* out, int rows_local, int vocab, void* stream) {
  SLIME_REQUIRE(comm != nullptr && comm->comm != nullptr && local_logits != nullptr && out != nullptr,
                "allgather_logits: null argument");
  SLIME_REQUIRE(rows_local > 0 && vocab > 0, "allgather_logits: empty block %d x %d", rows_local, vocab);
  SLIME_CHECK_NCCL(g_nccl.AllGather(local_logits, out, static_cast<size_t>(rows_local) * vocab, kNcclFloat32, comm->comm,
                                    static_cast<cudaStream_t>(stream)));
  slime_note_launch();
  return SLIME_OK;
}

void slime_comm_destroy(slime_comm* comm) {
  if (comm == nullptr) return;
  if (comm->comm != nullptr && g_nccl.CommDestroy != nullptr) g_nccl.CommDestroy(comm->comm);
  delete comm;
}

}  // extern "C"
