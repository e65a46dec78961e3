// Thread-local error message storage, launch accounting and the optional event profiler.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "errors.h"

static thread_local char g_err[1024] = "";

void slime_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* slime_get_error() { return g_err; }

namespace {
std::atomic<long long> g_launches{0};
std::atomic<bool> g_prof_on{false};
struct ProfRec {
  cudaEvent_t a, b;
  int cls;
  double work;
};
std::mutex g_prof_mu;
std::vector<ProfRec> g_recs;
ProfRec g_open;
bool g_has_open = false;
}  // namespace

namespace {
int g_pdl_mode = -1;  // -1 unset, 0 off, 1 on
}
bool slime_pdl_enabled() {
  if (g_pdl_mode < 0) {
    const char* e = getenv("SLIME_PDL");
    g_pdl_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pdl_mode != 0;
}
namespace {
int g_prefill_pdl = -1;
}
bool slime_prefill_pdl_enabled() {
  if (g_prefill_pdl < 0) {
    const char* e = getenv("SLIME_PREFILL_PDL");
    g_prefill_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;  // on by default: bit-identical, batch-1 latency -4 %
  }
  return g_prefill_pdl != 0 && slime_pdl_enabled();
}
extern "C" int slime_set_prefill_pdl(int mode) {
  g_prefill_pdl = mode < 0 ? -1 : (mode != 0 ? 1 : 0);
  return SLIME_OK;
}
void slime_carveout_once(const void* kernel) {
  static std::mutex mu;
  static std::vector<const void*> seen;
  static int pct = -2;
  std::lock_guard<std::mutex> lk(mu);
  if (pct == -2) {  // SLIME_CARVEOUT_PCT: 0..100 = preferred shared-memory share for every kernel of the chain, -1 = driver's choice
    const char* e = getenv("SLIME_CARVEOUT_PCT");
    pct = (e != nullptr && (e[0] == '-' || (e[0] >= '0' && e[0] <= '9'))) ? atoi(e) : 44;
  }
  if (pct < 0) return;  // (default 44 = the 100 KB configuration: measured best, profiles/r01_decode_bench.txt)
  for (const void* k : seen)
    if (k == kernel) return;
  seen.push_back(kernel);
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
  cudaGetLastError();  // a preference only: never fatal
}
extern "C" int slime_set_pdl_mode(int mode) {
  g_pdl_mode = mode < 0 ? -1 : (mode != 0 ? 1 : 0);
  return SLIME_OK;
}

void slime_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool slime_prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void slime_prof_begin(int cls, double work, cudaStream_t stream) {
  if (!slime_prof_enabled()) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.cls = cls;
  r.work = work;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, stream);
  g_open = r;
  g_has_open = true;
}
void slime_prof_end(cudaStream_t stream) {
  if (!slime_prof_enabled()) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_has_open) return;
  cudaEventRecord(g_open.b, stream);
  g_recs.push_back(g_open);
  g_has_open = false;
}

extern "C" {

long long slime_launch_count(void) { return g_launches.load(); }

int slime_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  g_has_open = false;
  g_prof_on.store(on != 0);
  return SLIME_OK;
}

// Sums per class: ms[3], work[3], launches[3].  Synchronises on the recorded events.
int slime_profile_collect(double* ms, double* work, long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < 3; ++i) {
    ms[i] = 0.0;
    work[i] = 0.0;
    launches[i] = 0;
  }
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    const int c = r.cls < 0 || r.cls > 2 ? 2 : r.cls;
    ms[c] += t;
    work[c] += r.work;
    launches[c] += 1;
  }
  return SLIME_OK;
}

}  // extern "C"
