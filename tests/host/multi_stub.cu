// TEST SHIM: pose_estimation_b200/csrc/multi.cu (the peb_multi_* host layer) linked against FAKE single-device entry
// points, so that its sharding, threading and error propagation can be exercised without a GPU.  The fakes do no
// arithmetic of the path: a "context" records what it was handed, and peb_icp_align_batch stamps every record with
// (context ordinal, first float of the hypothesis' guess) so the test can see who refined what.
// Nothing here is linked into libpe_b200.so.
#include <atomic>
#include <chrono>
#include <thread>

#include "../../pose_estimation_b200/csrc/multi.cu"

namespace {
std::atomic<int> g_live{0};
std::atomic<int> g_created{0};
std::atomic<int> g_concurrent{0};
std::atomic<int> g_max_concurrent{0};
int g_fail_ordinal = -1;  // the context with this ordinal fails its next align
thread_local std::string g_err;
}  // namespace

extern "C" {

PEB_API int peb_ctx_create(int device, peb_ctx** out) {
  if (device < 0 || device >= 8) {
    g_err = "stub: device index out of range";
    return PEB_E_INVALID_ARG;
  }
  peb_ctx* c = new peb_ctx();
  c->device = device;
  c->launches = 0;
  c->nn_group = g_created++;  // the ordinal, kept in a field the stub does not otherwise use
  ++g_live;
  *out = c;
  return PEB_OK;
}
PEB_API void peb_ctx_destroy(peb_ctx* c) {
  if (!c) return;
  --g_live;
  delete c;
}
PEB_API const char* peb_last_error(const peb_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
PEB_API uint64_t peb_ctx_launch_count(const peb_ctx* c) { return c ? c->launches : 0; }
PEB_API int peb_ctx_set_int(peb_ctx* c, const char* key, int value) {
  if (std::string(key) != "batch_streams") {
    c->err = std::string("unknown option '") + key + "'";
    return PEB_E_INVALID_ARG;
  }
  c->batch_streams = value;
  return PEB_OK;
}
PEB_API int peb_target_set(peb_ctx* c, const void*, size_t n, size_t, const void*, size_t) {
  c->n_tgt = n;
  c->launches += 1;
  return PEB_OK;
}
PEB_API int peb_source_set(peb_ctx* c, const void*, size_t n, size_t) {
  c->n_src = n;
  c->launches += 1;
  return PEB_OK;
}
PEB_API int peb_icp_align_batch(peb_ctx* c, const float* guesses, size_t n, const peb_icp_params* p, peb_icp_result* results) {
  const int now = ++g_concurrent;
  int seen = g_max_concurrent.load();
  while (now > seen && !g_max_concurrent.compare_exchange_weak(seen, now)) {
  }
  std::this_thread::sleep_for(std::chrono::milliseconds(30));  // long enough for the shards to overlap
  int rc = PEB_OK;
  if (c->nn_group == g_fail_ordinal) {
    c->err = "stub: injected failure";
    rc = PEB_E_CUDA;
  } else {
    for (size_t h = 0; h < n; ++h) {
      results[h] = peb_icp_result{};
      results[h].T[0] = guesses[16 * h];
      results[h].iterations = p->max_iterations;
      results[h].state = c->nn_group;          // who
      results[h].n_correspondences = c->device;  // where
      results[h].fitness = static_cast<double>(c->n_src) + 1e-3 * static_cast<double>(c->n_tgt);
    }
    c->launches += n;
  }
  --g_concurrent;
  return rc;
}

// test controls
PEB_API int stub_live_contexts() { return g_live.load(); }
PEB_API int stub_max_concurrent() { return g_max_concurrent.exchange(0); }
PEB_API void stub_fail_ordinal(int k) { g_fail_ordinal = k; }
PEB_API void stub_reset_ordinals() { g_created = 0; }
PEB_API int stub_ctx_batch_streams(peb_ctx* c) { return c->batch_streams; }

}  // extern "C"
