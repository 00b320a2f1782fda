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
std::atomic<int> g_uploads{0};  // host -> device stagings
std::atomic<int> g_clones{0};   // device -> device replicas
int g_fail_ordinal = -1;  // the context with this ordinal fails its next align
thread_local std::string g_err;
unsigned long long stub_thread_id() { return static_cast<unsigned long long>(std::hash<std::thread::id>{}(std::this_thread::get_id())); }
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
// the halves of the two setters and the replicas (api.cu): staged -> built; a clone needs a staged original
PEB_API int peb_target_stage(peb_ctx* c, const void*, size_t n, size_t, const void*, size_t) {
  c->n_tgt = n;
  c->tgt_staged = true;
  c->launches += 1;
  ++g_uploads;
  return PEB_OK;
}
PEB_API int peb_target_build(peb_ctx* c) {
  if (!c->tgt_staged) {
    c->err = "stub: nothing staged";
    return PEB_E_NO_TARGET;
  }
  c->tgt_grid.valid = true;
  return PEB_OK;
}
PEB_API int peb_target_clone(peb_ctx* dst, peb_ctx* src) {
  if (!src->tgt_staged) {
    dst->err = "stub: the original has nothing staged";
    return PEB_E_NO_TARGET;
  }
  dst->n_tgt = src->n_tgt;
  dst->tgt_staged = true;
  dst->tgt_grid.valid = true;
  dst->launches += 1;
  ++g_clones;
  return PEB_OK;
}
PEB_API int peb_source_stage(peb_ctx* c, const void*, size_t n, size_t) {
  c->n_src = n;
  c->src_staged = true;
  c->launches += 1;
  ++g_uploads;
  return PEB_OK;
}
PEB_API int peb_source_build(peb_ctx* c) {
  if (!c->src_staged) {
    c->err = "stub: nothing staged";
    return PEB_E_NO_SOURCE;
  }
  c->src_set = true;
  return PEB_OK;
}
PEB_API int peb_source_clone(peb_ctx* dst, peb_ctx* src) {
  if (!src->src_staged) {
    dst->err = "stub: the original has nothing staged";
    return PEB_E_NO_SOURCE;
  }
  dst->n_src = src->n_src;
  dst->src_staged = true;
  dst->src_set = true;
  dst->launches += 1;
  ++g_clones;
  return PEB_OK;
}
PEB_API int peb_ctx_enable_peer(peb_ctx*, const peb_ctx*) { return PEB_OK; }
PEB_API int peb_icp_align_batch(peb_ctx* c, const float* guesses, size_t n, const peb_icp_params* p, peb_icp_result* results) {
  const int now = ++g_concurrent;
  int seen = g_max_concurrent.load();
  while (now > seen && !g_max_concurrent.compare_exchange_weak(seen, now)) {
  }
  std::this_thread::sleep_for(std::chrono::milliseconds(30));  // long enough for the shards to overlap
  int rc = PEB_OK;
  c->coop_max_rows = static_cast<int>(stub_thread_id() & 0x7FFFFFFF);  // which host thread drove this context (a field the stub does not otherwise use)
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
PEB_API int stub_uploads() { return g_uploads.exchange(0); }
PEB_API int stub_clones() { return g_clones.exchange(0); }
PEB_API int stub_live_contexts() { return g_live.load(); }
PEB_API int stub_max_concurrent() { return g_max_concurrent.exchange(0); }
PEB_API void stub_fail_ordinal(int k) { g_fail_ordinal = k; }
PEB_API void stub_reset_ordinals() { g_created = 0; }
PEB_API int stub_ctx_batch_streams(peb_ctx* c) { return c->batch_streams; }
PEB_API int stub_ctx_thread(peb_ctx* c) { return c->coop_max_rows; }
PEB_API int stub_ctx_ready(peb_ctx* c) { return (c->tgt_grid.valid ? 1 : 0) + (c->src_set ? 2 : 0); }

}  // extern "C"
