// multi.cu — several B200s behind one handle: the peb_multi_* entry points (include/pe_b200.h,
// SURVEY.md 8b layer 2 / 8e).  The reference node is ONE process (a component container with a
// single-threaded executor, launch/pose_estimation.launch.py:17-35), so a drop-in that wants all the
// GPUs of the box cannot rely on one process per GPU: this file gives that process one context per
// device, a replica of the scene grid and of the model on each, and shards the H initial poses of
// registerModelToScene(model, scene, poses) (opencv_surface_match.cpp:94) into contiguous blocks,
// one per device — the same rule as pose_estimation_b200/multi.py : shard_range.  Hypotheses are
// independent: no collective; every device copies its block of result records straight into the
// caller's `results` array.
//
// Host code only (no kernel of its own): one short-lived host thread per extra device drives the
// blocking single-device entry points of api.cu, so every device runs the tested path unchanged.
#include <algorithm>
#include <exception>
#include <thread>

#include "common.cuh"

struct peb_multi {
  std::vector<peb_ctx*> ctx;
  std::string err;
};

namespace {

thread_local std::string g_multi_create_error;

int multi_fail(peb_multi* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  return code;
}

// fn(i) on every context, device 0's on the calling thread; the first failing device (lowest index) is reported
template <typename Fn>
int on_all_devices(peb_multi* m, const char* what, Fn fn) {
  const size_t n = m->ctx.size();
  std::vector<int> rc(n, PEB_OK);
  try {
    std::vector<std::thread> workers;
    workers.reserve(n);
    try {
      for (size_t i = 1; i < n; ++i) workers.emplace_back([&rc, &fn, i]() { rc[i] = fn(i); });
    } catch (...) {
      for (std::thread& t : workers) t.join();
      throw;
    }
    rc[0] = fn(0);
    for (std::thread& t : workers) t.join();
  } catch (const std::exception& e) {
    return multi_fail(m, PEB_E_OOM, std::string(what) + ": " + e.what());
  } catch (...) {
    return multi_fail(m, PEB_E_OOM, std::string(what) + ": could not start the per-device host threads");
  }
  for (size_t i = 0; i < n; ++i)
    if (rc[i] != PEB_OK)
      return multi_fail(m, rc[i], std::string(what) + " [context " + std::to_string(i) + ", device " +
                                      std::to_string(m->ctx[i]->device) + "]: " + peb_last_error(m->ctx[i]));
  return PEB_OK;
}

}  // namespace

extern "C" {

PEB_API int peb_multi_create(int ndev, const int* devices, peb_multi** out) {
  if (!out) return PEB_E_INVALID_ARG;
  *out = nullptr;
  if (ndev < 1 || ndev > 64) {
    g_multi_create_error = "peb_multi_create: ndev must be in [1, 64]";
    return PEB_E_INVALID_ARG;
  }
  peb_multi* m = nullptr;
  try {
    m = new peb_multi();
    m->ctx.reserve(static_cast<size_t>(ndev));
  } catch (...) {
    delete m;
    g_multi_create_error = "peb_multi_create: out of host memory";
    return PEB_E_OOM;
  }
  for (int i = 0; i < ndev; ++i) {
    peb_ctx* c = nullptr;
    const int dev = devices ? devices[i] : i;
    const int rc = peb_ctx_create(dev, &c);
    if (rc != PEB_OK) {
      g_multi_create_error = std::string("peb_multi_create [context ") + std::to_string(i) + ", device " + std::to_string(dev) +
                             "]: " + peb_last_error(nullptr);
      peb_multi_destroy(m);
      return rc;
    }
    m->ctx.push_back(c);
  }
  *out = m;
  return PEB_OK;
}

PEB_API void peb_multi_destroy(peb_multi* m) {
  if (!m) return;
  for (peb_ctx* c : m->ctx) peb_ctx_destroy(c);
  delete m;
}

PEB_API const char* peb_multi_last_error(const peb_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
PEB_API int peb_multi_size(const peb_multi* m) { return m ? static_cast<int>(m->ctx.size()) : 0; }
PEB_API peb_ctx* peb_multi_ctx(peb_multi* m, int i) {
  return (m && i >= 0 && static_cast<size_t>(i) < m->ctx.size()) ? m->ctx[static_cast<size_t>(i)] : nullptr;
}

PEB_API int peb_multi_set_int(peb_multi* m, const char* key, int value) {
  if (!m || !key) return PEB_E_INVALID_ARG;
  for (size_t i = 0; i < m->ctx.size(); ++i) {
    const int rc = peb_ctx_set_int(m->ctx[i], key, value);
    if (rc != PEB_OK) return multi_fail(m, rc, peb_last_error(m->ctx[i]));
  }
  return PEB_OK;
}

PEB_API void peb_multi_shard_range(size_t n_items, int ndev, int i, size_t* lo, size_t* hi) {
  size_t a = 0, b = 0;
  if (ndev >= 1 && i >= 0 && i < ndev) {
    const size_t per = (n_items + static_cast<size_t>(ndev) - 1) / static_cast<size_t>(ndev);
    a = std::min(static_cast<size_t>(i) * per, n_items);
    b = std::min((static_cast<size_t>(i) + 1) * per, n_items);
  }
  if (lo) *lo = a;
  if (hi) *hi = b;
}

// every device builds its grid from the same host buffer: deterministic, so the replicas are identical
PEB_API int peb_multi_target_set(peb_multi* m, const void* pts, size_t n, size_t stride, const void* normals, size_t nstride) {
  if (!m) return PEB_E_INVALID_ARG;
  return on_all_devices(m, "peb_multi_target_set",
                        [&](size_t i) { return peb_target_set(m->ctx[i], pts, n, stride, normals, nstride); });
}

PEB_API int peb_multi_source_set(peb_multi* m, const void* pts, size_t n, size_t stride) {
  if (!m) return PEB_E_INVALID_ARG;
  return on_all_devices(m, "peb_multi_source_set", [&](size_t i) { return peb_source_set(m->ctx[i], pts, n, stride); });
}

PEB_API int peb_multi_icp_align_batch(peb_multi* m, const float* guesses, size_t n_guesses, const peb_icp_params* params,
                                      peb_icp_result* results) {
  if (!m || !params) return PEB_E_INVALID_ARG;
  if (n_guesses == 0) return PEB_OK;
  if (!guesses || !results) return multi_fail(m, PEB_E_INVALID_ARG, "peb_multi_icp_align_batch: null guesses / results");
  const int ndev = static_cast<int>(m->ctx.size());
  return on_all_devices(m, "peb_multi_icp_align_batch", [&](size_t i) {
    size_t lo = 0, hi = 0;
    peb_multi_shard_range(n_guesses, ndev, static_cast<int>(i), &lo, &hi);
    if (hi == lo) return static_cast<int>(PEB_OK);  // more devices than hypotheses
    return peb_icp_align_batch(m->ctx[i], guesses + 16 * lo, hi - lo, params, results + lo);
  });
}

PEB_API uint64_t peb_multi_launch_count(const peb_multi* m) {
  uint64_t total = 0;
  if (m)
    for (const peb_ctx* c : m->ctx) total += peb_ctx_launch_count(c);
  return total;
}

}  // extern "C"
