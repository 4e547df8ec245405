// hs_service: the serving front of the engine.
//
// The reference serves one query per HTTP request: every /query handler thread calls
// hnsw_slim.searchKnn(vec, k, out) on its own (hnsw_slim_server.cc:69-98, hnsw_slim_server_patch.cc:133-160),
// /setEf calls setEf (:100-115) and the update path swaps neighbour lists under the searches' feet
// (patchFromStream, slim.h:2206-2388).  A GPU answers a BATCH per launch, so the handler's one call becomes:
// copy the query into the batch that is currently filling, sleep, wake up with the answer.  A dispatcher
// thread launches the filling batch as soon as the previous one has completed (or when it is full / has waited
// max_wait_us): while batch A runs on the GPU the requests that arrive collect in batch B — the batch size
// adapts to the load by itself, an idle server answers a lone query at once and a busy one fills its batches.
// The two batches live in page-locked, mapped host memory: hs_search_batch reads the queries and writes the
// result rows in place (no staging copies).
//
// Same-k batching: a launch has one k, and ef = max(ef_, k) (slim.h:2080) depends on it, so a batch only takes
// requests with the k of its first request; a request with another k waits for the next batch.
// hs_service_set_ef takes effect with the next batch; hs_service_patch drains both batches, applies the patch
// (hs_patch_apply) and lets the requests continue — what the reference does NOT guarantee (its patchFromStream
// runs unsynchronised with searchKnn) is guaranteed here: a query sees the index before or after a patch, never
// in between.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "hs_index.h"

using namespace hs;

namespace {

using Clock = std::chrono::steady_clock;

enum class State { Filling, Running, Draining };

struct Batch {
  float *q = nullptr;
  uint32_t *lab = nullptr;
  float *dist = nullptr;
  size_t count = 0, k = 0, readers = 0;
  State state = State::Filling;
  unsigned long long ticket = 0;       // bumped every time the batch is handed to the backend
  int rc = HS_OK;
  std::string err;
  Clock::time_point t0;
};

}  // namespace

struct hs_service {
  // backend: an hs_index, or (tests) a callback with hs_search_batch's meaning
  hs_index *ix = nullptr;
  hs_service_backend_fn fn = nullptr;
  void *fn_ctx = nullptr;
  size_t dim = 0, max_batch = 0, k_max = 0;
  unsigned max_wait_us = 0;
  bool pinned = false;

  std::mutex mu;
  std::condition_variable cv_join, cv_work, cv_done, cv_idle;
  Batch b[2];
  int fill = 0;
  bool closing = false, paused = false;
  size_t pending_ef = 0;
  std::thread dispatcher;
  hs_service_stats st{};

  void run();
  bool idle() const { return b[0].state == State::Filling && b[1].state == State::Filling && b[0].count == 0 && b[1].count == 0; }
};

void hs_service::run() {
  std::unique_lock<std::mutex> lk(mu);
  auto ready = [&] { return b[fill].state == State::Filling && b[fill].count > 0; };
  for (;;) {
    cv_work.wait(lk, [&] { return closing || ready(); });
    if (!ready()) {
      if (closing) return;
      continue;
    }
    Batch &cur = b[fill];
    // the previous batch has completed (this thread ran it synchronously): launch at once unless the caller asked
    // for a minimum collection window and the batch is neither full nor held up by a patch
    if (max_wait_us > 0) {
      const auto deadline = cur.t0 + std::chrono::microseconds(max_wait_us);
      cv_work.wait_until(lk, deadline, [&] { return closing || paused || cur.count >= max_batch; });
    }
    const int me = fill;
    cur.state = State::Running;
    cur.ticket++;
    fill ^= 1;                          // requests now collect in the other batch (once its readers are done)
    cv_join.notify_all();
    const size_t count = cur.count, k = cur.k, ef = pending_ef;
    pending_ef = 0;
    lk.unlock();
    int rc = HS_OK;
    std::string err;
    const auto t0 = Clock::now();
    if (ix) {
      if (ef) rc = hs_set_ef(ix, ef);
      if (rc == HS_OK) rc = hs_search_batch(ix, cur.q, count, k, cur.lab, cur.dist);
      if (rc != HS_OK) err = hs_last_error();
    } else {
      rc = fn(fn_ctx, cur.q, count, k, cur.lab, cur.dist);
      if (rc != HS_OK) err = "backend callback failed";
    }
    const double busy = std::chrono::duration<double>(Clock::now() - t0).count();
    lk.lock();
    Batch &done = b[me];
    done.rc = rc;
    done.err = err;
    done.state = State::Draining;
    done.readers = count;
    st.batches++;
    st.queries += count;
    st.max_batch = std::max<uint64_t>(st.max_batch, count);
    st.busy_seconds += busy;
    cv_done.notify_all();
  }
}

extern "C" {

static int service_create(hs_index *ix, hs_service_backend_fn fn, void *ctx, size_t dim, size_t max_batch,
                          unsigned max_wait_us, size_t k_max, hs_service **out) {
  if (!out || (!ix && !fn) || dim == 0 || max_batch == 0 || k_max == 0 || max_batch > (1u << 20) || k_max > 4096) {
    set_error("hs_service_create: bad argument");
    return HS_ERR_ARG;
  }
  *out = nullptr;
  hs_service *s = nullptr;
  try {
    s = new hs_service;
  } catch (const std::bad_alloc &) {
    set_error("hs_service_create: out of host memory");
    return HS_ERR_NOMEM;
  }
  s->ix = ix;
  s->fn = fn;
  s->fn_ctx = ctx;
  s->dim = dim;
  s->max_batch = max_batch;
  s->max_wait_us = max_wait_us;
  s->k_max = k_max;
  const size_t qb = max_batch * dim * sizeof(float), rb = max_batch * k_max * 4;
  bool ok = true;
  if (ix) {
    // page-locked + mapped: hs_search_batch uses the buffers in place
    ok = cudaSetDevice(ix->device) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
      ok = cudaHostAlloc(reinterpret_cast<void **>(&s->b[i].q), qb, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
           cudaHostAlloc(reinterpret_cast<void **>(&s->b[i].lab), rb, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
           cudaHostAlloc(reinterpret_cast<void **>(&s->b[i].dist), rb, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
    }
    s->pinned = true;
    if (!ok) {
      set_error(std::string("hs_service_create: cudaHostAlloc: ") + cudaGetErrorString(cudaGetLastError()));
    }
  } else {
    for (int i = 0; i < 2 && ok; ++i) {
      s->b[i].q = static_cast<float *>(std::malloc(qb));
      s->b[i].lab = static_cast<uint32_t *>(std::malloc(rb));
      s->b[i].dist = static_cast<float *>(std::malloc(rb));
      ok = s->b[i].q && s->b[i].lab && s->b[i].dist;
    }
    if (!ok) set_error("hs_service_create: out of host memory");
  }
  if (!ok) {
    hs_service_free(s);
    return ix ? HS_ERR_CUDA : HS_ERR_NOMEM;
  }
  try {
    s->dispatcher = std::thread([s] { s->run(); });
  } catch (const std::exception &e) {
    set_error(std::string("hs_service_create: cannot start the dispatcher thread: ") + e.what());
    hs_service_free(s);
    return HS_ERR_NOMEM;
  }
  *out = s;
  return HS_OK;
}

int hs_service_create(hs_index *ix, size_t max_batch, unsigned max_wait_us, size_t k_max, hs_service **out) {
  if (!ix) {
    set_error("hs_service_create: null index");
    return HS_ERR_ARG;
  }
  return service_create(ix, nullptr, nullptr, ix->info.dim, max_batch, max_wait_us, k_max, out);
}

int hs_debug_service_create(hs_service_backend_fn fn, void *ctx, size_t dim, size_t max_batch, unsigned max_wait_us,
                            size_t k_max, hs_service **out) {
  return service_create(nullptr, fn, ctx, dim, max_batch, max_wait_us, k_max, out);
}

int hs_service_query(hs_service *s, const float *vec, size_t k, uint32_t *labels_out, float *dists_out) {
  if (!s || !vec || !labels_out || k == 0 || k > s->k_max) {
    set_error("hs_service_query: null argument or k out of range");
    return HS_ERR_ARG;
  }
  std::unique_lock<std::mutex> lk(s->mu);
  Batch *bt = nullptr;
  for (;;) {
    if (s->closing) {
      set_error("hs_service_query: the service is shutting down");
      return HS_ERR_ARG;
    }
    bt = &s->b[s->fill];
    if (!s->paused && bt->state == State::Filling && bt->count < s->max_batch && (bt->count == 0 || bt->k == k)) break;
    if (!s->paused && bt->state == State::Filling && bt->count > 0) s->cv_work.notify_one();   // full, or another k: flush it
    s->cv_join.wait(lk);
  }
  const size_t slot = bt->count++;
  if (slot == 0) {
    bt->k = k;
    bt->t0 = Clock::now();
  }
  std::memcpy(bt->q + slot * s->dim, vec, s->dim * sizeof(float));
  const unsigned long long want = bt->ticket + 1;
  if (slot == 0 || bt->count == s->max_batch) s->cv_work.notify_one();
  s->cv_done.wait(lk, [&] { return bt->ticket >= want && bt->state == State::Draining; });
  const int rc = bt->rc;
  if (rc == HS_OK) {
    std::memcpy(labels_out, bt->lab + slot * k, k * 4);
    if (dists_out) std::memcpy(dists_out, bt->dist + slot * k, k * 4);
  } else {
    set_error(bt->err);
  }
  if (--bt->readers == 0) {             // last reader: the batch may fill again
    bt->count = 0;
    bt->state = State::Filling;
    s->cv_join.notify_all();
    s->cv_work.notify_one();
    if (s->idle()) s->cv_idle.notify_all();
  }
  return rc;
}

int hs_service_set_ef(hs_service *s, size_t ef) {
  if (!s || ef == 0 || ef > 4096) {
    set_error("hs_service_set_ef: ef must be in [1, 4096]");
    return HS_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(s->mu);
  s->pending_ef = ef;                   // applied by the dispatcher before the next launch
  return HS_OK;
}

int hs_service_patch(hs_service *s, const void *patch, size_t patch_bytes, unsigned flags, const float *rows,
                     const uint64_t *row_labels, size_t n_rows, hs_patch_info *info_out) {
  if (!s || !s->ix) {
    set_error("hs_service_patch: no index behind this service");
    return HS_ERR_ARG;
  }
  std::unique_lock<std::mutex> lk(s->mu);
  if (s->paused) {
    set_error("hs_service_patch: another patch is being applied");
    return HS_ERR_ARG;
  }
  s->paused = true;                     // no request joins a batch; what has joined is answered first
  s->cv_work.notify_one();
  s->cv_idle.wait(lk, [&] { return s->idle(); });
  if (s->pending_ef) {
    hs_set_ef(s->ix, s->pending_ef);
    s->pending_ef = 0;
  }
  lk.unlock();
  const int rc = hs_patch_apply(s->ix, patch, patch_bytes, flags, rows, row_labels, n_rows, info_out);
  lk.lock();
  s->paused = false;
  if (rc == HS_OK) s->st.patches++;
  s->cv_join.notify_all();
  return rc;
}

int hs_service_get_stats(hs_service *s, hs_service_stats *out) {
  if (!s || !out) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(s->mu);
  *out = s->st;
  return HS_OK;
}

void hs_service_free(hs_service *s) {
  if (!s) return;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->closing = true;
    s->cv_work.notify_all();
    s->cv_join.notify_all();
  }
  if (s->dispatcher.joinable()) s->dispatcher.join();
  for (auto &bt : s->b) {
    if (s->pinned) {
      cudaFreeHost(bt.q);
      cudaFreeHost(bt.lab);
      cudaFreeHost(bt.dist);
    } else {
      std::free(bt.q);
      std::free(bt.lab);
      std::free(bt.dist);
    }
  }
  delete s;
}

}  // extern "C"
