// hs_service: the serving front of the engine.
//
// The reference serves one query per HTTP request: every /query handler thread calls
// hnsw_slim.searchKnn(vec, k, out) on its own (hnsw_slim_server.cc:69-98, hnsw_slim_server_patch.cc:133-160),
// /setEf calls setEf (:100-115) and the update path swaps neighbour lists under the searches' feet
// (patchFromStream, slim.h:2206-2388).  A GPU answers a BATCH per launch, so the handler's one call becomes:
// copy the query into the batch that is currently filling, sleep, wake up with the answer.
//
// Batches live in a ring of 32 page-locked, mapped host buffers (hs_search_batch_submit reads the queries
// and writes the result rows in place, no staging copies).  Two threads drive the ring:
//   dispatcher  hands the filling batch to hs_search_batch_submit as soon as it holds a query (or, with
//               max_wait_us > 0, once it is full / its first query has waited that long) and moves the fill
//               point to the next free buffer.  Launches go to the handle's stream with batch overlap on
//               (hs_set_overlap), so the small batches of a lightly loaded server run side by side instead of
//               queueing behind each other: a request's latency is its own batch's.
//   completer   waits for the batches in submission order (hs_search_batch_wait_oldest) and wakes their callers.
// When every buffer of the ring is in flight, arrivals wait for the next free one and then share it: the batch
// size follows the load by itself — a lone query is answered at once, a busy server fills its batches.
//
// Same-k batching: a launch has one k, and ef = max(ef_, k) (slim.h:2080) depends on it, so a batch only takes
// requests with the k of its first request; a request with another k goes to the next buffer.
// hs_service_set_ef takes effect with the next launch; hs_service_patch lets the ring drain, applies the patch
// (hs_patch_apply) and lets the requests continue — what the reference does NOT guarantee (its patchFromStream
// runs unsynchronised with searchKnn) is guaranteed here: a query sees the index before or after a patch, never
// in between.  The service owns the handle's submit queue: no other caller may use hs_search_batch_submit /
// _wait* on the index while the service exists.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

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
  std::condition_variable done;        // the callers of THIS batch sleep here: a completion wakes them, not everybody
};

// Ring buffers: at most kDepth batches between "filling" and "answered" (below hs_index::kEventRing).  A closed loop
// of T clients keeps T queries in flight; the more batches they may spread over, the closer a request's latency is
// to the ~0.3 ms of its own traversal (small launches run side by side: the GPU holds 3552 query warps at once).
constexpr int kDepth = 32;
static_assert(kDepth < hs_index::kEventRing, "the service must not fill the handle's event ring");
// ... but a launch costs the host tens of microseconds (submit, event, completion hand-over), so a batch of one or
// two queries is only worth launching while few batches are in flight: beyond kEagerDepth of them a batch waits
// until it holds kEagerBatch queries (or a batch in flight completes).  Few clients thus get ring-of-8 behaviour
// (their batches grow while the ring is busy), many clients the deep ring.
constexpr int kEagerDepth = 8;
constexpr size_t kEagerBatch = 8;

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
  std::condition_variable cv_join, cv_work, cv_flight, cv_idle;
  Batch b[kDepth];
  int fill = 0;
  std::deque<int> inflight;            // submitted, not yet completed — in submission order
  bool closing = false, closing_completer = false, paused = false;
  size_t pending_ef = 0;
  std::thread dispatcher, completer;
  hs_service_stats st{};

  void dispatch_loop();
  void complete_loop();
  bool ready() const { return b[fill].state == State::Filling && b[fill].count > 0; }
  bool idle() const {
    if (!inflight.empty()) return false;
    for (const Batch &x : b)
      if (x.state != State::Filling || x.count != 0) return false;
    return true;
  }
  // the fill point moves to a free buffer, if there is one (else arrivals wait for the next to free up)
  void advance_fill() {
    for (int i = 1; i <= kDepth; ++i) {
      const int c = (fill + i) % kDepth;
      if (b[c].state == State::Filling && b[c].count == 0) {
        fill = c;
        return;
      }
    }
  }
  void finish(int me, int rc, const std::string &err, double seconds) {     // mu held
    Batch &done = b[me];
    done.rc = rc;
    done.err = err;
    done.state = State::Draining;
    done.readers = done.count;
    st.batches++;
    st.queries += done.count;
    st.max_batch = std::max<uint64_t>(st.max_batch, done.count);
    st.busy_seconds += seconds;
    done.done.notify_all();
    cv_work.notify_one();               // a batch left the ring: the dispatcher may launch a small one again
  }
};

void hs_service::dispatch_loop() {
  std::unique_lock<std::mutex> lk(mu);
  for (;;) {
    cv_work.wait(lk, [&] { return closing || ready(); });
    if (!ready()) {
      if (closing) return;
      continue;
    }
    Batch &cur = b[fill];
    if (max_wait_us > 0) {             // the caller asked for a collection window
      const auto deadline = cur.t0 + std::chrono::microseconds(max_wait_us);
      cv_work.wait_until(lk, deadline, [&] { return closing || paused || cur.count >= max_batch; });
    }
    // small batches only while the ring is nearly empty (see kEagerDepth); completions and arrivals wake us
    cv_work.wait(lk, [&] {
      return closing || paused || (int)inflight.size() < kEagerDepth || cur.count >= std::min(kEagerBatch, max_batch);
    });
    const int me = fill;
    cur.state = State::Running;
    cur.ticket++;
    advance_fill();                     // requests now collect in another buffer
    cv_join.notify_all();
    const size_t count = cur.count, k = cur.k, ef = pending_ef;
    pending_ef = 0;
    int rc = HS_OK;
    std::string err;
    if (ix) {
      lk.unlock();
      if (ef) rc = hs_set_ef(ix, ef);
      if (rc == HS_OK) rc = hs_search_batch_submit(ix, cur.q, count, k, cur.lab, cur.dist);     // asynchronous
      if (rc != HS_OK) err = hs_last_error();
      lk.lock();
    }
    if (rc != HS_OK) {
      finish(me, rc, err, 0.0);         // nothing was enqueued: the callers get the error
    } else {
      inflight.push_back(me);           // only now: the completer must not wait for a batch that is not enqueued yet
      cv_flight.notify_one();
    }
  }
}

void hs_service::complete_loop() {
  std::unique_lock<std::mutex> lk(mu);
  for (;;) {
    cv_flight.wait(lk, [&] { return closing_completer || !inflight.empty(); });
    if (inflight.empty()) {
      if (closing_completer) return;
      continue;
    }
    const int me = inflight.front();
    Batch &cur = b[me];
    const size_t count = cur.count, k = cur.k;
    lk.unlock();
    int rc;
    std::string err;
    const auto t0 = Clock::now();
    if (ix) {
      rc = hs_search_batch_wait_oldest(ix);        // batches complete in submission order
      if (rc != HS_OK) err = hs_last_error();
    } else {
      rc = fn(fn_ctx, cur.q, count, k, cur.lab, cur.dist);
      if (rc != HS_OK) err = "backend callback failed";
    }
    const double seconds = std::chrono::duration<double>(Clock::now() - t0).count();
    lk.lock();
    inflight.pop_front();
    finish(me, rc, err, seconds);
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
    // page-locked + mapped: hs_search_batch_submit uses the buffers in place
    ok = cudaSetDevice(ix->device) == cudaSuccess;
    s->pinned = true;
    for (int i = 0; i < kDepth && ok; ++i) {
      ok = cudaHostAlloc(reinterpret_cast<void **>(&s->b[i].q), qb, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
           cudaHostAlloc(reinterpret_cast<void **>(&s->b[i].lab), rb, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
           cudaHostAlloc(reinterpret_cast<void **>(&s->b[i].dist), rb, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
    }
    if (!ok) set_error(std::string("hs_service_create: cudaHostAlloc: ") + cudaGetErrorString(cudaGetLastError()));
    if (ok) hs_set_overlap(ix, 1);        // consecutive small launches run side by side (see the header comment)
  } else {
    for (int i = 0; i < kDepth && ok; ++i) {
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
    s->dispatcher = std::thread([s] { s->dispatch_loop(); });
    s->completer = std::thread([s] { s->complete_loop(); });
  } catch (const std::exception &e) {
    set_error(std::string("hs_service_create: cannot start the service threads: ") + e.what());
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
  if (slot == 0 || bt->count == s->max_batch || bt->count == kEagerBatch) s->cv_work.notify_one();
  bt->done.wait(lk, [&] { return bt->ticket >= want && bt->state == State::Draining; });
  const int rc = bt->rc;
  if (rc == HS_OK) {
    std::memcpy(labels_out, bt->lab + slot * k, k * 4);
    if (dists_out) std::memcpy(dists_out, bt->dist + slot * k, k * 4);
  } else {
    set_error(bt->err);
  }
  if (--bt->readers == 0) {             // last reader: the buffer may fill again
    bt->count = 0;
    bt->state = State::Filling;
    if (s->b[s->fill].state != State::Filling) s->fill = (int)(bt - s->b);     // arrivals were waiting for a buffer
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
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->closing_completer = true;        // after the dispatcher: what it submitted is still waited for
    s->cv_flight.notify_all();
  }
  if (s->completer.joinable()) s->completer.join();
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
