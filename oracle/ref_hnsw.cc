// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around two more UNMODIFIED reference index classes under /root/reference:
//   * hnswlib::HierarchicalNSW<float>          (third_party/hnswlib/hnswalg.h) — the `hnsw`
//     strategy, include/strategy/hnsw_strategy.h:15-61: loadIndex + searchKnn(query, k)
//   * hnswlib::HierarchicalNSWSlimZero<float>  (third_party/hnswlib/hnswalg_slimzero.h) — the
//     `hnsw_slimzero` strategy, include/strategy/hnsw_slimzero_strategy.h:38-140:
//     convertFromHNSW + saveIndex + loadIndex + searchKnn(query, k, tableint*)
// Nothing is copied: this TU only #includes the reference where it lies and forwards calls.
// Built by oracle/Makefile into oracle/_ref/libhsref_hnsw_{v3,v4}.so; loaded only by tests/.
#include "core.h"
#include "hnswlib/hnswlib.h"
#include "hnswlib/hnswalg.h"
#include "hnswlib/hnswalg_slimzero.h"

#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {
std::string g_err;

std::unique_ptr<hnswlib::SpaceInterface<float>> make_space(size_t dim, int metric) {
  if (metric == 1) return std::unique_ptr<hnswlib::SpaceInterface<float>>(new hnswlib::InnerProductSpace(dim));
  return std::unique_ptr<hnswlib::SpaceInterface<float>>(new hnswlib::L2Space(dim));
}

struct HnswHandle {
  std::unique_ptr<hnswlib::SpaceInterface<float>> space;
  std::unique_ptr<hnswlib::HierarchicalNSW<float>> index;
  size_t dim;
};
struct ZeroHandle {
  std::unique_ptr<hnswlib::SpaceInterface<float>> space;
  std::unique_ptr<hnswlib::HierarchicalNSWSlimZero<float>> index;
  size_t dim;
};
}  // namespace

extern "C" {

const char *refh_last_error() { return g_err.c_str(); }

// hnsw_strategy.h:24-45: omp addPoint loop + saveIndex.  labels may be NULL (label = row).
int refh_hnsw_build(const float *base, size_t n, size_t dim, int metric, size_t M, size_t ef_construction,
                    const char *branching, int threads, const uint64_t *labels, const char *out_graph) {
  try {
    auto space = make_space(dim, metric);
    hnswlib::HierarchicalNSW<float> hnsw(space.get(), n, M, ef_construction, std::string(branching));
    if (threads <= 0) threads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t i = 0; i < n; ++i) hnsw.addPoint(base + i * dim, labels ? labels[i] : i);
    hnsw.saveIndex(out_graph);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

void *refh_hnsw_open(const char *graph, size_t dim, int metric, size_t max_elements) {
  try {
    auto *h = new HnswHandle;
    h->dim = dim;
    h->space = make_space(dim, metric);
    h->index.reset(new hnswlib::HierarchicalNSW<float>(h->space.get(), graph, false, max_elements));
    return h;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void refh_hnsw_close(void *hv) { delete static_cast<HnswHandle *>(hv); }

// hnsw_strategy.h:49-58: per query searchKnn(q, K) -> max-heap of (dist, label); written here
// NEAREST first.  threads == 1 serial, else omp dynamic over queries.
int refh_hnsw_search(void *hv, const float *q, size_t nq, size_t k, size_t ef, int threads, uint32_t *out_labels,
                     float *out_dists, double *seconds) {
  auto *h = static_cast<HnswHandle *>(hv);
  h->index->setEf(ef);
  const size_t dim = h->dim;
  if (threads <= 0) threads = omp_get_num_procs();
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
  for (size_t i = 0; i < nq; ++i) {
    auto res = h->index->searchKnn(q + i * dim, k);
    size_t m = res.size();
    for (size_t j = 0; j < k; ++j) {
      out_labels[i * k + j] = 0xFFFFFFFFu;
      if (out_dists) out_dists[i * k + j] = __builtin_inff();
    }
    while (!res.empty()) {
      --m;
      out_labels[i * k + m] = (uint32_t)res.top().second;
      if (out_dists) out_dists[i * k + m] = res.top().first;
      res.pop();
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

// hnsw_slimzero_strategy.h:38-103: HNSW build, convertFromHNSW, saveIndex.
int refh_slimzero_build(const float *base, size_t n, size_t dim, int metric, size_t M, size_t ef_construction,
                        const char *branching, int threshold_level, float top_degree_percent0,
                        float top_degree_percent, size_t top_M0, size_t low_m0, size_t top_M, size_t low_m,
                        size_t min_indegree0, size_t min_indegree, int threads, const uint64_t *labels,
                        const char *out_graph) {
  try {
    auto space = make_space(dim, metric);
    hnswlib::HierarchicalNSW<float> hnsw(space.get(), n, M, ef_construction, std::string(branching));
    if (threads <= 0) threads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t i = 0; i < n; ++i) hnsw.addPoint(base + i * dim, labels ? labels[i] : i);
    hnswlib::HierarchicalNSWSlimZero<float> zero(space.get(), n, M, ef_construction, threshold_level,
                                                 top_degree_percent0, top_degree_percent, top_M0, low_m0, top_M,
                                                 low_m, min_indegree0, min_indegree);
    zero.convertFromHNSW(&hnsw);
    zero.saveIndex(out_graph);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

void *refh_slimzero_open(const char *graph, size_t dim, int metric, size_t max_elements) {
  try {
    auto *h = new ZeroHandle;
    h->dim = dim;
    h->space = make_space(dim, metric);
    h->index.reset(new hnswlib::HierarchicalNSWSlimZero<float>(h->space.get()));
    h->index->loadIndex(graph, h->space.get(), max_elements);
    return h;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void refh_slimzero_close(void *hv) { delete static_cast<ZeroHandle *>(hv); }

// hnsw_slimzero_strategy.h:131-133: the serial loop over searchKnn(q, K, tableint*); k labels per
// query, unordered within a row (hnswalg_slimzero.h:1766-1770).
int refh_slimzero_search(void *hv, const float *q, size_t nq, size_t k, size_t ef, uint32_t *out_labels,
                         double *seconds) {
  auto *h = static_cast<ZeroHandle *>(hv);
  K = k;   // the reference's global (include/core.h:30)
  h->index->setEf(ef);
  auto t0 = std::chrono::steady_clock::now();
  for (size_t i = 0; i < nq; ++i) h->index->searchKnn(q + i * h->dim, k, out_labels + i * k);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

}  // extern "C"
