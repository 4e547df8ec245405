// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference headers under /root/reference
// (hnswlib fork: hnswalg.h, hnswalg_slim.h, bruteforce.h, space_l2.h, space_ip.h).
// Nothing is copied: this TU only #includes the reference where it lies and
// forwards calls.  Built by oracle/Makefile into oracle/_ref/libhsref_slim_*.so.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it.  It is (1) the graph builder of the parity tests
// (HierarchicalNSW::addPoint -> HierarchicalNSWSlim::convertFromHNSW -> saveIndex),
// (2) the reference search the CUDA path and oracle/hs_oracle.c are pinned to,
// (3) the CPU baseline timed beside the GPU numbers.
//
// Reference call sites mirrored here:
//   build        include/strategy/hnsw_slim_strategy.h:60-95
//   search loop  include/strategy/hnsw_slim_strategy.h:112-114 (serial) and
//                hnsw_slim_client_update_patch.cc:223-226 (omp dynamic)
//   brute force  include/strategy/brute_force_strategy.h:15-45
#include "core.h"
#include "hnswlib/hnswlib.h"
#include "hnswlib/hnswalg.h"
#include "hnswlib/hnswalg_slim.h"
#include "hnswlib/bruteforce.h"
#include "strategy/solve_strategy.h"   // SolveStrategy::recall, executed as is (solve_strategy.h:67-103)

#include <omp.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <iostream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local uint64_t g_dist_calls = 0;
hnswlib::DISTFUNC<float> g_inner = nullptr;

float counting_dist(const void *a, const void *b, const void *p) {
  ++g_dist_calls;
  return g_inner(a, b, p);
}

// SpaceInterface whose distance function is the reference's own, optionally
// wrapped by a call counter.
struct Space : hnswlib::SpaceInterface<float> {
  std::unique_ptr<hnswlib::SpaceInterface<float>> inner;
  bool counting;
  Space(size_t dim, int metric, bool counting_) : counting(counting_) {
    if (metric == 1)
      inner.reset(new hnswlib::InnerProductSpace(dim));
    else
      inner.reset(new hnswlib::L2Space(dim));
    if (counting) g_inner = inner->get_dist_func();
  }
  size_t get_data_size() override { return inner->get_data_size(); }
  hnswlib::DISTFUNC<float> get_dist_func() override {
    return counting ? counting_dist : inner->get_dist_func();
  }
  void *get_dist_func_param() override { return inner->get_dist_func_param(); }
};

struct SlimHandle {
  std::unique_ptr<Space> space;
  std::unique_ptr<hnswlib::HierarchicalNSWSlim<float>> index;
  size_t dim;
};

thread_local std::string g_err;

}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

int ref_num_procs() { return omp_get_num_procs(); }

// Build the full HNSW with OpenMP addPoint, prune it into the CHAL layout and
// save it exactly as HnswSlimStrategy::solve does.  labels[i] (or i when
// labels == NULL) is the external label of row i.  Returns 0 on success.
int ref_slim_build(const float *base, size_t n, size_t dim, int metric,
                   size_t M, size_t ef_construction, const char *branching,
                   int threshold_level, float top_degree_percent0,
                   float top_degree_percent, size_t top_M0, size_t low_m0,
                   size_t top_M, size_t low_m, int threads,
                   const uint64_t *labels, const char *out_slim_graph,
                   const char *out_hnsw_graph, double *build_s,
                   double *convert_s) {
  try {
    Space space(dim, metric, false);
    hnswlib::HierarchicalNSW<float> hnsw(&space, n, M, ef_construction,
                                         std::string(branching));
    auto t0 = std::chrono::steady_clock::now();
    if (threads <= 0) threads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t i = 0; i < n; ++i) {
      hnsw.addPoint(base + i * dim, labels ? labels[i] : i);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (out_hnsw_graph && out_hnsw_graph[0]) hnsw.saveIndex(out_hnsw_graph);

    hnswlib::HierarchicalNSWSlim<float> slim(
        &space, n, M, ef_construction, threshold_level, top_degree_percent0,
        top_degree_percent, top_M0, low_m0, top_M, low_m);
    auto t2 = std::chrono::steady_clock::now();
    slim.convertFromHNSW(&hnsw);
    auto t3 = std::chrono::steady_clock::now();
    slim.saveIndex(out_slim_graph);
    if (build_s) *build_s = std::chrono::duration<double>(t1 - t0).count();
    if (convert_s) *convert_s = std::chrono::duration<double>(t3 - t2).count();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

void *ref_slim_open(const char *graph, size_t dim, int metric,
                    size_t max_elements, int counting) {
  try {
    auto *h = new SlimHandle;
    h->dim = dim;
    h->space.reset(new Space(dim, metric, counting != 0));
    h->index.reset(new hnswlib::HierarchicalNSWSlim<float>(h->space.get()));
    h->index->loadIndex(graph, h->space.get(), max_elements);
    // patchFromStream's vector<vector> and to_add overloads free() the neighbour pointer of NEW nodes too
    // (slim.h:2221-2236, :2303-2319), i.e. they rely on the unused tail of elements_ (a plain malloc,
    // slim.h:785) reading as zero — true for the client's fresh multi-hundred-MB mapping, not for a
    // small block recycled inside a test process.  Give the tail the state it has there.
    auto &ix = *h->index;
    if (ix.max_elements_ > ix.cur_element_count_)
      std::memset(ix.elements_ + ix.cur_element_count_ * ix.size_data_per_element_, 0,
                  (ix.max_elements_ - ix.cur_element_count_) * ix.size_data_per_element_);
    return h;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void ref_slim_close(void *hv) { delete static_cast<SlimHandle *>(hv); }

// header fields as loaded by the reference (for loader parity tests)
void ref_slim_info(void *hv, uint64_t *out /*[10]*/) {
  auto *h = static_cast<SlimHandle *>(hv);
  auto &ix = *h->index;
  out[0] = ix.cur_element_count_;
  out[1] = ix.size_data_per_element_;
  out[2] = (uint64_t)(int64_t)ix.maxlevel_;
  out[3] = (uint64_t)(int64_t)ix.threshold_level_;
  out[4] = ix.enterpoint_node_;
  out[5] = ix.maxM_;
  out[6] = ix.maxM0_;
  out[7] = ix.M_;
  out[8] = ix.ef_construction_;
  out[9] = ix.has_deleted_elements_ ? 1 : 0;
}

// level of a node, its external label, and its level-`level` neighbour slice
// read through the reference's own accessors (slim.h:620-661, :245-260).
int ref_slim_node(void *hv, uint32_t node, int level, int *node_level,
                  uint64_t *label, uint32_t *out, int cap) {
  auto *h = static_cast<SlimHandle *>(hv);
  auto &ix = *h->index;
  char *element = ix.elements_ + (size_t)node * ix.size_data_per_element_;
  int el = ix.get_element_level(element);
  if (node_level) *node_level = el;
  if (label) *label = ix.getExternalLabel(node);
  char *nb = ix.get_neighbors(element);
  if (nb == nullptr || level > el) return 0;
  size_t off = ix.get_neighbor_offset_at_level(nb, level);
  size_t end = (level == el) ? (size_t)ix.get_total_neighbor(element)
                             : (size_t)((hnswlib::offsetint *)nb)[level];
  const hnswlib::tableint *ids =
      (const hnswlib::tableint *)(nb + sizeof(hnswlib::offsetint) * el) + off;
  int cnt = (int)(end - off);
  for (int i = 0; i < cnt && i < cap; ++i) out[i] = ids[i];
  return cnt;
}

// The reference's query loop.  threads == 1: the serial loop of
// hnsw_slim_strategy.h:112-114.  threads > 1 (or <= 0 for all cores): the
// OpenMP dynamic loop of hnsw_slim_client_update_patch.cc:223-226.
// out: nq*k labels, unordered within a row (slim.h:2126-2130).
// Returns elapsed seconds of the loop in *seconds.
int ref_slim_search(void *hv, const float *q, size_t nq, size_t k, size_t ef,
                    int threads, uint32_t *out, double *seconds,
                    uint64_t *dist_calls) {
  auto *h = static_cast<SlimHandle *>(hv);
  K = k;  // the reference's global (include/core.h:30)
  h->index->setEf(ef);
  size_t dim = h->dim;
  uint64_t total_calls = 0;
  auto t0 = std::chrono::steady_clock::now();
  if (threads == 1) {
    g_dist_calls = 0;
    for (size_t i = 0; i < nq; ++i)
      h->index->searchKnn(q + i * dim, k, out + i * k);
    total_calls = g_dist_calls;
  } else {
    if (threads <= 0) threads = omp_get_num_procs();
#pragma omp parallel num_threads(threads) reduction(+ : total_calls)
    {
      g_dist_calls = 0;
#pragma omp for schedule(dynamic)
      for (size_t i = 0; i < nq; ++i)
        h->index->searchKnn(q + i * dim, k, out + i * k);
      total_calls += g_dist_calls;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (dist_calls) *dist_calls = total_calls;
  return 0;
}

// per-query distance-evaluation counts (index must be opened with counting=1)
int ref_slim_counts(void *hv, const float *q, size_t nq, size_t k, size_t ef,
                    uint32_t *out, uint64_t *per_query) {
  auto *h = static_cast<SlimHandle *>(hv);
  K = k;
  h->index->setEf(ef);
  for (size_t i = 0; i < nq; ++i) {
    g_dist_calls = 0;
    h->index->searchKnn(q + i * h->dim, k, out + i * k);
    per_query[i] = g_dist_calls;
  }
  return 0;
}

// the DISTFUNC the reference picks for (dim, metric) — space_l2.h:214-238,
// space_ip.h:342-398
float ref_dist(const float *a, const float *b, size_t dim, int metric) {
  Space s(dim, metric, false);
  return s.get_dist_func()(a, b, s.get_dist_func_param());
}

// Ground truth exactly as BruteForce::solve produces it: BruteforceSearch::
// searchKnn (bruteforce.h:106-135) per query under omp, heap drained
// farthest-first (brute_force_strategy.h:24-36).  out_dists may be NULL.
int ref_bruteforce(const float *base, size_t n, size_t dim, int metric,
                   const float *q, size_t nq, size_t k, int threads,
                   uint32_t *out_labels, float *out_dists, double *seconds) {
  try {
    Space space(dim, metric, false);
    hnswlib::BruteforceSearch<float> bf(&space, n);
    for (size_t i = 0; i < n; ++i) bf.addPoint(base + i * dim, i);
    if (threads <= 0) threads = omp_get_num_procs();
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t i = 0; i < nq; ++i) {
      auto res = bf.searchKnn(q + i * dim, k);
      size_t j = 0;
      while (!res.empty() && j < k) {
        out_labels[i * k + j] = (uint32_t)res.top().second;
        if (out_dists) out_dists[i * k + j] = res.top().first;
        res.pop();
        ++j;
      }
      for (; j < k; ++j) {
        out_labels[i * k + j] = (uint32_t)-1;
        if (out_dists) out_dists[i * k + j] = 0.f;
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// ---- delta patches ----
// SERVER side of the protocol, hnsw_slim_server_patch.cc:186-228 (/updateIndex) and :232-279
// (/getLastBatch): the full HNSW receives new rows with addPoint, convertFromHNSWWithDiff re-prunes it
// and writes what changed (slim.h:1110-1424) — the response body a client feeds to patchFromStream.
// Here: rows [0, n0) make the partial index (saved to out_partial_graph); rows [n0, n) arrive in
// `rounds` equal updates, round r's stream goes to "<patch_prefix><r>.bin".  inline_last != 0: the last
// round is produced the /getLastBatch way instead — convertFromHNSWWithDiff(hnsw, old_cnt, new_cnt) +
// genPatch(..., to_add = true) (slim.h:1427-1476, :1478-1751), new rows INLINE in the stream; the file
// holds what the client hands to patchFromStream(in, true) (i.e. without the leading `finished` word).
// The server's own index after the last round is saved to out_final_graph (may be empty).
int ref_slim_make_patches(const float *base, size_t n, size_t dim, int metric, size_t M, size_t ef_construction,
                          const char *branching, int threshold_level, float top_degree_percent0,
                          float top_degree_percent, size_t top_M0, size_t low_m0, size_t top_M, size_t low_m,
                          int threads, size_t n0, size_t rounds, int inline_last, const char *out_partial_graph,
                          const char *patch_prefix, const char *out_final_graph) {
  try {
    if (n0 == 0 || n0 >= n || rounds == 0) {
      g_err = "ref_slim_make_patches: need 0 < n0 < n and rounds > 0";
      return -1;
    }
    Space space(dim, metric, false);
    hnswlib::HierarchicalNSW<float> hnsw(&space, n, M, ef_construction, std::string(branching));
    if (threads <= 0) threads = omp_get_num_procs();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t i = 0; i < n0; ++i) hnsw.addPoint(base + i * dim, i);
    hnswlib::HierarchicalNSWSlim<float> slim(&space, n, M, ef_construction, threshold_level, top_degree_percent0,
                                             top_degree_percent, top_M0, low_m0, top_M, low_m);
    std::ostringstream quiet;
    std::streambuf *old = std::cout.rdbuf(quiet.rdbuf());       // "lookup", "changed nodes count: ..." chatter
    slim.convertFromHNSW(&hnsw);
    slim.saveIndex(out_partial_graph);
    // convertFromHNSWWithDiff reads the previous neighbour pointer of EVERY node below the new count, new
    // ones included (slim.h:1251-1254, :1360-1375), i.e. it relies on the unused tail of elements_ (a plain
    // malloc of max_elements records, slim.h:827-828) reading as zero — true for the server's fresh
    // multi-hundred-MB mapping, not for a small recycled block.  Give the tail the state it has there.
    std::memset(slim.elements_ + slim.cur_element_count_ * slim.size_data_per_element_, 0,
                (slim.max_elements_ - slim.cur_element_count_) * slim.size_data_per_element_);
    for (size_t r = 0; r < rounds; ++r) {
      const size_t lo = n0 + (n - n0) * r / rounds, hi = n0 + (n - n0) * (r + 1) / rounds;
#pragma omp parallel for schedule(dynamic) num_threads(threads)
      for (size_t i = lo; i < hi; ++i) hnsw.addPoint(base + i * dim, i, false);
      std::ostringstream oss(std::ios::binary);
      if (inline_last && r + 1 == rounds) {
        size_t changed_old = 0, changed_new = 0;
        slim.convertFromHNSWWithDiff(&hnsw, changed_old, changed_new);
        std::ostringstream body(std::ios::binary);
        size_t old_written = 0, new_written = 0;
        const uint32_t finished = slim.genPatch(body, old_written, new_written, (size_t)1 << 62, true);
        if (finished != 1) {
          std::cout.rdbuf(old);
          g_err = "genPatch did not finish in one batch";
          return -1;
        }
        const size_t cur = slim.cur_element_count_;
        hnswlib::writeBinaryPOD(oss, cur);
        hnswlib::writeBinaryPOD(oss, old_written);
        hnswlib::writeBinaryPOD(oss, new_written);
        oss << body.str();
      } else {
        slim.convertFromHNSWWithDiff(&hnsw, oss);
      }
      const std::string bytes = oss.str();
      const std::string path = std::string(patch_prefix) + std::to_string(r) + ".bin";
      std::ofstream f(path, std::ios::binary);
      f.write(bytes.data(), (std::streamsize)bytes.size());
      if (!f) {
        std::cout.rdbuf(old);
        g_err = "cannot write " + path;
        return -1;
      }
    }
    if (out_final_graph && out_final_graph[0]) slim.saveIndex(out_final_graph);
    std::cout.rdbuf(old);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// CLIENT side: patchFromStream on an index opened with max_elements >= the patched count
// (hnsw_slim_client_update_patch.cc:113, :56-81).  mode 0: the unordered_map overload the client uses
// (slim.h:2343-2388; rows = the new vectors by LABEL: rows[label * dim] for label < n_rows);
// mode 1: the vector<vector<float>> overload (slim.h:2206-2253); mode 2: rows inline, patchFromStream(in, true)
// (slim.h:2292-2340, the /getLastBatch path).
int ref_slim_patch(void *hv, const void *bytes, size_t len, int mode, const float *rows, size_t n_rows) {
  auto *h = static_cast<SlimHandle *>(hv);
  try {
    std::istringstream in(std::string(static_cast<const char *>(bytes), len), std::ios::binary);
    if (mode == 2) {
      h->index->patchFromStream(in, true);
    } else if (mode == 1) {
      std::vector<std::vector<float>> data(n_rows);
      for (size_t i = 0; i < n_rows; ++i) data[i].assign(rows + i * h->dim, rows + (i + 1) * h->dim);
      h->index->patchFromStream(in, data);
    } else {
      std::unordered_map<uint32_t, std::vector<float>> data;
      data.reserve(n_rows);
      for (size_t i = 0; i < n_rows; ++i) data[(uint32_t)i].assign(rows + i * h->dim, rows + (i + 1) * h->dim);
      h->index->patchFromStream(in, data);
    }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

int ref_slim_save(void *hv, const char *path) {
  auto *h = static_cast<SlimHandle *>(hv);
  try {
    h->index->saveIndex(path);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// SolveStrategy::recall (include/strategy/solve_strategy.h:67-103) EXECUTED, not restated: the strategy
// object reads base / queries (.fvecs), the answers (.ivecs, read_knn :58-61) and the ground truth (.ivecs)
// exactly as `main` would, and prints "Recall: x" — which is captured and returned.  k = the global K.
double ref_strategy_recall(const char *source_fvecs, const char *query_fvecs, const char *knn_ivecs,
                           const char *gt_ivecs, size_t k) {
  struct RecallOnly : SolveStrategy {
    using SolveStrategy::SolveStrategy;
    void solve() override {}
  };
  try {
    K = k;
    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    double value = -1.0;
    {
      RecallOnly s(source_fvecs, query_fvecs, "");
      s.read_knn(knn_ivecs);
      s.recall(gt_ivecs);
    }
    std::cout.rdbuf(old);
    const std::string out = captured.str();
    const size_t at = out.rfind("Recall: ");
    if (at == std::string::npos) {
      g_err = "SolveStrategy::recall printed no result";
      return -1.0;
    }
    value = std::stod(out.substr(at + 8));
    return value;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1.0;
  }
}

}  // extern "C"
