// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the UNMODIFIED reference headers for the hnsw_slimq path
// (third_party/hnswlib/hnswalg_slimq.h + the vendored rabitqlib).  Nothing is copied:
// this TU #includes the reference where it lies under /root/reference and forwards
// calls.  Built by oracle/Makefile into oracle/_ref/libhsref_slimq_v4.so (rabitqlib
// hard-requires AVX-512: rq/quantization/pack_excode.hpp:56-102).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.
//
// Reference call sites mirrored here:
//   build   include/strategy/hnsw_slimq_strategy.h:100-142  (rabitqlib HNSW construct ->
//           HierarchicalNSWSlimQ::convertFromHNSW -> saveIndex)
//   search  include/strategy/hnsw_slimq_strategy.h:145-159  (setDataset, setEf, serial loop)
//   prep    third_party/hnswlib/hnswalg_slimq.h:1814-1856   (rotate, SplitSingleQuery,
//           centroid distances, get_bin_est)
//
// rabitqlib (fht_avx.hpp) and core.h define non-inline symbols, so this is its own TU /
// shared object, separate from ref_slim.cc.
#include "core.h"
#include "hnswlib/hnswlib.h"
#include "hnswlib/hnswalg.h"
#include "hnswlib/hnswalg_slimq.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct SlimQHandle {
  std::unique_ptr<hnswlib::L2Space> space;
  std::unique_ptr<hnswlib::HierarchicalNSWSlimQ<float>> index;
  std::vector<std::vector<float>> dataset;   // what SolveStrategy keeps in data_set_
  size_t dim = 0;
};

}  // namespace

extern "C" {

const char *refq_last_error() { return g_err.c_str(); }

// hnsw_slimq_strategy.h:100-142.  The strategy hard-codes total_bits=4, M=32,
// ef_construction=128, seed=100 for the rabitqlib HNSW (:106-108); M / efc are
// parameters here so small test graphs can use other values.  `threads` = 1 keeps
// internal id == label.  Returns 0 on success.
int refq_build(const float *base, size_t n, size_t dim, const float *centroids,
               size_t num_cluster, const uint32_t *cluster_ids, size_t M,
               size_t ef_construction, int threshold_level,
               float top_degree_percent0, float top_degree_percent,
               size_t top_M0, size_t low_m0, size_t top_M, size_t low_m,
               int threads, const char *out_slimq_graph, double *build_s,
               double *convert_s) {
  try {
    hnswlib::L2Space l2space(dim);
    hnswlib::HierarchicalNSWSlimQ<float> slimq(
        &l2space, n, M, ef_construction, threshold_level, top_degree_percent0,
        top_degree_percent, top_M0, low_m0, top_M, low_m);
    auto *hnsw = new rabitqlib::hnsw::HierarchicalNSW(
        n, dim, 4, M, ef_construction, 100, rabitqlib::METRIC_L2);
    std::vector<uint32_t> cids(cluster_ids, cluster_ids + n);
    auto t0 = std::chrono::steady_clock::now();
    hnsw->construct(num_cluster, centroids, n, base, cids.data(),
                    (size_t)(threads <= 0 ? 0 : threads), true);
    auto t1 = std::chrono::steady_clock::now();
    slimq.convertFromHNSW(hnsw);
    auto t2 = std::chrono::steady_clock::now();
    slimq.saveIndex(out_slimq_graph);
    if (build_s) *build_s = std::chrono::duration<double>(t1 - t0).count();
    if (convert_s) *convert_s = std::chrono::duration<double>(t2 - t1).count();
    delete hnsw;
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// loadIndex + setDataset (hnsw_slimq_strategy.h:72-73,145).  `base` is copied into the
// vector<vector<float>> form the strategy owns.
void *refq_open(const char *graph, size_t dim, size_t n, const float *base) {
  try {
    auto h = std::make_unique<SlimQHandle>();
    h->dim = dim;
    h->space.reset(new hnswlib::L2Space(dim));
    h->index.reset(new hnswlib::HierarchicalNSWSlimQ<float>(h->space.get(), n));
    h->index->loadIndex(graph, h->space.get(), n);
    h->dataset.resize(n);
    for (size_t i = 0; i < n; ++i)
      h->dataset[i].assign(base + i * dim, base + (i + 1) * dim);
    h->index->setDataset(&h->dataset);
    return h.release();
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void refq_close(void *hp) { delete static_cast<SlimQHandle *>(hp); }

// info[0..]: n, size_data_per_element, maxM, maxM0, M, ef_construction, maxlevel,
// threshold_level, enterpoint, num_cluster, dim, padded_dim, ex_bits, metric_type
void refq_info(void *hp, uint64_t *info) {
  auto &ix = *static_cast<SlimQHandle *>(hp)->index;
  info[0] = ix.cur_element_count_;
  info[1] = ix.size_data_per_element_;
  info[2] = ix.maxM_;
  info[3] = ix.maxM0_;
  info[4] = ix.M_;
  info[5] = ix.ef_construction_;
  info[6] = (uint64_t)(int64_t)ix.maxlevel_;
  info[7] = (uint64_t)(int64_t)ix.threshold_level_;
  info[8] = ix.enterpoint_node_;
  info[9] = ix.num_cluster_;
  info[10] = ix.dim_;
  info[11] = ix.padded_dim_;
  info[12] = ix.ex_bits_;
  info[13] = (uint64_t)ix.metric_type_;
}

// The query-quantiser constant drawn at load time (slimq.h:1274-1276 ->
// rabitq.hpp:27-34 -> rabitq_impl.hpp:363-377: 100 Gaussian vectors from
// std::random_device, hence different on every load).
double refq_get_tconst(void *hp) {
  return static_cast<SlimQHandle *>(hp)->index->query_config_.t_const;
}
void refq_set_tconst(void *hp, double t) {
  static_cast<SlimQHandle *>(hp)->index->query_config_.t_const = t;
}

// Serial query loop of hnsw_slimq_strategy.h:157-159 (the slimq search is not
// re-entrant: member search_pool_, slimq.h:220,1814).  Sets the global K = k first
// (the rerank heap is bounded by K, slimq.h:1917).  out: nq*k external labels in the
// reference's heap order.
int refq_search(void *hp, const float *queries, size_t nq, size_t k, size_t ef,
                uint32_t *out_labels, double *seconds) {
  try {
    auto *h = static_cast<SlimQHandle *>(hp);
    K = k;
    h->index->setEf(ef);
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < nq; ++i)
      h->index->searchKnn(queries + i * h->dim, k, out_labels + i * k);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// The per-query preparation of searchKnn (slimq.h:1816-1847) through the same library
// calls: rotated query [padded_dim], query bit planes [padded_dim/64*4], delta, vl,
// k1xsumq -> scal[0..2], q_to_centroids [num_cluster] (L2 metric).
int refq_prep(void *hp, const float *query, float *rotated, uint64_t *planes,
              float *scal, float *q_to_centroids) {
  try {
    auto &ix = *static_cast<SlimQHandle *>(hp)->index;
    std::vector<float> rq(ix.padded_dim_);
    ix.rotator_->rotate(query, rq.data());
    rabitqlib::SplitSingleQuery<float> qw(rq.data(), ix.padded_dim_, ix.ex_bits_,
                                          ix.query_config_, ix.metric_type_);
    std::memcpy(rotated, rq.data(), sizeof(float) * ix.padded_dim_);
    std::memcpy(planes, qw.query_bin(), sizeof(uint64_t) * ix.padded_dim_ / 64 * 4);
    scal[0] = qw.delta();
    scal[1] = qw.vl();
    scal[2] = qw.k1xsumq();
    for (size_t c = 0; c < ix.num_cluster_; ++c)
      q_to_centroids[c] = std::sqrt(ix.raw_dist_func_(
          rq.data(),
          reinterpret_cast<float *>(ix.centroids_memory_) + c * ix.padded_dim_,
          ix.padded_dim_));
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// get_bin_est (slimq.h:408-440) of `query` against the listed internal ids.
int refq_est(void *hp, const float *query, const uint32_t *ids, size_t n_ids,
             float *est_out) {
  try {
    auto &ix = *static_cast<SlimQHandle *>(hp)->index;
    std::vector<float> rq(ix.padded_dim_);
    ix.rotator_->rotate(query, rq.data());
    rabitqlib::SplitSingleQuery<float> qw(rq.data(), ix.padded_dim_, ix.ex_bits_,
                                          ix.query_config_, ix.metric_type_);
    std::vector<float> q2c(ix.num_cluster_);
    for (size_t c = 0; c < ix.num_cluster_; ++c)
      q2c[c] = std::sqrt(ix.raw_dist_func_(
          rq.data(),
          reinterpret_cast<float *>(ix.centroids_memory_) + c * ix.padded_dim_,
          ix.padded_dim_));
    for (size_t i = 0; i < n_ids; ++i) {
      hnswlib::HierarchicalNSWSlimQ<float>::EstimateRecord rec;
      ix.get_bin_est(q2c, qw, ids[i], rec);
      est_out[i] = rec.est_dist;
    }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// FhtKacRotator::rotate alone (rotator.hpp:370-423) with the index's flips.
int refq_rotate(void *hp, const float *vecs, size_t nv, float *out) {
  auto *h = static_cast<SlimQHandle *>(hp);
  auto &ix = *h->index;
  for (size_t i = 0; i < nv; ++i)
    ix.rotator_->rotate(vecs + i * h->dim, out + i * ix.padded_dim_);
  return 0;
}

// accessors over the reference record (slimq.h:388-405): cluster id, code words,
// f_add, f_rescale of one node; level-0 neighbour list.
int refq_node(void *hp, uint32_t node, uint32_t *cluster, uint64_t *code,
              float *factors, uint32_t *nbr_out, int cap) {
  auto &ix = *static_cast<SlimQHandle *>(hp)->index;
  *cluster = ix.get_clusterid_by_internalid(node);
  const char *bin = ix.get_bindata_by_internalid(node);
  std::memcpy(code, bin, ix.padded_dim_ / 8);
  std::memcpy(factors, bin + ix.padded_dim_ / 8, 12);
  char *element = ix.elements_ + (size_t)node * ix.size_data_per_element_;
  char *neighbors = ix.get_neighbors(element);
  if (!neighbors) return 0;
  int lvl = ix.get_element_level(element);
  size_t size = lvl == 0 ? ix.get_total_neighbor(element)
                         : ((hnswlib::offsetint *)(neighbors))[0];
  hnswlib::tableint *data =
      (hnswlib::tableint *)(neighbors + sizeof(hnswlib::offsetint) * lvl);
  int m = (int)size < cap ? (int)size : cap;
  for (int j = 0; j < m; ++j) nbr_out[j] = data[j];
  return (int)size;
}

// All-core baseline for hnsw_slimq ("one index copy per thread — state which").  The
// reference's slimq searchKnn is not re-entrant (member search_pool_, slimq.h:220,1814), so `copies` =
// `threads` independent HierarchicalNSWSlimQ objects are loaded from the same file (loadIndex, slimq.h:1218)
// over ONE shared raw dataset (setDataset only borrows it, slimq.h:303-305), and every OpenMP thread runs
// the serial loop of hnsw_slimq_strategy.h:157-159 on its own copy over its share of the queries
// (schedule(dynamic), as hnsw_slim_client_update_patch.cc:223-226 does for hnsw_slim).  passes > 1 repeats the
// batch; *seconds = median pass.  t_const > 0 fixes every copy's query-quantiser constant.
int refq_search_copies(const char *graph, size_t dim, size_t n, const float *base, int threads,
                       const float *queries, size_t nq, size_t k, size_t ef, double t_const, int passes,
                       uint32_t *out_labels, double *seconds) {
  try {
    if (threads <= 0) threads = omp_get_num_procs();
    K = k;
    std::vector<std::vector<float>> dataset(n);
    for (size_t i = 0; i < n; ++i) dataset[i].assign(base + i * dim, base + (i + 1) * dim);
    hnswlib::L2Space space(dim);
    std::vector<std::unique_ptr<hnswlib::HierarchicalNSWSlimQ<float>>> copies(threads);
    std::string err;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; ++t) {
      try {
        copies[t].reset(new hnswlib::HierarchicalNSWSlimQ<float>(&space, n));
        copies[t]->loadIndex(graph, &space, n);
        copies[t]->setDataset(&dataset);
        copies[t]->setEf(ef);
        if (t_const > 0) copies[t]->query_config_.t_const = t_const;
      } catch (const std::exception &e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) {
      g_err = err;
      return -1;
    }
    std::vector<double> times;
    for (int pass = 0; pass < std::max(1, passes); ++pass) {
      auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic)
      for (size_t i = 0; i < nq; ++i)
        copies[omp_get_thread_num()]->searchKnn(queries + i * dim, k, out_labels + i * k);
      auto t1 = std::chrono::steady_clock::now();
      times.push_back(std::chrono::duration<double>(t1 - t0).count());
    }
    std::sort(times.begin(), times.end());
    if (seconds) *seconds = times[times.size() / 2];
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
