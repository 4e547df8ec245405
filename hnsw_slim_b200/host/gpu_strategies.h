// GPU strategies behind the reference's strategy names:
//   hnsw_slim      -> HnswSlimGpuStrategy     (include/strategy/hnsw_slim_strategy.h:34-120)
//   hnsw_slimq     -> HnswSlimQGpuStrategy    (include/strategy/hnsw_slimq_strategy.h:48-165)
//   hnsw           -> HnswGpuStrategy         (include/strategy/hnsw_strategy.h:15-61)
//   hnsw_slimzero  -> HnswSlimZeroGpuStrategy (include/strategy/hnsw_slimzero_strategy.h:38-140; load only)
//   bruteforce     -> BruteForceGpu           (include/strategy/brute_force_strategy.h:15-45)
// Each solve() keeps the reference's build-or-load logic and console output; the serial
// per-query loop becomes ONE hs_search_batch call.
#pragma once
#include <filesystem>

#include "solve_strategy.h"

struct PruneParams {      // main.cc:58-70
  int threshold_level = 0;
  float top_degree_percent0 = 0.02f, top_degree_percent = 0.02f;
  size_t top_M0 = 32, low_m0 = 8, top_M = 16, low_m = 4;
};

inline hs_build_params make_build_params(size_t M, size_t efc, const std::string &bf, const PruneParams &pp) {
  hs_build_params p;
  hs_build_params_default(&p);
  p.M = M;
  p.ef_construction = efc;
  p.branching_factor = bf.c_str();
  p.threshold_level = pp.threshold_level;
  p.top_degree_percent0 = pp.top_degree_percent0;
  p.top_degree_percent = pp.top_degree_percent;
  p.top_M0 = pp.top_M0;
  p.low_m0 = pp.low_m0;
  p.top_M = pp.top_M;
  p.low_m = pp.low_m;
  return p;
}

inline void check(int rc) {
  if (rc != HS_OK) throw std::runtime_error(hs_last_error());
}

class HnswSlimGpuStrategy : public SolveStrategy {
 public:
  HnswSlimGpuStrategy(std::string source_path, std::string query_path, std::string index_path, PruneParams pp,
                      int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device), pp_(pp) {}

  void solve() override {
    std::cout << "index path: " << index_path_ << std::endl;
    if (!std::filesystem::exists(index_path_)) {            // hnsw_slim_strategy.h:56-95: build, prune, save
      auto s_build = std::chrono::system_clock::now();
      std::filesystem::path p(index_path_);
      if (p.has_parent_path()) std::filesystem::create_directories(p.parent_path());
      hs_build_params bp = make_build_params(M_, ef_construction_, branching_factor_, pp_);
      check(hs_build_slim_graph(data_set_.data(), data_num_, data_dim_, HS_METRIC_L2, &bp, nullptr,
                                index_path_.c_str()));
      auto e_build = std::chrono::system_clock::now();
      std::cout << "build cost: " << time_cost(s_build, e_build) << " (ms)\n";
      std::cout << "save index: " + index_path_ << std::endl;
    }
    hs_index *ix = nullptr;                                  // loadIndex, slim.h:753
    check(hs_load(index_path_.c_str(), HS_KIND_SLIM, HS_METRIC_L2, data_dim_, nullptr, 0, device_, &ix));
    hs_index_info info;
    hs_get_info(ix, &info);
    std::cout << "hnsw_slim index size: " << info.device_bytes << " bytes\n";
    check(hs_set_ef(ix, ef_search_));                        // setEf, slim.h:193
    auto s_solve = std::chrono::system_clock::now();
    const int rc = hs_search_batch(ix, query_set_.data(), query_num_, K_, knn_results_.data(), nullptr);
    auto e_solve = std::chrono::system_clock::now();         // the loop of hnsw_slim_strategy.h:112-114
    if (rc != HS_OK) {
      hs_free(ix);
      check(rc);
    }
    std::cout << "solve cost: " << time_cost(s_solve, e_solve) << " (ms)\n";
    std::cout << "query cost: " << std::chrono::duration<double>(e_solve - s_solve).count() << "\n";
    hs_free(ix);
  }

 private:
  PruneParams pp_;
};

// hnsw_strategy.h:15-61: the un-pruned hnswlib index.  Same search kernel (HierarchicalNSW::searchKnn,
// hnsw.h:1378-1440, is the slim search with threshold_level 0 on full lists); the file format is
// HierarchicalNSW::saveIndex's, read by hs_load(kind = HS_KIND_HNSW) and written by hs_build_hnsw_graph.
class HnswGpuStrategy : public SolveStrategy {
 public:
  HnswGpuStrategy(std::string source_path, std::string query_path, std::string index_path, int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device) {}

  void solve() override {
    if (!std::filesystem::exists(index_path_)) {            // hnsw_strategy.h:24-45
      auto s_build = std::chrono::system_clock::now();
      std::filesystem::path p(index_path_);
      if (p.has_parent_path()) std::filesystem::create_directories(p.parent_path());
      hs_build_params bp = make_build_params(M_, ef_construction_, branching_factor_, PruneParams());
      check(hs_build_hnsw_graph(data_set_.data(), data_num_, data_dim_, HS_METRIC_L2, &bp, nullptr,
                                index_path_.c_str()));
      auto e_build = std::chrono::system_clock::now();
      std::cout << "build cost: " << time_cost(s_build, e_build) << " (ms)\n";
      std::cout << "save index: " + index_path_ << std::endl;
    }
    hs_index *ix = nullptr;                                  // loadIndex, hnsw.h:781
    check(hs_load(index_path_.c_str(), HS_KIND_HNSW, HS_METRIC_L2, data_dim_, nullptr, 0, device_, &ix));
    hs_index_info info;
    hs_get_info(ix, &info);
    std::cout << "hnsw index size: " << info.device_bytes << " bytes\n";
    check(hs_set_ef(ix, ef_search_));
    auto s_solve = std::chrono::system_clock::now();
    const int rc = hs_search_batch(ix, query_set_.data(), query_num_, K_, knn_results_.data(), nullptr);
    auto e_solve = std::chrono::system_clock::now();         // the loop of hnsw_strategy.h:49-58
    if (rc != HS_OK) {
      hs_free(ix);
      check(rc);
    }
    std::cout << "solve cost: " << time_cost(s_solve, e_solve) << " (ms)\n";
    hs_free(ix);
  }
};

// hnsw_slimzero_strategy.h:38-140.  The index class differs from hnsw_slim only in how
// convertFromHNSW prunes (min in-degree, hnswalg_slimzero.h:928-1158) — a build-time step of the
// reference that stays on the CPU; the file format (:701-735) and searchKnn (:1675-1771) are
// hnsw_slim's, so an index the reference built is loaded and searched here as HS_KIND_SLIM.
class HnswSlimZeroGpuStrategy : public SolveStrategy {
 public:
  HnswSlimZeroGpuStrategy(std::string source_path, std::string query_path, std::string index_path, int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device) {}

  void solve() override {
    std::cout << "index path: " << index_path_ << std::endl;
    if (!std::filesystem::exists(index_path_))
      throw std::runtime_error("hnsw_slimzero: " + index_path_ + " not found — build it with the reference "
                               "(HierarchicalNSWSlimZero::convertFromHNSW); the GPU engine loads and searches it");
    hs_index *ix = nullptr;                                  // loadIndex, hnswalg_slimzero.h:737
    check(hs_load(index_path_.c_str(), HS_KIND_SLIM, HS_METRIC_L2, data_dim_, nullptr, 0, device_, &ix));
    hs_index_info info;
    hs_get_info(ix, &info);
    std::cout << "hnsw_slim_zero index size: " << info.device_bytes << " bytes\n";
    check(hs_set_ef(ix, ef_search_));
    auto s_solve = std::chrono::system_clock::now();
    const int rc = hs_search_batch(ix, query_set_.data(), query_num_, K_, knn_results_.data(), nullptr);
    auto e_solve = std::chrono::system_clock::now();         // the loop of hnsw_slimzero_strategy.h:131-133
    if (rc != HS_OK) {
      hs_free(ix);
      check(rc);
    }
    std::cout << "solve cost: " << time_cost(s_solve, e_solve) << " (ms)\n";
    std::cout << "query cost: " << std::chrono::duration<double>(e_solve - s_solve).count() << "\n";
    hs_free(ix);
  }
};

class HnswSlimQGpuStrategy : public SolveStrategy {
 public:
  HnswSlimQGpuStrategy(std::string source_path, std::string query_path, std::string index_path, PruneParams pp,
                       int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device), pp_(pp), source_path_(source_path) {
    // hnsw_slimq_strategy.h:42-45
    centroid_path_ = source_path;
    cluster_path_ = source_path;
    const size_t pos = source_path.find("_base.fvecs");
    if (pos != std::string::npos) {
      centroid_path_.replace(pos, 11, "_centroids_16.fvecs");
      cluster_path_.replace(pos, 11, "_clusterids_16.ivecs");
    }
  }

  void solve() override {
    std::cout << "index path: " << index_path_ << std::endl;
    if (!std::filesystem::exists(index_path_)) {            // hnsw_slimq_strategy.h:75-142
      std::filesystem::path p(index_path_);
      if (p.has_parent_path()) std::filesystem::create_directories(p.parent_path());
      hs_build_params bp = make_build_params(M_, ef_construction_, branching_factor_, pp_);
      std::vector<float> cent;
      std::vector<uint32_t> cid;
      uint32_t nc = 16, cd = 0, cn = 0, c1 = 0;
      const bool have = std::filesystem::exists(centroid_path_) && std::filesystem::exists(cluster_path_);
      if (have) {
        ReadData(centroid_path_, cent, nc, cd);
        ReadData(cluster_path_, cid, cn, c1);
        if (cd != data_dim_ || cn != data_num_) throw std::runtime_error("centroid / cluster-id files do not match the base set");
      }
      auto s_build = std::chrono::system_clock::now();
      check(hs_build_slimq_graph(data_set_.data(), data_num_, data_dim_, &bp, have ? cent.data() : nullptr, nc,
                                 have ? cid.data() : nullptr, nullptr, index_path_.c_str()));
      auto e_build = std::chrono::system_clock::now();
      std::cout << "convert hnsw to hnsw_slimq cost: " << time_cost(s_build, e_build) << " (ms)\n";
      std::cout << "save index: " + index_path_ << std::endl;
    }
    hs_index *ix = nullptr;                                  // loadIndex + setDataset, slimq.h:1218, :303
    check(hs_load(index_path_.c_str(), HS_KIND_SLIMQ, HS_METRIC_L2, data_dim_, data_set_.data(), data_num_, device_,
                  &ix));
    hs_index_info info;
    hs_get_info(ix, &info);
    std::cout << "hnsw_slimq index size: " << info.device_bytes << " bytes\n";
    check(hs_set_ef(ix, ef_search_));
    auto s_solve = std::chrono::system_clock::now();
    const int rc = hs_search_batch(ix, query_set_.data(), query_num_, K_, knn_results_.data(), nullptr);
    auto e_solve = std::chrono::system_clock::now();         // the loop of hnsw_slimq_strategy.h:157-159
    if (rc != HS_OK) {
      hs_free(ix);
      check(rc);
    }
    std::cout << "solve cost: " << time_cost(s_solve, e_solve) << " (ms)\n";
    std::cout << "query cost: " << std::chrono::duration<double>(e_solve - s_solve).count() << "\n";
    hs_free(ix);
  }

 private:
  PruneParams pp_;
  std::string source_path_, centroid_path_, cluster_path_;
};

// brute_force_strategy.h:15-45: exact k-NN of every query, rows written FARTHEST first to gt_path
class BruteForceGpu : public SolveStrategy {
 public:
  BruteForceGpu(std::string source_path, std::string query_path, std::string index_path, std::string gt_path,
                size_t gt_k = 100, int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device), gt_path_(gt_path), gt_k_(gt_k) {}

  void solve() override {
    const size_t k = std::min<size_t>(gt_k_, data_num_);
    std::vector<uint32_t> rows((size_t)query_num_ * k);
    auto s = std::chrono::system_clock::now();
    check(hs_bruteforce_knn(data_set_.data(), data_num_, data_dim_, query_set_.data(), query_num_, k, HS_METRIC_L2,
                            device_, rows.data(), nullptr));
    auto e = std::chrono::system_clock::now();
    std::cout << "solve cost: " << time_cost(s, e) << " (ms)\n";
    for (size_t i = 0; i < query_num_; ++i) {                // nearest first -> the reference's farthest first
      uint32_t *r = rows.data() + i * k;
      for (size_t a = 0, b = k - 1; a < b; ++a, --b) std::swap(r[a], r[b]);
      for (size_t j = 0; j < K_ && j < k; ++j) knn_results_[i * K_ + j] = r[k - 1 - j];
    }
    WriteData(gt_path_, rows, query_num_, (uint32_t)k);
  }

 private:
  std::string gt_path_;
  size_t gt_k_;
};
