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

// hnsw_slim over a corpus that is SHARDED into per-GPU sub-graphs (no reference analogue: the reference
// builds one graph; SURVEY.md §8(e), north_star (4)).  --shards S --gpus N: the base set is cut into S
// contiguous ranges (labels stay global), rank r = GPU device_ + r holds S / N of them, every sub-graph is
// built on its GPU (hs_build_slim_index_gpu), and ONE host thread drives all ranks: each query batch is
// submitted to every rank's hs_shardgroup (peer memory between the GPUs, hs_shardgroup_connect_local), whose
// traversal kernels exchange the rows and whose merge kernels write the global top-k; rank 0's rows are the
// result.  ef_search is the PER-SHARD ef.
class HnswSlimShardedGpuStrategy : public SolveStrategy {
 public:
  HnswSlimShardedGpuStrategy(std::string source_path, std::string query_path, std::string index_path, PruneParams pp,
                             size_t shards, size_t gpus, size_t batch, int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device), pp_(pp), shards_(shards), gpus_(gpus), batch_(batch) {}

  void solve() override {
    if (shards_ == 0 || gpus_ == 0 || shards_ % gpus_ != 0)
      throw std::runtime_error("--shards must be a positive multiple of --gpus");
    const size_t per = shards_ / gpus_, nq = query_num_, k = K_, dim = data_dim_;
    const size_t batch = std::max<size_t>(1, std::min(batch_, nq));
    std::vector<hs_index *> ix(shards_, nullptr);
    std::vector<hs_shardgroup *> groups(gpus_, nullptr);
    std::vector<std::vector<uint32_t>> side(gpus_);            // result rows of the ranks other than 0
    auto cleanup = [&]() {
      for (auto *g : groups) hs_shardgroup_free(g);
      for (auto *i : ix) hs_free(i);
      for (auto &v : side)
        if (!v.empty()) hs_unpin_host(v.data());
    };
    try {
      auto s_build = std::chrono::system_clock::now();
      hs_build_params bp = make_build_params(M_, ef_construction_, branching_factor_, pp_);
      size_t bytes = 0;
      for (size_t s = 0; s < shards_; ++s) {                   // contiguous label ranges, sizes differing by <= 1
        const size_t lo = data_num_ * s / shards_, hi = data_num_ * (s + 1) / shards_;
        std::vector<uint64_t> labels(hi - lo);
        for (size_t i = lo; i < hi; ++i) labels[i - lo] = i;
        check(hs_build_slim_index_gpu(data_set_.data() + lo * dim, hi - lo, dim, HS_METRIC_L2, &bp, labels.data(),
                                      device_ + (int)(s / per), &ix[s]));
        check(hs_set_ef(ix[s], ef_search_));
        hs_index_info info;
        hs_get_info(ix[s], &info);
        bytes += info.device_bytes;
      }
      auto e_build = std::chrono::system_clock::now();
      std::cout << "build cost: " << time_cost(s_build, e_build) << " (ms) for " << shards_ << " shards on " << gpus_
                << " GPUs\n";
      std::cout << "hnsw_slim index size: " << bytes << " bytes\n";
      for (size_t r = 0; r < gpus_; ++r)
        check(hs_shardgroup_create(&ix[r * per], per, (int)gpus_, (int)r, batch, k, 4, &groups[r]));
      check(hs_shardgroup_connect_local(groups.data(), gpus_));
      for (size_t r = 1; r < gpus_; ++r) {
        side[r].resize(nq * k);
        check(hs_pin_host(side[r].data(), side[r].size() * sizeof(uint32_t)));
      }
      auto s_solve = std::chrono::system_clock::now();
      for (size_t q0 = 0; q0 < nq; q0 += batch) {              // every rank gets every batch, in the same order
        const size_t n = std::min(batch, nq - q0);
        for (size_t r = 0; r < gpus_; ++r) {
          uint32_t *out = (r == 0 ? knn_results_.data() : side[r].data()) + q0 * k;
          check(hs_shardgroup_submit(groups[r], query_set_.data() + q0 * dim, n, out, nullptr));
        }
      }
      for (auto *g : groups) check(hs_shardgroup_wait(g));
      auto e_solve = std::chrono::system_clock::now();
      std::cout << "solve cost: " << time_cost(s_solve, e_solve) << " (ms)\n";
      std::cout << "query cost: " << std::chrono::duration<double>(e_solve - s_solve).count() << "\n";
      for (size_t r = 1; r < gpus_; ++r)                        // every rank merged the same rows
        if (side[r] != knn_results_) throw std::runtime_error("ranks disagree on the merged result");
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  }

 private:
  PruneParams pp_;
  size_t shards_, gpus_, batch_;
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
