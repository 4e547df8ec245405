// The strategy plugin surface of the reference (include/strategy/solve_strategy.h:9-127) over the
// C ABI of the B200 engine: same constructor, solve(), recall(gt_path), setEf, save_knn/read_knn.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hnswslim_b200.h"
#include "core.h"
#include "util.h"

class SolveStrategy {
 public:
  // solve_strategy.h:11-25
  SolveStrategy(std::string source_path, std::string query_path, std::string index_path, int device = 0)
      : device_(device) {
    ReadData(source_path, data_set_, data_num_, data_dim_);
    ReadData(query_path, query_set_, query_num_, query_dim_);
    knn_results_.resize((size_t)query_num_ * K);
    // page-lock the query and result blocks once: hs_search_batch then reads / writes them in place
    // (no staging copies).  Best effort — without a device the search itself reports the error.
    pinned_ = hs_pin_host(query_set_.data(), query_set_.size() * sizeof(float)) == HS_OK &&
              hs_pin_host(knn_results_.data(), knn_results_.size() * sizeof(uint32_t)) == HS_OK;
    M_ = M;
    M0_ = M0;
    ef_construction_ = EF_CONSTRUCTION;
    ef_search_ = EF_SEARCH;
    K_ = K;
    branching_factor_ = BRANCHING_FACTOR;
    threshold_level_ = THRESHOLD_LEVEL;
    index_path_ = index_path;
  }
  virtual ~SolveStrategy() {
    hs_unpin_host(query_set_.data());
    hs_unpin_host(knn_results_.data());
  }

  virtual void solve() = 0;

  void setEf(size_t ef) { ef_search_ = ef; }

  void read_knn(std::string knn_path) {
    uint32_t num, dim;
    hs_unpin_host(knn_results_.data());              // ReadData may reallocate the block
    ReadData(knn_path, knn_results_, num, dim);
    hs_pin_host(knn_results_.data(), knn_results_.size() * sizeof(uint32_t));
  }
  void save_knn(std::string knn_path) { WriteData(knn_path, knn_results_, query_num_, (uint32_t)K_); }

  // solve_strategy.h:67-103 — ground-truth ids re-ranked by exact L2, ties -> smaller id, first K;
  // hit / (nq * K).  Runs on the device (hs_recall).
  void recall(std::string gt_path) {
    uint32_t gt_num, gt_dim;
    std::vector<uint32_t> gt_set;
    ReadData(gt_path, gt_set, gt_num, gt_dim);
    if (gt_num < query_num_ || gt_dim < K_) throw std::runtime_error("ground truth has too few rows / columns");
    double r = 0;
    if (hs_recall(data_set_.data(), data_num_, data_dim_, query_set_.data(), query_num_, knn_results_.data(), K_,
                  gt_set.data(), gt_dim, HS_METRIC_L2, device_, &r) != HS_OK)
      throw std::runtime_error(hs_last_error());
    std::cout << "Recall: " << (float)r << std::endl;
  }

  const std::vector<uint32_t> &knn_results() const { return knn_results_; }

 protected:
  std::vector<float> data_set_;      // data_num_ x data_dim_
  uint32_t data_num_ = 0, data_dim_ = 0;
  size_t M_, M0_, ef_construction_;
  std::string branching_factor_;
  size_t threshold_level_;
  std::vector<float> query_set_;     // query_num_ x query_dim_
  uint32_t query_num_ = 0, query_dim_ = 0;
  size_t ef_search_;
  std::vector<uint32_t> knn_results_;   // query_num_ x K_
  bool pinned_ = false;
  size_t K_;
  std::string index_path_;
  int device_;
};
