// Serving + delta patches behind the reference's server/client pair:
//   server handlers   /query, /setEf                    hnsw_slim_server.cc:69-115, hnsw_slim_server_patch.cc:133-180
//   client update     updateIndex -> patchFromStream    hnsw_slim_client_update_patch.cc:56-81,113,179
// HnswSlimGpuService is what a handler thread holds: query() is the body of the /query handler (one vector in, k
// labels out, thread-safe), setEf() the /setEf handler, patchFromStream() the client's update step — on the
// HBM-resident index, through hs_service (requests of many threads become batches) and hs_patch_apply.
// HnswSlimServeGpuStrategy drives it the way the reference's client does: the query set goes through `threads`
// concurrent single-query callers; --patches=a.bin,b.bin applies the server's patch streams first (vectors of the
// new rows come from the client's own base file by label, as hnsw_slim_client_update_patch.cc:150-162 has it;
// --patch_inline_last: the last stream is the /getLastBatch form with the vectors inline).
#pragma once
#include <atomic>
#include <mutex>
#include <fstream>
#include <sstream>
#include <thread>

#include "gpu_strategies.h"

class HnswSlimGpuService {
 public:
  // HierarchicalNSWSlim(&space, index_path, false, max_elements) as the server / client construct it
  // (hnsw_slim_server.cc:63, hnsw_slim_client_update_patch.cc:113) + the serving front
  HnswSlimGpuService(const std::string &index_path, size_t dim, size_t max_elements, size_t ef_search, size_t k_max,
                     size_t max_batch = 4096, unsigned max_wait_us = 0, int device = 0)
      : dim_(dim) {
    check(hs_load_reserve(index_path.c_str(), HS_KIND_SLIM, HS_METRIC_L2, dim, max_elements, device, &ix_));
    int rc = hs_set_ef(ix_, ef_search);
    if (rc == HS_OK) rc = hs_service_create(ix_, max_batch, max_wait_us, k_max, &svc_);
    if (rc != HS_OK) {
      hs_free(ix_);
      check(rc);
    }
  }
  ~HnswSlimGpuService() {
    hs_service_free(svc_);
    hs_free(ix_);
  }
  HnswSlimGpuService(const HnswSlimGpuService &) = delete;
  HnswSlimGpuService &operator=(const HnswSlimGpuService &) = delete;

  // hnsw_slim.searchKnn(vec.data(), k, knn_results.data()) of the /query handler
  void query(const float *vec, size_t k, uint32_t *labels_out) { check(hs_service_query(svc_, vec, k, labels_out, nullptr)); }
  void setEf(size_t ef) { check(hs_service_set_ef(svc_, ef)); }
  // patchFromStream(in, new_data) / (in, data_set) / (in, true): slim.h:2343-2388, :2206-2253, :2292-2340
  hs_patch_info patchFromStream(std::istream &in, const float *rows, const uint64_t *row_labels, size_t n_rows,
                                bool rows_inline = false) {
    std::ostringstream body;
    body << in.rdbuf();
    const std::string bytes = body.str();
    hs_patch_info info{};
    check(hs_service_patch(svc_, bytes.data(), bytes.size(), rows_inline ? HS_PATCH_INLINE_ROWS : 0u, rows, row_labels,
                           n_rows, &info));
    return info;
  }
  hs_index_info info() const {
    hs_index_info i;
    hs_get_info(ix_, &i);
    return i;
  }
  hs_service_stats stats() const {
    hs_service_stats s{};
    hs_service_get_stats(svc_, &s);
    return s;
  }

 private:
  hs_index *ix_ = nullptr;
  hs_service *svc_ = nullptr;
  size_t dim_;
};

class HnswSlimServeGpuStrategy : public SolveStrategy {
 public:
  HnswSlimServeGpuStrategy(std::string source_path, std::string query_path, std::string index_path, size_t threads,
                           size_t max_batch, unsigned max_wait_us, std::string patches, bool patch_inline_last,
                           int device = 0)
      : SolveStrategy(source_path, query_path, index_path, device), threads_(std::max<size_t>(1, threads)),
        max_batch_(max_batch), max_wait_us_(max_wait_us), patch_inline_last_(patch_inline_last) {
    std::stringstream ss(patches);
    for (std::string item; std::getline(ss, item, ',');)
      if (!item.empty()) patch_files_.push_back(item);
  }

  void solve() override {
    std::cout << "index path: " << index_path_ << std::endl;
    HnswSlimGpuService svc(index_path_, data_dim_, data_num_, ef_search_, K_, max_batch_, max_wait_us_, device_);
    std::cout << "hnsw_slim index size: " << svc.info().device_bytes << " bytes, " << svc.info().n << " of " << data_num_
              << " elements\n";
    for (size_t p = 0; p < patch_files_.size(); ++p) {        // the update loop of hnsw_slim_client_update_patch.cc:147-168
      std::ifstream in(patch_files_[p], std::ios::binary);
      if (!in) throw std::runtime_error("cannot open patch " + patch_files_[p]);
      const bool inl = patch_inline_last_ && p + 1 == patch_files_.size();
      auto s = std::chrono::system_clock::now();
      const hs_patch_info pi = svc.patchFromStream(in, inl ? nullptr : data_set_.data(), nullptr, inl ? 0 : data_num_, inl);
      auto e = std::chrono::system_clock::now();
      std::cout << "patch " << patch_files_[p] << ": " << pi.changed_old << " + " << pi.changed_new << " nodes, "
                << pi.n_before << " -> " << pi.n_after << " elements, " << time_cost(s, e) << " (ms)\n";
    }
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    std::string error;
    std::mutex err_mu;
    auto s_solve = std::chrono::system_clock::now();
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads_; ++t)
      pool.emplace_back([&] {                                 // one handler thread: a query at a time
        try {
          for (size_t i = next++; i < query_num_ && !failed; i = next++)
            svc.query(query_set_.data() + i * (size_t)data_dim_, K_, knn_results_.data() + i * K_);
        } catch (const std::exception &e) {
          std::lock_guard<std::mutex> g(err_mu);
          failed = true;
          error = e.what();
        }
      });
    for (auto &t : pool) t.join();
    auto e_solve = std::chrono::system_clock::now();
    if (failed) throw std::runtime_error(error);
    const hs_service_stats st = svc.stats();
    std::cout << "solve cost: " << time_cost(s_solve, e_solve) << " (ms)\n";
    std::cout << "served " << st.queries << " queries from " << threads_ << " threads in " << st.batches
              << " batches (largest " << st.max_batch << ")\n";
  }

 private:
  size_t threads_, max_batch_;
  unsigned max_wait_us_;
  bool patch_inline_last_;
  std::vector<std::string> patch_files_;
};
