// Exact kNN (ground-truth path), cross-part top-k merge and recall — launch interface.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace hs {

// BruteforceSearch<float>::searchKnn for a batch (bruteforce.h:106-135): exact top-k by
// (distance, label) ascending, label of base row i = i.  Device buffers, async on stream.
int bruteforce_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq,
                      size_t k, int metric, uint32_t *d_labels, float *d_dists, cudaStream_t stream);

// tcgen05 path (bruteforce_tc.cu): same results as the scan kernel.  Queries it cannot certify are
// returned in *fallback_list / *fallback_count (device memory inside *scratch_to_free, which the
// caller releases with cudaFreeAsync after the fallback scan has been enqueued).
bool bruteforce_tc_applicable(size_t n, size_t dim, size_t nq, size_t k);
int bruteforce_tc_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq, size_t k,
                         int metric, uint32_t *d_labels, float *d_dists, cudaStream_t stream,
                         uint32_t **fallback_list, unsigned int **fallback_count, void **scratch_to_free);

long long bruteforce_last_tc_fallback();

// n_parts consecutive [nq x k] (label, dist) tables -> global top-k per query.
int topk_merge_device(const uint32_t *d_labels_in, const float *d_dists_in, size_t n_parts, size_t nq,
                      size_t k, uint32_t *d_labels_out, float *d_dists_out, cudaStream_t stream);

// SolveStrategy::recall (solve_strategy.h:67-103): writes the hit count to *d_hits.
int recall_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq,
                  const uint32_t *d_knn, size_t K, const uint32_t *d_gt, size_t gt_k, int metric,
                  unsigned long long *d_hits, cudaStream_t stream);

}  // namespace hs
