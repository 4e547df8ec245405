// Launch interface of the hnsw_slimq traversal kernel (traverse_slimq.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "hs_internal.h"

namespace hs {

struct TraverseQParams {
  // HBM-resident index
  const uint2 *qrec;                       // n x rec_words uint2: [code u64[W]][f_add, f_rescale][cluster, 0]
  const float4 *vec;                       // n x row_chunks float4: raw rows for the exact rerank
  const uint32_t *adj0;                    // n x deg0_stride
  const int32_t *upper_slot;               // n
  const uint32_t *upper_adj[kMaxLevels];   // [l] -> level_count[l] x upper_stride
  const uint32_t *labels;                  // n
  const float *centroids;                  // num_cluster x padded_dim (rotated)
  const uint8_t *flip;                     // 4 * padded_dim / 8 bytes
  uint32_t n, row_chunks, deg0_stride, upper_stride, enterpoint;
  int32_t maxlevel, threshold_level;
  uint32_t padded_dim, trunc_dim, num_cluster, words, rec_words;
  float fht_fac;                           // 1 / sqrt(trunc_dim)
  double t_const;                          // query-quantiser constant
  // query batch
  const float *queries;                    // nq x dim
  uint32_t nq, dim, k, ef;
  uint32_t *out_labels;                    // nq x k
  float *out_dists;                        // nq x k or null
  ScatterDst scatter;                      // extra destinations (sharded path), see hs_internal.h
  // optional dump of the per-query preparation (hs_slimq_prepare); search is skipped when set
  float *prep_rotated;                     // nq x padded_dim
  unsigned long long *prep_planes;         // nq x words*4
  float *prep_scal;                        // nq x 3
  float *prep_q2c;                         // nq x num_cluster
  // scratch / counters
  unsigned long long *work_counter;        // one slot of the ring of tagged counters (next_ticket)
  uint32_t launch_tag;                     // launch sequence number: the counter's tag
  uint32_t overlap;                        // 1: launched with programmatic stream serialization
  unsigned long long *stats;               // [0] n_est [1] n_hops [2] n_rerank
  uint32_t *per_query;                     // optional nq x 2 (n_est, n_hops)
  // shared-memory carve-up per warp (bytes)
  uint32_t smem_per_warp, off_buf, off_g2c, off_planes, off_topk;
  uint32_t flags;                          // bit0: speculative L2 prefetch of the next hop's code records
};

struct TraverseQLaunch {
  int warps_per_cta;
  int grid;
  size_t smem_bytes;
};

int plan_traverse_slimq(TraverseQParams &p, int sm_count, int nq, TraverseQLaunch *out);
int launch_traverse_slimq(const TraverseQParams &p, const TraverseQLaunch &l, cudaStream_t stream);

}  // namespace hs
