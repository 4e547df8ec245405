// Launch interface of the fp32 traversal kernel (traverse_fp32.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "hs_internal.h"

namespace hs {

struct TraverseParams {
  // HBM-resident index
  const float4 *vec;                       // n x row_chunks float4
  const uint32_t *adj0;                    // n x deg0_stride
  const int32_t *upper_slot;               // n
  const uint32_t *upper_adj[kMaxLevels];   // [l] -> level_count[l] x upper_stride
  const uint32_t *labels;                  // n
  const uint8_t *deleted;                  // n (only read when has_deleted)
  uint32_t n, row_chunks, deg0_stride, upper_stride, enterpoint;
  uint32_t level_count[kMaxLevels];        // rows of upper_adj[l] (nodes with level >= l)
  int32_t maxlevel, threshold_level, has_deleted;
  // query batch
  const float *queries;                    // nq x dim
  uint32_t nq, dim, k, ef;
  uint32_t *out_labels;                    // nq x k
  float *out_dists;                        // nq x k or null
  ScatterDst scatter;                      // extra destinations (sharded path), see hs_internal.h
  // scratch / counters
  unsigned long long *work_counter;        // one slot of the ring of tagged counters (next_ticket)
  uint32_t launch_tag;                     // launch sequence number: the counter's tag
  uint32_t overlap;                        // 1: launched with programmatic stream serialization
  unsigned long long *stats;               // [0] n_dist  [1] n_hops
  uint32_t *per_query;                     // optional nq x 2 (n_dist, n_hops) or null
  // shared-memory carve-up per warp (bytes)
  uint32_t hash_bits, smem_per_warp, off_hash, off_stage, off_query;
  uint32_t flags;                          // bit0: L2 row prefetch, bit1: speculative next-hop prefetch
  // visited hash in global memory (large ef / large dim, see plan_traverse): one table of
  // 2^hash_bits slots per resident warp; null when the tables live in shared memory
  uint32_t *ghash;
};

struct TraverseLaunch {
  int warps_per_cta;
  int resident;          // CTAs the GPU holds at once with this shape (sm_count x CTAs per SM): the grid's upper bound
  int grid;
  size_t smem_bytes;
  bool cvtab;            // compact 16-bit visited table in shared memory (traverse_fp32_c.cu)
  bool ghash;            // visited tables in global memory: the caller points p.ghash at ghash_bytes of scratch
  size_t ghash_bytes;    // grid * warps_per_cta * 4 << hash_bits
  bool may_overlap;      // programmatic stream serialization allowed for this launch shape
};

// Fills the smem carve-up fields of `p` and returns the launch shape.
// hash_bits_override = 0 picks the default for p.ef; ghash_mode: -1 = decide from the shared-memory
// budget (compact 16-bit tables in shared memory where they apply: ef 129..256, dim 96 / 128 rows, n <= 2^24),
// 0 = 32-bit visited tables in shared memory, 1 = in global memory, 2 = compact where it applies else as -1,
// 3 = as -1 but never compact.
int plan_traverse(TraverseParams &p, int metric, int hash_bits_override, int ghash_mode, int sm_count, int nq,
                  TraverseLaunch *out);
// Adapts a plan made for one batch size to another (p must carry the plan's hash_bits).
void resize_traverse_launch(const TraverseParams &p, int nq, TraverseLaunch *l);
int launch_traverse(const TraverseParams &p, int metric, const TraverseLaunch &l, cudaStream_t stream);

}  // namespace hs
