// Exact kNN (the recall_knn / bruteforce ground-truth path), cross-part top-k merge, recall.
//
// bf_scan_kernel restates BruteforceSearch<float>::searchKnn (bruteforce.h:106-135) for a
// 128-query tile per CTA: each thread owns one query and a size-k max-heap ordered by the
// (distance, label) pair — the order of the reference's priority_queue<pair<float,size_t>> —
// so a row replaces the heap top iff its pair is smaller, which is the reference's
// "dist <= lastdist, emplace, pop" at the k-th boundary.  Distances are fp32 FMA chains in
// index order (oracle: HSO_ORDER_SEQFMA), so ids AND distances are bit-reproducible.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>

#include "bruteforce.cuh"
#include "hs_internal.h"

namespace hs {
namespace {

constexpr int QT = 128;     // queries per CTA (one per thread)
constexpr int RT = 32;      // base rows per tile
constexpr int SLAB = 64;    // dims staged per pass
constexpr int QPAD = QT + 1;

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// max-heap of keys in thread-private global memory
__device__ __forceinline__ void heap_push(uint64_t *h, uint32_t &sz, uint64_t key) {
  uint32_t i = sz++;
  while (i > 0) {
    const uint32_t parent = (i - 1) >> 1;
    const uint64_t pk = h[parent];
    if (pk >= key) break;
    h[i] = pk;
    i = parent;
  }
  h[i] = key;
}
__device__ __forceinline__ void heap_replace_top(uint64_t *h, uint32_t sz, uint64_t key) {
  uint32_t i = 0;
  for (;;) {
    const uint32_t l = 2 * i + 1, r = l + 1;
    if (l >= sz) break;
    uint32_t c = l;
    uint64_t ck = h[l];
    if (r < sz) {
      const uint64_t rk = h[r];
      if (rk > ck) {
        ck = rk;
        c = r;
      }
    }
    if (ck <= key) break;
    h[i] = ck;
    i = c;
  }
  h[i] = key;
}

template <int METRIC>
__global__ void __launch_bounds__(QT) bf_scan_kernel(const float *__restrict__ base, uint32_t n, uint32_t dim,
                                                      const float *__restrict__ queries, uint32_t nq, uint32_t k,
                                                      uint32_t rows_per_split, uint64_t *__restrict__ partial,
                                                      const uint32_t *__restrict__ qmap,
                                                      const unsigned int *__restrict__ qcount) {
  const uint32_t heap_stride_nq = nq;
  // qmap / qcount (optional): only the *qcount queries listed in qmap are scanned (the fallback
  // list of the tcgen05 path); heaps are indexed by list position
  if (qcount) nq = min(nq, *qcount);
  if (blockIdx.x * QT >= nq) return;
  extern __shared__ __align__(16) float sm[];
  float *qs = sm;                     // [SLAB][QPAD]   qs[d][query]
  float *xs = sm + SLAB * QPAD;       // [RT][SLAB]
  const int t = threadIdx.x;
  const uint32_t q0 = blockIdx.x * QT;
  const uint32_t my_q = q0 + t;
  const uint32_t split = blockIdx.y;
  const uint32_t r_begin = split * rows_per_split;
  const uint32_t r_end = min(n, r_begin + rows_per_split);
  const uint32_t n_slabs = (dim + SLAB - 1) / SLAB;

  uint64_t *heap = partial + ((size_t)split * heap_stride_nq + min(my_q, nq - 1)) * k;
  uint32_t hsz = 0;
  uint64_t top = ~0ull;

  auto load_q_slab = [&](uint32_t s) {
    // coalesced: consecutive threads read consecutive dims of one query
    for (uint32_t i = t; i < QT * SLAB; i += QT) {
      const uint32_t qq = i / SLAB, d = i % SLAB;
      const uint32_t gd = s * SLAB + d, gq = q0 + qq;
      float v = 0.f;
      if (gd < dim && gq < nq) v = __ldg(queries + (size_t)(qmap ? __ldg(qmap + gq) : gq) * dim + gd);
      qs[d * QPAD + qq] = v;
    }
  };
  if (n_slabs == 1) load_q_slab(0);

  for (uint32_t r0 = r_begin; r0 < r_end; r0 += RT) {
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.f;
    for (uint32_t s = 0; s < n_slabs; ++s) {
      __syncthreads();
      if (n_slabs > 1) load_q_slab(s);
      for (uint32_t i = t; i < RT * SLAB; i += QT) {
        const uint32_t r = i / SLAB, d = i % SLAB;
        const uint32_t gd = s * SLAB + d, gr = r0 + r;
        xs[r * SLAB + d] = (gd < dim && gr < r_end) ? __ldg(base + (size_t)gr * dim + gd) : 0.f;
      }
      __syncthreads();
#pragma unroll 2
      for (int d = 0; d < SLAB; d += 4) {
        const float qa = qs[(d + 0) * QPAD + t], qb = qs[(d + 1) * QPAD + t];
        const float qc = qs[(d + 2) * QPAD + t], qd = qs[(d + 3) * QPAD + t];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float4 x = *reinterpret_cast<const float4 *>(xs + r * SLAB + d);
          if (METRIC == HS_METRIC_L2) {
            float e;
            e = __fsub_rn(qa, x.x); acc[r] = __fmaf_rn(e, e, acc[r]);
            e = __fsub_rn(qb, x.y); acc[r] = __fmaf_rn(e, e, acc[r]);
            e = __fsub_rn(qc, x.z); acc[r] = __fmaf_rn(e, e, acc[r]);
            e = __fsub_rn(qd, x.w); acc[r] = __fmaf_rn(e, e, acc[r]);
          } else {
            acc[r] = __fmaf_rn(qa, x.x, acc[r]);
            acc[r] = __fmaf_rn(qb, x.y, acc[r]);
            acc[r] = __fmaf_rn(qc, x.z, acc[r]);
            acc[r] = __fmaf_rn(qd, x.w, acc[r]);
          }
        }
      }
    }
    if (my_q < nq) {
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const uint32_t row = r0 + r;
        if (row < r_end) {
          const float dist = METRIC == HS_METRIC_IP ? __fsub_rn(1.0f, acc[r]) : acc[r];
          const uint64_t key = ((uint64_t)f2ord(dist) << 32) | row;
          if (hsz < k) {
            heap_push(heap, hsz, key);
            if (hsz == k) top = heap[0];
          } else if (key < top) {
            heap_replace_top(heap, hsz, key);
            top = heap[0];
          }
        }
      }
    }
  }
  if (my_q < nq)
    for (uint32_t i = hsz; i < k; ++i) heap[i] = ~0ull;
}

// ---- CTA-wide bitonic sort of P (power of two) 64-bit keys in shared memory ----
__device__ __forceinline__ void bitonic_sort(uint64_t *keys, uint32_t P) {
  for (uint32_t size = 2; size <= P; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < P / 2; i += blockDim.x) {
        const uint32_t lo = 2 * i - (i & (stride - 1));
        const uint32_t hi = lo + stride;
        const bool up = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a > b) == up) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// one CTA per query: sort n_parts*k candidate keys, emit the k smallest
__global__ void merge_keys_kernel(const uint64_t *__restrict__ partial, uint32_t n_parts, uint32_t nq, uint32_t k,
                                  uint32_t P, uint32_t *__restrict__ out_labels, float *__restrict__ out_dists,
                                  const uint32_t *__restrict__ qmap, const unsigned int *__restrict__ qcount,
                                  uint32_t heap_stride_nq) {
  extern __shared__ __align__(16) uint64_t keys[];
  uint32_t q = blockIdx.x;
  if (qcount && q >= *qcount) return;
  const uint32_t slot = q;
  if (qmap) q = qmap[slot];
  (void)nq;
  const uint32_t total = n_parts * k;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t v = ~0ull;
    if (i < total) v = partial[((size_t)(i / k) * heap_stride_nq + slot) * k + (i % k)];
    keys[i] = v;
  }
  bitonic_sort(keys, P);
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t e = keys[i];
    const bool ok = e != ~0ull;
    out_labels[(size_t)q * k + i] = ok ? (uint32_t)e : 0xFFFFFFFFu;
    if (out_dists) out_dists[(size_t)q * k + i] = ok ? ord2f((uint32_t)(e >> 32)) : __int_as_float(0x7f800000);
  }
}

__global__ void merge_pairs_kernel(const uint32_t *__restrict__ labels_in, const float *__restrict__ dists_in,
                                   uint32_t n_parts, uint32_t nq, uint32_t k, uint32_t P,
                                   uint32_t *__restrict__ out_labels, float *__restrict__ out_dists) {
  extern __shared__ __align__(16) uint64_t keys[];
  const uint32_t q = blockIdx.x;
  const uint32_t total = n_parts * k;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t v = ~0ull;
    if (i < total) {
      const size_t src = ((size_t)(i / k) * nq + q) * k + (i % k);
      const uint32_t lab = labels_in[src];
      if (lab != 0xFFFFFFFFu) v = ((uint64_t)f2ord(dists_in[src]) << 32) | lab;
    }
    keys[i] = v;
  }
  bitonic_sort(keys, P);
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t e = keys[i];
    const bool ok = e != ~0ull;
    out_labels[(size_t)q * k + i] = ok ? (uint32_t)e : 0xFFFFFFFFu;
    if (out_dists) out_dists[(size_t)q * k + i] = ok ? ord2f((uint32_t)(e >> 32)) : __int_as_float(0x7f800000);
  }
}

// SolveStrategy::recall (solve_strategy.h:67-103), one CTA per query
template <int METRIC>
__global__ void recall_kernel(const float *__restrict__ base, uint32_t n, uint32_t dim,
                              const float *__restrict__ queries, const uint32_t *__restrict__ knn, uint32_t K,
                              const uint32_t *__restrict__ gt, uint32_t gt_k, uint32_t P,
                              unsigned long long *hits) {
  extern __shared__ __align__(16) uint64_t keys[];
  const uint32_t q = blockIdx.x;
  const float *qv = queries + (size_t)q * dim;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t v = ~0ull;
    if (i < gt_k) {
      const uint32_t g = gt[(size_t)q * gt_k + i];
      if (g < n) {
        const float *x = base + (size_t)g * dim;
        float acc = 0.f;
        for (uint32_t d = 0; d < dim; ++d) {
          if (METRIC == HS_METRIC_L2) {
            const float e = __fsub_rn(__ldg(qv + d), __ldg(x + d));
            acc = __fmaf_rn(e, e, acc);
          } else {
            acc = __fmaf_rn(__ldg(qv + d), __ldg(x + d), acc);
          }
        }
        const float dist = METRIC == HS_METRIC_IP ? __fsub_rn(1.0f, acc) : acc;
        v = ((uint64_t)f2ord(dist) << 32) | g;      // pair sort: ties -> smaller id (:87)
      }
    }
    keys[i] = v;
  }
  bitonic_sort(keys, P);
  // |knn ∩ first K of re-ranked GT|; duplicates inside knn count once per GT id, as
  // std::set_intersection on sorted ranges of distinct GT ids does
  unsigned local = 0;
  for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) {
    const uint64_t e = keys[i];
    if (e == ~0ull) continue;
    const uint32_t g = (uint32_t)e;
    bool hit = false;
    for (uint32_t j = 0; j < K; ++j) hit |= knn[(size_t)q * K + j] == g;
    local += hit;
  }
  if (local) atomicAdd(hits, (unsigned long long)local);
}

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

#define BF_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__));         \
      return HS_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

}  // namespace

static int bruteforce_scan_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq,
                                  size_t k, int metric, uint32_t *d_labels, float *d_dists, cudaStream_t stream,
                                  const uint32_t *qmap, const unsigned int *qcount);

// -1: the last call did not take the tcgen05 path (or HS_BF_TC_STATS is unset); else the number
// of queries that went through the fallback scan
static long long g_last_tc_fallback = -1;
long long bruteforce_last_tc_fallback() { return g_last_tc_fallback; }

int bruteforce_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq, size_t k,
                      int metric, uint32_t *d_labels, float *d_dists, cudaStream_t stream) {
  if (nq == 0 || n == 0) return HS_OK;
  if (k == 0 || k > 2048 || n >= (1ull << 32) || nq >= (1ull << 31)) {
    set_error("bruteforce: k must be in [1, 2048], n < 2^32");
    return HS_ERR_ARG;
  }
  g_last_tc_fallback = -1;
  if (bruteforce_tc_applicable(n, dim, nq, k)) {
    // tensor-core filter + exact re-scoring; queries it could not certify come back in a list
    // and go through the scan kernel (same results either way)
    uint32_t *fb_list = nullptr;
    unsigned int *fb_count = nullptr;
    void *scratch = nullptr;
    int rc = bruteforce_tc_device(d_base, n, dim, d_queries, nq, k, metric, d_labels, d_dists, stream, &fb_list,
                                  &fb_count, &scratch);
    if (rc == HS_OK)
      rc = bruteforce_scan_device(d_base, n, dim, d_queries, nq, k, metric, d_labels, d_dists, stream, fb_list,
                                  fb_count);
    if (rc == HS_OK && std::getenv("HS_BF_TC_STATS")) {     // debugging aid: synchronous
      unsigned int c = 0;
      if (cudaMemcpyAsync(&c, fb_count, sizeof c, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
          cudaStreamSynchronize(stream) == cudaSuccess)
        g_last_tc_fallback = (long long)c;
    }
    if (scratch) cudaFreeAsync(scratch, stream);
    return rc;
  }
  return bruteforce_scan_device(d_base, n, dim, d_queries, nq, k, metric, d_labels, d_dists, stream, nullptr, nullptr);
}

static int bruteforce_scan_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq,
                                  size_t k, int metric, uint32_t *d_labels, float *d_dists, cudaStream_t stream,
                                  const uint32_t *qmap, const unsigned int *qcount) {
  int dev = 0, sms = 148;
  BF_CUDA(cudaGetDevice(&dev));
  BF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t q_tiles = (uint32_t)((nq + QT - 1) / QT);
  // enough (query tile, base split) CTAs for ~2 waves, but each split keeps >= max(4k, 4096) rows
  uint32_t splits = (uint32_t)((2 * sms + q_tiles - 1) / q_tiles);
  const size_t min_rows = std::max<size_t>(4096, 4 * k);
  if ((size_t)splits * min_rows > n) splits = (uint32_t)std::max<size_t>(1, n / min_rows);
  while ((size_t)splits * k > 8192 && splits > 1) --splits;
  uint32_t rows_per_split = (uint32_t)((n + splits - 1) / splits);
  rows_per_split = (rows_per_split + RT - 1) / RT * RT;
  splits = (uint32_t)((n + rows_per_split - 1) / rows_per_split);

  uint64_t *partial = nullptr;
  BF_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&partial), (size_t)splits * nq * k * sizeof(uint64_t), stream));
  const size_t smem = (size_t)(SLAB * QPAD + RT * SLAB) * sizeof(float);
  dim3 grid(q_tiles, splits);
  if (metric == HS_METRIC_IP) {
    BF_CUDA(cudaFuncSetAttribute(bf_scan_kernel<HS_METRIC_IP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bf_scan_kernel<HS_METRIC_IP><<<grid, QT, smem, stream>>>(d_base, (uint32_t)n, (uint32_t)dim, d_queries,
                                                             (uint32_t)nq, (uint32_t)k, rows_per_split, partial, qmap, qcount);
  } else {
    BF_CUDA(cudaFuncSetAttribute(bf_scan_kernel<HS_METRIC_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bf_scan_kernel<HS_METRIC_L2><<<grid, QT, smem, stream>>>(d_base, (uint32_t)n, (uint32_t)dim, d_queries,
                                                             (uint32_t)nq, (uint32_t)k, rows_per_split, partial, qmap, qcount);
  }
  BF_CUDA(cudaGetLastError());
  const uint32_t P = next_pow2(std::max<uint32_t>(2, splits * (uint32_t)k));
  BF_CUDA(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(P * 8)));
  merge_keys_kernel<<<(uint32_t)nq, 256, P * 8, stream>>>(partial, splits, (uint32_t)nq, (uint32_t)k, P, d_labels,
                                                          d_dists, qmap, qcount, (uint32_t)nq);
  BF_CUDA(cudaGetLastError());
  BF_CUDA(cudaFreeAsync(partial, stream));
  return HS_OK;
}

int topk_merge_device(const uint32_t *d_labels_in, const float *d_dists_in, size_t n_parts, size_t nq, size_t k,
                      uint32_t *d_labels_out, float *d_dists_out, cudaStream_t stream) {
  if (nq == 0) return HS_OK;
  if (n_parts == 0 || k == 0 || n_parts * k > 16384) {
    set_error("topk_merge: n_parts * k must be in [1, 16384]");
    return HS_ERR_ARG;
  }
  const uint32_t P = next_pow2(std::max<uint32_t>(2, (uint32_t)(n_parts * k)));
  BF_CUDA(cudaFuncSetAttribute(merge_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(P * 8)));
  merge_pairs_kernel<<<(uint32_t)nq, P >= 512 ? 256 : 64, P * 8, stream>>>(
      d_labels_in, d_dists_in, (uint32_t)n_parts, (uint32_t)nq, (uint32_t)k, P, d_labels_out, d_dists_out);
  BF_CUDA(cudaGetLastError());
  return HS_OK;
}

int recall_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq, const uint32_t *d_knn,
                  size_t K, const uint32_t *d_gt, size_t gt_k, int metric, unsigned long long *d_hits,
                  cudaStream_t stream) {
  if (gt_k < K || gt_k > 8192 || K == 0) {
    set_error("recall: need K <= gt_k <= 8192");
    return HS_ERR_ARG;
  }
  BF_CUDA(cudaMemsetAsync(d_hits, 0, sizeof(unsigned long long), stream));
  if (nq == 0) return HS_OK;
  const uint32_t P = next_pow2(std::max<uint32_t>(2, (uint32_t)gt_k));
  BF_CUDA(cudaFuncSetAttribute(recall_kernel<HS_METRIC_IP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(P * 8)));
  BF_CUDA(cudaFuncSetAttribute(recall_kernel<HS_METRIC_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(P * 8)));
  if (metric == HS_METRIC_IP)
    recall_kernel<HS_METRIC_IP><<<(uint32_t)nq, 128, P * 8, stream>>>(d_base, (uint32_t)n, (uint32_t)dim, d_queries,
                                                                      d_knn, (uint32_t)K, d_gt, (uint32_t)gt_k, P, d_hits);
  else
    recall_kernel<HS_METRIC_L2><<<(uint32_t)nq, 128, P * 8, stream>>>(d_base, (uint32_t)n, (uint32_t)dim, d_queries,
                                                                      d_knn, (uint32_t)K, d_gt, (uint32_t)gt_k, P, d_hits);
  BF_CUDA(cudaGetLastError());
  return HS_OK;
}

}  // namespace hs
