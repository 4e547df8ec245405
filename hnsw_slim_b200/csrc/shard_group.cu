// hs_shardgroup — one rank's part of a corpus that is sharded into per-GPU sub-graphs (SURVEY.md §8e;
// the reference builds one graph, so this has no reference analogue beyond "every shard answers the
// query like hnsw_slim_client_update_patch.cc:223-226 would, the k best of all answers win").
//
// A batch costs this rank: one traversal launch per local shard (each with programmatic stream
// serialization, so consecutive launches — of this batch AND of the next one — overlap tail to head),
// whose result rows go straight into slot [shard] of the gather table of EVERY rank (own memory +
// peer memory over NVLink), and one small merge kernel on a second, higher-priority stream.  Nothing
// else: no collective call, no stream memory operation between two traversal launches.
//   * "my rows of batch s are in place everywhere" is signalled by the LAST warp of the batch's
//     launches to finish (scatter_signal_done, traverse_common.cuh): flag word [rank] of every rank;
//   * the merge stream waits (cuStreamWaitValue32) until all flag words of ITS rank reached s,
//     merges table slot s % depth, and acknowledges to every rank (cuStreamWriteValue32);
//   * a traversal warp of batch s checks the acknowledgements of batch s - depth before its first
//     row store (scatter_wait_acks): the slot it overwrites has been consumed everywhere.
// So the search stream never waits for another rank: the step time of a stream of batches is the
// traversal time, and the exchange + merge of batch s ride on the tail of batch s + 1.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "bruteforce.cuh"
#include "hs_index.h"
#include "traverse_common.cuh"

using namespace hs;

#define SG_CUDA(call)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                 \
      return HS_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

struct hs_shardgroup {
  int device = 0, world = 1, rank = 0, depth = 4;
  size_t slots = 0, nq_max = 0, k = 0;          // slots = shards over all ranks
  std::vector<hs_index *> shards;               // local shards (borrowed)
  size_t first_slot = 0;                        // global slot of shards[0]
  uint8_t *base = nullptr;                      // [depth x labels][depth x dists][control block]
  size_t bytes = 0;
  std::vector<uint8_t *> peer;                  // every rank's base in THIS process (peer[rank] == base)
  bool peer_is_ipc = false;
  cudaStream_t s_search = nullptr, s_merge = nullptr;
  unsigned int seq = 0;
  static constexpr int kEvRing = 32;
  cudaEvent_t ev[kEvRing] = {};
  unsigned long long ev_head = 0, ev_tail = 0;
  CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;

  size_t table_elems() const { return slots * nq_max * k; }
  size_t off_labels(unsigned d) const { return (size_t)d * table_elems() * 4; }
  size_t off_dists(unsigned d) const { return ((size_t)depth + d) * table_elems() * 4; }
  size_t off_ctrl() const { return (2 * (size_t)depth * table_elems() * 4 + 255) / 256 * 256; }
  size_t off_flags() const { return off_ctrl(); }              // kMaxScatter words: flags[r] = last batch of rank r
  size_t off_acks() const { return off_ctrl() + 64; }          // kMaxScatter words: acks[r] = last batch rank r merged
  size_t off_done() const { return off_ctrl() + 128; }         // depth words: finished warps of the batch in flight
  size_t off_status() const { return off_ctrl() + 192; }       // 1 word: a traversal warp gave up waiting for a peer
  unsigned long long ack_timeout_ns = 20ull * 1000 * 1000 * 1000;
};

namespace hs {
namespace {

// One warp per query: the slots x k candidate rows of a query -> its k best by (distance, label).
// Labels of different shards are disjoint, so every key is unique.
__global__ void __launch_bounds__(128) merge_slots_kernel(const uint32_t *__restrict__ labs,
                                                          const float *__restrict__ dsts, uint32_t slots,
                                                          uint32_t nq, uint32_t k, uint32_t *__restrict__ out_l,
                                                          float *__restrict__ out_d) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
  const uint32_t total = slots * k;
  for (uint32_t q = warp; q < nq; q += nwarps) {
    uint64_t key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t i = (uint32_t)lane + 32u * j;
      key[j] = NONE;
      if (i < total) {
        const size_t src = ((size_t)(i / k) * nq + q) * k + (i % k);
        const uint32_t lab = labs[src];
        if (lab != 0xFFFFFFFFu) key[j] = make_key(dsts[src], lab);
      }
    }
    uint64_t last = 0, mine_out = NONE;      // every key is > 0 (f2ord(+0.0f) has the top bit set)
    for (uint32_t i = 0; i < k; ++i) {
      uint64_t m = NONE;
      if (last != NONE) {
#pragma unroll
        for (int j = 0; j < 4; ++j) m = (key[j] > last && key[j] < m) ? key[j] : m;
      }
      const int o = warp_argmin_key(m);
      last = o >= 0 ? __shfl_sync(FULL, m, o) : NONE;
      if ((uint32_t)lane == i) mine_out = last;
    }
    if ((uint32_t)lane < k) {
      const bool have = mine_out != NONE;
      out_l[(size_t)q * k + lane] = have ? (uint32_t)mine_out : 0xFFFFFFFFu;
      if (out_d) out_d[(size_t)q * k + lane] = have ? ord2f((uint32_t)(mine_out >> 32)) : __int_as_float(0x7f800000);
    }
  }
}

// device pointer for `p`: a device allocation as is, a pinned + mapped host range through its alias
void *device_view(const void *p, size_t bytes) {
  if (!p) return nullptr;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return const_cast<void *>(p);
  return mapped_alias(p, bytes);
}

int driver_entry(const char *name, void **fn) {
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || !*fn) {
    cudaGetLastError();
    set_error(std::string(name) + " is not available from this driver");
    return HS_ERR_CUDA;
  }
  return HS_OK;
}

}  // namespace
}  // namespace hs

extern "C" {

int hs_shardgroup_create(hs_index *const *shards, size_t n_local, int world, int rank, size_t nq_max, size_t k,
                         int depth, hs_shardgroup **out) {
  if (!out || !shards || n_local == 0 || world < 1 || world > kMaxScatter || rank < 0 || rank >= world ||
      nq_max == 0 || k == 0 || k > 4096 || depth < 2 || depth > 16) {
    set_error("hs_shardgroup_create: bad argument (1 <= world <= " + std::to_string(kMaxScatter) +
              ", 2 <= depth <= 16, at least one local shard)");
    return HS_ERR_ARG;
  }
  *out = nullptr;
  for (size_t i = 0; i < n_local; ++i) {
    if (!shards[i] || shards[i]->device != shards[0]->device) {
      set_error("hs_shardgroup_create: the local shards must be loaded on one device");
      return HS_ERR_ARG;
    }
  }
  try {
    std::unique_ptr<hs_shardgroup> g(new hs_shardgroup);
    g->device = shards[0]->device;
    int rc = select_device(g->device);
    if (rc != HS_OK) return rc;
    g->world = world;
    g->rank = rank;
    g->depth = depth;
    g->slots = n_local * (size_t)world;
    g->first_slot = n_local * (size_t)rank;
    g->nq_max = nq_max;
    g->k = k;
    g->shards.assign(shards, shards + n_local);
    void *fn = nullptr;
    if ((rc = driver_entry("cuStreamWriteValue32", &fn)) != HS_OK) return rc;
    g->write32 = reinterpret_cast<decltype(g->write32)>(fn);
    if ((rc = driver_entry("cuStreamWaitValue32", &fn)) != HS_OK) return rc;
    g->wait32 = reinterpret_cast<decltype(g->wait32)>(fn);
    g->bytes = g->off_ctrl() + 256;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&g->base), g->bytes);
    if (e != cudaSuccess) {
      set_error(std::string("hs_shardgroup_create: cudaMalloc(") + std::to_string(g->bytes) + "): " + cudaGetErrorString(e));
      cudaGetLastError();
      return HS_ERR_NOMEM;
    }
    int lo = 0, hi = 0;
    if (cudaMemset(g->base, 0, g->bytes) != cudaSuccess ||
        cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&g->s_search, cudaStreamNonBlocking, lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&g->s_merge, cudaStreamNonBlocking, hi) != cudaSuccess) {
      set_error(std::string("hs_shardgroup_create: ") + cudaGetErrorString(cudaGetLastError()));
      cudaFree(g->base);
      if (g->s_search) cudaStreamDestroy(g->s_search);
      if (g->s_merge) cudaStreamDestroy(g->s_merge);
      return HS_ERR_CUDA;
    }
    g->peer.assign(world, nullptr);
    g->peer[rank] = g->base;
    for (hs_index *ix : g->shards) ix->overlap = 1;      // the launches of a group always chain programmatically
    *out = g.release();
    return HS_OK;
  } catch (const std::bad_alloc &) {
    set_error("hs_shardgroup_create: out of host memory");
    return HS_ERR_NOMEM;
  }
}

int hs_shardgroup_handle(hs_shardgroup *g, void *handle64) {
  if (!g || !handle64) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  SG_CUDA(cudaSetDevice(g->device));
  SG_CUDA(cudaIpcGetMemHandle(&h, g->base));
  std::memcpy(handle64, &h, 64);
  return HS_OK;
}

int hs_shardgroup_connect(hs_shardgroup *g, const void *handles) {
  if (!g || (!handles && g->world > 1)) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  SG_CUDA(cudaSetDevice(g->device));
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank || g->peer[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const uint8_t *>(handles) + 64 * (size_t)r, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(rank " + std::to_string(r) + "): " + cudaGetErrorString(e));
      cudaGetLastError();
      return HS_ERR_CUDA;
    }
    g->peer[r] = static_cast<uint8_t *>(p);
    g->peer_is_ipc = true;
  }
  return HS_OK;
}

int hs_shardgroup_connect_local(hs_shardgroup *const *groups, size_t world) {
  if (!groups || world == 0 || world > (size_t)kMaxScatter) {
    set_error("hs_shardgroup_connect_local: bad argument");
    return HS_ERR_ARG;
  }
  for (size_t r = 0; r < world; ++r) {
    if (!groups[r] || groups[r]->world != (int)world || groups[r]->rank != (int)r ||
        groups[r]->bytes != groups[0]->bytes) {
      set_error("hs_shardgroup_connect_local: groups[r] must be rank r of `world` equally shaped groups");
      return HS_ERR_ARG;
    }
  }
  for (size_t a = 0; a < world; ++a) {
    SG_CUDA(cudaSetDevice(groups[a]->device));
    for (size_t b = 0; b < world; ++b) {
      if (a == b) continue;
      if (groups[a]->device != groups[b]->device) {
        int can = 0;
        SG_CUDA(cudaDeviceCanAccessPeer(&can, groups[a]->device, groups[b]->device));
        if (!can) {
          set_error("device " + std::to_string(groups[a]->device) + " cannot access device " +
                    std::to_string(groups[b]->device) + " (no peer access)");
          return HS_ERR_UNSUPPORTED;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(groups[b]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
          cudaGetLastError();
          return HS_ERR_CUDA;
        }
        cudaGetLastError();
      }
      groups[a]->peer[b] = groups[b]->base;
    }
  }
  return HS_OK;
}

int hs_shardgroup_submit(hs_shardgroup *g, const float *queries, size_t nq, uint32_t *labels_out, float *dists_out) {
  if (!g || (!queries && nq) || (!labels_out && nq)) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  if (nq > g->nq_max) {
    set_error("hs_shardgroup_submit: nq exceeds the group's nq_max");
    return HS_ERR_ARG;
  }
  if (nq == 0) return HS_OK;
  for (int r = 0; r < g->world; ++r) {
    if (!g->peer[r]) {
      set_error("hs_shardgroup_submit: group not connected (hs_shardgroup_connect / _connect_local)");
      return HS_ERR_ARG;
    }
  }
  SG_CUDA(cudaSetDevice(g->device));
  const size_t dim = g->shards[0]->info.dim, k = g->k;
  const float *dq = static_cast<const float *>(device_view(queries, nq * dim * 4));
  uint32_t *dl = static_cast<uint32_t *>(device_view(labels_out, nq * k * 4));
  float *dd = dists_out ? static_cast<float *>(device_view(dists_out, nq * k * 4)) : nullptr;
  if (!dq || !dl || (dists_out && !dd)) {
    set_error("hs_shardgroup_submit: buffers must be device memory or page-locked + mapped host memory "
              "(cudaHostAlloc / hs_pin_host)");
    return HS_ERR_ARG;
  }
  // launch shapes first: every launch of the batch carries the batch's total warp count
  uint32_t total_warps = 0;
  int rc;
  for (hs_index *ix : g->shards) {
    uint32_t w = 0;
    ScatterDst probe;        // plan-only calls validate like a launch: rows need a destination
    probe.n = 1;
    if ((rc = search_device(ix, dq, nq, k, nullptr, nullptr, nullptr, g->s_search, &probe, &w)) != HS_OK) return rc;
    total_warps += w;
  }
  const unsigned int seq = ++g->seq;
  const unsigned d = seq % (unsigned)g->depth;
  // Back-pressure on the host: batch `seq` reuses the table slot of batch seq - depth, and its warps wait
  // in the kernel until every rank has merged that batch.  This rank's own merge must be complete BEFORE
  // the launch: a grid of waiting warps holds every SM slot, including the ones that merge kernel needs.
  // (Induction over the rank that has merged least shows that some rank can always proceed.)
  if (seq > (unsigned)g->depth) {
    const unsigned long long need = (unsigned long long)seq - (unsigned)g->depth;   // 1-based batch number
    if (need > g->ev_head) {
      SG_CUDA(cudaEventSynchronize(g->ev[(need - 1) % hs_shardgroup::kEvRing]));
      g->ev_head = need;
    }
  }
  for (size_t i = 0; i < g->shards.size(); ++i) {
    ScatterDst sc;
    sc.n = (uint32_t)g->world;
    sc.row0 = (unsigned long long)(g->first_slot + i) * nq;      // tables are [slot][nq][k] with the call's nq
    for (int r = 0; r < g->world; ++r) {
      sc.labels[r] = reinterpret_cast<uint32_t *>(g->peer[r] + g->off_labels(d));
      sc.dists[r] = reinterpret_cast<float *>(g->peer[r] + g->off_dists(d));
      sc.flags[r] = reinterpret_cast<uint32_t *>(g->peer[r] + g->off_flags()) + g->rank;
    }
    sc.n_flags = (uint32_t)g->world;
    sc.done_ctr = reinterpret_cast<unsigned int *>(g->base + g->off_done()) + d;
    sc.done_target = total_warps;
    sc.seq = seq;
    sc.acks = reinterpret_cast<const uint32_t *>(g->base + g->off_acks());
    sc.n_acks = (uint32_t)g->world;
    sc.ack_need = seq > (unsigned)g->depth ? seq - (unsigned)g->depth : 0u;
    sc.ack_timeout_ns = g->ack_timeout_ns;
    sc.status = reinterpret_cast<unsigned int *>(g->base + g->off_status());
    if ((rc = search_device(g->shards[i], dq, nq, k, nullptr, nullptr, nullptr, g->s_search, &sc)) != HS_OK) return rc;
  }
  // merge stream: all ranks' rows of `seq` in place -> merge -> acknowledge -> completion event
  CUstream ms = reinterpret_cast<CUstream>(g->s_merge);
  if (total_warps == 0) {      // only empty shards here: nothing was launched, announce from the stream
    for (int r = 0; r < g->world; ++r) {
      const CUdeviceptr a = reinterpret_cast<CUdeviceptr>(g->peer[r] + g->off_flags()) + 4u * (unsigned)g->rank;
      if (g->write32(reinterpret_cast<CUstream>(g->s_search), a, seq, 0) != CUDA_SUCCESS) {
        set_error("cuStreamWriteValue32 failed");
        return HS_ERR_CUDA;
      }
    }
  }
  for (int r = 0; r < g->world; ++r) {
    const CUdeviceptr a = reinterpret_cast<CUdeviceptr>(g->base + g->off_flags()) + 4u * (unsigned)r;
    const CUresult cr = g->wait32(ms, a, seq, 0 /* CU_STREAM_WAIT_VALUE_GEQ: (int32)(*a - seq) >= 0 */);
    if (cr != CUDA_SUCCESS) {
      set_error("cuStreamWaitValue32 failed (" + std::to_string((int)cr) + ")");
      return HS_ERR_CUDA;
    }
  }
  const uint32_t *tl = reinterpret_cast<const uint32_t *>(g->base + g->off_labels(d));
  const float *td = reinterpret_cast<const float *>(g->base + g->off_dists(d));
  if (g->slots * k <= 128 && k <= 32) {
    const int sm = g->shards[0]->sm_count;
    const uint32_t want = (uint32_t)((nq + 3) / 4);
    merge_slots_kernel<<<std::min<uint32_t>(want, (uint32_t)sm * 2u), 128, 0, g->s_merge>>>(
        tl, td, (uint32_t)g->slots, (uint32_t)nq, (uint32_t)k, dl, dd);
    SG_CUDA(cudaGetLastError());
  } else if ((rc = topk_merge_device(tl, td, g->slots, nq, k, dl, dd, g->s_merge)) != HS_OK) {
    return rc;
  }
  for (int r = 0; r < g->world; ++r) {
    const CUdeviceptr a = reinterpret_cast<CUdeviceptr>(g->peer[r] + g->off_acks()) + 4u * (unsigned)g->rank;
    const CUresult cr = g->write32(ms, a, seq, 0 /* CU_STREAM_WRITE_VALUE_DEFAULT */);
    if (cr != CUDA_SUCCESS) {
      set_error("cuStreamWriteValue32 failed (" + std::to_string((int)cr) + ")");
      return HS_ERR_CUDA;
    }
  }
  if (g->ev_tail - g->ev_head == hs_shardgroup::kEvRing) {
    SG_CUDA(cudaEventSynchronize(g->ev[g->ev_head % hs_shardgroup::kEvRing]));
    g->ev_head++;
  }
  cudaEvent_t &ev = g->ev[g->ev_tail % hs_shardgroup::kEvRing];
  if (!ev) SG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  SG_CUDA(cudaEventRecord(ev, g->s_merge));
  g->ev_tail++;
  return HS_OK;
}

int hs_shardgroup_wait_oldest(hs_shardgroup *g) {
  if (!g) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  if (g->ev_head == g->ev_tail) return HS_OK;
  SG_CUDA(cudaSetDevice(g->device));
  SG_CUDA(cudaEventSynchronize(g->ev[g->ev_head % hs_shardgroup::kEvRing]));
  g->ev_head++;
  return HS_OK;
}

int hs_shardgroup_wait(hs_shardgroup *g) {
  if (!g) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  SG_CUDA(cudaSetDevice(g->device));
  SG_CUDA(cudaStreamSynchronize(g->s_merge));
  SG_CUDA(cudaStreamSynchronize(g->s_search));
  g->ev_head = g->ev_tail;
  unsigned int status = 0;
  SG_CUDA(cudaMemcpy(&status, g->base + g->off_status(), 4, cudaMemcpyDeviceToHost));
  if (status != 0) {
    set_error("hs_shardgroup: a traversal warp gave up waiting for another rank's acknowledgement (peer dead or "
              "not submitting the same batches); results since then are unreliable");
    return HS_ERR_CUDA;
  }
  return HS_OK;
}

int hs_shardgroup_streams(hs_shardgroup *g, void **search_stream, void **merge_stream) {
  if (!g) {
    set_error("null argument");
    return HS_ERR_ARG;
  }
  if (search_stream) *search_stream = g->s_search;
  if (merge_stream) *merge_stream = g->s_merge;
  return HS_OK;
}

void hs_shardgroup_free(hs_shardgroup *g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaStreamSynchronize(g->s_merge);
  cudaStreamSynchronize(g->s_search);
  if (g->peer_is_ipc)
    for (int r = 0; r < g->world; ++r)
      if (r != g->rank && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
  cudaFree(g->base);
  for (auto &e : g->ev)
    if (e) cudaEventDestroy(e);
  if (g->s_search) cudaStreamDestroy(g->s_search);
  if (g->s_merge) cudaStreamDestroy(g->s_merge);
  delete g;
}

}  // extern "C"
