// Delta patches on a device-resident hnsw_slim index.
//
// The reference's server re-prunes its full HNSW after every /updateIndex and answers with the nodes whose CHAL
// record changed (convertFromHNSWWithDiff, slim.h:1110-1424; genPatch, slim.h:1427-1476); the client applies that
// stream to its own index with HierarchicalNSWSlim::patchFromStream (slim.h:2206-2253 vector<vector> rows,
// :2292-2340 rows inline, :2343-2388 unordered_map rows; call sites hnsw_slim_client_update_patch.cc:41,73,179)
// and keeps answering queries.  Stream layout, no padding:
//   size_t cur_element_count            (the count AFTER the patch)
//   size_t changed_old_cnt, changed_new_cnt
//   changed_old_cnt records  [uint32 id][int32 level][uint32 total_nbr]                [uint32 blob_size][blob]
//   changed_new_cnt records  [uint32 id][int32 level][uint32 total_nbr][uint64 label]  [uint32 blob_size][blob]
//                            (+ [float vec[dim]] when the rows travel inline)
//   blob = [uint16 offsets[level]][uint32 ids[total_nbr]]  (slim.h:1096-1106), blob_size = 2*level + 4*total_nbr
// patchFromStream overwrites exactly those fields of the element record and swaps the blob; it does NOT touch
// enterpoint_node_ / maxlevel_ — the client keeps searching from its original entry point — and neither does
// this file.
//
// Here the index lives in HBM as fixed-stride rows, so a patch is
//   level 0   one staged upload of the changed rows + ONE scatter kernel (a warp per row) into adj0 — in place;
//             the same for the vectors and labels of new nodes;
//   levels>0  the upper-level arrays are small (a quarter of the nodes at branching factor 4, a few ids each):
//             the host keeps a mirror of them (levels, slots, rows), re-derives the level-descending slot order
//             when a patch touches a node with level > 0 and re-uploads them.
// An index that is to receive patches is loaded with room for them (hs_load_reserve = loadIndex's max_elements
// argument, slim.h:753-761,784).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <unordered_map>

#include "hs_index.h"

namespace hs {

uint32_t PatchNode::slice(int l, const uint8_t **ids) const {
  *ids = nullptr;
  if (!blob || l < 0 || l > level || total == 0) return 0;
  auto off = [&](int j) {
    uint16_t v;
    std::memcpy(&v, blob + 2 * (size_t)j, 2);
    return (uint32_t)v;
  };
  const uint32_t begin = l == 0 ? 0u : off(l - 1);
  const uint32_t end = l == level ? total : off(l);
  *ids = blob + 2 * (size_t)level + 4 * (size_t)begin;
  return end - begin;
}

int parse_patch(const uint8_t *bytes, size_t size, size_t dim, bool inline_rows, uint64_t n_before, uint64_t capacity,
                PatchSet *out) {
  size_t pos = 0;
  bool ok = true;
  auto get = [&](void *dst, size_t nb) {
    if (!ok || nb > size - pos) {
      ok = false;
      std::memset(dst, 0, nb);
      return;
    }
    std::memcpy(dst, bytes + pos, nb);
    pos += nb;
  };
  auto take = [&](size_t nb) -> const uint8_t * {
    if (!ok || nb > size - pos) {
      ok = false;
      return nullptr;
    }
    const uint8_t *r = bytes + pos;
    pos += nb;
    return r;
  };
  get(&out->n_after, 8);
  get(&out->n_old, 8);
  get(&out->n_new, 8);
  if (!ok) {
    set_error("truncated patch header");
    return HS_ERR_IO;
  }
  if (out->n_after < n_before) {
    set_error("patch: element count " + std::to_string(out->n_after) + " is below the index's " + std::to_string(n_before));
    return HS_ERR_IO;
  }
  if (out->n_after > capacity) {
    set_error("patch grows the index to " + std::to_string(out->n_after) + " elements, beyond the " +
              std::to_string(capacity) + " it was loaded with room for (hs_load_reserve max_elements)");
    return HS_ERR_ARG;
  }
  // the shortest record is 4 + 8 + 4 bytes: a count that cannot fit in the stream is a corrupt header
  const uint64_t max_records = (size - pos) / 16;
  if (out->n_old > max_records || out->n_new > max_records || out->n_old + out->n_new > max_records) {
    set_error("patch: record counts exceed the stream length (corrupt header?)");
    return HS_ERR_IO;
  }
  const uint64_t total_records = out->n_old + out->n_new;
  out->nodes.clear();
  out->nodes.reserve(total_records);
  for (uint64_t i = 0; i < total_records; ++i) {
    PatchNode nd;
    nd.is_new = i >= out->n_old;
    get(&nd.id, 4);
    get(&nd.level, 4);
    get(&nd.total, 4);
    if (nd.is_new) get(&nd.label, 8);
    uint32_t bsz = 0;
    get(&bsz, 4);
    if (!ok) break;
    if (nd.id >= out->n_after) {
      set_error("patch: node id " + std::to_string(nd.id) + " out of range");
      return HS_ERR_IO;
    }
    if (nd.level < 0 || nd.level >= kMaxLevels) {
      set_error("patch: level of node " + std::to_string(nd.id) + " out of range");
      return HS_ERR_IO;
    }
    if (nd.total > (1u << 24) || bsz != 2u * (uint32_t)nd.level + 4u * nd.total) {     // get_neighbor_size, slim.h:652-661
      set_error("patch: neighbour blob size mismatch (node " + std::to_string(nd.id) + ")");
      return HS_ERR_IO;
    }
    if (bsz) nd.blob = take(bsz);
    if (nd.is_new && inline_rows) nd.row = take(4 * dim);
    if (!ok) break;
    if (nd.blob) {
      uint32_t prev = 0;
      for (int l = 0; l <= nd.level; ++l) {
        uint32_t end = nd.total;
        if (l < nd.level) {
          uint16_t v;
          std::memcpy(&v, nd.blob + 2 * (size_t)l, 2);
          end = v;
        }
        if (end < prev || end > nd.total) {
          set_error("patch: corrupt level offsets (node " + std::to_string(nd.id) + ")");
          return HS_ERR_IO;
        }
        prev = end;
      }
      const uint8_t *ids = nd.blob + 2 * (size_t)nd.level;
      for (uint32_t j = 0; j < nd.total; ++j) {
        uint32_t v;
        std::memcpy(&v, ids + 4 * (size_t)j, 4);
        if (v >= out->n_after) {
          set_error("patch: neighbour id out of range (node " + std::to_string(nd.id) + ")");
          return HS_ERR_IO;
        }
      }
    }
    out->nodes.push_back(nd);
  }
  if (!ok) {
    set_error("truncated patch stream");
    return HS_ERR_IO;
  }
  out->consumed = pos;
  return HS_OK;
}

void PatchRows::prepare() {
  sorted.clear();
  if (!rows || !row_labels) return;
  sorted.reserve(n_rows);
  for (size_t i = 0; i < n_rows; ++i) sorted.emplace_back(row_labels[i], i);
  std::sort(sorted.begin(), sorted.end());
}

const uint8_t *PatchRows::find(const PatchNode &nd) const {
  if (nd.row) return nd.row;
  if (!rows) return nullptr;
  if (!row_labels) return nd.label < n_rows ? reinterpret_cast<const uint8_t *>(rows + nd.label * dim) : nullptr;
  auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(nd.label, (size_t)0));
  if (it == sorted.end() || it->first != nd.label) return nullptr;
  return reinterpret_cast<const uint8_t *>(rows + it->second * dim);
}

namespace {

// the record that is in force for each patched id (a later record of the same id supersedes an earlier one,
// as the sequential loop of patchFromStream has it)
std::unordered_map<uint32_t, size_t> last_record(const PatchSet &ps) {
  std::unordered_map<uint32_t, size_t> m;
  m.reserve(ps.nodes.size() * 2);
  for (size_t i = 0; i < ps.nodes.size(); ++i) m[ps.nodes[i].id] = i;
  return m;
}

void copy_ids(uint32_t *dst, uint32_t stride, const uint8_t *ids, uint32_t cnt) {
  for (uint32_t j = 0; j < cnt; ++j) std::memcpy(dst + j, ids + 4 * (size_t)j, 4);
  for (uint32_t j = cnt; j < stride; ++j) dst[j] = kInvalid;
}

// wider rows for a host-only image (hs_debug_patch on a graph flattened with the tight strides of its file)
void restride_host(HostGraph *g, uint32_t stride0, uint32_t stride_up) {
  auto widen = [](std::vector<uint32_t> &a, uint32_t from, uint32_t to) {
    if (from == to || a.empty()) return;
    const size_t rows = a.size() / from;
    std::vector<uint32_t> b(rows * (size_t)to, kInvalid);
    for (size_t r = 0; r < rows; ++r) std::memcpy(&b[r * to], &a[r * from], 4 * (size_t)from);
    a.swap(b);
  };
  widen(g->adj0, g->deg0_stride, stride0);
  for (auto &lv : g->upper_adj) widen(lv, g->upper_stride, stride_up);
  g->deg0_stride = stride0;
  g->upper_stride = stride_up;
}

}  // namespace

int apply_patch_host(HostGraph *g, const PatchSet &ps, const PatchRows &rows, bool *upper_changed) {
  if (upper_changed) *upper_changed = false;
  if (g->kind != HS_KIND_SLIM) {
    set_error("delta patches exist for hnsw_slim indices only (patchFromStream, slim.h:2206-2388)");
    return HS_ERR_UNSUPPORTED;
  }
  const bool big = !g->mirror_only;       // host-only inspection keeps the level-0 rows and vectors
  const size_t n_before = g->n, n_after = ps.n_after;
  // 1. everything is checked before anything is changed
  uint32_t need0 = 0, need_up = 0;
  for (const PatchNode &nd : ps.nodes) {
    const uint8_t *ids;
    for (int l = 0; l <= nd.level; ++l) {
      const uint32_t c = nd.slice(l, &ids);
      if (l == 0) need0 = std::max(need0, c);
      else need_up = std::max(need_up, c);
    }
    if (nd.is_new && !rows.find(nd)) {
      set_error("patch: no vector for new node " + std::to_string(nd.id) + " (label " + std::to_string(nd.label) + ")");
      return HS_ERR_ARG;
    }
  }
  if (need0 > g->deg0_stride || need_up > g->upper_stride) {
    if (!big) {
      // a device-resident index is not re-strided in place; hs_load_reserve sizes rows for maxM0 / maxM ids
      // (the re-prune limits of convertFromHNSW, slim.h:1036-1058), so this is a list beyond those limits
      set_error("patch: a neighbour list is longer than the index's row stride (" + std::to_string(need0) + " / " +
                std::to_string(need_up) + " ids, strides " + std::to_string(g->deg0_stride) + " / " +
                std::to_string(g->upper_stride) + ")");
      return HS_ERR_UNSUPPORTED;
    }
    restride_host(g, std::max(g->deg0_stride, (need0 + 31) / 32 * 32), std::max(g->upper_stride, (need_up + 7) / 8 * 8));
  }
  const auto in_force = last_record(ps);

  // 2. per-node fields
  bool touches_upper = false;
  for (const auto &kv : in_force) {
    const PatchNode &nd = ps.nodes[kv.second];
    const int old_level = nd.id < n_before ? g->levels[nd.id] : 0;
    touches_upper |= old_level > 0 || nd.level > 0;
  }
  g->levels.resize(n_after, 0);
  g->labels.resize(n_after, 0u);
  g->deleted.resize(n_after, 0);
  const std::vector<int32_t> old_slot = g->upper_slot;
  g->upper_slot.resize(n_after, -1);
  if (big) {
    g->adj0.resize(n_after * (size_t)g->deg0_stride, kInvalid);
    if (g->kind == HS_KIND_SLIM) g->vec.resize(n_after * g->dim_padded, 0.f);
  }
  for (const auto &kv : in_force) {
    const PatchNode &nd = ps.nodes[kv.second];
    g->levels[nd.id] = (int8_t)nd.level;
    if (nd.is_new) g->labels[nd.id] = (uint32_t)nd.label;       // slim.h:2129 truncation, as the loader
    if (big) {
      const uint8_t *ids;
      const uint32_t c = nd.slice(0, &ids);
      copy_ids(&g->adj0[nd.id * (size_t)g->deg0_stride], g->deg0_stride, ids, c);
      if (nd.is_new) {
        float *dst = &g->vec[nd.id * g->dim_padded];
        std::memcpy(dst, rows.find(nd), 4 * g->dim);
        for (size_t j = g->dim; j < g->dim_padded; ++j) dst[j] = 0.f;
      }
    }
  }
  g->n = n_after;

  // 3. upper levels: slots are the nodes with level > 0 sorted by level descending (ties by id), so the rows of
  //    level l are the dense prefix [0, level_count[l]) — the loader's order, hence a patched image equals the
  //    image of the patched file.  Rows of untouched nodes move with their node.
  if (touches_upper || n_after != n_before) {
    int top = 0;
    for (size_t i = 0; i < n_after; ++i) top = std::max(top, (int)g->levels[i]);
    top = std::max(top, (int)g->maxlevel);
    std::vector<uint32_t> level_count(top + 2, 0);
    std::vector<uint32_t> upper_nodes;
    for (size_t i = 0; i < n_after; ++i) {
      for (int l = 0; l <= g->levels[i]; ++l) level_count[l]++;
      if (g->levels[i] > 0) upper_nodes.push_back((uint32_t)i);
    }
    if (touches_upper) {
      std::stable_sort(upper_nodes.begin(), upper_nodes.end(),
                       [&](uint32_t a, uint32_t b) { return g->levels[a] > g->levels[b]; });
      std::vector<std::vector<uint32_t>> old_adj = std::move(g->upper_adj);
      const std::vector<uint32_t> old_count = g->level_count;
      g->upper_adj.assign(top + 1, {});
      for (int l = 1; l <= top; ++l) g->upper_adj[l].assign((size_t)level_count[l] * g->upper_stride, kInvalid);
      std::fill(g->upper_slot.begin(), g->upper_slot.end(), -1);
      for (uint32_t s = 0; s < upper_nodes.size(); ++s) {
        const uint32_t i = upper_nodes[s];
        g->upper_slot[i] = (int32_t)s;
        const auto it = in_force.find(i);
        for (int l = 1; l <= g->levels[i]; ++l) {
          uint32_t *dst = &g->upper_adj[l][(size_t)s * g->upper_stride];
          if (it != in_force.end()) {
            const uint8_t *ids;
            const uint32_t c = ps.nodes[it->second].slice(l, &ids);
            copy_ids(dst, g->upper_stride, ids, c);
          } else {
            const int32_t os = old_slot[i];      // untouched: it existed before with the same level
            if (os >= 0 && (size_t)l < old_adj.size() && (size_t)l < old_count.size() && (uint32_t)os < old_count[l])
              std::memcpy(dst, &old_adj[l][(size_t)os * g->upper_stride], 4 * (size_t)g->upper_stride);
          }
        }
      }
      g->n_upper = (uint32_t)upper_nodes.size();
      if (upper_changed) *upper_changed = true;
    }
    g->level_count = level_count;
  }
  if (big) {
    uint64_t sum = 0;
    uint32_t mx = 0;
    for (size_t i = 0; i < n_after; ++i) {
      const uint32_t *row = &g->adj0[i * (size_t)g->deg0_stride];
      uint32_t c = 0;
      while (c < g->deg0_stride && row[c] != kInvalid) ++c;
      sum += c;
      mx = std::max(mx, c);
    }
    g->sum_deg0 = sum;
    g->max_deg0 = mx;
  }
  return HS_OK;
}

namespace {

// dst[ids[r] * stride + j] = src[r * stride + j]: one warp per row, 16 bytes per lane and step
// (strides are multiples of 32 words, both buffers 128-byte aligned)
__global__ void scatter_rows_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, const uint32_t *__restrict__ ids,
                                    uint32_t count, uint32_t stride_vec4) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp; r < count; r += n_warps) {
    const size_t to = (size_t)ids[r] * stride_vec4, from = (size_t)r * stride_vec4;
    for (uint32_t j = lane; j < stride_vec4; j += 32) dst[to + j] = src[from + j];
  }
}

__global__ void scatter_words_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src,
                                     const uint32_t *__restrict__ ids, uint32_t count) {
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x) dst[ids[r]] = src[r];
}

// level-0 degree statistics of the rows [0, n): out[0] += sum, out[1] = max
__global__ void degree_stats_kernel(const uint32_t *__restrict__ adj0, uint32_t n, uint32_t stride, unsigned long long *out) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long sum = 0, mx = 0;
  for (uint32_t r = warp; r < n; r += n_warps) {
    uint32_t c = 0;
    for (uint32_t j = lane; j < stride; j += 32) c += adj0[(size_t)r * stride + j] != kInvalid;
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    sum += c;
    mx = mx > c ? mx : c;
  }
  if (lane == 0) {
    atomicAdd(out + 0, sum);
    atomicMax(out + 1, mx);
  }
}

struct DeviceBuf {
  void *p = nullptr;
  ~DeviceBuf() { cudaFree(p); }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) {
      set_error(std::string("cudaMalloc (patch staging): ") + cudaGetErrorString(e));
      return HS_ERR_NOMEM;
    }
    return HS_OK;
  }
};

#define HS_CUDA_P(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__));        \
      return HS_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

}  // namespace

// The device side of patchFromStream.  Synchronous; the caller (hs_patch_apply) holds the handle's mutex and has
// drained the handle's stream.  Launches on other streams must have been quiesced by the caller, as the reference
// requires of its own unsynchronised patchFromStream.
int apply_patch_device(hs_index *ix, const PatchSet &ps, const PatchRows &rows, hs_patch_info *info) {
  HostGraph *m = ix->mirror.get();
  if (!m) {
    set_error("hs_patch_apply: the index was not loaded with hs_load_reserve");
    return HS_ERR_UNSUPPORTED;
  }
  const size_t n_before = m->n;
  bool upper_changed = false;
  int rc = apply_patch_host(m, ps, rows, &upper_changed);      // validates everything before changing anything
  if (rc != HS_OK) return rc;
  const auto in_force = last_record(ps);
  const size_t cnt = in_force.size();
  const uint32_t stride0 = m->deg0_stride;
  const size_t dp = m->dim_padded, dim = m->dim;
  cudaStream_t st = ix->stream;
  uint64_t rows_written = 0;

  // level-0 rows, and the vectors / labels of the new nodes: staged in slabs of at most ~64 MB
  std::vector<uint32_t> order;
  order.reserve(cnt);
  for (const auto &kv : in_force) order.push_back((uint32_t)kv.second);
  std::sort(order.begin(), order.end());            // stream order: deterministic, and new nodes come out id-ascending
  const size_t slab_rows = std::max<size_t>(1, (64u << 20) / (std::max<size_t>(stride0, dp) * 4));
  DeviceBuf d_ids, d_rows;
  if ((rc = d_ids.alloc(std::min(cnt, slab_rows) * 4)) != HS_OK) return rc;
  if ((rc = d_rows.alloc(std::min(cnt, slab_rows) * std::max<size_t>(stride0, dp) * 4)) != HS_OK) return rc;
  std::vector<uint32_t> h_ids, h_words;
  std::vector<float> h_vec;
  for (size_t at = 0; at < cnt; at += slab_rows) {
    const size_t c = std::min(slab_rows, cnt - at);
    h_ids.resize(c);
    h_words.assign(c * stride0, kInvalid);
    for (size_t r = 0; r < c; ++r) {
      const PatchNode &nd = ps.nodes[order[at + r]];
      h_ids[r] = nd.id;
      const uint8_t *ids;
      const uint32_t deg = nd.slice(0, &ids);
      copy_ids(&h_words[r * stride0], stride0, ids, deg);
    }
    HS_CUDA_P(cudaMemcpyAsync(d_ids.p, h_ids.data(), c * 4, cudaMemcpyHostToDevice, st));
    HS_CUDA_P(cudaMemcpyAsync(d_rows.p, h_words.data(), c * stride0 * 4, cudaMemcpyHostToDevice, st));
    scatter_rows_kernel<<<(unsigned)std::min<size_t>((c + 7) / 8, 148 * 8), 256, 0, st>>>(
        reinterpret_cast<uint4 *>(ix->d_adj0), static_cast<const uint4 *>(d_rows.p), static_cast<const uint32_t *>(d_ids.p),
        (uint32_t)c, stride0 / 4);
    HS_CUDA_P(cudaGetLastError());
    HS_CUDA_P(cudaStreamSynchronize(st));          // the host staging vectors are reused by the next slab
    rows_written += c;
    // new nodes of this slab
    h_ids.clear();
    h_vec.clear();
    std::vector<uint32_t> h_lab;
    for (size_t r = 0; r < c; ++r) {
      const PatchNode &nd = ps.nodes[order[at + r]];
      if (!nd.is_new) continue;
      h_ids.push_back(nd.id);
      h_lab.push_back((uint32_t)nd.label);
      const size_t o = h_vec.size();
      h_vec.resize(o + dp, 0.f);
      std::memcpy(&h_vec[o], rows.find(nd), 4 * dim);
    }
    if (!h_ids.empty()) {
      const size_t cn = h_ids.size();
      HS_CUDA_P(cudaMemcpyAsync(d_ids.p, h_ids.data(), cn * 4, cudaMemcpyHostToDevice, st));
      HS_CUDA_P(cudaMemcpyAsync(d_rows.p, h_vec.data(), cn * dp * 4, cudaMemcpyHostToDevice, st));
      scatter_rows_kernel<<<(unsigned)std::min<size_t>((cn + 7) / 8, 148 * 8), 256, 0, st>>>(
          reinterpret_cast<uint4 *>(ix->d_vec), static_cast<const uint4 *>(d_rows.p), static_cast<const uint32_t *>(d_ids.p),
          (uint32_t)cn, (uint32_t)(dp / 4));
      HS_CUDA_P(cudaGetLastError());
      HS_CUDA_P(cudaStreamSynchronize(st));
      HS_CUDA_P(cudaMemcpyAsync(d_rows.p, h_lab.data(), cn * 4, cudaMemcpyHostToDevice, st));
      scatter_words_kernel<<<(unsigned)((cn + 255) / 256), 256, 0, st>>>(ix->d_labels, static_cast<const uint32_t *>(d_rows.p),
                                                                        static_cast<const uint32_t *>(d_ids.p), (uint32_t)cn);
      HS_CUDA_P(cudaGetLastError());
      HS_CUDA_P(cudaStreamSynchronize(st));
    }
  }

  // upper levels from the mirror
  if (upper_changed) {
    HS_CUDA_P(cudaMemcpyAsync(ix->d_upper_slot, m->upper_slot.data(), m->n * 4, cudaMemcpyHostToDevice, st));
    for (int l = 1; l < (int)m->upper_adj.size() && l < kMaxLevels; ++l) {
      const size_t words = m->upper_adj[l].size();
      if (words > ix->cap_upper_words[l]) {
        const size_t cap = words + words / 4 + 1024;        // head-room: the next patches grow it again
        uint32_t *fresh = nullptr;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&fresh), cap * 4);
        if (e != cudaSuccess) {
          set_error(std::string("cudaMalloc (upper-level rows): ") + cudaGetErrorString(e));
          return HS_ERR_NOMEM;
        }
        cudaFree(ix->d_upper_adj[l]);          // nothing of this handle is in flight (see above)
        ix->info.device_bytes += (cap - ix->cap_upper_words[l]) * 4;
        ix->d_upper_adj[l] = fresh;
        ix->cap_upper_words[l] = cap;
      }
      if (words) HS_CUDA_P(cudaMemcpyAsync(ix->d_upper_adj[l], m->upper_adj[l].data(), words * 4, cudaMemcpyHostToDevice, st));
    }
  }
  // level-0 degree statistics (hs_index_info.sum_deg0 / max_deg0 feed the algorithmic-bytes figure)
  DeviceBuf d_stat;
  if ((rc = d_stat.alloc(16)) != HS_OK) return rc;
  HS_CUDA_P(cudaMemsetAsync(d_stat.p, 0, 16, st));
  if (m->n) {
    degree_stats_kernel<<<148 * 4, 256, 0, st>>>(ix->d_adj0, (uint32_t)m->n, stride0, static_cast<unsigned long long *>(d_stat.p));
    HS_CUDA_P(cudaGetLastError());
  }
  unsigned long long h_stat[2] = {0, 0};
  HS_CUDA_P(cudaMemcpyAsync(h_stat, d_stat.p, 16, cudaMemcpyDeviceToHost, st));
  HS_CUDA_P(cudaStreamSynchronize(st));

  m->sum_deg0 = h_stat[0];
  m->max_deg0 = (uint32_t)h_stat[1];
  ix->info.n = m->n;
  ix->info.n_upper = m->n_upper;
  ix->info.sum_deg0 = m->sum_deg0;
  ix->info.max_deg0 = m->max_deg0;
  for (int l = 0; l < kMaxLevels; ++l) ix->level_count[l] = l < (int)m->level_count.size() ? m->level_count[l] : 0u;
  ix->plan_ok = false;          // cached launch plans carry n and the array pointers
  ix->planq_ok = false;
  if (info) {
    info->n_before = n_before;
    info->n_after = m->n;
    info->changed_old = ps.n_old;
    info->changed_new = ps.n_new;
    info->bytes_consumed = ps.consumed;
    info->rows_written = rows_written;
    info->upper_rebuilt = upper_changed ? 1 : 0;
  }
  return HS_OK;
}

}  // namespace hs
