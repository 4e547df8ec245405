// Host side of hs_load: parse the reference's .graph image and flatten it.
//
// File format (written by HierarchicalNSWSlim::saveIndex, slim.h:717-751; read by
// loadIndex, slim.h:753-815), no padding between fields:
//   size_t cur_element_count, size_data_per_element, label_offset, offsetTotalNeighbor,
//          offsetData, offsetNeighbor; int maxlevel, threshold_level; uint32 enterpoint;
//   size_t maxM, maxM0, M, ef_construction; bool has_deleted;
//   [hnsw_slimq only: RaBitQ metadata, centroids, rotator — slimq.h:1184-1203]
//   cur_element_count records of size_data_per_element bytes:
//       [int32 level @0][uint32 total_nbr @4][uint64 label @8][stale pointer @16][payload @24]
//   per node: uint32 blobSize, then blobSize bytes iff blobSize != 0 && total_nbr != 0:
//       [uint16 offsets[level]][uint32 ids[total_nbr]]           (slim.h:1096-1106)
//   level-l slice of a node = ids[(l ? offsets[l-1] : 0) .. (l == level ? total : offsets[l]))
//
// Flattened form (DESIGN.md "HBM layout"): fixed-stride, 128-byte-aligned rows so one
// hop costs one dependent load instead of the reference's record -> blob pointer chase.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>
#include <random>

#include "hs_internal.h"

namespace hs {

namespace {

struct Reader {
  const uint8_t *p;
  size_t size, pos = 0;
  bool ok = true;
  template <typename T> T get() {
    T v{};
    if (pos + sizeof(T) > size) {
      ok = false;
      return v;
    }
    std::memcpy(&v, p + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
  const uint8_t *take(size_t nbytes) {
    if (pos + nbytes > size || pos + nbytes < pos) {
      ok = false;
      return nullptr;
    }
    const uint8_t *r = p + pos;
    pos += nbytes;
    return r;
  }
};

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
// the blobs sit at arbitrary byte offsets of the file image: no aligned access
inline uint32_t load_u16(const uint8_t *p, size_t idx) {
  uint16_t v;
  std::memcpy(&v, p + 2 * idx, 2);
  return v;
}

}  // namespace

int parse_hnsw_graph(const uint8_t *bytes, size_t size, size_t dim, HostGraph *g);

// The query-quantiser constant of hnsw_slimq.  The reference computes it at load time
// (slimq.h:1274-1276 -> faster_config, rq/quantization/rabitq.hpp:27-34 ->
// get_const_scaling_factors, rabitq_impl.hpp:363-377): the mean, over 100 random unit vectors,
// of the rescale factor t that maximises <o, quantised(t*o)> / |quantised(t*o)|
// (best_rescale_factor, rabitq_impl.hpp:275-333).  The reference seeds the vectors from
// std::random_device (a different constant on every load); here the seed is fixed so that an
// index answers the same way every time, and hs_set_query_tconst can override the value.
double slimq_default_tconst(size_t padded_dim, size_t ex_bits) {
  static const double kTightStart[9] = {0, 0.15, 0.20, 0.52, 0.59, 0.71, 0.75, 0.77, 0.81};
  constexpr int kConstNum = 100, kNEnum = 10;
  constexpr double kEps = 1e-5;
  std::mt19937_64 gen(0x5eedULL * 1000003ULL + padded_dim * 31ULL + ex_bits);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> o(padded_dim);
  std::vector<int> cur(padded_dim);
  const int top = (1 << ex_bits) - 1;
  double sum = 0;
  for (int v = 0; v < kConstNum; ++v) {
    double nrm = 0;
    for (auto &x : o) {
      x = nd(gen);
      nrm += x * x;
    }
    nrm = std::sqrt(nrm);
    double max_o = 0;
    for (auto &x : o) {
      x = std::fabs(x / nrm);
      max_o = std::max(max_o, x);
    }
    const double t_end = (double)(top + kNEnum) / max_o;
    const double t_start = t_end * kTightStart[ex_bits];
    double sqr_den = (double)padded_dim * 0.25, num = 0;
    using Ev = std::pair<double, size_t>;
    std::priority_queue<Ev, std::vector<Ev>, std::greater<>> next_t;
    for (size_t i = 0; i < padded_dim; ++i) {
      cur[i] = (int)(t_start * o[i] + kEps);
      sqr_den += (double)cur[i] * cur[i] + cur[i];
      num += (cur[i] + 0.5) * o[i];
      next_t.emplace((double)(cur[i] + 1) / o[i], i);
    }
    double max_ip = 0, t = 0;
    while (!next_t.empty()) {
      const double cur_t = next_t.top().first;
      const size_t id = next_t.top().second;
      next_t.pop();
      cur[id]++;
      sqr_den += 2 * cur[id];
      num += o[id];
      const double ip = num / std::sqrt(sqr_den);
      if (ip > max_ip) {
        max_ip = ip;
        t = cur_t;
      }
      if (cur[id] < top) {
        const double t_next = (double)(cur[id] + 1) / o[id];
        if (t_next < t_end) next_t.emplace(t_next, id);
      }
    }
    sum += t;
  }
  return sum / kConstNum;
}

int read_file(const char *path, std::vector<uint8_t> *out) {
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) {
    set_error(std::string("Cannot open file ") + path);   // slim.h:757-758
    return HS_ERR_IO;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) {
    ::close(fd);
    set_error(std::string("Cannot stat file ") + path);
    return HS_ERR_IO;
  }
  out->resize((size_t)st.st_size);
  size_t got = 0;
  while (got < out->size()) {
    ssize_t r = ::read(fd, out->data() + got, std::min<size_t>(out->size() - got, 1u << 30));
    if (r <= 0) break;
    got += (size_t)r;
  }
  ::close(fd);
  if (got != out->size()) {
    set_error(std::string("short read on ") + path);
    return HS_ERR_IO;
  }
  return HS_OK;
}

// Shared tail of the parsers: from "neighbour list of node i on level l" to the fixed-stride
// rows of DESIGN.md "HBM layout".  list(i, l, &ids) returns the length and points ids at the
// (unaligned) uint32 ids; levels[] is filled.
template <typename ListFn>
int flatten_lists(HostGraph *g, ListFn &&list) {
  const size_t n = g->n;
  uint32_t max_deg0 = 0, max_deg_up = 0;
  uint64_t sum_deg0 = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t *ids = nullptr;
    for (int l = 0; l <= g->levels[i]; ++l) {
      const uint32_t deg = list(i, l, &ids);
      if (l == 0) {
        max_deg0 = std::max(max_deg0, deg);
        sum_deg0 += deg;
      } else {
        max_deg_up = std::max(max_deg_up, deg);
      }
    }
  }
  g->max_deg0 = max_deg0;
  g->max_deg_upper = max_deg_up;
  g->sum_deg0 = sum_deg0;
  g->deg0_stride = std::max<uint32_t>(32, round_up(max_deg0, 32));
  g->upper_stride = std::max<uint32_t>(8, round_up(max_deg_up, 8));
  if (g->reserve_strides) {      // room for the longest lists later patches may bring (re-prune limits, slim.h:1036-1058)
    g->deg0_stride = std::max(g->deg0_stride, round_up((uint32_t)g->maxM0, 32));
    g->upper_stride = std::max(g->upper_stride, round_up((uint32_t)g->maxM, 8));
  }

  // upper-level slots: nodes sorted by level descending, so the rows of level l are the dense
  // slot prefix [0, level_count[l])
  // (a client index that received delta patches may hold nodes above the header's maxlevel: patchFromStream,
  // slim.h:2206-2388, leaves maxlevel_ alone — the search never looks at those levels, the rows are kept anyway)
  int top = g->maxlevel;
  for (size_t i = 0; i < n; ++i) top = std::max(top, (int)g->levels[i]);
  g->level_count.assign(top + 2, 0);
  std::vector<uint32_t> upper_nodes;
  for (size_t i = 0; i < n; ++i) {
    for (int l = 0; l <= g->levels[i]; ++l) g->level_count[l]++;
    if (g->levels[i] > 0) upper_nodes.push_back((uint32_t)i);
  }
  std::stable_sort(upper_nodes.begin(), upper_nodes.end(),
                   [&](uint32_t a, uint32_t b) { return g->levels[a] > g->levels[b]; });
  g->n_upper = (uint32_t)upper_nodes.size();
  g->upper_slot.assign(n, -1);
  for (uint32_t s = 0; s < upper_nodes.size(); ++s) g->upper_slot[upper_nodes[s]] = (int32_t)s;

  g->adj0.assign(n * (size_t)g->deg0_stride, kInvalid);
  g->upper_adj.assign(top + 1, {});
  for (int l = 1; l <= top; ++l)
    g->upper_adj[l].assign((size_t)g->level_count[l] * g->upper_stride, kInvalid);
  for (size_t i = 0; i < n; ++i) {
    for (int l = 0; l <= g->levels[i]; ++l) {
      const uint8_t *ids = nullptr;
      const uint32_t deg = list(i, l, &ids);
      uint32_t *dst = (l == 0) ? &g->adj0[i * (size_t)g->deg0_stride]
                               : &g->upper_adj[l][(size_t)g->upper_slot[i] * g->upper_stride];
      for (uint32_t j = 0; j < deg; ++j) {
        uint32_t id;
        std::memcpy(&id, ids + 4 * (size_t)j, 4);
        if (id >= n) {
          set_error("neighbour id out of range in .graph (node " + std::to_string(i) + ")");
          return HS_ERR_IO;
        }
        dst[j] = id;
      }
    }
  }
  return HS_OK;
}

int parse_graph(const uint8_t *bytes, size_t size, int kind, size_t dim, HostGraph *g) {
  if (kind == HS_KIND_HNSW) return parse_hnsw_graph(bytes, size, dim, g);
  Reader r{bytes, size};
  g->kind = kind;
  g->dim = dim;
  g->n = r.get<uint64_t>();
  g->size_data_per_element = r.get<uint64_t>();
  g->label_offset = r.get<uint64_t>();
  g->offset_total = r.get<uint64_t>();
  g->offset_data = r.get<uint64_t>();
  g->offset_nbr = r.get<uint64_t>();
  g->maxlevel = r.get<int32_t>();
  g->threshold_level = r.get<int32_t>();
  g->enterpoint = r.get<uint32_t>();
  g->maxM = r.get<uint64_t>();
  g->maxM0 = r.get<uint64_t>();
  g->M = r.get<uint64_t>();
  g->ef_construction = r.get<uint64_t>();
  g->has_deleted = r.get<uint8_t>() != 0;
  if (!r.ok) {
    set_error("truncated .graph header");
    return HS_ERR_IO;
  }
  // hnsw_slimq stores offsetData_ = offset_bin_data_ = 28 (slimq.h:1503), hnsw_slim 24 (slim.h:130)
  const uint64_t want_offset_data = kind == HS_KIND_SLIMQ ? 28 : 24;
  if (g->offset_total != 4 || g->label_offset != 8 || g->offset_nbr != 16 || g->offset_data != want_offset_data) {
    set_error("unexpected record offsets in .graph header (not a HNSW-Slim index?)");
    return HS_ERR_IO;
  }
  if (g->n >= (1ull << 31)) {
    set_error("index has >= 2^31 nodes");
    return HS_ERR_UNSUPPORTED;
  }
  if (g->n > 0 && (g->maxlevel < 0 || g->maxlevel >= kMaxLevels || g->enterpoint >= g->n)) {
    set_error("bad maxlevel / enterpoint in .graph header");
    return HS_ERR_IO;
  }
  if (g->n == 0) g->maxlevel = std::max(0, std::min(g->maxlevel, kMaxLevels - 1));   // empty index: the header value is -1
  if (g->maxM > 65535 || g->maxM0 > 65535) {
    set_error("maxM / maxM0 above 65535 in .graph header");
    return HS_ERR_IO;
  }

  size_t payload_off = 24;
  if (kind == HS_KIND_SLIM) {
    if (g->size_data_per_element != 24 + 4 * dim) {
      set_error("size_data_per_element " + std::to_string(g->size_data_per_element) +
                " does not match dim " + std::to_string(dim) + " (expected 24 + 4*dim)");
      return HS_ERR_IO;
    }
  } else if (kind == HS_KIND_SLIMQ) {
    // slimq.h:1184-1203
    g->num_cluster = r.get<uint64_t>();
    uint64_t qdim = r.get<uint64_t>();
    g->padded_dim_q = r.get<uint64_t>();
    uint64_t off_cluster = r.get<uint64_t>(), off_bin = r.get<uint64_t>(), off_ex = r.get<uint64_t>();
    uint64_t size_bin = r.get<uint64_t>(), size_ex = r.get<uint64_t>();
    g->ex_bits = r.get<uint64_t>();
    g->metric_type_q = r.get<uint8_t>();
    // slimq.h:1498-1505: size_data_per_element = offset_ex_data + size_ex_data, offset_ex_data = 28 + size_bin_data
    if (!r.ok || qdim != dim || g->padded_dim_q % 64 != 0 || g->padded_dim_q < dim || g->padded_dim_q > (1u << 20) ||
        off_cluster != 24 || off_bin != 28 || size_bin != g->padded_dim_q / 8 + 12 || off_ex != 28 + size_bin ||
        size_ex > (1ull << 32) || g->size_data_per_element != 28 + size_bin + size_ex) {
      set_error("inconsistent RaBitQ metadata in hnsw_slimq .graph header");
      return HS_ERR_IO;
    }
    if (g->num_cluster == 0 || g->num_cluster > (1u << 20)) {
      set_error("cluster count out of range in hnsw_slimq .graph header");
      return HS_ERR_IO;
    }
    const uint8_t *c = r.take(g->num_cluster * g->padded_dim_q * sizeof(float));
    const uint8_t *f = r.take(4 * g->padded_dim_q / 8);   // FhtKacRotator::save, rotator.hpp:263-275
    if (!r.ok) {
      set_error("truncated centroids / rotator in hnsw_slimq .graph");
      return HS_ERR_IO;
    }
    g->centroids.resize(g->num_cluster * g->padded_dim_q);
    std::memcpy(g->centroids.data(), c, g->centroids.size() * sizeof(float));
    g->rotator_flip.assign(f, f + 4 * g->padded_dim_q / 8);
  } else {
    set_error("unknown index kind");
    return HS_ERR_ARG;
  }

  const size_t n = g->n, rec = g->size_data_per_element;
  size_t elements_bytes = 0;
  if (__builtin_mul_overflow(n, rec, &elements_bytes)) {
    set_error("element count x record size overflows in .graph header");
    return HS_ERR_IO;
  }
  const uint8_t *elements = r.take(elements_bytes);
  if (!r.ok) {
    set_error("truncated element records in .graph");
    return HS_ERR_IO;
  }

  g->dim_padded = (dim + kRowAlignFloats - 1) / kRowAlignFloats * kRowAlignFloats;
  g->labels.resize(n);
  g->levels.resize(n);
  g->deleted.assign(n, 0);
  std::vector<uint32_t> total(n);
  if (kind == HS_KIND_SLIM) g->vec.assign(n * g->dim_padded, 0.f);
  if (kind == HS_KIND_SLIMQ) {
    const size_t words = g->padded_dim_q / 64;
    g->cluster_id.resize(n);
    g->bin_code.resize(n * words);
    g->f_add.resize(n);
    g->f_rescale.resize(n);
    g->f_error.resize(n);
  }
  for (size_t i = 0; i < n; ++i) {
    const uint8_t *e = elements + i * rec;
    int32_t lvl;
    uint64_t label;
    std::memcpy(&lvl, e, 4);
    std::memcpy(&total[i], e + 4, 4);
    std::memcpy(&label, e + 8, 8);
    if (lvl < 0 || lvl >= kMaxLevels) {
      set_error("node level out of range in .graph");
      return HS_ERR_IO;
    }
    g->levels[i] = (int8_t)lvl;
    g->labels[i] = (uint32_t)label;     // slim.h:2129 result[i] = (tableint) label
    g->deleted[i] = e[6] & 1;           // slim.h:1776-1781
    if (kind == HS_KIND_SLIM) {
      std::memcpy(&g->vec[i * g->dim_padded], e + payload_off, 4 * dim);
    } else {
      // [uint32 cluster @24][bin @28: uint64 code[pd/64], float f_add, f_rescale, f_error]
      const size_t words = g->padded_dim_q / 64;
      std::memcpy(&g->cluster_id[i], e + 24, 4);
      if (g->cluster_id[i] >= g->num_cluster) {       // the kernel indexes its centroid-distance table with it
        set_error("cluster id out of range in hnsw_slimq .graph (node " + std::to_string(i) + ")");
        return HS_ERR_IO;
      }
      std::memcpy(&g->bin_code[i * words], e + 28, 8 * words);
      std::memcpy(&g->f_add[i], e + 28 + 8 * words, 4);
      std::memcpy(&g->f_rescale[i], e + 28 + 8 * words + 4, 4);
      std::memcpy(&g->f_error[i], e + 28 + 8 * words + 8, 4);
    }
  }

  // ---- the blobs: one per node, level slices inside ----
  struct BlobRef {
    const uint8_t *p;
    uint32_t size;
  };
  std::vector<BlobRef> blobs(n, BlobRef{nullptr, 0});
  for (size_t i = 0; i < n; ++i) {
    uint32_t bsz = r.get<uint32_t>();
    if (!r.ok) {
      set_error("truncated neighbour blobs in .graph");
      return HS_ERR_IO;
    }
    if (bsz == 0 || total[i] == 0) continue;          // slim.h:741-748 / :796-808
    const uint8_t *b = r.take(bsz);
    const int lvl = g->levels[i];
    if (!r.ok || bsz != 2u * lvl + 4u * total[i]) {
      set_error("neighbour blob size mismatch in .graph (node " + std::to_string(i) + ")");
      return HS_ERR_IO;
    }
    blobs[i] = BlobRef{b, bsz};
    uint32_t prev = 0;
    for (int l = 0; l <= lvl; ++l) {
      uint32_t end = (l == lvl) ? total[i] : load_u16(b, l);
      if (end < prev || end > total[i]) {
        set_error("corrupt level offsets in .graph (node " + std::to_string(i) + ")");
        return HS_ERR_IO;
      }
      prev = end;
    }
  }
  return flatten_lists(g, [&](size_t i, int l, const uint8_t **ids) -> uint32_t {
    if (!blobs[i].p) return 0;
    const int lvl = g->levels[i];
    const uint32_t begin = l == 0 ? 0 : load_u16(blobs[i].p, l - 1);
    const uint32_t end = l == lvl ? total[i] : load_u16(blobs[i].p, l);
    *ids = blobs[i].p + 2 * (size_t)lvl + 4 * (size_t)begin;
    return end - begin;
  });
}

// Upstream-format HNSW index, as this fork's HierarchicalNSW::saveIndex writes it
// (hnsw.h:748-779; read by loadIndex, hnsw.h:781-893) — the `--solve_strategy=hnsw` baseline of
// the reference (hnsw_strategy.h:15-61):
//   size_t offsetLevel0, max_elements, cur_element_count, size_data_per_element, label_offset,
//          offsetData; int maxlevel; uint32 enterpoint; size_t maxM, maxM0, M; double mult;
//   size_t ef_construction;
//   cur_element_count level-0 records of size_data_per_element bytes (hnsw.h:116-121):
//       [uint32 header: low uint16 = list length, byte 2 bit 0 = deleted (hnsw.h:1007-1018)]
//       [uint32 ids[maxM0]] [float vec[dim] @offsetData] [uint64 label @label_offset]
//   per node: uint32 linkListSize, then linkListSize = level * (4 + 4*maxM) bytes: the list of
//       level l >= 1 at (l-1) * (4 + 4*maxM): [uint32 header][uint32 ids[maxM]]
// searchKnn (hnsw.h:1378-1440) is the slim search with threshold_level 0 on these lists: the same
// greedy descent over levels maxlevel..1 and the same bare-bone searchBaseLayerST (hnsw.h:325-480).
int parse_hnsw_graph(const uint8_t *bytes, size_t size, size_t dim, HostGraph *g) {
  Reader r{bytes, size};
  g->kind = HS_KIND_HNSW;
  g->dim = dim;
  const uint64_t offset_level0 = r.get<uint64_t>();
  (void)r.get<uint64_t>();                       // max_elements
  g->n = r.get<uint64_t>();
  g->size_data_per_element = r.get<uint64_t>();
  g->label_offset = r.get<uint64_t>();
  g->offset_data = r.get<uint64_t>();
  g->maxlevel = r.get<int32_t>();
  g->enterpoint = r.get<uint32_t>();
  g->maxM = r.get<uint64_t>();
  g->maxM0 = r.get<uint64_t>();
  g->M = r.get<uint64_t>();
  (void)r.get<double>();                         // mult_
  g->ef_construction = r.get<uint64_t>();
  g->threshold_level = 0;
  if (!r.ok) {
    set_error("truncated .graph header");
    return HS_ERR_IO;
  }
  if (g->maxM > 65535 || g->maxM0 > 65535) {          // list lengths are uint16 (hnsw.h:170-172); also keeps 4 * maxM from wrapping
    set_error("maxM / maxM0 above 65535 in .graph header");
    return HS_ERR_IO;
  }
  const uint64_t links0 = 4 + 4 * g->maxM0, links = 4 + 4 * g->maxM;
  if (offset_level0 != 0 || g->offset_data != links0 || g->label_offset != links0 + 4 * dim ||
      g->size_data_per_element != links0 + 4 * dim + 8 || g->maxM0 == 0 || g->maxM == 0) {
    set_error("unexpected record layout in .graph header (not an hnswlib HNSW index of this dim?)");
    return HS_ERR_IO;
  }
  if (g->n >= (1ull << 31)) {
    set_error("index has >= 2^31 nodes");
    return HS_ERR_UNSUPPORTED;
  }
  if (g->n > 0 && (g->maxlevel < 0 || g->maxlevel >= kMaxLevels || g->enterpoint >= g->n)) {
    set_error("bad maxlevel / enterpoint in .graph header");
    return HS_ERR_IO;
  }
  if (g->n == 0) g->maxlevel = std::max(0, std::min(g->maxlevel, kMaxLevels - 1));   // empty index: maxlevel_ = -1
  const size_t n = g->n, rec = g->size_data_per_element;
  size_t elements_bytes = 0;
  if (__builtin_mul_overflow(n, rec, &elements_bytes)) {
    set_error("element count x record size overflows in .graph header");
    return HS_ERR_IO;
  }
  const uint8_t *elements = r.take(elements_bytes);
  if (!r.ok) {
    set_error("truncated level-0 records in .graph");
    return HS_ERR_IO;
  }
  g->dim_padded = (dim + kRowAlignFloats - 1) / kRowAlignFloats * kRowAlignFloats;
  g->labels.resize(n);
  g->levels.resize(n);
  g->deleted.assign(n, 0);
  g->vec.assign(n * g->dim_padded, 0.f);
  std::vector<const uint8_t *> upper(n, nullptr);
  for (size_t i = 0; i < n; ++i) {
    const uint8_t *e = elements + i * rec;
    uint64_t label;
    std::memcpy(&label, e + g->label_offset, 8);
    g->labels[i] = (uint32_t)label;
    g->deleted[i] = e[2] & 1;                    // DELETE_MARK, hnsw.h:1007-1018
    if (g->deleted[i]) g->has_deleted = true;
    uint16_t cnt;
    std::memcpy(&cnt, e, 2);
    if (cnt > g->maxM0) {
      set_error("level-0 list longer than maxM0 in .graph (node " + std::to_string(i) + ")");
      return HS_ERR_IO;
    }
    std::memcpy(&g->vec[i * g->dim_padded], e + g->offset_data, 4 * dim);
  }
  for (size_t i = 0; i < n; ++i) {
    const uint32_t lsz = r.get<uint32_t>();
    if (!r.ok) {
      set_error("truncated link lists in .graph");
      return HS_ERR_IO;
    }
    int lvl = 0;
    if (lsz) {
      if (lsz % links != 0 || lsz / links > (uint64_t)g->maxlevel) {
        set_error("link list size is not a multiple of the per-level size (node " + std::to_string(i) + ")");
        return HS_ERR_IO;
      }
      lvl = (int)(lsz / links);                  // hnsw.h:866
      upper[i] = r.take(lsz);
      if (!r.ok) {
        set_error("truncated link lists in .graph");
        return HS_ERR_IO;
      }
      for (int l = 1; l <= lvl; ++l) {
        uint16_t cnt;
        std::memcpy(&cnt, upper[i] + (size_t)(l - 1) * links, 2);
        if (cnt > g->maxM) {
          set_error("upper list longer than maxM in .graph (node " + std::to_string(i) + ")");
          return HS_ERR_IO;
        }
      }
    }
    g->levels[i] = (int8_t)lvl;
  }
  if (r.pos != size) {
    set_error("Index seems to be corrupted or unsupported");      // hnsw.h:833-835
    return HS_ERR_IO;
  }
  return flatten_lists(g, [&](size_t i, int l, const uint8_t **ids) -> uint32_t {
    const uint8_t *list = l == 0 ? elements + i * rec : upper[i] + (size_t)(l - 1) * links;
    uint16_t cnt;
    std::memcpy(&cnt, list, 2);                  // getListCount: the low uint16, hnsw.h:170-172
    *ids = list + 4;
    return cnt;
  });
}

}  // namespace hs
