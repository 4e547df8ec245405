/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See hs_oracle.h.
 *
 * Plain-C restatement of the reference's HNSW-Slim search path.  Written from the
 * algorithm; every function names the reference file:line it follows
 * (slim.h = third_party/hnswlib/hnswalg_slim.h).
 */
#include "hs_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static _Thread_local char g_err[256];
const char *hso_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ index -- */
/* In-memory image identical to the reference's: one record per node
 * [int32 level @0][uint32 total_nbr @4][uint64 label @8][ptr @16][float vec @24]
 * (slim.h:127-131) plus one blob per node [uint16 offsets[level]][uint32 ids[total]]
 * (slim.h:1096-1106). */
struct hso_index {
  uint64_t n, size_data_per_element, label_offset, offset_total, offset_data, offset_nbr;
  uint64_t maxM, maxM0, M, ef_construction;
  int32_t maxlevel, threshold_level;
  uint32_t enterpoint;
  uint8_t has_deleted;
  size_t dim;
  int metric;
  char *elements;
  char **blobs;
};

static int rd(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }

/* slim.h:753-815 */
hso_index *hso_load(const char *path, size_t dim, int metric) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    snprintf(g_err, sizeof g_err, "Cannot open file %s", path);
    return NULL;
  }
  hso_index *ix = (hso_index *)calloc(1, sizeof *ix);
  ix->dim = dim;
  ix->metric = metric;
  int bad = 0;
  bad |= rd(f, &ix->n, 8);
  bad |= rd(f, &ix->size_data_per_element, 8);
  bad |= rd(f, &ix->label_offset, 8);
  bad |= rd(f, &ix->offset_total, 8);
  bad |= rd(f, &ix->offset_data, 8);
  bad |= rd(f, &ix->offset_nbr, 8);
  bad |= rd(f, &ix->maxlevel, 4);
  bad |= rd(f, &ix->threshold_level, 4);
  bad |= rd(f, &ix->enterpoint, 4);
  bad |= rd(f, &ix->maxM, 8);
  bad |= rd(f, &ix->maxM0, 8);
  bad |= rd(f, &ix->M, 8);
  bad |= rd(f, &ix->ef_construction, 8);
  bad |= rd(f, &ix->has_deleted, 1);
  if (bad || ix->size_data_per_element != ix->offset_data + 4 * dim) {
    snprintf(g_err, sizeof g_err, "bad header in %s (size_data_per_element=%llu, dim=%zu)", path,
             (unsigned long long)ix->size_data_per_element, dim);
    fclose(f);
    free(ix);
    return NULL;
  }
  ix->elements = (char *)malloc(ix->n * ix->size_data_per_element + 1);
  ix->blobs = (char **)calloc(ix->n + 1, sizeof(char *));
  if (rd(f, ix->elements, ix->n * ix->size_data_per_element)) bad = 1;
  for (uint64_t i = 0; i < ix->n && !bad; i++) {
    uint32_t sz;
    if (rd(f, &sz, 4)) { bad = 1; break; }
    uint32_t total = *(uint32_t *)(ix->elements + i * ix->size_data_per_element + ix->offset_total);
    if (sz == 0 || total == 0) continue;      /* slim.h:800-802 */
    ix->blobs[i] = (char *)malloc(sz);
    if (rd(f, ix->blobs[i], sz)) bad = 1;
  }
  fclose(f);
  if (bad) {
    snprintf(g_err, sizeof g_err, "truncated graph file %s", path);
    hso_free(ix);
    return NULL;
  }
  return ix;
}

/* The un-pruned hnswlib index of the `hnsw` strategy (hnsw_strategy.h:15-61): file written by
 * HierarchicalNSW::saveIndex (hnsw.h:748-779), read by loadIndex (hnsw.h:781-893).  Its
 * searchKnn (hnsw.h:1378-1440) is the same greedy descent over levels maxlevel..1 followed by the
 * same bare-bone searchBaseLayerST (hnsw.h:325-480) that hso_search restates for hnsw_slim with
 * threshold_level 0, so the lists are re-packed into the record + blob image above and searched
 * by the same code.  Level-0 record: [uint32 header, low uint16 = count][uint32 ids[maxM0]]
 * [float vec[dim]][uint64 label]; per node `level` upper lists of 4 + 4*maxM bytes each. */
hso_index *hso_load_hnsw(const char *path, size_t dim, int metric) {
  FILE *f = fopen(path, "rb");
  if (!f) {
    snprintf(g_err, sizeof g_err, "Cannot open file %s", path);
    return NULL;
  }
  uint64_t offset_level0, max_elements, n, sdpe, label_offset, offset_data, maxM, maxM0, M, efc;
  int32_t maxlevel;
  uint32_t enterpoint;
  double mult;
  int bad = 0;
  bad |= rd(f, &offset_level0, 8);
  bad |= rd(f, &max_elements, 8);
  bad |= rd(f, &n, 8);
  bad |= rd(f, &sdpe, 8);
  bad |= rd(f, &label_offset, 8);
  bad |= rd(f, &offset_data, 8);
  bad |= rd(f, &maxlevel, 4);
  bad |= rd(f, &enterpoint, 4);
  bad |= rd(f, &maxM, 8);
  bad |= rd(f, &maxM0, 8);
  bad |= rd(f, &M, 8);
  bad |= rd(f, &mult, 8);
  bad |= rd(f, &efc, 8);
  if (bad || offset_data != 4 + 4 * maxM0 || sdpe != offset_data + 4 * dim + 8 || label_offset != offset_data + 4 * dim) {
    snprintf(g_err, sizeof g_err, "bad hnsw header in %s", path);
    fclose(f);
    return NULL;
  }
  char *lvl0 = (char *)malloc(n * sdpe + 1);
  if (rd(f, lvl0, n * sdpe)) bad = 1;
  hso_index *ix = (hso_index *)calloc(1, sizeof *ix);
  ix->dim = dim;
  ix->metric = metric;
  ix->n = n;
  ix->offset_total = 4;
  ix->label_offset = 8;
  ix->offset_nbr = 16;
  ix->offset_data = 24;
  ix->size_data_per_element = 24 + 4 * dim;
  ix->maxlevel = maxlevel;
  ix->threshold_level = 0;
  ix->enterpoint = enterpoint;
  ix->maxM = maxM;
  ix->maxM0 = maxM0;
  ix->M = M;
  ix->ef_construction = efc;
  ix->elements = (char *)calloc(n * ix->size_data_per_element + 1, 1);
  ix->blobs = (char **)calloc(n + 1, sizeof(char *));
  const size_t links = 4 + 4 * maxM;
  char *up = (char *)malloc(links * 64);
  for (uint64_t i = 0; i < n && !bad; i++) {
    uint32_t lsz;
    if (rd(f, &lsz, 4) || lsz % links || lsz / links > 63) { bad = 1; break; }
    const int level = (int)(lsz / links);                       /* hnsw.h:866 */
    if (lsz && rd(f, up, lsz)) { bad = 1; break; }
    const char *rec = lvl0 + i * sdpe;
    if (((const unsigned char *)rec)[2] & 1) ix->has_deleted = 1;   /* hnsw.h:1007-1018 */
    uint16_t cnt0;
    memcpy(&cnt0, rec, 2);                                      /* getListCount, hnsw.h:170-172 */
    uint32_t total = cnt0;
    uint16_t cnt[64];
    for (int l = 1; l <= level; l++) {
      memcpy(&cnt[l], up + (size_t)(l - 1) * links, 2);
      total += cnt[l];
    }
    char *e = ix->elements + i * ix->size_data_per_element;
    int32_t lv = level;
    memcpy(e, &lv, 4);
    memcpy(e + 4, &total, 4);
    memcpy(e + 8, rec + label_offset, 8);
    memcpy(e + 24, rec + offset_data, 4 * dim);
    if (total == 0) continue;
    char *b = (char *)malloc(2 * (size_t)level + 4 * (size_t)total);
    uint16_t *offs = (uint16_t *)b;
    char *ids = b + 2 * (size_t)level;
    memcpy(ids, rec + 4, 4 * (size_t)cnt0);
    uint32_t run = cnt0;
    for (int l = 1; l <= level; l++) {
      offs[l - 1] = (uint16_t)run;                              /* cumulative count through level l-1 */
      memcpy(ids + 4 * (size_t)run, up + (size_t)(l - 1) * links + 4, 4 * (size_t)cnt[l]);
      run += cnt[l];
    }
    ix->blobs[i] = b;
  }
  free(up);
  free(lvl0);
  fclose(f);
  if (bad) {
    snprintf(g_err, sizeof g_err, "truncated or inconsistent hnsw graph file %s", path);
    hso_free(ix);
    return NULL;
  }
  return ix;
}

void hso_free(hso_index *ix) {
  if (!ix) return;
  if (ix->blobs) {
    for (uint64_t i = 0; i < ix->n; i++) free(ix->blobs[i]);
    free(ix->blobs);
  }
  free(ix->elements);
  free(ix);
}

/* patchFromStream: slim.h:2206-2253 (rows by label, mode HSO_PATCH_ROWS), :2292-2340 (rows inline,
 * HSO_PATCH_INLINE), :2343-2388 (rows by label from a map — the same lookup as far as this restatement goes).
 * Stream: size_t cur_element_count, changed_old_cnt, changed_new_cnt; then per record uint32 id, the first
 * label_offset (old: level + total) or offsetNeighbor (new: + label) bytes of the element record, uint32
 * blob size, the blob, and for new records of the inline form the vector.  The element array grows to the new
 * count (the reference allocated max_elements up front, slim.h:784); enterpoint and maxlevel stay as they are. */
int hso_patch(hso_index *ix, const void *bytes, size_t len, int mode, const float *rows, size_t n_rows) {
  const unsigned char *p = (const unsigned char *)bytes;
  size_t pos = 0;
#define HSO_NEED(nb)                                                   \
  do {                                                                 \
    if ((size_t)(nb) > len - pos) {                                    \
      snprintf(g_err, sizeof g_err, "truncated patch stream");         \
      return -1;                                                       \
    }                                                                  \
  } while (0)
  uint64_t cur = 0, n_old = 0, n_new = 0;
  HSO_NEED(24);
  memcpy(&cur, p, 8);
  memcpy(&n_old, p + 8, 8);
  memcpy(&n_new, p + 16, 8);
  pos = 24;
  if (cur < ix->n) {
    snprintf(g_err, sizeof g_err, "patch shrinks the index");
    return -1;
  }
  if (cur > ix->n) {
    char *e = (char *)realloc(ix->elements, cur * ix->size_data_per_element + 1);
    char **b = (char **)realloc(ix->blobs, (cur + 1) * sizeof(char *));
    if (!e || !b) {
      snprintf(g_err, sizeof g_err, "out of memory");
      return -1;
    }
    memset(e + ix->n * ix->size_data_per_element, 0, (cur - ix->n) * ix->size_data_per_element);
    for (uint64_t i = ix->n; i <= cur; i++) b[i] = NULL;
    ix->elements = e;
    ix->blobs = b;
    ix->n = cur;                                   /* slim.h:2209, :2294, :2346 */
  }
  for (uint64_t i = 0; i < n_old + n_new; i++) {
    uint32_t id, bsz;
    HSO_NEED(4);
    memcpy(&id, p + pos, 4);
    pos += 4;
    if (id >= ix->n) {
      snprintf(g_err, sizeof g_err, "patch: node id out of range");
      return -1;
    }
    char *e = ix->elements + (size_t)id * ix->size_data_per_element;
    const size_t head = i < n_old ? ix->label_offset : ix->offset_nbr;     /* slim.h:2223-2226 */
    HSO_NEED(head);
    memcpy(e, p + pos, head);
    pos += head;
    HSO_NEED(4);
    memcpy(&bsz, p + pos, 4);
    pos += 4;
    free(ix->blobs[id]);
    ix->blobs[id] = NULL;
    if (bsz) {                                                              /* slim.h:2238-2247 */
      HSO_NEED(bsz);
      ix->blobs[id] = (char *)malloc(bsz);
      memcpy(ix->blobs[id], p + pos, bsz);
      pos += bsz;
    }
    if (i >= n_old) {
      if (mode == HSO_PATCH_INLINE) {                                       /* slim.h:2332-2334 */
        HSO_NEED(4 * ix->dim);
        memcpy(e + ix->offset_data, p + pos, 4 * ix->dim);
        pos += 4 * ix->dim;
      } else {                                                              /* slim.h:2227-2229, :2362-2364 */
        uint64_t label;
        memcpy(&label, e + ix->label_offset, 8);
        if (!rows || label >= n_rows) {
          snprintf(g_err, sizeof g_err, "patch: no row for label %llu", (unsigned long long)label);
          return -1;
        }
        memcpy(e + ix->offset_data, rows + label * ix->dim, 4 * ix->dim);
      }
    }
  }
#undef HSO_NEED
  return 0;
}

void hso_get_info(const hso_index *ix, hso_info *o) {
  o->n = ix->n;
  o->size_data_per_element = ix->size_data_per_element;
  o->maxM = ix->maxM;
  o->maxM0 = ix->maxM0;
  o->M = ix->M;
  o->ef_construction = ix->ef_construction;
  o->dim = ix->dim;
  o->maxlevel = ix->maxlevel;
  o->threshold_level = ix->threshold_level;
  o->enterpoint = ix->enterpoint;
  o->has_deleted = ix->has_deleted;
}

static inline const char *elem(const hso_index *ix, uint32_t i) {
  return ix->elements + (size_t)i * ix->size_data_per_element;
}
int hso_node_level(const hso_index *ix, uint32_t i) { return *(const int32_t *)elem(ix, i); }
static inline uint32_t node_total(const hso_index *ix, uint32_t i) {
  return *(const uint32_t *)(elem(ix, i) + ix->offset_total);
}
uint64_t hso_node_label(const hso_index *ix, uint32_t i) {
  uint64_t l;
  memcpy(&l, elem(ix, i) + ix->label_offset, 8);
  return l;
}
const float *hso_node_vector(const hso_index *ix, uint32_t i) {
  return (const float *)(elem(ix, i) + ix->offset_data);
}
/* slim.h:1776-1781: bit 0 of byte 6 of the record */
static inline int node_deleted(const hso_index *ix, uint32_t i) {
  return ((const unsigned char *)elem(ix, i))[6] & 1;
}

/* slim.h:245-260 (level slice) and :363-369 (level-0 slice) */
int hso_node_neighbors(const hso_index *ix, uint32_t i, int level, const uint32_t **ids) {
  const char *blob = ix->blobs[i];
  *ids = NULL;
  if (!blob) return 0;
  int el = hso_node_level(ix, i);
  if (level > el) return 0;
  const uint16_t *offs = (const uint16_t *)blob;
  uint32_t begin = level == 0 ? 0 : offs[level - 1];
  uint32_t end = level == el ? node_total(ix, i) : offs[level];
  *ids = (const uint32_t *)(blob + 2 * (size_t)el) + begin;
  return (int)(end - begin);
}

/* -------------------------------------------------------------- distances -- */
/* space_l2.h:6-20 / space_ip.h:6-19 */
static float dist_seq(const float *a, const float *b, size_t dim, int metric) {
  float res = 0;
  if (metric == HSO_IP) {
    for (size_t i = 0; i < dim; i++) res += a[i] * b[i];
    return 1.0f - res;
  }
  for (size_t i = 0; i < dim; i++) {
    float t = a[i] - b[i];
    res += t * t;
  }
  return res;
}

/* space_l2.h:25-54 (mul then add per lane) and space_ip.h:146-204 (fma per lane), 16 lanes, chunks in
 * order.  The source then adds the 16 lane sums left to right (space_l2.h:49-51) / _mm512_reduce_add_ps,
 * but the reference is compiled -Ofast (CMakeLists.txt:14), which lets the compiler re-associate: the
 * x86-64-v4 build of oracle/_ref (g++ 13, `objdump -d libhsref_slim_v4.so`) folds the lanes as a tree,
 * (j, j+8) -> (j, j+4) -> (j, j+2) -> (j, j+1), for both metrics.  This restates THAT association, so
 * that distances — and with them every near-tie of the search — are bit-identical to the reference as
 * built here (tests/test_oracle.py::test_ref_order_is_bit_exact).  Requires dim % 16 == 0, else SEQ. */
static float dist_ref(const float *a, const float *b, size_t dim, int metric) {
  if (dim % 16) return dist_seq(a, b, dim, metric);
  float lane[16];
  for (int j = 0; j < 16; j++) lane[j] = 0.f;
  for (size_t i = 0; i < dim; i += 16)
    for (int j = 0; j < 16; j++) {
      if (metric == HSO_IP) {
        lane[j] = fmaf(a[i + j], b[i + j], lane[j]);
      } else {
        float d = a[i + j] - b[i + j];
        lane[j] = lane[j] + d * d;
      }
    }
  for (int s = 8; s >= 1; s >>= 1)
    for (int j = 0; j < s; j++) lane[j] = lane[j] + lane[j + s];
  return metric == HSO_IP ? 1.0f - lane[0] : lane[0];
}

/* The CUDA kernel's association: `team` lanes per row, lane t owns the 4-float
 * chunks t, t+team, ...; elements past dim are zeros (rows are zero-padded in HBM).
 * Every lane keeps TWO partial sums (the packed FADD2/FFMA2 pair of traverse_common.cuh
 * acc4): `lo` takes elements 0 and 2 of each chunk, `hi` elements 1 and 3, in chunk order;
 * the lane sum is lo + hi; lanes are then combined by xor-shuffles team/2 ... 1. */
static float dist_gpu(const float *a, const float *b, size_t dim, int metric, int team) {
  float lane[32];
  size_t chunks = (dim + 3) / 4;
  for (int t = 0; t < team; t++) {
    float acc[2] = {0.f, 0.f};
    for (size_t c = (size_t)t; c < chunks; c += (size_t)team)
      for (int e = 0; e < 4; e++) {
        size_t i = 4 * c + e;
        float x = i < dim ? a[i] : 0.f, y = i < dim ? b[i] : 0.f;
        if (metric == HSO_IP) {
          acc[e & 1] = fmaf(x, y, acc[e & 1]);
        } else {
          float d = x - y;
          acc[e & 1] = fmaf(d, d, acc[e & 1]);
        }
      }
    lane[t] = acc[0] + acc[1];
  }
  for (int off = team / 2; off >= 1; off >>= 1) {
    float nxt[32];
    for (int t = 0; t < team; t++) nxt[t] = lane[t] + lane[t ^ off];
    memcpy(lane, nxt, sizeof(float) * team);
  }
  return metric == HSO_IP ? 1.0f - lane[0] : lane[0];
}

/* one fp32 fma chain in index order: the association of the exact-kNN CUDA kernel
 * (hnsw_slim_b200/csrc/bruteforce.cu) */
static float dist_seqfma(const float *a, const float *b, size_t dim, int metric) {
  float acc = 0.f;
  if (metric == HSO_IP) {
    for (size_t i = 0; i < dim; i++) acc = fmaf(a[i], b[i], acc);
    return 1.0f - acc;
  }
  for (size_t i = 0; i < dim; i++) {
    float t = a[i] - b[i];
    acc = fmaf(t, t, acc);
  }
  return acc;
}

float hso_dist(const float *a, const float *b, size_t dim, int metric, int order, int team) {
  if (order == HSO_ORDER_SEQFMA) return dist_seqfma(a, b, dim, metric);
  if (order == HSO_ORDER_REF) return dist_ref(a, b, dim, metric);
  if (order == HSO_ORDER_GPU) return dist_gpu(a, b, dim, metric, team > 0 ? team : 8);
  return dist_seq(a, b, dim, metric);
}

/* ------------------------------------------------------------------ heaps -- */
typedef struct {
  float d;
  uint32_t id;
} pairfu;

/* Binary heap on an array with "less" = (a.d < b.d) for a max-heap (maxheap=1)
 * or (a.d > b.d) for a min-heap, as std::push_heap / std::pop_heap with
 * compare_by_first / compare_by_first_rev (slim.h:169-183). */
static inline int heap_less(pairfu a, pairfu b, int maxheap) { return maxheap ? a.d < b.d : a.d > b.d; }

static void heap_push(pairfu *h, size_t n /* new size, value at h[n-1] */, int maxheap) {
  pairfu v = h[n - 1];
  size_t hole = n - 1;
  while (hole > 0) {
    size_t parent = (hole - 1) / 2;
    if (!heap_less(h[parent], v, maxheap)) break;
    h[hole] = h[parent];
    hole = parent;
  }
  h[hole] = v;
}

/* moves the top to h[n-1], leaves a heap in h[0..n-1) */
static void heap_pop(pairfu *h, size_t n, int maxheap) {
  if (n <= 1) return;
  pairfu top = h[0], v = h[n - 1];
  size_t len = n - 1, hole = 0, child = 0;
  while (child < (len - 1) / 2) {       /* walk the hole down along the better child */
    child = 2 * (child + 1);
    if (heap_less(h[child], h[child - 1], maxheap)) child--;
    h[hole] = h[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h[hole] = h[child - 1];
    hole = child - 1;
  }
  while (hole > 0) {                     /* then sift the displaced value up */
    size_t parent = (hole - 1) / 2;
    if (!heap_less(h[parent], v, maxheap)) break;
    h[hole] = h[parent];
    hole = parent;
  }
  h[hole] = v;
  h[n - 1] = top;
}

/* ----------------------------------------------------------------- search -- */
/* optional trace of the expanded nodes of ONE query (hso_debug_trace) */
static _Thread_local uint32_t *g_trace = NULL;
static _Thread_local size_t g_trace_cap = 0, g_trace_n = 0;
static inline void trace_node(uint32_t id) {
  if (g_trace && g_trace_n < g_trace_cap) g_trace[g_trace_n] = id;
  if (g_trace) g_trace_n++;
}

typedef struct {
  pairfu *top;     /* max-heap, ef+1 */
  pairfu *cand;    /* min-heap, grows */
  size_t cand_cap;
  uint16_t *visited;
  uint16_t tag;
  size_t n;
} scratch;

static void cand_reserve(scratch *s, size_t need) {
  if (need <= s->cand_cap) return;
  size_t c = s->cand_cap ? s->cand_cap : 1024;
  while (c < need) c *= 2;
  s->cand = (pairfu *)realloc(s->cand, c * sizeof(pairfu));
  s->cand_cap = c;
}

/* One beam pass over `layer` (slim.h:222-316 for layer > 0; slim.h:321-457 for
 * layer 0).  stop_needs_full: the stop rule also requires |top| == ef
 * (slim.h:237 and the non-bare-bone rule :346-347).  check_deleted: skip
 * delete-marked nodes when filling `top` (slim.h:297, :418). */
/* n_ties (may be NULL) counts the exact ties at the ef boundary: a result trimmed from `top` while the entry that
 * stays behind as the new worst carries the bit-identical distance.  Only then can the reference go on to EXPAND
 * an entry that is no longer among its results (the stop rule is a strict `>`, slim.h:237, :339-340) — the one
 * place where an engine that keeps candidates and results in a single pool has to take special care. */
static void beam_layer(const hso_index *ix, const float *q, int layer, scratch *s, size_t *top_size,
                       size_t ef, float *lower_bound, int stop_needs_full, int check_deleted,
                       int order, int team, uint32_t *n_dist, uint32_t *n_hops, uint32_t *n_ties) {
  size_t csz = *top_size;
  cand_reserve(s, csz + 1);
  memcpy(s->cand, s->top, csz * sizeof(pairfu));
  /* std::make_heap on the copy */
  for (size_t i = 1; i <= csz; i++) heap_push(s->cand, i, 0);

  while (csz > 0) {
    pairfu cur = s->cand[0];
    int stop = cur.d > *lower_bound;
    if (stop_needs_full) stop = stop && (*top_size == ef);
    if (stop) break;
    heap_pop(s->cand, csz, 0);
    csz--;

    const uint32_t *ids;
    int cnt = hso_node_neighbors(ix, cur.id, layer, &ids);
    if (cnt == 0) continue;
    if (n_hops) (*n_hops)++;
    trace_node(cur.id);
    for (int j = 0; j < cnt; j++) {
      uint32_t c = ids[j];
      if (s->visited[c] == s->tag) continue;
      s->visited[c] = s->tag;
      float d = hso_dist(q, hso_node_vector(ix, c), ix->dim, ix->metric, order, team);
      if (n_dist) (*n_dist)++;
      if (*top_size < ef || *lower_bound > d) {
        cand_reserve(s, csz + 1);
        s->cand[csz].d = d;
        s->cand[csz].id = c;
        csz++;
        heap_push(s->cand, csz, 0);
        if (!check_deleted || !node_deleted(ix, c)) {
          s->top[*top_size].d = d;
          s->top[*top_size].id = c;
          (*top_size)++;
          heap_push(s->top, *top_size, 1);
        }
        while (*top_size > ef) {
          heap_pop(s->top, *top_size, 1);
          (*top_size)--;
          if (n_ties && *top_size > 0 && s->top[*top_size].d == s->top[0].d) (*n_ties)++;
        }
        if (*top_size > 0) *lower_bound = s->top[0].d;
      }
    }
  }
}

static int cmp_pair(const void *a, const void *b) {
  const pairfu *x = (const pairfu *)a, *y = (const pairfu *)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}

/* slim.h:2030-2131 */
static void search_one(const hso_index *ix, const float *q, size_t k, size_t ef_in, scratch *s,
                       int order, int team, uint32_t *out_labels, float *out_dists,
                       uint32_t *n_dist, uint32_t *n_hops, uint32_t *n_ties) {
  uint32_t nd = 0, nh = 0, nt = 0;
  uint32_t cur = ix->enterpoint;
  float curdist = hso_dist(q, hso_node_vector(ix, cur), ix->dim, ix->metric, order, team);
  nd++;
  /* VisitedList::reset, visited_list_pool.h:22-28 */
  s->tag++;
  if (s->tag == 0) {
    memset(s->visited, 0, sizeof(uint16_t) * s->n);
    s->tag++;
  }
  /* greedy descent, slim.h:2040-2078: strict improvement, rescan while improved,
   * no visited marking */
  for (int level = ix->maxlevel; level > ix->threshold_level; level--) {
    int changed = 1;
    while (changed) {
      changed = 0;
      const uint32_t *ids;
      int cnt = hso_node_neighbors(ix, cur, level, &ids);
      if (cnt == 0) continue;
      nh++;
      for (int i = 0; i < cnt; i++) {
        float d = hso_dist(q, hso_node_vector(ix, ids[i]), ix->dim, ix->metric, order, team);
        nd++;
        if (d < curdist) {
          curdist = d;
          cur = ids[i];
          changed = 1;
        }
      }
    }
  }
  size_t ef = ef_in > k ? ef_in : k;     /* slim.h:2080 */
  size_t top_size = 1;
  s->top[0].d = curdist;
  s->top[0].id = cur;
  s->visited[cur] = s->tag;
  float lower = node_deleted(ix, cur) ? 3.402823466e+38f : curdist;   /* slim.h:2104-2106 */

  int thr = ix->threshold_level < ix->maxlevel ? ix->threshold_level : ix->maxlevel;
  for (int level = thr; level > 0; level--)     /* slim.h:2108-2113 */
    beam_layer(ix, q, level, s, &top_size, ef, &lower, 1, 1, order, team, &nd, &nh, &nt);
  int bare = !ix->has_deleted;                  /* slim.h:2114-2123 */
  beam_layer(ix, q, 0, s, &top_size, ef, &lower, !bare, !bare, order, team, &nd, &nh, &nt);

  /* slim.h:2126-2130 returns an unordered k-subset of labels; we sort so that the
   * result is canonical: (dist, internal id) ascending */
  qsort(s->top, top_size, sizeof(pairfu), cmp_pair);
  for (size_t i = 0; i < k; i++) {
    if (i < top_size) {
      out_labels[i] = (uint32_t)hso_node_label(ix, s->top[i].id);
      if (out_dists) out_dists[i] = s->top[i].d;
    } else {
      out_labels[i] = 0xFFFFFFFFu;
      if (out_dists) out_dists[i] = INFINITY;
    }
  }
  if (n_dist) *n_dist = nd;
  if (n_hops) *n_hops = nh;
  if (n_ties) *n_ties = nt;
}

int hso_search(const hso_index *ix, const float *queries, size_t nq, size_t k, size_t ef, int order,
               int team, int threads, uint32_t *out_labels, float *out_dists, uint32_t *n_dist,
               uint32_t *n_hops) {
  return hso_search_ties(ix, queries, nq, k, ef, order, team, threads, out_labels, out_dists, n_dist, n_hops, NULL);
}

int hso_search_ties(const hso_index *ix, const float *queries, size_t nq, size_t k, size_t ef, int order,
                    int team, int threads, uint32_t *out_labels, float *out_dists, uint32_t *n_dist,
                    uint32_t *n_hops, uint32_t *n_ties) {
  if (ix->n == 0) return 0;                     /* slim.h:2031-2032 */
  size_t efx = ef > k ? ef : k;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_num_procs();
#else
  threads = 1;
#endif
#pragma omp parallel num_threads(threads)
  {
    scratch s;
    memset(&s, 0, sizeof s);
    s.n = ix->n;
    s.top = (pairfu *)malloc((efx + 2) * sizeof(pairfu));
    s.visited = (uint16_t *)calloc(ix->n, sizeof(uint16_t));
    s.tag = 0;
#pragma omp for schedule(dynamic)
    for (size_t i = 0; i < nq; i++)
      search_one(ix, queries + i * ix->dim, k, ef, &s, order, team, out_labels + i * k,
                 out_dists ? out_dists + i * k : NULL, n_dist ? n_dist + i : NULL,
                 n_hops ? n_hops + i : NULL, n_ties ? n_ties + i : NULL);
    free(s.top);
    free(s.cand);
    free(s.visited);
  }
  return 0;
}

/* ------------------------------------------------- the ENGINE's algorithm -- */
/* hso_search_pool: a CPU restatement, not of the reference, but of what the CUDA traversal kernel does
 * (hnsw_slim_b200/csrc/traverse_fp32.cu + RegPool32 / RegPool32C in traverse_common.cuh) — one pool of at most ef
 * (distance, id) entries spread over 32 columns ("lanes") of `slots` entries, an "expanded" mark per entry, the
 * seven-entry ghost list for exact ties — with the kernel's placement and tie rules spelled out step by step:
 *   append      while the pool has room, entry number e goes to lane e % 32, slot e / 32
 *   worst       the largest distance word; among equal ones the LOWEST lane, then the lowest slot, is displaced
 *   pop         the smallest distance word among unexpanded entries; lowest lane, then lowest slot
 *   ghost       when the displaced entry is unexpanded and ANOTHER LANE's column maximum carries the same
 *               distance word it is remembered, and expanded once the pool has no unexpanded entry left, if its
 *               distance still equals the pool's worst one and the pool is full
 * It exists to measure, on the CPU and at any scale, where a single-pool engine parts from the reference's two
 * heaps (slim.h:321-457): everywhere except at exact fp32 ties at the ef boundary the two must agree to the bit,
 * ids, distances and per-query counters (tests/test_oracle.py), and what remains at the ties — the one blind spot
 * of the ghost rule, a tie inside ONE column — can be counted.  Register pools only (ef <= 256), threshold_level 0,
 * no delete marks; the visited set is exact (the kernel's tables are, short of resets that only add evaluations). */
typedef struct {
  uint32_t kd, id;        /* ordered-uint distance (0 = empty slot), node id */
  uint8_t expanded;
} pslot;

static inline uint32_t f2ord_u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float ord2f_u(uint32_t u) {
  uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &v, 4);
  return f;
}
static int pool_slots_for(size_t ef) {   /* slots_variant, traverse_fp32.cu */
  return ef <= 64 ? 2 : (ef <= 128 ? 4 : (ef <= 160 ? 5 : (ef <= 192 ? 6 : (ef <= 224 ? 7 : (ef <= 256 ? 8 : 0)))));
}

static void search_one_pool(const hso_index *ix, const float *q, size_t k, size_t ef_in, uint16_t *visited,
                            uint16_t tag, pslot *pool, int team, uint32_t *out_labels, float *out_dists,
                            uint32_t *n_dist, uint32_t *n_hops, uint32_t *n_ghosts) {
  const int order = HSO_ORDER_GPU;
  uint32_t nd = 0, nh = 0, ng = 0, nblind = 0;
  uint32_t cur = ix->enterpoint;
  float curdist = hso_dist(q, hso_node_vector(ix, cur), ix->dim, ix->metric, order, team);
  nd++;
  for (int level = ix->maxlevel; level > ix->threshold_level; level--) {     /* the descent is the reference's */
    int changed = 1;
    while (changed) {
      changed = 0;
      const uint32_t *ids;
      int cnt = hso_node_neighbors(ix, cur, level, &ids);
      if (cnt == 0) continue;
      nh++;
      for (int i = 0; i < cnt; i++) {
        float d = hso_dist(q, hso_node_vector(ix, ids[i]), ix->dim, ix->metric, order, team);
        nd++;
        if (d < curdist) {
          curdist = d;
          cur = ids[i];
          changed = 1;
        }
      }
    }
  }
  const size_t ef = ef_in > k ? ef_in : k;
  const int slots = pool_slots_for(ef);
  const int cap = slots * 32;
  memset(pool, 0, sizeof(pslot) * (size_t)cap);
#define PS(lane, slot) pool[(slot) * 32 + (lane)]
  size_t size = 1;
  PS(0, 0).kd = f2ord_u(curdist);
  PS(0, 0).id = cur;
  visited[cur] = tag;
  uint32_t gd[7], gi[7], gcnt = 0;          /* the ghost list (kGhostCap = 7) */

  for (;;) {
    /* pop_closest_unexpanded */
    uint32_t g = 0xffffffffu;
    for (int e = 0; e < cap; e++)
      if (pool[e].kd && !pool[e].expanded && pool[e].kd < g) g = pool[e].kd;
    uint32_t node = 0xffffffffu;
    if (g != 0xffffffffu) {
      for (int lane = 0; lane < 32 && node == 0xffffffffu; lane++)
        for (int sl = 0; sl < slots; sl++)
          if (PS(lane, sl).kd == g && !PS(lane, sl).expanded) {
            PS(lane, sl).expanded = 1;
            node = PS(lane, sl).id;
            break;
          }
    } else if (gcnt) {
      /* take_ghost */
      uint32_t worst = 0;
      for (int e = 0; e < cap; e++)
        if (pool[e].kd > worst) worst = pool[e].kd;
      int o = -1;
      for (uint32_t j = 0; j < gcnt; j++)
        if (size >= ef && gd[j] == worst) {
          o = (int)j;
          break;
        }
      if (o < 0) {
        gcnt = 0;
      } else {
        gd[o] = 0;             /* taken */
        node = gi[o];
        ng++;
      }
    }
    if (node == 0xffffffffu) break;

    const uint32_t *ids;
    int cnt = hso_node_neighbors(ix, node, 0, &ids);
    if (cnt > 0) nh++;
    if (cnt > 0) trace_node(node);
    for (int seg = 0; seg < cnt; seg += 32) {      /* the kernel walks the row 32 ids at a time */
      uint32_t cid[32];
      float cdist[32];
      int count = 0;
      for (int j = seg; j < cnt && j < seg + 32; j++) {
        uint32_t c = ids[j];
        if (visited[c] == tag) continue;
        visited[c] = tag;
        cid[count++] = c;
      }
      for (int j = 0; j < count; j++)
        cdist[j] = hso_dist(q, hso_node_vector(ix, cid[j]), ix->dim, ix->metric, order, team);
      nd += (uint32_t)count;
      /* admit, candidate by candidate in list order */
      int j = 0;
      while (j < count && size < ef) {              /* room left: appended unconditionally */
        const size_t e = size++;
        PS(e & 31, e >> 5).kd = f2ord_u(cdist[j]);
        PS(e & 31, e >> 5).id = cid[j];
        PS(e & 31, e >> 5).expanded = 0;
        j++;
      }
      for (; j < count; j++) {
        uint32_t cm[32], worst = 0;
        for (int lane = 0; lane < 32; lane++) {
          cm[lane] = 0;
          for (int sl = 0; sl < slots; sl++)
            if (PS(lane, sl).kd > cm[lane]) cm[lane] = PS(lane, sl).kd;
          if (cm[lane] > worst) worst = cm[lane];
        }
        const uint32_t cd = f2ord_u(cdist[j]);
        if (!(cd < worst)) continue;                /* the reference's strict lowerBound > dist */
        int owner = -1, lanes_at_worst = 0;
        for (int lane = 0; lane < 32; lane++)
          if (cm[lane] == worst) {
            if (owner < 0) owner = lane;
            lanes_at_worst++;
          }
        int sl = 0;
        while (PS(owner, sl).kd != worst) sl++;
        if (lanes_at_worst > 1 && !PS(owner, sl).expanded && gcnt < 7) {     /* ghost_append */
          gd[gcnt] = worst;
          gi[gcnt] = PS(owner, sl).id;
          gcnt++;
        }
        if (lanes_at_worst == 1 && !PS(owner, sl).expanded) {     /* the blind spot: the tie partner sits in the same column */
          for (int s2 = sl + 1; s2 < slots; s2++)
            if (PS(owner, s2).kd == worst) {
              nblind++;
              break;
            }
        }
        PS(owner, sl).kd = cd;
        PS(owner, sl).id = cid[j];
        PS(owner, sl).expanded = 0;
      }
    }
  }
#undef PS
  /* results: the k smallest (distance word, id) keys */
  size_t used = 0;
  pairfu *res = (pairfu *)malloc(sizeof(pairfu) * (size_t)cap);
  for (int e = 0; e < cap; e++)
    if (pool[e].kd) {
      res[used].d = ord2f_u(pool[e].kd);
      res[used].id = pool[e].id;
      used++;
    }
  qsort(res, used, sizeof(pairfu), cmp_pair);
  for (size_t i = 0; i < k; i++) {
    if (i < used) {
      out_labels[i] = (uint32_t)hso_node_label(ix, res[i].id);
      if (out_dists) out_dists[i] = res[i].d;
    } else {
      out_labels[i] = 0xFFFFFFFFu;
      if (out_dists) out_dists[i] = INFINITY;
    }
  }
  free(res);
  if (n_dist) *n_dist = nd;
  if (n_hops) *n_hops = nh;
  if (n_ghosts) *n_ghosts = ng | (nblind << 16);
}

int hso_search_pool(const hso_index *ix, const float *queries, size_t nq, size_t k, size_t ef, int team, int threads,
                    uint32_t *out_labels, float *out_dists, uint32_t *n_dist, uint32_t *n_hops, uint32_t *n_ghosts) {
  if (ix->n == 0) return 0;
  const size_t efx = ef > k ? ef : k;
  if (pool_slots_for(efx) == 0 || ix->threshold_level != 0 || ix->has_deleted) {
    snprintf(g_err, sizeof g_err, "hso_search_pool: register pools only (ef <= 256), threshold_level 0, no deletes");
    return -1;
  }
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_num_procs();
#else
  threads = 1;
#endif
#pragma omp parallel num_threads(threads)
  {
    uint16_t *visited = (uint16_t *)calloc(ix->n, sizeof(uint16_t));
    pslot *pool = (pslot *)malloc(sizeof(pslot) * 8 * 32);
    uint16_t tag = 0;
#pragma omp for schedule(dynamic)
    for (size_t i = 0; i < nq; i++) {
      tag++;
      if (tag == 0) {
        memset(visited, 0, sizeof(uint16_t) * ix->n);
        tag++;
      }
      search_one_pool(ix, queries + i * ix->dim, k, ef, visited, tag, pool, team, out_labels + i * k,
                      out_dists ? out_dists + i * k : NULL, n_dist ? n_dist + i : NULL, n_hops ? n_hops + i : NULL,
                      n_ghosts ? n_ghosts + i : NULL);
    }
    free(visited);
    free(pool);
  }
  return 0;
}

/* The base-layer expansion order of one query under hso_search (which = 0) or hso_search_pool (which = 1): node
 * ids in the order they were expanded; returns how many there were (may exceed cap). */
size_t hso_debug_trace(const hso_index *ix, const float *query, size_t k, size_t ef, int which, int team,
                       uint32_t *out_ids, size_t cap) {
  uint32_t lab[4096];
  if (k > 4096) return 0;
  g_trace = out_ids;
  g_trace_cap = cap;
  g_trace_n = 0;
  if (which == 0)
    hso_search_ties(ix, query, 1, k, ef, HSO_ORDER_GPU, team, 1, lab, NULL, NULL, NULL, NULL);
  else
    hso_search_pool(ix, query, 1, k, ef, team, 1, lab, NULL, NULL, NULL, NULL);
  g_trace = NULL;
  return g_trace_n;
}

/* ------------------------------------------------------------ brute force -- */
/* std::priority_queue<std::pair<float, size_t>>: max-heap ordered by the PAIR
 * (dist, then label) — bruteforce.h:109.  We keep the pair order explicitly. */
typedef struct {
  float d;
  uint64_t label;
} pairfl;
static inline int pl_less(pairfl a, pairfl b) { return a.d < b.d || (a.d == b.d && a.label < b.label); }
static void pl_push(pairfl *h, size_t n) {
  pairfl v = h[n - 1];
  size_t hole = n - 1;
  while (hole > 0) {
    size_t p = (hole - 1) / 2;
    if (!pl_less(h[p], v)) break;
    h[hole] = h[p];
    hole = p;
  }
  h[hole] = v;
}
static void pl_pop(pairfl *h, size_t n) {
  if (n <= 1) return;
  pairfl top = h[0], v = h[n - 1];
  size_t len = n - 1, hole = 0;
  for (;;) {
    size_t l = 2 * hole + 1, r = l + 1, c;
    if (l >= len) break;
    c = (r < len && pl_less(h[l], h[r])) ? r : l;
    if (!pl_less(v, h[c])) break;
    h[hole] = h[c];
    hole = c;
  }
  h[hole] = v;
  h[n - 1] = top;
}

/* bruteforce.h:106-135; rows drained farthest-first (brute_force_strategy.h:27-31) */
int hso_bruteforce(const float *base, size_t n, size_t dim, int metric, int order, int team,
                   const float *queries, size_t nq, size_t k, int threads, uint32_t *out_labels,
                   float *out_dists) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_num_procs();
#else
  threads = 1;
#endif
  if (k > n) k = n;
#pragma omp parallel for schedule(dynamic) num_threads(threads)
  for (size_t qi = 0; qi < nq; qi++) {
    const float *q = queries + qi * dim;
    pairfl *h = (pairfl *)malloc((k + 2) * sizeof(pairfl));
    size_t sz = 0;
    for (size_t i = 0; i < k; i++) {           /* first k unconditionally, :112-118 */
      h[sz].d = hso_dist(q, base + i * dim, dim, metric, order, team);
      h[sz].label = i;
      sz++;
      pl_push(h, sz);
    }
    float last = sz ? h[0].d : 3.402823466e+38f;
    for (size_t i = k; i < n; i++) {           /* :120-133, admits dist <= lastdist */
      float d = hso_dist(q, base + i * dim, dim, metric, order, team);
      if (d <= last) {
        h[sz].d = d;
        h[sz].label = i;
        sz++;
        pl_push(h, sz);
        if (sz > k) {
          pl_pop(h, sz);
          sz--;
        }
        if (sz) last = h[0].d;
      }
    }
    size_t j = 0;
    while (sz > 0) {                           /* drain: farthest first */
      out_labels[qi * k + j] = (uint32_t)h[0].label;
      if (out_dists) out_dists[qi * k + j] = h[0].d;
      pl_pop(h, sz);
      sz--;
      j++;
    }
    free(h);
  }
  return 0;
}

/* ----------------------------------------------------------------- recall -- */
static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : x > y;
}

/* solve_strategy.h:67-103 */
double hso_recall(const float *base, size_t dim, const float *queries, size_t nq, const uint32_t *knn,
                  size_t K, const uint32_t *gt, size_t gt_k, int metric) {
  long hit = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : hit)
  for (size_t i = 0; i < nq; i++) {
    pairfu *t = (pairfu *)malloc(gt_k * sizeof(pairfu));
    for (size_t j = 0; j < gt_k; j++) {
      uint32_t g = gt[i * gt_k + j];
      t[j].id = g;
      t[j].d = dist_seq(queries + i * dim, base + (size_t)g * dim, dim, metric);  /* :85 L2Sqr */
    }
    qsort(t, gt_k, sizeof(pairfu), cmp_pair);     /* :87 sort pairs: ties -> smaller id */
    uint32_t *a = (uint32_t *)malloc(K * sizeof(uint32_t)), *b = (uint32_t *)malloc(K * sizeof(uint32_t));
    for (size_t j = 0; j < K; j++) {
      a[j] = knn[i * K + j];
      b[j] = t[j].id;
    }
    qsort(a, K, 4, cmp_u32);
    qsort(b, K, 4, cmp_u32);
    size_t x = 0, y = 0;                          /* std::set_intersection */
    while (x < K && y < K) {
      if (a[x] < b[y]) x++;
      else if (b[y] < a[x]) y++;
      else { hit++; x++; y++; }
    }
    free(a);
    free(b);
    free(t);
  }
  return (double)hit / (double)(nq * K);
}
