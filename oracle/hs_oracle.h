/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * hs_oracle: a plain-C CPU restatement of the HNSW-Slim search hot path of the
 * reference (InfiniteNightmare/HNSW-Slim), written from the algorithm, citing the
 * reference file:line each function follows.  It is the checker for the CUDA path.
 * Pinned (tests/test_oracle_vs_reference.py) against the reference itself, compiled
 * unmodified into oracle/_ref/ by oracle/Makefile.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 */
#ifndef HS_ORACLE_H
#define HS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hso_index hso_index;

enum { HSO_L2 = 0, HSO_IP = 1 };

/* Floating-point association of the distance sum.
 *   HSO_ORDER_SEQ : plain left-to-right scalar loop (space_l2.h:6-20)
 *   HSO_ORDER_REF : the AVX-512 kernel's association — 16 partial sums, lane j
 *                   takes elements j, j+16, ...; L2 uses mul+add, IP uses fma;
 *                   lanes added left to right (space_l2.h:25-54, space_ip.h:146-204)
 *   HSO_ORDER_GPU : the CUDA kernel's association — `team` lanes, lane t takes the
 *                   4-float chunks t, t+team, ...; one fma chain per lane; xor
 *                   butterfly team/2 ... 1 (hnsw_slim_b200/csrc/traverse_fp32.cu)
 */
/*   HSO_ORDER_SEQFMA : one fma chain in index order (the exact-kNN CUDA kernel, bruteforce.cu) */
enum { HSO_ORDER_SEQ = 0, HSO_ORDER_REF = 1, HSO_ORDER_GPU = 2, HSO_ORDER_SEQFMA = 3 };

typedef struct {
  uint64_t n, size_data_per_element, maxM, maxM0, M, ef_construction, dim;
  int32_t maxlevel, threshold_level;
  uint32_t enterpoint;
  int32_t has_deleted;
} hso_info;

/* slim.h:753-815 (file written by slim.h:717-751) */
hso_index *hso_load(const char *graph_path, size_t dim, int metric);
/* hnsw.h:781-893 (file written by hnsw.h:748-779): the `hnsw` strategy's un-pruned index */
hso_index *hso_load_hnsw(const char *graph_path, size_t dim, int metric);
void hso_free(hso_index *);
/* patchFromStream, slim.h:2206-2388: rows[label * dim] supply the vectors of new nodes (HSO_PATCH_ROWS) or they
 * sit inline in the stream (HSO_PATCH_INLINE).  0 on success. */
enum { HSO_PATCH_ROWS = 0, HSO_PATCH_INLINE = 2 };
int hso_patch(hso_index *, const void *bytes, size_t len, int mode, const float *rows, size_t n_rows);
void hso_get_info(const hso_index *, hso_info *out);
const char *hso_last_error(void);

/* accessors over the reference memory layout (slim.h:195-210,620-661) */
int hso_node_level(const hso_index *, uint32_t node);
uint64_t hso_node_label(const hso_index *, uint32_t node);
const float *hso_node_vector(const hso_index *, uint32_t node);
/* level-`level` neighbour slice; returns count, *ids points into the blob */
int hso_node_neighbors(const hso_index *, uint32_t node, int level, const uint32_t **ids);

float hso_dist(const float *a, const float *b, size_t dim, int metric, int order, int team);

/* slim.h:2030-2131.  out_labels/out_dists: nq*k, sorted by (dist, internal id)
 * ascending (the reference returns the same k-subset unordered and no distances).
 * Rows with fewer than k results are padded with label 0xFFFFFFFF / dist +inf.
 * n_dist / n_hops (may be NULL): per-query counters with the meaning of
 * metric_distance_computations / metric_hops (slim.h:70-71,371-374,2064-2065):
 * evaluated distances (entry point, upper-layer scans, unvisited base-layer
 * neighbours) and expanded nodes. */
int hso_search(const hso_index *, const float *queries, size_t nq, size_t k, size_t ef,
               int order, int team, int threads, uint32_t *out_labels, float *out_dists,
               uint32_t *n_dist, uint32_t *n_hops);

/* hso_search that also reports, per query, how many results were trimmed from the result heap while tying bit
 * for bit with the entry that stayed behind as the new worst one (n_ties, nq entries): the exact-tie events at
 * the ef boundary after which the reference may still expand an entry that has left its results (slim.h:237,
 * :339-340).  A query with n_ties == 0 has no such event. */
int hso_search_ties(const hso_index *, const float *queries, size_t nq, size_t k, size_t ef,
                    int order, int team, int threads, uint32_t *out_labels, float *out_dists,
                    uint32_t *n_dist, uint32_t *n_hops, uint32_t *n_ties);

/* The ENGINE's traversal restated on the CPU (not the reference's): one pool of <= ef entries in 32 columns with
 * the CUDA kernel's placement, tie and ghost rules (hnsw_slim_b200/csrc/traverse_fp32.cu, traverse_common.cuh
 * RegPool32 / ghost_append), distances in the kernel's association (HSO_ORDER_GPU).  Same outputs as hso_search
 * plus n_ghosts (low 16 bits: ghost entries expanded per query; high 16 bits: unexpanded entries displaced while
 * their tie partner sat in the SAME column, which the ghost rule does not see).  It must equal hso_search to the bit on every query without an
 * exact tie at the ef boundary (hso_search_ties); what differs at the ties is the measured reach of the engine's
 * documented blind spot.  ef <= 256, threshold_level 0, no delete marks; returns -1 otherwise. */
int hso_search_pool(const hso_index *, const float *queries, size_t nq, size_t k, size_t ef, int team, int threads,
                    uint32_t *out_labels, float *out_dists, uint32_t *n_dist, uint32_t *n_hops, uint32_t *n_ghosts);

/* diagnostics: the base-layer expansion order of ONE query under hso_search (which = 0) / hso_search_pool (1) */
size_t hso_debug_trace(const hso_index *, const float *query, size_t k, size_t ef, int which, int team,
                       uint32_t *out_ids, size_t cap);

/* bruteforce.h:106-135 + brute_force_strategy.h:24-36: k labels per query,
 * FARTHEST first (the order the strategy writes to *_groundtruth.ivecs);
 * labels[i] of base row i is i. */
int hso_bruteforce(const float *base, size_t n, size_t dim, int metric, int order, int team,
                   const float *queries, size_t nq, size_t k, int threads,
                   uint32_t *out_labels, float *out_dists);

/* solve_strategy.h:67-103: recall@K of `knn` (nq x K labels, any order) against a
 * ground-truth table gt (nq x gt_k labels, any order, gt_k >= K): GT rows are
 * re-ranked with scalar L2Sqr against base, ties -> smaller id, first K taken. */
double hso_recall(const float *base, size_t dim, const float *queries, size_t nq,
                  const uint32_t *knn, size_t K, const uint32_t *gt, size_t gt_k, int metric);

/* ------------------------------------------------------------------------------------
 * hnsw_slimq (hs_oracle_slimq.c): slimq.h = third_party/hnswlib/hnswalg_slimq.h,
 * rq/ = third_party/rabitqlib/ */
typedef struct hsoq_index hsoq_index;
typedef struct {
  uint64_t n, size_data_per_element, maxM, maxM0, M, ef_construction, dim, padded_dim, num_cluster, ex_bits;
  int32_t maxlevel, threshold_level;
  uint32_t enterpoint;
  int32_t metric_type;
} hsoq_info;

/* slimq.h:1218-1313 */
hsoq_index *hsoq_load(const char *graph_path, size_t dim);
void hsoq_free(hsoq_index *);
void hsoq_get_info(const hsoq_index *, hsoq_info *out);
const char *hsoq_last_error(void);
/* the query-quantiser constant the reference draws at random at load time (slimq.h:1274-1276) */
void hsoq_set_tconst(hsoq_index *, double t_const);
/* record accessors (slimq.h:388-405): cluster id, code words, (f_add, f_rescale, f_error),
 * level-0 neighbour ids; returns the level-0 degree */
int hsoq_node(const hsoq_index *, uint32_t node, uint32_t *cluster, uint64_t *code, float *factors,
              uint32_t *nbr_out, int cap);
/* FhtKacRotator::rotate, rq/utils/rotator.hpp:370-423: q[dim] -> out[padded_dim] */
void hsoq_rotate(const hsoq_index *, const float *q, float *out);
/* SplitSingleQuery ctor, rq/index/query.hpp:127-156: rotated query -> planes[padded_dim/64*4],
 * scal = (delta, vl, k1xsumq); codes_out (may be NULL): the 4-bit code of every dimension */
void hsoq_quantize_query(const hsoq_index *, const float *rotated, uint64_t *planes, float *scal,
                         uint16_t *codes_out);
/* slimq.h:1816-1847: rotate + quantise + distances to the cluster centroids */
void hsoq_prep(const hsoq_index *, const float *q, float *rotated, uint64_t *planes, float *scal,
               float *q2c);
/* get_bin_est, slimq.h:408-440 */
float hsoq_est(const hsoq_index *, uint32_t node, const uint64_t *planes, const float *scal,
               const float *q2c);
/* slimq.h:1810-1924 for a batch.  raw_base: the n x dim rows setDataset() points at (indexed by
 * internal id, slimq.h:747-749).  inj_* (all or none, may be NULL): per-query preparation to use
 * instead of hsoq_prep (planes nq x padded_dim/64*4, scal nq x 3, q2c nq x num_cluster) — this is
 * how the search is pinned to the reference independently of the preparation's rounding.
 * out rows: the k nearest EXPANDED nodes by exact distance, sorted by (dist, internal id);
 * padded with 0xFFFFFFFF / +inf.  Counters per query (may be NULL): estimates computed,
 * nodes expanded (upper scans + base expansions), exact reranks. */
int hsoq_search(const hsoq_index *, const float *raw_base, const float *queries, size_t nq, size_t k,
                size_t ef, int order, int team, int threads, const uint64_t *inj_planes,
                const float *inj_scal, const float *inj_q2c, uint32_t *out_labels, float *out_dists,
                uint32_t *n_est, uint32_t *n_hops, uint32_t *n_rerank);

#ifdef __cplusplus
}
#endif
#endif
