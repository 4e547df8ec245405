// fp32 graph traversal for hnsw_slim on sm_100a: one warp per query.
//
// Replaces, for a whole query batch, HierarchicalNSWSlim<float>::searchKnn
// (slim.h:2030-2131): entry distance, greedy descent over the thinned upper levels
// (slim.h:2040-2078), then the ef-bounded best-first search of searchBaseLayerST<true>
// (slim.h:321-457) with L2Sqr / InnerProductDistance (space_l2.h:25-54, space_ip.h:146-204).
//
// Equivalence with the CPU algorithm (DESIGN.md "Sequential semantics"): the reference
// keeps a min-heap of candidates and an ef-bounded max-heap of results and tightens
// lowerBound after every admitted neighbour.  Here both are ONE pool of at most ef
// (distance,id) keys with an "expanded" bit; a hop scores all unvisited neighbours of the
// closest unexpanded entry at once and admits each one iff it beats the pool's current
// worst entry (which it replaces) — the reference's `top_size < ef || lowerBound > dist`
// followed by the trim to ef.  An entry pushed out of the pool can never be expanded by
// the reference either (its distance exceeds lowerBound from then on), so visited sets,
// distance counts and results coincide except where two distances tie bit-for-bit.
//
// The pool is UNSORTED and column-distributed: lane l owns entries l, l+32, ... and caches
// its column's closest-unexpanded and worst keys in registers; the warp-wide best / worst
// are single REDUX (__reduce_min/max_sync) instructions plus a ballot, so a hop needs no
// binary search and no shifting.  Per-warp shared memory: the pool, an open-addressing
// visited hash (replaces the N-entry tag array of visited_list_pool.h), 32 staging ids and
// — for large dim — the query.  Vector rows are read with 128-bit loads, 8 lanes per row (4 rows per warp
// instruction), fp32 FMA chains per lane, xor-shuffle reduction 4,2,1.
#include <cuda_runtime.h>

#include <cstdint>

#include "traverse_fp32.cuh"

namespace hs {
namespace {

// tuning knobs (see profiles/ for the measurements behind the defaults)
#ifndef HS_TRAVERSE_MIN_CTAS
#define HS_TRAVERSE_MIN_CTAS 6     // 128-thread CTAs per SM the register budget must allow
#endif
#ifndef HS_TRAVERSE_U
#define HS_TRAVERSE_U 1            // x4 rows whose loads are in flight per warp
#endif

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t FLAG = 0x80000000u;             // "expanded" bit inside the id half of a key
constexpr uint64_t KEYMASK = ~(uint64_t)FLAG;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;

// monotone float -> uint32 map (IP distances 1 - <a,b> can be negative)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
  return ((uint64_t)f2ord(d) << 32) | id;
}

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    uint64_t o = __shfl_xor_sync(FULL, v, off);
    v = o < v ? o : v;
  }
  return v;
}

template <int METRIC>
__device__ __forceinline__ float acc4(float acc, const float4 q, const float4 x) {
  if (METRIC == HS_METRIC_L2) {
    float d;
    d = __fsub_rn(q.x, x.x); acc = __fmaf_rn(d, d, acc);
    d = __fsub_rn(q.y, x.y); acc = __fmaf_rn(d, d, acc);
    d = __fsub_rn(q.z, x.z); acc = __fmaf_rn(d, d, acc);
    d = __fsub_rn(q.w, x.w); acc = __fmaf_rn(d, d, acc);
  } else {
    acc = __fmaf_rn(q.x, x.x, acc);
    acc = __fmaf_rn(q.y, x.y, acc);
    acc = __fmaf_rn(q.z, x.z, acc);
    acc = __fmaf_rn(q.w, x.w, acc);
  }
  return acc;
}
template <int METRIC>
__device__ __forceinline__ float finish(float acc) {
  return METRIC == HS_METRIC_IP ? __fsub_rn(1.0f, acc) : acc;
}
__device__ __forceinline__ float team_reduce(float v) {
  v = __fadd_rn(v, __shfl_xor_sync(FULL, v, 4));
  v = __fadd_rn(v, __shfl_xor_sync(FULL, v, 2));
  v = __fadd_rn(v, __shfl_xor_sync(FULL, v, 1));
  return v;
}

// Distances from the query to `count` (<= 32) rows; lane j holds id j, gets d_j back.
// Register-resident query, CPL float4 chunks per lane, U x 4 rows in flight per warp.
template <int CPL, int METRIC, int U>
__device__ __forceinline__ float eval_rows_reg(const float4 *__restrict__ vec, uint32_t row_chunks,
                                               const float4 (&q)[CPL], uint32_t my_id, int count,
                                               int lane) {
  const int team = lane >> 3, t = lane & 7;
  float my_d = 0.f;
  for (int it0 = 0; it0 * 4 < count; it0 += U) {
    float4 x[U][CPL];
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int src = (it0 + u) * 4 + team;
      const uint32_t id = __shfl_sync(FULL, my_id, src & 31);
      act[u] = src < count;
      if (act[u]) {
        const float4 *row = vec + (size_t)id * row_chunks + t;
#pragma unroll
        for (int j = 0; j < CPL; ++j) x[u][j] = __ldg(row + 8 * j);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc = 0.f;
      if (act[u]) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc = acc4<METRIC>(acc, q[j], x[u][j]);
      }
      acc = team_reduce(acc);
      const float v = __shfl_sync(FULL, acc, (lane & 3) * 8);
      if ((lane >> 2) == it0 + u) my_d = finish<METRIC>(v);
    }
  }
  return my_d;
}

// Same with the query in shared memory and a run-time chunk count (large / odd dims).
template <int METRIC>
__device__ __forceinline__ float eval_rows_smem(const float4 *__restrict__ vec, uint32_t row_chunks,
                                                const float4 *qs, uint32_t my_id, int count, int lane) {
  const int team = lane >> 3, t = lane & 7;
  const int cpl = (int)(row_chunks >> 3);
  float my_d = 0.f;
  for (int it = 0; it * 4 < count; ++it) {
    const int src = it * 4 + team;
    const uint32_t id = __shfl_sync(FULL, my_id, src & 31);
    float acc = 0.f;
    if (src < count) {
      const float4 *row = vec + (size_t)id * row_chunks + t;
      const float4 *qq = qs + t;
      int j = 0;
      for (; j + 8 <= cpl; j += 8) {
        float4 x[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) x[jj] = __ldg(row + 8 * (j + jj));
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc = acc4<METRIC>(acc, qq[8 * (j + jj)], x[jj]);
      }
      for (; j < cpl; ++j) acc = acc4<METRIC>(acc, qq[8 * j], __ldg(row + 8 * j));
    }
    acc = team_reduce(acc);
    const float v = __shfl_sync(FULL, acc, (lane & 3) * 8);
    if ((lane >> 2) == it) my_d = finish<METRIC>(v);
  }
  return my_d;
}

// visited set: open addressing, linear probing, 32-bit keys (replaces the per-thread
// uint16 tag array of visited_list_pool.h:10-31).  Returns true if id was already present.
__device__ __forceinline__ bool visited_test_and_set(uint32_t *hash, uint32_t hbits, uint32_t hmask,
                                                     uint32_t id) {
  uint32_t h = (id * 0x9E3779B1u) >> (32 - hbits);
  volatile uint32_t *vh = hash;
  for (;;) {
    const uint32_t v = vh[h];
    if (v == id) return true;
    if (v == EMPTY) {
      const uint32_t old = atomicCAS(hash + h, EMPTY, id);
      if (old == EMPTY) return false;
      if (old == id) return true;
    }
    h = (h + 1) & hmask;
  }
}

__device__ __forceinline__ void hash_clear(uint32_t *hash, uint32_t hsize, int lane) {
  const uint4 e = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
  for (uint32_t i = lane * 4; i < hsize; i += 128) *reinterpret_cast<uint4 *>(hash + i) = e;
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// one instruction pulls a whole vector row (bytes % 16 == 0) into L2; no registers are held
__device__ __forceinline__ void prefetch_row_l2(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
constexpr uint64_t NONE = ~0ull;

// per-lane summary of the pool column this lane owns (entries lane, lane+32, ...)
struct ColStat {
  uint64_t min_un;    // smallest key among unexpanded entries, NONE if there is none
  uint64_t max_all;   // largest key in the column, 0 if the column is empty
  uint32_t min_e, max_e;   // their pool indices
};

__device__ __forceinline__ void col_rescan(const uint64_t *pool, uint32_t size, int lane, ColStat &cs) {
  cs.min_un = NONE;
  cs.max_all = 0;
  cs.min_e = cs.max_e = 0;
  for (uint32_t e = lane; e < size; e += 32) {
    const uint64_t k = pool[e];
    const uint64_t km = k & KEYMASK;
    if (!((uint32_t)k & FLAG) && km < cs.min_un) {
      cs.min_un = km;
      cs.min_e = e;
    }
    if (km >= cs.max_all) {
      cs.max_all = km;
      cs.max_e = e;
    }
  }
}

// lane holding the warp-wide smallest `key` ((dist,id) order; NONE = no entry); -1 if none
__device__ __forceinline__ int warp_argmin_key(uint64_t key) {
  const uint32_t hi = (uint32_t)(key >> 32);
  const uint32_t ghi = __reduce_min_sync(FULL, hi);
  if (ghi == 0xffffffffu) return -1;
  unsigned b = __ballot_sync(FULL, hi == ghi);
  if (b & (b - 1)) {   // equal distances in several lanes: smaller id first
    const uint32_t lo = hi == ghi ? (uint32_t)key : 0xffffffffu;
    const uint32_t glo = __reduce_min_sync(FULL, lo);
    b = __ballot_sync(FULL, hi == ghi && lo == glo);
  }
  return __ffs(b) - 1;
}
// lane holding the warp-wide largest `key` (0 = no entry)
__device__ __forceinline__ int warp_argmax_key(uint64_t key) {
  const uint32_t hi = (uint32_t)(key >> 32);
  const uint32_t ghi = __reduce_max_sync(FULL, hi);
  unsigned b = __ballot_sync(FULL, hi == ghi);
  if (b & (b - 1)) {
    const uint32_t lo = hi == ghi ? (uint32_t)key : 0u;
    const uint32_t glo = __reduce_max_sync(FULL, lo);
    b = __ballot_sync(FULL, hi == ghi && lo == glo);
  }
  return __ffs(b) - 1;
}

// Admit up to 32 scored neighbours (one per lane, `valid`) into the pool: each enters iff the
// pool is not full or it beats the current worst entry, which it then replaces — exactly
// `top_size < ef || lowerBound > dist` + trim (slim.h:403-452), candidate by candidate.
// Returns the ballot of lanes whose candidate entered (it may be displaced again later).
__device__ __forceinline__ unsigned pool_admit(uint64_t *pool, uint32_t &size, uint32_t ef, bool valid,
                                               uint64_t key, ColStat &cs, int lane) {
  unsigned entered = 0;
  const unsigned vmask = __ballot_sync(FULL, valid);
  if (vmask == 0) return 0;
  const uint32_t n_valid = (uint32_t)__popc(vmask);
  unsigned todo = vmask;
  if (size < ef) {
    // room left: the first (ef - size) candidates are appended unconditionally
    const uint32_t room = ef - size;
    const uint32_t rank = (uint32_t)__popc(vmask & ((1u << lane) - 1));
    const bool app = valid && rank < room;
    if (app) pool[size + rank] = key;
    const unsigned am = __ballot_sync(FULL, app);
    entered |= am;
    todo &= ~am;
    size += min(room, n_valid);
    __syncwarp();
    col_rescan(pool, size, lane, cs);
    if (todo == 0) return entered;
  }
  // pool full: pre-filter against the current worst (it only gets smaller), then one by one
  int owner = warp_argmax_key(cs.max_all);
  uint64_t worst = __shfl_sync(FULL, cs.max_all, owner);
  todo &= __ballot_sync(FULL, valid && key < worst);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint64_t ck = __shfl_sync(FULL, key, src);
    if (ck < worst) {
      if (lane == owner) {
        pool[cs.max_e] = ck;
        col_rescan(pool, size, lane, cs);
      }
      entered |= 1u << src;
      if (todo) {
        owner = warp_argmax_key(cs.max_all);
        worst = __shfl_sync(FULL, cs.max_all, owner);
      }
    }
  }
  return entered;
}

template <int CPL, int METRIC>
__global__ void __launch_bounds__(128, HS_TRAVERSE_MIN_CTAS) traverse_kernel(const __grid_constant__ TraverseParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  unsigned char *wbase = smem + (size_t)wid * p.smem_per_warp;
  uint64_t *pool = reinterpret_cast<uint64_t *>(wbase);
  uint32_t *hash = reinterpret_cast<uint32_t *>(wbase + p.off_hash);
  uint32_t *stage_ids = reinterpret_cast<uint32_t *>(wbase + p.off_stage);
  float4 *qs = reinterpret_cast<float4 *>(wbase + p.off_query);

  const uint32_t hbits = p.hash_bits, hsize = 1u << hbits, hmask = hsize - 1;
  const uint32_t hlimit = hsize - hsize / 4;
  const uint32_t ef = p.ef;
  const int t8 = lane & 7;
  constexpr int QN = CPL > 0 ? CPL : 1;

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(p.work_counter, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= p.nq) break;

    // ---- query: lane t of each team owns chunks t, t+8, ... (zero-padded past dim) ----
    const float *qptr = p.queries + (size_t)qi * p.dim;
    float4 q[QN];
    if (CPL > 0) {
#pragma unroll
      for (int j = 0; j < QN; ++j) {
        const uint32_t c = 4u * (uint32_t)(t8 + 8 * j);
        q[j].x = c + 0 < p.dim ? __ldg(qptr + c + 0) : 0.f;
        q[j].y = c + 1 < p.dim ? __ldg(qptr + c + 1) : 0.f;
        q[j].z = c + 2 < p.dim ? __ldg(qptr + c + 2) : 0.f;
        q[j].w = c + 3 < p.dim ? __ldg(qptr + c + 3) : 0.f;
      }
    } else {
      for (uint32_t ch = lane; ch < p.row_chunks; ch += 32) {
        const uint32_t c = 4u * ch;
        float4 v;
        v.x = c + 0 < p.dim ? __ldg(qptr + c + 0) : 0.f;
        v.y = c + 1 < p.dim ? __ldg(qptr + c + 1) : 0.f;
        v.z = c + 2 < p.dim ? __ldg(qptr + c + 2) : 0.f;
        v.w = c + 3 < p.dim ? __ldg(qptr + c + 3) : 0.f;
        qs[ch] = v;
      }
    }
    hash_clear(hash, hsize, lane);
    __syncwarp();

    auto eval = [&](uint32_t my_id, int count) -> float {
      if constexpr (CPL > 0) {
        return eval_rows_reg<QN, METRIC, HS_TRAVERSE_U>(p.vec, p.row_chunks, q, my_id, count, lane);
      } else {
        return eval_rows_smem<METRIC>(p.vec, p.row_chunks, qs, my_id, count, lane);
      }
    };

    uint32_t nd = 0, nh = 0;

    // ---- entry point and greedy descent, slim.h:2033-2078 ----
    uint32_t cur = p.enterpoint;
    float curdist = __shfl_sync(FULL, eval(cur, 1), 0);
    nd = 1;
    for (int level = p.maxlevel; level > p.threshold_level; --level) {
      const uint32_t *ladj = p.upper_adj[level];
      bool changed = true;
      while (changed) {
        changed = false;
        const int slot = __ldg(p.upper_slot + cur);
        if (slot < 0) break;
        const uint32_t *row = ladj + (size_t)slot * p.upper_stride;
        bool any = false;
        for (uint32_t seg = 0; seg < p.upper_stride; seg += 32) {
          const uint32_t id = (seg + lane < p.upper_stride) ? __ldg(row + seg + lane) : kInvalid;
          const unsigned vm = __ballot_sync(FULL, id != kInvalid);
          if (vm == 0) break;                       // rows are packed front to back
          const int count = __popc(vm);
          any = true;
          const float d = eval(id, count);
          nd += (uint32_t)count;
          // sequential scan with strict '<' == first-index argmin
          uint64_t key = lane < count ? (((uint64_t)f2ord(d) << 32) | (uint32_t)lane) : ~0ull;
          key = warp_min_u64(key);
          const float best = ord2f((uint32_t)(key >> 32));
          const uint32_t best_id = __shfl_sync(FULL, id, (int)(key & 31));
          if (best < curdist) {
            curdist = best;
            cur = best_id;
            changed = true;
          }
        }
        if (any) nh++;
      }
    }

    // ---- seed, slim.h:2100-2106 ----
    uint32_t size = 1, hcount = 1;
    ColStat cs;
    if (lane == 0) {
      pool[0] = make_key(curdist, cur);
      visited_test_and_set(hash, hbits, hmask, cur);
    }
    __syncwarp();
    col_rescan(pool, size, lane, cs);

    // ---- base layer, slim.h:321-457 ----
    const uint32_t row_bytes = p.row_chunks * 16u;
    const bool opt_prefetch = p.flags & 1u;
    for (;;) {
      // closest unexpanded entry (the reference pops its candidate min-heap, slim.h:335-354)
      const int bo = warp_argmin_key(cs.min_un);
      if (bo < 0) break;
      const uint32_t node = (uint32_t)__shfl_sync(FULL, (uint32_t)cs.min_un, bo);
      if (lane == bo) {
        pool[cs.min_e] |= (uint64_t)FLAG;
        col_rescan(pool, size, lane, cs);
      }
      const uint32_t *row = p.adj0 + (size_t)node * p.deg0_stride;
      uint32_t id = __ldg(row + lane);

      if (hcount + p.deg0_stride > hlimit) {
        // visited hash nearly full: keep only the pool entries (results are unchanged:
        // a node scored before was rejected or displaced and will be again)
        __syncwarp();
        hash_clear(hash, hsize, lane);
        __syncwarp();
        for (uint32_t t = lane; t < size; t += 32)
          visited_test_and_set(hash, hbits, hmask, (uint32_t)pool[t] & ~FLAG);
        hcount = size;
        __syncwarp();
      }

      bool any = false;
      for (uint32_t seg = 0; seg < p.deg0_stride; seg += 32) {
        if (seg) id = __ldg(row + seg + lane);
        const unsigned vm = __ballot_sync(FULL, id != kInvalid);
        if (vm == 0) break;
        any = true;
        bool fresh = false;
        if (id != kInvalid) fresh = !visited_test_and_set(hash, hbits, hmask, id);
        if (fresh && opt_prefetch) prefetch_row_l2(p.vec + (size_t)id * p.row_chunks, row_bytes);
        const unsigned fm = __ballot_sync(FULL, fresh);
        const int count = __popc(fm);
        if (count == 0) continue;
        hcount += (uint32_t)count;
        if (fresh) stage_ids[__popc(fm & ((1u << lane) - 1))] = id;
        __syncwarp();
        const uint32_t cid = lane < count ? stage_ids[lane] : 0u;
        __syncwarp();
        const float d = eval(cid, count);
        nd += (uint32_t)count;
        const unsigned entered = pool_admit(pool, size, ef, lane < count, make_key(d, cid), cs, lane);
        if ((entered >> lane) & 1u) prefetch_l2(p.adj0 + (size_t)cid * p.deg0_stride);
      }
      if (any) nh++;
    }

    // ---- results: the k closest, ascending (the reference nth_element's the same set,
    //      slim.h:2126-2130): k rounds of warp-wide extract-min over the pool ----
    __syncwarp();
    {
      uint64_t last = 0;   // every key is > 0 (f2ord(+0.0f) has the top bit set)
      for (uint32_t i = 0; i < p.k; ++i) {
        uint64_t mine = NONE;
        for (uint32_t e = lane; e < size; e += 32) {
          const uint64_t km = pool[e] & KEYMASK;
          if (km > last && km < mine) mine = km;
        }
        const int o = warp_argmin_key(mine);
        uint32_t lab = 0xFFFFFFFFu;
        float d = __int_as_float(0x7f800000);
        if (o >= 0) {
          last = __shfl_sync(FULL, mine, o);
          lab = (uint32_t)last;
          d = ord2f((uint32_t)(last >> 32));
        } else {
          last = NONE;
        }
        if (lane == 0) {
          p.out_labels[(size_t)qi * p.k + i] = o >= 0 ? __ldg(p.labels + lab) : 0xFFFFFFFFu;
          if (p.out_dists) p.out_dists[(size_t)qi * p.k + i] = d;
        }
      }
    }
    if (lane == 0) {
      atomicAdd(p.stats + 0, (unsigned long long)nd);
      atomicAdd(p.stats + 1, (unsigned long long)nh);
      if (p.per_query) {
        p.per_query[2 * (size_t)qi + 0] = nd;
        p.per_query[2 * (size_t)qi + 1] = nh;
      }
    }
    __syncwarp();
  }
}

template <int CPL, int METRIC>
int launch_t(const TraverseParams &p, const TraverseLaunch &l, cudaStream_t stream) {
  auto kern = traverse_kernel<CPL, METRIC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem_bytes);
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    return HS_ERR_CUDA;
  }
  kern<<<l.grid, l.warps_per_cta * 32, l.smem_bytes, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("traverse_kernel launch: ") + cudaGetErrorString(e));
    return HS_ERR_CUDA;
  }
  return HS_OK;
}

template <int METRIC>
int launch_m(const TraverseParams &p, const TraverseLaunch &l, cudaStream_t stream) {
  switch (p.row_chunks / kTeam) {
    case 1: return launch_t<1, METRIC>(p, l, stream);
    case 2: return launch_t<2, METRIC>(p, l, stream);
    case 3: return launch_t<3, METRIC>(p, l, stream);   // dim 96  (DEEP / MSTuring)
    case 4: return launch_t<4, METRIC>(p, l, stream);   // dim 128 (SIFT)
    default: return launch_t<0, METRIC>(p, l, stream);  // dim 768 / 960 / anything else
  }
}

template <int CPL, int METRIC>
int occupancy_t(int threads, size_t smem) {
  auto kern = traverse_kernel<CPL, METRIC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) nb = 0;
  return nb;
}
template <int METRIC>
int occupancy_m(uint32_t cpl, int threads, size_t smem) {
  switch (cpl) {
    case 1: return occupancy_t<1, METRIC>(threads, smem);
    case 2: return occupancy_t<2, METRIC>(threads, smem);
    case 3: return occupancy_t<3, METRIC>(threads, smem);
    case 4: return occupancy_t<4, METRIC>(threads, smem);
    default: return occupancy_t<0, METRIC>(threads, smem);
  }
}

inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace

int plan_traverse(TraverseParams &p, int metric, int hash_bits_override, int sm_count, int nq,
                  TraverseLaunch *out) {
  // visited-hash capacity: ~16 slots per ef entry (measured ~12 evaluations per ef entry on
  // 1M x 128), clamped to [1024, 16384]; the kernel resets the table when it
  // passes 75 % so a smaller table only costs repeated evaluations, never correctness.
  uint32_t bits = 10;
  while ((1u << bits) < 16u * p.ef && bits < 14) ++bits;
  if (hash_bits_override > 0) bits = (uint32_t)hash_bits_override;
  if (bits < 8) bits = 8;
  if (bits > 16) bits = 16;
  while ((1u << bits) - (1u << bits) / 4 < p.ef + 2 * p.deg0_stride + 32 && bits < 16) ++bits;
  p.hash_bits = bits;
  const uint32_t list_bytes = align_up(p.ef * 8u, 16);
  const uint32_t hash_bytes = 4u << bits;
  const uint32_t stage_bytes = 32 * 4;
  const bool generic = (p.row_chunks / kTeam) > 4 || (p.row_chunks / kTeam) == 0;
  const uint32_t query_bytes = generic ? p.row_chunks * 16u : 0u;
  p.off_hash = list_bytes;
  p.off_stage = p.off_hash + hash_bytes;
  p.off_query = p.off_stage + stage_bytes;
  p.smem_per_warp = align_up(p.off_query + query_bytes, 16);
  if (p.smem_per_warp > 227u * 1024u) {
    set_error("ef / dim too large for the per-warp shared-memory working set");
    return HS_ERR_UNSUPPORTED;
  }
  int wpc = 4;
  while (wpc > 1 && (size_t)wpc * p.smem_per_warp > 227u * 1024u / 2) wpc >>= 1;
  out->warps_per_cta = wpc;
  out->smem_bytes = (size_t)wpc * p.smem_per_warp;
  const uint32_t cpl = p.row_chunks / kTeam;
  int per_sm = metric == HS_METRIC_IP ? occupancy_m<HS_METRIC_IP>(cpl, wpc * 32, out->smem_bytes)
                                      : occupancy_m<HS_METRIC_L2>(cpl, wpc * 32, out->smem_bytes);
  if (per_sm <= 0) {
    set_error("traverse_kernel does not fit on an SM (cudaOccupancyMaxActiveBlocksPerMultiprocessor)");
    return HS_ERR_CUDA;
  }
  const int resident = sm_count * per_sm;
  const int need = (nq + wpc - 1) / wpc;
  out->grid = need < resident ? (need > 0 ? need : 1) : resident;
  return HS_OK;
}

int launch_traverse(const TraverseParams &p, int metric, const TraverseLaunch &l, cudaStream_t stream) {
  return metric == HS_METRIC_IP ? launch_m<HS_METRIC_IP>(p, l, stream)
                                : launch_m<HS_METRIC_L2>(p, l, stream);
}

}  // namespace hs
