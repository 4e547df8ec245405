// Exact kNN on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// The ground-truth path of the reference (BruteForce::solve, brute_force_strategy.h:15-45 ->
// BruteforceSearch<float>::searchKnn, bruteforce.h:106-135) is the one dense contraction of the
// system: nq x n x dim multiply-adds.  This file computes it as a GEMM on tcgen05 and still
// returns EXACTLY what the fp32 scan kernel (bruteforce.cu) returns, bit for bit:
//
//   1. bf_split_kernel   every fp32 value v is split into hi = v with the low 13 mantissa bits
//                        cleared (a TF32 number) and lo = v - hi (exact); row norms and the
//                        largest row norm are reduced on the way.
//   2. bf_tc_kernel      <q,x> ~= hi.hi + hi.lo + lo.hi as ONE tcgen05.mma.kind::tf32 chain per
//                        128 x 128 (queries x rows) tile, fp32 accumulators in TMEM.  Operands
//                        arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle, 3-stage mbarrier
//                        pipeline, each stage = one 32-column k-block of q_hi, q_lo, x_hi, x_lo); one elected thread issues the MMAs; four epilogue warps read
//                        the accumulators back with tcgen05.ld — one TMEM lane == one query == one
//                        thread.  The operands are augmented ([q,1] . [-2x,|x|^2]) so that the
//                        accumulator IS the approximate score; each thread keeps the kp best
//                        (score,row) pairs in a max-heap (shared memory for kp <= 64), exactly the
//                        structure of the scan kernel.  Warp roles: 0 = TMA producer, 1 = TMEM
//                        allocator + MMA issuer, 2..5 = epilogue; TMEM accumulators are double
//                        buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//   3. bf_finish_kernel  per query: the k-th best approximate score plus twice a rigorous error
//                        bound of the approximation gives a threshold below which every true
//                        top-k row must lie; those candidates are re-scored with the scan kernel's
//                        own fp32 arithmetic (one fma chain in index order) and the k smallest
//                        (distance, label) pairs are emitted.  If a partial heap was too small to
//                        prove that (it is full and its worst entry is under the threshold), the
//                        query is appended to a fallback list ...
//   4. ... which bruteforce.cu's scan kernel then handles (normally empty).
//
// Error bound of step 2: |lo| < 2^-10 |v|, the hardware keeps >= 10 mantissa bits
// of lo, products of TF32 numbers are exact in fp32, the 3*Kpad-term accumulation loses at most
// 2^-22 of the running magnitude per term  =>  |dot~ - dot| <= (3*2^-20 + 3*Kpad*2^-22) |q||x|.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>

#include "bruteforce.cuh"
#include "hs_internal.h"

namespace hs {
namespace {

constexpr int TM = 128;          // queries per CTA  (UMMA M, one TMEM lane each)
constexpr int TN = 128;          // base rows per tile (UMMA N, one TMEM column each)
constexpr int KB = 32;           // fp32 per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32 elements per tcgen05.mma
constexpr int STAGES = 3;
constexpr int ACC = 2;           // TMEM accumulator buffers
// one pipeline stage = one k-block of all four operand tiles: q_hi | q_lo | x_hi | x_lo.  Each tile
// is loaded ONCE per k-block and used by the three MMA groups hi.hi, hi.lo, lo.hi (24.6 MAC per
// shared-memory byte filled instead of 16: the fill traffic from L2 is what bounds this kernel)
constexpr uint32_t A_BYTES = TM * KB * 4, B_BYTES = TN * KB * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int THREADS = 192;
constexpr uint32_t kMaxSmemKp = 32;     // 128 threads x 32 x 8 B = 32 KB next to the 192 KB pipeline

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// per-thread max-heap of keys; storage is either thread-private global memory (stride 1) or a
// shared-memory array interleaved over the 128 epilogue threads (stride 128: conflict-free when
// the threads of a warp touch the same heap level)
struct HeapRef {
  uint64_t *base;
  uint32_t stride;
  __device__ __forceinline__ uint64_t &operator[](uint32_t i) const { return base[(size_t)i * stride]; }
};
__device__ __forceinline__ void heap_push(const HeapRef h, uint32_t &sz, uint64_t key) {
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
__device__ __forceinline__ void heap_replace_top(const HeapRef h, uint32_t sz, uint64_t key) {
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

// one out-of-line copy for the epilogue's (rare) insertions
__device__ __noinline__ void heap_admit(const HeapRef h, uint32_t &sz, uint32_t kp, uint64_t key, float pre,
                                        float &cur) {
  if (sz < kp) {
    heap_push(h, sz, key);
  } else {
    heap_replace_top(h, sz, key);
  }
  if (sz == kp) cur = fminf(pre, ord2f((uint32_t)(h[0] >> 32)));
}

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a pipeline bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B
// apart (stride byte offset), start address in 16-byte units, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // version
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B TF32, both K-major, M x N
constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------- 1. augment + split + norms
// The approximate score is made a pure inner product so that the GEMM delivers it directly:
//   L2:  |x|^2 - 2<q,x> = <[q, 1], [-2x, |x|^2]>        IP:  -<q,x> = <q, -x>
// One warp per row writes the augmented row, zero-padded to kpad columns, as hi + lo
// (hi = low 13 mantissa bits cleared, lo = v - hi exactly), and reduces the largest |row|^2.
enum { SPLIT_QUERY_L2 = 0, SPLIT_BASE_L2 = 1, SPLIT_QUERY_IP = 2, SPLIT_BASE_IP = 3 };

__global__ void bf_split_kernel(const float *__restrict__ src, uint32_t rows, uint32_t dim, uint32_t kpad, int mode,
                                float *__restrict__ hi, float *__restrict__ lo,
                                unsigned int *__restrict__ max_norm2_bits) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float scale = mode == SPLIT_BASE_L2 ? -2.0f : (mode == SPLIT_BASE_IP ? -1.0f : 1.0f);
  float acc = 0.f;
  for (uint32_t c = lane; c < kpad; c += 32) {
    if (c == dim && mode == SPLIT_BASE_L2) continue;         // written below, once |x|^2 is known
    const float raw = c < dim ? __ldg(src + (size_t)row * dim + c) : 0.f;
    acc = __fmaf_rn(raw, raw, acc);
    float v = raw * scale;                                   // exact: a power of two
    if (c == dim && mode == SPLIT_QUERY_L2) v = 1.0f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[(size_t)row * kpad + c] = h;
    lo[(size_t)row * kpad + c] = __fsub_rn(v, h);
  }
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    if (mode == SPLIT_BASE_L2) {                             // the |x|^2 column
      const float h = __uint_as_float(__float_as_uint(acc) & 0xFFFFE000u);
      hi[(size_t)row * kpad + dim] = h;
      lo[(size_t)row * kpad + dim] = __fsub_rn(acc, h);
    }
    if (max_norm2_bits) atomicMax(max_norm2_bits, __float_as_uint(acc));   // acc >= 0: uint order == float order
  }
}

// ---------------------------------------------------------------- 2. tcgen05 GEMM + per-query heaps
struct TcParams {
  uint64_t *cand;            // [splits][nq][kp] max-heaps of (ord(score) << 32 | row), ~0-filled to kp
  const float *qthr;         // optional per-query prefilter: rows with score >= qthr[q] cannot matter
  uint32_t n, nq, kblocks, rows_per_split, kp;   // n = rows to scan (a prefix of the base)
  int metric;
  uint32_t smem_heap;        // 1: heaps live in shared memory during the scan (kp <= kMaxSmemKp)
  // Resident-query mode (few k-blocks, i.e. dim <= ~128; opt-in, see bruteforce_tc_device): the CTA's
  // query tiles q_hi | q_lo of ALL k-blocks are loaded once and stay in shared memory; only the base
  // tiles x_hi | x_lo stream through `xstages` pipeline slots.  Halves the L2 -> shared-memory fill
  // per MMA.
  uint32_t qres, xstages;
};

__global__ void __launch_bounds__(THREADS, 1)
bf_tc_kernel(const __grid_constant__ CUtensorMap tm_qhi, const __grid_constant__ CUtensorMap tm_qlo,
             const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
             const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // data region: streaming mode  STAGES x [q_hi | q_lo | x_hi | x_lo]
  //              resident mode   kblocks x [q_hi | q_lo]  then  xstages x [x_hi | x_lo]
  const uint32_t n_stages = p.qres ? p.xstages : (uint32_t)STAGES;
  const uint32_t q_region = p.qres ? p.kblocks * 2 * A_BYTES : 0u;
  const uint32_t data_bytes = p.qres ? q_region + p.xstages * 2 * B_BYTES : (uint32_t)(STAGES * STAGE_BYTES);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + data_bytes);
  uint64_t *empty = full + STAGES;
  uint64_t *tfull = empty + STAGES;
  uint64_t *tempty = tfull + ACC;
  uint64_t *qfull = tempty + ACC;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qfull + 1);
  uint64_t *smem_heaps = reinterpret_cast<uint64_t *>(smem + data_bytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q0 = blockIdx.x * TM;
  const uint32_t split = blockIdx.y;
  const uint32_t r_begin = split * p.rows_per_split;
  const uint32_t r_end = min(p.n, r_begin + p.rows_per_split);
  const uint32_t n_tiles = r_end > r_begin ? (r_end - r_begin + TN - 1) / TN : 0;
  const uint32_t k_iters = p.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128);
    }
    mbar_init(qfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {     // TMEM: ACC x TN fp32 columns for 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(ACC * TN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      if (p.qres && n_tiles) {                      // the query tiles of every k-block, once
        mbar_expect_tx(qfull, q_region);
        for (uint32_t kb = 0; kb < k_iters; ++kb) {
          tma_load_2d(smem + kb * 2 * A_BYTES, &tm_qhi, (int)(kb * KB), (int)q0, qfull);
          tma_load_2d(smem + kb * 2 * A_BYTES + A_BYTES, &tm_qlo, (int)(kb * KB), (int)q0, qfull);
        }
      }
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const int row0 = (int)(r_begin + t * TN);
        for (uint32_t kb = 0; kb < k_iters; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (p.qres) {
            uint8_t *st = smem + q_region + stage * 2 * B_BYTES;
            mbar_expect_tx(&full[stage], 2 * B_BYTES);
            tma_load_2d(st, &tm_xhi, (int)(kb * KB), row0, &full[stage]);
            tma_load_2d(st + B_BYTES, &tm_xlo, (int)(kb * KB), row0, &full[stage]);
          } else {
            uint8_t *st = smem + stage * STAGE_BYTES;
            mbar_expect_tx(&full[stage], STAGE_BYTES);
            tma_load_2d(st, &tm_qhi, (int)(kb * KB), (int)q0, &full[stage]);
            tma_load_2d(st + A_BYTES, &tm_qlo, (int)(kb * KB), (int)q0, &full[stage]);
            tma_load_2d(st + 2 * A_BYTES, &tm_xhi, (int)(kb * KB), row0, &full[stage]);
            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tm_xlo, (int)(kb * KB), row0, &full[stage]);
          }
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
      uint32_t stage = 0, phase = 0;
      if (p.qres && n_tiles) {
        mbar_wait(qfull, 0);
        tc_fence_after();
      }
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + acc * TN;
        for (uint32_t it = 0; it < k_iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a = p.qres ? smem_u32(smem + it * 2 * A_BYTES) : smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b = p.qres ? smem_u32(smem + q_region + stage * 2 * B_BYTES) : a + 2 * A_BYTES;
          const uint64_t qhi = umma_desc_sw128(a), qlo = umma_desc_sw128(a + A_BYTES);
          const uint64_t xhi = umma_desc_sw128(b), xlo = umma_desc_sw128(b + B_BYTES);
#pragma unroll
          for (int g = 0; g < 3; ++g) {               // hi.hi, hi.lo, lo.hi
            const uint64_t ad = g == 2 ? qlo : qhi, bd = g == 1 ? xlo : xhi;
#pragma unroll
            for (int k = 0; k < KB / UMMA_K; ++k)     // +32 bytes (2 x 16 B) along K inside the swizzle row
              tc_mma_tf32(d, ad + 2 * k, bd + 2 * k, idesc, (it | g | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty[stage]);                   // frees the smem slot when these MMAs retire
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull[acc]);                       // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TMEM lane == query =====
    const uint32_t quad = warp & 3;                   // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const uint32_t my_q = q0 + quad * 32 + lane;
    const bool live = my_q < p.nq;
    uint64_t *gheap = p.cand + ((size_t)split * p.nq + (live ? my_q : 0)) * p.kp;
    const HeapRef heap = p.smem_heap ? HeapRef{smem_heaps + (quad * 32 + lane), 128u} : HeapRef{gheap, 1u};
    uint32_t hsz = 0;
    // admission threshold on the approximate score: the sampled prefilter (if any) until the
    // heap is full, then also the heap's worst entry.  Rows that tie with it are dropped:
    // bf_finish_kernel treats a full heap whose worst entry is under ITS threshold as
    // insufficient, so exactness never depends on a dropped row.
    const float pre = (p.qthr && live) ? __ldg(p.qthr + my_q) : __int_as_float(0x7f800000);
    float cur = pre;
    for (uint32_t t = 0; t < n_tiles; ++t) {
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t row0 = r_begin + t * TN;
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((quad * 32u) << 16) + acc * TN + c0, v);
        if (!live) continue;
        // the accumulator IS the approximate score.  Common case: none of the 32 scores passes the
        // threshold — 32 predicated compares into a mask and ONE branch.  The rare insertions go
        // through a single out-of-line copy of the heap code (inlining it per element made the
        // loop body ~50 KB of SASS: ncu showed the warps starved for instructions).
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) < cur) ? (1u << j) : 0u;
        if (mask) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (mask & (1u << j)) {
              const float s = __uint_as_float(v[j]);
              const uint32_t row = row0 + c0 + j;
              if (s < cur && row < r_end)             // cur may have tightened; rows past the edge are TMA zero fill
                heap_admit(heap, hsz, p.kp, ((uint64_t)f2ord(s) << 32) | row, pre, cur);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
    if (live)
      for (uint32_t i = 0; i < p.kp; ++i) gheap[i] = i < hsz ? heap[i] : ~0ull;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACC * TN));
  }
}

// ---------------------------------------------------------------- 3. exact re-scoring of the candidates
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

// rigorous bound on |approximate score - exact score| (see the file comment), doubled
__device__ __forceinline__ float score_eps(float qn, float xmax, uint32_t kpad, int metric) {
  const float delta = 3.0f * 9.5367431640625e-07f /* 2^-20 */ + 3.0f * (float)kpad * 2.384185791015625e-07f /* 2^-22 */;
  float eps;
  if (metric == HS_METRIC_IP) {
    eps = delta * qn * xmax + 1.2e-7f * (1.0f + qn * xmax);                    // + rounding of 1 - <q,x>
  } else {
    const float sum = qn + xmax;
    // |[q,1]| |[-2x,|x|^2]| terms: 2|q||x| + |x|^2; + split of the |x|^2 column and scan-kernel rounding
    eps = delta * (2.0f * qn * xmax + xmax * xmax) + (float)kpad * 1.2e-7f * (xmax * xmax + sum * sum);
  }
  return 2.0f * eps;                                                            // safety factor
}

// Prefilter thresholds from a SAMPLE of the base (its first rows): the k-th best approximate score
// inside the sample is >= the k-th best over all rows, so a row scoring above it (+ 2 eps) can
// neither be in the exact top-k nor move the finish kernel's threshold.  One warp per query;
// cand holds one heap of kp entries per query (a single split).
__global__ void bf_threshold_kernel(const uint64_t *__restrict__ cand, const float *__restrict__ queries,
                                    const unsigned int *__restrict__ max_norm2_bits, uint32_t nq, uint32_t dim,
                                    uint32_t kpad, uint32_t k, uint32_t kp, int metric, float *__restrict__ qthr) {
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const float *qv = queries + (size_t)q * dim;
  float a = 0.f;
  for (uint32_t d = lane; d < dim; d += 32) a = __fmaf_rn(qv[d], qv[d], a);
  for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
  // k-th smallest key of the heap = the key with exactly k-1 smaller ones (keys are distinct: row ids)
  const uint64_t *h = cand + (size_t)q * kp;
  uint64_t kth = ~0ull;
  for (uint32_t i = lane; i < kp; i += 32) {
    const uint64_t e = h[i];
    if (e == ~0ull) continue;
    uint32_t smaller = 0;
    for (uint32_t j = 0; j < kp; ++j) smaller += h[j] < e;
    if (smaller == k - 1) kth = e;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    const uint64_t o = __shfl_xor_sync(0xffffffffu, kth, off);
    kth = o < kth ? o : kth;
  }
  if (lane == 0) {
    float t = __int_as_float(0x7f800000);          // fewer than k rows in the sample: no filter
    if (kth != ~0ull)
      t = ord2f((uint32_t)(kth >> 32)) + 2.0f * score_eps(sqrtf(a), sqrtf(__uint_as_float(*max_norm2_bits)), kpad, metric);
    qthr[q] = t;
  }
}

struct FinishParams {
  const float *base, *queries;
  const uint64_t *cand;
  const unsigned int *max_norm2_bits;
  uint32_t n, nq, dim, kpad, k, kp, splits, P;
  int metric;
  uint32_t *out_labels;
  float *out_dists;
  uint32_t *fallback_list;
  unsigned int *fallback_count;
};

__global__ void bf_finish_kernel(const FinishParams p) {
  extern __shared__ __align__(16) uint64_t keys[];
  __shared__ float s_qn2;
  __shared__ int s_overflow;
  __shared__ uint32_t s_m;
  const uint32_t q = blockIdx.x;
  const float *qv = p.queries + (size_t)q * p.dim;
  const uint32_t total = p.splits * p.kp;
  if (threadIdx.x == 0) {
    s_overflow = 0;
    s_m = 0;
  }
  if (threadIdx.x < 32) {
    float a = 0.f;
    for (uint32_t d = threadIdx.x; d < p.dim; d += 32) a = __fmaf_rn(qv[d], qv[d], a);
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (threadIdx.x == 0) s_qn2 = a;
  }
  for (uint32_t i = threadIdx.x; i < p.P; i += blockDim.x) {
    uint64_t v = ~0ull;
    if (i < total) v = p.cand[((size_t)(i / p.kp) * p.nq + q) * p.kp + (i % p.kp)];
    keys[i] = v;
  }
  bitonic_sort(keys, p.P);

  // threshold: every row of the exact top-k has an approximate score <= sigma_k + 2 eps
  const float eps = score_eps(sqrtf(s_qn2), sqrtf(__uint_as_float(*p.max_norm2_bits)), p.kpad, p.metric);
  const uint64_t kth = keys[min(p.k, total) - 1];
  uint64_t thr_key = ~0ull;                                                     // fewer than k rows: keep all
  if (kth != ~0ull) {
    const float tau = ord2f((uint32_t)(kth >> 32)) + 2.0f * eps;
    thr_key = ((uint64_t)f2ord(tau) << 32) | 0xFFFFFFFFull;
  }
  // a partial heap that is full and whose worst entry is under the threshold may have dropped a
  // row that matters: cannot prove sufficiency
  for (uint32_t s = threadIdx.x; s < p.splits; s += blockDim.x) {
    const uint64_t *h = p.cand + ((size_t)s * p.nq + q) * p.kp;
    if (h[p.kp - 1] != ~0ull && h[0] <= thr_key) s_overflow = 1;
  }
  __syncthreads();
  if (s_overflow) {
    if (threadIdx.x == 0) p.fallback_list[atomicAdd(p.fallback_count, 1u)] = q;
    return;
  }
  // candidates = sorted prefix under the threshold; re-score with the scan kernel's arithmetic
  for (uint32_t i = threadIdx.x; i < p.P; i += blockDim.x) {
    const uint64_t e = keys[i];
    uint64_t out = ~0ull;
    if (e != ~0ull && e <= thr_key) {
      const uint32_t row = (uint32_t)e;
      const float *x = p.base + (size_t)row * p.dim;
      float acc = 0.f;
      for (uint32_t d = 0; d < p.dim; ++d) {
        if (p.metric == HS_METRIC_L2) {
          const float t = __fsub_rn(__ldg(qv + d), __ldg(x + d));
          acc = __fmaf_rn(t, t, acc);
        } else {
          acc = __fmaf_rn(__ldg(qv + d), __ldg(x + d), acc);
        }
      }
      const float dist = p.metric == HS_METRIC_IP ? __fsub_rn(1.0f, acc) : acc;
      out = ((uint64_t)f2ord(dist) << 32) | row;
      atomicAdd(&s_m, 1u);
    }
    keys[i] = out;       // same slot: entries past the prefix become padding
  }
  __syncthreads();
  bitonic_sort(keys, p.P);
  for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x) {
    const uint64_t e = keys[i];
    const bool ok = e != ~0ull;
    p.out_labels[(size_t)q * p.k + i] = ok ? (uint32_t)e : 0xFFFFFFFFu;
    if (p.out_dists) p.out_dists[(size_t)q * p.k + i] = ok ? ord2f((uint32_t)(e >> 32)) : __int_as_float(0x7f800000);
  }
}

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rows x kpad fp32 matrix, box = 32 floats x 128 rows, 128-byte swizzle, zero fill past the edge
int make_map(CUtensorMap *map, const float *ptr, uint64_t rows, uint64_t kpad) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return HS_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {kpad, rows};
  const cuuint64_t gstride[1] = {kpad * sizeof(float)};
  const cuuint32_t box[2] = {KB, TM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return HS_ERR_CUDA;
  }
  return HS_OK;
}

#define TC_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__));         \
      return HS_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

}  // namespace

bool bruteforce_tc_applicable(size_t n, size_t dim, size_t nq, size_t k) {
  if (const char *e = std::getenv("HS_BF_TC")) {
    if (e[0] == '0') return false;
    if (e[0] == '1') return n >= TN && k <= 512 && dim <= 8192;
  }
  // below this the scan kernel is launch-latency bound anyway
  return n >= 32768 && nq >= 64 && k <= 512 && dim <= 8192;
}

int bruteforce_tc_device(const float *d_base, size_t n, size_t dim, const float *d_queries, size_t nq, size_t k,
                         int metric, uint32_t *d_labels, float *d_dists, cudaStream_t stream,
                         uint32_t **fallback_list, unsigned int **fallback_count, void **scratch_to_free) {
  int dev = 0, sms = 148;
  TC_CUDA(cudaGetDevice(&dev));
  TC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t aug = metric == HS_METRIC_IP ? 0u : 1u;          // the |x|^2 / 1 column
  const uint32_t kpad = (uint32_t)((dim + aug + KB - 1) / KB * KB);
  const uint32_t kp = (uint32_t)std::max<size_t>(32, 2 * k);
  const uint32_t q_tiles = (uint32_t)((nq + TM - 1) / TM);
  // one CTA per SM (TMEM + >= 128 KB of shared memory): pick the number of base splits that
  // wastes the least of the last wave, preferring more (shorter) CTAs
  const size_t min_rows = std::max<size_t>(8 * TN, 8 * kp);
  uint32_t max_splits = (uint32_t)std::max<size_t>(1, std::min<size_t>(n / min_rows, 16384 / kp));
  max_splits = std::min<uint32_t>(max_splits, std::max<uint32_t>(1, (uint32_t)((8 * sms + q_tiles - 1) / q_tiles)));
  uint32_t splits = 1;
  double best_eff = 0.0;
  for (uint32_t s = 1; s <= max_splits; ++s) {
    const double ctas = (double)q_tiles * s;
    const double eff = ctas / (std::ceil(ctas / sms) * sms);
    if (eff >= best_eff - 0.02) {
      if (eff > best_eff) best_eff = eff;
      splits = s;
    }
  }
  uint32_t rows_per_split = (uint32_t)((n + splits - 1) / splits);
  rows_per_split = (rows_per_split + TN - 1) / TN * TN;
  splits = (uint32_t)((n + rows_per_split - 1) / rows_per_split);

  // one scratch allocation: x_hi, x_lo, q_hi, q_lo, |x|^2, heaps, fallback list + counters
  const size_t xb = n * (size_t)kpad * 4, qb = nq * (size_t)kpad * 4;
  const size_t off_xhi = 0, off_xlo = off_xhi + xb, off_qhi = off_xlo + xb, off_qlo = off_qhi + qb;
  const size_t off_cand = (off_qlo + qb + 1023) / 1024 * 1024;
  const size_t off_fb = off_cand + (size_t)splits * nq * kp * 8, off_thr = off_fb + nq * 4, off_cnt = off_thr + nq * 4;
  const size_t bytes = off_cnt + 64;
  uint8_t *scratch = nullptr;
  TC_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), bytes, stream));
  *scratch_to_free = scratch;
  float *x_hi = reinterpret_cast<float *>(scratch + off_xhi), *x_lo = reinterpret_cast<float *>(scratch + off_xlo);
  float *q_hi = reinterpret_cast<float *>(scratch + off_qhi), *q_lo = reinterpret_cast<float *>(scratch + off_qlo);
  uint64_t *cand = reinterpret_cast<uint64_t *>(scratch + off_cand);
  uint32_t *fb = reinterpret_cast<uint32_t *>(scratch + off_fb);
  float *qthr = reinterpret_cast<float *>(scratch + off_thr);
  unsigned int *cnt = reinterpret_cast<unsigned int *>(scratch + off_cnt);     // [0] fallback count, [1] max |x|^2 bits
  TC_CUDA(cudaMemsetAsync(cnt, 0, 64, stream));
  *fallback_list = fb;
  *fallback_count = cnt;

  const bool ip = metric == HS_METRIC_IP;
  bf_split_kernel<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, stream>>>(
      d_base, (uint32_t)n, (uint32_t)dim, kpad, ip ? SPLIT_BASE_IP : SPLIT_BASE_L2, x_hi, x_lo, cnt + 1);
  bf_split_kernel<<<(uint32_t)((nq * 32 + 255) / 256), 256, 0, stream>>>(
      d_queries, (uint32_t)nq, (uint32_t)dim, kpad, ip ? SPLIT_QUERY_IP : SPLIT_QUERY_L2, q_hi, q_lo, nullptr);
  TC_CUDA(cudaGetLastError());

  CUtensorMap m_qhi, m_qlo, m_xhi, m_xlo;
  int rc;
  if ((rc = make_map(&m_qhi, q_hi, nq, kpad)) != HS_OK || (rc = make_map(&m_qlo, q_lo, nq, kpad)) != HS_OK ||
      (rc = make_map(&m_xhi, x_hi, n, kpad)) != HS_OK || (rc = make_map(&m_xlo, x_lo, n, kpad)) != HS_OK)
    return rc;

  // resident-query mode when the query tiles of all k-blocks plus at least two base-tile slots fit
  // (dim + 1 <= 160: SIFT / DEEP / MSTuring shapes); the heaps then live in global memory
  const uint32_t kblocks = kpad / KB;
  const size_t smem_budget = 227 * 1024 - 1024 - 256;
  uint32_t qres = 0, xstages = 0;
  if ((size_t)kblocks * 2 * A_BYTES + 2 * 2 * B_BYTES <= smem_budget) {
    qres = 1;
    xstages = (uint32_t)std::min<size_t>(STAGES, (smem_budget - (size_t)kblocks * 2 * A_BYTES) / (2 * B_BYTES));
  }
  // Measured on 1M x 128, 10k queries (main launch): streaming 16.5 ms, resident 19.0 ms — with the
  // query tiles resident only two base-tile slots fit and the heaps move to global memory, which
  // costs more than the halved fill saves.  Resident mode therefore stays opt-in (HS_BF_QRES=1).
  {
    const char *e = std::getenv("HS_BF_QRES");
    if (!(e && e[0] == '1')) qres = 0;
  }
  const size_t data_bytes = qres ? (size_t)kblocks * 2 * A_BYTES + (size_t)xstages * 2 * B_BYTES : (size_t)STAGES * STAGE_BYTES;
  size_t smem_heap_bytes = kp <= kMaxSmemKp ? (size_t)128 * kp * 8 : 0;
  if (data_bytes + 1024 + 256 + smem_heap_bytes > 227 * 1024) smem_heap_bytes = 0;
  const size_t smem = data_bytes + 1024 /* alignment slack */ + 256 /* barriers */ + smem_heap_bytes;
  TC_CUDA(cudaFuncSetAttribute(bf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auto launch_tc = [&](uint32_t rows, uint32_t n_splits, uint32_t rps, const float *thr) {
    TcParams tp{};
    tp.cand = cand;
    tp.qthr = thr;
    tp.n = rows;
    tp.nq = (uint32_t)nq;
    tp.kblocks = kpad / KB;
    tp.rows_per_split = rps;
    tp.kp = kp;
    tp.metric = metric;
    tp.smem_heap = smem_heap_bytes ? 1u : 0u;
    tp.qres = qres;
    tp.xstages = xstages;
    bf_tc_kernel<<<dim3(q_tiles, n_splits), THREADS, smem, stream>>>(m_qhi, m_qlo, m_xhi, m_xlo, tp);
  };
  // phase A: a sample (the first rows) yields per-query prefilter thresholds, which keep the
  // heaps of phase B out of their insert-heavy warm-up (lanes of a warp insert at different rows)
  const float *thr = nullptr;
  const size_t sample = std::max<size_t>(16384, 512 * k);
  if (n >= 8 * sample) {
    const uint32_t srows = (uint32_t)(sample / TN * TN);
    launch_tc(srows, 1, srows, nullptr);
    bf_threshold_kernel<<<(uint32_t)((nq * 32 + 255) / 256), 256, 0, stream>>>(
        cand, d_queries, cnt + 1, (uint32_t)nq, (uint32_t)dim, kpad, (uint32_t)k, kp, metric, qthr);
    thr = qthr;
  }
  // phase B: everything
  launch_tc((uint32_t)n, splits, rows_per_split, thr);
  TC_CUDA(cudaGetLastError());

  FinishParams fp{};
  fp.base = d_base;
  fp.queries = d_queries;
  fp.cand = cand;
  fp.max_norm2_bits = cnt + 1;
  fp.n = (uint32_t)n;
  fp.nq = (uint32_t)nq;
  fp.dim = (uint32_t)dim;
  fp.kpad = kpad;
  fp.k = (uint32_t)k;
  fp.kp = kp;
  fp.splits = splits;
  fp.P = next_pow2(std::max<uint32_t>(2, splits * kp));
  fp.metric = metric;
  fp.out_labels = d_labels;
  fp.out_dists = d_dists;
  fp.fallback_list = fb;
  fp.fallback_count = cnt;
  TC_CUDA(cudaFuncSetAttribute(bf_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fp.P * 8)));
  bf_finish_kernel<<<(uint32_t)nq, 256, fp.P * 8, stream>>>(fp);
  TC_CUDA(cudaGetLastError());
  return HS_OK;
}

}  // namespace hs
