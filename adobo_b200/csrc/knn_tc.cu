// K8 (tensor-core shortlist)  knn_scores_tc -- the distance "GEMM" of the kNN search on tcgen05.
//
// Replaces the candidate generation of adobo/clustering.py:47-50 (sklearn ball tree).  For a tile
// of 128 queries and every tile of 128 references it forms, on the 5th-generation tensor cores,
//
//        s(q, r) = q . r - 1/2 ||r||^2          (largest s  <=>  smallest ||q - r||^2)
//
// with fp32 accuracy out of TF32 MMAs: every centred fp32 coordinate is split into two TF32
// numbers x = hi + lo, the operands are stored as 2 Kp columns each,
//        A' = [ q_hi | q_lo ],   B' = [ r_hi | r_lo ],
// and three products are accumulated per k-block: q_hi.r_hi and q_lo.r_hi against ONE staged block of r_hi, then
// q_hi.r_lo (hi.hi + lo.hi + hi.lo).  The reference side is what streams through L2 -- every CTA reads all of B' for
// its 128 queries, 6.5 TB/s at 1 M points when r_hi was staged twice -- so a block of B' is loaded once per use
// pair: two thirds of the traffic for the same MMAs.
// Two spare columns of the hi segments carry the norm:  A' = 1, 1 ;  B' = -rn_hi/2, -rn_lo/2.
// The epilogue therefore does NO arithmetic: a score is a candidate iff it exceeds the thread's
// current 32nd/64th best, one compare per element.
//
// Structure (one CTA = one 128-query tile, persistent over all reference tiles, 6 warps):
//   warp 0     TMA producer: A' (resident, K'/32 blocks of 128 x 32 tf32 = 16 KB, 128B swizzle), then a
//              ring of B' blocks, cp.async.bulk.tensor.2d -> mbarrier complete_tx
//   warp 1     MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M=128, N=128, K=8 per instruction,
//              accumulators in TMEM (2 x 128 columns, double buffered), tcgen05.commit -> mbarriers
//   warps 2-5  epilogue: tcgen05.ld 32x32b.x16 (thread = query row).  A chunk of 16 scores whose maximum does not
//              exceed the row's threshold costs a max tree and one compare; survivors are appended to a 32-slot pool
//              per row in shared memory.  A pool is merged into the row's sorted shortlist only when the next chunk
//              could overflow it: the warp sorts the pool across its lanes (bitonic network) and merges it with the
//              list by one reversed min step + a bitonic merge -- ~20 batched merges per row at 68 k points where
//              an insertion per survivor took 250 (and was what the kernel spent its time on: tensor pipe 21 % busy)
// The shortlist is re-ranked in fp64 and certified by knn_rerank (knn.cu), exactly as before.
#include <cuda.h>

#include <cfloat>
#include <climits>

#include "common.cuh"
#include "knn_common.cuh"

namespace adobo {

constexpr int TC_M = 128, TC_N = 128, TC_KB = 32;   // tile rows, tile refs, tf32 per k-block (128 B)
constexpr int TC_BLOCK_BYTES = TC_M * TC_KB * 4;     // 16 KB
constexpr int TC_THREADS = 192;

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// The producer and the MMA issuer are single threads that share a scheduler with an epilogue warp: a bare try_wait
// loop re-issues every ~9 cycles (ncu: 3.7e8 try_waits in an 11 ms launch, the XU pipe 87 % busy with them), so
// they back off between polls
#ifndef ADOBO_KNN_SPIN_NS
#define ADOBO_KNN_SPIN_NS 32
#endif
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (ADOBO_KNN_SPIN_NS > 0) __nanosleep(ADOBO_KNN_SPIN_NS);
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (CUTLASS elect_one_sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// A from tensor memory (lane = row of A, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile [128 rows][32 tf32] written by TMA with 128-byte swizzle, 1024-byte aligned:
// start address >> 4, LBO (unused for swizzled K-major) = 1, SBO = 1024 B (8 rows x 128 B), version 1,
// layout SWIZZLE_128B (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- operand preparation ------------------------------------------------------------------------
// x: n x p fp64; writes A' = [q_hi | q_lo] and B' = [r_hi | r_lo] (npad x 2 Kp fp32), fp32 norms of the centred points
// One warp per point: the lanes stride the columns, so every row of A' / B' is written with coalesced stores (a thread
// per point wrote 32 different rows per store instruction: 0.36 ms at 68 k points for 100 MB).  The fp32 norm is
// accumulated in coordinate order by every lane (broadcast loads), the same bits as before.
__global__ void __launch_bounds__(256) knn_tc_prep(const double* __restrict__ x, i64 n, i64 npad, int p, int Kp,
                                                   const double* __restrict__ ctr, float* __restrict__ Aop,
                                                   float* __restrict__ Bop, float* __restrict__ norms,
                                                   unsigned* __restrict__ maxnorm_bits) {
  const int lane = threadIdx.x & 31;
  const i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= npad) return;
  const int Kt = 2 * Kp;
  float* a = Aop + (size_t)i * Kt;
  float* b = Bop + (size_t)i * Kt;
  if (i < n) {
    const double* xr = x + i * p;
    float nrm = 0.f;
    for (int d = 0; d < p; ++d) {
      const float v = (float)(xr[d] - ctr[d]);
      nrm = fmaf(v, v, nrm);
    }
    const float hn = -0.5f * nrm, hh = tf32_round(hn), hl = tf32_round(hn - hh);
    for (int d = lane; d < Kt; d += 32) {
      const int dd = d < Kp ? d : d - Kp;
      float av = 0.f, bv = 0.f;
      if (dd < p) {
        const float v = (float)(xr[dd] - ctr[dd]);
        const float hi = tf32_round(v);
        av = bv = (d < Kp) ? hi : tf32_round(v - hi);
      } else if (d == p) {
        av = 1.f, bv = hh;
      } else if (d == p + 1) {
        av = 1.f, bv = hl;
      }
      a[d] = av;
      b[d] = bv;
    }
    if (lane == 0) {
      norms[i] = nrm;
      atomicMax(maxnorm_bits, __float_as_uint(nrm));
    }
  } else {
    for (int d = lane; d < Kt; d += 32) {
      a[d] = (d == p) ? 1.f : 0.f;
      b[d] = (d == p) ? -1e30f : 0.f;  // padding rows score -1e30: never a candidate
    }
    if (lane == 0) norms[i] = INFINITY;
  }
}

// ---- the tensor-core kernel ------------------------------------------------------------------------
constexpr int TC_ACC = 3;                            // most accumulators in TMEM (3 x 128 columns when A' leaves room, else 2):
                                                     // the epilogue's time per tile varies (pool merges)
constexpr int TC_CHUNK = 16;                         // columns per append / flush-check step
constexpr int TC_POOL = 32;                          // pool slots per query row; flushed when a chunk might not fit
constexpr int TC_POOL_STRIDE = TC_M * 8 + 8;         // bytes between slots: [slot][row] of {score, ref}, padded so that
                                                     // both the appends (same slot, 32 rows) and the flush (32 slots,
                                                     // one row) are free of bank conflicts
constexpr int TC_POOL_BYTES = TC_POOL * TC_POOL_STRIDE;

__device__ __forceinline__ bool better(float ka, int ia, float kb, int ib) { return ka > kb || (ka == kb && ia < ib); }
// compare-exchange with the lane at distance d: this lane keeps the better candidate iff keep_better
__device__ __forceinline__ void cmpx_lane(float& k, int& i, int d, bool keep_better) {
  const float ok = __shfl_xor_sync(0xffffffffu, k, d);
  const int oi = __shfl_xor_sync(0xffffffffu, i, d);
  if (better(k, i, ok, oi) != keep_better) {
    k = ok;
    i = oi;
  }
}

// E: shortlist length KP = 32 E.  KS = ceil(p / 8): the k-slices of 8 columns that hold q_lo / r_lo data; the hi
// segments (p + 2 columns) have ks_hi = KS or KS + 1.  KS is a template parameter so that the MMA schedule of a tile
// is straight-line code (see the MMA issuer below).
template <int E, int KS>
__global__ void __launch_bounds__(TC_THREADS, 1)
    knn_scores_tc(const float* __restrict__ Aop, const __grid_constant__ CUtensorMap mapB,
                  int KB /* k-blocks per segment: A' and B' hold 2 KB each */, int ks_hi, int nstages, int lists_in_smem, int dbg, i64 ntiles, i64 qtile0, i64 q_begin, i64 q_end,
                  float* __restrict__ out_key, int* __restrict__ out_idx, float* __restrict__ gl_key,
                  int* __restrict__ gl_idx) {
  constexpr int KP = 32 * E;
  constexpr int ks_lo = KS;
  constexpr int NB_LO = (KS + 3) / 4;            // 32-column blocks of a lo segment that hold data
  const int nb_hi = (ks_hi + 3) / 4;             // ... of a hi segment (NB_LO or NB_LO + 1)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand blocks first, then shortlists / pools / barriers
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = base;                                   // nstages blocks of B'
  unsigned char* pool = sB + (size_t)nstages * TC_BLOCK_BYTES;          // [TC_POOL][128] {score, ref}, padded stride
  uint64_t* bars = reinterpret_cast<uint64_t*>(pool + TC_POOL_BYTES);
  uint64_t* full = bars;                 // [nstages]
  uint64_t* empty = bars + 8;            // [nstages]
  uint64_t* tmem_full = bars + 17;       // [TC_ACC]
  uint64_t* tmem_empty = bars + 21;      // [TC_ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  float* sm_lkey = reinterpret_cast<float*>(bars + 28);       // [128][KP] when lists_in_smem
  int* sm_lidx = reinterpret_cast<int*>(sm_lkey + TC_M * KP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 qtile = qtile0 + blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < TC_ACC; ++b) {
      mbar_init(tmem_full + b, 1);
      mbar_init(tmem_empty + b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM): nacc accumulators of 128, then A'
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // A' = [q_hi | q_lo] lives in TENSOR MEMORY (lane = query row, one tf32 per column), not in shared memory: with
  // both operands in shared memory an M = N = 128, K = 8 MMA reads 8 KB for 64 cycles of math and the operand fetch
  // (~57 B/clk measured) paces the tensor pipe at 143 cycles per MMA.  Only the k-slices that hold data are stored:
  // ks_hi slices of q_hi (p + 2 columns), then ks_lo slices of q_lo.
  const int a_cols = (ks_hi + ks_lo) * 8;
  const int nacc = (512 - a_cols) / TC_N >= 3 ? 3 : 2;
  const uint32_t tmem_a = tmem_base + (uint32_t)nacc * TC_N;
  if (warp >= 2) {
    const int slab = warp & 3;
    const float* arow = Aop + (size_t)(qtile * TC_M + slab * 32 + lane) * (size_t)(2 * KB * TC_KB);
    for (int sl = 0; sl < ks_hi + ks_lo; ++sl) {
      const float* src = sl < ks_hi ? arow + sl * 8 : arow + KB * TC_KB + (sl - ks_hi) * 8;
      const float4 x0 = *reinterpret_cast<const float4*>(src), x1 = *reinterpret_cast<const float4*>(src + 4);
      const uint32_t taddr = tmem_a + ((uint32_t)(slab * 32) << 16) + (uint32_t)sl * 8;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                   "r"(__float_as_uint(x0.x)), "r"(__float_as_uint(x0.y)), "r"(__float_as_uint(x0.z)),
                   "r"(__float_as_uint(x0.w)), "r"(__float_as_uint(x1.x)), "r"(__float_as_uint(x1.y)),
                   "r"(__float_as_uint(x1.z)), "r"(__float_as_uint(x1.w))
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      int s = 0;
      uint32_t ph = 0;
      for (i64 t = 0; t < ntiles; ++t)
        for (int kb = 0; kb < nb_hi + NB_LO; ++kb) {  // the r_hi blocks that hold data, then the r_lo blocks
          const int col = kb < nb_hi ? kb * TC_KB : (KB + kb - nb_hi) * TC_KB;
          mbar_wait_backoff(empty + s, ph ^ 1);
          if (dbg & 4) {
            mbar_arrive(full + s);
          } else {
            mbar_expect_tx(full + s, TC_BLOCK_BYTES);
            tma_load_2d(sB + (size_t)s * TC_BLOCK_BYTES, &mapB, full + s, col, (int)(t * TC_N));
          }
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    // The issue rate of this one warp is what the tensor pipe sees.  Two things made it the bottleneck of the kernel
    // (measured with the producer and the epilogue switched off: 178, then 122 cycles per MMA where the pipe needs
    // 64): (1) inside `if (lane == 0)` the compiler cannot prove that the operands of tcgen05.mma are warp-uniform
    // and wrapped every MMA in an elect / broadcast / branch waterfall -- so the whole warp walks the loop and ONE
    // elected lane issues; (2) with run-time slice counts a k-block cost ~150 dependent scalar instructions (predicates,
    // 64-bit descriptor arithmetic, vector -> uniform moves) -- so the schedule of a tile is unrolled at compile time
    // (KS): a barrier wait, then up to 8 MMAs whose operands differ by immediates.
    const uint64_t descB = umma_desc_sw128(smem_u32(sB));   // + byte offset >> 4 selects block / k-slice
    const bool extra = ks_hi > KS;                          // the hi segments hold one slice more than the lo ones
    const uint32_t tal0 = tmem_a + (uint32_t)ks_hi * 8;     // q_lo slices follow the q_hi slices
    int s = 0, acc = 0;
    uint32_t ph = 0, accph = 0;
    for (i64 t = 0; t < ntiles; ++t) {
      mbar_wait_backoff(tmem_empty + acc, accph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)acc * TC_N;
#pragma unroll
      for (int b = 0; b < NB_LO + 1; ++b) {        // r_hi blocks: q_hi . r_hi and q_lo . r_hi
        if (b < NB_LO || (b * 4 == KS && extra)) { // (the last one exists only when slice KS starts a new block)
          mbar_wait_backoff(full + s, ph);
          tc_fence_after();
          const uint64_t db = descB + (uint64_t)((s * TC_BLOCK_BYTES) >> 4);
          if (elect_one()) {
            if (!(dbg & 2)) {
#pragma unroll
              for (int k = 0; k < TC_KB / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes along the swizzled row
                const int sl = b * 4 + k;
                if (sl < KS) {
                  umma_tf32_ts(tmem_d, tmem_a + 8 * sl, db + 2 * k, TC_IDESC, sl ? 1u : 0u);
                  umma_tf32_ts(tmem_d, tal0 + 8 * sl, db + 2 * k, TC_IDESC, 1u);
                } else if (sl == KS) {
                  if (extra) umma_tf32_ts(tmem_d, tmem_a + 8 * sl, db + 2 * k, TC_IDESC, 1u);
                }
              }
            }
            umma_commit(empty + s);                // frees the B block when these MMAs have read it
          }
          __syncwarp();
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
#pragma unroll
      for (int b = 0; b < NB_LO; ++b) {            // r_lo blocks: q_hi . r_lo
        mbar_wait_backoff(full + s, ph);
        tc_fence_after();
        const uint64_t db = descB + (uint64_t)((s * TC_BLOCK_BYTES) >> 4);
        if (elect_one()) {
          if (!(dbg & 2)) {
#pragma unroll
            for (int k = 0; k < TC_KB / 8; ++k)
              if (b * 4 + k < KS) umma_tf32_ts(tmem_d, tmem_a + 8 * (b * 4 + k), db + 2 * k, TC_IDESC, 1u);
          }
          umma_commit(empty + s);
          if (b == NB_LO - 1) umma_commit(tmem_full + acc);   // accumulator complete
        }
        __syncwarp();
        if (++s == nstages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (++acc == nacc) {
        acc = 0;
        accph ^= 1;
      }
    }
  } else {
    // ---- epilogue: thread = query row; the warp owns 32 rows ----
    const int slab = warp & 3;                       // TMEM lanes 32 slab .. 32 slab + 31
    const int row = slab * 32 + lane;
    const i64 q = qtile * TC_M + row;
    const bool mine = (q >= q_begin && q < q_end);
    // shortlists of this CTA's 128 rows, best first, entry i of a row at [row][i] with i = 32 e + lane: shared memory
    // when they fit, else a global scratch
    float* lkey = lists_in_smem ? sm_lkey : gl_key + (size_t)blockIdx.x * TC_M * KP;
    int* lidx = lists_in_smem ? sm_lidx : gl_idx + (size_t)blockIdx.x * TC_M * KP;
    for (int r = 0; r < 32; ++r)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lkey[(size_t)(slab * 32 + r) * KP + e * 32 + lane] = -INFINITY;
        lidx[(size_t)(slab * 32 + r) * KP + e * 32 + lane] = INT_MAX;
      }
    __syncwarp();
    float thr = mine ? -INFINITY : INFINITY;         // rows outside this rank's range never pass
    const uint32_t pool_row = smem_u32(pool) + (uint32_t)row * 8;   // slot 0 of this thread's row
    const uint32_t pool_slab = smem_u32(pool) + (uint32_t)(slab * 32) * 8;
    uint32_t pw = pool_row;                           // where the next survivor goes
    // merges the pool of row r (warp-uniform) into its shortlist; n_new entries
    auto flush_row = [&](int r) {
      const int n_new = (int)((__shfl_sync(0xffffffffu, pw, r) - (pool_slab + (uint32_t)r * 8)) / TC_POOL_STRIDE);
      float pk = -INFINITY;
      int pi = INT_MAX;
      if (lane < n_new) {
        const uint32_t a = pool_slab + (uint32_t)r * 8 + (uint32_t)lane * TC_POOL_STRIDE;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(pk), "=r"(pi) : "r"(a));
      }
      // bitonic sort of the pool across the lanes, best first
#pragma unroll
      for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) cmpx_lane(pk, pi, j, ((lane & j) == 0) == ((lane & k2) == 0));
      // reversed pool against the LAST 32 entries of the list: the better of each pair; with the entries before
      // them this is a bitonic sequence holding the best KP of list + pool
      const float rk = __shfl_sync(0xffffffffu, pk, 31 - lane);
      const int ri = __shfl_sync(0xffffffffu, pi, 31 - lane);
      const size_t lo = (size_t)(slab * 32 + r) * KP + lane;
      float key[E];
      int idx[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        key[e] = lkey[lo + e * 32];
        idx[e] = lidx[lo + e * 32];
      }
      if (better(rk, ri, key[E - 1], idx[E - 1])) {
        key[E - 1] = rk;
        idx[E - 1] = ri;
      }
      if (E == 2) {                                   // half-cleaner at distance 32 lives in registers
        if (better(key[E - 1], idx[E - 1], key[0], idx[0])) {
          const float tk = key[0];
          const int ti = idx[0];
          key[0] = key[E - 1];
          idx[0] = idx[E - 1];
          key[E - 1] = tk;
          idx[E - 1] = ti;
        }
      }
#pragma unroll
      for (int j = 16; j > 0; j >>= 1)
#pragma unroll
        for (int e = 0; e < E; ++e) cmpx_lane(key[e], idx[e], j, (lane & j) == 0);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lkey[lo + e * 32] = key[e];
        lidx[lo + e * 32] = idx[e];
      }
      const float worst = __shfl_sync(0xffffffffu, key[E - 1], 31);
      if (lane == r) {
        thr = worst;
        pw = pool_row;
      }
    };
    // appends the survivors of 16 scores (registers x[o .. o + 15], references ref0 ..) to the thread's pool
    auto append16 = [&](const uint32_t (&x)[32], int o, int ref0) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (__uint_as_float(x[o + j]) > thr) {
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(pw), "r"(x[o + j]), "r"(ref0 + j) : "memory");
          pw += TC_POOL_STRIDE;
        }
    };
    // a pool that the next 16 scores could overflow is merged now (warp-uniform loop over the rows that need it)
    auto flush_check = [&]() {
      unsigned pending = __ballot_sync(0xffffffffu, pw > pool_row + (TC_POOL - TC_CHUNK) * TC_POOL_STRIDE);
      if (pending) {
        __syncwarp();                                 // pool writes visible to the whole warp
        while (pending) {
          const int r = __ffs(pending) - 1;
          pending &= pending - 1;
          flush_row(r);
        }
        __syncwarp();
      }
    };
    auto max16 = [](const uint32_t (&x)[32], int o) {
      const float a = fmaxf(fmaxf(__uint_as_float(x[o + 0]), __uint_as_float(x[o + 1])), __uint_as_float(x[o + 2]));
      const float b = fmaxf(fmaxf(__uint_as_float(x[o + 3]), __uint_as_float(x[o + 4])), __uint_as_float(x[o + 5]));
      const float c = fmaxf(fmaxf(__uint_as_float(x[o + 6]), __uint_as_float(x[o + 7])), __uint_as_float(x[o + 8]));
      const float d = fmaxf(fmaxf(__uint_as_float(x[o + 9]), __uint_as_float(x[o + 10])), __uint_as_float(x[o + 11]));
      const float e = fmaxf(fmaxf(__uint_as_float(x[o + 12]), __uint_as_float(x[o + 13])), __uint_as_float(x[o + 14]));
      return fmaxf(fmaxf(fmaxf(a, b), c), fmaxf(fmaxf(d, e), __uint_as_float(x[o + 15])));
    };
    // 32 scores of the row: a max tree and ONE vote when no row of the warp has a survivor (the common case at
    // large n); otherwise the halves are appended by the lanes that have one
    auto process32 = [&](const uint32_t (&x)[32], int ref0) {
      const float m0 = max16(x, 0), m1 = max16(x, 16);
      if (__any_sync(0xffffffffu, fmaxf(m0, m1) > thr)) {
        if (m0 > thr) append16(x, 0, ref0);
        flush_check();
        if (m1 > thr) append16(x, 16, ref0 + 16);
        flush_check();
      }
    };
    auto tmem_ld32 = [&](uint32_t (&d)[32], uint32_t taddr) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
            "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]),
            "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]),
            "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
          : "r"(taddr));
    };
    // the loaded registers are operands of the wait: nothing that reads them may be scheduled above it
    auto tmem_wait = [&](uint32_t (&d)[32]) {
      asm volatile("tcgen05.wait::ld.sync.aligned;"
                   : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]),
                     "+r"(d[8]), "+r"(d[9]), "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15]),
                     "+r"(d[16]), "+r"(d[17]), "+r"(d[18]), "+r"(d[19]), "+r"(d[20]), "+r"(d[21]), "+r"(d[22]), "+r"(d[23]),
                     "+r"(d[24]), "+r"(d[25]), "+r"(d[26]), "+r"(d[27]), "+r"(d[28]), "+r"(d[29]), "+r"(d[30]), "+r"(d[31])
                   :
                   : "memory");
    };
    int acc = 0;
    uint32_t accph = 0;
    for (i64 t = 0; t < ntiles; ++t) {
      mbar_wait(tmem_full + acc, accph);
      tc_fence_after();
      // two register sets, loads one step ahead: set B is in flight while set A is compared and vice versa
      uint32_t va[32], vb[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(slab * 32) << 16) + (uint32_t)(acc * TC_N);
      const int ref0 = (int)(t * TC_N);
      if (!(dbg & 1)) tmem_ld32(va, taddr);
#pragma unroll 1
      for (int c = 0; c < ((dbg & 1) ? 0 : TC_N / 64); ++c) {
        tmem_wait(va);
        tmem_ld32(vb, taddr + c * 64 + 32);
        process32(va, ref0 + c * 64);
        tmem_wait(vb);
        if (c + 1 < TC_N / 64) tmem_ld32(va, taddr + c * 64 + 64);
        process32(vb, ref0 + c * 64 + 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
      if (++acc == nacc) {
        acc = 0;
        accph ^= 1;
      }
    }
    {  // what is still in the pools
      unsigned pending = __ballot_sync(0xffffffffu, pw != pool_row);
      __syncwarp();
      while (pending) {
        const int r = __ffs(pending) - 1;
        pending &= pending - 1;
        flush_row(r);
      }
    }
    // hand the shortlists to knn_rerank in its convention: key = ||r||^2 - 2 q.r = -2 s, ascending, worst last
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      const i64 qq = qtile * TC_M + slab * 32 + r;
      if (qq >= q_begin && qq < q_end) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const size_t src = (size_t)(slab * 32 + r) * KP + e * 32 + lane;
          out_key[(size_t)(qq - q_begin) * KP + e * 32 + lane] = -2.0f * lkey[src];
          out_idx[(size_t)(qq - q_begin) * KP + e * 32 + lane] = lidx[src];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    REQUIRE(p && q == cudaDriverEntryPointSuccess, ADOBO_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

static CUtensorMap make_map(float* ptr, i64 rows, int Kt) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)Kt, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kt * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_M};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, ADOBO_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return m;
}

struct TcLaunch {
  const float* Aop;
  CUtensorMap mapB;
  int KB, ks_hi, nstages, lists_in_smem, dbg;
  i64 ntiles, qt0, q_begin, q_end;
  float* skey;
  int* sidx;
  float* glk;
  int* gli;
  unsigned grid;
  size_t smem;
  cudaStream_t st;
};
// one instantiation per KS = ceil(p / 8) = 1 .. 16 (p <= 126)
template <int E, int KS>
static bool tc_dispatch(adobo_ctx* c, const TcLaunch& L, int ks) {
  if constexpr (KS >= 1) {
    if (ks == KS) {
      CUDA_CHECK(cudaFuncSetAttribute(knn_scores_tc<E, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
      LAUNCH(c, "knn_scores_tc", (knn_scores_tc<E, KS><<<L.grid, TC_THREADS, L.smem, L.st>>>(
                                     L.Aop, L.mapB, L.KB, L.ks_hi, L.nstages, L.lists_in_smem, L.dbg, L.ntiles, L.qt0, L.q_begin,
                                     L.q_end, L.skey, L.sidx, L.glk, L.gli)));
      return true;
    }
    return tc_dispatch<E, KS - 1>(c, L, ks);
  } else {
    return false;
  }
}

// Can the tensor-core shortlist take this problem?
bool knn_tc_supported(int p) {
  return 8 * ((p + 2 + 7) / 8 + (p + 7) / 8) + 2 * TC_N <= 512;   // A' and two accumulators in tensor memory
}

// x: all points (n x p fp64, device).  Fills skey/sidx ((q_end-q_begin) x KP) for queries [q_begin, q_end),
// norms (npad) and maxnorm like the CUDA-core path.
void run_knn_shortlist_tc(adobo_ctx* c, const double* x, i64 n, int p, const double* ctr, i64 q_begin, i64 q_end, int E,
                          float* skey, int* sidx, DevBuf<float>& norms, unsigned* maxnorm) {
  cudaStream_t st = c->stream;
  const int Kp = (p + 2 + TC_KB - 1) / TC_KB * TC_KB;
  const int Kt = 2 * Kp, KB = Kp / TC_KB;
  const i64 npad = (n + TC_N - 1) / TC_N * TC_N;
  DevBuf<float> Aop, Bop;
  Aop.ensure((size_t)npad * Kt);
  Bop.ensure((size_t)npad * Kt);
  norms.ensure((size_t)npad);
  LAUNCH(c, "knn_tc_prep", (knn_tc_prep<<<(unsigned)((npad + 7) / 8), 256, 0, st>>>(x, n, npad, p, Kp, ctr, Aop.p, Bop.p,
                                                                                  norms.p, maxnorm)));
  const CUtensorMap mapB = make_map(Bop.p, npad, Kt);
  const int KP = 32 * E;
  const i64 qt0 = q_begin / TC_M, qt1 = (q_end + TC_M - 1) / TC_M;
  const i64 ntiles = npad / TC_N;
  // shared memory: A' resident + B' ring + pools + barriers (+ shortlists when they still leave >= 2 stages)
  const size_t cap = 227 * 1024, fixed = 1024 /*align*/ + (size_t)TC_POOL_BYTES + 256;
  const size_t lists = (size_t)TC_M * KP * 8;
  const size_t a_bytes = 0;   // A' is kept in tensor memory
  const int lists_in_smem = (a_bytes + 2 * (size_t)TC_BLOCK_BYTES + fixed + lists <= cap) ? 1 : 0;
  int nstages = (int)((cap - a_bytes - fixed - (lists_in_smem ? lists : 0)) / TC_BLOCK_BYTES);
  if (nstages > 8) nstages = 8;  // the barrier block holds 8 full / empty pairs
  // diagnostic: switch pipeline roles off (1 = epilogue, 2 = MMAs, 4 = TMA loads) to time the others alone -- how the
  // issue-bound MMA loop was found (profiles/r02_knn_stages.md); results are meaningless with any bit set
  const int dbg = getenv("ADOBO_B200_KNN_DBG") ? atoi(getenv("ADOBO_B200_KNN_DBG")) : 0;
  REQUIRE(nstages >= 2, ADOBO_ERR_LIMIT, "knn: embedding too wide for the tensor-core shortlist");
  const size_t smem = a_bytes + (size_t)nstages * TC_BLOCK_BYTES + fixed + (lists_in_smem ? lists : 0);
  DevBuf<float> glk;
  DevBuf<int> gli;
  if (!lists_in_smem) {
    glk.ensure((size_t)(qt1 - qt0) * TC_M * KP);
    gli.ensure((size_t)(qt1 - qt0) * TC_M * KP);
  }
  WORK(c, 2.0 * (double)(qt1 - qt0) * TC_M * (double)npad * 8.0 * ((p + 2 + 7) / 8 + 2 * ((p + 7) / 8)));  // tf32 flops actually issued (three products, k-slices of 8)
  TcLaunch L{Aop.p, mapB, KB, (p + 2 + 7) / 8, nstages, lists_in_smem, dbg, ntiles, qt0, q_begin, q_end, skey, sidx, glk.p, gli.p,
             (unsigned)(qt1 - qt0), smem, st};
  const int KS = (p + 7) / 8;
  bool done = (E == 1) ? tc_dispatch<1, 16>(c, L, KS) : tc_dispatch<2, 16>(c, L, KS);
  REQUIRE(done, ADOBO_ERR_LIMIT, "knn: embedding too wide for the tensor-core shortlist");
  CUDA_CHECK(cudaStreamSynchronize(st));  // Aop / Bop die here
}

}  // namespace adobo
