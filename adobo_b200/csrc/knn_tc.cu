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
//   warps 2-5  epilogue: tcgen05.ld 32x32b.x16 (thread = query row), one compare per score against the
//              row's threshold; the rare survivors go to a 16-slot pool per row in shared memory and
//              the warp then merges the pools of its 32 rows cooperatively into the sorted shortlists
//              (lanes hold the list, shuffle insertion -- no divergent per-thread insertion)
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
__global__ void knn_tc_prep(const double* __restrict__ x, i64 n, i64 npad, int p, int Kp, const double* __restrict__ ctr,
                            float* __restrict__ Aop, float* __restrict__ Bop, float* __restrict__ norms,
                            unsigned* __restrict__ maxnorm_bits) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  const int Kt = 2 * Kp;
  float* a = Aop + (size_t)i * Kt;
  float* b = Bop + (size_t)i * Kt;
  for (int d = 0; d < Kt; ++d) {
    a[d] = 0.f;
    b[d] = 0.f;
  }
  if (i < n) {
    float nrm = 0.f;
    for (int d = 0; d < p; ++d) {
      const float v = (float)(x[i * p + d] - ctr[d]);
      nrm = fmaf(v, v, nrm);
      const float hi = tf32_round(v), lo = tf32_round(v - hi);
      a[d] = hi;
      a[Kp + d] = lo;
      b[d] = hi;
      b[Kp + d] = lo;
    }
    norms[i] = nrm;
    atomicMax(maxnorm_bits, __float_as_uint(nrm));
    const float hn = -0.5f * nrm, hh = tf32_round(hn), hl = tf32_round(hn - hh);
    a[p] = 1.f;
    a[p + 1] = 1.f;
    b[p] = hh;
    b[p + 1] = hl;
  } else {
    norms[i] = INFINITY;
    a[p] = 1.f;
    b[p] = -1e30f;  // padding rows score -1e30: never a candidate
  }
}

// ---- the tensor-core kernel ------------------------------------------------------------------------
constexpr int TC_POOL = 16;  // pool slots per query row = columns per tcgen05.ld step

template <int E>  // shortlist length KP = 32 E
__global__ void __launch_bounds__(TC_THREADS, 1)
    knn_scores_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  int KB /* k-blocks per segment: A' and B' hold 2 KB each */, int nstages, int lists_in_smem, i64 ntiles, i64 qtile0, i64 q_begin, i64 q_end,
                  float* __restrict__ out_key, int* __restrict__ out_idx, float* __restrict__ gl_key,
                  int* __restrict__ gl_idx) {
  constexpr int KP = 32 * E;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand blocks first, then shortlists / pools / barriers
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;                                   // 2 KB blocks: q_hi, then q_lo
  unsigned char* sB = base + (size_t)2 * KB * TC_BLOCK_BYTES; // nstages blocks
  unsigned char* after = sB + (size_t)nstages * TC_BLOCK_BYTES;
  float* pool_key = reinterpret_cast<float*>(after);          // [128][TC_POOL]
  int* pool_idx = reinterpret_cast<int*>(pool_key + TC_M * TC_POOL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(pool_idx + TC_M * TC_POOL);
  uint64_t* full = bars;                 // [nstages]
  uint64_t* empty = bars + 8;            // [nstages]
  uint64_t* a_full = bars + 16;
  uint64_t* tmem_full = bars + 17;       // [2]
  uint64_t* tmem_empty = bars + 19;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* sm_lkey = reinterpret_cast<float*>(bars + 24);       // [128][KP] when lists_in_smem
  int* sm_lidx = reinterpret_cast<int*>(sm_lkey + TC_M * KP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const i64 qtile = qtile0 + blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(a_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full + b, 1);
      mbar_init(tmem_empty + b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 accumulators x 128 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      mbar_expect_tx(a_full, (uint32_t)(2 * KB) * TC_BLOCK_BYTES);
      for (int kb = 0; kb < 2 * KB; ++kb)
        tma_load_2d(sA + (size_t)kb * TC_BLOCK_BYTES, &mapA, a_full, kb * TC_KB, (int)(qtile * TC_M));
      int s = 0;
      uint32_t ph = 0;
      for (i64 t = 0; t < ntiles; ++t)
        for (int kb = 0; kb < 2 * KB; ++kb) {  // r_hi blocks, then r_lo blocks
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, TC_BLOCK_BYTES);
          tma_load_2d(sB + (size_t)s * TC_BLOCK_BYTES, &mapB, full + s, kb * TC_KB, (int)(t * TC_N));
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      mbar_wait(a_full, 0);
      tc_fence_after();
      int s = 0;
      uint32_t ph = 0;
      for (i64 t = 0; t < ntiles; ++t) {
        const int acc = (int)(t & 1);
        const uint32_t accph = (uint32_t)((t >> 1) & 1);
        mbar_wait(tmem_empty + acc, accph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * TC_N;
        for (int kb = 0; kb < 2 * KB; ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const bool hi = kb < KB;                   // staged block: r_hi[kb] or r_lo[kb - KB]
          const int ka = hi ? kb : kb - KB;
          const uint32_t ah_addr = smem_u32(sA + (size_t)ka * TC_BLOCK_BYTES);         // q_hi
          const uint32_t al_addr = smem_u32(sA + (size_t)(KB + ka) * TC_BLOCK_BYTES);  // q_lo
          const uint32_t b_addr = smem_u32(sB + (size_t)s * TC_BLOCK_BYTES);
#pragma unroll
          for (int k = 0; k < TC_KB / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes along the swizzled row
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
            umma_tf32(tmem_d, umma_desc_sw128(ah_addr + k * 32), db, TC_IDESC, (kb | k) ? 1u : 0u);  // q_hi.r_hi / q_hi.r_lo
            if (hi) umma_tf32(tmem_d, umma_desc_sw128(al_addr + k * 32), db, TC_IDESC, 1u);         // q_lo.r_hi
          }
          umma_commit(empty + s);  // frees the B block when these MMAs have read it
          if (++s == nstages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(tmem_full + acc);  // accumulator complete
      }
    }
  } else {
    // ---- epilogue: thread = query row; the warp owns 32 rows ----
    const int slab = warp & 3;                       // TMEM lanes 32 slab .. 32 slab + 31
    const int row = slab * 32 + lane;
    const i64 q = qtile * TC_M + row;
    const bool mine = (q >= q_begin && q < q_end);
    // shortlists of this CTA's 128 rows: shared memory when they fit, else a global scratch
    float* lkey = lists_in_smem ? sm_lkey : gl_key + (size_t)blockIdx.x * TC_M * KP;
    int* lidx = lists_in_smem ? sm_lidx : gl_idx + (size_t)blockIdx.x * TC_M * KP;
    for (int r = 0; r < 32; ++r)                     // keys are -score: ascending = best first
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lkey[(size_t)(slab * 32 + r) * KP + lane * E + e] = INFINITY;
        lidx[(size_t)(slab * 32 + r) * KP + lane * E + e] = INT_MAX;
      }
    __syncwarp();
    float thr = mine ? -INFINITY : INFINITY;         // rows outside this rank's range never pass
    float* my_pk = pool_key + row * TC_POOL;
    int* my_pi = pool_idx + row * TC_POOL;
    for (i64 t = 0; t < ntiles; ++t) {
      const int acc = (int)(t & 1);
      const uint32_t accph = (uint32_t)((t >> 1) & 1);
      mbar_wait(tmem_full + acc, accph);
      tc_fence_after();
      // the TMEM loads are pipelined one chunk deep: chunk c + 1 is in flight while chunk c is compared
      uint32_t v[16], w[16];
      auto tmem_ld16 = [&](uint32_t (&d)[16], int c) {
        const uint32_t taddr = tmem_base + ((uint32_t)(slab * 32) << 16) + (uint32_t)(acc * TC_N + c * TC_POOL);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
              "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
            : "r"(taddr));
      };
      tmem_ld16(w, 0);
#pragma unroll 1
      for (int c = 0; c < TC_N / TC_POOL; ++c) {
        // the loaded registers are operands of the wait: nothing that reads them may be scheduled above it
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]),
                       "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15])
                     :
                     : "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = w[j];
        if (c + 1 < TC_N / TC_POOL) tmem_ld16(w, c + 1);
        int cnt = 0;
        const int ref0 = (int)(t * TC_N + c * TC_POOL);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float sc = __uint_as_float(v[j]);
          if (sc > thr) {
            my_pk[cnt] = -sc;
            my_pi[cnt] = ref0 + j;
            ++cnt;
          }
        }
        unsigned pending = __ballot_sync(0xffffffffu, cnt > 0);
        if (pending) __syncwarp();                   // pool writes visible to the whole warp
        while (pending) {                            // warp-uniform: merge one row's pool at a time
          const int r = __ffs(pending) - 1;
          pending &= pending - 1;
          const int n_new = __shfl_sync(0xffffffffu, cnt, r);
          const size_t lo = (size_t)(slab * 32 + r) * KP + lane * E;
          float key[E];
          int idx[E];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            key[e] = lkey[lo + e];
            idx[e] = lidx[lo + e];
          }
          const float* pk = pool_key + (slab * 32 + r) * TC_POOL;
          const int* pi = pool_idx + (slab * 32 + r) * TC_POOL;
          for (int s2 = 0; s2 < n_new; ++s2) list_insert<float, E>(key, idx, pk[s2], pi[s2], lane);
#pragma unroll
          for (int e = 0; e < E; ++e) {
            lkey[lo + e] = key[e];
            lidx[lo + e] = idx[e];
          }
          const float worst = __shfl_sync(0xffffffffu, key[E - 1], 31);
          if (lane == r) thr = -worst;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
    }
    // hand the shortlists to knn_rerank in its convention: key = ||r||^2 - 2 q.r = -2 s, ascending
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      const i64 qq = qtile * TC_M + slab * 32 + r;
      if (qq >= q_begin && qq < q_end) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const size_t src = (size_t)(slab * 32 + r) * KP + lane * E + e;
          out_key[(size_t)(qq - q_begin) * KP + lane * E + e] = 2.0f * lkey[src];
          out_idx[(size_t)(qq - q_begin) * KP + lane * E + e] = lidx[src];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
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

// Can the tensor-core shortlist take this problem?  A' must stay resident next to >= 2 B' stages.
bool knn_tc_supported(int p) {
  const int Kp = (p + 2 + TC_KB - 1) / TC_KB * TC_KB;
  const int KB = Kp / TC_KB;
  return (size_t)(2 * KB + 2) * TC_BLOCK_BYTES + 1024 + (size_t)TC_M * TC_POOL * 8 + 256 <= 227 * 1024;
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
  LAUNCH(c, "knn_tc_prep", (knn_tc_prep<<<(unsigned)((npad + 127) / 128), 128, 0, st>>>(x, n, npad, p, Kp, ctr, Aop.p, Bop.p,
                                                                                      norms.p, maxnorm)));
  const CUtensorMap mapA = make_map(Aop.p, npad, Kt), mapB = make_map(Bop.p, npad, Kt);
  const int KP = 32 * E;
  const i64 qt0 = q_begin / TC_M, qt1 = (q_end + TC_M - 1) / TC_M;
  const i64 ntiles = npad / TC_N;
  // shared memory: A' resident + B' ring + pools + barriers (+ shortlists when they still leave >= 2 stages)
  const size_t cap = 227 * 1024, fixed = 1024 /*align*/ + (size_t)TC_M * TC_POOL * 8 + 256;
  const size_t lists = (size_t)TC_M * KP * 8;
  const size_t a_bytes = (size_t)2 * KB * TC_BLOCK_BYTES;
  const int lists_in_smem = (a_bytes + 2 * (size_t)TC_BLOCK_BYTES + fixed + lists <= cap) ? 1 : 0;
  int nstages = (int)((cap - a_bytes - fixed - (lists_in_smem ? lists : 0)) / TC_BLOCK_BYTES);
  if (nstages > 8) nstages = 8;  // the barrier block holds 8 full / empty pairs
  REQUIRE(nstages >= 2, ADOBO_ERR_LIMIT, "knn: embedding too wide for the tensor-core shortlist");
  const size_t smem = a_bytes + (size_t)nstages * TC_BLOCK_BYTES + fixed + (lists_in_smem ? lists : 0);
  DevBuf<float> glk;
  DevBuf<int> gli;
  if (!lists_in_smem) {
    glk.ensure((size_t)(qt1 - qt0) * TC_M * KP);
    gli.ensure((size_t)(qt1 - qt0) * TC_M * KP);
  }
  WORK(c, 2.0 * (double)(qt1 - qt0) * TC_M * (double)npad * 3.0 * Kp);  // tf32 flops actually issued (three products)
  if (E == 1) {
    CUDA_CHECK(cudaFuncSetAttribute(knn_scores_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(c, "knn_scores_tc", (knn_scores_tc<1><<<(unsigned)(qt1 - qt0), TC_THREADS, smem, st>>>(
                                   mapA, mapB, KB, nstages, lists_in_smem, ntiles, qt0, q_begin, q_end, skey, sidx, glk.p,
                                   gli.p)));
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(knn_scores_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(c, "knn_scores_tc", (knn_scores_tc<2><<<(unsigned)(qt1 - qt0), TC_THREADS, smem, st>>>(
                                   mapA, mapB, KB, nstages, lists_in_smem, ntiles, qt0, q_begin, q_end, skey, sidx, glk.p,
                                   gli.p)));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));  // Aop / Bop die here
}

}  // namespace adobo
