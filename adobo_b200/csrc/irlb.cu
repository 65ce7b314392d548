// K4-K7  Augmented implicitly restarted Lanczos bidiagonalization, device resident.
//
// Replaces adobo/irlbpy/irlb.py:95-258 (`lanczos`, with `multA` :25-34, `orthog` :73-79,
// `invcheck` :84-92) as dr.irlb drives it (adobo/dr.py:154-171).  The reference densifies the
// HVG subset and rewrites it with sklearn.preprocessing.scale; here the subset stays sparse
// (fp32 values, CSR for A.v and CSC for A^T.u) and the centring / scaling is applied inside the
// mat-vecs:   A_s v = X (v/sigma) - mu.(v/sigma) 1      A_s^T u = (X^T u - mu (1^T u)) / sigma
// All basis arithmetic (V: h x m_b replicated, W: n_local x m_b row-sharded, B: m_b x m_b) is
// fp64 with the reference's operation order: one pass of classical Gram-Schmidt against the
// whole basis, explicit norms, the same restart rule.  The small SVD of B (numpy.linalg.svd in
// the reference, irlb.py:224) is a one-sided Jacobi SVD in one CTA's shared memory; the host only
// reads two integers (conv, k) per restart.
//
// Reductions are two-stage and ordered (per-CTA partials, summed by the last CTA to finish), so
// a given shard layout gives bit-identical results run to run.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <atomic>

#include "common.cuh"

namespace adobo {
namespace cg = cooperative_groups;

// svd_arrow.cu: structured SVD of the restarted B (diag + spike + bidiagonal tail)
bool svd_arrow_supported(int m, int k);
void svd_arrow_prepare(int m);
void launch_svd_arrow(adobo_ctx* c, const double* B, int m, int k, double* Su, double* sv, double* Sv, int* info,
                      cudaStream_t st, LastDiag ldg = LastDiag());
int svd_arrow_cluster();

enum { SC_S = 0, SC_FN, SC_FNINV, SC_SU, SC_MVS, SC_SMAX, SC_NRM2, SC_AUX, SC_COUNT = 16 };
constexpr int MAXB = 128;  // widest basis the dot kernels handle (4 columns per lane)
constexpr double EPS2 = 2.0 * 2.220446049250313e-16;

__device__ __forceinline__ double inv_or_zero(double x) { return x > EPS2 ? 1.0 / x : 0.0; }  // irlb.py:84-92

// Diagnostic build (-DADOBO_B200_TIMELINE): CTA 0 stamps entry and the end of the dependency wait, every CTA
// its exit, per (kernel, column) slot -- the timeline of the last replay of the iteration graph.
#ifdef ADOBO_B200_TIMELINE
__device__ unsigned long long tl_buf[6][1024];
__device__ __forceinline__ unsigned long long tl_gt() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ unsigned long long tl_cta[8][256];  // per-CTA stamps of the last w_step launch
__device__ unsigned long long tl_ctak[2][10][256];  // ... of the last spmvT_csc_sm (0) and v_step_one (1) launches
#define TL_CTA(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) tl_cta[i][blockIdx.x] = tl_gt(); } while (0)
#define TL_CTAK(k, i) do { if (threadIdx.x == 0 && blockIdx.x < 256 && tl_on) tl_ctak[k][i][blockIdx.x] = tl_gt(); } while (0)
#define TL_ENTRY(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { tl_buf[0][slot] = tl_gt(); tl_buf[3][slot] = clock64(); } } while (0)
// kernels with a dependency wait: entry stamps are kept in registers and stored after the halt check
#define TL_ENTRY_REG() const unsigned long long tl_e0 = tl_gt(), tl_e1 = clock64()
#define TL_SYNC(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { tl_buf[0][slot] = tl_e0; tl_buf[3][slot] = tl_e1; tl_buf[1][slot] = tl_gt(); tl_buf[4][slot] = clock64(); } } while (0)
#define TL_EXIT(slot) do { if (threadIdx.x == 0) { atomicMax(&tl_buf[2][slot], tl_gt()); atomicMax(&tl_buf[5][slot], (unsigned long long)clock64()); } } while (0)
#else
#define TL_CTA(i) do { } while (0)
#define TL_CTAK(k, i) do { } while (0)
#define TL_ENTRY(slot) do { } while (0)
#define TL_ENTRY_REG() do { } while (0)
#define TL_SYNC(slot) do { } while (0)
#define TL_EXIT(slot) do { } while (0)
#endif

// Programmatic dependent launch: wait for the previous kernel of the stream, THEN release the next one.  Releasing
// only after the own wait means that when a kernel starts, everything two or more kernels back has completed.
__device__ __forceinline__ void pdl_sync() {
  cudaGridDependencySynchronize();         // no-op without the launch attribute
  cudaTriggerProgrammaticLaunchCompletion();
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

// ---- ordered cross-CTA reductions (row kernels run 1024-thread CTAs, at most one per SM) --------
constexpr int RT_THREADS = 1024;
constexpr int RT_WARPS = RT_THREADS / 32;

// vector of nd (<= 128) sums.  Every warp first deposits its partial vector in shared memory (deposit_*),
// then reduce_vec_finish sums over warps, writes the CTA partial, and the last CTA to arrive sums the CTA
// partials in a fixed order (and, when ranks > 1, exchanges the result over NVLink peer memory).
__device__ __forceinline__ double (*vec_stage())[MAXB] {
  __shared__ double sh[RT_WARPS][MAXB];
  return sh;
}
__device__ __forceinline__ double (*part_stage())[MAXB] {  // second level: [part][column]
  __shared__ double ps[RT_THREADS / MAXB][MAXB];
  return ps;
}
// acc[t] is this lane's partial for column lane + 32 t
__device__ __forceinline__ void deposit_lane32(double (&acc)[4]) {
  double(*sh)[MAXB] = vec_stage();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 4; ++t) sh[warp][lane + 32 * t] = acc[t];
}
// acc[q] is the partial for column (lane & 7) + 8 q, held by all four 8-lane groups of the warp
template <int NQ>
__device__ __forceinline__ void deposit_lane8(double (&acc)[NQ]) {
  double(*sh)[MAXB] = vec_stage();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double v = acc[q];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 8) sh[warp][lane + 8 * q] = v;
  }
  if (NQ * 8 < MAXB)
    for (int c = NQ * 8 + lane; c < MAXB; c += 32) sh[warp][c] = 0.0;
}
constexpr int SP_PARTS = 8, SP_PER = 19;  // 8 x 19 >= 148 CTAs: a part's loads are issued as one batch
// the part sum of one column: CTAs part, part + SP_PARTS, ... in ascending order
template <int PER = SP_PER>
__device__ __forceinline__ double sum_partials_part(const double* __restrict__ src, int part, int nparts, int nblk,
                                                    int stride = MAXB) {
  double p = 0;
  for (int base = part; base < nblk; base += nparts * PER) {  // one trip for <= nparts x PER CTAs
    double v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int b = base + nparts * i;
      v[i] = b < nblk ? __ldcg(src + (size_t)b * stride) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) p += v[i];
  }
  return p;
}
// Both levels are laid out for coalescing: thread = (column, part); a part sums every nparts-th warp (CTA) partial
// of its column with independent loads, then one thread per column adds the parts in a fixed order.  (Lanes striding
// the CTA partials of ONE column -- 1 KB apart -- cost the last CTA 15 k separate sector requests: 8.7 us measured.)
__device__ __forceinline__ void reduce_vec_finish(int nd, double* __restrict__ partials, double* __restrict__ out,
                                                  unsigned* __restrict__ counter, const PeerComm& pc) {
  double(*sh)[MAXB] = vec_stage();
  double(*ps)[MAXB] = part_stage();
  __shared__ bool is_last;
  const int tid = threadIdx.x, col = tid & (MAXB - 1), part = tid >> 7;
  const int nwarps = blockDim.x >> 5, nparts = blockDim.x >> 7, wpp = nwarps / nparts;  // blockDim: multiple of 128
  __syncthreads();
  TL_CTA(3);
  {
    double p = 0;
    for (int w = 0; w < wpp; ++w) p += sh[part * wpp + w][col];
    ps[part][col] = p;
  }
  __syncthreads();
  if (tid < nd) {
    double s = 0;
    for (int g = 0; g < nparts; ++g) s += ps[g][tid];
    partials[(size_t)blockIdx.x * MAXB + tid] = s;
  }
  TL_CTA(4);
  __threadfence();
  __syncthreads();
  TL_CTA(5);
  if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  TL_CTA(6);
  if (is_last) {
    const double s = col < nd ? sum_partials_part(partials + col, part, nparts, (int)gridDim.x) : 0.0;  // batched loads
    ps[part][col] = s;
    __syncthreads();
    if (tid < nd) {
      double t = 0;
      for (int g = 0; g < nparts; ++g) t += ps[g][tid];
      out[tid] = t;
    }
    if (tid == 0) *counter = 0;
    if (pc.world > 1) {  // sum over ranks, fused: this CTA exchanges the vector over NVLink peer memory
      __syncthreads();
      peer_allreduce_block(pc, out, nd);
    }
  }
}
// Split reduction (single GPU): the producer stops at the per-CTA partials -- no fence, no atomic, no serial last
// CTA on the step's critical path -- and every consumer sums the CTA partials of the columns it needs itself, in
// one fixed order: part p = CTAs p, p + 8, ... (ascending), then ((P0 + P1) + ...) + P7.
__device__ __forceinline__ void reduce_vec_partials(int nd, double* __restrict__ partials) {
  double(*sh)[MAXB] = vec_stage();
  double(*ps)[MAXB] = part_stage();
  const int tid = threadIdx.x, col = tid & (MAXB - 1), part = tid >> 7;
  const int nwarps = blockDim.x >> 5, nparts = blockDim.x >> 7, wpp = nwarps / nparts;
  __syncthreads();
  {
    double p = 0;
    for (int w = 0; w < wpp; ++w) p += sh[part * wpp + w][col];
    ps[part][col] = p;
  }
  __syncthreads();
  if (tid < nd) {
    double s = 0;
    for (int g = 0; g < nparts; ++g) s += ps[g][tid];
    partials[(size_t)blockIdx.x * MAXB + tid] = s;
  }
}
// one column summed by a full warp: lanes stride the CTAs (<= 5 batched loads each for <= 160 CTAs), butterfly
__device__ __forceinline__ double sum_partials_warp(const double* __restrict__ src, int nblk) {
  const int lane = threadIdx.x & 31;
  double v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = (lane + 32 * i < nblk) ? __ldcg(src + (size_t)(lane + 32 * i) * MAXB) : 0.0;
  double p = ((v[0] + v[1]) + (v[2] + v[3])) + v[4];
  for (int b = lane + 160; b < nblk; b += 32) p += __ldcg(src + (size_t)b * MAXB);
  return warp_sum(p);
}
// columns col0 and col0 + 1 (||y||^2 and sum y of w_step's vector): called by one full warp, results in all lanes;
// 16 parts per column (lanes 0-15 / 16-31), summed in lane order
__device__ __forceinline__ void sum_partials_two(const double* __restrict__ partials, int nblk, int col0, double& t0,
                                                 double& t1) {
  const int lane = threadIdx.x & 31;
  const double p = sum_partials_part<10>(partials + col0 + (lane >> 4), lane & 15, 16, nblk);  // 16 x 10 >= 148 CTAs
  t0 = __shfl_sync(0xffffffffu, p, 0);
  t1 = __shfl_sync(0xffffffffu, p, 16);
#pragma unroll
  for (int l = 1; l < 16; ++l) {
    t0 += __shfl_sync(0xffffffffu, p, l);
    t1 += __shfl_sync(0xffffffffu, p, 16 + l);
  }
}
__device__ __forceinline__ void reduce_vec_blocks(double (&acc)[4], int nd, double* __restrict__ partials,
                                                  double* __restrict__ out, unsigned* __restrict__ counter,
                                                  const PeerComm& pc) {
  deposit_lane32(acc);
  reduce_vec_finish(nd, partials, out, counter, pc);
}

// two scalars
__device__ __forceinline__ bool reduce_two_blocks(double a, double b, double* __restrict__ partials,
                                                  double* __restrict__ out, unsigned* __restrict__ counter,
                                                  const PeerComm& pc) {
  __shared__ double sh2[RT_WARPS][2];
  __shared__ bool is_last2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh2[warp][0] = a;
    sh2[warp][1] = b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0;
    const int nwarps = blockDim.x >> 5;
    for (int w = 0; w < nwarps; ++w) s += sh2[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last2 = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last2) {
    if (warp == 0) {
      double s0 = 0, s1 = 0;
      for (unsigned bb = lane; bb < gridDim.x; bb += 32) {
        s0 += __ldcg(partials + (size_t)bb * 2);
        s1 += __ldcg(partials + (size_t)bb * 2 + 1);
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        out[0] = s0;
        out[1] = s1;
        *counter = 0;
      }
    }
    if (pc.world > 1) {
      __syncthreads();
      peer_allreduce_block(pc, out, 2);
    }
  }
  return is_last2;
}

// Epilogue of the norm reduction, run by the LAST CTA only (all its threads): normalise the vector and
// store it -- the work of v_finish / w_finish without another kernel launch.  kind 1: F -> V[:, jn], x = F/sigma,
// mu.x, B[j,j], B[j,j+1] (irlb.py:191-197, 221);  kind 2: w -> W[:, jn] and its contiguous copy (irlb.py:176-178, 217-219).
struct Finish {
  int kind = 0;  // 0: none (a separate finish kernel follows)
  const double* rs = nullptr;
  double* sc = nullptr;
  double* M = nullptr;  // V or W
  int ld = 0, jn = 0, m_b = 0, j = -1, centered = 0;  // kind 2: m_b = row length of B
  double* fbuf = nullptr;  // kind 1: normalised F;  kind 2: wcur
  double* xs = nullptr;
  double* B = nullptr;
};
__device__ __forceinline__ void finish_in_last_cta(const Finish& f, const double* __restrict__ ybuf, i64 rows) {
  __syncthreads();  // out2 (written by lane 0 / the peer exchange) is visible to the whole CTA
  const double nrm = sqrt(__ldcg(f.sc + SC_NRM2));
  const double inv = inv_or_zero(nrm);
  const double aux = __ldcg(f.sc + SC_AUX);
  if (f.kind == 1) {
    for (i64 r = threadIdx.x; r < rows; r += blockDim.x) {
      const double v = inv * __ldcg(ybuf + r);
      f.fbuf[r] = v;
      if (f.jn < f.m_b) f.M[(size_t)r * f.ld + f.jn] = v;
      f.xs[r] = f.rs ? v * f.rs[r] : v;
    }
    if (threadIdx.x == 0) {
      f.sc[SC_FN] = nrm;
      f.sc[SC_FNINV] = inv;
      f.sc[SC_MVS] = f.centered ? inv * aux : 0.0;
      if (f.j >= 0 && f.j < f.m_b - 1) f.B[(size_t)f.j * f.m_b + f.j + 1] = nrm;  // irlb.py:197
    }
  } else {
    for (i64 r = threadIdx.x; r < rows; r += blockDim.x) {
      const double v = inv * __ldcg(ybuf + r);
      f.M[(size_t)r * f.ld + f.jn] = v;
      f.fbuf[r] = v;
    }
    if (threadIdx.x == 0) {
      f.sc[SC_S] = nrm;
      f.sc[SC_SU] = inv * aux;
      f.B[(size_t)f.jn * f.m_b + f.jn] = nrm;  // irlb.py:196 / 221
    }
  }
}

// ---- K5  A^T u on the CSC copy: one CTA per gene column ---------------------------------------
// With `wf.W != nullptr` the kernel also does the work of w_finish for the column of W it consumes: u = ybuf / ||w||
// is applied as one factor on the finished sum (A^T is linear), and CTA `col` stores its slice of W[:, jn].
struct WFinish {
  double* W = nullptr;  // nullptr: u is already normalised (w_finish ran)
  int ldw = 0, jn = 0, m_b = 0;
  i64 n = 0;
  double* sc = nullptr;
  double* B = nullptr;
  const double* n2p = nullptr;  // (||w||^2, sum w) of the column: sc + SC_NRM2, or the tail of w_step's vector
  const double* parts = nullptr;  // split reduction: per-CTA partials of w_step (columns jn, jn + 1 hold the two)
  int nblk = 0;
  double* gout = nullptr;  // split reduction: the summed vector (g = W^T y, ||y||^2, sum y) for v_step
};
constexpr int ST_EPF = 4;  // entries per thread fetched before the grid-dependency wait (1024 per column)
// (5 CTAs per SM: the partial-sum duty of warps 0 / 1 must not cost the streaming loop its occupancy -- at 72
// registers the kernel ran 73 instead of 34 us on a 160 k-cell shard)
template <typename VT>
__global__ void __launch_bounds__(256, 5) spmvT_csc(const i64* __restrict__ cptr, const int* __restrict__ ridx,
                                                 const VT* __restrict__ cval, const double* __restrict__ u,
                                                 double* __restrict__ t, int h, unsigned* __restrict__ counter,
                                                 const PeerComm pc, const WFinish wf, const int* __restrict__ halt) {
  __shared__ double red[8];
  __shared__ bool last_cta;
  const int col = blockIdx.x;
  TL_ENTRY_REG();
  // prologue: the static CSC arrays only (see w_step for the programmatic-launch rules)
  const i64 b = __ldg(cptr + col), e1 = __ldg(cptr + col + 1);
  int ri[ST_EPF];
  VT cv[ST_EPF];
#pragma unroll
  for (int q = 0; q < ST_EPF; ++q) {
    const i64 e = b + threadIdx.x + 256 * q;
    ri[q] = e < e1 ? __ldg(ridx + e) : 0;
    cv[q] = e < e1 ? __ldg(cval + e) : VT(0);
  }
  pdl_sync();
  if (halt && *halt) return;
  TL_SYNC(128 + (wf.W ? wf.jn : 127));
  __shared__ double s_n2[2];
  // the gathers of the prefetched entries go out first: they are in flight while warps 0 / 1 sum the partials
  double uv[ST_EPF];
#pragma unroll
  for (int q = 0; q < ST_EPF; ++q) uv[q] = (b + threadIdx.x + 256 * q < e1) ? u[ri[q]] : 0.0;
  // Warp 0 owns the norm of the W column this launch consumes: it sums the two scalars from w_step's partials
  // (split reduction) or reads them, stores the CTA's slice of W[:, jn] = u / ||u|| and publishes the pair for the
  // epilogue.  Nobody waits for it before the end of the dot product; warp 1 of CTA b sums column b of g for v_step.
  if (wf.W) {
    if (threadIdx.x < 32) {
      double t0, t1;
      if (wf.parts)
        sum_partials_two(wf.parts, wf.nblk, wf.jn, t0, t1);
      else
        t0 = wf.n2p[0], t1 = wf.n2p[1];
      const double nrm0 = sqrt(t0), sinv0 = inv_or_zero(nrm0);
      const i64 chunk = (wf.n + gridDim.x - 1) / gridDim.x, lo = (i64)col * chunk, hi = min(wf.n, lo + chunk);
      for (i64 r = lo + threadIdx.x; r < hi; r += 32) wf.W[(size_t)r * wf.ldw + wf.jn] = sinv0 * u[r];
      if (threadIdx.x == 0) {
        s_n2[0] = t0, s_n2[1] = t1;
        if (wf.parts && wf.gout && col == 0) wf.gout[wf.jn] = t0, wf.gout[wf.jn + 1] = t1;
        if (col == 0) {
          wf.sc[SC_S] = nrm0;
          wf.sc[SC_SU] = sinv0 * t1;
          wf.B[(size_t)wf.jn * wf.m_b + wf.jn] = nrm0;  // irlb.py:196 / 221
        }
      }
    } else if (wf.parts && wf.gout && threadIdx.x < 64) {
      for (int gc = col; gc < wf.jn; gc += gridDim.x) {
        const double g = sum_partials_warp(wf.parts + gc, wf.nblk);
        if ((threadIdx.x & 31) == 0) wf.gout[gc] = g;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < ST_EPF; ++q) s += (double)cv[q] * uv[q];  // out of range: 0 * 0
#pragma unroll 4
  for (i64 e = b + threadIdx.x + 256 * ST_EPF; e < e1; e += 256) s += (double)__ldg(cval + e) * u[__ldg(ridx + e)];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();  // also publishes s_n2
  if (threadIdx.x == 0) {
    const double sinv = wf.W ? inv_or_zero(sqrt(s_n2[0])) : 1.0;
    double tt = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tt += red[w];
    t[col] = tt * sinv;
  }
  if (pc.world > 1) {  // the last column CTA to finish sums t over ranks (fused NVLink exchange)
    if (threadIdx.x == 0) {
      __threadfence();
      last_cta = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last_cta) {
      peer_allreduce_block(pc, t, h);
      if (threadIdx.x == 0) *counter = 0;
    }
  }
  TL_EXIT(128 + (wf.W ? wf.jn : 127));
}

// ---- F = A_s^T w - s V[:,j]  and  c = V[:, :nd]^T F  (irlb.py:182-190, first half of orthog) ----
__global__ void __launch_bounds__(RT_THREADS, 1) v_prep_dots(const double* __restrict__ t, const double* __restrict__ mu,
                                                   const double* __restrict__ rs, const double* __restrict__ sc,
                                                   const double* __restrict__ V, int ldv, int j, int nd, int h,
                                                   double* __restrict__ ybuf, double* __restrict__ partials,
                                                   double* __restrict__ cvec, unsigned* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const PeerComm nopeer;  // V is replicated: no exchange
  const double s = sc[SC_S], su = sc[SC_SU];
  double acc[4] = {0, 0, 0, 0};
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < h; r += warps) {
    const double* vr = V + (size_t)r * ldv;
    double f = t[r];
    if (mu) f = (f - mu[r] * su) * rs[r];
    f -= s * vr[j];
    if (lane == 0) ybuf[r] = f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = lane + 32 * q;
      if (col < nd) acc[q] += vr[col] * f;
    }
  }
  reduce_vec_blocks(acc, nd, partials, cvec, counter, nopeer);
}

// ---- K4  w = A_s x - fn W[:,jprev]  and  c = W[:, :nd]^T w  (irlb.py:165 / 204-216) -------------
// One warp per row: lowest latency, used while the shard has few rows per resident warp (n < 64 k).
template <typename VT>
__global__ void __launch_bounds__(RT_THREADS, 1) spmv_dots_w32(const i64* __restrict__ rptr, const int* __restrict__ cidx,
                                                 const VT* __restrict__ rval, const double* __restrict__ xs,
                                                 const double* __restrict__ sc, const double* __restrict__ W,
                                                 int ldw, int jprev, int nd, i64 n, double* __restrict__ ybuf,
                                                 double* __restrict__ partials, double* __restrict__ cvec,
                                                 unsigned* __restrict__ counter, const PeerComm pc) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  const double mvs = sc[SC_MVS], fn = sc[SC_FN];
  double acc[4] = {0, 0, 0, 0};
  // a warp takes ~4 rows one after the other and every row is a chain of dependent loads (row pointers ->
  // entries -> gather): the next row's pointers and this row's slice of W are requested before the chain starts
  i64 r = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  i64 b = 0, e1 = 0;
  if (r < n) {
    b = __ldg(rptr + r);
    e1 = __ldg(rptr + r + 1);
  }
  for (; r < n; r += warps) {
    const i64 rn = r + warps;
    i64 nb = 0, ne = 0;
    if (rn < n) {
      nb = __ldg(rptr + rn);
      ne = __ldg(rptr + rn + 1);
    }
    const double* wr = W + (size_t)r * ldw;
    double wv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wv[q] = (lane + 32 * q < nd) ? wr[lane + 32 * q] : 0.0;
    const double wp = (jprev >= 0) ? wr[jprev] : 0.0;
    double a = 0;
#pragma unroll 4
    for (i64 e = b + lane; e < e1; e += 32) a += (double)__ldg(rval + e) * xs[__ldg(cidx + e)];
    a = warp_sum(a);
    double w = a - mvs;
    if (jprev >= 0) w -= fn * wp;
    if (lane == 0) ybuf[r] = w;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += wv[q] * w;
    b = nb;
    e1 = ne;
  }
  if (nd > 0) reduce_vec_blocks(acc, nd, partials, cvec, counter, pc);
}

// Four rows per warp (8 lanes each), used for large shards: the subset holds ~50-100 non-zeros per cell, so a full warp per row
// leaves lanes idle and only one dependent load chain (row pointer -> entries -> gather) in flight per warp.
constexpr int SD_THREADS = 512;
template <typename VT, int NQ>  // NQ = ceil(nd / 8) column slots per lane (12: basis <= 96 wide, 16: <= 128)
__global__ void __launch_bounds__(SD_THREADS, 2) spmv_dots(const i64* __restrict__ rptr, const int* __restrict__ cidx,
                                                 const VT* __restrict__ rval, const double* __restrict__ xs,
                                                 const double* __restrict__ sc, const double* __restrict__ W,
                                                 int ldw, int jprev, int nd, i64 n, double* __restrict__ ybuf,
                                                 double* __restrict__ partials, double* __restrict__ cvec,
                                                 unsigned* __restrict__ counter, const PeerComm pc) {
  const int lane = threadIdx.x & 31, sub = lane >> 3, sl = lane & 7;
  const i64 wg = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  const double mvs = sc[SC_MVS], fn = sc[SC_FN];
  double acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
  for (i64 base = wg * 4; base < n; base += warps * 4) {  // warp-uniform trip count (shuffles inside)
    const i64 r = base + sub;
    const bool valid = r < n;
    double a = 0;
    if (valid) {
      const i64 b = rptr[r], e1 = rptr[r + 1];
#pragma unroll 4
      for (i64 e = b + sl; e < e1; e += 8) a += (double)__ldg(rval + e) * xs[__ldg(cidx + e)];
    }
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (valid) {
      const double* wr = W + (size_t)r * ldw;
      double w = a - mvs;
      if (jprev >= 0) w -= fn * wr[jprev];
      if (sl == 0) ybuf[r] = w;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int col = sl + 8 * q;
        if (col < nd) acc[q] = fma(wr[col], w, acc[q]);
      }
    }
  }
  if (nd > 0) {
    deposit_lane8<NQ>(acc);
    reduce_vec_finish(nd, partials, cvec, counter, pc);
  }
}

// ---- y -= Q[:, :nd] c ;  nrm2 = y.y ;  aux = coef.y  (second half of orthog + the norm) ---------
__global__ void __launch_bounds__(RT_THREADS, 1) update_norm_w32(const double* __restrict__ Q, int ld, int nd,
                                                   const double* __restrict__ cvec, double* __restrict__ ybuf,
                                                   const double* __restrict__ coef, i64 rows,
                                                   double* __restrict__ partials, double* __restrict__ out2,
                                                   unsigned* __restrict__ counter, const PeerComm pc, const Finish fin) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  double creg[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) creg[q] = (lane + 32 * q < nd) ? cvec[lane + 32 * q] : 0.0;
  double n2 = 0, ax = 0;
  // two rows per trip: their loads and the two warp reductions overlap (same summation order as one by one)
  for (i64 r = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < rows; r += 2 * warps) {
    const i64 r2 = r + warps;
    const bool two = r2 < rows;
    double y0 = ybuf[r], y1 = two ? ybuf[r2] : 0.0;
    const double c0 = coef ? coef[r] : 1.0, c1 = (coef && two) ? coef[r2] : 1.0;
    if (nd > 0) {
      const double* q0 = Q + (size_t)r * ld;
      const double* q1 = Q + (size_t)(two ? r2 : r) * ld;
      double p0 = 0, p1 = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = lane + 32 * q;
        if (col < nd) {
          p0 += q0[col] * creg[q];
          p1 += q1[col] * creg[q];
        }
      }
      p0 = warp_sum(p0);
      p1 = warp_sum(p1);
      y0 -= p0;
      y1 -= p1;
      if (lane == 0) {
        ybuf[r] = y0;
        if (two) ybuf[r2] = y1;
      }
    }
    n2 += y0 * y0;
    ax += c0 * y0;
    if (two) {
      n2 += y1 * y1;
      ax += c1 * y1;
    }
  }
  const bool last = reduce_two_blocks(n2, ax, partials, out2, counter, pc);
  if (last && fin.kind) finish_in_last_cta(fin, ybuf, rows);
}

// same, four rows per warp (8 lanes each): more loads in flight, used for large row counts
__global__ void __launch_bounds__(RT_THREADS, 1) update_norm(const double* __restrict__ Q, int ld, int nd,
                                                   const double* __restrict__ cvec, double* __restrict__ ybuf,
                                                   const double* __restrict__ coef, i64 rows,
                                                   double* __restrict__ partials, double* __restrict__ out2,
                                                   unsigned* __restrict__ counter, const PeerComm pc) {
  __shared__ double csh[MAXB];
  const int lane = threadIdx.x & 31, sub = lane >> 3, sl = lane & 7;
  const i64 wg = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  if (threadIdx.x < MAXB) csh[threadIdx.x] = (threadIdx.x < nd) ? cvec[threadIdx.x] : 0.0;
  __syncthreads();
  double n2 = 0, ax = 0;
  for (i64 base = wg * 4; base < rows; base += warps * 4) {  // four rows per warp, 8 lanes each
    const i64 r = base + sub;
    const bool valid = r < rows;
    double y = valid ? ybuf[r] : 0.0;
    if (nd > 0) {
      double p = 0;
      if (valid) {
        const double* qr = Q + (size_t)r * ld;
        for (int col = sl; col < nd; col += 8) p = fma(qr[col], csh[col], p);
      }
      p += __shfl_xor_sync(0xffffffffu, p, 4);
      p += __shfl_xor_sync(0xffffffffu, p, 2);
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      y -= p;
      if (valid && sl == 0) ybuf[r] = y;
    }
    if (valid && sl == 0) {
      n2 = fma(y, y, n2);
      ax = fma(coef ? coef[r] : 1.0, y, ax);
    }
  }
  // leaders of the four row groups -> lane 0
  n2 += __shfl_xor_sync(0xffffffffu, n2, 8);
  n2 += __shfl_xor_sync(0xffffffffu, n2, 16);
  ax += __shfl_xor_sync(0xffffffffu, ax, 8);
  ax += __shfl_xor_sync(0xffffffffu, ax, 16);
  reduce_two_blocks(n2, ax, partials, out2, counter, pc);
}

// ---- w_step: the whole W side of a Lanczos step in ONE pass over W ---------------------------------
// The reference forms w = A_s x - fn w_prev, THEN the Gram-Schmidt coefficients c = W^T w, THEN w - W c
// (irlb.py:204-216): two passes over W with a grid-wide reduction between them.  But
//     W^T (A_s x - fn w_prev) = Z^T xs - mvs (W^T 1) - fn (W^T w_prev),       Z = A^T W  (h x m_b),
// and every column of Z is the vector t = A^T w_j that spmvT_csc computes anyway for the V side, W^T 1 are the
// column sums the finish step already has, and W^T w_prev is a by-product of the pass that produced w_prev (g
// below).  So c is known BEFORE the pass (v_step computes it from Z, h ~ 1000 rows) and the W side is one kernel:
//     y_r = (A xs)_r - mvs - fn W[r, jprev] - sum_c W[r, c] c_c,
// with one reduction of the vector (g = W^T y, ||y||^2, sum y).  Same arithmetic as the reference up to rounding
// (no orthogonality of W is assumed: g carries the actual W^T w_prev); Z is rotated with W at a restart.
//
// Latency: a warp takes its rows one after the other and each row is a chain row pointers -> entries -> gather.
// The chain is software-pipelined two rows deep: while row i is reduced, the entries (first 32 EPF of them) and
// the W slice of row i + 1 and the pointers of row i + 2 are already in flight in registers.
//
// Programmatic dependent launch: everything above pdl_sync() reads only the static CSR arrays and columns of W
// that were written at least two kernels earlier (spmvT_csc of the previous step; every step kernel releases its
// dependents only AFTER its own grid-dependency wait, so "two kernels back" has completed).  The first row's
// whole load chain therefore overlaps the tail of v_step.  W is read with ld.global.cg: no line of it may sit
// in L1 across the dependency.
constexpr int W2_THREADS = 512;  // CTA of the one-pass W-side kernels

// ---- w_step3: w_step2 with a HALF-WARP per row --------------------------------------------------------
// A subset row holds 50-90 entries and 70-96 basis columns: a full warp per row leaves a third of its entry
// slots empty and pays a 5-level reduction for every single row.  Here the two halves of a warp take two
// consecutive rows at once -- the same instruction stream serves both (6 entry slots and NH basis columns per
// lane, a 4-level reduction) -- which nearly halves the issue slots per row; ncu had w_step2 at 173 instructions
// per row, 0.5 IPC per scheduler, stalled on its own dependent latencies rather than on memory.
constexpr int W3_EPF = 6;  // entry slots per lane: 96 per row in registers, the rest in a loop
template <typename VT, int NH>
struct W3Row {
  int ci[W3_EPF];
  VT va[W3_EPF];
  double wv[NH], wp;
  int b, len;
};
template <typename VT, int NH>
__device__ __forceinline__ void w3_load(W3Row<VT, NH>& R, bool ok, const i64* __restrict__ rptr, int row,
                                        const int* __restrict__ cl, const VT* __restrict__ vl,
                                        const double* __restrict__ wr, int jprev, const bool (&cm)[NH], int hl) {
  int b = 0, e1 = 0;
  if (ok) {
    b = (int)__ldg(rptr + row);
    e1 = (int)__ldg(rptr + row + 1);
  }
  R.b = b;
  R.len = e1 - b;
#pragma unroll
  for (int q = 0; q < W3_EPF; ++q) {
    const bool in = hl + 16 * q < R.len;
    R.ci[q] = in ? __ldg(cl + b + 16 * q) : 0;
    R.va[q] = in ? __ldg(vl + b + 16 * q) : VT(0);
  }
#pragma unroll
  for (int q = 0; q < NH; ++q) R.wv[q] = (ok && cm[q]) ? __ldcg(wr + 16 * q) : 0.0;
  R.wp = (ok && jprev >= 0) ? __ldcg(wr - hl + jprev) : 0.0;
}
// NH = basis columns per lane: 6 (<= 96 columns including the two scalars) or 8
template <typename VT, int NH>
__global__ void __launch_bounds__(W2_THREADS, 1)
    w_step3(const i64* __restrict__ rptr, const int* __restrict__ cidx, const VT* __restrict__ rval,
            const double* __restrict__ xs, const double* __restrict__ sc, const double* __restrict__ W, int ldw, int jprev,
            int nd, int n, const double* __restrict__ cvec, double* __restrict__ ybuf, double* __restrict__ partials,
            double* __restrict__ gout, unsigned* __restrict__ counter, const PeerComm pc, const int* __restrict__ halt,
            int h, int split) {
  extern __shared__ __align__(16) double xsm[];  // [h]
  const int lane = threadIdx.x & 31, hl = lane & 15, hf = lane >> 4, warp = threadIdx.x >> 5;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5), wg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int per = (n + nwarps - 1) / nwarps;
  per += per & 1;                       // even: a warp takes its rows two at a time
  const int r0 = min(n, wg * per), rend = min(n, r0 + per);
  int r = r0 + hf;                      // this half-warp's row; advances by 2
  TL_ENTRY_REG();
  // ---- prologue (independent of the previous kernel) ----
  bool cm[NH];
#pragma unroll
  for (int q = 0; q < NH; ++q) cm[q] = hl + 16 * q < nd;
  const int* cl = cidx + hl;
  const VT* vl = rval + hl;
  const double* wr = W + (size_t)r * ldw + hl;
  W3Row<VT, NH> RA, RB;
  w3_load<VT, NH>(RA, r < rend, rptr, r, cl, vl, wr, jprev, cm, hl);
  pdl_sync();
  if (halt && *halt) return;
  TL_SYNC(nd);
#ifdef ADOBO_B200_TIMELINE
  if (threadIdx.x == 0 && blockIdx.x < 256) tl_cta[0][blockIdx.x] = tl_e0;
#endif
  TL_CTA(1);
  for (int i = threadIdx.x; i < h; i += W2_THREADS) xsm[i] = xs[i];
  const double mvs = sc[SC_MVS], fn = (jprev >= 0) ? sc[SC_FN] : 0.0;
  double creg[NH], acc[NH];
#pragma unroll
  for (int q = 0; q < NH; ++q) {
    creg[q] = cm[q] ? cvec[hl + 16 * q] : 0.0;
    acc[q] = 0.0;
  }
  double n2 = 0, ax = 0;
  __syncthreads();
  // one pair of rows: issue the loads of the next pair (into NX), then reduce the current one (CU)
  auto step = [&](W3Row<VT, NH>& CU, W3Row<VT, NH>& NX) {
    wr += 2 * (size_t)ldw;
    w3_load<VT, NH>(NX, r + 2 < rend, rptr, r + 2, cl, vl, wr, jprev, cm, hl);
    double a = 0;
#pragma unroll
    for (int q = 0; q < W3_EPF; ++q)
      if (hl + 16 * q < CU.len) a += (double)CU.va[q] * xsm[CU.ci[q]];
    for (int e = 16 * W3_EPF + hl; e < CU.len; e += 16) a += (double)__ldg(vl + CU.b + e - hl) * xsm[__ldg(cl + CU.b + e - hl)];
#pragma unroll
    for (int q = 0; q < NH; ++q) a -= CU.wv[q] * creg[q];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);  // within the half-warp
    const bool ok = r < rend;
    const double w = ok ? (a - mvs) - fn * CU.wp : 0.0;
    if (hl == 0 && ok) ybuf[r] = w;
#pragma unroll
    for (int q = 0; q < NH; ++q) acc[q] += CU.wv[q] * w;
    n2 += w * w;
    ax += w;
    r += 2;
  };
  while (r - hf < rend) {  // warp-uniform: r - hf is the pair's first row
    step(RA, RB);
    if (r - hf >= rend) break;
    step(RB, RA);
  }
  TL_CTA(2);
  // the halves hold partial sums of the same columns (hl + 16 q): add them, then one half deposits
#pragma unroll
  for (int q = 0; q < NH; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 16);
  n2 += __shfl_xor_sync(0xffffffffu, n2, 16);
  ax += __shfl_xor_sync(0xffffffffu, ax, 16);
  {
    double(*sh)[MAXB] = vec_stage();
    if (hf == 0) {
#pragma unroll
      for (int q = 0; q < NH; ++q) {
        const int col = hl + 16 * q;
        double v = acc[q];
        if (col == nd) v = n2;        // the two scalars ride in columns nd, nd + 1 (w is uniform in a half-warp)
        if (col == nd + 1) v = ax;
        sh[warp][col] = v;
      }
    } else {
      for (int col = 16 * NH + hl; col < MAXB; col += 16) sh[warp][col] = 0.0;
    }
  }
  if (split)
    reduce_vec_partials(nd + 2, partials);
  else
    reduce_vec_finish(nd + 2, partials, gout, counter, pc);
  TL_CTA(7);
  TL_EXIT(nd);
}

// ---- v_step: the whole V side of a Lanczos step in ONE launch on a thread-block cluster -----------
// F = A_s^T w - s V[:, j]  (from t = A^T w),  c = V[:, :j+1]^T F,  F -= V c,  ||F||, mu.F,  V[:, j+1] = F / ||F||,
// x = F / (||F|| sigma)   (irlb.py:185-197, 203-221).  V has only h ~ 1000 rows: v_prep_dots + update_norm + v_finish
// spend their time in three launches and two cross-CTA reductions through global memory.  Here eight CTAs of a
// cluster take h/8 rows each and reduce through DISTRIBUTED SHARED MEMORY: every CTA pushes its partial
// dot products (then its two partial norms) into the shared memory of all eight, one cluster barrier, and every
// CTA sums the eight partials in the same order -- bitwise the same c and ||F|| everywhere, no atomics.
constexpr int VS_THREADS = 1024, VS_WARPS = VS_THREADS / 32, VS_CLUSTER = 8, VS_ROWS = 4;
static_assert(VS_THREADS == 8 * MAXB, "v_step sums its per-warp partials with 8 threads per column");

// Z formulation (see w_step): the kernel also stores Z[:, j] = t and leaves the Gram-Schmidt coefficients of
// the NEXT W column in cvec:  c = Z[:, :ncc]^T xs - mvs sumW - [prev] fn g,  g = W^T w_j from w_step's vector.
struct ZArgs {
  double* Z = nullptr;  // h x ldv; nullptr: formulation off
  double* sumW = nullptr;
  const double* gprev = nullptr;  // w_step output for column j: W^T y (j values), ||y||^2, sum y
  double* cvec = nullptr;
  int ncc = 0, prev = 0;
};
#ifdef ADOBO_B200_VSTAMPS  // phase timing of v_step (clock64 of CTA 0), compiled out of the product build
__device__ long long vs_stamps[16];
#define VSTAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) vs_stamps[i] = clock64(); } while (0)
#else
#define VSTAMP(i) do { } while (0)
#endif
template <int NQ>  // columns per lane: 3 (m_b <= 96) or 4
__global__ void __launch_bounds__(VS_THREADS, 1)
    v_step(const double* __restrict__ t, const double* __restrict__ mu, const double* __restrict__ rs,
           const double* __restrict__ murs, double* __restrict__ sc, double* __restrict__ V, int ldv, int j, int m_b, int h,
           double* __restrict__ ybuf, double* __restrict__ fbuf, double* __restrict__ xs, double* __restrict__ B,
           const ZArgs z, const int* __restrict__ halt) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  VSTAMP(0);
  pdl_sync();
  if (halt && *halt) return;  // uniform over the cluster
  cluster.barrier_arrive();  // matched below, before the first write into a neighbour's shared memory
  extern __shared__ __align__(16) double vs_dyn[];
  double(*wacc)[MAXB] = reinterpret_cast<double(*)[MAXB]>(vs_dyn);                    // [VS_WARPS][MAXB]
  double(*psum)[MAXB] = reinterpret_cast<double(*)[MAXB]>(vs_dyn + VS_WARPS * MAXB);  // [8][MAXB]
  __shared__ double call[VS_CLUSTER][MAXB];   // partial dot products of every CTA of the cluster
  __shared__ double call2[VS_CLUSTER][MAXB];  // partial Z^T (F / sigma) of every CTA
  __shared__ double ctot[MAXB];
  __shared__ double red2[VS_WARPS][2];
  __shared__ double nall[VS_CLUSTER][2];  // partial (||F||^2, mu.F) of every CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nd = j + 1;
  const int per = (h + csize - 1) / csize;
  const int r0 = crank * per, r1 = min(h, r0 + per);
  const double s = sc[SC_S], su = sc[SC_SU];
  // phase A: F = prep(t), partial c.  Rows go to warps in batches of VS_ROWS independent rows so that their
  // loads (and, in phase B, their warp reductions) are in flight together.
  double acc[NQ] = {};
  for (int rb = r0 + warp; rb < r1; rb += VS_WARPS * VS_ROWS) {
    double f[VS_ROWS], vj[VS_ROWS], vq[VS_ROWS][NQ];
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = rb + i * VS_WARPS;
      const bool ok = r < r1;
      const double* vr = V + (size_t)(ok ? r : r0) * ldv;
      f[i] = ok ? t[r] : 0.0;
      if (z.Z && ok && lane == 0) z.Z[(size_t)r * ldv + j] = f[i];  // Z[:, j] = A^T w_j
      if (mu) f[i] = ok ? (f[i] - mu[r] * su) * rs[r] : 0.0;
      vj[i] = ok ? vr[j] : 0.0;
#pragma unroll
      for (int q = 0; q < NQ; ++q) vq[i][q] = (ok && lane + 32 * q < nd) ? vr[lane + 32 * q] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = rb + i * VS_WARPS;
      f[i] -= s * vj[i];
      if (lane == 0 && r < r1) ybuf[r] = f[i];
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] += vq[i][q] * f[i];
    }
  }
  VSTAMP(1);
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (lane + 32 * q < MAXB) wacc[warp][lane + 32 * q] = acc[q];
  __syncthreads();
  VSTAMP(2);
  cluster.barrier_wait();  // every CTA of the cluster is running
  VSTAMP(3);
  // sum over the 32 warps in two levels (8 groups of 4 warps, then 8 partials) instead of 32 serial loads
  {
    const int col = tid & (MAXB - 1), part = tid >> 7;  // MAXB = 128 columns x 8 parts = 1024 threads
    double p = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) p += wacc[part * 4 + w][col];
    psum[part][col] = p;
  }
  __syncthreads();
  if (tid < nd) {
    double p = 0;
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) p += psum[g8][tid];
    for (int q = 0; q < csize; ++q) cluster.map_shared_rank(&call[0][0], q)[crank * MAXB + tid] = p;
  }
  VSTAMP(4);
  cluster.sync();
  VSTAMP(5);
  if (tid < nd) {
    double p = 0;
    for (int q = 0; q < csize; ++q) p += call[q][tid];
    ctot[tid] = p;
  }
  __syncthreads();
  VSTAMP(6);
  // phase B: F -= V c, partial ||F||^2 and mu.F
  double creg[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) creg[q] = (lane + 32 * q < nd) ? ctot[lane + 32 * q] : 0.0;
  double n2 = 0, ax = 0;
  double zacc[NQ] = {};
  for (int rb = r0 + warp; rb < r1; rb += VS_WARPS * VS_ROWS) {
    double p[VS_ROWS], y[VS_ROWS], cf[VS_ROWS];
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = rb + i * VS_WARPS;
      const bool ok = r < r1;
      const double* vr = V + (size_t)(ok ? r : r0) * ldv;
      p[i] = 0;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int col = lane + 32 * q;
        if (ok && col < nd) p[i] += vr[col] * creg[q];
      }
      y[i] = ok ? __ldcg(ybuf + r) : 0.0;
      cf[i] = (ok && murs) ? murs[r] : (ok ? 1.0 : 0.0);
    }
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) p[i] = warp_sum(p[i]);
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = rb + i * VS_WARPS;
      y[i] -= p[i];
      if (lane == 0 && r < r1) ybuf[r] = y[i];
      n2 += y[i] * y[i];
      ax += cf[i] * y[i];
    }
    if (z.Z) {
#pragma unroll
      for (int i = 0; i < VS_ROWS; ++i) {
        const int r = rb + i * VS_WARPS;
        if (r < r1) {
          const double* zr = z.Z + (size_t)r * ldv;
          const double yz = rs ? y[i] * rs[r] : y[i];
#pragma unroll
          for (int q = 0; q < NQ; ++q)
            if (lane + 32 * q < z.ncc) zacc[q] += zr[lane + 32 * q] * yz;
        }
      }
    }
  }
  VSTAMP(7);
  if (z.Z) {  // wacc is free again: every CTA passed the first cluster barrier after reading it
#pragma unroll
    for (int q = 0; q < NQ; ++q) wacc[warp][lane + 32 * q] = zacc[q];
  }
  if (lane == 0) {
    red2[warp][0] = n2;
    red2[warp][1] = ax;
  }
  __syncthreads();
  if (tid < 2) {
    double p = 0;
#pragma unroll 8
    for (int w = 0; w < VS_WARPS; ++w) p += red2[w][tid];
    for (int q = 0; q < csize; ++q) cluster.map_shared_rank(&nall[0][0], q)[crank * 2 + tid] = p;
  }
  if (z.Z) {
    {
      const int col = tid & (MAXB - 1), part = tid >> 7;
      double p = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) p += wacc[part * 4 + w][col];
      psum[part][col] = p;
    }
    __syncthreads();
    if (tid >= 32 && tid < 32 + z.ncc) {
      const int col = tid - 32;
      double p = 0;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) p += psum[g8][col];
      for (int q = 0; q < csize; ++q) cluster.map_shared_rank(&call2[0][0], q)[crank * MAXB + col] = p;
    }
  }
  VSTAMP(8);
  cluster.sync();  // also orders this CTA's ybuf writes before its own reads below
  VSTAMP(9);
  double tot0 = 0, tot1 = 0;
  for (int q = 0; q < csize; ++q) {
    tot0 += nall[q][0];
    tot1 += nall[q][1];
  }
  // finish: V[:, j+1] = F / ||F||, x = F / (||F|| sigma)
  const double nrm = sqrt(tot0);
  const double inv = inv_or_zero(nrm);
  const int jn = j + 1;
  for (int r = r0 + tid; r < r1; r += VS_THREADS) {
    const double v = inv * __ldcg(ybuf + r);
    fbuf[r] = v;
    if (jn < m_b) V[(size_t)r * ldv + jn] = v;
    xs[r] = rs ? v * rs[r] : v;
  }
  if (z.Z && crank == 0 && tid >= 32 && tid < 32 + z.ncc) {
    // coefficients of the next W column; column j itself: g = 1, sumW[j] = (sum y) / ||y|| of w_step's vector
    const int col = tid - 32;
    double zs = 0;
    for (int q = 0; q < csize; ++q) zs += call2[q][col];
    const double sinvp = inv_or_zero(sqrt(z.gprev[j]));
    const double sw = (col == j) ? z.gprev[j + 1] * sinvp : z.sumW[col];
    if (col == j) z.sumW[j] = sw;
    const double mvs = murs ? inv * tot1 : 0.0;
    double cc = inv * zs - mvs * sw;
    if (z.prev) cc -= nrm * ((col == j) ? 1.0 : z.gprev[col] * sinvp);
    z.cvec[col] = cc;
  }
  if (crank == 0 && tid == 0) {
    sc[SC_NRM2] = tot0;
    sc[SC_AUX] = tot1;
    sc[SC_FN] = nrm;
    sc[SC_FNINV] = inv;
    sc[SC_MVS] = murs ? inv * tot1 : 0.0;
    if (j < m_b - 1) B[(size_t)j * m_b + j + 1] = nrm;  // irlb.py:197
  }
#ifdef ADOBO_B200_VSTAMPS
  __syncthreads();
#endif
  VSTAMP(10);
}

// ---- v_step_one: the same step when a CTA's rows fit ONE batch (h <= 8 x 128) -----------------------
// With h ~ 1000 every warp owns at most VS_ROWS rows, so the rows of V stay in registers from the dot products
// (phase A) to the update (phase B) and F never goes through memory; the CTA's rows of Z are staged in shared
// memory with cp.async.  Both fetches are issued BEFORE the grid-dependency wait (V[:, :j+1] and Z[:, :j] were
// written by earlier v_steps, three or more kernels back -- see w_step for the rule), so that after the wait
// one round trip (t, mu, sigma) separates the kernel from its first reduction.  Arithmetic and summation orders
// are those of v_step: both give bit-identical results.
template <int NQ>
__global__ void __launch_bounds__(VS_THREADS, 1)
    v_step_one(const double* __restrict__ t, const double* __restrict__ mu, const double* __restrict__ rs,
               const double* __restrict__ murs, double* __restrict__ sc, double* __restrict__ V, int ldv, int j, int m_b,
               int h, double* __restrict__ ybuf, double* __restrict__ fbuf, double* __restrict__ xs,
               double* __restrict__ B, const ZArgs z, const int* __restrict__ halt) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  extern __shared__ __align__(16) double vs_dyn[];
  double(*wacc)[MAXB] = reinterpret_cast<double(*)[MAXB]>(vs_dyn);                    // [VS_WARPS][MAXB]
  double(*psum)[MAXB] = reinterpret_cast<double(*)[MAXB]>(vs_dyn + VS_WARPS * MAXB);  // [8][MAXB]
  double* zs = vs_dyn + (VS_WARPS + 8) * MAXB;                                        // [rows of this CTA][ldv]
  __shared__ double call[VS_CLUSTER][MAXB];
  __shared__ double call2[VS_CLUSTER][MAXB];
  __shared__ double ctot[MAXB];
  __shared__ double red2[VS_WARPS][2];
  __shared__ double nall[VS_CLUSTER][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nd = j + 1;
  const int per = (h + csize - 1) / csize;
  const int r0 = crank * per, r1 = min(h, r0 + per);
  TL_ENTRY_REG();
#ifdef ADOBO_B200_TIMELINE
  const bool tl_on = true;
#endif
  // ---- prologue: nothing here depends on the previous two kernels ----
  if (z.Z) {
    const int c2n = (z.ncc + 1) >> 1, nrow = max(0, r1 - r0);
    for (int idx = tid; idx < nrow * c2n; idx += VS_THREADS) {
      const int row = idx / c2n, c2 = idx - row * c2n;
      cp_async16_cg(zs + (size_t)row * ldv + 2 * c2, z.Z + (size_t)(r0 + row) * ldv + 2 * c2);
    }
    cp_async_commit();
  }
  double vq[VS_ROWS][NQ], vj[VS_ROWS];
#pragma unroll
  for (int i = 0; i < VS_ROWS; ++i) {
    const int r = r0 + warp + i * VS_WARPS;
    const bool ok = r < r1;
    const double* vr = V + (size_t)(ok ? r : 0) * ldv;
    vj[i] = ok ? __ldcg(vr + j) : 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) vq[i][q] = (ok && lane + 32 * q < nd) ? __ldcg(vr + lane + 32 * q) : 0.0;
  }
  pdl_sync();
  if (halt && *halt) {  // uniform over the cluster
    cp_async_wait_all();
    return;
  }
  TL_SYNC(256 + j);
  TL_CTAK(1, 0);
  cluster.barrier_arrive();  // matched below, before the first write into a neighbour's shared memory
  const double s = sc[SC_S], su = sc[SC_SU];
  // ---- phase A: F = prep(t) - s V[:, j], partial c = V^T F ----
  double f[VS_ROWS], rr[VS_ROWS], acc[NQ] = {};
  {
    double tv[VS_ROWS], mm[VS_ROWS];
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = r0 + warp + i * VS_WARPS;
      const bool ok = r < r1;
      tv[i] = ok ? t[r] : 0.0;
      mm[i] = (ok && mu) ? mu[r] : 0.0;
      rr[i] = (ok && rs) ? rs[r] : 1.0;
    }
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = r0 + warp + i * VS_WARPS;
      f[i] = tv[i];
      if (mu) f[i] = (r < r1) ? (f[i] - mm[i] * su) * rr[i] : 0.0;
      f[i] -= s * vj[i];
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] += vq[i][q] * f[i];
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (lane + 32 * q < MAXB) wacc[warp][lane + 32 * q] = acc[q];
  TL_CTAK(1, 1);
  cp_async_wait_all();
  __syncthreads();  // the staged rows of Z have landed (every thread waited for its own copies)
  TL_CTAK(1, 2);
  if (z.Z && lane == 0) {  // Z[:, j] = A^T w_j, in memory and in the staged copy
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = r0 + warp + i * VS_WARPS;
      if (r < r1) {
        const double tv = t[r];
        z.Z[(size_t)r * ldv + j] = tv;
        zs[(size_t)(r - r0) * ldv + j] = tv;
      }
    }
  }
  cluster.barrier_wait();  // every CTA of the cluster is running
  {
    const int col = tid & (MAXB - 1), part = tid >> 7;
    double p = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) p += wacc[part * 4 + w][col];
    psum[part][col] = p;
  }
  __syncthreads();
  if (tid < nd) {
    double p = 0;
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) p += psum[g8][tid];
    for (int q = 0; q < csize; ++q) cluster.map_shared_rank(&call[0][0], q)[crank * MAXB + tid] = p;
  }
  TL_CTAK(1, 3);
  cluster.sync();
  TL_CTAK(1, 4);
  if (tid < nd) {
    double p = 0;
    for (int q = 0; q < csize; ++q) p += call[q][tid];
    ctot[tid] = p;
  }
  __syncthreads();
  // ---- phase B: F -= V c from the registers, partial ||F||^2, mu.F and Z^T (F / sigma) ----
  double y[VS_ROWS];
  {
    double creg[NQ], p[VS_ROWS];
#pragma unroll
    for (int q = 0; q < NQ; ++q) creg[q] = (lane + 32 * q < nd) ? ctot[lane + 32 * q] : 0.0;
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      p[i] = 0;
#pragma unroll
      for (int q = 0; q < NQ; ++q) p[i] += vq[i][q] * creg[q];
    }
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) p[i] = warp_sum(p[i]);
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) y[i] = f[i] - p[i];
  }
  double n2 = 0, ax = 0, zacc[NQ] = {};
#pragma unroll
  for (int i = 0; i < VS_ROWS; ++i) {
    const int r = r0 + warp + i * VS_WARPS;
    const bool ok = r < r1;
    const double cf = (ok && murs) ? murs[r] : (ok ? 1.0 : 0.0);
    if (lane == 0 && ok) ybuf[r] = y[i];
    n2 += y[i] * y[i];
    ax += cf * y[i];
    if (z.Z && ok) {
      const double* zr = zs + (size_t)(r - r0) * ldv;
      const double yz = rs ? y[i] * rr[i] : y[i];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (lane + 32 * q < z.ncc) zacc[q] += zr[lane + 32 * q] * yz;
    }
  }
  if (z.Z) {  // wacc is free again: every CTA passed the first cluster barrier after reading it
#pragma unroll
    for (int q = 0; q < NQ; ++q) wacc[warp][lane + 32 * q] = zacc[q];
  }
  if (lane == 0) {
    red2[warp][0] = n2;
    red2[warp][1] = ax;
  }
  __syncthreads();
  if (warp == 0) {  // the 32 warp partials of (||F||^2, mu.F): one butterfly each instead of two serial chains
    const double p0 = warp_sum(red2[lane][0]), p1 = warp_sum(red2[lane][1]);
    if (lane < 2) {
      const double p = lane == 0 ? p0 : p1;
      for (int q = 0; q < csize; ++q) cluster.map_shared_rank(&nall[0][0], q)[crank * 2 + lane] = p;
    }
  }
  if (z.Z) {
    {
      const int col = tid & (MAXB - 1), part = tid >> 7;
      double p = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) p += wacc[part * 4 + w][col];
      psum[part][col] = p;
    }
    __syncthreads();
    if (tid >= 32 && tid < 32 + z.ncc) {
      const int col = tid - 32;
      double p = 0;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) p += psum[g8][col];
      for (int q = 0; q < csize; ++q) cluster.map_shared_rank(&call2[0][0], q)[crank * MAXB + col] = p;
    }
  }
  // inputs of the coefficient update are fetched before the barrier (w_step / spmvT_csc wrote them, not this kernel)
  const bool ccthread = z.Z && crank == 0 && tid >= 32 && tid < 32 + z.ncc;
  double g_j = 0, g_j1 = 0, g_col = 0, sw_col = 0;
  if (ccthread) {
    g_j = z.gprev[j], g_j1 = z.gprev[j + 1];
    if (tid - 32 != j) g_col = z.gprev[tid - 32], sw_col = z.sumW[tid - 32];
  }
  TL_CTAK(1, 5);
  cluster.sync();
  TL_CTAK(1, 6);
  double tot0 = (lane < csize) ? nall[lane][0] : 0.0, tot1 = (lane < csize) ? nall[lane][1] : 0.0;
  tot0 = warp_sum(tot0);  // fixed butterfly, identical in every warp of every CTA
  tot1 = warp_sum(tot1);
  // finish from the registers: V[:, j+1] = F / ||F||, x = F / (||F|| sigma)
  const double nrm = sqrt(tot0);
  const double inv = inv_or_zero(nrm);
  const int jn = j + 1;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = r0 + warp + i * VS_WARPS;
      if (r < r1) {
        const double v = inv * y[i];
        fbuf[r] = v;
        if (jn < m_b) V[(size_t)r * ldv + jn] = v;
        xs[r] = rs ? v * rr[i] : v;
      }
    }
  }
  if (ccthread) {
    const int col = tid - 32;
    double zsum = 0;
    for (int q = 0; q < csize; ++q) zsum += call2[q][col];
    const double sinvp = inv_or_zero(sqrt(g_j));
    const double sw = (col == j) ? g_j1 * sinvp : sw_col;
    if (col == j) z.sumW[j] = sw;
    const double mvs = murs ? inv * tot1 : 0.0;
    double cc = inv * zsum - mvs * sw;
    if (z.prev) cc -= nrm * ((col == j) ? 1.0 : g_col * sinvp);
    z.cvec[col] = cc;
  }
  if (crank == 0 && tid == 0) {
    sc[SC_NRM2] = tot0;
    sc[SC_AUX] = tot1;
    sc[SC_FN] = nrm;
    sc[SC_FNINV] = inv;
    sc[SC_MVS] = murs ? inv * tot1 : 0.0;
    if (j < m_b - 1) B[(size_t)j * m_b + j + 1] = nrm;  // irlb.py:197
  }
  TL_CTAK(1, 7);
  TL_EXIT(256 + j);
}

// ---- normalise F, store V[:, jn], prepare x = F / sigma and mu.x, write B[j,j], B[j,j+1] ---------
__global__ void v_finish(const double* __restrict__ ybuf, const double* __restrict__ rs, double* __restrict__ sc,
                         double* __restrict__ V, int ldv, int jn, int m_b, int h, double* __restrict__ fbuf,
                         double* __restrict__ xs, double* __restrict__ B, int j, int centered) {
  const double fn = sqrt(sc[SC_NRM2]);
  const double fninv = inv_or_zero(fn);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < h) {
    const double f = fninv * ybuf[r];
    fbuf[r] = f;
    if (jn < m_b) V[(size_t)r * ldv + jn] = f;
    xs[r] = rs ? f * rs[r] : f;
  }
  if (r == 0) {
    sc[SC_FN] = fn;
    sc[SC_FNINV] = fninv;
    sc[SC_MVS] = centered ? fninv * sc[SC_AUX] : 0.0;
    if (j >= 0 && j < m_b - 1) B[(size_t)j * m_b + j + 1] = fn;  // irlb.py:197
  }
}

// ---- normalise w, store W[:, jn] and the contiguous copy the next A^T u gathers from -----------
// B[jn, jn] = ||w|| is complete here, so the SVD of B can start while the last F is still being computed
__global__ void w_finish(const double* __restrict__ ybuf, const double* __restrict__ n2p, double* __restrict__ sc,
                         double* __restrict__ W, int ldw, int jn, i64 n, double* __restrict__ wcur, double* __restrict__ B,
                         int m_b, const int* __restrict__ halt, const double* __restrict__ parts, int nblk,
                         double* __restrict__ gout) {
  if (halt && *halt) return;
  __shared__ double s_n2[2];
  if (parts) {  // split reduction: columns jn, jn + 1 of w_step's per-CTA partials
    if (threadIdx.x < 32) {
      double t0, t1;
      sum_partials_two(parts, nblk, jn, t0, t1);
      if (threadIdx.x == 0) {
        s_n2[0] = t0, s_n2[1] = t1;
        if (blockIdx.x == 0) gout[jn] = t0, gout[jn + 1] = t1;  // v_step reads them for sumW[jn]
      }
    }
    __syncthreads();
  }
  const double n2v = parts ? s_n2[0] : n2p[0], syv = parts ? s_n2[1] : n2p[1];
  const double s = sqrt(n2v);
  const double sinv = inv_or_zero(s);
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) {
    const double w = sinv * ybuf[r];
    W[(size_t)r * ldw + jn] = w;
    wcur[r] = w;
  }
  if (r == 0) {
    sc[SC_S] = s;
    sc[SC_SU] = sinv * syv;
    B[(size_t)jn * m_b + jn] = s;  // irlb.py:196 / 221
  }
  TL_ENTRY(387);
  TL_EXIT(387);
}

#include "irlb_step.cuh"

// ---- one-sided Jacobi SVD of B (m x m, row-major, m <= 128) in one CTA ---------------------------
// G = B, column-major in shared memory (column stride LD = 16 R rows, zero padded; an extra zero
// column when m is odd).  Columns are orthogonalised by plane rotations until B Vr = U diag(sv).
// The kernel is bound by the issue rate and the shared-memory bandwidth of ONE SM -- a round of the
// classic parallel ordering streams all of G through the register file for one rotation per column
// -- so the ordering is BLOCKED: columns are grouped in blocks of two, the round-robin tournament
// runs over blocks, and the 16 lanes that own a block pair (P, Q) load the four columns once
// (128-bit LDS, bank-conflict free), apply the four cross rotations (a0,b0) (a1,b1) (a0,b1) (a1,b0)
// in registers and store them once: half the shared-memory traffic and half the barriers per
// sweep.  The intra-block pairs (a0,a1) are rotated at the start of each sweep.  Further trims: no
// bounds predicates (padding), cached column norms (refreshed exactly every sweep), rotation angle
// in fp32 with an exactly orthogonal fp64 (c, s), no accumulation of the right vectors:
// Vr = B^T U diag(1/sv) is one small GEMM at the end (accurate for the leading columns, the only
// ones the restart uses: k <= m - 3).  Outputs column-major, singular values descending.
constexpr int JL = 16;        // lanes per block pair
constexpr int JTHREADS = 512; // 32 groups >= 128 / 4 block pairs

// two independent rotations (x0,y0) and (x1,y1), interleaved; columns are H double2 per lane, the
// cached squared norms are updated in place; returns the OR of the convergence flags
template <int H>
__device__ __forceinline__ int jacobi_rotate2(double2 (&x0)[H], double2 (&y0)[H], double& nx0, double& ny0,
                                              double2 (&x1)[H], double2 (&y1)[H], double& nx1, double& ny1) {
  double p0 = 0, p1 = 0, q0 = 0, q1 = 0;
#pragma unroll
  for (int r = 0; r < H; ++r) {
    p0 = fma(x0[r].x, y0[r].x, p0);
    p1 = fma(x0[r].y, y0[r].y, p1);
    q0 = fma(x1[r].x, y1[r].x, q0);
    q1 = fma(x1[r].y, y1[r].y, q1);
  }
  double ga = p0 + p1, gb = q0 + q1;
#pragma unroll
  for (int o = JL / 2; o > 0; o >>= 1) {
    ga += __shfl_xor_sync(0xffffffffu, ga, o);
    gb += __shfl_xor_sync(0xffffffffu, gb, o);
  }
  // late sweeps: most pairs are already orthogonal -- when no pair of this warp needs a rotation the
  // whole warp skips the angle / update / store work (warp-uniform branch, no divergence)
  const bool need = (ga * ga > 1e-30 * (nx0 * ny0)) || (gb * gb > 1e-30 * (nx1 * ny1));
  if (!__any_sync(0xffffffffu, need)) return 0;
  double ca, sa, cb, sb;
  const int fa = jacobi_angle(ga, nx0, ny0, ca, sa);
  const int fb = jacobi_angle(gb, nx1, ny1, cb, sb);
#pragma unroll
  for (int r = 0; r < H; ++r) {
    const double2 a = x0[r], b = y0[r], c = x1[r], d = y1[r];
    x0[r].x = ca * a.x - sa * b.x;
    x0[r].y = ca * a.y - sa * b.y;
    y0[r].x = sa * a.x + ca * b.x;
    y0[r].y = sa * a.y + ca * b.y;
    x1[r].x = cb * c.x - sb * d.x;
    x1[r].y = cb * c.y - sb * d.y;
    y1[r].x = sb * c.x + cb * d.x;
    y1[r].y = sb * c.y + cb * d.y;
  }
  {
    const double c2 = ca * ca, s2 = sa * sa, xx = 2.0 * ca * sa * ga;
    const double n = c2 * nx0 - xx + s2 * ny0;
    ny0 = s2 * nx0 + xx + c2 * ny0;
    nx0 = n;
  }
  {
    const double c2 = cb * cb, s2 = sb * sb, xx = 2.0 * cb * sb * gb;
    const double n = c2 * nx1 - xx + s2 * ny1;
    ny1 = s2 * nx1 + xx + c2 * ny1;
    nx1 = n;
  }
  return fa | fb;
}

template <int R>  // rows per lane (even); LD = JL * R >= m
__global__ void __launch_bounds__(JTHREADS, 1) jacobi_svd(const double* __restrict__ B, int m, double* __restrict__ Su,
                                                          double* __restrict__ svout, double* __restrict__ Sv,
                                                          int* __restrict__ info) {
  constexpr int LD = JL * R;
  constexpr int H = R / 2;  // double2 per lane and column
  constexpr int NGRP = JTHREADS / JL;
  extern __shared__ __align__(16) double smem[];
  const int nb = (m + 1) >> 1;             // column blocks (the last one is half empty when m is odd)
  const int mc = 2 * nb;                   // columns held (column m is all zero when m is odd)
  double* G = smem;                        // [mc][LD] column-major, rows >= m are zero
  double* nrm = smem + (size_t)mc * LD;    // [mc] squared column norms
  double* inv = nrm + 130;                 // [m] 1 / sv
  __shared__ int rotated;
  __shared__ int rankof[128];
  const int tid = threadIdx.x;
  const int grp = tid / JL, sl = tid % JL;
  if (info[4] != 0) return;  // halt flag (see conv_check)
  for (int idx = tid; idx < mc * LD; idx += blockDim.x) {
    const int col = idx / LD, row = idx - col * LD;
    G[idx] = (row < m && col < m) ? B[(size_t)row * m + col] : 0.0;
  }
  __syncthreads();
  const int nbp = nb + (nb & 1);           // players of the block tournament (dummy block when nb is odd)
  const int nbpairs = nbp >> 1;            // <= 32 = NGRP
  const bool warp_live = ((tid >> 5) * (32 / JL)) < nbpairs;
  int sweeps = 0;
  for (; sweeps < 60; ++sweeps) {
    // exact column norms at the start of every sweep
    for (int base = 0; base < mc; base += NGRP) {  // warp-uniform trip count (shuffles inside)
      const int col = base + grp;
      double sq = 0;
      if (col < mc) {
        const double* g = G + (size_t)col * LD;
#pragma unroll
        for (int r = 0; r < R; ++r) sq = fma(g[sl + JL * r], g[sl + JL * r], sq);
      }
#pragma unroll
      for (int o = JL / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (col < mc && sl == 0) nrm[col] = sq;
    }
    if (tid == 0) rotated = 0;
    __syncthreads();
    int flags = 0;
    // ---- pairs inside a block: each group takes two blocks per pass ----
    for (int base = 0; base < nb; base += 2 * NGRP) {  // warp-uniform
      const int blk0 = base + 2 * grp, blk1 = blk0 + 1;
      const bool act0 = blk0 < nb, act1 = blk1 < nb;
      double2* c0 = reinterpret_cast<double2*>(G + (size_t)(act0 ? 2 * blk0 : 0) * LD) + sl;
      double2* c1 = c0 + LD / 2;
      double2* d0 = reinterpret_cast<double2*>(G + (size_t)(act1 ? 2 * blk1 : 0) * LD) + sl;
      double2* d1 = d0 + LD / 2;
      double2 x0[H], y0[H], x1[H], y1[H];
#pragma unroll
      for (int r = 0; r < H; ++r) {
        x0[r] = act0 ? c0[JL * r] : make_double2(0.0, 0.0);  // idle slots rotate zeros: no flags
        y0[r] = act0 ? c1[JL * r] : make_double2(0.0, 0.0);
        x1[r] = act1 ? d0[JL * r] : make_double2(0.0, 0.0);
        y1[r] = act1 ? d1[JL * r] : make_double2(0.0, 0.0);
      }
      double nx0 = act0 ? nrm[2 * blk0] : 0.0, ny0 = act0 ? nrm[2 * blk0 + 1] : 0.0;
      double nx1 = act1 ? nrm[2 * blk1] : 0.0, ny1 = act1 ? nrm[2 * blk1 + 1] : 0.0;
      const int f = jacobi_rotate2<H>(x0, y0, nx0, ny0, x1, y1, nx1, ny1);
      if (act0) {
#pragma unroll
        for (int r = 0; r < H; ++r) {
          c0[JL * r] = x0[r];
          c1[JL * r] = y0[r];
        }
        if (sl == 0) {
          nrm[2 * blk0] = nx0;
          nrm[2 * blk0 + 1] = ny0;
        }
      }
      if (act1) {
#pragma unroll
        for (int r = 0; r < H; ++r) {
          d0[JL * r] = x1[r];
          d1[JL * r] = y1[r];
        }
        if (sl == 0) {
          nrm[2 * blk1] = nx1;
          nrm[2 * blk1 + 1] = ny1;
        }
      }
      flags |= f;
    }
    __syncthreads();
    // ---- round-robin over blocks, kept incrementally: slot i of round rd holds block
    // 1 + (i-1-rd) mod (nbp-1) (slot 0 holds block 0), pairs are slots (g, nbp-1-g); the next round
    // moves every block >= 1 one step down the circle (1 wraps to nbp-1).
    int pa = (grp < nbpairs) ? grp : 0, pb = (grp < nbpairs) ? nbp - 1 - grp : 1;
    for (int rd = 0; rd < nbp - 1; ++rd) {
      if (warp_live) {
        const int P = min(pa, pb), Q = max(pa, pb);
        const bool act = (grp < nbpairs) && (Q < nb);
        const int Qs = act ? Q : P;  // idle groups read block P twice and store nothing
        double2* a0p = reinterpret_cast<double2*>(G + (size_t)(2 * P) * LD) + sl;
        double2* a1p = a0p + LD / 2;
        double2* b0p = reinterpret_cast<double2*>(G + (size_t)(2 * Qs) * LD) + sl;
        double2* b1p = b0p + LD / 2;
        double2 a0[H], a1[H], b0[H], b1[H];
#pragma unroll
        for (int r = 0; r < H; ++r) {
          a0[r] = a0p[JL * r];
          a1[r] = a1p[JL * r];
          b0[r] = act ? b0p[JL * r] : make_double2(0.0, 0.0);  // idle groups rotate against zeros: no-op
          b1[r] = act ? b1p[JL * r] : make_double2(0.0, 0.0);
        }
        double na0 = nrm[2 * P], na1 = nrm[2 * P + 1], nb0 = nrm[2 * Qs], nb1 = nrm[2 * Qs + 1];
        if (!act) nb0 = nb1 = 0.0;  // zero norm -> no rotation
        int f = jacobi_rotate2<H>(a0, b0, na0, nb0, a1, b1, na1, nb1);
        f |= jacobi_rotate2<H>(a0, b1, na0, nb1, a1, b0, na1, nb0);
        if (act && f) {
#pragma unroll
          for (int r = 0; r < H; ++r) {
            a0p[JL * r] = a0[r];
            a1p[JL * r] = a1[r];
            b0p[JL * r] = b0[r];
            b1p[JL * r] = b1[r];
          }
          if (sl == 0) {
            nrm[2 * P] = na0;
            nrm[2 * P + 1] = na1;
            nrm[2 * Q] = nb0;
            nrm[2 * Q + 1] = nb1;
          }
          flags |= f;
        }
        if (pa > 0) pa = (pa == 1) ? nbp - 1 : pa - 1;
        pb = (pb == 1) ? nbp - 1 : pb - 1;
      }
      __syncthreads();
    }
    if (sl == 0 && flags) atomicOr(&rotated, flags);
    __syncthreads();
    const int all_flags = rotated;
    __syncthreads();
    // no rotation at all, or every cosine seen in this sweep was below 1e-8 (quadratic convergence:
    // the rotations of this sweep already took them below 1e-15)
    if ((all_flags & 2) == 0) {
      ++sweeps;
      break;
    }
  }
  // singular values = exact column norms; rank by descending value (ties: column order)
  for (int base = 0; base < m; base += NGRP) {
    const int col = base + grp;
    double sq = 0;
    if (col < m) {
      const double* g = G + (size_t)col * LD;
#pragma unroll
      for (int r = 0; r < R; ++r) sq = fma(g[sl + JL * r], g[sl + JL * r], sq);
    }
#pragma unroll
    for (int o = JL / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (col < m && sl == 0) nrm[col] = sqrt(sq);
  }
  __syncthreads();
  for (int col = tid; col < m; col += blockDim.x) {
    const double sv = nrm[col];
    int rank = 0;
    for (int o = 0; o < m; ++o) rank += (nrm[o] > sv) || (nrm[o] == sv && o < col);
    rankof[col] = rank;
    inv[col] = sv > 0 ? 1.0 / sv : 0.0;
    svout[rank] = sv;
  }
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += blockDim.x) {  // U = G diag(1/sv), kept in shared memory too
    const int col = idx / m, row = idx - col * m;
    const double u = G[(size_t)col * LD + row] * inv[col];
    G[(size_t)col * LD + row] = u;
    Su[(size_t)rankof[col] * m + row] = u;
  }
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += blockDim.x) {  // Vr = B^T U diag(1/sv)
    const int col = idx / m, i = idx - col * m;
    const double* u = G + (size_t)col * LD;
    double a = 0;
#pragma unroll 4
    for (int r = 0; r < m; ++r) a = fma(__ldg(B + (size_t)r * m + i), u[r], a);
    Sv[(size_t)rankof[col] * m + i] = a * inv[col];
  }
  if (tid == 0) info[2] = sweeps;
}

typedef void (*jacobi_fn)(const double*, int, double*, double*, double*, int*);
static int jacobi_rows_per_lane(int m) { return 2 * ((m + 2 * JL - 1) / (2 * JL)); }  // even R with 16 R >= m
static jacobi_fn jacobi_for(int m) {
  switch (jacobi_rows_per_lane(m)) {
    case 2: return jacobi_svd<2>;
    case 4: return jacobi_svd<4>;
    case 6: return jacobi_svd<6>;
    default: return jacobi_svd<8>;
  }
}

// ---- residuals, convergence count and the next k (irlb.py:225-234) ------------------------------
// info[0] = converged values (-1: svd_arrow rejected its own result), info[1] = next k, info[4] = halt flag
// (1 converged, 2 SVD rejected: every kernel of a later, speculatively launched iteration returns at once),
// info[5] = outer iterations completed.  `force`: run although the halt flag is set (host fallback path).
// With St != nullptr all CTAs also lay out the two rotation matrices of the coming restart (transpose_s, for
// the k this kernel has just decided), so the restart starts with the rotation itself.
// The host reads info without a copy node or an event in the stream: thread 0 of CTA 0 publishes the words the host
// needs into the context's page-locked mailbox (device-visible host memory).  Every word travels as ONE 8-byte store
// {value, version} -- the version (info[7], advanced per publication) makes the set self-validating, so no system
// fence sits in the kernel (two fences were 2 us of conv_check): the reader takes the five words when their versions
// agree.  A D2H copy node + event record between two iteration graphs cost ~8 us per restart at 68 k cells.
constexpr int INFO_PUB = 5;
__constant__ int info_pub_field[INFO_PUB] = {0, 1, 4, 5, 6};
__device__ __forceinline__ void publish_info(int* __restrict__ info, unsigned long long* __restrict__ hinfo) {
  if (hinfo == nullptr) return;
  const unsigned ver = (unsigned)info[7] + 1u;
#pragma unroll
  for (int i = 0; i < INFO_PUB; ++i) {
    const unsigned long long w = ((unsigned long long)ver << 32) | (unsigned)info[info_pub_field[i]];
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(hinfo + i), "l"(w) : "memory");
  }
  info[7] = (int)ver;
}

__global__ void conv_check(const double* __restrict__ Su, const double* __restrict__ sv, double* __restrict__ sc,
                           int m, int nu, int k, int it, double tol, double* __restrict__ R, int* __restrict__ info,
                           int honour_flag, int force, const double* __restrict__ Sv, double* __restrict__ St,
                           const double* __restrict__ cvec, const double* __restrict__ sumW, double* __restrict__ crot,
                           unsigned long long* __restrict__ hinfo) {
  __shared__ int cnt, s_kn;
  if (!force && info[4] != 0) return;
  TL_ENTRY(386);
  if (honour_flag && info[3] == 0) {  // svd_arrow rejected its own result: the host re-runs jacobi_svd
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      info[0] = -1;
      info[1] = k;
      info[4] = 2;
      publish_info(info, hinfo);
    }
    return;
  }
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const double fn = sc[SC_FN];
  double smax = sv[0];
  if (it > 0) smax = fmax(smax, sc[SC_SMAX]);  // CTA 0 may already have stored the new maximum: same value
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double r = fn * Su[(size_t)i * m + (m - 1)];
    if (blockIdx.x == 0) R[i] = r;
    if (i < nu && fabs(r) < tol * smax) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int kn = max(cnt + nu, k);
    kn = min(kn, m - 3);
    s_kn = kn;
    if (blockIdx.x == 0) {
      sc[SC_SMAX] = smax;
      info[0] = cnt;
      info[1] = kn;
      info[4] = (cnt >= nu) ? 1 : 0;
      info[5] += 1;
      publish_info(info, hinfo);
    }
  }
  if (St == nullptr) return;
  __syncthreads();
  const int kc = s_kn, kcp = (kc + 7) & ~7, per = m * kcp;
  if (cvec && blockIdx.x > 0) {
    // Z formulation (see w_step): the pre-computed coefficients and the column sums of W in the rotated basis,
    // c <- Su[:, :k]^T c, sumW <- Su[:, :k]^T sumW (Su[c * m + i] = component i of left vector c); a warp per
    // output, into crot[0..MAXB) / crot[MAXB..) -- set_restart copies them in place after the rotation
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int q = ((int)blockIdx.x - 1) * nw + (threadIdx.x >> 5); q < kc; q += ((int)gridDim.x - 1) * nw) {
      double nc = 0, ns = 0;
#pragma unroll
      for (int ee = 0; ee < MAXB / 32; ++ee) {
        const int e = lane + 32 * ee;
        if (e < m) {
          const double u = Su[(size_t)q * m + e];
          nc = fma(u, cvec[e], nc);
          ns = fma(u, sumW[e], ns);
        }
      }
      nc = warp_sum(nc);
      ns = warp_sum(ns);
      if (lane == 0) {
        crot[q] = nc;
        crot[MAXB + q] = ns;
      }
    }
  }
  TL_EXIT(386);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * per; idx += gridDim.x * blockDim.x) {
    const double* S = (idx < per) ? Sv : Su;
    const int e = (idx < per) ? idx : idx - per;
    const int i = e / kcp, c = e - i * kcp;
    St[idx] = (c < kc) ? S[(size_t)c * m + i] : 0.0;
  }
}

// ---- Q[:, :kc] <- Q[:, :m] S[:, :kc]  (irlb.py:238, 246, 249-250); S column-major m x m ----------
// Tall-skinny fp64 GEMM, row-local, in place.  transpose_s lays S out as [i][kcp] (kcp = kc rounded
// up to 8, zero filled) once; rotate_basis keeps it in shared memory and streams 64-row tiles of Q
// through a double buffer filled with 8-byte cp.async (transposed to [i][row]: conflict-free column
// reads).  Warp w owns output columns 8w .. 8w+7, each lane two rows of the tile: per inner step
// 2 shared loads of Q + 4 broadcast LDS.128 of S feed 16 DFMA, so the loop runs at the SM's fp64
// rate rather than its shared-memory rate.  blockDim = 32 * kcp / 8.
constexpr int RB_ROWS = 64;
constexpr int RB_QP = RB_ROWS + 1;

// both rotation matrices of a restart (Sv for V, Su for W) in one launch: St0 | St1
__global__ void transpose_s(const double* __restrict__ S0, const double* __restrict__ S1, int m, int kc, int kcp,
                            double* __restrict__ St) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = m * kcp;
  if (idx >= 2 * per) return;
  const double* S = (idx < per) ? S0 : S1;
  double* dst = St + ((idx < per) ? 0 : per);
  if (idx >= per) idx -= per;
  const int i = idx / kcp, c = idx - i * kcp;
  dst[idx] = (c < kc) ? S[(size_t)c * m + i] : 0.0;
}

struct RotJob {
  double* Q;               // rows x ld, row-major
  i64 rows;
  const double* St;        // [m][kcp] transposed rotation
  double* out;             // nullptr: in place into Q[:, :kc]
  int ldo;
  const double* colscale;  // optional per-column factor (var_weigh)
};


// The rotations of a restart -- V (h rows), Z = A^T W (h rows, optional) and W (n rows) -- run as ONE launch: CTAs
// [0, nba) take job a, [nba, nba + nbb) job b, the rest job c, so the small jobs do not occupy a nearly empty GPU.
__global__ void __launch_bounds__(512) rotate_basis(const RotJob ja, const RotJob jb, const RotJob jc, int nba, int nbb,
                                                    int ld, int m, int kc, int nbuf, const int* __restrict__ halt) {
  if (halt && *halt) return;
  const int which = ((int)blockIdx.x >= nba) + ((int)blockIdx.x >= nba + nbb);
  const RotJob& job = which == 0 ? ja : (which == 1 ? jb : jc);
  double* __restrict__ Q = job.Q;
  const i64 rows = job.rows;
  const double* __restrict__ St = job.St;
  double* __restrict__ out = job.out;
  const int ldo = job.ldo;
  const double* __restrict__ colscale = job.colscale;
  const unsigned bid = which == 0 ? blockIdx.x : (which == 1 ? blockIdx.x - nba : blockIdx.x - nba - nbb);
  const unsigned nblk = which == 0 ? nba : (which == 1 ? nbb : gridDim.x - nba - nbb);
  extern __shared__ __align__(16) double sm[];
  const int kcp = (kc + 7) & ~7;
  double* Ssm = sm;                   // [m][kcp]
  double* Qs = sm + (size_t)m * kcp;  // [nbuf][m][RB_QP]
  const size_t qstride = (size_t)m * RB_QP;
  for (int idx = threadIdx.x; idx < m * kcp; idx += blockDim.x) Ssm[idx] = St[idx];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const i64 ntiles = (rows + RB_ROWS - 1) / RB_ROWS;
  auto fetch = [&](i64 t, int buf) {
    const i64 row0 = t * RB_ROWS;
    double* dst = Qs + buf * qstride;
    for (int idx = threadIdx.x; idx < RB_ROWS * m; idx += blockDim.x) {
      const int r = idx / m, i = idx - r * m;
      if (row0 + r < rows) cp_async8(dst + (size_t)i * RB_QP + r, Q + (size_t)(row0 + r) * ld + i);
    }
    cp_async_commit();
  };
  int buf = 0;
  if ((i64)bid < ntiles) fetch(bid, 0);
  for (i64 t = bid; t < ntiles; t += nblk) {
    cp_async_wait_all();
    __syncthreads();  // tile t landed (and Ssm written); everybody is done with the other buffer
    const i64 tn = t + nblk;
    if (nbuf == 2 && tn < ntiles) fetch(tn, buf ^ 1);
    const i64 row0 = t * RB_ROWS;
    const double* qp = Qs + buf * qstride + lane;
    const double* sp = Ssm + warp * 8;
    double a0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
    for (int i = 0; i < m; ++i) {
      const double x0 = qp[(size_t)i * RB_QP], x1 = qp[(size_t)i * RB_QP + 32];
      const double2 s0 = *reinterpret_cast<const double2*>(sp + (size_t)i * kcp);
      const double2 s1 = *reinterpret_cast<const double2*>(sp + (size_t)i * kcp + 2);
      const double2 s2 = *reinterpret_cast<const double2*>(sp + (size_t)i * kcp + 4);
      const double2 s3 = *reinterpret_cast<const double2*>(sp + (size_t)i * kcp + 6);
      a0[0] = fma(x0, s0.x, a0[0]);
      a0[1] = fma(x0, s0.y, a0[1]);
      a0[2] = fma(x0, s1.x, a0[2]);
      a0[3] = fma(x0, s1.y, a0[3]);
      a0[4] = fma(x0, s2.x, a0[4]);
      a0[5] = fma(x0, s2.y, a0[5]);
      a0[6] = fma(x0, s3.x, a0[6]);
      a0[7] = fma(x0, s3.y, a0[7]);
      a1[0] = fma(x1, s0.x, a1[0]);
      a1[1] = fma(x1, s0.y, a1[1]);
      a1[2] = fma(x1, s1.x, a1[2]);
      a1[3] = fma(x1, s1.y, a1[3]);
      a1[4] = fma(x1, s2.x, a1[4]);
      a1[5] = fma(x1, s2.y, a1[5]);
      a1[6] = fma(x1, s3.x, a1[6]);
      a1[7] = fma(x1, s3.y, a1[7]);
    }
    if (nbuf == 1) {
      __syncthreads();  // single buffer: everybody is done reading before it is refilled
      if (tn < ntiles) fetch(tn, 0);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const i64 r = row0 + lane + 32 * half;
      if (r < rows) {
        double* dst = out ? out + (size_t)r * ldo : Q + (size_t)r * ld;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = warp * 8 + c;
          const double v = half ? a1[c] : a0[c];
          if (col < kc) dst[col] = colscale ? v * colscale[col] : v;
        }
      }
    }
    if (nbuf == 2) buf ^= 1;
  }
  cp_async_wait_all();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- rotate_mma: the same product on the fp64 tensor-core path (mma.sync m8n8k4) -------------------------------------
// The SIMT kernels above feed 16 DFMA from 6 shared-memory loads per inner step and keep the fp64 pipe ~29 % busy
// (rotate_basis, ncu) -- the shared-memory pipe and instruction issue are what they wait for, and a tile of 64 rows is
// the least a CTA can take (14 us for the 1000-row V and Z of every restart: 32 CTAs on 148 SMs).  DMMA has the same
// peak as DFMA on this part (64 FMA/clk/SM, profiles/micro) but does 256 FMAs per instruction from two operand
// registers: a warp tile of WM x WN 8 x 8 blocks needs WM + WN loads for 256 WM WN FMAs, and the tile height is a
// free multiple of 8 rows.  Q arrives ROW-MAJOR through 16-byte cp.async (no transposing 8-byte copies); both operand
// tiles have a row stride = 4 (mod 16) doubles, so the 8 x 4 / 4 x 8 fragment loads are free of bank conflicts.
//   <MB, WM, WN> = <8, 4, 2>: 64-row tiles streamed through a double buffer (long shards: W)
//   <MB, WM, WN> = <2, 2, 1>: 16-row tiles, one per CTA (V and Z of a restart on all SMs)
__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
template <int MB, int WM, int WN>
__global__ void __launch_bounds__(512) rotate_mma(const RotJob ja, const RotJob jb, const RotJob jc, int nba, int nbb,
                                                  int ld, int m, int kc, int nbuf, const int* __restrict__ halt) {
  // launched with a full dependency on its predecessor: a kernel behind it with a programmatic edge (lazy_p0, which
  // touches nothing this kernel writes until its own dependency wait) may start at once
  cudaTriggerProgrammaticLaunchCompletion();
  if (halt && *halt) return;
  TL_ENTRY(384);
  constexpr int TR = 8 * MB, NWM = MB / WM;
  const int which = ((int)blockIdx.x >= nba) + ((int)blockIdx.x >= nba + nbb);
  const RotJob& job = which == 0 ? ja : (which == 1 ? jb : jc);
  double* __restrict__ Q = job.Q;
  const i64 rows = job.rows;
  const double* __restrict__ St = job.St;
  double* __restrict__ out = job.out;
  const int ldo = job.ldo;
  const double* __restrict__ colscale = job.colscale;
  const unsigned bid = which == 0 ? blockIdx.x : (which == 1 ? blockIdx.x - nba : blockIdx.x - nba - nbb);
  const unsigned nblk = which == 0 ? nba : (which == 1 ? nbb : gridDim.x - nba - nbb);
  extern __shared__ __align__(16) double sm[];
  const int kcp = (kc + 7) & ~7, m4 = (m + 3) & ~3;
  const int ldp = kcp + 4, ldq = ((m + 7) & ~7) + 4;
  double* Ssm = sm;                       // [m4][ldp], rows >= m zero
  double* Qs = sm + (size_t)m4 * ldp;     // [nbuf][TR][ldq]
  const size_t qstride = (size_t)TR * ldq;
  const int nthr = (int)blockDim.x;
  for (int idx = threadIdx.x; idx < m4 * (kcp >> 1); idx += nthr) {
    const int i = idx / (kcp >> 1), c2 = idx - i * (kcp >> 1);
    double2 v = make_double2(0.0, 0.0);
    if (i < m) v = *reinterpret_cast<const double2*>(St + (size_t)i * kcp + 2 * c2);
    *reinterpret_cast<double2*>(Ssm + (size_t)i * ldp + 2 * c2) = v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp % NWM, wn = warp / NWM;
  const i64 ntiles = (rows + TR - 1) / TR;
  // the columns m .. m4 - 1 of a tile are whatever the buffer holds there (finite: every basis is zero-filled when the
  // decomposition starts): they meet the zero rows of Ssm
  auto fetch = [&](i64 t, int buf) {
    const i64 row0 = t * TR;
    double* dst = Qs + buf * qstride;
    const int nch = m4 >> 1;
    for (int idx = threadIdx.x; idx < TR * nch; idx += nthr) {
      const int r = idx / nch, c2 = idx - r * nch;
      double* d = dst + (size_t)r * ldq + 2 * c2;
      if (row0 + r < rows) cp_async16(d, Q + (size_t)(row0 + r) * ld + 2 * c2);
      else *reinterpret_cast<double2*>(d) = make_double2(0.0, 0.0);
    }
    cp_async_commit();
  };
  int buf = 0;
  if ((i64)bid < ntiles) fetch(bid, 0);
  const int ar = lane >> 2, ak = lane & 3;   // fragment coordinates: A[row ar][k ak], B[k ak][col ar], C[row ar][cols 2 ak, 2 ak + 1]
  for (i64 t = bid; t < ntiles; t += nblk) {
    cp_async_wait_all();
    __syncthreads();  // tile t landed (and Ssm written); everybody is done with the other buffer
    const i64 tn = t + nblk;
    if (nbuf == 2 && tn < ntiles) fetch(tn, buf ^ 1);
    const double* qa = Qs + buf * qstride + (size_t)(8 * wm * WM + ar) * ldq + ak;
    const double* sb = Ssm + (size_t)ak * ldp + 8 * wn * WN + ar;
    double acc[WM][WN][2];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    bool nbok[WN];
#pragma unroll
    for (int j = 0; j < WN; ++j) nbok[j] = 8 * (wn * WN + j) < kcp;
#pragma unroll 2
    for (int k0 = 0; k0 < m4; k0 += 4) {
      double a[WM], b[WN];
#pragma unroll
      for (int i = 0; i < WM; ++i) a[i] = qa[(size_t)(8 * i) * ldq + k0];
#pragma unroll
      for (int j = 0; j < WN; ++j) b[j] = nbok[j] ? sb[(size_t)k0 * ldp + 8 * j] : 0.0;
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma_884(acc[i][j], a[i], b[j]);
    }
    if (nbuf == 1) {
      __syncthreads();  // single buffer: everybody is done reading before it is refilled
      if (tn < ntiles) fetch(tn, 0);
    }
    const i64 row0 = t * TR;
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const i64 r = row0 + 8 * (wm * WM + i) + ar;
      if (r < rows) {
        double* dst = out ? out + (size_t)r * ldo : Q + (size_t)r * ld;
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          const int col = 8 * (wn * WN + j) + 2 * ak;
          double v0 = acc[i][j][0], v1 = acc[i][j][1];
          if (colscale) {
            if (col < kc) v0 *= colscale[col];
            if (col + 1 < kc) v1 *= colscale[col + 1];
          }
          if (col + 1 < kc) *reinterpret_cast<double2*>(dst + col) = make_double2(v0, v1);
          else if (col < kc) dst[col] = v0;
        }
      }
    }
    if (nbuf == 2) buf ^= 1;
  }
  cp_async_wait_all();
}

// ---- V[:,k] = F ; B = diag(sv[:k]) + R[:k] in column k  (irlb.py:239-244) ------------------------
// With the Z formulation (see w_step) CTA 0 also installs the pre-computed coefficients and the column sums of W
// in the rotated basis (computed by conv_check).
__global__ void set_restart(double* __restrict__ V, int ldv, int h, const double* __restrict__ fbuf,
                            double* __restrict__ B, int m, int k, const double* __restrict__ sv,
                            const double* __restrict__ R, double* __restrict__ cvec, double* __restrict__ sumW,
                            const double* __restrict__ crot, const int* __restrict__ halt, const LazyArgs lz, int nds) {
  if (halt && *halt) return;
  if (cvec && blockIdx.x == 0) {  // rotated by the conv_check of the previous iteration
    const int t = threadIdx.x;
    if (t < k) sumW[t] = crot[MAXB + t];
    if (lz.P0 == nullptr) {
      if (t < k) cvec[t] = crot[t];
    } else if (t >= lz.U0 && t < nds) {  // lazy rotation: c_s = [P0 c[:kp] ; c[kp:]] -- lazy_p0 has left the first part
      cvec[t] = crot[lz.kp + t - lz.U0];
    }
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < h) V[(size_t)i * ldv + k] = fbuf[i];
  if (i < m * m) {
    const int r = i / m, c = i - r * m;
    double v = 0.0;
    if (r < k && c == r) v = sv[r];
    if (r < k && c == k) v = R[r];
    B[i] = v;
  }
  TL_ENTRY(385);
  TL_EXIT(385);
}

__global__ void mul_vec(const double* a, const double* b, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

static unsigned row_blocks(adobo_ctx* c, i64 rows) {  // rotate_basis: 64-row tiles, at most one CTA per SM
  i64 b = (rows + RB_ROWS - 1) / RB_ROWS;
  i64 cap = (i64)c->sm_count;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

static unsigned red_blocks(adobo_ctx* c, i64 rows) {  // 1024-thread CTAs of the reducing row kernels
  i64 b = (rows + RT_WARPS - 1) / RT_WARPS;
  if (b > c->sm_count) b = c->sm_count;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// Launch with optional attributes: thread-block cluster, programmatic dependent launch (the kernel may start
// while the previous kernel of the stream drains; it calls pdl_sync() before touching that kernel's output).
template <typename... KA, typename... A>
static cudaError_t launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster,
                            A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

#ifdef ADOBO_B200_TIMELINE
// diagnostic build: the timeline of the last replay of the iteration graph (CTA 0's stamps per kernel and column)
static void dump_timeline() {
  static unsigned long long tl[6][1024];
  cudaMemcpyFromSymbol(tl, tl_buf, sizeof(tl));
  struct Ev { unsigned long long t0, t1, t2; char name[24]; };
  std::vector<Ev> evs;
  auto add = [&](int slot, const char* nm, int j) {
    if (!tl[0][slot]) return;
    Ev e{tl[0][slot], tl[1][slot], tl[2][slot], {}};
    snprintf(e.name, sizeof(e.name), "%s[%d]", nm, j);
    evs.push_back(e);
  };
  for (int j = 0; j < 128; ++j) add(j, "w_step", j), add(128 + j, "spmvT", j), add(256 + j, "v_step", j);
  add(384, "rotate", 0), add(385, "set_restart", 0), add(386, "conv_check", 0), add(387, "w_finish", 0);
  std::sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.t0 < b.t0; });
  const size_t from = evs.size() > 24 ? evs.size() - 24 : 0;
  const unsigned long long base = evs.empty() ? 0 : evs[from].t0;
  {
    static unsigned long long ck[2][10][256];
    cudaMemcpyFromSymbol(ck, tl_ctak, sizeof(ck));
    static const char* nm[10] = {"sync", "phase A done", "Z landed", "c pushed", "cluster sync 1", "phase B pushed", "cluster sync 2",
                                 "exit", "slice summed", "exchange done"};
    static const int order[10] = {0, 8, 9, 1, 2, 3, 4, 5, 6, 7};
    fprintf(stderr, "last v_step_t, stamps of the cluster's 8 CTAs relative to the earliest sync return (us): min / max\n");
    unsigned long long b0 = ~0ull;
    for (int cta = 0; cta < 8; ++cta)
      if (ck[1][0][cta]) b0 = std::min(b0, ck[1][0][cta]);
    for (int oi = 0; oi < 10; ++oi) {
      const int i = order[oi];
      double lo = 1e30, hi = -1e30;
      for (int cta = 0; cta < 8; ++cta)
        if (ck[1][i][cta]) lo = std::min(lo, (ck[1][i][cta] - b0) * 1e-3), hi = std::max(hi, (ck[1][i][cta] - b0) * 1e-3);
      if (hi > -1e29) fprintf(stderr, "  %-16s %8.2f %8.2f\n", nm[i], lo, hi);
    }
  }
  fprintf(stderr, "timeline of the last kernels (us from the first, %%globaltimer)\n");
  for (size_t q = from; q < evs.size(); ++q)
    fprintf(stderr, "%-16s entry %8.2f sync %8.2f exit %8.2f\n", evs[q].name, (evs[q].t0 - base) * 1e-3,
            evs[q].t1 ? (evs[q].t1 - base) * 1e-3 : -1.0, (evs[q].t2 - base) * 1e-3);
}
#endif

// Test / diagnosis hooks of the driver: ONE environment variable, ADOBO_B200_IRLB="hook,hook,...".  Every hook
// selects a fallback that exists for a reason (a size the fast path does not cover, or the synchronous form of an
// asynchronous mechanism); tests/test_gpu_parity.py::test_irlb_alternative_paths runs each of them.
enum : unsigned {
  HK_SVD_JACOBI = 1u << 0,       // generic Jacobi SVD of B for every restart (default: first iteration + fallback)
  HK_SVD_FALLBACK = 1u << 1,     // exercise the fallback that runs when svd_arrow rejects its own result
  HK_NO_GRAPH = 1u << 2,         // launch kernel by kernel instead of replaying CUDA graphs
  HK_NO_FORK = 1u << 3,          // SVD of B in the main stream instead of beside the last step
  HK_NO_FUSEDT = 1u << 5,        // three-kernel step (w_step3 + spmvT_csc + v_step*): the path for h > 1024
  HK_NO_VONE = 1u << 6,          // v_step instead of v_step_one: the path for 1024 < h <= 8192
  HK_NO_ZPATH = 1u << 7,         // two-pass Gram-Schmidt kernels: the path for h > 8192 or m_b > 126
  HK_NO_VSTEP = 1u << 8,         // grid-wide V-side kernels: the path for h > 8192
  HK_NO_WFOLD = 1u << 9,         // separate w_finish launch for every column
  HK_NO_SPLIT = 1u << 10,        // last-CTA reduction of w_step3's vector (what several GPUs use on that path)
  HK_NO_PDL = 1u << 11,          // no programmatic dependent launch
  HK_NO_RUNAHEAD = 1u << 12,     // wait for every convergence read-back
  HK_TRACE = 1u << 13,           // one line per outer iteration on stderr
  HK_NO_LAZY = 1u << 14,         // rotate W at every restart (two-kernel step without the lazy rotation)
  HK_ROTATE_SIMT = 1u << 4,      // DFMA rotation kernel (rotate_basis) instead of the tensor-core one
};
static unsigned irlb_hooks() {
  static const struct { const char* name; unsigned bit; } names[] = {
      {"svd_jacobi", HK_SVD_JACOBI}, {"svd_fallback", HK_SVD_FALLBACK}, {"no_graph", HK_NO_GRAPH},
      {"no_fork", HK_NO_FORK}, {"rotate_simt", HK_ROTATE_SIMT}, {"no_fusedt", HK_NO_FUSEDT},
      {"no_vone", HK_NO_VONE}, {"no_zpath", HK_NO_ZPATH}, {"no_vstep", HK_NO_VSTEP}, {"no_wfold", HK_NO_WFOLD},
      {"no_split", HK_NO_SPLIT}, {"no_pdl", HK_NO_PDL}, {"no_runahead", HK_NO_RUNAHEAD}, {"trace", HK_TRACE},
      {"no_lazy", HK_NO_LAZY}};
  const char* e = getenv("ADOBO_B200_IRLB");
  unsigned bits = 0;
  if (!e) return 0;
  std::string s(e);
  size_t pos = 0;
  while (pos <= s.size()) {
    size_t end = s.find(',', pos);
    if (end == std::string::npos) end = s.size();
    const std::string tok = s.substr(pos, end - pos);
    bool known = tok.empty();
    for (auto& nm : names)
      if (tok == nm.name) bits |= nm.bit, known = true;
    REQUIRE(known, ADOBO_ERR_ARG, "ADOBO_B200_IRLB: unknown hook '%s'", tok.c_str());
    pos = end + 1;
  }
  return bits;
}

// Workspace of the Lanczos driver.  It lives in the context across calls so that (a) repeated PCA
// calls on same-shaped data do not re-allocate and (b) the CUDA graphs of one outer iteration --
// keyed by (first iteration?, k) -- stay valid and are replayed: one graph launch and one 32-byte
// read-back per restart instead of ~70 kernel launches.
struct IrlbWs {
  const void* sig_ptr[8] = {};
  i64 sig_n = -1;
  int sig_i[6] = {};
  double sig_tol = 0;
  DevBuf<double> V, W, B, sc, tbuf, ybuf_h, ybuf_n, fbuf, xs, wcur, cvec, partials, tpart, Su, Sv, svd_s, R, murs, St, Z, sumW, gout, crot;
  DevBuf<double> P0, P0t, Pmat;  // lazy rotation of W (irlb_step.cuh)
  DevBuf<unsigned> counter;
  DevBuf<int> info;
  std::map<i64, cudaGraphExec_t> graphs;  // key: first iteration?, k, state of the lazy rotation
  std::map<i64, i64> graph_launches;
  void clear_graphs() {
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear();
    graph_launches.clear();
  }
  ~IrlbWs() { clear_graphs(); }
};
static void free_irlb_ws(void* p) { delete (IrlbWs*)p; }

template <typename VT>
IrlbResult run_irlb(adobo_ctx* c, SparsePair<VT>& A, const double* mu, const double* rs, int nu, double tol,
                    int maxit, const double* v0_host, bool var_weigh, DevBuf<double>& outU, DevBuf<double>& outS,
                    DevBuf<double>& outV) {
  const i64 n = A.rows;
  const int h = A.cols;
  const i64 n_glob = (c->world > 1) ? c->n_total : n;
  REQUIRE(std::min<i64>(n_glob, h) >= 2, ADOBO_ERR_ARG, "The input matrix must be at least 2x2.");  // irlb.py:116
  REQUIRE(nu >= 1, ADOBO_ERR_ARG, "lanczos: nval must be >= 1");
  const int m_b = std::min(std::min(nu + 20, 3 * nu), h);  // irlb.py:138
  REQUIRE(m_b <= MAXB, ADOBO_ERR_LIMIT, "lanczos: working dimension %d exceeds %d (ncomp too large)", m_b, MAXB);
  REQUIRE(nu <= m_b, ADOBO_ERR_ARG, "lanczos: nval %d exceeds the working dimension %d", nu, m_b);
  REQUIRE(m_b >= 4, ADOBO_ERR_ARG, "lanczos: working dimension %d too small (need >= 4 columns)", m_b);
  REQUIRE(n < (i64(1) << 31) && A.nnz < (i64(1) << 31), ADOBO_ERR_LIMIT, "lanczos: shard too large for 32-bit offsets");
  const int ld = (m_b + 15) & ~15;
  const bool centered = (mu != nullptr);
  cudaStream_t st = c->stream;
  const unsigned hooks = irlb_hooks();
  auto hook = [&](unsigned bit) { return (hooks & bit) != 0; };

  if (!c->irlb_ws) {
    c->irlb_ws = new IrlbWs();
    c->irlb_ws_free = free_irlb_ws;
  }
  IrlbWs& ws = *(IrlbWs*)c->irlb_ws;
  {
    const void* sp[8] = {A.rptr.p, A.cidx.p, A.rval.p, A.cptr.p, A.ridx.p, A.cval.p, mu, rs};
    // hooks change what a captured iteration contains
    const int si[6] = {h, nu, m_b, c->world, (int)sizeof(VT), (int)(hooks | ((unsigned)svd_arrow_cluster() << 20))};
    bool same = ws.sig_n == n && ws.sig_tol == tol;
    for (int q = 0; q < 8; ++q) same = same && ws.sig_ptr[q] == sp[q];
    for (int q = 0; q < 6; ++q) same = same && ws.sig_i[q] == si[q];
    if (!same) {
      ws.clear_graphs();
      for (int q = 0; q < 8; ++q) ws.sig_ptr[q] = sp[q];
      for (int q = 0; q < 6; ++q) ws.sig_i[q] = si[q];
      ws.sig_n = n;
      ws.sig_tol = tol;
    }
  }
  DevBuf<double>&V = ws.V, &W = ws.W, &B = ws.B, &sc = ws.sc, &tbuf = ws.tbuf, &ybuf_h = ws.ybuf_h, &ybuf_n = ws.ybuf_n,
  &fbuf = ws.fbuf, &xs = ws.xs, &wcur = ws.wcur, &cvec = ws.cvec, &partials = ws.partials, &Su = ws.Su, &Sv = ws.Sv,
  &svd_s = ws.svd_s, &R = ws.R, &murs = ws.murs, &Z = ws.Z, &sumW = ws.sumW, &gout = ws.gout, &tpart = ws.tpart;
  DevBuf<unsigned>& counter = ws.counter;
  DevBuf<int>& info = ws.info;

  // ---- which step kernels ----
  const bool pdl = !hook(HK_NO_PDL);  // programmatic dependent launch of the step kernels
  const bool fused = c->world > 1 && c->peer.world == c->world;  // exchanges inside the kernels (peer mailboxes)
  const bool fused_t = fused && h <= PEER_SLOT;
  // Z formulation of the W side (one pass over W, see w_step3): needs the cluster V kernel and two spare columns
  const bool use_vstep = h <= 8192 && !hook(HK_NO_VSTEP);
  const bool zpath = use_vstep && m_b + 2 <= MAXB && !hook(HK_NO_ZPATH);
  // one batch of rows per warp and the CTA's rows of Z staged in shared memory: v_step_one / v_step_t
  const int vs_per = (h + VS_CLUSTER - 1) / VS_CLUSTER;
  const size_t vs_one_smem = sizeof(double) * ((size_t)(VS_WARPS + 8) * MAXB + (zpath ? (size_t)vs_per * ld : 0));
  const bool vs_one = use_vstep && vs_per <= VS_WARPS * VS_ROWS && vs_one_smem <= 200 * 1024 && !hook(HK_NO_VONE);
  // the two-kernel step (irlb_step.cuh): every BASELINE configuration (h = 1000) on one or several GPUs
  const int hp = (h + 15) & ~15;
  const bool fusedt = zpath && vs_one && h <= WT_HMAX && (c->world == 1 || fused) && !hook(HK_NO_FUSEDT);
  // three-kernel step: split reduction of w_step3's vector (single GPU: the consuming spmvT_csc sums the partials)
  const bool fold_wt_ok = !hook(HK_NO_WFOLD);
  const bool split = !fusedt && c->world == 1 && zpath && vs_one && fold_wt_ok && !hook(HK_NO_SPLIT);
  const bool wide_rows = n >= 65536;  // two-pass kernels: enough rows per resident warp to prefer throughput
  // lazy rotation of W (irlb_step.cuh): W_s holds up to Ms stored columns, 16 per basis slot of w_step_t's lanes, two
  // slots short because the row kernel's vector carries ||y||^2 and sum y behind g_s
  const int wt_nh = m_b + 1 <= 80 ? 6 : (m_b + 1 <= 96 ? 7 : 8);
  const bool lazy = fusedt && !hook(HK_NO_LAZY) && 16 * wt_nh - 2 - m_b >= 3;
  const int wt_nh_used = lazy ? wt_nh : (m_b + 1 <= 96 ? 6 : 8);
  const int Ms = lazy ? 16 * wt_nh - 2 : m_b;
  const int ldw = lazy ? 16 * wt_nh : ld;               // row stride of W / W_s
  const int ldp = (m_b + 7) & ~7;                       // row stride of P0
  {
    // a re-allocation moves a buffer: graphs that captured the old pointer must go
    const double* before[6] = {V.p, W.p, partials.p, murs.p, Z.p, tpart.p};
    V.ensure((size_t)h * ld);
    W.ensure((size_t)std::max<i64>(n, 1) * ldw);
    B.ensure((size_t)MAXB * MAXB);
    sc.ensure(SC_COUNT);
    tbuf.ensure((size_t)h);
    ybuf_h.ensure((size_t)h);
    ybuf_n.ensure((size_t)std::max<i64>(n, 1));
    fbuf.ensure((size_t)h);
    xs.ensure((size_t)h);
    wcur.ensure((size_t)std::max<i64>(n, 1));
    cvec.ensure(MAXB);
    partials.ensure((size_t)2 * c->sm_count * MAXB);
    Su.ensure((size_t)MAXB * MAXB);
    Sv.ensure((size_t)MAXB * MAXB);
    svd_s.ensure((size_t)MAXB);
    R.ensure((size_t)MAXB);
    counter.ensure(2);
    info.ensure(8);
    murs.ensure((size_t)h);
    ws.St.ensure((size_t)2 * MAXB * MAXB);
    ws.crot.ensure((size_t)2 * MAXB);
    if (zpath) {
      Z.ensure((size_t)h * ld);
      sumW.ensure(MAXB);
      gout.ensure(MAXB);
    }
    if (fusedt) tpart.ensure((size_t)c->sm_count * hp);
    if (lazy) {
      ws.P0.ensure((size_t)MAXB * MAXB);
      ws.P0t.ensure((size_t)MAXB * MAXB);
      ws.Pmat.ensure((size_t)MAXB * MAXB);
    }
    const double* after[6] = {V.p, W.p, partials.p, murs.p, Z.p, tpart.p};
    for (int q = 0; q < 6; ++q)
      if (before[q] != after[q]) ws.clear_graphs();
  }
  DevBuf<double>& St = ws.St;
  const unsigned gb_n = red_blocks(c, n), gb_h = red_blocks(c, h);
  // spmv_dots (wide): 512-thread CTAs, two per SM, four rows per warp
  const unsigned gb_sd = (unsigned)std::max<i64>(1, std::min<i64>((n + 63) / 64, 2 * (i64)c->sm_count));
  CUDA_CHECK(cudaMemsetAsync(V.p, 0, sizeof(double) * (size_t)h * ld, st));
  CUDA_CHECK(cudaMemsetAsync(W.p, 0, sizeof(double) * (size_t)std::max<i64>(n, 1) * ldw, st));
  CUDA_CHECK(cudaMemsetAsync(B.p, 0, sizeof(double) * m_b * m_b, st));
  CUDA_CHECK(cudaMemsetAsync(sc.p, 0, sizeof(double) * SC_COUNT, st));
  CUDA_CHECK(cudaMemsetAsync(counter.p, 0, sizeof(unsigned) * 2, st));
  CUDA_CHECK(cudaMemsetAsync(cvec.p, 0, sizeof(double) * MAXB, st));
  CUDA_CHECK(cudaMemsetAsync(info.p, 0, sizeof(int) * 8, st));
  if (zpath) {
    CUDA_CHECK(cudaMemsetAsync(Z.p, 0, sizeof(double) * (size_t)h * ld, st));
    CUDA_CHECK(cudaMemsetAsync(sumW.p, 0, sizeof(double) * MAXB, st));
    CUDA_CHECK(cudaMemsetAsync(gout.p, 0, sizeof(double) * MAXB, st));
  }
  if (centered) LAUNCH(c, "mul_vec", (mul_vec<<<(h + 255) / 256, 256, 0, st>>>(mu, rs, murs.p, h)));
  // SVD kernel resources
  const size_t svd_smem = sizeof(double) * ((size_t)(m_b + 1) * JL * jacobi_rows_per_lane(m_b) + 260);
  jacobi_fn jacobi = jacobi_for(m_b);
  CUDA_CHECK(cudaFuncSetAttribute(jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svd_smem));
  // rotate_basis: S tile + one or two Q tiles in shared memory (two when they fit in 227 KB)
  const size_t rot_s = sizeof(double) * (size_t)m_b * ((m_b + 7) & ~7), rot_q = sizeof(double) * (size_t)m_b * RB_QP;
  const int rot_nbuf = (rot_s + 2 * rot_q <= 225 * 1024) ? 2 : 1;
  const size_t rot_smem = rot_s + rot_nbuf * rot_q;
  CUDA_CHECK(cudaFuncSetAttribute(rotate_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot_smem));
  // W-side kernels of the one-pass formulations; under programmatic launch their CTAs start while the V-side
  // cluster still holds VS_CLUSTER SMs: a CTA that had to wait for one of those would run its whole prologue after
  // everybody else's -- leave them out of the grid
  auto w3_fn = m_b + 1 <= 96 ? w_step3<VT, 6> : w_step3<VT, 8>;
  // entry slots per lane of w_step_t: enough for a typical row (1.5 x the mean length), the rest of a long row loops
  const double mean_len = (double)A.nnz / (double)std::max<i64>(n, 1);
  const int wt_epf = mean_len * 1.5 <= 32 ? 2 : (mean_len * 1.5 <= 64 ? 4 : 6);
  auto wt_pick = [&](auto nh) {
    constexpr int NHc = decltype(nh)::value;
    return wt_epf == 2 ? w_step_t<VT, NHc, 2> : (wt_epf == 4 ? w_step_t<VT, NHc, 4> : w_step_t<VT, NHc, 6>);
  };
  auto wt_fn = wt_nh_used == 6 ? wt_pick(std::integral_constant<int, 6>()) : (wt_nh_used == 7 ? wt_pick(std::integral_constant<int, 7>()) : wt_pick(std::integral_constant<int, 8>()));
  const size_t wt_smem = sizeof(double) * (size_t)(1 + WT_WARPS) * hp;
  if (zpath && !fusedt) CUDA_CHECK(cudaFuncSetAttribute(w3_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * h)));
  if (fusedt) CUDA_CHECK(cudaFuncSetAttribute(wt_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wt_smem));
  const i64 w2_sms = (pdl && use_vstep && c->sm_count > 4 * VS_CLUSTER) ? c->sm_count - VS_CLUSTER : c->sm_count;
  const unsigned gb_w2 = (unsigned)std::max<i64>(1, std::min<i64>(w2_sms, (n + W2_THREADS / 32 - 1) / (W2_THREADS / 32)));
  if (use_vstep) {  // static + dynamic shared memory of the cluster kernels exceeds the 48 KB default
    CUDA_CHECK(cudaFuncSetAttribute(v_step<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (VS_WARPS + 8) * MAXB)));
    CUDA_CHECK(cudaFuncSetAttribute(v_step<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (VS_WARPS + 8) * MAXB)));
    if (vs_one) {
      CUDA_CHECK(cudaFuncSetAttribute(v_step_one<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs_one_smem));
      CUDA_CHECK(cudaFuncSetAttribute(v_step_one<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs_one_smem));
      CUDA_CHECK(cudaFuncSetAttribute(v_step_t<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs_one_smem));
      CUDA_CHECK(cudaFuncSetAttribute(v_step_t<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs_one_smem));
    }
  }
  const int* hp_halt = nullptr;  // halt flag handed to the kernels of an outer iteration (run-ahead), else none
  // The tall-skinny GEMMs Q[:, :kc] (or out) = Q[:, :inner] S[:, :kc] of a restart / of the final vectors: up to three
  // jobs in one launch (rows = 0: no CTAs), the SMs split between them in proportion to their 64-row tiles.
  auto ldo_even = [&](const RotJob& j) { return j.rows == 0 || j.out == nullptr || (j.ldo % 2 == 0 && ((uintptr_t)j.out & 15) == 0); };
  auto rotate_launch = [&](RotJob jv, RotJob jz, RotJob jw, int ldq, int inner, int kc) {
    const int kcp = (kc + 7) & ~7;
    const i64 sms = c->sm_count;
    const i64 tv = (jv.rows + RB_ROWS - 1) / RB_ROWS, tz = (jz.rows + RB_ROWS - 1) / RB_ROWS, tw = (jw.rows + RB_ROWS - 1) / RB_ROWS;
    const i64 tall = std::max<i64>(1, tv + tz + tw);
    i64 bv = tv ? std::max<i64>(1, std::min<i64>((tv * sms + tall / 2) / tall, tv)) : 0;
    i64 bz = tz ? std::max<i64>(1, std::min<i64>((tz * sms + tall / 2) / tall, tz)) : 0;
    i64 bw = tw ? std::max<i64>(1, std::min<i64>(tw, sms - bv - bz)) : 0;
    if (bv + bz + bw == 0) return;
    const unsigned nbv = (unsigned)bv, nbz = (unsigned)bz, nbw = (unsigned)bw;
    WORK(c, 8.0 * (jv.rows + jz.rows + jw.rows) * (inner + kc));
    // tensor-core path (rotate_mma): 16-row tiles when every SM gets at most a few of them, else 64-row tiles
    // streamed through a double buffer
    if (!hook(HK_ROTATE_SIMT) && ldo_even(jv) && ldo_even(jz) && ldo_even(jw)) {
      const int m4 = (inner + 3) & ~3, ldp = kcp + 4, ldqs = ((inner + 7) & ~7) + 4;
      const i64 rows_all = jv.rows + jz.rows + jw.rows;
      const bool small = rows_all <= 16 * 3 * sms;
      const int TR = small ? 16 : 64;
      const i64 uv = (jv.rows + TR - 1) / TR, uz = (jz.rows + TR - 1) / TR, uw = (jw.rows + TR - 1) / TR;
      const i64 uall = std::max<i64>(1, uv + uz + uw);
      const i64 cap = small ? uall : sms;   // small: one tile per CTA (several CTAs per SM); long: persistent CTAs
      i64 cv = uv ? std::max<i64>(1, std::min<i64>((uv * cap + uall / 2) / uall, uv)) : 0;
      i64 cz = uz ? std::max<i64>(1, std::min<i64>((uz * cap + uall / 2) / uall, uz)) : 0;
      i64 cw = uw ? std::max<i64>(1, std::min<i64>(uw, std::max<i64>(1, cap - cv - cz))) : 0;
      const size_t s_b = sizeof(double) * (size_t)m4 * ldp, q_b = sizeof(double) * (size_t)TR * ldqs;
      const int nbuf = (!small && s_b + 2 * q_b <= 225 * 1024) ? 2 : 1;
      const size_t smem = s_b + nbuf * q_b;
      REQUIRE(smem <= 227 * 1024, ADOBO_ERR_LIMIT, "rotate_mma: %d x %d rotation does not fit shared memory", inner, kcp);
      if (small) {
        CUDA_CHECK(cudaFuncSetAttribute(rotate_mma<2, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(c, "rotate_basis", (rotate_mma<2, 2, 1><<<(unsigned)(cv + cz + cw), 32 * (kcp / 8), smem, st>>>(
                                     jv, jz, jw, (int)cv, (int)cz, ldq, inner, kc, nbuf, hp_halt)));
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(rotate_mma<8, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(c, "rotate_basis", (rotate_mma<8, 4, 2><<<(unsigned)(cv + cz + cw), 32 * 2 * ((kcp / 8 + 1) / 2), smem, st>>>(
                                     jv, jz, jw, (int)cv, (int)cz, ldq, inner, kc, nbuf, hp_halt)));
      }
      return;
    }
    // fallback (rotate_simt hook; an output whose rows are not 16-byte aligned): the DFMA kernel, 64-row tiles
    const size_t s_bytes = sizeof(double) * (size_t)inner * kcp;
    {
      const size_t q_bytes = sizeof(double) * (size_t)inner * RB_QP;
      const int nbuf = (s_bytes + 2 * q_bytes <= 225 * 1024) ? 2 : 1;
      const size_t smem = s_bytes + nbuf * q_bytes;
      REQUIRE(smem <= 227 * 1024, ADOBO_ERR_LIMIT, "rotate_basis: %d x %d rotation does not fit shared memory", inner, kcp);
      CUDA_CHECK(cudaFuncSetAttribute(rotate_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      LAUNCH(c, "rotate_basis", (rotate_basis<<<nbv + nbz + nbw, 32 * (kcp / 8), smem, st>>>(jv, jz, jw, (int)nbv, (int)nbz, ldq, inner,
                                                                                       kc, nbuf, hp_halt)));
    }
  };
  const RotJob no_job{nullptr, 0, nullptr, nullptr, 0, nullptr};
  // V[:, :kc] (or outv) = V Sv[:, :kc]  and  W[:, :kc] (or outw) = W Su[:, :kc]  (and Z = A^T W, which follows W at a
  // restart); with_w = false: V and Z only (lazy rotation of W)
  auto rotate2 = [&](int kc, double* outv, double* outw, int ldo, const double* scale_w, bool st_ready, bool with_w) {
    const int kcp = (kc + 7) & ~7;
    if (!st_ready)  // restarts: conv_check has laid the two matrices out already
      LAUNCH(c, "transpose_s", (transpose_s<<<(2 * m_b * kcp + 255) / 256, 256, 0, st>>>(Sv.p, Su.p, m_b, kc, kcp, St.p)));
    const bool with_z = zpath && outv == nullptr;  // restart rotation: Z = A^T W follows W
    RotJob jv{V.p, (i64)h, St.p, outv, ldo, nullptr};
    RotJob jz = with_z ? RotJob{Z.p, (i64)h, St.p + (size_t)m_b * kcp, nullptr, 0, nullptr} : no_job;
    RotJob jw = with_w ? RotJob{W.p, n, St.p + (size_t)m_b * kcp, outw, ldo, scale_w} : no_job;
    rotate_launch(jv, jz, jw, ld, m_b, kc);
  };
  // ---- lazy rotation of W: state of the stored columns (see irlb_step.cuh) ----
  struct LState {
    int U0 = 0, kp = 0, S = 0;  // S: stored columns once the iteration's steps are done
  };
  auto lazy_next = [&](const LState& a, int k, bool* materialise) {  // state of the iteration that restarts with k
    LState b;
    const bool mat = a.S + (m_b - k) > Ms;
    if (materialise) *materialise = mat;
    if (mat) b.U0 = 0, b.kp = 0, b.S = m_b;
    else b.U0 = a.S, b.kp = k, b.S = a.S + (m_b - k);
    return b;
  };
  auto lazy_args = [&](const LState& a) {
    LazyArgs lz;
    if (lazy) lz.P0 = ws.P0.p, lz.P0t = ws.P0t.p, lz.ldp = ldp, lz.U0 = a.U0, lz.kp = a.kp;
    return lz;
  };
  // W_s[:, :kc] (or out) = W_s[:, :S] Pmat, the real rotation (Pmat: S x kcp from lazy_p0)
  auto rotate_w_stored = [&](int S, int kc, double* out, int ldo, const double* scale_w) {
    RotJob jw{W.p, n, ws.Pmat.p, out, ldo, scale_w};
    rotate_launch(no_job, no_job, jw, ldw, S, kc);
  };
  const bool use_arrow = !hook(HK_SVD_JACOBI);
  const bool force_fallback = hook(HK_SVD_FALLBACK);
  if (use_arrow) svd_arrow_prepare(m_b);
  const unsigned fin_h = (unsigned)((h + 255) / 256), fin_n = (unsigned)std::max<i64>(1, (n + 255) / 256);
  // page-locked, device-visible: conv_check publishes the iteration's info words here (publish_info)
  unsigned long long* h_info = (unsigned long long*)c->pinned;
  for (int i = 0; i < 8; ++i) h_info[i] = 0;   // (no device work of this context is in flight here)
  int hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto read_info = [&]() {  // consistent snapshot of the mailbox: five {value, version} words with one version
    static const int field[5] = {0, 1, 4, 5, 6};
    volatile unsigned long long* hv = h_info;
    for (;;) {
      unsigned long long w[5];
      for (int i = 0; i < 5; ++i) w[i] = hv[i];
      bool same = true;
      for (int i = 1; i < 5; ++i) same = same && (w[i] >> 32) == (w[0] >> 32);
      if (!same) continue;
      for (int i = 0; i < 5; ++i) hi[field[i]] = (int)(unsigned)(w[i] & 0xffffffffull);
      return;
    }
  };
  // run-ahead: the host polls the mailbox instead of waiting on an event recorded between the iteration graphs
  auto wait_iteration = [&](int iter) {
    unsigned spins = 0;
    for (;;) {
      read_info();
      if (hi[5] > iter || hi[4] != 0) return;
      if ((++spins & 255u) == 0) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) {  // everything enqueued has run: the words are final
          read_info();
          REQUIRE(hi[5] > iter || hi[4] != 0, ADOBO_ERR_STATE, "lanczos: iteration %d finished without publishing its result", iter);
          return;
        }
        if (q != cudaErrorNotReady) CUDA_CHECK(q);
      }
    }
  };

  // ---- start vector (irlb.py:151-153): V[:,0] = v0 / ||v0|| ----
  CUDA_CHECK(cudaMemcpyAsync(ybuf_h.p, v0_host, sizeof(double) * h, cudaMemcpyHostToDevice, st));
  PeerComm pcw = fused ? c->peer : PeerComm();
  pcw.err = info.p + 6;
  const PeerComm pc1;  // replicated (V-side) reductions
  // the finish step (normalise + store) runs in the last CTA of the norm reduction when one CTA can do it
  // in about a launch latency; larger vectors keep the separate grid-wide kernels
  const bool fold_v = h <= 16384, fold_w = !wide_rows && n <= 8192 && !(c->world > 1 && !fused);
  // otherwise the W finish rides in the spmvT_csc that consumes the column -- except for the last column, whose
  // norm completes B for the SVD that is forked off before that spmvT_csc
  const bool fold_wt = (zpath || !fold_w) && fold_wt_ok;
  auto v_fin = [&](int jn, int j) {
    Finish f;
    if (!fold_v) return f;
    f.kind = 1, f.rs = rs, f.sc = sc.p, f.M = V.p, f.ld = ld, f.jn = jn, f.m_b = m_b, f.j = j, f.centered = centered ? 1 : 0;
    f.fbuf = fbuf.p, f.xs = xs.p, f.B = B.p;
    return f;
  };
  auto w_fin = [&](int jn) {
    Finish f;
    if (!fold_w) return f;
    f.kind = 2, f.sc = sc.p, f.M = W.p, f.ld = ld, f.jn = jn, f.fbuf = wcur.p, f.B = B.p, f.m_b = m_b;
    return f;
  };
  LAUNCH(c, "update_norm", (update_norm_w32<<<gb_h, RT_THREADS, 0, st>>>(V.p, ld, 0, cvec.p, ybuf_h.p,
                                                            centered ? murs.p : nullptr, h, partials.p,
                                                            sc.p + SC_NRM2, counter.p, pc1, v_fin(0, -1))));
  if (!fold_v)
    LAUNCH(c, "v_finish", (v_finish<<<fin_h, 256, 0, st>>>(ybuf_h.p, rs, sc.p, V.p, ld, 0, m_b, h, fbuf.p, xs.p, B.p, -1,
                                                         centered ? 1 : 0)));

  // algorithmic bytes of the step kernels: one pass over the subset CSR and over W[:, :nd]
  auto w_bytes = [&](int nd, int jprev) {
    return (double)A.nnz * (4.0 + sizeof(VT)) + 8.0 * (n + 1) + 8.0 * n * (nd + (jprev >= 0 ? 1 : 0) + 1) + 8.0 * h;
  };
  // ---- W side of a step on the three-kernel / two-pass paths ----
  auto w_column = [&](int jprev, int nd, int jn) {
    // W[:, jn] = normalise( orthog( A_s x - fn W[:, jprev], W[:, :nd] ) )
    WORK(c, w_bytes(nd, jprev));
    if (zpath) {  // one pass: the coefficients were left in cvec by v_step (set_restart after a restart)
      // programmatic launch behind v_step; the first column after a restart follows ordinary kernels
      LAUNCH(c, "w_step", CUDA_CHECK(launch_k(w3_fn, dim3(gb_w2), dim3(W2_THREADS), sizeof(double) * h, st, pdl && jprev >= 0, 1,
                                              A.rptr.p, A.cidx.p, A.rval.p, xs.p, sc.p, W.p, ld, jprev, nd, (int)n, cvec.p,
                                              ybuf_n.p, partials.p, gout.p, counter.p, pcw, hp_halt, h, split ? 1 : 0)));
      if (c->world > 1 && !fused) allreduce_sum_f64(c, gout.p, (size_t)nd + 2);
      if (!(fold_wt && jn != m_b - 1))
        LAUNCH(c, "w_finish", (w_finish<<<fin_n, 256, 0, st>>>(ybuf_n.p, gout.p + nd, sc.p, W.p, ld, jn, n, wcur.p, B.p, m_b, hp_halt,
                                                               split ? partials.p : nullptr, (int)gb_w2, gout.p)));
      return;
    }
    if (!wide_rows)
      LAUNCH(c, "spmv_dots", (spmv_dots_w32<VT><<<gb_n, RT_THREADS, 0, st>>>(A.rptr.p, A.cidx.p, A.rval.p, xs.p, sc.p, W.p, ld,
                                                                    jprev, nd, n, ybuf_n.p, partials.p, cvec.p,
                                                                    counter.p, pcw)));
    else if (m_b <= 96)
      LAUNCH(c, "spmv_dots", (spmv_dots<VT, 12><<<gb_sd, SD_THREADS, 0, st>>>(A.rptr.p, A.cidx.p, A.rval.p, xs.p, sc.p, W.p, ld,
                                                                     jprev, nd, n, ybuf_n.p, partials.p, cvec.p,
                                                                     counter.p, pcw)));
    else
      LAUNCH(c, "spmv_dots", (spmv_dots<VT, 16><<<gb_sd, SD_THREADS, 0, st>>>(A.rptr.p, A.cidx.p, A.rval.p, xs.p, sc.p, W.p, ld,
                                                                     jprev, nd, n, ybuf_n.p, partials.p, cvec.p,
                                                                     counter.p, pcw)));
    if (nd > 0 && c->world > 1 && !fused) allreduce_sum_f64(c, cvec.p, (size_t)nd);
    WORK(c, 8.0 * n * (nd + 2));
    if (wide_rows)
      LAUNCH(c, "update_norm", (update_norm<<<gb_n, RT_THREADS, 0, st>>>(W.p, ld, nd, cvec.p, ybuf_n.p, nullptr, n, partials.p,
                                                                sc.p + SC_NRM2, counter.p, pcw)));
    else
      LAUNCH(c, "update_norm", (update_norm_w32<<<gb_n, RT_THREADS, 0, st>>>(W.p, ld, nd, cvec.p, ybuf_n.p, nullptr, n,
                                                                    partials.p, sc.p + SC_NRM2, counter.p, pcw,
                                                                    w_fin(jn))));
    if (c->world > 1 && !fused) allreduce_sum_f64(c, sc.p + SC_NRM2, 2);
    if ((wide_rows || !fold_w) && !(fold_wt && jn != m_b - 1))
      LAUNCH(c, "w_finish", (w_finish<<<fin_n, 256, 0, st>>>(ybuf_n.p, sc.p + SC_NRM2, sc.p, W.p, ld, jn, n, wcur.p, B.p, m_b, nullptr,
                                                             nullptr, 0, nullptr)));
  };
  // restart update (irlb.py:238-246): Ritz vectors into the leading k columns, V[:,k] = F, new B.
  // `prev`: lazy-rotation state of the iteration that just ended (ignored without the lazy rotation)
  bool restart_mat = true;  // whether the kernel before the iteration's first w_step_t is NOT lazy_p0 (no programmatic edge then)
  auto restart_update = [&](int k, const LState& prev) {
    const unsigned sr_grid = (unsigned)((std::max(h, m_b * m_b) + 255) / 256);
    if (!lazy) {
      rotate2(k, nullptr, nullptr, 0, nullptr, true, true);
      LAUNCH(c, "set_restart", (set_restart<<<sr_grid, 256, 0, st>>>(V.p, ld, h, fbuf.p, B.p, m_b, k, svd_s.p, R.p,
                                                                  zpath ? cvec.p : nullptr, sumW.p, ws.crot.p, hp_halt,
                                                                  LazyArgs(), 0)));
      return;
    }
    bool mat = false;
    const LState next = lazy_next(prev, k, &mat);
    const int kcp = (k + 7) & ~7;
    rotate2(k, nullptr, nullptr, 0, nullptr, true, false);  // V and Z
    // lazy_p0 + set_restart in one kernel, beside the rotation (programmatic edge; see lazy_p0)
    RestartArgs ra;
    ra.V = V.p, ra.ldv = ld, ra.h = h, ra.m = m_b, ra.k = k, ra.fbuf = fbuf.p, ra.B = B.p, ra.sv = svd_s.p, ra.R = R.p;
    ra.sumW = sumW.p, ra.cvec = cvec.p, ra.crot = ws.crot.p;
    ra.U0n = next.U0, ra.kpn = next.kp, ra.nds = next.U0 + (k - next.kp);
    LAUNCH(c, "lazy_p0", CUDA_CHECK(launch_k(lazy_p0, dim3(prev.S), dim3(128), 0, st, pdl, 1, ws.P0.p, ws.P0t.p, ldp, prev.U0,
                                             prev.kp, (const double*)Su.p, m_b, k, mat ? ws.Pmat.p : (double*)nullptr, kcp,
                                             hp_halt, mat ? (const double*)nullptr : (const double*)ws.crot.p, cvec.p, ra)));
    if (mat) rotate_w_stored(prev.S, k, nullptr, 0, nullptr);  // W_s[:, :k] = the Ritz vectors themselves again
    restart_mat = mat;
  };
  const bool use_graphs = !c->profiling && !hook(HK_NO_GRAPH);
  // SVD of B (irlb.py:224): restarts have the diag + spike + bidiagonal-tail structure (svd_arrow), the first
  // iteration is a plain bidiagonal matrix (Jacobi).  B is complete once the last column of W is normalised,
  // so under graph capture the SVD is forked onto the side stream and overlaps the last V step.
  auto enqueue_svd = [&](bool first, int k, cudaStream_t s, const LastDiag& ldg = LastDiag()) {
    if (!first && use_arrow && svd_arrow_supported(m_b, k))
      launch_svd_arrow(c, B.p, m_b, k, Su.p, svd_s.p, Sv.p, info.p, s, ldg);
    else
      LAUNCH(c, "jacobi_svd", (jacobi<<<1, JTHREADS, svd_smem, s>>>(B.p, m_b, Su.p, svd_s.p, Sv.p, info.p)));
  };
  const bool fork_svd = use_graphs && !hook(HK_NO_FORK);
  auto fork_svd_here = [&](bool first, int k, const LastDiag& ldg = LastDiag()) {
    CUDA_CHECK(cudaEventRecord(c->fork_ev, st));
    CUDA_CHECK(cudaStreamWaitEvent(c->side, c->fork_ev, 0));
    enqueue_svd(first, k, c->side, ldg);
    CUDA_CHECK(cudaEventRecord(c->join_ev, c->side));
  };
  constexpr int RUN_DEPTH = 3, CC_BLOCKS = 16;
  const bool trace = hook(HK_TRACE);
  const bool runahead = use_graphs && zpath && (c->world == 1 || (fused && fused_t)) && !trace && !force_fallback &&
                        !hook(HK_NO_RUNAHEAD);
  auto zargs = [&](int j) {
    ZArgs za;
    if (zpath)
      za.Z = Z.p, za.sumW = sumW.p, za.gprev = gout.p, za.cvec = cvec.p, za.ncc = (j < m_b - 1) ? j + 1 : m_b,
      za.prev = (j < m_b - 1) ? 1 : 0;
    return za;
  };
  // ---- one outer iteration (irlb.py:155-231) up to the convergence read-back: two-kernel steps ----
  auto enqueue_steps_fused = [&](bool first, int k, const LState& ls) {
    const LazyArgs lz = lazy_args(ls);
    auto slot = [&](int col) { return ls.U0 + (col - ls.kp); };  // where column col of W is stored
    // the first column of an iteration follows lazy_p0 (programmatic edge: its prologue reads the static CSR and rows
    // of W_s only) unless W_s has just been rotated for real
    const bool first_pdl = !first && lazy && !restart_mat;
    auto w_launch = [&](int jprev, int nd) {                     // effective indices in, stored indices to the kernel
      const int nds = slot(nd), jps = jprev >= 0 ? nds - 1 : -1;
      WORK(c, w_bytes(nds, jps));
      LAUNCH(c, "w_step_t", CUDA_CHECK(launch_k(wt_fn, dim3(gb_w2), dim3(W2_THREADS), wt_smem, st, pdl && (jprev >= 0 || first_pdl), 1, A.rptr.p,
                                                A.cidx.p, A.rval.p, xs.p, sc.p, W.p, ldw, jps, nds, (int)n, cvec.p, ybuf_n.p,
                                                partials.p, tpart.p, hp_halt, h, hp)));
    };
    int j = first ? 0 : k;
    w_launch(-1, j);
    while (j < m_b) {
      const bool last = j == m_b - 1;
      const int sj = slot(j);
      if (last) {
        // W[:, m_b-1] and B[m_b-1, m_b-1] now: B is complete and its SVD runs beside the last V step
        if (c->world > 1)
          LAUNCH(c, "w_norm_exchange", (w_norm_exchange<<<1, 256, 0, st>>>(partials.p, (int)gb_w2, sj, gout.p, hp_halt, pcw)));
        // the structured SVD takes B[m_b-1, m_b-1] = ||y|| from the partials / the exchanged pair itself (LastDiag) and
        // is forked BEFORE w_finish_t: it is the longer branch beside the last V step
        const bool early = fork_svd && !first && use_arrow && svd_arrow_supported(m_b, k);
        if (early) {
          LastDiag ldg;
          if (c->world > 1) ldg.n2p = gout.p;
          else ldg.parts = partials.p, ldg.nblk = (int)gb_w2, ldg.col0 = sj, ldg.stride = MAXB;
          fork_svd_here(first, k, ldg);
        }
        LAUNCH(c, "w_finish", (w_finish_t<<<fin_n, 256, 0, st>>>(ybuf_n.p, gout.p, sc.p, W.p, ldw, sj, j, n, B.p, m_b, hp_halt,
                                                                 c->world > 1 ? nullptr : partials.p, (int)gb_w2)));
        if (fork_svd && !early) fork_svd_here(first, k);
      }
      WORK(c, 8.0 * gb_w2 * (hp + sj + 2.0) + 8.0 * h * (2.0 * (j + 1) + 8));
      LAUNCH(c, "v_step_t", CUDA_CHECK(launch_k(m_b <= 96 ? v_step_t<3> : v_step_t<4>, dim3(VS_CLUSTER), dim3(VS_THREADS),
                                                vs_one_smem, st, pdl, VS_CLUSTER, (const double*)tpart.p, hp,
                                                (const double*)partials.p, (int)gb_w2, mu, rs,
                                                (const double*)(centered ? murs.p : nullptr), sc.p, V.p, ld, j, m_b, h,
                                                ybuf_h.p, fbuf.p, xs.p, B.p, zargs(j), hp_halt, pcw, lz)));
      if (!last) w_launch(j, j + 1);
      ++j;
    }
  };
  // ---- the same on the three-kernel / two-pass paths ----
  auto enqueue_steps = [&](bool first, int k) {
    int j = first ? 0 : k;
    w_column(-1, j, j);
    while (j < m_b) {
      if (j == m_b - 1 && fork_svd) fork_svd_here(first, k);
      WORK(c, (double)A.nnz * (4.0 + sizeof(VT)) + 8.0 * (h + 1) + 8.0 * h + 8.0 * n);
      {
        // column j of W arrives normalised (w_finish / folded finish) or as ybuf + ||w||^2 (finish folded in here)
        WFinish wf;
        const bool infold = fold_wt && j != m_b - 1;
        if (infold)
          wf.W = W.p, wf.ldw = ld, wf.jn = j, wf.m_b = m_b, wf.n = n, wf.sc = sc.p, wf.B = B.p,
          wf.n2p = zpath ? gout.p + j : sc.p + SC_NRM2,  // column j came from w_step with nd = j
          wf.parts = split ? partials.p : nullptr, wf.nblk = (int)gb_w2, wf.gout = split ? gout.p : nullptr;
        const bool spdl = pdl && infold && zpath && (fused || c->world == 1);
        LAUNCH(c, "spmvT_csc", CUDA_CHECK(launch_k(spmvT_csc<VT>, dim3(h), dim3(256), 0, st, spdl, 1, A.cptr.p, A.ridx.p, A.cval.p,
                                                   (const double*)(infold ? ybuf_n.p : wcur.p), tbuf.p, h, counter.p + 1,
                                                   fused_t ? pcw : pc1, wf, hp_halt)));
      }
      if (c->world > 1 && !fused_t) allreduce_sum_f64(c, tbuf.p, (size_t)h);
      if (use_vstep) {
        WORK(c, 8.0 * h * (2.0 * (j + 1) + 8));
        const bool vpdl = pdl && (fused_t || c->world == 1);
        if (vs_one)
          LAUNCH(c, "v_step", CUDA_CHECK(launch_k(m_b <= 96 ? v_step_one<3> : v_step_one<4>, dim3(VS_CLUSTER), dim3(VS_THREADS),
                                                  vs_one_smem, st, vpdl, VS_CLUSTER, (const double*)tbuf.p, mu, rs,
                                                  (const double*)(centered ? murs.p : nullptr), sc.p, V.p, ld, j, m_b, h,
                                                  ybuf_h.p, fbuf.p, xs.p, B.p, zargs(j), hp_halt)));
        else
          LAUNCH(c, "v_step", CUDA_CHECK(launch_k(m_b <= 96 ? v_step<3> : v_step<4>, dim3(VS_CLUSTER), dim3(VS_THREADS),
                                                  sizeof(double) * (VS_WARPS + 8) * MAXB, st, vpdl, VS_CLUSTER,
                                                  (const double*)tbuf.p, mu, rs, (const double*)(centered ? murs.p : nullptr),
                                                  sc.p, V.p, ld, j, m_b, h, ybuf_h.p, fbuf.p, xs.p, B.p, zargs(j), hp_halt)));
      } else {
        WORK(c, 8.0 * h * (j + 1 + 4));
        LAUNCH(c, "v_prep_dots", (v_prep_dots<<<gb_h, RT_THREADS, 0, st>>>(tbuf.p, mu, rs, sc.p, V.p, ld, j, j + 1, h, ybuf_h.p,
                                                                  partials.p, cvec.p, counter.p)));
        WORK(c, 8.0 * h * (j + 1 + 3));
        LAUNCH(c, "update_norm_v", (update_norm_w32<<<gb_h, RT_THREADS, 0, st>>>(V.p, ld, j + 1, cvec.p, ybuf_h.p,
                                                                  centered ? murs.p : nullptr, h, partials.p,
                                                                  sc.p + SC_NRM2, counter.p, pc1, v_fin(j + 1, j))));
        if (!fold_v)
          LAUNCH(c, "v_finish", (v_finish<<<fin_h, 256, 0, st>>>(ybuf_h.p, rs, sc.p, V.p, ld, j + 1, m_b, h, fbuf.p, xs.p,
                                                               B.p, j, centered ? 1 : 0)));
      }
      if (j < m_b - 1) w_column(j, j + 1, j + 1);
      ++j;
    }
  };
  // `prev`: lazy-rotation state of the iteration before (the fresh state for the first iteration)
  auto enqueue_iteration = [&](bool first, int k, const LState& prev) {
    hp_halt = runahead ? info.p + 4 : nullptr;
    if (!first) restart_update(k, prev);
    if (fusedt)
      enqueue_steps_fused(first, k, (first || !lazy) ? prev : lazy_next(prev, k, nullptr));
    else
      enqueue_steps(first, k);
    const bool arrow = !first && use_arrow && svd_arrow_supported(m_b, k);
    if (fork_svd)
      CUDA_CHECK(cudaStreamWaitEvent(st, c->join_ev, 0));
    else
      enqueue_svd(first, k, st);
    LAUNCH(c, "conv_check", (conv_check<<<CC_BLOCKS, 256, 0, st>>>(Su.p, svd_s.p, sc.p, m_b, nu, k, first ? 0 : 1, tol, R.p,
                                                                 info.p, arrow ? 1 : 0, 0, Sv.p, St.p,
                                                                 zpath ? cvec.p : nullptr, sumW.p, ws.crot.p, h_info)));
    hp_halt = nullptr;
  };
  // fallback when svd_arrow rejected its own result: generic Jacobi SVD of the same B
  auto svd_fallback = [&](bool first, int k) {
    CUDA_CHECK(cudaMemsetAsync(info.p + 4, 0, sizeof(int), st));  // clear the halt flag: jacobi_svd honours it
    LAUNCH(c, "jacobi_svd", (jacobi<<<1, JTHREADS, svd_smem, st>>>(B.p, m_b, Su.p, svd_s.p, Sv.p, info.p)));
    LAUNCH(c, "conv_check", (conv_check<<<CC_BLOCKS, 256, 0, st>>>(Su.p, svd_s.p, sc.p, m_b, nu, k, first ? 0 : 1, tol, R.p,
                                                                 info.p, 0, 1, Sv.p, St.p, zpath ? cvec.p : nullptr,
                                                                 sumW.p, ws.crot.p, h_info)));
    CUDA_CHECK(cudaStreamSynchronize(st));
    read_info();
  };
  auto run_iteration = [&](bool first, int k, const LState& prev) {
    if (!use_graphs) {
      enqueue_iteration(first, k, prev);
      return;
    }
    const i64 key = (first ? 1 : 0) | ((i64)k << 1) | ((i64)prev.kp << 9) | ((i64)prev.U0 << 17);
    auto it = ws.graphs.find(key);
    if (it == ws.graphs.end()) {
      const i64 l0 = c->launches;
      cudaGraph_t graph = nullptr;
      CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      try {
        enqueue_iteration(first, k, prev);
      } catch (...) {
        cudaStreamEndCapture(st, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      CUDA_CHECK(cudaStreamEndCapture(st, &graph));
      cudaGraphExec_t exec = nullptr;
      cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(e));
        throw Fail{ADOBO_ERR_CUDA};
      }
      ws.graph_launches[key] = c->launches - l0;
      c->launches = l0;
      it = ws.graphs.emplace(key, exec).first;
    }
    CUDA_CHECK(cudaGraphLaunch(it->second, st));
    c->launches += ws.graph_launches[key];
  };

  IrlbResult res;
  // Outer loop.  k only grows and stays at m_b - 3 once it gets there (conv_check), which is where almost all
  // restarts happen: from then on the host does not wait for the 32-byte read-back before enqueueing the next
  // iteration.  It runs up to RUN_DEPTH iterations ahead; when an iteration converges (or its SVD is rejected)
  // conv_check raises the halt flag on the device and every kernel of the iterations already enqueued behind it
  // returns at once, leaving V, W, B and the SVD of the converged iteration untouched.
  std::vector<int> ks;       // k of every enqueued iteration
  std::vector<LState> lst;   // and the lazy-rotation state its steps run in (fresh without the lazy rotation)
  LState fresh;
  fresh.S = m_b;
  int it = 0, k = nu, enq = 0;
  bool converged = false;
  while (it < maxit) {
    while (enq < maxit && (enq == it || (runahead && enq - it < RUN_DEPTH && it > 0 && k == m_b - 3))) {
      const LState prev = enq == 0 ? fresh : lst[enq - 1];
      run_iteration(enq == 0, k, prev);
      ks.push_back(k);
      lst.push_back((enq == 0 || !lazy) ? fresh : lazy_next(prev, k, nullptr));
      ++enq;
    }
    if (runahead) {
      wait_iteration(it);
    } else {
      CUDA_CHECK(cudaStreamSynchronize(st));
      read_info();
    }
    REQUIRE(hi[6] == 0, ADOBO_ERR_NCCL, "lanczos: rank %d did not answer an exchange over NVLink peer memory within %d s",
            hi[6] - 1, (int)(PEER_TIMEOUT_NS / 1000000000ull));
    if (runahead && hi[4] != 0) {
      // something stopped: drain the (no-op) iterations behind it and read the final state
      CUDA_CHECK(cudaStreamSynchronize(st));
      read_info();
      it = (hi[4] == 1) ? hi[5] - 1 : hi[5];
      k = ks[it];
      enq = it + 1;
      ks.resize(enq);
      lst.resize(enq);
    }
    if (hi[0] < 0 || (force_fallback && it > 0)) {  // structured SVD failed its self-check (kept for safety)
      svd_fallback(it == 0, k);
      ++res.svd_fallbacks;
    }
    if (trace) fprintf(stderr, "irlb it=%d k=%d conv=%d next_k=%d\n", it, k, hi[0], hi[1]);
    if (hi[0] >= nu) {  // irlb.py:232-236
      converged = true;
      break;
    }
    k = hi[1];
    ++it;
  }
  int nmult = 0;  // 1 + sum over j of (1 + [j < m_b-1]) mat-vecs per outer iteration
  for (int i = 0; i < (int)ks.size() && i <= it; ++i) nmult += 2 * (m_b - (i == 0 ? 0 : ks[i]));
  // the reference applies the restart update before `it` reaches maxit (irlb.py:238-247)
  const LState fin = lst.empty() ? fresh : lst[std::min<size_t>(it, lst.size() - 1)];  // state the vectors were left in
  const bool exhausted = !converged && it > 0;  // the reference applies the restart update before `it` reaches maxit
  // ---- U = W Su[:, :nu] (x diag(s) when var_weigh), V = V Sv[:, :nu]  (irlb.py:249-250, dr.py:164-168) ----
  outU.ensure((size_t)std::max<i64>(n, 1) * nu);
  outV.ensure((size_t)h * nu);
  outS.ensure((size_t)nu);
  if (!lazy) {
    if (exhausted) restart_update(k, fin);  // St for this k was laid out by the last conv_check
    rotate2(nu, outV.p, outU.p, nu, var_weigh ? svd_s.p : nullptr, false, true);
  } else if (!exhausted) {
    // U = W_s [P0 0; 0 I] Su[:, :nu]: the small product first, then one real rotation of the stored columns
    rotate2(nu, outV.p, nullptr, nu, nullptr, false, false);
    LAUNCH(c, "lazy_p0", (lazy_p0<<<fin.S, 128, 0, st>>>(ws.P0.p, ws.P0t.p, ldp, fin.U0, fin.kp, Su.p, m_b, nu, ws.Pmat.p,
                                                        (nu + 7) & ~7, nullptr, nullptr, nullptr, RestartArgs())));
    rotate_w_stored(fin.S, nu, outU.p, nu, var_weigh ? svd_s.p : nullptr);
  } else {
    // Not converged within maxit.  The reference then rotates the first k columns of V and W once more and multiplies
    // the MIXED basis (new Ritz vectors | old columns k..m_b-1) by Su[:, :nu] (irlb.py:238-250).  V is rotated in
    // place as always; for W the two products are folded into one S x nu matrix on the host (off the hot path).
    rotate2(k, nullptr, nullptr, 0, nullptr, true, false);
    LAUNCH(c, "set_restart", (set_restart<<<(unsigned)((std::max(h, m_b * m_b) + 255) / 256), 256, 0, st>>>(
                                 V.p, ld, h, fbuf.p, B.p, m_b, k, svd_s.p, R.p, nullptr, sumW.p, ws.crot.p, nullptr, LazyArgs(), 0)));
    rotate2(nu, outV.p, nullptr, nu, nullptr, false, false);
    std::vector<double> hsu((size_t)m_b * m_b), hp0((size_t)MAXB * MAXB);
    CUDA_CHECK(cudaMemcpyAsync(hsu.data(), Su.p, sizeof(double) * m_b * m_b, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(hp0.data(), ws.P0.p, sizeof(double) * (size_t)MAXB * MAXB, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const int S = fin.S, kcp = (nu + 7) & ~7;
    auto su = [&](int e, int col) { return hsu[(size_t)col * m_b + e]; };          // component e of left vector col
    auto pfull = [&](int sidx, int e) {                                            // [P0 0; 0 I]
      if (e < fin.kp) return sidx < fin.U0 ? hp0[(size_t)sidx * ldp + e] : 0.0;
      return sidx == fin.U0 + (e - fin.kp) ? 1.0 : 0.0;
    };
    std::vector<double> ritz((size_t)S * k), comb((size_t)S * kcp, 0.0);
    for (int sidx = 0; sidx < S; ++sidx)
      for (int col = 0; col < k; ++col) {
        double a = 0;
        for (int e = 0; e < m_b; ++e) a += pfull(sidx, e) * su(e, col);
        ritz[(size_t)sidx * k + col] = a;                                           // W_new[:, col] = W_s ritz[:, col]
      }
    for (int sidx = 0; sidx < S; ++sidx)
      for (int col = 0; col < nu; ++col) {
        double a = 0;
        for (int e = 0; e < k; ++e) a += ritz[(size_t)sidx * k + e] * su(e, col);
        for (int e = k; e < m_b; ++e) a += pfull(sidx, e) * su(e, col);
        comb[(size_t)sidx * kcp + col] = a;
      }
    CUDA_CHECK(cudaMemcpyAsync(ws.Pmat.p, comb.data(), sizeof(double) * (size_t)S * kcp, cudaMemcpyHostToDevice, st));
    rotate_w_stored(S, nu, outU.p, nu, var_weigh ? svd_s.p : nullptr);
    CUDA_CHECK(cudaStreamSynchronize(st));  // comb goes out of scope
  }
  CUDA_CHECK(cudaMemcpyAsync(outS.p, svd_s.p, sizeof(double) * nu, cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
#ifdef ADOBO_B200_TIMELINE
  if (getenv("ADOBO_B200_TIMELINE")) dump_timeline();
#endif
  res.steps = it;
  res.nmult = nmult;
  if (getenv("ADOBO_B200_DEBUG"))
    fprintf(stderr, "[adobo_b200] irlb: %d restarts, %d mat-vecs, %d structured-SVD fallbacks, %s step\n", it, nmult,
            res.svd_fallbacks, fusedt ? "two-kernel" : (zpath ? "three-kernel" : "two-pass"));
  return res;
}

template IrlbResult run_irlb<float>(adobo_ctx*, SparsePair<float>&, const double*, const double*, int, double, int,
                                    const double*, bool, DevBuf<double>&, DevBuf<double>&, DevBuf<double>&);
template IrlbResult run_irlb<double>(adobo_ctx*, SparsePair<double>&, const double*, const double*, int, double, int,
                                     const double*, bool, DevBuf<double>&, DevBuf<double>&, DevBuf<double>&);

}  // namespace adobo
