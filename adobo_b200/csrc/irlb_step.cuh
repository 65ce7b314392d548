// The two kernels of a Lanczos step (h <= 1024 selected genes, the configuration every BASELINE config uses).
// Included by irlb.cu after the shared reduction helpers, W3Row / w3_load and ZArgs.
//
//   w_step_t   W side (irlb.py:204-217) AND the products of A^T with the new column (irlb.py:182), all ranks' SMs
//   v_step_t   V side (irlb.py:185-197, 203-221) on a cluster of 8 CTAs, exchange between GPUs included
//
// The column-wise kernel spmvT_csc that used to sit between them is gone from this path: it ended with its longest
// gene column (HVG columns hold from 20 to n non-zeros: max / mean 22 on the benchmark matrix) and cost a third
// launch, a third dependency and a second exchange between GPUs per step.
#pragma once

// ---- w_step_t: one pass over the rows -----------------------------------------------------------------------------
// y_r = (A xs)_r - mvs - fn W[r, jprev] - sum_c W[r, c] c_c  as in w_step3 (half-warp per row, two-deep pipeline),
// plus, while the row's entries are still in registers,
//     t += y_r * A[r, :]                   (A^T is linear: v_step_t applies 1 / ||y|| to the finished sum)
// into a PRIVATE copy of t per warp in shared memory (h doubles each): plain load / fma / store, no atomics.  A warp
// walks its contiguous rows in order and its two half-warps take turns (row r, then row r + 1), so every copy is
// summed in one fixed order; the 16 copies of a CTA are added in warp order and the CTA's partial goes to
// tpart[cta][hp] -- v_step_t adds the CTA partials in a fixed order too.  Bit-reproducible for a given shard.
//
// The column W[:, jprev] = y_prev / ||y_prev|| is written HERE, by the warp that owns the row (||y_prev|| was not
// known when y_prev was formed: it needs the sum over all rows and ranks, which v_step_t has made since); the last
// column of an outer iteration is written by w_finish.
constexpr int WT_WARPS = W2_THREADS / 32;
constexpr int WT_HMAX = 1024;  // (1 + WT_WARPS) copies of h doubles in shared memory

// EPF: entry slots per lane held in registers (16 EPF entries of a row; longer rows finish in a loop).  Every slot
// costs its loads, its fma and its scatter whether the row fills it or not, so the host picks EPF from the mean row
// length of the subset (15 entries at 68 k x 1000 HVGs: EPF = 2; six slots were 355 instructions per pair of rows).
template <typename VT, int NH, int EPF>
struct WTRow {
  int ci[EPF];
  VT va[EPF];
  double wv[NH], yp;
  int b, len;
};
// (b, e1): the row's pointers -- loaded one pair EARLIER than the entries they delimit, so that the chain
// row pointers -> entries -> gather never stalls a prefetch on two dependent round trips (ncu: the first use of the
// prefetched entries and of the row length were the two hottest stall sites of the kernel)
// The basis slice of a row is loaded WITHOUT per-lane predicates: every column of W_s holds a finite number (the
// buffer is zeroed when the decomposition starts and only ever written with finite values), the coefficient of a
// column that is not in use is zero, and what such a column adds to g_s is never read -- so only whole groups of 16
// columns beyond nd are skipped (a warp-uniform test).  A row past the end reads row n - 1 (its y is forced to 0).
template <typename VT, int NH, int EPF>
__device__ __forceinline__ void wt_load(WTRow<VT, NH, EPF>& R, bool ok, int b, int e1, int rowc,
                                        const int* __restrict__ cl, const VT* __restrict__ vl,
                                        const double* __restrict__ wr, const double* __restrict__ ybuf, bool lazy,
                                        int nd, int hl) {
  R.b = b;
  R.len = ok ? e1 - b : 0;
#pragma unroll
  for (int q = 0; q < EPF; ++q) {
    const bool in = hl + 16 * q < R.len;
    R.ci[q] = in ? __ldg(cl + b + 16 * q) : 0;
    R.va[q] = in ? __ldg(vl + b + 16 * q) : VT(0);
  }
#pragma unroll
  for (int q = 0; q < NH; ++q) R.wv[q] = (16 * q < nd) ? __ldcg(wr + 16 * q) : 0.0;
  R.yp = lazy ? __ldcg(ybuf + rowc) : 0.0;
}

template <typename VT, int NH, int EPF>
__global__ void __launch_bounds__(W2_THREADS, 1)
    w_step_t(const i64* __restrict__ rptr, const int* __restrict__ cidx, const VT* __restrict__ rval,
             const double* __restrict__ xs, const double* __restrict__ sc, double* __restrict__ W, int ldw, int jprev,
             int nd, int n, const double* __restrict__ cvec, double* __restrict__ ybuf, double* __restrict__ partials,
             double* __restrict__ tpart, const int* __restrict__ halt, int h, int hp) {
  extern __shared__ __align__(16) double wt_dyn[];  // xsm[hp] | tp[WT_WARPS][hp]
  double* xsm = wt_dyn;
  const int lane = threadIdx.x & 31, hl = lane & 15, hf = lane >> 4, warp = threadIdx.x >> 5;
  double* tp = wt_dyn + (size_t)(1 + warp) * hp;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5), wg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int per = (n + nwarps - 1) / nwarps;
  per += per & 1;                       // even: a warp takes its rows two at a time
  const int r0 = min(n, wg * per), rend = min(n, r0 + per);
  int r = r0 + hf;                      // this half-warp's row; advances by 2
  TL_ENTRY_REG();
  // ---- prologue (independent of the previous kernel) ----
  {
    double2* tz = reinterpret_cast<double2*>(tp);
    for (int i = lane; i < (hp >> 1); i += 32) tz[i] = make_double2(0.0, 0.0);
  }
  const bool lazy = jprev >= 0;
  const int jq = lazy ? (jprev >> 4) : -1;
  const bool mine = lazy && hl == (jprev & 15);   // the lane that holds column jprev of the basis slice
  const int* cl = cidx + hl;
  const VT* vl = rval + hl;
  const int rlast = max(n - 1, 0);
  WTRow<VT, NH, EPF> RA, RB;
  int pb = 0, pe = 0, qb = 0, qe = 0;  // pointers of this half-warp's rows r and r + 2
  if (r < rend) pb = (int)__ldg(rptr + r), pe = (int)__ldg(rptr + r + 1);
  if (r + 2 < rend) qb = (int)__ldg(rptr + r + 2), qe = (int)__ldg(rptr + r + 3);
  wt_load<VT, NH, EPF>(RA, r < rend, pb, pe, min(r, rlast), cl, vl, W + (size_t)min(r, rlast) * ldw + hl, ybuf, lazy, nd, hl);
  pdl_sync();
  if (halt && *halt) return;
  TL_SYNC(nd);
  for (int i = threadIdx.x; i < h; i += W2_THREADS) xsm[i] = xs[i];
  const double mvs = sc[SC_MVS], fn = lazy ? sc[SC_FN] : 0.0;
  const double sinvp = lazy ? inv_or_zero(sc[SC_S]) : 0.0;  // 1 / ||y_prev||, all rows and ranks (v_step_t)
  double creg[NH], acc[NH];
#pragma unroll
  for (int q = 0; q < NH; ++q) {
    creg[q] = (hl + 16 * q < nd) ? cvec[hl + 16 * q] : 0.0;
    acc[q] = 0.0;
  }
  double n2 = 0, ax = 0;
  __syncthreads();
  // one pair of rows: issue the loads of the next pair (into NX), then reduce and scatter the current one (CU)
  auto step = [&](WTRow<VT, NH, EPF>& CU, WTRow<VT, NH, EPF>& NX) {
    {
      const int rn = min(r + 2, rlast);
      wt_load<VT, NH, EPF>(NX, r + 2 < rend, qb, qe, rn, cl, vl, W + (size_t)rn * ldw + hl, ybuf, lazy, nd, hl);
    }
    qb = qe = 0;
    if (r + 4 < rend) qb = (int)__ldg(rptr + r + 4), qe = (int)__ldg(rptr + r + 5);  // two pairs ahead
    const bool ok = r < rend;
    const double wp = sinvp * CU.yp;  // W[r, jprev]: not stored yet (the loaded value is stale)
    if (mine) {
#pragma unroll
      for (int q = 0; q < NH; ++q)
        if (q == jq) CU.wv[q] = wp;
      if (ok) W[(size_t)r * ldw + jprev] = wp;
    }
    double a = 0;
#pragma unroll
    for (int q = 0; q < EPF; ++q)
      if (hl + 16 * q < CU.len) a += (double)CU.va[q] * xsm[CU.ci[q]];
    for (int e = 16 * EPF + hl; e < CU.len; e += 16) a += (double)__ldg(vl + CU.b + e - hl) * xsm[__ldg(cl + CU.b + e - hl)];
#pragma unroll
    for (int q = 0; q < NH; ++q) a -= CU.wv[q] * creg[q];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);  // within the half-warp
    const double w = ok ? (a - mvs) - fn * wp : 0.0;
    if (hl == 0 && ok) ybuf[r] = w;
#pragma unroll
    for (int q = 0; q < NH; ++q) acc[q] += CU.wv[q] * w;
    n2 += w * w;
    ax += w;
    // t += y_r A[r, :]: the half-warps take turns, so two rows never meet on one column of the warp's copy
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      if (hf == ph) {
#pragma unroll
        for (int q = 0; q < EPF; ++q)
          if (hl + 16 * q < CU.len) tp[CU.ci[q]] = fma((double)CU.va[q], w, tp[CU.ci[q]]);
        for (int e = 16 * EPF + hl; e < CU.len; e += 16) {
          const int cc = __ldg(cl + CU.b + e - hl);
          tp[cc] = fma((double)__ldg(vl + CU.b + e - hl), w, tp[cc]);
        }
      }
      __syncwarp();
    }
    r += 2;
  };
  while (r - hf < rend) {  // warp-uniform: r - hf is the pair's first row
    step(RA, RB);
    if (r - hf >= rend) break;
    step(RB, RA);
  }
  __syncthreads();
  // the CTA's partial of t: the warp copies in warp order
  for (int col = threadIdx.x; col < h; col += W2_THREADS) {
    const double* p = wt_dyn + hp + col;
    double s = 0;
#pragma unroll
    for (int w = 0; w < WT_WARPS; ++w) s += p[(size_t)w * hp];
    tpart[(size_t)blockIdx.x * hp + col] = s;
  }
  // g = W^T y and the two scalars, as in w_step3: the halves hold partial sums of the same columns (hl + 16 q)
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
  reduce_vec_partials(nd + 2, partials);
  TL_EXIT(nd);
}

// ---- v_step_t: the V side of the step from w_step_t's per-CTA partials ----------------------------------------------
// v_step_one (registers for the CTA's rows of V, Z staged with cp.async before the dependency wait, two reductions
// through distributed shared memory) with a new front end: CTA c of the cluster first adds the CTA partials of ITS
// h / 8 rows of t, of the two scalars (||y||^2, sum y) and of every 8th column of g = W^T y -- 18 coalesced loads per
// thread, in one batch -- and, with several GPUs, exchanges that slice with the same CTA of every peer over NVLink
// (channel 1 + c of the mailbox, self-validating 16-byte words -- peer_allreduce_ll: eight exchanges in flight, ONE
// NVLink crossing per Lanczos step where there were two round trips, each behind a serial last-CTA sum).  Every rank adds the slices in rank order: V, Z, B stay bit-identical.
// Lazy rotation of W (the tall-skinny GEMM W <- W Su[:, :k] of every restart, irlb.py:246, was the largest kernel of
// the pass: 130 us of 330 per restart at 68 k cells).  W is kept as STORED columns W_s and a small matrix instead:
//     W[:, c] = W_s[:, :U0] P0[:, c]      for c <  kp   (the Ritz vectors of the last restart)
//     W[:, c] = W_s[:, U0 + c - kp]       for c >= kp   (columns generated since, stored raw)
// A restart only multiplies the small matrix (lazy_p0: P0 <- [P0 0; 0 I] Su[:, :k], S x m x k flops instead of
// n x m x k) and appends the next columns; W_s is rotated for real only when its columns run out (every 6th restart
// at m_b = 95) and for the final U.  The row kernel works in stored space unchanged -- it gets c_s = [P0 c; c_raw]
// and returns g_s = W_s^T y -- and v_step_t maps g_s to W^T y and the next coefficients back (two S x k mat-vecs
// in the cluster's first CTA).  V, Z = A^T W, B and every coefficient stay in the rotated basis as before.
struct LazyArgs {
  const double* P0 = nullptr;   // [U0][ldp] row-major, columns >= kp zero
  const double* P0t = nullptr;  // [kp][MAXB] its transpose
  int ldp = 0, U0 = 0, kp = 0;  // fresh / just rotated: U0 = kp = 0, every column raw at its own index
};

constexpr int VT_G0 = MAXB;                  // msg[0 .. 128): rows of t;  msg[128], msg[129]: ||y||^2, sum y;
constexpr int VT_MSG = VT_G0 + 2 + 16;       // msg[130 .. 146): columns c, c + 8, ... of g
static_assert(VT_MSG <= PEER_LL_MAX, "a cluster CTA's slice must fit its mailbox channel");

template <int NQ>
__global__ void __launch_bounds__(VS_THREADS, 1)
    v_step_t(const double* __restrict__ tpart, int hp, const double* __restrict__ wparts, int nblk,
             const double* __restrict__ mu, const double* __restrict__ rs, const double* __restrict__ murs,
             double* __restrict__ sc, double* __restrict__ V, int ldv, int j, int m_b, int h, double* __restrict__ ybuf,
             double* __restrict__ fbuf, double* __restrict__ xs, double* __restrict__ B, const ZArgs z,
             const int* __restrict__ halt, const PeerComm pc, const LazyArgs lz) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int slotj = lz.U0 + (j - lz.kp);  // where column j of W is stored = the width of w_step_t's vector g_s
  extern __shared__ __align__(16) double vs_dyn[];
  double(*wacc)[MAXB] = reinterpret_cast<double(*)[MAXB]>(vs_dyn);                    // [VS_WARPS][MAXB]
  double(*psum)[MAXB] = reinterpret_cast<double(*)[MAXB]>(vs_dyn + VS_WARPS * MAXB);  // [8][MAXB]
  double* zs = vs_dyn + (VS_WARPS + 8) * MAXB;                                        // [rows of this CTA][ldv]
  __shared__ double call[VS_CLUSTER][MAXB];
  __shared__ double call2[VS_CLUSTER][MAXB];
  __shared__ double ctot[MAXB];
  __shared__ double red2[VS_WARPS][2];
  __shared__ double nall[VS_CLUSTER][2];
  __shared__ double msg[VT_MSG + 14];
  __shared__ double gall[MAXB];  // CTA 0: g_s = W_s^T y, every stored column
  __shared__ double geff[MAXB];  // CTA 0: g = W^T y in the rotated basis
  __shared__ double cce[MAXB];   // CTA 0: the next coefficients in the rotated basis
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
  // the exchange's sequence number: last written by the v_step_t two kernels back, read here off the critical path
  const unsigned seq_pre = pc.world > 1 ? pc.seq[1 + crank] + 1u : 0u;
  pdl_sync();
  if (halt && *halt) {  // uniform over the cluster
    cp_async_wait_all();
    return;
  }
  TL_SYNC(256 + j);
  TL_CTAK(1, 0);
  cluster.barrier_arrive();  // matched below, before the first write into a neighbour's shared memory
  // ---- this CTA's slice of the step's vector: CTA partials of w_step_t in a fixed order ----
  {
    const int nrow = max(0, r1 - r0);
    const int col = tid & (MAXB - 1), part = tid >> 7;  // thread = (row of the slice, part): parts stride the CTAs
    psum[part][col] = (col < nrow) ? sum_partials_part<10>(tpart + r0 + col, part, 8, nblk, hp) : 0.0;  // two batches
    if (warp == 0) {  // ||y||^2 and sum y: the summation order w_finish / w_norm_exchange use (same bits)
      double t0, t1;
      sum_partials_two(wparts, nblk, slotj, t0, t1);
      if (lane == 0) msg[VT_G0] = t0, msg[VT_G0 + 1] = t1;
    } else if (warp >= 2 && warp < 18) {  // warp 2 + q: column c + 8 q of g
      const int gcol = crank + VS_CLUSTER * (warp - 2);
      const double v = gcol < slotj ? sum_partials_warp(wparts + gcol, nblk) : 0.0;
      if (lane == 0) msg[VT_G0 + warp] = v;
    }
    __syncthreads();
    if (tid < VT_G0) {
      double s = 0;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) s += psum[g8][tid];
      msg[tid] = s;
    }
    __syncthreads();
    TL_CTAK(1, 8);
    if (pc.world > 1) peer_allreduce_ll(pc, msg, VT_MSG, 1 + crank, seq_pre);  // ends with a CTA barrier
    TL_CTAK(1, 9);
  }
  // the W column this step consumes was left unnormalised: s = ||y||, u = y / s  (irlb.py:176-178, 217-219)
  const double n2w = msg[VT_G0], syw = msg[VT_G0 + 1];
  const double s = sqrt(n2w), sinvw = inv_or_zero(s), su = sinvw * syw;
  // ---- phase A: F = prep(t) - s V[:, j], partial c = V^T F ----
  double f[VS_ROWS], rr[VS_ROWS], acc[NQ] = {};
  {
    double mm[VS_ROWS], tv[VS_ROWS];
#pragma unroll
    for (int i = 0; i < VS_ROWS; ++i) {
      const int r = r0 + warp + i * VS_WARPS;
      const bool ok = r < r1;
      tv[i] = ok ? sinvw * msg[r - r0] : 0.0;  // t = A^T u
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
        const double tvr = sinvw * msg[r - r0];
        z.Z[(size_t)r * ldv + j] = tvr;
        zs[(size_t)(r - r0) * ldv + j] = tvr;
      }
    }
  }
  cluster.barrier_wait();  // every CTA of the cluster is running
  if (tid < 16) {          // this CTA's columns of g to CTA 0, which forms the coefficients of the next W column
    const int gc = crank + VS_CLUSTER * tid;
    if (gc < slotj) cluster.map_shared_rank(&gall[0], 0)[gc] = msg[VT_G0 + 2 + tid];
  }
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
  const bool ccthread = z.Z && crank == 0 && tid >= 32 && tid < 32 + z.ncc;
  double sw_col = 0;
  if (ccthread && tid - 32 != j) sw_col = z.sumW[tid - 32];
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
  if (z.Z && crank == 0) {
    // ---- g = W^T y from g_s (lazy rotation: the kp Ritz columns are combinations of the first U0 stored ones) ----
    {
      const int col = tid & (MAXB - 1), part = tid >> 7;
      double p = 0;
      if (col < lz.kp)
        for (int sidx = part; sidx < lz.U0; sidx += 8) p = fma(__ldg(lz.P0 + (size_t)sidx * lz.ldp + col), gall[sidx], p);
      psum[part][col] = p;
    }
    __syncthreads();
    if (tid < j) {
      double g = 0;
      if (tid < lz.kp) {
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) g += psum[g8][tid];
      } else {
        g = gall[lz.U0 + tid - lz.kp];
      }
      geff[tid] = g;
    }
    __syncthreads();
    if (ccthread) {
      // coefficients of the next W column; column j itself: g = 1, sumW[j] = (sum y) / ||y||
      const int col = tid - 32;
      double zsum = 0;
      for (int q = 0; q < csize; ++q) zsum += call2[q][col];
      const double sw = (col == j) ? su : sw_col;
      if (col == j) z.sumW[j] = sw;
      const double mvs = murs ? inv * tot1 : 0.0;
      double cc = inv * zsum - mvs * sw;
      if (z.prev) cc -= nrm * ((col == j) ? 1.0 : geff[col] * sinvw);
      cce[col] = cc;
      if (!z.prev) z.cvec[col] = cc;  // last column of the iteration: conv_check takes them to the next basis
    }
    __syncthreads();
    if (z.prev) {
      // ---- the coefficients in stored space for w_step_t: c_s = [P0 c[:kp] ; c[kp:]] ----
      const int sidx = tid & (MAXB - 1), part = tid >> 7;
      double p = 0;
      if (sidx < lz.U0)
        for (int col = part; col < lz.kp; col += 8) p = fma(__ldg(lz.P0t + (size_t)col * MAXB + sidx), cce[col], p);
      psum[part][sidx] = p;
      __syncthreads();
      if (tid <= slotj) {
        double cs = 0;
        if (tid < lz.U0) {
#pragma unroll
          for (int g8 = 0; g8 < 8; ++g8) cs += psum[g8][tid];
        } else {
          cs = cce[lz.kp + tid - lz.U0];
        }
        z.cvec[tid] = cs;
      }
    }
  }
  if (crank == 0 && tid == 0) {
    sc[SC_S] = s;              // ||w_j||: w_step_t normalises W[:, j] with it, B[j, j]  (irlb.py:196 / 221)
    sc[SC_SU] = su;
    B[(size_t)j * m_b + j] = s;
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

// ---- several GPUs, last column of an outer iteration: ||y|| over all ranks before w_finish -----------------------
// (B[m_b-1, m_b-1] completes B for the SVD that is forked beside the last V step; on one GPU w_finish sums the two
// columns of w_step_t's partials itself.)  One CTA: local sum, exchange on channel 0, result to out2[0..1].
__global__ void __launch_bounds__(256) w_norm_exchange(const double* __restrict__ wparts, int nblk, int col0,
                                                       double* __restrict__ out2, const int* __restrict__ halt,
                                                       const PeerComm pc) {
  if (halt && *halt) return;
  __shared__ double two[2];
  if (threadIdx.x < 32) {
    double t0, t1;
    sum_partials_two(wparts, nblk, col0, t0, t1);
    if (threadIdx.x == 0) two[0] = t0, two[1] = t1;
  }
  __syncthreads();
  if (pc.world > 1) peer_allreduce_ll(pc, two, 2, 0);
  if (threadIdx.x < 2) out2[threadIdx.x] = two[threadIdx.x];
}

// ---- last column of an outer iteration: W_s[:, slot] = y / ||y||, B[jn, jn] = ||y|| ------------------------------------
// (the columns before it were written by the w_step_t that followed them).  One GPU: the norm is summed from
// w_step_t's partials here (columns slot, slot + 1); several GPUs: n2p holds the exchanged pair.
__global__ void w_finish_t(const double* __restrict__ ybuf, const double* __restrict__ n2p, double* __restrict__ sc,
                           double* __restrict__ W, int ldw, int slot, int jn, i64 n, double* __restrict__ B, int m_b,
                           const int* __restrict__ halt, const double* __restrict__ parts, int nblk) {
  // launched with a full dependency on w_step_t; the last v_step_t of the iteration follows with a programmatic edge
  // (its prologue reads V and Z only, and its dependency wait covers this kernel)
  cudaTriggerProgrammaticLaunchCompletion();
  if (halt && *halt) return;
  __shared__ double s_n2[2];
  if (parts) {
    if (threadIdx.x < 32) {
      double t0, t1;
      sum_partials_two(parts, nblk, slot, t0, t1);
      if (threadIdx.x == 0) s_n2[0] = t0, s_n2[1] = t1;
    }
    __syncthreads();
  }
  const double n2v = parts ? s_n2[0] : n2p[0], syv = parts ? s_n2[1] : n2p[1];
  const double s = sqrt(n2v);
  const double sinv = inv_or_zero(s);
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) W[(size_t)r * ldw + slot] = sinv * ybuf[r];
  if (r == 0) {
    sc[SC_S] = s;
    sc[SC_SU] = sinv * syv;
    B[(size_t)jn * m_b + jn] = s;  // irlb.py:196 / 221
  }
}

// ---- lazy_p0: the small matrix of the lazy rotation after a restart -------------------------------------------------
// P0_new = [P0 0; 0 I] Su[:, :kn]   (S rows: the U0 stored columns behind the old Ritz vectors, then the m - kp raw ones).
// One CTA per row s, in place (row s of the product needs row s of P0 only); outputs: P0 (row stride ldp, zero
// padded), its transpose P0t, and -- for the launches that rotate W_s for real -- the same matrix with row stride
// kcp = kn rounded up to 8, the layout the rotation kernels read their matrix in.
// With crot != nullptr the CTA also leaves c_s[s] = P0_new[s, :] . crot, the coefficients of the restart's first new
// column in stored space (crot: conv_check's coefficients in the new rotated basis).
// set_restart's work (irlb.py:239-244) folded into lazy_p0, for the restarts of the two-kernel path: V[:, k] = F,
// B = diag(sv[:k]) + R[:k] in column k, the rotated column sums and the raw tail of the coefficients.
struct RestartArgs {
  double* V = nullptr;  // nullptr: lazy_p0 only
  int ldv = 0, h = 0, m = 0, k = 0;
  const double* fbuf = nullptr;
  double* B = nullptr;
  const double* sv = nullptr;
  const double* R = nullptr;
  double* sumW = nullptr;
  double* cvec = nullptr;
  const double* crot = nullptr;   // conv_check's coefficients [0, MAXB) and column sums [MAXB, 2 MAXB) in the new basis
  int U0n = 0, kpn = 0, nds = 0;  // lazy state of the coming iteration, stored index of its first new column
};

// Launched with a PROGRAMMATIC edge behind the rotation of V and Z (rotate_mma triggers at its start): nothing below
// reads what the rotation writes, so the two run side by side; only V[:, k] = F waits for the rotation to complete
// (the rotation reads whole rows of V, column k included).  The trigger at the top lets the first w_step_t of the
// iteration start its prologue (static CSR, rows of W_s) in turn; its own dependency wait covers this kernel.
__global__ void __launch_bounds__(128) lazy_p0(double* __restrict__ P0, double* __restrict__ P0t, int ldp, int U0, int kp,
                                               const double* __restrict__ Su, int m, int kn, double* __restrict__ Pmat,
                                               int kcp, const int* __restrict__ halt, const double* __restrict__ crot,
                                               double* __restrict__ cvec_s, const RestartArgs ra) {
  cudaTriggerProgrammaticLaunchCompletion();
  if (halt && *halt) return;
  __shared__ double row[MAXB];
  __shared__ double red[4];
  const int sidx = blockIdx.x, c = threadIdx.x;
  if (sidx < U0)
    for (int e = c; e < kp; e += 128) row[e] = P0[(size_t)sidx * ldp + e];
  __syncthreads();
  double v = 0.0;
  if (c < kn) {
    const double* su = Su + (size_t)c * m;  // left singular vector c (column-major)
    if (sidx < U0) {
      for (int e = 0; e < kp; ++e) v = fma(row[e], su[e], v);
    } else {
      v = su[kp + (sidx - U0)];
    }
  }
  if (c < ldp) P0[(size_t)sidx * ldp + c] = v;
  if (c < kn) P0t[(size_t)c * MAXB + sidx] = v;
  if (Pmat && c < kcp) Pmat[(size_t)sidx * kcp + c] = v;
  if (crot) {
    double p = warp_sum(c < kn ? v * crot[c] : 0.0);
    if ((c & 31) == 0) red[c >> 5] = p;
    __syncthreads();
    if (c == 0) cvec_s[sidx] = (red[0] + red[1]) + (red[2] + red[3]);
  }
  if (ra.V == nullptr) return;
  // ---- set_restart ----
  if (blockIdx.x == 0) {  // the raw tail of c_s (disjoint from the entries written above: sidx < U0n) and sum W
    for (int t = c; t < MAXB; t += 128) {
      if (t < ra.k) ra.sumW[t] = ra.crot[MAXB + t];
      if (t >= ra.U0n && t < ra.nds) ra.cvec[t] = ra.crot[ra.kpn + t - ra.U0n];
    }
  }
  const int nthr = (int)gridDim.x * 128, gid = (int)blockIdx.x * 128 + c;
  for (int i = gid; i < ra.m * ra.m; i += nthr) {
    const int r = i / ra.m, cc = i - r * ra.m;
    double b = 0.0;
    if (r < ra.k && cc == r) b = ra.sv[r];
    if (r < ra.k && cc == ra.k) b = ra.R[r];
    ra.B[i] = b;
  }
  cudaGridDependencySynchronize();  // the rotation has read its last row of V
  for (int i = gid; i < ra.h; i += nthr) ra.V[(size_t)i * ra.ldv + ra.k] = ra.fbuf[i];
}
