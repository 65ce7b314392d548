// svd_arrow -- SVD of the restarted Lanczos matrix B by its structure instead of by Jacobi sweeps.
//
// After a restart (irlb.py:238-246) B is not a general matrix:
//
//        B = [ D   z e1^T ]     D = diag(sigma_1..sigma_k)   (the kept Ritz values)
//            [ 0   T      ]     z = residual spike in column k, T = t x t upper bidiagonal tail
//                               (t = m - k <= 20 new Lanczos steps, alpha = T[0,0])
//
// so that          B B^T = diag(D^2, T' T'^T) + [z; alpha e1] [z; alpha e1]^T ,   T' = T without its
// first column.  The left singular vectors therefore come from (1) the eigen-decomposition of the
// SMALL t x t matrix T' T'^T = Y Theta Y^T (one-sided Jacobi on T'^T, a few columns), and (2) ONE
// diagonal-plus-rank-one eigenproblem diag(D^2, Theta) + w w^T, w = [z; alpha Y^T e1], solved like
// a divide-and-conquer merge step: deflation, the m secular roots found independently (one warp per
// root, rational two-pole iteration in coordinates shifted to the nearest pole), Gu-Eisenstat
// recomputation of the weights so that the eigenvectors w_i / (d_i - mu_j) are orthogonal to working
// precision.  U = diag(I, Y) X, sv = sqrt(mu), and the right vectors follow from the sparsity of B:
// V = B^T U diag(1/sv) costs O(m^2).  Everything is O(m^2) work spread over one CTA -- ~10x less
// than the ~7 Jacobi sweeps of jacobi_svd (irlb.cu), which stays in use for the first iteration
// (B bidiagonal, no diagonal block) and as the fallback when the cheap self-check at the end fails.
//
// Replaces numpy.linalg.svd(B) at adobo/irlbpy/irlb.py:224 for restarts (it > 0).
#include <cooperative_groups.h>

#include "common.cuh"

namespace adobo {
namespace cg = cooperative_groups;

constexpr int SA_THREADS = 512;
constexpr int SA_WARPS = SA_THREADS / 32;
constexpr int SA_MAXM = 128;
constexpr int SA_MAXT = 32;
constexpr double SA_EPS = 2.220446049250313e-16;

// stage timing for profiles/micro/svd_arrow_stages.cu (never defined in the library build)
#ifdef ADOBO_SA_TIMING
__device__ long long sa_stamps[16];
#define SA_STAMP(i)                              \
  do {                                           \
    __syncthreads();                             \
    if (threadIdx.x == 0) sa_stamps[i] = clock64(); \
  } while (0)
#else
#define SA_STAMP(i)
#endif

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum / product over the 8 lanes of a group (four groups per warp); every lane of the warp must take part
__device__ __forceinline__ double gsum8(double v) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double gprod8(double v) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// info[3] = 1 when the decomposition passed its self-check, 0 -> caller falls back to jacobi_svd
__global__ void __launch_bounds__(SA_THREADS, 1) svd_arrow(const double* __restrict__ B, int m, int k,
                                                           double* __restrict__ Su, double* __restrict__ svout,
                                                           double* __restrict__ Sv, int* __restrict__ info,
                                                           const LastDiag ldg) {
  extern __shared__ __align__(16) double sm[];
  const int t = m - k;
  double* X = sm;                          // [m][m] column-major: X[j*m + i] = coordinate i of eigenvector j
  double* dD = X + (size_t)m * m;          // poles, original coordinate order           [SA_MAXM]
  double* wv = dD + SA_MAXM;               // weights, original order
  double* ds = wv + SA_MAXM;               // poles sorted ascending
  double* ws = ds + SA_MAXM;               // weights in sorted order (after deflation rotations)
  double* da = ws + SA_MAXM;               // active poles (ascending)
  double* w2a = da + SA_MAXM;              // active weights squared
  double* wsa = w2a + SA_MAXM;             // active weights (signed)
  double* tau = wsa + SA_MAXM;             // root j = da[base[j]] + tau[j]
  double* mu = tau + SA_MAXM;              // eigenvalue of every sorted coordinate
  double* what = mu + SA_MAXM;             // recomputed weights
  double* sig = what + SA_MAXM;            // singular value of eigenvector j (sorted coordinate j)
  double* dg = sig + SA_MAXM;              // diagonal of D (original order): B[i][i]
  double* rootv = dg + SA_MAXM;            // secular root of active position a (published by its owner CTA)
  double* ta = rootv + SA_MAXM;            // tail diagonal a[0..t)
  double* tb = ta + SA_MAXT;               // tail super-diagonal b[0..t) (b[t-1] = 0)
  double* theta = tb + SA_MAXT;            // eigenvalues of T' T'^T
  double* rotc = theta + SA_MAXT;          // deflation rotations (close poles)       [SA_MAXM]
  double* rots = rotc + SA_MAXM;
  double* G = rots + SA_MAXM;              // [t][32] columns of T'^T
  double* Y = G + SA_MAXT * 32;            // [t][32] accumulated rotations: Y[c*32 + r]
  int* perm = reinterpret_cast<int*>(Y + SA_MAXT * 32);   // sorted position -> original coordinate [SA_MAXM]
  int* actidx = perm + SA_MAXM;            // active position -> sorted position
  int* base = actidx + SA_MAXM;            // root j is measured from active pole base[j]
  int* rota = base + SA_MAXM;              // rotation pairs (sorted positions)
  int* rotb = rota + SA_MAXM;
  int* rankof = rotb + SA_MAXM;            // eigenvector (sorted coordinate) -> descending rank of its sv
  int* posact = rankof + SA_MAXM;          // sorted position -> active position or -1
  int* aorig = posact + SA_MAXM;           // active position -> original coordinate
  __shared__ int s_na, s_nrot, s_flag, s_bad;
  __shared__ double s_alpha, s_frob, s_gtiny;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  if (info[4] != 0) return;  // halt flag of the Lanczos driver (conv_check): uniform over the cluster
  if (crank == 0 && tid == 0) info[3] = 1;
  cluster.barrier_arrive();  // matched by the wait before the first write into a neighbour's shared memory

  SA_STAMP(0);
  // ---- S0: pieces of B ----
  for (int i = tid; i < k; i += SA_THREADS) {
    const double d = B[(size_t)i * m + i];
    dg[i] = d;
    dD[i] = d * d;
    wv[i] = B[(size_t)i * m + k];
  }
  const bool own_last = ldg.parts != nullptr || ldg.n2p != nullptr;  // B[m-1][m-1] is not in memory yet (LastDiag)
  for (int j = tid; j < t; j += SA_THREADS) {
    if (!(own_last && j == t - 1)) ta[j] = B[(size_t)(k + j) * m + (k + j)];
    tb[j] = (j + 1 < t) ? B[(size_t)(k + j) * m + (k + j + 1)] : 0.0;
  }
  if (own_last && warp == SA_THREADS / 32 - 1) {
    double t0;
    if (ldg.parts) {  // the order of sum_partials_two: 16 parts of every 16th CTA (ascending), then the parts in order
      double p = 0;
      if (lane < 16) {
        for (int base = lane; base < ldg.nblk; base += 160) {
          double v[10];
#pragma unroll
          for (int i = 0; i < 10; ++i) {
            const int b = base + 16 * i;
            v[i] = b < ldg.nblk ? __ldcg(ldg.parts + (size_t)b * ldg.stride + ldg.col0) : 0.0;
          }
#pragma unroll
          for (int i = 0; i < 10; ++i) p += v[i];
        }
      }
      t0 = __shfl_sync(0xffffffffu, p, 0);
#pragma unroll
      for (int l = 1; l < 16; ++l) t0 += __shfl_sync(0xffffffffu, p, l);
    } else {
      t0 = ldg.n2p[0];
    }
    if (lane == 0) ta[t - 1] = sqrt(t0);
  }
  for (int idx = tid; idx < SA_MAXT * 32; idx += SA_THREADS) {
    G[idx] = 0.0;
    Y[idx] = ((idx >> 5) == (idx & 31)) ? 1.0 : 0.0;
  }
  if (tid == 0) {
    s_bad = 0;
    s_nrot = 0;
  }
  __syncthreads();
  if (tid == 0) s_alpha = ta[0];
  if (warp == 1) {  // a column of G below 1e-15 ||G||_F is a null direction of T' T'^T (there is always one)
    double f2 = 0;
    if (lane < t) f2 = (lane >= 1 ? ta[lane] * ta[lane] : 0.0) + (lane + 1 < t ? tb[lane] * tb[lane] : 0.0);
    f2 = wsum(f2);
    if (lane == 0) s_gtiny = 1e-30 * f2;
  }
  // G = T' (t rows, t-1 columns): column c is column c+1 of T -- b_c at row c, a_{c+1} at row c+1
  for (int c = tid; c + 1 < t; c += SA_THREADS) {
    G[c * 32 + c] = tb[c];
    G[c * 32 + c + 1] = ta[c + 1];
  }
  __syncthreads();

  SA_STAMP(1);
  // ---- S1: eigen-decomposition of T' T'^T = Y diag(theta) Y^T from the left singular vectors of T' ----
  // One-sided Jacobi on the t-1 columns of T' (round-robin pairs, no rotation matrix to accumulate): at
  // convergence the columns are sigma_c y_c.  T' T'^T has rank t-1; its null vector completes the basis.
  // The usual restart adds only a few Lanczos steps (t <= 8): then ONE warp holds a whole round -- eight lanes
  // (rows) per pair, up to four pairs -- and runs without CTA barriers; larger tails use a warp per pair.
  // A sweep whose rotations were all below 1e-8 ends the iteration (quadratic convergence).
  {
    const int nc = t - 1;
    const int np = nc + (nc & 1), npairs = np >> 1;
    if (t <= 8) {
      if (warp == 0 && nc >= 2) {
        const int grp = lane >> 3, row = lane & 7;
        for (int sweep = 0; sweep < 40; ++sweep) {
          int flags = 0;
          for (int rd = 0; rd < np - 1; ++rd) {
            int p = 0, q = 0;
            if (grp < npairs) {
              const int i0 = grp, i1 = np - 1 - grp;
              p = (i0 == 0) ? 0 : 1 + ((i0 - 1 - rd) % (np - 1) + (np - 1)) % (np - 1);
              q = 1 + ((i1 - 1 - rd) % (np - 1) + (np - 1)) % (np - 1);
              if (p > q) {
                const int tmp = p;
                p = q;
                q = tmp;
              }
            }
            const bool act = grp < npairs && q < nc;
            if (!act) p = q = 0;
            const double gp = G[p * 32 + row], gq = G[q * 32 + row];
            const double al = gsum8(gp * gp), be = gsum8(gq * gq), ga = gsum8(gp * gq);
            double cs, sn;
            int f = jacobi_angle(ga, al, be, cs, sn);
            if (!act || !(al > s_gtiny && be > s_gtiny)) f = 0;  // a null column has no direction to converge to
            if (f) {
              G[p * 32 + row] = cs * gp - sn * gq;
              G[q * 32 + row] = sn * gp + cs * gq;
            }
            flags |= f;
            __syncwarp();
          }
          if (!__any_sync(0xffffffffu, (flags & 2) != 0)) break;
        }
      }
      __syncthreads();
    } else {
      for (int sweep = 0; sweep < 40; ++sweep) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (int rd = 0; rd < np - 1; ++rd) {
          if (warp < npairs) {
            const int i0 = warp, i1 = np - 1 - warp;
            int p = (i0 == 0) ? 0 : 1 + ((i0 - 1 - rd) % (np - 1) + (np - 1)) % (np - 1);
            int q = 1 + ((i1 - 1 - rd) % (np - 1) + (np - 1)) % (np - 1);
            if (p > q) {
              const int tmp = p;
              p = q;
              q = tmp;
            }
            if (q < nc) {  // warp-uniform
              const double gp = G[p * 32 + lane], gq = G[q * 32 + lane];
              const double al = wsum(gp * gp), be = wsum(gq * gq), ga = wsum(gp * gq);
              double cs, sn;
              int f = jacobi_angle(ga, al, be, cs, sn);
              if (!(al > s_gtiny && be > s_gtiny)) f = 0;
              if (f) {
                G[p * 32 + lane] = cs * gp - sn * gq;
                G[q * 32 + lane] = sn * gp + cs * gq;
                if ((f & 2) && lane == 0) s_flag = 1;
              }
            }
          }
          __syncthreads();
        }
        const int any = s_flag;
        __syncthreads();
        if (!any) break;
      }
    }
    if (warp == 0) {
      // theta_c = ||column c||^2, y_c = column c / sigma_c  (a lane per column)
      if (lane < nc) {
        double n2 = 0;
        for (int r = 0; r < t; ++r) {
          const double g = G[lane * 32 + r];
          n2 = fma(g, g, n2);
        }
        theta[lane] = n2;
        const double inv = 1.0 / sqrt(n2);
        for (int r = 0; r < t; ++r) Y[lane * 32 + r] = G[lane * 32 + r] * inv;
      }
      if (lane == nc) theta[lane] = 0.0;
      __syncwarp();
      // null vector: (I - Y Y^T) e_i for the unit vector with the largest remainder (>= 1/t, no cancellation)
      double pr = 2.0;
      if (lane < t) {
        pr = 0;
        for (int c = 0; c < nc; ++c) pr = fma(Y[c * 32 + lane], Y[c * 32 + lane], pr);
      }
      double best = pr;
      int bi = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      double v = 0;
      if (lane < t) {
        v = (lane == bi) ? 1.0 : 0.0;
        for (int c = 0; c < nc; ++c) v = fma(-Y[c * 32 + bi], Y[c * 32 + lane], v);
      }
      const double vn = wsum(v * v);
      if (lane < t) Y[nc * 32 + lane] = v / sqrt(vn);
    }
  }
  __syncthreads();

  SA_STAMP(2);
  // ---- S2: poles and weights of the rank-one problem ----
  for (int c = tid; c < t; c += SA_THREADS) {
    dD[k + c] = theta[c];
    wv[k + c] = s_alpha * Y[c * 32 + 0];  // alpha * (first row of Y)
  }
  __syncthreads();
  SA_STAMP(3);
  // ---- S3: sort poles ascending (rank sort, ties by index) ----
  for (int i = tid; i < m; i += SA_THREADS) {
    const double d = dD[i];
    int r = 0;
#pragma unroll 8
    for (int j = 0; j < m; ++j) r += (dD[j] < d) || (dD[j] == d && j < i);
    ds[r] = d;
    ws[r] = wv[i];
    perm[r] = i;
  }
  __syncthreads();
  SA_STAMP(4);
  // ---- S4: deflation.  Negligible weights drop out; the usual case has no (numerically) equal poles and
  // is resolved by warp 0 with ballots, the rotation of equal poles stays sequential (rare) ----
  if (warp == 0) {
    double wn2 = 0, dmax = 0, fro = 0;
    for (int i = lane; i < m; i += 32) {
      wn2 += ws[i] * ws[i];
      dmax = fmax(dmax, fabs(ds[i]));
      fro += ds[i];
    }
    wn2 = wsum(wn2);
    fro = wsum(fro);
    dmax = wmax(dmax);
    if (lane == 0) s_frob = fro + wn2;  // trace of B B^T
    const double tolw = SA_EPS * sqrt(wn2);
    const double tol = 8.0 * SA_EPS * fmax(dmax, wn2);
    bool anyclose = false;
    int carry = -1, cnt = 0;
    for (int q0 = 0; q0 < m; q0 += 32) {  // warp-uniform trip count
      const int j = q0 + lane;
      const bool keep = j < m && fabs(ws[j]) > tolw;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      const unsigned lower = bal & ((1u << lane) - 1u);
      const int prev = lower ? q0 + 31 - __clz(lower) : carry;
      const bool close = keep && prev >= 0 && fabs(ds[j] - ds[prev]) <= tol;
      anyclose = __any_sync(0xffffffffu, close) || anyclose;
      if (j < m) posact[j] = keep ? cnt + __popc(lower) : -1;
      if (keep) actidx[cnt + __popc(lower)] = j;
      cnt += __popc(bal);
      if (bal) carry = q0 + 31 - __clz(bal);
    }
    if (lane == 0) {
      s_na = cnt;
      s_nrot = 0;
    }
    __syncwarp();
    if (anyclose && lane == 0) {
      int last = -1, na = 0, nrot = 0;
      for (int j = 0; j < m; ++j) posact[j] = -1;
      for (int j = 0; j < m; ++j) {
        if (!(fabs(ws[j]) > tolw)) continue;  // negligible weight: (ds[j], e_j) is an eigenpair as it is
        if (last >= 0 && fabs(ds[j] - ds[last]) <= tol) {
          // two (numerically) equal poles: rotate the weight of `last` into j, `last` deflates
          const double r = hypot(ws[last], ws[j]);
          const double c = ws[j] / r, s = ws[last] / r;
          rota[nrot] = last;
          rotb[nrot] = j;
          rotc[nrot] = c;
          rots[nrot] = s;
          ++nrot;
          ws[last] = 0.0;
          ws[j] = r;
          posact[last] = -1;
          --na;  // `last` was appended below in an earlier iteration: remove it (it is the newest entry)
        }
        actidx[na] = j;
        posact[j] = na;
        ++na;
        last = j;
      }
      s_na = na;
      s_nrot = nrot;
    }
  }
  __syncthreads();
  const int na = s_na;
  for (int a = tid; a < na; a += SA_THREADS) {
    const int j = actidx[a];
    da[a] = ds[j];
    aorig[a] = perm[j];
    wsa[a] = ws[j];
    w2a[a] = ws[j] * ws[j];
  }
  for (int i = tid; i < m; i += SA_THREADS) mu[i] = ds[i];  // deflated coordinates keep their pole
  __syncthreads();

  cluster.barrier_wait();  // every CTA of the cluster is running
  SA_STAMP(5);
  // From here on the O(m^2) stages are split over the CTAs of the cluster: every CTA has run S0-S4 on its own
  // copy (same inputs, same code, same result), CTA c takes roots / weights / eigenvectors 16c .. 16c+15 -- one
  // per warp, so that with eight CTAs every stage is a single pass -- and the m-vectors the next stage needs
  // (roots, weights) are written into the shared memory of all CTAs of the cluster.
  // ---- S5: secular roots, one warp per root; root j lies in (da[j], da[j+1]) (last: (da[na-1], +||w||^2)) ----
  for (int j = crank * SA_WARPS + warp; j < na; j += csize * SA_WARPS) {
    const bool lastroot = (j + 1 >= na);
#ifdef ADOBO_SA_TIMING
    const long long tq0 = clock64();
#endif
    // this lane's (up to four) poles and squared weights stay in registers for the whole iteration
    double dai[4], w2[4];
    bool low[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = lane + 32 * r;
      dai[r] = (i < na) ? da[i] : 1e300;  // padding: weight 0 at a far-away pole
      w2[r] = (i < na) ? w2a[i] : 0.0;
      low[r] = i <= j;
    }
    const double daj = da[j];
    double nrm2 = wsum((w2[0] + w2[1]) + (w2[2] + w2[3]));
    // origin: the pole nearer to the root (sign of f at the midpoint decides)
    int k0 = j;
    double lo = 0.0, hi = nrm2;
    if (!lastroot) {  // warp-uniform
      const double mid = 0.5 * (da[j + 1] - daj);
      double f = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) f += w2[r] / ((dai[r] - daj) - mid);
      f = 1.0 + wsum(f);
      if (f < 0) {
        k0 = j + 1;
        lo = -mid;
        hi = 0.0;
      } else {
        lo = 0.0;
        hi = mid;
      }
    }
    const double org = da[k0];
    const double dl = daj - org;                         // lower pole, shifted
    const double du = lastroot ? 0.0 : da[j + 1] - org;  // upper pole, shifted
#pragma unroll
    for (int r = 0; r < 4; ++r) dai[r] -= org;
    double x = 0.5 * (lo + hi);
#ifdef ADOBO_SA_TIMING
    const long long tq1 = clock64();
    int nev = 0;
#endif
    for (int it = 0; it < 60; ++it) {
#ifdef ADOBO_SA_TIMING
      ++nev;
#endif
      double psi = 0, dpsi = 0, phi = 0, dphi = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double ri = 1.0 / (dai[r] - x);
        const double q = w2[r] * ri, q2 = q * ri;
        if (low[r]) {
          psi += q;
          dpsi += q2;
        } else {
          phi += q;
          dphi += q2;
        }
      }
      psi = wsum(psi);
      dpsi = wsum(dpsi);
      phi = wsum(phi);
      dphi = wsum(dphi);
      const double abst = phi - psi;  // every term of psi is <= 0 and every term of phi >= 0 inside (dl, du)
      const double f = 1.0 + psi + phi;
      if (f > 0) hi = fmin(hi, x);
      else lo = fmax(lo, x);
      if (fabs(f) <= 8.0 * SA_EPS * (1.0 + abst)) break;
      if (hi - lo <= 4.0 * SA_EPS * fmax(fabs(lo), fabs(hi))) break;
      // two-pole rational model: psi ~ s + p/(dl - x), phi ~ r + q/(du - x), matched in value and slope
      const double el = dl - x;
      const double p = dpsi * el * el, sp = psi - dpsi * el;
      double xn;
      if (!lastroot) {
        const double eu = du - x;
        const double q = dphi * eu * eu, r = phi - dphi * eu;
        const double c = 1.0 + sp + r;
        // c (dl - x)(du - x) + p (du - x) + q (dl - x) = 0
        const double qa = c, qb = -(c * (dl + du) + p + q), qc = c * dl * du + p * du + q * dl;
        if (qa == 0.0) {
          xn = -qc / qb;
        } else {
          double disc = qb * qb - 4.0 * qa * qc;
          if (disc < 0) disc = 0;
          const double sq = sqrt(disc);
          // the two roots without cancellation: tq / (2 qa) and 2 qc / tq, tq = -qb -+ sq
          const double tq = (qb > 0) ? (-qb - sq) : (-qb + sq);
          const double xa = tq / (2.0 * qa);
          const double xb = (tq != 0.0) ? (2.0 * qc) / tq : xa;
          const double x1 = (qb > 0) ? xa : xb, x2 = (qb > 0) ? xb : xa;
          xn = (x1 > lo && x1 < hi) ? x1 : x2;
        }
      } else {
        const double c = 1.0 + sp;
        xn = (c != 0.0) ? dl + p / c : 0.5 * (lo + hi);
      }
      if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
      if (xn == x) break;
      x = xn;
    }
#ifdef ADOBO_SA_TIMING
    if (crank == 0 && warp == 3 && lane == 0) {
      sa_stamps[13] = tq1 - tq0;
      sa_stamps[14] = clock64() - tq1;
      sa_stamps[15] = nev;
    }
#endif
    if (lane < csize) {  // lane r publishes the root to CTA r
      cluster.map_shared_rank(tau, lane)[j] = x;
      cluster.map_shared_rank(base, lane)[j] = k0;
      cluster.map_shared_rank(rootv, lane)[j] = org + x;
    }
  }
  cluster.sync();
  for (int a = tid; a < na; a += SA_THREADS) mu[actidx[a]] = rootv[a];  // read again after the next barrier
  SA_STAMP(6);
  // ---- S6: Gu-Eisenstat weights: what_i^2 = prod_j (mu_j - d_i) / prod_{j != i} (d_j - d_i) ----
  for (int i = crank * SA_WARPS + warp; i < na; i += csize * SA_WARPS) {  // one warp per weight, lanes over the roots
    const double di = da[i];
    double p = 1.0;
    for (int j = lane; j < na; j += 32) {
      const double num = (da[base[j]] - di) + tau[j];  // mu_j - d_i
      p *= (j == i) ? num : num / (da[j] - di);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
    const double v = sqrt(fabs(p));
    if (lane < csize) cluster.map_shared_rank(what, lane)[i] = (wsa[i] >= 0) ? v : -v;
  }
  // own eigenvector columns start from zero (nobody else touches this CTA's X)
  for (int js = crank * SA_WARPS + warp; js < m; js += csize * SA_WARPS)
    for (int i = lane; i < m; i += 32) X[(size_t)js * m + i] = 0.0;
  cluster.sync();
  SA_STAMP(7);
  // ---- S7: eigenvectors in ORIGINAL coordinates: X[:, js] for the sorted coordinates js of this CTA ----
  for (int js = crank * SA_WARPS + warp; js < m; js += csize * SA_WARPS) {
    const int a = posact[js];
    double* xcol = X + (size_t)js * m;
    if (a < 0) {
      if (lane == 0) xcol[perm[js]] = 1.0;  // deflated: unit vector
    } else {
      const double org = da[base[a]], tj = tau[a];
      double vals[4];
      double n2 = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        double v = 0;
        if (i < na) v = what[i] / ((da[i] - org) - tj);
        vals[r] = v;
        n2 += v * v;
      }
      n2 = wsum(n2);
      const double inv = rsqrt(n2);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        if (i < na) xcol[aorig[i]] = vals[r] * inv;
      }
    }
    __syncwarp();
    // close-pole rotations (rare): X <- R_1 ... R_n X, rows in original coordinates
    for (int r = s_nrot - 1; r >= 0; --r) {
      if (lane == 0) {
        const int ra = perm[rota[r]], rb = perm[rotb[r]];
        const double c = rotc[r], sn = rots[r];
        const double xa = xcol[ra], xb = xcol[rb];
        xcol[ra] = c * xa + sn * xb;
        xcol[rb] = -sn * xa + c * xb;
      }
    }
    __syncwarp();
    // ---- S8: U = diag(I, Y) X : rows k..m-1 of the column ----
    const double xc = (lane < t) ? xcol[k + lane] : 0.0;
    double u = 0;
    for (int c = 0; c < t; ++c) {
      const double xv = __shfl_sync(0xffffffffu, xc, c);
      if (lane < t) u = fma(Y[c * 32 + lane], xv, u);
    }
    __syncwarp();
    if (lane < t) xcol[k + lane] = u;
  }
  // ---- S9: singular values ----
  for (int j = tid; j < m; j += SA_THREADS) sig[j] = sqrt(fmax(mu[j], 0.0));
  cluster.sync();  // every column of X is final (the self-check reads one column of the next CTA)
  SA_STAMP(8);
  SA_STAMP(9);
  SA_STAMP(10);
  // ---- S10: Su = U, Sv = B^T U diag(1/sv) using the sparsity of B; rows in descending order of sv ----
  for (int j = crank * SA_WARPS + warp; j < m; j += csize * SA_WARPS) {
    const double* u = X + (size_t)j * m;
    const double s = sig[j];
    int r = 0;  // descending rank, ties by index
    for (int o = lane; o < m; o += 32) r += (sig[o] > s) || (sig[o] == s && o < j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) svout[r] = s;
    const double inv = s > 0 ? 1.0 / s : 0.0;
    double zk = 0;  // row k of B^T U: sum_{i<k} z_i u_i + alpha u_k
    for (int i = lane; i < k; i += 32) zk = fma(wv[i], u[i], zk);
    zk = wsum(zk) + s_alpha * u[k];
    for (int i = lane; i < m; i += 32) {
      Su[(size_t)r * m + i] = u[i];
      double v;
      if (i < k) v = dg[i] * u[i];
      else if (i == k) v = zk;
      else v = tb[i - k - 1] * u[i - 1] + ta[i - k] * u[i];
      Sv[(size_t)r * m + i] = v * inv;
    }
    // ---- S11: self-check: finite values, orthogonality of neighbouring (sorted) vectors; the neighbour of the
    // last column of a CTA lives in the next CTA ----
    if (j + 1 < m) {
      const int owner = ((j + 1) / SA_WARPS) % csize;
      const double* un = (owner == crank ? X : cluster.map_shared_rank(X, owner)) + (size_t)(j + 1) * m;
      double dot = 0, nn = 0;
      for (int i = lane; i < m; i += 32) {
        dot = fma(u[i], un[i], dot);
        nn = fma(u[i], u[i], nn);
      }
      dot = wsum(dot);
      nn = wsum(nn);
      if (lane == 0 && !(fabs(dot) <= 1e-11 && fabs(nn - 1.0) <= 1e-11)) s_bad = 1;
    }
  }
  SA_STAMP(11);
  if (crank == 0 && warp == SA_WARPS - 1) {  // trace of B B^T
    double tr = 0;
    for (int j = lane; j < m; j += 32) tr += mu[j];
    tr = wsum(tr);
    if (lane == 0 && !(fabs(tr - s_frob) <= 1e-11 * fabs(s_frob))) s_bad = 1;
  }
  __syncthreads();
  SA_STAMP(12);
  if (tid == 0) {
    if (crank == 0) info[2] = -na;  // diagnostic: number of non-deflated roots (negative marks the structured path)
    if (s_bad) atomicExch(&info[3], 0);  // info[3] was set to 1 by CTA 0 before the first cluster barrier
  }
  cluster.sync();  // nobody leaves while a neighbour may still read its X
}

size_t svd_arrow_smem(int m) {
  return sizeof(double) * ((size_t)m * m + 14 * SA_MAXM + 3 * SA_MAXT + 2 * SA_MAXM + 2 * SA_MAXT * 32) +
         sizeof(int) * 8 * SA_MAXM + 64;
}

bool svd_arrow_supported(int m, int k) { return m <= SA_MAXM && k >= 1 && (m - k) >= 2 && (m - k) <= SA_MAXT; }

int svd_arrow_cluster() {  // CTAs per cluster (test hook: ADOBO_B200_SVD_CLUSTER = 1, 2, 4; default 8)
  const char* e = getenv("ADOBO_B200_SVD_CLUSTER");
  const int v = e ? atoi(e) : 8;
  return (v == 1 || v == 2 || v == 4) ? v : 8;
}

void svd_arrow_prepare(int m) {
  static int prepared = 0;
  const int need = (int)svd_arrow_smem(m);
  if (need > prepared) {
    CUDA_CHECK(cudaFuncSetAttribute(svd_arrow, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
    prepared = need;
  }
}

void launch_svd_arrow(adobo_ctx* c, const double* B, int m, int k, double* Su, double* sv, double* Sv, int* info,
                      cudaStream_t st, LastDiag ldg) {
  WORK(c, 24.0 * m * m);
  const int cs = svd_arrow_cluster();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs);
  cfg.blockDim = dim3(SA_THREADS);
  cfg.dynamicSmemBytes = svd_arrow_smem(m);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LAUNCH(c, "svd_arrow", CUDA_CHECK(cudaLaunchKernelEx(&cfg, svd_arrow, B, m, k, Su, sv, Sv, info, ldg)));
}

}  // namespace adobo
