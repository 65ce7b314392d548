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
#include "common.cuh"

namespace adobo {

constexpr int SA_THREADS = 512;
constexpr int SA_WARPS = SA_THREADS / 32;
constexpr int SA_MAXM = 128;
constexpr int SA_MAXT = 32;
constexpr double SA_EPS = 2.220446049250313e-16;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// info[3] = 1 when the decomposition passed its self-check, 0 -> caller falls back to jacobi_svd
__global__ void __launch_bounds__(SA_THREADS, 1) svd_arrow(const double* __restrict__ B, int m, int k,
                                                           double* __restrict__ Su, double* __restrict__ svout,
                                                           double* __restrict__ Sv, int* __restrict__ info) {
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
  double* ta = sig + SA_MAXM;              // tail diagonal a[0..t)
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
  __shared__ int s_na, s_nrot, s_flag, s_bad;
  __shared__ double s_alpha, s_frob;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- S0: pieces of B ----
  for (int i = tid; i < k; i += SA_THREADS) {
    const double d = B[(size_t)i * m + i];
    dD[i] = d * d;
    wv[i] = B[(size_t)i * m + k];
  }
  for (int j = tid; j < t; j += SA_THREADS) {
    ta[j] = B[(size_t)(k + j) * m + (k + j)];
    tb[j] = (j + 1 < t) ? B[(size_t)(k + j) * m + (k + j + 1)] : 0.0;
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
  // G = T'^T : column c holds row c of T' = T[c][1..t): a_c at row c-1, b_c at row c
  for (int c = tid; c < t; c += SA_THREADS) {
    if (c >= 1) G[c * 32 + (c - 1)] = ta[c];
    if (c + 1 < t) G[c * 32 + c] = tb[c];
  }
  __syncthreads();

  // ---- S1: eigen-decomposition of T' T'^T by one-sided Jacobi on the t columns of G ----
  // round-robin pairs, one warp per pair (t/2 <= 16 pairs), lane = row (G has t-1 rows, Y has t)
  {
    const int np = t + (t & 1), npairs = np >> 1;
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
          if (q < t) {  // warp-uniform
            const double gp = G[p * 32 + lane], gq = G[q * 32 + lane];
            const double al = wsum(gp * gp), be = wsum(gq * gq), ga = wsum(gp * gq);
            if (ga * ga > 1e-30 * al * be) {
              const double d = be - al;
              const double rh = rsqrt(fma(d, d, 4.0 * ga * ga));
              const double c2 = fma(0.5 * fabs(d), rh, 0.5);
              const double rc = rsqrt(c2);
              const double cs = c2 * rc, sn = (d >= 0 ? ga : -ga) * rh * rc;
              G[p * 32 + lane] = cs * gp - sn * gq;
              G[q * 32 + lane] = sn * gp + cs * gq;
              const double yp = Y[p * 32 + lane], yq = Y[q * 32 + lane];
              Y[p * 32 + lane] = cs * yp - sn * yq;
              Y[q * 32 + lane] = sn * yp + cs * yq;
              if (lane == 0) s_flag = 1;
            }
          }
        }
        __syncthreads();
      }
      const int any = s_flag;
      __syncthreads();
      if (!any) break;
    }
    if (warp < 1)
      for (int c = 0; c < t; ++c) {
        const double g = G[c * 32 + lane];
        const double n2 = wsum(g * g);
        if (lane == 0) theta[c] = n2;
      }
  }
  __syncthreads();

  // ---- S2: poles and weights of the rank-one problem ----
  for (int c = tid; c < t; c += SA_THREADS) {
    dD[k + c] = theta[c];
    wv[k + c] = s_alpha * Y[c * 32 + 0];  // alpha * (first row of Y)
  }
  __syncthreads();
  // ---- S3: sort poles ascending (rank sort, ties by index) ----
  for (int i = tid; i < m; i += SA_THREADS) {
    const double d = dD[i];
    int r = 0;
    for (int j = 0; j < m; ++j) r += (dD[j] < d) || (dD[j] == d && j < i);
    ds[r] = d;
    ws[r] = wv[i];
    perm[r] = i;
  }
  __syncthreads();
  // ---- S4: deflation (sequential, O(m)) ----
  if (tid == 0) {
    double wn2 = 0, dmax = 0, fro = 0;
    for (int i = 0; i < m; ++i) {
      wn2 += ws[i] * ws[i];
      dmax = fmax(dmax, fabs(ds[i]));
    }
    for (int i = 0; i < m; ++i) fro += ds[i];
    s_frob = fro + wn2;  // trace of B B^T
    const double tolw = SA_EPS * sqrt(wn2);
    const double tol = 8.0 * SA_EPS * fmax(dmax, wn2);
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
  __syncthreads();
  const int na = s_na;
  for (int a = tid; a < na; a += SA_THREADS) {
    const int j = actidx[a];
    da[a] = ds[j];
    wsa[a] = ws[j];
    w2a[a] = ws[j] * ws[j];
  }
  for (int i = tid; i < m; i += SA_THREADS) mu[i] = ds[i];  // deflated coordinates keep their pole
  __syncthreads();

  // ---- S5: secular roots, one warp per root; root j lies in (da[j], da[j+1]) (last: (da[na-1], +||w||^2)) ----
  for (int j = warp; j < na; j += SA_WARPS) {
    const bool lastroot = (j + 1 >= na);
    double nrm2 = 0;
    for (int i = lane; i < na; i += 32) nrm2 += w2a[i];
    nrm2 = wsum(nrm2);
    // origin: the pole nearer to the root (sign of f at the midpoint decides)
    int k0 = j;
    double lo = 0.0, hi = nrm2;
    if (!lastroot) {
      const double gap = da[j + 1] - da[j], mid = 0.5 * gap;
      double f = 0;
      for (int i = lane; i < na; i += 32) f += w2a[i] / ((da[i] - da[j]) - mid);
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
    const double dl = da[j] - org;                            // lower pole, shifted
    const double du = lastroot ? 0.0 : da[j + 1] - org;       // upper pole, shifted
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 60; ++it) {
      double psi = 0, dpsi = 0, phi = 0, dphi = 0, abst = 0;
      for (int i = lane; i < na; i += 32) {
        const double ri = 1.0 / ((da[i] - org) - x);
        const double q = w2a[i] * ri, q2 = q * ri;
        abst += fabs(q);
        if (i <= j) {
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
      abst = wsum(abst);
      const double f = 1.0 + psi + phi;
      if (f > 0) hi = fmin(hi, x);
      else lo = fmax(lo, x);
      if (fabs(f) <= 8.0 * SA_EPS * (1.0 + abst)) break;
      if (hi - lo <= 4.0 * SA_EPS * fmax(fabs(lo), fabs(hi))) break;
      // two-pole rational model: psi ~ s + p/(dl - x), phi ~ r + q/(du - x), matched in value and slope
      const double el = dl - x;
      const double p = dpsi * el * el, s = psi - dpsi * el;
      double xn;
      if (!lastroot) {
        const double eu = du - x;
        const double q = dphi * eu * eu, r = phi - dphi * eu;
        const double c = 1.0 + s + r;
        // c (dl - x)(du - x) + p (du - x) + q (dl - x) = 0
        const double qa = c, qb = -(c * (dl + du) + p + q), qc = c * dl * du + p * du + q * dl;
        if (qa == 0.0) {
          xn = -qc / qb;
        } else {
          double disc = qb * qb - 4.0 * qa * qc;
          if (disc < 0) disc = 0;
          const double sq = sqrt(disc);
          double x1, x2;
          if (qb > 0) {
            x1 = (-qb - sq) / (2.0 * qa);
            x2 = ((-qb - sq) != 0.0) ? (2.0 * qc) / (-qb - sq) : x1;
          } else {
            x2 = (-qb + sq) / (2.0 * qa);
            x1 = ((-qb + sq) != 0.0) ? (2.0 * qc) / (-qb + sq) : x2;
          }
          xn = (x1 > lo && x1 < hi) ? x1 : x2;
        }
      } else {
        const double c = 1.0 + s;
        xn = (c != 0.0) ? dl + p / c : 0.5 * (lo + hi);
      }
      if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
      if (xn == x) break;
      x = xn;
    }
    if (lane == 0) {
      tau[j] = x;
      base[j] = k0;
      mu[actidx[j]] = org + x;
    }
  }
  __syncthreads();

  // ---- S6: Gu-Eisenstat weights: what_i^2 = prod_j (mu_j - d_i) / prod_{j != i} (d_j - d_i) ----
  for (int i = warp; i < na; i += SA_WARPS) {  // one warp per weight, lanes over the roots
    const double di = da[i];
    double p = 1.0;
    for (int j = lane; j < na; j += 32) {
      const double num = (da[base[j]] - di) + tau[j];  // mu_j - d_i
      p *= (j == i) ? num : num / (da[j] - di);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
    const double v = sqrt(fabs(p));
    if (lane == 0) what[i] = (wsa[i] >= 0) ? v : -v;
  }
  // ---- S7: eigenvectors in ORIGINAL coordinates: X[:, js] for every sorted coordinate js ----
  for (int idx = tid; idx < m * m; idx += SA_THREADS) X[idx] = 0.0;
  __syncthreads();
  for (int js = warp; js < m; js += SA_WARPS) {  // eigenvector js (sorted position)
    const int a = posact[js];
    if (a < 0) {
      if (lane == 0) X[(size_t)js * m + perm[js]] = 1.0;  // deflated: unit vector
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
        if (i < na) X[(size_t)js * m + perm[actidx[i]]] = vals[r] * inv;
      }
    }
  }
  __syncthreads();
  // close-pole rotations (rare): X <- R_1 ... R_n X, rows in original coordinates
  {
    const int nrot = s_nrot;
    for (int r = nrot - 1; r >= 0; --r) {
      const int ra = perm[rota[r]], rb = perm[rotb[r]];
      const double c = rotc[r], s = rots[r];
      for (int j = tid; j < m; j += SA_THREADS) {
        const double xa = X[(size_t)j * m + ra], xb = X[(size_t)j * m + rb];
        X[(size_t)j * m + ra] = c * xa + s * xb;
        X[(size_t)j * m + rb] = -s * xa + c * xb;
      }
      __syncthreads();
    }
  }
  // ---- S8: U = diag(I, Y) X : rows k..m-1 of every column ----
  for (int j = warp; j < m; j += SA_WARPS) {
    const double xc = (lane < t) ? X[(size_t)j * m + k + lane] : 0.0;
    double u = 0;
    for (int c = 0; c < t; ++c) {
      const double xv = __shfl_sync(0xffffffffu, xc, c);
      if (lane < t) u = fma(Y[c * 32 + lane], xv, u);
    }
    __syncwarp();
    if (lane < t) X[(size_t)j * m + k + lane] = u;
  }
  // ---- S9: singular values, descending rank ----
  for (int j = tid; j < m; j += SA_THREADS) sig[j] = sqrt(fmax(mu[j], 0.0));
  __syncthreads();
  for (int j = tid; j < m; j += SA_THREADS) {
    const double s = sig[j];
    int r = 0;
    for (int o = 0; o < m; ++o) r += (sig[o] > s) || (sig[o] == s && o < j);
    rankof[j] = r;
    svout[r] = s;
  }
  __syncthreads();
  // ---- S10: Su = U, Sv = B^T U diag(1/sv) using the sparsity of B ----
  for (int j = warp; j < m; j += SA_WARPS) {
    const double* u = X + (size_t)j * m;
    const double s = sig[j];
    const double inv = s > 0 ? 1.0 / s : 0.0;
    const int r = rankof[j];
    double zk = 0;  // row k of B^T U: sum_{i<k} z_i u_i + alpha u_k
    for (int i = lane; i < k; i += 32) zk = fma(B[(size_t)i * m + k], u[i], zk);
    zk = wsum(zk) + s_alpha * u[k];
    for (int i = lane; i < m; i += 32) {
      Su[(size_t)r * m + i] = u[i];
      double v;
      if (i < k) v = B[(size_t)i * m + i] * u[i];
      else if (i == k) v = zk;
      else v = tb[i - k - 1] * u[i - 1] + ta[i - k] * u[i];
      Sv[(size_t)r * m + i] = v * inv;
    }
  }
  // ---- S11: self-check: trace, finite values, orthogonality of neighbouring (sorted) vectors ----
  {
    double tr = 0;
    for (int j = lane; j < m; j += 32) tr += mu[j];
    if (warp == 0) {
      tr = wsum(tr);
      if (lane == 0 && !(fabs(tr - s_frob) <= 1e-11 * fabs(s_frob))) s_bad = 1;
    }
    for (int j = warp; j + 1 < m; j += SA_WARPS) {  // sorted positions j, j+1 are neighbours in mu as well
      double dot = 0, nn = 0;
      for (int i = lane; i < m; i += 32) {
        dot = fma(X[(size_t)j * m + i], X[(size_t)(j + 1) * m + i], dot);
        nn = fma(X[(size_t)j * m + i], X[(size_t)j * m + i], nn);
      }
      dot = wsum(dot);
      nn = wsum(nn);
      if (lane == 0 && !(fabs(dot) <= 1e-11 && fabs(nn - 1.0) <= 1e-11)) s_bad = 1;
    }
  }
  __syncthreads();
  if (tid == 0) {
    info[2] = -na;  // diagnostic: number of non-deflated roots (negative marks the structured path)
    info[3] = s_bad ? 0 : 1;
  }
}

size_t svd_arrow_smem(int m) {
  return sizeof(double) * ((size_t)m * m + 12 * SA_MAXM + 3 * SA_MAXT + 2 * SA_MAXM + 2 * SA_MAXT * 32) +
         sizeof(int) * 7 * SA_MAXM + 64;
}

bool svd_arrow_supported(int m, int k) { return m <= SA_MAXM && k >= 1 && (m - k) >= 2 && (m - k) <= SA_MAXT; }

void svd_arrow_prepare(int m) {
  static int prepared = 0;
  const int need = (int)svd_arrow_smem(m);
  if (need > prepared) {
    CUDA_CHECK(cudaFuncSetAttribute(svd_arrow, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
    prepared = need;
  }
}

void launch_svd_arrow(adobo_ctx* c, const double* B, int m, int k, double* Su, double* sv, double* Sv, int* info,
                      cudaStream_t st) {
  WORK(c, 24.0 * m * m);
  LAUNCH(c, "svd_arrow", (svd_arrow<<<1, SA_THREADS, svd_arrow_smem(m), st>>>(B, m, k, Su, sv, Sv, info)));
}

}  // namespace adobo
