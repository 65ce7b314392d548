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
#include "common.cuh"

namespace adobo {

enum { SC_S = 0, SC_FN, SC_FNINV, SC_SU, SC_MVS, SC_SMAX, SC_NRM2, SC_AUX, SC_COUNT = 16 };
constexpr int MAXB = 128;  // widest basis the dot kernels handle (4 columns per lane)
constexpr double EPS2 = 2.0 * 2.220446049250313e-16;

__device__ __forceinline__ double inv_or_zero(double x) { return x > EPS2 ? 1.0 / x : 0.0; }  // irlb.py:84-92

// ---- ordered cross-CTA reductions ------------------------------------------------------------
// vector of nd (<= 128) sums; acc[t] is this lane's partial for column lane + 32 t
__device__ __forceinline__ void reduce_vec_blocks(double (&acc)[4], int nd, double* __restrict__ partials,
                                                  double* __restrict__ out, unsigned* __restrict__ counter) {
  __shared__ double sh[8][MAXB];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 4; ++t) sh[warp][lane + 32 * t] = acc[t];
  __syncthreads();
  if (threadIdx.x < nd) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
    partials[(size_t)blockIdx.x * MAXB + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < nd) {
      double s = 0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(partials + (size_t)b * MAXB + threadIdx.x);
      out[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

// two scalars
__device__ __forceinline__ void reduce_two_blocks(double a, double b, double* __restrict__ partials,
                                                  double* __restrict__ out, unsigned* __restrict__ counter) {
  __shared__ double sh2[8][2];
  __shared__ bool is_last2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh2[warp][0] = a;
    sh2[warp][1] = b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh2[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last2 = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last2 && warp == 0) {
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
}

// ---- K5  A^T u on the CSC copy: one CTA per gene column ---------------------------------------
template <typename VT>
__global__ void __launch_bounds__(256) spmvT_csc(const i64* __restrict__ cptr, const int* __restrict__ ridx,
                                                 const VT* __restrict__ cval, const double* __restrict__ u,
                                                 double* __restrict__ t) {
  __shared__ double red[8];
  const int col = blockIdx.x;
  const i64 b = cptr[col], e1 = cptr[col + 1];
  double s = 0;
#pragma unroll 4
  for (i64 e = b + threadIdx.x; e < e1; e += 256) s += (double)__ldg(cval + e) * u[__ldg(ridx + e)];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tt += red[w];
    t[col] = tt;
  }
}

// ---- F = A_s^T w - s V[:,j]  and  c = V[:, :nd]^T F  (irlb.py:182-190, first half of orthog) ----
__global__ void __launch_bounds__(256) v_prep_dots(const double* __restrict__ t, const double* __restrict__ mu,
                                                   const double* __restrict__ rs, const double* __restrict__ sc,
                                                   const double* __restrict__ V, int ldv, int j, int nd, int h,
                                                   double* __restrict__ ybuf, double* __restrict__ partials,
                                                   double* __restrict__ cvec, unsigned* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
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
  reduce_vec_blocks(acc, nd, partials, cvec, counter);
}

// ---- K4  w = A_s x - fn W[:,jprev]  and  c = W[:, :nd]^T w  (irlb.py:165 / 204-216) -------------
template <typename VT>
__global__ void __launch_bounds__(256) spmv_dots(const i64* __restrict__ rptr, const int* __restrict__ cidx,
                                                 const VT* __restrict__ rval, const double* __restrict__ xs,
                                                 const double* __restrict__ sc, const double* __restrict__ W,
                                                 int ldw, int jprev, int nd, i64 n, double* __restrict__ ybuf,
                                                 double* __restrict__ partials, double* __restrict__ cvec,
                                                 unsigned* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  const double mvs = sc[SC_MVS], fn = sc[SC_FN];
  double acc[4] = {0, 0, 0, 0};
  for (i64 r = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n; r += warps) {
    const i64 b = rptr[r], e1 = rptr[r + 1];
    double a = 0;
#pragma unroll 4
    for (i64 e = b + lane; e < e1; e += 32) a += (double)__ldg(rval + e) * xs[__ldg(cidx + e)];
    a = warp_sum(a);
    const double* wr = W + (size_t)r * ldw;
    double w = a - mvs;
    if (jprev >= 0) w -= fn * wr[jprev];
    if (lane == 0) ybuf[r] = w;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = lane + 32 * q;
      if (col < nd) acc[q] += wr[col] * w;
    }
  }
  if (nd > 0) reduce_vec_blocks(acc, nd, partials, cvec, counter);
}

// ---- y -= Q[:, :nd] c ;  nrm2 = y.y ;  aux = coef.y  (second half of orthog + the norm) ---------
__global__ void __launch_bounds__(256) update_norm(const double* __restrict__ Q, int ld, int nd,
                                                   const double* __restrict__ cvec, double* __restrict__ ybuf,
                                                   const double* __restrict__ coef, i64 rows,
                                                   double* __restrict__ partials, double* __restrict__ out2,
                                                   unsigned* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  double creg[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) creg[q] = (lane + 32 * q < nd) ? cvec[lane + 32 * q] : 0.0;
  double n2 = 0, ax = 0;
  for (i64 r = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < rows; r += warps) {
    double y = ybuf[r];
    if (nd > 0) {
      const double* qr = Q + (size_t)r * ld;
      double p = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = lane + 32 * q;
        if (col < nd) p += qr[col] * creg[q];
      }
      p = warp_sum(p);
      y -= p;
      if (lane == 0) ybuf[r] = y;
    }
    n2 += y * y;
    ax += (coef ? coef[r] : 1.0) * y;
  }
  reduce_two_blocks(n2, ax, partials, out2, counter);
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
    if (j >= 0) {
      B[(size_t)j * m_b + j] = sc[SC_S];               // irlb.py:196 / 221
      if (j < m_b - 1) B[(size_t)j * m_b + j + 1] = fn;  // irlb.py:197
    }
  }
}

// ---- normalise w, store W[:, jn] and the contiguous copy the next A^T u gathers from -----------
__global__ void w_finish(const double* __restrict__ ybuf, double* __restrict__ sc, double* __restrict__ W, int ldw,
                         int jn, i64 n, double* __restrict__ wcur) {
  const double s = sqrt(sc[SC_NRM2]);
  const double sinv = inv_or_zero(s);
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) {
    const double w = sinv * ybuf[r];
    W[(size_t)r * ldw + jn] = w;
    wcur[r] = w;
  }
  if (r == 0) {
    sc[SC_S] = s;
    sc[SC_SU] = sinv * sc[SC_AUX];
  }
}

// ---- one-sided Jacobi SVD of B (m x m, row-major) in one CTA -----------------------------------
// G = B (column-major in shared memory), columns rotated pairwise (round-robin ordering, one pair
// per half-warp) until mutually orthogonal: B Vr = U diag(sv).  Outputs column-major, singular
// values descending.  vglobal != null keeps the accumulated Vr in global memory (large m).
__global__ void __launch_bounds__(1024, 1) jacobi_svd(const double* __restrict__ B, int m, double* __restrict__ Su,
                                                      double* __restrict__ svout, double* __restrict__ Sv,
                                                      double* vglobal, int* __restrict__ info) {
  extern __shared__ double smem[];
  double* G = smem;
  volatile double* Vr = vglobal ? vglobal : (smem + (size_t)m * m);
  __shared__ int rotated;
  __shared__ double svs[256];
  const int tid = threadIdx.x;
  const int hw = tid >> 4, sl = tid & 15;
  const int nhw = blockDim.x >> 4;
  for (int idx = tid; idx < m * m; idx += blockDim.x) {
    const int col = idx / m, row = idx - col * m;
    G[idx] = B[(size_t)row * m + col];
    Vr[idx] = (row == col) ? 1.0 : 0.0;
  }
  if (tid == 0) rotated = 0;
  __syncthreads();
  const int np = m + (m & 1);
  const int npairs = np >> 1;
  int sweeps = 0;
  for (; sweeps < 60; ++sweeps) {
    for (int rd = 0; rd < np - 1; ++rd) {
      for (int base = 0; base < npairs; base += nhw) {
        const int pr = base + hw;
        int p = -1, q = -1;
        if (pr < npairs) {
          const int i0 = pr, i1 = np - 1 - pr;
          p = (i0 == 0) ? 0 : 1 + (((i0 - 1 - rd) % (np - 1)) + (np - 1)) % (np - 1);
          q = 1 + (((i1 - 1 - rd) % (np - 1)) + (np - 1)) % (np - 1);
          if (p > q) {
            const int tmp = p;
            p = q;
            q = tmp;
          }
        }
        const bool act = (pr < npairs) && (q < m);
        double al = 0, be = 0, ga = 0;
        if (act) {
          const double* gp = G + (size_t)p * m;
          const double* gq = G + (size_t)q * m;
          for (int i = sl; i < m; i += 16) {
            const double a = gp[i], b = gq[i];
            al += a * a;
            be += b * b;
            ga += a * b;
          }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        if (act && fabs(ga) > 1e-15 * sqrt(al * be)) {
          const double zeta = (be - al) / (2.0 * ga);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          double* gp = G + (size_t)p * m;
          double* gq = G + (size_t)q * m;
          volatile double* vp = Vr + (size_t)p * m;
          volatile double* vq = Vr + (size_t)q * m;
          for (int i = sl; i < m; i += 16) {
            const double a = gp[i], b = gq[i];
            gp[i] = cs * a - sn * b;
            gq[i] = sn * a + cs * b;
            const double c = vp[i], d = vq[i];
            vp[i] = cs * c - sn * d;
            vq[i] = sn * c + cs * d;
          }
          if (sl == 0) rotated = 1;
        }
      }
      __syncthreads();
    }
    const int any = rotated;
    __syncthreads();
    if (tid == 0) rotated = 0;
    __syncthreads();
    if (!any) break;
  }
  // singular values = column norms
  for (int base = 0; base < m; base += nhw) {
    const int col = base + hw;
    double s = 0;
    if (col < m)
      for (int i = sl; i < m; i += 16) s += G[(size_t)col * m + i] * G[(size_t)col * m + i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (col < m && sl == 0) svs[col] = sqrt(s);
  }
  __syncthreads();
  for (int base = 0; base < m; base += nhw) {
    const int col = base + hw;
    if (col < m) {
      const double sv = svs[col];
      int rank = 0;
      for (int o = 0; o < m; ++o) rank += (svs[o] > sv) || (svs[o] == sv && o < col);
      const double inv = sv > 0 ? 1.0 / sv : 0.0;
      for (int i = sl; i < m; i += 16) {
        Su[(size_t)rank * m + i] = G[(size_t)col * m + i] * inv;
        Sv[(size_t)rank * m + i] = Vr[(size_t)col * m + i];
      }
      if (sl == 0) svout[rank] = sv;
    }
  }
  if (tid == 0) info[2] = sweeps;
}

// ---- residuals, convergence count and the next k (irlb.py:225-234) ------------------------------
__global__ void conv_check(const double* __restrict__ Su, const double* __restrict__ sv, double* __restrict__ sc,
                           int m, int nu, int k, int it, double tol, double* __restrict__ R, int* __restrict__ info) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const double fn = sc[SC_FN];
  double smax = sv[0];
  if (it > 0) smax = fmax(smax, sc[SC_SMAX]);
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double r = fn * Su[(size_t)i * m + (m - 1)];
    R[i] = r;
    if (i < nu && fabs(r) < tol * smax) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sc[SC_SMAX] = smax;
    int kn = max(cnt + nu, k);
    kn = min(kn, m - 3);
    info[0] = cnt;
    info[1] = kn;
  }
}

// ---- Q[:, :kc] <- Q[:, :m] S[:, :kc]  (irlb.py:238, 246, 249-250); S column-major m x m ----------
__global__ void __launch_bounds__(256) rotate_basis(double* __restrict__ Q, int ld, i64 rows, const double* __restrict__ S,
                                                    int m, int kc, double* __restrict__ out, int ldo,
                                                    const double* __restrict__ colscale) {
  extern __shared__ double sm[];
  double* Ssm = sm;                                     // [m][kc]
  double* rowbuf = sm + (size_t)m * kc + (threadIdx.x >> 5) * m;  // per warp [m]
  for (int idx = threadIdx.x; idx < m * kc; idx += blockDim.x) {
    const int c = idx / m, i = idx - c * m;
    Ssm[(size_t)i * kc + c] = S[(size_t)c * m + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 r = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < rows; r += warps) {
    double* qr = Q + (size_t)r * ld;
    for (int i = lane; i < m; i += 32) rowbuf[i] = qr[i];
    __syncwarp();
    for (int c = lane; c < kc; c += 32) {
      double a = 0;
      for (int i = 0; i < m; ++i) a += rowbuf[i] * Ssm[(size_t)i * kc + c];
      if (out)
        out[(size_t)r * ldo + c] = colscale ? a * colscale[c] : a;
      else
        qr[c] = a;
    }
    __syncwarp();
  }
}

// ---- V[:,k] = F ; B = diag(sv[:k]) + R[:k] in column k  (irlb.py:239-244) ------------------------
__global__ void set_restart(double* __restrict__ V, int ldv, int h, const double* __restrict__ fbuf,
                            double* __restrict__ B, int m, int k, const double* __restrict__ sv,
                            const double* __restrict__ R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < h) V[(size_t)i * ldv + k] = fbuf[i];
  if (i < m * m) {
    const int r = i / m, c = i - r * m;
    double v = 0.0;
    if (r < k && c == r) v = sv[r];
    if (r < k && c == k) v = R[r];
    B[i] = v;
  }
}

__global__ void mul_vec(const double* a, const double* b, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

static unsigned row_blocks(adobo_ctx* c, i64 rows) {
  i64 b = (rows + 7) / 8;
  i64 cap = (i64)c->sm_count * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

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
  const int ld = (m_b + 15) & ~15;
  const bool centered = (mu != nullptr);
  cudaStream_t st = c->stream;

  DevBuf<double> V, W, B, sc, tbuf, ybuf_h, ybuf_n, fbuf, xs, wcur, cvec, partials, Su, Sv, svd_s, R, murs, vglob;
  DevBuf<unsigned> counter;
  DevBuf<int> info;
  V.ensure((size_t)h * ld);
  W.ensure((size_t)std::max<i64>(n, 1) * ld);
  B.ensure((size_t)m_b * m_b);
  sc.ensure(SC_COUNT);
  tbuf.ensure((size_t)h);
  ybuf_h.ensure((size_t)h);
  ybuf_n.ensure((size_t)std::max<i64>(n, 1));
  fbuf.ensure((size_t)h);
  xs.ensure((size_t)h);
  wcur.ensure((size_t)std::max<i64>(n, 1));
  cvec.ensure(MAXB);
  const unsigned gb_n = row_blocks(c, n), gb_h = row_blocks(c, h);
  partials.ensure((size_t)std::max(gb_n, gb_h) * MAXB);
  Su.ensure((size_t)m_b * m_b);
  Sv.ensure((size_t)m_b * m_b);
  svd_s.ensure((size_t)m_b);
  R.ensure((size_t)m_b);
  counter.ensure(1);
  info.ensure(4);
  CUDA_CHECK(cudaMemsetAsync(V.p, 0, sizeof(double) * (size_t)h * ld, st));
  CUDA_CHECK(cudaMemsetAsync(W.p, 0, sizeof(double) * (size_t)std::max<i64>(n, 1) * ld, st));
  CUDA_CHECK(cudaMemsetAsync(B.p, 0, sizeof(double) * m_b * m_b, st));
  CUDA_CHECK(cudaMemsetAsync(sc.p, 0, sizeof(double) * SC_COUNT, st));
  CUDA_CHECK(cudaMemsetAsync(counter.p, 0, sizeof(unsigned), st));
  CUDA_CHECK(cudaMemsetAsync(cvec.p, 0, sizeof(double) * MAXB, st));
  if (centered) {
    murs.ensure((size_t)h);
    LAUNCH(c, "mul_vec", (mul_vec<<<(h + 255) / 256, 256, 0, st>>>(mu, rs, murs.p, h)));
  }
  // SVD kernel resources
  size_t svd_smem = sizeof(double) * 2 * (size_t)m_b * m_b;
  double* vg = nullptr;
  if (svd_smem > 200 * 1024) {
    vglob.ensure((size_t)m_b * m_b);
    vg = vglob.p;
    svd_smem = sizeof(double) * (size_t)m_b * m_b;
  }
  CUDA_CHECK(cudaFuncSetAttribute(jacobi_svd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)svd_smem));
  const size_t rot_smem = sizeof(double) * ((size_t)m_b * m_b + 8 * (size_t)m_b);
  CUDA_CHECK(cudaFuncSetAttribute(rotate_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot_smem));

  const unsigned fin_h = (unsigned)((h + 255) / 256), fin_n = (unsigned)std::max<i64>(1, (n + 255) / 256);
  int* h_info = (int*)c->pinned;

  // ---- start vector (irlb.py:151-153): V[:,0] = v0 / ||v0|| ----
  CUDA_CHECK(cudaMemcpyAsync(ybuf_h.p, v0_host, sizeof(double) * h, cudaMemcpyHostToDevice, st));
  LAUNCH(c, "update_norm", (update_norm<<<gb_h, 256, 0, st>>>(V.p, ld, 0, cvec.p, ybuf_h.p,
                                                            centered ? murs.p : nullptr, h, partials.p,
                                                            sc.p + SC_NRM2, counter.p)));
  LAUNCH(c, "v_finish", (v_finish<<<fin_h, 256, 0, st>>>(ybuf_h.p, rs, sc.p, V.p, ld, 0, m_b, h, fbuf.p, xs.p, B.p, -1,
                                                       centered ? 1 : 0)));

  auto w_column = [&](int jprev, int nd, int jn) {
    // W[:, jn] = normalise( orthog( A_s x - fn W[:, jprev], W[:, :nd] ) )
    WORK(c, (double)A.nnz * (4.0 + sizeof(VT)) + 8.0 * (n + 1) + 8.0 * n * (nd + (jprev >= 0 ? 1 : 0) + 1) + 8.0 * h);
    LAUNCH(c, "spmv_dots", (spmv_dots<VT><<<gb_n, 256, 0, st>>>(A.rptr.p, A.cidx.p, A.rval.p, xs.p, sc.p, W.p, ld, jprev,
                                                              nd, n, ybuf_n.p, partials.p, cvec.p, counter.p)));
    if (nd > 0) allreduce_sum_f64(c, cvec.p, (size_t)nd);
    WORK(c, 8.0 * n * (nd + 2));
    LAUNCH(c, "update_norm", (update_norm<<<gb_n, 256, 0, st>>>(W.p, ld, nd, cvec.p, ybuf_n.p, nullptr, n, partials.p,
                                                              sc.p + SC_NRM2, counter.p)));
    allreduce_sum_f64(c, sc.p + SC_NRM2, 2);
    LAUNCH(c, "w_finish", (w_finish<<<fin_n, 256, 0, st>>>(ybuf_n.p, sc.p, W.p, ld, jn, n, wcur.p)));
  };

  IrlbResult res;
  int it = 0, k = nu, nmult = 0;
  while (it < maxit) {
    int j = (it > 0) ? k : 0;
    w_column(-1, j, j);
    ++nmult;
    while (j < m_b) {
      WORK(c, (double)A.nnz * (4.0 + sizeof(VT)) + 8.0 * (h + 1) + 8.0 * h + 8.0 * n);
      LAUNCH(c, "spmvT_csc", (spmvT_csc<VT><<<h, 256, 0, st>>>(A.cptr.p, A.ridx.p, A.cval.p, wcur.p, tbuf.p)));
      ++nmult;
      allreduce_sum_f64(c, tbuf.p, (size_t)h);
      WORK(c, 8.0 * h * (j + 1 + 4));
      LAUNCH(c, "v_prep_dots", (v_prep_dots<<<gb_h, 256, 0, st>>>(tbuf.p, mu, rs, sc.p, V.p, ld, j, j + 1, h, ybuf_h.p,
                                                                partials.p, cvec.p, counter.p)));
      WORK(c, 8.0 * h * (j + 1 + 3));
      LAUNCH(c, "update_norm_v", (update_norm<<<gb_h, 256, 0, st>>>(V.p, ld, j + 1, cvec.p, ybuf_h.p,
                                                                centered ? murs.p : nullptr, h, partials.p,
                                                                sc.p + SC_NRM2, counter.p)));
      LAUNCH(c, "v_finish", (v_finish<<<fin_h, 256, 0, st>>>(ybuf_h.p, rs, sc.p, V.p, ld, j + 1, m_b, h, fbuf.p, xs.p,
                                                           B.p, j, centered ? 1 : 0)));
      if (j < m_b - 1) {
        w_column(j, j + 1, j + 1);
        ++nmult;
      }
      ++j;
    }
    LAUNCH(c, "jacobi_svd", (jacobi_svd<<<1, 1024, svd_smem, st>>>(B.p, m_b, Su.p, svd_s.p, Sv.p, vg, info.p)));
    LAUNCH(c, "conv_check", (conv_check<<<1, 128, 0, st>>>(Su.p, svd_s.p, sc.p, m_b, nu, k, it, tol, R.p, info.p)));
    CUDA_CHECK(cudaMemcpyAsync(h_info, info.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const int conv = h_info[0];
    if (conv >= nu) break;  // irlb.py:232-236
    k = h_info[1];
    LAUNCH(c, "rotate_basis", (rotate_basis<<<row_blocks(c, h), 256, rot_smem, st>>>(V.p, ld, h, Sv.p, m_b, k, nullptr, 0,
                                                                                 nullptr)));
    LAUNCH(c, "set_restart", (set_restart<<<(std::max(h, m_b * m_b) + 255) / 256, 256, 0, st>>>(V.p, ld, h, fbuf.p, B.p,
                                                                                            m_b, k, svd_s.p, R.p)));
    WORK(c, 8.0 * n * (m_b + k));
    LAUNCH(c, "rotate_basis", (rotate_basis<<<row_blocks(c, n), 256, rot_smem, st>>>(W.p, ld, n, Su.p, m_b, k, nullptr, 0,
                                                                                 nullptr)));
    ++it;
  }
  // ---- U = W Su[:, :nu] (x diag(s) when var_weigh), V = V Sv[:, :nu]  (irlb.py:249-250, dr.py:164-168) ----
  outU.ensure((size_t)std::max<i64>(n, 1) * nu);
  outV.ensure((size_t)h * nu);
  outS.ensure((size_t)nu);
  LAUNCH(c, "rotate_basis", (rotate_basis<<<row_blocks(c, n), 256, rot_smem, st>>>(W.p, ld, n, Su.p, m_b, nu, outU.p, nu,
                                                                               var_weigh ? svd_s.p : nullptr)));
  LAUNCH(c, "rotate_basis", (rotate_basis<<<row_blocks(c, h), 256, rot_smem, st>>>(V.p, ld, h, Sv.p, m_b, nu, outV.p, nu,
                                                                               nullptr)));
  CUDA_CHECK(cudaMemcpyAsync(outS.p, svd_s.p, sizeof(double) * nu, cudaMemcpyDeviceToDevice, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  res.steps = it;
  res.nmult = nmult;
  return res;
}

template IrlbResult run_irlb<float>(adobo_ctx*, SparsePair<float>&, const double*, const double*, int, double, int,
                                    const double*, bool, DevBuf<double>&, DevBuf<double>&, DevBuf<double>&);
template IrlbResult run_irlb<double>(adobo_ctx*, SparsePair<double>&, const double*, const double*, int, double, int,
                                     const double*, bool, DevBuf<double>&, DevBuf<double>&, DevBuf<double>&);

}  // namespace adobo
