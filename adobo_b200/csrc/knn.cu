// K8  exact k-nearest-neighbour search over the PCA embedding.
//
// Replaces adobo/clustering.py:47-52 (sklearn NearestNeighbors(algorithm='ball_tree',
// metric='euclidean').kneighbors on the fit set itself; +1 to make the ids 1-based).  The
// reference evaluates true fp64 Euclidean distances; the result here is the same neighbour set
// and order (ties: see tests/parity.py), obtained in three steps:
//   1. knn_shortlist  tiled fp32 distance "GEMM" (||r||^2 - 2 q.r on centred fp32 coordinates,
//                     64 queries x 128 references per tile, 4x8 register blocking) with a fused
//                     threshold filter and a per-query sorted shortlist of KP > k candidates;
//   2. knn_rerank     fp64 direct sum of squared differences for the KP candidates on the
//                     ORIGINAL fp64 coordinates, exact ordering, and a certificate: if the k-th
//                     exact distance is not provably below everything the fp32 filter dropped
//                     (rounding bound eps), the query is flagged;
//   3. knn_exact_rows fp64 brute force for flagged queries only (duplicates / massive ties).
// There is no approximate result: a query is either certified or recomputed exactly.
#include <cfloat>
#include <cstdlib>
#include <climits>

#include "common.cuh"
#include "knn_common.cuh"

namespace adobo {

constexpr int QT = 64;    // queries per CTA
constexpr int RT = 128;   // references per shared-memory tile
constexpr int MAXP = 128; // embedding dimensions supported

// ---- preparation ------------------------------------------------------------------------------
// centre = column mean of the first `m` rows (any fixed vector works; it only tightens eps)
__global__ void __launch_bounds__(1024) knn_center(const double* __restrict__ x, i64 m, int p, double* __restrict__ ctr) {
  __shared__ double sh[8][MAXP];
  const int c = threadIdx.x & 127, grp = threadIdx.x >> 7;
  double s = 0;
  if (c < p)
    for (i64 r = grp; r < m; r += 8) s += x[r * p + c];
  sh[grp][c] = s;
  __syncthreads();
  if (threadIdx.x < p) {
    double t = 0;
    for (int g2 = 0; g2 < 8; ++g2) t += sh[g2][threadIdx.x];
    ctr[threadIdx.x] = t / (double)m;
  }
}

// fp32 centred coordinates in tile-transposed layout refT[(tile*P + d)*RT + i], norms, max norm
__global__ void knn_prep(const double* __restrict__ x, i64 n, i64 npad, int p, int P, const double* __restrict__ ctr,
                         float* __restrict__ refT, float* __restrict__ norms, unsigned* __restrict__ maxnorm_bits) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  const i64 tile = i / RT;
  const int li = (int)(i - tile * RT);
  float* dst = refT + (size_t)tile * P * RT + li;
  float nrm = 0.f;
  if (i < n) {
    for (int d = 0; d < p; ++d) {
      const float v = (float)(x[i * p + d] - ctr[d]);
      dst[(size_t)d * RT] = v;
      nrm = fmaf(v, v, nrm);
    }
    for (int d = p; d < P; ++d) dst[(size_t)d * RT] = 0.f;
    norms[i] = nrm;
    atomicMax(maxnorm_bits, __float_as_uint(nrm));  // nrm >= 0: bit pattern order == value order
  } else {
    for (int d = 0; d < P; ++d) dst[(size_t)d * RT] = 0.f;
    norms[i] = INFINITY;  // padding never enters a shortlist
  }
}

// ---- 1. shortlist ------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(256, 1) knn_shortlist(const float* __restrict__ refT, const float* __restrict__ norms,
                                                        i64 npad, int P, i64 qtile0, float* __restrict__ out_key,
                                                        int* __restrict__ out_idx, i64 q_begin, i64 q_end) {
  constexpr int KP = 32 * E;
  extern __shared__ __align__(16) unsigned char raw[];
  float* Rs = reinterpret_cast<float*>(raw);        // [P][RT]
  float* Qs = Rs + (size_t)P * RT;                  // [P][QT]
  float* rn = Qs + (size_t)P * QT;                  // [RT]
  float* thr = rn + RT;                             // [QT]
  int* cnt = reinterpret_cast<int*>(thr + QT);      // [QT]
  Cand* pool = reinterpret_cast<Cand*>(cnt + QT);   // [QT][RT]
  Cand* lists = pool + (size_t)QT * RT;             // [QT][KP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const i64 qtile = qtile0 + blockIdx.x;            // 64-query tile index (global numbering)
  const i64 qbase = qtile * QT;
  // query tile = half of reference tile qbase / RT
  {
    const i64 rt = qbase / RT;
    const int half = (int)((qbase - rt * RT) / QT);
    const float* src = refT + (size_t)rt * P * RT + half * QT;
    for (int idx = tid; idx < P * QT; idx += 256) {
      const int d = idx / QT, i = idx - d * QT;
      Qs[idx] = src[(size_t)d * RT + i];
    }
  }
  for (int i = tid; i < QT; i += 256) {
    thr[i] = INFINITY;
    cnt[i] = 0;
  }
  for (int i = tid; i < QT * KP; i += 256) {
    lists[i].key = INFINITY;
    lists[i].idx = INT_MAX;
  }
  __syncthreads();
  const i64 ntiles = npad / RT;
  for (i64 t = 0; t < ntiles; ++t) {
    // ---- stage the reference tile (contiguous P*RT floats) ----
    {
      const float4* src = reinterpret_cast<const float4*>(refT + (size_t)t * P * RT);
      float4* dst = reinterpret_cast<float4*>(Rs);
      for (int idx = tid; idx < P * RT / 4; idx += 256) dst[idx] = __ldg(src + idx);
      if (tid < RT) rn[tid] = norms[t * RT + tid];
    }
    __syncthreads();
    // ---- 4 queries x 8 references per thread ----
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int d = 0; d < P; ++d) {
      const float4 q = *reinterpret_cast<const float4*>(Qs + (size_t)d * QT + ty * 4);
      const float4 r0 = *reinterpret_cast<const float4*>(Rs + (size_t)d * RT + tx * 4);
      const float4 r1 = *reinterpret_cast<const float4*>(Rs + (size_t)d * RT + 64 + tx * 4);
      const float qa[4] = {q.x, q.y, q.z, q.w};
      const float rb[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(qa[a], rb[b], acc[a][b]);
    }
    // ---- threshold filter -> per-query pool ----
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ql = ty * 4 + a;
      const float th = thr[ql];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int rl = (b < 4) ? (tx * 4 + b) : (64 + tx * 4 + (b - 4));
        const float key = fmaf(-2.f, acc[a][b], rn[rl]);  // ||r||^2 - 2 q.r  (+ ||q||^2 is constant per query)
        if (key < th) {
          const int slot = atomicAdd(&cnt[ql], 1);
          pool[(size_t)ql * RT + slot].key = key;
          pool[(size_t)ql * RT + slot].idx = (int)(t * RT + rl);
        }
      }
    }
    __syncthreads();
    // ---- drain pools into the sorted lists: warp w owns queries 8w .. 8w+7 ----
    for (int qq = 0; qq < 8; ++qq) {
      const int ql = warp * 8 + qq;
      const int c = cnt[ql];
      if (c == 0) continue;
      float key[E];
      int idx[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        key[e] = lists[(size_t)ql * KP + lane * E + e].key;
        idx[e] = lists[(size_t)ql * KP + lane * E + e].idx;
      }
      for (int s = 0; s < c; ++s) {
        const Cand cd = pool[(size_t)ql * RT + s];
        list_insert<float, E>(key, idx, cd.key, cd.idx, lane);
      }
#pragma unroll
      for (int e = 0; e < E; ++e) {
        lists[(size_t)ql * KP + lane * E + e].key = key[e];
        lists[(size_t)ql * KP + lane * E + e].idx = idx[e];
      }
      if (lane == 31) thr[ql] = key[E - 1];
      if (lane == 0) cnt[ql] = 0;
    }
    __syncthreads();
  }
  for (int i = tid; i < QT * KP; i += 256) {
    const i64 q = qbase + i / KP;
    if (q >= q_begin && q < q_end) {
      out_key[(size_t)(q - q_begin) * KP + (i % KP)] = lists[i].key;
      out_idx[(size_t)(q - q_begin) * KP + (i % KP)] = lists[i].idx;
    }
  }
}

// ---- exact ordering of 32*E fp64 candidates, self first on a zero-distance tie --------------------
template <int E>
__device__ __forceinline__ void rank_and_write(const double (&d)[E], const int (&id)[E], int lane, i64 qglobal, int k,
                                               i64* __restrict__ out1, int* __restrict__ nn0, double* kth) {
  double kd = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int myself = (id[e] == (int)qglobal) ? 0 : 1;
    int rank = 0;
    for (int o = 0; o < 32; ++o) {
#pragma unroll
      for (int f = 0; f < E; ++f) {
        const double od = __shfl_sync(0xffffffffu, d[f], o);
        const int oi = __shfl_sync(0xffffffffu, id[f], o);
        const int oself = (oi == (int)qglobal) ? 0 : 1;
        const bool before = od < d[e] || (od == d[e] && (oself < myself || (oself == myself && oi < id[e])));
        rank += before ? 1 : 0;
      }
    }
    if (rank < k) {
      if (out1) out1[rank] = (i64)id[e] + 1;  // 1-based, clustering.py:51
      nn0[rank] = id[e];
    }
    if (rank == k - 1) kd = d[e];
  }
  // exactly one (lane, e) holds rank k-1
  double tot = kd;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  *kth = tot;
}

// ---- 2. fp64 re-rank + certificate: one warp per query --------------------------------------------
template <int E>
__global__ void __launch_bounds__(256) knn_rerank(const double* __restrict__ x, int p, const float* __restrict__ sk,
                                                  const int* __restrict__ si, const float* __restrict__ norms,
                                                  const unsigned* __restrict__ maxnorm_bits, i64 q_begin, i64 nq,
                                                  i64 n, int k, i64* __restrict__ out1, int* __restrict__ nn0,
                                                  int* __restrict__ flagged, int* __restrict__ nflag) {
  constexpr int KP = 32 * E;
  const int lane = threadIdx.x & 31;
  const i64 warps = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 ql = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); ql < nq; ql += warps) {
    const i64 q = q_begin + ql;
    const double* xq = x + q * p;
    double d[E];
    int id[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      id[e] = si[(size_t)ql * KP + lane * E + e];
      if (id[e] >= n) id[e] = INT_MAX;  // padding rows of the last reference tile (only when n < KP)
      if (id[e] == INT_MAX) {
        d[e] = INFINITY;
      } else {
        const double* xr = x + (i64)id[e] * p;
        double s = 0;
        for (int t = 0; t < p; ++t) {
          const double df = xq[t] - xr[t];
          s += df * df;
        }
        d[e] = s;
      }
    }
    double kth;
    rank_and_write<E>(d, id, lane, q, k, out1 ? out1 + (size_t)ql * k : nullptr, nn0 + (size_t)ql * k, &kth);
    // certificate: every reference the fp32 filter dropped has computed distance >= T32
    const float lastkey = __shfl_sync(0xffffffffu, sk[(size_t)ql * KP + 31 * E + (E - 1)], 0);
    const double qn = (double)norms[q];
    const double T32 = (double)lastkey + qn;
    // Rounding bound of the fp32 scores, C eps (|q|^2 + |r|^2), over the references the filter dropped.  A reference
    // with |r| > |q| + sqrt(kth) is farther than the k-th candidate by the triangle inequality whatever its computed
    // score, so only norms up to that radius need to be covered -- not the largest norm of the whole embedding (a
    // few outlying cells made that bound fail thousands of certificates at 80 k cells: 87 ms of knn_exact_rows).
    const double rq = (sqrt(qn) + sqrt(kth)) * (1.0 + 1e-5);
    const double eps = (3.0 * p + 16.0) * 5.9604644775390625e-08 * (qn + fmin((double)__uint_as_float(*maxnorm_bits), rq * rq));
    const bool ok = (lastkey == INFINITY) || (kth < T32 - eps);
    if (!ok && lane == 0) flagged[atomicAdd(nflag, 1)] = (int)ql;
  }
}

// ---- 3. exact fp64 brute force for flagged queries: one CTA per query ------------------------------
__global__ void __launch_bounds__(256) knn_exact_rows(const double* __restrict__ x, i64 n, int p,
                                                      const int* __restrict__ flagged, i64 q_begin, int k,
                                                      i64* __restrict__ out1, int* __restrict__ nn0) {
  constexpr int E = 2;
  __shared__ double skey[8][64];
  __shared__ int sidx[8][64];
  __shared__ double xq[MAXP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const i64 ql = flagged[blockIdx.x];
  const i64 q = q_begin + ql;
  for (int t = threadIdx.x; t < p; t += 256) xq[t] = x[q * p + t];
  __syncthreads();
  double key[E] = {INFINITY, INFINITY};
  int idx[E] = {INT_MAX, INT_MAX};
  for (i64 base = (i64)warp * 32; base < n; base += 256) {  // warp-uniform trip count
    const i64 r = base + lane;
    double dd = INFINITY;
    int ri = INT_MAX;
    if (r < n) {
      const double* xr = x + r * p;
      double s = 0;
      for (int t = 0; t < p; ++t) {
        const double df = xq[t] - xr[t];
        s += df * df;
      }
      dd = s;
      ri = (int)r;
      // self wins a zero-distance tie: order by (d, notself, idx) -> encode by keeping self's index
      // comparison separate below
    }
    // self-first rule: give the query itself index -1 inside the list
    if (ri == (int)q) ri = -1;
    const double wk = __shfl_sync(0xffffffffu, key[E - 1], 31);
    const int wi = __shfl_sync(0xffffffffu, idx[E - 1], 31);
    unsigned m = __ballot_sync(0xffffffffu, cand_less(dd, ri, wk, wi));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const double nk = __shfl_sync(0xffffffffu, dd, src);
      const int ni = __shfl_sync(0xffffffffu, ri, src);
      list_insert<double, E>(key, idx, nk, ni, lane);
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    skey[warp][lane * E + e] = key[e];
    sidx[warp][lane * E + e] = idx[e];
  }
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < 8; ++w)
      for (int s = 0; s < 64; ++s) list_insert<double, E>(key, idx, skey[w][s], sidx[w][s], lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int rank = lane * E + e;
      if (rank < k) {
        const int id = (idx[e] == -1) ? (int)q : idx[e];
        if (out1) out1[(size_t)ql * k + rank] = (i64)id + 1;
        nn0[(size_t)ql * k + rank] = id;
      }
    }
  }
}


// ---- 4. other ball-tree metrics (clustering.py:47-48 forwards `distance` to sklearn's ball tree) --------------------
// Exact brute force in fp64 on the CUDA cores: 32 queries per CTA (four per warp), references streamed through a
// transposed shared-memory tile (lane = reference, the query coordinate is a broadcast), coordinates accumulated in
// index order like sklearn's DistanceMetric does, so the distances have the same bits.  Order: (distance, index),
// the query itself first.  n^2 p fp64 operations: fine at 68 k cells (tens of ms), not the path for 1 M.
constexpr int MQ = 32;   // queries per CTA
constexpr int MQW = 4;   // queries per warp
constexpr int MRT = 64;  // references per tile (two per lane)
constexpr int MRS = MRT + 1;

template <int METRIC>
struct MetricAcc {
  double a = 0, b = 0;
  __device__ __forceinline__ void add(double x, double y) {
    if (METRIC == ADOBO_METRIC_EUCLIDEAN) {
      const double d = x - y;
      a = fma(d, d, a);
    } else if (METRIC == ADOBO_METRIC_MANHATTAN) {
      a += fabs(x - y);
    } else if (METRIC == ADOBO_METRIC_CHEBYSHEV) {
      a = fmax(a, fabs(x - y));
    } else if (METRIC == ADOBO_METRIC_HAMMING) {
      a += (x != y) ? 1.0 : 0.0;
    } else if (METRIC == ADOBO_METRIC_CANBERRA) {
      const double den = fabs(x) + fabs(y);
      if (den > 0) a += fabs(x - y) / den;
    } else {  // Bray-Curtis as sklearn defines it: |x| + |y| in the denominator (scipy has |x + y|)
      a += fabs(x - y);
      b += fabs(x) + fabs(y);
    }
  }
  __device__ __forceinline__ double value() const {
    if (METRIC == ADOBO_METRIC_BRAYCURTIS) return b > 0 ? a / b : 0.0;
    return a;  // squared (Euclidean) / un-normalised (Hamming): the order is the metric's
  }
};

template <int METRIC>
__global__ void __launch_bounds__(256) knn_metric_tile(const double* __restrict__ x, i64 n, int p, i64 q_begin,
                                                       i64 n_local, int k, int* __restrict__ nn0) {
  constexpr int E = 2;
  extern __shared__ __align__(16) double km_dyn[];  // xq[MQ][p] | rt[p][MRS]
  double* xq = km_dyn;
  double* rt = km_dyn + (size_t)MQ * p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const i64 ql0 = (i64)blockIdx.x * MQ;
  for (int e = threadIdx.x; e < MQ * p; e += 256) {
    const i64 ql = ql0 + e / p;
    xq[e] = ql < n_local ? x[(q_begin + ql) * p + (e % p)] : 0.0;
  }
  double key[MQW][E];
  int idx[MQW][E];
#pragma unroll
  for (int q = 0; q < MQW; ++q)
#pragma unroll
    for (int e = 0; e < E; ++e) key[q][e] = INFINITY, idx[q][e] = INT_MAX;
  const double* myq = xq + (size_t)warp * MQW * p;
  for (i64 base = 0; base < n; base += MRT) {
    __syncthreads();
    for (int e = threadIdx.x; e < MRT * p; e += 256) {  // coalesced along the coordinates, transposed on the way in
      const int i = e / p, t = e - i * p;
      const i64 r = base + i;
      rt[(size_t)t * MRS + i] = r < n ? x[r * p + t] : 0.0;
    }
    __syncthreads();
    MetricAcc<METRIC> acc[MQW][2];
    for (int t = 0; t < p; ++t) {
      const double ra = rt[(size_t)t * MRS + lane], rb = rt[(size_t)t * MRS + lane + 32];
#pragma unroll
      for (int q = 0; q < MQW; ++q) {
        const double qv = myq[(size_t)q * p + t];
        acc[q][0].add(qv, ra);
        acc[q][1].add(qv, rb);
      }
    }
#pragma unroll
    for (int q = 0; q < MQW; ++q) {
      const i64 qg = q_begin + ql0 + warp * MQW + q;
#pragma unroll
      for (int hq = 0; hq < 2; ++hq) {
        const i64 r = base + lane + 32 * hq;
        double dd = r < n ? acc[q][hq].value() : INFINITY;
        int ri = r < n ? (int)r : INT_MAX;
        if (r == qg) ri = -1;  // the query itself wins a zero-distance tie
        const double wk = __shfl_sync(0xffffffffu, key[q][E - 1], 31);
        const int wi = __shfl_sync(0xffffffffu, idx[q][E - 1], 31);
        unsigned m = __ballot_sync(0xffffffffu, cand_less(dd, ri, wk, wi));
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const double nk = __shfl_sync(0xffffffffu, dd, src);
          const int ni = __shfl_sync(0xffffffffu, ri, src);
          list_insert<double, E>(key[q], idx[q], nk, ni, lane);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < MQW; ++q) {
    const i64 ql = ql0 + warp * MQW + q;
    if (ql >= n_local) continue;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int rank = lane * E + e;
      if (rank < k) nn0[(size_t)ql * k + rank] = (idx[q][e] == -1) ? (int)(q_begin + ql) : idx[q][e];
    }
  }
}

template <int METRIC>
static void launch_knn_metric(adobo_ctx* c, const double* x, i64 n, int p, i64 q_begin, i64 n_local, int k) {
  const size_t smem = sizeof(double) * ((size_t)MQ * p + (size_t)p * MRS);
  CUDA_CHECK(cudaFuncSetAttribute(knn_metric_tile<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WORK(c, (double)n_local * (double)n * p);
  LAUNCH(c, "knn_metric_tile", (knn_metric_tile<METRIC><<<(unsigned)((n_local + MQ - 1) / MQ), 256, smem, c->stream>>>(
                                   x, n, p, q_begin, n_local, k, c->nn.p)));
}

bool knn_tc_supported(int p);  // knn_tc.cu
void run_knn_shortlist_tc(adobo_ctx* c, const double* x, i64 n, int p, const double* ctr, i64 q_begin, i64 q_end, int E,
                          float* skey, int* sidx, DevBuf<float>& norms, unsigned* maxnorm);

// comp_local: this rank's rows (n_local x p, device, fp64).  Leaves c->nn (n_local x k, 0-based global ids).
void run_knn(adobo_ctx* c, const double* comp_local, i64 n_local, int p, int k, int metric) {
  REQUIRE(metric >= ADOBO_METRIC_EUCLIDEAN && metric <= ADOBO_METRIC_EUCLIDEAN_FP64, ADOBO_ERR_ARG, "knn: unknown metric %d",
          metric);
  REQUIRE(p >= 1 && p <= MAXP, ADOBO_ERR_LIMIT, "knn: %d dimensions (supported: 1..%d)", p, MAXP);
  REQUIRE(k >= 1 && k <= 48, ADOBO_ERR_LIMIT, "knn: k=%d (supported: 1..48)", k);
  cudaStream_t st = c->stream;
  // ---- all-gather the embedding (every rank searches its queries against all cells) ----
  std::vector<i64> all_n;
  allgather_i64_host(c, n_local, all_n);
  i64 n = 0, q_begin = 0;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) q_begin = n;
    n += all_n[r];
  }
  REQUIRE(n < INT_MAX - 256, ADOBO_ERR_LIMIT, "knn: %lld points exceed the int32 id range", (long long)n);
  REQUIRE(k <= n, ADOBO_ERR_ARG, "Expected n_neighbors <= n_samples, but n_samples = %lld, n_neighbors = %d",
          (long long)n, k);
  DevBuf<double> gathered;
  const double* x = comp_local;
  if (c->world > 1) {
    gathered.ensure((size_t)n * p);
    std::vector<size_t> bytes(c->world);
    for (int r = 0; r < c->world; ++r) bytes[r] = (size_t)all_n[r] * p * sizeof(double);
    allgatherv_bytes(c, comp_local, gathered.p, bytes);
    x = gathered.p;
  }
  if (metric != ADOBO_METRIC_EUCLIDEAN) {
    c->nn.ensure((size_t)std::max<i64>(n_local, 1) * k);
    c->k = k;
    c->knn_n = n_local;
    if (n_local > 0) {
      switch (metric) {
        case ADOBO_METRIC_MANHATTAN: launch_knn_metric<ADOBO_METRIC_MANHATTAN>(c, x, n, p, q_begin, n_local, k); break;
        case ADOBO_METRIC_CHEBYSHEV: launch_knn_metric<ADOBO_METRIC_CHEBYSHEV>(c, x, n, p, q_begin, n_local, k); break;
        case ADOBO_METRIC_HAMMING: launch_knn_metric<ADOBO_METRIC_HAMMING>(c, x, n, p, q_begin, n_local, k); break;
        case ADOBO_METRIC_CANBERRA: launch_knn_metric<ADOBO_METRIC_CANBERRA>(c, x, n, p, q_begin, n_local, k); break;
        case ADOBO_METRIC_BRAYCURTIS: launch_knn_metric<ADOBO_METRIC_BRAYCURTIS>(c, x, n, p, q_begin, n_local, k); break;
        default: launch_knn_metric<ADOBO_METRIC_EUCLIDEAN>(c, x, n, p, q_begin, n_local, k); break;  // fp64 brute force
      }
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    c->have_knn = true;
    return;
  }
  const int P = (p + 3) & ~3;
  const i64 npad = (n + RT - 1) / RT * RT;
  const int E = (k <= 16) ? 1 : 2;
  const int KP = 32 * E;
  DevBuf<double> ctr;
  DevBuf<float> refT, norms, skey;
  DevBuf<int> sidx, flagged, nflag;
  DevBuf<unsigned> maxnorm;
  // The shortlist comes from the tcgen05 kernel (knn_tc.cu) whenever its resident A' operand fits in
  // shared memory (p <= 126); wider embeddings use the CUDA-core tile kernel above.
  const bool use_tc = knn_tc_supported(p) && getenv("ADOBO_B200_KNN_SIMT") == nullptr;
  ctr.ensure(MAXP);
  maxnorm.ensure(1);
  CUDA_CHECK(cudaMemsetAsync(maxnorm.p, 0, sizeof(unsigned), st));
  LAUNCH(c, "knn_center", (knn_center<<<1, 1024, 0, st>>>(x, std::min<i64>(n, 1024), p, ctr.p)));
  skey.ensure((size_t)std::max<i64>(n_local, 1) * KP);
  sidx.ensure((size_t)std::max<i64>(n_local, 1) * KP);
  const i64 q_end = q_begin + n_local;
  c->nn.ensure((size_t)std::max<i64>(n_local, 1) * k);
  c->k = k;
  c->knn_n = n_local;
  if (n_local > 0) {
    if (use_tc) {
      run_knn_shortlist_tc(c, x, n, p, ctr.p, q_begin, q_end, E, skey.p, sidx.p, norms, maxnorm.p);
    } else {
      refT.ensure((size_t)npad * P);
      norms.ensure((size_t)npad);
      LAUNCH(c, "knn_prep", (knn_prep<<<(unsigned)((npad + 127) / 128), 128, 0, st>>>(x, n, npad, p, P, ctr.p, refT.p,
                                                                                   norms.p, maxnorm.p)));
      const i64 qt0 = q_begin / QT, qt1 = (q_end + QT - 1) / QT;
      const size_t smem = sizeof(float) * ((size_t)P * RT + (size_t)P * QT + RT + QT) + sizeof(int) * QT +
                          sizeof(Cand) * ((size_t)QT * RT + (size_t)QT * KP);
      WORK(c, 2.0 * (double)(qt1 - qt0) * QT * (double)npad * P);  // flops
      if (E == 1) {
        CUDA_CHECK(cudaFuncSetAttribute(knn_shortlist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(c, "knn_shortlist", (knn_shortlist<1><<<(unsigned)(qt1 - qt0), 256, smem, st>>>(
                                       refT.p, norms.p, npad, P, qt0, skey.p, sidx.p, q_begin, q_end)));
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(knn_shortlist<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(c, "knn_shortlist", (knn_shortlist<2><<<(unsigned)(qt1 - qt0), 256, smem, st>>>(
                                       refT.p, norms.p, npad, P, qt0, skey.p, sidx.p, q_begin, q_end)));
      }
    }
    flagged.ensure((size_t)n_local);
    nflag.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(nflag.p, 0, sizeof(int), st));
    const unsigned rb = (unsigned)std::min<i64>((n_local + 7) / 8, (i64)c->sm_count * 8);
    if (E == 1)
      LAUNCH(c, "knn_rerank", (knn_rerank<1><<<rb, 256, 0, st>>>(x, p, skey.p, sidx.p, norms.p, maxnorm.p, q_begin,
                                                               n_local, n, k, nullptr, c->nn.p, flagged.p, nflag.p)));
    else
      LAUNCH(c, "knn_rerank", (knn_rerank<2><<<rb, 256, 0, st>>>(x, p, skey.p, sidx.p, norms.p, maxnorm.p, q_begin,
                                                               n_local, n, k, nullptr, c->nn.p, flagged.p, nflag.p)));
    int nf = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nf, nflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (getenv("ADOBO_B200_DEBUG")) fprintf(stderr, "[adobo_b200] knn: %d of %lld queries failed the fp32 certificate\n", nf, (long long)n_local);
    if (nf > 0)
      LAUNCH(c, "knn_exact_rows",
             (knn_exact_rows<<<(unsigned)nf, 256, 0, st>>>(x, n, p, flagged.p, q_begin, k, nullptr, c->nn.p)));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  c->have_knn = true;
}

}  // namespace adobo
