// K9  shared-nearest-neighbour graph with Jaccard pruning -- pure integer work.
//
// Replaces adobo/clustering.py:86-119: the reference builds the 0/1 matrix M (ones at
// (nn_idx[i,0]-1, nn_idx[i,c]-1) for c >= 1 plus the diagonal, duplicates summed by coo->csr),
// forms V = M M^T with a scipy sparse product and walks the non-zeros in a Python loop keeping
// (i, j) when v / (k + (k - v)) > prune_snn.  Here:
//   snn_histogram / snn_fill   forward lists F(a) = columns of row a of M and reverse lists
//                              R(g) = rows of column g, as multisets (general in nn_idx[i,0]);
//   snn_rows                   one warp per row a: gathers R(g) for every g in F(a) into shared
//                              memory, bitonic-sorts the ids, run lengths are the integer
//                              shared-neighbour counts V[a, b].  Three tiers by candidate count (k^2 on
//                              average, heavy-tailed: hub cells sit in many neighbour lists): a warp with a
//                              2048-entry buffer, a CTA with 8192 entries (32 KB: several CTAs per SM --
//                              at k = 30 a third of the rows land here), a CTA with 32768 entries (128 KB).
// The compact edge list sorted by (source, target) needs every row's number of kept edges before any edge can be
// placed.  One pass over the rows does the expensive part (gather + sort) once: it counts the row's kept edges,
// reserves that many slots of a staging buffer with one atomic and writes (target, V) there; a copy kernel then moves
// every row's segment to its final place (snn_emit).  The staging buffer is sized for 64 kept edges per row on
// average (a BASELINE-shaped graph keeps ~11); if a pruning level keeps more, the second pass of the two-pass
// scheme (MODE 1) recomputes the rows instead.
#include "common.cuh"

namespace adobo {

__global__ void snn_histogram(const int* __restrict__ nn, i64 n, int k, i64* __restrict__ fcnt, i64* __restrict__ rcnt,
                              int* __restrict__ bad) {
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n * k; e += (i64)gridDim.x * blockDim.x) {
    const i64 i = e / k;
    const int c = (int)(e - i * k);
    const int row = (c == 0) ? (int)i : nn[i * k];
    const int col = (c == 0) ? (int)i : nn[e];
    if (row < 0 || row >= n || col < 0 || col >= n) {
      *bad = 1;
      continue;
    }
    atomicAdd((unsigned long long*)(fcnt + row), 1ull);
    atomicAdd((unsigned long long*)(rcnt + col), 1ull);
  }
}

__global__ void snn_fill(const int* __restrict__ nn, i64 n, int k, const i64* __restrict__ fptr,
                         const i64* __restrict__ rptr, unsigned* __restrict__ fcur, unsigned* __restrict__ rcur,
                         int* __restrict__ fcol, int* __restrict__ rrow) {
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n * k; e += (i64)gridDim.x * blockDim.x) {
    const i64 i = e / k;
    const int c = (int)(e - i * k);
    const int row = (c == 0) ? (int)i : nn[i * k];
    const int col = (c == 0) ? (int)i : nn[e];
    fcol[fptr[row] + atomicAdd(fcur + row, 1u)] = col;
    rrow[rptr[col] + atomicAdd(rcur + col, 1u)] = row;
  }
}

template <int T>
__device__ __forceinline__ void group_sync() {
  if (T == 32)
    __syncwarp();
  else
    __syncthreads();
}

template <int T>
__device__ __forceinline__ void bitonic_sort(unsigned* buf, int N, int tid) {
  for (int sz = 2; sz <= N; sz <<= 1)
    for (int stp = sz >> 1; stp > 0; stp >>= 1) {
      for (int i = tid; i < (N >> 1); i += T) {
        const int lo = ((i & ~(stp - 1)) << 1) | (i & (stp - 1));
        const int hi = lo + stp;
        const bool asc = (lo & sz) == 0;
        const unsigned a = buf[lo], b = buf[hi];
        if ((a > b) == asc) {
          buf[lo] = b;
          buf[hi] = a;
        }
      }
      group_sync<T>();
    }
}

// T threads cooperate on one row.  rows: [row_begin, row_end) when list == nullptr, else list[0..nlist).
// MODE 0: rowkept / rowmiss / totals (+ big-row list in warp mode); MODE 1: edges at their final place (eptr);
// MODE 2: MODE 0 and the row's kept (target, V) pairs into the staging buffer (SnnStage).
struct SnnStage {
  unsigned long long* cursor = nullptr;  // next free slot
  i64 cap = 0;
  int* dst = nullptr;
  int* v = nullptr;
  i64* rowoff = nullptr;  // per local row: where its segment starts
  int* overflow = nullptr;
};
template <int T, int CAP, int MODE>
__global__ void __launch_bounds__(T == 32 ? 128 : 256) snn_rows(
    const i64* __restrict__ fptr, const int* __restrict__ fcol, const i64* __restrict__ rptr,
    const int* __restrict__ rrow, i64 row_begin, i64 row_end, const int* __restrict__ list, int nlist, int k,
    double prune, i64* __restrict__ rowkept, i64* __restrict__ rowmiss, unsigned long long* __restrict__ totals,
    int* __restrict__ biglist, int* __restrict__ nbig, const i64* __restrict__ eptr, const i64* __restrict__ mptr,
    i64 e_regular, int* __restrict__ e_src, int* __restrict__ e_dst, int* __restrict__ e_v, int* __restrict__ err,
    const SnnStage sg) {
  constexpr bool FILL = MODE == 1;
  extern __shared__ unsigned sbuf[];
  constexpr int GROUPS = (T == 32) ? 4 : 1;
  const int grp = (T == 32) ? (threadIdx.x >> 5) : 0;
  const int tid = (T == 32) ? (threadIdx.x & 31) : threadIdx.x;
  const int lane = threadIdx.x & 31;
  unsigned* buf = sbuf + (size_t)grp * CAP;
  const i64 nwork = list ? (i64)nlist : (row_end - row_begin);
  for (i64 wi = (i64)blockIdx.x * GROUPS + grp; wi < nwork; wi += (i64)gridDim.x * GROUPS) {
    const i64 a = list ? (i64)list[wi] : (row_begin + wi);
    const i64 la = a - row_begin;  // local row number (output arrays are per local row)
    const i64 f0 = fptr[a], f1 = fptr[a + 1];
    // ---- candidate total ----
    i64 total = 0;
    for (i64 t = f0; t < f1; ++t) {
      const int g = fcol[t];
      total += rptr[g + 1] - rptr[g];
    }
    if (total > CAP) {  // too long for this tier: hand the row to the next one (count pass), or give up
      if (biglist) {
        if (!FILL && tid == 0) biglist[atomicAdd(nbig, 1)] = (int)a;
      } else if (!FILL && tid == 0) {
        *err = 1;
      }
      continue;  // uniform for the whole group
    }
    // ---- gather reverse lists ----
    int off = 0;
    for (i64 t = f0; t < f1; ++t) {
      const int g = fcol[t];
      const i64 rb = rptr[g];
      const int len = (int)(rptr[g + 1] - rb);
      for (int s = tid; s < len; s += T) buf[off + s] = (unsigned)rrow[rb + s];
      off += len;
    }
    int N = 2;
    while (N < (int)total) N <<= 1;
    for (int s = (int)total + tid; s < N; s += T) buf[s] = 0xffffffffu;
    group_sync<T>();
    bitonic_sort<T>(buf, N, tid);
    // ---- run lengths = shared-neighbour counts; the group's first warp walks the sorted ids ----
    if (T == 32 || threadIdx.x < 32) {
      i64 kept = 0, uniq = 0;
      const i64 ebase = FILL ? eptr[la] : 0;
      // one walk over the sorted ids; stage_base >= 0: write the kept pairs to the staging segment that starts there
      auto walk = [&](i64 stage_base) {
        kept = 0, uniq = 0;
        for (int base = 0; base < (int)total; base += 32) {  // warp-uniform trip count
          const int idx = base + lane;
          bool head = false, keep = false;
          unsigned b = 0;
          int v = 0;
          if (idx < (int)total) {
            b = buf[idx];
            head = (idx == 0) || (buf[idx - 1] != b);
            if (head) {
              v = 1;
              while (idx + v < (int)total && buf[idx + v] == b) ++v;
              const double strength = (double)v / (double)(k + (k - v));  // clustering.py:108
              keep = strength > prune;                                     // clustering.py:109
            }
          }
          const unsigned hm = __ballot_sync(0xffffffffu, head);
          const unsigned km = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            const i64 rank = kept + __popc(km & ((1u << lane) - 1u));
            if (FILL) {
              e_src[ebase + rank] = (int)a;
              e_dst[ebase + rank] = (int)b;
              e_v[ebase + rank] = v;  // V_ab = (M M^T)_ab, clustering.py:99
            } else if (MODE == 2 && stage_base >= 0) {
              sg.dst[stage_base + rank] = (int)b;
              sg.v[stage_base + rank] = v;
            }
          }
          uniq += __popc(hm);
          kept += __popc(km);
        }
      };
      walk(-1);
      if (MODE == 2) {
        i64 sb = -1;
        if (lane == 0 && kept > 0) {
          sb = (i64)atomicAdd(sg.cursor, (unsigned long long)kept);
          if (sb + kept > sg.cap) {
            atomicExch(sg.overflow, 1);
            sb = -1;
          }
        }
        sb = __shfl_sync(0xffffffffu, sb, 0);
        if (lane == 0) sg.rowoff[la] = sb;
        if (sb >= 0) walk(sb);
      }
      if (lane == 0) {
        if (!FILL) {
          rowkept[la] = kept;
          rowmiss[la] = (kept == 0) ? 1 : 0;
          atomicAdd(totals, (unsigned long long)uniq);
        } else if (kept == 0) {  // node without any kept edge -> self loop, clustering.py:114-118
          e_src[e_regular + mptr[la]] = (int)a;
          e_dst[e_regular + mptr[la]] = (int)a;
          e_v[e_regular + mptr[la]] = 0;  // not an entry of V that passed the threshold
        }
      }
    }
    group_sync<T>();
  }
}

// every row's staged segment to its final place (rows in order, targets ascending inside a row: the walk's order);
// a row without a kept edge gets its self loop (clustering.py:114-118).  One warp per row.
__global__ void __launch_bounds__(256) snn_emit(const SnnStage sg, const i64* __restrict__ rowkept,
                                                const i64* __restrict__ eptr, const i64* __restrict__ mptr, i64 row_begin,
                                                i64 nloc, i64 e_regular, int* __restrict__ e_src, int* __restrict__ e_dst,
                                                int* __restrict__ e_v) {
  const int lane = threadIdx.x & 31;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 la = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; la < nloc; la += nw) {
    const i64 kept = rowkept[la];
    const int a = (int)(row_begin + la);
    if (kept == 0) {
      if (lane == 0) {
        const i64 pos = e_regular + mptr[la];
        e_src[pos] = a;
        e_dst[pos] = a;
        e_v[pos] = 0;
      }
      continue;
    }
    const i64 from = sg.rowoff[la], to = eptr[la];
    for (i64 e = lane; e < kept; e += 32) {
      e_src[to + e] = a;
      e_dst[to + e] = sg.dst[from + e];
      e_v[to + e] = sg.v[from + e];
    }
  }
}

// nn_all: (n_all x k) 0-based ids on the device; this rank emits rows [row_begin, row_end).
void run_snn(adobo_ctx* c, const int* nn, i64 n, i64 row_begin, i64 row_end, int k, double prune, i64* n_edges,
             i64* pruned, i64* total_links) {
  REQUIRE(k >= 1, ADOBO_ERR_ARG, "snn: k must be >= 1");
  REQUIRE(n < INT_MAX, ADOBO_ERR_LIMIT, "snn: too many cells");
  cudaStream_t st = c->stream;
  const i64 nloc = row_end - row_begin;
  DevBuf<i64> fcnt, rcnt, fptr, rptr, rowkept, rowmiss, eptr, mptr;
  DevBuf<unsigned> fcur, rcur;
  DevBuf<int> fcol, rrow, flags, biglist;
  DevBuf<unsigned long long> totals;
  fcnt.ensure((size_t)n + 2);
  rcnt.ensure((size_t)n + 2);
  fptr.ensure((size_t)n + 2);
  rptr.ensure((size_t)n + 2);
  fcur.ensure((size_t)n + 1);
  rcur.ensure((size_t)n + 1);
  fcol.ensure((size_t)n * k);
  rrow.ensure((size_t)n * k);
  flags.ensure(4);  // [0] bad input  [1] nbig  [2] cta-capacity overflow
  totals.ensure(2);
  rowkept.ensure((size_t)nloc + 2);
  rowmiss.ensure((size_t)nloc + 2);
  eptr.ensure((size_t)nloc + 2);
  mptr.ensure((size_t)nloc + 2);
  biglist.ensure((size_t)std::max<i64>(nloc, 1));
  CUDA_CHECK(cudaMemsetAsync(fcnt.p, 0, sizeof(i64) * (n + 2), st));
  CUDA_CHECK(cudaMemsetAsync(rcnt.p, 0, sizeof(i64) * (n + 2), st));
  CUDA_CHECK(cudaMemsetAsync(fcur.p, 0, sizeof(unsigned) * (n + 1), st));
  CUDA_CHECK(cudaMemsetAsync(rcur.p, 0, sizeof(unsigned) * (n + 1), st));
  CUDA_CHECK(cudaMemsetAsync(flags.p, 0, sizeof(int) * 4, st));
  CUDA_CHECK(cudaMemsetAsync(totals.p, 0, sizeof(unsigned long long) * 2, st));
  CUDA_CHECK(cudaMemsetAsync(rowkept.p, 0, sizeof(i64) * (nloc + 2), st));
  CUDA_CHECK(cudaMemsetAsync(rowmiss.p, 0, sizeof(i64) * (nloc + 2), st));
  const unsigned gb = (unsigned)std::max<i64>(1, std::min<i64>((n * k + 255) / 256, (i64)c->sm_count * 16));
  LAUNCH(c, "snn_histogram", (snn_histogram<<<gb, 256, 0, st>>>(nn, n, k, fcnt.p, rcnt.p, flags.p)));
  exclusive_scan_i64(c, fcnt.p, fptr.p, n);
  exclusive_scan_i64(c, rcnt.p, rptr.p, n);
  int h_flags[4];
  CUDA_CHECK(cudaMemcpyAsync(h_flags, flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  REQUIRE(h_flags[0] == 0, ADOBO_ERR_ARG, "snn: nn_idx holds an index outside 1..n");
  LAUNCH(c, "snn_fill", (snn_fill<<<gb, 256, 0, st>>>(nn, n, k, fptr.p, rptr.p, fcur.p, rcur.p, fcol.p, rrow.p)));

  // staging for the single pass (see the header): 64 kept edges per row on average + slack
  DevBuf<int> stage_dst, stage_v;
  DevBuf<i64> rowoff;
  DevBuf<unsigned long long> cursor;
  SnnStage sg;
  sg.cap = 64 * std::max<i64>(nloc, 1) + (1 << 16);
  stage_dst.ensure((size_t)sg.cap);
  stage_v.ensure((size_t)sg.cap);
  rowoff.ensure((size_t)nloc + 2);
  cursor.ensure(2);
  CUDA_CHECK(cudaMemsetAsync(cursor.p, 0, sizeof(unsigned long long) * 2, st));
  sg.cursor = cursor.p, sg.dst = stage_dst.p, sg.v = stage_v.p, sg.rowoff = rowoff.p;
  sg.overflow = (int*)(cursor.p + 1);
  constexpr int WCAP = 2048, MCAP = 8192, CCAP = 32768;
  const size_t wsmem = sizeof(unsigned) * WCAP * 4, msmem = sizeof(unsigned) * MCAP, csmem = sizeof(unsigned) * CCAP;
  CUDA_CHECK(cudaFuncSetAttribute(snn_rows<256, CCAP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
  CUDA_CHECK(cudaFuncSetAttribute(snn_rows<256, CCAP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
  const unsigned wb = (unsigned)std::max<i64>(1, std::min<i64>((nloc + 3) / 4, (i64)c->sm_count * 16));
  DevBuf<int> hugelist;
  hugelist.ensure((size_t)std::max<i64>(nloc, 1));
  int nbig = 0, nhuge = 0;  // flags: [1] rows for the middle tier, [3] rows for the top tier
  if (nloc > 0) {
    LAUNCH(c, "snn_rows", (snn_rows<32, WCAP, 2><<<wb, 128, wsmem, st>>>(
                              fptr.p, fcol.p, rptr.p, rrow.p, row_begin, row_end, nullptr, 0, k, prune, rowkept.p,
                              rowmiss.p, totals.p, biglist.p, flags.p + 1, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                              flags.p + 2, sg)));
    CUDA_CHECK(cudaMemcpyAsync(h_flags, flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    nbig = h_flags[1];
    if (nbig > 0) {
      LAUNCH(c, "snn_rows_mid", (snn_rows<256, MCAP, 2><<<(unsigned)nbig, 256, msmem, st>>>(
                                    fptr.p, fcol.p, rptr.p, rrow.p, row_begin, row_end, biglist.p, nbig, k, prune,
                                    rowkept.p, rowmiss.p, totals.p, hugelist.p, flags.p + 3, nullptr, nullptr, 0, nullptr,
                                    nullptr, nullptr, flags.p + 2, sg)));
      CUDA_CHECK(cudaMemcpyAsync(h_flags, flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      nhuge = h_flags[3];
    }
    if (nhuge > 0) {
      LAUNCH(c, "snn_rows_big", (snn_rows<256, CCAP, 2><<<(unsigned)nhuge, 256, csmem, st>>>(
                                    fptr.p, fcol.p, rptr.p, rrow.p, row_begin, row_end, hugelist.p, nhuge, k, prune,
                                    rowkept.p, rowmiss.p, totals.p, nullptr, nullptr, nullptr, nullptr, 0, nullptr,
                                    nullptr, nullptr, flags.p + 2, sg)));
      CUDA_CHECK(cudaMemcpyAsync(h_flags, flags.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      REQUIRE(h_flags[2] == 0, ADOBO_ERR_LIMIT, "snn: a cell has more than %d shared-neighbour candidates", CCAP);
    }
  }
  exclusive_scan_i64(c, rowkept.p, eptr.p, nloc);
  exclusive_scan_i64(c, rowmiss.p, mptr.p, nloc);
  i64 e_regular = 0, n_missing = 0;
  unsigned long long h_tot[2], h_cur[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(h_cur, cursor.p, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(&e_regular, eptr.p + nloc, sizeof(i64), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(&n_missing, mptr.p + nloc, sizeof(i64), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaMemcpyAsync(h_tot, totals.p, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  const i64 E = e_regular + n_missing;
  c->e_src.ensure((size_t)std::max<i64>(E, 1));
  c->e_dst.ensure((size_t)std::max<i64>(E, 1));
  c->e_v.ensure((size_t)std::max<i64>(E, 1));
  const bool staged = ((int)(h_cur[1] & 0xffffffffull) == 0) && getenv("ADOBO_B200_SNN_TWOPASS") == nullptr;
  if (nloc > 0 && staged) {
    const unsigned eb = (unsigned)std::max<i64>(1, std::min<i64>((nloc + 7) / 8, (i64)c->sm_count * 16));
    LAUNCH(c, "snn_emit", (snn_emit<<<eb, 256, 0, st>>>(sg, rowkept.p, eptr.p, mptr.p, row_begin, nloc, e_regular, c->e_src.p,
                                                        c->e_dst.p, c->e_v.p)));
  } else if (nloc > 0) {  // more kept edges than the staging buffer holds: second pass over the rows
    LAUNCH(c, "snn_rows", (snn_rows<32, WCAP, 1><<<wb, 128, wsmem, st>>>(
                              fptr.p, fcol.p, rptr.p, rrow.p, row_begin, row_end, nullptr, 0, k, prune, nullptr, nullptr,
                              nullptr, nullptr, nullptr, eptr.p, mptr.p, e_regular, c->e_src.p, c->e_dst.p, c->e_v.p,
                              flags.p + 2, sg)));
    if (nbig > 0)  // rows that overflow this tier are skipped here (uniformly) and filled by the next launch
      LAUNCH(c, "snn_rows_mid", (snn_rows<256, MCAP, 1><<<(unsigned)nbig, 256, msmem, st>>>(
                                    fptr.p, fcol.p, rptr.p, rrow.p, row_begin, row_end, biglist.p, nbig, k, prune,
                                    nullptr, nullptr, nullptr, hugelist.p, nullptr, eptr.p, mptr.p, e_regular, c->e_src.p,
                                    c->e_dst.p, c->e_v.p, flags.p + 2, sg)));
    if (nhuge > 0)
      LAUNCH(c, "snn_rows_big", (snn_rows<256, CCAP, 1><<<(unsigned)nhuge, 256, csmem, st>>>(
                                    fptr.p, fcol.p, rptr.p, rrow.p, row_begin, row_end, hugelist.p, nhuge, k, prune,
                                    nullptr, nullptr, nullptr, nullptr, nullptr, eptr.p, mptr.p, e_regular, c->e_src.p,
                                    c->e_dst.p, c->e_v.p, flags.p + 2, sg)));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  c->n_edges = E;
  *n_edges = E;
  *total_links = (i64)h_tot[0];
  *pruned = (i64)h_tot[0] - e_regular;
}

}  // namespace adobo
