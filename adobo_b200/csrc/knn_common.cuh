// Shared by the kNN kernels: candidate ordering and the warp-held sorted shortlist.
#pragma once
#include "common.cuh"

namespace adobo {

struct Cand {
  float key;
  int idx;
};
__device__ __forceinline__ bool cand_less(float ka, int ia, float kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ bool cand_less(double ka, int ia, double kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// Sorted list of 32*E (key, idx) held E per lane (lane l owns global slots l*E .. l*E+E-1).
template <typename KT, int E>
__device__ __forceinline__ void list_insert(KT (&key)[E], int (&idx)[E], KT nk, int ni, int lane) {
  // skip when not better than the current worst
  const KT wk = __shfl_sync(0xffffffffu, key[E - 1], 31);
  const int wi = __shfl_sync(0xffffffffu, idx[E - 1], 31);
  if (!cand_less(nk, ni, wk, wi)) return;
  int less = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) less += cand_less(key[e], idx[e], nk, ni) ? 1 : 0;
  const int pos = __reduce_add_sync(0xffffffffu, less);
  KT pk = __shfl_up_sync(0xffffffffu, key[E - 1], 1);
  int pi = __shfl_up_sync(0xffffffffu, idx[E - 1], 1);
  KT ok[E];
  int oi[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    ok[e] = key[e];
    oi[e] = idx[e];
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int gi = lane * E + e;
    const KT prevk = (e == 0) ? pk : ok[e - 1];
    const int previ = (e == 0) ? pi : oi[e - 1];
    if (gi == pos) {
      key[e] = nk;
      idx[e] = ni;
    } else if (gi > pos) {
      key[e] = prevk;
      idx[e] = previ;
    }
  }
}


}  // namespace adobo
