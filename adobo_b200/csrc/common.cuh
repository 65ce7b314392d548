// adobo_b200 -- shared declarations of the sm_100a CUDA library (see include/adobo_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/adobo_b200.h"

typedef long long i64;

namespace adobo {

// ---- errors -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
struct Fail {
  int code;
};

#define CUDA_CHECK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      adobo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      throw adobo::Fail{ADOBO_ERR_CUDA};                                                      \
    }                                                                                         \
  } while (0)

#define REQUIRE(cond, code, ...)    \
  do {                              \
    if (!(cond)) {                  \
      adobo::set_error(__VA_ARGS__); \
      throw adobo::Fail{code};      \
    }                               \
  } while (0)

// ---- device memory ------------------------------------------------------------------------
// Caching allocator (api.cu): blocks released by DevBuf go back to a per-device free list instead
// of cudaFree, so a steady-state pass of the hot path performs no cudaMalloc / cudaFree at all
// (both serialise with the device and cost anywhere from microseconds to milliseconds).
void* pool_alloc(size_t bytes);
void pool_free(void* p);
void pool_trim();  // cudaFree every cached block of the current device

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  void ensure(size_t n) {
    if (n <= cap && p) return;
    release();
    size_t want = n ? n : 1;
    p = (T*)pool_alloc(want * sizeof(T));
    cap = want;
  }
  void release() {
    if (p) pool_free(p);
    p = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

struct ProfItem {
  double ms = 0;
  i64 launches = 0;
  double work = 0;  // algorithmic bytes (or flops for the distance kernel) of the launches
};

struct PendingEvent {
  const char* name;
  cudaEvent_t a, b;
  double work;
};

// Matrix (sparse, cells x columns) in the two layouts the Lanczos mat-vecs read.
template <typename VT>
struct SparsePair {
  i64 rows = 0;
  int cols = 0;
  i64 nnz = 0;
  DevBuf<i64> rptr;   // CSR row pointers (rows+1)
  DevBuf<int> cidx;   // CSR column ids
  DevBuf<VT> rval;    // CSR values
  DevBuf<i64> cptr;   // CSC column pointers (cols+1)
  DevBuf<int> ridx;   // CSC row ids (ascending inside a column)
  DevBuf<VT> cval;    // CSC values
};

struct Comm;  // NCCL wrapper, api.cu

// One-shot all-reduce over NVLink peer memory (api.cu sets it up with CUDA IPC when world > 1).
// Every rank owns a mailbox data[parity][rank][2 PEER_SLOT] + flags[parity][rank][channel]; an exchanging CTA writes
// its vector into ITS slot of EVERY peer's mailbox, releases a sequence flag, waits for the flags of all ranks in
// its own mailbox and sums the slots in rank order -- the result is bitwise identical on all ranks, and the exchange
// is part of the kernel that needs the sum (no separate collective launch).  Channels keep concurrent exchanges
// apart: channel 0 is the one-CTA exchange of the last-CTA reductions (first PEER_SLOT doubles of a rank slot),
// channels 1..PEER_CTAS belong to CTA c-1 of the V-side cluster kernel, which exchange their slices of the step's
// vector in parallel (PEER_CH_SLOT doubles each, in the second half of the rank slot).
constexpr int PEER_MAX = 8;
constexpr int PEER_SLOT = 4096;    // doubles per rank slot on channel 0: larger vectors go through NCCL
constexpr int PEER_CTAS = 8;       // cluster CTAs with a channel of their own
constexpr int PEER_CH = 1 + PEER_CTAS;
constexpr int PEER_CH_SLOT = PEER_SLOT / PEER_CTAS;  // 512 doubles per cluster-CTA channel
constexpr int PEER_LL_MAX = 256;   // doubles per low-latency message (peer_allreduce_ll)
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;  // a peer that is 20 s late is dead
struct PeerComm {
  int world = 1, rank = 0;
  double* data[PEER_MAX] = {};     // mailbox of every rank (peer pointers through CUDA IPC)
  unsigned* flags[PEER_MAX] = {};  // [2][world][PEER_CH] sequence flags, 128 bytes apart
  uint4* ll[PEER_MAX] = {};        // [2][world][PEER_CH][PEER_LL_MAX] flagged words of the low-latency exchange
  unsigned* seq = nullptr;         // [PEER_CH] this rank's sequence counters (device)
  int* err = nullptr;              // raised when a peer did not answer within PEER_TIMEOUT_NS (the kernel goes on
                                   // with garbage instead of hanging the GPU; the host turns it into an error)
};

}  // namespace adobo

struct adobo_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // second stream + fork/join events: the small SVD of an IRLB restart runs beside the last Lanczos step
  cudaStream_t side = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  // ---- count matrix shard (CSR cells x genes) ----
  i64 n = 0;        // local cells
  i64 n_total = 0;  // cells over all ranks
  i64 row_offset = 0;
  int g = 0;
  i64 nnz = 0;
  adobo::DevBuf<i64> indptr;
  adobo::DevBuf<int> indices;
  adobo::DevBuf<int> counts;
  adobo::DevBuf<float> vals;  // normalized values, fp32, same pattern
  bool have_counts = false, have_norm = false;
  // ---- HVG ----
  adobo::DevBuf<i64> gsum, gsq;     // per gene sum x 2^e1, sum x^2 2^e2, fixed point (all ranks after all-reduce)
  double val_bound = 0;             // max |normalized value| when known analytically (0: measure it)
  double zero_level = 0;            // value of the implicit zeros of the normalized matrix: vals hold y - zero_level
  bool have_moments = false;
  adobo::DevBuf<int> suboffs;       // gene moments: per row, where each column sub-range starts (moments.cu)
  bool have_suboffs = false;
  int suboffs_gw = 0;
  std::vector<double> h_mean, h_var;
  std::vector<int> hvg_order;  // reference output order
  // ---- PCA ----
  adobo::SparsePair<float> sub;  // HVG column subset
  adobo::DevBuf<int> sub_genes;  // ascending gene ids (h)
  std::vector<int> h_sub_genes;
  adobo::DevBuf<double> mu, rs;  // per-column mean and 1/sigma
  adobo::SparsePair<float> perm; // dr.jackstraw: the subset with some genes permuted across cells
  adobo::DevBuf<double> perm_u, perm_s, perm_v;  // its IRLB outputs (the real PCA stays in comp / contr / sv)
  bool sub_scaled = false;       // statistics of `sub` are those of scale=True
  int p = 0;                     // components of the last PCA
  adobo::DevBuf<double> comp;    // n x p row-major
  adobo::DevBuf<double> contr;   // h x p row-major
  adobo::DevBuf<double> sv;      // p
  bool have_pca = false;
  // ---- kNN / SNN ----
  int k = 0;
  i64 knn_n = 0;
  adobo::DevBuf<int> nn;  // knn_n x k, 0-based global ids
  bool have_knn = false;
  adobo::DevBuf<int> e_src, e_dst, e_v;  // kept edges and their shared-neighbour counts V_ij
  i64 n_edges = 0;
  // ---- plumbing ----
  adobo::Comm* comm = nullptr;
  int world = 1, rank = 0;
  adobo::PeerComm peer;         // world == 1 unless the peer-memory all-reduce is set up
  void* peer_mailbox = nullptr;
  cudaEvent_t events[64] = {};
  bool profiling = false;
  std::map<std::string, adobo::ProfItem> prof;
  std::vector<adobo::PendingEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  i64 launches = 0;
  double next_work = 0;  // algorithmic bytes/flops of the next LAUNCH (profiling only)
  adobo::DevBuf<char> flush_buf;
  adobo::DevBuf<char> scratch;  // general scratch (cub temp storage etc.)
  void* pinned = nullptr;       // 4 KB pinned host mailbox
  void* irlb_ws = nullptr;      // persistent Lanczos workspace + CUDA graphs (irlb.cu)
  void (*irlb_ws_free)(void*) = nullptr;
};

namespace adobo {

// ---- dr.jackstraw: keyed bijection of [0, n) (adobo_perm_row) ---------------------------------------------------
// Four rounds of (odd multiply + add, xor-shift) on the smallest power-of-two domain that holds n, cycle-walked back
// into [0, n): every step is invertible, so distinct cells go to distinct cells.  Keys come from splitmix64 of
// (seed, gene); the same code runs on the host (tests, adobo_perm_row) and on the device.
struct PermKey {
  unsigned long long mul[4], add[4];
};
#ifdef __CUDACC__
#define ADOBO_HD __host__ __device__ __forceinline__
#else
#define ADOBO_HD inline
#endif
ADOBO_HD unsigned long long perm_splitmix(unsigned long long& x) {
  unsigned long long z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
ADOBO_HD PermKey perm_key(unsigned long long seed, int col) {
  unsigned long long st = seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(col + 1));
  PermKey k;
  for (int r = 0; r < 4; ++r) {
    k.mul[r] = perm_splitmix(st) | 1ull;
    k.add[r] = perm_splitmix(st);
  }
  return k;
}
ADOBO_HD long long perm_apply(const PermKey& k, long long n, long long row) {
  int bits = 1;
  while ((1ll << bits) < n) ++bits;
  const unsigned long long mask = (1ull << bits) - 1ull;
  const int sh = bits > 1 ? bits / 2 : 1;
  unsigned long long x = (unsigned long long)row;
  do {
    for (int r = 0; r < 4; ++r) {
      x = (x * k.mul[r] + k.add[r]) & mask;
      x ^= x >> sh;
    }
  } while ((long long)x >= n);
  return (long long)x;
}

// ---- launch bookkeeping ---------------------------------------------------------------------
void prof_begin(adobo_ctx* c, const char* name);
void prof_end(adobo_ctx* c);

// WORK(ctx, x): algorithmic bytes (flops for knn_shortlist) of the next launch, for the roofline
#define WORK(ctx, x) ((ctx)->next_work = (double)(x))
#define LAUNCH(ctx, name, ...)                                                              \
  do {                                                                                      \
    if ((ctx)->profiling) adobo::prof_begin((ctx), name);                                   \
    __VA_ARGS__;                                                                            \
    (ctx)->launches++;                                                                      \
    if ((ctx)->profiling) adobo::prof_end((ctx));                                           \
    cudaError_t _le = cudaGetLastError();                                                   \
    if (_le != cudaSuccess) {                                                               \
      adobo::set_error("%s:%d: launch %s -> %s", __FILE__, __LINE__, name,                  \
                       cudaGetErrorString(_le));                                            \
      throw adobo::Fail{ADOBO_ERR_CUDA};                                                    \
    }                                                                                       \
  } while (0)

// ---- collectives (no-ops when world == 1) ---------------------------------------------------
void allreduce_sum_f64(adobo_ctx* c, double* dev, size_t count);
void allreduce_sum_i64(adobo_ctx* c, i64* dev, size_t count);
// gathers `count_local` elements of every rank into out (rank order); counts[] per rank in elements
void allgatherv_bytes(adobo_ctx* c, const void* src, void* dst, const std::vector<size_t>& bytes_per_rank);
void allgather_i64_host(adobo_ctx* c, i64 mine, std::vector<i64>& all);

// ---- stage entry points (one per .cu file) ---------------------------------------------------
void run_normalize(adobo_ctx* c, double scaling, int log_mode, double small_const, i64* out_lib,
                   double* out_vals);
void run_moments(adobo_ctx* c);
void run_seurat_host(adobo_ctx* c, int ngenes, int num_bins, std::vector<int>& order, std::vector<double>& z);
void run_gather_subset(adobo_ctx* c, const std::vector<int>& genes_sorted, bool scale);
void run_permute_subset(adobo_ctx* c, const int* perm_cols, int n_perm, uint64_t seed, bool need_csc);  // gather.cu

struct IrlbResult {
  int steps = 0, nmult = 0, svd_fallbacks = 0;
};
// generic driver on a SparsePair; mu/rs may be null (no centring / scaling)
template <typename VT>
IrlbResult run_irlb(adobo_ctx* c, SparsePair<VT>& A, const double* mu_dev, const double* rs_dev, int nval,
                    double tol, int maxit, const double* v0_host, bool var_weigh, DevBuf<double>& outU,
                    DevBuf<double>& outS, DevBuf<double>& outV);
template <typename VT>
void build_csc(adobo_ctx* c, SparsePair<VT>& A);

// B[m-1][m-1] = ||y|| of an outer iteration's last W column for svd_arrow, which then does not wait for w_finish_t:
// from w_step_t's per-CTA partials (one GPU; column col0 of [nblk][stride], summed exactly as w_finish_t sums it --
// sum_partials_two in irlb.cu -- so both get the same bits) or from the pair w_norm_exchange left (several GPUs).
struct LastDiag {
  const double* parts = nullptr;
  int nblk = 0, col0 = 0, stride = 0;
  const double* n2p = nullptr;
};
void run_count_stats(adobo_ctx* c, i64* out_total, i64* out_detected, i64* out_expressed);
void run_knn(adobo_ctx* c, const double* comp_dev_local, i64 n_local, int p, int k, int metric = 0);
void run_snn(adobo_ctx* c, const int* nn_dev_all, i64 n_all, i64 row_begin, i64 row_end, int k, double prune,
             i64* n_edges, i64* pruned, i64* total);

// exclusive scan of int64 counts (n elements) into out[n+1] (out[n] = total); in-place allowed
void exclusive_scan_i64(adobo_ctx* c, const i64* in, i64* out, i64 n);

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ i64 warp_sum(i64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// ---- fused all-reduce over peer memory: called by ALL threads of ONE CTA ------------------------------------
// `vec` (global or shared memory of the calling CTA, `count` doubles) is replaced by its sum over the ranks.
// ch = 0: one CTA per rank exchanges up to PEER_SLOT doubles; ch = 1..PEER_CTAS: up to PEER_CH_SLOT doubles.
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <bool SMEM = false>  // SMEM: vec is shared memory of the calling CTA (else global, read through L2)
__device__ __forceinline__ void peer_allreduce_block(const PeerComm& pc, double* vec, int count, int ch = 0) {
  __shared__ unsigned s_seq;
  const int W = pc.world, me = pc.rank;
  if (threadIdx.x == 0) s_seq = pc.seq[ch] + 1u;
  __syncthreads();
  const unsigned seq = s_seq;
  const int par = (int)(seq & 1u);
  const size_t choff = ch == 0 ? 0 : (size_t)PEER_SLOT + (size_t)(ch - 1) * PEER_CH_SLOT;
  // phase 1: my vector into my slot of every rank's mailbox (NVLink stores), then release the flag
  for (int r = 0; r < W; ++r) {
    double* dst = pc.data[r] + ((size_t)(par * W + me)) * (2 * PEER_SLOT) + choff;
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = SMEM ? vec[i] : __ldcg(vec + i);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < W) {
    unsigned* f = pc.flags[threadIdx.x] + ((size_t)(par * W + me) * PEER_CH + ch) * 32;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
  }
  // phase 2: wait for every rank's contribution in MY mailbox (bounded), sum in rank order
  if ((int)threadIdx.x < W) {
    const unsigned* f = pc.flags[me] + ((size_t)(par * W + threadIdx.x) * PEER_CH + ch) * 32;
    unsigned v;
    unsigned long long t0 = 0;
    unsigned polls = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == seq) break;
      if ((++polls & 1023u) == 0) {
        const unsigned long long now = peer_now_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > PEER_TIMEOUT_NS) {
          if (pc.err) atomicExch(pc.err, 1 + (int)threadIdx.x);
          break;
        }
      }
    }
  }
  __syncthreads();
  const double* mine = pc.data[me] + (size_t)(par * W) * (2 * PEER_SLOT) + choff;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double s = 0;
    for (int r = 0; r < W; ++r) s += __ldcv(mine + (size_t)r * (2 * PEER_SLOT) + i);
    vec[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) pc.seq[ch] = seq;
}

// ---- the same exchange without fence and flag round trip: every double travels as ONE 16-byte store {lo, seq, hi, seq}
// (each 8-byte half carries its own copy of the sequence number, so the message is self-validating however the
// store is split on the way: the protocol of NCCL's LL).  The receiver polls the words themselves: one NVLink
// crossing per exchange instead of data + system fence + flag.  vec: shared memory of the calling CTA, count <=
// PEER_LL_MAX; parity double buffering as above.
// seq_pre: this exchange's sequence number (pc.seq[ch] + 1) when the caller has read it already -- v_step_t does in its
// prologue, off the critical path -- else 0.
__device__ __forceinline__ void peer_allreduce_ll(const PeerComm& pc, double* vec, int count, int ch, unsigned seq_pre = 0) {
  __shared__ unsigned s_seq_ll;
  const int W = pc.world, me = pc.rank;
  unsigned seq = seq_pre;
  if (seq_pre == 0) {
    if (threadIdx.x == 0) s_seq_ll = pc.seq[ch] + 1u;
    __syncthreads();
    seq = s_seq_ll;
  }
  const int par = (int)(seq & 1u);
  const int i = threadIdx.x;
  if (i < count) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vec[i]);
    const unsigned lo = (unsigned)bits, hi = (unsigned)(bits >> 32);
    for (int r = 0; r < W; ++r) {
      uint4* dst = pc.ll[r] + ((size_t)(par * W + me) * PEER_CH + ch) * PEER_LL_MAX + i;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
    }
    // every rank's word is polled in the SAME round (independent loads: one L2 round trip per round, not one per
    // rank -- eight serial polls were 3 us of the exchange at N = 8); the sum is then taken in rank order
    const uint4* src0 = pc.ll[me] + ((size_t)(par * W) * PEER_CH + ch) * PEER_LL_MAX + i;
    unsigned wa[PEER_MAX], wc[PEER_MAX];
    unsigned pending = (W >= 32) ? 0xffffffffu : ((1u << W) - 1u);
    unsigned long long t0 = 0;
    unsigned polls = 0;
    while (pending) {
      unsigned b[PEER_MAX], d[PEER_MAX];
#pragma unroll
      for (int r = 0; r < PEER_MAX; ++r)
        if ((pending >> r) & 1u)
          asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(wa[r]), "=r"(b[r]), "=r"(wc[r]), "=r"(d[r])
                       : "l"(src0 + (size_t)r * PEER_CH * PEER_LL_MAX)
                       : "memory");
#pragma unroll
      for (int r = 0; r < PEER_MAX; ++r)
        if (((pending >> r) & 1u) && b[r] == seq && d[r] == seq) pending &= ~(1u << r);
      if (pending && (++polls & 1023u) == 0) {
        const unsigned long long now = peer_now_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > PEER_TIMEOUT_NS) {
          if (pc.err) atomicExch(pc.err, 1 + (__ffs(pending) - 1));
          break;
        }
      }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < PEER_MAX; ++r)   // rank order: the same sum on every rank
      if (r < W) s += __longlong_as_double((long long)(((unsigned long long)wc[r] << 32) | wa[r]));
    vec[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) pc.seq[ch] = seq;
}

// (c, s) of the Jacobi rotation that orthogonalises two columns with squared norms nx, ny and inner
// product ga; identity when the pair is already orthogonal to 1e-15.  The angle comes from fp32
// arithmetic (the fp64 test decides when to stop, so fp32 only costs an occasional extra rotation);
// (c, s) is then renormalised in fp64 so that the rotation is orthogonal to working precision.
// Straight-line code: two independent calls interleave in the instruction stream.
__device__ __forceinline__ float rsqrt_approx(float x) {  // one MUFU.RSQ, no denormal fix-up
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ int jacobi_angle(double ga, double nx, double ny, double& cs, double& sn) {
  const double g2 = ga * ga, ab = nx * ny;
  const bool rot = g2 > 1e-30 * ab;
  const float d32 = (float)(ny - nx), g32 = (float)ga;
  const float h2 = fmaf(d32, d32, 4.0f * g32 * g32);
  const float rh = rsqrt_approx(h2);                           // 1 / hypot(d, 2 gamma)
  const float c2 = fmaf(0.5f * fabsf(d32), rh, 0.5f);          // cos^2 of the rotation angle
  const float rc = rsqrt_approx(c2);
  double c = (double)(c2 * rc), s_ = (double)((d32 >= 0.0f ? g32 : -g32) * rh * rc);
  if (rot && !(h2 > 1e-30f && h2 < 1e30f)) {  // fp32 range exhausted (wildly different norms): fp64 angle
    const double d = ny - nx;
    const double rhd = rsqrt(fma(d, d, 4.0 * g2));
    const double c2d = fma(0.5 * fabs(d), rhd, 0.5);
    const double rcd = rsqrt(c2d);
    c = c2d * rcd;
    s_ = (d >= 0 ? ga : -ga) * rhd * rcd;
  }
  // renormalise: y = rsqrt(c^2 + s^2), two Newton steps from 1 (c^2 + s^2 = 1 +- 1e-6)
  const double w = fma(c, c, s_ * s_);
  double y = fma(-0.5, w, 1.5);
  y = y * fma(-0.5 * w, y * y, 1.5);
  cs = rot ? c * y : 1.0;
  sn = rot ? s_ * y : 0.0;
  return rot ? ((g2 > 1e-16 * ab) ? 3 : 1) : 0;
}

// streaming 128-bit loads/stores that do not pollute L1
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}
#endif

}  // namespace adobo
