#include <chrono>
// C-ABI of adobo_b200 (include/adobo_b200.h): context, NCCL plumbing, stopwatch, and the entry
// points that mirror the reference's functions.  Everything computes on the GPU; there is no
// host fallback -- adobo_ctx_create fails when no CUDA device is usable.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>

#include <algorithm>
#include <cstdlib>
#include <climits>
#include <mutex>
#include <numeric>

#include "common.cuh"

namespace adobo {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void upload_normalized_values(adobo_ctx* c, const double* values);  // normalize.cu
void run_synth_counts(adobo_ctx* c, i64 n_rows, int g, i64 row_begin, uint64_t seed, const float* table_host, int n_clusters,
                      i64* out_nnz);  // synth.cu

// ---- caching device allocator ---------------------------------------------------------------------
namespace {
struct Pool {
  std::mutex mu;
  std::map<int, std::multimap<size_t, void*>> free_blocks;  // device -> size -> block
  std::map<void*, std::pair<int, size_t>> live;             // block -> (device, size)
};
Pool& pool() {
  static Pool* p = new Pool();  // leaked on purpose: DevBufs may outlive static destruction order
  return *p;
}
}  // namespace

void* pool_alloc(size_t bytes) {
  const size_t size = (bytes + 511) & ~(size_t)511;
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  Pool& P = pool();
  {
    std::lock_guard<std::mutex> g(P.mu);
    auto& fl = P.free_blocks[dev];
    auto it = fl.lower_bound(size);
    if (it != fl.end() && it->first <= size + size / 4 + 4096) {  // reuse when at most 25 % larger
      void* p = it->second;
      const size_t sz = it->first;
      fl.erase(it);
      P.live[p] = std::make_pair(dev, sz);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, size);
  if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
    cudaGetLastError();
    pool_trim();
    e = cudaMalloc(&p, size);
  }
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) -> %s", size, cudaGetErrorString(e));
    throw Fail{ADOBO_ERR_CUDA};
  }
  std::lock_guard<std::mutex> g(P.mu);
  P.live[p] = std::make_pair(dev, size);
  return p;
}

void pool_free(void* p) {
  if (!p) return;
  Pool& P = pool();
  std::lock_guard<std::mutex> g(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) return;
  P.free_blocks[it->second.first].insert(std::make_pair(it->second.second, p));
  P.live.erase(it);
}

void pool_trim() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  Pool& P = pool();
  std::vector<void*> blocks;
  {
    std::lock_guard<std::mutex> g(P.mu);
    for (auto& kv : P.free_blocks[dev]) blocks.push_back(kv.second);
    P.free_blocks[dev].clear();
  }
  for (void* b : blocks) cudaFree(b);
}

// ---- NCCL, resolved at run time so that single-GPU use does not need the library -------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static void load_nccl() {
  if (g_nccl.lib) return;
  // libnccl.so.2 is already mapped when torch.distributed (nccl backend) is in the process
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  REQUIRE(g_nccl.lib, ADOBO_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                    \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                       \
  REQUIRE(g_nccl.field, ADOBO_ERR_NCCL, "libnccl is missing symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

#define NCCL_CHECK(expr)                                                                         \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != ncclSuccess) {                                                                     \
      adobo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      throw adobo::Fail{ADOBO_ERR_NCCL};                                                         \
    }                                                                                            \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
};

void allreduce_sum_f64(adobo_ctx* c, double* dev, size_t count) {
  if (c->world <= 1 || count == 0) return;
  NCCL_CHECK(g_nccl.AllReduce(dev, dev, count, ncclFloat64, ncclSum, c->comm->comm, c->stream));
  c->launches++;
}
void allreduce_sum_i64(adobo_ctx* c, i64* dev, size_t count) {
  if (c->world <= 1 || count == 0) return;
  NCCL_CHECK(g_nccl.AllReduce(dev, dev, count, ncclInt64, ncclSum, c->comm->comm, c->stream));
  c->launches++;
}
void allgatherv_bytes(adobo_ctx* c, const void* src, void* dst, const std::vector<size_t>& bytes) {
  if (c->world <= 1) {
    if (src != dst) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes[0], cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  size_t off = 0;
  NCCL_CHECK(g_nccl.GroupStart());
  for (int r = 0; r < c->world; ++r) {
    char* d = (char*)dst + off;
    if (bytes[r] > 0) NCCL_CHECK(g_nccl.Broadcast(r == c->rank ? src : d, d, bytes[r], ncclChar, r, c->comm->comm, c->stream));
    off += bytes[r];
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  c->launches++;
}
void allgather_i64_host(adobo_ctx* c, i64 mine, std::vector<i64>& all) {
  all.assign(c->world, 0);
  if (c->world <= 1) {
    all[0] = mine;
    return;
  }
  DevBuf<i64> buf;
  buf.ensure((size_t)c->world + 1);
  CUDA_CHECK(cudaMemcpyAsync(buf.p + c->world, &mine, sizeof(i64), cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(buf.p + c->world, buf.p, 1, ncclInt64, c->comm->comm, c->stream));
  CUDA_CHECK(cudaMemcpyAsync(all.data(), buf.p, sizeof(i64) * c->world, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// ---- profiling ----------------------------------------------------------------------------------
static cudaEvent_t get_event(adobo_ctx* c) {
  if (!c->event_pool.empty()) {
    cudaEvent_t e = c->event_pool.back();
    c->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CUDA_CHECK(cudaEventCreate(&e));
  return e;
}
void prof_begin(adobo_ctx* c, const char* name) {
  PendingEvent pe;
  pe.name = name;
  pe.a = get_event(c);
  pe.b = get_event(c);
  pe.work = c->next_work;
  c->next_work = 0;
  CUDA_CHECK(cudaEventRecord(pe.a, c->stream));
  c->pending.push_back(pe);
}
void prof_end(adobo_ctx* c) { CUDA_CHECK(cudaEventRecord(c->pending.back().b, c->stream)); }
static void prof_collect(adobo_ctx* c) {
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (PendingEvent& pe : c->pending) {
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, pe.a, pe.b));
    ProfItem& it = c->prof[pe.name];
    it.ms += ms;
    it.launches += 1;
    it.work += pe.work;
    c->event_pool.push_back(pe.a);
    c->event_pool.push_back(pe.b);
  }
  c->pending.clear();
}

// Mailboxes of the fused all-reduce: one cudaMalloc per rank, handles exchanged over NCCL, peers mapped
// with CUDA IPC (one process per GPU on one node).
static void setup_peer(adobo_ctx* c) {
  const int W = c->world;
  const size_t data_bytes = sizeof(double) * 2 * (size_t)W * (2 * PEER_SLOT);
  const size_t flag_bytes = sizeof(unsigned) * 2 * (size_t)W * PEER_CH * 32;
  const size_t ll_bytes = sizeof(uint4) * 2 * (size_t)W * PEER_CH * PEER_LL_MAX;
  const size_t bytes = data_bytes + flag_bytes + 256 + ll_bytes;  // 256: the PEER_CH sequence counters
  void* mb = nullptr;
  CUDA_CHECK(cudaMalloc(&mb, bytes));
  CUDA_CHECK(cudaMemset(mb, 0, bytes));
  c->peer_mailbox = mb;
  cudaIpcMemHandle_t mine;
  CUDA_CHECK(cudaIpcGetMemHandle(&mine, mb));
  DevBuf<char> hbuf;
  hbuf.ensure(sizeof(cudaIpcMemHandle_t) * (size_t)(W + 1));
  CUDA_CHECK(cudaMemcpyAsync(hbuf.p + sizeof(mine) * W, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  NCCL_CHECK(g_nccl.AllGather(hbuf.p + sizeof(mine) * W, hbuf.p, sizeof(mine), ncclChar, c->comm->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all(W);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), hbuf.p, sizeof(mine) * W, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PeerComm pc;
  pc.world = W;
  pc.rank = c->rank;
  for (int r = 0; r < W; ++r) {
    void* ptr = mb;
    if (r != c->rank) CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess));
    pc.data[r] = (double*)ptr;
    pc.flags[r] = (unsigned*)((char*)ptr + data_bytes);
    pc.ll[r] = (uint4*)((char*)ptr + data_bytes + flag_bytes + 256);
  }
  pc.seq = (unsigned*)((char*)mb + data_bytes + flag_bytes);
  // nobody may write into a mailbox before every rank has zeroed and mapped it
  DevBuf<double> one;
  one.ensure(1);
  CUDA_CHECK(cudaMemsetAsync(one.p, 0, sizeof(double), c->stream));
  NCCL_CHECK(g_nccl.AllReduce(one.p, one.p, 1, ncclFloat64, ncclSum, c->comm->comm, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->peer = pc;
}

static void refresh_partition(adobo_ctx* c) {
  std::vector<i64> all;
  allgather_i64_host(c, c->n, all);
  c->n_total = 0;
  c->row_offset = 0;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) c->row_offset = c->n_total;
    c->n_total += all[r];
  }
}

static void upload_pattern(adobo_ctx* c, i64 n, int g, const i64* indptr, const int* indices) {
  REQUIRE(n >= 0 && g >= 1, ADOBO_ERR_ARG, "upload: bad shape %lld x %d", (long long)n, g);
  REQUIRE(indptr && indptr[0] == 0, ADOBO_ERR_ARG, "upload: indptr[0] must be 0");
  const i64 nnz = indptr[n];
  REQUIRE(nnz >= 0, ADOBO_ERR_ARG, "upload: negative nnz");
  c->n = n;
  c->g = g;
  c->nnz = nnz;
  c->indptr.ensure((size_t)n + 1);
  c->indices.ensure((size_t)nnz);
  CUDA_CHECK(cudaMemcpyAsync(c->indptr.p, indptr, sizeof(i64) * (n + 1), cudaMemcpyHostToDevice, c->stream));
  if (nnz) CUDA_CHECK(cudaMemcpyAsync(c->indices.p, indices, sizeof(int) * nnz, cudaMemcpyHostToDevice, c->stream));
  c->have_counts = c->have_norm = c->have_moments = c->have_pca = c->have_knn = false;
  c->have_suboffs = false;
  refresh_partition(c);
}

static void nn_to_host(adobo_ctx* c, int64_t* out_idx) {
  const size_t cnt = (size_t)c->knn_n * c->k;
  std::vector<int> tmp(cnt);
  CUDA_CHECK(cudaMemcpyAsync(tmp.data(), c->nn.p, sizeof(int) * cnt, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < cnt; ++i) out_idx[i] = (int64_t)tmp[i] + 1;  // 1-based, clustering.py:51
}

static void do_snn(adobo_ctx* c, const int* nn_local_dev, i64 n_local, int k, double prune, i64* ne, i64* pr, i64* tot) {
  std::vector<i64> all_n;
  allgather_i64_host(c, n_local, all_n);
  i64 n = 0, begin = 0;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) begin = n;
    n += all_n[r];
  }
  DevBuf<int> gathered;
  const int* nn = nn_local_dev;
  if (c->world > 1) {
    gathered.ensure((size_t)n * k);
    std::vector<size_t> bytes(c->world);
    for (int r = 0; r < c->world; ++r) bytes[r] = (size_t)all_n[r] * k * sizeof(int);
    allgatherv_bytes(c, nn_local_dev, gathered.p, bytes);
    nn = gathered.p;
  }
  run_snn(c, nn, n, begin, begin + n_local, k, prune, ne, pr, tot);
  if (c->world > 1) {
    DevBuf<i64> t;
    t.ensure(2);
    i64 v[2] = {*pr, *tot};
    CUDA_CHECK(cudaMemcpyAsync(t.p, v, sizeof(v), cudaMemcpyHostToDevice, c->stream));
    allreduce_sum_i64(c, t.p, 2);
    CUDA_CHECK(cudaMemcpyAsync(v, t.p, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *pr = v[0];
    *tot = v[1];
  }
}

static void do_pca(adobo_ctx* c, std::vector<int> genes, int ncomp, double tol, int maxit, const double* v0, bool scale,
                   bool var_weigh, int* steps, int* nmult) {
  std::sort(genes.begin(), genes.end());
  run_gather_subset(c, genes, scale);
  IrlbResult r = run_irlb<float>(c, c->sub, scale ? c->mu.p : nullptr, scale ? c->rs.p : nullptr, ncomp, tol, maxit, v0,
                                 var_weigh, c->comp, c->sv, c->contr);
  c->p = ncomp;
  c->have_pca = true;
  if (steps) *steps = r.steps;
  if (nmult) *nmult = r.nmult;
}

}  // namespace adobo

using namespace adobo;

#define API_BEGIN try {
#define API_END                                      \
  }                                                  \
  catch (const adobo::Fail& f) { return f.code; }    \
  catch (const std::exception& e) {                  \
    adobo::set_error("exception: %s", e.what());     \
    return ADOBO_ERR_STATE;                          \
  }                                                  \
  return ADOBO_OK;

extern "C" {

int adobo_version(void) { return 100; }
const char* adobo_last_error(void) { return g_err; }

int adobo_ctx_create(int device, adobo_ctx** out) {
  API_BEGIN
  REQUIRE(out, ADOBO_ERR_ARG, "adobo_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  REQUIRE(e == cudaSuccess && ndev > 0, ADOBO_ERR_CUDA,
          "no CUDA device available (%s); adobo_b200 has no CPU fallback", cudaGetErrorString(e));
  REQUIRE(device >= 0 && device < ndev, ADOBO_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  REQUIRE(prop.major >= 10, ADOBO_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
          prop.major, prop.minor);
  adobo_ctx* c = new adobo_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
  CUDA_CHECK(cudaMallocHost(&c->pinned, 4096));
  *out = c;
  API_END
}

void adobo_ctx_destroy(adobo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  // graphs that captured NCCL kernels must be gone before the communicator is destroyed
  if (c->irlb_ws && c->irlb_ws_free) c->irlb_ws_free(c->irlb_ws);
  c->irlb_ws = nullptr;
  if (c->peer_mailbox) {
    for (int r = 0; r < c->peer.world; ++r)
      if (r != c->peer.rank && c->peer.data[r]) cudaIpcCloseMemHandle(c->peer.data[r]);
    cudaFree(c->peer_mailbox);
    c->peer_mailbox = nullptr;
  }
  if (c->comm) {
    if (c->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->comm);
    delete c->comm;
  }
  for (cudaEvent_t& e : c->events)
    if (e) cudaEventDestroy(e);
  for (PendingEvent& pe : c->pending) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  if (c->join_ev) cudaEventDestroy(c->join_ev);
  cudaStream_t st = c->stream, side = c->side;
  delete c;
  if (side) cudaStreamDestroy(side);
  cudaStreamDestroy(st);
  pool_trim();
}

int adobo_sync(adobo_ctx* c) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  API_END
}

int adobo_comm_unique_id(void* id128) {
  API_BEGIN
  load_nccl();
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  API_END
}

int adobo_comm_init(adobo_ctx* c, int world, int rank, const void* id128) {
  API_BEGIN
  REQUIRE(world >= 1 && rank >= 0 && rank < world, ADOBO_ERR_ARG, "comm_init: bad world/rank %d/%d", world, rank);
  CUDA_CHECK(cudaSetDevice(c->device));
  if (world > 1) {
    load_nccl();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    c->comm = new Comm();
    NCCL_CHECK(g_nccl.CommInitRank(&c->comm->comm, world, id, rank));
  }
  c->world = world;
  c->rank = rank;
  if (world > 1 && world <= PEER_MAX && getenv("ADOBO_B200_NO_PEER") == nullptr) setup_peer(c);
  if (c->have_counts || c->have_norm) refresh_partition(c);
  API_END
}

int adobo_event_record(adobo_ctx* c, int slot) {
  API_BEGIN
  REQUIRE(slot >= 0 && slot < 64, ADOBO_ERR_ARG, "event slot %d out of range", slot);
  CUDA_CHECK(cudaSetDevice(c->device));
  if (!c->events[slot]) CUDA_CHECK(cudaEventCreate(&c->events[slot]));
  CUDA_CHECK(cudaEventRecord(c->events[slot], c->stream));
  API_END
}

int adobo_event_elapsed_ms(adobo_ctx* c, int a, int b, float* ms) {
  API_BEGIN
  REQUIRE(a >= 0 && a < 64 && b >= 0 && b < 64 && c->events[a] && c->events[b], ADOBO_ERR_ARG, "bad event slots");
  CUDA_CHECK(cudaEventSynchronize(c->events[b]));
  CUDA_CHECK(cudaEventElapsedTime(ms, c->events[a], c->events[b]));
  API_END
}

// Calibration of the profiler itself: an empty kernel between the same pair of events.  Its time is what every
// profiled launch carries on top of its own duration (event record + launch latency); reported as "_event_overhead".
__global__ void prof_noop() {}

int adobo_profile(adobo_ctx* c, int on) {
  API_BEGIN
  prof_collect(c);
  if (on) c->prof.clear();
  c->profiling = on != 0;
  if (on) {
    const i64 launches = c->launches;
    for (int q = 0; q < 64; ++q) LAUNCH(c, "_event_overhead", (prof_noop<<<1, 32, 0, c->stream>>>()));
    c->launches = launches;  // not part of the path
    prof_collect(c);
  }
  API_END
}

int adobo_profile_read(adobo_ctx* c, int max_items, char* names, float* ms, int64_t* launches, double* work,
                       int* n_items) {
  API_BEGIN
  prof_collect(c);
  int i = 0;
  for (auto& kv : c->prof) {
    if (i >= max_items) break;
    strncpy(names + 32 * i, kv.first.c_str(), 31);
    names[32 * i + 31] = 0;
    ms[i] = (float)kv.second.ms;
    launches[i] = kv.second.launches;
    work[i] = kv.second.work;
    ++i;
  }
  *n_items = i;
  API_END
}

int64_t adobo_launch_count(adobo_ctx* c) { return c->launches; }

int adobo_host_register(void* ptr, size_t bytes) {
  API_BEGIN
  CUDA_CHECK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  API_END
}
int adobo_host_unregister(void* ptr) {
  API_BEGIN
  CUDA_CHECK(cudaHostUnregister(ptr));
  API_END
}

int adobo_l2_flush(adobo_ctx* c) {
  API_BEGIN
  const size_t sz = 256u << 20;
  c->flush_buf.ensure(sz);
  CUDA_CHECK(cudaMemsetAsync(c->flush_buf.p, (int)(c->launches & 0xff), sz, c->stream));
  API_END
}

int adobo_upload_counts(adobo_ctx* c, int64_t n, int32_t g, const int64_t* indptr, const int32_t* indices,
                        const int32_t* counts) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  upload_pattern(c, n, g, (const i64*)indptr, indices);
  c->counts.ensure((size_t)c->nnz);
  if (c->nnz) CUDA_CHECK(cudaMemcpyAsync(c->counts.p, counts, sizeof(int) * c->nnz, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->have_counts = true;
  API_END
}

int adobo_synth_counts(adobo_ctx* c, int64_t n_rows, int32_t n_genes, int64_t row_begin, uint64_t seed,
                       const float* lam_table, int32_t n_clusters, int64_t* out_nnz) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  i64 nnz = 0;
  run_synth_counts(c, n_rows, n_genes, row_begin, seed, lam_table, n_clusters, &nnz);
  c->have_counts = true;
  c->have_norm = c->have_moments = c->have_pca = c->have_knn = false;
  c->have_suboffs = false;
  refresh_partition(c);
  if (out_nnz) *out_nnz = nnz;
  API_END
}

int adobo_download_counts(adobo_ctx* c, int64_t* indptr, int32_t* indices, int32_t* counts) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(c->have_counts, ADOBO_ERR_STATE, "download_counts: no count matrix on the device");
  if (indptr) CUDA_CHECK(cudaMemcpyAsync(indptr, c->indptr.p, sizeof(i64) * (c->n + 1), cudaMemcpyDeviceToHost, c->stream));
  if (indices && c->nnz) CUDA_CHECK(cudaMemcpyAsync(indices, c->indices.p, sizeof(int) * c->nnz, cudaMemcpyDeviceToHost, c->stream));
  if (counts && c->nnz) CUDA_CHECK(cudaMemcpyAsync(counts, c->counts.p, sizeof(int) * c->nnz, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  API_END
}

int adobo_upload_normalized(adobo_ctx* c, int64_t n, int32_t g, const int64_t* indptr, const int32_t* indices,
                            const double* values) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  upload_pattern(c, n, g, (const i64*)indptr, indices);
  upload_normalized_values(c, values);
  API_END
}

int adobo_normalize(adobo_ctx* c, double scaling_factor, int log_mode, double small_const, int64_t* out_libsize,
                    double* out_values) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(log_mode >= 0 && log_mode <= 3, ADOBO_ERR_ARG, "normalize: unknown log mode %d", log_mode);
  run_normalize(c, scaling_factor, log_mode, small_const, (i64*)out_libsize, out_values);
  API_END
}

int adobo_count_stats(adobo_ctx* c, int64_t* out_total_reads, int64_t* out_detected_genes, int64_t* out_expressed) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  run_count_stats(c, (i64*)out_total_reads, (i64*)out_detected_genes, (i64*)out_expressed);
  if (c->world > 1 && out_expressed) {  // cells are sharded: the per-gene counts are sums over the ranks
    DevBuf<i64> e;
    e.ensure((size_t)std::max<i64>(c->g, 1));
    CUDA_CHECK(cudaMemcpyAsync(e.p, out_expressed, sizeof(i64) * c->g, cudaMemcpyHostToDevice, c->stream));
    allreduce_sum_i64(c, e.p, (size_t)c->g);
    CUDA_CHECK(cudaMemcpyAsync(out_expressed, e.p, sizeof(i64) * c->g, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  API_END
}

int adobo_gene_moments(adobo_ctx* c, double* out_mean, double* out_var) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  run_moments(c);
  if (out_mean) memcpy(out_mean, c->h_mean.data(), sizeof(double) * c->g);
  if (out_var) memcpy(out_var, c->h_var.data(), sizeof(double) * c->g);
  API_END
}

int adobo_hvg_seurat(adobo_ctx* c, int32_t ngenes, int32_t num_bins, int32_t* out_idx, int32_t* out_n, double* out_z) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(ngenes >= 0, ADOBO_ERR_ARG, "seurat: ngenes must be >= 0");
  if (!c->have_moments) run_moments(c);
  std::vector<double> z;
  run_seurat_host(c, ngenes, num_bins, c->hvg_order, z);
  const int m = std::min<int>(ngenes, (int)c->hvg_order.size());
  if (out_idx)
    for (int i = 0; i < m; ++i) out_idx[i] = c->hvg_order[i];
  if (out_n) *out_n = m;
  if (out_z) memcpy(out_z, z.data(), sizeof(double) * c->g);
  API_END
}

int adobo_irlb(adobo_ctx* c, const int32_t* gene_idx, int32_t h, int32_t ncomp, double tol, int32_t maxit,
               const double* v0, int32_t scale, int32_t var_weigh, double* out_comp, double* out_s, double* out_contr,
               int32_t* out_steps, int32_t* out_nmult) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(gene_idx && h >= 1 && v0, ADOBO_ERR_ARG, "irlb: gene_idx / v0 missing");
  std::vector<int> genes(gene_idx, gene_idx + h);
  do_pca(c, genes, ncomp, tol, maxit, v0, scale != 0, var_weigh != 0, out_steps, out_nmult);
  if (out_comp) CUDA_CHECK(cudaMemcpyAsync(out_comp, c->comp.p, sizeof(double) * c->n * ncomp, cudaMemcpyDeviceToHost, c->stream));
  if (out_s) CUDA_CHECK(cudaMemcpyAsync(out_s, c->sv.p, sizeof(double) * ncomp, cudaMemcpyDeviceToHost, c->stream));
  if (out_contr) CUDA_CHECK(cudaMemcpyAsync(out_contr, c->contr.p, sizeof(double) * (size_t)h * ncomp, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  API_END
}

int adobo_gather_subset(adobo_ctx* c, const int32_t* gene_idx, int32_t h, int32_t scale) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(gene_idx && h >= 1, ADOBO_ERR_ARG, "gather_subset: gene_idx missing");
  std::vector<int> genes(gene_idx, gene_idx + h);
  std::sort(genes.begin(), genes.end());
  run_gather_subset(c, genes, scale != 0);
  API_END
}

int adobo_irlb_permuted(adobo_ctx* c, const int32_t* perm_cols, int32_t n_perm, uint64_t perm_seed, int32_t ncomp, double tol,
                        int32_t maxit, const double* v0, int32_t var_weigh, double* out_contr, int32_t* out_steps,
                        int32_t* out_nmult) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(v0 && n_perm >= 0 && (perm_cols || n_perm == 0), ADOBO_ERR_ARG, "irlb_permuted: bad arguments");
  const int h = c->sub.cols;
  // the two-kernel Lanczos step reads the CSR only; the other paths (h > 1024, hooks) also need the CSC copy
  const bool need_csc = h > 1024 || getenv("ADOBO_B200_IRLB") != nullptr;
  run_permute_subset(c, perm_cols, n_perm, perm_seed, need_csc);
  if (!need_csc) {  // never read on this path; keep the signature of the workspace stable across permutations
    c->perm.cptr.ensure(1), c->perm.ridx.ensure(1), c->perm.cval.ensure(1);
  }
  IrlbResult r = run_irlb<float>(c, c->perm, c->sub_scaled ? c->mu.p : nullptr, c->sub_scaled ? c->rs.p : nullptr, ncomp, tol,
                                 maxit, v0, var_weigh != 0, c->perm_u, c->perm_s, c->perm_v);
  if (out_contr) CUDA_CHECK(cudaMemcpyAsync(out_contr, c->perm_v.p, sizeof(double) * (size_t)h * ncomp, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (out_steps) *out_steps = r.steps;
  if (out_nmult) *out_nmult = r.nmult;
  API_END
}

int64_t adobo_perm_row(int64_t n_cells, uint64_t perm_seed, int32_t col, int64_t row) {
  if (n_cells < 1 || row < 0 || row >= n_cells) return -1;
  return perm_apply(perm_key(perm_seed, col), n_cells, row);
}

int adobo_lanczos_csr(adobo_ctx* c, int64_t rows, int32_t cols, const int64_t* indptr, const int32_t* indices,
                      const double* values, int32_t nval, double tol, int32_t maxit, const double* v0, double* out_U,
                      double* out_s, double* out_V, int32_t* out_steps, int32_t* out_nmult) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(rows >= 1 && cols >= 1 && indptr && v0, ADOBO_ERR_ARG, "lanczos: bad arguments");
  REQUIRE(c->world == 1, ADOBO_ERR_STATE, "adobo_lanczos_csr is single-GPU; use adobo_irlb for sharded data");
  SparsePair<double> A;
  A.rows = rows;
  A.cols = cols;
  A.nnz = indptr[rows];
  A.rptr.ensure((size_t)rows + 1);
  A.cidx.ensure((size_t)A.nnz);
  A.rval.ensure((size_t)A.nnz);
  CUDA_CHECK(cudaMemcpyAsync(A.rptr.p, indptr, sizeof(i64) * (rows + 1), cudaMemcpyHostToDevice, c->stream));
  if (A.nnz) {
    CUDA_CHECK(cudaMemcpyAsync(A.cidx.p, indices, sizeof(int) * A.nnz, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(A.rval.p, values, sizeof(double) * A.nnz, cudaMemcpyHostToDevice, c->stream));
  }
  build_csc<double>(c, A);
  DevBuf<double> U, S, V;
  IrlbResult r = run_irlb<double>(c, A, nullptr, nullptr, nval, tol, maxit, v0, false, U, S, V);
  if (out_U) CUDA_CHECK(cudaMemcpyAsync(out_U, U.p, sizeof(double) * rows * nval, cudaMemcpyDeviceToHost, c->stream));
  if (out_s) CUDA_CHECK(cudaMemcpyAsync(out_s, S.p, sizeof(double) * nval, cudaMemcpyDeviceToHost, c->stream));
  if (out_V) CUDA_CHECK(cudaMemcpyAsync(out_V, V.p, sizeof(double) * (size_t)cols * nval, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (out_steps) *out_steps = r.steps;
  if (out_nmult) *out_nmult = r.nmult;
  API_END
}

int adobo_knn_metric(adobo_ctx* c, const double* comp, int64_t n, int32_t p, int32_t k, int32_t metric, int64_t* out_idx) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  if (comp) {
    REQUIRE(n >= 1 && p >= 1, ADOBO_ERR_ARG, "knn: bad shape");
    DevBuf<double> x;
    x.ensure((size_t)n * p);
    CUDA_CHECK(cudaMemcpyAsync(x.p, comp, sizeof(double) * n * p, cudaMemcpyHostToDevice, c->stream));
    run_knn(c, x.p, n, p, k, metric);
  } else {
    REQUIRE(c->have_pca, ADOBO_ERR_STATE, "Compute principal components first using adobo.dr.pca(...)");
    run_knn(c, c->comp.p, c->n, c->p, k, metric);
  }
  if (out_idx) nn_to_host(c, out_idx);
  API_END
}

int adobo_knn(adobo_ctx* c, const double* comp, int64_t n, int32_t p, int32_t k, int64_t* out_idx) {
  return adobo_knn_metric(c, comp, n, p, k, ADOBO_METRIC_EUCLIDEAN, out_idx);
}

int adobo_upload_embedding(adobo_ctx* c, const double* comp, int64_t n, int32_t p) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(comp && n >= 1 && p >= 1, ADOBO_ERR_ARG, "upload_embedding: bad arguments");
  c->comp.ensure((size_t)n * p);
  CUDA_CHECK(cudaMemcpyAsync(c->comp.p, comp, sizeof(double) * n * p, cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->n = n;
  c->p = p;
  c->have_pca = true;
  c->h_sub_genes.clear();
  refresh_partition(c);
  API_END
}

int adobo_snn(adobo_ctx* c, const int64_t* nn_idx, int64_t n, int32_t k, double prune_snn, int64_t* out_n_edges,
              int64_t* out_pruned, int64_t* out_total_links) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  i64 ne = 0, pr = 0, tot = 0;
  if (nn_idx) {
    REQUIRE(n >= 1 && k >= 1, ADOBO_ERR_ARG, "snn: bad shape");
    std::vector<int> h((size_t)n * k);
    for (size_t i = 0; i < h.size(); ++i) {
      const int64_t v = nn_idx[i] - 1;
      REQUIRE(v >= 0 && v < INT_MAX, ADOBO_ERR_ARG, "snn: nn_idx holds an index outside 1..n");
      h[i] = (int)v;
    }
    DevBuf<int> d;
    d.ensure(h.size());
    CUDA_CHECK(cudaMemcpyAsync(d.p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, c->stream));
    do_snn(c, d.p, n, k, prune_snn, &ne, &pr, &tot);
  } else {
    REQUIRE(c->have_knn, ADOBO_ERR_STATE, "snn: run adobo_knn first");
    do_snn(c, c->nn.p, c->knn_n, c->k, prune_snn, &ne, &pr, &tot);
  }
  if (out_n_edges) *out_n_edges = ne;
  if (out_pruned) *out_pruned = pr;
  if (out_total_links) *out_total_links = tot;
  API_END
}

int adobo_snn_fetch(adobo_ctx* c, int32_t* out_source, int32_t* out_target) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  if (c->n_edges > 0) {
    CUDA_CHECK(cudaMemcpyAsync(out_source, c->e_src.p, sizeof(int) * c->n_edges, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(out_target, c->e_dst.p, sizeof(int) * c->n_edges, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  API_END
}

int adobo_snn_fetch_counts(adobo_ctx* c, int32_t* out_v) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(out_v, ADOBO_ERR_ARG, "snn_fetch_counts: out_v is NULL");
  if (c->n_edges > 0) CUDA_CHECK(cudaMemcpyAsync(out_v, c->e_v.p, sizeof(int) * c->n_edges, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  API_END
}

int adobo_result_sizes(adobo_ctx* c, int64_t* out8) {
  API_BEGIN
  REQUIRE(out8, ADOBO_ERR_ARG, "result_sizes: out is NULL");
  out8[0] = c->n;
  out8[1] = c->have_pca ? c->p : 0;
  out8[2] = c->have_pca ? (int64_t)c->h_sub_genes.size() : 0;
  out8[3] = c->have_knn ? c->k : 0;
  out8[4] = c->have_knn ? c->knn_n : 0;
  out8[5] = c->n_edges;
  out8[6] = c->nnz;
  out8[7] = c->g;
  API_END
}

int adobo_embed_path(adobo_ctx* c, double scaling_factor, int32_t ngenes, int32_t ncomp, const double* v0, int32_t k,
                     double prune_snn, int32_t* out_steps, int32_t* out_nmult, int64_t* out_n_edges) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  // ADOBO_B200_DEBUG: wall-clock of the stages (each synchronised at its end) on stderr
  const bool dbg = getenv("ADOBO_B200_DEBUG") != nullptr;
  double t[7];
  auto stamp = [&](int i) {
    if (!dbg) return;
    cudaStreamSynchronize(c->stream);
    t[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  stamp(0);
  run_normalize(c, scaling_factor, ADOBO_LOG_2, 1.0, nullptr, nullptr);
  stamp(1);
  run_moments(c);
  stamp(2);
  std::vector<double> z;
  run_seurat_host(c, ngenes, 20, c->hvg_order, z);
  const int h = std::min<int>(ngenes, (int)c->hvg_order.size());
  std::vector<int> genes(c->hvg_order.begin(), c->hvg_order.begin() + h);
  stamp(3);
  do_pca(c, genes, ncomp, 0.0001, 1000, v0, true, true, out_steps, out_nmult);
  stamp(4);
  run_knn(c, c->comp.p, c->n, c->p, k);
  stamp(5);
  i64 ne = 0, pr = 0, tot = 0;
  do_snn(c, c->nn.p, c->knn_n, k, prune_snn, &ne, &pr, &tot);
  stamp(6);
  if (dbg)
    fprintf(stderr, "[adobo_b200] embed_path ms: normalize %.3f  moments %.3f  seurat(host) %.3f  pca %.3f  knn %.3f  snn %.3f\n",
            t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]);
  if (out_n_edges) *out_n_edges = ne;
  API_END
}

int adobo_fetch_pca(adobo_ctx* c, double* out_comp, double* out_s, double* out_contr, int32_t* out_genes) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(c->have_pca, ADOBO_ERR_STATE, "no PCA result on the device");
  const int h = (int)c->h_sub_genes.size();
  if (out_comp) CUDA_CHECK(cudaMemcpyAsync(out_comp, c->comp.p, sizeof(double) * c->n * c->p, cudaMemcpyDeviceToHost, c->stream));
  if (out_s) CUDA_CHECK(cudaMemcpyAsync(out_s, c->sv.p, sizeof(double) * c->p, cudaMemcpyDeviceToHost, c->stream));
  if (out_contr) CUDA_CHECK(cudaMemcpyAsync(out_contr, c->contr.p, sizeof(double) * (size_t)h * c->p, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (out_genes) memcpy(out_genes, c->h_sub_genes.data(), sizeof(int) * h);
  API_END
}

int adobo_fetch_knn(adobo_ctx* c, int64_t* out_idx) {
  API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->device));
  REQUIRE(c->have_knn, ADOBO_ERR_STATE, "no kNN result on the device");
  nn_to_host(c, out_idx);
  API_END
}

}  // extern "C"
