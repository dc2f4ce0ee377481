// pxr_api.cu — context, error handling, NCCL plumbing and the host-side integer algorithms of
// the C-ABI (include/pxr.h).
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <cstring>
#include <limits>
#include <set>
#include <tuple>
#include <unordered_map>

#include "pxr_internal.h"
#include "pxr_device.cuh"

namespace pxr {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}
int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return status;
}

// ---- NCCL through dlopen (libnccl.so.2: torch's bundled copy if already loaded, else the system one)
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  void* InitRank = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
struct Uid128 { char b[128]; };
static int load_nccl() {
  if (g_nccl.h) return PXR_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(PXR_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.h = h;
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.InitRank = dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.InitRank || !g_nccl.AllReduce)
    return fail(PXR_ERR_NCCL, "libnccl.so.2 lacks the expected symbols");
  return PXR_OK;
}

int allreduce_f64(pxr_ctx* ctx, double* dptr, size_t count, bool max_op) {
  if (!ctx->nccl_comm || ctx->world <= 1 || count == 0) return PXR_OK;
  // ncclFloat64 = 8, ncclSum = 0, ncclMax = 2
  const int rc = g_nccl.AllReduce(dptr, dptr, count, 8, max_op ? 2 : 0, ctx->nccl_comm, ctx->stream);
  if (rc != 0) return fail(PXR_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->nccl_collectives++;
  return PXR_OK;
}

int allreduce_f64_oop(pxr_ctx* ctx, const double* send, double* recv, size_t count) {
  if (count == 0) return PXR_OK;
  if (!ctx->nccl_comm || ctx->world <= 1) {
    if (send != recv) PXR_CUDA(cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return PXR_OK;
  }
  const int rc = g_nccl.AllReduce(send, recv, count, 8, 0, ctx->nccl_comm, ctx->stream);
  if (rc != 0) return fail(PXR_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->nccl_collectives++;
  return PXR_OK;
}

int allgather_bytes(pxr_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes) {
  if (ctx->world <= 1 || !ctx->nccl_comm) {
    if (bytes) PXR_CUDA(cudaMemcpyAsync(recv_dev, send_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return PXR_OK;
  }
  if (!g_nccl.AllGather) return fail(PXR_ERR_NCCL, "libnccl.so.2 lacks ncclAllGather");
  if (bytes == 0) return PXR_OK;
  const int rc = g_nccl.AllGather(send_dev, recv_dev, bytes, 0 /* ncclInt8 */, ctx->nccl_comm, ctx->stream);
  if (rc != 0) return fail(PXR_ERR_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->nccl_collectives++;
  return PXR_OK;
}

// ---- peer mailboxes -------------------------------------------------------------------------------------------
// One launch, one warp: lane q < world writes this rank's payload into rank q's mailbox row [parity][rank] (peer
// memory over NVLink, or local memory for q == rank), fences, then publishes the epoch in q's sequence word with a
// system-scope release.  Lane q then acquires "source q has delivered epoch e" from the local sequence words and lane 0
// folds the `world` rows in rank order.  Two parities: a rank can only start epoch e+2 after every peer has SENT e+1,
// i.e. after every peer finished READING e, so the row of parity (e & 1) is free again.
struct MboxArgs {
  double* peer[16]; unsigned long long* seq_peer[16];
  const double* local; const unsigned long long* seq_local;
  const double* payload; double* out;
  int rank, world, n_sum, n_max;
  unsigned long long epoch;
  int* fail_flag;
};
static __global__ void __launch_bounds__(32) mailbox_exchange_kernel(MboxArgs a) {
  const int q = threadIdx.x;
  const int n = a.n_sum + a.n_max;
  const int par = (int)(a.epoch & 1ull);
  if (q < a.world) {
    double* dst = a.peer[q] + ((size_t)par * a.world + a.rank) * kMboxSlots;
    for (int i = 0; i < n; ++i) dst[i] = a.payload[i];
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.seq_peer[q] + a.rank), "l"(a.epoch) : "memory");
    // wait for source q
    const unsigned long long* sq = a.seq_local + q;
    long long start = 0; unsigned spins = 0; bool ok = true;
    while (true) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(sq) : "memory");
      if (v >= a.epoch) break;
      if ((++spins & 255u) == 0) {
        const long long now = clock64();
        if (start == 0) start = now;
        else if (now - start > 40000000000LL) { ok = false; break; }     // ~20 s: a peer died; fail instead of hanging
      }
    }
    if (!ok) *a.fail_flag = 1;
  }
  __syncwarp();
  if (q == 0) {
    // lane 0 re-acquires every sequence word itself (they are all up): its reads below are ordered after the peers' writes
    for (int r = 0; r < a.world; ++r) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.seq_local + r) : "memory");
      (void)v;
    }
    const double* rows = a.local + (size_t)par * a.world * kMboxSlots;
    for (int i = 0; i < n; ++i) {
      double v = __ldcv(rows + i);
      for (int r = 1; r < a.world; ++r) {
        const double w = __ldcv(rows + (size_t)r * kMboxSlots + i);
        v = i < a.n_sum ? v + w : fmax(v, w);
      }
      a.out[i] = v;
    }
  }
}

static int mailbox_setup(pxr_ctx* ctx) {
  if (ctx->mbox_ready) return PXR_OK;
  const int world = ctx->world;
  if (world > 16) return fail(PXR_ERR_UNSUPPORTED, "peer mailboxes support up to 16 ranks per node");
  const size_t nb = (size_t)2 * world * kMboxSlots * sizeof(double), ns = (size_t)world * sizeof(unsigned long long);
  PXR_CUDA(cudaMalloc((void**)&ctx->mbox_local, nb));
  PXR_CUDA(cudaMalloc((void**)&ctx->mbox_seq_local, ns));
  PXR_CUDA(cudaMemset(ctx->mbox_local, 0, nb));
  PXR_CUDA(cudaMemset(ctx->mbox_seq_local, 0, ns));
  PXR_CUDA(cudaDeviceSynchronize());
  for (int r = 0; r < 16; ++r) { ctx->mbox_peer[r] = nullptr; ctx->mbox_seq_peer[r] = nullptr; }
  ctx->mbox_peer[ctx->rank] = ctx->mbox_local; ctx->mbox_seq_peer[ctx->rank] = ctx->mbox_seq_local;
  if (world > 1) {
    // exchange the IPC handles of the two buffers through NCCL itself (no extra host-side plumbing)
    struct Handles { cudaIpcMemHandle_t box, seq; int device; int pad[3]; };
    Handles mine;
    std::memset(&mine, 0, sizeof(mine));
    PXR_CUDA(cudaIpcGetMemHandle(&mine.box, ctx->mbox_local));
    PXR_CUDA(cudaIpcGetMemHandle(&mine.seq, ctx->mbox_seq_local));
    mine.device = ctx->device;
    DevBuf<uint8_t> d_send, d_recv;
    PXR_TRY(d_send.alloc(sizeof(Handles))); PXR_TRY(d_recv.alloc(sizeof(Handles) * world));
    PXR_CUDA(cudaMemcpyAsync(d_send.p, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
    PXR_TRY(allgather_bytes(ctx, d_send.p, d_recv.p, sizeof(Handles)));
    std::vector<Handles> all(world);
    PXR_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, sizeof(Handles) * world, cudaMemcpyDeviceToHost, ctx->stream));
    PXR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < world; ++r) {
      if (r == ctx->rank) continue;
      int can = 0;
      PXR_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, all[r].device));
      if (!can) return fail(PXR_ERR_UNSUPPORTED, "device %d cannot access device %d: peer mailboxes need P2P", ctx->device, all[r].device);
      void* pb = nullptr; void* ps = nullptr;
      PXR_CUDA(cudaIpcOpenMemHandle(&pb, all[r].box, cudaIpcMemLazyEnablePeerAccess));
      PXR_CUDA(cudaIpcOpenMemHandle(&ps, all[r].seq, cudaIpcMemLazyEnablePeerAccess));
      ctx->mbox_peer[r] = (double*)pb; ctx->mbox_seq_peer[r] = (unsigned long long*)ps;
    }
  }
  ctx->mbox_ready = true;
  return PXR_OK;
}

int mailbox_exchange(pxr_ctx* ctx, const double* payload, int n_sum, int n_max, double* out, int* fail_flag) {
  const int n = n_sum + n_max;
  if (n <= 0 || n > kMboxSlots) return fail(PXR_ERR_INTERNAL, "mailbox payload of %d slots", n);
  if (ctx->world <= 1) {
    if (out != payload) PXR_CUDA(cudaMemcpyAsync(out, payload, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return PXR_OK;
  }
  if (!ctx->mbox_ready) return fail(PXR_ERR_INTERNAL, "peer mailboxes are not set up");
  MboxArgs a;
  for (int r = 0; r < 16; ++r) { a.peer[r] = ctx->mbox_peer[r]; a.seq_peer[r] = ctx->mbox_seq_peer[r]; }
  a.local = ctx->mbox_local; a.seq_local = ctx->mbox_seq_local; a.payload = payload; a.out = out;
  a.rank = ctx->rank; a.world = ctx->world; a.n_sum = n_sum; a.n_max = n_max; a.epoch = ++ctx->mbox_epoch; a.fail_flag = fail_flag;
  PXR_LAUNCH(ctx, mailbox_exchange_kernel, 1, 32, 0, a);
  PXR_CUDA(cudaGetLastError());
  return PXR_OK;
}

}  // namespace pxr

using namespace pxr;

// ---- interrupt callback (process-wide; see include/pxr.h)
static std::mutex g_interrupt_mutex;
static pxr_interrupt_fn g_interrupt_fn = nullptr;
static void* g_interrupt_user = nullptr;
static std::chrono::steady_clock::time_point g_interrupt_last;

static int call_interrupt_callback() {
  pxr_interrupt_fn fn; void* user;
  { std::lock_guard<std::mutex> lock(g_interrupt_mutex); fn = g_interrupt_fn; user = g_interrupt_user; }
  return fn ? (fn(user) != 0) : 0;
}

namespace pxr {
bool interrupt_pending() {
  const auto now = std::chrono::steady_clock::now();
  {
    std::lock_guard<std::mutex> lock(g_interrupt_mutex);
    if (!g_interrupt_fn || now - g_interrupt_last < std::chrono::milliseconds(200)) return false;
    g_interrupt_last = now;
  }
  return call_interrupt_callback() != 0;
}
}  // namespace pxr

extern "C" {

int pxr_set_interrupt_callback(pxr_interrupt_fn fn, void* user) {
  std::lock_guard<std::mutex> lock(g_interrupt_mutex);
  g_interrupt_fn = fn; g_interrupt_user = user;
  g_interrupt_last = std::chrono::steady_clock::time_point();
  return PXR_OK;
}
int pxr_poll_interrupt(void) { return call_interrupt_callback(); }

const char* pxr_last_error(void) { return g_error.c_str(); }
int pxr_version(void) { return PXR_VERSION_MAJOR * 100 + PXR_VERSION_MINOR; }

int pxr_ctx_create(int device, pxr_ctx** out) {
  if (!out) return fail(PXR_ERR_INVALID_ARGUMENT, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(PXR_ERR_NO_DEVICE, "no CUDA device available (%s); libpxr has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= count) return fail(PXR_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", device, count);
  PXR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PXR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 9)
    return fail(PXR_ERR_NO_DEVICE, "device %d is sm_%d%d; libpxr is built for sm_100a only", device, prop.major, prop.minor);
  pxr_ctx* c = new pxr_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  PXR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  *out = c;
  return PXR_OK;
}

int pxr_ctx_destroy(pxr_ctx* ctx) {
  if (!ctx) return PXR_OK;
  cudaSetDevice(ctx->device);
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
  pxr::stager_destroy(ctx);
  if (ctx->slab_cache) cudaFree(ctx->slab_cache);
  for (int r = 0; r < 16; ++r)
    if (r != ctx->rank) { if (ctx->mbox_peer[r]) cudaIpcCloseMemHandle(ctx->mbox_peer[r]); if (ctx->mbox_seq_peer[r]) cudaIpcCloseMemHandle(ctx->mbox_seq_peer[r]); }
  if (ctx->mbox_local) cudaFree(ctx->mbox_local);
  if (ctx->mbox_seq_local) cudaFree(ctx->mbox_seq_local);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PXR_OK;
}

int pxr_nccl_unique_id(void* id128) {
  PXR_TRY(load_nccl());
  const int rc = g_nccl.GetUniqueId(id128);
  if (rc != 0) return fail(PXR_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  return PXR_OK;
}

int pxr_ctx_init_comm(pxr_ctx* ctx, int rank, int world, const void* id128) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(PXR_ERR_INVALID_ARGUMENT, "bad comm arguments");
  ctx->rank = rank; ctx->world = world;
  if (world == 1) return PXR_OK;
  PXR_TRY(load_nccl());
  PXR_CUDA(cudaSetDevice(ctx->device));
  Uid128 uid;
  std::memcpy(uid.b, id128, 128);
  typedef int (*init_t)(void**, int, Uid128, int);
  const int rc = ((init_t)g_nccl.InitRank)(&ctx->nccl_comm, world, uid, rank);
  if (rc != 0) return fail(PXR_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  // peer mailboxes for the per-iteration scalar exchange (PXR_NO_MAILBOX=1: keep every exchange on NCCL)
  if (!getenv("PXR_NO_MAILBOX")) {
    const int mrc = mailbox_setup(ctx);
    if (mrc != PXR_OK) {
      // every rank of a node sees the same P2P topology, so all of them fall back together
      ctx->mbox_ready = false;
    }
  }
  return PXR_OK;
}
int64_t pxr_ctx_nccl_collectives(pxr_ctx* ctx) { return ctx ? ctx->nccl_collectives : 0; }
int pxr_ctx_mailbox_ready(pxr_ctx* ctx) { return ctx && ctx->mbox_ready ? 1 : 0; }

int pxr_ctx_sync(pxr_ctx* ctx) {
  if (!ctx) return fail(PXR_ERR_INVALID_ARGUMENT, "ctx is NULL");
  PXR_CUDA(cudaStreamSynchronize(ctx->stream));
  return PXR_OK;
}
int pxr_ctx_timer_start(pxr_ctx* ctx) {
  if (!ctx) return fail(PXR_ERR_INVALID_ARGUMENT, "ctx is NULL");
  PXR_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->ev0) { PXR_CUDA(cudaEventCreate(&ctx->ev0)); PXR_CUDA(cudaEventCreate(&ctx->ev1)); }
  PXR_CUDA(cudaStreamSynchronize(ctx->stream));
  PXR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return PXR_OK;
}
int pxr_ctx_timer_stop(pxr_ctx* ctx, double* elapsed_ms) {
  if (!ctx || !ctx->ev0 || !elapsed_ms) return fail(PXR_ERR_INVALID_ARGUMENT, "timer not started");
  PXR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  PXR_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  PXR_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *elapsed_ms = ms;
  return PXR_OK;
}
int64_t pxr_ctx_kernel_launches(pxr_ctx* ctx) { return ctx ? ctx->launches : 0; }

void pxr_default_interp_config(pxr_interp_config* c) {
  c->l2_normalize = 1; c->use_float_simd = 0; c->check_bounds = 0; c->reserved = 0;
}
void pxr_default_ba_options(pxr_solver_options* o) {
  o->loss_type = PXR_LOSS_CAUCHY; o->loss_scale = 0.25; o->linear_solver = PXR_SOLVER_AUTO;
  o->max_num_iterations = 100; o->max_linear_solver_iterations = 200;
  o->max_num_consecutive_invalid_steps = 10;
  o->function_tolerance = 0.0; o->gradient_tolerance = 0.0; o->parameter_tolerance = 0.0;
  o->use_inner_iterations = 1; o->inner_iteration_tolerance = 1e-3;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1; o->deterministic = 0;
  o->use_nonmonotonic_steps = 0; o->max_consecutive_nonmonotonic_steps = 5;
}
void pxr_default_ka_options(pxr_solver_options* o) {
  pxr_default_ba_options(o);
  o->use_inner_iterations = 0;
  o->parameter_tolerance = 1e-5;
}

int pxr_device_free(pxr_ctx* ctx, void* dptr) {
  if (ctx) cudaSetDevice(ctx->device);
  if (dptr) PXR_CUDA(cudaFree(dptr));
  return PXR_OK;
}
int pxr_host_alloc_pinned(void** out, size_t bytes) {
  if (!out) return fail(PXR_ERR_INVALID_ARGUMENT, "out is NULL");
  int device = 0;
  PXR_CUDA(cudaGetDevice(&device));
  pxr::NumaLocalScope numa(device);      // pages on the NUMA node of the current device (see pxr_internal.h)
  PXR_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return PXR_OK;
}
int pxr_host_free_pinned(void* p) {
  if (p) PXR_CUDA(cudaFreeHost(p));
  return PXR_OK;
}
int pxr_memcpy_d2h(pxr_ctx* ctx, void* host, const void* dev, size_t bytes) {
  if (!ctx) return fail(PXR_ERR_INVALID_ARGUMENT, "ctx is NULL");
  PXR_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PXR_CUDA(cudaStreamSynchronize(ctx->stream));
  return PXR_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-side integer algorithms (bit-exact targets).  Product code: written against the
// reference's behaviour (base/src/graph.cc:126-256), independent of oracle/.
static int64_t uf_root(int64_t i, std::vector<int64_t>& parent) {
  int64_t r = i;
  while (parent[r] != -1) r = parent[r];
  while (parent[i] != -1) { const int64_t nx = parent[i]; parent[i] = r; i = nx; }  // path compression
  return r;
}

int pxr_graph_track_labels(int64_t n_nodes, const int32_t* node_image, int64_t n_edges, const int64_t* e_src,
                           const int64_t* e_dst, const double* e_sim, int64_t* out) {
  if (n_nodes < 0 || n_edges < 0 || (n_nodes && (!node_image || !out)) || (n_edges && (!e_src || !e_dst)))
    return fail(PXR_ERR_INVALID_ARGUMENT, "bad graph arguments");
  // edges in DESCENDING lexicographic (sim, node1, node2) order == sort ascending + reverse
  std::vector<int64_t> order(n_edges);
  for (int64_t e = 0; e < n_edges; ++e) order[e] = e;
  auto sim_of = [&](int64_t e) { return e_sim ? e_sim[e] : 1.0; };
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return std::make_tuple(sim_of(a), (uint64_t)e_src[a], (uint64_t)e_dst[a]) >
           std::make_tuple(sim_of(b), (uint64_t)e_src[b], (uint64_t)e_dst[b]);
  });
  std::vector<int64_t> parent(n_nodes, -1);
  // image sets as sorted vectors (merged on union)
  std::vector<std::vector<int32_t>> images(n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i) images[i].push_back(node_image[i]);
  std::vector<int32_t> merged;
  for (int64_t e : order) {
    const int64_t a = e_src[e], b = e_dst[e];
    if (a < 0 || b < 0 || a >= n_nodes || b >= n_nodes) return fail(PXR_ERR_INVALID_ARGUMENT, "edge endpoint out of range");
    const int64_t r1 = uf_root(a, parent), r2 = uf_root(b, parent);
    if (r1 == r2) continue;
    const std::vector<int32_t>& s1 = images[r1];
    const std::vector<int32_t>& s2 = images[r2];
    bool intersects = false;
    for (size_t i = 0, j = 0; i < s1.size() && j < s2.size();) {
      if (s1[i] < s2[j]) ++i; else if (s2[j] < s1[i]) ++j; else { intersects = true; break; }
    }
    if (intersects) continue;
    merged.resize(s1.size() + s2.size());
    std::merge(s1.begin(), s1.end(), s2.begin(), s2.end(), merged.begin());
    if (s1.size() < s2.size()) { parent[r1] = r2; images[r2] = merged; images[r1].clear(); images[r1].shrink_to_fit(); }
    else { parent[r2] = r1; images[r1] = merged; images[r2].clear(); images[r2].shrink_to_fit(); }
  }
  int64_t n_tracks = 0;
  for (int64_t i = 0; i < n_nodes; ++i) out[i] = parent[i] == -1 ? n_tracks++ : -1;
  for (int64_t i = 0; i < n_nodes; ++i) if (out[i] == -1) out[i] = out[uf_root(i, parent)];
  return PXR_OK;
}

int pxr_graph_score_labels(int64_t n_nodes, int64_t n_edges, const int64_t* e_src, const int64_t* e_dst,
                           const double* e_sim, const int64_t* labels, double* scores) {
  for (int64_t i = 0; i < n_nodes; ++i) scores[i] = 0.0;
  for (int64_t e = 0; e < n_edges; ++e) {
    const double s = e_sim ? e_sim[e] : 1.0;
    if (labels[e_src[e]] == labels[e_dst[e]]) { scores[e_src[e]] += s; scores[e_dst[e]] += s; }
  }
  return PXR_OK;
}

int pxr_graph_root_labels(int64_t n_nodes, const int64_t* labels, const double* scores, uint8_t* is_root) {
  int64_t n_tracks = 0;
  for (int64_t i = 0; i < n_nodes; ++i) n_tracks = std::max(n_tracks, labels[i] + 1);
  std::vector<int64_t> order(n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return std::make_pair(scores[a], (uint64_t)a) > std::make_pair(scores[b], (uint64_t)b);
  });
  std::vector<char> has_root(n_tracks, 0);
  for (int64_t i = 0; i < n_nodes; ++i) is_root[i] = 0;
  for (int64_t n : order) {
    if (has_root[labels[n]]) continue;
    is_root[n] = 1; has_root[labels[n]] = 1;
  }
  return PXR_OK;
}

// keypoint_adjustment/main.py:13-57 (first-fit-decreasing; Counter.most_common order = count
// descending, ties by first appearance)
int pxr_ka_problem_labels(int64_t n_nodes, const int64_t* labels, int32_t max_per_problem, int32_t* out,
                          int32_t* n_problems_out) {
  std::vector<int64_t> order;
  std::unordered_map<int64_t, int64_t> count;
  for (int64_t i = 0; i < n_nodes; ++i) {
    auto it = count.find(labels[i]);
    if (it == count.end()) { count.emplace(labels[i], 1); order.push_back(labels[i]); } else ++it->second;
  }
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return count[a] > count[b]; });
  int64_t maxpp = max_per_problem;
  if (maxpp == -1) for (auto& kv : count) maxpp = std::max(maxpp, kv.second);
  std::vector<int64_t> bins;
  std::unordered_map<int64_t, int32_t> t2p;
  size_t start = 0;
  int64_t last_v = std::numeric_limits<int64_t>::max();
  for (int64_t k : order) {
    const int64_t v = count[k];
    if (v < last_v) { start = 0; last_v = v; }
    bool found = false;
    if (v < maxpp) {
      for (size_t i = start; i < bins.size(); ++i)
        if (bins[i] + v <= maxpp) { bins[i] += v; t2p[k] = (int32_t)i; found = true; start = i; break; }
    }
    if (!found) { t2p[k] = (int32_t)bins.size(); start = bins.size(); bins.push_back(v); }
  }
  for (int64_t i = 0; i < n_nodes; ++i) out[i] = t2p[labels[i]];
  if (n_problems_out) *n_problems_out = (int32_t)bins.size();
  return PXR_OK;
}

// Contiguous point ranges balanced by observation count.
int pxr_shard_points(int64_t n_points, int64_t n_obs, const int64_t* obs_pt, int world, int64_t* point_begin,
                     int64_t* obs_begin) {
  if (world < 1 || n_points < 0 || n_obs < 0) return fail(PXR_ERR_INVALID_ARGUMENT, "bad shard arguments");
  std::vector<int64_t> pt_begin(n_points + 1, 0);
  for (int64_t o = 0; o < n_obs; ++o) {
    if (obs_pt[o] < 0 || obs_pt[o] >= n_points) return fail(PXR_ERR_INVALID_ARGUMENT, "obs_pt out of range");
    if (o && obs_pt[o] < obs_pt[o - 1]) return fail(PXR_ERR_INVALID_ARGUMENT, "observations must be sorted by point");
    pt_begin[obs_pt[o] + 1]++;
  }
  for (int64_t p = 0; p < n_points; ++p) pt_begin[p + 1] += pt_begin[p];
  point_begin[0] = 0; obs_begin[0] = 0;
  int64_t p = 0;
  for (int r = 1; r < world; ++r) {
    const int64_t target = (n_obs * r) / world;
    while (p < n_points && pt_begin[p] < target) ++p;
    point_begin[r] = p; obs_begin[r] = pt_begin[p];
  }
  point_begin[world] = n_points; obs_begin[world] = n_obs;
  return PXR_OK;
}

int pxr_ba_estimate_device_bytes(const pxr_ba_desc* d, const pxr_solver_options* opt, double* patch_bytes,
                                 double* state_bytes, double* reduced_bytes) {
  if (!d) return fail(PXR_ERR_INVALID_ARGUMENT, "desc is NULL");
  if (d->n_obs < 0 || d->n_points < 0 || d->n_images < 0 || d->n_cameras < 0 || (d->n_obs > 0 && !d->obs_pt))
    return fail(PXR_ERR_INVALID_ARGUMENT, "bad problem sizes");
  const double esz = d->patch_dtype == PXR_F16 ? 2 : (d->patch_dtype == PXR_F32 ? 4 : 8);
  double n_patches = (double)(d->obs_patch ? d->n_patches : std::max(d->n_patches, d->n_obs));
  if (d->n_patch_blocks > 0 && d->patch_block_counts) {
    n_patches = 0;
    for (int b = 0; b < d->n_patch_blocks; ++b) n_patches += (double)d->patch_block_counts[b];
  }
  const double slab = d->patches_on_device ? 0.0 : n_patches * d->ph * d->pw * d->channels * esz;
  int K = 0;
  for (int i = 0; i < d->n_cameras; ++i) K = std::max(K, d->cam_model ? cam_num_params(d->cam_model[i]) : 0);
  const double dcmax = 6 + K, juv = 2 * (9 + K), n_obs = (double)d->n_obs, n_pts = (double)d->n_points;
  // observation pairs (i >= j) of every point: the static structure of the Schur complement
  double pairs = 0;
  for (int64_t o = 0, run = 0; o < d->n_obs; ++o) {
    run = (o > 0 && d->obs_pt[o] == d->obs_pt[o - 1]) ? run + 1 : 1;
    pairs += (double)run;
  }
  const double per_obs = 2.0 * (2 + 8 + juv) * 8       // uv, obs_out, juv: the linearisation and the trial point
                         + 2.0 * dcmax * 3 * 8          // W and T = W (Hpp + D)^-1
                         + dcmax * 4 + 4                // column table
                         + 4 + 8 + 8;                   // obs_img, obs_pt, obs_patch
  const double state = n_obs * per_obs + pairs * 8 + n_pts * ((9 + 3 + 6) * 8 + 2 * 3 * 8 + 8 + 8 + 1)
                       + n_patches * (8 + 16) + (d->refs ? n_pts * d->channels * 8 : 0.0);
  // upper bound of the camera unknowns: 6 per image + K per camera (constant blocks only make it smaller)
  const double nc = 6.0 * d->n_images + (double)K * d->n_cameras;
  const int solver = opt ? opt->linear_solver : PXR_SOLVER_AUTO;
  const bool iterative = solver == PXR_SOLVER_ITERATIVE_SCHUR || (solver == PXR_SOLVER_AUTO && d->n_images > 1000);
  double reduced = 2.0 * nc * nc * 8;                  // Hcc and [S; rhs]
  if (iterative && nc * nc * 8 > 4e9) reduced = ((double)d->n_images + pairs / 16.0) * 64 * 8;   // image blocks + co-visible pairs (estimate)
  if (patch_bytes) *patch_bytes = slab;
  if (state_bytes) *state_bytes = state;
  if (reduced_bytes) *reduced_bytes = reduced;
  return PXR_OK;
}

int pxr_shard_ka_problems(int32_t n_problems, const int64_t* weight, int world, int32_t* rank_of_problem) {
  if (world < 1 || n_problems < 0 || (n_problems > 0 && (!weight || !rank_of_problem)))
    return fail(PXR_ERR_INVALID_ARGUMENT, "bad shard arguments");
  // longest-processing-time-first: heaviest problem to the least loaded rank; ties by lower index / lower rank,
  // so every rank computes the same plan
  std::vector<int32_t> order(n_problems);
  for (int32_t i = 0; i < n_problems; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return weight[a] > weight[b]; });
  std::vector<int64_t> load(world, 0);
  for (int32_t i : order) {
    int best = 0;
    for (int r = 1; r < world; ++r) if (load[r] < load[best]) best = r;
    rank_of_problem[i] = best;
    load[best] += std::max<int64_t>(weight[i], 0);
  }
  return PXR_OK;
}

}  // extern "C"
