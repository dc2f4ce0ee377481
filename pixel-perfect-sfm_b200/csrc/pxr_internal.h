// pxr_internal.h — host-side plumbing shared by the translation units of libpxr.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pxr.h"

namespace pxr { struct Stager; }

struct pxr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  int sm_count = 148;
  // multi-GPU (NCCL resolved with dlopen at pxr_ctx_init_comm)
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // peer mailboxes (pxr_api.cu: mailbox_*): small scalar exchanges between the ranks of one node go through peer memory
  // over NVLink (cudaIpc-mapped buffers), not through NCCL, so an LM iteration issues exactly ONE NCCL collective
  double* mbox_local = nullptr;            // [2 parities][world][kMboxSlots] payloads written BY the peers
  unsigned long long* mbox_seq_local = nullptr;   // [world] sequence number of the last exchange each peer has delivered
  double* mbox_peer[16] = {};              // peer r's mbox_local as mapped into this process (own entry = mbox_local)
  unsigned long long* mbox_seq_peer[16] = {};
  unsigned long long mbox_epoch = 0;       // exchanges issued so far (same on every rank)
  bool mbox_ready = false;
  int64_t nccl_collectives = 0;            // NCCL calls issued through this context (bench.py reports them per LM iteration)
  pxr::Stager* stager = nullptr;          // pinned ring of the pageable-memory upload pipeline (pxr_upload.cu)
  cudaStream_t upload_stream = nullptr;   // carries the patch slab so that host-side setup (and its small syncs) overlaps it
  // one cached patch slab: cudaMalloc / cudaFree of tens of GB cost 0.1-0.5 s each, a multi-level refinement (or a
  // reference extraction followed by the adjustment) would pay them per call; released with the context
  void* slab_cache = nullptr;
  size_t slab_cache_bytes = 0;
};

namespace pxr {

void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define PXR_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return pxr::fail(PXR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define PXR_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != PXR_OK) return _s; \
  } while (0)

#define PXR_LAUNCH(ctx, kernel, grid, block, smem, ...)          \
  do {                                                           \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                           \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return PXR_OK;
    PXR_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    return PXR_OK;
  }
  int upload(const T* host, size_t count, cudaStream_t s) {
    PXR_TRY(alloc(count));
    if (count) PXR_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    return PXR_OK;
  }
  int zero(cudaStream_t s) {
    if (n) PXR_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    return PXR_OK;
  }
};

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// The patch slab of one optimizer: taken from / returned to the context's one-entry cache when it is large.
struct PatchSlab {
  pxr_ctx* ctx = nullptr;
  uint8_t* p = nullptr;
  size_t n = 0;
  static constexpr size_t kCacheFrom = (size_t)256 << 20;
  PatchSlab() {}
  PatchSlab(const PatchSlab&) = delete;
  PatchSlab& operator=(const PatchSlab&) = delete;
  ~PatchSlab() { release(); }
  int alloc(pxr_ctx* c, size_t bytes) {
    release();
    ctx = c; n = bytes;
    if (bytes == 0) return PXR_OK;
    if (c->slab_cache && c->slab_cache_bytes >= bytes) {
      p = (uint8_t*)c->slab_cache; n = c->slab_cache_bytes; c->slab_cache = nullptr; c->slab_cache_bytes = 0;
      return PXR_OK;
    }
    if (c->slab_cache && bytes >= kCacheFrom) { cudaFree(c->slab_cache); c->slab_cache = nullptr; c->slab_cache_bytes = 0; }
    PXR_CUDA(cudaMalloc((void**)&p, bytes));
    return PXR_OK;
  }
  void release() {
    if (!p) return;
    if (ctx && n >= kCacheFrom && n > ctx->slab_cache_bytes) {
      if (ctx->slab_cache) cudaFree(ctx->slab_cache);
      ctx->slab_cache = p; ctx->slab_cache_bytes = n;
    } else {
      cudaFree(p);
    }
    p = nullptr; n = 0;
  }
};

// large input upload on ctx->stream: any source (pinned, pageable, device); pageable sources are pipelined through a
// pinned ring (pxr_upload.cu).  *h2d_bytes (optional) is incremented by the bytes that crossed PCIe.
// `stream` (default: ctx->stream) is the stream the copies are ordered on.
int upload_bytes(pxr_ctx* ctx, void* dst, const void* src, size_t bytes, double* h2d_bytes = nullptr, cudaStream_t stream = nullptr);
int upload_segments(pxr_ctx* ctx, void* dst, const void* const* srcs, const size_t* sizes, int n, double* h2d_bytes = nullptr,
                    cudaStream_t stream = nullptr);
// window residency (pxr_resident.cuh): the rectangle h_rect[p] of every patch -> its place in the full-layout slab, packed
// on the host, moved by DMA, scattered on the device; pinned or pageable sources given as blocks of whole patches
int upload_windows(pxr_ctx* ctx, uint8_t* slab, const void* const* srcs, const int64_t* block_first, int n_blocks,
                   const uint32_t* h_rect, const uint32_t* d_rect, int64_t n_patches, int ph, int pw, int tap_bytes,
                   double* h2d_bytes, cudaStream_t stream = nullptr);
size_t staged_chunk_bytes();     // bytes of one staging buffer (PXR_STAGED_CHUNK, default 4 MB): a window / patch must fit
// the context's side stream for the patch slab (created on first use)
int upload_stream(pxr_ctx* ctx, cudaStream_t* out);
void stager_destroy(pxr_ctx* ctx);

// Host memory the GPU reads over PCIe should sit on the GPU's NUMA node (a copy that crosses the socket interconnect
// runs at 30 instead of 50 GB/s on the 2-socket B200 hosts).  While alive, the calling thread is bound to the CPUs
// local to `device` (sysfs local_cpulist of its PCI function, intersected with the CPUs the process may use) and prefers
// that node for new pages; the destructor restores both.  Does nothing when the topology cannot be read.
struct NumaLocalScope {
  explicit NumaLocalScope(int device);
  ~NumaLocalScope();
  NumaLocalScope(const NumaLocalScope&) = delete;
  NumaLocalScope& operator=(const NumaLocalScope&) = delete;
  bool bound = false;
 private:
  unsigned long old_mask_[16] = {};
  bool have_old_ = false, policy_set_ = false;
};

// true when the host's interrupt callback (pxr_set_interrupt_callback) asks to stop; rate-limited to one call per 200 ms
bool interrupt_pending();

// allreduce (sum, fp64) on ctx->stream when a communicator is attached; no-op otherwise
int allreduce_f64(pxr_ctx* ctx, double* dptr, size_t count, bool max_op = false);
// out-of-place sum (send untouched: a rejected LM step re-reduces the same local blocks with new damping)
int allreduce_f64_oop(pxr_ctx* ctx, const double* send, double* recv, size_t count);
// every rank contributes `bytes` bytes (device memory); recv holds world * bytes, rank-major.  Setup-time use only.
int allgather_bytes(pxr_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes);
// Scalar exchange over peer memory: out[i] = sum over ranks (fixed rank order: bitwise identical everywhere) of
// payload[i] for i < n_sum, and max over ranks for n_sum <= i < n_sum + n_max.  payload/out are device pointers,
// n_sum + n_max <= kMboxSlots.  world == 1: a device copy.  Stream-ordered on ctx->stream; *fail_flag (device int)
// is raised if a peer does not show up within the spin budget (no hang).
constexpr int kMboxSlots = 16;
int mailbox_exchange(pxr_ctx* ctx, const double* payload, int n_sum, int n_max, double* out, int* fail_flag);

}  // namespace pxr
