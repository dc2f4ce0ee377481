// pxr_upload.cu — host -> device transfer of the feature patches, the only large input of the path.
//
// The reference keeps its feature maps in pageable host memory (numpy arrays behind FeatureMap / FeaturePatch,
// features/src/featurepatch.h:35-70).  cudaMemcpy from pageable memory goes through the driver's single staging
// buffer (~11 GB/s measured on the B200 box, profiles/README.md "Keypoint adjustment"); PCIe Gen5 x16 carries
// ~50 GB/s from pinned memory.  upload_bytes() therefore
//   * forwards pinned / registered / device / managed sources to ONE cudaMemcpyAsync on the context stream, and
//   * pipelines large PAGEABLE sources itself: kThreads host threads each copy 4 MB chunks into their own pinned
//     double buffer and issue the H2D copy on their own stream, so the host memcpy of one chunk overlaps the DMA of
//     the others; the context stream then waits on every worker stream.
// On return the source has been read completely (same contract as cudaMemcpyAsync from pageable memory).
// Switches: PXR_STAGED_UPLOAD=0 disables the pipeline, PXR_STAGED_UPLOAD_MIN / PXR_STAGED_CHUNK (bytes) override the
// 16 MB threshold and the 4 MB chunk (used by the tests to push small inputs through it).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "pxr_internal.h"

namespace pxr {

// ------------------------------------------------------------------------------------------------ NUMA placement
static bool read_small_file(const char* path, char* buf, size_t cap) {
  FILE* f = std::fopen(path, "r");
  if (!f) return false;
  const size_t n = std::fread(buf, 1, cap - 1, f);
  std::fclose(f);
  buf[n] = 0;
  return n > 0;
}

NumaLocalScope::NumaLocalScope(int device) {
  static const bool off = []() { const char* v = std::getenv("PXR_NUMA_LOCAL"); return v && v[0] == '0'; }();
  if (off) return;
  char bus[32] = {};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return; }
  for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
  char path[160], buf[4096];
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  if (!read_small_file(path, buf, sizeof(buf))) return;
  cpu_set_t local; CPU_ZERO(&local);
  for (const char* c = buf; *c && *c != '\n';) {        // "0-31,64-95"
    char* e; long a = std::strtol(c, &e, 10); if (e == c) break;
    long b = a; c = e;
    if (*c == '-') { b = std::strtol(c + 1, &e, 10); c = e; }
    for (long k = a; k <= b && k < CPU_SETSIZE; ++k) CPU_SET((int)k, &local);
    if (*c == ',') ++c;
  }
  cpu_set_t old; CPU_ZERO(&old);
  if (sched_getaffinity(0, sizeof(old), &old) != 0) return;
  cpu_set_t both; CPU_AND(&both, &old, &local);
  if (CPU_COUNT(&both) == 0) return;                     // the process may not run there (cpuset): leave it alone
  static_assert(sizeof(cpu_set_t) <= sizeof(old_mask_), "cpu_set_t larger than expected");
  std::memcpy(old_mask_, &old, sizeof(old));
  if (sched_setaffinity(0, sizeof(both), &both) != 0) return;
  have_old_ = true; bound = true;
  // prefer the node for pages faulted / allocated by this thread (MPOL_PREFERRED = 1); raw syscall: no libnuma here
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  if (read_small_file(path, buf, sizeof(buf))) {
    const long node = std::strtol(buf, nullptr, 10);
    if (node >= 0 && node < 64) {
      unsigned long mask = 1ul << node;
      if (syscall(SYS_set_mempolicy, 1 /*MPOL_PREFERRED*/, &mask, 65ul) == 0) policy_set_ = true;
    }
  }
}

NumaLocalScope::~NumaLocalScope() {
  if (policy_set_) syscall(SYS_set_mempolicy, 0 /*MPOL_DEFAULT*/, nullptr, 0ul);
  if (have_old_) { cpu_set_t old; std::memcpy(&old, old_mask_, sizeof(old)); sched_setaffinity(0, sizeof(old), &old); }
}

struct Stager {
  static constexpr int kThreads = 16, kBufs = 2;     // 6 threads carry the plain slab upload, up to 16 the window packing
  static constexpr int kPlainThreads = 6;
  size_t chunk = 0;
  uint8_t* pin[kThreads][kBufs] = {};
  uint8_t* dev[kThreads][kBufs] = {};                // device side of the window upload (packed windows before the scatter)
  cudaStream_t st[kThreads] = {};
  cudaEvent_t ev[kThreads][kBufs] = {};
  cudaEvent_t start = nullptr;
  ~Stager() {
    for (int t = 0; t < kThreads; ++t) {
      for (int b = 0; b < kBufs; ++b) {
        if (ev[t][b]) cudaEventDestroy(ev[t][b]);
        if (pin[t][b]) cudaFreeHost(pin[t][b]);
        if (dev[t][b]) cudaFree(dev[t][b]);
      }
      if (st[t]) cudaStreamDestroy(st[t]);
    }
    if (start) cudaEventDestroy(start);
  }
};

static size_t env_bytes(const char* name, size_t dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  const long long x = std::atoll(v);
  return x > 0 ? (size_t)x : dflt;
}

size_t staged_chunk_bytes() { return env_bytes("PXR_STAGED_CHUNK", (size_t)4 << 20); }

static int stager_get(pxr_ctx* ctx, Stager** out) {
  const size_t chunk = env_bytes("PXR_STAGED_CHUNK", (size_t)4 << 20);
  if (ctx->stager && ctx->stager->chunk == chunk) { *out = ctx->stager; return PXR_OK; }
  delete ctx->stager;
  ctx->stager = nullptr;
  Stager* s = new Stager();
  s->chunk = chunk;
  NumaLocalScope numa(ctx->device);      // the ring the DMA engine reads from: on the GPU's node
  cudaError_t e = cudaEventCreateWithFlags(&s->start, cudaEventDisableTiming);
  for (int t = 0; t < Stager::kThreads && e == cudaSuccess; ++t) {
    e = cudaStreamCreateWithFlags(&s->st[t], cudaStreamNonBlocking);
    for (int b = 0; b < Stager::kBufs && e == cudaSuccess; ++b) {
      e = cudaHostAlloc((void**)&s->pin[t][b], chunk, cudaHostAllocDefault);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev[t][b], cudaEventDisableTiming);
    }
  }
  if (e != cudaSuccess) {
    delete s;
    return fail(PXR_ERR_CUDA, "staged upload: cannot create the pinned ring (%s)", cudaGetErrorString(e));
  }
  ctx->stager = s;
  *out = s;
  return PXR_OK;
}

void stager_destroy(pxr_ctx* ctx) {
  delete ctx->stager;
  ctx->stager = nullptr;
  if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
  ctx->upload_stream = nullptr;
}

static cudaMemoryType source_type(const void* src) {
  cudaPointerAttributes pa;
  cudaMemoryType type = cudaMemoryTypeUnregistered;
  if (cudaPointerGetAttributes(&pa, src) == cudaSuccess) type = pa.type;
  cudaGetLastError();   // plain host memory makes the query fail on old drivers: not an error here
  return type;
}

// dst <- srcs[0] | srcs[1] | ... (the per-image blocks of a feature set land in one slab).  h2d_bytes (optional) is
// incremented by the bytes that crossed PCIe.
int upload_segments(pxr_ctx* ctx, void* dst, const void* const* srcs, const size_t* sizes, int n, double* h2d_bytes,
                    cudaStream_t stream) {
  if (!stream) stream = ctx->stream;
  static const bool enabled = []() { const char* v = std::getenv("PXR_STAGED_UPLOAD"); return !(v && v[0] == '0'); }();
  const size_t min_bytes = env_bytes("PXR_STAGED_UPLOAD_MIN", (size_t)16 << 20);
  std::vector<size_t> begin((size_t)n + 1, 0);
  bool all_pageable = true;
  for (int i = 0; i < n; ++i) {
    begin[i + 1] = begin[i] + sizes[i];
    if (sizes[i] == 0) continue;
    const cudaMemoryType type = source_type(srcs[i]);
    if (type != cudaMemoryTypeUnregistered) all_pageable = false;
    if (h2d_bytes && type != cudaMemoryTypeDevice) *h2d_bytes += (double)sizes[i];
  }
  const size_t bytes = begin[n];
  if (bytes == 0) return PXR_OK;
  if (!all_pageable || !enabled || bytes < min_bytes) {
    // in pieces: small copies other streams make meanwhile (the problem tables) are not queued behind 32 GB
    const size_t piece = (size_t)256 << 20;
    for (int i = 0; i < n; ++i)
      for (size_t o = 0; o < sizes[i]; o += piece)
        PXR_CUDA(cudaMemcpyAsync((uint8_t*)dst + begin[i] + o, (const uint8_t*)srcs[i] + o, std::min(piece, sizes[i] - o),
                                 cudaMemcpyDefault, stream));
    return PXR_OK;
  }
  Stager* s = nullptr;
  PXR_TRY(stager_get(ctx, &s));
  const size_t chunk = s->chunk;
  const size_t n_chunks = (bytes + chunk - 1) / chunk;
  const int n_threads = (int)std::min<size_t>(Stager::kPlainThreads, n_chunks);
  // the destination may still be in use by earlier work on the consumer stream
  PXR_CUDA(cudaEventRecord(s->start, stream));
  cudaError_t errs[Stager::kThreads];
  const int device = ctx->device;
  auto work = [&](int t) {
    NumaLocalScope numa(device);         // the copy threads run next to the ring they fill
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->st[t], s->start, 0);
    size_t k = 0;
    for (size_t c = t; c < n_chunks && e == cudaSuccess; c += n_threads, ++k) {
      const int b = (int)(k % Stager::kBufs);
      const size_t off = c * chunk, len = std::min(chunk, bytes - off);
      e = cudaEventSynchronize(s->ev[t][b]);     // the previous DMA out of this buffer (this call or the last) is done
      if (e != cudaSuccess) break;
      // gather [off, off+len) of the concatenation: first segment with begin[i+1] > off
      size_t i = (size_t)(std::upper_bound(begin.begin(), begin.end(), off) - begin.begin()) - 1;
      for (size_t done = 0; done < len; ++i) {
        const size_t so = off + done - begin[i], take = std::min(len - done, sizes[i] - so);
        if (take) std::memcpy(s->pin[t][b] + done, (const uint8_t*)srcs[i] + so, take);
        done += take;
      }
      e = cudaMemcpyAsync((uint8_t*)dst + off, s->pin[t][b], len, cudaMemcpyHostToDevice, s->st[t]);
      if (e == cudaSuccess) e = cudaEventRecord(s->ev[t][b], s->st[t]);
    }
    errs[t] = e;
  };
  std::vector<std::thread> pool;
  int started = 1;
  try {
    for (int t = 1; t < n_threads; ++t) { pool.emplace_back(work, t); ++started; }
  } catch (const std::exception&) {
    // no more threads to be had (container limits): the remaining shares run on this thread
  }
  work(0);
  for (int t = started; t < n_threads; ++t) work(t);
  for (auto& th : pool) th.join();
  for (int t = 0; t < n_threads; ++t)
    if (errs[t] != cudaSuccess) {
      cudaDeviceSynchronize();
      return fail(PXR_ERR_CUDA, "staged upload failed: %s", cudaGetErrorString(errs[t]));
    }
  // make the consumer stream wait for the copies of every worker stream
  for (int t = 0; t < n_threads; ++t) {
    const size_t mine = (n_chunks - t + n_threads - 1) / n_threads;
    const int last = (int)((mine - 1) % Stager::kBufs);
    PXR_CUDA(cudaStreamWaitEvent(stream, s->ev[t][last], 0));   // worker streams are in order: the last event covers all
  }
  return PXR_OK;
}

// ------------------------------------------------------------------------------------------------ window upload
// Packed windows -> their places in the slab.  One CTA per patch of the chunk; the window of patch p sits at
// packed + (p - first) * win_bytes, row-major rows x (cols * tap_bytes).
static __global__ void __launch_bounds__(256) scatter_windows_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ slab,
                                                                     const uint32_t* __restrict__ rect, const int64_t* __restrict__ chunk_off,
                                                                     int64_t first, int count, int ph, int pw, int tap_bytes) {
  const int64_t patch_bytes = (int64_t)ph * pw * tap_bytes;
  for (int k = blockIdx.x; k < count; k += gridDim.x) {
    const int64_t p = first + k;
    const uint32_t rc = rect[p];
    const int r0 = (int)(rc & 255u), c0 = (int)((rc >> 8) & 255u), rows = (int)((rc >> 16) & 255u), cols = (int)(rc >> 24);
    const int row_words = cols * tap_bytes / 16;
    const uint8_t* src = packed + (chunk_off[p] - chunk_off[first]);
    uint8_t* dst = slab + p * patch_bytes;
    for (int w = threadIdx.x; w < rows * row_words; w += blockDim.x) {
      const int r = w / row_words, q = w - r * row_words;
      *reinterpret_cast<uint4*>(dst + ((int64_t)(r0 + r) * pw + c0) * tap_bytes + (int64_t)q * 16) =
          *reinterpret_cast<const uint4*>(src + (int64_t)w * 16);
    }
  }
}

// Brings the rectangle h_rect[p] (r0 | c0<<8 | rows<<16 | cols<<24) of every patch into the slab (full [n][ph][pw][tap]
// layout in device memory): host threads pack the rectangles of a chunk of patches into a pinned buffer (the only
// bytes of the source they touch), the DMA engine moves the packed chunk, a small kernel on the same stream scatters
// it into place.  Works for pinned and pageable sources alike; `srcs` as in upload_segments (blocks of whole patches).
int upload_windows(pxr_ctx* ctx, uint8_t* slab, const void* const* srcs, const int64_t* block_first, int n_blocks,
                   const uint32_t* h_rect, const uint32_t* d_rect, int64_t n_patches, int ph, int pw, int tap_bytes,
                   double* h2d_bytes, cudaStream_t stream) {
  if (!stream) stream = ctx->stream;
  if (n_patches <= 0) return PXR_OK;
  Stager* s = nullptr;
  PXR_TRY(stager_get(ctx, &s));
  const size_t chunk = s->chunk;
  // packed offsets of the patches and the chunk boundaries (chunks of whole patches, at most `chunk` bytes)
  std::vector<int64_t> off((size_t)n_patches + 1, 0);
  for (int64_t p = 0; p < n_patches; ++p) {
    const uint32_t rc = h_rect[p];
    off[p + 1] = off[p] + (int64_t)((rc >> 16) & 255u) * (int64_t)(rc >> 24) * tap_bytes;
  }
  std::vector<int64_t> cfirst;       // first patch of every chunk
  for (int64_t p = 0; p < n_patches;) {
    cfirst.push_back(p);
    int64_t q = p;
    while (q < n_patches && off[q + 1] - off[p] <= (int64_t)chunk) ++q;
    if (q == p) return fail(PXR_ERR_INTERNAL, "window of patch %lld exceeds the staging chunk", (long long)p);
    p = q;
  }
  cfirst.push_back(n_patches);
  const size_t n_chunks = cfirst.size() - 1;
  DevBuf<int64_t> d_off;
  PXR_TRY(d_off.upload(off.data(), off.size(), stream));
  PXR_CUDA(cudaEventRecord(s->start, stream));          // d_off (and the slab's earlier users) before the worker streams
  int n_threads = (int)std::min<size_t>(Stager::kThreads, n_chunks);
  {
    cpu_set_t cs; CPU_ZERO(&cs);
    if (sched_getaffinity(0, sizeof(cs), &cs) == 0) n_threads = std::max(1, std::min(n_threads, CPU_COUNT(&cs)));
    long quota = 0, period = 0;
    if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
      char q[32] = {};
      if (std::fscanf(f, "%31s %ld", q, &period) == 2 && q[0] != 'm' && period > 0) quota = std::atol(q);
      std::fclose(f);
    }
    if (quota > 0) n_threads = std::max(1, std::min<int>(n_threads, (int)(quota / period)));
  }
  const size_t patch_bytes = (size_t)ph * pw * tap_bytes;
  cudaError_t errs[Stager::kThreads];
  const int device = ctx->device;
  for (int t = 0; t < n_threads; ++t)
    for (int b = 0; b < Stager::kBufs; ++b)
      if (!s->dev[t][b]) PXR_CUDA(cudaMalloc((void**)&s->dev[t][b], chunk));
  auto work = [&](int t) {
    NumaLocalScope numa(device);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->st[t], s->start, 0);
    size_t k = 0;
    for (size_t c = t; c < n_chunks && e == cudaSuccess; c += n_threads, ++k) {
      const int b = (int)(k % Stager::kBufs);
      e = cudaEventSynchronize(s->ev[t][b]);     // the previous DMA out of this pinned buffer and the scatter out of its device twin
      if (e != cudaSuccess) break;
      const int64_t p0 = cfirst[c], p1 = cfirst[c + 1];
      uint8_t* out = s->pin[t][b];
      int blk = 0;
      if (n_blocks > 1) blk = (int)(std::upper_bound(block_first, block_first + n_blocks + 1, p0) - block_first) - 1;
      for (int64_t p = p0; p < p1; ++p) {
        while (n_blocks > 1 && p >= block_first[blk + 1]) ++blk;
        const uint8_t* src = (const uint8_t*)srcs[blk] + (size_t)(p - (n_blocks > 0 ? block_first[blk] : 0)) * patch_bytes;
        const uint32_t rc = h_rect[p];
        const int r0 = (int)(rc & 255u), c0 = (int)((rc >> 8) & 255u), rows = (int)((rc >> 16) & 255u), cols = (int)(rc >> 24);
        const size_t row_bytes = (size_t)cols * tap_bytes;
        if (p + 1 < p1 && (n_blocks <= 1 || p + 1 < block_first[blk + 1])) {
          // the next patch's rows: 2 KiB pieces 4 KiB apart defeat the hardware prefetcher, ask for them a patch ahead
          const uint32_t rn = h_rect[p + 1];
          const int nr0 = (int)(rn & 255u), nc0 = (int)((rn >> 8) & 255u), nrows = (int)((rn >> 16) & 255u), ncols = (int)(rn >> 24);
          const uint8_t* nsrc = src + patch_bytes;
          const size_t nrb = (size_t)ncols * tap_bytes;
          for (int r = 0; r < nrows; ++r) {
            const uint8_t* q = nsrc + ((size_t)(nr0 + r) * pw + nc0) * tap_bytes;
            for (size_t b = 0; b < nrb; b += 64) __builtin_prefetch(q + b, 0, 0);
          }
        }
        if (cols == pw) { std::memcpy(out, src + (size_t)r0 * pw * tap_bytes, row_bytes * rows); out += row_bytes * rows; }
        else for (int r = 0; r < rows; ++r) { std::memcpy(out, src + ((size_t)(r0 + r) * pw + c0) * tap_bytes, row_bytes); out += row_bytes; }
      }
      const size_t len = (size_t)(off[p1] - off[p0]);
      e = cudaMemcpyAsync(s->dev[t][b], s->pin[t][b], len, cudaMemcpyHostToDevice, s->st[t]);
      if (e == cudaSuccess) {
        scatter_windows_kernel<<<(unsigned)std::min<int64_t>(p1 - p0, 1024), 256, 0, s->st[t]>>>(s->dev[t][b], slab, d_rect, d_off.p, p0,
                                                                                            (int)(p1 - p0), ph, pw, tap_bytes);
        e = cudaGetLastError();
      }
      if (e == cudaSuccess) e = cudaEventRecord(s->ev[t][b], s->st[t]);
    }
    errs[t] = e;
  };
  std::vector<std::thread> pool;
  int started = 1;
  try {
    for (int t = 1; t < n_threads; ++t) { pool.emplace_back(work, t); ++started; }
  } catch (const std::exception&) {
  }
  work(0);
  for (int t = started; t < n_threads; ++t) work(t);
  for (auto& th : pool) th.join();
  for (int t = 0; t < n_threads; ++t)
    if (errs[t] != cudaSuccess) {
      cudaDeviceSynchronize();
      return fail(PXR_ERR_CUDA, "window upload failed: %s", cudaGetErrorString(errs[t]));
    }
  for (int t = 0; t < n_threads; ++t) {
    const size_t mine = (n_chunks - t + n_threads - 1) / n_threads;
    const int last = (int)((mine - 1) % Stager::kBufs);
    PXR_CUDA(cudaStreamWaitEvent(stream, s->ev[t][last], 0));
  }
  ctx->launches += (int64_t)n_chunks;
  // d_off is read by the scatter kernels: keep it until they are done (the caller's stream now depends on all of them)
  PXR_CUDA(cudaStreamSynchronize(stream));
  if (h2d_bytes) *h2d_bytes += (double)off[n_patches];
  return PXR_OK;
}

int upload_bytes(pxr_ctx* ctx, void* dst, const void* src, size_t bytes, double* h2d_bytes, cudaStream_t stream) {
  return upload_segments(ctx, dst, &src, &bytes, 1, h2d_bytes, stream);
}

int upload_stream(pxr_ctx* ctx, cudaStream_t* out) {
  if (std::getenv("PXR_UPLOAD_SAME_STREAM")) { *out = ctx->stream; return PXR_OK; }   // A/B switch: no overlap with host setup
  if (!ctx->upload_stream) PXR_CUDA(cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
  *out = ctx->upload_stream;
  return PXR_OK;
}

}  // namespace pxr
