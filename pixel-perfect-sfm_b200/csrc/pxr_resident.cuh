// pxr_resident.cuh — window residency of the patch slab.
//
// A featuremetric BA evaluation touches the 4x4 tap window under the projected point — 4 KiB of a 64 KiB patch (16x16,
// C=128, fp16) — and over a whole solve the point moves by a pixel or two.  The reference keeps every patch in host RAM
// (features/src/featurepatch.h:35-70) and its solver reads what it needs; a device solver that first copies EVERY byte
// over PCIe spends 10x longer on the copy than on the 20 LM iterations (32.8 GB at ~40 GB/s vs 75 ms at BASELINE
// configs[2]).  So pxr_ba_run brings over only a W x W window per observation, centred on its initial projection
// (W = 8: 1/4 of the bytes): host threads pack the windows into pinned staging buffers — the only bytes of the source
// they touch, pinned or pageable —, the DMA engine moves the packed chunks and a small kernel scatters them into place
// (pxr_upload.cu::upload_windows).  The evaluation kernels check every tap window against the resident rectangle of its
// patch: an observation that walks out of it is reported, its whole patch is fetched, and the observation is evaluated
// again before anything consumes the result — same arithmetic on the same taps as a solve on fully resident patches
// (tests/test_gpu_resident.py).  (A first version let the SMs read the windows straight out of the pinned buffer;
// zero-copy reads of 2 KiB rows ran at 9 GB/s on the B200 box, the packed DMA runs at PCIe speed.)
//
// Layout: the slab keeps its full [n_patches][ph][pw][C] shape in device memory (allocated, mostly never written),
// so no kernel changes its addressing; res_rect[patch] = r0 | c0<<8 | rows<<16 | cols<<24 says what is valid.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pxr {

// rectangle of every patch: a W x W window centred on the 4x4 taps of the initial projection (margin (W-4)/2 on the low
// side), the whole patch when several observations read it
static __global__ void __launch_bounds__(256) resident_rect_kernel(const double* __restrict__ uv, const int64_t* __restrict__ obs_patch,
                                                                   const uint8_t* __restrict__ patch_shared, int64_t n_obs,
                                                                   int ph, int pw, int window, uint32_t* __restrict__ rect) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_obs) return;
  const int64_t p = obs_patch ? obs_patch[o] : o;
  int r0 = 0, c0 = 0, rows = ph, cols = pw;
  if (!(patch_shared && patch_shared[p])) {
    const double u = uv[2 * o], v = uv[2 * o + 1];
    const int m = (window - 4) / 2;
    const int fc = (int)fmin(fmax(floor(u), -4.0), (double)pw + 4.0), fr = (int)fmin(fmax(floor(v), -4.0), (double)ph + 4.0);
    c0 = min(max(fc - 1 - m, 0), pw - window); r0 = min(max(fr - 1 - m, 0), ph - window);
    rows = window; cols = window;
  }
  rect[p] = (uint32_t)r0 | ((uint32_t)c0 << 8) | ((uint32_t)rows << 16) | ((uint32_t)cols << 24);   // shared: same value from every reader
}

// the listed observations' patches are whole now
static __global__ void __launch_bounds__(256) resident_mark_full_kernel(const int64_t* __restrict__ list, int64_t n,
                                                                        const int64_t* __restrict__ obs_patch, int ph, int pw,
                                                                        uint32_t* __restrict__ rect) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t o = list[k];
  rect[obs_patch ? obs_patch[o] : o] = ((uint32_t)ph << 16) | ((uint32_t)pw << 24);
}

}  // namespace pxr
