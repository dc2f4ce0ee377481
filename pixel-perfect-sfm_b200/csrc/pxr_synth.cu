// pxr_synth.cu — device-side synthetic feature patches for bench.py ("data": "synthetic").
// Same field model as pixsfm/util/synthetic.py: per smooth field j and channel c
// F_j(du,dv)[c] = a + b*du + g*dv + h*du*dv with a~N(0,1), b,g,h~N(0,0.15^2), L2-normalised per
// pixel, plus N(0, noise^2), stored fp16 HWC.  Counter-based hash RNG => reproducible.
#include <cuda_fp16.h>

#include "pxr_internal.h"

namespace pxr {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float normal_from(uint64_t key) {
  const uint64_t h = mix64(key);
  const float u1 = ((uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777217.0f);  // (0,1]
  const float u2 = (uint32_t)(h & 0xFFFFFF) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
}

// one CTA per patch, one thread per pixel
__global__ void __launch_bounds__(256) synth_patches_kernel(__half* __restrict__ out, int64_t n_patches, int ps, int C,
                                                            const double* __restrict__ uv0, const int64_t* __restrict__ field,
                                                            uint64_t seed, float noise) {
  extern __shared__ float coef[];  // [4][C]
  const int64_t pi = blockIdx.x;
  if (pi >= n_patches) return;
  const int64_t fid = field[pi];
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    const int k = i / C, c = i % C;
    const float sd = k == 0 ? 1.0f : 0.15f;
    coef[i] = sd * normal_from(seed ^ mix64((uint64_t)fid * 4099ull + (uint64_t)k * 1000003ull + (uint64_t)c * 7919ull + 17ull));
  }
  __syncthreads();
  const int npx = ps * ps;
  for (int px = threadIdx.x; px < npx; px += blockDim.x) {
    const int row = px / ps, col = px % ps;
    const float du = (float)((double)col - uv0[2 * pi]);
    const float dv = (float)((double)row - uv0[2 * pi + 1]);
    float n2 = 0.f;
    for (int c = 0; c < C; ++c) {
      const float f = coef[c] + coef[C + c] * du + coef[2 * C + c] * dv + coef[3 * C + c] * du * dv;
      n2 += f * f;
    }
    const float ninv = rsqrtf(n2);
    __half* dst = out + ((size_t)pi * npx + px) * C;
    for (int c = 0; c < C; c += 2) {
      float f0 = (coef[c] + coef[C + c] * du + coef[2 * C + c] * dv + coef[3 * C + c] * du * dv) * ninv;
      float f1 = (coef[c + 1] + coef[C + c + 1] * du + coef[2 * C + c + 1] * dv + coef[3 * C + c + 1] * du * dv) * ninv;
      if (noise > 0.f) {
        const uint64_t key = seed * 0x100000001B3ull + ((uint64_t)pi * npx + px) * (uint64_t)C + c;
        f0 += noise * normal_from(key);
        f1 += noise * normal_from(key + 1);
      }
      *reinterpret_cast<__half2*>(dst + c) = __floats2half2_rn(f0, f1);
    }
  }
}

}  // namespace pxr

using namespace pxr;

extern "C" int pxr_synth_patches_device(pxr_ctx* ctx, void** d_out, int64_t n_patches, int ps, int channels,
                                        const double* uv0, const int64_t* field_id, uint64_t seed, double noise_sigma) {
  if (!ctx || !d_out || !uv0 || !field_id || n_patches < 1 || ps < 1 || channels < 2 || (channels & 1))
    return fail(PXR_ERR_INVALID_ARGUMENT, "bad synth arguments");
  PXR_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> duv;
  DevBuf<int64_t> dfield;
  PXR_TRY(duv.upload(uv0, (size_t)n_patches * 2, ctx->stream));
  PXR_TRY(dfield.upload(field_id, (size_t)n_patches, ctx->stream));
  void* out = nullptr;
  const size_t bytes = (size_t)n_patches * ps * ps * channels * sizeof(__half);
  PXR_CUDA(cudaMalloc(&out, bytes));
  PXR_LAUNCH(ctx, synth_patches_kernel, (unsigned)n_patches, 256, (size_t)4 * channels * sizeof(float),
             (__half*)out, n_patches, ps, channels, duv.p, dfield.p, seed, (float)noise_sigma);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { cudaFree(out); return fail(PXR_ERR_CUDA, "synth kernel failed: %s", cudaGetErrorString(e)); }
  *d_out = out;
  return PXR_OK;
}
