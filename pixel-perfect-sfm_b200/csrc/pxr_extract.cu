// pxr_extract.cu — dense feature map of one image -> per-keypoint patch slab, on the device.
//
// Reference: FeatureExtractor.tensor_to_fmap, sparse branch (pixsfm/features/extractor.py:176-201): the [1,C,H,W]
// CNN output is L2-normalised over C, cast to the storage dtype, and one ps x ps window per keypoint is cut out by
// extract_patches_numpy (features/extract_patches.py:36-44), which moves the windows GPU -> CPU -> numpy — the copy its
// authors flag as "main performance bottleneck" — before the optimizers upload them again.  Here the windows are
// gathered straight into the HWC-interleaved slab layout of FeaturePatch (features/src/featurepatch.h:244-262) in
// device memory, ready for pxr_ba_create / pxr_ka_run with patches_on_device / device patch blocks.
//
// One warp per output pixel; lanes stride the channels (for a channels-first map the reads are H*W apart: this is a
// gather, sized by the patches, not a hot-path kernel).  Normalisation in fp32 like torch.nn.functional.normalize
// (x / max(||x||_2, 1e-12)).
#include <cuda_fp16.h>

#include "pxr_internal.h"
#include "pxr_device.cuh"

namespace pxr {

template <typename T> __device__ __forceinline__ float ex_load(const T* p);
template <> __device__ __forceinline__ float ex_load<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ex_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ex_load<double>(const double* p) { return (float)*p; }
template <typename T> __device__ __forceinline__ T ex_store(float v);
template <> __device__ __forceinline__ __half ex_store<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float ex_store<float>(float v) { return v; }
template <> __device__ __forceinline__ double ex_store<double>(float v) { return (double)v; }

struct ExtractArgs {
  const void* dense; int C, H, W, channels_first;
  const int32_t* corners;        // [n][2] (x0, y0), device
  int64_t n; int ps; int l2_normalize;
  void* out;                     // [n][ps][ps][C]
};

template <typename TI, typename TO>
static __global__ void __launch_bounds__(256) extract_patches_kernel(ExtractArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t pix = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t per = (int64_t)a.ps * a.ps;
  if (pix >= a.n * per) return;
  const int64_t k = pix / per;
  const int r = (int)((pix - k * per) / a.ps), c = (int)((pix - k * per) % a.ps);
  const int x = a.corners[2 * k] + c, y = a.corners[2 * k + 1] + r;
  const TI* src = reinterpret_cast<const TI*>(a.dense);
  const int64_t plane = (int64_t)a.H * a.W;
  const int64_t base = a.channels_first ? (int64_t)y * a.W + x : ((int64_t)y * a.W + x) * a.C;
  const int64_t cstride = a.channels_first ? plane : 1;
  float sq = 0.f;
  if (a.l2_normalize) {
    for (int ch = lane; ch < a.C; ch += 32) { const float v = ex_load<TI>(src + base + ch * cstride); sq += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  const float denom = a.l2_normalize ? fmaxf(sqrtf(sq), 1e-12f) : 1.f;
  TO* dst = reinterpret_cast<TO*>(a.out) + pix * a.C;
  for (int ch = lane; ch < a.C; ch += 32) dst[ch] = ex_store<TO>(ex_load<TI>(src + base + ch * cstride) / denom);
}

template <typename TI>
static int launch_extract(pxr_ctx* ctx, const ExtractArgs& a, int out_dtype) {
  const int64_t warps = a.n * a.ps * a.ps;
  const unsigned grid = (unsigned)cdiv(warps * 32, 256);
  if (out_dtype == PXR_F16) PXR_LAUNCH(ctx, (extract_patches_kernel<TI, __half>), grid, 256, 0, a);
  else if (out_dtype == PXR_F32) PXR_LAUNCH(ctx, (extract_patches_kernel<TI, float>), grid, 256, 0, a);
  else PXR_LAUNCH(ctx, (extract_patches_kernel<TI, double>), grid, 256, 0, a);
  PXR_CUDA(cudaGetLastError());
  return PXR_OK;
}

}  // namespace pxr

using namespace pxr;

extern "C" int pxr_extract_patches(pxr_ctx* ctx, const void* dense, int32_t dense_on_device, int32_t in_dtype, int32_t C,
                                   int32_t H, int32_t W, int32_t channels_first, const int32_t* corners, int64_t n, int32_t ps,
                                   int32_t l2_normalize, int32_t out_dtype, void* out_host, void** out_device) {
  if (!ctx || !dense || (n > 0 && !corners) || (!out_host && !out_device))
    return fail(PXR_ERR_INVALID_ARGUMENT, "NULL argument");
  if (C < 1 || H < 1 || W < 1 || ps < 1 || ps > H || ps > W || n < 0)
    return fail(PXR_ERR_INVALID_ARGUMENT, "bad map / patch sizes (C=%d H=%d W=%d ps=%d)", C, H, W, ps);
  if (in_dtype < PXR_F16 || in_dtype > PXR_F64 || out_dtype < PXR_F16 || out_dtype > PXR_F64)
    return fail(PXR_ERR_INVALID_ARGUMENT, "bad dtype");
  for (int64_t k = 0; k < n; ++k)
    if (corners[2 * k] < 0 || corners[2 * k + 1] < 0 || corners[2 * k] + ps > W || corners[2 * k + 1] + ps > H)
      return fail(PXR_ERR_INVALID_ARGUMENT, "patch %lld leaves the map (corner %d,%d)", (long long)k, corners[2 * k], corners[2 * k + 1]);
  PXR_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t isz = in_dtype == PXR_F16 ? 2 : (in_dtype == PXR_F32 ? 4 : 8), osz = out_dtype == PXR_F16 ? 2 : (out_dtype == PXR_F32 ? 4 : 8);
  DevBuf<uint8_t> d_dense;
  DevBuf<int32_t> d_corners;
  ExtractArgs a;
  a.dense = dense;
  if (!dense_on_device) {
    PXR_TRY(d_dense.alloc((size_t)C * H * W * isz));
    PXR_TRY(upload_bytes(ctx, d_dense.p, dense, (size_t)C * H * W * isz));
    a.dense = d_dense.p;
  }
  const size_t out_bytes = (size_t)n * ps * ps * C * osz;
  uint8_t* d_out = nullptr;
  if (n > 0) {
    PXR_TRY(d_corners.upload(corners, (size_t)n * 2, s));
    PXR_CUDA(cudaMalloc((void**)&d_out, out_bytes));
  }
  a.C = C; a.H = H; a.W = W; a.channels_first = channels_first; a.corners = d_corners.p; a.n = n; a.ps = ps;
  a.l2_normalize = l2_normalize; a.out = d_out;
  int rc = PXR_OK;
  if (n > 0) {
    if (in_dtype == PXR_F16) rc = launch_extract<__half>(ctx, a, out_dtype);
    else if (in_dtype == PXR_F32) rc = launch_extract<float>(ctx, a, out_dtype);
    else rc = launch_extract<double>(ctx, a, out_dtype);
  }
  if (rc == PXR_OK && out_host && n > 0) {
    const cudaError_t e = cudaMemcpyAsync(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) rc = fail(PXR_ERR_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
  }
  if (rc == PXR_OK) {
    const cudaError_t e = cudaStreamSynchronize(s);      // corners / staged dense map are locals of this call
    if (e != cudaSuccess) rc = fail(PXR_ERR_CUDA, "extract_patches failed: %s", cudaGetErrorString(e));
  }
  if (rc == PXR_OK && out_device) *out_device = d_out;
  else if (d_out) cudaFree(d_out);
  return rc;
}
