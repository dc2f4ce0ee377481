// pxr_inner.cu — instantiations of the per-point inner-iteration kernel (pxr_inner.cuh).
#include "pxr_ba_host.h"

namespace pxr {

template <typename T, int C>
static int launch_inner_tc(pxr_ctx* ctx, bool fs, const InnerArgs& a) {
  const unsigned grid = (unsigned)cdiv(a.n_points * 32, 128);
  if (fs) PXR_LAUNCH(ctx, (ba_inner_kernel<T, C, true>), grid, 128, 0, a);
  else PXR_LAUNCH(ctx, (ba_inner_kernel<T, C, false>), grid, 128, 0, a);
  PXR_CUDA(cudaGetLastError());
  return PXR_OK;
}

int launch_inner(pxr_ctx* ctx, int dtype, int C, bool float_simd, const InnerArgs& a) {
#define PXR_CASE(T, CC) \
  if (C == CC) return launch_inner_tc<T, CC>(ctx, float_simd, a);
  if (dtype == PXR_F16) {
    PXR_CASE(__half, 128) PXR_CASE(__half, 64) PXR_CASE(__half, 32) PXR_CASE(__half, 16) PXR_CASE(__half, 8)
    PXR_CASE(__half, 256)
  } else if (dtype == PXR_F32) {
    PXR_CASE(float, 128) PXR_CASE(float, 64) PXR_CASE(float, 16)
  } else if (dtype == PXR_F64) {
    PXR_CASE(double, 128) PXR_CASE(double, 16)
  }
#undef PXR_CASE
  return fail(PXR_ERR_UNSUPPORTED, "Unsupported dimensions (CHANNELS=%d, dtype=%d, N_NODES=1).", C, dtype);
}

}  // namespace pxr
