// pxr_ka.cu — featuremetric keypoint adjustment (placeholder until the KA kernels land this round).
#include "pxr_internal.h"
using namespace pxr;
extern "C" int pxr_ka_run(pxr_ctx*, const pxr_ka_desc*, const pxr_interp_config*, const pxr_solver_options*, pxr_summary*) {
  return fail(PXR_ERR_UNSUPPORTED, "pxr_ka_run: KA kernels not built yet");
}
