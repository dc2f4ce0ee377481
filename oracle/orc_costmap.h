// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's cost-map extraction.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything under oracle/.
//
// CostMapExtractor::FillPointCostmap, pixsfm/bundle_adjustment/src/costmap_extractor.h:230-358,
// restricted to what the default configuration runs (CostMapConfig :18-40, CostMapBundleAdjuster
// defaults bundle_adjustment/main.py:227-238): upsampling_factor == 1, compute_cross_derivative ==
// false, so the "no interpolation" branches (:253-279 and :330-341) are taken:
//   f      = raw patch value at (y,x) cast to double  (NOT L2-normalised)
//   dfdr   = 0.5 * double( dtype(top) - dtype(bottom) )   -- the difference is formed in dtype (Eigen
//            evaluates Map<dtype> - Map<dtype> before .cast<double>()), rows clamped to the patch
//   dfdc   likewise with (right, left)
//   r = f - ref ; cost = 0.5 * rho0(|r|^2) ; if cost > 1e-8: d = rho1 * r.dfd{r,c}, optional sqrt (:303-322)
//   SetEntry stores dtype_o(value) (featurepatch.h:246-248) = half(float(double)) for fp16 outputs.
// The per-observation reference is the one ReferenceExtractor produced for the observation's 3D point
// (RunSubset :151-205).
#pragma once
#include "orc_ba.h"

namespace orc {

template <typename T> struct CmT;
template <> struct CmT<half_t> {
  static inline double val(half_t a) { return (double)a; }
  static inline double diff(half_t a, half_t b) { return (double)(half_t)((float)a - (float)b); }
  static inline half_t cast(double v) { return (half_t)(float)v; }
};
template <> struct CmT<float> {
  static inline double val(float a) { return (double)a; }
  static inline double diff(float a, float b) { volatile float d = a - b; return (double)d; }
  static inline float cast(double v) { return (float)v; }
};
template <> struct CmT<double> {
  static inline double val(double a) { return a; }
  static inline double diff(double a, double b) { return a - b; }
  static inline double cast(double v) { return v; }
};

template <typename T>
inline void FillPointCostmap(const T* src, int ph, int pw, int C, const double* ref, const Loss& loss,
                             bool as_gradientfield, bool apply_sqrt, T* out, int OC) {
  for (int y = 0; y < ph; ++y)
    for (int x = 0; x < pw; ++x) {
      const T* pc = src + ((size_t)y * pw + x) * C;
      T* op = out + ((size_t)y * pw + x) * OC;
      double s = 0, dr = 0, dc = 0;
      if (as_gradientfield) {
        const int top = std::min(ph - 1, y + 1), bottom = std::max(0, y - 1);
        const int right = std::min(pw - 1, x + 1), left = std::max(0, x - 1);
        const T* pt = src + ((size_t)top * pw + x) * C;
        const T* pb = src + ((size_t)bottom * pw + x) * C;
        const T* pr = src + ((size_t)y * pw + right) * C;
        const T* pl = src + ((size_t)y * pw + left) * C;
        for (int ch = 0; ch < C; ++ch) {
          const double r = CmT<T>::val(pc[ch]) - ref[ch];
          s += r * r;
          dr += r * (CmT<T>::diff(pt[ch], pb[ch]) * 0.5);
          dc += r * (CmT<T>::diff(pr[ch], pl[ch]) * 0.5);
        }
      } else {
        for (int ch = 0; ch < C; ++ch) { const double r = CmT<T>::val(pc[ch]) - ref[ch]; s += r * r; }
      }
      double rho[3];
      loss.Evaluate(s, rho);
      double cost = rho[0] * 0.5;
      if (as_gradientfield) {
        double dcostdr = 0, dcostdc = 0;
        if (cost > 1.0e-8) {
          dcostdr = rho[1] * dr;
          dcostdc = rho[1] * dc;
          if (apply_sqrt) { cost = std::sqrt(cost); dcostdr *= 0.5 / cost; dcostdc *= 0.5 / cost; }
        }
        op[0] = CmT<T>::cast(cost); op[1] = CmT<T>::cast(dcostdr); op[2] = CmT<T>::cast(dcostdc);
      } else {
        if (apply_sqrt) cost = std::sqrt(cost);
        op[0] = CmT<T>::cast(cost);
      }
    }
}

// All observations of the problem; `out` is [n_patches][ph][pw][OC] in the patches' dtype.
inline int ComputeCostmaps(const pxr_ba_desc& d, const Loss& loss, bool as_gradientfield, bool apply_sqrt, void* out) {
  if (!d.refs) return 1;
  const int OC = as_gradientfield ? 3 : 1;
  const size_t in_stride = (size_t)d.ph * d.pw * d.channels, out_stride = (size_t)d.ph * d.pw * OC;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t o = 0; o < d.n_obs; ++o) {
    const Patch p = MakePatch(d, o);
    const int64_t pi = d.obs_patch ? d.obs_patch[o] : o;
    const double* ref = d.refs + (size_t)d.obs_pt[o] * d.channels;
    switch (d.patch_dtype) {
      case PXR_F16: FillPointCostmap<half_t>((const half_t*)p.data, d.ph, d.pw, d.channels, ref, loss, as_gradientfield, apply_sqrt, (half_t*)out + pi * out_stride, OC); break;
      case PXR_F32: FillPointCostmap<float>((const float*)p.data, d.ph, d.pw, d.channels, ref, loss, as_gradientfield, apply_sqrt, (float*)out + pi * out_stride, OC); break;
      default: FillPointCostmap<double>((const double*)p.data, d.ph, d.pw, d.channels, ref, loss, as_gradientfield, apply_sqrt, (double*)out + pi * out_stride, OC); break;
    }
  }
  (void)in_stride;
  return 0;
}

}  // namespace orc
