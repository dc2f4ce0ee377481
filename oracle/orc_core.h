// oracle/orc_core.h — TEST INFRASTRUCTURE ONLY (never linked/imported by the product).
//
// CPU restatement of the reference's featuremetric interpolation / projection /
// loss arithmetic, written from the reference sources cited per function
// (paths relative to /root/reference).  Ceres and COLMAP are third-party
// dependencies that are NOT vendored in the reference (README.md:34: ceres >= 2.1,
// COLMAP 3.8); their published formulas are restated here and marked (ceres)/(colmap).
//
// Pinning: orc_core's splines are checked bit-for-bit against the reference's own
// AVX2 header compiled verbatim (oracle/_ref, see oracle/Makefile) and against the
// reference's known-answer tests (tests/test_oracle_kat.py).  Featuremetric BA/KA
// end results have no test in the reference => "parity unpinned" at that level.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

typedef _Float16 half_t;

enum DType { F16 = 0, F32 = 1, F64 = 2 };

// ---------------------------------------------------------------------------
// a1. CubicHermiteSplineSIMD — pixsfm/base/src/cubic_hermite_spline_simd.h
// ---------------------------------------------------------------------------
// f16/f32-in overload (:123-192): 8 fp32 lanes, op-for-op identical (fmaf == vfmadd).
template <typename IN_T>
inline void CubicHermiteF32(IN_T p0i, IN_T p1i, IN_T p2i, IN_T p3i, double x, float* f,
                            float* dfdx) {
  const float p0 = (float)p0i, p1 = (float)p1i, p2 = (float)p2i, p3 = (float)p3i;
  const float x2s = static_cast<float>(x * x);        // :129
  const float fourx = (float)(4.0f * x);              // :135 (double product, then to ps)
  const float xhalf = (float)(x * 0.5f);              // :136
  const float onefivex2 = 1.5f * x2s;                 // :138
  const float t1 = fmaf(3.0f, p1, -p0);               // :149 fmsub
  const float t2 = fmaf(3.0f, p2, -p3);
  const float t4 = fmaf(4.0f, p2, -p3);
  const float t5 = fmaf(2.5f, p1, -p0);
  const float t6 = fmaf(-1.0f, p0, p2);
  const float t3 = t1 - t2;
  const float b = fmaf(0.5f, t4, -t5);
  if (f) {
    const float t7 = fmaf(xhalf, t6, p1);
    const float t8 = fmaf(xhalf, t3, b);
    *f = fmaf(x2s, t8, t7);
  }
  if (dfdx) {
    const float t9 = fmaf(fourx, b, t6);
    const float t10 = onefivex2 * t3;
    *dfdx = fmaf(0.5f, t9, t10);
  }
}

// f64-in overload (:56-121): 4 fp64 lanes.
inline void CubicHermiteF64(double p0, double p1, double p2, double p3, double x, double* f,
                            double* dfdx) {
  const double x2s = x * x;
  const double fourx = 4.0 * x;
  const double xhalf = x * 0.5;
  const double onefivex2 = 1.5 * x2s;
  const double t1 = fma(3.0, p1, -p0);
  const double t2 = fma(3.0, p2, -p3);
  const double t4 = fma(4.0, p2, -p3);
  const double t5 = fma(2.5, p1, -p0);
  const double t6 = fma(-1.0, p0, p2);
  const double t3 = t1 - t2;
  const double b = fma(0.5, t4, -t5);
  if (f) {
    const double t7 = fma(xhalf, t6, p1);
    const double t8 = fma(xhalf, t3, b);
    *f = fma(x2s, t8, t7);
  }
  if (dfdx) {
    const double t9 = fma(fourx, b, t6);
    const double t10 = onefivex2 * t3;
    *dfdx = fma(0.5, t9, t10);
  }
}

// (ceres) ceres::CubicHermiteSpline, include/ceres/cubic_interpolation.h — used by the
// reference when C < 8 (interpolation.h:224,230-262) and as the KAT comparison
// (interpolation_test.cc:327-364).
inline void CubicHermiteCeres(double p0, double p1, double p2, double p3, double x, double* f,
                              double* dfdx) {
  const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
  const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
  const double c = 0.5 * (-p0 + p2);
  const double d = p1;
  if (f) *f = d + x * (c + x * (b + x * a));
  if (dfdx) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

// ---------------------------------------------------------------------------
// a2. Grid2D — pixsfm/base/src/grid2d.h:29-73 (row-major, interleaved, per-tap clamp)
// ---------------------------------------------------------------------------
struct Patch {
  const void* data;
  int dtype;  // DType
  int h, w, c;
  int corner[2];     // (x0, y0)
  double scale[2];   // (sx, sy)
  double upsampling; // 1.0
  inline size_t Offset(int r, int col) const {
    const int ri = std::min(std::max(0, r), h - 1);
    const int ci = std::min(std::max(0, col), w - 1);
    return (size_t)(w * ri + ci) * c;
  }
  inline double Value(size_t off) const {
    switch (dtype) {
      case F16: return (double)((const half_t*)data)[off];
      case F32: return (double)((const float*)data)[off];
      default: return ((const double*)data)[off];
    }
  }
};

// ---------------------------------------------------------------------------
// a3. BiCubicInterpolator — pixsfm/base/src/interpolation.h:168-274
//   C >= 8 : EvaluateSIMD (:177-218): horizontal pass in the input's SIMD
//            precision (f16/f32 -> fp32 lanes, f64 -> fp64 lanes), vertical pass
//            in `dtype` (double by default; float when use_float_simd, :619-623).
//   C <  8 : ceres::CubicHermiteSpline in double (:230-262).
// ---------------------------------------------------------------------------
// Optional inner kernel: the reference's own AVX2 spline (pixsfm/base/src/cubic_hermite_spline_simd.h, compiled
// verbatim into oracle/_ref/libpxref.so).  orc_use_reference_spline() resolves the three entry points with dlopen;
// the restatement below stays the default and is pinned bit-exact against it (tests/test_oracle_golden.py), so
// switching changes speed, not results.  Used by the timed CPU arm of bench.py.
struct RefSpline {
  int (*f16)(int, const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, double, double*, double*) = nullptr;
  int (*f32)(int, const float*, const float*, const float*, const float*, double, double*, double*) = nullptr;
  int (*f64)(int, const double*, const double*, const double*, const double*, double, double*, double*) = nullptr;
};
inline RefSpline& GlobalRefSpline() { static RefSpline r; return r; }

// BiCubicInterpolator::EvaluateSIMD (interpolation.h:177-218) through the verbatim header: four horizontal splines over
// the C-contiguous taps of each window row, then the vertical pass in fp64 over the four rows of values / x-derivatives.
inline bool BiCubicRef(const Patch& g, double r, double c, double* f, double* dfdr, double* dfdc) {
  const RefSpline& R = GlobalRefSpline();
  const int C = g.c;
  if (C < 8 || C > 256 || !R.f64) return false;
  if ((g.dtype == F16 && !R.f16) || (g.dtype == F32 && !R.f32)) return false;
  const int row = (int)std::floor(r), col = (int)std::floor(c);
  const double xc = c - col, xr = r - row;
  double fi[4][256], di[4][256];
  for (int i = 0; i < 4; ++i) {
    size_t o[4];
    for (int j = 0; j < 4; ++j) o[j] = g.Offset(row - 1 + i, col - 1 + j);
    int rc;
    if (g.dtype == F16) { const uint16_t* d = (const uint16_t*)g.data; rc = R.f16(C, d + o[0], d + o[1], d + o[2], d + o[3], xc, fi[i], di[i]); }
    else if (g.dtype == F32) { const float* d = (const float*)g.data; rc = R.f32(C, d + o[0], d + o[1], d + o[2], d + o[3], xc, fi[i], di[i]); }
    else { const double* d = (const double*)g.data; rc = R.f64(C, d + o[0], d + o[1], d + o[2], d + o[3], xc, fi[i], di[i]); }
    if (rc != 0) return false;
  }
  double unused[256];
  if (R.f64(C, fi[0], fi[1], fi[2], fi[3], xr, f, dfdr) != 0) return false;
  if (dfdc && R.f64(C, di[0], di[1], di[2], di[3], xr, dfdc, unused) != 0) return false;
  return true;
}

inline void BiCubic(const Patch& g, double r, double c, bool use_float_simd, double* f,
                    double* dfdr, double* dfdc) {
  if (!use_float_simd && GlobalRefSpline().f64 && BiCubicRef(g, r, c, f, dfdr, dfdc)) return;
  const int row = (int)std::floor(r);
  const int col = (int)std::floor(c);
  const int C = g.c;
  const double xc = c - col, xr = r - row;
  size_t off[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) off[i][j] = g.Offset(row - 1 + i, col - 1 + j);
  for (int ch = 0; ch < C; ++ch) {
    if (C >= 8) {
      // NB: the SIMD loops cover floor(C/step)*step channels; the tail uses the
      // double formula (:174-190 of the spline header). All whitelisted C>=8 are
      // multiples of 8, so only the vector body is restated here.
      if (g.dtype == F64) {
        double fi[4], di[4];
        for (int i = 0; i < 4; ++i) {
          const double* d = (const double*)g.data;
          CubicHermiteF64(d[off[i][0] + ch], d[off[i][1] + ch], d[off[i][2] + ch],
                          d[off[i][3] + ch], xc, &fi[i], &di[i]);
        }
        if (use_float_simd) {
          // dtype=float: f0..f3 stored as float (_mm256_storeu_pd_T(float*)), vertical fp32
          float ff[4], dd[4];
          for (int i = 0; i < 4; ++i) { ff[i] = (float)fi[i]; dd[i] = (float)di[i]; }
          float o1, o2, o3;
          CubicHermiteF32<float>(ff[0], ff[1], ff[2], ff[3], xr, &o1, &o2);
          f[ch] = o1; dfdr[ch] = o2;
          if (dfdc) { CubicHermiteF32<float>(dd[0], dd[1], dd[2], dd[3], xr, &o3, nullptr); dfdc[ch] = o3; }
        } else {
          CubicHermiteF64(fi[0], fi[1], fi[2], fi[3], xr, &f[ch], &dfdr[ch]);
          if (dfdc) CubicHermiteF64(di[0], di[1], di[2], di[3], xr, &dfdc[ch], nullptr);
        }
      } else {
        float fi[4], di[4];
        for (int i = 0; i < 4; ++i) {
          if (g.dtype == F16) {
            const half_t* d = (const half_t*)g.data;
            CubicHermiteF32<half_t>(d[off[i][0] + ch], d[off[i][1] + ch], d[off[i][2] + ch],
                                    d[off[i][3] + ch], xc, &fi[i], &di[i]);
          } else {
            const float* d = (const float*)g.data;
            CubicHermiteF32<float>(d[off[i][0] + ch], d[off[i][1] + ch], d[off[i][2] + ch],
                                   d[off[i][3] + ch], xc, &fi[i], &di[i]);
          }
        }
        if (use_float_simd) {
          float o1, o2, o3;
          CubicHermiteF32<float>(fi[0], fi[1], fi[2], fi[3], xr, &o1, &o2);
          f[ch] = o1; dfdr[ch] = o2;
          if (dfdc) { CubicHermiteF32<float>(di[0], di[1], di[2], di[3], xr, &o3, nullptr); dfdc[ch] = o3; }
        } else {
          // _mm256_storeu_ps_T(double*) widens exactly; vertical pass in fp64 lanes
          CubicHermiteF64(fi[0], fi[1], fi[2], fi[3], xr, &f[ch], &dfdr[ch]);
          if (dfdc) CubicHermiteF64(di[0], di[1], di[2], di[3], xr, &dfdc[ch], nullptr);
        }
      }
    } else {
      double fi[4], di[4];
      for (int i = 0; i < 4; ++i)
        CubicHermiteCeres(g.Value(off[i][0] + ch), g.Value(off[i][1] + ch),
                          g.Value(off[i][2] + ch), g.Value(off[i][3] + ch), xc, &fi[i], &di[i]);
      CubicHermiteCeres(fi[0], fi[1], fi[2], fi[3], xr, &f[ch], &dfdr[ch]);
      if (dfdc) CubicHermiteCeres(di[0], di[1], di[2], di[3], xr, &dfdc[ch], nullptr);
    }
  }
}

// (ceres) ceres::BiCubicInterpolator::Evaluate — all-double separable Catmull-Rom; the
// comparison side of interpolation_test.cc:327-364 (TestSimilarToCeres).
inline void BiCubicCeres(const Patch& g, double r, double c, double* f, double* dfdr,
                         double* dfdc) {
  const int row = (int)std::floor(r);
  const int col = (int)std::floor(c);
  for (int ch = 0; ch < g.c; ++ch) {
    double fi[4], di[4];
    for (int i = 0; i < 4; ++i)
      CubicHermiteCeres(g.Value(g.Offset(row - 1 + i, col - 1) + ch),
                        g.Value(g.Offset(row - 1 + i, col) + ch),
                        g.Value(g.Offset(row - 1 + i, col + 1) + ch),
                        g.Value(g.Offset(row - 1 + i, col + 2) + ch), c - col, &fi[i], &di[i]);
    CubicHermiteCeres(fi[0], fi[1], fi[2], fi[3], r - row, &f[ch], &dfdr[ch]);
    CubicHermiteCeres(di[0], di[1], di[2], di[3], r - row, &dfdc[ch], nullptr);
  }
}

// ---------------------------------------------------------------------------
// a4. PixelInterpolator::Evaluate — pixsfm/base/src/interpolation.h:642-677
// ---------------------------------------------------------------------------
struct InterpConfig {
  bool l2_normalize = true;
  bool use_float_simd = false;
};

inline void PixelInterp(const Patch& g, const InterpConfig& cfg, double r, double c, double* f,
                        double* dfdr, double* dfdc) {
  BiCubic(g, r, c, cfg.use_float_simd, f, dfdr, dfdc);
  const int C = g.c;
  if (cfg.l2_normalize) {
    double n2 = 0;
    for (int i = 0; i < C; ++i) n2 += f[i] * f[i];
    const double norm_inv = 1.0 / std::sqrt(n2);
    for (int i = 0; i < C; ++i) f[i] *= norm_inv;
    if (dfdc) {
      double dot = 0;
      for (int i = 0; i < C; ++i) { dfdc[i] *= norm_inv; }
      for (int i = 0; i < C; ++i) dot += f[i] * dfdc[i];
      for (int i = 0; i < C; ++i) dfdc[i] -= dot * f[i];
    }
    if (dfdr) {
      double dot = 0;
      for (int i = 0; i < C; ++i) { dfdr[i] *= norm_inv; }
      for (int i = 0; i < C; ++i) dot += f[i] * dfdr[i];
      for (int i = 0; i < C; ++i) dfdr[i] -= dot * f[i];
    }
  }
}

// a5. FeaturePatch::ToPixelCoordinates — pixsfm/features/src/featurepatch.h:250-255
template <typename T>
inline void ToPixelCoordinates(const Patch& p, const T* xy, T* uv) {
  uv[0] = (xy[0] * p.scale[0] - 0.5 - double(p.corner[0])) * p.upsampling;
  uv[1] = (xy[1] * p.scale[1] - 0.5 - double(p.corner[1])) * p.upsampling;
}

// ---------------------------------------------------------------------------
// (ceres) Jet<N> — forward-mode dual number, the cost model of AutoDiffCostFunction
// (feature_reference.h:87-96: AutoDiffCostFunction<..., N_RESIDUALS, 4,3,3,kNumParams>)
// ---------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; const double yi = 1.0 / y.a; r.a = x.a * yi; for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * yi; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a += s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& x) { return x + s; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a -= s; return r; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& x) { return (-x) + s; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, double s) { return x * (1.0 / s); }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& y) { return Jet<N>(s) / y; }
template <int N> inline Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
template <int N> inline Jet<N>& operator/=(Jet<N>& x, const Jet<N>& y) { x = x / y; return x; }
template <int N> inline Jet<N> sqrt(const Jet<N>& x) { Jet<N> r; r.a = std::sqrt(x.a); const double t = 1.0 / (2.0 * r.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * t; return r; }
template <int N> inline Jet<N> atan(const Jet<N>& x) { Jet<N> r; r.a = std::atan(x.a); const double t = 1.0 / (1.0 + x.a * x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * t; return r; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double atan(double x) { return std::atan(x); }
template <int N> inline double scalar(const Jet<N>& x) { return x.a; }
inline double scalar(double x) { return x; }

// ---------------------------------------------------------------------------
// (colmap) camera models, colmap/base/camera_models.h @3.8 — WorldToImage
// ---------------------------------------------------------------------------
inline int CameraNumParams(int model) {
  switch (model) {
    case 0: return 3; case 1: return 4; case 2: return 4; case 3: return 5;
    case 4: return 8; case 5: return 8; case 6: return 12; default: return -1;
  }
}
// parameter groups (colmap FocalLengthIdxs / PrincipalPointIdxs / ExtraParamsIdxs), as bit masks
inline void CameraParamGroups(int model, uint32_t* focal, uint32_t* pp, uint32_t* extra) {
  switch (model) {
    case 0: *focal = 0x1; *pp = 0x6; *extra = 0; break;
    case 1: *focal = 0x3; *pp = 0xC; *extra = 0; break;
    case 2: *focal = 0x1; *pp = 0x6; *extra = 0x8; break;
    case 3: *focal = 0x1; *pp = 0x6; *extra = 0x18; break;
    case 4: case 5: *focal = 0x3; *pp = 0xC; *extra = 0xF0; break;
    case 6: *focal = 0x3; *pp = 0xC; *extra = 0xFF0; break;
    default: *focal = *pp = *extra = 0;
  }
}

template <typename T>
inline void CameraWorldToImage(int model, const T* p, const T& u, const T& v, T* x, T* y) {
  switch (model) {
    case 0: { *x = p[0] * u + p[1]; *y = p[0] * v + p[2]; } break;
    case 1: { *x = p[0] * u + p[2]; *y = p[1] * v + p[3]; } break;
    case 2: {
      const T r2 = u * u + v * v; const T radial = p[3] * r2;
      const T du = u * radial, dv = v * radial;
      *x = p[0] * (u + du) + p[1]; *y = p[0] * (v + dv) + p[2];
    } break;
    case 3: {
      const T r2 = u * u + v * v; const T radial = p[3] * r2 + p[4] * r2 * r2;
      const T du = u * radial, dv = v * radial;
      *x = p[0] * (u + du) + p[1]; *y = p[0] * (v + dv) + p[2];
    } break;
    case 4: {
      const T u2 = u * u, uv = u * v, v2 = v * v; const T r2 = u2 + v2;
      const T radial = p[4] * r2 + p[5] * r2 * r2;
      const T du = u * radial + T(2.0) * p[6] * uv + p[7] * (r2 + T(2.0) * u2);
      const T dv = v * radial + T(2.0) * p[7] * uv + p[6] * (r2 + T(2.0) * v2);
      *x = p[0] * (u + du) + p[2]; *y = p[1] * (v + dv) + p[3];
    } break;
    case 5: {
      const T r = sqrt(u * u + v * v);
      T du(0.0), dv(0.0);
      if (scalar(r) > std::numeric_limits<double>::epsilon()) {
        const T theta = atan(r);
        const T theta2 = theta * theta, theta4 = theta2 * theta2;
        const T theta6 = theta4 * theta2, theta8 = theta4 * theta4;
        const T thetad = theta * (T(1.0) + p[4] * theta2 + p[5] * theta4 + p[6] * theta6 + p[7] * theta8);
        du = u * thetad / r - u; dv = v * thetad / r - v;
      }
      *x = p[0] * (u + du) + p[2]; *y = p[1] * (v + dv) + p[3];
    } break;
    case 6: {
      const T u2 = u * u, uv = u * v, v2 = v * v; const T r2 = u2 + v2;
      const T r4 = r2 * r2, r6 = r4 * r2;
      const T radial = (T(1.0) + p[4] * r2 + p[5] * r4 + p[8] * r6) /
                       (T(1.0) + p[9] * r2 + p[10] * r4 + p[11] * r6);
      const T du = u * radial + T(2.0) * p[6] * uv + p[7] * (r2 + T(2.0) * u2) - u;
      const T dv = v * radial + T(2.0) * p[7] * uv + p[6] * (r2 + T(2.0) * v2) - v;
      *x = p[0] * (u + du) + p[2]; *y = p[1] * (v + dv) + p[3];
    } break;
    default: *x = u; *y = v;
  }
}

// (ceres) ceres::QuaternionRotatePoint, include/ceres/rotation.h @2.1: normalises q, then
// UnitQuaternionRotatePoint.
template <typename T>
inline void QuaternionRotatePoint(const T* q, const T* pt, T* result) {
  const T scale = T(1.0) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T u[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
  T uv0 = u[2] * pt[2] - u[3] * pt[1];
  T uv1 = u[3] * pt[0] - u[1] * pt[2];
  T uv2 = u[1] * pt[1] - u[2] * pt[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  result[0] = pt[0] + u[0] * uv0;
  result[1] = pt[1] + u[0] * uv1;
  result[2] = pt[2] + u[0] * uv2;
  result[0] += u[2] * uv2 - u[3] * uv1;
  result[1] += u[3] * uv0 - u[1] * uv2;
  result[2] += u[1] * uv1 - u[2] * uv0;
}

// a6. WorldToPixel — pixsfm/base/src/projection.h:60-75
template <typename T>
inline void WorldToPixel(int model, const T* cam, const T* q, const T* t, const T* X, T* xy) {
  T pr[3];
  QuaternionRotatePoint(q, X, pr);
  pr[0] += t[0]; pr[1] += t[1]; pr[2] += t[2];
  pr[0] /= pr[2];
  pr[1] /= pr[2];
  CameraWorldToImage(model, cam, pr[0], pr[1], &xy[0], &xy[1]);
}

// ---------------------------------------------------------------------------
// (ceres) loss functions, internal/ceres/loss_function.cc, and the Triggs corrector,
// internal/ceres/corrector.cc.
// ---------------------------------------------------------------------------
struct Loss {
  int type = 1;       // pxr_loss_type
  double a = 0.25;
  double weight = 1.0;  // ScaledLoss factor (featuremetric_keypoint_optimizer.h:191-196)
  inline void Evaluate(double s, double rho[3]) const {
    const double kMin = std::numeric_limits<double>::min();
    switch (type) {
      case 0: rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; break;
      case 1: {
        const double b = a * a, c = 1.0 / b;
        const double sum = 1.0 + s * c, inv = 1.0 / sum;
        rho[0] = b * std::log(sum); rho[1] = std::max(kMin, inv); rho[2] = -c * (inv * inv);
      } break;
      case 2: {
        const double b = a * a;
        if (s > b) {
          const double r = std::sqrt(s);
          rho[0] = 2.0 * a * r - b; rho[1] = std::max(kMin, a / r); rho[2] = -rho[1] / (2.0 * s);
        } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
      } break;
      case 3: {
        const double b = a * a, c = 1.0 / b;
        const double sum = 1.0 + s * c, tmp = std::sqrt(sum);
        rho[0] = 2.0 * b * (tmp - 1.0); rho[1] = std::max(kMin, 1.0 / tmp);
        rho[2] = -(c * rho[1]) / (2.0 * sum);
      } break;
      case 4: {
        const double b = 1.0 / (a * a);
        const double sum = 1.0 + s * s * b, inv = 1.0 / sum;
        rho[0] = a * std::atan2(s, a); rho[1] = std::max(kMin, inv); rho[2] = -2.0 * s * b * (inv * inv);
      } break;
      default: rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
    rho[0] *= weight; rho[1] *= weight; rho[2] *= weight;
  }
};

struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) {
      residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  // J <- sqrt_rho1 * (J - alpha_sq_norm * r * (r^T J)), J is nres x ncols row-major
  inline void CorrectJacobian(int nres, int ncols, const double* r, double* J) const {
    if (alpha_sq_norm == 0.0) {
      for (int i = 0; i < nres * ncols; ++i) J[i] *= sqrt_rho1;
      return;
    }
    for (int c = 0; c < ncols; ++c) {
      double rtj = 0;
      for (int k = 0; k < nres; ++k) rtj += J[k * ncols + c] * r[k];
      for (int k = 0; k < nres; ++k)
        J[k * ncols + c] = sqrt_rho1 * (J[k * ncols + c] - alpha_sq_norm * r[k] * rtj);
    }
  }
  inline void CorrectResiduals(int nres, double* r) const {
    for (int i = 0; i < nres; ++i) r[i] *= residual_scaling;
  }
};

}  // namespace orc
