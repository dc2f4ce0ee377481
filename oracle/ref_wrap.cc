// oracle/ref_wrap.cc — TEST INFRASTRUCTURE ONLY.
// Thin C wrapper around TWO reference files compiled verbatim from where they lie under
// /root/reference (never copied into this repo):
//   pixsfm/base/src/cubic_hermite_spline_simd.h   (AVX2/FMA/F16C spline, a1)
//   pixsfm/base/src/graph.cc                      (track/score/root labelling, a13)
// Output: oracle/_ref/libpxref.so (git-ignored, travels with gpurun).  Used to pin the
// restatement in orc_core.h / orc_refs_graph.h and as the inner kernel of the CPU baseline.
#define AVX2_ENABLED 1
#include "base/src/cubic_hermite_spline_simd.h"
#include "base/src/graph.h"

#include <cstdint>
#include <vector>

template <int C, typename IN>
static void SplineT(const IN* p0, const IN* p1, const IN* p2, const IN* p3, double x, double* f, double* d) {
  pixsfm::CubicHermiteSplineSIMD<C>(p0, p1, p2, p3, x, f, d);
}
template <typename IN>
static int SplineDispatch(int C, const IN* p0, const IN* p1, const IN* p2, const IN* p3, double x, double* f, double* d) {
  switch (C) {
    case 8: SplineT<8>(p0, p1, p2, p3, x, f, d); return 0;
    case 16: SplineT<16>(p0, p1, p2, p3, x, f, d); return 0;
    case 32: SplineT<32>(p0, p1, p2, p3, x, f, d); return 0;
    case 64: SplineT<64>(p0, p1, p2, p3, x, f, d); return 0;
    case 128: SplineT<128>(p0, p1, p2, p3, x, f, d); return 0;
    case 256: SplineT<256>(p0, p1, p2, p3, x, f, d); return 0;
  }
  return 1;
}

extern "C" {
int ref_spline_f16(int C, const uint16_t* p0, const uint16_t* p1, const uint16_t* p2, const uint16_t* p3,
                   double x, double* f, double* dfdx) {
  return SplineDispatch<half>(C, (const half*)p0, (const half*)p1, (const half*)p2, (const half*)p3, x, f, dfdx);
}
int ref_spline_f32(int C, const float* p0, const float* p1, const float* p2, const float* p3, double x,
                   double* f, double* dfdx) {
  return SplineDispatch<float>(C, p0, p1, p2, p3, x, f, dfdx);
}
int ref_spline_f64(int C, const double* p0, const double* p1, const double* p2, const double* p3, double x,
                   double* f, double* dfdx) {
  return SplineDispatch<double>(C, p0, p1, p2, p3, x, f, dfdx);
}

// graph.cc: build a pixsfm::Graph from flat arrays and run the three labelling passes.
int ref_graph_labels(int64_t n_nodes, const int32_t* node_image, const int32_t* node_feature,
                     int64_t n_edges, const int64_t* es, const int64_t* ed, const double* sim,
                     int64_t* track_labels, double* scores, uint8_t* is_root) {
  pixsfm::Graph g;
  for (int64_t i = 0; i < n_nodes; ++i) g.AddNode((colmap::image_t)node_image[i], (colmap::point2D_t)node_feature[i]);
  for (int64_t e = 0; e < n_edges; ++e) g.AddEdge(g.nodes[es[e]], g.nodes[ed[e]], sim[e]);
  std::vector<size_t> tl = pixsfm::ComputeTrackLabels(g);
  std::vector<double> sc = pixsfm::ComputeScoreLabels(g, tl);
  std::vector<bool> rt = pixsfm::ComputeRootLabels(g, tl, sc);
  for (int64_t i = 0; i < n_nodes; ++i) { track_labels[i] = (int64_t)tl[i]; scores[i] = sc[i]; is_root[i] = rt[i] ? 1 : 0; }
  return 0;
}
}
