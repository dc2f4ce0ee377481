// oracle/orc_ka.h — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of featuremetric keypoint adjustment:
//   residual : FeatureMetric2DCostFunctor (pixsfm/residuals/src/featuremetric.h:24-69)
//              r = interp(patch1, kp1) - interp(patch2, kp2), AutoDiff<..., C, 2, 2>
//   loss     : ceres::ScaledLoss(loss, weight) per edge
//              (pixsfm/keypoint_adjustment/src/featuremetric_keypoint_optimizer.h:191-196)
//   bounds   : KeypointOptimizerBase::ParameterizeKeypoints
//              (pixsfm/keypoint_adjustment/src/keypoint_optimizer.h:110-157)
//   solve    : ceres::Solve, SPARSE_NORMAL_CHOLESKY (exact), bounded trust region with
//              projected line search (keypoint_optimizer.h:77-104) -> orc_trust_region.h
//   fan-out  : one independent problem per label (base/src/parallel_optimizer.h:77-211)
#pragma once
#include <unordered_map>
#include <vector>

#include "orc_ba.h"

namespace orc {

inline Patch MakeKAPatch(const pxr_ka_desc& d, int64_t kp) {
  const int64_t pi = d.kp_patch ? d.kp_patch[kp] : kp;
  Patch p;
  const size_t esz = d.patch_dtype == PXR_F16 ? 2 : (d.patch_dtype == PXR_F32 ? 4 : 8);
  if (d.n_patch_blocks > 0) {   // host blocks laid out back to back (device blocks are not readable here)
    int64_t local = pi;
    int b = 0;
    while (b < d.n_patch_blocks - 1 && local >= d.patch_block_counts[b]) { local -= d.patch_block_counts[b]; ++b; }
    p.data = (const char*)d.patch_block_ptrs[b] + (size_t)local * d.ph * d.pw * d.channels * esz;
  } else {
    p.data = (const char*)d.patches + (size_t)pi * d.ph * d.pw * d.channels * esz;
  }
  p.dtype = d.patch_dtype; p.h = d.ph; p.w = d.pw; p.c = d.channels;
  p.corner[0] = d.corner[2 * pi]; p.corner[1] = d.corner[2 * pi + 1];
  p.scale[0] = d.scale[2 * pi]; p.scale[1] = d.scale[2 * pi + 1];
  p.upsampling = d.upsampling_factor;
  return p;
}

// descriptor + d(descriptor)/d(xy) at an image-space keypoint (PatchInterpolator::Evaluate with
// Jets of size 2: features/src/patch_interpolator.h:125-135)
inline void KAInterp(const Patch& patch, const InterpConfig& ic, const double* xy, double* f,
                     double* dfdx, double* dfdy, std::vector<double>& tmp) {
  const int C = patch.c;
  tmp.resize(2 * C);
  double uv[2];
  ToPixelCoordinates<double>(patch, xy, uv);
  double* dfdr = tmp.data(); double* dfdc = dfdr + C;
  PixelInterp(patch, ic, uv[1], uv[0], f, dfdr, dfdc);
  if (dfdx) {
    const double ax = patch.scale[0] * patch.upsampling, ay = patch.scale[1] * patch.upsampling;
    for (int i = 0; i < C; ++i) { dfdx[i] = dfdc[i] * ax; dfdy[i] = dfdr[i] * ay; }
  }
}

class KAProblemEvaluator : public TREvaluator {
 public:
  const pxr_ka_desc& d;
  InterpConfig ic;
  Loss base_loss;
  std::vector<int64_t> edges;               // edge ids of this problem
  std::vector<int64_t> var_kps;             // variable keypoints (global ids)
  std::unordered_map<int64_t, int> var_idx; // global kp -> local variable index
  std::vector<double> lower, upper;         // bounds per variable coordinate (2 per kp)
  bool constrained = false;
  std::vector<double> H, g;
  double fixed_cost = 0;

  KAProblemEvaluator(const pxr_ka_desc& desc, const InterpConfig& icfg, const Loss& l)
      : d(desc), ic(icfg), base_loss(l) {}

  void Finalize() {
    // will_be_optimized_: keypoints touched by at least one residual block
    for (int64_t e : edges) {
      for (int q = 0; q < (d.ref_desc ? 1 : 2); ++q) {    // query mode: edge_dst indexes ref_desc
        const int64_t kp = q == 0 ? d.edge_src[e] : d.edge_dst[e];
        if (d.kp_const[kp]) continue;
        if (var_idx.find(kp) == var_idx.end()) { var_idx[kp] = (int)var_kps.size(); var_kps.push_back(kp); }
      }
    }
    const int nv = (int)var_kps.size();
    lower.assign(2 * nv, -std::numeric_limits<double>::max());
    upper.assign(2 * nv, std::numeric_limits<double>::max());
    if (d.bound > 0.0 || d.patches_are_sparse) {
      constrained = nv > 0;
      for (int i = 0; i < nv; ++i) {
        const int64_t kp = var_kps[i];
        const Patch p = MakeKAPatch(d, kp);
        const double* k = d.keypoints + 2 * kp;
        const double dx = p.w / p.scale[0], dy = p.h / p.scale[1];
        double lowerx = (p.corner[0] + 0.5) / p.scale[0], lowery = (p.corner[1] + 0.5) / p.scale[1];
        double upperx = lowerx + dx, uppery = lowery + dy;
        if (d.bound > 0.0) {
          upperx = std::min(k[0] + d.bound / p.scale[0], upperx);
          uppery = std::min(k[1] + d.bound / p.scale[1], uppery);
          lowerx = std::max(k[0] - d.bound / p.scale[0], lowerx);
          lowery = std::max(k[1] - d.bound / p.scale[1], lowery);
        }
        lower[2 * i] = lowerx; lower[2 * i + 1] = lowery;
        upper[2 * i] = upperx; upper[2 * i + 1] = uppery;
      }
    }
    H.assign((size_t)4 * nv * nv, 0.0);
    g.assign(2 * nv, 0.0);
  }
  int NumParameters() const override { return 2 * (int)var_kps.size(); }
  int NumLocal() const override { return 2 * (int)var_kps.size(); }
  bool IsConstrained() const override { return constrained; }
  void Pack(double* x) const {
    for (size_t i = 0; i < var_kps.size(); ++i) { x[2 * i] = d.keypoints[2 * var_kps[i]]; x[2 * i + 1] = d.keypoints[2 * var_kps[i] + 1]; }
  }
  void Unpack(const double* x) const {
    for (size_t i = 0; i < var_kps.size(); ++i) { d.keypoints[2 * var_kps[i]] = x[2 * i]; d.keypoints[2 * var_kps[i] + 1] = x[2 * i + 1]; }
  }
  const double* KP(const double* x, int64_t kp, int* vi) const {
    auto it = var_idx.find(kp);
    if (it == var_idx.end()) { *vi = -1; return d.keypoints + 2 * kp; }
    *vi = it->second;
    return x + 2 * it->second;
  }
  bool Evaluate(const double* x, double* cost, bool with_jac) override {
    const int C = d.channels, n = NumLocal();
    if (with_jac) { std::fill(H.begin(), H.end(), 0.0); std::fill(g.begin(), g.end(), 0.0); }
    std::vector<double> f1(C), f2(C), a1(C), b1(C), a2(C), b2(C), r(C), tmp, J(4 * (size_t)C);
    double total = 0;
    for (int64_t e : edges) {
      const int64_t k1 = d.edge_src[e], k2 = d.edge_dst[e];
      if (!d.ref_desc && k1 == k2) continue;  // "Avoid optimizing a keypoint to itself" (topological_keypoint_optimizer.h:139-143)
      int v1, v2 = -1;
      const double* x1 = KP(x, k1, &v1);
      KAInterp(MakeKAPatch(d, k1), ic, x1, f1.data(), with_jac ? a1.data() : nullptr, b1.data(), tmp);
      if (d.ref_desc) {
        // FeatureReference2DCostFunctor (residuals/src/feature_reference.h:44-59): target - fixed reference
        for (int i = 0; i < C; ++i) { f2[i] = d.ref_desc[(size_t)k2 * C + i]; a2[i] = 0; b2[i] = 0; }
      } else {
        const double* x2 = KP(x, k2, &v2);
        KAInterp(MakeKAPatch(d, k2), ic, x2, f2.data(), with_jac ? a2.data() : nullptr, b2.data(), tmp);
      }
      double s = 0;
      for (int i = 0; i < C; ++i) { r[i] = f1[i] - f2[i]; s += r[i] * r[i]; }
      Loss l = base_loss; l.weight = d.edge_weight ? d.edge_weight[e] : 1.0;
      double rho[3];
      l.Evaluate(s, rho);
      total += 0.5 * rho[0];
      if (!with_jac) continue;
      for (int i = 0; i < C; ++i) { J[4 * i] = a1[i]; J[4 * i + 1] = b1[i]; J[4 * i + 2] = -a2[i]; J[4 * i + 3] = -b2[i]; }
      Corrector corr(s, rho);
      corr.CorrectJacobian(C, 4, r.data(), J.data());
      corr.CorrectResiduals(C, r.data());
      const int col[4] = {v1 >= 0 ? 2 * v1 : -1, v1 >= 0 ? 2 * v1 + 1 : -1, v2 >= 0 ? 2 * v2 : -1, v2 >= 0 ? 2 * v2 + 1 : -1};
      for (int a = 0; a < 4; ++a) {
        if (col[a] < 0) continue;
        double ga = 0;
        for (int i = 0; i < C; ++i) ga += J[4 * i + a] * r[i];
        g[col[a]] += ga;
        for (int b = 0; b < 4; ++b) {
          if (col[b] < 0) continue;
          double v = 0;
          for (int i = 0; i < C; ++i) v += J[4 * i + a] * J[4 * i + b];
          H[(size_t)col[a] * n + col[b]] += v;
        }
      }
    }
    *cost = total;
    return std::isfinite(total);
  }
  void Gradient(double* gg) const override { for (size_t i = 0; i < g.size(); ++i) gg[i] = g[i]; }
  void SquaredColumnNorm(double* dd) const override { const int n = NumLocal(); for (int i = 0; i < n; ++i) dd[i] = H[(size_t)i * n + i]; }
  bool SolveDamped(const double* D2, double* delta, int* iters) override {
    const int n = NumLocal();
    std::vector<double> A(H), b(n);
    for (int i = 0; i < n; ++i) { A[(size_t)i * n + i] += D2[i]; b[i] = -g[i]; }
    if (!CholeskySolveInPlace(n, A, b)) return false;
    for (int i = 0; i < n; ++i) delta[i] = b[i];
    *iters = 1;
    return true;
  }
  double ModelCostChange(const double* dl) const override {
    const int n = NumLocal();
    double gd = 0, dHd = 0;
    for (int i = 0; i < n; ++i) { gd += g[i] * dl[i]; double row = 0; for (int j = 0; j < n; ++j) row += H[(size_t)i * n + j] * dl[j]; dHd += dl[i] * row; }
    return -gd - 0.5 * dHd;
  }
  void Plus(const double* x, const double* dl, double* xp) const override {
    // ParameterBlock::Plus: x + delta then projection onto the box
    const int n = NumLocal();
    for (int i = 0; i < n; ++i) xp[i] = std::min(std::max(x[i] + dl[i], lower[i]), upper[i]);
  }
};

inline std::vector<std::vector<int64_t>> GroupEdgesByProblem(const pxr_ka_desc& d) {
  std::vector<std::vector<int64_t>> groups(d.n_problems);
  for (int64_t e = 0; e < d.n_edges; ++e) groups[d.edge_problem ? d.edge_problem[e] : 0].push_back(e);
  return groups;
}

inline int KAEvaluateAll(const pxr_ka_desc& d, const InterpConfig& ic, const pxr_solver_options& so,
                         double* sq_norm, double* cost) {
  const int C = d.channels;
  double total = 0;
#pragma omp parallel for reduction(+ : total) schedule(dynamic, 64)
  for (int64_t e = 0; e < d.n_edges; ++e) {
    std::vector<double> f1(C), f2(C), tmp;
    const int64_t k1 = d.edge_src[e], k2 = d.edge_dst[e];
    if (!d.ref_desc && k1 == k2) { if (sq_norm) sq_norm[e] = 0; continue; }
    KAInterp(MakeKAPatch(d, k1), ic, d.keypoints + 2 * k1, f1.data(), nullptr, nullptr, tmp);
    if (d.ref_desc) for (int i = 0; i < C; ++i) f2[i] = d.ref_desc[(size_t)k2 * C + i];
    else KAInterp(MakeKAPatch(d, k2), ic, d.keypoints + 2 * k2, f2.data(), nullptr, nullptr, tmp);
    double s = 0;
    for (int i = 0; i < C; ++i) { const double r = f1[i] - f2[i]; s += r * r; }
    if (sq_norm) sq_norm[e] = s;
    Loss l; l.type = so.loss_type; l.a = so.loss_scale; l.weight = d.edge_weight ? d.edge_weight[e] : 1.0;
    double rho[3];
    l.Evaluate(s, rho);
    total += 0.5 * rho[0];
  }
  if (cost) *cost = total;
  return 0;
}

inline void KASolveAll(const pxr_ka_desc& d, const InterpConfig& ic, const pxr_solver_options& so, TRSummary* acc) {
  auto groups = GroupEdgesByProblem(d);
  Loss l; l.type = so.loss_type; l.a = so.loss_scale;
  TROptions to = ToTROptions(so);
  to.use_inner_iterations = false;
  double init = 0, fin = 0; int ns = 0, nu = 0;
  // keypoints of different problems are disjoint; results are written after each solve
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : init, fin, ns, nu)
  for (int p = 0; p < d.n_problems; ++p) {
    if (groups[p].empty()) continue;
    KAProblemEvaluator ev(d, ic, l);
    ev.edges = groups[p];
    ev.Finalize();
    if (ev.NumLocal() == 0) {
      double c; std::vector<double> x0(1); ev.Evaluate(x0.data(), &c, false); init += c; fin += c; continue;
    }
    std::vector<double> x(ev.NumParameters());
    ev.Pack(x.data());
    TrustRegionMinimizer tr(to);
    TRSummary s;
    tr.Minimize(&ev, x.data(), &s);
    ev.Unpack(x.data());
    init += s.initial_cost; fin += s.final_cost; ns += s.num_successful_steps; nu += s.num_unsuccessful_steps;
  }
  acc->initial_cost = init; acc->final_cost = fin; acc->num_successful_steps = ns; acc->num_unsuccessful_steps = nu;
}

}  // namespace orc
