// oracle/orc_refs_graph.h — TEST INFRASTRUCTURE ONLY.
//
// (1) Reference extraction: RobustMeanIRLS (pixsfm/base/src/irls_optim.h:23-71) +
//     ReferenceExtractor::ComputeReference / FillDescriptorTrack / GetVisibleObservations
//     (pixsfm/bundle_adjustment/src/reference_extractor.h:171-318).
// (2) Track / score / root labelling (pixsfm/base/src/graph.cc:126-256) and KA problem
//     packing (pixsfm/keypoint_adjustment/main.py:13-57).  Integer algorithms: the
//     restatement must be bit-exact; graph.cc itself is compiled from /root/reference into
//     oracle/_ref to pin this file (tests/test_oracle_ref_pin.py).
#pragma once
#include <algorithm>
#include <map>
#include <set>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "orc_ba.h"

namespace orc {

// irls_optim.h:23-71. descriptors: n x C row-major. Returns robust mean in `mean`;
// if some weight denominator rho<=0, returns that descriptor (irls_optim.h:63-67).
inline void RobustMeanIRLS(const std::vector<double>& desc, int n, int C, const Loss& loss,
                           int num_iterations, bool l2_normalize, std::vector<double>* mean_out) {
  std::vector<double>& mean = *mean_out;
  mean.assign(C, 0.0);
  std::vector<double> w(n, 1.0);
  for (int k = 0; k < num_iterations; ++k) {
    double wsum = 0;
    for (int i = 0; i < n; ++i) wsum += w[i];
    for (int i = 0; i < n; ++i) w[i] = w[i] / wsum;
    std::fill(mean.begin(), mean.end(), 0.0);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < C; ++c) mean[c] += desc[(size_t)i * C + c] * w[i];
    if (l2_normalize) {
      double nn = 0;
      for (int c = 0; c < C; ++c) nn += mean[c] * mean[c];
      nn = std::sqrt(nn);
      if (nn > 0) for (int c = 0; c < C; ++c) mean[c] /= nn;  // Eigen normalize(): no-op on zero norm
    }
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int c = 0; c < C; ++c) { const double dd = desc[(size_t)i * C + c] - mean[c]; s += dd * dd; }
      double rho[3];
      loss.Evaluate(s, rho);
      if (rho[0] > 0.0) w[i] = 1.0 / rho[0];
      else { for (int c = 0; c < C; ++c) mean[c] = desc[(size_t)i * C + c]; return; }
    }
  }
}

// reference_extractor.h:239-272 for every point (closest_to_robust_mean = true).
inline void ComputeReferences(const pxr_ba_desc& d, const InterpConfig& icfg, const Loss& loss,
                              int iters, double* refs_out, int64_t* src_obs_out) {
  BALayout L = MakeLayout(d);
  const int C = d.channels;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t p = 0; p < d.n_points; ++p) {
    const int n = (int)(L.pt_begin[p + 1] - L.pt_begin[p]);
    if (n == 0) { src_obs_out[p] = -1; continue; }
    std::vector<double> desc((size_t)n * C), dfdr(C), dfdc(C), mean;
    for (int i = 0; i < n; ++i) {
      const int64_t o = L.pt_begin[p] + i;
      const int img = d.obs_img[o];
      const int cam = d.img_cam[img];
      double xy[2], uv[2];
      WorldToPixel<double>(d.cam_model[cam], d.cam_params + (size_t)cam * PXR_MAX_CAM_PARAMS,
                           d.qvec + 4 * img, d.tvec + 3 * img, d.xyz + 3 * p, xy);
      const Patch patch = MakePatch(d, o);
      ToPixelCoordinates<double>(patch, xy, uv);
      PixelInterp(patch, icfg, uv[1], uv[0], &desc[(size_t)i * C], dfdr.data(), dfdc.data());
    }
    RobustMeanIRLS(desc, n, C, loss, iters, icfg.l2_normalize, &mean);
    int ref_idx = 0;
    double best = 0;
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int c = 0; c < C; ++c) { const double dd = desc[(size_t)i * C + c] - mean[c]; s += dd * dd; }
      if (i == 0 || s < best) { best = s; ref_idx = i; }  // minCoeff: first minimum
    }
    for (int c = 0; c < C; ++c) refs_out[(size_t)p * C + c] = desc[(size_t)ref_idx * C + c];
    src_obs_out[p] = L.pt_begin[p] + ref_idx;
  }
}

// ---- graph.cc:208-223 union_find_get_root
inline int64_t UFRoot(int64_t i, std::vector<int64_t>& parent) {
  if (parent[i] == -1) return i;
  parent[i] = UFRoot(parent[i], parent);
  return parent[i];
}

// graph.cc:225-256... (ComputeTrackLabels :126-206)
inline void TrackLabels(int64_t n_nodes, const int32_t* node_image, int64_t n_edges,
                        const int64_t* es, const int64_t* ed, const double* sim, int64_t* labels) {
  typedef std::tuple<double, size_t, size_t> ET;
  std::vector<ET> edges;
  edges.reserve(n_edges);
  for (int64_t e = 0; e < n_edges; ++e) edges.push_back(std::make_tuple(sim[e], (size_t)es[e], (size_t)ed[e]));
  std::sort(edges.begin(), edges.end());
  std::reverse(edges.begin(), edges.end());
  std::vector<int64_t> parent(n_nodes, -1);
  std::vector<std::set<int32_t>> images(n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i) images[i].insert(node_image[i]);
  for (auto& it : edges) {
    const int64_t r1 = UFRoot((int64_t)std::get<1>(it), parent);
    const int64_t r2 = UFRoot((int64_t)std::get<2>(it), parent);
    if (r1 == r2) continue;
    bool intersects = false;
    {
      auto a = images[r1].begin(); auto b = images[r2].begin();
      while (a != images[r1].end() && b != images[r2].end()) {
        if (*a < *b) ++a; else if (*b < *a) ++b; else { intersects = true; break; }
      }
    }
    if (intersects) continue;
    if (images[r1].size() < images[r2].size()) {
      parent[r1] = r2;
      images[r2].insert(images[r1].begin(), images[r1].end());
      images[r1].clear();
    } else {
      parent[r2] = r1;
      images[r1].insert(images[r2].begin(), images[r2].end());
      images[r2].clear();
    }
  }
  int64_t n_tracks = 0;
  for (int64_t i = 0; i < n_nodes; ++i) labels[i] = -1;
  for (int64_t i = 0; i < n_nodes; ++i) if (parent[i] == -1) labels[i] = n_tracks++;
  for (int64_t i = 0; i < n_nodes; ++i) if (labels[i] == -1) labels[i] = labels[UFRoot(i, parent)];
}

// graph.cc ComputeScoreLabels. Edge order must be the out_matches traversal order
// (node by node, match by match) because of floating-point accumulation order.
inline void ScoreLabels(int64_t n_nodes, int64_t n_edges, const int64_t* es, const int64_t* ed,
                        const double* sim, const int64_t* labels, double* scores) {
  for (int64_t i = 0; i < n_nodes; ++i) scores[i] = 0.0;
  for (int64_t e = 0; e < n_edges; ++e)
    if (labels[es[e]] == labels[ed[e]]) { scores[es[e]] += sim[e]; scores[ed[e]] += sim[e]; }
}

// graph.cc ComputeRootLabels
inline void RootLabels(int64_t n_nodes, const int64_t* labels, const double* scores, uint8_t* is_root) {
  int64_t n_tracks = 0;
  for (int64_t i = 0; i < n_nodes; ++i) n_tracks = std::max(n_tracks, labels[i] + 1);
  std::vector<std::pair<double, size_t>> sc;
  for (int64_t i = 0; i < n_nodes; ++i) sc.push_back(std::make_pair(scores[i], (size_t)i));
  std::sort(sc.begin(), sc.end());
  std::reverse(sc.begin(), sc.end());
  std::vector<bool> has_root(n_tracks, false);
  for (int64_t i = 0; i < n_nodes; ++i) is_root[i] = 0;
  for (auto& it : sc) {
    const size_t n = it.second;
    if (has_root[labels[n]]) continue;
    is_root[n] = 1; has_root[labels[n]] = true;
  }
}

// keypoint_adjustment/main.py:13-57 find_problem_labels (track_edge_counts=None).
// Counter.most_common(): sorted by count descending, stable w.r.t. first-appearance order.
inline int KAProblemLabels(int64_t n_nodes, const int64_t* labels, int max_per_problem, int32_t* out) {
  std::vector<int64_t> order;  // first-appearance order of track labels
  std::unordered_map<int64_t, int64_t> count;
  for (int64_t i = 0; i < n_nodes; ++i) {
    auto it = count.find(labels[i]);
    if (it == count.end()) { count[labels[i]] = 1; order.push_back(labels[i]); } else it->second++;
  }
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return count[a] > count[b]; });
  if (max_per_problem == -1) { int64_t m = 0; for (auto& kv : count) m = std::max(m, kv.second); max_per_problem = (int)m; }
  std::vector<int64_t> bins;
  int64_t n_labels = (int64_t)count.size();
  std::vector<int32_t> t2p(n_labels, -1);  // python: [-1]*len(track_count), indexed by label
  size_t start = 0;
  int64_t last_v = std::numeric_limits<int64_t>::max();
  for (int64_t k : order) {
    const int64_t v = count[k];
    if (v < last_v) { start = 0; last_v = v; }
    bool found = false;
    if (v < max_per_problem) {
      for (size_t i = start; i < bins.size(); ++i) {
        if (bins[i] + v <= max_per_problem) { bins[i] += v; t2p[k] = (int32_t)i; found = true; start = i; break; }
      }
    }
    if (!found) { t2p[k] = (int32_t)bins.size(); start = bins.size(); bins.push_back(v); }
  }
  for (int64_t i = 0; i < n_nodes; ++i) out[i] = t2p[labels[i]];
  return (int)bins.size();
}

}  // namespace orc
