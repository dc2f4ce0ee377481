// pxr_problem.cu — BundleOptimizer::SetUp + Parameterize in C++ (host code only; a .cu file so that one compiler line
// builds the library).
//
// Reference: pixsfm/bundle_adjustment/src/bundle_optimizer.h
//   :139-165  SetUp           images of the setup, then the setup's points, then Parameterize
//   :247-275  AddImageToProblem     every 2D point with a 3D point (track length >= min_track_length) -> one residual
//   :277-313  AddPointToProblem     observations of a setup point in images OUTSIDE the setup -> constant-pose residuals,
//                                   cameras seen only that way become constant
//   :315-331  RegisterPoint3DObservation   linear search of the track for (image_id, point2D_idx)
//   :335-358  ParameterizePoints    constant when fewer of its track elements take part than min(track, min_track_length)
//   :360-398  ParameterizeImages    constant pose: !refine_extrinsics | setup.HasConstantPose | image outside the setup
//   :400-442  ParameterizeCameras   constant camera or the focal / principal point / extra parameter subsets
// and pixsfm/bundle_adjustment/src/reference_extractor.h:171-205 (GetVisibleObservations: every track element that has
// a feature patch, by ascending point id) for the reference extractor's observation list.
//
// The reference walks colmap::Reconstruction; this walks a structure-of-arrays view of the same data
// (pxr_recon_view) and produces the flat problem IR of pxr_ba_desc: observations sorted by point, index maps, masks.
// Integer logic only — bit-exact against the Python restatement it replaces (tests/test_problem_builder.py).
#include <algorithm>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "pxr_internal.h"

struct pxr_problem_ir {
  std::vector<int64_t> obs_point3D_id, obs_image_id, obs_point2D_idx, obs_pt;
  std::vector<int32_t> obs_img, img_cam;
  std::vector<int64_t> image_ids, camera_ids, point_ids;
  std::vector<uint8_t> pose_const, tvec_const_mask, point_const;
  std::vector<uint32_t> cam_const_mask;
};

namespace {

// (focal, principal point, extra) parameter-index bit masks per COLMAP camera model 0..6 (colmap camera_models.h
// focal_length_idxs / principal_point_idxs / extra_params_idxs)
const uint32_t kParamGroups[7][3] = {{0x1, 0x6, 0x0}, {0x3, 0xC, 0x0}, {0x1, 0x6, 0x8}, {0x1, 0x6, 0x18},
                                     {0x3, 0xC, 0xF0}, {0x3, 0xC, 0xF0}, {0x3, 0xC, 0xFF0}};

template <typename T>
std::unordered_map<int64_t, int64_t> index_of(const T* ids, int64_t n) {
  std::unordered_map<int64_t, int64_t> m;
  m.reserve((size_t)n * 2 + 1);
  for (int64_t i = 0; i < n; ++i) m.emplace((int64_t)ids[i], i);
  return m;
}

std::vector<int64_t> sorted_copy(const int64_t* p, int64_t n) {
  std::vector<int64_t> v(p, p + n);
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  return v;
}

}  // namespace

extern "C" {

int pxr_problem_build(const pxr_recon_view* rec, const pxr_ba_setup_view* setup, const pxr_ba_build_options* opt,
                      pxr_problem_ir** out) {
  using pxr::fail;
  if (!rec || !opt || !out) return fail(PXR_ERR_INVALID_ARGUMENT, "pxr_problem_build: NULL argument");
  const bool refs_mode = opt->mode == 1;
  if (!refs_mode && !setup) return fail(PXR_ERR_INVALID_ARGUMENT, "pxr_problem_build: a setup is required in BA mode");
  const auto img_index = index_of(rec->image_id, rec->n_images);
  const auto cam_index = index_of(rec->camera_id, rec->n_cameras);
  const auto pt_index = index_of(rec->point3D_id, rec->n_points);
  auto track_len = [&](int64_t pi) { return rec->track_begin[pi + 1] - rec->track_begin[pi]; };

  std::vector<int64_t> o_pid, o_img, o_p2d;            // in enumeration order
  std::unique_ptr<pxr_problem_ir> ir(new pxr_problem_ir());

  // per-image / per-camera residual counters and per-point registered track elements
  std::vector<int64_t> image_num_residuals((size_t)rec->n_images, 0), camera_num_residuals((size_t)rec->n_cameras, 0);
  std::vector<uint8_t> image_touched((size_t)rec->n_images, 0);
  std::vector<uint8_t> track_registered(refs_mode ? 0 : (size_t)rec->track_begin[rec->n_points], 0);
  std::vector<int64_t> n_registered((size_t)rec->n_points, 0);
  std::vector<uint8_t> point_in_problem((size_t)rec->n_points, 0);
  std::vector<uint8_t> extra_const_camera((size_t)rec->n_cameras, 0);

  std::unordered_set<int64_t> setup_images, const_poses, const_cameras, const_points;
  std::unordered_map<int64_t, uint8_t> const_tvecs;
  if (setup) {
    setup_images.insert(setup->image_ids, setup->image_ids + setup->n_images);
    const_poses.insert(setup->const_pose_ids, setup->const_pose_ids + setup->n_const_poses);
    const_cameras.insert(setup->const_camera_ids, setup->const_camera_ids + setup->n_const_cameras);
    const_points.insert(setup->const_point_ids, setup->const_point_ids + setup->n_const_points);
    for (int64_t i = 0; i < setup->n_const_tvecs; ++i) const_tvecs[setup->const_tvec_ids[i]] = setup->const_tvec_masks[i];
  }

  if (!refs_mode) {
    int rc_err = PXR_OK;
    auto add_residual = [&](int64_t ii /*image index*/, int64_t p2d_idx) {
      const int64_t pid = rec->p2d_point3D_id[rec->p2d_begin[ii] + p2d_idx];
      if (pid < 0) return;
      const int64_t image_id = rec->image_id[ii];
      auto pit = pt_index.find(pid);
      if (pit == pt_index.end()) { rc_err = fail(PXR_ERR_INVALID_ARGUMENT, "2D point refers to the unknown 3D point %lld", (long long)pid); return; }
      const int64_t pi = pit->second;
      o_pid.push_back(pid); o_img.push_back(image_id); o_p2d.push_back(p2d_idx);
      const bool constant_pose = !opt->refine_extrinsics || const_poses.count(image_id) > 0;
      if (!constant_pose) image_num_residuals[ii]++;
      image_touched[ii] = 1;
      // RegisterPoint3DObservation: first track element equal to (image_id, point2D_idx)
      bool found = false;
      for (int64_t k = rec->track_begin[pi]; k < rec->track_begin[pi + 1]; ++k)
        if (rec->track_image_id[k] == image_id && rec->track_point2D_idx[k] == p2d_idx) {
          if (!track_registered[k]) { track_registered[k] = 1; n_registered[pi]++; }
          found = true;
          break;
        }
      if (!found) { rc_err = fail(PXR_ERR_INVALID_ARGUMENT, "Failed to register track element."); return; }
      point_in_problem[pi] = 1;
      auto cit = cam_index.find(rec->image_camera_id[ii]);
      if (cit == cam_index.end()) { rc_err = fail(PXR_ERR_INVALID_ARGUMENT, "image %lld refers to an unknown camera", (long long)image_id); return; }
      camera_num_residuals[cit->second]++;
    };
    // AddImageToProblem, ascending image id
    for (int64_t image_id : sorted_copy(setup->image_ids, setup->n_images)) {
      auto it = img_index.find(image_id);
      if (it == img_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "setup image %lld is not in the reconstruction", (long long)image_id);
      const int64_t ii = it->second;
      const int64_t np2d = rec->p2d_begin[ii + 1] - rec->p2d_begin[ii];
      for (int64_t k = 0; k < np2d && rc_err == PXR_OK; ++k) {
        const int64_t pid = rec->p2d_point3D_id[rec->p2d_begin[ii] + k];
        if (pid < 0) continue;
        auto pit = pt_index.find(pid);
        if (pit == pt_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "2D point refers to the unknown 3D point %lld", (long long)pid);
        if (track_len(pit->second) < (int64_t)opt->min_track_length) continue;
        add_residual(ii, k);
      }
      if (rc_err != PXR_OK) return rc_err;
    }
    // AddPointToProblem: variable points, then constant points, each ascending
    std::vector<int64_t> pts = sorted_copy(setup->var_point_ids, setup->n_var_points);
    const std::vector<int64_t> cpts = sorted_copy(setup->const_point_ids, setup->n_const_points);
    pts.insert(pts.end(), cpts.begin(), cpts.end());
    for (int64_t pid : pts) {
      auto pit = pt_index.find(pid);
      if (pit == pt_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "setup point %lld is not in the reconstruction", (long long)pid);
      const int64_t pi = pit->second;
      if (n_registered[pi] == track_len(pi)) continue;
      for (int64_t k = rec->track_begin[pi]; k < rec->track_begin[pi + 1]; ++k) {
        const int64_t image_id = rec->track_image_id[k];
        if (setup_images.count(image_id)) continue;
        auto it = img_index.find(image_id);
        if (it == img_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "track element refers to the unknown image %lld", (long long)image_id);
        auto cit = cam_index.find(rec->image_camera_id[it->second]);
        if (cit == cam_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "image %lld refers to an unknown camera", (long long)image_id);
        if (camera_num_residuals[cit->second] == 0) extra_const_camera[cit->second] = 1;
        add_residual(it->second, rec->track_point2D_idx[k]);
        if (rc_err != PXR_OK) return rc_err;
      }
    }
  } else {
    // GetVisibleObservations: ascending point id, track order, only elements that have a feature patch
    std::vector<int64_t> ids = sorted_copy(opt->ref_point_ids, opt->n_ref_points);
    for (int64_t pid : ids) {
      auto pit = pt_index.find(pid);
      if (pit == pt_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "point %lld is not in the reconstruction", (long long)pid);
      const int64_t pi = pit->second;
      for (int64_t k = rec->track_begin[pi]; k < rec->track_begin[pi + 1]; ++k) {
        if (opt->track_has_patch && !opt->track_has_patch[k]) continue;
        o_pid.push_back(pid); o_img.push_back(rec->track_image_id[k]); o_p2d.push_back(rec->track_point2D_idx[k]);
      }
    }
  }

  // ---- canonical order: stable sort by point id; dense index maps
  const int64_t n_obs = (int64_t)o_pid.size();
  std::vector<int64_t> order((size_t)n_obs);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return o_pid[a] < o_pid[b]; });
  ir->obs_point3D_id.resize(n_obs); ir->obs_image_id.resize(n_obs); ir->obs_point2D_idx.resize(n_obs);
  for (int64_t k = 0; k < n_obs; ++k) {
    ir->obs_point3D_id[k] = o_pid[order[k]]; ir->obs_image_id[k] = o_img[order[k]]; ir->obs_point2D_idx[k] = o_p2d[order[k]];
  }
  ir->point_ids = sorted_copy(o_pid.data(), n_obs);
  ir->image_ids = sorted_copy(o_img.data(), n_obs);
  {
    std::vector<int64_t> cams;
    for (int64_t image_id : ir->image_ids) {
      auto it = img_index.find(image_id);
      if (it == img_index.end()) return fail(PXR_ERR_INVALID_ARGUMENT, "observation in the unknown image %lld", (long long)image_id);
      cams.push_back(rec->image_camera_id[it->second]);
    }
    ir->camera_ids = sorted_copy(cams.data(), (int64_t)cams.size());
  }
  const auto pidx = index_of(ir->point_ids.data(), (int64_t)ir->point_ids.size());
  const auto iidx = index_of(ir->image_ids.data(), (int64_t)ir->image_ids.size());
  const auto cidx = index_of(ir->camera_ids.data(), (int64_t)ir->camera_ids.size());
  ir->obs_img.resize(n_obs); ir->obs_pt.resize(n_obs);
  for (int64_t k = 0; k < n_obs; ++k) {
    ir->obs_img[k] = (int32_t)iidx.at(ir->obs_image_id[k]);
    ir->obs_pt[k] = pidx.at(ir->obs_point3D_id[k]);
  }
  const size_t n_img = ir->image_ids.size(), n_cam = ir->camera_ids.size(), n_pts = ir->point_ids.size();
  ir->img_cam.resize(n_img);
  for (size_t i = 0; i < n_img; ++i) ir->img_cam[i] = (int32_t)cidx.at(rec->image_camera_id[img_index.at(ir->image_ids[i])]);
  ir->pose_const.assign(n_img, 1); ir->tvec_const_mask.assign(n_img, 0); ir->point_const.assign(n_pts, 0);
  ir->cam_const_mask.assign(n_cam, 0xFFFFFFFFu);
  if (!refs_mode) {
    // ParameterizeImages
    for (int64_t ii = 0; ii < rec->n_images; ++ii) {
      if (image_num_residuals[ii] <= 0) continue;
      const int64_t image_id = rec->image_id[ii];
      const bool constant_pose = !opt->refine_extrinsics || const_poses.count(image_id) > 0 || setup_images.count(image_id) == 0;
      if (constant_pose) continue;
      const int64_t k = iidx.at(image_id);
      ir->pose_const[k] = 0;
      auto tv = const_tvecs.find(image_id);
      if (tv != const_tvecs.end()) ir->tvec_const_mask[k] = (uint8_t)(tv->second & 7u);
    }
    // ParameterizeCameras
    const bool constant_camera = !(opt->refine_focal_length || opt->refine_principal_point || opt->refine_extra_params);
    for (int64_t ci = 0; ci < rec->n_cameras; ++ci) {
      if (camera_num_residuals[ci] <= 0) continue;
      const int64_t camera_id = rec->camera_id[ci];
      auto ck = cidx.find(camera_id);
      if (ck == cidx.end()) continue;
      uint32_t mask = 0xFFFFFFFFu;
      if (!extra_const_camera[ci] && !constant_camera && const_cameras.count(camera_id) == 0) {
        const int model = rec->camera_model[ci];
        if (model < 0 || model > 6) return fail(PXR_ERR_UNSUPPORTED, "camera model %d", model);
        mask = 0;
        if (!opt->refine_focal_length) mask |= kParamGroups[model][0];
        if (!opt->refine_principal_point) mask |= kParamGroups[model][1];
        if (!opt->refine_extra_params) mask |= kParamGroups[model][2];
      }
      ir->cam_const_mask[ck->second] = mask;
    }
    // ParameterizePoints
    for (int64_t pi = 0; pi < rec->n_points; ++pi) {
      if (!point_in_problem[pi]) continue;
      const int64_t tl = track_len(pi);
      const int64_t mtl = opt->min_track_length > 0 ? std::min<int64_t>(opt->min_track_length, tl) : tl;
      if (mtl > n_registered[pi]) ir->point_const[pidx.at(rec->point3D_id[pi])] = 1;
    }
    for (int64_t pid : const_points) {
      auto it = pidx.find(pid);
      if (it != pidx.end()) ir->point_const[it->second] = 1;
    }
  }
  *out = ir.release();
  return PXR_OK;
}

int pxr_problem_sizes(const pxr_problem_ir* ir, int64_t* n_obs, int64_t* n_images, int64_t* n_cameras, int64_t* n_points) {
  if (!ir) return pxr::fail(PXR_ERR_INVALID_ARGUMENT, "ir is NULL");
  if (n_obs) *n_obs = (int64_t)ir->obs_pt.size();
  if (n_images) *n_images = (int64_t)ir->image_ids.size();
  if (n_cameras) *n_cameras = (int64_t)ir->camera_ids.size();
  if (n_points) *n_points = (int64_t)ir->point_ids.size();
  return PXR_OK;
}

int pxr_problem_copy(const pxr_problem_ir* ir, int64_t* obs_point3D_id, int64_t* obs_image_id, int64_t* obs_point2D_idx,
                     int32_t* obs_img, int64_t* obs_pt, int64_t* image_ids, int64_t* camera_ids, int64_t* point_ids,
                     int32_t* img_cam, uint8_t* pose_const, uint8_t* tvec_const_mask, uint8_t* point_const,
                     uint32_t* cam_const_mask) {
  if (!ir) return pxr::fail(PXR_ERR_INVALID_ARGUMENT, "ir is NULL");
  auto cp = [](auto* dst, const auto& v) { if (dst && !v.empty()) std::copy(v.begin(), v.end(), dst); };
  cp(obs_point3D_id, ir->obs_point3D_id); cp(obs_image_id, ir->obs_image_id); cp(obs_point2D_idx, ir->obs_point2D_idx);
  cp(obs_img, ir->obs_img); cp(obs_pt, ir->obs_pt); cp(image_ids, ir->image_ids); cp(camera_ids, ir->camera_ids);
  cp(point_ids, ir->point_ids); cp(img_cam, ir->img_cam); cp(pose_const, ir->pose_const);
  cp(tvec_const_mask, ir->tvec_const_mask); cp(point_const, ir->point_const); cp(cam_const_mask, ir->cam_const_mask);
  return PXR_OK;
}

void pxr_problem_destroy(pxr_problem_ir* ir) { delete ir; }

}  // extern "C"
