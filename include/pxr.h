/*
 * pxr.h — C-ABI of the B200-native featuremetric refinement engine (libpxr.so).
 *
 * Drop-in boundary for the featuremetric keypoint-adjustment (KA) and
 * bundle-adjustment (BA) hot path of cvg/pixel-perfect-sfm.  The reference has
 * no C-ABI: its boundary is the pybind11 module `pixsfm._pixsfm`
 * (reference pixsfm/_pixsfm/bindings.cc:34-63).  Every entry point below names
 * the reference interface it replaces; the pybind/ctypes stub a maintainer
 * would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, `extern "C"`, no C++/torch types; all sizes explicit.
 *  - every function returns a pxr_status; pxr_last_error() gives the message
 *    for the calling thread.  No exception crosses the boundary, nothing
 *    aborts the process (the reference's glog CHECKs abort:
 *    bundle_optimizer.h:222).
 *  - the caller owns all host memory; the library owns all device memory.
 *  - there is NO CPU fallback: every compute entry point fails with
 *    PXR_ERR_NO_DEVICE when no CUDA device is usable.
 *  - one handle is used by one host thread at a time (thread-compatible).
 *
 * Problem IR (SoA, independent of COLMAP objects).  The host adapter resolves
 * the reference's BundleAdjustmentSetup + BundleOptimizerOptions into plain
 * constancy masks exactly as BundleOptimizer::Parameterize{Points,Images,
 * Cameras} does (reference bundle_adjustment/src/bundle_optimizer.h:335-442).
 */
#ifndef PXR_H_
#define PXR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXR_VERSION_MAJOR 0
#define PXR_VERSION_MINOR 1
#define PXR_MAX_CAM_PARAMS 12

typedef enum {
  PXR_OK = 0,
  PXR_ERR_INVALID_ARGUMENT = 1, /* -> Python ValueError / std::invalid_argument */
  PXR_ERR_UNSUPPORTED = 2,      /* unsupported (channels, nodes, model) combination */
  PXR_ERR_NO_DEVICE = 3,        /* no CUDA device: the product never falls back to CPU */
  PXR_ERR_CUDA = 4,             /* CUDA runtime error (message has the cudaError string) */
  PXR_ERR_NCCL = 5,
  PXR_ERR_NUMERIC = 6,          /* solver failure (non-PD reduced system that damping could not fix) */
  PXR_ERR_INTERRUPTED = 7,      /* interrupt callback asked to stop (PyInterrupt, util/src/py_interrupt.h) */
  PXR_ERR_INTERNAL = 8
} pxr_status;

/* COLMAP 3.8 camera model ids (colmap/base/camera_models.h; used through
 * CAMERA_MODEL_SWITCH_CASES in reference residuals/src/feature_reference.h:224-252). */
typedef enum {
  PXR_CAM_SIMPLE_PINHOLE = 0, /* f, cx, cy */
  PXR_CAM_PINHOLE = 1,        /* fx, fy, cx, cy */
  PXR_CAM_SIMPLE_RADIAL = 2,  /* f, cx, cy, k */
  PXR_CAM_RADIAL = 3,         /* f, cx, cy, k1, k2 */
  PXR_CAM_OPENCV = 4,         /* fx, fy, cx, cy, k1, k2, p1, p2 */
  PXR_CAM_OPENCV_FISHEYE = 5, /* fx, fy, cx, cy, k1, k2, k3, k4 */
  PXR_CAM_FULL_OPENCV = 6     /* fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6 */
} pxr_camera_model;

typedef enum { PXR_F16 = 0, PXR_F32 = 1, PXR_F64 = 2 } pxr_dtype;

/* ceres loss functions reachable from the reference's {"name","params"} dicts
 * (bundle_adjustment/main.py:37-40). */
typedef enum {
  PXR_LOSS_TRIVIAL = 0,
  PXR_LOSS_CAUCHY = 1, /* default, scale 0.25 (bundle_adjustment_options.h:49) */
  PXR_LOSS_HUBER = 2,
  PXR_LOSS_SOFT_L1 = 3,
  PXR_LOSS_ARCTAN = 4
} pxr_loss_type;

typedef enum {
  PXR_SOLVER_AUTO = 0,        /* by #images as bundle_optimizer.h:181-191 */
  PXR_SOLVER_DENSE_SCHUR = 1, /* explicit reduced camera system + dense Cholesky */
  PXR_SOLVER_SPARSE_SCHUR = 2,/* exact too: runs the dense-Cholesky path */
  PXR_SOLVER_ITERATIVE_SCHUR = 3 /* block-Jacobi PCG on the reduced system */
} pxr_linear_solver;

/* InterpolationConfig subset on the named path (base/src/interpolation.h:39-51):
 * mode BICUBIC, one node {0,0}. */
typedef struct {
  int32_t l2_normalize;   /* default 1 */
  int32_t use_float_simd; /* 0: fp32 horizontal + fp64 vertical (reference default), 1: all fp32 */
  int32_t check_bounds;   /* must be 0 on this path (reference default) */
  int32_t reserved;
} pxr_interp_config;

/* Solver options = ceres::Solver::Options fields the reference sets
 * (bundle_adjustment_options.h:48-64, base/main.py:9-22). */
typedef struct {
  int32_t loss_type;                /* pxr_loss_type */
  double loss_scale;                /* "params"[0] */
  int32_t linear_solver;            /* pxr_linear_solver */
  int32_t max_num_iterations;       /* 100 */
  int32_t max_linear_solver_iterations; /* 200 */
  int32_t max_num_consecutive_invalid_steps; /* 10 */
  double function_tolerance;        /* 0 */
  double gradient_tolerance;        /* 0 */
  double parameter_tolerance;       /* 0 (KA: 1e-5) */
  int32_t use_inner_iterations;     /* BA python default: 1 (bundle_adjustment/main.py:41-44) */
  double inner_iteration_tolerance; /* 1e-3 (ceres default) */
  double initial_trust_region_radius; /* 1e4 */
  double max_trust_region_radius;   /* 1e16 */
  double min_trust_region_radius;   /* 1e-32 */
  double min_relative_decrease;     /* 1e-3 */
  double min_lm_diagonal;           /* 1e-6 */
  double max_lm_diagonal;           /* 1e32 */
  int32_t jacobi_scaling;           /* 1 */
  int32_t deterministic;            /* 1: bit-reproducible solve.  The normal equations are assembled by fixed-order reductions
                                     * instead of fp64 atomics (image / pair chunk partials summed in chunk order, point blocks
                                     * observation by observation, scalar sums block by block: csrc/pxr_block.cuh) on the
                                     * block-mode driver; needs <= 8 camera columns per image.  0: atomics (run-to-run
                                     * differences ~1e-16 relative per sum).  tests/test_gpu_deterministic.py */
  int32_t use_nonmonotonic_steps;   /* ceres::Solver::Options::use_nonmonotonic_steps (the reference's configs/default.yaml
                                     * sets it for KA, BA and QKA); 0 */
  int32_t max_consecutive_nonmonotonic_steps; /* 5 */
} pxr_solver_options;

/* One featuremetric BA problem (reference: what BundleOptimizer::SetUp turns a
 * colmap::Reconstruction + FeatureView + references into,
 * bundle_optimizer.h:139-165, feature_reference_bundle_optimizer.h:90-149).
 * Parameter arrays are IN/OUT: pxr_ba_solve updates them in place, as the
 * reference updates the Reconstruction through raw double*
 * (feature_reference_bundle_optimizer.h:111-114). */
typedef struct {
  /* cameras (intrinsics) */
  int32_t n_cameras;
  const int32_t* cam_model;       /* [n_cameras] pxr_camera_model */
  double* cam_params;             /* [n_cameras][PXR_MAX_CAM_PARAMS], unused tail ignored */
  const uint32_t* cam_const_mask; /* [n_cameras] bit i: param i constant; 0xFFFFFFFF: block constant */
  /* images (poses); qvec is (w,x,y,z), world->camera, as COLMAP */
  int32_t n_images;
  double* qvec;                   /* [n_images][4] */
  double* tvec;                   /* [n_images][3] */
  const int32_t* img_cam;         /* [n_images] index into cameras */
  const uint8_t* pose_const;      /* [n_images] 1: qvec and tvec constant */
  const uint8_t* tvec_const_mask; /* [n_images] bit i: tvec[i] constant (SubsetManifold) */
  /* points */
  int64_t n_points;
  double* xyz;                    /* [n_points][3] */
  const uint8_t* point_const;     /* [n_points] 1: constant */
  /* observations = residual blocks, sorted by point index (ties: any fixed order) */
  int64_t n_obs;
  const int32_t* obs_img;         /* [n_obs] */
  const int64_t* obs_pt;          /* [n_obs] non-decreasing */
  const int64_t* obs_patch;       /* [n_obs] index into the patch slab, or NULL: patch i = obs i */
  /* feature patches: slab [n_patches][ph][pw][channels], HWC interleaved like
   * FeaturePatch (features/src/featurepatch.h:244-262, grid2d.h:29-73) */
  int64_t n_patches;
  const void* patches;            /* host pointer (or device pointer if patches_on_device) */
  int32_t patches_on_device;
  int32_t patch_dtype;            /* pxr_dtype */
  int32_t ph, pw, channels;
  const int32_t* corner;          /* [n_patches][2] (x0,y0) */
  const double* scale;            /* [n_patches][2] (sx,sy) */
  double upsampling_factor;       /* FeaturePatch::upsampling_factor_, 1.0 */
  /* per-point reference descriptors, fp64 (Reference::descriptor, references.h:32-65);
   * NULL => residual = interpolated feature (costmap use, feature_reference.h:128-130) */
  const double* refs;             /* [n_points][channels] */
  /* Optional: the patch slab given as several host blocks (one per FeatureMap numpy array, no host-side
   * concatenation: featuremap.cc:38-44 keeps references to the arrays).  When n_patch_blocks > 0,
   * `patches` is ignored and patch index i lives in the block b with offsets[b] <= i < offsets[b+1]. */
  int32_t n_patch_blocks;
  const void* const* patch_block_ptrs;   /* [n_patch_blocks]; each block may be host OR device memory */
  const int64_t* patch_block_counts;     /* [n_patch_blocks] patches per block */
} pxr_ba_desc;

typedef struct {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_successful;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  int32_t linear_solver_iterations;
  double iteration_time_s;
} pxr_iteration_summary;

/* Mirrors the fields of ceres::Solver::Summary that the reference reads
 * (bundle_optimizer.h:236-241, util/src/statistics.h:54-129). */
typedef struct {
  double initial_cost;
  double final_cost;
  int32_t num_residual_blocks;
  int64_t num_residuals;          /* num_residual_blocks * channels */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t num_inner_iteration_steps;
  int32_t termination_type;       /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE, 3 USER_FAILURE */
  double total_time_s;            /* wall, includes uploads */
  double solve_time_s;            /* wall, LM loop only (inputs resident) */
  double h2d_bytes, d2h_bytes;
  int32_t num_iterations;         /* entries valid in `iterations` */
  int32_t iterations_capacity;    /* in: capacity of `iterations` (may be 0) */
  pxr_iteration_summary* iterations; /* caller-allocated */
  int64_t kernel_launches;        /* CUDA kernels of this library launched by the call */
  char message[256];
  /* window residency of the patch slab (pxr_ba_run on a pinned host buffer, csrc/pxr_resident.cuh): side of the window
   * brought over per observation (0: the whole slab was uploaded), observations whose whole patch was fetched later
   * because the point left its window, and evaluation passes repeated for that; h2d_bytes counts what really crossed */
  int32_t resident_window;
  int32_t resident_passes_repeated;
  int64_t resident_refetched;
} pxr_summary;

typedef struct pxr_ctx pxr_ctx;
typedef struct pxr_ba pxr_ba;

/* ---- context ----------------------------------------------------------- */
const char* pxr_last_error(void);
int pxr_version(void);
/* device < 0 → current device. Fails with PXR_ERR_NO_DEVICE when no GPU. */
int pxr_ctx_create(int device, pxr_ctx** out);
int pxr_ctx_destroy(pxr_ctx* ctx);
/* Multi-GPU: point-sharded BA, one process per GPU.  nccl_unique_id is the 128-byte
 * ncclUniqueId made on rank 0 (pxr_nccl_unique_id) and distributed by the host's own
 * plumbing (torch.distributed broadcast in bench.py).  No reference counterpart
 * (the reference is single-process: base/src/parallel_optimizer.h:77-211). */
int pxr_nccl_unique_id(void* id128);
int pxr_ctx_init_comm(pxr_ctx* ctx, int rank, int world, const void* id128);
int pxr_ctx_sync(pxr_ctx* ctx);
/* NCCL collectives issued through this context so far (bench.py divides by the LM iterations it timed: the contract is
 * ONE all-reduce of the reduced camera blocks per LM iteration), and whether the per-iteration SCALAR exchange runs over
 * peer memory (cudaIpc-mapped mailboxes over NVLink, 1) or had to fall back to NCCL (0). */
int64_t pxr_ctx_nccl_collectives(pxr_ctx* ctx);
int pxr_ctx_mailbox_ready(pxr_ctx* ctx);
/* CUDA-event stopwatch on the library stream (bench.py times K LM iterations with it) */
int pxr_ctx_timer_start(pxr_ctx* ctx);
int pxr_ctx_timer_stop(pxr_ctx* ctx, double* elapsed_ms);
int64_t pxr_ctx_kernel_launches(pxr_ctx* ctx);

/* Interruptibility.  The reference polls PyErr_CheckSignals from a ceres::IterationCallback and from its thread-pool
 * wait loop (util/src/py_interrupt.h:29-38, base/src/callbacks.h:10-20, parallel_optimizer.h:186).  Here the host
 * registers ONE process-wide callback; the LM drivers call it between iterations (at most every 200 ms; the reference polls every 2 s) and stop with
 * PXR_ERR_INTERRUPTED when it returns non-zero.  fn == NULL removes it.  pxr_poll_interrupt() calls it right away
 * (no device needed) and returns its answer, 0 without a callback. */
typedef int (*pxr_interrupt_fn)(void* user);
int pxr_set_interrupt_callback(pxr_interrupt_fn fn, void* user);
int pxr_poll_interrupt(void);

/* ---- default option blocks --------------------------------------------- */
void pxr_default_interp_config(pxr_interp_config* c);  /* base/main.py:1-7 */
void pxr_default_ba_options(pxr_solver_options* o);    /* bundle_adjustment/main.py:30-62 */
void pxr_default_ka_options(pxr_solver_options* o);    /* keypoint_adjustment/main.py:60-83 */

/* ---- featuremetric BA ---------------------------------------------------
 * replaces _bundle_adjustment.FeatureReferenceBundleOptimizer.run / set_up /
 * solve_problem / summary (bundle_adjustment/bindings.cc:36-51,137-141). */
int pxr_ba_create(pxr_ctx* ctx, const pxr_ba_desc* desc, const pxr_interp_config* interp,
                  const pxr_solver_options* opt, pxr_ba** out);   /* uploads (set_up) */
int pxr_ba_solve(pxr_ba* ba, pxr_summary* summary);               /* solve_problem */
/* Continues (or starts) the LM trajectory of `ba` for n more iterations (bench.py: W warm-up, then K timed). */
int pxr_ba_iterate(pxr_ba* ba, int n_iterations, pxr_summary* summary);
/* CUDA-event timing of the residual/Jacobian kernel launches inside solve/iterate.
 * enable: 1/0 switches recording and clears the log, -1 leaves it; which: 1 Jacobian-mode K1, 0 cost-only K1.
 * total_ms/count (optional) receive the summed device time and number of recorded launches. */
int pxr_ba_kernel_timing(pxr_ba* ba, int enable, int which, double* total_ms, int* count);
int pxr_ba_read_params(pxr_ba* ba, double* cam_params, double* qvec, double* tvec, double* xyz);
/* Puts the resident problem back to the given parameters (same array shapes as in the desc) and forgets the LM
 * trajectory; patches, topology, work space and captured graphs stay.  bench.py: W warm-up iterations, reset, then K
 * timed iterations FROM ITERATION ZERO, like the CPU arm.  (The reference's optimizers are one-shot objects,
 * bundle_optimizer.h:121-122; this is what constructing a second one on the same FeatureView amounts to.) */
int pxr_ba_reset(pxr_ba* ba, const double* cam_params, const double* qvec, const double* tvec, const double* xyz);
int pxr_ba_destroy(pxr_ba* ba);
/* one-shot convenience = FeatureReferenceBundleOptimizer.run: upload, solve, write back into desc arrays */
int pxr_ba_run(pxr_ctx* ctx, const pxr_ba_desc* desc, const pxr_interp_config* interp,
               const pxr_solver_options* opt, pxr_summary* summary);

/* Device memory pxr_ba_create / pxr_ba_run will take for this problem, in bytes (host computation, no device needed) —
 * the counterpart of the RAM estimate the reference logs before it solves (bundle_optimizer.h:200-208,
 * util/src/memory.h:45-69).  patch_bytes: the patch slab (0 when it is already resident); state_bytes: per-observation
 * buffers, Schur pair lists and point blocks; reduced_bytes: the reduced camera system (dense, or its block-sparse form
 * when the solver is ITERATIVE_SCHUR and dense would exceed 4 GB).  Any output may be NULL. */
int pxr_ba_estimate_device_bytes(const pxr_ba_desc* desc, const pxr_solver_options* opt, double* patch_bytes,
                                 double* state_bytes, double* reduced_bytes);

/* Parity / introspection: evaluates every residual block at the current parameters.
 * Replaces calling FeatureReferenceCostFunctor::operator() per block
 * (residuals/src/feature_reference.h:98-137).  Any output may be NULL.
 *  sq_norm [n_obs]      ||r||^2 (uncorrected)
 *  gtr     [n_obs][2]   G^T r,  G = [d r/d u, d r/d v] in patch pixel units (u=col, v=row)
 *  gtg     [n_obs][3]   G^T G (uu, uv, vv)
 *  xy      [n_obs][2]   projected image point
 *  residuals [n_obs][C] r as double (large! optional)
 *  cost    scalar       sum 0.5*rho(||r||^2) */
int pxr_ba_evaluate(pxr_ba* ba, double* sq_norm, double* gtr, double* gtg, double* xy,
                    double* residuals, double* cost);
/* The cost-functor surface: what a ceres::CostFunction built by the reference's `_residuals` factories returns from
 * Evaluate() (residuals/bindings.cc:14-30; CreateFeatureReferenceCostFunctor / ...ConstantPoseCostFunctor,
 * residuals/src/feature_reference.h:256-321; parameter order of operator(), :98-137), for every residual block of the
 * problem at once and in factored form: the Jacobian of block o is  J_o = G_o^T-less product  G_o (C x 2) * P_o (2 x p).
 *  residuals [n_obs][C]       r (f - reference, or f when the problem has no references)
 *  grad      [n_obs][2][C]    d r/d u, d r/d v in patch pixel units (after the L2-normalisation chain rule)
 *  juv       [n_obs][2][W]    d(u,v)/d(3 rotation tangent (QuaternionManifold, left-multiplicative) | 3 t | 3 X | K
 *                             camera parameters), W = 9 + K returned through juv_cols, K = largest kNumParams in use
 *  xy        [n_obs][2]       projected image point.   Any output may be NULL. */
int pxr_ba_evaluate_jacobians(pxr_ba* ba, double* residuals, double* grad, double* juv, int32_t* juv_cols, double* xy);
/* Introspection for parity tests: linearise at the current parameters (camera blocks Hcc [nc*nc,
 * lower triangle], gc, point blocks Hpp [n_points*9], gp), assemble the damped Schur system S/rhs at
 * `radius` (Jacobi scaling taken from this linearisation, as in LM iteration 0), solve it and
 * return the full step `delta` [n_local] and the model cost change.  Any output may be NULL. */
int pxr_ba_debug_linearize(pxr_ba* ba, double radius, double* cost, double* Hcc, double* gc, double* Hpp,
                           double* gp, double* S, double* rhs, double* delta, double* model_cost_change);
/* Runs the per-point inner iterations (ceres CoordinateDescentMinimizer equivalent) once on the
 * current parameters (parity tests). */
int pxr_ba_debug_inner_iterations(pxr_ba* ba);
/* Timing hooks for bench.py: run `iters` LM iterations' worth of the named stage on the
 * resident problem and return the average device time per launch in ms (CUDA events on the
 * library stream).  stage: 0 residual/Jacobian kernel (K1), 1 cost-only kernel, 2 projection+K1+block build,
 * 3 projection kernel (K0), 4 block build, 5 inner-iteration kernel */
int pxr_ba_time_stage(pxr_ba* ba, int stage, int iters, double* ms_per_launch);
/* Device-side synthetic scene generator used by bench.py (data: synthetic). Fills a patch slab
 * on the device from per-patch smooth fields; see DESIGN.md §bench. */
int pxr_synth_patches_device(pxr_ctx* ctx, void** d_patches_out, int64_t n_patches, int ps, int channels,
                             const double* uv0 /*[n][2] true location in patch px*/,
                             const int64_t* field_id /*[n] which smooth field (point id)*/,
                             uint64_t seed, double noise_sigma);
int pxr_device_free(pxr_ctx* ctx, void* dptr);
/* pinned host staging buffers for callers that want full PCIe bandwidth on the patch upload */
int pxr_host_alloc_pinned(void** out, size_t bytes);
int pxr_host_free_pinned(void* p);
int pxr_memcpy_d2h(pxr_ctx* ctx, void* host, const void* dev, size_t bytes);

/* ---- reference extraction ----------------------------------------------
 * replaces _bundle_adjustment.ReferenceExtractor.run (bindings.cc:28-34,172-177;
 * reference_extractor.h:171-318, irls_optim.h:23-71).
 *  refs_out    [n_points][C] fp64  reference descriptor (closest_to_robust_mean)
 *  src_obs_out [n_points]    index (into obs) of the observation picked as reference, -1 if none */
int pxr_refs_compute(pxr_ctx* ctx, const pxr_ba_desc* desc, const pxr_interp_config* interp,
                     int loss_type, double loss_scale, int iters,
                     double* refs_out, int64_t* src_obs_out, pxr_summary* summary);

/* descriptors of every observation at its current projection = ReferenceExtractor::FillDescriptorTrack
 * (reference_extractor.h:207-237), what Reference.observations holds with keep_observations=True (needed by
 * FindNearestReferences).  out_desc [n_obs][channels] fp64, in observation order. */
int pxr_obs_descriptors(pxr_ctx* ctx, const pxr_ba_desc* desc, const pxr_interp_config* interp, double* out_desc);

/* ---- cost maps ---------------------------------------
 * replaces _bundle_adjustment.CostMapExtractor.run (bundle_adjustment/bindings.cc:20-26,179-184;
 * costmap_extractor.h:93-228 Run/RunSubset, :230-358 FillPointCostmap) followed, on the caller's side,
 * by CostMapBundleOptimizer (costmap_bundle_optimizer.h:76-132) = pxr_ba_* on the returned
 * 3-channel patches with refs == NULL and l2_normalize == 0 (bundle_adjustment/main.py:262-281).
 * Config mirrors CostMapConfig (costmap_extractor.h:18-40) + the ReferenceConfig used by the
 * extractor's embedded ReferenceExtractor (reference_extractor.h:30-50). */
typedef struct pxr_costmap_config {
  int32_t loss_type;                /* pxr_loss_type, default TRIVIAL */
  double loss_scale;
  int32_t as_gradientfield;         /* 1: (cost, dcost/dr, dcost/dc), 0: cost only */
  int32_t compute_cross_derivative; /* must be 0 (4-channel variant not built) */
  int32_t apply_sqrt;
  double upsampling_factor;         /* must be 1.0 */
  int32_t compute_refs;             /* 1: run the reference extraction first, on the same upload */
  int32_t ref_loss_type;            /* CAUCHY 0.25, 100 IRLS iterations (bundle_adjustment/main.py:46-57) */
  double ref_loss_scale;
  int32_t ref_iters;
} pxr_costmap_config;
int pxr_default_costmap_config(pxr_costmap_config* cfg);
/*  desc        the feature problem (C-channel patches, geometry); desc->refs used when !compute_refs
 *  refs_io     [n_points][C] fp64: written when compute_refs, else read if desc->refs is NULL; may be NULL
 *  src_obs_out [n_points] or NULL (see pxr_refs_compute)
 *  out_host    [n_patches][ph][pw][OC] in desc->patch_dtype (the reference instantiates Run<dtype,dtype>), or NULL
 *  out_device  if not NULL receives a device pointer to the same array (release with pxr_device_free):
 *              feed it to pxr_ba_create with patches_on_device = 1 to skip the host round trip */
int pxr_costmaps_compute(pxr_ctx* ctx, const pxr_ba_desc* desc, const pxr_interp_config* interp,
                         const pxr_costmap_config* cfg, double* refs_io, int64_t* src_obs_out,
                         void* out_host, void** out_device, pxr_summary* summary);

/* ---- problem construction -------------------------------------------------------
 * replaces BundleOptimizer::SetUp / AddImageToProblem / AddPointToProblem / Parameterize{Points,Images,Cameras}
 * (bundle_adjustment/src/bundle_optimizer.h:139-165,247-331,335-442) and ReferenceExtractor::GetVisibleObservations
 * (bundle_adjustment/src/reference_extractor.h:171-205).  The reference walks a colmap::Reconstruction; the caller hands
 * the same data as a structure-of-arrays view (host memory, caller-owned).  Host code only, no device needed. */
typedef struct pxr_recon_view {
  int64_t n_images;
  const int64_t* image_id;          /* [n_images] */
  const int64_t* image_camera_id;   /* [n_images] */
  const int64_t* p2d_begin;         /* [n_images+1] offsets of the image's 2D points in p2d_point3D_id */
  const int64_t* p2d_point3D_id;    /* [p2d_begin[n_images]] 3D point of every 2D point, -1 = none */
  int64_t n_cameras;
  const int64_t* camera_id;         /* [n_cameras] */
  const int32_t* camera_model;      /* [n_cameras] COLMAP model id 0..6 */
  int64_t n_points;
  const int64_t* point3D_id;        /* [n_points] */
  const int64_t* track_begin;       /* [n_points+1] */
  const int64_t* track_image_id;    /* [track_begin[n_points]] */
  const int64_t* track_point2D_idx; /* [track_begin[n_points]] */
} pxr_recon_view;
/* colmap::BundleAdjustmentConfig as pixsfm::BundleAdjustmentSetup exposes it (bundle_adjustment/bindings.cc:81-111) */
typedef struct pxr_ba_setup_view {
  int64_t n_images; const int64_t* image_ids;
  int64_t n_const_poses; const int64_t* const_pose_ids;
  int64_t n_const_tvecs; const int64_t* const_tvec_ids; const uint8_t* const_tvec_masks;   /* bit k: tvec[k] constant */
  int64_t n_const_cameras; const int64_t* const_camera_ids;
  int64_t n_var_points; const int64_t* var_point_ids;
  int64_t n_const_points; const int64_t* const_point_ids;
} pxr_ba_setup_view;
typedef struct pxr_ba_build_options {
  int32_t refine_focal_length, refine_principal_point, refine_extra_params, refine_extrinsics;   /* BundleOptimizerOptions */
  int32_t min_track_length;          /* -1: off */
  int32_t mode;                      /* 0: bundle adjustment (setup required), 1: reference extraction */
  int64_t n_ref_points;              /* mode 1: the points to extract references for ...                        */
  const int64_t* ref_point_ids;
  const uint8_t* track_has_patch;    /* ... and, per track element, whether a feature patch exists (NULL: all do) */
} pxr_ba_build_options;
typedef struct pxr_problem_ir pxr_problem_ir;
int pxr_problem_build(const pxr_recon_view* rec, const pxr_ba_setup_view* setup, const pxr_ba_build_options* opt,
                      pxr_problem_ir** out);
int pxr_problem_sizes(const pxr_problem_ir* ir, int64_t* n_obs, int64_t* n_images, int64_t* n_cameras, int64_t* n_points);
/* observations sorted by point (stable), ids of the blocks in index order, index maps and constancy masks as
 * pxr_ba_desc takes them; any output may be NULL */
int pxr_problem_copy(const pxr_problem_ir* ir, int64_t* obs_point3D_id, int64_t* obs_image_id, int64_t* obs_point2D_idx,
                     int32_t* obs_img, int64_t* obs_pt, int64_t* image_ids, int64_t* camera_ids, int64_t* point_ids,
                     int32_t* img_cam, uint8_t* pose_const, uint8_t* tvec_const_mask, uint8_t* point_const,
                     uint32_t* cam_const_mask);
void pxr_problem_destroy(pxr_problem_ir* ir);

/* ---- descriptor interpolation ------------------------------------------------
 * replaces _features.PatchInterpolator(cfg).interpolate / interpolate_nodes for one node
 * (features/bindings.cc, features/src/patch_interpolator.h:125-135, dynamic_patch_interpolator.h): the bicubic
 * (+ L2-normalised) descriptor of patch item_patch[i] at image coordinates xy[i], the quantity
 * FindNearestReferences (localization/src/nearest_references.h:20-53) and find_feature_inliers compare.
 *  patches [n_patches][ph][pw][channels] host memory; corner/scale per patch; out_desc [n_items][channels] fp64 */
int pxr_interpolate_descriptors(pxr_ctx* ctx, const void* patches, int64_t n_patches, int32_t patch_dtype,
                                int32_t ph, int32_t pw, int32_t channels, const int32_t* corner, const double* scale,
                                double upsampling_factor, int64_t n_items, const int64_t* item_patch, const double* xy,
                                const pxr_interp_config* interp, double* out_desc);

/* ---- dense map -> patch slab -----------------------------------------------
 * replaces the sparse branch of FeatureExtractor.tensor_to_fmap + extract_patches_numpy
 * (features/extractor.py:176-201, features/extract_patches.py:36-44: the GPU -> CPU copy the authors flag as their
 * bottleneck): L2-normalise every pixel's C-vector (fp32, x / max(||x||, 1e-12)), cast, and gather one ps x ps window
 * per keypoint into the slab layout [n][ps][ps][C] the optimizers take.
 *  dense     one image's map, [C][H][W] (channels_first = 1, the CNN's NCHW) or [H][W][C]; host or device memory
 *  corners   host [n][2] (x0, y0), every window must lie inside the map (features/extractor.py:192-193 clamps them)
 *  out_host  optional host copy; out_device optional: receives a device pointer (release with pxr_device_free) that
 *            pxr_ba_create / pxr_ka_run accept as device patches or as a device patch block */
int pxr_extract_patches(pxr_ctx* ctx, const void* dense, int32_t dense_on_device, int32_t in_dtype, int32_t channels,
                        int32_t height, int32_t width, int32_t channels_first, const int32_t* corners, int64_t n_patches,
                        int32_t patch_size, int32_t l2_normalize, int32_t out_dtype, void* out_host, void** out_device);

/* ---- featuremetric KA ---------------------------------------------------
 * replaces _keypoint_adjustment.FeatureMetricKeypointOptimizer.run
 * (keypoint_adjustment/bindings.cc:17-34; featuremetric_keypoint_optimizer.h:50-202;
 * keypoint_optimizer.h:110-157). Keypoints are IN/OUT. */
typedef struct {
  int64_t n_keypoints;
  double* keypoints;            /* [n_keypoints][2] COLMAP image coords, in/out */
  const uint8_t* kp_const;      /* [n_keypoints] 1: constant (root nodes) */
  const int64_t* kp_patch;      /* [n_keypoints] patch index, or NULL: identity */
  int64_t n_edges;
  const int64_t* edge_src;      /* [n_edges] keypoint index */
  const int64_t* edge_dst;      /* [n_edges] */
  const double* edge_weight;    /* [n_edges] ScaledLoss weight (similarity) */
  const int32_t* edge_problem;  /* [n_edges] non-decreasing problem label (RunParallel groups) */
  int32_t n_problems;
  int64_t n_patches;
  const void* patches;
  int32_t patches_on_device;
  int32_t patch_dtype;
  int32_t ph, pw, channels;
  const int32_t* corner;        /* [n_patches][2] */
  const double* scale;          /* [n_patches][2] */
  double upsampling_factor;
  double bound;                 /* KeypointOptimizerOptions::bound, 4.0 from python */
  int32_t patches_are_sparse;   /* FeatureMap::IsSparse() */
  /* optional: patches given as blocks (one per FeatureMap numpy array, featuremap.cc:38-44) that are laid out
   * back to back on the device, as in pxr_ba_desc; each block may be host OR device memory */
  int32_t n_patch_blocks;
  const void* const* patch_block_ptrs;
  const int64_t* patch_block_counts;
  /* optional "query" mode (localization/src/single_query_keypoint_optimizer.h:84-203, residual
   * FeatureReference2DCostFunctor, residuals/src/feature_reference.h:14-62): when ref_desc != NULL EVERY edge is
   * (keypoint edge_src[e]) against the FIXED descriptor ref_desc[edge_dst[e]]; edge_dst then indexes ref_desc, not
   * keypoints.  The normal equations are block diagonal (2x2 per keypoint), so a problem may hold any number of
   * keypoints; all keypoints of one problem still share one trust region, as in the reference's single ceres::Problem. */
  const double* ref_desc;       /* [n_ref_desc][channels] fp64 or NULL */
  int64_t n_ref_desc;
} pxr_ka_desc;

int pxr_ka_run(pxr_ctx* ctx, const pxr_ka_desc* desc, const pxr_interp_config* interp,
               const pxr_solver_options* opt, pxr_summary* summary);

/* ---- host-side integer algorithms (bit-exact targets) -------------------
 * replace _base.compute_track_labels / compute_score_labels / compute_root_labels
 * (base/src/graph.cc:126-256) and keypoint_adjustment/main.py:13-57 find_problem_labels. */
int pxr_graph_track_labels(int64_t n_nodes, const int32_t* node_image, int64_t n_edges,
                           const int64_t* e_src, const int64_t* e_dst, const double* e_sim,
                           int64_t* track_labels_out);
int pxr_graph_score_labels(int64_t n_nodes, int64_t n_edges, const int64_t* e_src,
                           const int64_t* e_dst, const double* e_sim,
                           const int64_t* track_labels, double* scores_out);
int pxr_graph_root_labels(int64_t n_nodes, const int64_t* track_labels, const double* scores,
                          uint8_t* is_root_out);
int pxr_ka_problem_labels(int64_t n_nodes, const int64_t* track_labels, int32_t max_per_problem,
                          int32_t* problem_labels_out, int32_t* n_problems_out);
/* BA: reference bundle_adjustment/main.py:21-27 (label = p3D_id // max_tracks_per_problem) lives in Python. */

/* Multi-GPU sharding plan: contiguous point ranges balanced by observation count. */
int pxr_shard_points(int64_t n_points, int64_t n_obs, const int64_t* obs_pt, int world,
                     int64_t* point_begin_out /*[world+1]*/, int64_t* obs_begin_out /*[world+1]*/);
/* KA: the packed problems are independent (ParallelOptimizer::RunParallel, base/src/parallel_optimizer.h:77-211,
 * runs one task per label), so ranks take whole problems and need no collective.  `weight` = cost of a problem
 * (patch bytes or edge count); heaviest-first onto the least loaded rank, deterministic. */
int pxr_shard_ka_problems(int32_t n_problems, const int64_t* weight, int world, int32_t* rank_of_problem_out);

#ifdef __cplusplus
}
#endif
#endif /* PXR_H_ */
