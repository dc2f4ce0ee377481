// pxr_pcg.cuh — ITERATIVE_SCHUR: preconditioned conjugate gradients on the explicit reduced camera system
// with the SCHUR_JACOBI preconditioner (block diagonal of S), the reference's choice above 1000 images
// (pixsfm/bundle_adjustment/src/bundle_optimizer.h:188-190).  Restates ceres' ConjugateGradientsSolver
// (internal/ceres/conjugate_gradients_solver.cc @2.1): Q-based termination zeta = k (Q1 - Q0) / Q1 < eta,
// residual refresh every 10 iterations, x0 = 0.  All scalar logic lives in a device-side state so that the
// host only polls a flag every few iterations.
#pragma once
#include "pxr_device.cuh"

namespace pxr {

struct CGState {
  double rho, last_rho, alpha, beta, Q0, pq, rz, xbr;
  int it, done, failed, max_iter;
  double q_tol;
};

// full symmetric copy of the lower-stored S
static __global__ void cg_mirror_kernel(double* S, int n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * n) return;
  const int r = (int)(i / n), c = (int)(i % n);
  if (c > r) S[i] = S[(int64_t)c * n + r];
}

// inverse of every diagonal block (dim <= 12), one thread per block; Minv row-wise [n][12]
static __global__ void cg_block_inverse_kernel(const double* S, int n, const int32_t* blk_off, const int32_t* blk_dim, int nblk,
                                               double* Minv, int32_t* row_off, int32_t* row_dim, int* fail) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblk) return;
  const int o = blk_off[b], d = blk_dim[b];
  double A[12][24];
  for (int i = 0; i < d; ++i) { for (int j = 0; j < d; ++j) { A[i][j] = S[(int64_t)(o + i) * n + o + j]; A[i][d + j] = i == j ? 1.0 : 0.0; } }
  for (int k = 0; k < d; ++k) {
    const double pv = A[k][k];
    if (!(pv > 0.0)) { *fail = 1; return; }
    for (int j = 0; j < 2 * d; ++j) A[k][j] /= pv;
    for (int i = 0; i < d; ++i) {
      if (i == k) continue;
      const double f = A[i][k];
      for (int j = 0; j < 2 * d; ++j) A[i][j] -= f * A[k][j];
    }
  }
  for (int i = 0; i < d; ++i) {
    row_off[o + i] = o; row_dim[o + i] = d;
    for (int j = 0; j < d; ++j) Minv[(int64_t)(o + i) * 12 + j] = A[i][d + j];
  }
}

// out = S v (S full, row-major), one warp per row
static __global__ void __launch_bounds__(256) cg_gemv_kernel(const double* __restrict__ S, const double* __restrict__ v, double* out, int n,
                                                             const CGState* st) {
  if (st && st->done) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* r = S + (int64_t)row * n;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += r[j] * v[j];
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// single-CTA vector kernels (n is the reduced system size, at most a few 10^4)
__device__ __forceinline__ double cta_sum(double v) {
  __shared__ double sh[32];
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
  return t;
}

// init: x = 0, r = b, state
static __global__ void __launch_bounds__(1024) cg_init_kernel(const double* b, double* x, double* r, int n, CGState* st, int max_iter, double q_tol) {
  double nb = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { x[i] = 0.0; r[i] = b[i]; nb += b[i] * b[i]; }
  nb = cta_sum(nb);
  if (threadIdx.x == 0) {
    st->rho = 1.0; st->last_rho = 1.0; st->alpha = 0; st->beta = 0; st->Q0 = 0.0; st->it = 0; st->failed = 0;
    st->done = nb == 0.0 ? 1 : 0; st->max_iter = max_iter; st->q_tol = q_tol;
  }
}

// z = M^-1 r ; rho = r.z ; p = z (+ beta p)
static __global__ void __launch_bounds__(1024) cg_precond_kernel(const double* Minv, const int32_t* row_off, const int32_t* row_dim, const double* r,
                                                                 double* z, double* p, int n, CGState* st) {
  if (st->done) return;
  double rz = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    const int o = row_off[i], d = row_dim[i];
    for (int j = 0; j < d; ++j) s += Minv[(int64_t)i * 12 + j] * r[o + j];
    z[i] = s; rz += r[i] * s;
  }
  rz = cta_sum(rz);
  __shared__ double beta_s; __shared__ int stop;
  if (threadIdx.x == 0) {
    st->it += 1;
    st->last_rho = st->rho; st->rho = rz;
    stop = 0;
    if (rz == 0.0 || !isfinite(rz)) { st->failed = 1; st->done = 1; stop = 1; }
    double beta = 0.0;
    if (st->it > 1) { beta = rz / st->last_rho; if (beta == 0.0 || !isfinite(beta)) { st->failed = 1; st->done = 1; stop = 1; } }
    st->beta = beta; beta_s = beta;
  }
  __syncthreads();
  if (stop) return;
  const double beta = beta_s;
  const bool first = st->it == 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = first ? z[i] : z[i] + beta * p[i];
}

// alpha = rho / p.q ; x += alpha p ; r -= alpha q (or flagged for refresh) ; Q test
static __global__ void __launch_bounds__(1024) cg_update_kernel(const double* b, const double* p, const double* q, double* x, double* r, int n, CGState* st) {
  if (st->done) return;
  double pq = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) pq += p[i] * q[i];
  pq = cta_sum(pq);
  __shared__ double alpha_s; __shared__ int stop;
  if (threadIdx.x == 0) {
    stop = 0;
    if (pq <= 0.0 || !isfinite(pq)) { st->done = 1; stop = 1; }          // matrix not positive definite along p: keep x
    const double alpha = st->rho / pq;
    if (!stop && !isfinite(alpha)) { st->failed = 1; st->done = 1; stop = 1; }
    st->alpha = alpha; alpha_s = alpha;
  }
  __syncthreads();
  if (stop) return;
  const double alpha = alpha_s;
  const bool refresh = (st->it % 10) == 0;   // residual_reset_period: r = b - S x is recomputed by the caller's next launches
  for (int i = threadIdx.x; i < n; i += blockDim.x) { x[i] += alpha * p[i]; if (!refresh) r[i] -= alpha * q[i]; }
}

// after an optional residual refresh: Q1 = -0.5 x.(b + r), zeta test, iteration cap
static __global__ void __launch_bounds__(1024) cg_check_kernel(const double* b, const double* x, const double* r, int n, CGState* st) {
  if (st->done) return;
  double xbr = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xbr += x[i] * (b[i] + r[i]);
  xbr = cta_sum(xbr);
  if (threadIdx.x == 0) {
    const double Q1 = -0.5 * xbr;
    const double zeta = st->it * (Q1 - st->Q0) / Q1;
    if (zeta < st->q_tol) st->done = 1;
    st->Q0 = Q1;
    if (st->it >= st->max_iter) st->done = 1;
  }
}

// r = b - tmp (only on refresh iterations)
static __global__ void __launch_bounds__(1024) cg_refresh_kernel(const double* b, const double* Sx, double* r, int n, const CGState* st) {
  if (st->done || (st->it % 10) != 0) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) r[i] = b[i] - Sx[i];
}

}  // namespace pxr
