// Dense Cholesky of the reduced camera system as ONE persistent kernel (tile DAG with flags).
//
// Replaces, for the exact DENSE_SCHUR / SPARSE_SCHUR step of ceres::Solve (reference call site
// bundle_adjustment/src/bundle_optimizer.h:181-191,224), the chain of ~3*n/32 dependent launches
// (pxr_ba_kernels.cuh: chol_diag/panel/update + single-CTA back-substitution) whose cost was launch
// and DRAM-round-trip latency, not arithmetic.
//
// Storage is the same (n+1) x n row-major array: rows 0..n-1 = lower triangle of S, row n = rhs.
// Tiles are 32x32; column tiles j in [0,nb), row tiles i in [0,nb] where row tile nb is the rhs row.
//
//   CTA 0 ("panel CTA") walks the critical path alone and entirely out of shared memory:
//        factor L_kk (one warp, registers) -> publish -> solve L_{k+1,k} -> publish ->
//        A_{k+1,k+1} -= L_{k+1,k} L_{k+1,k}^T -> next k.  Its idle warps prefetch the two tiles
//        it needs next while warp 0 factors.
//   CTAs 1.. ("workers") own the remaining tiles round-robin in column-major tile order, and for
//        every panel k, in that order: solve their tiles of column k (i >= k+2), then apply panel k
//        to their tiles of columns > k.  Every tile is only ever written by its owner (the panel CTA
//        takes over the diagonal / sub-diagonal tile after the owner has published `upd`), so there
//        are no atomics on matrix data.
//   Dependencies travel through three flag arrays in global memory (release/acquire at gpu scope):
//        diag_ready[k], ready[i,k] (L_ik final), upd[i,j] (panels applied by the owner; only read for
//        diagonal and sub-diagonal tiles).  Every CTA executes its tasks in an order compatible with
//        the DAG ((k, tile) lexicographic), all CTAs are co-resident (grid <= occupancy x SMs), so
//        the waits cannot deadlock; a cycle-count bail-out sets `abort`+fail instead of hanging.
//   The same kernel then runs the back-substitution L^T x = y column block by column block, one CTA
//   per block, chained through xready[c].
//
// Matrix data written by another SM is always read with ld.global.cg (L1 is not coherent).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pxr_chol {

constexpr int TB = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
constexpr unsigned long long kXSentinel = 0xFFFFFFFFFFFFFFFFull;   // memset(0xFF): a NaN no computation produces

struct Args {
  double* A;        // (n+1) x n
  double* x;        // n, solution
  int n, nb;
  int* diag_ready;  // [nb]
  int* ready;       // [(nb+1)*nb]
  int* upd;         // [(nb+1)*nb]
  int* xready;      // [nb]
  int* abort;       // [1]
  int* fail_flag;   // set when a pivot is not positive / on bail-out
  long long* trace; // optional [nb+2][8] globaltimer stamps of the panel CTA (PXR_CHOL_TRACE), else null
};

static inline size_t sync_ints(int nb) { return (size_t)2 * (nb + 1) * nb + 2 * (size_t)nb + 1; }

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PXR_CHOL_STAMP(slot) do { if (a.trace && tid == 0) a.trace[(int64_t)k * 8 + (slot)] = gtime(); } while (0)

struct Geo {
  int n, nb;
  __device__ __forceinline__ int row0(int i) const { return i < nb ? i * TB : n; }
  __device__ __forceinline__ int rows(int i) const { return i < nb ? min(TB, n - i * TB) : 1; }
  __device__ __forceinline__ int cols(int j) const { return min(TB, n - j * TB); }
  // column-major enumeration of the tiles i >= j, i in [j, nb]
  __device__ __forceinline__ int64_t off(int j) const { return (int64_t)j * (nb + 1) - (int64_t)j * (j - 1) / 2; }
};

// One thread spins until *p >= target (or the abort flag is up / the cycle budget is gone).
__device__ __forceinline__ void spin_until(const int* p, int target, int* abort_flag, int* fail_flag) {
  long long start = 0;
  unsigned spins = 0;
  while (ld_acquire(p) < target) {
    if ((++spins & 1023u) == 0) {
      if (ld_acquire(abort_flag)) break;
      const long long now = clock64();
      if (start == 0) start = now;
      else if (now - start > 4000000000LL) { st_release(abort_flag, 1); *fail_flag = 1; break; }
    }
  }
}
// Poll a double that the host pre-set to kXSentinel until its producer has stored the value.
__device__ __forceinline__ double poll_value(const double* p, int* abort_flag, int* fail_flag) {
  long long start = 0;
  unsigned spins = 0;
  unsigned long long bits;
  while (true) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory");
    if (bits != kXSentinel) break;
    if ((++spins & 1023u) == 0) {
      if (ld_acquire(abort_flag)) break;
      const long long now = clock64();
      if (start == 0) start = now;
      else if (now - start > 4000000000LL) { st_release(abort_flag, 1); *fail_flag = 1; break; }
    }
  }
  return __longlong_as_double((long long)bits);
}

// Spin (thread `who` only) until *p >= target or the abort flag is up; all threads then pass a barrier.
template <int BAR, int NTHREADS>
__device__ __forceinline__ void wait_flags(const int* p0, int t0, const int* p1, int t1, int* abort_flag, int* fail_flag,
                                           bool leader) {
  if (leader) {
    long long start = 0;
    unsigned spins = 0;
    while (true) {
      const bool ok0 = ld_acquire(p0) >= t0;
      const bool ok1 = (p1 == nullptr) || ld_acquire(p1) >= t1;
      if (ok0 && ok1) break;
      if ((++spins & 1023u) == 0) {
        if (ld_acquire(abort_flag)) break;
        const long long now = clock64();
        if (start == 0) start = now;
        else if (now - start > 4000000000LL) { st_release(abort_flag, 1); *fail_flag = 1; break; }
      }
    }
  }
  if (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(BAR), "r"(NTHREADS) : "memory");
}

// Tile (i, j) -> shared [32][33], zero-filled outside the valid rows/cols; `lower_only` zeroes c > r.
template <int NT>
__device__ __forceinline__ void load_tile(double (*dst)[TB + 1], const double* A, const Geo& g, int i, int j, int tid,
                                          bool lower_only = false) {
  const int r0 = g.row0(i), nr = g.rows(i), c0 = j * TB, ncl = g.cols(j);
  for (int e = tid; e < TB * TB; e += NT) {
    const int r = e >> 5, c = e & 31;
    double v = 0.0;
    if (r < nr && c < ncl && !(lower_only && c > r)) v = __ldcg(A + (int64_t)(r0 + r) * g.n + c0 + c);
    dst[r][c] = v;
  }
}
template <int NT>
__device__ __forceinline__ void store_tile(double (*src)[TB + 1], double* A, const Geo& g, int i, int j, int tid,
                                           bool lower_only = false) {
  const int r0 = g.row0(i), nr = g.rows(i), c0 = j * TB, ncl = g.cols(j);
  for (int e = tid; e < TB * TB; e += NT) {
    const int r = e >> 5, c = e & 31;
    if (r < nr && c < ncl && !(lower_only && c > r)) __stcg(A + (int64_t)(r0 + r) * g.n + c0 + c, src[r][c]);
  }
}

// The whole CTA factors the 32x32 block in `a` (lower triangle valid, identity-padded), two-level:
// four 8-column panels; warp 0 factors a panel with lane = row and the 8 entries of the row in
// registers (8 dependent steps of shuffle + rsqrt + <= 7 rank-1 column updates: ~70 instructions per
// step), then all warps apply the rank-8 update to the trailing block out of shared memory.
// Leaves L (zeros above the diagonal) in `a`, reciprocal pivots in rd; returns false on a bad pivot
// (meaningful in warp 0).  While warp 0 works on the first panel the other warps run `prefetch`.
//
// Why not one warp with the full row in registers: the measured cost of that was the warp's own
// instruction stream (400 instructions per column with a rotating window, or 10k straight-line
// instructions = 160 KB of i-cache when fully unrolled): ~20 us per tile against a ~2.5 us dependency
// chain (32 x [shuffle, rsqrt, multiply, shuffle, fma]).
template <class TryPrefetchFn>
__device__ __forceinline__ bool factor_diag_cta(double (*a)[TB + 1], double* rd, int kb, int tid, TryPrefetchFn try_prefetch) {
  // While warp 0 factors an 8-column panel, warps 1..7 call try_prefetch(blocking=false): one non-blocking look
  // at the flags per panel, the tile loads (224 threads) as soon as they are up.  A late flag therefore overlaps
  // the whole factorisation; only what is still missing at the end is waited for (blocking=true).
  const int warp = tid >> 5, lane = tid & 31;
  bool ok = true;
  bool loaded = false;                               // uniform over warps 1..7
  if (tid >= kb && tid < TB) a[tid][tid] = 1.0;      // identity padding of a ragged last tile (loads zero-fill it)
  __syncthreads();
#pragma unroll 1
  for (int jb = 0; jb < 4; ++jb) {
    const int j0 = jb * 8;
    if (warp == 0) {
      double r[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) r[c] = a[lane][j0 + c];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int j = j0 + c;
        const double dj = __shfl_sync(0xffffffffu, r[c], j);
        if (!(dj > 0.0) || !isfinite(dj)) ok = false;
        const double rj = rsqrt(dj);
        double lij = r[c];
        if (lane == j) { lij = dj * rj; rd[j] = rj; }
        else if (lane > j) lij *= rj;
        r[c] = lij;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
          const double lcj = __shfl_sync(0xffffffffu, lij, j0 + c2);
          if (lane >= j0 + c2) r[c2] -= lij * lcj;
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) a[lane][j0 + c] = (lane >= j0 + c) ? r[c] : 0.0;
    } else if (!loaded) {
      loaded = try_prefetch(false);
    }
    __syncthreads();
    const int n0 = j0 + 8, m = TB - n0;              // trailing m x m block (lower part)
    for (int e = tid; e < m * m; e += kThreads) {
      const int rr = n0 + e / m, cc = n0 + e % m;
      if (rr >= cc) {
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) sacc += a[rr][j0 + q] * a[cc][j0 + q];
        a[rr][cc] -= sacc;
      }
    }
    __syncthreads();
  }
  if (warp != 0 && !loaded) try_prefetch(true);
  __syncthreads();
  return ok;
}

// Rows of `t` (32 rows, lane = column) <- t L^-T, L in `l` with reciprocal pivots rd. Each warp takes
// TB/NW rows and runs them interleaved for ILP.
template <int NW>
__device__ __forceinline__ void solve_rows(double (*t)[TB + 1], double (*l)[TB + 1], const double* rd, int kb, int warp, int lane) {
  constexpr int R = TB / NW;
  double v[R];
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = t[warp * R + m][lane];
  for (int j = 0; j < kb; ++j) {
    const double rj = rd[j];
    const double lj = l[lane][j];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const double xj = __shfl_sync(0xffffffffu, v[m], j) * rj;
      if (lane == j) v[m] = xj;
      else if (lane > j) v[m] -= xj * lj;
    }
  }
#pragma unroll
  for (int m = 0; m < R; ++m) t[warp * R + m][lane] = v[m];
}

// Blocked triangular solve t <- t L^-T with the four 8x8 diagonal blocks of L inverted up front: per 8-column block
// X_b = R_b (L_bb^-1)^T is a small mat-mul and the later blocks are updated with X_b L_cb^T, all 256 threads busy and
// only 4 dependent stages instead of a 32-step substitution chain (2.2 us -> ~0.5 us per tile on B200).
__device__ __forceinline__ void diag_block_inverses(double (*l)[TB + 1], double (*dinv)[8][9], int tid) {
  if (tid < TB) {
    const int b = tid >> 3, c = tid & 7, o = b * 8;
    double rinv[8], x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) rinv[r] = 1.0 / l[o + r][o + r];     // independent: off the substitution chain
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      double sacc = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < r) sacc -= l[o + r][o + k] * x[k];
      x[r] = (r < c) ? 0.0 : sacc * rinv[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) dinv[b][r][c] = x[r];
  }
}
// Thread (r, j) = tid/8, tid%8 only ever touches row r of `t`, and the 8 threads of a row sit in one warp: the stages
// are separated by __syncwarp(), not CTA barriers (l and dinv are read-only here).
__device__ __forceinline__ void solve_rows_blocked(double (*t)[TB + 1], double (*l)[TB + 1], double (*dinv)[8][9], int tid) {
  const int r = tid >> 3, j = tid & 7;              // 256 threads: row r, column j of the current block
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    double rv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rv[k] = t[r][b * 8 + k];
    double xv = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) xv += rv[k] * dinv[b][j][k];      // (L_bb^-1)^T: entries k <= j
    __syncwarp();
    t[r][b * 8 + j] = xv;
    __syncwarp();
    if (b < 3) {
      double xr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) xr[k] = t[r][b * 8 + k];
#pragma unroll
      for (int c = b + 1; c < 4; ++c) {
        double u = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) u += xr[k] * l[c * 8 + j][b * 8 + k];
        t[r][c * 8 + j] -= u;
      }
      __syncwarp();
    }
  }
}

// out[r][c] -= sum_q li[r][q] lj[c][q] for a 2x2 micro-tile per thread (256 threads), out in shared.
__device__ __forceinline__ void syrk_tile_smem(double (*out)[TB + 1], double (*li)[TB + 1], double (*lj)[TB + 1], int tid) {
  const int ty = tid >> 4, tx = tid & 15;
  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll 8
  for (int q = 0; q < TB; ++q) {
    const double a0 = li[ty][q], a1 = li[ty + 16][q], b0 = lj[tx][q], b1 = lj[tx + 16][q];
    s00 += a0 * b0; s01 += a0 * b1; s10 += a1 * b0; s11 += a1 * b1;
  }
  out[ty][tx] -= s00; out[ty][tx + 16] -= s01; out[ty + 16][tx] -= s10; out[ty + 16][tx + 16] -= s11;
}

static __global__ void __launch_bounds__(kThreads, 2) chol_persistent_kernel(Args a) {
  __shared__ double sbuf[3][TB][TB + 1];
  __shared__ double sdinv[4][8][9];
  __shared__ double srd[TB];
  __shared__ double sx[TB];
  __shared__ double sred[kWarps][TB];
  __shared__ int s_ok;
  __shared__ int s_pf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo g{a.n, a.nb};
  const int nb = a.nb;
  const int W = (int)gridDim.x - 1;                 // workers
  double* A = a.A;
  auto RD = [&](int i, int j) { return a.ready + (int64_t)i * nb + j; };
  auto UP = [&](int i, int j) { return a.upd + (int64_t)i * nb + j; };

  if (blockIdx.x == 0) {
    // ------------------------------------------------------------------ panel CTA
    int ia = 0, ib = 1, ic = 2;                      // sbuf roles: A_kk / L_{k+1,k} / next diagonal tile
    load_tile<kThreads>(sbuf[ia], A, g, 0, 0, tid, true);
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
      const int kb = g.cols(k);
      PXR_CHOL_STAMP(0);
      {
        // tiles (k+1,k) and (k+1,k+1) — the owners have applied panels 0..k-1 — land while warp 0 factors
        auto try_prefetch = [&](bool blocking) -> bool {   // warps 1..7 (224 threads, named barrier 1)
          constexpr int NT = kThreads - 32;
          const int t2 = tid - 32;
          if (t2 == 0) {
            if (blocking) {
              spin_until(UP(k + 1, k), k, a.abort, a.fail_flag);
              if (k + 1 < nb) spin_until(UP(k + 1, k + 1), k, a.abort, a.fail_flag);
              s_pf = 1;
            } else {
              const int f0 = ld_acquire(UP(k + 1, k));
              const int f1 = (k + 1 < nb) ? ld_acquire(UP(k + 1, k + 1)) : k;
              s_pf = (f0 >= k && f1 >= k) ? 1 : 0;
            }
            if (s_pf && a.trace) a.trace[(int64_t)k * 8 + 6] = gtime();
          }
          asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
          const bool up = s_pf != 0;
          if (up) {
            load_tile<NT>(sbuf[ib], A, g, k + 1, k, t2);
            if (k + 1 < nb) load_tile<NT>(sbuf[ic], A, g, k + 1, k + 1, t2, true);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");   // s_pf may be rewritten by the next attempt
          return up;
        };
        const bool ok = factor_diag_cta(sbuf[ia], srd, kb, tid, try_prefetch);
        if (!ok && tid == 0) *a.fail_flag = 1;
        PXR_CHOL_STAMP(1);
      }
      __syncthreads();
      PXR_CHOL_STAMP(2);                              // factor done AND both prefetches landed
      diag_block_inverses(sbuf[ia], sdinv, tid);      // warp 0, while the other warps already store L_kk
      store_tile<kThreads>(sbuf[ia], A, g, k, k, tid, true);
      __syncthreads();                                // barrier + release store by one thread is cumulative
      if (tid == 0) st_release(a.diag_ready + k, 1);
      PXR_CHOL_STAMP(3);
      solve_rows_blocked(sbuf[ib], sbuf[ia], sdinv, tid);
      __syncthreads();
      PXR_CHOL_STAMP(4);
      store_tile<kThreads>(sbuf[ib], A, g, k + 1, k, tid);
      __syncthreads();
      if (tid == 0) st_release(RD(k + 1, k), 1);
      PXR_CHOL_STAMP(5);
      if (k + 1 < nb) {
        syrk_tile_smem(sbuf[ic], sbuf[ib], sbuf[ib], tid);
        __syncthreads();
        const int t = ia; ia = ic; ic = t;
      }
    }
  } else if (W > 0) {
    // ------------------------------------------------------------------ workers
    // Per stage k a worker runs, for the tiles it owns:
    //   F   finish the deferred ("bulk") updates that column k+1 still waits for
    //   A   panel k-1 -> columns k and k+1      (what the panel CTA and the solves of stage k need next)
    //   B   bulk updates (panel p -> columns >= p+3, oldest first), one tile at a time, until diag_ready[k]
    //   C   solve its tiles of column k against L_kk, publish them
    // so the latency-critical work never queues behind the O(n^2) trailing updates, which are pre-emptible
    // at tile granularity and may lag by one panel.  Per tile the panels still arrive in increasing order
    // (bulk p <= j-3 is forced before A applies p = j-2), which keeps `upd` a prefix count.
    const int w = (int)blockIdx.x - 1;
    const int64_t total = g.off(nb);                 // number of tiles
    double (*sK)[TB + 1] = sbuf[0];
    double (*sI)[TB + 1] = sbuf[1];
    double (*sJ)[TB + 1] = sbuf[2];
    __shared__ int s_poll[2];
    const int ty = tid >> 4, tx = tid & 15;
    auto first_tile = [&](int j) -> int64_t { const int64_t o = g.off(j); return o + ((w - o) % W + W) % W; };

    // tile (i, j) -= L_ip L_jp^T
    auto do_update = [&](int i, int j, int p) {
      const int r0 = g.row0(i), nr = g.rows(i), c0 = j * TB, ncl = g.cols(j);
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
      const bool v0 = ty < nr, v1 = ty + 16 < nr, u0 = tx < ncl, u1 = tx + 16 < ncl;
      double* p0 = A + (int64_t)(r0 + ty) * g.n + c0 + tx;
      double* p1 = p0 + (int64_t)16 * g.n;
      // the tile itself is owned (no flag): read-modify-write straight from global, 2x2 per thread
      if (v0 && u0) c00 = __ldcg(p0);
      if (v0 && u1) c01 = __ldcg(p0 + 16);
      if (v1 && u0) c10 = __ldcg(p1);
      if (v1 && u1) c11 = __ldcg(p1 + 16);
      wait_flags<0, kThreads>(RD(i, p), 1, (i != j) ? RD(j, p) : nullptr, 1, a.abort, a.fail_flag, tid == 0);
      load_tile<kThreads>(sI, A, g, i, p, tid);
      if (i != j) load_tile<kThreads>(sJ, A, g, j, p, tid);
      __syncthreads();
      double (*lj)[TB + 1] = (i != j) ? sJ : sI;
#pragma unroll 8
      for (int q = 0; q < TB; ++q) {
        const double a0 = sI[ty][q], a1 = sI[ty + 16][q], b0 = lj[tx][q], b1 = lj[tx + 16][q];
        c00 -= a0 * b0; c01 -= a0 * b1; c10 -= a1 * b0; c11 -= a1 * b1;
      }
      if (v0 && u0) __stcg(p0, c00);
      if (v0 && u1) __stcg(p0 + 16, c01);
      if (v1 && u0) __stcg(p1, c10);
      if (v1 && u1) __stcg(p1 + 16, c11);
      const bool watched = (i == j) || (i == j + 1);
      __syncthreads();                                // sI/sJ free again; orders the tile stores before the release
      if (watched && tid == 0) st_release(UP(i, j), p + 1);
    };

    // bulk cursor: panel bp, my next tile bt (column bj), columns >= bp + 3
    int bp = 0, bj = 3;
    int64_t bt = (3 < nb) ? first_tile(3) : total;
    auto bulk_has = [&](int limit_p) -> bool {        // a deferred tile of a panel <= limit_p is left
      while (bp <= limit_p && bt >= total) { ++bp; bj = bp + 3; bt = (bj < nb) ? first_tile(bj) : total; }
      return bp <= limit_p;
    };
    auto bulk_col = [&]() -> int { while (bt >= g.off(bj + 1)) ++bj; return bj; };
    auto bulk_step = [&]() {
      const int j = bulk_col();
      do_update(j + (int)(bt - g.off(j)), j, bp);
      bt += W;
    };

    for (int k = 0; k < nb; ++k) {
      const int kb = g.cols(k);
      // F: everything deferred that column k+1 (and older) still needs: bulk through (panel k-2, column k+1)
      while (bulk_has(k - 2)) {
        if (bp == k - 2 && bulk_col() > k + 1) break;
        bulk_step();
      }
      // A: panel k-1 -> columns k, k+1
      if (k >= 1) {
        for (int j = k; j <= k + 1 && j < nb; ++j)
          for (int64_t t = first_tile(j); t < g.off(j + 1); t += W) {
            const int i = j + (int)(t - g.off(j));
            if (i == j && j == k) continue;           // (k,k) -= L_{k,k-1} L_{k,k-1}^T is the panel CTA's
            do_update(i, j, k - 1);
          }
      }
      // B: deferred updates until L_kk is there; one look at the flag per tile, so the solves below start at
      // most one tile (~2.5 us) after diag_ready[k]
      while (bulk_has(k - 1)) {
        if (tid == 0) s_poll[0] = ld_acquire(a.diag_ready + k);
        __syncthreads();
        const int rdy = s_poll[0];
        __syncthreads();                              // every thread has its copy before s_poll can be rewritten (a worker
        if (rdy != 0) break;                          // without tiles in the next stages reaches the next poll barrier-free)
        bulk_step();
      }
      // C: solves of column k
      bool have_lkk = false;
      for (int64_t t = first_tile(k); t < g.off(k + 1); t += W) {
        const int i = k + (int)(t - g.off(k));
        if (i < k + 2) continue;                      // diagonal and sub-diagonal tile: panel CTA
        load_tile<kThreads>(sI, A, g, i, k, tid);    // own tile (all its updates are this CTA's): no flag needed
        if (!have_lkk) {
          wait_flags<0, kThreads>(a.diag_ready + k, 1, nullptr, 0, a.abort, a.fail_flag, tid == 0);
          load_tile<kThreads>(sK, A, g, k, k, tid, true);
          __syncthreads();
          if (tid >= kb && tid < TB) sK[tid][tid] = 1.0;      // identity padding of a ragged last tile
          __syncthreads();
          diag_block_inverses(sK, sdinv, tid);
          have_lkk = true;
        }
        __syncthreads();
        solve_rows_blocked(sI, sK, sdinv, tid);
        __syncthreads();
        store_tile<kThreads>(sI, A, g, i, k, tid);
        __syncthreads();
        if (tid == 0) st_release(RD(i, k), 1);
      }
    }
    while (bulk_has(nb - 1)) bulk_step();            // nothing is left in a complete run; keeps the invariant explicit
  }

  // -------------------------------------------------------------------- back-substitution L^T x = y
  // y_c = row n, columns of block c  (tile (nb, c));  x_c = L_cc^-T (y_c - sum_{j>c} L_jc^T x_j)
  // One CTA per column block.  Everything that does not depend on x is done up front (flags of column c,
  // L_cc^-T formed explicitly so the last step is a mat-vec, not a 32-step substitution).  x itself is
  // the message: the host pre-fills x with an all-ones NaN pattern, the producer stores plain doubles,
  // every consumer warp polls the 32 values it needs — one L2 round trip per block, no flag, no fence.
  const int G = (int)gridDim.x;
  if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[(int64_t)nb * 8] = gtime();
  for (int c = nb - 1 - (int)blockIdx.x; c >= 0; c -= G) {
    const int kb = g.cols(c), c0 = c * TB;
    const int wr = tid >> 5;                         // row group: rows wr, wr+8, wr+16, wr+24 of a tile
    // ---- phase A: column c is final once diag_ready[c] and ready[j,c] for j in (c, nb]
    for (int t = tid; t <= nb - c; t += kThreads)
      spin_until(t == 0 ? a.diag_ready + c : RD(c + t, c), 1, a.abort, a.fail_flag);
    __syncthreads();
    load_tile<kThreads>(sbuf[0], A, g, c, c, tid, true);
    for (int e = tid; e < TB * TB; e += kThreads) sbuf[1][e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
    __syncthreads();
    if (tid < TB) srd[tid] = 1.0 / (tid < kb ? sbuf[0][tid][tid] : 1.0);
    __syncthreads();
    solve_rows<kWarps>(sbuf[1], sbuf[0], srd, kb, warp, lane);      // I L^-T : sbuf[1] = L_cc^-T (upper triangular)
    const double yv = (warp == 0 && lane < kb) ? __ldcg(A + (int64_t)g.n * g.n + c0 + lane) : 0.0;
    // ---- phase B: accumulate L_jc^T x_j as the x_j arrive (j descending)
    double acc = 0.0;
    for (int j = nb - 1; j > c; --j) {
      const int r0 = j * TB, nr = g.rows(j);
      double l[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int r = wr + 8 * m;
        l[m] = (r < nr && lane < kb) ? __ldcg(A + (int64_t)(r0 + r) * g.n + c0 + lane) : 0.0;
      }
      const double xv = lane < nr ? poll_value(a.x + r0 + lane, a.abort, a.fail_flag) : 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) acc += l[m] * __shfl_sync(0xffffffffu, xv, wr + 8 * m);
    }
    sred[wr][lane] = acc;
    __syncthreads();
    // ---- phase C: x_c = L_cc^-T (y_c - s)
    if (warp == 0) {
      double v = yv;
#pragma unroll
      for (int m = 0; m < kWarps; ++m) v -= sred[m][lane];
      double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
      for (int q = 0; q < TB; q += 4) {
        x0 += sbuf[1][lane][q] * __shfl_sync(0xffffffffu, v, q);
        x1 += sbuf[1][lane][q + 1] * __shfl_sync(0xffffffffu, v, q + 1);
        x2 += sbuf[1][lane][q + 2] * __shfl_sync(0xffffffffu, v, q + 2);
        x3 += sbuf[1][lane][q + 3] * __shfl_sync(0xffffffffu, v, q + 3);
      }
      if (lane < kb) st_relaxed_f64(a.x + c0 + lane, (x0 + x1) + (x2 + x3));
      if (a.trace && lane == 0) a.trace[(int64_t)nb * 8 + 8 + c] = gtime();   // x_c published
    }
    __syncthreads();
  }
  (void)s_ok;
}

}  // namespace pxr_chol
