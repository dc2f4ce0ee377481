// pxr_chol2.cuh — dense Cholesky of the reduced camera system, second design: a BAND CTA walks the critical path, the
// rest of the grid trails it.
//
// Same job and storage as pxr_chol.cuh (exact DENSE_SCHUR / SPARSE_SCHUR step of ceres::Solve, reference call site
// bundle_adjustment/src/bundle_optimizer.h:181-191,224): (n+1) x n row-major, rows 0..n-1 = lower triangle of S,
// row n = rhs; 32x32 tiles; one persistent launch factors, forward- and back-substitutes.
//
// What limited the first design (profiles/chol_trace_r01.txt): per 32-column step the panel CTA needed two tiles whose
// last update came from the step before — two dependent hops through other CTAs and global memory (~10 us) — and its
// own factorisation shared barriers with the prefetch of those tiles.  Here
//   * CTA 0 owns the band i - j <= 2 and works row by row, left-looking inside the band:
//         a  L(r,r-2) = A'(r,r-2) L(r-2,r-2)^-T                      helper warps
//         b  A'(r,r-1) -= L(r,r-2) L(r-1,r-2)^T                      helper warps
//         c  A'(r,r)   -= L(r,r-2) L(r,r-2)^T                        helper warps
//         d  L(r,r-1) = A'(r,r-1) L(r-1,r-1)^-T                      chain warps
//         e  A'(r,r)  -= L(r,r-1) L(r,r-1)^T                         chain warps
//         f  L(r,r) = chol(A'(r,r)), inverses of its 8x8 diagonal blocks   chain warps
//     so the chain (4 warps, shared memory only, no global memory access at all) is  d e f d e f ...  and what it needs
//     from outside — row r's three tiles with the panels p <= r-3 applied — was finished by the workers two rows earlier.
//     The helper warps (4) run a small event loop: publish finished tiles first (diagonal, then L(r,r-1), L(r,r-2)),
//     then steps a-c for the next row, then fetch the row after that when its flags are up.
//   * workers (all other CTAs) own the remaining tiles round-robin.  Per panel p, in this order: forced remainder of
//     old trailing updates, trailing updates of panel p-1 while polling diag_ready[p], the solves of column p
//     (rows >= p+3), the "urgent" updates of panel p (columns p+1..p+4, the band's next row first).  Updates of one tile
//     commute, so there is no per-tile order to keep: band tiles carry a COUNT of applied panels, the band waits for
//     count == r-2.
// Flags (global, release/acquire at gpu scope): diag_ready[k], ready[i,k] (L_ik final), upd[i,j] (panels applied to a
// band tile).  Every wait has the cycle-budget bail-out of pxr_chol.cuh (abort word + fail flag -> the host falls back
// to the launch-per-panel path), all CTAs must be co-resident.
#pragma once
#include "pxr_chol.cuh"

namespace pxr_chol2 {

using pxr_chol::Args;
using pxr_chol::Geo;
using pxr_chol::TB;
using pxr_chol::kThreads;
using pxr_chol::kWarps;
using pxr_chol::ld_acquire;
using pxr_chol::st_release;
using pxr_chol::spin_until;
using pxr_chol::poll_value;
using pxr_chol::st_relaxed_f64;
using pxr_chol::gtime;

constexpr int kGroup = 128;              // threads per warp group of the band CTA
constexpr int kWorker = 64;              // threads per worker
constexpr int kWorkersPerCta = kThreads / kWorker;
constexpr int kRowSlots = 4;             // rows of band tiles kept in shared memory
typedef double Tile[TB][TB + 1];

struct Smem {
  Tile t[kRowSlots][3];                  // [row & 3][0: (r,r-2), 1: (r,r-1), 2: (r,r)]   (workers: t[0][0..2] = K, I, J)
  double dinv[kRowSlots][4][8][9];       // inverses of the 8x8 diagonal blocks of L(r,r)
  double rd[TB];
  double sx[TB];
  double sred[kWarps][TB];
  volatile int fact, dsolved, helped;    // band CTA: rows factored, rows whose step d is done, rows whose steps a-c are done
  int bc[4];                             // broadcast words (one per worker of the CTA / the helper group's action)
};
static inline size_t smem_bytes() { return sizeof(Smem); }

__device__ __forceinline__ void bar_group(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kGroup) : "memory"); }

// one thread of the group spins on a shared-memory word, then the group's barrier
__device__ __forceinline__ void wait_smem(volatile int* p, int target, int* abort_flag, int* fail_flag, bool leader, int bar) {
  if (leader) {
    long long start = 0; unsigned spins = 0;
    while (*p < target) {
      if ((++spins & 4095u) == 0) {
        if (ld_acquire(abort_flag)) break;
        const long long now = clock64();
        if (start == 0) start = now;
        else if (now - start > 4000000000LL) { st_release(abort_flag, 1); *fail_flag = 1; break; }
      }
    }
    __threadfence_block();       // acquire side: the other group's tile writes precede its flag write
  }
  bar_group(bar);
}

template <int NT>
__device__ __forceinline__ void ld_tile(Tile dst, const double* A, const Geo& g, int i, int j, int tid, bool lower_only = false) {
  const int r0 = g.row0(i), nr = g.rows(i), c0 = j * TB, ncl = g.cols(j);
  for (int e = tid; e < TB * TB; e += NT) {
    const int r = e >> 5, c = e & 31;
    double v = 0.0;
    if (r < nr && c < ncl && !(lower_only && c > r)) v = __ldcg(A + (int64_t)(r0 + r) * g.n + c0 + c);
    dst[r][c] = v;
  }
}
template <int NT>
__device__ __forceinline__ void st_tile(Tile src, double* A, const Geo& g, int i, int j, int tid, bool lower_only = false) {
  const int r0 = g.row0(i), nr = g.rows(i), c0 = j * TB, ncl = g.cols(j);
  for (int e = tid; e < TB * TB; e += NT) {
    const int r = e >> 5, c = e & 31;
    if (r < nr && c < ncl && !(lower_only && c > r)) __stcg(A + (int64_t)(r0 + r) * g.n + c0 + c, src[r][c]);
  }
}

// inverses of the four 8x8 diagonal blocks of the lower-triangular L (32 threads: t = block*8 + column)
__device__ __forceinline__ void diag_block_inverses(Tile l, double (*dinv)[8][9], int t) {
  if (t < TB) {
    const int b = t >> 3, c = t & 7, o = b * 8;
    double rinv[8], x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) rinv[r] = 1.0 / l[o + r][o + r];
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

// t <- t L^-T for the 32 rows of t, NT threads (NT/8 rows at a time, 8 threads per row inside one warp: stages are
// separated by __syncwarp; l and dinv are read-only here)
template <int NT>
__device__ __forceinline__ void solve_rows_blocked(Tile t, Tile l, double (*dinv)[8][9], int tid) {
  const int j = tid & 7;
  for (int r = tid >> 3; r < TB; r += NT / 8) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double rv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) rv[k] = t[r][b * 8 + k];
      double xv = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) xv += rv[k] * dinv[b][j][k];
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
}

// out -= a b^T (all 32x32 in shared memory), NT threads, 2x2 micro-tiles
template <int NT>
__device__ __forceinline__ void gemm_nt_sub(Tile out, Tile a, Tile b, int tid) {
  for (int m = tid; m < 256; m += NT) {
    const int ty = m >> 4, tx = m & 15;
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll 8
    for (int q = 0; q < TB; ++q) {
      const double a0 = a[ty][q], a1 = a[ty + 16][q], b0 = b[tx][q], b1 = b[tx + 16][q];
      s00 += a0 * b0; s01 += a0 * b1; s10 += a1 * b0; s11 += a1 * b1;
    }
    out[ty][tx] -= s00; out[ty][tx + 16] -= s01; out[ty + 16][tx] -= s10; out[ty + 16][tx + 16] -= s11;
  }
}

// The chain group (128 threads, barrier `bar`) factors the 32x32 block in `a` (lower triangle valid): four 8-column
// panels, warp 0 factors a panel with lane = row and the row's 8 entries in registers, all four warps apply the rank-8
// update to the trailing block.  Leaves L (zeros above the diagonal) in `a`; returns false on a bad pivot (warp 0).
// (A variant with the 8x8 diagonal blocks factored redundantly in registers, an fp32-seeded Newton rsqrt and a
//  look-ahead by the chain warp was written and did not pass the parity tests in the one run it got; this is the
//  version that did.)
__device__ __forceinline__ bool factor_diag_group(Tile a, double* rd, int kb, int tid, int bar) {
  const int warp = tid >> 5, lane = tid & 31;
  bool ok = true;
  if (tid >= kb && tid < TB) a[tid][tid] = 1.0;      // identity padding of a ragged last tile
  bar_group(bar);
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
    }
    bar_group(bar);
    const int n0 = j0 + 8, m = TB - n0;
    for (int e = tid; e < m * m; e += kGroup) {
      const int rr = n0 + e / m, cc = n0 + e % m;
      if (rr >= cc) {
        double sacc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) sacc += a[rr][j0 + q] * a[cc][j0 + q];
        a[rr][cc] -= sacc;
      }
    }
    bar_group(bar);
  }
  return ok;
}

#define PXR_CHOL2_STAMP(row, slot) do { if (a.trace && gt == 0) a.trace[(int64_t)(row) * 8 + (slot)] = gtime(); } while (0)

static __global__ void __launch_bounds__(kThreads, 2) chol_band_kernel(Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo g{a.n, a.nb};
  const int nb = a.nb;
  const int W = ((int)gridDim.x - 1) * kWorkersPerCta;   // workers
  double* A = a.A;
  auto RD = [&](int i, int j) { return a.ready + (int64_t)i * nb + j; };
  auto UP = [&](int i, int j) { return a.upd + (int64_t)i * nb + j; };

  if (blockIdx.x == 0) {
    if (tid == 0) { S.fact = 0; S.dsolved = 0; S.helped = 0; }
    __syncthreads();
    const int gt = tid & (kGroup - 1);               // thread index inside its group
    if (tid < kGroup) {
      // ---------------------------------------------------------------- chain group: d e f, shared memory only
      for (int r = 0; r <= nb; ++r) {
        const int slot = r & 3;
        wait_smem(&S.helped, r + 1, a.abort, a.fail_flag, gt == 0, 1);
        PXR_CHOL2_STAMP(r, 0);
        if (r >= 1) {
          solve_rows_blocked<kGroup>(S.t[slot][1], S.t[(r - 1) & 3][2], S.dinv[(r - 1) & 3], gt);    // d
          bar_group(1);
        }
        if (gt == 0) { __threadfence_block(); S.dsolved = r + 1; }
        PXR_CHOL2_STAMP(r, 1);
        if (r < nb) {
          if (r >= 1) { gemm_nt_sub<kGroup>(S.t[slot][2], S.t[slot][1], S.t[slot][1], gt); bar_group(1); }   // e
          PXR_CHOL2_STAMP(r, 2);
          const bool ok = factor_diag_group(S.t[slot][2], S.rd, g.cols(r), gt, 1);                   // f
          if (!ok && gt == 0) *a.fail_flag = 1;
          diag_block_inverses(S.t[slot][2], S.dinv[slot], gt);
          bar_group(1);
          if (gt == 0) { __threadfence_block(); S.fact = r + 1; }
          PXR_CHOL2_STAMP(r, 3);
        }
      }
    } else {
      // ---------------------------------------------------------------- helper group: event loop
      int pub_diag = 0;          // rows whose diagonal tile has been stored / published
      int pub_x1 = 1;            // next row whose L(r,r-1) is to be published (rows 1..nb)
      int hr = 0;                // next row for steps a-c
      bool a_done = false;       // step a of row hr done
      int loaded = 0;            // rows fetched into shared memory so far
      const bool leader = gt == 0;
      long long idle_start = 0;
      while (pub_diag < nb || pub_x1 <= nb || hr <= nb) {
        // ---- decide (leader), broadcast
        if (leader) {
          const int fact = S.fact, dsolved = S.dsolved;
          __threadfence_block();   // acquire side of the chain group's flag writes
          int act = 0;
          if (pub_diag < fact) act = 1;                                           // publish L(r,r)
          else if (pub_x1 < dsolved && pub_x1 <= nb) act = 2;                      // publish L(r,r-1)
          else if (hr <= nb && hr < loaded && !a_done && (hr < 2 || fact >= hr - 1)) act = 3;          // step a (+ publish L(r,r-2))
          else if (hr <= nb && hr < loaded && a_done && (hr < 1 || dsolved >= hr)) act = 4;            // steps b, c
          else if (loaded <= nb && loaded <= hr + 1 && (loaded < 4 || (fact >= loaded - 2 && pub_diag >= loaded - 3 && pub_x1 >= loaded - 2))) {
            // fetch row `loaded` when the workers have applied panels 0..loaded-3 to its tiles
            const int L = loaded, need = L - 2;
            bool up = true;
            if (need > 0) {
              if (L >= 2 && ld_acquire(UP(L, L - 2)) < need) up = false;
              if (up && L >= 1 && ld_acquire(UP(L, L - 1)) < need) up = false;
              if (up && L < nb && ld_acquire(UP(L, L)) < need) up = false;
            }
            if (up) act = 5;
          }
          if (act == 0) {
            if (ld_acquire(a.abort)) act = 9;
            else {
              const long long now = clock64();
              if (idle_start == 0) idle_start = now;
              else if (now - idle_start > 4000000000LL) { st_release(a.abort, 1); *a.fail_flag = 1; act = 9; }
            }
          } else idle_start = 0;
          S.bc[0] = act;
        }
        bar_group(2);
        const int act = S.bc[0];
        bar_group(2);
        if (act == 9) break;
        if (act == 1) {
          const int r = pub_diag;
          st_tile<kGroup>(S.t[r & 3][2], A, g, r, r, gt, true);
          bar_group(2);
          if (leader) { st_release(a.diag_ready + r, 1); if (a.trace) a.trace[(int64_t)r * 8 + 4] = gtime(); }
          ++pub_diag;
        } else if (act == 2) {
          const int r = pub_x1;
          st_tile<kGroup>(S.t[r & 3][1], A, g, r, r - 1, gt);
          bar_group(2);
          if (leader) st_release(RD(r, r - 1), 1);
          ++pub_x1;
        } else if (act == 3) {
          const int r = hr;
          if (r >= 2) {
            solve_rows_blocked<kGroup>(S.t[r & 3][0], S.t[(r - 2) & 3][2], S.dinv[(r - 2) & 3], gt);     // a
            bar_group(2);
            st_tile<kGroup>(S.t[r & 3][0], A, g, r, r - 2, gt);
            bar_group(2);
            if (leader) st_release(RD(r, r - 2), 1);
          }
          a_done = true;
        } else if (act == 4) {
          const int r = hr;
          if (r >= 2) {
            gemm_nt_sub<kGroup>(S.t[r & 3][1], S.t[r & 3][0], S.t[(r - 1) & 3][1], gt);                  // b
            if (r < nb) gemm_nt_sub<kGroup>(S.t[r & 3][2], S.t[r & 3][0], S.t[r & 3][0], gt);            // c
            bar_group(2);
          }
          if (leader) { __threadfence_block(); S.helped = r + 1; if (a.trace) a.trace[(int64_t)r * 8 + 5] = gtime(); }
          ++hr; a_done = false;
        } else if (act == 5) {
          const int L = loaded;
          if (L >= 2) ld_tile<kGroup>(S.t[L & 3][0], A, g, L, L - 2, gt);
          if (L >= 1) ld_tile<kGroup>(S.t[L & 3][1], A, g, L, L - 1, gt);
          if (L < nb) ld_tile<kGroup>(S.t[L & 3][2], A, g, L, L, gt, true);
          bar_group(2);
          if (leader && a.trace) a.trace[(int64_t)L * 8 + 6] = gtime();
          ++loaded;
        }
      }
    }
    __syncthreads();
  } else if (W > 0) {
    // ------------------------------------------------------------------ workers: four independent 64-thread workers per CTA
    // (a 32x32x32 tile update is latency-, not throughput-bound: 8 workers per SM instead of 2 CTAs hide it)
    const int wg = tid >> 6, wt = tid & (kWorker - 1);
    const int w = ((int)blockIdx.x - 1) * kWorkersPerCta + wg;
    const int bar_id = 1 + wg;
    auto bar_w = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kWorker) : "memory"); };
    const int64_t total = g.off(nb);                 // number of tiles (i >= j, i in [j, nb])
    Tile& sK = S.t[wg][0];
    Tile& sI = S.t[wg][1];
    Tile& sJ = S.t[wg][2];
    const int ty = wt >> 3, tx = wt & 7;             // 4x4 micro-tile: rows ty + 8a, columns tx + 8b
    // my tiles of column j: t = first_tile(j), t += W while t < off(j+1); row i = j + (t - off(j))
    auto first_tile = [&](int j) -> int64_t { const int64_t o = g.off(j); return o + ((w - o) % W + W) % W; };
    // one thread spins until both flags are up (or abort), then the worker's barrier
    auto wait2 = [&](const int* p0, const int* p1) {
      if (wt == 0) {
        long long start = 0; unsigned spins = 0;
        while (true) {
          const bool ok0 = ld_acquire(p0) >= 1;
          const bool ok1 = (p1 == nullptr) || ld_acquire(p1) >= 1;
          if (ok0 && ok1) break;
          if ((++spins & 1023u) == 0) {
            if (ld_acquire(a.abort)) break;
            const long long now = clock64();
            if (start == 0) start = now;
            else if (now - start > 4000000000LL) { st_release(a.abort, 1); *a.fail_flag = 1; break; }
          }
        }
      }
      bar_w();
    };

    // tile (i, j) -= L_ip L_jp^T   (i - p >= 3; j in (p, i])
    auto do_update = [&](int i, int j, int p) {
      const int r0 = g.row0(i), nr = g.rows(i), c0 = j * TB, ncl = g.cols(j);
      double c[4][4];
      double* base = A + (int64_t)r0 * g.n + c0;
#pragma unroll
      for (int ra = 0; ra < 4; ++ra)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          const int r = ty + 8 * ra, cc = tx + 8 * cb;
          c[ra][cb] = (r < nr && cc < ncl) ? __ldcg(base + (int64_t)r * g.n + cc) : 0.0;
        }
      wait2(RD(i, p), (i != j) ? RD(j, p) : nullptr);
      ld_tile<kWorker>(sI, A, g, i, p, wt);
      if (i != j) ld_tile<kWorker>(sJ, A, g, j, p, wt);
      bar_w();
      Tile& lj = (i != j) ? sJ : sI;
#pragma unroll 4
      for (int q = 0; q < TB; ++q) {
        double av[4], bv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { av[k] = sI[ty + 8 * k][q]; bv[k] = lj[tx + 8 * k][q]; }
#pragma unroll
        for (int ra = 0; ra < 4; ++ra)
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) c[ra][cb] -= av[ra] * bv[cb];
      }
#pragma unroll
      for (int ra = 0; ra < 4; ++ra)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          const int r = ty + 8 * ra, cc = tx + 8 * cb;
          if (r < nr && cc < ncl) __stcg(base + (int64_t)r * g.n + cc, c[ra][cb]);
        }
      bar_w();                                        // sI/sJ free again; orders the tile stores before the release
      if (i - j <= 2 && wt == 0) st_release(UP(i, j), *UP(i, j) + 1);     // band tile: one more panel applied (single writer)
    };

    // trailing ("bulk") updates: panel bp -> columns >= bp + 5, my tiles in column-major order
    int bp = 0, bj = 5;
    int64_t bt = (5 < nb) ? first_tile(5) : total;
    auto bulk_has = [&](int limit_p) -> bool {        // a deferred tile of a panel <= limit_p is left
      while (bp <= limit_p && bt >= total) { ++bp; bj = bp + 5; bt = (bj < nb) ? first_tile(bj) : total; }
      return bp <= limit_p;
    };
    auto bulk_step = [&]() {
      while (bt >= g.off(bj + 1)) ++bj;
      do_update(bj + (int)(bt - g.off(bj)), bj, bp);
      bt += W;
    };

    for (int p = 0; p < nb; ++p) {
      const int kb = g.cols(p);
      // F: column p is solved below: every trailing update it still misses (panels <= p-5) first
      while (bulk_has(p - 5)) bulk_step();
      // B: trailing updates of the panels before p until L_pp is there (one look at the flag per tile)
      while (bulk_has(p - 1)) {
        if (wt == 0) S.bc[wg] = ld_acquire(a.diag_ready + p);
        bar_w();
        const int rdy = S.bc[wg];
        bar_w();
        if (rdy != 0) break;
        bulk_step();
      }
      // C: solves of column p, rows >= p + 3
      bool have_lpp = false;
      for (int64_t t = first_tile(p); t < g.off(p + 1); t += W) {
        const int i = p + (int)(t - g.off(p));
        if (i < p + 3) continue;                      // band tiles
        ld_tile<kWorker>(sI, A, g, i, p, wt);        // own tile (all its updates are this worker's): no flag needed
        if (!have_lpp) {
          wait2(a.diag_ready + p, nullptr);
          ld_tile<kWorker>(sK, A, g, p, p, wt, true);
          bar_w();
          if (wt >= kb && wt < TB) sK[wt][wt] = 1.0;
          bar_w();
          diag_block_inverses(sK, S.dinv[wg], wt);
          have_lpp = true;
        }
        bar_w();
        solve_rows_blocked<kWorker>(sI, sK, S.dinv[wg], wt);
        bar_w();
        st_tile<kWorker>(sI, A, g, i, p, wt);
        bar_w();
        if (wt == 0) st_release(RD(i, p), 1);
      }
      // A: urgent updates of panel p: columns p+1 .. p+4, the band's next row (i = p + 3) first
      for (int pass = 0; pass < 2; ++pass)
        for (int j = p + 1; j <= p + 4 && j < nb; ++j)
          for (int64_t t = first_tile(j); t < g.off(j + 1); t += W) {
            const int i = j + (int)(t - g.off(j));
            if (i < p + 3) continue;
            if ((pass == 0) != (i == p + 3)) continue;
            do_update(i, j, p);
          }
    }
    while (bulk_has(nb - 1)) bulk_step();
    __syncthreads();
  }

  // -------------------------------------------------------------------- back-substitution L^T x = y  (as pxr_chol.cuh)
  const int G = (int)gridDim.x;
  if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[(int64_t)(nb + 1) * 8] = gtime();
  Tile& b0 = S.t[0][0];
  Tile& b1 = S.t[0][1];
  for (int c = nb - 1 - (int)blockIdx.x; c >= 0; c -= G) {
    const int kb = g.cols(c), c0 = c * TB;
    const int wr = tid >> 5;
    for (int t = tid; t <= nb - c; t += kThreads)
      spin_until(t == 0 ? a.diag_ready + c : RD(c + t, c), 1, a.abort, a.fail_flag);
    __syncthreads();
    ld_tile<kThreads>(b0, A, g, c, c, tid, true);
    for (int e = tid; e < TB * TB; e += kThreads) b1[e >> 5][e & 31] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
    __syncthreads();
    if (tid < TB) S.rd[tid] = 1.0 / (tid < kb ? b0[tid][tid] : 1.0);
    __syncthreads();
    pxr_chol::solve_rows<kWarps>(b1, b0, S.rd, kb, warp, lane);      // I L^-T : b1 = L_cc^-T
    const double yv = (warp == 0 && lane < kb) ? __ldcg(A + (int64_t)g.n * g.n + c0 + lane) : 0.0;
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
    S.sred[wr][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      double v = yv;
#pragma unroll
      for (int m = 0; m < kWarps; ++m) v -= S.sred[m][lane];
      double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
      for (int q = 0; q < TB; q += 4) {
        x0 += b1[lane][q] * __shfl_sync(0xffffffffu, v, q);
        x1 += b1[lane][q + 1] * __shfl_sync(0xffffffffu, v, q + 1);
        x2 += b1[lane][q + 2] * __shfl_sync(0xffffffffu, v, q + 2);
        x3 += b1[lane][q + 3] * __shfl_sync(0xffffffffu, v, q + 3);
      }
      if (lane < kb) st_relaxed_f64(a.x + c0 + lane, (x0 + x1) + (x2 + x3));
      if (a.trace && lane == 0) a.trace[(int64_t)(nb + 1) * 8 + 8 + c] = gtime();
    }
    __syncthreads();
  }
}

}  // namespace pxr_chol2
