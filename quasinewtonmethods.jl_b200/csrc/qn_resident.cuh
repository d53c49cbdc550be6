// qn_resident.cuh — resident BFGS kernels, thread-per-row layout: one CTA owns one
// problem for its whole run (QN.jl:237-292 entirely on the device), B⁻¹ never leaves
// the SM.
//
// Thread c owns element c of every vector and ROW c of B⁻¹ (the reference's
// unit-stride index, QN.jl:48,61-62).  Two storage policies for the row:
//
//   NREG > 0  register-resident, n = NREG compile-time, row in NREG fp64 registers.
//             Used for small n (<= 32); for 33..64 the register-TILE kernel in
//             qn_resident_tile.cuh is faster (it re-uses every shared-memory
//             broadcast for four rows).
//   NREG == 0 shared-memory-resident: runtime n, B⁻¹ in shared memory in the
//             reference's own column-major layout H[c + n r]; a warp's access to one
//             column is 32 consecutive doubles (conflict-free).
//             Bound: shared-memory bandwidth (24 n² bytes per iteration).
//
// The grid is persistent: gridDim = (#SMs × resident CTAs per SM) and CTAs pull
// problem indices from an atomic counter, so the spread of iteration counts
// across starts (≈300-680 at n = 60) does not idle SMs.
#pragma once

#include "qn_device.cuh"
#include "qn_objectives.cuh"

namespace qn {

// packed per-column operands of the fused update+GEMV pass (QN.jl:56-67):
// w = c₁ s[r], nu = -u[r], s = s[r], g = ∇_new[r]; 32 bytes = two LDS.128.
struct __align__(16) Pass2Col {
  double w, nu, s, g;
};

__host__ __device__ inline int resident_block_threads(int n) { return (n + 31) & ~31; }

// shared memory carve-up (in doubles).  nthreads = blockDim.x.  Shared-memory policy: B⁻¹ is
// stored column-major with an EVEN column stride (resident_ldh) so a thread's pair of rows is one
// 16-byte access; with nthreads > B the rule-C2 blocks of the rows are dealt out to thread
// groups whose block partials meet in `part` ([blocks][B]).
struct ResidentSmem {
  int red, xs, ys, cols, scratch, part, part2, rowp, h, total;
};
__host__ __device__ inline int resident_ldh(int n, int pr = 2) { return (n + pr - 1) & ~(pr - 1); }
template <class Obj>
__host__ __device__ inline ResidentSmem resident_smem_layout(int n, bool h_in_smem, int nthreads,
                                                             const typename Obj::Params& op, bool onepass = false,
                                                             int pr = 2, int hybrid_rows = 0) {
  const int B = resident_block_threads(n);
  const int nblocks = (n + chain_block(n) - 1) / chain_block(n);
  ResidentSmem L;
  int o = 0;
  L.cols = o; o += 4 * B;   // Pass2Col[B], 16-byte aligned at offset 0
  L.ys = o;   o += B;
  L.xs = o;   o += B;
  L.red = o;  o += onepass ? 8 * B : 2 * B; // two alternating staging buffers (+ two of three values each for csum3)
  L.scratch = o; o += Obj::scratch_doubles(n, op);
  L.part = o; o += (nthreads > B) ? nblocks * B : 0;
  L.part2 = o; o += onepass ? nblocks * B : 0; // one-pass mode: partials of the second GEMV (w = B⁻¹∇)
  L.rowp = o;  o += onepass ? 2 * B : 0;       //   and the pending s / -u by element (row operands)
  // B⁻¹ itself (shared-memory policy), or — register-tile policy with hybrid_rows of every thread's
  // pr rows kept here — those rows, packed: [(B/pr)·hybrid_rows] × n
  L.h = o;    o += h_in_smem ? resident_ldh(n, pr) * n : (B / pr) * hybrid_rows * n;
  L.total = o + 2; // + problem index broadcast slot
  return L;
}

// linesearch!  QN.jl:165-232.  Every thread of the team runs the same scalar code on
// identical values, so control flow is uniform.  x, s: this thread's elements of
// x_old and the search direction.  Returns α (0 = failure, QN.jl:193).
template <class Obj>
__device__ __forceinline__ double team_linesearch(Team& tm, const typename Obj::Params& op,
                                                  const LsParams& ls, const double x,
                                                  const double s, const double l0,
                                                  const double m, int& fevals) {
  // step! (QN.jl:157-162): x_new = x_old + α s (one fma, rule C3), then logdensity.
  // Objectives with kMemoizeTrials (an evaluation is a GEMV against parameters in L2): when the
  // trial point is BIT FOR BIT the previous trial point, the previous value is the value and the
  // evaluation is skipped.  That is the fate of a search that is going to fail: after a step too
  // small to move x, y = 0 makes ρ = 1/sᵀy infinite (QN.jl:42 has no guard), the direction is NaN,
  // and all 1052 trials (QN.jl:178-193) evaluate the same all-NaN point.  Same results, same
  // evaluation count reported.
  long long xt_prev = 0x7ff8dead00000001LL; // no trial point has these bits
  double l_prev = 0.0;
  auto trial = [&](const double a) {
    ++fevals;
    const double xt = fma(a, s, x);
    if (Obj::kMemoizeTrials) {
      const long long bits = __double_as_longlong(xt);
      if (__syncthreads_and(tm.c >= tm.nvec || bits == xt_prev)) return l_prev;
      xt_prev = bits;
    }
    double dummy;
    if (tm.c < tm.nvec) tm.xs[tm.c] = xt;
    __syncthreads();
    l_prev = Obj::template eval<false>(tm, op, dummy);
    return l_prev;
  };
  double lx_last;
  return backtracking_linesearch(ls, l0, m, trial, lx_last);
}

__device__ __forceinline__ void write_problem_scalars(const BatchParams& p, long long j, int c,
                                                      double result, int iters, int fevals,
                                                      int status) {
  if (c == 0) {
    if (p.ell_out) p.ell_out[j] = result;
    if (p.iters_out) p.iters_out[j] = iters;
    if (p.fevals_out) p.fevals_out[j] = fevals;
    if (p.status_out) p.status_out[j] = status;
  }
}

// Two fp64 values of the calling thread's own tensor-memory lane (tcgen05.ld/st .32x32b.x4: four
// 32-bit columns); see qn_resident_tile.cuh for what was measured about this home.
__device__ __forceinline__ void tm2_load(double (&v)[2], const uint32_t addr) {
  asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b;\n"
               "tcgen05.ld.sync.aligned.32x32b.x4.b32 {t0a, t0b, t1a, t1b}, [%2];\n"
               "mov.b64 %0, {t0a, t0b};\nmov.b64 %1, {t1a, t1b};\n}\n"
               : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void tm2_store(const uint32_t addr, const double a, const double b) {
  asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b;\n"
               "mov.b64 {t0a, t0b}, %1;\nmov.b64 {t1a, t1b}, %2;\n"
               "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {t0a, t0b, t1a, t1b};\n}\n"
               :: "r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void tm2_wait_load() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tm2_wait_store() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// PAIRN > 0 (with NREG == 0; PAIRN = n at compile time, or 1 for n at run time): the register-tile policy — every thread's PROWS × rb tile of B⁻¹
// (PROWS rows, one rule-C2 block of rb <= 16 columns) lives in REGISTERS (or, PREG < PROWS, partly
// in shared memory); needs blockDim >= (B/PROWS) · #blocks so that a thread owns exactly one block.
// ONEPASS (shared-memory and register-pair policies only): the deferred-update order of
// streaming path — ONE sweep applies the previous
// iteration's rank-2 update and accumulates both GEMVs, t = B⁻¹y and w = B⁻¹∇; 16 n² instead of
// 24 n² bytes of B⁻¹ traffic per iteration.  Not the reference's rounding (QN_FLAG_ONE_PASS).
// PREG (register-tile policy): how many of a thread's PROWS rows live in registers; the others
// (0 or 2) live in shared memory, which halves the registers of a CTA so that TWO problems are
// resident per SM at n = 128.
// PTM (hybrid tiles of a full-size CTA only: every thread owns a complete tile, so the .sync.aligned
// tensor-memory instructions are reached by whole warps): the rows that are not in registers live
// in the thread's tensor-memory lane (`tmem`: lane and first column of this warp) instead of
// shared memory; column k of the two rows at columns 4k … 4k+3, the next column in flight while
// this one is used.
template <class Obj, int NREG, int PAIRN = 0, bool ONEPASS = false, int PROWS = 2, int PREG = PROWS, bool PTM = false>
__device__ __forceinline__ void resident_solve_one(const BatchParams& p,
                                                   const typename Obj::Params& op,
                                                   const long long j, Team& tm, double* ys,
                                                   Pass2Col* cols, double* hs, double* part,
                                                   double* part2 = nullptr, double* rowp = nullptr,
                                                   const uint32_t tmem = 0) {
  const int n = (NREG > 0) ? NREG : (PAIRN > 1 ? PAIRN : p.n); // PAIRN = 1: tiles with n at run time
  const int c = tm.c;          // thread index: the owner of element c when c < n
  const bool active = c < n;
  // shared-memory policy with more threads than elements ("pair" policy): thread c works on the
  // two rows 2·(c mod B/2), +1 and on the rule-C2 blocks bgrp, bgrp + G, ... of them
  const int B = resident_block_threads(n);
  const bool multi = (NREG == 0) && ((int)blockDim.x > B || PAIRN > 0);
  constexpr int PR = PROWS; // rows of B⁻¹ per thread (shared-memory and register-tile policies): 2 or 4
  static_assert(PR == 2 || PR == 4, "rows per thread");
  const int half = B / PR;                    // threads per column-block group
  const int G = (NREG > 0) ? 1 : (int)blockDim.x / half;
  const int row0 = (NREG > 0) ? c : PR * (c % half);
  const int bgrp = (NREG > 0) ? 0 : c / half;
  const int ldh = (NREG > 0) ? 0 : resident_ldh(n, PR);
  (void)part; (void)multi; (void)G; (void)row0; (void)bgrp; (void)ldh;
  constexpr int HR = (NREG > 0) ? NREG : 1;
  const int rb = chain_block(n); // rule C2 block length
  double h[HR]; // row c of B⁻¹ (register policy)
  (void)h; (void)hs;
  constexpr int PRB = (PAIRN > 0) ? 16 : 1; // a rule-C2 block has at most 16 columns
  const int pcol0 = bgrp * rb;                        // first column of this thread's block
  const int pcols = (PAIRN > 0) ? min(rb, n - pcol0) : 0;  // ... and how many it has (<= 0: none)
  (void)pcol0; (void)pcols;
  constexpr int PX = (PAIRN > 0) ? PR - PREG : 0; // rows of the tile kept in shared memory
  static_assert(PX == 0 || PX == 2, "hybrid tiles keep 0 or 2 rows in shared memory");
  double hp[PREG > 0 ? PREG : 1][PRB]; // rows row0 … row0+PREG-1 × the columns of block bgrp (register-tile policy)
  (void)hp;
  double* hx = hs + PX * (c % half);  // this thread's shared-memory rows of column 0; column stride ldx
  const int ldx = half * PX;
  (void)hx; (void)ldx;
  static_assert(!PTM || (PX == 2 && !ONEPASS), "tensor-memory rows: hybrid tiles in the reference order");
  double tmv[2][2]; // PTM: the tensor-memory rows of this column and of the next one
  (void)tmv;
  auto tile_begin = [&]() { // before a sweep over the columns k = 0, 1, …: column 0 on its way
    if constexpr (PTM) tm2_load(tmv[0], tmem);
  };
  auto load_tile_col = [&](const int k, double (&hh)[PR]) {
#pragma unroll
    for (int i = 0; i < PREG; ++i) hh[i] = hp[i][k];
    if (PX) {
      if constexpr (PTM) {
        tm2_wait_load();
        if (k + 1 < PRB) tm2_load(tmv[(k + 1) & 1], tmem + 4 * (k + 1));
        hh[PR - 2] = tmv[k & 1][0]; hh[PR - 1] = tmv[k & 1][1];
      } else {
        const double2 v = *reinterpret_cast<const double2*>(hx + ldx * (pcol0 + k));
        hh[PR - 2] = v.x; hh[PR - 1] = v.y;
      }
    }
  };
  auto store_tile_col = [&](const int k, const double (&hh)[PR]) {
#pragma unroll
    for (int i = 0; i < PREG; ++i) hp[i][k] = hh[i];
    if (PX) {
      if constexpr (PTM) tm2_store(tmem + 4 * k, hh[PR - 2], hh[PR - 1]);
      else *reinterpret_cast<double2*>(hx + ldx * (pcol0 + k)) = make_double2(hh[PR - 2], hh[PR - 1]);
    }
  };
  auto tile_end = [&]() { // after a sweep that stored
    if constexpr (PTM) tm2_wait_store();
  };

  double x = active ? p.theta0[j * p.theta_stride + c] : 0.0; // copyto!(x_old, x), QN.jl:248
  double g = 0.0, g_prev = 0.0, s = 0.0;
  double ps = 0.0, pu = 0.0, pc1 = 0.0; // one-pass mode: the pending update (s, u of this element; c₁)
  bool pending = false;
  (void)ps; (void)pu; (void)pc1; (void)pending; (void)part2; (void)rowp;
  // one sweep over this thread's part of B⁻¹ (pair policies): applies the pending update when
  // APPLY, accumulates t and w when GEMV, block partials to part / part2
  auto onepass_sweep = [&](const bool apply, const bool gemv) {
    if (!(bgrp < G && row0 < n)) return;
    if (PAIRN > 0 && pcols <= 0) return;
    if (PAIRN > 0) {
      double sR[PR], nR[PR], tt[PR], ww[PR];
#pragma unroll
      for (int i = 0; i < PR; ++i) { sR[i] = rowp[row0 + i]; nR[i] = rowp[B + row0 + i]; tt[i] = 0.0; ww[i] = 0.0; }
#pragma unroll
      for (int k = 0; k < PRB; ++k) {
        if (k < pcols) {
          const double2 wn = *reinterpret_cast<const double2*>(&cols[pcol0 + k].w);
          const double2 sg = *reinterpret_cast<const double2*>(&cols[pcol0 + k].s);
          const double yr = ys[pcol0 + k];
          double hh[PR];
          load_tile_col(k, hh);
#pragma unroll
          for (int i = 0; i < PR; ++i) {
            if (apply) {
              hh[i] = fma(wn.x, sR[i], hh[i]);
              hh[i] = fma(wn.y, sR[i], hh[i]);
              hh[i] = fma(nR[i], sg.x, hh[i]);
            }
            if (gemv) {
              tt[i] = fma(yr, hh[i], tt[i]);
              ww[i] = fma(hh[i], sg.y, ww[i]);
            }
          }
          if (apply) store_tile_col(k, hh);
        }
      }
      if (gemv) {
#pragma unroll
        for (int i = 0; i < PR; i += 2) {
          *reinterpret_cast<double2*>(part + bgrp * B + row0 + i) = make_double2(tt[i], tt[i + 1]);
          *reinterpret_cast<double2*>(part2 + bgrp * B + row0 + i) = make_double2(ww[i], ww[i + 1]);
        }
      }
    } else {
      double* hrow = hs + row0;
      double sR[PR], nR[PR];
#pragma unroll
      for (int i = 0; i < PR; ++i) { sR[i] = rowp[row0 + i]; nR[i] = rowp[B + row0 + i]; }
      for (int r0 = bgrp * rb, blk = bgrp; r0 < n; r0 += G * rb, blk += G) {
        const int r1 = min(r0 + rb, n);
        double tt[PR], ww[PR];
#pragma unroll
        for (int i = 0; i < PR; ++i) { tt[i] = 0.0; ww[i] = 0.0; }
#pragma unroll 4
        for (int r = r0; r < r1; ++r) {
          const double2 wn = *reinterpret_cast<const double2*>(&cols[r].w);
          const double2 sg = *reinterpret_cast<const double2*>(&cols[r].s);
          const double yr = ys[r];
          double hh[PR];
#pragma unroll
          for (int i = 0; i < PR; i += 2) {
            const double2 h2 = *reinterpret_cast<const double2*>(hrow + ldh * r + i);
            hh[i] = h2.x; hh[i + 1] = h2.y;
          }
#pragma unroll
          for (int i = 0; i < PR; ++i) {
            if (apply) {
              hh[i] = fma(wn.x, sR[i], hh[i]);
              hh[i] = fma(wn.y, sR[i], hh[i]);
              hh[i] = fma(nR[i], sg.x, hh[i]);
            }
            if (gemv) {
              tt[i] = fma(yr, hh[i], tt[i]);
              ww[i] = fma(hh[i], sg.y, ww[i]);
            }
          }
          if (apply) {
#pragma unroll
            for (int i = 0; i < PR; i += 2)
              *reinterpret_cast<double2*>(hrow + ldh * r + i) = make_double2(hh[i], hh[i + 1]);
          }
        }
        if (gemv) {
#pragma unroll
          for (int i = 0; i < PR; i += 2) {
            *reinterpret_cast<double2*>(part + blk * B + row0 + i) = make_double2(tt[i], tt[i + 1]);
            *reinterpret_cast<double2*>(part2 + blk * B + row0 + i) = make_double2(ww[i], ww[i + 1]);
          }
        }
      }
    }
  };
  double result = __longlong_as_double(0x7ff8000000000000LL); // NaN, QN.jl:291
  int iters = 0, fevals = 0, status = QN_MAX_ITER;

  for (int it = 1; it <= p.max_iter; ++it) { // QN.jl:250,253
    // ---- ℓ₀ = ∂logdensity!(∇_new, obj, x_old)   QN.jl:254
    if (c < B) tm.xs[c] = x;
    __syncthreads();
    const double l0 = Obj::template eval<true>(tm, op, g);
    ++fevals;
    if (!isfinite(l0)) { status = QN_NONFINITE; if (!(it & 1)) g = g_prev; break; } // QN.jl:255

    // publish what the update needs, and test convergence, on one barrier.
    // y = ∇_old - ∇_new, QN.jl:38
    const double y = g_prev - g;
    if (c < B) {
      ys[c] = y;
      cols[c].s = s;
      cols[c].g = g;
    }
    tm.stage(s * y);
    // maximum(abs, ∇_new) < tol, QN.jl:257 (a NaN entry makes it false)
    const int conv = __syncthreads_and(!active || fabs(g) < p.tol);
    if (conv) { status = QN_CONVERGED; result = l0; break; } // QN.jl:258-261

    double m;
    if (it == 1) {
      tm.phase ^= 1; // staged value unused
      m = -1.0;      // QN.jl:263-264
    } else if (ONEPASS) {
      // ---- deferred-update BFGS step (qo_bfgs_update_onepass): same mathematics as QN.jl:34-69
      const double sy = tm.finish(n);  // sᵀy, QN.jl:39
      const double rho = 1.0 / sy;     // QN.jl:42 (no guard)
      if (c < B) { // column operands of the PENDING update and of the two GEMVs; y is already in ys
        cols[c].w = pc1 * ps;   // (c₁ s[r]) of the pending update, rounded once per r (rule C4)
        cols[c].nu = -pu;
        cols[c].s = ps;
        cols[c].g = g;
        rowp[c] = ps;
        rowp[B + c] = -pu;
      }
      __syncthreads();
      onepass_sweep(pending, true);
      __syncthreads();
      double T = 0.0, W = 0.0;
      if (active) { // block partials in ascending order (rule C2)
        T = part[c];
        W = part2[c];
        for (int r0 = rb, blk = 1; r0 < n; r0 += rb, ++blk) { T = T + part[blk * B + c]; W = W + part2[blk * B + c]; }
      }
      const double u = T * rho;                       // B⁻¹y[c] = t ρ, QN.jl:50
      double yHy, sg, ug;                             // yᵀB⁻¹y (QN.jl:51), sᵀ∇, uᵀ∇ on one barrier
      tm.csum3(T * y, s * g, u * g, n, yHy, sg, ug);
      const double c1 = fma(yHy, rho, 1.0) * rho;     // QN.jl:54
      // d = H⁺∇ = ((w + (c₁ s)(sᵀ∇)) - u (sᵀ∇)) - s (uᵀ∇); this update becomes the pending one
      double d = fma(c1 * s, sg, W);
      d = fma(-u, sg, d);
      d = fma(-s, ug, d);
      ps = s; pu = u; pc1 = c1; pending = true;
      m = tm.csum(d * g, n); // m = dᵀ∇_new, QN.jl:66-68
      s = d;                 // the new direction; `s, y = y, s`, QN.jl:65,268
    } else {
      // ---- BFGS_update!  QN.jl:34-69
      const double sy = tm.finish(n);  // sᵀy, QN.jl:39
      const double rho = 1.0 / sy;     // QN.jl:42 (no guard)
      // pass 1: T = Σ_r y[r] H[c,r]   QN.jl:44-49: block chains, partials added in order (rule C2)
      double T = 0.0;
      if (NREG > 0) {
        constexpr int RB = chain_block(HR);
#pragma unroll
        for (int r0 = 0; r0 < HR; r0 += RB) {
          double t = 0.0;
#pragma unroll
          for (int r = r0; r < r0 + RB && r < HR; ++r) t = fma(ys[r], h[r], t);
          T = (r0 == 0) ? t : T + t;
        }
      } else if (PAIRN > 0) {
        if (row0 < n && pcols > 0) { // (threads beyond the last column block: the rounding of the CTA to whole warps)
          double tt[PR];
#pragma unroll
          for (int i = 0; i < PR; ++i) tt[i] = 0.0;
          tile_begin();
#pragma unroll
          for (int k = 0; k < PRB; ++k) {
            if (k < pcols) {
              const double yr = ys[pcol0 + k];
              double hh[PR];
              load_tile_col(k, hh);
#pragma unroll
              for (int i = 0; i < PR; ++i) tt[i] = fma(yr, hh[i], tt[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < PR; i += 2)
            *reinterpret_cast<double2*>(part + bgrp * B + row0 + i) = make_double2(tt[i], tt[i + 1]);
        }
        __syncthreads();
        if (active) { // block partials in ascending order (rule C2)
          T = part[c];
          for (int r0 = rb, blk = 1; r0 < n; r0 += rb, ++blk) T = T + part[blk * B + c];
        }
      } else if (!multi) {
        if (active) {
          const double* hrow = hs + c;
          for (int r0 = 0; r0 < n; r0 += rb) {
            const int r1 = min(r0 + rb, n);
            double t = 0.0;
#pragma unroll 4
            for (int r = r0; r < r1; ++r) t = fma(ys[r], hrow[ldh * r], t);
            T = (r0 == 0) ? t : T + t;
          }
        }
      } else {
        if (bgrp < G && row0 < n) {
          const double* hrow = hs + row0;
          for (int r0 = bgrp * rb, blk = bgrp; r0 < n; r0 += G * rb, blk += G) {
            const int r1 = min(r0 + rb, n);
            double tt[PR];
#pragma unroll
            for (int i = 0; i < PR; ++i) tt[i] = 0.0;
#pragma unroll 8
            for (int r = r0; r < r1; ++r) {
              const double yr = ys[r];
#pragma unroll
              for (int i = 0; i < PR; i += 2) {
                const double2 hh = *reinterpret_cast<const double2*>(hrow + ldh * r + i);
                tt[i] = fma(yr, hh.x, tt[i]);
                tt[i + 1] = fma(yr, hh.y, tt[i + 1]);
              }
            }
#pragma unroll
            for (int i = 0; i < PR; i += 2)
              *reinterpret_cast<double2*>(part + blk * B + row0 + i) = make_double2(tt[i], tt[i + 1]);
          }
        }
        __syncthreads();
        if (active) { // block partials in ascending order (rule C2)
          T = part[c];
          for (int r0 = rb, blk = 1; r0 < n; r0 += rb, ++blk) T = T + part[blk * B + c];
        }
      }
      const double u = T * rho;       // B⁻¹y[c] = t ρ, QN.jl:50
      if (c < B) cols[c].nu = -u;
      const double yHy = tm.csum(T * y, n);          // QN.jl:51
      const double c1 = fma(yHy, rho, 1.0) * rho;    // QN.jl:54
      if (c < B) {
        cols[c].w = c1 * s;                          // (c₁ s[r]), rounded once per r (rule C4)
        if (multi) { ys[c] = s; tm.xs[c] = -u; }     // row operands of a pair as one 16-byte read
      }                                              //   (y and the point are dead until the line search)
      __syncthreads();
      // pass 2: H ← H + c₁ssᵀ - usᵀ - suᵀ fused with d = H⁺∇_new   QN.jl:56-67
      const double nu_c = -u;
      double D = 0.0;
      if (NREG > 0) {
        constexpr int RB = chain_block(HR);
#pragma unroll
        for (int r0 = 0; r0 < HR; r0 += RB) {
          double acc = 0.0;
#pragma unroll
          for (int r = r0; r < r0 + RB && r < HR; ++r) {
            const double2 wn = *reinterpret_cast<const double2*>(&cols[r].w);
            const double2 sg = *reinterpret_cast<const double2*>(&cols[r].s);
            double hh = h[r];
            hh = fma(wn.x, s, hh);   // + (c₁ s[r]) s[c]
            hh = fma(wn.y, s, hh);   // - u[r] s[c]
            hh = fma(nu_c, sg.x, hh);// - u[c] s[r]
            h[r] = hh;
            acc = fma(hh, sg.y, acc);
          }
          D = (r0 == 0) ? acc : D + acc;
        }
      } else if (PAIRN > 0) {
        if (row0 < n && pcols > 0) { // (threads beyond the last column block: the rounding of the CTA to whole warps)
          double sR[PR], nR[PR], acc[PR]; // s and -u of this thread's rows (see below), chain accumulators
#pragma unroll
          for (int i = 0; i < PR; i += 2) {
            const double2 s2 = *reinterpret_cast<const double2*>(ys + row0 + i);
            const double2 n2 = *reinterpret_cast<const double2*>(tm.xs + row0 + i);
            sR[i] = s2.x; sR[i + 1] = s2.y; nR[i] = n2.x; nR[i + 1] = n2.y;
            acc[i] = 0.0; acc[i + 1] = 0.0;
          }
          tile_begin();
#pragma unroll
          for (int k = 0; k < PRB; ++k) {
            if (k < pcols) {
              const double2 wn = *reinterpret_cast<const double2*>(&cols[pcol0 + k].w);
              const double2 sg = *reinterpret_cast<const double2*>(&cols[pcol0 + k].s);
              double hh[PR];
              load_tile_col(k, hh);
#pragma unroll
              for (int i = 0; i < PR; ++i) {
                hh[i] = fma(wn.x, sR[i], hh[i]);   // + (c₁ s[r]) s[c]
                hh[i] = fma(wn.y, sR[i], hh[i]);   // - u[r] s[c]
                hh[i] = fma(nR[i], sg.x, hh[i]);   // - u[c] s[r]
                acc[i] = fma(hh[i], sg.y, acc[i]);
              }
              store_tile_col(k, hh);
            }
          }
          tile_end();
#pragma unroll
          for (int i = 0; i < PR; i += 2)
            *reinterpret_cast<double2*>(part + bgrp * B + row0 + i) = make_double2(acc[i], acc[i + 1]);
        }
        __syncthreads();
        if (active) {
          D = part[c];
          for (int r0 = rb, blk = 1; r0 < n; r0 += rb, ++blk) D = D + part[blk * B + c];
        }
      } else if (!multi) {
        if (active) {
          double* hrow = hs + c;
          for (int r0 = 0; r0 < n; r0 += rb) {
            const int r1 = min(r0 + rb, n);
            double acc = 0.0;
#pragma unroll 4
            for (int r = r0; r < r1; ++r) {
              const double2 wn = *reinterpret_cast<const double2*>(&cols[r].w);
              const double2 sg = *reinterpret_cast<const double2*>(&cols[r].s);
              double hh = hrow[ldh * r];
              hh = fma(wn.x, s, hh);
              hh = fma(wn.y, s, hh);
              hh = fma(nu_c, sg.x, hh);
              hrow[ldh * r] = hh;
              acc = fma(hh, sg.y, acc);
            }
            D = (r0 == 0) ? acc : D + acc;
          }
        }
      } else {
        if (bgrp < G && row0 < n) {
          // row operands of this thread's rows, published by the rows' owners (zero beyond n)
          double sR[PR], nR[PR];
#pragma unroll
          for (int i = 0; i < PR; i += 2) {
            const double2 s2 = *reinterpret_cast<const double2*>(ys + row0 + i);
            const double2 n2 = *reinterpret_cast<const double2*>(tm.xs + row0 + i);
            sR[i] = s2.x; sR[i + 1] = s2.y; nR[i] = n2.x; nR[i + 1] = n2.y;
          }
          double* hrow = hs + row0;
          for (int r0 = bgrp * rb, blk = bgrp; r0 < n; r0 += G * rb, blk += G) {
            const int r1 = min(r0 + rb, n);
            double acc[PR];
#pragma unroll
            for (int i = 0; i < PR; ++i) acc[i] = 0.0;
#pragma unroll 4
            for (int r = r0; r < r1; ++r) {
              const double2 wn = *reinterpret_cast<const double2*>(&cols[r].w);
              const double2 sg = *reinterpret_cast<const double2*>(&cols[r].s);
#pragma unroll
              for (int i = 0; i < PR; i += 2) {
                double2 hh = *reinterpret_cast<const double2*>(hrow + ldh * r + i);
                hh.x = fma(wn.x, sR[i], hh.x);  hh.y = fma(wn.x, sR[i + 1], hh.y);   // + (c₁ s[r]) s[c]
                hh.x = fma(wn.y, sR[i], hh.x);  hh.y = fma(wn.y, sR[i + 1], hh.y);   // - u[r] s[c]
                hh.x = fma(nR[i], sg.x, hh.x);  hh.y = fma(nR[i + 1], sg.x, hh.y);   // - u[c] s[r]
                *reinterpret_cast<double2*>(hrow + ldh * r + i) = hh;
                acc[i] = fma(hh.x, sg.y, acc[i]);
                acc[i + 1] = fma(hh.y, sg.y, acc[i + 1]);
              }
            }
#pragma unroll
            for (int i = 0; i < PR; i += 2)
              *reinterpret_cast<double2*>(part + blk * B + row0 + i) = make_double2(acc[i], acc[i + 1]);
          }
        }
        __syncthreads();
        if (active) {
          D = part[c];
          for (int r0 = rb, blk = 1; r0 < n; r0 += rb, ++blk) D = D + part[blk * B + c];
        }
      }
      m = tm.csum(D * g, n); // m = dᵀ∇_new, QN.jl:66-68
      s = D;                 // the new direction; `s, y = y, s`, QN.jl:65,268
    }
    if (m <= 0.0) { // QN.jl:272-280: H ← I, s ← ∇_new, m ← ∇ᵀ∇
      if (NREG > 0) {
#pragma unroll
        for (int r = 0; r < HR; ++r) h[r] = (r == c) ? 1.0 : 0.0;
      } else if (PAIRN > 0) {
#pragma unroll
        for (int k = 0; k < PRB; ++k) {
          if (k < pcols || PX == 0) {
            double hh[PR];
#pragma unroll
            for (int i = 0; i < PR; ++i) hh[i] = (row0 + i == pcol0 + k) ? 1.0 : 0.0;
            store_tile_col(k, hh);
          }
        }
        tile_end();
      } else if (active) {
        for (int r = 0; r < n; ++r) hs[c + ldh * r] = (r == c) ? 1.0 : 0.0;
      }
      s = g;
      m = tm.csum(g * g, n);
      pending = false; // one-pass mode: the reset discards the pending update too
    }

    const double alpha = team_linesearch<Obj>(tm, op, p.ls, x, s, l0, m, fevals); // QN.jl:283
    if (alpha == 0.0) { status = QN_LINESEARCH_FAILED; if (!(it & 1)) g = g_prev; break; } // QN.jl:284
    // update_state!: s ← αs ; x_old += s   QN.jl:21-27
    s = alpha * s;
    x = x + s;
    ++iters;
    if (p.trace != nullptr && it <= p.ntrace && active)
      p.trace[(j * p.ntrace + (it - 1)) * (long long)n + c] = x;
    if (it == p.max_iter) { if (!(it & 1)) g = g_prev; break; } // N iterations exhausted, QN.jl:250,291
    g_prev = g; // ∇_old, ∇_new = ∇_new, ∇_old, QN.jl:286
  }

  // ---- results: θ* at offset 0 and ∇ at offset L of the state block (QN.jl:135,138).
  // The reference swaps its two gradient buffers every iteration (QN.jl:286) and only
  // copies into the canonical slot on convergence (QN.jl:258-259): after a failure in an
  // even iteration `gradient(state)` still shows the previous iteration's ∇ (selected above).
  if (ONEPASS && p.h_out && pending) { // the stored matrix lags one update behind: apply it (qo_flush_pending)
    __syncthreads();
    if (c < B) {
      cols[c].w = pc1 * ps; cols[c].nu = -pu; cols[c].s = ps; cols[c].g = 0.0;
      rowp[c] = ps; rowp[B + c] = -pu;
    }
    __syncthreads();
    onepass_sweep(true, false);
    __syncthreads();
  }
  if (PAIRN > 0 && p.h_out && row0 < n && pcols > 0) {
    double* ho = p.h_out + j * p.h_stride;
    tile_begin();
#pragma unroll
    for (int k = 0; k < PRB; ++k) {
      if (k < pcols) {
        double hh[PR];
        load_tile_col(k, hh);
#pragma unroll
        for (int i = 0; i < PR; ++i)
          if (row0 + i < n) ho[row0 + i + (long long)p.ldh * (pcol0 + k)] = hh[i];
      }
    }
  }
  if (active) {
    if (p.x_out) p.x_out[j * p.x_stride + c] = x;
    if (p.g_out) p.g_out[j * p.g_stride + c] = g;
    if (p.h_out && PAIRN == 0) {
      double* ho = p.h_out + j * p.h_stride + c;
      if (NREG > 0) {
#pragma unroll
        for (int r = 0; r < HR; ++r) ho[(long long)p.ldh * r] = h[r];
      } else {
        for (int r = 0; r < n; ++r) ho[(long long)p.ldh * r] = hs[c + ldh * r];
      }
    }
  }
  write_problem_scalars(p, j, c, result, iters, fevals, status);
}

template <class Obj, int NREG, int BLOCK, int MINB, int PAIRN = 0, bool ONEPASS = false, int PROWS = 2,
          int PREG = PROWS, bool PTM = false>
__global__ void __launch_bounds__(BLOCK, MINB)
qn_resident_kernel(const BatchParams p, const typename Obj::Params op) {
  // launch: blockDim >= resident_block_threads(n); see launch_resident for the policy
  extern __shared__ __align__(16) double smem[];
  const int n = (NREG > 0) ? NREG : (PAIRN > 1 ? PAIRN : p.n); // PAIRN = 1: tiles with n at run time
  const int B = resident_block_threads(n);
  const ResidentSmem L = resident_smem_layout<Obj>(n, NREG == 0 && PAIRN == 0, (int)blockDim.x, op, ONEPASS, PROWS,
                                                   (PAIRN > 0 && !PTM) ? PROWS - PREG : 0);
  // PTM: 4 columns per tile column and thread; warps 4-7 of the CTA sit on the lanes of warps 0-3,
  // one 64-column band further (n = 128: 256 threads, 16 tile columns, 128 columns per CTA)
  uint32_t tmem = 0, tmem_base = 0;
  if constexpr (PTM) {
    static_assert(BLOCK <= 256, "tensor-memory rows: at most two bands of four warps");
    __shared__ uint32_t tmem_base_s;
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n"
                   :: "l"((unsigned long long)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    tmem_base = tmem_base_s;
    const uint32_t w = threadIdx.x >> 5;
    tmem = __reduce_max_sync(0xffffffffu, tmem_base + (((w & 3) * 32u) << 16) + (w >> 2) * 64u);
  }

  Team tm;
  tm.n = n;
  tm.c = threadIdx.x;
  tm.lane = threadIdx.x & 31;
  tm.red = smem + L.red;
  tm.stride = B;
  tm.phase = 0;
  tm.xs = smem + L.xs;
  tm.nvec = B;
  tm.tpr = (NREG > 0) ? 1 : (int)blockDim.x / B;
  tm.part = smem + L.part;
  tm.scratch = smem + L.scratch;
  double* ys = smem + L.ys;
  Pass2Col* cols = reinterpret_cast<Pass2Col*>(smem + L.cols);
  double* hs = smem + L.h;
  volatile long long* next = reinterpret_cast<volatile long long*>(smem + L.total - 2);
  Obj::prologue(tm, op); // once per CTA: e.g. the dense Gaussian's Λ into shared memory
  if (NREG == 0 && PAIRN == 0) { // the padding row of an odd-n B⁻¹ is read and written by its pair: keep it finite
    const int ldh = resident_ldh(n, PROWS);
    for (int pad = n; pad < ldh; ++pad)
      for (int r = threadIdx.x; r < n; r += blockDim.x) hs[pad + ldh * r] = 0.0;
  }

  for (;;) {
    __syncthreads(); // previous problem's last reads of shared memory are done
    if (threadIdx.x == 0) *next = (long long)atomicAdd(p.counter, 1ULL);
    __syncthreads();
    const long long j = *next;
    if (j >= p.batch) break;
    resident_solve_one<Obj, NREG, PAIRN, ONEPASS, PROWS, PREG, PTM>(p, op, j, tm, ys, cols, hs, smem + L.part,
                                                                    smem + L.part2, smem + L.rowp, tmem);
  }
  if constexpr (PTM) {
    __syncthreads(); // every warp is done with the tensor memory
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace qn
