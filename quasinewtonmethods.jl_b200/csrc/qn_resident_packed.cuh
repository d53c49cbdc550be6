// qn_resident_packed.cuh — small problems (n = 8, 16, 32): SEVERAL problems per warp.
//
// The thread-per-row kernel of qn_resident.cuh gives a warp to one problem, so at n = 8 only
// 8 of 32 lanes work and every shared-memory broadcast of the two O(n²) passes serves one
// problem; those kernels are bound by the shared-memory pipe (one broadcast per DFMA).  Here a
// warp is cut into 32/n SUB-GROUPS of n lanes, one problem each (BASELINE.json configs[3]:
// "n <= 32 should pack several problems per CTA").  A quarter-warp-addressed
// LDS.128 costs the same as a fully broadcast one (profiles/ubench_r1.txt), so the column
// operands of all sub-groups arrive with the same instruction, and all lanes do useful DFMAs.
//
// Sub-groups never synchronise with each other: every barrier is __syncwarp(mask) and every
// shuffle / vote carries the sub-group's lane mask, so their control flow (iteration counts,
// line-search trip counts, the work queue) is independent.  A sub-group that finishes early
// waits at the warp's reconvergence point and then pulls its next problem.
//
// Arithmetic: exactly resident_solve_one<Rosenbrock, NREG> — the same operations on the same
// operands in the same order, hence the same bits as the oracle.  The rule-C1 butterfly over
// 32 lane accumulators, of which only the first n are non-zero, is reproduced inside n lanes:
// the levels that would add a +0.0 partner are the single addition p + 0.0 (it only turns -0.0
// into +0.0, as the full butterfly does), the remaining levels are shuffles inside the sub-group.
#pragma once

#include "qn_resident.cuh"

namespace qn {

// N = problem dimension (even, <= 32), RPT = rows (= elements) per lane; a sub-group has
// SG = N / RPT lanes and a warp 32 / SG problems.  Lane l of a sub-group owns elements and rows
// e_k = l + SG·k, k < RPT — so a Rosenbrock pair (i, i + N/2) lies in one lane whenever RPT is
// even, and every broadcast column operand is used for RPT rows.
template <int N, int RPT>
struct PackedCfg {
  static constexpr int SG = N / RPT;
  static_assert(N % RPT == 0 && (SG == 2 || SG == 4 || SG == 8 || SG == 16 || SG == 32), "sub-group of 2 … 32 lanes");
  static_assert(N % 2 == 0 && N <= 32, "even n <= 32");
  static constexpr int RB = chain_block(N);      // rule-C2 block: N for N <= 16, 16 for N = 32
  static_assert(N % RB == 0 && RB % 2 == 0, "whole, even rule-C2 blocks");
  static constexpr int kSub = 32 / SG;           // problems per warp
  static constexpr int kWarps = 4;               // warps per CTA (never synchronised)
  // doubles per sub-group: Pass2Col[N] + ys[N], padded so that consecutive sub-groups start 4
  // banks apart: the LDS.128 that broadcasts column r to every sub-group of the warp then touches
  // each bank once (unpadded, n = 8: 8 sub-groups on 2 bank groups, 4-way conflicts; ncu showed
  // 44 % of the kernel's shared-memory wavefronts were conflicts)
  static constexpr int kPerSub = 5 * N + ((2 - (5 * N) % 16) + 16) % 16;
  static constexpr int kSmemDoubles = kWarps * kSub * kPerSub;
};

// rule C1 for the sub-group's N <= 32 lane accumulators: v[k] is the accumulator of element
// l + SG·k (+0.0 where the count is smaller).  Butterfly levels 16 … SG pair accumulators that
// live in the same lane (or a +0.0 beyond N); levels SG/2 … 1 are shuffles inside the sub-group.
template <int N, int RPT>
__device__ __forceinline__ double sub_csum(const unsigned mask, const double (&v)[RPT]) {
  constexpr int SG = N / RPT, K = 32 / SG;
  double a[K];
#pragma unroll
  for (int k = 0; k < K; ++k) a[k] = (k < RPT) ? v[k] : 0.0;
#pragma unroll
  for (int half = K / 2; half >= 1; half >>= 1)
#pragma unroll
    for (int k = 0; k < half; ++k) a[k] = a[k] + a[k + half];
  double p = a[0];
#pragma unroll
  for (int off = SG / 2; off >= 1; off >>= 1) p = p + __shfl_xor_sync(mask, p, off);
  return p;
}

// Rosenbrock (test/runtests.jl:4-33) on a sub-group.  Element e = l + SG·k pairs with e ± N/2:
// the same lane (k ± RPT/2) when RPT is even, the lane l ^ N/2 when RPT = 1.
// Returns ℓ in every lane; with GRAD leaves ∇ in g.  N is even: no tail term.
template <int N, int RPT, bool GRAD>
__device__ __forceinline__ double sub_rosenbrock(const unsigned mask, const int l, const double (&x)[RPT],
                                                 double (&g)[RPT]) {
  constexpr int N2 = N / 2;
  double term[RPT];
  if (RPT == 1) {
    const double xp = __shfl_xor_sync(mask, x[0], N2); // the other element of my pair
    const bool first = l < N2;
    const double t1 = first ? x[0] : xp, t2 = first ? xp : x[0];
    const double a = t2 - t1 * t1;
    const double b = 1.0 - t1;
    term[0] = first ? 100.0 * (a * a) + b * b : 0.0;            // runtests.jl:9,23
    if (GRAD) g[0] = first ? (400.0 * a) * t1 + 2.0 * b          // runtests.jl:24
                           : 200.0 * (t1 * t1 - t2);             // runtests.jl:25
  } else {
    constexpr int H = RPT / 2; // elements k < H are θ1's, k >= H their θ2's
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const double t1 = x[k], t2 = x[k + H];
      const double a = t2 - t1 * t1;
      const double b = 1.0 - t1;
      term[k] = 100.0 * (a * a) + b * b;
      term[k + H] = 0.0;
      if (GRAD) {
        g[k] = (400.0 * a) * t1 + 2.0 * b;
        g[k + H] = 200.0 * (t1 * t1 - t2);
      }
    }
  }
  return 0.0 - sub_csum<N, RPT>(mask, term); // s -= term, @simd-reassociated
}

template <int N, int RPT>
__device__ __forceinline__ void packed_solve_one(const BatchParams& p, const long long j, const unsigned mask,
                                                 const int l, double* ys, Pass2Col* cols) {
  constexpr int SG = N / RPT;
  double h[RPT][N]; // rows l + SG·k of B⁻¹
  double x[RPT], g[RPT], g_prev[RPT], s[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    x[k] = p.theta0[j * p.theta_stride + l + SG * k]; // copyto!(x_old, x), QN.jl:248
    g[k] = g_prev[k] = s[k] = 0.0;
  }
  double result = __longlong_as_double(0x7ff8000000000000LL); // NaN, QN.jl:291
  int iters = 0, fevals = 0, status = QN_MAX_ITER;
  const LsParams& ls = p.ls;
  auto restore_prev = [&]() {
#pragma unroll
    for (int k = 0; k < RPT; ++k) g[k] = g_prev[k];
  };

  for (int it = 1; it <= p.max_iter; ++it) { // QN.jl:250,253
    const double l0 = sub_rosenbrock<N, RPT, true>(mask, l, x, g); // ℓ₀ = ∂logdensity!(∇_new, obj, x_old), QN.jl:254
    ++fevals;
    if (!isfinite(l0)) { status = QN_NONFINITE; if (!(it & 1)) restore_prev(); break; } // QN.jl:255
    double y[RPT], v[RPT];
    __syncwarp(mask); // the previous iteration's reads of ys / cols are done
    bool small = true;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int e = l + SG * k;
      y[k] = g_prev[k] - g[k]; // y = ∇_old - ∇_new, QN.jl:38
      ys[e] = y[k];
      cols[e].s = s[k];
      cols[e].g = g[k];
      small = small && (fabs(g[k]) < p.tol);
    }
    // maximum(abs, ∇_new) < tol, QN.jl:257 (a NaN entry makes it false)
    if (__all_sync(mask, small)) { status = QN_CONVERGED; result = l0; break; } // QN.jl:258-261

    double m;
    if (it == 1) {
      m = -1.0; // QN.jl:263-264
    } else {
      // ---- BFGS_update!  QN.jl:34-69
#pragma unroll
      for (int k = 0; k < RPT; ++k) v[k] = s[k] * y[k];
      const double sy = sub_csum<N, RPT>(mask, v); // sᵀy, QN.jl:39
      const double rho = 1.0 / sy;                 // QN.jl:42 (no guard)
      __syncwarp(mask);
      // pass 1: T = Σ_r y[r] H[e,r]   QN.jl:44-49: one fma chain per rule-C2 block, partials added in order
      constexpr int RB = PackedCfg<N, RPT>::RB;
      double T[RPT], u[RPT];
#pragma unroll
      for (int r0 = 0; r0 < N; r0 += RB) {
        double t[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) t[k] = 0.0;
#pragma unroll
        for (int r = r0; r < r0 + RB; r += 2) {
          const double2 yy = *reinterpret_cast<const double2*>(ys + r);
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            t[k] = fma(yy.x, h[k][r], t[k]);
            t[k] = fma(yy.y, h[k][r + 1], t[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) T[k] = (r0 == 0) ? t[k] : T[k] + t[k];
      }
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        u[k] = T[k] * rho; // B⁻¹y[e] = t ρ, QN.jl:50
        cols[l + SG * k].nu = -u[k];
        v[k] = T[k] * y[k];
      }
      const double yHy = sub_csum<N, RPT>(mask, v); // QN.jl:51
      const double c1 = fma(yHy, rho, 1.0) * rho;   // QN.jl:54
#pragma unroll
      for (int k = 0; k < RPT; ++k) cols[l + SG * k].w = c1 * s[k]; // (c₁ s[r]), rounded once per r (rule C4)
      __syncwarp(mask);
      // pass 2: H ← H + c₁ssᵀ - usᵀ - suᵀ fused with d = H⁺∇_new   QN.jl:56-67
      double acc[RPT];
#pragma unroll
      for (int r0 = 0; r0 < N; r0 += RB) {
        double a[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) a[k] = 0.0;
#pragma unroll
        for (int r = r0; r < r0 + RB; ++r) {
          const double2 wn = *reinterpret_cast<const double2*>(&cols[r].w);
          const double2 sg = *reinterpret_cast<const double2*>(&cols[r].s);
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            double hh = h[k][r];
            hh = fma(wn.x, s[k], hh);   // + (c₁ s[r]) s[c]
            hh = fma(wn.y, s[k], hh);   // - u[r] s[c]
            hh = fma(-u[k], sg.x, hh);  // - u[c] s[r]
            h[k][r] = hh;
            a[k] = fma(hh, sg.y, a[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) acc[k] = (r0 == 0) ? a[k] : acc[k] + a[k];
      }
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        v[k] = acc[k] * g[k];
        s[k] = acc[k]; // the new direction; `s, y = y, s`, QN.jl:65,268
      }
      m = sub_csum<N, RPT>(mask, v); // m = dᵀ∇_new, QN.jl:66-68
    }
    if (m <= 0.0) { // QN.jl:272-280: H ← I, s ← ∇_new, m ← ∇ᵀ∇
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
#pragma unroll
        for (int r = 0; r < N; ++r) h[k][r] = (r == l + SG * k) ? 1.0 : 0.0;
        s[k] = g[k];
        v[k] = g[k] * g[k];
      }
      m = sub_csum<N, RPT>(mask, v);
    }

    // ---- linesearch!  QN.jl:165-232 (every lane of the sub-group runs the same scalar code)
    double alpha;
    {
      double dummy[RPT];
      auto trial = [&](const double a) { // step! (QN.jl:157-162): x + a s as one fma, then ℓ
        ++fevals;
        double xt[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) xt[k] = fma(a, s[k], x[k]);
        return sub_rosenbrock<N, RPT, false>(mask, l, xt, dummy);
      };
      double lx_last;
      alpha = backtracking_linesearch(ls, l0, m, trial, lx_last);
    }
    if (alpha == 0.0) { status = QN_LINESEARCH_FAILED; if (!(it & 1)) restore_prev(); break; } // QN.jl:284
    // update_state!: s ← αs ; x_old += s   QN.jl:21-27
    ++iters;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      s[k] = alpha * s[k];
      x[k] = x[k] + s[k];
      if (p.trace != nullptr && it <= p.ntrace)
        p.trace[(j * p.ntrace + (it - 1)) * (long long)N + l + SG * k] = x[k];
    }
    if (it == p.max_iter) { if (!(it & 1)) restore_prev(); break; } // N iterations exhausted, QN.jl:250,291
#pragma unroll
    for (int k = 0; k < RPT; ++k) g_prev[k] = g[k]; // ∇_old, ∇_new = ∇_new, ∇_old, QN.jl:286
  }

  // results: θ* at offset 0 and ∇ at offset L of the state block (QN.jl:135,138); which ∇ the
  // canonical slot shows after a failure: see qn_resident.cuh
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int e = l + SG * k;
    if (p.x_out) p.x_out[j * p.x_stride + e] = x[k];
    if (p.g_out) p.g_out[j * p.g_stride + e] = g[k];
    if (p.h_out) {
      double* ho = p.h_out + j * p.h_stride + e;
#pragma unroll
      for (int r = 0; r < N; ++r) ho[(long long)p.ldh * r] = h[k][r];
    }
  }
  write_problem_scalars(p, j, l, result, iters, fevals, status);
}

template <int N, int RPT, int MINB>
__global__ void __launch_bounds__(PackedCfg<N, RPT>::kWarps * 32, MINB)
qn_resident_packed_kernel(const BatchParams p) {
  using C = PackedCfg<N, RPT>;
  constexpr int SG = C::SG;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / SG, l = lane % SG;
  const unsigned mask = (SG == 32 ? 0xffffffffu : ((1u << SG) - 1u)) << (sub * SG);
  double* mine = smem + (size_t)(warp * C::kSub + sub) * C::kPerSub;
  Pass2Col* cols = reinterpret_cast<Pass2Col*>(mine); // 16-byte aligned: kPerSub is even
  double* ys = mine + 4 * N;
  for (;;) {
    long long j = 0;
    if (l == 0) j = (long long)atomicAdd(p.counter, 1ULL);
    j = __shfl_sync(mask, j, sub * SG);
    if (j >= p.batch) break;
    packed_solve_one<N, RPT>(p, j, mask, l, ys, cols);
  }
}

}  // namespace qn
