// qn_device.cuh — device-side building blocks shared by the resident kernels.
//
// Compile this translation unit with -fmad=false: the canonical arithmetic
// order fixes where a fused multiply-add happens, so nvcc must
// never contract a*b+c on its own.  Every fma() below is deliberate.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qn_cuda.h"

namespace qn {

constexpr double kEps = 2.220446049250313e-16;     // eps(Float64)
constexpr double kSqrtTol = 1.0 / (double)(1 << 26); // sqrttolerance(Float64), QN.jl:155

// BackTracking fields, QN.jl:72-77
struct LsParams {
  double c1, rho_hi, rho_lo;
  long long iterations;
  int order;
};

// Arguments of one batched solve (all pointers are device pointers).
struct BatchParams {
  int n;
  int max_iter;
  long long batch;
  const double* theta0;
  long long theta_stride;
  double tol;
  LsParams ls;
  double* x_out;   // θ* of problem j at x_out + j*x_stride
  long long x_stride;
  double* g_out;   // ∇  of problem j at g_out + j*g_stride
  long long g_stride;
  double* h_out;   // B⁻¹ of problem j at h_out + j*h_stride, column stride ldh (or null)
  long long h_stride;
  int ldh;
  double* ell_out;
  int* iters_out;
  int* fevals_out;
  int* status_out;
  int ntrace;
  double* trace;   // [batch][ntrace][n]
  unsigned long long* counter; // persistent-grid work queue
  // optional (option "record_exit"): [0] earliest start, [1] latest end, [2 + slot] the %globaltimer
  // stamp at which problem slot `slot` found the work queue empty — the tail of a launch made visible
  unsigned long long* exit_ns;
  int exit_cap;
  int lead_div;  // tile kernels: warp (blockIdx / lead_div) & 1 of a pair leads; 0 = the same warp in every CTA
  int lead_warp; // one-CTA-per-problem tile kernels (n = 96, 128, 160): the warp that leads
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void stamp_start(const BatchParams& p) {
  if (p.exit_ns) atomicMin(p.exit_ns, global_ns());
}
__device__ __forceinline__ void stamp_exit(const BatchParams& p, const int slot) {
  if (p.exit_ns) {
    const unsigned long long t = global_ns();
    if (slot < p.exit_cap) p.exit_ns[2 + slot] = t;
    atomicMax(p.exit_ns + 1, t);
  }
}

// Rule C2: a row sum over K terms is cut into balanced blocks of
// ceil(K / ceil(K/16)) <= 16 consecutive terms; one sequential FMA chain per block,
// block partials added in ascending order (in super-blocks of 16 blocks for K > 256).
__host__ __device__ constexpr int chain_block(int K) {
  return K <= 0 ? 1 : (K + (K + 15) / 16 - 1) / ((K + 15) / 16);
}

// One rule-C2 block: a sequential fma chain over r0 <= r < r1 from +0.0.  The operands of eight
// steps are fetched before their fmas run, so eight (global-memory) loads are in flight.
template <class F>
__device__ __forceinline__ double block_chain(int r0, int r1, F&& term) {
  double t = 0.0;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    double a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) term(r + k, a[k], b[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) t = fma(a[k], b[k], t);
  }
  for (; r < r1; ++r) {
    double a, b;
    term(r, a, b);
    t = fma(a, b, t);
  }
  return t;
}

// Blocked row sum of rule C2 for one row: term(r) returns the two fma operands through
// references.  Handles the two-level (super-block) combine for any K.
template <class F>
__device__ __forceinline__ double chain_rowsum(int K, F&& term) {
  const int rb = chain_block(K);
  double T = 0.0, S = 0.0;
  int blk = 0;
  for (int r0 = 0; r0 < K; r0 += rb, ++blk) {
    const int r1 = min(r0 + rb, K);
    const double t = block_chain(r0, r1, term);
    S = ((blk & 15) == 0) ? t : S + t;
    if ((blk & 15) == 15 || r1 == K) T = (blk < 16) ? S : T + S;
  }
  return T;
}

// One rule-C2 SUPER-block (16 blocks) of a row sum over K terms: block partials added in order.
template <class F>
__device__ __forceinline__ double superblock_sum(int K, int sb, F&& term) {
  const int rb = chain_block(K);
  double S = 0.0;
  const int rend = min(K, (sb + 1) * 16 * rb);
  bool first = true;
  for (int r0 = sb * 16 * rb; r0 < rend; r0 += rb) {
    const double t = block_chain(r0, min(r0 + rb, rend), term);
    S = first ? t : S + t;
    first = false;
  }
  return S;
}

// nanmin / nanmax, QN.jl:152-153
__device__ __forceinline__ double nanmin(double a, double b) { return a < b ? a : (isnan(b) ? a : b); }
__device__ __forceinline__ double nanmax(double a, double b) { return a < b ? b : (isnan(a) ? b : a); }

// linesearch!  QN.jl:165-232, the scalar part shared by every kernel: every thread that calls it
// runs the same code on the same values, so control flow is uniform across the team.
// trial(a) evaluates ℓ(x + a s) (step!, QN.jl:157-162) and counts the evaluation; if the caller
// already has the first trial ℓ(x + 1·s) it passes it as `first` with have_first = true.
// Returns α (0 = failure, QN.jl:193) and leaves the ℓ of the last trial in lx_last.
template <class Trial>
__device__ __forceinline__ double backtracking_linesearch(const LsParams& ls, const double l0, const double m,
                                                          Trial&& trial, double& lx_last,
                                                          const bool have_first = false, const double first = 0.0) {
  double lx0 = l0;
  double a1 = 1.0, a2 = 1.0;
  double lx1 = have_first ? first : trial(a1); // QN.jl:174
  int iterfinite = 0;
  while (!isfinite(lx1) && iterfinite < 52) { // QN.jl:178-184
    ++iterfinite;
    a1 = a2;
    a2 = 0.5 * a2;
    lx1 = trial(a2);
  }
  long long iteration = 0;
  while (!(lx1 >= fma(a2 * ls.c1, m, l0))) { // QN.jl:187
    ++iteration;
    if (iteration > ls.iterations) { lx_last = lx1; return 0.0; } // QN.jl:193
    double at;
    if (ls.order == 2 || iteration == 1) {
      at = -((m * a2) * a2) / (2.0 * fma(-m, a2, lx1 - l0)); // QN.jl:205
    } else {
      const double denom = 1.0 / ((((a1 * a1) * a2) * a2) * (a2 - a1)); // QN.jl:207
      const double r1 = fma(-m, a2, lx1 - l0);
      const double r0 = fma(-m, a1, lx0 - l0);
      const double a = ((a1 * a1) * r1 - (a2 * a2) * r0) * denom;                   // QN.jl:208
      const double b = ((((-a1) * a1) * a1) * r1 + ((a2 * a2) * a2) * r0) * denom;  // QN.jl:209
      if (fabs(a) <= kEps + kSqrtTol * fabs(a)) {  // QN.jl:211
        at = m / (2.0 * b);                        // QN.jl:212
      } else {
        const double d = nanmax(b * b - (3.0 * a) * m, 0.0); // QN.jl:215
        at = (sqrt(d) + b) / (-3.0 * a);                     // QN.jl:218
      }
    }
    a1 = a2;                           // QN.jl:222
    at = nanmin(at, a2 * ls.rho_hi);   // QN.jl:224
    a2 = nanmax(at, a2 * ls.rho_lo);   // QN.jl:225
    lx0 = lx1;                         // QN.jl:228
    lx1 = trial(a2);
  }
  lx_last = lx1;
  return a2;
}

// Butterfly of the canonical reduction (rule C1): off = 16,8,4,2,1.
// IEEE addition is commutative, so every lane ends with the same bits.
__device__ __forceinline__ double warp_butterfly(double p) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, off);
  return p;
}

// A "team" is the set of threads that owns one problem: thread c owns element c
// of every vector and row c of B⁻¹.  Here a team is one CTA.
struct Team {
  int n;        // problem dimension
  int c;        // this thread's element / row
  int lane;     // c & 31
  double* red;  // two staging buffers of `stride` doubles each
  int stride;
  int phase;
  double* xs;      // the point the objective is evaluated at (shared, nvec doubles)
  int nvec;        // length of the per-element shared arrays (threads >= nvec own no element)
  double* scratch; // objective scratch (shared)
  double* part;    // [rule-C2 blocks][nvec] block partials, present when tpr > 1
  int tpr;         // threads per row: blockDim = nvec * tpr (1 = one thread per element)

  // canonical csum (rule C1) split around a barrier the caller provides
  __device__ __forceinline__ void stage(double v) {
    if (c < stride) red[phase * stride + c] = v;
  }
  __device__ __forceinline__ double finish(int count) {
    const double* r = red + phase * stride;
    phase ^= 1;
    double p = (lane < count) ? r[lane] : 0.0;
    for (int i = lane + 32; i < count; i += 32) p = p + r[i];
    return warp_butterfly(p);
  }
  // Rule-C2 row sum over K terms for every row c < n, using ALL threads of the team: the
  // blocks of row (c mod nvec) are dealt out to its tpr threads, the partials meet in `part`,
  // and the owner (c < n) adds them in ascending order.  term(row, r, a, b) yields the two fma
  // operands.  Returns the row sum in the owner thread (0 elsewhere).  Contains a barrier
  // when tpr > 1; the next use of `part` must be separated from this one by another barrier.
  template <class F>
  __device__ __forceinline__ double rowsum(int K, F&& term) {
    if (tpr == 1) {
      if (c >= n) return 0.0;
      const int row = c;
      return chain_rowsum(K, [&](int r, double& a, double& b) { term(row, r, a, b); });
    }
    const int rb = chain_block(K);
    const int row = c % nvec, hgrp = c / nvec;
    if (row < n && hgrp < tpr) {
      for (int r0 = hgrp * rb, blk = hgrp; r0 < K; r0 += tpr * rb, blk += tpr) {
        const int r1 = min(r0 + rb, K);
        part[blk * nvec + row] = block_chain(r0, r1, [&](int r, double& a, double& b) { term(row, r, a, b); });
      }
    }
    __syncthreads();
    double T = 0.0;
    if (c < n) {
      T = part[c];
      for (int r0 = rb, blk = 1; r0 < K; r0 += rb, ++blk) T = T + part[blk * nvec + c];
    }
    return T;
  }

  // Three independent canonical sums at once (one barrier, interleaved butterflies).  Its two
  // alternating staging regions (3 × stride doubles each) lie BEHIND the two of csum/stage, so the
  // two kinds of reduction never reuse each other's buffer (resident_smem_layout with onepass).
  int phase3 = 0;
  __device__ __forceinline__ void csum3(double v0, double v1, double v2, int count, double& r0, double& r1, double& r2) {
    double* base = red + 2 * stride + phase3 * 3 * stride;
    if (c < stride) { base[c] = v0; base[stride + c] = v1; base[2 * stride + c] = v2; }
    __syncthreads();
    phase3 ^= 1;
    double p0 = (lane < count) ? base[lane] : 0.0;
    double p1 = (lane < count) ? base[stride + lane] : 0.0;
    double p2 = (lane < count) ? base[2 * stride + lane] : 0.0;
    for (int i = lane + 32; i < count; i += 32) {
      p0 = p0 + base[i];
      p1 = p1 + base[stride + i];
      p2 = p2 + base[2 * stride + i];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double t0 = __shfl_xor_sync(0xffffffffu, p0, off);
      const double t1 = __shfl_xor_sync(0xffffffffu, p1, off);
      const double t2 = __shfl_xor_sync(0xffffffffu, p2, off);
      p0 = p0 + t0; p1 = p1 + t1; p2 = p2 + t2;
    }
    r0 = p0; r1 = p1; r2 = p2;
  }

  // Thread c contributes v when c < count.  Result is identical in every thread.
  // Two staging buffers alternate, so the barrier inside the NEXT csum is what
  // protects this one's buffer from being overwritten while a slow warp reads.
  __device__ __forceinline__ double csum(double v, int count) {
    stage(v);
    __syncthreads();
    return finish(count);
  }
};

}  // namespace qn
