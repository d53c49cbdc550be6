// qn_objectives.cuh — registered device objectives.
//
// The reference's objective protocol is two generic functions,
// `logdensity(obj, θ)` and `∂logdensity!(∇, obj, θ)` (QN.jl:8-9), called at
// QN.jl:161 and QN.jl:254.  Here an objective is a type with
//
//   template <bool GRAD> static double eval(Team&, const Params&, double& g)
//
// called by every thread of the team after the point θ has been published in
// team.xs (and a barrier).  It returns ℓ(θ) in every thread and, when GRAD,
// leaves ∇[c] in g for the thread that owns element c.  eval must end with a
// barrier-containing reduction before which all of its reads of team.xs are
// done (the caller overwrites xs right after).
#pragma once

#include "qn_device.cuh"

namespace qn {

// ---------------------------------------------------------------- Rosenbrock
// test/runtests.jl:4-33 (= README.md:19-48).  Pairs (θ[i], θ[i+N]), N = n>>1,
// plus a tail term for odd n.  No fma except the explicit muladd (rule C5).
struct Rosenbrock {
  struct Params {};
  static constexpr int kKind = QN_OBJ_ROSENBROCK;
  static constexpr bool kRegisterPaths = true; // instantiate the register-resident kernels
  static constexpr bool kPairRegister = true;  // n = 128: B⁻¹ in 2×16 register tiles
  static constexpr bool kMemoizeTrials = false; // an evaluation is O(n): cheaper than the vote
  __host__ __device__ static int scratch_doubles(int, const Params&) { return 0; }
  __device__ static void prologue(Team&, const Params&) {}

  template <bool GRAD>
  __device__ __forceinline__ static double eval(Team& t, const Params&, double& g) {
    const int n = t.n, N = n >> 1, c = t.c;
    const double* xs = t.xs;
    double term = 0.0;
    if (c < N) {
      const double t1 = xs[c], t2 = xs[c + N];
      const double a = t2 - t1 * t1;
      const double b = 1.0 - t1;
      term = 100.0 * (a * a) + b * b;               // runtests.jl:9,23
      if (GRAD) g = (400.0 * a) * t1 + 2.0 * b;     // runtests.jl:24
    } else if (c < 2 * N) {
      if (GRAD) {
        const double t1 = xs[c - N], t2 = xs[c];
        g = 200.0 * (t1 * t1 - t2);                 // runtests.jl:25
      }
    } else if (c < n) {
      if (GRAD) g = 2.0 * (1.0 - xs[c]);            // runtests.jl:30
    }
    double s = 0.0 - t.csum(term, N);               // s -= term, @simd-reassociated
    if (n & 1) {
      const double d = 1.0 - xs[n - 1];
      s = -fma(d, d, -s);                           // -muladd(δ, δ, -s), runtests.jl:13,29
    }
    return s;
  }
};

// ------------------------------------------------------------ dense Gaussian
// Build-defined (the reference ships no such objective):
// ℓ(θ) = -½ (θ-μ)ᵀ Λ (θ-μ), ∇ = -Λ(θ-μ).  z = Λ(θ-μ) is a row-chain GEMV
// (rule C2) with Λ read from global memory / L2, shared by the whole batch.
struct DenseGaussian {
  struct Params {
    const double* mu;     // n
    const double* lambda; // n×n column-major, ld = n
    int lam_in_smem;      // the launcher found room for Λ beside the problem's shared memory
  };
  static constexpr int kKind = QN_OBJ_DENSE_GAUSSIAN;
  static constexpr bool kRegisterPaths = false; // no thread-per-row / 4×15-tile instantiations
  static constexpr bool kPairRegister = true;   // n = 128: B⁻¹ in 2×16 register tiles
  static constexpr bool kMemoizeTrials = true;  // an evaluation is an n×n GEMV
  // θ-μ (B doubles), then Λ itself when B⁻¹ lives in registers and the 227 KB of shared memory
  // are free: every CTA solves many problems against the same Λ, so it is fetched from L2 once
  // per CTA instead of once per objective evaluation (128 KB at n = 128).
  __host__ __device__ static int scratch_doubles(int n, const Params& p) {
    return ((n + 31) & ~31) + (p.lam_in_smem ? n * n : 0);
  }
  __device__ static void prologue(Team& t, const Params& p) {
    if (!p.lam_in_smem) return;
    double* ls = t.scratch + ((t.n + 31) & ~31);
    const int total = t.n * t.n;
    for (int i = threadIdx.x; i < total; i += blockDim.x) ls[i] = __ldg(p.lambda + i);
    __syncthreads();
  }

  template <bool GRAD>
  __device__ __forceinline__ static double eval(Team& t, const Params& p, double& g) {
    const int n = t.n, c = t.c;
    double* ds = t.scratch;
    double d = 0.0;
    if (c < n) {
      d = t.xs[c] - __ldg(p.mu + c);
      ds[c] = d;
    }
    __syncthreads();
    // z = Λ(θ-μ): rule-C2 row sums, the blocks of a row shared by the team's tpr threads
    // (K = n <= 256: one super-block).  The csum below is the barrier after this use of `part`.
    double z;
    if (p.lam_in_smem) {
      const double* lam = t.scratch + ((n + 31) & ~31);
      z = t.rowsum(n, [&](int row, int r, double& a, double& b) {
        a = lam[row + n * r];
        b = ds[r];
      });
    } else {
      const double* lam = p.lambda;
      z = t.rowsum(n, [&](int row, int r, double& a, double& b) {
        a = __ldg(lam + row + (size_t)n * r);
        b = ds[r];
      });
    }
    const double q = t.csum(d * z, n);
    if (GRAD) g = -z;
    return -0.5 * q;
  }
};

// ------------------------------------------------------- logistic regression
// Build-defined: ℓ(θ) = Σᵢ [yᵢ zᵢ - log(1+e^{zᵢ})] - ½λ‖θ‖²,
// z = Xθ, X m×n.  exp/log1p come from the CUDA math library, which is not
// bit-identical to glibc's: parity for this objective is to a tolerance.
struct Logistic {
  struct Params {
    int m;
    const double* X;  // m×n column-major (ld = m)
    const double* Xt; // n×m column-major (ld = n)
    const double* y;  // m
    double lam;
  };
  static constexpr int kKind = QN_OBJ_LOGISTIC;
  static constexpr bool kRegisterPaths = false;
  // the two GEMVs with X want the registers for loads in flight (measured: 132 ms with B⁻¹ in
  // shared memory against 170 ms with the register tiles, 1024 problems at n = 128)
  static constexpr bool kPairRegister = false;
  static constexpr bool kMemoizeTrials = true;  // an evaluation is an m×n GEMV against X in L2
  // yᵢzᵢ - softplus(zᵢ) and yᵢ - σ(zᵢ) (m each), then the super-block partials of Xᵀr ([nsb][B])
  __host__ __device__ static int nsb(int m) {
    const int rb = chain_block(m);
    return ((m + rb - 1) / rb + 15) / 16;
  }
  // [vs | rs | super-block partials of Xᵀr | the point vs/rs were computed at | its valid flag]
  __host__ __device__ static int scratch_doubles(int n, const Params& p) {
    return 2 * ((p.m + 31) & ~31) + (nsb(p.m) + 1) * ((n + 31) & ~31) + 2;
  }
  __device__ static double* memo_point(Team& t, const Params& p) {
    return t.scratch + 2 * ((p.m + 31) & ~31) + nsb(p.m) * ((t.n + 31) & ~31);
  }
  __device__ static void prologue(Team& t, const Params& p) {
    if (threadIdx.x == 0) *reinterpret_cast<int*>(memo_point(t, p) + ((t.n + 31) & ~31)) = 0;
    __syncthreads();
  }
  // e = exp(-|z|) serves both: softplus(z) = max(z, 0) + log1p(e), σ(z) = 1/(1+e) or e/(1+e)
  __device__ __forceinline__ static void softplus_sigmoid(const double z, double& sp, double& sg) {
    const double e = exp(-fabs(z));
    sp = (z > 0.0 ? z : 0.0) + log1p(e);
    sg = (z >= 0.0 ? 1.0 : e) / (1.0 + e);
  }

  template <bool GRAD>
  __device__ __forceinline__ static double eval(Team& t, const Params& p, double& g) {
    const int n = t.n, c = t.c, m = p.m;
    const int nthr = blockDim.x;
    double* vs = t.scratch;                     // yᵢzᵢ - softplus(zᵢ)
    double* rs = t.scratch + ((m + 31) & ~31);  // yᵢ - σ(zᵢ)
    // The point of the previous evaluation, bit for bit?  Then vs and rs are already those of this
    // point and the pass over X is skipped.  That is every iteration whose first trial is accepted:
    // QN.jl:254 asks for the gradient exactly where the line search just evaluated the value
    // (x + 1·s and fma(1, s, x) are the same number) — 2 MB of L2 reads per iteration instead of 3.
    double* xc = memo_point(t, p);
    int* valid = reinterpret_cast<int*>(xc + ((n + 31) & ~31));
    const bool mine = c < n;
    const bool eq = *valid != 0 && (!mine || __double_as_longlong(t.xs[c]) == __double_as_longlong(xc[c]));
    const bool same = __syncthreads_and(eq);
    if (!same) {
      if (mine) xc[c] = t.xs[c];
      if (threadIdx.x == 0) *valid = 1;
      // z = Xθ, two rows per thread at a time (i and i + nthr): θ[j] is read once for both and two
      // rule-C2 chains are in flight (chain_rowsum2: each row's bits are those of chain_rowsum)
      auto finish_row = [&](const int i, const double z) {
        const double yi = __ldg(p.y + i);
        double sp, sg;
        softplus_sigmoid(z, sp, sg);
        vs[i] = yi * z - sp;
        rs[i] = yi - sg;
      };
      for (int i = c; i < m; i += 2 * nthr) {
        const double* col = p.X + i;
        if (i + nthr < m) {
          double z0, z1;
          chain_rowsum2(n, [&](int j, double& a0, double& a1, double& b) {
            const double* e = col + (size_t)m * j;
            a0 = __ldg(e);
            a1 = __ldg(e + nthr);
            b = t.xs[j];
          }, z0, z1);
          finish_row(i, z0);
          finish_row(i + nthr, z1);
        } else {
          finish_row(i, chain_rowsum(n, [&](int j, double& a, double& b) {
            a = __ldg(col + (size_t)m * j);
            b = t.xs[j];
          }));
        }
      }
    }
    const double th = (c < n) ? t.xs[c] : 0.0;
    t.stage(th * th);
    __syncthreads();
    const double pen = t.finish(n);
    // csum over the m staged values (rule C1), every warp redundantly
    double pp = (t.lane < m) ? vs[t.lane] : 0.0;
    for (int i = t.lane + 32; i < m; i += 32) pp = pp + vs[i];
    const double S = warp_butterfly(pp);
    const double ell = fma(-(0.5 * p.lam), pen, S);
    if (GRAD) {
      // T = Xᵀr: the rule-C2 super-blocks of a row are dealt out to the team's tpr threads
      const double* Xt = p.Xt;
      auto term = [&](int row) {
        return [=](int i, double& a, double& b) {
          a = __ldg(Xt + row + (size_t)n * i);
          b = rs[i];
        };
      };
      if (t.tpr == 1) {
        if (c < n) g = fma(-p.lam, th, chain_rowsum(m, term(c)));
      } else {
        double* sp = t.scratch + 2 * ((m + 31) & ~31); // [nsb][nvec]
        const int row = c % t.nvec, hgrp = c / t.nvec, ns = nsb(m);
        if (row < n && hgrp < t.tpr)
          for (int sb = hgrp; sb < ns; sb += t.tpr) sp[sb * t.nvec + row] = superblock_sum_ilp2(m, sb, term(row));
        __syncthreads();
        if (c < n) {
          double T = sp[c];
          for (int sb = 1; sb < ns; ++sb) T = T + sp[sb * t.nvec + c];
          g = fma(-p.lam, th, T);
        }
      }
    }
    // every read of xs / vs / rs above happens before the caller's next barrier
    // only if we close with one: the next eval overwrites vs/rs straight away.
    __syncthreads();
    return ell;
  }
};

}  // namespace qn
