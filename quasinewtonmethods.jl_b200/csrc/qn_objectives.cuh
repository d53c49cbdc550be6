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

  // The two GEMVs walk X (and Xᵀ) with RUNNING POINTERS: with the element index multiplied out
  // ((size_t)m * j) every load cost five integer instructions — IMAD, LEA and VIADD were 46 % of the
  // kernel's instructions, DFMA 12 % (profiles/r2_logistic_hot_sass.txt).
  // rows2: z0 = Σ_j a[j·ld] b[j], z1 = Σ_j a[second + j·ld] b[j] as rule-C2 row sums (the bits of
  // chain_rowsum for each row), b[j] read once for both rows, two fma chains in flight.
  __device__ __forceinline__ static void rows2(const double* a, const int second, const int ld, const int K,
                                               const double* b, double& out0, double& out1) {
    const int rb = chain_block(K);
    double T0 = 0.0, S0 = 0.0, T1 = 0.0, S1 = 0.0;
    int blk = 0;
    for (int r0 = 0; r0 < K; r0 += rb, ++blk) {
      const int r1 = min(r0 + rb, K);
      double t0 = 0.0, t1 = 0.0;
      const double* bb = b + r0;
      int r = r0;
      for (; r + 8 <= r1; r += 8, bb += 8) {
        double a0[8], a1[8], bv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a0[k] = __ldg(a); a1[k] = __ldg(a + second); a += ld; bv[k] = bb[k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { t0 = fma(a0[k], bv[k], t0); t1 = fma(a1[k], bv[k], t1); }
      }
      for (; r < r1; ++r, ++bb) {
        const double a0 = __ldg(a), a1 = __ldg(a + second), bv = *bb;
        a += ld;
        t0 = fma(a0, bv, t0);
        t1 = fma(a1, bv, t1);
      }
      const bool first = (blk & 15) == 0, last = (blk & 15) == 15 || r1 == K;
      S0 = first ? t0 : S0 + t0;
      S1 = first ? t1 : S1 + t1;
      if (last) { T0 = (blk < 16) ? S0 : T0 + S0; T1 = (blk < 16) ? S1 : T1 + S1; }
    }
    out0 = T0;
    out1 = T1;
  }
  // superblock2: rule-C2 super-block sb of Σ_i a[i·ld] b[i] over K terms (the bits of superblock_sum),
  // two consecutive blocks' fma chains in flight, their partials added in ascending block order.
  __device__ __forceinline__ static double superblock2(const double* a, const int ld, const int K, const int sb,
                                                       const double* b) {
    const int rb = chain_block(K);
    const int rend = min(K, (sb + 1) * 16 * rb);
    int r0 = sb * 16 * rb;
    const double* pa = a + (size_t)ld * r0;
    const size_t skip = (size_t)ld * rb;
    double S = 0.0;
    bool first = true;
    for (; r0 + 2 * rb <= rend; r0 += 2 * rb) {
      const double* pb = pa + skip;
      const double* ba = b + r0;
      const double* bb = ba + rb;
      double ta = 0.0, tb = 0.0;
      int r = 0;
      for (; r + 8 <= rb; r += 8, ba += 8, bb += 8) {
        double x0[8], y0[8], x1[8], y1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { x0[k] = __ldg(pa); x1[k] = __ldg(pb); pa += ld; pb += ld; y0[k] = ba[k]; y1[k] = bb[k]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { ta = fma(x0[k], y0[k], ta); tb = fma(x1[k], y1[k], tb); }
      }
      for (; r < rb; ++r, ++ba, ++bb) {
        const double x0 = __ldg(pa), x1 = __ldg(pb);
        pa += ld; pb += ld;
        ta = fma(x0, *ba, ta);
        tb = fma(x1, *bb, tb);
      }
      pa = pb; // block b + 1 has been consumed as well
      S = first ? ta : S + ta;
      first = false;
      S = S + tb;
    }
    for (; r0 < rend; r0 += rb) {
      const int r1 = min(r0 + rb, rend);
      double tt = 0.0;
      const double* ba = b + r0;
      for (int r = r0; r < r1; ++r, ++ba) { tt = fma(__ldg(pa), *ba, tt); pa += ld; }
      S = first ? tt : S + tt;
      first = false;
    }
    return S;
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
      // rule-C2 chains are in flight (rows2: each row's bits are those of chain_rowsum)
      auto finish_row = [&](const int i, const double z) {
        const double yi = __ldg(p.y + i);
        double sp, sg;
        softplus_sigmoid(z, sp, sg);
        vs[i] = yi * z - sp;
        rs[i] = yi - sg;
      };
      for (int i = c; i < m; i += 2 * nthr) {
        const bool two = i + nthr < m;
        double z0, z1;
        rows2(p.X + i, two ? nthr : 0, m, n, t.xs, z0, z1);
        finish_row(i, z0);
        if (two) finish_row(i + nthr, z1);
      }
    }
    const double th = (c < n) ? t.xs[c] : 0.0;
    t.stage(th * th);
    __syncthreads();
    const double pen = t.finish(n);
    // csum over the m staged values (rule C1), every warp redundantly
    double pp = (t.lane < m) ? vs[t.lane] : 0.0;
#pragma unroll 8
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
          for (int sb = hgrp; sb < ns; sb += t.tpr) sp[sb * t.nvec + row] = superblock2(Xt + row, n, m, sb, rs);
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
