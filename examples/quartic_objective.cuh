// examples/quartic_objective.cuh — a complete plug-in objective for include/qn_objective.cuh:
//   ℓ(θ) = -Σ_i [ a_i (θ_i - c_i)⁴ + ½ b_i (θ_i - c_i)² ],   ∇_i = -(4 a_i d_i³ + b_i d_i),  d = θ - c
// (separable, strictly concave for a, b > 0: the maximiser is c).  Parameter blob: a[n], b[n], c[n].
#pragma once

namespace qn {

struct UserObjective {
  struct Params {
    const double* a;
    const double* b;
    const double* c;
  };
  static Params make(const void* d_blob, size_t bytes, int n) {
    const double* v = static_cast<const double*>(d_blob);
    (void)bytes; // 3 n doubles
    return Params{v, v + n, v + 2 * n};
  }
  static constexpr int kKind = QN_OBJ_USER;
  static constexpr bool kRegisterPaths = false;
  static constexpr bool kPairRegister = false;
  static constexpr bool kMemoizeTrials = false; // true: skip a line-search evaluation whose point repeats the previous one
  __host__ __device__ static int scratch_doubles(int, const Params&) { return 0; }
  __device__ static void prologue(Team&, const Params&) {}

  template <bool GRAD>
  __device__ __forceinline__ static double eval(Team& t, const Params& p, double& g) {
    const int c = t.c;
    double term = 0.0;
    if (c < t.n) {
      const double d = t.xs[c] - p.c[c];
      const double d2 = d * d;
      const double a = p.a[c], b = p.b[c];
      term = (a * d2) * d2 + (0.5 * b) * d2;
      if (GRAD) g = -(((4.0 * a) * d2) * d + b * d);
    }
    return 0.0 - t.csum(term, t.n); // canonical sum; contains the closing barrier
  }
};

}  // namespace qn
