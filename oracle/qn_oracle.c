/*
 * qn_oracle.c — CPU oracle for the BFGS hot path.  TEST INFRASTRUCTURE ONLY
 * (see qn_oracle.h for who may load it and for the parity-pinning status).
 *
 * Every function cites the reference lines it restates.  "QN.jl" abbreviates
 * /root/reference/src/QuasiNewtonMethods.jl.
 *
 * Build with -ffp-contract=off: the compiler must never fuse or reassociate.
 * Every fused multiply-add below is an explicit fma() and is part of the
 * canonical order.
 *
 * CANONICAL ORDER (what LoopVectorization's @turbo decides in the reference
 * and what is therefore fixed here; the CUDA kernels do exactly the same):
 *
 *  C1  csum(v, n): 32 lane accumulators.  p[l] = v[l] (+0.0 when l >= n);
 *      then p[i mod 32] += v[i] for i = 32, 33, ... in ascending i; then a
 *      butterfly  p[l] = p[l] + p[l xor off]  for off = 16, 8, 4, 2, 1; the
 *      result is p[0].  Dot products are csum over individually rounded
 *      products a[i]*b[i].  (The SIMD analogue: vector accumulators, then a
 *      horizontal reduce.)
 *  C2  Row sums over K terms (the GEMV at QN.jl:44-52, the fused update+GEMV
 *      at QN.jl:56-67, and the GEMVs inside the build-defined objectives):
 *      the terms are cut into balanced blocks of rb = ceil(K / ceil(K/16))
 *      (<= 16) consecutive columns.  Inside a block: one sequential FMA chain
 *      in ascending order started from +0.0.  Blocks are grouped 16 at a time
 *      into super-blocks; inside a super-block the block partials are added
 *      in ascending order (S = b0; S = S + b1; ...), and the super-block sums
 *      are added in ascending order the same way.  K <= 256 has one
 *      super-block.  (@turbo vectorises over the unit-stride index c and
 *      carries the r-sum in each lane with several accumulators; this is the
 *      same shape with the accumulator count fixed.  On the GPU a block is the
 *      16-or-fewer columns one thread's register tile covers.)
 *  C3  t += a*b, muladd(a,b,c), x + alpha*s under @turbo, and x +- a*b under
 *      @fastmath with exactly one product operand are one fma.  Expressions
 *      with two product operands (a*b - c*d) round both products.
 *  C4  The rank-2 update is ((H + (c1*s[r])*s[c]) - u[r]*s[c]) - u[c]*s[r]
 *      as three chained fma, in the source's left-to-right order (QN.jl:61);
 *      c1*s[r] is rounded once per r.
 *  C5  Objective code that the reference writes without @turbo/@fastmath
 *      (Rosenbrock, test/runtests.jl:4-33) uses no fma except its explicit
 *      muladd; its @simd sum is a csum.
 */
#define _GNU_SOURCE
#include "qn_oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(QO_NO_CLONES)
#define QO_HOT __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define QO_HOT
#endif

#define QO_W 32
#define QO_CB 32 /* rows per register block in the row-chain loops */
#define QO_EPS 2.220446049250313e-16 /* eps(Float64) */

/* ------------------------------------------------------------------ C1 */
static double qo_csum(const double* v, int n) {
  double p[QO_W];
  for (int l = 0; l < QO_W; ++l) p[l] = (l < n) ? v[l] : 0.0;
  for (int i = QO_W; i < n; ++i) p[i & (QO_W - 1)] = p[i & (QO_W - 1)] + v[i];
  /* butterfly p[l] + p[l ^ off]: after each level lanes l and l^off hold the same
   * value (IEEE addition commutes), so only the low half needs computing. */
  for (int l = 0; l < 16; ++l) p[l] = p[l] + p[l + 16];
  for (int l = 0; l < 8; ++l) p[l] = p[l] + p[l + 8];
  for (int l = 0; l < 4; ++l) p[l] = p[l] + p[l + 4];
  for (int l = 0; l < 2; ++l) p[l] = p[l] + p[l + 2];
  return p[0] + p[1];
}

/* rule C2 block length for a chain of K terms */
int qo_chain_block(int K) {
  if (K <= 0) return 1;
  const int nb = (K + 15) / 16;
  return (K + nb - 1) / nb;
}

/* rule C2 combine step: fold the chain partial t of block `blk` into the super-block
 * sums S and the row sums T.  `last` marks the final block of the row. */
static inline void qo_combine(int n, int blk, int last, const double* restrict t,
                              double* restrict S, double* restrict T) {
  const int in_sb = blk & 15;
  if (in_sb == 0) for (int c = 0; c < n; ++c) S[c] = t[c];
  else for (int c = 0; c < n; ++c) S[c] = S[c] + t[c];
  if (in_sb == 15 || last) {
    if (blk < 16) for (int c = 0; c < n; ++c) T[c] = S[c];
    else for (int c = 0; c < n; ++c) T[c] = T[c] + S[c];
  }
}

int qo_state_ld(int n, int native_w) {
  /* pick_vector_width(Val(P), Float64) = min(native width, nextpow2(P));
   * align(P, W) rounds P up to a multiple of W.  QN.jl:96-99 */
  int w = 1;
  while (w < n && w < native_w) w <<= 1;
  return ((n + w - 1) / w) * w;
}

int64_t qo_state_len(int n, int L) { return 7 * (int64_t)L + (int64_t)n * L; }

/* ---------------------------------------------- objectives (QN.jl:8-9) */

/* Rosenbrock value: test/runtests.jl:5-17 */
static double rosen_value(int n, const double* th, double* scratch) {
  const int N = n >> 1;
  for (int i = 0; i < N; ++i) {
    const double t1 = th[i], t2 = th[i + N];
    const double a = t2 - t1 * t1;
    const double b = 1.0 - t1;
    scratch[i] = 100.0 * (a * a) + b * b; /* 100(θ[i+N]-θ[i]^2)^2 + (1-θ[i])^2 */
  }
  double s = 0.0 - qo_csum(scratch, N); /* s -= term, lane-reassociated (@simd) */
  if (n & 1) {
    const double d = 1.0 - th[n - 1];
    s = -fma(d, d, -s); /* -muladd(δ, δ, -s), runtests.jl:13 */
  }
  return s;
}

/* Rosenbrock value + gradient: test/runtests.jl:19-33 */
static double rosen_grad(double* g, int n, const double* th, double* scratch) {
  const int N = n >> 1;
  for (int i = 0; i < N; ++i) {
    const double t1 = th[i], t2 = th[i + N];
    const double a = t2 - t1 * t1;
    const double b = 1.0 - t1;
    scratch[i] = 100.0 * (a * a) + b * b;
    g[i] = (400.0 * a) * t1 + 2.0 * b;    /* 400(θ[i+N]-θ[i]^2)*θ[i] + 2(1-θ[i]) */
    g[i + N] = 200.0 * (t1 * t1 - t2);    /* 200(θ[i]^2 - θ[i+N]) */
  }
  double s = 0.0 - qo_csum(scratch, N);
  if (n & 1) {
    const double d = 1.0 - th[n - 1];
    s = -fma(d, d, -s);
    g[n - 1] = 2.0 * d;
  }
  return s;
}

/* Dense Gaussian (build-defined): z = Λ(θ-μ) by rule C2
 * row chains, ∇ = -z, ell = -1/2 csum((θ-μ)[c] z[c]).  scratch: 4n doubles. */
QO_HOT static double gauss_eval(double* g, int n, const qo_gaussian* p, const double* th,
                                int rblock, double* scratch) {
  double* restrict d = scratch;
  double* restrict z = scratch + n;
  double* restrict t = scratch + 2 * n;
  double* restrict S = scratch + 3 * n;
  const int rb = rblock > 0 ? rblock : qo_chain_block(n);
  for (int r = 0; r < n; ++r) d[r] = th[r] - p->mu[r];
  int blk = 0;
  for (int r0 = 0; r0 < n; r0 += rb, ++blk) {
    const int r1 = (r0 + rb < n) ? r0 + rb : n;
    for (int c = 0; c < n; ++c) t[c] = 0.0;
    for (int r = r0; r < r1; ++r) {
      const double dr = d[r];
      const double* restrict col = p->lambda + (size_t)n * r;
      for (int c = 0; c < n; ++c) t[c] = fma(col[c], dr, t[c]);
    }
    qo_combine(n, blk, r1 == n, t, S, z);
  }
  for (int c = 0; c < n; ++c) t[c] = d[c] * z[c];
  const double q = qo_csum(t, n);
  if (g) for (int c = 0; c < n; ++c) g[c] = -z[c];
  return -0.5 * q;
}

static double qo_softplus(double z) { return z > 0.0 ? z + log1p(exp(-z)) : log1p(exp(z)); }
static double qo_sigmoid(double z) {
  if (z >= 0.0) return 1.0 / (1.0 + exp(-z));
  const double e = exp(z);
  return e / (1.0 + e);
}

/* Logistic regression (build-defined).  Both GEMVs (z = Xθ over
 * n terms, Xᵀr over m terms) are rule C2 row chains.  scratch: 4m + 4n doubles. */
QO_HOT static double logit_eval(double* g, int n, const qo_logistic* p, const double* th,
                                int rblock, double* scratch) {
  const int m = p->m;
  double* restrict z = scratch;
  double* restrict v = scratch + m;
  double* restrict tm = scratch + 2 * m;
  double* restrict Sm = scratch + 3 * m;
  double* restrict w = scratch + 4 * m;
  double* restrict tn = scratch + 4 * m + n;
  double* restrict Sn = scratch + 4 * m + 2 * n;
  double* restrict Tn = scratch + 4 * m + 3 * n;
  {
    const int rb = rblock > 0 ? rblock : qo_chain_block(n);
    int blk = 0;
    for (int j0 = 0; j0 < n; j0 += rb, ++blk) {
      const int j1 = (j0 + rb < n) ? j0 + rb : n;
      for (int i = 0; i < m; ++i) tm[i] = 0.0;
      for (int j = j0; j < j1; ++j) {
        const double tj = th[j];
        const double* restrict col = p->X + (size_t)m * j;
        for (int i = 0; i < m; ++i) tm[i] = fma(col[i], tj, tm[i]);
      }
      qo_combine(m, blk, j1 == n, tm, Sm, z);
    }
  }
  for (int i = 0; i < m; ++i) v[i] = p->y[i] * z[i] - qo_softplus(z[i]);
  const double S = qo_csum(v, m);
  for (int j = 0; j < n; ++j) w[j] = th[j] * th[j];
  const double pen = qo_csum(w, n);
  const double ell = fma(-(0.5 * p->lam), pen, S);
  if (g) {
    for (int i = 0; i < m; ++i) v[i] = p->y[i] - qo_sigmoid(z[i]);
    const int rb = rblock > 0 ? rblock : qo_chain_block(m);
    int blk = 0;
    for (int i0 = 0; i0 < m; i0 += rb, ++blk) {
      const int i1 = (i0 + rb < m) ? i0 + rb : m;
      for (int j = 0; j < n; ++j) tn[j] = 0.0;
      for (int i = i0; i < i1; ++i) {
        const double ri = v[i];
        const double* restrict col = p->Xt + (size_t)n * i;
        for (int j = 0; j < n; ++j) tn[j] = fma(col[j], ri, tn[j]);
      }
      qo_combine(n, blk, i1 == m, tn, Sn, Tn);
    }
    for (int j = 0; j < n; ++j) g[j] = fma(-p->lam, th[j], Tn[j]);
  }
  return ell;
}

static size_t qo_scratch_len(const qo_model* model, int n) {
  size_t len = 10 * (size_t)n + 64;
  if (model->kind == QO_LOGISTIC) {
    const qo_logistic* p = (const qo_logistic*)model->params;
    len += 4 * (size_t)p->m;
  }
  return len;
}

static double qo_eval(double* g, const qo_model* model, int n, const double* th, int rblock,
                      double* scratch) {
  switch (model->kind) {
    case QO_ROSENBROCK:
      return g ? rosen_grad(g, n, th, scratch) : rosen_value(n, th, scratch);
    case QO_DENSE_GAUSSIAN:
      return gauss_eval(g, n, (const qo_gaussian*)model->params, th, rblock, scratch);
    case QO_LOGISTIC:
      return logit_eval(g, n, (const qo_logistic*)model->params, th, rblock, scratch);
    default:
      return NAN;
  }
}

double qo_logdensity(const qo_model* model, int n, const double* theta) {
  double* scratch = (double*)malloc(sizeof(double) * qo_scratch_len(model, n));
  const double v = qo_eval(NULL, model, n, theta, 0, scratch);
  free(scratch);
  return v;
}

double qo_dlogdensity(double* grad, const qo_model* model, int n, const double* theta) {
  double* scratch = (double*)malloc(sizeof(double) * qo_scratch_len(model, n));
  const double v = qo_eval(grad, model, n, theta, 0, scratch);
  free(scratch);
  return v;
}

/* ------------------------------------------- BFGS_update! (QN.jl:34-69) */
QO_HOT double qo_bfgs_update(int n, int L, double* H, double* u, double* y, const double* s,
                             const double* g_new, const double* g_old, int rblock,
                             double* scratch) {
  double* restrict t = scratch;         /* chain accumulators, one per row c */
  double* restrict T = scratch + n;     /* row sums */
  double* restrict w = scratch + 2 * n; /* c1*s[r] */
  double* restrict v = scratch + 3 * n; /* products for csum */
  double* restrict S = scratch + 4 * n; /* super-block sums (rule C2) */
  const double* restrict sv = s;        /* s, u, H and the scratch rows never overlap */
  double* restrict uv = u;
  const int rb = rblock > 0 ? rblock : qo_chain_block(n);

  /* QN.jl:36-41  y = ∇_old - ∇_new ; sᵀy */
  for (int i = 0; i < n; ++i) {
    const double yi = g_old[i] - g_new[i];
    y[i] = yi;
    v[i] = s[i] * yi;
  }
  const double sy = qo_csum(v, n);
  const double rho = 1.0 / sy; /* QN.jl:42, no guard */

  /* QN.jl:44-52  t[c] = Σ_r y[r] H[c,r] ; u = t ρ ; yᵀHy.
   * Loop nest: strips of QO_CB rows keep their chain accumulators in registers while r
   * runs over one block; each row's chain is sequential in ascending r (rule C2). */
  int blk = 0;
  for (int r0 = 0; r0 < n; r0 += rb, ++blk) {
    const int r1 = (r0 + rb < n) ? r0 + rb : n;
    for (int c0 = 0; c0 < n; c0 += QO_CB) {
      double acc[QO_CB];
      if (c0 + QO_CB <= n) {
        for (int k = 0; k < QO_CB; ++k) acc[k] = 0.0;
        for (int r = r0; r < r1; ++r) {
          const double yr = y[r];
          const double* restrict col = H + (size_t)L * r + c0;
          for (int k = 0; k < QO_CB; ++k) acc[k] = fma(yr, col[k], acc[k]);
        }
        for (int k = 0; k < QO_CB; ++k) t[c0 + k] = acc[k];
      } else {
        const int cb = n - c0;
        for (int k = 0; k < cb; ++k) acc[k] = 0.0;
        for (int r = r0; r < r1; ++r) {
          const double yr = y[r];
          const double* restrict col = H + (size_t)L * r + c0;
          for (int k = 0; k < cb; ++k) acc[k] = fma(yr, col[k], acc[k]);
        }
        for (int k = 0; k < cb; ++k) t[c0 + k] = acc[k];
      }
    }
    qo_combine(n, blk, r1 == n, t, S, T);
  }
  for (int c = 0; c < n; ++c) {
    u[c] = T[c] * rho;
    v[c] = T[c] * y[c];
  }
  const double yHy = qo_csum(v, n);

  /* QN.jl:54 */
  const double c1 = fma(yHy, rho, 1.0) * rho;

  /* QN.jl:56-67  rank-2 update fused with d = H⁺ ∇_new */
  for (int r = 0; r < n; ++r) w[r] = c1 * s[r];
  blk = 0;
  for (int r0 = 0; r0 < n; r0 += rb, ++blk) {
    const int r1 = (r0 + rb < n) ? r0 + rb : n;
    for (int c0 = 0; c0 < n; c0 += QO_CB) {
      double acc[QO_CB], sc[QO_CB], nuc[QO_CB];
      const int cb = (c0 + QO_CB <= n) ? QO_CB : n - c0;
      for (int k = 0; k < cb; ++k) { acc[k] = 0.0; sc[k] = sv[c0 + k]; nuc[k] = -uv[c0 + k]; }
      if (cb == QO_CB) {
        for (int r = r0; r < r1; ++r) {
          const double wr = w[r], nur = -uv[r], sr = sv[r], gr = g_new[r];
          double* restrict col = H + (size_t)L * r + c0;
          for (int k = 0; k < QO_CB; ++k) {
            double h = col[k];
            h = fma(wr, sc[k], h);   /*  + (c1 s[r]) s[c] */
            h = fma(nur, sc[k], h);  /*  - u[r] s[c]      */
            h = fma(nuc[k], sr, h);  /*  - u[c] s[r]      */
            col[k] = h;
            acc[k] = fma(h, gr, acc[k]);
          }
        }
      } else {
        for (int r = r0; r < r1; ++r) {
          const double wr = w[r], nur = -uv[r], sr = sv[r], gr = g_new[r];
          double* restrict col = H + (size_t)L * r + c0;
          for (int k = 0; k < cb; ++k) {
            double h = col[k];
            h = fma(wr, sc[k], h);
            h = fma(nur, sc[k], h);
            h = fma(nuc[k], sr, h);
            col[k] = h;
            acc[k] = fma(h, gr, acc[k]);
          }
        }
      }
      for (int k = 0; k < cb; ++k) t[c0 + k] = acc[k];
    }
    qo_combine(n, blk, r1 == n, t, S, T);
  }
  for (int c = 0; c < n; ++c) {
    y[c] = T[c]; /* new direction overwrites y, QN.jl:65 */
    v[c] = T[c] * g_new[c];
  }
  return qo_csum(v, n); /* m, QN.jl:66-68 */
}

/* ------------------------------------------------------------------------------------
 * One-pass deferred-update variant — NOT the reference order.
 * The stored matrix lags one rank-2 update behind: `pend` = {ps[n], pu[n], c1, flag} is the
 * update of the previous iteration, not yet applied.  One sweep over H applies it (rule C4
 * order) and accumulates BOTH GEMVs, t = H y and w = H ∇_new (rule C2 chains); the new
 * direction then needs no second sweep:
 *     d = H⁺∇ = ((w + (c₁ s)(sᵀ∇)) - u (sᵀ∇)) - s (uᵀ∇)         (three fma per element)
 * and this iteration's update (s, u, c₁) becomes the pending one.  16 n² bytes of matrix
 * traffic per iteration instead of 24 n².  Same mathematics as QN.jl:34-69, different
 * rounding; selected with rblock = -1 and gated by tolerances against the reference order. */
QO_HOT static double qo_bfgs_update_onepass(int n, int L, double* H, double* u, double* y,
                                            const double* s, const double* g_new,
                                            const double* g_old, double* pend, double* scratch) {
  double* restrict t = scratch;
  double* restrict T = scratch + n;
  double* restrict wv = scratch + 2 * n;
  double* restrict W = scratch + 3 * n;
  double* restrict S1 = scratch + 4 * n;
  double* restrict S2 = scratch + 5 * n;
  double* restrict v = scratch + 6 * n;
  double* restrict pw = scratch + 7 * n;
  double* restrict ps = pend;
  double* restrict pu = pend + n;
  const double pc1 = pend[2 * n];
  const int pending = pend[2 * n + 1] != 0.0;
  const int rb = qo_chain_block(n);

  for (int i = 0; i < n; ++i) {
    const double yi = g_old[i] - g_new[i];
    y[i] = yi;
    v[i] = s[i] * yi;
  }
  const double sy = qo_csum(v, n);
  const double rho = 1.0 / sy;
  if (pending) for (int r = 0; r < n; ++r) pw[r] = pc1 * ps[r];

  int blk = 0;
  for (int r0 = 0; r0 < n; r0 += rb, ++blk) {
    const int r1 = (r0 + rb < n) ? r0 + rb : n;
    for (int c = 0; c < n; ++c) { t[c] = 0.0; wv[c] = 0.0; }
    for (int r = r0; r < r1; ++r) {
      double* restrict col = H + (size_t)L * r;
      const double yr = y[r], gr = g_new[r];
      if (pending) {
        const double wr = pw[r], nur = -pu[r], sr = ps[r];
        for (int c = 0; c < n; ++c) {
          double h = col[c];
          h = fma(wr, ps[c], h);
          h = fma(nur, ps[c], h);
          h = fma(-pu[c], sr, h);
          col[c] = h;
          t[c] = fma(yr, h, t[c]);
          wv[c] = fma(h, gr, wv[c]);
        }
      } else {
        for (int c = 0; c < n; ++c) {
          const double h = col[c];
          t[c] = fma(yr, h, t[c]);
          wv[c] = fma(h, gr, wv[c]);
        }
      }
    }
    qo_combine(n, blk, r1 == n, t, S1, T);
    qo_combine(n, blk, r1 == n, wv, S2, W);
  }
  for (int c = 0; c < n; ++c) {
    u[c] = T[c] * rho;
    v[c] = T[c] * y[c];
  }
  const double yHy = qo_csum(v, n);
  const double c1 = fma(yHy, rho, 1.0) * rho;
  for (int c = 0; c < n; ++c) v[c] = s[c] * g_new[c];
  const double sg = qo_csum(v, n);
  for (int c = 0; c < n; ++c) v[c] = u[c] * g_new[c];
  const double ug = qo_csum(v, n);
  for (int c = 0; c < n; ++c) {
    double d = fma(c1 * s[c], sg, W[c]);
    d = fma(-u[c], sg, d);
    d = fma(-s[c], ug, d);
    y[c] = d;
    v[c] = d * g_new[c];
  }
  for (int c = 0; c < n; ++c) { ps[c] = s[c]; pu[c] = u[c]; }
  pend[2 * n] = c1;
  pend[2 * n + 1] = 1.0;
  return qo_csum(v, n);
}

/* apply a pending update to the stored matrix (end of a one-pass solve) */
static void qo_flush_pending(int n, int L, double* H, double* pend) {
  if (pend[2 * n + 1] == 0.0) return;
  const double* ps = pend;
  const double* pu = pend + n;
  const double pc1 = pend[2 * n];
  for (int r = 0; r < n; ++r) {
    const double wr = pc1 * ps[r], nur = -pu[r], sr = ps[r];
    double* col = H + (size_t)L * r;
    for (int c = 0; c < n; ++c) {
      double h = col[c];
      h = fma(wr, ps[c], h);
      h = fma(nur, ps[c], h);
      h = fma(-pu[c], sr, h);
      col[c] = h;
    }
  }
  pend[2 * n + 1] = 0.0;
}

/* nanmin / nanmax: QN.jl:152-153 */
static double qo_nanmin(double a, double b) { return a < b ? a : (isnan(b) ? a : b); }
static double qo_nanmax(double a, double b) { return a < b ? b : (isnan(a) ? b : a); }

/* step!: QN.jl:157-162 */
static double qo_step(int n, double* x_new, const double* x_old, const double* s,
                      const qo_model* model, double alpha, int rblock, double* scratch) {
  for (int i = 0; i < n; ++i) x_new[i] = fma(alpha, s[i], x_old[i]);
  return qo_eval(NULL, model, n, x_new, rblock, scratch);
}

/* linesearch!: QN.jl:165-232 */
static double qo_linesearch_impl(int n, double* x_new, const double* x_old, const double* s,
                                 const qo_model* model, double l0, double m,
                                 const qo_backtracking* ls, int64_t* fevals, int rblock,
                                 double* scratch) {
  const double c1 = ls->c1, rho_hi = ls->rho_hi, rho_lo = ls->rho_lo;
  const int64_t iterations = ls->iterations;
  const double sqrttol = 1.0 / (double)(1 << 26); /* sqrttolerance(Float64), QN.jl:155 */

  double lx0 = l0, lx1 = l0;
  double a1 = 1.0, a2 = 1.0;
  lx1 = qo_step(n, x_new, x_old, s, model, a1, rblock, scratch); /* QN.jl:174 */
  ++*fevals;

  /* QN.jl:178-184: halve until finite, at most significand_bits(Float64)=52 times */
  int iterfinite = 0;
  while (!isfinite(lx1) && iterfinite < 52) {
    ++iterfinite;
    a1 = a2;
    a2 = 0.5 * a2;
    lx1 = qo_step(n, x_new, x_old, s, model, a2, rblock, scratch);
    ++*fevals;
  }

  int64_t iteration = 0;
  /* QN.jl:187: while !(ℓx₁ ≥ ℓ₀ + α₂ c₁ m) */
  while (!(lx1 >= fma(a2 * c1, m, l0))) {
    ++iteration;
    if (iteration > iterations) return 0.0; /* QN.jl:193 */
    double at;
    if (ls->order == 2 || iteration == 1) {
      /* QN.jl:205: -(m α₂ α₂) / (2 (ℓx₁ - ℓ₀ - m α₂)) */
      at = -((m * a2) * a2) / (2.0 * fma(-m, a2, lx1 - l0));
    } else {
      /* QN.jl:207-209 */
      const double denom = 1.0 / ((((a1 * a1) * a2) * a2) * (a2 - a1));
      const double r1 = fma(-m, a2, lx1 - l0);
      const double r0 = fma(-m, a1, lx0 - l0);
      const double a = ((a1 * a1) * r1 - (a2 * a2) * r0) * denom;
      const double b = ((((-a1) * a1) * a1) * r1 + ((a2 * a2) * a2) * r0) * denom;
      if (fabs(a) <= QO_EPS + sqrttol * fabs(a)) { /* QN.jl:211 */
        at = m / (2.0 * b);                        /* QN.jl:212 */
      } else {
        const double d = qo_nanmax(b * b - (3.0 * a) * m, 0.0); /* QN.jl:215 */
        at = (sqrt(d) + b) / (-3.0 * a);                         /* QN.jl:218 */
      }
    }
    a1 = a2;                                   /* QN.jl:222 */
    at = qo_nanmin(at, a2 * rho_hi);           /* QN.jl:224 */
    a2 = qo_nanmax(at, a2 * rho_lo);           /* QN.jl:225 */
    lx0 = lx1;                                 /* QN.jl:228 */
    lx1 = qo_step(n, x_new, x_old, s, model, a2, rblock, scratch);
    ++*fevals;
  }
  return a2;
}

double qo_linesearch(int n, double* x_new, const double* x_old, const double* s,
                     const qo_model* model, double l0, double m, const qo_backtracking* ls,
                     int64_t* fevals) {
  double* scratch = (double*)malloc(sizeof(double) * qo_scratch_len(model, n));
  const double a = qo_linesearch_impl(n, x_new, x_old, s, model, l0, m, ls, fevals, 0, scratch);
  free(scratch);
  return a;
}

/* optimize!: QN.jl:237-292 */
static double qo_optimize_impl(int n, int L, double* state, const qo_model* model,
                               const double* x0, const qo_backtracking* ls, double tol,
                               int64_t max_iter, int rblock, int ntrace, double* trace,
                               double* trace_scalars, qo_info* info, double* scratch) {
  /* field order of BFGSState, QN.jl:84-92; views QN.jl:135-142 */
  double* x_old = state;
  double* g_new = state + (size_t)L;
  double* x_new = state + 2 * (size_t)L;
  double* g_old = state + 3 * (size_t)L;
  double* y = state + 4 * (size_t)L;
  double* s = state + 5 * (size_t)L;
  double* u = state + 6 * (size_t)L;
  double* H = state + 7 * (size_t)L;
  double* const g_canon = g_new;

  int64_t fevals = 0, gevals = 0, iters = 0;
  int status = QO_MAX_ITER;
  double result = NAN;
  const int onepass = (rblock == -1); /* deferred-update variant, see qo_bfgs_update_onepass */
  double* pend = NULL;
  if (onepass) {
    rblock = 0;
    pend = (double*)calloc((size_t)2 * n + 2, sizeof(double));
  }

  memcpy(x_old, x0, sizeof(double) * (size_t)n); /* QN.jl:248 */
  for (int64_t it = 1; it <= max_iter; ++it) {   /* QN.jl:250,253 */
    const double l0 = qo_eval(g_new, model, n, x_old, rblock, scratch); /* QN.jl:254 */
    ++fevals;
    ++gevals;
    if (!isfinite(l0)) { status = QO_NONFINITE; break; } /* QN.jl:255 */
    /* QN.jl:257: maximum(abs, ∇) < tol  (a NaN entry makes it false) */
    int conv = 1;
    for (int i = 0; i < n; ++i) conv &= (fabs(g_new[i]) < tol);
    if (conv) {
      if (g_new != g_canon) memcpy(g_canon, g_new, sizeof(double) * (size_t)n); /* QN.jl:258-259 */
      status = QO_CONVERGED;
      result = l0;
      break;
    }
    double m;
    if (it == 1) {
      m = -1.0; /* QN.jl:263-264 */
    } else {
      m = onepass ? qo_bfgs_update_onepass(n, L, H, u, y, s, g_new, g_old, pend, scratch)
                  : qo_bfgs_update(n, L, H, u, y, s, g_new, g_old, rblock, scratch); /* QN.jl:267 */
      double* tmp = s; s = y; y = tmp;                                          /* QN.jl:268 */
    }
    if (m <= 0.0) { /* QN.jl:272-280 */
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) H[c + (size_t)L * r] = (r == c) ? 1.0 : 0.0; /* QN.jl:144-148 */
      if (onepass) pend[2 * n + 1] = 0.0; /* the reset discards the pending update too */
      double* v = scratch;
      for (int i = 0; i < n; ++i) {
        const double si = g_new[i];
        s[i] = si;
        v[i] = si * si;
      }
      m = qo_csum(v, n);
    }
    int64_t fe = 0;
    const double alpha = qo_linesearch_impl(n, x_new, x_old, s, model, l0, m, ls, &fe, rblock,
                                            scratch); /* QN.jl:283 */
    fevals += fe;
    if (alpha == 0.0) { status = QO_LINESEARCH_FAILED; break; } /* QN.jl:284 */
    /* update_state!: QN.jl:21-27 */
    for (int i = 0; i < n; ++i) {
      const double si = alpha * s[i];
      s[i] = si;
      x_old[i] = x_old[i] + si;
    }
    ++iters;
    if (trace && it <= ntrace) {
      memcpy(trace + (size_t)(it - 1) * n, x_old, sizeof(double) * (size_t)n);
      if (trace_scalars) {
        double* ts = trace_scalars + (size_t)(it - 1) * 4;
        ts[0] = l0; ts[1] = m; ts[2] = alpha; ts[3] = (double)fe;
      }
    }
    double* tmp = g_old; g_old = g_new; g_new = tmp; /* QN.jl:286 */
  }
  if (onepass) {
    qo_flush_pending(n, L, H, pend);
    free(pend);
  }
  if (info) {
    info->iters = iters;
    info->fevals = fevals;
    info->gevals = gevals;
    info->status = status;
  }
  return result; /* NaN unless converged, QN.jl:291 */
}

double qo_optimize(int n, int L, double* state, const qo_model* model, const double* x0,
                   const qo_backtracking* ls, double tol, int64_t max_iter, int rblock,
                   int ntrace, double* trace, double* trace_scalars, qo_info* info) {
  double* scratch = (double*)malloc(sizeof(double) * qo_scratch_len(model, n));
  const double r = qo_optimize_impl(n, L, state, model, x0, ls, tol, max_iter, rblock, ntrace,
                                    trace, trace_scalars, info, scratch);
  free(scratch);
  return r;
}

/* ------------------------------------------------------- batch driver */
int qo_num_cores(void) {
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) return CPU_COUNT(&set);
  return 1;
}

typedef struct {
  int n;
  int64_t batch;
  const double* theta0;
  const qo_model* model;
  const qo_backtracking* ls;
  double tol;
  int64_t max_iter;
  int rblock;
  double *x_out, *g_out, *ell_out;
  int64_t *iters_out, *fevals_out;
  int32_t* status_out;
  int ntrace;
  double* trace;
  atomic_long next;
  atomic_long converged;
} qo_batch_job;

static void* qo_batch_worker(void* arg) {
  qo_batch_job* job = (qo_batch_job*)arg;
  const int n = job->n;
  const int L = (n + 7) & ~7;
  double* state = (double*)aligned_alloc(64, sizeof(double) * (size_t)qo_state_len(n, L));
  double* scratch = (double*)malloc(sizeof(double) * qo_scratch_len(job->model, n));
  for (;;) {
    /* grab a small block of problems at a time (dynamic schedule) */
    const long j0 = atomic_fetch_add(&job->next, 16);
    if (j0 >= job->batch) break;
    const long j1 = (j0 + 16 < job->batch) ? j0 + 16 : job->batch;
    for (long j = j0; j < j1; ++j) {
      qo_info info;
      double* tr = job->trace ? job->trace + (size_t)j * job->ntrace * n : NULL;
      const double ell = qo_optimize_impl(n, L, state, job->model, job->theta0 + (size_t)j * n,
                                          job->ls, job->tol, job->max_iter, job->rblock,
                                          job->ntrace, tr, NULL, &info, scratch);
      if (job->x_out) memcpy(job->x_out + (size_t)j * n, state, sizeof(double) * (size_t)n);
      if (job->g_out) memcpy(job->g_out + (size_t)j * n, state + L, sizeof(double) * (size_t)n);
      if (job->ell_out) job->ell_out[j] = ell;
      if (job->iters_out) job->iters_out[j] = info.iters;
      if (job->fevals_out) job->fevals_out[j] = info.fevals;
      if (job->status_out) job->status_out[j] = info.status;
      if (info.status == QO_CONVERGED) atomic_fetch_add(&job->converged, 1);
    }
  }
  free(scratch);
  free(state);
  return NULL;
}

int64_t qo_optimize_batch(int n, int64_t batch, const double* theta0, const qo_model* model,
                          const qo_backtracking* ls, double tol, int64_t max_iter, int rblock,
                          double* x_out, double* g_out, double* ell_out, int64_t* iters_out,
                          int64_t* fevals_out, int32_t* status_out, int ntrace, double* trace,
                          int nthreads) {
  qo_batch_job job;
  job.n = n; job.batch = batch; job.theta0 = theta0; job.model = model; job.ls = ls;
  job.tol = tol; job.max_iter = max_iter; job.rblock = rblock;
  job.x_out = x_out; job.g_out = g_out; job.ell_out = ell_out;
  job.iters_out = iters_out; job.fevals_out = fevals_out; job.status_out = status_out;
  job.ntrace = ntrace; job.trace = trace;
  atomic_init(&job.next, 0);
  atomic_init(&job.converged, 0);
  if (nthreads <= 0) nthreads = qo_num_cores();
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, qo_batch_worker, &job);
  qo_batch_worker(&job);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  return atomic_load(&job.converged);
}
