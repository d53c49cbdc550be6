/*
 * qn_fastport.c — the FAST CPU build of the reference's hot path.  TEST / BASELINE INFRASTRUCTURE ONLY
 * (bench.py's cpu_baseline and --impl reference legs; tests check it against the strict oracle).
 *
 * Same algorithm as oracle/qn_oracle.c (QuasiNewtonMethods.jl:21-27, 34-69, 144-162, 165-232,
 * 237-292; objective test/runtests.jl:4-33), but with the freedoms the reference's own macros take:
 * `@turbo` (LoopVectorization) vectorises over the unit-stride index c of B⁻¹[c, r], keeps ONE
 * accumulator vector per c-lane over the whole r loop, contracts a*b+c into FMAs and reassociates
 * reductions; `@fastmath` licenses the same in the line search.  So: no rule-C2 blocks, no fixed
 * reduction tree, built with -O3 -ffast-math and per-ISA clones (AVX-512 / AVX2+FMA / baseline).
 * Its rounding therefore differs from the strict oracle's — as the real Julia build's does — and it
 * is NEVER the parity checker: tests/test_fastport.py holds it to the north-star tolerances against
 * the strict oracle (first five iterates 1e-12 relative, optimum 1e-8, max|∇| < 1e-8).
 *
 * Why it exists: the strict oracle spends its time in 15-long sequential fma chains
 * (-ffp-contract=off) and ran the README problem 2.6x slower than the reference's published
 * 330.8 µs (README.md:80).  A baseline tuned down flatters the GPU; this is the honest one.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "qn_oracle.h"

#if defined(__x86_64__) && defined(__GNUC__) && !defined(QO_NO_CLONES)
#define QF_HOT __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define QF_HOT
#endif
#define QF_CB 32 /* c-lanes per register block: 4 AVX-512 vectors of accumulators */

/* Rosenbrock, test/runtests.jl:4-33 (value; value + gradient) */
QF_HOT static double qf_rosen(double* restrict g, int n, const double* restrict th) {
  const int N = n >> 1;
  double s = 0.0;
  if (g) {
    for (int i = 0; i < N; ++i) {
      const double t1 = th[i], t2 = th[i + N];
      const double a = t2 - t1 * t1, b = 1.0 - t1;
      s -= 100.0 * a * a + b * b;
      g[i] = 400.0 * a * t1 + 2.0 * b;
      g[i + N] = 200.0 * (t1 * t1 - t2);
    }
  } else {
    for (int i = 0; i < N; ++i) {
      const double t1 = th[i], t2 = th[i + N];
      const double a = t2 - t1 * t1, b = 1.0 - t1;
      s -= 100.0 * a * a + b * b;
    }
  }
  if (n & 1) {
    const double d = 1.0 - th[n - 1];
    s -= d * d;
    if (g) g[n - 1] = 2.0 * d;
  }
  return s;
}

/* BFGS_update!, QN.jl:34-69.  H is column-major with column stride L (a multiple of QF_CB here, the
 * padding rows are zero and stay zero), vectors are padded with zeros to L. */
QF_HOT static double qf_bfgs_update(int n, int L, double* restrict H, double* restrict u,
                                    double* restrict y, const double* restrict s,
                                    const double* restrict g_new, const double* restrict g_old) {
  double sy = 0.0;
  for (int i = 0; i < n; ++i) { /* QN.jl:36-41 */
    const double yi = g_old[i] - g_new[i];
    y[i] = yi;
    sy += s[i] * yi;
  }
  const double rho = 1.0 / sy; /* QN.jl:42 */
  double yHy = 0.0;
  for (int c0 = 0; c0 < L; c0 += QF_CB) { /* QN.jl:44-52 */
    double acc[QF_CB];
    for (int k = 0; k < QF_CB; ++k) acc[k] = 0.0;
    for (int r = 0; r < n; ++r) {
      const double yr = y[r];
      const double* restrict col = H + (size_t)L * r + c0;
      for (int k = 0; k < QF_CB; ++k) acc[k] += yr * col[k];
    }
    for (int k = 0; k < QF_CB; ++k) {
      u[c0 + k] = acc[k] * rho;
      yHy += acc[k] * y[c0 + k];
    }
  }
  const double c1 = (yHy * rho + 1.0) * rho; /* QN.jl:54 */
  double m = 0.0;
  for (int c0 = 0; c0 < L; c0 += QF_CB) { /* QN.jl:56-67 */
    double acc[QF_CB], sc[QF_CB], uc[QF_CB];
    for (int k = 0; k < QF_CB; ++k) { acc[k] = 0.0; sc[k] = s[c0 + k]; uc[k] = u[c0 + k]; }
    for (int r = 0; r < n; ++r) {
      const double c1sr = c1 * s[r], ur = u[r], sr = s[r], gr = g_new[r];
      double* restrict col = H + (size_t)L * r + c0;
      for (int k = 0; k < QF_CB; ++k) {
        const double h = col[k] + c1sr * sc[k] - ur * sc[k] - uc[k] * sr;
        col[k] = h;
        acc[k] += h * gr;
      }
    }
    for (int k = 0; k < QF_CB; ++k) {
      y[c0 + k] = acc[k]; /* the new direction goes into y, QN.jl:65 */
      m += acc[k] * g_new[c0 + k];
    }
  }
  return m;
}

static inline double qf_nanmin(double a, double b) { return a < b ? a : (b != b ? a : b); }
static inline double qf_nanmax(double a, double b) { return a < b ? b : (a != a ? b : a); }
/* -ffast-math assumes no NaN/Inf: classify by bits so the guards of QN.jl:178,187,255 survive */
static inline int qf_finite(double x) {
  uint64_t b;
  memcpy(&b, &x, 8);
  return ((b >> 52) & 0x7ff) != 0x7ff;
}

QF_HOT static double qf_step(int n, double* restrict x_new, const double* restrict x_old,
                             const double* restrict s, double alpha) { /* QN.jl:157-162 */
  for (int i = 0; i < n; ++i) x_new[i] = x_old[i] + alpha * s[i];
  return qf_rosen(NULL, n, x_new);
}

/* linesearch!, QN.jl:165-232 */
static double qf_linesearch(int n, double* x_new, const double* x_old, const double* s, double l0,
                            double m, const qo_backtracking* ls, int64_t* fevals) {
  const double c1 = ls->c1, rho_hi = ls->rho_hi, rho_lo = ls->rho_lo;
  const double sqrttol = 1.0 / (double)(1 << 26);
  double lx0 = l0, a1 = 1.0, a2 = 1.0;
  double lx1 = qf_step(n, x_new, x_old, s, a1);
  ++*fevals;
  int iterfinite = 0;
  while (!qf_finite(lx1) && iterfinite < 52) { /* QN.jl:178-184 */
    ++iterfinite;
    a1 = a2;
    a2 = 0.5 * a2;
    lx1 = qf_step(n, x_new, x_old, s, a2);
    ++*fevals;
  }
  int64_t iteration = 0;
  while (!(qf_finite(lx1) && lx1 >= l0 + a2 * c1 * m)) { /* QN.jl:187 */
    if (++iteration > ls->iterations) return 0.0;       /* QN.jl:193 */
    double at;
    if (ls->order == 2 || iteration == 1) {
      at = -(m * a2 * a2) / (2.0 * (lx1 - l0 - m * a2)); /* QN.jl:205 */
    } else {
      const double denom = 1.0 / (a1 * a1 * a2 * a2 * (a2 - a1)); /* QN.jl:207-209 */
      const double r1 = lx1 - l0 - m * a2, r0 = lx0 - l0 - m * a1;
      const double a = (a1 * a1 * r1 - a2 * a2 * r0) * denom;
      const double b = (-a1 * a1 * a1 * r1 + a2 * a2 * a2 * r0) * denom;
      if (fabs(a) <= 2.220446049250313e-16 + sqrttol * fabs(a)) at = m / (2.0 * b); /* QN.jl:211-212 */
      else at = (sqrt(qf_nanmax(b * b - 3.0 * a * m, 0.0)) + b) / (-3.0 * a);        /* QN.jl:215-218 */
    }
    a1 = a2;
    at = qf_nanmin(at, a2 * rho_hi); /* QN.jl:224 */
    a2 = qf_nanmax(at, a2 * rho_lo); /* QN.jl:225 */
    lx0 = lx1;
    lx1 = qf_step(n, x_new, x_old, s, a2); /* QN.jl:228 */
    ++*fevals;
  }
  return a2;
}

/* optimize!, QN.jl:237-292.  state: 7 L + n L doubles, L a multiple of QF_CB. */
static double qf_optimize_impl(int n, int L, double* state, const double* x0,
                               const qo_backtracking* ls, double tol, int64_t max_iter, int ntrace,
                               double* trace, qo_info* info) {
  double* x_old = state;
  double* g_new = state + (size_t)L;
  double* x_new = state + 2 * (size_t)L;
  double* g_old = state + 3 * (size_t)L;
  double* y = state + 4 * (size_t)L;
  double* s = state + 5 * (size_t)L;
  double* u = state + 6 * (size_t)L;
  double* H = state + 7 * (size_t)L;
  double* const g_canon = g_new;
  memset(state, 0, sizeof(double) * (7 * (size_t)L + (size_t)n * L)); /* zero padding lanes */
  memcpy(x_old, x0, sizeof(double) * (size_t)n);                      /* QN.jl:248 */
  int64_t fevals = 0, iters = 0;
  int status = QO_MAX_ITER;
  double result = NAN;
  for (int64_t it = 1; it <= max_iter; ++it) {
    const double l0 = qf_rosen(g_new, n, x_old); /* QN.jl:254 */
    ++fevals;
    if (!qf_finite(l0)) { status = QO_NONFINITE; break; }
    double gmax = 0.0;
    int bad = 0;
    for (int i = 0; i < n; ++i) {
      bad |= !qf_finite(g_new[i]);
      gmax = fmax(gmax, fabs(g_new[i]));
    }
    if (!bad && gmax < tol) { /* QN.jl:257-261 */
      if (g_new != g_canon) memcpy(g_canon, g_new, sizeof(double) * (size_t)n);
      status = QO_CONVERGED;
      result = l0;
      break;
    }
    double m;
    if (it == 1) {
      m = -1.0;
    } else {
      m = qf_bfgs_update(n, L, H, u, y, s, g_new, g_old); /* QN.jl:267 */
      double* tmp = s; s = y; y = tmp;                    /* QN.jl:268 */
    }
    if (!(m > 0.0)) { /* QN.jl:272-280 (m <= 0; a NaN m also resets nothing in the reference — it fails the line search) */
      if (m != m) { /* NaN: the reference goes on with the NaN direction and ends in the line search */
      } else {
        for (int r = 0; r < n; ++r) {
          memset(H + (size_t)L * r, 0, sizeof(double) * (size_t)L);
          H[r + (size_t)L * r] = 1.0;
        }
        m = 0.0;
        for (int i = 0; i < n; ++i) { s[i] = g_new[i]; m += g_new[i] * g_new[i]; }
      }
    }
    int64_t fe = 0;
    const double alpha = qf_linesearch(n, x_new, x_old, s, l0, m, ls, &fe); /* QN.jl:283 */
    fevals += fe;
    if (alpha == 0.0) { status = QO_LINESEARCH_FAILED; break; }
    for (int i = 0; i < n; ++i) { /* update_state!, QN.jl:21-27 */
      const double si = alpha * s[i];
      s[i] = si;
      x_old[i] += si;
    }
    ++iters;
    if (trace && it <= ntrace) memcpy(trace + (size_t)(it - 1) * n, x_old, sizeof(double) * (size_t)n);
    double* tmp = g_old; g_old = g_new; g_new = tmp; /* QN.jl:286 */
  }
  if (info) { info->iters = iters; info->fevals = fevals; info->gevals = 0; info->status = status; }
  return result;
}

static int qf_ld(int n) { return (n + QF_CB - 1) / QF_CB * QF_CB; }

/* single problem; x_out/g_out: n doubles each (may be NULL) */
double qf_optimize(int n, const double* x0, const qo_backtracking* ls, double tol, int64_t max_iter,
                   double* x_out, double* g_out, int ntrace, double* trace, qo_info* info) {
  const int L = qf_ld(n);
  double* state = (double*)aligned_alloc(64, sizeof(double) * (7 * (size_t)L + (size_t)n * L));
  const double r = qf_optimize_impl(n, L, state, x0, ls, tol, max_iter, ntrace, trace, info);
  if (x_out) memcpy(x_out, state, sizeof(double) * (size_t)n);
  if (g_out) memcpy(g_out, state + L, sizeof(double) * (size_t)n);
  free(state);
  return r;
}

/* repeated single-problem timing without the ctypes call overhead: median-able per-call seconds */
void qf_time_single(int n, const double* x0, const qo_backtracking* ls, double tol, int reps,
                    double* seconds) {
  const int L = qf_ld(n);
  double* state = (double*)aligned_alloc(64, sizeof(double) * (7 * (size_t)L + (size_t)n * L));
  for (int k = 0; k < reps; ++k) {
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    qf_optimize_impl(n, L, state, x0, ls, tol, 10000, 0, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    seconds[k] = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
  }
  free(state);
}

typedef struct {
  int n;
  int64_t batch;
  const double* theta0;
  const qo_backtracking* ls;
  double tol;
  int64_t max_iter;
  double *x_out, *g_out, *ell_out;
  int64_t *iters_out, *fevals_out;
  int32_t* status_out;
  int ntrace;
  double* trace;
  atomic_long next, converged;
} qf_job;

static void* qf_worker(void* arg) { /* the Threads.@threads analogue: one private state per thread */
  qf_job* job = (qf_job*)arg;
  const int n = job->n, L = qf_ld(n);
  double* state = (double*)aligned_alloc(64, sizeof(double) * (7 * (size_t)L + (size_t)n * L));
  for (;;) {
    const long j0 = atomic_fetch_add(&job->next, 16);
    if (j0 >= job->batch) break;
    const long j1 = (j0 + 16 < job->batch) ? j0 + 16 : job->batch;
    for (long j = j0; j < j1; ++j) {
      qo_info info;
      double* tr = job->trace ? job->trace + (size_t)j * job->ntrace * n : NULL;
      const double ell = qf_optimize_impl(n, L, state, job->theta0 + (size_t)j * n, job->ls, job->tol,
                                          job->max_iter, job->ntrace, tr, &info);
      if (job->x_out) memcpy(job->x_out + (size_t)j * n, state, sizeof(double) * (size_t)n);
      if (job->g_out) memcpy(job->g_out + (size_t)j * n, state + L, sizeof(double) * (size_t)n);
      if (job->ell_out) job->ell_out[j] = ell;
      if (job->iters_out) job->iters_out[j] = info.iters;
      if (job->fevals_out) job->fevals_out[j] = info.fevals;
      if (job->status_out) job->status_out[j] = info.status;
      if (info.status == QO_CONVERGED) atomic_fetch_add(&job->converged, 1);
    }
  }
  free(state);
  return NULL;
}

int64_t qf_optimize_batch(int n, int64_t batch, const double* theta0, const qo_backtracking* ls,
                          double tol, int64_t max_iter, double* x_out, double* g_out, double* ell_out,
                          int64_t* iters_out, int64_t* fevals_out, int32_t* status_out, int ntrace,
                          double* trace, int nthreads) {
  qf_job job;
  job.n = n; job.batch = batch; job.theta0 = theta0; job.ls = ls; job.tol = tol; job.max_iter = max_iter;
  job.x_out = x_out; job.g_out = g_out; job.ell_out = ell_out; job.iters_out = iters_out;
  job.fevals_out = fevals_out; job.status_out = status_out; job.ntrace = ntrace; job.trace = trace;
  atomic_init(&job.next, 0);
  atomic_init(&job.converged, 0);
  if (nthreads <= 0) {
    cpu_set_t set;
    nthreads = (sched_getaffinity(0, sizeof set, &set) == 0) ? CPU_COUNT(&set) : 1;
  }
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < nthreads; ++t)
    if (pthread_create(&th[started], NULL, qf_worker, &job) == 0) ++started;
  if (started == 0) qf_worker(&job);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  return atomic_load(&job.converged);
}
