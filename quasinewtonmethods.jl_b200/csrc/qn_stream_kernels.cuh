// qn_stream_kernels.cuh — device side of the large-n streaming path: the sweeps over B⁻¹ in HBM and
// the O(n) vector kernels between them.  Host side (graphs, row-sharded setup, C entry points):
// qn_stream.cu, which also carries the description of the path.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "qn_device.cuh"
#include "qn_internal.h"

namespace qn {

constexpr int kVecThreads = 1024;
constexpr int kPassThreads = 256;             // most threads of a sweep CTA (2 rows each)
constexpr int kShardRows = 2 * kPassThreads;  // the row-sharded mode splits B⁻¹ in units of 512 rows

// B⁻¹ is streamed: every element is read once and written at most once per sweep; the cache
// hint of those accesses is a kernel template parameter (0: default, 1: .cs evict-first,
// 2: .lu / .cg, 3: .cv / .wt).  scripts/exp_stream_hints.sh measures the combinations.
template <int HINT>
__device__ __forceinline__ double2 ld_h(const double* ptr) {
  if (HINT == 1) return __ldcs(reinterpret_cast<const double2*>(ptr));
  if (HINT == 2) return __ldlu(reinterpret_cast<const double2*>(ptr));
  if (HINT == 3) return __ldcv(reinterpret_cast<const double2*>(ptr));
  if (HINT == 4) return __ldcg(reinterpret_cast<const double2*>(ptr));
  return *reinterpret_cast<const double2*>(ptr);
}
template <int HINT>
__device__ __forceinline__ void st_h(double* ptr, const double2 v) {
  if (HINT == 1) __stcs(reinterpret_cast<double2*>(ptr), v);
  else if (HINT == 2) __stcg(reinterpret_cast<double2*>(ptr), v);
  else if (HINT == 3) __stwt(reinterpret_cast<double2*>(ptr), v);
  else *reinterpret_cast<double2*>(ptr) = v;
}

// The leading `persist_sb` super-blocks of B⁻¹ (a contiguous run of columns) are kept RESIDENT IN L2
// across iterations (an access-policy window with the persisting property, qn_stream.cu): their
// accesses carry no eviction hint, the rest of the matrix streams with the kernel's hints.
template <int HINT>
__device__ __forceinline__ double2 ld_hk(const double* ptr, const bool keep) {
  return keep ? *reinterpret_cast<const double2*>(ptr) : ld_h<HINT>(ptr);
}
template <int HINT>
__device__ __forceinline__ void st_hk(double* ptr, const double2 v, const bool keep) {
  if (keep) *reinterpret_cast<double2*>(ptr) = v; else st_h<HINT>(ptr, v);
}

struct StreamScalars {
  double l0, sy, rho, c1, m, result;
  int it;          // completed outer iterations
  int iters, fevals, status;
  int done;        // 1: the remaining kernels are no-ops
  int skip_update; // iteration 1: m = -1, no BFGS update (QN.jl:263-264)
  int need_reset;  // vec_end -> identity
  int exit_it;     // iteration in which the solve ended (selects which ∇ buffer is canonical)
  int pending;     // one-pass mode: the rank-2 update (ps, pu, pc1) is not yet applied to H
  double pc1;
  int spec;        // cluster vector kernel: evaluate the lower-clamp chain of step sizes speculatively
};

struct StreamParams {
  int n, ld, rb, sbcols, nsb, max_iter;
  // columns of B⁻¹ one sweep CTA covers, and how many such groups a row has: a SUPER-block (16 rule-C2
  // blocks, ngrp = nsb) when that fills the GPU, ONE block (ngrp = blocks of a row) when the matrix —
  // or this GPU's slab of it — is small: 16× the CTAs, each with all its loads in flight at once
  // instead of a walk of 16 dependent batches.  The partial sums are per group either way, and
  // pass_tail adds them in the canonical two-level order, so the bits do not depend on the choice.
  int grp_cols, ngrp;
  double tol;
  LsParams ls;
  double* H;
  double *x, *g, *gp, *s, *y, *nu, *w, *xs, *tmp, *dvec; // length-n vectors
  double *ps, *pu, *tvec, *wvec;                          // one-pass mode: pending s, u; t = Hy, w = H∇
  double* tmp2;                                           // pass2's terms of dᵀ∇ (tmp holds pass1's)
  int nvstride;                                           // distance between consecutive work vectors
  double* partial2;                                        // one-pass mode: second [nsb][n] array
  int onepass;
  double* partial;                                  // [nsb][n]
  unsigned int* tile_count;                         // one arrival counter per row tile
  int stage_in_smem;                                // vec kernels stage reductions in shared memory
  StreamScalars* sc;
  int ntrace;
  double* trace; // [ntrace][n]
  // ---- row-sharded mode: this GPU holds rows [row0, row0 + 512·ntiles) of
  // B⁻¹ (H points at that slab, column stride ld) and every vector replicated.  A sweep's row
  // sums are stored into EVERY replica through peer pointers (same layout on each GPU:
  // peer address = peer_base[g] + (address - base)), then the shard raises its flag on every
  // peer; the vector kernel that consumes them spins on its LOCAL flags.  nshard = 1: one GPU.
  int nshard, shard;
  int row0, rows;   // this GPU's slab: rows [row0, row0 + rows) of B⁻¹ (all of them when nshard = 1)
  int ntiles;       // CTA row tiles of a sweep on this GPU: 2·blockDim rows each
  // the objective (QN.jl:8-9): kind + device parameters.  Rosenbrock needs none; dense Gaussian:
  // obj_a = μ (n), obj_b = Λ (n×n column-major).  The cluster vector kernel and the one-pass mode
  // are Rosenbrock's; any other registered objective runs the single-CTA vector kernels.
  int obj_kind;
  const double* obj_a;
  const double* obj_b;
  int persist_sb; // leading super-blocks of B⁻¹ inside the L2-persisting window (0: none)
  int fused2; // two-pass mode in three kernels: pass1 (+ yᵀB⁻¹y, c₁), pass2 (c₁s on the fly), cluster kernel
  double* base;
  double* peer_base[8];
  unsigned int* tiles_done;  // local: row tiles of the current sweep that have been broadcast
  unsigned int* flags;       // local: flags[2·g + w] = last epoch shard g finished sweep w
};

constexpr int kMaxShards = 8;

// store v at `ptr` (an address inside this GPU's work block) in every replica
__device__ __forceinline__ void bstore(const StreamParams& p, double* ptr, double v) {
  if (p.nshard == 1) { *ptr = v; return; }
  const ptrdiff_t off = ptr - p.base;
  for (int g = 0; g < p.nshard; ++g) p.peer_base[g][off] = v;
}
// after a CTA has broadcast the sums of one row tile: count it, and when it was this shard's
// last tile publish the epoch on every peer (which = 0: pass1 / the one-pass sweep, 1: pass2).
// Returns (to every thread of the CTA) whether this CTA completed the shard's last tile.
__device__ __forceinline__ bool shard_signal(const StreamParams& p, int which, unsigned epoch, bool count_always = false) {
  if (p.nshard == 1 && !count_always) return false;
  __shared__ int last_tile;
  __threadfence_system(); // this thread's (peer) stores are visible system-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(p.tiles_done, 1u);
    last_tile = (done == (unsigned)p.ntiles - 1);
    if (last_tile) {
      *p.tiles_done = 0;
      __threadfence_system();
      if (p.nshard > 1) {
        const ptrdiff_t off = reinterpret_cast<double*>(p.flags) - p.base;
        for (int g = 0; g < p.nshard; ++g) {
          volatile unsigned* f = reinterpret_cast<volatile unsigned*>(p.peer_base[g] + off);
          f[2 * p.shard + which] = epoch;
        }
      }
    }
  }
  __syncthreads();
  return last_tile != 0;
}
// the consumer side: wait until every shard has published `epoch` for sweep `which`
__device__ __forceinline__ void shard_wait(const StreamParams& p, int which, unsigned epoch) {
  if (p.nshard == 1) return;
  if ((int)threadIdx.x < p.nshard) {
    const volatile unsigned* f = p.flags + 2 * threadIdx.x + which;
    while (*f < epoch) { }
  }
  __threadfence_system();
  __syncthreads();
}


// Element-wise pass over n elements by one CTA with 8 independent loads in flight per thread
// (a single SM sees ~700 cycles of L2 latency per dependent round trip): ld(i) loads/computes,
// st(i, v) stores.
template <class L, class S>
__device__ __forceinline__ void cta_map(int n, L&& ld, S&& st) {
  const int T = blockDim.x;
  int i = threadIdx.x;
  for (; i + 7 * T < n; i += 8 * T) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ld(i + k * T);
#pragma unroll
    for (int k = 0; k < 8; ++k) st(i + k * T, v[k]);
  }
  for (; i < n; i += T) st(i, ld(i));
}

// ---- canonical csum (rule C1) by one CTA.  `stage` is an n-double scratch array (shared
// memory when it fits, else global); `red` is one shared double.  Every thread calls it.
__device__ __forceinline__ double cta_csum_staged(int n, const double* stage, double* red) {
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    double p = (l < n) ? stage[l] : 0.0;
    int i = l + 32;
    for (; i + 224 < n; i += 256) { // 8 loads in flight; the adds stay in ascending order
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = stage[i + 32 * k];
#pragma unroll
      for (int k = 0; k < 8; ++k) p = p + v[k];
    }
    for (; i < n; i += 32) p = p + stage[i];
    p = warp_butterfly(p);
    if (l == 0) *red = p;
  }
  __syncthreads();
  const double r = *red;
  __syncthreads(); // stage and red may be reused right away
  return r;
}
template <class F>
__device__ __forceinline__ double cta_csum(int n, double* stage, double* red, F&& v) {
  // four independent elements per thread per round: the L2 loads inside v() overlap
  const int T = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * T < n; i += 4 * T) {
    const double a = v(i), b = v(i + T), c = v(i + 2 * T), d = v(i + 3 * T);
    stage[i] = a; stage[i + T] = b; stage[i + 2 * T] = c; stage[i + 3 * T] = d;
  }
  for (; i < n; i += T) stage[i] = v(i);
  __syncthreads();
  return cta_csum_staged(n, stage, red);
}

// Reduce K staged arrays (array k has cnt[k] elements at stage + k*ns) by rule C1, warp k on
// array k, so the K sequential chains run side by side.  Contains the barriers around it.
template <int K>
__device__ __forceinline__ void cta_reduce_multi(const int (&cnt)[K], const double* stage, int ns, double* red,
                                                 double (&out)[K]) {
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (w < K) {
    const int n = cnt[w];
    const double* st = stage + w * ns;
    double p = (l < n) ? st[l] : 0.0;
    int j = l + 32;
    for (; j + 224 < n; j += 256) {
      double x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = st[j + 32 * k];
#pragma unroll
      for (int k = 0; k < 8; ++k) p = p + x[k];
    }
    for (; j < n; j += 32) p = p + st[j];
    p = warp_butterfly(p);
    if (l == 0) red[w] = p;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = red[k];
  __syncthreads();
}

// K independent canonical csums at once: v(i, out[K]) produces the K values of element i;
// warp k reduces array k, so the K sequential chains run side by side.  stage: K*ns doubles.
template <int K, class F>
__device__ __forceinline__ void cta_csum_multi(int n, double* stage, int ns, double* red, F&& v, double (&out)[K]) {
  const int T = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * T < n; i += 4 * T) { // four independent elements per thread per round
    double a[4][K];
#pragma unroll
    for (int q = 0; q < 4; ++q) v(i + q * T, a[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < K; ++k) stage[k * ns + i + q * T] = a[q][k];
  }
  for (; i < n; i += T) {
    double a[K];
    v(i, a);
#pragma unroll
    for (int k = 0; k < K; ++k) stage[k * ns + i] = a[k];
  }
  int cnt[K];
#pragma unroll
  for (int k = 0; k < K; ++k) cnt[k] = n;
  cta_reduce_multi<K>(cnt, stage, ns, red, out);
}

// Rosenbrock (test/runtests.jl:4-33) by one CTA at the point X(i); writes ∇ to g when
// g != nullptr.  X is a functor so a line-search trial can evaluate at x + a·s without
// materialising the point (step!, QN.jl:157-162).
template <class XF>
__device__ __forceinline__ double cta_rosenbrock(int n, XF&& X, double* g, double* stage, double* red) {
  const int N = n >> 1;
  const double sum = cta_csum(N, stage, red, [&](int i) {
    const double t1 = X(i), t2 = X(i + N);
    const double a = t2 - t1 * t1;
    const double b = 1.0 - t1;
    if (g) {
      g[i] = (400.0 * a) * t1 + 2.0 * b;
      g[i + N] = 200.0 * (t1 * t1 - t2);
    }
    return 100.0 * (a * a) + b * b;
  });
  double s = 0.0 - sum;
  if (n & 1) {
    const double d = 1.0 - X(n - 1);
    s = -fma(d, d, -s);
    if (g && threadIdx.x == 0) g[n - 1] = 2.0 * d;
  }
  return s;
}

// The objective protocol for the single-CTA vector kernels: eval(p, X, g, stage, red) returns ℓ at
// the point X(i) and writes ∇ to g when g != nullptr (logdensity / ∂logdensity!, QN.jl:8-9).
struct SRosenbrock {
  template <class XF>
  __device__ __forceinline__ static double eval(const StreamParams& p, XF&& X, double* g, double* stage, double* red) {
    return cta_rosenbrock(p.n, X, g, stage, red);
  }
};
// Dense Gaussian ℓ = -½ (θ-μ)ᵀΛ(θ-μ), ∇ = -Λ(θ-μ), CTA-cooperative: d = θ-μ to
// p.xs, z = Λd by rule-C2 row chains (thread c: rows c, c + T, …; a column of Λ is read by
// consecutive threads) to p.w — dead between pass2 and the next vec_mid, which is when the
// objective runs —, ℓ = -½ csum(d·z).  Same order as the oracle's gauss_eval: the same bits.
struct SGaussian {
  template <class XF>
  __device__ __forceinline__ static double eval(const StreamParams& p, XF&& X, double* g, double* stage, double* red) {
    const int n = p.n, T = blockDim.x;
    const double* mu = p.obj_a;
    const double* lam = p.obj_b;
    double* d = p.xs;
    double* z = p.w;
    for (int i = threadIdx.x; i < n; i += T) d[i] = X(i) - mu[i];
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += T) {
      const double zc = chain_rowsum(n, [&](int r, double& a, double& b) { a = lam[c + (size_t)n * r]; b = d[r]; });
      z[c] = zc;
      if (g) g[c] = -zc;
    }
    __syncthreads();
    const double q = cta_csum(n, stage, red, [&](int i) { return d[i] * z[i]; });
    return -0.5 * q;
  }
};

template <class SObj>
__global__ void __launch_bounds__(kVecThreads) stream_vec_begin(const StreamParams p) {
  extern __shared__ __align__(16) double vsm[];
  __shared__ double red;
  __shared__ int flag;
  StreamScalars* sc = p.sc;
  if (sc->done) return;
  const int n = p.n;
  const int it = sc->it + 1;
  double* stage = p.stage_in_smem ? vsm : p.tmp;
  const double* xv = p.x;
  const double l0 = SObj::eval(p, [&](int i) { return xv[i]; }, p.g, stage, &red); // QN.jl:254
  __syncthreads();
  if (threadIdx.x == 0) { sc->l0 = l0; sc->fevals += 1; flag = 1; }
  if (!isfinite(l0)) { // QN.jl:255
    if (threadIdx.x == 0) { sc->status = QN_NONFINITE; sc->done = 1; sc->exit_it = it; }
    return;
  }
  __syncthreads();
  bool small = true; // maximum(abs, ∇) < tol, QN.jl:257
  cta_map(n, [&](int i) { return p.g[i]; }, [&](int, double gi) { small = small && (fabs(gi) < p.tol); });
  if (!small) flag = 0;
  __syncthreads();
  if (flag) {
    if (threadIdx.x == 0) { sc->status = QN_CONVERGED; sc->result = l0; sc->done = 1; sc->exit_it = it; }
    return;
  }
  if (it == 1) { // QN.jl:263-264
    if (threadIdx.x == 0) { sc->m = -1.0; sc->skip_update = 1; }
    return;
  }
  // y = ∇_old - ∇_new ; sᵀy   QN.jl:36-41
  const double sy = cta_csum(n, stage, &red, [&](int i) {
    const double yi = p.gp[i] - p.g[i];
    p.y[i] = yi;
    return p.s[i] * yi;
  });
  if (threadIdx.x == 0) { sc->sy = sy; sc->rho = 1.0 / sy; sc->skip_update = 0; } // QN.jl:42
}

// Tail of a pass: the LAST of the nsb CTAs of a row tile to finish adds that tile's
// super-block partials in ascending order (rule C2) — 16 SMs do this in parallel instead of
// one.  finish(c, T) consumes the row sum.  Returns after the tile is done.
template <class F>
__device__ __forceinline__ bool pass_tail(const StreamParams& p, int which, unsigned epoch, F&& finish,
                                          bool count_always = false) {
  __shared__ int is_last;
  __threadfence(); // this CTA's partials are visible device-wide before it signs off
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(p.tile_count + blockIdx.x, 1u);
    is_last = (prev == (unsigned)p.ngrp - 1);
    if (is_last) p.tile_count[blockIdx.x] = 0; // ready for the next pass
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  if (c0 < p.n) {
    const bool two = (c0 + 1 < p.n);
    const volatile double* part = p.partial + c0; // written by other CTAs of this launch
    double T0, T1;
    if (p.ngrp == p.nsb) { // one partial per super-block: added in ascending order (rule C2, second level)
      T0 = part[0]; T1 = two ? part[1] : 0.0;
      for (int j = 1; j < p.nsb; ++j) {
        T0 = T0 + part[(size_t)j * p.n];
        if (two) T1 = T1 + part[(size_t)j * p.n + 1];
      }
    } else { // one partial per block: blocks in ascending order inside a super-block of 16, then the super-blocks
      T0 = 0.0; T1 = 0.0;
      for (int g0 = 0, sbi = 0; g0 < p.ngrp; g0 += 16, ++sbi) {
        const int g1 = min(g0 + 16, p.ngrp);
        double v0[16], v1[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { // the loads of a super-block are in flight together
          const bool ok = g0 + k < g1;
          v0[k] = ok ? part[(size_t)(g0 + k) * p.n] : 0.0;
          v1[k] = (ok && two) ? part[(size_t)(g0 + k) * p.n + 1] : 0.0;
        }
        double S0 = v0[0], S1 = v1[0];
#pragma unroll
        for (int k = 1; k < 16; ++k)
          if (g0 + k < g1) { S0 = S0 + v0[k]; S1 = S1 + v1[k]; }
        T0 = (sbi == 0) ? S0 : T0 + S0;
        T1 = (sbi == 0) ? S1 : T1 + S1;
      }
    }
    finish(c0, T0);
    if (two) finish(c0 + 1, T1);
  }
  return shard_signal(p, which, epoch, count_always); // row-sharded mode: this tile's sums are in every replica
}

__global__ void __launch_bounds__(kVecThreads) stream_vec_mid(const StreamParams p) {
  extern __shared__ __align__(16) double vsm[];
  __shared__ double red;
  StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  const int n = p.n;
  const double rho = sc->rho;
  shard_wait(p, 0, (unsigned)sc->it + 1u); // row-sharded mode: every shard's slice of t has arrived
  // pass1's tail left t[c]·y[c] in tmp and -u in nu; yᵀB⁻¹y = csum, QN.jl:51
  double* stage = p.stage_in_smem ? vsm : p.tmp;
  if (p.stage_in_smem) {
    cta_map(n, [&](int i) { return p.tmp[i]; }, [&](int i, double v) { stage[i] = v; });
    __syncthreads();
  }
  const double yHy = cta_csum_staged(n, stage, &red);
  const double c1 = fma(yHy, rho, 1.0) * rho; // QN.jl:54
  cta_map(n, [&](int c) { return p.s[c]; }, [&](int c, double sc_) { p.w[c] = c1 * sc_; });
  if (threadIdx.x == 0) sc->c1 = c1;
}

// Second half of an iteration, shared by the two-pass and the one-pass mode: the reset
// decision (QN.jl:272-280), the whole line search (QN.jl:165-232), update_state! (QN.jl:21-27)
// and the bookkeeping.  Runs in one CTA; s holds the direction, m = dᵀ∇.
template <class SObj = SRosenbrock>
__device__ __forceinline__ void stream_finish_iteration(const StreamParams& p, StreamScalars* sc, const int it,
                                                        const double l0, double m, double* stage, double* redp) {
  const int n = p.n;
  double& red = *redp;
  int need_reset = 0;
  if (m <= 0.0) { // QN.jl:272-280
    need_reset = 1;
    if (threadIdx.x == 0) sc->pending = 0; // one-pass mode: the reset discards the pending update
    m = cta_csum(n, stage, &red, [&](int i) {
      const double gi = p.g[i];
      p.s[i] = gi;
      return gi * gi;
    });
  }
  __syncthreads();
  // ---- linesearch!  QN.jl:165-232
  int fevals = 0;
  const double* xv = p.x;
  const double* sv = p.s;
  auto trial = [&](double a) { // step! (QN.jl:157-162): x_new[i] = x_old[i] + a s[i] as one fma
    ++fevals;
    return SObj::eval(p, [&](int i) { return fma(a, sv[i], xv[i]); }, nullptr, stage, &red);
  };
  double lx_last;
  const double alpha = backtracking_linesearch(p.ls, l0, m, trial, lx_last);
  if (alpha == 0.0) { // QN.jl:284
    if (threadIdx.x == 0) { sc->status = QN_LINESEARCH_FAILED; sc->done = 1; sc->fevals += fevals; sc->need_reset = need_reset; sc->exit_it = it; }
    return;
  }
  // update_state!  QN.jl:21-27 ; ∇_old ← ∇_new  QN.jl:286
  cta_map(n, [&](int i) { return alpha * p.s[i]; }, [&](int i, double si) { p.s[i] = si; });
  cta_map(n, [&](int i) { return p.x[i] + p.s[i]; }, [&](int i, double xi) {
    p.x[i] = xi;
    if (p.trace != nullptr && it <= p.ntrace) p.trace[(size_t)(it - 1) * n + i] = xi;
  });
  if (it < p.max_iter) cta_map(n, [&](int i) { return p.g[i]; }, [&](int i, double gi) { p.gp[i] = gi; });
  if (threadIdx.x == 0) {
    sc->m = m;
    sc->it = it;
    sc->iters += 1;
    sc->fevals += fevals;
    sc->need_reset = need_reset;
    if (it >= p.max_iter) { sc->status = QN_MAX_ITER; sc->done = 1; sc->exit_it = it; }
  }
}


template <class SObj>
__global__ void __launch_bounds__(kVecThreads) stream_vec_end(const StreamParams p) {
  extern __shared__ __align__(16) double vsm[];
  __shared__ double red;
  StreamScalars* sc = p.sc;
  if (sc->done) return;
  const int n = p.n;
  const int it = sc->it + 1;
  const double l0 = sc->l0;
  double* stage = p.stage_in_smem ? vsm : p.tmp;            // n doubles
  double m;
  if (!sc->skip_update) {
    shard_wait(p, 1, (unsigned)it); // row-sharded mode: every shard's slice of d has arrived
    // pass2's tail left d in dvec and d[c]·∇[c] in tmp2; m = dᵀ∇_new = csum, QN.jl:66-68
    if (p.stage_in_smem) {
      cta_map(n, [&](int i) { return p.tmp2[i]; }, [&](int i, double v) { stage[i] = v; });
      __syncthreads();
    }
    m = cta_csum_staged(n, p.stage_in_smem ? stage : p.tmp2, &red);
    cta_map(n, [&](int i) { return p.dvec[i]; }, [&](int i, double v) { p.s[i] = v; }); // `s, y = y, s`, QN.jl:65,268
    __syncthreads();
  } else {
    m = sc->m;
  }
  stream_finish_iteration<SObj>(p, sc, it, l0, m, stage, &red);
}

#ifndef QN_PASS2_MINB
#define QN_PASS2_MINB 2
#endif
__device__ __forceinline__ void pass1_finish(const StreamParams& p, const StreamScalars* sc, const int sb,
                                             const int c0, const bool two, const double S0, const double S1);
__device__ __forceinline__ void pass2_finish(const StreamParams& p, const StreamScalars* sc, const int sb,
                                             const int c0, const bool two, const double S0, const double S1);
#ifndef QN_PASS1_MINB
#define QN_PASS1_MINB 4
#endif
// ---- pass 1: partial[sb][c] = Σ_{blocks of sb} chain_b(c),  chain_b(c) = Σ_{r in b} y[r] H[c,r]
// DESC: the rule-C2 blocks of the super-block are walked from the LAST to the first (the chain inside a
// block still runs in ascending column order, and the block partials are still added in ascending
// order — they wait in registers), so that the sweep starts where the previous sweep over B⁻¹ ended:
// what that sweep left in the 126 MB L2 is read before it is evicted.
template <int LDH, int STH, int DEPTH = 8, bool DESC = false>
__global__ void __launch_bounds__(kPassThreads, DEPTH > 8 ? 1 : (DESC ? 2 : QN_PASS1_MINB)) stream_pass1(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(16) double sm[];
  double* ys = sm; // y of this super-block's columns
  const int n = p.n, ld = p.ld, rb = p.rb;
  const int sb = blockIdx.y;
  const int r_begin = sb * p.grp_cols;
  const int r_end = min(r_begin + p.grp_cols, n);
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) ys[r - r_begin] = p.y[r];
  __syncthreads();
  const bool reset = p.fused2 && sc->need_reset != 0; // B⁻¹ ← I (QN.jl:144-148) folded into the sweeps
  const bool keep = r_begin / p.sbcols < p.persist_sb; // this super-block stays in L2
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  const bool two = (c0 + 1 < n);
  const double* hp = p.H + (c0 - p.row0);
  double S0 = 0.0, S1 = 0.0;
  auto block_chain2 = [&](const int b0, double& t0, double& t1) { // one rule-C2 block: ascending fma chains of both rows
    const int b1 = min(b0 + rb, r_end);
    t0 = 0.0; t1 = 0.0;
    int r = b0;
    for (; r + DEPTH <= b1; r += DEPTH) { // 8 independent 128-bit loads in flight
      double2 h[DEPTH];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k)
        h[k] = reset ? make_double2(c0 == r + k ? 1.0 : 0.0, c0 + 1 == r + k ? 1.0 : 0.0)
                     : ld_hk<LDH>(hp + (size_t)ld * (r + k), keep);
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        const double yr = ys[r + k - r_begin];
        t0 = fma(yr, h[k].x, t0);
        t1 = fma(yr, h[k].y, t1);
      }
    }
    for (; r < b1; ++r) {
      const double2 h = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0)
                              : ld_hk<LDH>(hp + (size_t)ld * r, keep);
      const double yr = ys[r - r_begin];
      t0 = fma(yr, h.x, t0);
      t1 = fma(yr, h.y, t1);
    }
  };
  if constexpr (DESC) {
    double T0[16], T1[16]; // a super-block has at most 16 blocks (sbcols = 16 rb)
#pragma unroll
    for (int bi = 15; bi >= 0; --bi) {
      T0[bi] = 0.0; T1[bi] = 0.0;
      if (r_begin + bi * rb < r_end && c0 < n) block_chain2(r_begin + bi * rb, T0[bi], T1[bi]);
    }
    S0 = T0[0]; S1 = T1[0];
#pragma unroll
    for (int bi = 1; bi < 16; ++bi)
      if (r_begin + bi * rb < r_end) { S0 = S0 + T0[bi]; S1 = S1 + T1[bi]; }
  } else {
    bool first = true;
    for (int b0 = r_begin; b0 < r_end && c0 < n; b0 += rb) {
      double t0, t1;
      block_chain2(b0, t0, t1);
      S0 = first ? t0 : S0 + t0;
      S1 = first ? t1 : S1 + t1;
      first = false;
    }
  }
  pass1_finish(p, sc, sb, c0, two, S0, S1);
}

// End of pass 1 for one CTA: its super-block partials, then (last CTA of the row tile) the row sums.
__device__ __forceinline__ void pass1_finish(const StreamParams& p, const StreamScalars* sc, const int sb,
                                             const int c0, const bool two, const double S0, const double S1) {
  const int n = p.n;
  if (c0 < n) {
    double* out = p.partial + (size_t)sb * n + c0;
    out[0] = S0;
    if (two) out[1] = S1;
  }
  const double rho = sc->rho;
  const unsigned epoch = (unsigned)sc->it + 1u;
  const bool last_tile = pass_tail(p, 0, epoch, [&](int c, double T) {
    bstore(p, p.nu + c, -(T * rho));     // -B⁻¹y[c] = -(t ρ), QN.jl:50
    bstore(p, p.tmp + c, T * p.y[c]);    // term of yᵀB⁻¹y, QN.jl:51
  }, p.fused2 != 0);
  if (!p.fused2 || !last_tile) return;
  // fused two-pass mode: the CTA that completed this GPU's last row tile stands in for vec_mid —
  // once every shard's slice has arrived, yᵀB⁻¹y = csum of the terms (rule C1, one warp), then c₁
  shard_wait(p, 0, epoch);
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    const volatile double* tv = p.tmp; // written by other CTAs (and peers) during this launch
    double acc = (l < n) ? tv[l] : 0.0;
    int j = l + 32;
    for (; j + 224 < n; j += 256) {
      double x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = tv[j + 32 * k];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = acc + x[k];
    }
    for (; j < n; j += 32) acc = acc + tv[j];
    const double yHy = warp_butterfly(acc);
    if (l == 0) p.sc->c1 = fma(yHy, rho, 1.0) * rho; // QN.jl:54
  }
}

// ---- pass 2: H[c,r] ← ((H + (c₁s[r])s[c]) - u[r]s[c]) - u[c]s[r] ; partials of d = H⁺∇
// Residency matters for these sweeps, and FEWER resident CTAs are better: each CTA walks its own
// 4 KB-per-column stream, and 2 CTAs per SM (296 concurrent streams, all 128 registers spent on loads
// in flight) beat 3 or 4 (measured at n = 8192, profiles/r1_stream_residency.txt: pass2 172 µs at 2 CTAs/SM
// against 207 µs at 3-4; the one-pass sweep 201 / 245 / 253 µs per iteration at 2 / 3 / 4).
template <int LDH, int STH, int DEPTH = 8, bool DESC = false>
__global__ void __launch_bounds__(kPassThreads, DEPTH > 8 ? 1 : QN_PASS2_MINB) stream_pass2(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(16) double sm[];
  double4* ops = reinterpret_cast<double4*>(sm); // {c₁s[r], -u[r], s[r], ∇[r]} per column
  const int n = p.n, ld = p.ld, rb = p.rb;
  const int sb = blockIdx.y;
  const int r_begin = sb * p.grp_cols;
  const int r_end = min(r_begin + p.grp_cols, n);
  const bool reset = p.fused2 && sc->need_reset != 0;
  const bool keep = r_begin / p.sbcols < p.persist_sb; // this super-block stays in L2
  const double c1 = sc->c1;
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
    const double sr = p.s[r];
    ops[r - r_begin] = make_double4(p.fused2 ? c1 * sr : p.w[r], p.nu[r], sr, p.g[r]); // (c₁ s[r]) once per r (rule C4)
  }
  __syncthreads();
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  const bool two = (c0 + 1 < n);
  const double s0 = (c0 < n) ? p.s[c0] : 0.0, nu0 = (c0 < n) ? p.nu[c0] : 0.0;
  const double s1 = two ? p.s[c0 + 1] : 0.0, nu1 = two ? p.nu[c0 + 1] : 0.0;
  double* hp = p.H + (c0 - p.row0);
  double S0 = 0.0, S1 = 0.0;
  auto block_update2 = [&](const int b0, double& a0, double& a1) { // one rule-C2 block, ascending columns
    const int b1 = min(b0 + rb, r_end);
    a0 = 0.0; a1 = 0.0;
    int r = b0;
    for (; r + DEPTH <= b1; r += DEPTH) {
      double2 h[DEPTH];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k)
        h[k] = reset ? make_double2(c0 == r + k ? 1.0 : 0.0, c0 + 1 == r + k ? 1.0 : 0.0)
                     : ld_hk<LDH>(hp + (size_t)ld * (r + k), keep);
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        const double4 o = ops[r + k - r_begin];
        double h0 = h[k].x, h1 = h[k].y;
        h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);     // + (c₁ s[r]) s[c]
        h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);     // - u[r] s[c]
        h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);   // - u[c] s[r]
        st_hk<STH>(hp + (size_t)ld * (r + k), make_double2(h0, h1), keep);
        a0 = fma(h0, o.w, a0);
        a1 = fma(h1, o.w, a1);
      }
    }
    for (; r < b1; ++r) {
      const double2 h = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0)
                              : ld_hk<LDH>(hp + (size_t)ld * r, keep);
      const double4 o = ops[r - r_begin];
      double h0 = h.x, h1 = h.y;
      h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);
      h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);
      h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);
      st_hk<STH>(hp + (size_t)ld * r, make_double2(h0, h1), keep);
      a0 = fma(h0, o.w, a0);
      a1 = fma(h1, o.w, a1);
    }
  };
  if constexpr (DESC) { // last block first (see stream_pass1): the partials wait in registers
    double A0[16], A1[16];
#pragma unroll
    for (int bi = 15; bi >= 0; --bi) {
      A0[bi] = 0.0; A1[bi] = 0.0;
      if (r_begin + bi * rb < r_end && c0 < n) block_update2(r_begin + bi * rb, A0[bi], A1[bi]);
    }
    S0 = A0[0]; S1 = A1[0];
#pragma unroll
    for (int bi = 1; bi < 16; ++bi)
      if (r_begin + bi * rb < r_end) { S0 = S0 + A0[bi]; S1 = S1 + A1[bi]; }
  } else {
    bool first = true;
    for (int b0 = r_begin; b0 < r_end && c0 < n; b0 += rb) {
      double a0, a1;
      block_update2(b0, a0, a1);
      S0 = first ? a0 : S0 + a0;
      S1 = first ? a1 : S1 + a1;
      first = false;
    }
  }
  pass2_finish(p, sc, sb, c0, two, S0, S1);
}

__device__ __forceinline__ void pass2_finish(const StreamParams& p, const StreamScalars* sc, const int sb,
                                             const int c0, const bool two, const double S0, const double S1) {
  const int n = p.n;
  if (c0 < n) {
    double* out = p.partial + (size_t)sb * n + c0;
    out[0] = S0;
    if (two) out[1] = S1;
  }
  pass_tail(p, 1, (unsigned)sc->it + 1u, [&](int c, double d) {
    bstore(p, p.tmp2 + c, d * p.g[c]); // term of m = dᵀ∇_new, QN.jl:66
    bstore(p, p.dvec + c, d);          // the new direction (other CTAs of this launch still read s)
  });
}

// =====================================================================================================
// TMA-staged sweeps (option "stream_tma").  B⁻¹ is column-major, so the 512 rows of a CTA's tile are
// ONE contiguous 4 KB run per column: the bulk-copy engine (cp.async.bulk, the 1-D form of TMA) moves
// whole columns of the tile between HBM and a shared-memory ring — kTmaCols columns per stage,
// kTmaStages stages, completion through mbarriers (expect_tx / complete_tx) — while the threads do
// nothing but LDS.128 + fma on landed data.  No thread holds a global address or a load in flight:
// the registers the LDG version spends on 8-16 loads in flight are free, and the DMA keeps
// kTmaStages·kTmaCols·4 KB per CTA on their way.  pass 2 writes its updated columns back into the
// stage and the same engine stores them (cp.async.bulk.global.shared::cta + bulk groups); a stage is
// refilled once its stores have READ the shared memory (cp.async.bulk.wait_group.read).
// Same arithmetic, same order (rule C2) as stream_pass1 / stream_pass2: the same bits.
// columns per stage × stages, chosen by the host for the tile height so that the ring stays <= 96 KB
// (two CTAs per SM): 512-row tiles (4 KB per column) 8 × 3; 256 rows 16 × 3; 128 rows 32 × 3 — the
// shorter the run per column, the more columns have to be on their way to cover the HBM latency
// (a thread of the LDG sweeps walks its 256 columns in 16 dependent round trips of 16 loads).
constexpr int kTmaChunk = 8;   // columns a thread takes out of a landed stage at a time

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, const unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const unsigned parity) {
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\nbra.uni WAIT_LOOP;\nWAIT_DONE:\n}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, const unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the executing thread's bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, const unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               :: "l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" :: "n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;\n" :: "n"(N) : "memory"); }
// generic-proxy writes (st.shared, mbarrier.init) become visible to the async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// shared memory of a TMA sweep: [operands of the super-block | mbarriers | ring]
__host__ __device__ constexpr size_t tma_ring_offset(const int sbcols, const int ops_per_col) {
  return (((size_t)sbcols * ops_per_col * sizeof(double) + 8 * sizeof(uint64_t)) + 127) & ~(size_t)127;
}
__host__ __device__ constexpr size_t tma_smem_bytes(const int sbcols, const int ops_per_col, const int tile_rows,
                                                    const int cols, const int stages) {
  return tma_ring_offset(sbcols, ops_per_col) + (size_t)stages * cols * tile_rows * sizeof(double);
}

// ---- pass 1, TMA-staged
template <int kTmaCols, int kTmaStages>
__global__ void __launch_bounds__(kPassThreads, 2) stream_pass1_tma(const StreamParams p) {
  static_assert(kTmaCols % kTmaChunk == 0 && kTmaStages <= 8, "ring shape");
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = p.n, ld = p.ld, rb = p.rb;
  double* ys = reinterpret_cast<double*>(smraw); // y of this super-block's columns
  uint64_t* full = reinterpret_cast<uint64_t*>(smraw + (size_t)p.sbcols * sizeof(double));
  const int tile_rows = 2 * (int)blockDim.x;
  unsigned char* ring = smraw + tma_ring_offset(p.sbcols, 1);
  const int sb = blockIdx.y;
  const int r_begin = sb * p.grp_cols;
  const int r_end = min(r_begin + p.grp_cols, n);
  const int ncols = r_end - r_begin;
  const int nstg = (ncols + kTmaCols - 1) / kTmaCols;
  const bool reset = p.fused2 && sc->need_reset != 0; // B⁻¹ ← I: nothing to fetch
  const int tile0 = (int)blockIdx.x * tile_rows;       // first row of this tile inside the slab
  const int rows_here = min(tile_rows, ld - tile0);    // ld is a multiple of 16 rows = 128 bytes
  const unsigned col_bytes = (unsigned)rows_here * sizeof(double);
  const double* gsrc = p.H + tile0;                    // column r of the tile at gsrc + ld·r
  auto issue = [&](const int stg) { // one elected thread: the bulk copies of stage `stg`
    const int slot = stg % kTmaStages;
    const int cfirst = stg * kTmaCols, cn = min(kTmaCols, ncols - cfirst);
    mbar_expect_tx(full + slot, col_bytes * (unsigned)cn);
    for (int k = 0; k < cn; ++k)
      bulk_g2s(ring + ((size_t)slot * kTmaCols + k) * tile_rows * sizeof(double),
               gsrc + (size_t)ld * (r_begin + cfirst + k), col_bytes, full + slot);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(full + s, 1);
    fence_async_smem();
  }
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) ys[r - r_begin] = p.y[r];
  __syncthreads();
  if (threadIdx.x == 0 && !reset)
    for (int s = 0; s < kTmaStages && s < nstg; ++s) issue(s);
  const int c0 = p.row0 + tile0 + 2 * (int)threadIdx.x;
  const bool two = (c0 + 1 < n);
  const bool rowok = c0 < n;
  double S0 = 0.0, S1 = 0.0, t0 = 0.0, t1 = 0.0;
  bool first = true;
  int inblk = 0; // columns of the current rule-C2 block already folded into t0, t1
  for (int stg = 0; stg < nstg; ++stg) {
    const int slot = stg % kTmaStages;
    if (!reset) mbar_wait(full + slot, (unsigned)(stg / kTmaStages) & 1u);
    const double2* src = reinterpret_cast<const double2*>(ring + (size_t)slot * kTmaCols * tile_rows * sizeof(double)) + threadIdx.x;
#pragma unroll 1
    for (int ch = 0; ch < kTmaCols / kTmaChunk; ++ch) {
      double2 h[kTmaChunk];
#pragma unroll
      for (int k = 0; k < kTmaChunk; ++k) {
        const int r = r_begin + stg * kTmaCols + ch * kTmaChunk + k;
        h[k] = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0) : src[(size_t)(ch * kTmaChunk + k) * blockDim.x];
      }
#pragma unroll
      for (int k = 0; k < kTmaChunk; ++k) {
        const int cc = stg * kTmaCols + ch * kTmaChunk + k;
        if (cc < ncols && rowok) {
          const double yr = ys[cc];
          t0 = fma(yr, h[k].x, t0);
          t1 = fma(yr, h[k].y, t1);
          if (++inblk == rb || cc == ncols - 1) { // the block's chain is complete: add its partial in order
            S0 = first ? t0 : S0 + t0;
            S1 = first ? t1 : S1 + t1;
            first = false; t0 = 0.0; t1 = 0.0; inblk = 0;
          }
        }
      }
    }
    __syncthreads(); // every thread has read the stage: it can be refilled
    if (threadIdx.x == 0 && !reset && stg + kTmaStages < nstg) issue(stg + kTmaStages);
  }
  pass1_finish(p, sc, sb, c0, two, S0, S1);
}

// ---- pass 2, TMA-staged: columns land in the ring, are updated in place and stored by the engine
template <int kTmaCols, int kTmaStages>
__global__ void __launch_bounds__(kPassThreads, 2) stream_pass2_tma(const StreamParams p) {
  static_assert(kTmaCols % kTmaChunk == 0 && kTmaStages <= 8, "ring shape");
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = p.n, ld = p.ld, rb = p.rb;
  double4* ops = reinterpret_cast<double4*>(smraw); // {c₁s[r], -u[r], s[r], ∇[r]} per column
  uint64_t* full = reinterpret_cast<uint64_t*>(smraw + (size_t)p.sbcols * 4 * sizeof(double));
  const int tile_rows = 2 * (int)blockDim.x;
  unsigned char* ring = smraw + tma_ring_offset(p.sbcols, 4);
  const int sb = blockIdx.y;
  const int r_begin = sb * p.grp_cols;
  const int r_end = min(r_begin + p.grp_cols, n);
  const int ncols = r_end - r_begin;
  const int nstg = (ncols + kTmaCols - 1) / kTmaCols;
  const bool reset = p.fused2 && sc->need_reset != 0;
  const double c1 = sc->c1;
  const int tile0 = (int)blockIdx.x * tile_rows;
  const int rows_here = min(tile_rows, ld - tile0);
  const unsigned col_bytes = (unsigned)rows_here * sizeof(double);
  double* gsrc = p.H + tile0;
  auto stage_ptr = [&](const int slot, const int k) { return ring + ((size_t)slot * kTmaCols + k) * tile_rows * sizeof(double); };
  auto issue = [&](const int stg) {
    const int slot = stg % kTmaStages;
    const int cfirst = stg * kTmaCols, cn = min(kTmaCols, ncols - cfirst);
    mbar_expect_tx(full + slot, col_bytes * (unsigned)cn);
    for (int k = 0; k < cn; ++k)
      bulk_g2s(stage_ptr(slot, k), gsrc + (size_t)ld * (r_begin + cfirst + k), col_bytes, full + slot);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(full + s, 1);
    fence_async_smem();
  }
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
    const double sr = p.s[r];
    ops[r - r_begin] = make_double4(p.fused2 ? c1 * sr : p.w[r], p.nu[r], sr, p.g[r]); // (c₁ s[r]) once per r (rule C4)
  }
  __syncthreads();
  if (threadIdx.x == 0 && !reset)
    for (int s = 0; s < kTmaStages && s < nstg; ++s) issue(s);
  const int c0 = p.row0 + tile0 + 2 * (int)threadIdx.x;
  const bool two = (c0 + 1 < n);
  const bool rowok = c0 < n;
  const double s0 = rowok ? p.s[c0] : 0.0, nu0 = rowok ? p.nu[c0] : 0.0;
  const double s1 = two ? p.s[c0 + 1] : 0.0, nu1 = two ? p.nu[c0 + 1] : 0.0;
  double S0 = 0.0, S1 = 0.0, a0 = 0.0, a1 = 0.0;
  bool first = true;
  int inblk = 0;
  for (int stg = 0; stg < nstg; ++stg) {
    const int slot = stg % kTmaStages;
    if (!reset) mbar_wait(full + slot, (unsigned)(stg / kTmaStages) & 1u);
    double2* buf = reinterpret_cast<double2*>(stage_ptr(slot, 0)) + threadIdx.x;
#pragma unroll 1
    for (int ch = 0; ch < kTmaCols / kTmaChunk; ++ch) {
      double2 h[kTmaChunk];
#pragma unroll
      for (int k = 0; k < kTmaChunk; ++k) {
        const int r = r_begin + stg * kTmaCols + ch * kTmaChunk + k;
        h[k] = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0) : buf[(size_t)(ch * kTmaChunk + k) * blockDim.x];
      }
#pragma unroll
      for (int k = 0; k < kTmaChunk; ++k) {
        const int cc = stg * kTmaCols + ch * kTmaChunk + k;
        if (cc < ncols && rowok) {
          const double4 o = ops[cc];
          double h0 = h[k].x, h1 = h[k].y;
          h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);     // + (c₁ s[r]) s[c]
          h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);     // - u[r] s[c]
          h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);   // - u[c] s[r]
          buf[(size_t)(ch * kTmaChunk + k) * blockDim.x] = make_double2(h0, h1);
          a0 = fma(h0, o.w, a0);
          a1 = fma(h1, o.w, a1);
          if (++inblk == rb || cc == ncols - 1) {
            S0 = first ? a0 : S0 + a0;
            S1 = first ? a1 : S1 + a1;
            first = false; a0 = 0.0; a1 = 0.0; inblk = 0;
          }
        }
      }
    }
    fence_async_smem(); // this thread's updated elements are visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0) {
      const int cfirst = stg * kTmaCols, cn = min(kTmaCols, ncols - cfirst);
      for (int k = 0; k < cn; ++k)
        bulk_s2g(gsrc + (size_t)ld * (r_begin + cfirst + k), stage_ptr(slot, k), col_bytes);
      bulk_commit();
      // all but the newest group have READ their shared memory: the stage used one step ago is free
      // (in reset mode, where nothing is fetched, this is what lets the threads overwrite it two
      // steps from now); refill it with the stage that comes kTmaStages - 1 steps after this one
      bulk_wait_read<1>();
      if (!reset && stg >= 1 && stg - 1 + kTmaStages < nstg) issue(stg - 1 + kTmaStages);
    }
  }
  if (threadIdx.x == 0) bulk_wait_all<0>(); // the stores are complete (and visible) before this CTA signs off
  __syncthreads();
  pass2_finish(p, sc, sb, c0, two, S0, S1);
}

// ---- one-pass mode: ONE sweep applies the pending update of the previous
// iteration and accumulates both GEMVs, t = H y and w = H ∇ (rule C2 chains, same tiling as
// pass1/pass2).  16 n² bytes per iteration instead of 24 n²; not the reference's rounding.
#ifndef QN_PASSONE_MINB
#define QN_PASSONE_MINB 2
#endif
template <int LDH, int STH, int DEPTH = 8>
__global__ void __launch_bounds__(kPassThreads, DEPTH > 8 ? 1 : QN_PASSONE_MINB) stream_pass_one(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(16) double sm[];
  double4* ops = reinterpret_cast<double4*>(sm);         // {pc₁ps[r], -pu[r], ps[r], 0} per column
  double2* yg = reinterpret_cast<double2*>(sm + 4 * p.sbcols); // {y[r], ∇[r]}
  const int n = p.n, ld = p.ld, rb = p.rb;
  const int sb = blockIdx.y;
  const int r_begin = sb * p.sbcols;
  const int r_end = min(r_begin + p.sbcols, n);
  const bool pending = sc->pending != 0;
  const bool reset = sc->need_reset != 0; // B⁻¹ ← I (QN.jl:144-148) is folded into this sweep
  const bool keep = sb < p.persist_sb;    // this super-block stays in L2
  const double pc1 = sc->pc1;
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
    const double psr = p.ps[r];
    ops[r - r_begin] = make_double4(pc1 * psr, -p.pu[r], psr, 0.0);
    yg[r - r_begin] = make_double2(p.y[r], p.g[r]);
  }
  __syncthreads();
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  const bool two = (c0 + 1 < n);
  const double s0 = (c0 < n) ? p.ps[c0] : 0.0, nu0 = (c0 < n) ? -p.pu[c0] : 0.0;
  const double s1 = two ? p.ps[c0 + 1] : 0.0, nu1 = two ? -p.pu[c0 + 1] : 0.0;
  double* hp = p.H + (c0 - p.row0);
  double St0 = 0.0, St1 = 0.0, Sw0 = 0.0, Sw1 = 0.0;
  bool first = true;
  for (int b0 = r_begin; b0 < r_end && c0 < n; b0 += rb) {
    const int b1 = min(b0 + rb, r_end);
    double t0 = 0.0, t1 = 0.0, w0 = 0.0, w1 = 0.0;
    int r = b0;
    for (; r + DEPTH <= b1; r += DEPTH) {
      double2 h[DEPTH];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k)
        h[k] = reset ? make_double2(c0 == r + k ? 1.0 : 0.0, c0 + 1 == r + k ? 1.0 : 0.0)
                     : ld_hk<LDH>(hp + (size_t)ld * (r + k), keep);
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        double h0 = h[k].x, h1 = h[k].y;
        if (reset) st_hk<STH>(hp + (size_t)ld * (r + k), h[k], keep);
        if (pending) {
          const double4 o = ops[r + k - r_begin];
          h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);
          h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);
          h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);
          st_hk<STH>(hp + (size_t)ld * (r + k), make_double2(h0, h1), keep);
        }
        const double2 q = yg[r + k - r_begin];
        t0 = fma(q.x, h0, t0); t1 = fma(q.x, h1, t1);
        w0 = fma(h0, q.y, w0); w1 = fma(h1, q.y, w1);
      }
    }
    for (; r < b1; ++r) {
      const double2 h = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0)
                              : ld_hk<LDH>(hp + (size_t)ld * r, keep);
      double h0 = h.x, h1 = h.y;
      if (reset) st_hk<STH>(hp + (size_t)ld * r, h, keep);
      if (pending) {
        const double4 o = ops[r - r_begin];
        h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);
        h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);
        h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);
        st_hk<STH>(hp + (size_t)ld * r, make_double2(h0, h1), keep);
      }
      const double2 q = yg[r - r_begin];
      t0 = fma(q.x, h0, t0); t1 = fma(q.x, h1, t1);
      w0 = fma(h0, q.y, w0); w1 = fma(h1, q.y, w1);
    }
    St0 = first ? t0 : St0 + t0; St1 = first ? t1 : St1 + t1;
    Sw0 = first ? w0 : Sw0 + w0; Sw1 = first ? w1 : Sw1 + w1;
    first = false;
  }
  if (c0 < n) {
    double* o1 = p.partial + (size_t)sb * n + c0;
    double* o2 = p.partial2 + (size_t)sb * n + c0;
    o1[0] = St0; o2[0] = Sw0;
    if (two) { o1[1] = St1; o2[1] = Sw1; }
  }
  // tail: the last CTA of a row tile adds both sets of super-block partials in order
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(p.tile_count + blockIdx.x, 1u);
    is_last = (prev == (unsigned)p.nsb - 1);
    if (is_last) p.tile_count[blockIdx.x] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (c0 < n) {
    const volatile double* p1 = p.partial + c0;
    const volatile double* p2 = p.partial2 + c0;
    double T0 = p1[0], T1 = two ? p1[1] : 0.0, W0 = p2[0], W1 = two ? p2[1] : 0.0;
    for (int j = 1; j < p.nsb; ++j) {
      T0 = T0 + p1[(size_t)j * n]; W0 = W0 + p2[(size_t)j * n];
      if (two) { T1 = T1 + p1[(size_t)j * n + 1]; W1 = W1 + p2[(size_t)j * n + 1]; }
    }
    // row-sharded mode: the consumer may still be reading the previous iteration's sums on a
    // peer, so t and w alternate between two buffers by iteration parity
    const int par = (p.nshard > 1) ? ((sc->it + 1) & 1) * p.nvstride * 2 : 0;
    bstore(p, p.tvec + par + c0, T0); bstore(p, p.wvec + par + c0, W0);
    if (two) { bstore(p, p.tvec + par + c0 + 1, T1); bstore(p, p.wvec + par + c0 + 1, W1); }
  }
  shard_signal(p, 0, (unsigned)sc->it + 1u);
}

// one-pass mode: everything between two sweeps, in one CTA — the O(n) part of iteration `it`
// after its sweep (u, yᵀB⁻¹y, c₁, d, m; reset decision; line search; update_state!), then the
// head of iteration it+1 (∂logdensity!, convergence, y, sᵀy).  Independent reductions share a
// round: {yᵀB⁻¹y, sᵀ∇, uᵀ∇} and {ℓ, sᵀy}.
__global__ void __launch_bounds__(kVecThreads) stream_vec_onepass(const StreamParams p) {
  extern __shared__ __align__(16) double vsm[];
  __shared__ double red[4];
  __shared__ int flag;
  StreamScalars* sc = p.sc;
  if (sc->done) return;
  const int n = p.n;
  const int ns = (n + 1) & ~1;
  int it = sc->it + 1;
  const double l0 = sc->l0;
  double* stage = p.stage_in_smem ? vsm : p.tmp; // 3*ns doubles
  double m;
  if (!sc->skip_update) {
    const double rho = sc->rho;
    shard_wait(p, 0, (unsigned)it);
    const int par = (p.nshard > 1) ? (it & 1) * p.nvstride * 2 : 0;
    const double* tv = p.tvec + par;
    const double* wv = p.wvec + par;
    double r3[3];
    cta_csum_multi<3>(n, stage, ns, red, [&](int c, double (&o)[3]) {
      const double T = tv[c], gc = p.g[c];
      const double u = T * rho;   // B⁻¹y[c] = tρ, QN.jl:50
      p.nu[c] = u;
      o[0] = T * p.y[c];          // term of yᵀB⁻¹y, QN.jl:51
      o[1] = p.s[c] * gc;         // term of sᵀ∇
      o[2] = u * gc;              // term of uᵀ∇
    }, r3);
    const double c1 = fma(r3[0], rho, 1.0) * rho; // QN.jl:54
    const double sg = r3[1], ug = r3[2];
    // d = ((w + (c₁ s)(sᵀ∇)) - u (sᵀ∇)) - s (uᵀ∇) ; m = dᵀ∇ ; this update becomes the pending one
    m = cta_csum(n, stage, red, [&](int c) {
      const double sc_ = p.s[c], uc = p.nu[c];
      double d = fma(c1 * sc_, sg, wv[c]);
      d = fma(-uc, sg, d);
      d = fma(-sc_, ug, d);
      p.ps[c] = sc_;
      p.pu[c] = uc;
      p.s[c] = d; // `s, y = y, s`, QN.jl:65,268
      return d * p.g[c];
    });
    if (threadIdx.x == 0) { sc->pc1 = c1; sc->pending = 1; sc->c1 = c1; }
    __syncthreads();
  } else {
    m = sc->m;
  }
  stream_finish_iteration(p, sc, it, l0, m, stage, red);
  __syncthreads();
  if (sc->done) return;
  // ---- head of iteration it+1: ℓ₀, ∇ = ∂logdensity!(x_old) fused with y and sᵀy   QN.jl:254,36-41
  it = it + 1;
  const int N = n >> 1;
  {
    // stage the ℓ terms (N of them) and the sᵀy terms (n, in element order) in one sweep
    double* st0 = stage;
    double* st1 = stage + ns;
    const int T = blockDim.x;
    for (int i0 = threadIdx.x; i0 < N; i0 += 2 * T) {
      double t1[2], t2[2], gp1[2], gp2[2], s1[2], s2[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = i0 + q * T;
        if (i < N) { t1[q] = p.x[i]; t2[q] = p.x[i + N]; gp1[q] = p.gp[i]; gp2[q] = p.gp[i + N]; s1[q] = p.s[i]; s2[q] = p.s[i + N]; }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = i0 + q * T;
        if (i < N) {
          const double a = t2[q] - t1[q] * t1[q];
          const double b = 1.0 - t1[q];
          const double g1 = (400.0 * a) * t1[q] + 2.0 * b, g2 = 200.0 * (t1[q] * t1[q] - t2[q]);
          const double y1 = gp1[q] - g1, y2 = gp2[q] - g2; // y = ∇_old - ∇_new, QN.jl:38
          p.g[i] = g1; p.g[i + N] = g2;
          p.y[i] = y1; p.y[i + N] = y2;
          st0[i] = 100.0 * (a * a) + b * b;
          st1[i] = s1[q] * y1;
          st1[i + N] = s2[q] * y2;
        }
      }
    }
  }
  double tail_d = 0.0;
  if (n & 1) {
    tail_d = 1.0 - p.x[n - 1];
    if (threadIdx.x == 0) {
      const double gt = 2.0 * tail_d, yt = p.gp[n - 1] - gt;
      p.g[n - 1] = gt;
      p.y[n - 1] = yt;
      stage[ns + n - 1] = p.s[n - 1] * yt;
    }
  }
  double r2[2];
  const int cnt2[2] = {N, n};
  cta_reduce_multi<2>(cnt2, stage, ns, red, r2);
  double lnew = 0.0 - r2[0];
  if (n & 1) lnew = -fma(tail_d, tail_d, -lnew);
  const double sy = r2[1];
  __syncthreads();
  if (threadIdx.x == 0) { sc->l0 = lnew; sc->fevals += 1; flag = 1; }
  if (!isfinite(lnew)) { // QN.jl:255
    if (threadIdx.x == 0) { sc->status = QN_NONFINITE; sc->done = 1; sc->exit_it = it; }
    return;
  }
  __syncthreads();
  bool small = true; // maximum(abs, ∇) < tol, QN.jl:257
  cta_map(n, [&](int i) { return p.g[i]; }, [&](int, double gi) { small = small && (fabs(gi) < p.tol); });
  if (!small) flag = 0;
  __syncthreads();
  if (flag) {
    if (threadIdx.x == 0) { sc->status = QN_CONVERGED; sc->result = lnew; sc->done = 1; sc->exit_it = it; }
    return;
  }
  if (threadIdx.x == 0) { sc->sy = sy; sc->rho = 1.0 / sy; sc->skip_update = 0; } // QN.jl:39,42
}

// ---- one-pass mode, cluster form of stream_vec_onepass.  The O(n) work between two sweeps
// is latency-bound in a single CTA (every dependent round is an L2 round trip through one
// SM); here a thread-block CLUSTER of kCluCtas CTAs keeps the whole state in REGISTERS — a
// thread owns PPT Rosenbrock pairs (i, i+N) — and only the rule-C1 reductions leave the
// thread: every term is stored through distributed shared memory into the CTA that reduces
// that array (array k -> cluster rank k), one warp there runs the sequential lane chains, and
// the result is read back by everyone.  Two cluster barriers per round of reductions.
// Same operations on the same operands as stream_vec_onepass: the same bits.
constexpr int kCluCtas = 8;
constexpr int kCluThreads = 512;

struct CluShared {
  double red;            // this rank's reduction result
  int flags[kCluCtas];   // rank 0: per-CTA votes
  int all;               // rank 0: the combined vote
};

// rank k reduces the cnt[k] staged terms in ITS `stage` by rule C1; every thread gets all K sums.
// vote < 0: no vote; otherwise the CTAs' votes (CTA-uniform 0/1) are and-ed into *vote_out.
template <int K>
__device__ __forceinline__ void cluster_reduce(cooperative_groups::cluster_group& cl, const int (&cnt)[K],
                                               const double* stage, CluShared* sh, double (&out)[K],
                                               int vote = -1, int* vote_out = nullptr) {
  const unsigned rank = cl.block_rank();
  if (vote >= 0 && threadIdx.x == 0) cl.map_shared_rank(sh, 0)->flags[rank] = vote;
  cl.sync(); // every term (and vote) has landed in its reducer's shared memory
  if (rank < K && threadIdx.x < 32) {
    const int l = threadIdx.x, n = cnt[rank];
    double p = (l < n) ? stage[l] : 0.0;
    int j = l + 32;
    if (j + 224 < n) { // the loads of the next eight terms are in flight while these eight are added
      double x[8], z[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = stage[j + 32 * k];
      j += 256;
      for (; j + 224 < n; j += 256) {
#pragma unroll
        for (int k = 0; k < 8; ++k) z[k] = stage[j + 32 * k];
#pragma unroll
        for (int k = 0; k < 8; ++k) p = p + x[k];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = z[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) p = p + x[k];
    }
    for (; j < n; j += 32) p = p + stage[j];
    p = warp_butterfly(p);
    if (l == 0) sh->red = p;
  }
  if (vote >= 0 && rank == 0 && threadIdx.x == 32) {
    int a = 1;
#pragma unroll
    for (int k = 0; k < kCluCtas; ++k) a &= sh->flags[k];
    sh->all = a;
  }
  cl.sync();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = cl.map_shared_rank(sh, k)->red;
  if (vote >= 0) *vote_out = cl.map_shared_rank(sh, 0)->all;
}

// ONEPASS = false: the same kernel between pass2 of iteration `it` and pass1 of iteration it+1 of
// the two-pass (reference-order) mode — it replaces vec_end + vec_begin: d and the terms of m
// come from pass2's tail instead of being formed here.
template <int PPT, bool ONEPASS>
__global__ void __cluster_dims__(kCluCtas, 1, 1) __launch_bounds__(kCluThreads)
stream_vec_cluster(const StreamParams p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ __align__(16) double vsm[]; // ns doubles: the array this rank reduces
  __shared__ CluShared sh;
  StreamScalars* sc = p.sc;
  if (sc->done) return; // cluster-uniform: nobody writes sc before the first cluster barrier
  const int n = p.n, N = n >> 1;
  const unsigned rank = cl.block_rank();
  const int tau = (int)rank * kCluThreads + (int)threadIdx.x;
  const bool leader = (tau == 0);
  constexpr int NE = 2 * PPT + 1; // elements of this thread: PPT pairs (i, i+N) and the odd tail
  int idx[NE];
  bool ok[NE], mine[NE];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int i = tau + k * kCluCtas * kCluThreads;
    idx[2 * k] = i; idx[2 * k + 1] = i + N;
    ok[2 * k] = ok[2 * k + 1] = mine[2 * k] = mine[2 * k + 1] = (i < N);
  }
  idx[NE - 1] = n - 1;
  ok[NE - 1] = (n & 1) != 0;     // the tail element is replicated in every thread
  mine[NE - 1] = ok[NE - 1] && leader; // ... and stored by one
  double* st[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) st[k] = cl.map_shared_rank(vsm, k);

  double x[NE], g[NE], s[NE], y[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    x[j] = g[j] = s[j] = y[j] = 0.0;
    if (ok[j]) { x[j] = p.x[idx[j]]; g[j] = p.g[idx[j]]; s[j] = p.s[idx[j]]; y[j] = p.y[idx[j]]; }
  }
  int it = sc->it + 1;
  const double l0 = sc->l0;
  const int skip_update = sc->skip_update;
  int fevals = 0;
  double m;
  if (!skip_update && ONEPASS) {
    const double rho = sc->rho;
    shard_wait(p, 0, (unsigned)it); // row-sharded mode: every shard's slices of t and w have arrived
    const int par = (p.nshard > 1) ? (it & 1) * p.nvstride * 2 : 0;
    double T[NE], W[NE], u[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      T[j] = W[j] = 0.0;
      if (ok[j]) { T[j] = p.tvec[par + idx[j]]; W[j] = p.wvec[par + idx[j]]; }
    }
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      u[j] = T[j] * rho; // B⁻¹y[c] = tρ, QN.jl:50
      if (mine[j]) {
        st[0][idx[j]] = T[j] * y[j]; // term of yᵀB⁻¹y, QN.jl:51
        st[1][idx[j]] = s[j] * g[j]; // term of sᵀ∇
        st[2][idx[j]] = u[j] * g[j]; // term of uᵀ∇
      }
    }
    double r3[3];
    const int cnt3[3] = {n, n, n};
    cluster_reduce<3>(cl, cnt3, vsm, &sh, r3);
    const double c1 = fma(r3[0], rho, 1.0) * rho; // QN.jl:54
    const double sg = r3[1], ug = r3[2];
    // d = ((w + (c₁ s)(sᵀ∇)) - u (sᵀ∇)) - s (uᵀ∇) ; m = dᵀ∇ ; this update becomes the pending one
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      double d = fma(c1 * s[j], sg, W[j]);
      d = fma(-u[j], sg, d);
      d = fma(-s[j], ug, d);
      if (mine[j]) {
        p.ps[idx[j]] = s[j];
        p.pu[idx[j]] = u[j];
        st[0][idx[j]] = d * g[j];
      }
      s[j] = d; // `s, y = y, s`, QN.jl:65,268
    }
    double r1[1];
    const int cnt1[1] = {n};
    cluster_reduce<1>(cl, cnt1, vsm, &sh, r1);
    m = r1[0];
    if (leader) { sc->pc1 = c1; sc->pending = 1; sc->c1 = c1; }
  } else if (!skip_update) {
    // two-pass mode: pass2's tail left the new direction in dvec (`s, y = y, s`, QN.jl:65,268);
    // m = dᵀ∇_new = csum of d[c]·∇[c], QN.jl:66-68
    shard_wait(p, 1, (unsigned)it); // row-sharded mode: every shard's slice of d has arrived
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      s[j] = ok[j] ? p.dvec[idx[j]] : 0.0;
      if (mine[j]) st[0][idx[j]] = s[j] * g[j];
    }
    double r1[1];
    const int cnt1[1] = {n};
    cluster_reduce<1>(cl, cnt1, vsm, &sh, r1);
    m = r1[0];
  } else {
    m = sc->m;
  }
  int need_reset = 0;
  if (m <= 0.0) { // QN.jl:272-280
    need_reset = 1;
    if (leader && ONEPASS) sc->pending = 0; // the reset discards the pending update
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      s[j] = g[j];
      if (mine[j]) st[0][idx[j]] = g[j] * g[j];
    }
    double r1[1];
    const int cnt1[1] = {n};
    cluster_reduce<1>(cl, cnt1, vsm, &sh, r1);
    m = r1[0];
  }
  // ---- linesearch!  QN.jl:165-232 (the same scalar code on the same values in every thread)
  // A trial step! (QN.jl:157-162): x + a s as one fma, then ℓ.  Far from the optimum the
  // backtracking sequence is the lower-clamp chain a, aρₗ, aρₗ², … (QN.jl:225), so a round
  // evaluates the next three links of that chain as well — rank k reduces candidate k — and a
  // later trial whose step is bit-equal to a cached candidate costs no round at all.
  const LsParams& ls = p.ls;
  const int spec = sc->spec;
  double ca[4], cv[4];
  int ncache = 0;
  auto trial = [&](const double a) {
    ++fevals;
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (k < ncache && ca[k] == a) return cv[k];
    ca[0] = a;
#pragma unroll
    for (int k = 1; k < 4; ++k) ca[k] = ca[k - 1] * ls.rho_lo;
    const int nc = spec ? 4 : 1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < nc) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          if (mine[2 * k]) {
            const double t1 = fma(ca[c], s[2 * k], x[2 * k]), t2 = fma(ca[c], s[2 * k + 1], x[2 * k + 1]);
            const double aa = t2 - t1 * t1;
            const double bb = 1.0 - t1;
            st[c][idx[2 * k]] = 100.0 * (aa * aa) + bb * bb;
          }
        }
      }
    }
    double r4[4];
    if (spec) {
      const int cnt4[4] = {N, N, N, N};
      cluster_reduce<4>(cl, cnt4, vsm, &sh, r4);
    } else {
      double r1[1];
      const int cnt1[1] = {N};
      cluster_reduce<1>(cl, cnt1, vsm, &sh, r1);
      r4[0] = r1[0];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double v = 0.0 - r4[c < nc ? c : 0];
      if (n & 1) {
        const double d = 1.0 - fma(ca[c], s[NE - 1], x[NE - 1]);
        v = -fma(d, d, -v);
      }
      cv[c] = v;
    }
    ncache = nc;
    return cv[0];
  };
  double lx_last;
  const double alpha = backtracking_linesearch(ls, l0, m, trial, lx_last);
  if (alpha == 0.0) { // QN.jl:284
    if (leader) { sc->status = QN_LINESEARCH_FAILED; sc->done = 1; sc->fevals += fevals; sc->need_reset = need_reset; sc->exit_it = it; }
#pragma unroll
    for (int j = 0; j < NE; ++j) if (mine[j]) p.s[idx[j]] = s[j];
    cl.sync(); // nobody leaves while its shared memory may still be read remotely
    return;
  }
  // update_state!  QN.jl:21-27 ; ∇_old ← ∇_new  QN.jl:286
  const bool last = (it >= p.max_iter);
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    s[j] = alpha * s[j];
    x[j] = x[j] + s[j];
    if (mine[j]) {
      p.s[idx[j]] = s[j];
      p.x[idx[j]] = x[j];
      if (p.trace != nullptr && it <= p.ntrace) p.trace[(size_t)(it - 1) * n + idx[j]] = x[j];
      if (!last) p.gp[idx[j]] = g[j];
    }
  }
  if (last) {
    if (leader) {
      sc->m = m; sc->it = it; sc->iters += 1; sc->fevals += fevals; sc->need_reset = need_reset;
      sc->status = QN_MAX_ITER; sc->done = 1; sc->exit_it = it;
    }
    cl.sync();
    return;
  }
  // ---- head of iteration it+1: ℓ₀, ∇ = ∂logdensity!(x_old) fused with y and sᵀy   QN.jl:254,36-41
  const int it_done = it;
  it = it + 1;
  bool small = true;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const double t1 = x[2 * k], t2 = x[2 * k + 1];
    const double aa = t2 - t1 * t1;
    const double bb = 1.0 - t1;
    const double g1 = (400.0 * aa) * t1 + 2.0 * bb, g2 = 200.0 * (t1 * t1 - t2);
    const double y1 = g[2 * k] - g1, y2 = g[2 * k + 1] - g2; // y = ∇_old - ∇_new, QN.jl:38
    if (mine[2 * k]) {
      p.g[idx[2 * k]] = g1; p.g[idx[2 * k + 1]] = g2;
      p.y[idx[2 * k]] = y1; p.y[idx[2 * k + 1]] = y2;
      st[0][idx[2 * k]] = 100.0 * (aa * aa) + bb * bb;
      st[1][idx[2 * k]] = s[2 * k] * y1;
      st[1][idx[2 * k + 1]] = s[2 * k + 1] * y2;
      small = small && (fabs(g1) < p.tol) && (fabs(g2) < p.tol); // maximum(abs, ∇) < tol, QN.jl:257
    }
  }
  double tail_d = 0.0;
  if (n & 1) {
    tail_d = 1.0 - x[NE - 1];
    const double gt = 2.0 * tail_d, yt = g[NE - 1] - gt;
    if (mine[NE - 1]) {
      p.g[n - 1] = gt;
      p.y[n - 1] = yt;
      st[1][n - 1] = s[NE - 1] * yt;
    }
    small = small && (fabs(gt) < p.tol);
  }
  const int cta_small = __syncthreads_and(small ? 1 : 0);
  double r2[2];
  const int cnt2[2] = {N, n};
  int all_small = 0;
  cluster_reduce<2>(cl, cnt2, vsm, &sh, r2, cta_small, &all_small);
  double lnew = 0.0 - r2[0];
  if (n & 1) lnew = -fma(tail_d, tail_d, -lnew);
  const double sy = r2[1];
  if (leader) {
    sc->m = m; sc->it = it_done; sc->iters += 1; sc->fevals += fevals + 1; sc->need_reset = need_reset;
    sc->spec = (fevals > 1); // this line search backtracked: expect the next one to
    sc->l0 = lnew;
    if (!isfinite(lnew)) { // QN.jl:255
      sc->status = QN_NONFINITE; sc->done = 1; sc->exit_it = it;
    } else if (all_small) {
      sc->status = QN_CONVERGED; sc->result = lnew; sc->done = 1; sc->exit_it = it;
    } else {
      sc->sy = sy; sc->rho = 1.0 / sy; sc->skip_update = 0; // QN.jl:39,42
    }
  }
  cl.sync();
}

// one-pass mode, end of the solve: apply the pending update so H is the true B⁻¹
__global__ void __launch_bounds__(kPassThreads) stream_flush_pending(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if (!sc->pending) return;
  const int n = p.n, ld = p.ld;
  const double pc1 = sc->pc1;
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  if (c0 >= n) return;
  const bool two = (c0 + 1 < n);
  const double s0 = p.ps[c0], nu0 = -p.pu[c0];
  const double s1 = two ? p.ps[c0 + 1] : 0.0, nu1 = two ? -p.pu[c0 + 1] : 0.0;
  const int r_begin = blockIdx.y * p.sbcols, r_end = min(r_begin + p.sbcols, n);
  for (int r = r_begin; r < r_end; ++r) {
    double2* hp = reinterpret_cast<double2*>(p.H + (c0 - p.row0) + (size_t)ld * r);
    const double psr = p.ps[r], wr = pc1 * psr, nur = -p.pu[r];
    double2 h = *hp;
    h.x = fma(wr, s0, h.x); h.y = fma(wr, s1, h.y);
    h.x = fma(nur, s0, h.x); h.y = fma(nur, s1, h.y);
    h.x = fma(nu0, psr, h.x); h.y = fma(nu1, psr, h.y);
    *hp = h;
  }
}

// B⁻¹ ← I (QN.jl:144-148), only when vec_end asked for it
template <bool AT_EXIT>
__global__ void __launch_bounds__(256) stream_identity(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if ((!AT_EXIT && sc->done) || !sc->need_reset) return;
  const size_t ld2 = (size_t)p.ld / 2;
  const size_t total = ld2 * p.n; // double2 elements
  double2* H2 = reinterpret_cast<double2*>(p.H);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ld2, c = 2 * (i % ld2) + (size_t)p.row0; // global row of the slab's entry
    H2[i] = make_double2(c == r ? 1.0 : 0.0, c + 1 == r ? 1.0 : 0.0);
  }
}

}  // namespace qn
