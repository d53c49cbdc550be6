// qn_stream.cu — the large-n path (BASELINE.json configs[4], n in the thousands): B⁻¹ lives
// in HBM and one problem uses the whole GPU.
//
// Per BFGS iteration (QN.jl:253-288) six kernels run back to back, captured once in a CUDA
// graph and replayed; all control flow stays on the device (a `done` flag turns the
// remaining launches into no-ops; the host only polls it every few iterations):
//
//   vec_begin  (1 CTA)   ℓ₀, ∇ = ∂logdensity!(x_old); convergence; y, sᵀy, ρ        QN.jl:254-264,36-42
//   pass1      (grid)    block partials of t = B⁻¹y                                   QN.jl:44-49
//   vec_mid    (1 CTA)   t, u = tρ, yᵀB⁻¹y, c₁, c₁s                                   QN.jl:50-54
//   pass2      (grid)    B⁻¹ ← B⁻¹ + c₁ssᵀ - usᵀ - suᵀ fused with partials of d = B⁻¹∇   QN.jl:56-67
//   vec_end    (1 CTA)   d, m, reset decision, the whole line search, update_state!   QN.jl:66-68,272-286
//   identity   (grid)    B⁻¹ ← I when vec_end asked for a reset                       QN.jl:144-148
//
// The two passes are pure streams over the n×n matrix: pass1 reads 8n² bytes, pass2 reads
// and writes 16n² — the reference's own two-pass order (24n² bytes per iteration), which
// is what keeps this path bit-identical to the oracle.  Their bound is HBM bandwidth.
//
// Work split of a pass: a CTA owns 512 consecutive rows (2 per thread, one 128-bit access
// per column) and one SUPER-BLOCK of 16·rb columns (rule C2, rb = chain_block(n) = 16 at
// n = 8192).  Inside it each thread runs the rule-C2 chains of its two rows and adds the
// block partials in order; the per-super-block sums go to a small [nsb × n] array that the
// next single-CTA kernel adds up in ascending order.  Every global access of a warp is 512
// contiguous bytes; 8 independent 128-bit loads per thread are in flight.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

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
  const double l0 = cta_rosenbrock(n, [&](int i) { return xv[i]; }, p.g, stage, &red); // QN.jl:254
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
    is_last = (prev == (unsigned)p.nsb - 1);
    if (is_last) p.tile_count[blockIdx.x] = 0; // ready for the next pass
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  if (c0 < p.n) {
    const bool two = (c0 + 1 < p.n);
    const volatile double* part = p.partial + c0; // written by other CTAs of this launch
    double T0 = part[0], T1 = two ? part[1] : 0.0;
    for (int j = 1; j < p.nsb; ++j) {
      T0 = T0 + part[(size_t)j * p.n];
      if (two) T1 = T1 + part[(size_t)j * p.n + 1];
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
    return cta_rosenbrock(n, [&](int i) { return fma(a, sv[i], xv[i]); }, nullptr, stage, &red);
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
  stream_finish_iteration(p, sc, it, l0, m, stage, &red);
}

#ifndef QN_PASS2_MINB
#define QN_PASS2_MINB 2
#endif
#ifndef QN_PASS1_MINB
#define QN_PASS1_MINB 4
#endif
// ---- pass 1: partial[sb][c] = Σ_{blocks of sb} chain_b(c),  chain_b(c) = Σ_{r in b} y[r] H[c,r]
template <int LDH, int STH, int DEPTH = 8>
__global__ void __launch_bounds__(kPassThreads, DEPTH > 8 ? 1 : QN_PASS1_MINB) stream_pass1(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(16) double sm[];
  double* ys = sm; // y of this super-block's columns
  const int n = p.n, ld = p.ld, rb = p.rb;
  const int sb = blockIdx.y;
  const int r_begin = sb * p.sbcols;
  const int r_end = min(r_begin + p.sbcols, n);
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) ys[r - r_begin] = p.y[r];
  __syncthreads();
  const bool reset = p.fused2 && sc->need_reset != 0; // B⁻¹ ← I (QN.jl:144-148) folded into the sweeps
  const int c0 = p.row0 + (int)(blockIdx.x * 2 * blockDim.x) + 2 * (int)threadIdx.x;
  const bool two = (c0 + 1 < n);
  const double* hp = p.H + (c0 - p.row0);
  double S0 = 0.0, S1 = 0.0;
  bool first = true;
  for (int b0 = r_begin; b0 < r_end && c0 < n; b0 += rb) {
    const int b1 = min(b0 + rb, r_end);
    double t0 = 0.0, t1 = 0.0;
    int r = b0;
    for (; r + DEPTH <= b1; r += DEPTH) { // 8 independent 128-bit loads in flight
      double2 h[DEPTH];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k)
        h[k] = reset ? make_double2(c0 == r + k ? 1.0 : 0.0, c0 + 1 == r + k ? 1.0 : 0.0)
                     : ld_h<LDH>(hp + (size_t)ld * (r + k));
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        const double yr = ys[r + k - r_begin];
        t0 = fma(yr, h[k].x, t0);
        t1 = fma(yr, h[k].y, t1);
      }
    }
    for (; r < b1; ++r) {
      const double2 h = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0)
                              : ld_h<LDH>(hp + (size_t)ld * r);
      const double yr = ys[r - r_begin];
      t0 = fma(yr, h.x, t0);
      t1 = fma(yr, h.y, t1);
    }
    S0 = first ? t0 : S0 + t0;
    S1 = first ? t1 : S1 + t1;
    first = false;
  }
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
template <int LDH, int STH, int DEPTH = 8>
__global__ void __launch_bounds__(kPassThreads, DEPTH > 8 ? 1 : QN_PASS2_MINB) stream_pass2(const StreamParams p) {
  const StreamScalars* sc = p.sc;
  if (sc->done || sc->skip_update) return;
  extern __shared__ __align__(16) double sm[];
  double4* ops = reinterpret_cast<double4*>(sm); // {c₁s[r], -u[r], s[r], ∇[r]} per column
  const int n = p.n, ld = p.ld, rb = p.rb;
  const int sb = blockIdx.y;
  const int r_begin = sb * p.sbcols;
  const int r_end = min(r_begin + p.sbcols, n);
  const bool reset = p.fused2 && sc->need_reset != 0;
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
  bool first = true;
  for (int b0 = r_begin; b0 < r_end && c0 < n; b0 += rb) {
    const int b1 = min(b0 + rb, r_end);
    double a0 = 0.0, a1 = 0.0;
    int r = b0;
    for (; r + DEPTH <= b1; r += DEPTH) {
      double2 h[DEPTH];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k)
        h[k] = reset ? make_double2(c0 == r + k ? 1.0 : 0.0, c0 + 1 == r + k ? 1.0 : 0.0)
                     : ld_h<LDH>(hp + (size_t)ld * (r + k));
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        const double4 o = ops[r + k - r_begin];
        double h0 = h[k].x, h1 = h[k].y;
        h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);     // + (c₁ s[r]) s[c]
        h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);     // - u[r] s[c]
        h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);   // - u[c] s[r]
        st_h<STH>(hp + (size_t)ld * (r + k), make_double2(h0, h1));
        a0 = fma(h0, o.w, a0);
        a1 = fma(h1, o.w, a1);
      }
    }
    for (; r < b1; ++r) {
      const double2 h = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0)
                              : ld_h<LDH>(hp + (size_t)ld * r);
      const double4 o = ops[r - r_begin];
      double h0 = h.x, h1 = h.y;
      h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);
      h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);
      h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);
      st_h<STH>(hp + (size_t)ld * r, make_double2(h0, h1));
      a0 = fma(h0, o.w, a0);
      a1 = fma(h1, o.w, a1);
    }
    S0 = first ? a0 : S0 + a0;
    S1 = first ? a1 : S1 + a1;
    first = false;
  }
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
                     : ld_h<LDH>(hp + (size_t)ld * (r + k));
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) {
        double h0 = h[k].x, h1 = h[k].y;
        if (reset) st_h<STH>(hp + (size_t)ld * (r + k), h[k]);
        if (pending) {
          const double4 o = ops[r + k - r_begin];
          h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);
          h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);
          h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);
          st_h<STH>(hp + (size_t)ld * (r + k), make_double2(h0, h1));
        }
        const double2 q = yg[r + k - r_begin];
        t0 = fma(q.x, h0, t0); t1 = fma(q.x, h1, t1);
        w0 = fma(h0, q.y, w0); w1 = fma(h1, q.y, w1);
      }
    }
    for (; r < b1; ++r) {
      const double2 h = reset ? make_double2(c0 == r ? 1.0 : 0.0, c0 + 1 == r ? 1.0 : 0.0)
                              : ld_h<LDH>(hp + (size_t)ld * r);
      double h0 = h.x, h1 = h.y;
      if (reset) st_h<STH>(hp + (size_t)ld * r, h);
      if (pending) {
        const double4 o = ops[r - r_begin];
        h0 = fma(o.x, s0, h0); h1 = fma(o.x, s1, h1);
        h0 = fma(o.y, s0, h0); h1 = fma(o.y, s1, h1);
        h0 = fma(nu0, o.z, h0); h1 = fma(nu1, o.z, h1);
        st_h<STH>(hp + (size_t)ld * r, make_double2(h0, h1));
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

// -------------------------------------------------------------------------- host side
namespace {

struct StreamWork {
  qn::StreamParams p{};
  void* block = nullptr; // one allocation for everything but H
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  size_t vec_smem = 0;
  int pass_threads = 256; // threads of a sweep CTA
  int clu_ppt = 0;      // one-pass mode: pairs per thread of the cluster vector kernel (0: single-CTA kernel)
  size_t clu_smem = 0;
};

using PassKernel = void (*)(const qn::StreamParams);

// cache hints of the B⁻¹ accesses per kernel: measured defaults (profiles/r1_stream_cache_hints.txt,
// n = 8192: pass2 .lu loads + .cs stores 356 -> 316 µs per iteration; one-pass sweep .cg loads
// 216 -> 206 µs).  The full 5×4 sweep of hints was run once (scripts/exp_stream_hints.sh, git
// history); the combinations worth keeping are instantiated and selectable for A/B runs.
int hint_env(const char* name, int dflt, int hi) {
  const char* v = getenv(name);
  if (!v) return dflt;
  const int k = atoi(v);
  return (k < 0 || k > hi) ? dflt : k;
}
// deep = 16 instead of 8 loads in flight per thread (a whole rule-C2 block): for grids that leave
// most SMs with at most one CTA — the shards of the row-sharded mode — where each thread's walk
// over its 256 columns is a chain of exposed memory latencies
bool sweep_deep(const qn_context* ctx, const qn::StreamParams& p) {
  const int v = hint_env("QN_SWEEP_DEPTH", 0, 16);
  if (v) return v == 16;
  return p.ntiles * p.nsb <= ctx->prop.multiProcessorCount;
}
PassKernel pass1_kernel(bool deep = false) {
  if (deep) return qn::stream_pass1<0, 0, 16>;
  return hint_env("QN_HINT_P1_LD", 0, 4) == 4 ? qn::stream_pass1<4, 0> : qn::stream_pass1<0, 0>;
}
PassKernel pass2_kernel(bool deep = false) {
  if (deep) return qn::stream_pass2<2, 1, 16>;
  const int ld = hint_env("QN_HINT_P2_LD", 2, 4), st = hint_env("QN_HINT_P2_ST", 1, 3);
  if (ld == 0 && st == 0) return qn::stream_pass2<0, 0>; // no hints
  if (ld == 0) return qn::stream_pass2<0, 1>;            // .cs stores only
  if (st == 0) return qn::stream_pass2<2, 0>;            // .lu loads only
  return qn::stream_pass2<2, 1>;                         // default: .lu loads + .cs stores
}
PassKernel pass_one_kernel(bool deep = false) {
  if (deep) return qn::stream_pass_one<4, 0, 16>;
  const int ld = hint_env("QN_HINT_ONE_LD", 4, 4);
  if (ld == 0) return qn::stream_pass_one<0, 0>;         // no hints
  if (ld == 1) return qn::stream_pass_one<1, 0>;         // .cs loads
  return qn::stream_pass_one<4, 0>;                      // default: .cg loads
}

// extra dynamic shared memory for pass1, a residency knob (bytes): QN_P1_PAD_KB
size_t pass1_pad() { return (size_t)hint_env("QN_P1_PAD_KB", 0, 200) * 1024; }
using VecKernel = void (*)(const qn::StreamParams);
VecKernel cluster_kernel(int ppt, bool onepass) {
  if (onepass) return ppt == 1 ? qn::stream_vec_cluster<1, true> : ppt == 2 ? qn::stream_vec_cluster<2, true> : qn::stream_vec_cluster<3, true>;
  return ppt == 1 ? qn::stream_vec_cluster<1, false> : ppt == 2 ? qn::stream_vec_cluster<2, false> : qn::stream_vec_cluster<3, false>;
}

int enqueue_iteration(qn_context* ctx, const StreamWork& w, cudaStream_t st) {
  const qn::StreamParams& p = w.p;
  const dim3 grid(p.ntiles, p.nsb);
  const bool deep = sweep_deep(ctx, p);
  if (p.onepass) { // two kernels per iteration; the head of iteration 1 ran before the graph
    pass_one_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * 6 * p.sbcols, st>>>(p);
    if (w.clu_ppt) cluster_kernel(w.clu_ppt, true)<<<qn::kCluCtas, qn::kCluThreads, w.clu_smem, st>>>(p);
    else qn::stream_vec_onepass<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
    QN_CUDA_TRY(cudaGetLastError());
    return QN_OK;
  }
  if (p.fused2) { // three kernels per iteration; the head of iteration 1 ran before the graph
    pass1_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * p.sbcols + pass1_pad(), st>>>(p);
    pass2_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * 4 * p.sbcols, st>>>(p);
    cluster_kernel(w.clu_ppt, false)<<<qn::kCluCtas, qn::kCluThreads, w.clu_smem, st>>>(p);
    QN_CUDA_TRY(cudaGetLastError());
    return QN_OK;
  }
  qn::stream_vec_begin<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
  pass1_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * p.sbcols, st>>>(p);
  qn::stream_vec_mid<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
  pass2_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * 4 * p.sbcols, st>>>(p);
  qn::stream_vec_end<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
  qn::stream_identity<false><<<ctx->prop.multiProcessorCount * 8, 256, 0, st>>>(p);
  QN_CUDA_TRY(cudaGetLastError());
  return QN_OK;
}

constexpr int kItersPerGraph = 8; // iterations per graph replay; also the cadence of the done-flag poll

void stream_release(StreamWork& w) {
  if (w.exec) cudaGraphExecDestroy(w.exec);
  if (w.graph) cudaGraphDestroy(w.graph);
  if (w.block) cudaFree(w.block);
  if (w.p.H) cudaFree(w.p.H);
  w = StreamWork{};
}

#define QN_TRY_REL(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      return qn_fail(QN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                 \
    }                                                                                     \
  } while (0)

// Allocate and lay out the work space of one GPU (shard `shard` of `nshard`: row tiles
// [row0, row0 + rows) of B⁻¹), pick the sweep CTA size and initialise the scalars.
// The caller fills peer_base[] for nshard > 1 BEFORE calling stream_capture.
int stream_setup(qn_context* ctx, StreamWork& w, int n, const double* theta0, bool theta_on_host,
                 const qn_backtracking* ls, double tol, int max_iter, int ntrace, double* d_trace,
                 int onepass, int shard, int nshard, cudaStream_t st) {
  using namespace qn;
  StreamParams& p = w.p;
  QN_TRY_REL(cudaSetDevice(ctx->device));
  p.n = n;
  p.rb = chain_block(n);
  p.sbcols = 16 * p.rb;
  p.nsb = (n + p.sbcols - 1) / p.sbcols;
  p.max_iter = max_iter;
  p.tol = tol;
  p.ls = LsParams{ls->c1, ls->rho_hi, ls->rho_lo, (long long)ls->iterations, ls->order};
  p.ntrace = ntrace;
  p.trace = d_trace;
  p.onepass = onepass;
  const int row_tiles = (n + kShardRows - 1) / kShardRows;
  p.nshard = nshard;
  p.shard = shard;
  {
    const int t0 = (int)((long long)row_tiles * shard / nshard), t1 = (int)((long long)row_tiles * (shard + 1) / nshard);
    p.row0 = t0 * kShardRows;
    p.rows = std::min(n, t1 * kShardRows) - p.row0;
  }
  // sweep CTAs: 2 rows per thread; halve the CTA (down to 64 threads) while the grid would leave
  // SMs idle — a single SM cannot keep enough loads in flight to use its share of the bandwidth
  w.pass_threads = kPassThreads;
  while (w.pass_threads > 64 &&
         ((p.rows + 2 * w.pass_threads - 1) / (2 * w.pass_threads)) * p.nsb < 2 * ctx->prop.multiProcessorCount)
    w.pass_threads /= 2;
  p.ntiles = (p.rows + 2 * w.pass_threads - 1) / (2 * w.pass_threads);
  p.ld = (p.rows + 15) & ~15;
  QN_TRY_REL(cudaMalloc(&p.H, sizeof(double) * (size_t)p.ld * n));
  const size_t nv = ((size_t)n + 31) & ~(size_t)31;
  p.nvstride = (int)nv;
  const size_t words = nv * 22 + 2 * (size_t)p.nsb * n + 128 + 4 * (size_t)row_tiles; // + one counter per 128-row CTA tile
  cudaError_t e = cudaMalloc(&w.block, sizeof(double) * words);
  if (e != cudaSuccess)
    return qn_fail(QN_ERR_CUDA, "cudaMalloc failed for the streaming work space: %s", cudaGetErrorString(e));
  double* base = (double*)w.block;
  p.base = base;
  for (int g = 0; g < kMaxShards; ++g) p.peer_base[g] = base;
  p.x = base; p.g = base + nv; p.gp = base + 2 * nv; p.s = base + 3 * nv; p.y = base + 4 * nv;
  p.nu = base + 5 * nv; p.w = base + 6 * nv; p.xs = base + 7 * nv; p.tmp = base + 8 * nv;
  p.dvec = base + 12 * nv; // tmp spans 8..11 (three staged arrays when shared memory is too small)
  p.ps = base + 13 * nv; p.pu = base + 14 * nv;
  p.tvec = base + 15 * nv; p.wvec = base + 16 * nv; // and their odd-iteration copies at 17, 18 (row-sharded mode)
  p.tmp2 = base + 19 * nv;
  p.partial = base + 22 * nv;
  p.partial2 = p.partial + (size_t)p.nsb * n;
  double* tail = p.partial2 + (size_t)p.nsb * n;
  p.sc = reinterpret_cast<StreamScalars*>(tail);          // 32 doubles reserved
  p.flags = reinterpret_cast<unsigned int*>(tail + 32);   // 2·kMaxShards words
  p.tiles_done = reinterpret_cast<unsigned int*>(tail + 48);
  p.tile_count = reinterpret_cast<unsigned int*>(tail + 64);
  static_assert(sizeof(StreamScalars) <= 32 * sizeof(double), "scalars overflow their slot");
  // vec kernels stage their reductions in shared memory (n doubles; three arrays in one-pass mode)
  const size_t vec_smem = sizeof(double) * (onepass ? 3 : 1) * (((size_t)n + 1) & ~(size_t)1);
  p.stage_in_smem = vec_smem <= (size_t)ctx->prop.sharedMemPerBlockOptin - 1024;
  w.vec_smem = p.stage_in_smem ? vec_smem : 0;
  QN_TRY_REL(cudaMemsetAsync(w.block, 0, sizeof(double) * words, st));
  QN_TRY_REL(cudaMemcpyAsync(p.x, theta0, sizeof(double) * n, theta_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  StreamScalars init{};
  init.result = NAN;
  init.status = QN_MAX_ITER;
  init.spec = 1;
  if (max_iter <= 0) init.done = 1;
  QN_TRY_REL(cudaMemcpyAsync(p.sc, &init, sizeof init, cudaMemcpyHostToDevice, st));
  QN_TRY_REL(cudaStreamSynchronize(st)); // `init` is a stack variable
  return QN_OK;
}

int stream_capture(qn_context* ctx, StreamWork& w, cudaStream_t st) {
  const qn::StreamParams& p = w.p;
  const int n = p.n;
  QN_TRY_REL(cudaSetDevice(ctx->device));
  const bool deep = sweep_deep(ctx, p);
  QN_TRY_REL(cudaFuncSetAttribute(pass2_kernel(deep), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double) * 4 * p.sbcols)));
  if (pass1_pad())
    QN_TRY_REL(cudaFuncSetAttribute(pass1_kernel(deep), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(sizeof(double) * p.sbcols + pass1_pad())));
  if (w.vec_smem) {
    QN_TRY_REL(cudaFuncSetAttribute(qn::stream_vec_begin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    QN_TRY_REL(cudaFuncSetAttribute(qn::stream_vec_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    QN_TRY_REL(cudaFuncSetAttribute(qn::stream_vec_end, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    QN_TRY_REL(cudaFuncSetAttribute(qn::stream_vec_onepass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
  }
  if (!getenv("QN_STREAM_NO_CLUSTER")) {
    // the cluster vector kernel: a thread owns up to 3 Rosenbrock pairs; a rank stages n doubles
    const int pairs = n >> 1, cap = qn::kCluCtas * qn::kCluThreads;
    const int ppt = pairs <= cap ? 1 : (pairs + cap - 1) / cap;
    const size_t smem = sizeof(double) * (((size_t)n + 1) & ~(size_t)1);
    if (ppt <= 3 && smem + 1024 <= (size_t)ctx->prop.sharedMemPerBlockOptin) {
      w.clu_ppt = ppt;
      w.clu_smem = smem;
      QN_TRY_REL(cudaFuncSetAttribute(cluster_kernel(ppt, p.onepass != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (!p.onepass) w.p.fused2 = 1;
    }
  }
  // one graph = kItersPerGraph iterations: with several GPUs driven by one host thread the graph
  // launches themselves (≈ 6 µs each) would otherwise bound the iteration rate
  QN_TRY_REL(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = QN_OK;
  for (int k = 0; k < kItersPerGraph && rc == QN_OK; ++k) rc = enqueue_iteration(ctx, w, st);
  const cudaError_t ce = cudaStreamEndCapture(st, &w.graph);
  if (rc != QN_OK) return rc;
  if (ce != cudaSuccess) return qn_fail(QN_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
  QN_TRY_REL(cudaGraphInstantiate(&w.exec, w.graph, 0));
  if (p.onepass || p.fused2) qn::stream_vec_begin<<<1, qn::kVecThreads, w.vec_smem, st>>>(p); // head of iteration 1
  QN_TRY_REL(cudaGetLastError());
  return QN_OK;
}

// Replay the iteration graph on every shard until the device says done (or exactly fixed_iters
// times, timed on shard 0, the first two untimed).  The shards advance in lock step through their
// flags; shard 0's scalars are everybody's.
int stream_run(qn_context* const* ctxs, StreamWork* ws, cudaStream_t* sts, int G, int max_iter,
               int fixed_iters, float* ms_per_iter_out, qn::StreamScalars& host) {
  auto launch_all = [&]() -> int {
    for (int g = 0; g < G; ++g) {
      if (G > 1) QN_TRY_REL(cudaSetDevice(ctxs[g]->device));
      QN_TRY_REL(cudaGraphLaunch(ws[g].exec, sts[g]));
    }
    return QN_OK;
  };
  int rc;
  if (fixed_iters > 0) {
    // benchmark mode: one untimed graph (kItersPerGraph iterations), then enough timed graphs to
    // reach fixed_iters; the iteration count of the timed region is read back from the device
    QN_TRY_REL(cudaSetDevice(ctxs[0]->device));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    qn::StreamScalars warm_sc{};
    if ((rc = launch_all())) return rc;
    QN_TRY_REL(cudaSetDevice(ctxs[0]->device));
    QN_TRY_REL(cudaMemcpyAsync(&warm_sc, ws[0].p.sc, sizeof warm_sc, cudaMemcpyDeviceToHost, sts[0]));
    cudaEventRecord(e0, sts[0]);
    const int graphs = std::max(1, (fixed_iters - kItersPerGraph + kItersPerGraph - 1) / kItersPerGraph);
    for (int k = 0; k < graphs; ++k) if ((rc = launch_all())) return rc;
    QN_TRY_REL(cudaSetDevice(ctxs[0]->device));
    cudaEventRecord(e1, sts[0]);
    QN_TRY_REL(cudaMemcpyAsync(&host, ws[0].p.sc, sizeof host, cudaMemcpyDeviceToHost, sts[0]));
    for (int g = G - 1; g >= 0; --g) {
      QN_TRY_REL(cudaSetDevice(ctxs[g]->device));
      QN_TRY_REL(cudaStreamSynchronize(sts[g]));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const int timed = host.it - warm_sc.it;
    if (ms_per_iter_out) *ms_per_iter_out = timed > 0 ? ms / (float)timed : 0.0f;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return QN_OK;
  }
  // replay kItersPerGraph iterations at a time, then poll the done flag
  int launched = 0;
  for (;;) {
    if ((rc = launch_all())) return rc;
    launched += kItersPerGraph;
    QN_TRY_REL(cudaSetDevice(ctxs[0]->device));
    QN_TRY_REL(cudaMemcpyAsync(&host, ws[0].p.sc, sizeof host, cudaMemcpyDeviceToHost, sts[0]));
    QN_TRY_REL(cudaStreamSynchronize(sts[0]));
    if (host.done || launched >= max_iter) break;
  }
  for (int g = G - 1; g >= 1; --g) {
    QN_TRY_REL(cudaSetDevice(ctxs[g]->device));
    QN_TRY_REL(cudaStreamSynchronize(sts[g]));
  }
  QN_TRY_REL(cudaSetDevice(ctxs[0]->device));
  return QN_OK;
}

}  // namespace

// One large problem, device-resident x0 (n doubles at theta0).  Returns when finished.
int qn_stream_solve(qn_context* ctx, int n, const double* d_theta0, const qn_backtracking* ls,
                    double tol, int max_iter, double* d_x_out, double* d_g_out, double* d_h_out,
                    int ldh_out, double* ell, int* iters, int* fevals, int* status, int ntrace,
                    double* d_trace, cudaStream_t st, float* ms_per_iter_out, int fixed_iters,
                    int onepass) {
  using namespace qn;
  StreamWork w;
  StreamScalars host{};
  host.result = NAN;
  host.status = QN_MAX_ITER;
  int rc = stream_setup(ctx, w, n, d_theta0, false, ls, tol, max_iter, ntrace, d_trace, onepass, 0, 1, st);
  if (rc == QN_OK && max_iter > 0) rc = stream_capture(ctx, w, st);
  if (rc == QN_OK && max_iter > 0) rc = stream_run(&ctx, &w, &st, 1, max_iter, fixed_iters, ms_per_iter_out, host);
  if (rc != QN_OK) { stream_release(w); return rc; }
  const StreamParams& p = w.p;
  ctx->last_path = 2;
  ctx->last_grid = p.ntiles * p.nsb;
  ctx->last_block = w.pass_threads;
  ctx->last_smem = (int)(sizeof(double) * 4 * p.sbcols);
  {
    const int its = host.it + (host.done ? 1 : 0);
    ctx->last_launches = onepass ? 1 + 2 * its : p.fused2 ? 1 + 3 * its : 6 * its;
  }
  auto finish = [&]() -> int {
    if (d_h_out && host.need_reset) // the solve ended between a reset decision (QN.jl:272-280) and the next sweep
      qn::stream_identity<true><<<ctx->prop.multiProcessorCount * 8, 256, 0, st>>>(p);
    if (d_x_out) QN_TRY_REL(cudaMemcpyAsync(d_x_out, p.x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    // which ∇ the reference's canonical slot shows after a failure: see qn_resident.cuh
    const double* g_src = (host.status == QN_CONVERGED || (host.exit_it & 1)) ? p.g : p.gp;
    if (d_g_out) QN_TRY_REL(cudaMemcpyAsync(d_g_out, g_src, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    if (d_h_out && onepass) { // the stored matrix lags one update behind: apply it
      const dim3 grid(p.ntiles, p.nsb);
      qn::stream_flush_pending<<<grid, w.pass_threads, 0, st>>>(p);
    }
    if (d_h_out)
      QN_TRY_REL(cudaMemcpy2DAsync(d_h_out, sizeof(double) * ldh_out, p.H, sizeof(double) * p.ld,
                                   sizeof(double) * n, n, cudaMemcpyDeviceToDevice, st));
    QN_TRY_REL(cudaStreamSynchronize(st));
    return QN_OK;
  };
  rc = finish();
  stream_release(w);
  if (rc != QN_OK) return rc;
  if (ell) *ell = host.result;
  if (iters) *iters = host.iters;
  if (fevals) *fevals = host.fevals;
  if (status) *status = host.status;
  return QN_OK;
}

// Row-sharded large-n mode (BASELINE.json configs[4]): ONE problem on G GPUs of one
// process.  GPU g holds the rows of its row tiles of B⁻¹ and a replica of every vector; each sweep's
// row sums are stored straight into all replicas over NVLink (peer stores) and announced with
// per-shard epoch flags, so the only exchange per sweep is n doubles per vector and no host or NCCL
// call sits inside an iteration.  All O(n) work — the rule-C1 reductions, the line search, the
// objective — is replicated, so every GPU holds bit-identical scalars and the result is the
// single-GPU result bit for bit.  Host pointers in, host pointers out.
extern "C" int32_t qn_bfgs_optimize_rowsharded_f64(
    qn_context* const* ctxs, int32_t nctx, const qn_model* model, int32_t n, const double* theta0,
    const qn_backtracking* ls, double tol, int32_t max_iter, double* x_out, double* g_out, double* h_out,
    int32_t ldh, double* ell_out, int32_t* iters_out, int32_t* fevals_out, int32_t* status_out,
    int32_t ntrace, double* trace_out, uint32_t flags, int32_t bench_iters, double* ms_per_iter) {
  using namespace qn;
  if (!ctxs || nctx < 1 || nctx > kMaxShards) return qn_fail(QN_ERR_INVALID_ARGUMENT, "need 1..%d contexts", kMaxShards);
  if (!model || n < 2 || !theta0 || !ls) return qn_fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  if (model->kind != QN_OBJ_ROSENBROCK)
    return qn_fail(QN_ERR_UNREGISTERED_OBJECTIVE, "objective kind %d has no registered device implementation on the "
                   "streaming path; only Rosenbrock has one", model->kind);
  if (h_out && ldh < n) return qn_fail(QN_ERR_INVALID_ARGUMENT, "ldh < n");
  const int onepass = (flags & QN_FLAG_ONE_PASS) ? 1 : 0;
  const int row_tiles = (n + kShardRows - 1) / kShardRows;
  const int G = std::min((int)nctx, row_tiles); // a shard owns at least one 512-row unit
  for (int g = 0; g < G; ++g) if (!ctxs[g]) return qn_fail(QN_ERR_INVALID_ARGUMENT, "NULL context %d", g);
  // peer access between every pair of distinct devices
  for (int a = 0; a < G; ++a)
    for (int b = 0; b < G; ++b) {
      if (ctxs[a]->device == ctxs[b]->device) continue;
      int can = 0;
      QN_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device));
      if (!can) return qn_fail(QN_ERR_CUDA, "device %d cannot access device %d (no peer path)", ctxs[a]->device, ctxs[b]->device);
      QN_CUDA_TRY(cudaSetDevice(ctxs[a]->device));
      const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return qn_fail(QN_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctxs[a]->device, ctxs[b]->device, cudaGetErrorString(e));
      (void)cudaGetLastError();
    }
  StreamWork ws[kMaxShards];
  cudaStream_t sts[kMaxShards];
  double* d_trace = nullptr;
  auto release_all = [&]() {
    for (int g = 0; g < G; ++g) { cudaSetDevice(ctxs[g]->device); stream_release(ws[g]); }
    cudaSetDevice(ctxs[0]->device);
    if (d_trace) cudaFree(d_trace);
  };
  int rc = QN_OK;
  if (ntrace > 0 && trace_out) {
    QN_CUDA_TRY(cudaSetDevice(ctxs[0]->device));
    QN_CUDA_TRY(cudaMalloc(&d_trace, sizeof(double) * (size_t)ntrace * n));
    QN_CUDA_TRY(cudaMemset(d_trace, 0, sizeof(double) * (size_t)ntrace * n));
  }
  for (int g = 0; g < G && rc == QN_OK; ++g) {
    sts[g] = ctxs[g]->stream;
    rc = stream_setup(ctxs[g], ws[g], n, theta0, true, ls, tol, max_iter, g == 0 ? ntrace : 0,
                      g == 0 ? d_trace : nullptr, onepass, g, G, sts[g]);
  }
  for (int g = 0; g < G && rc == QN_OK; ++g) {
    for (int k = 0; k < G; ++k) ws[g].p.peer_base[k] = ws[k].p.base;
    if (max_iter > 0) rc = stream_capture(ctxs[g], ws[g], sts[g]);
  }
  StreamScalars host{};
  host.result = NAN;
  host.status = QN_MAX_ITER;
  float ms = 0;
  if (rc == QN_OK && max_iter > 0) rc = stream_run(ctxs, ws, sts, G, max_iter, bench_iters, &ms, host);
  if (rc != QN_OK) { release_all(); return rc; }
  auto finish = [&]() -> int {
    const StreamParams& p0 = ws[0].p;
    QN_TRY_REL(cudaSetDevice(ctxs[0]->device));
    if (x_out) QN_TRY_REL(cudaMemcpyAsync(x_out, p0.x, sizeof(double) * n, cudaMemcpyDeviceToHost, sts[0]));
    const double* g_src = (host.status == QN_CONVERGED || (host.exit_it & 1)) ? p0.g : p0.gp;
    if (g_out) QN_TRY_REL(cudaMemcpyAsync(g_out, g_src, sizeof(double) * n, cudaMemcpyDeviceToHost, sts[0]));
    if (trace_out && d_trace)
      QN_TRY_REL(cudaMemcpyAsync(trace_out, d_trace, sizeof(double) * (size_t)ntrace * n, cudaMemcpyDeviceToHost, sts[0]));
    QN_TRY_REL(cudaStreamSynchronize(sts[0]));
    if (h_out) {
      for (int g = 0; g < G; ++g) { // every shard returns its slab of rows
        const StreamParams& p = ws[g].p;
        QN_TRY_REL(cudaSetDevice(ctxs[g]->device));
        if (host.need_reset) qn::stream_identity<true><<<ctxs[g]->prop.multiProcessorCount * 8, 256, 0, sts[g]>>>(p);
        if (onepass) {
          const dim3 grid(p.ntiles, p.nsb);
          qn::stream_flush_pending<<<grid, ws[g].pass_threads, 0, sts[g]>>>(p);
        }
        const int row0 = p.row0, rows = p.rows;
        QN_TRY_REL(cudaMemcpy2DAsync(h_out + row0, sizeof(double) * ldh, p.H, sizeof(double) * p.ld,
                                     sizeof(double) * rows, n, cudaMemcpyDeviceToHost, sts[g]));
        QN_TRY_REL(cudaStreamSynchronize(sts[g]));
      }
    }
    return QN_OK;
  };
  rc = finish();
  for (int g = 0; g < G; ++g) {
    ctxs[g]->last_path = 2;
    ctxs[g]->last_grid = ws[g].p.ntiles * ws[g].p.nsb;
    ctxs[g]->last_block = ws[g].pass_threads;
    const int its = host.it + (host.done ? 1 : 0);
    ctxs[g]->last_launches = onepass ? 1 + 2 * its : ws[g].p.fused2 ? 1 + 3 * its : 6 * its;
  }
  release_all();
  if (rc != QN_OK) return rc;
  if (ell_out) *ell_out = host.result;
  if (iters_out) *iters_out = host.iters;
  if (fevals_out) *fevals_out = host.fevals;
  if (status_out) *status_out = host.status;
  if (ms_per_iter) *ms_per_iter = ms;
  return QN_OK;
}

// Time `iters` iterations of the streaming path on a fresh n-dimensional Rosenbrock problem
// (x0 = a fixed pseudo-random start) and report ms per iteration and the achieved ALGORITHMIC
// bandwidth (16 n² + 96 n bytes per iteration).
extern "C" int32_t qn_stream_bench_f64(qn_context* ctx, int32_t n, int32_t iters, int32_t order,
                                       uint32_t flags, double* ms_per_iter, double* gbs) {
  if (!ctx || !ms_per_iter || !gbs || n < 2 || iters < 3) return qn_fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  double* h = (double*)malloc(sizeof(double) * n);
  unsigned long long state = 88172645463325252ULL;
  for (int i = 0; i < n; ++i) { // xorshift + sum of uniforms: a deterministic roughly normal start
    double acc = 0;
    for (int k = 0; k < 12; ++k) {
      state ^= state << 13; state ^= state >> 7; state ^= state << 17;
      acc += (double)(state >> 11) / 9007199254740992.0;
    }
    h[i] = acc - 6.0;
  }
  double* d = nullptr;
  QN_CUDA_TRY(cudaMalloc(&d, sizeof(double) * n));
  cudaMemcpy(d, h, sizeof(double) * n, cudaMemcpyHostToDevice);
  free(h);
  qn_backtracking ls;
  qn_backtracking_default(&ls, order);
  float ms = 0;
  int it = 0, fe = 0, stt = 0;
  double ell = 0;
  int rc = qn_stream_solve(ctx, n, d, &ls, 1e-8, iters, nullptr, nullptr, nullptr, 0, &ell, &it, &fe, &stt,
                           0, nullptr, ctx->stream, &ms, iters, (flags & QN_FLAG_ONE_PASS) ? 1 : 0);
  cudaFree(d);
  if (rc) return rc;
  *ms_per_iter = ms;
  *gbs = (16.0 * n * (double)n + 96.0 * n) / (ms * 1e-3) / 1e9;
  return QN_OK;
}
