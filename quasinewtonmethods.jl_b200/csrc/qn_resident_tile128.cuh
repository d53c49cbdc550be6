// qn_resident_tile128.cuh — n = 128 Rosenbrock (BASELINE.json configs[2], configs[3]): the leader/helper
// tile kernel of qn_resident_tile.cuh at the size where B⁻¹ (128 KB) fills half of the SM's tensor memory.
//
// One CTA of 256 threads = ONE problem.  Thread (q, b) owns rows 4q..4q+3 of the rule-C2 column block b
// (16 columns): a 4×16 fp64 tile = 128 columns of its own tensor-memory lane (warps 0-3 on the lanes'
// first 128 columns, warps 4-7 on the next 128: 256 columns per CTA, two CTAs = two problems per SM).
// A warp is one column block — its 32 lanes read the block's column operands as one broadcast.
//
// Warp 0 is the LEADER: it holds every vector as FOUR elements per lane (element l + 32k in slot k —
// exactly the lane accumulators of rule C1) and runs the whole O(n) part alone: objective, line
// search, dot products, butterflies.  At n = 128 a Rosenbrock pair (i, i + 64) lives in ONE lane, so
// the objective needs no exchange at all.  The other seven warps only sweep their tile (pass 1,
// pass 2) between CTA barriers.  The kernel it replaces (qn_resident.cuh, `Team` protocol: one thread
// per element, every reduction staged through shared memory behind a CTA barrier) ran at 7.2e7
// iterations/s with FP64 37 % busy and its warps waiting on shared-memory scoreboards and barriers.
// Same canonical order: the same bits as the oracle and as that kernel.
#pragma once

#include "qn_resident_tile.cuh"

namespace qn {

// RT rows per thread = elements per lane of the leader: n = 32·RT (RT = 4: n = 128, two problems per SM;
// RT = 5: n = 160, 480 of the 512 columns, one problem per SM; RT = 3: n = 96).
template <int RT_>
struct TileW {
  static constexpr int RT = RT_, N = 32 * RT_, CT = 16, NB = N / 16, THREADS = 32 * NB, EPL = RT_;
  static constexpr int kTmemCols = 32 * RT;                                // 32-bit columns of one thread's tile
  static constexpr int kBands = (NB + 3) / 4;                              // groups of four warps share the 128 lanes
  static constexpr int kAlloc = kBands * kTmemCols <= 128 ? 128 : (kBands * kTmemCols <= 256 ? 256 : 512);
  static constexpr int kCtasPerSm = 512 / kAlloc;
  // shared memory (doubles)
  static constexpr int kRec = 4;                      // {c₁s, -u, s, ∇} per column
  static constexpr int oRec = 0;                      // [128 columns][4]
  static constexpr int oYb = oRec + N * kRec;         // y by column [128]
  static constexpr int oS = oYb + N;                  // s by row [128]
  static constexpr int oNu = oS + N;                  // -u by row [128]
  static constexpr int oPart1 = oNu + N;              // block partials [8 blocks][128 rows]
  static constexpr int oPart2 = oPart1 + NB * N;
  static constexpr int oD0 = oPart2 + NB * N;         // cooperative objectives: the point(s) being evaluated, θ-μ [N] each
  static constexpr int oD1 = oD0 + N;
  static constexpr int oCmd = oD1 + N;                // command word
  static constexpr int oEnd = oCmd + 2;               // (+ 2 doubles: the work-queue slot)
  static constexpr int kSmemDoubles = oEnd + 2;
};

using Tile128 = TileW<4>;
template <int R>
struct VecR { double v[R]; };

// Rosenbrock on the leader's layout: lane l holds θ[l + 32k] in slot k.  The pairs (i, i + n/2) of
// test/runtests.jl:8-9: for even RT both elements sit in ONE lane (slots k and k + RT/2) and the
// objective needs no exchange; for odd RT (n/2 = 32·hk + 16) the partner sits in lane ^ 16, slot
// k + hk (+ 1 for lanes >= 16), one 64-bit shuffle per pair slot each way.
// Returns this lane's rule-C1 accumulator of the @simd sum: Σ_k term[l + 32k] in ascending k.
template <int RT, bool GRAD>
__device__ __forceinline__ double rosenbrock_w_terms(const int lane, const VecR<RT>& x, VecR<RT>& g) {
  double acc = 0.0;
  if constexpr (RT % 2 == 0) {
#pragma unroll
    for (int k = 0; k < RT / 2; ++k) {
      const double t1 = x.v[k], t2 = x.v[k + RT / 2];
      const double a = t2 - t1 * t1;
      const double b = 1.0 - t1;
      const double term = 100.0 * (a * a) + b * b; // runtests.jl:9,23
      acc = (k == 0) ? term : acc + term;
      if (GRAD) {
        g.v[k] = (400.0 * a) * t1 + 2.0 * b;        // ∇[i],    runtests.jl:24
        g.v[k + RT / 2] = 200.0 * (t1 * t1 - t2);   // ∇[i+N],  runtests.jl:25
      }
    }
  } else {
    constexpr int hk = (RT - 1) / 2;
    const bool low = lane < 16;
#pragma unroll
    for (int k = 0; k <= hk; ++k) {
      // what lane ^ 16 needs from me for ITS pair slot k: my slot k + hk (+ 1 when I am a low lane)
      const double send = low ? (k + hk + 1 < RT ? x.v[k + hk + 1 < RT ? k + hk + 1 : 0] : 0.0) : x.v[k + hk];
      const double t2 = __shfl_xor_sync(0xffffffffu, send, 16);
      const bool have = (k < hk) || low;          // element lane + 32k is a θ1 (index < n/2)
      const double t1 = x.v[k];
      const double a = t2 - t1 * t1;
      const double b = 1.0 - t1;
      const double term = have ? 100.0 * (a * a) + b * b : 0.0;
      acc = (k == 0) ? term : acc + term;
      if (GRAD) {
        const double g2 = 200.0 * (t1 * t1 - t2); // ∇[i+N] goes back along the same route
        const double r = __shfl_xor_sync(0xffffffffu, g2, 16);
        if (have) g.v[k] = (400.0 * a) * t1 + 2.0 * b;
        if (low) { if (k + hk + 1 < RT) g.v[k + hk + 1 < RT ? k + hk + 1 : 0] = r; }
        else g.v[k + hk] = r;
      }
    }
  }
  return acc;
}

template <int RT>
__device__ __forceinline__ double dot_w(const VecR<RT>& a, const VecR<RT>& b) { // lane accumulator of rule C1
  double p = a.v[0] * b.v[0];
#pragma unroll
  for (int k = 1; k < RT; ++k) p = p + a.v[k] * b.v[k];
  return p;
}

// One tile column (RT rows) of the calling thread's tensor-memory lane.  The first four (RT = 3: two)
// rows of the 16 tile columns are one aligned group of columns each (.x8 / .x4 loads), the last row
// of every tile column follows behind them (.x2): a load never starts off its natural alignment.
template <int RT>
struct TileCol { double d[RT]; };
template <int RT>
struct TileColRaw { uint32_t w[2 * RT]; }; // as loaded; paired into doubles only after tcgen05.wait::ld
__device__ __forceinline__ void tm_load1_raw(uint32_t& lo, uint32_t& hi, const uint32_t addr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(lo), "=r"(hi) : "r"(addr));
}
// 32-bit column (relative to the thread's band) of row r of tile column k
template <int RT>
__device__ __forceinline__ constexpr uint32_t tile_word(const int k, const int r) {
  constexpr int G = (RT >= 4) ? 4 : 2;                 // rows in the aligned group
  return r < G ? (uint32_t)(2 * G * k + 2 * r) : (uint32_t)(2 * G * 16 + 2 * (RT - G) * k + 2 * (r - G));
}
template <int RT>
__device__ __forceinline__ void tile_col_load(TileColRaw<RT>& c, const uint32_t base, const int k) {
  static_assert(RT == 3 || RT == 4 || RT == 5, "rows per thread");
  if constexpr (RT >= 4) {
    TmRaw<4> q;
    tm_load_raw(q, base + tile_word<RT>(k, 0));
#pragma unroll
    for (int i = 0; i < 8; ++i) c.w[i] = q.w[i];
    if constexpr (RT == 5) tm_load1_raw(c.w[8], c.w[9], base + tile_word<RT>(k, 4));
  } else {
    TmRaw<2> q;
    tm_load_raw(q, base + tile_word<RT>(k, 0));
#pragma unroll
    for (int i = 0; i < 4; ++i) c.w[i] = q.w[i];
    tm_load1_raw(c.w[4], c.w[5], base + tile_word<RT>(k, 2));
  }
}
template <int RT>
__device__ __forceinline__ void tile_col_unpack(TileCol<RT>& c, const TileColRaw<RT>& r) { // after tm_wait_load()
#pragma unroll
  for (int k = 0; k < RT; ++k) c.d[k] = __hiloint2double((int)r.w[2 * k + 1], (int)r.w[2 * k]);
}

// The objective of this kernel (logdensity / ∂logdensity!, QN.jl:8-9).
//  - WRosenbrock: evaluated by the leader alone, in its own lanes.
//  - WGaussian (build-defined): ℓ = -½ dᵀΛd, ∇ = -Λd, d = θ-μ.  z = Λd is a sweep of the
//    SAME shape as pass 1 — thread (q, b) the rows RT·q… of column block b, Λ read through L1 (128 KB at
//    n = 128, shared by the SM's CTAs and by every problem) — so all warps take part: the leader
//    publishes d (for the accepted point, and for the trial point when they differ: one sweep serves
//    both), commands an EVAL, and adds the block partials in ascending order (rule C2), the oracle's
//    gauss_eval order.  One GEMV gives the value AND the gradient.
struct WRosenbrock {
  static constexpr bool kCoop = false;
  struct Params {};
  __device__ __forceinline__ static double finish(const double sum) { return 0.0 - sum; } // runtests.jl:6-10: s -= term
};
struct WGaussian {
  static constexpr bool kCoop = true;
  struct Params { const double* mu; const double* lam; };
  __device__ __forceinline__ static double finish(const double q) { return -0.5 * q; }
};
constexpr int kLhEval = 8, kLhEval2 = 16; // cooperative objective: evaluate at one / two published points

template <int RT_, class WObj>
__global__ void __launch_bounds__(TileW<RT_>::THREADS, TileW<RT_>::kCtasPerSm)
qn_resident_tilew_kernel(const __grid_constant__ BatchParams p, const __grid_constant__ typename WObj::Params op) {
  using C = TileW<RT_>;
  using Vec = VecR<RT_>;
  constexpr int N = C::N, RT = C::RT, CT = C::CT, NB = C::NB;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint32_t tmem_base_s;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "l"((unsigned long long)__cvta_generic_to_shared(&tmem_base_s)), "n"(C::kAlloc) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t tmem = __reduce_max_sync(0xffffffffu, tmem_base_s + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)C::kTmemCols);
  const int b = warp;        // column block of this thread's tile
  const int q = lane;        // row group
  const int row0 = RT * q, col0 = CT * b;
  double* rec = smem + C::oRec;
  double* yb_arr = smem + C::oYb;
  double* s_arr = smem + C::oS;
  double* nu_arr = smem + C::oNu;
  double* part1 = smem + C::oPart1;
  double* part2 = smem + C::oPart2;
  volatile int* cmd = reinterpret_cast<volatile int*>(smem + C::oCmd);
  volatile long long* next = reinterpret_cast<volatile long long*>(smem + C::kSmemDoubles - 2);
  const bool leader = (warp == p.lead_warp);
  constexpr int NCH = CT; // one tile column (RT rows) per step

  auto pass1 = [&]() { // t[i] = Σ_k y[col0+k] H[row0+i, col0+k], QN.jl:44-49
    double tt[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) tt[i] = 0.0;
    const double2* yb2 = reinterpret_cast<const double2*>(yb_arr + col0);
    double2 yy[CT / 2];
#pragma unroll
    for (int k = 0; k < CT / 2; ++k) yy[k] = yb2[k];
    TileColRaw<RT> ring[2];
    tile_col_load<RT>(ring[0], tmem, 0);
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      const double yk = (k & 1) ? yy[k >> 1].y : yy[k >> 1].x;
      tm_wait_load();
      TileCol<RT> cur;
      tile_col_unpack<RT>(cur, ring[k & 1]);
      if (k + 1 < NCH) tile_col_load<RT>(ring[(k + 1) & 1], tmem, k + 1);
#pragma unroll
      for (int r = 0; r < RT; ++r) tt[r] = fma(yk, cur.d[r], tt[r]);
    }
    double* pp = part1 + b * N + row0;
#pragma unroll
    for (int r = 0; r < RT; ++r) pp[r] = tt[r];
  };
  auto pass2 = [&]() { // H ← H + c₁ssᵀ - usᵀ - suᵀ fused with d = H⁺∇_new, QN.jl:56-67
    double sc_[RT], nuc[RT], acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) { sc_[r] = s_arr[row0 + r]; nuc[r] = nu_arr[row0 + r]; }
#pragma unroll
    for (int i = 0; i < RT; ++i) acc[i] = 0.0;
    const double* rb_ = rec + col0 * C::kRec;
    TileColRaw<RT> ring[2];
    tile_col_load<RT>(ring[0], tmem, 0);
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      const double2 wn = *reinterpret_cast<const double2*>(rb_ + k * C::kRec);
      const double2 sg = *reinterpret_cast<const double2*>(rb_ + k * C::kRec + 2);
      tm_wait_load();
      TileCol<RT> cur;
      tile_col_unpack<RT>(cur, ring[k & 1]);
      if (k + 1 < NCH) tile_col_load<RT>(ring[(k + 1) & 1], tmem, k + 1);
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        double hh = cur.d[r];
        hh = fma(wn.x, sc_[r], hh);  // + (c₁ s[r]) s[c]
        hh = fma(wn.y, sc_[r], hh);  // - u[r] s[c]
        hh = fma(nuc[r], sg.x, hh);  // - u[c] s[r]
        cur.d[r] = hh;
        acc[r] = fma(hh, sg.y, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < RT; ++r) tm_store1(tmem + tile_word<RT>(k, r), cur.d[r]); // one element per store: no staging moves
    }
    tm_wait_store();
    double* pp = part2 + b * N + row0;
#pragma unroll
    for (int r = 0; r < RT; ++r) pp[r] = acc[r];
  };
  auto reset_tile = [&]() { // B⁻¹ ← I, QN.jl:144-148
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
#pragma unroll
      for (int r = 0; r < RT; ++r) tm_store1(tmem + tile_word<RT>(k, r), (row0 + r == col0 + k) ? 1.0 : 0.0);
    }
    tm_wait_store();
  };
  auto write_h = [&](const long long j) {
    if (!p.h_out) return;
    double* ho = p.h_out + j * p.h_stride;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      TileColRaw<RT> hr;
      tile_col_load<RT>(hr, tmem, k);
      tm_wait_load();
      TileCol<RT> hc;
      tile_col_unpack<RT>(hc, hr);
#pragma unroll
      for (int r = 0; r < RT; ++r) ho[(row0 + r) + (long long)p.ldh * (col0 + k)] = hc.d[r];
    }
  };
  // Σ_b part[b][e] in ascending block order (rule C2) for the leader's four elements
  auto combine = [&](const double* part) {
    Vec T;
#pragma unroll
    for (int k = 0; k < RT; ++k) {
      const int e = lane + 32 * k;
      double acc = part[e];
#pragma unroll
      for (int bb = 1; bb < NB; ++bb) acc = acc + part[bb * N + e];
      T.v[k] = acc;
    }
    return T;
  };

  double* dvec0 = smem + C::oD0;
  double* dvec1 = smem + C::oD1;
  // cooperative objective: this thread's block partials of z = Λd for the published point(s) -> part1 (part2)
  auto eval_sweep = [&](const bool second) {
    if constexpr (WObj::kCoop) {
      double zo[RT], zn[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) { zo[r] = 0.0; zn[r] = 0.0; }
      const double* lamp = op.lam + row0 + (size_t)N * col0;
#pragma unroll 8
      for (int k = 0; k < CT; ++k) {
        double lam[RT];
        if constexpr (RT == 4) { // the thread's four rows of column k: 32 contiguous, 32-byte aligned bytes = two full-sector LDG.128
          const double2 l01 = __ldg(reinterpret_cast<const double2*>(lamp + (size_t)N * k));
          const double2 l23 = __ldg(reinterpret_cast<const double2*>(lamp + (size_t)N * k + 2));
          lam[0] = l01.x; lam[1] = l01.y; lam[2] = l23.x; lam[3] = l23.y;
        } else {
#pragma unroll
          for (int r = 0; r < RT; ++r) lam[r] = __ldg(lamp + (size_t)N * k + r);
        }
        const double d0 = dvec0[col0 + k];
#pragma unroll
        for (int r = 0; r < RT; ++r) zo[r] = fma(lam[r], d0, zo[r]);
        if (second) {
          const double d1 = dvec1[col0 + k];
#pragma unroll
          for (int r = 0; r < RT; ++r) zn[r] = fma(lam[r], d1, zn[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        part1[b * N + row0 + r] = zo[r];
        if (second) part2[b * N + row0 + r] = zn[r];
      }
    }
  };

  if (t == 0) stamp_start(p);
  for (;;) {
    __syncthreads(); // the previous problem's shared-memory traffic is finished
    if (t == 0) *next = (long long)atomicAdd(p.counter, 1ULL);
    __syncthreads();
    const long long j = *next;
    if (j >= p.batch) break;

    if (!leader) { // ---- helpers: their part of the sweeps, between CTA barriers
      for (;;) {
        __syncthreads();                 // (A) a command and its operands are published
        const int c = *cmd;
        if (c & kLhReset) reset_tile();
        if (c & kLhExit) break;
        if (c & kLhEval) {               // cooperative objective
          eval_sweep((c & kLhEval2) != 0);
          __syncthreads();               // (E) the partials of z
          continue;
        }
        pass1();
        __syncthreads();                 // (B) pass-1 partials
        __syncthreads();                 // (C) c₁s, -u published
        pass2();
        __syncthreads();                 // (D) pass-2 partials
      }
      write_h(j);
      continue;
    }

    // ---- leader
    Vec x, g, s, y, gp;
    double l0, sy = 0.0;
    {
      const double* th = p.theta0 + j * p.theta_stride; // copyto!(x_old, x), QN.jl:248
#pragma unroll
      for (int k = 0; k < RT; ++k) { x.v[k] = th[lane + 32 * k]; s.v[k] = y.v[k] = gp.v[k] = 0.0; }
    }
    Vec muv;
#pragma unroll
    for (int k = 0; k < RT; ++k) muv.v[k] = 0.0;
    if constexpr (WObj::kCoop) {
#pragma unroll
      for (int k = 0; k < RT; ++k) muv.v[k] = __ldg(op.mu + lane + 32 * k);
    }
    // ∂logdensity! at xo (gradient to gout, lane accumulator of the ℓ-sum returned in acc_o) and, when
    // `second`, logdensity at xn (acc_n): ℓ = WObj::finish(rule-C1 sum of the accumulators)
    auto eval_pair = [&](const Vec& xo, const bool second, const Vec& xn, Vec& gout, double& acc_o, double& acc_n) {
      if constexpr (!WObj::kCoop) {
        acc_o = rosenbrock_w_terms<RT, true>(lane, xo, gout);
        if (second) { Vec unused; acc_n = rosenbrock_w_terms<RT, false>(lane, xn, unused); }
      } else {
        Vec d_o, d_n;
#pragma unroll
        for (int k = 0; k < RT; ++k) {
          d_o.v[k] = xo.v[k] - muv.v[k];
          dvec0[lane + 32 * k] = d_o.v[k];
          d_n.v[k] = 0.0;
          if (second) { d_n.v[k] = xn.v[k] - muv.v[k]; dvec1[lane + 32 * k] = d_n.v[k]; }
        }
        if (lane == 0) *cmd = kLhEval | (second ? kLhEval2 : 0);
        __syncthreads(); // (A)
        eval_sweep(second);
        __syncthreads(); // (E)
        const Vec z = combine(part1);
#pragma unroll
        for (int k = 0; k < RT; ++k) gout.v[k] = -z.v[k];
        acc_o = dot_w<RT>(d_o, z);
        if (second) { const Vec zn = combine(part2); acc_n = dot_w<RT>(d_n, zn); }
      }
    };
    {
      Vec unused_x; double acc, unused_acc;
      eval_pair(x, false, unused_x, g, acc, unused_acc);      // QN.jl:254 of iteration 1
      l0 = WObj::finish(warp_butterfly(acc));
    }
    const LsParams ls = p.ls;
    const double tol = p.tol;
    const int max_iter = p.max_iter;
    const int ntrace = p.trace != nullptr ? p.ntrace : 0;
    double result = __longlong_as_double(0x7ff8000000000000LL); // NaN, QN.jl:291
    int iters = 0, fevals = 0, status = QN_MAX_ITER;
    bool helper_reset = false;
    auto command = [&](const int c) {
      if (lane == 0) *cmd = c;
      __syncthreads(); // (A)
    };
    struct Spec { Vec g; double term_next, sy; };
    // a line-search trial at step a together with what the next iteration needs if it is accepted
    // (see spec_eval in qn_resident_tile.cuh): the trial point is fma(a, d, x), the accepted point x + a·d
    auto spec_eval = [&](const double a, const Vec& d, Spec& sp, double& term_trial) {
      Vec sn, xo, yn;
#pragma unroll
      for (int k = 0; k < RT; ++k) { sn.v[k] = a * d.v[k]; xo.v[k] = x.v[k] + sn.v[k]; }
      if (a == 1.0) {
        eval_pair(xo, false, xo, sp.g, sp.term_next, term_trial);
        term_trial = sp.term_next;
      } else {
        Vec xn;
#pragma unroll
        for (int k = 0; k < RT; ++k) xn.v[k] = fma(a, d.v[k], x.v[k]);
        eval_pair(xo, true, xn, sp.g, sp.term_next, term_trial);
      }
#pragma unroll
      for (int k = 0; k < RT; ++k) yn.v[k] = g.v[k] - sp.g.v[k];
      sp.sy = dot_w<RT>(sn, yn);
    };

    for (int it = 1; it <= max_iter; ++it) { // QN.jl:250,253
      ++fevals;
      bool small = true;
#pragma unroll
      for (int k = 0; k < RT; ++k) small = small && (fabs(g.v[k]) < tol);
      const bool all_small = __all_sync(0xffffffffu, small);
      if (!isfinite(l0) || all_small) {
        iters = it - 1;
        if (!isfinite(l0)) { status = QN_NONFINITE; if (!(it & 1)) g = gp; }
        else { status = QN_CONVERGED; result = l0; }
        break;
      }
      double m;
      Vec d;
      Spec sp;
      double lx1 = 0.0;
      bool have_trial = false;
      if (it == 1) {
        m = -1.0; // QN.jl:263-264
      } else {
        // ---- BFGS_update!  QN.jl:34-69: the operands of every column block
#pragma unroll
        for (int k = 0; k < RT; ++k) {
          const int e = lane + 32 * k;
          s_arr[e] = s.v[k];
          yb_arr[e] = y.v[k];
          *reinterpret_cast<double2*>(rec + e * C::kRec + 2) = make_double2(s.v[k], g.v[k]);
        }
        command(kLhPass | (helper_reset ? kLhReset : 0));
        helper_reset = false;
        const double rho = 1.0 / sy; // QN.jl:42 (no guard)
        pass1();
        __syncthreads(); // (B)
        {
          const Vec T = combine(part1);
          const double yHy = warp_butterfly(dot_w<RT>(T, y)); // Σ_c t[c] y[c], QN.jl:51
          const double c1 = fma(yHy, rho, 1.0) * rho;      // QN.jl:54
#pragma unroll
          for (int k = 0; k < RT; ++k) {
            const int e = lane + 32 * k;
            const double nu = -(T.v[k] * rho);             // -B⁻¹y[c] = -(t ρ), QN.jl:50
            nu_arr[e] = nu;
            *reinterpret_cast<double2*>(rec + e * C::kRec) = make_double2(c1 * s.v[k], nu); // (c₁ s[r]) once per r (rule C4)
          }
        }
        __syncthreads(); // (C)
        pass2();
        __syncthreads(); // (D)
        d = combine(part2); // the new direction d = H⁺∇_new; `s, y = y, s`, QN.jl:65,268
        double tterm;
        spec_eval(1.0, d, sp, tterm);
        m = dot_w<RT>(d, g);
        warp_butterfly3(m, tterm, sp.sy, lane);
        sp.term_next = tterm;
        lx1 = WObj::finish(tterm);
        have_trial = true;
      }
      if (m <= 0.0) { // QN.jl:272-280: H ← I, s ← ∇_new, m ← ∇ᵀ∇
        reset_tile();
        helper_reset = true;
        d = g;
        m = warp_butterfly(dot_w<RT>(g, g));
        have_trial = false;
      }
      double alpha;
      {
        auto trial = [&](const double a) {
          double tterm;
          spec_eval(a, d, sp, tterm);
          warp_butterfly3(tterm, sp.term_next, sp.sy, lane);
          ++fevals;
          return WObj::finish(tterm);
        };
        if (have_trial) ++fevals;
        alpha = backtracking_linesearch(ls, l0, m, trial, lx1, have_trial, lx1);
      }
      if (alpha == 0.0) { status = QN_LINESEARCH_FAILED; iters = it - 1; if (!(it & 1)) g = gp; break; } // QN.jl:284
#pragma unroll
      for (int k = 0; k < RT; ++k) { s.v[k] = alpha * d.v[k]; x.v[k] = x.v[k] + s.v[k]; } // update_state!, QN.jl:21-27
      if (it <= ntrace) {
        double* tr = p.trace + (j * p.ntrace + (it - 1)) * (long long)N;
#pragma unroll
        for (int k = 0; k < RT; ++k) tr[lane + 32 * k] = x.v[k];
      }
      if (it == max_iter) { iters = it; if (!(it & 1)) g = gp; break; }
      gp = g;
#pragma unroll
      for (int k = 0; k < RT; ++k) y.v[k] = g.v[k] - sp.g.v[k]; // y = ∇_old - ∇_new, QN.jl:38
      sy = sp.sy;
      g = sp.g;
      l0 = (alpha == 1.0) ? lx1 : WObj::finish(sp.term_next);
    }
    command(kLhExit | (helper_reset ? kLhReset : 0));
#pragma unroll
    for (int k = 0; k < RT; ++k) {
      if (p.x_out) p.x_out[j * p.x_stride + lane + 32 * k] = x.v[k];
      if (p.g_out) p.g_out[j * p.g_stride + lane + 32 * k] = g.v[k];
    }
    write_h(j);
    write_problem_scalars(p, j, lane, result, iters, fevals, status);
  }
  if (t == 0) stamp_exit(p, (int)blockIdx.x);
  __syncthreads(); // every warp is done with the tensor memory
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base_s), "n"(C::kAlloc) : "memory");
}

}  // namespace qn
