// qn_resident_tile3w.cuh — the n = 36..64 tile kernel with THREE warps per problem and a speculative pass 1.
//
// In the leader/helper form of qn_resident_tile.cuh the leader warp also owns half of the tile, so its
// own sweeps sit on the problem's critical path between the serial O(n) phases.  Here the tile belongs
// to TWO helper warps (64 threads, the 4×CT tiles of qn_resident_tile.cuh, in tensor memory) and a
// THIRD warp, the leader, owns nothing but the vectors.  That frees the one overlap the algorithm
// allows: the y of the NEXT update, y' = ∇ - ∇(x + d), is known as soon as the first line-search
// trial (a = 1) has been evaluated — long before the trial is judged (m, ℓ(trial), sᵀy butterfly, the
// Armijo test, update_state!, the convergence vote).  The leader publishes y' and commands pass 1
// right there; the helpers sweep t = B⁻¹y' while the leader finishes the iteration.  When the trial is
// accepted (≈ 70 % of the iterations at n = 60) pass 1 is off the critical path: 480 of the 3 160
// cycles of an iteration.  When it is not (another step, a reset, an exit), the partials are simply
// not read: the leader waits for the acknowledgement, publishes the right y and commands pass 1 again —
// the same work as without speculation.  B⁻¹ is not written by pass 1, so a discarded sweep leaves no
// trace; the arithmetic and its order are those of tile_solve_lh: the same bits.
//
// One CTA = 12 warps = FOUR problems: warps 0-7 are the helpers (problem i: warps 2i, 2i+1 — lanes
// 64·(i mod 2)…+63 of the tensor memory, columns 128·(i div 2)…: 256 columns per CTA), warps 8-11 the four
// leaders, one per scheduler.  Two CTAs = 8 problems = 24 warps per SM, 85 registers per thread.
// Hand-shakes: two mbarriers per problem — CMD (leader -> helpers: PASS1 | PASS2 | EXIT | QUIT, with an
// optional RESET) and ACK (helpers -> leader) — and the leader always collects the acknowledgement of
// a command before it issues the next one.
#pragma once

#include "qn_resident_tile.cuh"

namespace qn {

template <int N>
struct Tile3wCfg {
  static constexpr int kRec = 6;
  static constexpr int oRec = 0;                 // [4 blocks][16 slots] column records {c₁s, -u, s, ∇, pad, pad}
  static constexpr int oYb = oRec + 64 * kRec;   // y by slot [4][16]
  static constexpr int oS = oYb + 64;            // s by element, quad layout (row operands)
  static constexpr int oNu = oS + 80;            // -u by element
  static constexpr int oPart1 = oNu + 80;        // block partials of pass 1 and pass 2, [4 blocks][80]
  static constexpr int oPart2 = oPart1 + 4 * 80;
  static constexpr int oBar = oPart2 + 4 * 80;   // two mbarriers: CMD, ACK
  static constexpr int oCmd = oBar + 2;          // command word (int) and the problem index (long long)
  static constexpr int oJ = oCmd + 2;
  static constexpr int kDoubles = oJ + 2;
};
constexpr int k3wPass1 = 1, k3wPass2 = 2, k3wExit = 4, k3wReset = 8, k3wQuit = 16;
constexpr int k3wProblems = 4, k3wThreads = 32 * 3 * k3wProblems;

template <int N>
__global__ void __launch_bounds__(k3wThreads, 2)
qn_resident_tile3w_kernel(const __grid_constant__ BatchParams p) {
  using C = TileCfg<N>;
  using L = Tile3wCfg<N>;
  constexpr int RT = C::RT, CT = C::CT, NB = C::NB;
  constexpr int EPC = 4, NCH = CT, WPC = 8; // one tile column (4 rows) per tensor-memory chunk
  static_assert(NCH * WPC <= 128, "the tile must fit its 128 columns");
  extern __shared__ __align__(16) double smem_all[];
  __shared__ uint32_t tmem_base;
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const bool is_leader = warp >= 2 * k3wProblems;
  const int prob = is_leader ? warp - 2 * k3wProblems : warp >> 1; // problem slot of this warp inside the CTA
  double* smem = smem_all + prob * L::kDoubles;
  double* rec = smem + L::oRec;
  double* yb_arr = smem + L::oYb;
  double* s_arr = smem + L::oS;
  double* nu_arr = smem + L::oNu;
  double* part1 = smem + L::oPart1;
  double* part2 = smem + L::oPart2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar); // 0 CMD (count 32), 1 ACK (count 64)
  volatile int* cmd = reinterpret_cast<volatile int*>(smem + L::oCmd);
  volatile long long* jword = reinterpret_cast<volatile long long*>(smem + L::oJ);
  if (is_leader && lane == 0) { lh_bar_init(bars + 0, 32); lh_bar_init(bars + 1, 64); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n"
                 :: "l"((unsigned long long)__cvta_generic_to_shared(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  if (!is_leader) {
    // ------------------------------------------------------------------ helper: the two sweeps of its tile
    const int t = (warp & 1) * 32 + lane;  // thread of this problem's tile grid
    const int b = t >> 4, q = t & 15;      // column block (one per half-warp), row group
    const int row0 = RT * q, col0 = CT * b;
    const uint32_t tmem = __reduce_max_sync(0xffffffffu, tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u);
    unsigned ph = 0;
    for (;;) {
      lh_wait(bars + 0, ph & 1u); ph ^= 1u;
      const int c = *cmd;
      if (c & k3wQuit) break;
      if (c & k3wReset) { // B⁻¹ ← I, QN.jl:144-148
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
#pragma unroll
          for (int r = 0; r < RT; ++r) tm_store1(tmem + WPC * k + 2 * r, (row0 + r == col0 + k) ? 1.0 : 0.0);
        }
        tm_wait_store();
      }
      if (c & k3wExit) { // final B⁻¹ (QN_FLAG_WRITE_H); whole warps take part in the .sync.aligned loads
        if (p.h_out) {
          double* ho = p.h_out + *jword * p.h_stride;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            TmRaw<EPC> hr;
            tm_load_raw(hr, tmem + WPC * k);
            tm_wait_load();
            TmChunk<EPC> hc;
            tm_unpack(hc, hr);
#pragma unroll
            for (int r = 0; r < RT; ++r)
              if (row0 + r < N && col0 + k < N) ho[(row0 + r) + (long long)p.ldh * (col0 + k)] = hc.d[r];
          }
        }
      } else if (c & k3wPass1) { // t[i] = Σ_k y[col0+k] H[row0+i, col0+k], QN.jl:44-49 -> part1
        double tt[RT];
#pragma unroll
        for (int i = 0; i < RT; ++i) tt[i] = 0.0;
        const double2* yb2 = reinterpret_cast<const double2*>(yb_arr + b * 16);
        double2 yy[(CT + 1) / 2];
#pragma unroll
        for (int k = 0; k < (CT + 1) / 2; ++k) yy[k] = yb2[k];
        __syncwarp(); // keeps the loads above it
        TmRaw<EPC> ring[2];
        tm_load_raw(ring[0], tmem);
#pragma unroll
        for (int k = 0; k < CT; ++k) {
          const double yk = (k & 1) ? yy[k >> 1].y : yy[k >> 1].x;
          tm_wait_load();
          TmChunk<EPC> cur;
          tm_unpack(cur, ring[k & 1]);
          if (k + 1 < NCH) tm_load_raw(ring[(k + 1) & 1], tmem + WPC * (k + 1));
          if (C::kFull || col0 + k < N) {
#pragma unroll
            for (int r = 0; r < RT; ++r) tt[r] = fma(yk, cur.d[r], tt[r]);
          }
        }
        double* pp = part1 + b * C::kQuad + 2 * q;
        *reinterpret_cast<double2*>(pp) = make_double2(tt[0], tt[1]);
        *reinterpret_cast<double2*>(pp + 40) = make_double2(tt[2], tt[3]);
      } else if (c & k3wPass2) { // H ← H + c₁ssᵀ - usᵀ - suᵀ fused with d = H⁺∇_new, QN.jl:56-67 -> part2
        double sc_[RT], nuc[RT], acc[RT];
        const double2 s01 = *reinterpret_cast<const double2*>(s_arr + 2 * q);
        const double2 s23 = *reinterpret_cast<const double2*>(s_arr + 40 + 2 * q);
        const double2 n01 = *reinterpret_cast<const double2*>(nu_arr + 2 * q);
        const double2 n23 = *reinterpret_cast<const double2*>(nu_arr + 40 + 2 * q);
        sc_[0] = s01.x; sc_[1] = s01.y; sc_[2] = s23.x; sc_[3] = s23.y;
        nuc[0] = n01.x; nuc[1] = n01.y; nuc[2] = n23.x; nuc[3] = n23.y;
#pragma unroll
        for (int i = 0; i < RT; ++i) acc[i] = 0.0;
        const double* rb_ = rec + b * 16 * L::kRec;
        TmRaw<EPC> ring[2];
        tm_load_raw(ring[0], tmem);
#pragma unroll
        for (int k = 0; k < CT; ++k) {
          const double2 wn = *reinterpret_cast<const double2*>(rb_ + k * L::kRec);
          const double2 sg = *reinterpret_cast<const double2*>(rb_ + k * L::kRec + 2);
          tm_wait_load();
          TmChunk<EPC> cur;
          tm_unpack(cur, ring[k & 1]);
          if (k + 1 < NCH) tm_load_raw(ring[(k + 1) & 1], tmem + WPC * (k + 1));
          if (C::kFull || col0 + k < N) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
              double hh = cur.d[r];
              hh = fma(wn.x, sc_[r], hh);  // + (c₁ s[r]) s[c]
              hh = fma(wn.y, sc_[r], hh);  // - u[r] s[c]
              hh = fma(nuc[r], sg.x, hh);  // - u[c] s[r]
              cur.d[r] = hh;
              acc[r] = fma(hh, sg.y, acc[r]);
            }
          }
          // by the whole warp (.sync.aligned), also where a half-warp has no column k; one element per
          // store: its two 32-bit words are the DFMA's own register pair (no staging moves)
#pragma unroll
          for (int r = 0; r < RT; ++r) tm_store1(tmem + WPC * k + 2 * r, cur.d[r]);
        }
        tm_wait_store();
        double* pp = part2 + b * C::kQuad + 2 * q;
        *reinterpret_cast<double2*>(pp) = make_double2(acc[0], acc[1]);
        *reinterpret_cast<double2*>(pp + 40) = make_double2(acc[2], acc[3]);
      }
      lh_arrive(bars + 1); // every command but QUIT is acknowledged
    }
  } else {
    // ------------------------------------------------------------------ leader: the vectors and the O(n) part
    const bool hi_valid = lane + 32 < N;
    auto slot_of = [&](const int el) { return el < N ? (el / CT) * 16 + el % CT : -1; };
    const int slot_lo = slot_of(lane), slot_hi = slot_of(lane + 32);
    const int q_lo = C::quad(lane), q_hi = q_lo + 16; // quad-layout index of elements lane, lane + 32
    auto dot = [&](const Pair& a, const Pair& bb) {   // lane accumulator of rule C1
      double pacc = a.lo * bb.lo;
      if (hi_valid) pacc = pacc + a.hi * bb.hi;
      return pacc;
    };
    auto combine = [&](const double* part) { // Σ_b part[b][e'] in ascending block order (rule C2)
      Pair T;
      T.lo = part[q_lo];
#pragma unroll
      for (int bb = 1; bb < NB; ++bb) T.lo = T.lo + part[bb * C::kQuad + q_lo];
      T.hi = 0.0;
      if (hi_valid) {
        T.hi = part[q_hi];
#pragma unroll
        for (int bb = 1; bb < NB; ++bb) T.hi = T.hi + part[bb * C::kQuad + q_hi];
      }
      return T;
    };
    unsigned ph_ack = 0;
    bool outstanding = false;  // a command is out whose acknowledgement has not been collected
    bool helper_reset = false; // the helpers' tile still has to be reset (rides on the next command)
    auto ack_wait = [&]() {
      if (outstanding) { lh_wait(bars + 1, ph_ack & 1u); ph_ack ^= 1u; outstanding = false; }
    };
    auto command = [&](const int c) { // everything written before it is visible to the helpers
      ack_wait();
      if (lane == 0) *cmd = c | (helper_reset ? k3wReset : 0);
      helper_reset = false;
      __syncwarp();
      lh_arrive(bars + 0);
      outstanding = true;
    };
    auto publish_y = [&](const Pair& yv) {
      yb_arr[slot_lo] = yv.lo;
      if (slot_hi >= 0) yb_arr[slot_hi] = yv.hi;
    };
    const LsParams ls = p.ls;
    const double tol = p.tol;
    const int max_iter = p.max_iter;
    const int ntrace = p.trace != nullptr ? p.ntrace : 0;
    if (lane == 0) stamp_start(p);

    for (;;) {
      long long j = 0;
      if (lane == 0) j = (long long)atomicAdd(p.counter, 1ULL);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= p.batch) break;
      ack_wait(); // the helpers are done with the previous problem (its B⁻¹ is written out)
      if (lane == 0) *jword = j;

      Pair x, g, s, y, gp;
      double l0, sy = 0.0;
      {
        const double* th = p.theta0 + j * p.theta_stride; // copyto!(x_old, x), QN.jl:248
        x.lo = th[lane];
        x.hi = hi_valid ? th[lane + 32] : 0.0;
        s.lo = s.hi = y.lo = y.hi = gp.lo = gp.hi = 0.0;
        const double term = rosenbrock_warp_terms<N, true>(lane, x, g); // QN.jl:254 of iteration 1
        l0 = rosenbrock_finish<N>(warp_butterfly(term), x);
      }
      double result = __longlong_as_double(0x7ff8000000000000LL); // NaN, QN.jl:291
      int iters = 0, fevals = 0, status = QN_MAX_ITER;
      bool spec_done = false; // pass 1 of THIS iteration's update is already under way (speculated last iteration)

      struct Spec { Pair g; double term_next, sy; };
      auto spec_eval = [&](const double a, const Pair& d, Spec& sp, Pair& xn, double& term_trial) { // see tile_solve_one
        Pair sn, xo, yn;
        sn.lo = a * d.lo;     sn.hi = a * d.hi;
        xo.lo = x.lo + sn.lo; xo.hi = x.hi + sn.hi;
        sp.term_next = rosenbrock_warp_terms<N, true>(lane, xo, sp.g);
        if (a == 1.0) {
          xn = xo;
          term_trial = sp.term_next;
        } else {
          xn.lo = fma(a, d.lo, x.lo);
          xn.hi = fma(a, d.hi, x.hi);
          Pair unused;
          term_trial = rosenbrock_warp_terms<N, false>(lane, xn, unused);
        }
        yn.lo = g.lo - sp.g.lo; yn.hi = g.hi - sp.g.hi;
        sp.sy = dot(sn, yn);
      };

      for (int it = 1; it <= max_iter; ++it) { // QN.jl:250,253
        ++fevals;
        const bool small = fabs(g.lo) < tol && (!hi_valid || fabs(g.hi) < tol);
        const bool all_small = __all_sync(0xffffffffu, small);
        if (!isfinite(l0) || all_small) {
          iters = it - 1;
          if (!isfinite(l0)) { status = QN_NONFINITE; if (!(it & 1)) g = gp; }
          else { status = QN_CONVERGED; result = l0; }
          break;
        }
        double m;
        Pair d;
        Spec sp;
        Pair xn;
        double lx1 = 0.0;
        bool have_trial = false;
        bool spec_pending = false; // a speculative pass 1 for the NEXT iteration has been commanded
        if (it == 1) {
          m = -1.0; // QN.jl:263-264
        } else {
          // ---- BFGS_update!  QN.jl:34-69
          if (!spec_done) { // pass 1 was not speculated (or its y turned out wrong): command it now
            ack_wait();     // ... once the helpers are done reading the y of a discarded sweep
            publish_y(y);
            command(k3wPass1);
          }
          spec_done = false;
          // the operands of pass 2 that are known now (the helpers read them only after PASS2)
          s_arr[q_lo] = s.lo;
          s_arr[q_hi] = s.hi;
          *reinterpret_cast<double2*>(rec + slot_lo * L::kRec + 2) = make_double2(s.lo, g.lo);
          if (slot_hi >= 0) *reinterpret_cast<double2*>(rec + slot_hi * L::kRec + 2) = make_double2(s.hi, g.hi);
          const double rho = 1.0 / sy; // QN.jl:42 (no guard)
          ack_wait(); // the pass-1 partials
          {
            const Pair T = combine(part1);
            const double yHy = warp_butterfly(dot(T, y)); // Σ_c t[c] y[c], QN.jl:51
            const double c1 = fma(yHy, rho, 1.0) * rho;   // QN.jl:54
            const double nu_lo = -(T.lo * rho), nu_hi = -(T.hi * rho); // -B⁻¹y[c] = -(t ρ), QN.jl:50
            nu_arr[q_lo] = nu_lo;
            nu_arr[q_hi] = nu_hi;
            // (c₁ s[r]) rounded once per r (rule C4)
            *reinterpret_cast<double2*>(rec + slot_lo * L::kRec) = make_double2(c1 * s.lo, nu_lo);
            if (slot_hi >= 0) *reinterpret_cast<double2*>(rec + slot_hi * L::kRec) = make_double2(c1 * s.hi, nu_hi);
          }
          command(k3wPass2);
          ack_wait(); // the pass-2 partials
          d = combine(part2); // the new direction d = H⁺∇_new; `s, y = y, s`, QN.jl:65,268
          double tterm;
          spec_eval(1.0, d, sp, xn, tterm);
          // SPECULATE: if this trial is accepted, update_state! (QN.jl:21-27) makes s = 1·d and the next
          // BFGS_update! starts from y = ∇ - ∇(x + d) (QN.jl:38) — known now.  Let the helpers sweep.
          if (it < max_iter) {
            Pair ys;
            ys.lo = g.lo - sp.g.lo; ys.hi = g.hi - sp.g.hi;
            publish_y(ys);
            command(k3wPass1);
            spec_pending = true;
          }
          m = dot(d, g);
          warp_butterfly3(m, tterm, sp.sy, lane);
          sp.term_next = tterm;
          lx1 = rosenbrock_finish<N>(tterm, xn);
          have_trial = true;
        }
        if (m <= 0.0) { // QN.jl:272-280: H ← I, s ← ∇_new, m ← ∇ᵀ∇
          helper_reset = true; // rides on the next command; a speculated sweep of the old B⁻¹ is discarded
          spec_pending = false;
          d = g;
          m = warp_butterfly(dot(g, g));
          have_trial = false;
        }
        double alpha;
        {
          auto trial = [&](const double a) {
            double tterm;
            spec_eval(a, d, sp, xn, tterm);
            warp_butterfly3(tterm, sp.term_next, sp.sy, lane);
            ++fevals;
            return rosenbrock_finish<N>(tterm, xn);
          };
          if (have_trial) ++fevals;
          alpha = backtracking_linesearch(ls, l0, m, trial, lx1, have_trial, lx1);
        }
        if (alpha == 0.0) { status = QN_LINESEARCH_FAILED; iters = it - 1; if (!(it & 1)) g = gp; break; } // QN.jl:284
        s.lo = alpha * d.lo; s.hi = alpha * d.hi; // s ← αs      QN.jl:23-24
        x.lo = x.lo + s.lo;  x.hi = x.hi + s.hi;  // x_old += s  QN.jl:25
        if (it <= ntrace) {
          double* tr = p.trace + (j * p.ntrace + (it - 1)) * (long long)N;
          tr[lane] = x.lo;
          if (hi_valid) tr[lane + 32] = x.hi;
        }
        if (it == max_iter) { iters = it; if (!(it & 1)) g = gp; break; }
        gp = g;
        y.lo = g.lo - sp.g.lo; y.hi = g.hi - sp.g.hi; // y = ∇_old - ∇_new, QN.jl:38
        sy = sp.sy;
        g = sp.g;
        l0 = (alpha == 1.0) ? lx1 : rosenbrock_finish<N>(sp.term_next, x);
        // the speculated sweep is the right one exactly when the first trial (a = 1) was the accepted one:
        // then y above is, operation for operation, the y that was published
        spec_done = spec_pending && have_trial && alpha == 1.0;
      }
      command(k3wExit); // (collects the acknowledgement of a sweep that may still be under way first)
      if (p.x_out) {
        double* xo = p.x_out + j * p.x_stride;
        xo[lane] = x.lo;
        if (hi_valid) xo[lane + 32] = x.hi;
      }
      if (p.g_out) {
        double* go = p.g_out + j * p.g_stride;
        go[lane] = g.lo;
        if (hi_valid) go[lane + 32] = g.hi;
      }
      write_problem_scalars(p, j, lane, result, iters, fevals, status);
    }
    ack_wait();
    if (lane == 0) *cmd = k3wQuit;
    __syncwarp();
    lh_arrive(bars + 0);
    if (lane == 0) stamp_exit(p, k3wProblems * (int)blockIdx.x + prob);
  }
  __syncthreads(); // every problem slot is done with the tensor memory
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(tmem_base) : "memory");
}

}  // namespace qn
