// qn_resident_tile.cuh — the register-TILE resident kernel (33 <= n <= 64; the n = 60
// headline configuration runs here).
//
// Why tiles.  With one thread per row of B⁻¹ every step of the two O(n²) passes needs
// the column operands (y[r]; c₁s[r], -u[r], s[r], ∇[r]) as warp-wide shared-memory
// broadcasts, and each broadcast feeds ONE row per thread: 300 LDS.128 per problem
// iteration against 300 DFMA.  On B200 a broadcast LDS.128 costs 2.14 SM-cycles and a
// warp DFMA 2.24 SMSP-cycles (tools/ubench.cu, profiles/ubench_r1.txt), so that layout is
// bound by the shared-memory pipe (ncu: 78 % of peak wavefronts, FP64 pipe 36 %).
// Here a thread owns an RT×CT = 4×15 register tile of B⁻¹ (rows 4q..4q+3, columns of
// block b), so every broadcast is re-used for four rows: 16 DFMA per pair of LDS.128.
// The 16 lanes of a half-warp share a column block, and one LDS.128 serves one distinct
// address per half-warp at no extra cost (measured), so a warp covers two column blocks.
//
// Summation order.  A row sum is the sum of its column-block chains — exactly rule C2
// of the canonical order: block length CT = chain_block(n), partials
// added in ascending block order.  The partials of a row live in NB different threads;
// they meet in shared memory, and the reduction that follows (yᵀB⁻¹y, m) reads them
// straight from there, so the exchange costs no extra barrier.
//
// Warp-replicated vectors.  Everything that is O(n) — x, ∇, s, y, the objective, the line
// search, the dot products — is computed by BOTH warps redundantly, each warp holding
// the whole vector as two elements per lane (element l in `lo`, element l+32 in `hi`,
// which is exactly the accumulator layout of the canonical reduction, rule C1).  The
// two warps therefore only meet at the four barriers around the two tile passes; the
// objective evaluations of the line search need no barrier and no shared memory, and
// reductions whose inputs are ready together (ℓ₀ with sᵀy; m with the first trial ℓ)
// share one butterfly.
#pragma once

#include "qn_resident.cuh"

namespace qn {

// Optional per-phase cycle accounting (build with -DQN_PHASE_TIMING; tools/phase_timing.py).
#ifdef QN_PHASE_TIMING
// Accumulated in SHARED memory (a global read-modify-write per tick would put an L2 round trip, several
// hundred cycles under load, into every phase) and flushed once per problem; a tick costs ~40 cycles.
__device__ unsigned long long g_phase_clk[16];
#define QN_TICK_ON ((threadIdx.x & 31) == 0 && threadIdx.x < 64 && blockIdx.x == 0)
#define QN_TICK(k) do { const long long _t = clock64(); if (QN_TICK_ON) _shclk[k] += (unsigned long long)(_t - _tprev); _tprev = clock64(); } while (0)
#define QN_COUNT(k) do { if (QN_TICK_ON) _shclk[k] += 1ULL; } while (0)
#define QN_TICK_INIT __shared__ unsigned long long _shclk[16]; if (QN_TICK_ON) { for (int _k = 0; _k < 16; ++_k) _shclk[_k] = 0; } long long _tprev = clock64()
#define QN_TICK_FLUSH do { if (QN_TICK_ON) { for (int _k = 0; _k < 16; ++_k) g_phase_clk[_k] += _shclk[_k]; } } while (0)
#else
#define QN_TICK(k) do { } while (0)
#define QN_COUNT(k) do { } while (0)
#define QN_TICK_INIT do { } while (0)
#define QN_TICK_FLUSH do { } while (0)
#endif

// two elements of a length-n vector per lane: lo = v[lane], hi = v[lane + 32]
struct Pair {
  double lo, hi;
};

// Warp-level Rosenbrock (test/runtests.jl:4-33) on a replicated vector.  Lane i < N
// evaluates pair i: θ1 = x[i] is its own `lo`; θ2 = x[i+N] comes from the lane that
// holds it, by one 64-bit shuffle (each lane is asked for at most one of lo/hi
// because N <= 32).  Returns this lane's term of the @simd sum (0 for lanes >= N);
// the caller reduces it by rule C1.  With GRAD the gradient is returned replicated.
template <int n, bool GRAD>
__device__ __forceinline__ double rosenbrock_warp_terms(const int lane, const Pair x, Pair& g) {
  constexpr int N = n >> 1;
  static_assert(N <= 32 && n > 32, "pair partner must be reachable by one shuffle");
  // lane `src` sends x[src] (lo) when that element is a θ2 (src >= N), else x[src+32]
  const double send = (lane >= N) ? x.lo : x.hi;
  const int src = (lane + N < 32) ? lane + N : lane + N - 32;
  const double t2 = __shfl_sync(0xffffffffu, send, src);
  const double t1 = x.lo;
  const double a = t2 - t1 * t1;
  const double b = 1.0 - t1;
  const double term = (lane < N) ? 100.0 * (a * a) + b * b : 0.0; // runtests.jl:9,23
  if (GRAD) {
    const double g1 = (400.0 * a) * t1 + 2.0 * b; // ∇[i],   runtests.jl:24
    const double g2 = 200.0 * (t1 * t1 - t2);     // ∇[i+N], runtests.jl:25
    // element e = i+N lives in lane src(i); send g2 back along the same route
    const int back = (lane >= N) ? lane - N : lane + 32 - N; // the pair index whose θ2 I hold
    const double r = __shfl_sync(0xffffffffu, g2, back & 31);
    g.lo = (lane < N) ? g1 : r;                    // lanes >= N hold a θ2 in lo
    g.hi = r;                                      // lanes <  N hold a θ2 (or nothing) in hi
    if (n & 1) {                                   // tail element n-1 = 2N, runtests.jl:27-31
      const double tail = 2.0 * (1.0 - ((2 * N < 32) ? x.lo : x.hi));
      if (2 * N < 32) { if (lane == 2 * N) g.lo = tail; }
      else { if (lane == 2 * N - 32) g.hi = tail; }
    }
  }
  return term;
}

// finish ℓ from the reduced sum of terms: s = 0 - Σ; odd n adds the tail (runtests.jl:11-16)
template <int n>
__device__ __forceinline__ double rosenbrock_finish(const double sum, const Pair x) {
  double s = 0.0 - sum;
  if (n & 1) {
    constexpr int N = n >> 1;
    const double xt = __shfl_sync(0xffffffffu, (2 * N < 32) ? x.lo : x.hi, (2 * N) & 31);
    const double d = 1.0 - xt;
    s = -fma(d, d, -s);
  }
  return s;
}

template <int N>
struct TileCfg {
  static constexpr int RT = 4;                       // rows per thread
  static constexpr int CT = chain_block(N);          // columns per thread = rule-C2 block
  static constexpr int NB = (N + CT - 1) / CT;       // column blocks (<= 4)
  static constexpr int NG = (N + RT - 1) / RT;       // row groups (<= 16)
  static constexpr int BLOCK = 64;                   // 4 half-warps
  static constexpr bool kFull = (N % CT == 0);       // every column block is CT wide
  static_assert(N > 32 && N <= 64, "tile kernel covers 33..64");
  static_assert(NB <= 4 && NG <= 16, "tile grid must fit four half-warps");
  // ---- shared memory (doubles) ----
  // per-warp PRIVATE region (only the owning warp touches it: __syncwarp is enough).  Few distinct
  // address patterns on purpose — every extra base register is spilled in the hot loop (168
  // registers, 120 of them B⁻¹): the column operands are ONE array of 48-byte records addressed by
  // slot, everything element-indexed uses the quad layout below.
  static constexpr int kRec = 6;                     // doubles per column record {c₁s, -u, s, ∇, pad, pad}
  static constexpr int wRec = 0;                     // records of this warp's two column blocks [2][16]
  static constexpr int wYb = wRec + 32 * kRec;       // y by slot [2][16]: pass 1 reads it two columns at a time
  static constexpr int wS = wYb + 32;                // s  by element (row operands), quad layout
  static constexpr int wNu = wS + 80;                // -u by element
  static constexpr int wX = wNu + 80;                // x_old, ∇, ∇_old, y parked here (quad layout) while
  static constexpr int wG = wX + 80;                 //   the passes need the registers
  static constexpr int wGp = wG + 80;
  static constexpr int wY = wGp + 80;
  static constexpr int wSize = wY + 80;
  // shared by the two warps: block partials of pass 1 and pass 2, [4 blocks][80] each.
  // Quad layout of a 64-element vector: element e = 4q+i sits at quad(e) = (i>>1)·40 + 2q + (i&1),
  // so a thread's four rows are two consecutive-by-q 16-byte words (conflict-free LDS/STS.128),
  // and a warp's by-element access (e = lane) touches banks 0-15 and 16-31 once each.
  static constexpr int kQuad = 80;
  static constexpr int oPart1 = 2 * wSize;
  static constexpr int oPart2 = oPart1 + 4 * kQuad;
  static constexpr int oEnd = oPart2 + 4 * kQuad;
  __host__ __device__ static constexpr int quad(int e) { return ((e & 3) >> 1) * 40 + 2 * (e >> 2) + (e & 1); }
};

template <int N>
__host__ __device__ constexpr int tile_smem_doubles() { return TileCfg<N>::oEnd + 2; }

// Three rule-C1 butterflies on shared shuffles.  A butterfly leaves the same tree sum in every
// lane, so only one lane per sum has to finish it: at offset 16 the sums a and b trade halves
// (lanes with bit 4 clear carry on with a, the others with b — one exchange instead of two), at
// offset 8 that register and c do the same (bit 3 clear: a or b; set: c), offsets 4, 2, 1 are
// ordinary levels, and three broadcasts hand the finished sums (lanes 0, 16 and 8) to everyone.
// Every partial is formed from the same two operands as in warp_butterfly (IEEE addition is
// commutative): the same bits with 9 exchanges and 6 additions instead of 15 and 15 — and with a
// dependent chain of 6 where the register-starved schedule of three interleaved butterflies had 15.
__device__ __forceinline__ void warp_butterfly3(double& a, double& b, double& c, const int lane) {
  const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
  const double r16 = __shfl_xor_sync(0xffffffffu, up16 ? a : b, 16);
  const double c16 = __shfl_xor_sync(0xffffffffu, c, 16);
  const double ab = (up16 ? b : a) + r16;
  c = c + c16;
  const double r8 = __shfl_xor_sync(0xffffffffu, up8 ? ab : c, 8);
  double v = (up8 ? c : ab) + r8;
#pragma unroll
  for (int off = 4; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  a = __shfl_sync(0xffffffffu, v, 0);
  b = __shfl_sync(0xffffffffu, v, 16);
  c = __shfl_sync(0xffffffffu, v, 8);
}


// ---- tensor memory as the home of the tile (TM = true) -------------------------------------------
// The register file caps the all-register kernel at 6 problems per SM (120 of 168 registers are
// B⁻¹).  Tensor memory is as large as the register file (256 KB per SM), has its own data path
// (measured on B200, tools/tmem_bench.cu: 475 B/clk/SM of tcgen05.ld, 890 B/clk/SM of tcgen05.st,
// 30 cycles for a dependent 4 KB load — the shared-memory pipe moves 128 B/clk), and a thread may
// use its own lane as plain storage: no MMA is involved.  With the tile there, a thread needs
// < 128 registers, so 8 problems are resident per SM (4 warps per scheduler instead of 3), x, ∇
// and y stay in registers through the passes, and nothing spills.
// A warp reaches only the lanes 32·(warp mod 4)…+31, so a CTA is 4 warps = TWO problems
// (threads 0-63 and 64-127, each pair with its own named barrier and its own pulls from the work
// queue) and owns 128 columns.  A thread's 4×CT tile is stored in the order e = 4k + i (column k,
// row i), two 32-bit columns per element, and moved in chunks of EPC consecutive elements: EPC = 4
// (one column, tcgen05.ld/st .32x32b.x8; the default) or 2 (half a column, .x4), QN_TM_DEPTH
// chunks in flight ahead of the one being used.  The 32-bit words of tcgen05.ld/st are paired into
// fp64 registers inside the asm block.  Every tcgen05 instruction is .sync.aligned: it must be
// reached by the whole warp, also by a half-warp whose column block lies beyond n.
#ifndef QN_TM_EPC
#define QN_TM_EPC 4 // measured: one column per chunk 77.2 ms, half a column 80.5 ms per 2^16 problems (n = 60)
#endif
#ifndef QN_TM_PARK
#define QN_TM_PARK 0 // 1: x, ∇, y wait in shared memory during the passes, as in the all-register kernel
#endif
#ifndef QN_TM_DEPTH
#define QN_TM_DEPTH (QN_TM_EPC == 2 ? 4 : (QN_TM_STW == 1 ? 1 : 2)) // chunks in flight ahead of the one being used
#endif
#ifndef QN_TM_STW
// elements per tcgen05.st of pass 2: 4 = the chunk (.x8), 2 (.x4), 1 (.x2).  Measured on B200 (n = 60, 2^16
// problems, tools/tile_bench.cu, same output hash): 76.2 / 74.7 / 73.3 ms, and 71.3 ms for .x2 with a
// prefetch distance of ONE chunk (the registers the staging block needed now hide nothing worth having).
// ptxas copies every element into an aligned 8-register block in front of an .x8 store (120 of pass
// 2's 480 instructions per thread); a .x2 store takes the DFMA's own register pair.
#define QN_TM_STW 1
#endif
#ifndef QN_TM_RAW
#define QN_TM_RAW 1 // 1: tcgen05.ld hands back raw 32-bit registers, paired into fp64 only AFTER tcgen05.wait::ld
#endif
template <int EPC>
struct TmChunk { double d[EPC]; };
// tcgen05.wait::ld emits no instruction: ptxas guards the destination registers of LDTM with a
// scoreboard like those of any load.  The PTX contract nevertheless leaves them undefined until the
// wait, so the load returns the raw 32-bit registers (TmRaw) and tm_unpack — called after
// tm_wait_load() — pairs them into doubles; the pairing is register naming, not an instruction.
template <int EPC>
struct TmRaw { uint32_t w[2 * EPC]; };
__device__ __forceinline__ void tm_wait_load() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
template <int EPC>
__device__ __forceinline__ void tm_load_raw(TmRaw<EPC>& c, const uint32_t addr) {
  static_assert(EPC == 2 || EPC == 4, "chunk of 2 or 4 elements");
  if constexpr (EPC == 2) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]) : "r"(addr));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]), "=r"(c.w[4]), "=r"(c.w[5]), "=r"(c.w[6]), "=r"(c.w[7])
                 : "r"(addr));
  }
}
template <int EPC>
__device__ __forceinline__ void tm_unpack(TmChunk<EPC>& c, const TmRaw<EPC>& r) {
#pragma unroll
  for (int k = 0; k < EPC; ++k) c.d[k] = __hiloint2double((int)r.w[2 * k + 1], (int)r.w[2 * k]);
}
template <int EPC>
__device__ __forceinline__ void tm_load(TmChunk<EPC>& c, const uint32_t addr) {
  static_assert(EPC == 2 || EPC == 4, "chunk of 2 or 4 elements");
  if constexpr (EPC == 2) {
    asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x4.b32 {t0a, t0b, t1a, t1b}, [%2];\n"
                 "mov.b64 %0, {t0a, t0b};\nmov.b64 %1, {t1a, t1b};\n}\n"
                 : "=d"(c.d[0]), "=d"(c.d[1]) : "r"(addr));
  } else {
    asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b, t2a, t2b, t3a, t3b;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {t0a, t0b, t1a, t1b, t2a, t2b, t3a, t3b}, [%4];\n"
                 "mov.b64 %0, {t0a, t0b};\nmov.b64 %1, {t1a, t1b};\nmov.b64 %2, {t2a, t2b};\nmov.b64 %3, {t3a, t3b};\n}\n"
                 : "=d"(c.d[0]), "=d"(c.d[1]), "=d"(c.d[2]), "=d"(c.d[3]) : "r"(addr));
  }
}
template <int EPC>
__device__ __forceinline__ void tm_store(const uint32_t addr, const TmChunk<EPC>& c) {
  if constexpr (EPC == 2) {
    asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b;\n"
                 "mov.b64 {t0a, t0b}, %1;\nmov.b64 {t1a, t1b}, %2;\n"
                 "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {t0a, t0b, t1a, t1b};\n}\n"
                 :: "r"(addr), "d"(c.d[0]), "d"(c.d[1]) : "memory");
  } else {
    asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b, t2a, t2b, t3a, t3b;\n"
                 "mov.b64 {t0a, t0b}, %1;\nmov.b64 {t1a, t1b}, %2;\nmov.b64 {t2a, t2b}, %3;\nmov.b64 {t3a, t3b}, %4;\n"
                 "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {t0a, t0b, t1a, t1b, t2a, t2b, t3a, t3b};\n}\n"
                 :: "r"(addr), "d"(c.d[0]), "d"(c.d[1]), "d"(c.d[2]), "d"(c.d[3]) : "memory");
  }
}
// One fp64 element per store (tcgen05.st .x2): its two 32-bit words are the register pair the DFMA
// wrote, so ptxas needs no staging block in front of the store (the .x8 form costs 8 register moves
// per column: 120 of pass 2's 480 instructions per thread).
__device__ __forceinline__ void tm_store1(const uint32_t addr, const double v) {
  asm volatile("{\n.reg .b32 ta, tb;\nmov.b64 {ta, tb}, %1;\n"
               "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {ta, tb};\n}\n" :: "r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void tm_store2(const uint32_t addr, const double v0, const double v1) {
  asm volatile("{\n.reg .b32 t0a, t0b, t1a, t1b;\nmov.b64 {t0a, t0b}, %1;\nmov.b64 {t1a, t1b}, %2;\n"
               "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {t0a, t0b, t1a, t1b};\n}\n" :: "r"(addr), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ void tm_wait_store() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// barrier of one problem's 64 threads (barrier 0 stays the CTA's)
// (immediate ids: with the id in a register ptxas reserves all 16 barriers and one CTA fills the SM)
__device__ __forceinline__ void pair_barrier(const int pair) {
  if (pair == 0) asm volatile("bar.sync 1, 64;\n" ::: "memory");
  else asm volatile("bar.sync 2, 64;\n" ::: "memory");
}

template <int N, int PF, bool TM = false>
__device__ __forceinline__ void tile_solve_one(const BatchParams& p, const long long j, double* smem, const uint32_t tmem = 0) {
  using C = TileCfg<N>;
  constexpr int RT = C::RT, CT = C::CT, NB = C::NB;
  const int t = TM ? (threadIdx.x & 63) : threadIdx.x; // thread of this problem
  const int lane = t & 31;
  const int warp = t >> 5;
  const int e = t;                    // the element this thread writes out at the end
  constexpr int EPC = QN_TM_EPC;                        // elements per chunk of the tile in tensor memory (TM)
  constexpr int SPC = 4 / EPC;                          // chunks per column
  constexpr int NCH = TileCfg<N>::CT * SPC;             // chunks
  constexpr int WPC = 2 * EPC;                          // 32-bit columns per chunk
  constexpr int TD = QN_TM_DEPTH;                       // prefetch distance in chunks
  static_assert(!TM || NCH * WPC <= 128, "the tile must fit the CTA's 128 columns");
  auto cta_sync = [&]() {
    if constexpr (TM) pair_barrier((int)(threadIdx.x >> 6)); else __syncthreads();
  };
  const bool owner = e < N;
  const bool hi_valid = lane + 32 < N;
  const int b = t >> 4;               // column block of this thread's tile (one per half-warp)
  const int hb = b & 1;               // ... counted inside this warp's private operand arrays
  const int q = t & 15;               // row group
  const int row0 = RT * q;
  const int col0 = CT * b;

  double* pw = smem + warp * C::wSize;             // this warp's private region
  double* rec = pw + C::wRec;                      // column records
  double* yb_arr = pw + C::wYb;
  double* s_arr = pw + C::wS;
  double* nu_arr = pw + C::wNu;
  double* part1 = smem + C::oPart1;
  double* part2 = smem + C::oPart2;

  // slot of element el in this warp's block-padded operand arrays, or -1 when the element
  // belongs to a column block of the other warp
  auto slot_of = [&](const int el) {
    const int blk = el / CT - 2 * warp;
    return (el < N && blk >= 0 && blk < 2) ? blk * 16 + el % CT : -1;
  };
  const int slot_lo = slot_of(lane), slot_hi = slot_of(lane + 32);
  const int q_lo = C::quad(lane), q_hi = q_lo + 16; // quad-layout index of elements lane, lane + 32

  double h[TM ? 1 : RT][TM ? 1 : CT]; // the tile: h[i][k] = B⁻¹[row0 + i, col0 + k] (TM: in tensor memory)
  (void)h;

  auto mine = [&](const Pair& v) { return warp ? v.hi : v.lo; }; // element e of a replicated vector
  auto dot = [&](const Pair& a, const Pair& bb) {                // lane accumulator of rule C1
    double pacc = a.lo * bb.lo;
    if (hi_valid) pacc = pacc + a.hi * bb.hi;
    return pacc;
  };
  auto park = [&](const int off, const Pair& v) { pw[off + q_lo] = v.lo; pw[off + q_hi] = v.hi; };
  auto unpark = [&](const int off) { Pair v; v.lo = pw[off + q_lo]; v.hi = pw[off + q_hi]; return v; };
  // Σ_b part[b][e'] in ascending block order (rule C2), for e' = lane and lane+32
  auto combine = [&](const double* part) {
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

  // ---- state.  x, g: x_old and ∇ at x_old.  s: the last step α·d.  y, sy: ∇_old - ∇ and sᵀy
  // for the coming update.  They are produced one iteration early, by the line-search trial
  // that gets accepted (see spec_eval).
  Pair x, g, s, y;
  double l0, sy = 0.0;
  {
    const double* th = p.theta0 + j * p.theta_stride; // copyto!(x_old, x), QN.jl:248
    x.lo = th[lane];
    x.hi = hi_valid ? th[lane + 32] : 0.0;
    s.lo = s.hi = y.lo = y.hi = 0.0;
    const double term = rosenbrock_warp_terms<N, true>(lane, x, g); // QN.jl:254 of iteration 1
    l0 = rosenbrock_finish<N>(warp_butterfly(term), x);
  }
  // ∇ of the previous iteration (what the canonical slot shows after a failure in an even
  // iteration): parked in shared memory, or kept in registers when the tile is in tensor memory
  constexpr bool kParkGp = !TM || QN_TM_PARK;
  Pair gp; gp.lo = gp.hi = 0.0;
  auto save_gp = [&](const Pair& v) { if (kParkGp) park(C::wGp, v); else gp = v; };
  auto load_gp = [&]() { return kParkGp ? unpark(C::wGp) : gp; };
  save_gp(gp);
  double result = __longlong_as_double(0x7ff8000000000000LL); // NaN, QN.jl:291
  int iters = 0, fevals = 0, status = QN_MAX_ITER; // iters (completed iterations) is set where the loop is left

  // A line-search trial at step a along d (QN.jl:157-162), evaluated TOGETHER with what the
  // next iteration needs if the trial is accepted: update_state! (QN.jl:21-27) gives
  // s' = a·d, x' = x + s'; then ∂logdensity! at x' (QN.jl:254), y' = ∇ - ∇' and s'ᵀy'
  // (QN.jl:36-41).  The trial point itself is x + a·d as ONE fma (rule C3) — the same point
  // as x' only for a = 1.  All sums are returned un-reduced so they share one butterfly.
  // Only ∇', the un-reduced ℓ' terms and s'ᵀy' are kept; s', x', y' are recomputed on acceptance
  // (the same operations on the same operands: the same bits).
  struct Spec { Pair g; double term_next, sy; };
  auto spec_eval = [&](const double a, const Pair& d, Spec& sp, Pair& xn, double& term_trial) {
    Pair sn, xo, yn;
    sn.lo = a * d.lo;     sn.hi = a * d.hi;
    xo.lo = x.lo + sn.lo; xo.hi = x.hi + sn.hi;
    sp.term_next = rosenbrock_warp_terms<N, true>(lane, xo, sp.g);
    if (a == 1.0) {           // x + 1·d and fma(1, d, x) are the same bits
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

  QN_TICK_INIT;
  for (int it = 1; it <= p.max_iter; ++it) { // QN.jl:250,253
    QN_TICK(0);
    ++fevals; // the ∂logdensity! of this iteration (QN.jl:254), evaluated by the accepted trial
    // QN.jl:255 (ℓ₀ not finite) and QN.jl:257-261 (maximum(abs, ∇_new) < tol; a NaN entry makes it
    // false) leave the loop through ONE branch: the vote is taken before either is looked at
    const bool small = fabs(g.lo) < p.tol && (!hi_valid || fabs(g.hi) < p.tol);
    const bool all_small = __all_sync(0xffffffffu, small);
    if (!isfinite(l0) || all_small) {
      iters = it - 1;
      if (!isfinite(l0)) { status = QN_NONFINITE; if (!(it & 1)) g = load_gp(); }
      else { status = QN_CONVERGED; result = l0; }
      break;
    }
    QN_TICK(1);

    double m;
    Pair d;       // search direction
    Spec sp;      // candidate next state of the latest trial
    Pair xn;      // the latest trial point
    double lx1 = 0.0;
    bool have_trial = false;
    if (it == 1) {
      m = -1.0; // QN.jl:263-264
    } else {
      // ---- BFGS_update!  QN.jl:34-69.  Operands go to this warp's private arrays.
      s_arr[q_lo] = s.lo;
      s_arr[q_hi] = s.hi;
      if (slot_lo >= 0) {
        *reinterpret_cast<double2*>(rec + slot_lo * C::kRec + 2) = make_double2(s.lo, g.lo);
        yb_arr[slot_lo] = y.lo;
      }
      if (slot_hi >= 0) {
        *reinterpret_cast<double2*>(rec + slot_hi * C::kRec + 2) = make_double2(s.hi, g.hi);
        yb_arr[slot_hi] = y.hi;
      }
      // (with the tile in tensor memory the registers can keep x, ∇ and y through the passes)
      constexpr bool kPark = !TM || QN_TM_PARK;
      if (kPark) {
        park(C::wX, x);
        park(C::wG, g);
        park(C::wY, y);
      }
      __syncwarp();
      QN_TICK(2);
      const double rho = 1.0 / sy; // QN.jl:42 (no guard)
      // pass 1 on the tile: t[i] = Σ_k y[col0+k] H[row0+i, col0+k]   QN.jl:44-49
      {
        double tt[RT];
#pragma unroll
        for (int i = 0; i < RT; ++i) tt[i] = 0.0;
        // y's block, two columns per LDS.128: eight loads issued back to back ahead of the fmas (the
        // warp barrier keeps ptxas from sinking each load to its first use), so their latency is
        // paid once and not per column
        const double2* yb2 = reinterpret_cast<const double2*>(yb_arr + hb * 16);
        double2 yy[(CT + 1) / 2];
#pragma unroll
        for (int k = 0; k < (CT + 1) / 2; ++k) yy[k] = yb2[k];
        __syncwarp(); // keeps the loads above it
        if constexpr (TM) {
          TmChunk<EPC> ring[TD + 1];
#pragma unroll
          for (int c = 0; c < TD && c < NCH; ++c) tm_load(ring[c], tmem + WPC * c);
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            const double yk = (k & 1) ? yy[k >> 1].y : yy[k >> 1].x;
#pragma unroll
            for (int sc = 0; sc < SPC; ++sc) {
              const int c = k * SPC + sc;
              tm_wait_load();
              if (c + TD < NCH) tm_load(ring[(c + TD) % (TD + 1)], tmem + WPC * (c + TD));
              if (C::kFull || col0 + k < N) {
#pragma unroll
                for (int r = 0; r < EPC; ++r) tt[sc * EPC + r] = fma(yk, ring[c % (TD + 1)].d[r], tt[sc * EPC + r]);
              }
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            const double yk = (k & 1) ? yy[k >> 1].y : yy[k >> 1].x;
            if (C::kFull || col0 + k < N) {
#pragma unroll
              for (int i = 0; i < RT; ++i) tt[i] = fma(yk, h[i][k], tt[i]);
            }
          }
        }
        double* pp = part1 + b * C::kQuad + 2 * q;
        *reinterpret_cast<double2*>(pp) = make_double2(tt[0], tt[1]);
        *reinterpret_cast<double2*>(pp + 40) = make_double2(tt[2], tt[3]);
      }
      cta_sync(); // pass-1 partials of both warps visible
      QN_TICK(3);
      {
        const Pair T = combine(part1);
        const Pair yv = kPark ? unpark(C::wY) : y;
        const double s_lo = s_arr[q_lo], s_hi = s_arr[q_hi]; // fetched ahead of the butterfly that c₁ waits for
        const double yHy = warp_butterfly(dot(T, yv)); // Σ_c t[c] y[c], QN.jl:51
        const double c1 = fma(yHy, rho, 1.0) * rho;    // QN.jl:54
        const double nu_lo = -(T.lo * rho), nu_hi = -(T.hi * rho); // -B⁻¹y[c] = -(t ρ), QN.jl:50
        nu_arr[q_lo] = nu_lo;
        nu_arr[q_hi] = nu_hi;
        // (c₁ s[r]) rounded once per r (rule C4)
        if (slot_lo >= 0) *reinterpret_cast<double2*>(rec + slot_lo * C::kRec) = make_double2(c1 * s_lo, nu_lo);
        if (slot_hi >= 0) *reinterpret_cast<double2*>(rec + slot_hi * C::kRec) = make_double2(c1 * s_hi, nu_hi);
      }
      __syncwarp();
      QN_TICK(4);
      // pass 2 on the tile: H ← H + c₁ssᵀ - usᵀ - suᵀ fused with d = H⁺∇_new   QN.jl:56-67
      {
        double sc_[RT], nuc[RT], acc[RT];
        const double2 s01 = *reinterpret_cast<const double2*>(s_arr + 2 * q);
        const double2 s23 = *reinterpret_cast<const double2*>(s_arr + 40 + 2 * q);
        const double2 n01 = *reinterpret_cast<const double2*>(nu_arr + 2 * q);
        const double2 n23 = *reinterpret_cast<const double2*>(nu_arr + 40 + 2 * q);
        sc_[0] = s01.x; sc_[1] = s01.y; sc_[2] = s23.x; sc_[3] = s23.y;
        nuc[0] = n01.x; nuc[1] = n01.y; nuc[2] = n23.x; nuc[3] = n23.y;
#pragma unroll
        for (int i = 0; i < RT; ++i) acc[i] = 0.0;
        const double* rb_ = rec + hb * 16 * C::kRec;
        auto ld_wn = [&](const int k) { return *reinterpret_cast<const double2*>(rb_ + k * C::kRec); };
        auto ld_sg = [&](const int k) { return *reinterpret_cast<const double2*>(rb_ + k * C::kRec + 2); };
        if constexpr (TM) {
          TmChunk<EPC> ring[TD + 1];
#pragma unroll
          for (int c = 0; c < TD && c < NCH; ++c) tm_load(ring[c], tmem + WPC * c);
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            const double2 wn = ld_wn(k), sg = ld_sg(k);
#pragma unroll
            for (int sc = 0; sc < SPC; ++sc) {
              const int c = k * SPC + sc;
              TmChunk<EPC>& cur = ring[c % (TD + 1)];
              tm_wait_load();
              if (c + TD < NCH) tm_load(ring[(c + TD) % (TD + 1)], tmem + WPC * (c + TD));
              if (C::kFull || col0 + k < N) {
#pragma unroll
                for (int r = 0; r < EPC; ++r) {
                  const int i = sc * EPC + r;
                  double hh = cur.d[r];
                  hh = fma(wn.x, sc_[i], hh);  // + (c₁ s[r]) s[c]
                  hh = fma(wn.y, sc_[i], hh);  // - u[r] s[c]
                  hh = fma(nuc[i], sg.x, hh);  // - u[c] s[r]
                  cur.d[r] = hh;
                  acc[i] = fma(hh, sg.y, acc[i]);
                }
              }
              // by the whole warp (.sync.aligned), also where a half-warp has no column k
              if constexpr (QN_TM_STW == 1) {
#pragma unroll
                for (int r = 0; r < EPC; ++r) tm_store1(tmem + WPC * c + 2 * r, cur.d[r]);
              } else if constexpr (QN_TM_STW == 2) {
#pragma unroll
                for (int r = 0; r < EPC; r += 2) tm_store2(tmem + WPC * c + 2 * r, cur.d[r], cur.d[r + 1]);
              } else {
                tm_store(tmem + WPC * c, cur);
              }
            }
          }
          tm_wait_store();
        } else {
          // operand ring: the loads of step k+PF are issued while step k computes
          double2 wnr[PF + 1], sgr[PF + 1];
  #pragma unroll
          for (int k = 0; k < PF && k < CT; ++k) { wnr[k] = ld_wn(k); sgr[k] = ld_sg(k); }
  #pragma unroll
          for (int k = 0; k < CT; ++k) {
            if (k + PF < CT) { wnr[(k + PF) % (PF + 1)] = ld_wn(k + PF); sgr[(k + PF) % (PF + 1)] = ld_sg(k + PF); }
            const double2 wn = wnr[k % (PF + 1)], sg = sgr[k % (PF + 1)];
            if (C::kFull || col0 + k < N) {
  #pragma unroll
              for (int i = 0; i < RT; ++i) {
                double hh = h[i][k];
                hh = fma(wn.x, sc_[i], hh);   // + (c₁ s[r]) s[c]
                hh = fma(wn.y, sc_[i], hh);   // - u[r] s[c]
                hh = fma(nuc[i], sg.x, hh);  // - u[c] s[r]
                h[i][k] = hh;
                acc[i] = fma(hh, sg.y, acc[i]);
              }
            }
          }
        }
        double* pp = part2 + b * C::kQuad + 2 * q;
        *reinterpret_cast<double2*>(pp) = make_double2(acc[0], acc[1]);
        *reinterpret_cast<double2*>(pp + 40) = make_double2(acc[2], acc[3]);
      }
      cta_sync(); // pass-2 partials of both warps visible
      QN_TICK(5);
      d = combine(part2); // the new direction d = H⁺∇_new; `s, y = y, s`, QN.jl:65,268
      if (kPark) {
        x = unpark(C::wX);
        g = unpark(C::wG);
      }
      // m = dᵀ∇_new (QN.jl:66-68) shares its butterfly with the first line-search trial
      // ℓ(x + 1·d) (QN.jl:174) and with s'ᵀy' of the iteration that follows if it is accepted.
      double tterm;
      spec_eval(1.0, d, sp, xn, tterm);
      m = dot(d, g);
      warp_butterfly3(m, tterm, sp.sy, lane);
      sp.term_next = tterm;
      lx1 = rosenbrock_finish<N>(tterm, xn);
      have_trial = true;
      QN_TICK(6);
    }
    if (m <= 0.0) { // QN.jl:272-280: H ← I, s ← ∇_new, m ← ∇ᵀ∇
      if constexpr (TM) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          TmChunk<EPC> hc;
#pragma unroll
          for (int r = 0; r < EPC; ++r) hc.d[r] = (row0 + (c % SPC) * EPC + r == col0 + c / SPC) ? 1.0 : 0.0;
          tm_store(tmem + WPC * c, hc);
        }
        tm_wait_store();
      } else {
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
          for (int k = 0; k < CT; ++k) h[i][k] = (row0 + i == col0 + k) ? 1.0 : 0.0;
      }
      d = g;
      m = warp_butterfly(dot(g, g));
      have_trial = false;
    }

    // ---- linesearch!  QN.jl:165-232; both warps run the same scalar code on the same values
    double alpha;
    {
      const LsParams& ls = p.ls;
      auto trial = [&](const double a) {
        double tterm;
        spec_eval(a, d, sp, xn, tterm);
        warp_butterfly3(tterm, sp.term_next, sp.sy, lane);
        ++fevals;
        return rosenbrock_finish<N>(tterm, xn);
      };
      if (have_trial) ++fevals; // the first trial was evaluated above, together with m
      alpha = backtracking_linesearch(ls, l0, m, trial, lx1, have_trial, lx1);
    }
    QN_TICK(7);
    if (alpha == 0.0) { status = QN_LINESEARCH_FAILED; iters = it - 1; if (!(it & 1)) g = load_gp(); break; } // QN.jl:284
    // ---- accept: update_state! (QN.jl:21-27) and the buffer swap (QN.jl:286) were already
    // applied to the candidate; its ∂logdensity! becomes the next iteration's (QN.jl:254).
    s.lo = alpha * d.lo; s.hi = alpha * d.hi; // s ← αs      QN.jl:23-24
    x.lo = x.lo + s.lo;  x.hi = x.hi + s.hi;  // x_old += s  QN.jl:25
    if (p.trace != nullptr && it <= p.ntrace && owner)
      p.trace[(j * p.ntrace + (it - 1)) * (long long)N + e] = mine(x);
    // N iterations exhausted (QN.jl:250,291); which ∇ the canonical slot shows: see qn_resident.cuh
    if (it == p.max_iter) { iters = it; if (!(it & 1)) g = load_gp(); break; }
    save_gp(g);
    y.lo = g.lo - sp.g.lo; y.hi = g.hi - sp.g.hi; // y = ∇_old - ∇_new, QN.jl:38
    sy = sp.sy;
    g = sp.g;
    l0 = (alpha == 1.0) ? lx1 : rosenbrock_finish<N>(sp.term_next, x);
  }

  QN_TICK_FLUSH;
  if (owner) {
    if (p.x_out) p.x_out[j * p.x_stride + e] = mine(x);
    if (p.g_out) p.g_out[j * p.g_stride + e] = mine(g);
  }
  // (TM: every thread of the warp takes part in the .sync.aligned loads, also a half-warp without a column block)
  if (p.h_out && (TM || b < NB)) {
    double* ho = p.h_out + j * p.h_stride;
    if constexpr (TM) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        TmChunk<EPC> hc;
        tm_load(hc, tmem + WPC * c);
        tm_wait_load();
#pragma unroll
        for (int r = 0; r < EPC; ++r) {
          const int k = c / SPC, i = (c % SPC) * EPC + r;
          if (row0 + i < N && col0 + k < N) ho[(row0 + i) + (long long)p.ldh * (col0 + k)] = hc.d[r];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int k = 0; k < CT; ++k)
          if (row0 + i < N && col0 + k < N) ho[(row0 + i) + (long long)p.ldh * (col0 + k)] = h[i][k];
    }
  }
  write_problem_scalars(p, j, t, result, iters, fevals, status);
}

// =====================================================================================================
// Leader / helper form of the same solve (QN_TILE_LH, the default).
//
// In tile_solve_one BOTH warps of a problem run the whole O(n) part — objective, line search, dot
// products, butterflies — on replicated vectors: 370 of a warp's 810 instructions per iteration, and
// twice the shuffles and FP64 adds the problem needs.  Here one warp, the LEADER, owns the vectors
// and runs that part alone; the HELPER only sweeps its half of the tile (pass 1, pass 2) and waits.
// Its issue slots, FP64 pipe time and shuffle traffic go to the other problems resident on the SM.
// The leader's arithmetic is the replicated warp's, operation for operation: the same bits.
//
// Hand-shakes: four mbarriers per problem in shared memory (arrive = release, wait = acquire; the
// SM's 16 named barriers do not stretch to 4 CTAs x 2 problems x 4 hand-shakes):
//   CMD   leader -> helper   a command word: PASS (operands of pass 1 are published) or EXIT, each
//                            with an optional RESET (B⁻¹ <- I in the helper's half of the tile first)
//   P1    helper -> leader   the helper's pass-1 partials are in shared memory
//   OPS2  leader -> helper   c₁s, -u (column records) and -u by row are published
//   P2    helper -> leader   the helper's pass-2 partials are in shared memory
// Exactly one command per iteration and the leader waits for P2 before the next one, so a hand-shake
// never overtakes the previous one.  The SAME warp of every pair leads (BatchParams::lead_div):
// leaders and helpers then sit on different schedulers of the SM, and a helper's idle polling
// never takes an issue slot from a leader's dependent chain.
// QN_TILE_LH: the tensor-memory kernel (65.9 ms per 2^16 n = 60 problems with, 71.3 without).
// QN_TILE_LH_REG: the all-register kernel, which is faster WITHOUT the split (80.3 ms against 87.9;
// profiles/r2_lead_variants.txt): its six resident problems already fill the schedulers.
#ifndef QN_TILE_LH
#define QN_TILE_LH 1
#endif
#ifndef QN_TILE_LH_REG
#define QN_TILE_LH_REG 0
#endif
#ifndef QN_LH_HOIST
#define QN_LH_HOIST 1
#endif
#ifndef QN_TM_DEPTH1
#define QN_TM_DEPTH1 QN_TM_DEPTH // chunks in flight in pass 1 (leader/helper kernels)
#endif
template <int N>
struct TileLhCfg {
  static constexpr int kRec = 6;                 // doubles per column record {c₁s, -u, s, ∇, pad, pad}
  static constexpr int oRec = 0;                 // [4 blocks][16 slots] records
  static constexpr int oYb = oRec + 64 * kRec;   // y by slot [4][16]
  static constexpr int oS = oYb + 64;            // s by element, quad layout (row operands)
  static constexpr int oNu = oS + 80;            // -u by element
  static constexpr int oX = oNu + 80;            // x_old, ∇, ∇_old, y parked here while the passes need the
  static constexpr int oG = oX + 80;             //   registers (tile in the register file only)
  static constexpr int oGp = oG + 80;
  static constexpr int oY = oGp + 80;
  static constexpr int oPart1 = oY + 80;         // block partials of pass 1 and pass 2, [4 blocks][80]
  static constexpr int oPart2 = oPart1 + 4 * 80;
  static constexpr int oBar = oPart2 + 4 * 80;   // four mbarriers
  static constexpr int oCmd = oBar + 4;          // the command word
  static constexpr int oEnd = oCmd + 2;          // (+ 2 doubles for the kernel's work-queue slot)
};
template <int N>
__host__ __device__ constexpr int tile_lh_smem_doubles() { return TileLhCfg<N>::oEnd + 2; }
// shared memory (doubles) of one problem in the tile kernels as they are built
template <int N, bool TM>
__host__ __device__ constexpr int tile_problem_smem_doubles() {
  return (TM ? QN_TILE_LH : QN_TILE_LH_REG) ? tile_lh_smem_doubles<N>() : tile_smem_doubles<N>();
}

constexpr int kLhPass = 1, kLhExit = 2, kLhReset = 4;
__device__ __forceinline__ void lh_bar_init(uint64_t* bar, const unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lh_arrive(uint64_t* bar) { // every lane of the warp: count 32, release
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// acquire.  The suspend-time hint lets the hardware park the waiting warp until the phase completes (or
// the hint, in ns, runs out) instead of returning after its short default limit: without it a waiting
// helper re-issued try_wait + branch ~60 times per wait — 38 % of the kernel's warp-stall samples sat on
// these branches and 1.7e9 of its instructions were spin iterations (profiles/r2_tile60_hot_sass.txt),
// issue slots taken from the leaders that share the scheduler.
#ifndef QN_LH_SUSPEND_NS
#define QN_LH_SUSPEND_NS 1000000
#endif
__device__ __forceinline__ void lh_wait(uint64_t* bar, const unsigned parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLH_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra.uni LH_DONE;\nbra.uni LH_WAIT;\nLH_DONE:\n}\n"
      :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity), "r"((uint32_t)QN_LH_SUSPEND_NS) : "memory");
}

// `ph`: this warp's parity bits of the four hand-shakes (bit k = barrier k), carried across problems.
template <int N, int PF, bool TM>
__device__ __forceinline__ void tile_solve_lh(const BatchParams& p, const long long j, double* smem,
                                              const uint32_t tmem, const bool leader, unsigned& ph) {
  using C = TileCfg<N>;
  using L = TileLhCfg<N>;
  constexpr int RT = C::RT, CT = C::CT, NB = C::NB;
  const int t = TM ? (threadIdx.x & 63) : threadIdx.x; // thread of this problem
  const int lane = t & 31;
  constexpr int EPC = QN_TM_EPC, SPC = 4 / EPC, NCH = CT * SPC, WPC = 2 * EPC, TD = QN_TM_DEPTH;
  static_assert(!TM || NCH * WPC <= 128, "the tile must fit the CTA's 128 columns");
  const bool hi_valid = lane + 32 < N;
  const int b = t >> 4;               // column block of this thread's tile (one per half-warp)
  const int q = t & 15;               // row group
  const int row0 = RT * q;
  const int col0 = CT * b;
  double* rec = smem + L::oRec;
  double* yb_arr = smem + L::oYb;
  double* s_arr = smem + L::oS;
  double* nu_arr = smem + L::oNu;
  double* part1 = smem + L::oPart1;
  double* part2 = smem + L::oPart2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar); // 0 CMD, 1 P1, 2 OPS2, 3 P2
  volatile int* cmd = reinterpret_cast<volatile int*>(smem + L::oCmd);

  double h[TM ? 1 : RT][TM ? 1 : CT]; // the tile: h[i][k] = B⁻¹[row0 + i, col0 + k] (TM: in tensor memory)
  (void)h;

  // ---- the two sweeps over this thread's tile, and the tile's reset and write-out: BOTH roles
  auto pass1 = [&]() { // t[i] = Σ_k y[col0+k] H[row0+i, col0+k], QN.jl:44-49 -> part1
    double tt[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) tt[i] = 0.0;
    const double2* yb2 = reinterpret_cast<const double2*>(yb_arr + b * 16);
    double2 yy[(CT + 1) / 2];
#pragma unroll
    for (int k = 0; k < (CT + 1) / 2; ++k) yy[k] = yb2[k];
    __syncwarp(); // keeps the loads above it
    if constexpr (TM) {
      // pass 1 does 4 DFMA per chunk and keeps few registers: its ring is deeper than pass 2's (TD1 chunks
      // in flight) — with one chunk ahead the sweep is a chain of 15 exposed tensor-memory load latencies
      constexpr int TD1 = QN_TM_DEPTH1;
#if QN_TM_RAW
      TmRaw<EPC> ring[TD1 + 1];
#pragma unroll
      for (int c = 0; c < TD1 && c < NCH; ++c) tm_load_raw(ring[c], tmem + WPC * c);
#else
      TmChunk<EPC> ring[TD1 + 1];
#pragma unroll
      for (int c = 0; c < TD1 && c < NCH; ++c) tm_load(ring[c], tmem + WPC * c);
#endif
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        const double yk = (k & 1) ? yy[k >> 1].y : yy[k >> 1].x;
#pragma unroll
        for (int sc = 0; sc < SPC; ++sc) {
          const int c = k * SPC + sc;
          tm_wait_load();
#if QN_TM_RAW
          TmChunk<EPC> cur;
          tm_unpack(cur, ring[c % (TD1 + 1)]);
          if (c + TD1 < NCH) tm_load_raw(ring[(c + TD1) % (TD1 + 1)], tmem + WPC * (c + TD1));
#else
          if (c + TD1 < NCH) tm_load(ring[(c + TD1) % (TD1 + 1)], tmem + WPC * (c + TD1));
          const TmChunk<EPC>& cur = ring[c % (TD1 + 1)];
#endif
          if (C::kFull || col0 + k < N) {
#pragma unroll
            for (int r = 0; r < EPC; ++r) tt[sc * EPC + r] = fma(yk, cur.d[r], tt[sc * EPC + r]);
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        const double yk = (k & 1) ? yy[k >> 1].y : yy[k >> 1].x;
        if (C::kFull || col0 + k < N) {
#pragma unroll
          for (int i = 0; i < RT; ++i) tt[i] = fma(yk, h[i][k], tt[i]);
        }
      }
    }
    double* pp = part1 + b * C::kQuad + 2 * q;
    *reinterpret_cast<double2*>(pp) = make_double2(tt[0], tt[1]);
    *reinterpret_cast<double2*>(pp + 40) = make_double2(tt[2], tt[3]);
  };
  auto pass2 = [&]() { // H ← H + c₁ssᵀ - usᵀ - suᵀ fused with d = H⁺∇_new, QN.jl:56-67 -> part2
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
    auto ld_wn = [&](const int k) { return *reinterpret_cast<const double2*>(rb_ + k * L::kRec); };
    auto ld_sg = [&](const int k) { return *reinterpret_cast<const double2*>(rb_ + k * L::kRec + 2); };
    if constexpr (TM) {
#if QN_TM_RAW
      TmRaw<EPC> ring[TD + 1];
#pragma unroll
      for (int c = 0; c < TD && c < NCH; ++c) tm_load_raw(ring[c], tmem + WPC * c);
#else
      TmChunk<EPC> ring[TD + 1];
#pragma unroll
      for (int c = 0; c < TD && c < NCH; ++c) tm_load(ring[c], tmem + WPC * c);
#endif
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        const double2 wn = ld_wn(k), sg = ld_sg(k);
#pragma unroll
        for (int sc = 0; sc < SPC; ++sc) {
          const int c = k * SPC + sc;
          tm_wait_load();
#if QN_TM_RAW
          TmChunk<EPC> cur;
          tm_unpack(cur, ring[c % (TD + 1)]);
          if (c + TD < NCH) tm_load_raw(ring[(c + TD) % (TD + 1)], tmem + WPC * (c + TD));
#else
          TmChunk<EPC>& cur = ring[c % (TD + 1)];
          if (c + TD < NCH) tm_load(ring[(c + TD) % (TD + 1)], tmem + WPC * (c + TD));
#endif
          if (C::kFull || col0 + k < N) {
#pragma unroll
            for (int r = 0; r < EPC; ++r) {
              const int i = sc * EPC + r;
              double hh = cur.d[r];
              hh = fma(wn.x, sc_[i], hh);  // + (c₁ s[r]) s[c]
              hh = fma(wn.y, sc_[i], hh);  // - u[r] s[c]
              hh = fma(nuc[i], sg.x, hh);  // - u[c] s[r]
              cur.d[r] = hh;
              acc[i] = fma(hh, sg.y, acc[i]);
            }
          }
          // by the whole warp (.sync.aligned), also where a half-warp has no column k
          if constexpr (QN_TM_STW == 1) {
#pragma unroll
            for (int r = 0; r < EPC; ++r) tm_store1(tmem + WPC * c + 2 * r, cur.d[r]);
          } else if constexpr (QN_TM_STW == 2) {
#pragma unroll
            for (int r = 0; r < EPC; r += 2) tm_store2(tmem + WPC * c + 2 * r, cur.d[r], cur.d[r + 1]);
          } else {
            tm_store(tmem + WPC * c, cur);
          }
        }
      }
      tm_wait_store();
    } else {
      double2 wnr[PF + 1], sgr[PF + 1];
#pragma unroll
      for (int k = 0; k < PF && k < CT; ++k) { wnr[k] = ld_wn(k); sgr[k] = ld_sg(k); }
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (k + PF < CT) { wnr[(k + PF) % (PF + 1)] = ld_wn(k + PF); sgr[(k + PF) % (PF + 1)] = ld_sg(k + PF); }
        const double2 wn = wnr[k % (PF + 1)], sg = sgr[k % (PF + 1)];
        if (C::kFull || col0 + k < N) {
#pragma unroll
          for (int i = 0; i < RT; ++i) {
            double hh = h[i][k];
            hh = fma(wn.x, sc_[i], hh);
            hh = fma(wn.y, sc_[i], hh);
            hh = fma(nuc[i], sg.x, hh);
            h[i][k] = hh;
            acc[i] = fma(hh, sg.y, acc[i]);
          }
        }
      }
    }
    double* pp = part2 + b * C::kQuad + 2 * q;
    *reinterpret_cast<double2*>(pp) = make_double2(acc[0], acc[1]);
    *reinterpret_cast<double2*>(pp + 40) = make_double2(acc[2], acc[3]);
  };
  auto reset_tile = [&]() { // B⁻¹ ← I, QN.jl:144-148
    if constexpr (TM) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        TmChunk<EPC> hc;
#pragma unroll
        for (int r = 0; r < EPC; ++r) hc.d[r] = (row0 + (c % SPC) * EPC + r == col0 + c / SPC) ? 1.0 : 0.0;
        tm_store(tmem + WPC * c, hc);
      }
      tm_wait_store();
    } else {
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int k = 0; k < CT; ++k) h[i][k] = (row0 + i == col0 + k) ? 1.0 : 0.0;
    }
  };
  auto write_h = [&]() { // final B⁻¹ (QN_FLAG_WRITE_H); TM: whole warps take part in the .sync.aligned loads
    if (!(p.h_out && (TM || b < NB))) return;
    double* ho = p.h_out + j * p.h_stride;
    if constexpr (TM) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        TmRaw<EPC> hr;
        tm_load_raw(hr, tmem + WPC * c);
        tm_wait_load();
        TmChunk<EPC> hc;
        tm_unpack(hc, hr);
#pragma unroll
        for (int r = 0; r < EPC; ++r) {
          const int k = c / SPC, i = (c % SPC) * EPC + r;
          if (row0 + i < N && col0 + k < N) ho[(row0 + i) + (long long)p.ldh * (col0 + k)] = hc.d[r];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int k = 0; k < CT; ++k)
          if (row0 + i < N && col0 + k < N) ho[(row0 + i) + (long long)p.ldh * (col0 + k)] = h[i][k];
    }
  };

  // ---- helper: nothing but its half of the sweeps
  if (!leader) {
    for (;;) {
      lh_wait(bars + 0, ph & 1u); ph ^= 1u;
      const int c = *cmd;
      if (c & kLhReset) reset_tile();
      if (c & kLhExit) break;
      pass1();
      lh_arrive(bars + 1);
      lh_wait(bars + 2, (ph >> 2) & 1u); ph ^= 4u;
      pass2();
      lh_arrive(bars + 3);
    }
    write_h();
    return;
  }

  // ---- leader: the vectors (element `lane` in lo, `lane + 32` in hi: the accumulator layout of rule C1)
  // slot of element el in the block-padded column arrays
  auto slot_of = [&](const int el) { return el < N ? (el / CT) * 16 + el % CT : -1; };
  const int slot_lo = slot_of(lane), slot_hi = slot_of(lane + 32);
  const int q_lo = C::quad(lane), q_hi = q_lo + 16; // quad-layout index of elements lane, lane + 32
  auto dot = [&](const Pair& a, const Pair& bb) {   // lane accumulator of rule C1
    double pacc = a.lo * bb.lo;
    if (hi_valid) pacc = pacc + a.hi * bb.hi;
    return pacc;
  };
  auto park = [&](const int off, const Pair& v) { smem[off + q_lo] = v.lo; smem[off + q_hi] = v.hi; };
  auto unpark = [&](const int off) { Pair v; v.lo = smem[off + q_lo]; v.hi = smem[off + q_hi]; return v; };
  auto combine = [&](const double* part) { // Σ_b part[b][e'] in ascending block order (rule C2), e' = lane, lane + 32
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
  auto command = [&](const int c) { // publish a command (and everything written before it) to the helper
    if (lane == 0) *cmd = c;
    __syncwarp();
    lh_arrive(bars + 0);
  };

  // The parameters the loop tests on its critical path, read ONCE per problem: a warp-wide reduction
  // of 32 equal values marks them warp-uniform, so they live in uniform registers instead of being
  // fetched from the constant bank (LDC + dependent compare + branch, ~40 cycles each) every iteration.
#if QN_LH_HOIST
  auto uni32 = [](const int v) { return __reduce_max_sync(0xffffffffu, v); };
  auto uni64 = [&](const double v) {
    return __hiloint2double(__reduce_max_sync(0xffffffffu, __double2hiint(v)), (int)__reduce_max_sync(0xffffffffu, (unsigned)__double2loint(v)));
  };
  LsParams ls;
  ls.c1 = uni64(p.ls.c1); ls.rho_hi = uni64(p.ls.rho_hi); ls.rho_lo = uni64(p.ls.rho_lo);
  ls.iterations = (long long)uni32((int)min(p.ls.iterations, 2000000000LL)); ls.order = uni32(p.ls.order);
  const double tol = uni64(p.tol);
  const int max_iter = uni32(p.max_iter);
  const int ntrace = p.trace != nullptr ? uni32(p.ntrace) : 0;
#else
  const LsParams& ls = p.ls;
  const double tol = p.tol;
  const int max_iter = p.max_iter;
  const int ntrace = p.trace != nullptr ? p.ntrace : 0;
#endif
  Pair x, g, s, y;
  double l0, sy = 0.0;
  {
    const double* th = p.theta0 + j * p.theta_stride; // copyto!(x_old, x), QN.jl:248
    x.lo = th[lane];
    x.hi = hi_valid ? th[lane + 32] : 0.0;
    s.lo = s.hi = y.lo = y.hi = 0.0;
    const double term = rosenbrock_warp_terms<N, true>(lane, x, g); // QN.jl:254 of iteration 1
    l0 = rosenbrock_finish<N>(warp_butterfly(term), x);
  }
  constexpr bool kParkGp = !TM || QN_TM_PARK;
  Pair gp; gp.lo = gp.hi = 0.0;
  auto save_gp = [&](const Pair& v) { if (kParkGp) park(L::oGp, v); else gp = v; };
  auto load_gp = [&]() { return kParkGp ? unpark(L::oGp) : gp; };
  save_gp(gp);
  double result = __longlong_as_double(0x7ff8000000000000LL); // NaN, QN.jl:291
  int iters = 0, fevals = 0, status = QN_MAX_ITER;
  bool helper_reset = false; // the helper's half of the tile still has to be reset (with the next command)

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

  QN_TICK_INIT;
  for (int it = 1; it <= max_iter; ++it) { // QN.jl:250,253
    QN_TICK(0);
    ++fevals;
    const bool small = fabs(g.lo) < tol && (!hi_valid || fabs(g.hi) < tol);
    const bool all_small = __all_sync(0xffffffffu, small);
    if (!isfinite(l0) || all_small) {
      iters = it - 1;
      if (!isfinite(l0)) { status = QN_NONFINITE; if (!(it & 1)) g = load_gp(); }
      else { status = QN_CONVERGED; result = l0; }
      break;
    }
    QN_TICK(1);
    double m;
    Pair d;
    Spec sp;
    Pair xn;
    double lx1 = 0.0;
    bool have_trial = false;
    if (it == 1) {
      m = -1.0; // QN.jl:263-264
    } else {
      // ---- BFGS_update!  QN.jl:34-69.  Operands for BOTH warps' column blocks.
      s_arr[q_lo] = s.lo;
      s_arr[q_hi] = s.hi;
      *reinterpret_cast<double2*>(rec + slot_lo * L::kRec + 2) = make_double2(s.lo, g.lo);
      yb_arr[slot_lo] = y.lo;
      if (slot_hi >= 0) {
        *reinterpret_cast<double2*>(rec + slot_hi * L::kRec + 2) = make_double2(s.hi, g.hi);
        yb_arr[slot_hi] = y.hi;
      }
      constexpr bool kPark = !TM || QN_TM_PARK;
      if (kPark) {
        park(L::oX, x);
        park(L::oG, g);
        park(L::oY, y);
      }
      command(kLhPass | (helper_reset ? kLhReset : 0));
      helper_reset = false;
      QN_TICK(2);
      const double rho = 1.0 / sy; // QN.jl:42 (no guard)
      pass1();
      lh_wait(bars + 1, (ph >> 1) & 1u); ph ^= 2u; // the helper's pass-1 partials
      QN_TICK(3);
      {
        const Pair T = combine(part1);
        const Pair yv = kPark ? unpark(L::oY) : y;
        const double s_lo = s_arr[q_lo], s_hi = s_arr[q_hi];
        const double yHy = warp_butterfly(dot(T, yv)); // Σ_c t[c] y[c], QN.jl:51
        const double c1 = fma(yHy, rho, 1.0) * rho;    // QN.jl:54
        const double nu_lo = -(T.lo * rho), nu_hi = -(T.hi * rho); // -B⁻¹y[c] = -(t ρ), QN.jl:50
        nu_arr[q_lo] = nu_lo;
        nu_arr[q_hi] = nu_hi;
        // (c₁ s[r]) rounded once per r (rule C4)
        *reinterpret_cast<double2*>(rec + slot_lo * L::kRec) = make_double2(c1 * s_lo, nu_lo);
        if (slot_hi >= 0) *reinterpret_cast<double2*>(rec + slot_hi * L::kRec) = make_double2(c1 * s_hi, nu_hi);
      }
      __syncwarp();
      lh_arrive(bars + 2);
      QN_TICK(4);
      pass2();
      lh_wait(bars + 3, (ph >> 3) & 1u); ph ^= 8u; // the helper's pass-2 partials
      QN_TICK(5);
      d = combine(part2); // the new direction d = H⁺∇_new; `s, y = y, s`, QN.jl:65,268
      if (kPark) {
        x = unpark(L::oX);
        g = unpark(L::oG);
      }
      double tterm;
      spec_eval(1.0, d, sp, xn, tterm);
      m = dot(d, g);
      warp_butterfly3(m, tterm, sp.sy, lane);
      sp.term_next = tterm;
      lx1 = rosenbrock_finish<N>(tterm, xn);
      have_trial = true;
      QN_TICK(6);
    }
    if (m <= 0.0) { // QN.jl:272-280: H ← I, s ← ∇_new, m ← ∇ᵀ∇
      reset_tile();
      helper_reset = true;
      d = g;
      m = warp_butterfly(dot(g, g));
      have_trial = false;
    }
    double alpha;
    {
      auto trial = [&](const double a) {
        QN_TICK(7); // the scalar part of the line search up to this trial
        double tterm;
        spec_eval(a, d, sp, xn, tterm);
        warp_butterfly3(tterm, sp.term_next, sp.sy, lane);
        ++fevals;
        const double v = rosenbrock_finish<N>(tterm, xn);
        QN_TICK(8); // one extra trial
        QN_COUNT(9);
        return v;
      };
      if (have_trial) ++fevals;
      alpha = backtracking_linesearch(ls, l0, m, trial, lx1, have_trial, lx1);
    }
    QN_TICK(7);
    if (alpha == 0.0) { status = QN_LINESEARCH_FAILED; iters = it - 1; if (!(it & 1)) g = load_gp(); break; } // QN.jl:284
    s.lo = alpha * d.lo; s.hi = alpha * d.hi; // s ← αs      QN.jl:23-24
    x.lo = x.lo + s.lo;  x.hi = x.hi + s.hi;  // x_old += s  QN.jl:25
    if (it <= ntrace) {
      double* tr = p.trace + (j * p.ntrace + (it - 1)) * (long long)N;
      tr[lane] = x.lo;
      if (hi_valid) tr[lane + 32] = x.hi;
    }
    if (it == max_iter) { iters = it; if (!(it & 1)) g = load_gp(); break; }
    save_gp(g);
    y.lo = g.lo - sp.g.lo; y.hi = g.hi - sp.g.hi; // y = ∇_old - ∇_new, QN.jl:38
    sy = sp.sy;
    g = sp.g;
    l0 = (alpha == 1.0) ? lx1 : rosenbrock_finish<N>(sp.term_next, x);
  }
  QN_TICK_FLUSH;
  command(kLhExit | (helper_reset ? kLhReset : 0));
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
  write_h();
  write_problem_scalars(p, j, lane, result, iters, fevals, status);
}

// registers per thread: 168 lets 6 CTAs stay resident; MINB = 7 asks for 144 (experiment)
template <int N, int MINB, int PF>
__global__ void __launch_bounds__(64) __maxnreg__(MINB >= 7 ? 144 : (MINB == 6 ? 168 : (MINB == 5 ? 200 : 255)))
qn_resident_tile_kernel(const __grid_constant__ BatchParams p) {
  extern __shared__ __align__(16) double smem[];
#if QN_TILE_LH_REG
  volatile long long* next = reinterpret_cast<volatile long long*>(smem + tile_lh_smem_doubles<N>() - 2);
  if (threadIdx.x == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TileLhCfg<N>::oBar);
    for (int k = 0; k < 4; ++k) lh_bar_init(bars + k, 32);
  }
  const bool leader = (threadIdx.x >> 5) == (p.lead_div > 0 ? (blockIdx.x / (unsigned)p.lead_div) & 1u : 0u);
  unsigned ph = 0;
#else
  volatile long long* next = reinterpret_cast<volatile long long*>(smem + tile_smem_doubles<N>() - 2);
#endif
  if (threadIdx.x == 0) stamp_start(p);
  for (;;) {
    __syncthreads(); // the previous problem's shared-memory traffic is finished
    if (threadIdx.x == 0) *next = (long long)atomicAdd(p.counter, 1ULL);
    __syncthreads();
    const long long j = *next;
    if (j >= p.batch) break;
#if QN_TILE_LH_REG
    tile_solve_lh<N, PF, false>(p, j, smem, 0u, leader, ph);
#else
    tile_solve_one<N, PF>(p, j, smem);
#endif
  }
  if (threadIdx.x == 0) stamp_exit(p, (int)blockIdx.x);
}

// The tensor-memory variant: 4 warps = two problems per CTA, 128 columns of tensor memory per CTA,
// 4 CTAs = 8 problems per SM.  Each pair of warps runs its own persistent loop.
template <int N, int PF>
__global__ void __launch_bounds__(128, 4)
qn_resident_tile_tmem_kernel(const __grid_constant__ BatchParams p) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint32_t tmem_base;
#if QN_TILE_LH
  constexpr int kPairDoubles = tile_lh_smem_doubles<N>();
#else
  constexpr int kPairDoubles = tile_smem_doubles<N>();
#endif
  const int pair = (int)(threadIdx.x >> 6);
  double* sm = smem + pair * kPairDoubles;
#if QN_TILE_LH
  if ((threadIdx.x & 63) == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TileLhCfg<N>::oBar);
    for (int k = 0; k < 4; ++k) lh_bar_init(bars + k, 32);
  }
  // Which warp of a pair leads.  The same one in every CTA (lead_div = 0) is the fastest: the
  // helpers' idle polling then shares its two scheduler partitions with other helpers only, not
  // with a leader's dependent chain (n = 60: 69.0 ms alternating per residency slot, 66.2 ms
  // alternating per CTA, 65.3 ms the same everywhere; profiles/r2_lead_parity.txt).
  const bool leader = ((threadIdx.x >> 5) & 1u) == (p.lead_div > 0 ? (blockIdx.x / (unsigned)p.lead_div) & 1u : 0u);
  unsigned ph = 0;
#endif
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n"
                 :: "l"((unsigned long long)__cvta_generic_to_shared(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // this warp's 32 lanes; the warp-wide reduction (of 32 equal values) tells the compiler that the
  // address is warp-uniform, so it stays in a uniform register instead of being moved there
  // (R2UR) in front of every tcgen05.ld/st
  const uint32_t tmem = __reduce_max_sync(0xffffffffu, tmem_base + (((threadIdx.x >> 5) * 32u) << 16));
  volatile long long* next = reinterpret_cast<volatile long long*>(sm + kPairDoubles - 2);
  if (threadIdx.x == 0) stamp_start(p);
  for (;;) {
    pair_barrier(pair); // the previous problem's shared-memory traffic is finished
    if ((threadIdx.x & 63) == 0) *next = (long long)atomicAdd(p.counter, 1ULL);
    pair_barrier(pair);
    const long long j = *next;
    if (j >= p.batch) break;
#if QN_TILE_LH
    tile_solve_lh<N, PF, true>(p, j, sm, tmem, leader, ph);
#else
    tile_solve_one<N, PF, true>(p, j, sm, tmem);
#endif
  }
  if ((threadIdx.x & 63) == 0) stamp_exit(p, 2 * (int)blockIdx.x + pair);
  __syncthreads(); // both problems are done with the tensor memory
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" :: "r"(tmem_base) : "memory");
}

}  // namespace qn
