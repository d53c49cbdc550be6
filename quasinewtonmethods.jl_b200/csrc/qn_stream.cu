// qn_stream.cu — the large-n path (BASELINE.json configs[4], n in the thousands): B⁻¹ lives
// in HBM and one problem uses the whole GPU.
//
// All control flow stays on the device: one BFGS iteration (QN.jl:253-288) is a fixed sequence
// of kernels, eight iterations are captured in a CUDA graph and replayed, a `done` flag turns the
// remaining launches into no-ops, and the host polls that flag after every replay.
//
// Two-pass mode (the reference's order, bit-exact with the oracle), three kernels per iteration:
//   pass1    (grid)     block partials of t = B⁻¹y; its last CTA forms yᵀB⁻¹y and c₁       QN.jl:44-54
//   pass2    (grid)     B⁻¹ ← B⁻¹ + c₁ssᵀ - usᵀ - suᵀ fused with partials of d = B⁻¹∇       QN.jl:56-67
//   cluster  (8 CTAs)   m, reset decision, the whole line search, update_state!, then the
//                       head of the next iteration: ∂logdensity!, convergence, y, sᵀy, ρ   QN.jl:66-68,272-286,254-264,36-42
// One-pass mode (QN_FLAG_ONE_PASS): one sweep (pending update + both GEMVs) and the cluster kernel.
// Beyond the cluster kernel's reach (n > 24 577) the original sequence of single-CTA vector
// kernels runs instead: vec_begin, pass1, vec_mid, pass2, vec_end, identity.
// Row-sharded mode: the same graphs on up to 8 GPUs, each sweeping its own rows of B⁻¹ and
// storing its row sums into every replica (peer stores + epoch flags, qn_stream_kernels.cuh).
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
#include <cstring>
#include <vector>

#include "qn_device.cuh"
#include "qn_internal.h"

#include "qn_stream_kernels.cuh"

#ifndef QN_SWEEP_ORDER_DEFAULT
#define QN_SWEEP_ORDER_DEFAULT 0
#endif


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
  size_t persist_bytes = 0; // leading bytes of H inside the L2-persisting access-policy window
  cudaStream_t window_stream = nullptr; // the stream that carries the window attribute
};

// The L2-persisting window of one work space: the leading `persist_sb` super-blocks of B⁻¹ (a
// contiguous run of columns) get the persisting access property on every sweep launch, so they are
// read from and written to the 126 MB L2 instead of HBM in every iteration after the first
// (B200_PROFILING: "keep tensors resident instead of re-reading them").  Option "l2_persist_mb":
// -1 = as much as the device allows (cudaDevAttrMaxPersistingL2CacheSize), 0 = off.
void window_value(const StreamWork& w, cudaLaunchAttributeValue& v) {
  memset(&v, 0, sizeof v);
  v.accessPolicyWindow.base_ptr = w.p.H;
  v.accessPolicyWindow.num_bytes = w.persist_bytes;
  v.accessPolicyWindow.hitRatio = 1.0f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
}
int persist_setup(qn_context* ctx, StreamWork& w) {
  qn::StreamParams& p = w.p;
  p.persist_sb = 0;
  w.persist_bytes = 0;
  // A slab that fits the L2 with room to spare (every shard of the row-sharded mode at n = 8192 on 8
  // GPUs: 64 MiB of 126 MB; one GPU up to n ≈ 2 800) needs no window at all: its sweeps just drop
  // their eviction hints (.lu loads, .cs stores) and the matrix stays cached from sweep to sweep
  // and from iteration to iteration.  Option "l2_resident": -1 auto, 0 never, 1 always.
  {
    const size_t hb = sizeof(double) * (size_t)p.ld * p.n;
    const int v = ctx->opt.l2_resident;
    if (v > 0 || (v < 0 && hb * 20 <= (size_t)ctx->prop.l2CacheSize * 11)) { p.persist_sb = p.nsb; return QN_OK; }
  }
  const long long want_mb = ctx->opt.l2_persist_mb;
  if (want_mb == 0) return QN_OK;
  const size_t cap = (size_t)ctx->prop.persistingL2CacheMaxSize;
  const size_t maxwin = (size_t)ctx->prop.accessPolicyMaxWindowSize;
  if (cap == 0 || maxwin == 0) return QN_OK;
  size_t want = want_mb < 0 ? cap : std::min(cap, (size_t)want_mb << 20);
  want = std::min(want, maxwin);
  const size_t sb_bytes = sizeof(double) * (size_t)p.ld * p.sbcols;
  const size_t h_bytes = sizeof(double) * (size_t)p.ld * p.n;
  int k = (int)std::min<size_t>((size_t)p.nsb, want / sb_bytes);
  if (k == 0) return QN_OK;
  p.persist_sb = k;
  w.persist_bytes = std::min(h_bytes, (size_t)k * sb_bytes);
  // set aside exactly what the window needs: the set-aside is taken from the L2 of every other access
  QN_CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(cap, w.persist_bytes)));
  ctx->last_l2_persist_bytes = (long long)w.persist_bytes;
  return QN_OK;
}
int persist_apply_stream(StreamWork& w, cudaStream_t st) {
  if (!w.persist_bytes) return QN_OK;
  w.window_stream = st;
  cudaLaunchAttributeValue v;
  window_value(w, v);
  QN_CUDA_TRY(cudaStreamSetAttribute(st, cudaLaunchAttributeAccessPolicyWindow, &v));
  return QN_OK;
}
int persist_apply_graph(const StreamWork& w) { // every kernel node of the captured graph
  if (!w.persist_bytes) return QN_OK;
  size_t nn = 0;
  QN_CUDA_TRY(cudaGraphGetNodes(w.graph, nullptr, &nn));
  std::vector<cudaGraphNode_t> nodes(nn);
  QN_CUDA_TRY(cudaGraphGetNodes(w.graph, nodes.data(), &nn));
  cudaLaunchAttributeValue v;
  window_value(w, v);
  for (size_t i = 0; i < nn; ++i) {
    cudaGraphNodeType t;
    QN_CUDA_TRY(cudaGraphNodeGetType(nodes[i], &t));
    if (t == cudaGraphNodeTypeKernel)
      QN_CUDA_TRY(cudaGraphKernelNodeSetAttribute(nodes[i], cudaLaunchAttributeAccessPolicyWindow, &v));
  }
  return QN_OK;
}
void persist_release(const StreamWork& w) { // the device of the work space is current
  if (!w.persist_bytes) return;
  cudaLaunchAttributeValue v;
  memset(&v, 0, sizeof v);
  if (w.window_stream) cudaStreamSetAttribute(w.window_stream, cudaLaunchAttributeAccessPolicyWindow, &v); // no window
  cudaCtxResetPersistingL2Cache();
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0); // give the set-aside back
}

using PassKernel = void (*)(const qn::StreamParams);

// cache hints of the B⁻¹ accesses per kernel: measured defaults (profiles/r1_stream_cache_hints.txt,
// n = 8192: pass2 .lu loads + .cs stores 356 -> 316 µs per iteration; one-pass sweep .cg loads
// 216 -> 206 µs).  The full 5×4 sweep of hints was run once (scripts/exp_stream_hints.sh, git
// history); the combinations worth keeping are instantiated and selectable for A/B runs.
// The hints are options of the context (qn_set_option "hint_p1_ld" ...; the QN_HINT_* environment
// variables are read once, in qn_create).  t_opt: the options of the context this thread is serving.
thread_local const qn_options* t_opt = nullptr;
const qn_options& opts() { static const qn_options dflt; return t_opt ? *t_opt : dflt; }
int hint_opt(int k, int dflt, int hi) { return (k < 0 || k > hi) ? dflt : k; }
// deep = 16 instead of 8 loads in flight per thread (a whole rule-C2 block): for grids that leave
// most SMs with at most one CTA — the shards of the row-sharded mode — where each thread's walk
// over its 256 columns is a chain of exposed memory latencies
bool sweep_deep(const qn_context* ctx, const qn::StreamParams& p) {
  const int v = hint_opt(ctx->opt.sweep_depth, 0, 16);
  if (v) return v == 16;
  if (p.grp_cols < p.sbcols) return true; // one block per CTA: its <= 16 columns are one batch of loads
  return p.ntiles * p.nsb <= ctx->prop.multiProcessorCount;
}
// option "sweep_order": 0 both sweeps walk their column blocks in ascending order (round 1), 1 pass1
// descending, 2 pass2 descending — a descending sweep starts where the ascending one before it ended,
// on what that one left in L2 (single-GPU, 8-deep sweeps only)
int sweep_order() { return hint_opt(opts().sweep_order, QN_SWEEP_ORDER_DEFAULT, 2); }
PassKernel pass1_kernel(bool deep = false) {
  if (deep) return qn::stream_pass1<0, 0, 16>;
  if (sweep_order() == 1) return qn::stream_pass1<0, 0, 8, true>;
  return hint_opt(opts().hint_p1_ld, 0, 4) == 4 ? qn::stream_pass1<4, 0> : qn::stream_pass1<0, 0>;
}
PassKernel pass2_kernel(bool deep = false) {
  if (deep) return qn::stream_pass2<2, 1, 16>;
  const int ld = hint_opt(opts().hint_p2_ld, 2, 4), st = hint_opt(opts().hint_p2_st, 1, 3);
  if (sweep_order() == 2) return qn::stream_pass2<2, 1, 8, true>;
  if (ld == 0 && st == 0) return qn::stream_pass2<0, 0>; // no hints
  if (ld == 0) return qn::stream_pass2<0, 1>;            // .cs stores only
  if (st == 0) return qn::stream_pass2<2, 0>;            // .lu loads only
  return qn::stream_pass2<2, 1>;                         // default: .lu loads + .cs stores
}
PassKernel pass_one_kernel(bool deep = false) {
  if (deep) return qn::stream_pass_one<4, 0, 16>;
  const int ld = hint_opt(opts().hint_one_ld, 4, 4);
  if (ld == 0) return qn::stream_pass_one<0, 0>;         // no hints
  if (ld == 1) return qn::stream_pass_one<1, 0>;         // .cs loads
  return qn::stream_pass_one<4, 0>;                      // default: .cg loads
}

// option "stream_tma": the two-pass sweeps staged through shared memory by the bulk-copy engine
// (stream_pass1_tma / stream_pass2_tma); -1 = the measured default
#ifndef QN_STREAM_TMA_DEFAULT
#define QN_STREAM_TMA_DEFAULT 0
#endif
// auto (-1) = QN_STREAM_TMA_DEFAULT = off.  Measured (profiles/r2_stream_tma_vs_ldg.txt, r2_tma_small.txt):
// equal to the LDG.128 sweeps for 512-row tiles (4 KB per bulk copy, n >= 4096) and SLOWER for the
// shorter tiles of small n and of the row-sharded slabs (n = 1024: 77 against 47 µs per iteration) —
// one elected thread issues one bulk copy per 0.5-1 KB column run, and that issue rate, not the
// bytes in flight, bounds the sweep; a 2-D tensor map (one copy per stage) would lift it.
bool sweep_tma(const qn_context* ctx, const qn::StreamParams& p, int pass_threads) {
  const int v = ctx->opt.stream_tma;
  if (p.onepass) return false;
  return v < 0 ? (QN_STREAM_TMA_DEFAULT != 0 && pass_threads <= 128) : v != 0;
}
// the ring shape for a tile of 2·threads rows: <= 96 KB, two CTAs per SM
void tma_shape(int pass_threads, int& cols, int& stages) {
  stages = 3;
  cols = pass_threads >= 256 ? 8 : pass_threads >= 128 ? 16 : 32;
}
PassKernel pass1_tma_kernel(int pass_threads) {
  return pass_threads >= 256 ? qn::stream_pass1_tma<8, 3> : pass_threads >= 128 ? qn::stream_pass1_tma<16, 3> : qn::stream_pass1_tma<32, 3>;
}
PassKernel pass2_tma_kernel(int pass_threads) {
  return pass_threads >= 256 ? qn::stream_pass2_tma<8, 3> : pass_threads >= 128 ? qn::stream_pass2_tma<16, 3> : qn::stream_pass2_tma<32, 3>;
}
// extra dynamic shared memory for pass1, a residency knob (bytes): QN_P1_PAD_KB
size_t pass1_pad() { return (size_t)hint_opt(opts().p1_pad_kb, 0, 200) * 1024; }
using VecKernel = void (*)(const qn::StreamParams);
// the single-CTA vector kernels of the registered objective (the objective registry of this path)
VecKernel vec_begin_kernel(int kind) {
  return kind == QN_OBJ_DENSE_GAUSSIAN ? qn::stream_vec_begin<qn::SGaussian> : qn::stream_vec_begin<qn::SRosenbrock>;
}
VecKernel vec_end_kernel(int kind) {
  return kind == QN_OBJ_DENSE_GAUSSIAN ? qn::stream_vec_end<qn::SGaussian> : qn::stream_vec_end<qn::SRosenbrock>;
}
VecKernel cluster_kernel(int ppt, bool onepass) {
  if (onepass) return ppt == 1 ? qn::stream_vec_cluster<1, true> : ppt == 2 ? qn::stream_vec_cluster<2, true> : qn::stream_vec_cluster<3, true>;
  return ppt == 1 ? qn::stream_vec_cluster<1, false> : ppt == 2 ? qn::stream_vec_cluster<2, false> : qn::stream_vec_cluster<3, false>;
}

int enqueue_iteration(qn_context* ctx, const StreamWork& w, cudaStream_t st) {
  t_opt = &ctx->opt;
  const qn::StreamParams& p = w.p;
  const dim3 grid(p.ntiles, p.ngrp);
  const bool deep = sweep_deep(ctx, p);
  if (p.onepass) { // two kernels per iteration; the head of iteration 1 ran before the graph
    pass_one_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * 6 * p.sbcols, st>>>(p);
    if (w.clu_ppt) cluster_kernel(w.clu_ppt, true)<<<qn::kCluCtas, qn::kCluThreads, w.clu_smem, st>>>(p);
    else qn::stream_vec_onepass<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
    QN_CUDA_TRY(cudaGetLastError());
    return QN_OK;
  }
  if (p.fused2) { // three kernels per iteration; the head of iteration 1 ran before the graph
    if (sweep_tma(ctx, p, w.pass_threads)) {
      int tc, ts;
      tma_shape(w.pass_threads, tc, ts);
      pass1_tma_kernel(w.pass_threads)<<<grid, w.pass_threads, qn::tma_smem_bytes(p.sbcols, 1, 2 * w.pass_threads, tc, ts), st>>>(p);
      pass2_tma_kernel(w.pass_threads)<<<grid, w.pass_threads, qn::tma_smem_bytes(p.sbcols, 4, 2 * w.pass_threads, tc, ts), st>>>(p);
    } else {
      pass1_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * p.sbcols + pass1_pad(), st>>>(p);
      pass2_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * 4 * p.sbcols, st>>>(p);
    }
    cluster_kernel(w.clu_ppt, false)<<<qn::kCluCtas, qn::kCluThreads, w.clu_smem, st>>>(p);
    QN_CUDA_TRY(cudaGetLastError());
    return QN_OK;
  }
  vec_begin_kernel(p.obj_kind)<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
  pass1_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * p.sbcols, st>>>(p);
  qn::stream_vec_mid<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
  pass2_kernel(deep)<<<grid, w.pass_threads, sizeof(double) * 4 * p.sbcols, st>>>(p);
  vec_end_kernel(p.obj_kind)<<<1, qn::kVecThreads, w.vec_smem, st>>>(p);
  qn::stream_identity<false><<<ctx->prop.multiProcessorCount * 8, 256, 0, st>>>(p);
  QN_CUDA_TRY(cudaGetLastError());
  return QN_OK;
}

constexpr int kItersPerGraph = 8; // iterations per graph replay; also the cadence of the done-flag poll

void stream_release(StreamWork& w) {
  persist_release(w);
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
                 int onepass, int shard, int nshard, cudaStream_t st, int obj_kind = QN_OBJ_ROSENBROCK,
                 const double* obj_a = nullptr, const double* obj_b = nullptr) {
  using namespace qn;
  StreamParams& p = w.p;
  p.obj_kind = obj_kind;
  p.obj_a = obj_a;
  p.obj_b = obj_b;
  if (obj_kind != QN_OBJ_ROSENBROCK) onepass = 0; // the deferred-update order is a permission; only Rosenbrock has its kernels
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
  p.grp_cols = p.sbcols;
  p.ngrp = p.nsb;
  auto tiles_of = [&](int threads) { return (p.rows + 2 * threads - 1) / (2 * threads); };
  // A grid of super-block CTAs that leaves SMs idle (small n; the slabs of the row-sharded mode): one
  // rule-C2 BLOCK per CTA instead — 16× the CTAs, every load of a thread in flight at once (option
  // "sweep_fine": -1 auto, 0, 1; two-pass mode only)
  // Measured (profiles/r2_stream_sweep_fine.txt): n = 512 43.5 -> 30.7 µs per iteration, n = 1024 46.9 -> 35.9;
  // from n = 2048 on the super-block grid already has >= 128 CTAs and one block per CTA LOSES (66 against
  // 47 µs; n = 6144 332 against 184: 16x the CTAs, each paying its prologue, its arrival atomic and a
  // two-level tail over n/16 partials) — so auto means: only while even 128-row tiles leave half the SMs idle.
  const bool fine = !onepass && (ctx->opt.sweep_fine > 0 ||
                                 (ctx->opt.sweep_fine < 0 && tiles_of(64) * p.nsb <= ctx->prop.multiProcessorCount / 2));
  if (fine) {
    p.grp_cols = p.rb;
    p.ngrp = (n + p.rb - 1) / p.rb;
  }
  w.pass_threads = kPassThreads;
  while (w.pass_threads > 64 && tiles_of(w.pass_threads) * p.ngrp < 2 * ctx->prop.multiProcessorCount)
    w.pass_threads /= 2;
  p.ntiles = tiles_of(w.pass_threads);
  p.ld = (p.rows + 15) & ~15;
  QN_TRY_REL(cudaMalloc(&p.H, sizeof(double) * (size_t)p.ld * n));
  { const int prc = persist_setup(ctx, w); if (prc) return prc; }
  const size_t nv = ((size_t)n + 31) & ~(size_t)31;
  p.nvstride = (int)nv;
  const size_t words = nv * 22 + 2 * (size_t)p.ngrp * n + 192 + 4 * (size_t)row_tiles; // + one counter per 128-row CTA tile
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
  // Everything a PEER addresses (the vectors, the flags) sits in front of the partial-sum arrays, whose
  // size depends on this shard's own choice of sweep granularity (ngrp): the row-sharded mode computes
  // peer addresses as peer_base + (address - base), so those offsets must be the same on every shard.
  double* tail = base + 22 * nv;
  p.sc = reinterpret_cast<StreamScalars*>(tail);          // 32 doubles reserved
  p.flags = reinterpret_cast<unsigned int*>(tail + 32);   // 2·kMaxShards words
  p.tiles_done = reinterpret_cast<unsigned int*>(tail + 48);
  p.tile_count = reinterpret_cast<unsigned int*>(tail + 64); // 4 counters per 512-row tile (one per 128-row CTA tile)
  p.partial = tail + 64 + (((size_t)4 * row_tiles + 63) & ~(size_t)63);
  p.partial2 = p.partial + (size_t)p.ngrp * n;
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
  t_opt = &ctx->opt;
  const qn::StreamParams& p = w.p;
  const int n = p.n;
  QN_TRY_REL(cudaSetDevice(ctx->device));
  const bool deep = sweep_deep(ctx, p);
  QN_TRY_REL(cudaFuncSetAttribute(pass2_kernel(deep), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double) * 4 * p.sbcols)));
  if (pass1_pad())
    QN_TRY_REL(cudaFuncSetAttribute(pass1_kernel(deep), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(sizeof(double) * p.sbcols + pass1_pad())));
  if (sweep_tma(ctx, p, w.pass_threads)) {
    int tc, ts;
    tma_shape(w.pass_threads, tc, ts);
    QN_TRY_REL(cudaFuncSetAttribute(pass1_tma_kernel(w.pass_threads), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)qn::tma_smem_bytes(p.sbcols, 1, 2 * w.pass_threads, tc, ts)));
    QN_TRY_REL(cudaFuncSetAttribute(pass2_tma_kernel(w.pass_threads), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)qn::tma_smem_bytes(p.sbcols, 4, 2 * w.pass_threads, tc, ts)));
  }
  if (w.vec_smem) {
    QN_TRY_REL(cudaFuncSetAttribute(vec_begin_kernel(p.obj_kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    if (p.obj_kind != QN_OBJ_ROSENBROCK)
      QN_TRY_REL(cudaFuncSetAttribute(vec_begin_kernel(QN_OBJ_ROSENBROCK), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    QN_TRY_REL(cudaFuncSetAttribute(qn::stream_vec_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    QN_TRY_REL(cudaFuncSetAttribute(vec_end_kernel(p.obj_kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
    QN_TRY_REL(cudaFuncSetAttribute(qn::stream_vec_onepass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.vec_smem));
  }
  if (ctx->opt.stream_cluster && p.obj_kind == QN_OBJ_ROSENBROCK) {
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
  { const int prc = persist_apply_stream(w, st); if (prc) return prc; }
  QN_TRY_REL(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = QN_OK;
  for (int k = 0; k < kItersPerGraph && rc == QN_OK; ++k) rc = enqueue_iteration(ctx, w, st);
  const cudaError_t ce = cudaStreamEndCapture(st, &w.graph);
  if (rc != QN_OK) return rc;
  if (ce != cudaSuccess) return qn_fail(QN_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
  { const int prc = persist_apply_graph(w); if (prc) return prc; }
  QN_TRY_REL(cudaGraphInstantiate(&w.exec, w.graph, 0));
  if (p.onepass || p.fused2) vec_begin_kernel(p.obj_kind)<<<1, qn::kVecThreads, w.vec_smem, st>>>(p); // head of iteration 1
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
                    int onepass, int obj_kind, const double* obj_a, const double* obj_b) {
  using namespace qn;
  if (obj_kind != QN_OBJ_ROSENBROCK) onepass = 0;
  StreamWork w;
  StreamScalars host{};
  host.result = NAN;
  host.status = QN_MAX_ITER;
  int rc = stream_setup(ctx, w, n, d_theta0, false, ls, tol, max_iter, ntrace, d_trace, onepass, 0, 1, st,
                        obj_kind, obj_a, obj_b);
  if (rc == QN_OK && max_iter > 0) rc = stream_capture(ctx, w, st);
  if (rc == QN_OK && max_iter > 0) rc = stream_run(&ctx, &w, &st, 1, max_iter, fixed_iters, ms_per_iter_out, host);
  if (rc != QN_OK) { stream_release(w); return rc; }
  const StreamParams& p = w.p;
  ctx->last_path = 2;
  ctx->last_grid = p.ntiles * p.ngrp;
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
    ctxs[g]->last_grid = ws[g].p.ntiles * ws[g].p.ngrp;
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
