// qn_internal.h — host-side internals shared by the .cu files of libqn_cuda.so.
#pragma once

#include <cuda_runtime.h>

#include <mutex>

#include "../../include/qn_cuda.h"

// Records a thread-local error message (printf-style) and returns `code`.
int qn_fail(int code, const char* fmt, ...);

#define QN_CUDA_TRY(expr)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return qn_fail(QN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                 \
  } while (0)

// Tuning / ablation switches of one context (qn_set_option; the QN_* environment variables of the
// same names are read ONCE, in qn_create, as defaults — dispatch never calls getenv).
struct qn_options {
  int tile_tmem = -1;     // "tile_tmem":  -1 auto, 0 tile in registers, 1 tile in tensor memory   (QN_TILE_TMEM)
  int tile_3w = 0;        // "tile_3w":    tensor-memory tiles with three warps per problem and a speculative pass 1; measured 3 % SLOWER than leader/helper pairs, off by default (QN_TILE_3W)
  int pack = 1;           // "pack":        several small problems per warp (n = 8, 16, 32)        (QN_NO_PACK clears)
  int pack_variant = 0;   // "pack_variant": rows per lane 1 | 2 | 4, 0 = measured default         (QN_PACK_VARIANT)
  int pair_rows = 0;      // "pair_rows":   0 auto, else 'h' | 't' | '4' | '2' | 's'                 (QN_PAIR_ROWS)
  int smem_rows = 0;      // "smem_rows":   0 auto, 2 or 4 rows per thread                          (QN_SMEM_ROWS)
  int lambda_smem = 1;    // "lambda_smem": dense Gaussian n = 128 keeps Λ in shared memory         (QN_NO_LAMBDA_SMEM clears)
  int stream_cluster = 1; // "stream_cluster": cluster vector kernel of the streaming path          (QN_STREAM_NO_CLUSTER clears)
  int stream_tma = -1;    // "stream_tma":  -1 auto, 0 LDG sweeps, 1 cp.async.bulk-staged sweeps    (QN_STREAM_TMA)
  int l2_persist_mb = 0;  // "l2_persist_mb": leading part of B⁻¹ kept resident in L2 by the streaming path; -1 device maximum, 0 off = default: measured SLOWER beyond 16 MB (QN_L2_PERSIST_MB)
  int l2_resident = -1;   // "l2_resident": sweeps without eviction hints when B⁻¹ (or this GPU's slab) fits the L2; -1 auto, 0, 1 (QN_L2_RESIDENT)
  int sweep_fine = -1;    // "sweep_fine": one rule-C2 block per sweep CTA instead of a super-block; -1 auto (small grids), 0, 1 (QN_SWEEP_FINE)
  int sweep_order = -1;   // "sweep_order": -1 default, 0 ascending, 1 pass1 descending, 2 pass2 descending (QN_SWEEP_ORDER)
  int sweep_depth = 0;    // "sweep_depth": 0 auto, 8 or 16 loads in flight                         (QN_SWEEP_DEPTH)
  int hint_p1_ld = 0, hint_p2_ld = 2, hint_p2_st = 1, hint_one_ld = 4; // cache hints, QN_HINT_*
  int p1_pad_kb = 0;      // "p1_pad_kb"                                                            (QN_P1_PAD_KB)
  int host_chunks = -1;   // "host_chunks": host-pointer entry: column ranges pipelined over two streams, -1 auto (QN_HOST_CHUNKS)
  int lead_parity = 0;    // "lead_parity": tile kernels, which warp of a pair leads: 0 the same in every CTA, d > 0 warp (blockIdx / d) & 1, -1 d = #SM (QN_LEAD_PARITY)
  int lead_warp = -1;     // "lead_warp": n = 96, 128, 160 tiles (one CTA per problem): the leading warp, -1 auto (QN_LEAD_WARP)
  int record_exit = 0;    // "record_exit": tile kernels stamp %globaltimer when a problem slot runs dry
};

struct qn_context {
  int device = 0;
  qn_options opt;
  static constexpr int kRangeStreams = 8;
  cudaStream_t range_stream[kRangeStreams] = {}; // one per pipelined column range of a host-pointer entry
  cudaStream_t stream = nullptr;                 // = range_stream[0]: the context's stream
  cudaDeviceProp prop{};
  unsigned long long* counter = nullptr; // ring of work-queue counters
  int counter_slot = 0;
  static constexpr int kCounterSlots = 64;
  void* stage = nullptr; // grow-only device staging for the host-pointer entry points
  size_t stage_bytes = 0;
  void* hstage = nullptr; // grow-only PINNED host staging (multi-GPU gather of the per-problem scalars)
  size_t hstage_bytes = 0;
  unsigned long long* exit_ns = nullptr; // [kExitSlots + 2] start, end and per-slot exit stamps (record_exit)
  static constexpr int kExitSlots = 4096;
  int exit_slots_used = 0;
  int last_path = -1, last_grid = 0, last_block = 0, last_smem = 0, last_launches = 0;
  long long last_l2_persist_bytes = 0; // streaming path: size of the L2-persisting window of the last solve
  std::mutex mu;
};

struct qn_model {
  qn_context* ctx = nullptr;
  int kind = -1;
  int n = 0; // 0 = any dimension
  int m = 0;
  double* d_a = nullptr; // gaussian: mu      | logistic: X
  double* d_b = nullptr; // gaussian: lambda  | logistic: Xt
  double* d_c = nullptr; //                   | logistic: y
  double lam = 0.0;
  size_t bytes = 0;      // user objective: size of the parameter blob at d_a
};

// grow the context's device / pinned-host staging buffer (caller holds ctx->mu)
int qn_ensure_stage(qn_context* ctx, size_t bytes);
int qn_ensure_hstage(qn_context* ctx, size_t bytes);
// qn_bfgs_optimize_batched_device_f64 for a caller that already holds ctx->mu
int qn_optimize_device_locked(qn_context* ctx, const qn_model* model, int32_t n, int64_t batch,
                              const double* theta0, int64_t theta_stride, const qn_backtracking* ls,
                              double tol, int32_t max_iter, double* states, int64_t state_stride,
                              int32_t ld, double* x_out, double* g_out, double* ell_out,
                              int32_t* iters_out, int32_t* fevals_out, int32_t* status_out,
                              int32_t ntrace, double* trace_out, uint32_t flags, cudaStream_t st);
// how many column ranges a host-pointer entry pipelines `batch` problems in (option "host_chunks")
int qn_host_ranges(const qn_context* ctx, int64_t batch, int ntrace);
// next zeroed work-queue counter, zeroing enqueued on `st`
unsigned long long* qn_next_counter(qn_context* ctx, cudaStream_t st);

// L = align(P, pick_vector_width(P, Float64)), QN.jl:96-99
inline int qn_host_state_ld(int n, int native_w) {
  int w = 1;
  while (w < n && w < native_w) w <<= 1;
  return ((n + w - 1) / w) * w;
}

// qn_stream.cu: one large problem with B⁻¹ in HBM (device pointers; blocks until finished).
int qn_stream_solve(qn_context* ctx, int n, const double* d_theta0, const qn_backtracking* ls,
                    double tol, int max_iter, double* d_x_out, double* d_g_out, double* d_h_out,
                    int ldh_out, double* ell, int* iters, int* fevals, int* status, int ntrace,
                    double* d_trace, cudaStream_t st, float* ms_per_iter_out, int fixed_iters,
                    int onepass, int obj_kind = QN_OBJ_ROSENBROCK, const double* obj_a = nullptr,
                    const double* obj_b = nullptr);
