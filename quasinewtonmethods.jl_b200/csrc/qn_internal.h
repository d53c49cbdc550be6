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

struct qn_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  unsigned long long* counter = nullptr; // ring of work-queue counters
  int counter_slot = 0;
  static constexpr int kCounterSlots = 64;
  void* stage = nullptr; // grow-only device staging for the host-pointer entry points
  size_t stage_bytes = 0;
  int last_path = -1, last_grid = 0, last_block = 0, last_smem = 0, last_launches = 0;
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

// grow the context's device staging buffer (caller holds ctx->mu or is single-threaded)
int qn_ensure_stage(qn_context* ctx, size_t bytes);
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
                    int onepass);
