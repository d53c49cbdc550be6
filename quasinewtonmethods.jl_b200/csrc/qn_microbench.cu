// qn_microbench.cu — microbenchmarks that MEASURE the roofline denominators the
// resident path is reported against: sustained fp64 FMA rate,
// shared-memory load bandwidth, and HBM copy bandwidth.  MEASURED_PEAKS.json
// (driver-written) carries only HBM and bf16-tensor peaks; the BFGS kernels
// are bounded by the FP64 pipe and shared memory instead.
#include "qn_internal.h"

namespace {

// 8 independent DFMA chains per thread; 256 threads; many CTAs per SM.
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// conflict-free LDS.128: lane l of a warp reads the 16 bytes at (window + l)*16, so a warp
// request is 512 consecutive bytes = 4 wavefronts of 128 B.  The window moves with every
// load and with the loop counter, so nothing can be hoisted or folded (same pattern as
// tools/ubench.cu, which measured 4.01 SM-cycles per such load = 127.6 B/clk/SM).
__global__ void __launch_bounds__(256) lds_kernel(double* out, int iters) {
  __shared__ __align__(16) unsigned char buf[16384];
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<int*>(buf)[i] = i;
  __syncthreads();
  const unsigned off = (threadIdx.x & 31) * 16;
  long long bits = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const unsigned a = (off + (unsigned)(k * 528 + i * 16)) & (16384 - 16);
      const longlong2 v = *reinterpret_cast<const longlong2*>(buf + a);
      bits ^= v.x ^ v.y; // integer xor keeps the loads alive without touching the FP64 pipe
    }
  }
  if (bits == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = (double)bits;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src,
                                                   double2* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

template <class F>
int time_best(cudaStream_t st, int reps, float* best_ms, F&& launch) {
  cudaEvent_t e0, e1;
  QN_CUDA_TRY(cudaEventCreate(&e0));
  QN_CUDA_TRY(cudaEventCreate(&e1));
  *best_ms = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {
    QN_CUDA_TRY(cudaEventRecord(e0, st));
    launch();
    QN_CUDA_TRY(cudaEventRecord(e1, st));
    QN_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    QN_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2 && ms < *best_ms) *best_ms = ms; // first two are warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  QN_CUDA_TRY(cudaGetLastError());
  return QN_OK;
}

}  // namespace

extern "C" int32_t qn_microbench_dfma(qn_context* ctx, double* tflops) {
  if (!ctx || !tflops) return qn_fail(QN_ERR_INVALID_ARGUMENT, "NULL argument");
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  const int grid = ctx->prop.multiProcessorCount * 8, block = 256, iters = 2048;
  double* out = nullptr;
  QN_CUDA_TRY(cudaMalloc(&out, sizeof(double) * grid * block));
  float ms = 0;
  int rc = time_best(ctx->stream, 5, &ms, [&] { dfma_kernel<<<grid, block, 0, ctx->stream>>>(out, iters, 0.999999, 1e-9); });
  cudaFree(out);
  if (rc) return rc;
  const double fmas = (double)grid * block * iters * 64.0;
  *tflops = 2.0 * fmas / (ms * 1e-3) / 1e12;
  return QN_OK;
}

extern "C" int32_t qn_microbench_lds(qn_context* ctx, double* gbs) {
  if (!ctx || !gbs) return qn_fail(QN_ERR_INVALID_ARGUMENT, "NULL argument");
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  const int grid = ctx->prop.multiProcessorCount * 8, block = 256, iters = 2048;
  double* out = nullptr;
  QN_CUDA_TRY(cudaMalloc(&out, sizeof(double) * grid * block));
  float ms = 0;
  int rc = time_best(ctx->stream, 5, &ms, [&] { lds_kernel<<<grid, block, 0, ctx->stream>>>(out, iters); });
  cudaFree(out);
  if (rc) return rc;
  const double bytes = (double)grid * block * iters * 16.0 * 16.0;
  *gbs = bytes / (ms * 1e-3) / 1e9;
  return QN_OK;
}

extern "C" int32_t qn_microbench_hbm(qn_context* ctx, double* gbs) {
  if (!ctx || !gbs) return qn_fail(QN_ERR_INVALID_ARGUMENT, "NULL argument");
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)1 << 30; // 1 GiB each way, far larger than the 126 MB L2
  double2 *a = nullptr, *b = nullptr;
  QN_CUDA_TRY(cudaMalloc(&a, bytes));
  if (cudaMalloc(&b, bytes) != cudaSuccess) {
    cudaFree(a);
    return qn_fail(QN_ERR_CUDA, "cudaMalloc failed in qn_microbench_hbm");
  }
  cudaMemsetAsync(a, 1, bytes, ctx->stream);
  cudaMemsetAsync(b, 2, bytes, ctx->stream);
  const size_t n = bytes / sizeof(double2);
  float ms = 0;
  int rc = time_best(ctx->stream, 5, &ms, [&] {
    copy_kernel<<<ctx->prop.multiProcessorCount * 16, 256, 0, ctx->stream>>>(a, b, n);
  });
  cudaFree(a);
  cudaFree(b);
  if (rc) return rc;
  *gbs = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
  return QN_OK;
}
