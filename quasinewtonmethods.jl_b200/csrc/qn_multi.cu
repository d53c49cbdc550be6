// qn_multi.cu — one-process multi-GPU batch sharding:
// problem j goes to context j mod G; the GPUs never communicate.  Every device
// gets its columns with one strided 2-D copy, runs its own persistent grid and
// work queue, and the host gathers the results with strided copies back.
#include <vector>

#include "qn_internal.h"

extern "C" int32_t qn_bfgs_optimize_batched_multi_f64(
    qn_context* const* ctxs, int32_t nctx, const qn_model* const* models, int32_t n, int64_t batch,
    const double* theta0, const qn_backtracking* ls, double tol, int32_t max_iter, double* x_out,
    double* g_out, double* ell_out, int32_t* iters_out, int32_t* fevals_out, int32_t* status_out,
    uint32_t flags) {
  if (!ctxs || !models || nctx < 1) return qn_fail(QN_ERR_INVALID_ARGUMENT, "need >= 1 context");
  if (n < 1 || batch < 0 || (batch > 0 && !theta0) || !ls)
    return qn_fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  for (int g = 0; g < nctx; ++g)
    if (!ctxs[g] || !models[g]) return qn_fail(QN_ERR_INVALID_ARGUMENT, "NULL context/model %d", g);
  if (batch == 0) return QN_OK;

  const int G = nctx;
  struct Shard {
    int64_t count = 0;
    char* base = nullptr;
    size_t o_theta, o_x, o_g, o_ell, o_it, o_fe, o_st;
  };
  std::vector<Shard> sh(G);
  // phase 1: scatter (interleaved columns) and launch on every device, no waiting
  for (int g = 0; g < G; ++g) {
    qn_context* ctx = ctxs[g];
    Shard& s = sh[g];
    s.count = (batch - g + G - 1) / G;
    if (s.count <= 0) continue;
    QN_CUDA_TRY(cudaSetDevice(ctx->device));
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t nb = sizeof(double) * (size_t)n * s.count;
    s.o_theta = carve(nb); s.o_x = carve(nb); s.o_g = carve(nb);
    s.o_ell = carve(sizeof(double) * s.count);
    s.o_it = carve(sizeof(int32_t) * s.count);
    s.o_fe = carve(sizeof(int32_t) * s.count);
    s.o_st = carve(sizeof(int32_t) * s.count);
    int rc = qn_ensure_stage(ctx, off);
    if (rc) return rc;
    s.base = (char*)ctx->stage;
    // columns g, g+G, g+2G, ... of theta0: pitch G*n doubles, width n doubles
    QN_CUDA_TRY(cudaMemcpy2DAsync(s.base + s.o_theta, sizeof(double) * n, theta0 + (size_t)g * n,
                                  sizeof(double) * n * G, sizeof(double) * n, s.count,
                                  cudaMemcpyHostToDevice, ctx->stream));
    rc = qn_bfgs_optimize_batched_device_f64(
        ctx, models[g], n, s.count, (const double*)(s.base + s.o_theta), n, ls, tol, max_iter,
        nullptr, 0, 0, (double*)(s.base + s.o_x), (double*)(s.base + s.o_g),
        (double*)(s.base + s.o_ell), (int32_t*)(s.base + s.o_it), (int32_t*)(s.base + s.o_fe),
        (int32_t*)(s.base + s.o_st), 0, nullptr, flags, ctx->stream);
    if (rc) return rc;
  }
  // phase 2: gather
  for (int g = 0; g < G; ++g) {
    qn_context* ctx = ctxs[g];
    Shard& s = sh[g];
    if (s.count <= 0) continue;
    QN_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (x_out)
      QN_CUDA_TRY(cudaMemcpy2DAsync(x_out + (size_t)g * n, sizeof(double) * n * G, s.base + s.o_x,
                                    sizeof(double) * n, sizeof(double) * n, s.count,
                                    cudaMemcpyDeviceToHost, st));
    if (g_out)
      QN_CUDA_TRY(cudaMemcpy2DAsync(g_out + (size_t)g * n, sizeof(double) * n * G, s.base + s.o_g,
                                    sizeof(double) * n, sizeof(double) * n, s.count,
                                    cudaMemcpyDeviceToHost, st));
    if (ell_out)
      QN_CUDA_TRY(cudaMemcpy2DAsync(ell_out + g, sizeof(double) * G, s.base + s.o_ell, sizeof(double),
                                    sizeof(double), s.count, cudaMemcpyDeviceToHost, st));
    if (iters_out)
      QN_CUDA_TRY(cudaMemcpy2DAsync(iters_out + g, sizeof(int32_t) * G, s.base + s.o_it,
                                    sizeof(int32_t), sizeof(int32_t), s.count,
                                    cudaMemcpyDeviceToHost, st));
    if (fevals_out)
      QN_CUDA_TRY(cudaMemcpy2DAsync(fevals_out + g, sizeof(int32_t) * G, s.base + s.o_fe,
                                    sizeof(int32_t), sizeof(int32_t), s.count,
                                    cudaMemcpyDeviceToHost, st));
    if (status_out)
      QN_CUDA_TRY(cudaMemcpy2DAsync(status_out + g, sizeof(int32_t) * G, s.base + s.o_st,
                                    sizeof(int32_t), sizeof(int32_t), s.count,
                                    cudaMemcpyDeviceToHost, st));
  }
  for (int g = 0; g < G; ++g) {
    if (sh[g].count <= 0) continue;
    QN_CUDA_TRY(cudaSetDevice(ctxs[g]->device));
    QN_CUDA_TRY(cudaStreamSynchronize(ctxs[g]->stream));
  }
  return QN_OK;
}
