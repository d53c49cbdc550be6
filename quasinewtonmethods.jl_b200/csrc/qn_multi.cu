// qn_multi.cu — one-process multi-GPU batch sharding:
// problem j goes to context j mod G; the GPUs never communicate.  Every device
// gets its columns with one strided 2-D copy, runs its own persistent grid and
// work queue, and the host gathers the results: θ* and ∇ with one strided copy per device (rows
// of n doubles), the per-problem scalars with ONE contiguous copy per device into pinned host
// staging followed by a host de-interleave (a strided copy of 4- or 8-byte rows would be one DMA
// descriptor per problem).  Every context's mutex is held from staging to the last synchronise.
#include <cstring>
#include <memory>
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
  for (int g = 0; g < G; ++g)
    for (int h = 0; h < g; ++h)
      if (ctxs[g] == ctxs[h]) return qn_fail(QN_ERR_INVALID_ARGUMENT, "context %d is passed twice", g);
  // all contexts locked, in index order, for the whole call (their staging buffers and streams are in use)
  std::vector<std::unique_lock<std::mutex>> locks;
  locks.reserve(G);
  for (int g = 0; g < G; ++g) locks.emplace_back(ctxs[g]->mu);
  struct Shard {
    int64_t count = 0;
    char* base = nullptr;
    char* hbase = nullptr; // pinned host staging: ell | iters | fevals | status, contiguous
    size_t o_theta, o_x, o_g, o_ell, o_it, o_fe, o_st, scalar_bytes;
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
    // the four scalar arrays sit back to back (8-byte aligned), so one copy brings them all home
    s.o_ell = carve(0);
    s.o_it = s.o_ell + sizeof(double) * s.count;
    s.o_fe = s.o_it + sizeof(int32_t) * ((s.count + 1) & ~(int64_t)1);
    s.o_st = s.o_fe + sizeof(int32_t) * ((s.count + 1) & ~(int64_t)1);
    s.scalar_bytes = s.o_st + sizeof(int32_t) * s.count - s.o_ell;
    off = s.o_ell + s.scalar_bytes;
    int rc = qn_ensure_stage(ctx, off);
    if (rc) return rc;
    rc = qn_ensure_hstage(ctx, s.scalar_bytes);
    if (rc) return rc;
    s.base = (char*)ctx->stage;
    s.hbase = (char*)ctx->hstage;
    // columns g, g+G, g+2G, ... of theta0: pitch G*n doubles, width n doubles
    QN_CUDA_TRY(cudaMemcpy2DAsync(s.base + s.o_theta, sizeof(double) * n, theta0 + (size_t)g * n,
                                  sizeof(double) * n * G, sizeof(double) * n, s.count,
                                  cudaMemcpyHostToDevice, ctx->stream));
    rc = qn_optimize_device_locked(
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
    if (ell_out || iters_out || fevals_out || status_out)
      QN_CUDA_TRY(cudaMemcpyAsync(s.hbase, s.base + s.o_ell, s.scalar_bytes, cudaMemcpyDeviceToHost, st));
  }
  for (int g = 0; g < G; ++g) {
    Shard& s = sh[g];
    if (s.count <= 0) continue;
    QN_CUDA_TRY(cudaSetDevice(ctxs[g]->device));
    QN_CUDA_TRY(cudaStreamSynchronize(ctxs[g]->stream));
    // de-interleave this device's scalars: its problem i is problem g + i·G of the batch
    const double* h_ell = (const double*)s.hbase;
    const int32_t* h_it = (const int32_t*)(s.hbase + (s.o_it - s.o_ell));
    const int32_t* h_fe = (const int32_t*)(s.hbase + (s.o_fe - s.o_ell));
    const int32_t* h_st = (const int32_t*)(s.hbase + (s.o_st - s.o_ell));
    if (ell_out) for (int64_t i = 0; i < s.count; ++i) ell_out[g + i * G] = h_ell[i];
    if (iters_out) for (int64_t i = 0; i < s.count; ++i) iters_out[g + i * G] = h_it[i];
    if (fevals_out) for (int64_t i = 0; i < s.count; ++i) fevals_out[g + i * G] = h_fe[i];
    if (status_out) for (int64_t i = 0; i < s.count; ++i) status_out[g + i * G] = h_st[i];
  }
  return QN_OK;
}
