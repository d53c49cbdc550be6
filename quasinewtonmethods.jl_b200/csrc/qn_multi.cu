// qn_multi.cu — one-process multi-GPU batch sharding:
// problem j goes to context j mod G; the GPUs never communicate.  Every device
// gets its columns with one strided 2-D copy, runs its own persistent grid and
// work queue, and the host gathers the results: θ* and ∇ with one strided copy per device (rows
// of n doubles), the per-problem scalars with ONE contiguous copy per device into pinned host
// staging followed by a host de-interleave (a strided copy of 4- or 8-byte rows would be one DMA
// descriptor per problem).  A device's share is pipelined in ranges over the context's streams
// (option "host_chunks").  Every context's mutex is held from staging to the last synchronise.
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
  // Every device's share is cut into ranges (option "host_chunks"), each on a stream of its own, so
  // the copies of one range run under the kernels of the others and the kernel of range k+1 fills
  // the SM slots range k frees while its last problems finish.
  struct Range {
    int64_t i0 = 0, count = 0; // local problems [i0, i0+count) of the device
    size_t o_theta, o_x, o_g, o_ell, o_it, o_fe, o_st, o_h, scalar_bytes;
  };
  struct Shard {
    int64_t count = 0;
    int ranges = 0;
    char* base = nullptr;
    char* hbase = nullptr; // pinned host staging: per range ell | iters | fevals | status, contiguous
    Range r[qn_context::kRangeStreams];
  };
  std::vector<Shard> sh(G);
  int max_ranges = 0;
  for (int g = 0; g < G; ++g) {
    qn_context* ctx = ctxs[g];
    Shard& s = sh[g];
    s.count = (batch - g + G - 1) / G;
    if (s.count <= 0) continue;
    s.ranges = qn_host_ranges(ctx, s.count, 0);
    if (s.ranges > max_ranges) max_ranges = s.ranges;
    QN_CUDA_TRY(cudaSetDevice(ctx->device));
    size_t off = 0, hoff = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    for (int c = 0; c < s.ranges; ++c) {
      Range& r = s.r[c];
      r.i0 = s.count * c / s.ranges;
      r.count = s.count * (c + 1) / s.ranges - r.i0;
      const size_t nb = sizeof(double) * (size_t)n * r.count;
      r.o_theta = carve(nb); r.o_x = carve(nb); r.o_g = carve(nb);
      // the four scalar arrays sit back to back (8-byte aligned), so one copy brings them all home
      r.o_ell = carve(0);
      r.o_it = r.o_ell + sizeof(double) * r.count;
      r.o_fe = r.o_it + sizeof(int32_t) * ((r.count + 1) & ~(int64_t)1);
      r.o_st = r.o_fe + sizeof(int32_t) * ((r.count + 1) & ~(int64_t)1);
      r.scalar_bytes = r.o_st + sizeof(int32_t) * ((r.count + 1) & ~(int64_t)1) - r.o_ell;
      off = r.o_ell + r.scalar_bytes;
      carve(0);
      r.o_h = hoff;
      hoff += r.scalar_bytes;
    }
    int rc = qn_ensure_stage(ctx, off);
    if (rc) return rc;
    rc = qn_ensure_hstage(ctx, hoff);
    if (rc) return rc;
    s.base = (char*)ctx->stage;
    s.hbase = (char*)ctx->hstage;
  }
  // phase 1: scatter (interleaved columns) and launch, range by range on every device, no waiting
  for (int c = 0; c < max_ranges; ++c)
    for (int g = 0; g < G; ++g) {
      qn_context* ctx = ctxs[g];
      Shard& s = sh[g];
      if (c >= s.ranges) continue;
      const Range& r = s.r[c];
      QN_CUDA_TRY(cudaSetDevice(ctx->device));
      cudaStream_t st = ctx->range_stream[c];
      // columns g + (i0+i)·G of theta0: pitch G*n doubles, width n doubles
      QN_CUDA_TRY(cudaMemcpy2DAsync(s.base + r.o_theta, sizeof(double) * n,
                                    theta0 + ((size_t)g + (size_t)r.i0 * G) * n, sizeof(double) * n * G,
                                    sizeof(double) * n, r.count, cudaMemcpyHostToDevice, st));
      int rc = qn_optimize_device_locked(
          ctx, models[g], n, r.count, (const double*)(s.base + r.o_theta), n, ls, tol, max_iter,
          nullptr, 0, 0, (double*)(s.base + r.o_x), (double*)(s.base + r.o_g),
          (double*)(s.base + r.o_ell), (int32_t*)(s.base + r.o_it), (int32_t*)(s.base + r.o_fe),
          (int32_t*)(s.base + r.o_st), 0, nullptr, flags, st);
      if (rc) return rc;
    }
  // phase 2: gather
  for (int c = 0; c < max_ranges; ++c)
    for (int g = 0; g < G; ++g) {
      qn_context* ctx = ctxs[g];
      Shard& s = sh[g];
      if (c >= s.ranges) continue;
      const Range& r = s.r[c];
      QN_CUDA_TRY(cudaSetDevice(ctx->device));
      cudaStream_t st = ctx->range_stream[c];
      const size_t col = ((size_t)g + (size_t)r.i0 * G) * n;
      if (x_out)
        QN_CUDA_TRY(cudaMemcpy2DAsync(x_out + col, sizeof(double) * n * G, s.base + r.o_x,
                                      sizeof(double) * n, sizeof(double) * n, r.count,
                                      cudaMemcpyDeviceToHost, st));
      if (g_out)
        QN_CUDA_TRY(cudaMemcpy2DAsync(g_out + col, sizeof(double) * n * G, s.base + r.o_g,
                                      sizeof(double) * n, sizeof(double) * n, r.count,
                                      cudaMemcpyDeviceToHost, st));
      if (ell_out || iters_out || fevals_out || status_out)
        QN_CUDA_TRY(cudaMemcpyAsync(s.hbase + r.o_h, s.base + r.o_ell, r.scalar_bytes,
                                    cudaMemcpyDeviceToHost, st));
    }
  for (int c = 0; c < max_ranges; ++c)
    for (int g = 0; g < G; ++g) {
      Shard& s = sh[g];
      if (c >= s.ranges) continue;
      const Range& r = s.r[c];
      QN_CUDA_TRY(cudaSetDevice(ctxs[g]->device));
      QN_CUDA_TRY(cudaStreamSynchronize(ctxs[g]->range_stream[c]));
      // de-interleave this range's scalars: local problem i0+i is problem g + (i0+i)·G of the batch
      const char* hb = s.hbase + r.o_h;
      const double* h_ell = (const double*)hb;
      const int32_t* h_it = (const int32_t*)(hb + (r.o_it - r.o_ell));
      const int32_t* h_fe = (const int32_t*)(hb + (r.o_fe - r.o_ell));
      const int32_t* h_st = (const int32_t*)(hb + (r.o_st - r.o_ell));
      const int64_t j0 = g + r.i0 * G;
      if (ell_out) for (int64_t i = 0; i < r.count; ++i) ell_out[j0 + i * G] = h_ell[i];
      if (iters_out) for (int64_t i = 0; i < r.count; ++i) iters_out[j0 + i * G] = h_it[i];
      if (fevals_out) for (int64_t i = 0; i < r.count; ++i) fevals_out[j0 + i * G] = h_fe[i];
      if (status_out) for (int64_t i = 0; i < r.count; ++i) status_out[j0 + i * G] = h_st[i];
    }
  return QN_OK;
}
