// qn_capi.cu — the C ABI of libqn_cuda.so (include/qn_cuda.h) and the host-side
// dispatch onto the sm_100a kernels.  No CPU fallback exists anywhere in this
// file: every entry either runs a CUDA kernel or returns an error.
#include <cuda_runtime.h>

#include <algorithm>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <type_traits>
#include <vector>

#include "../../include/qn_cuda.h"
#include "qn_device.cuh"
#include "qn_objectives.cuh"
#ifdef QN_USER_OBJECTIVE_HEADER
#include QN_USER_OBJECTIVE_HEADER  // the build's plug-in objective (include/qn_objective.cuh)
#endif
#include "qn_resident.cuh"
#include "qn_internal.h"
#include "qn_launch.h"

namespace {
thread_local char g_err[512] = "";
}  // namespace

int qn_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define fail qn_fail

int qn_ensure_stage(qn_context* ctx, size_t bytes) {
  if (bytes <= ctx->stage_bytes) return QN_OK;
  if (ctx->stage) cudaFree(ctx->stage);
  ctx->stage = nullptr;
  ctx->stage_bytes = 0;
  QN_CUDA_TRY(cudaMalloc(&ctx->stage, bytes));
  ctx->stage_bytes = bytes;
  return QN_OK;
}

int qn_ensure_hstage(qn_context* ctx, size_t bytes) {
  if (bytes <= ctx->hstage_bytes) return QN_OK;
  if (ctx->hstage) cudaFreeHost(ctx->hstage);
  ctx->hstage = nullptr;
  ctx->hstage_bytes = 0;
  QN_CUDA_TRY(cudaMallocHost(&ctx->hstage, bytes));
  ctx->hstage_bytes = bytes;
  return QN_OK;
}

unsigned long long* qn_next_counter(qn_context* ctx, cudaStream_t st) {
  unsigned long long* c = ctx->counter + ctx->counter_slot;
  ctx->counter_slot = (ctx->counter_slot + 1) % qn_context::kCounterSlots;
  cudaMemsetAsync(c, 0, sizeof(unsigned long long), st);
  return c;
}

// Number of column ranges a host-pointer entry pipelines a batch of `batch` problems in.
int qn_host_ranges(const qn_context* ctx, int64_t batch, int ntrace) {
  int r = ctx->opt.host_chunks;
  // auto: ranges of at least 2048 problems — a range must still fill the SMs' resident problem
  // slots (1184 on B200 for the tensor-memory tiles), or the dispatch would pick a smaller kernel
  if (r < 0) r = (int)std::min<int64_t>(batch / 2048, qn_context::kRangeStreams);
  if (r < 1 || ntrace > 0 || ctx->opt.record_exit) r = 1;
  if (r > qn_context::kRangeStreams) r = qn_context::kRangeStreams;
  if ((int64_t)r > batch) r = (int)batch;
  return r < 1 ? 1 : r;
}

namespace {

// ---- kernel launchers ------------------------------------------------------
// ONEPASS: the deferred-update order (QN_FLAG_ONE_PASS) — shared-memory and register-pair policies
// with more threads than elements only; a request that cannot be honoured (too small an n, or no
// room for the second partial-sum array) silently runs the reference order.
template <class Obj, int NREG, int BLOCK, int MINB, int PAIRN = 0, bool ONEPASS = false, int PROWS = 2,
          int PREG = PROWS, bool PTM = false>
int launch_resident(qn_context* ctx, qn::BatchParams& p, const typename Obj::Params& op,
                    cudaStream_t st, int* probe_occ = nullptr) {
  // probe_occ != nullptr: do not launch, report how many CTAs (= problems) would be resident per SM
  const int n = p.n;
  // shared-memory policy: a thread owns a PAIR of rows of B⁻¹ (one 16-byte access per column)
  // and the rule-C2 column blocks are dealt out to G thread groups of B/2 threads, as many as
  // the CTA's thread limit and the partial-sum array beside B⁻¹ allow
  const int B = qn::resident_block_threads(n);
  int nthreads = (NREG > 0) ? BLOCK : B;
  if (PAIRN > 0) { // register-tile policy: PROWS rows × one rule-C2 block per thread
    nthreads = (B / PROWS) * ((n + qn::chain_block(n) - 1) / qn::chain_block(n));
    nthreads = (nthreads + 31) & ~31; // whole warps (the reductions shuffle with a full mask)
  } else if (NREG == 0 && n >= 48) {
    const int half = B / PROWS;
    const int nblocks = (n + qn::chain_block(n) - 1) / qn::chain_block(n);
    for (int G = std::min(nblocks, BLOCK / half); G >= 2; --G) {
      const int cand = ((G * half + 31) / 32) * 32;
      if (cand <= B) break;
      const size_t need = sizeof(double) * (size_t)qn::resident_smem_layout<Obj>(n, true, cand, op, ONEPASS, PROWS).total;
      if (cand <= BLOCK && need <= (size_t)ctx->prop.sharedMemPerBlockOptin) { nthreads = cand; break; }
    }
  }
  if (ONEPASS && PAIRN == 0 && nthreads <= B) // no row-tile policy possible: the reference order
    return launch_resident<Obj, NREG, BLOCK, MINB, PAIRN, false, PROWS>(ctx, p, op, st, probe_occ);
  if (NREG == 0 && PAIRN == 0 && PROWS == 4 && nthreads <= B) // 4-row groups need more threads than elements
    return launch_resident<Obj, 0, 1024, 1, 0, ONEPASS, 2>(ctx, p, op, st, probe_occ);
  if (NREG == 0 && PAIRN == 0 && BLOCK > 512 && nthreads <= 512) // the 512-thread build may use 128 registers
    return launch_resident<Obj, 0, 512, 1, 0, ONEPASS, PROWS>(ctx, p, op, st, probe_occ);
  const qn::ResidentSmem L = qn::resident_smem_layout<Obj>(n, NREG == 0 && PAIRN == 0, nthreads, op, ONEPASS, PROWS,
                                                           (PAIRN > 0 && !PTM) ? PROWS - PREG : 0);
  const size_t smem = sizeof(double) * (size_t)L.total;
  if (smem > (size_t)ctx->prop.sharedMemPerBlockOptin)
    return fail(QN_ERR_UNSUPPORTED_SIZE,
                "n=%d needs %zu bytes of shared memory per problem; the device allows %zu "
                "(use the streaming path)",
                n, smem, (size_t)ctx->prop.sharedMemPerBlockOptin);
  auto kern = qn::qn_resident_kernel<Obj, NREG, BLOCK, MINB, PAIRN, ONEPASS, PROWS, PREG, PTM>;
  QN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int block = nthreads;
  int occ = 0;
  QN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
  if (PTM) occ = tmem_residency(ctx, kern, block, smem, 128); // the occupancy query answers 1 for a kernel that allocates tensor memory
  if (probe_occ) { *probe_occ = occ; return QN_OK; }
  if (occ < 1) return fail(QN_ERR_UNSUPPORTED_SIZE, "resident kernel does not fit on an SM (n=%d)", n);
  long long grid = (long long)occ * ctx->prop.multiProcessorCount;
  if (grid > p.batch) grid = p.batch;
  if (grid < 1) grid = 1;
  p.counter = qn_next_counter(ctx, st);
  kern<<<(unsigned)grid, block, smem, st>>>(p, op);
  QN_CUDA_TRY(cudaGetLastError());
  ctx->last_path = (NREG > 0 || PAIRN > 0) ? 0 : 1;
  ctx->last_grid = (int)grid;
  ctx->last_block = block;
  ctx->last_smem = (int)smem;
  ctx->last_launches = 1;
  return QN_OK;
}

template <class Obj>
int dispatch_size(qn_context* ctx, qn::BatchParams& p, const typename Obj::Params& op,
                  uint32_t flags, cudaStream_t st) {
  const int n = p.n;
  if constexpr (Obj::kRegisterPaths) {
    if (!(flags & QN_FLAG_FORCE_SMEM)) {
      if constexpr (std::is_same<Obj, qn::Rosenbrock>::value) {
        // the Rosenbrock-only kernels (packed small n, 4×CT tiles, leader/helper tensor-memory tiles)
        int rc = qn_launch_rosenbrock_small(ctx, p, st, (flags & QN_FLAG_ONE_PASS) != 0);
        if (rc >= 0) return rc;
        rc = qn_launch_rosenbrock_tiles(ctx, p, st);
        if (rc >= 0) return rc;
      }
      switch (n) {
#define QN_REG_CASE(N, BLOCK, MINB) \
  case N: return launch_resident<Obj, N, BLOCK, MINB>(ctx, p, op, st);
#define QN_TILE_CASE(N, MINB)
#include "qn_reg_sizes.inc"
#undef QN_REG_CASE
#undef QN_TILE_CASE
        default: break;
      }
    }
  }
  // Register tiles beyond the 4×15-tile sizes (qn_resident.cuh, PAIRN): a thread owns 4 rows × one
  // rule-C2 block.  option "pair_rows" selects the variant: 'h' hybrid — two of the four rows in registers,
  // two in shared memory, which fits every n <= 160 and makes TWO problems resident per SM up to
  // n = 128 (Rosenbrock: n = 128 6.1e7 iterations/s, n = 96 7.0e7 against 4.1e7 and n = 160 2.2e7
  // against 1.8e7 with B⁻¹ in shared memory) —, '4' all four rows in registers (n = 128 only: 5.4e7), '2' 2×16 tiles of 512 threads
  // (n = 128 only: 4.9e7).  Default: hybrid for cheap objectives at every 64 < n <= 160; for the
  // others the second resident problem buys nothing (dense Gaussian 2.21e7 against 2.19e7) and
  // all-register tiles at n = 128 leave the shared memory to the objective (Λ).  't' (n = 128, the
  // default for cheap objectives): the hybrid tiles with their two other rows in TENSOR MEMORY instead
  // of shared memory (tcgen05.ld/st .x4 per tile column, the next column in flight): 6.95e7 against
  // 6.05e7 iterations/s, 16 KB instead of 82 KB of shared memory per problem.
  const bool onepass = (flags & QN_FLAG_ONE_PASS) != 0; // honoured by these policies (and the streaming path)
  if constexpr (Obj::kPairRegister) {
    if (!(flags & QN_FLAG_FORCE_SMEM)) {
      const bool lh_size = (n == 96 || n == 128 || n == 160);
      const char variant = ctx->opt.pair_rows ? (char)ctx->opt.pair_rows
                                              : (Obj::kRegisterPaths ? (lh_size && !onepass ? 'L' : 'h') : '4');
      // n = 128 (BASELINE.json configs[2], [3]) has its own instantiations with n at compile time
      // (hybrid 6.3e7 against 4.7e7 iterations/s with n at run time; dense Gaussian 1.04e6 against
      // 0.68e6 solves/s); the other sizes share the run-time-n build
      if (variant == 't' && n == 128 && !onepass) // hybrid tiles, the other two rows in tensor memory
        return launch_resident<Obj, 0, 256, 2, 128, false, 4, 2, true>(ctx, p, op, st);
      if ((variant == 'h' || variant == 't') && n == 128)
        return onepass ? launch_resident<Obj, 0, 256, 2, 128, true, 4, 2>(ctx, p, op, st)
                       : launch_resident<Obj, 0, 256, 2, 128, false, 4, 2>(ctx, p, op, st);
      if (variant == 'h' && n > 64 && n <= 160) {
        const int B = qn::resident_block_threads(n);
        const int threads = ((B / 4) * ((n + qn::chain_block(n) - 1) / qn::chain_block(n)) + 31) & ~31;
        if (threads <= 256)
          return onepass ? launch_resident<Obj, 0, 256, 2, 1, true, 4, 2>(ctx, p, op, st)
                         : launch_resident<Obj, 0, 256, 2, 1, false, 4, 2>(ctx, p, op, st);
        return onepass ? launch_resident<Obj, 0, 512, 1, 1, true, 4, 2>(ctx, p, op, st)
                       : launch_resident<Obj, 0, 512, 1, 1, false, 4, 2>(ctx, p, op, st);
      }
      if (n == 128 && variant == '2')
        return onepass ? launch_resident<Obj, 0, 512, 1, 128, true>(ctx, p, op, st)
                       : launch_resident<Obj, 0, 512, 1, 128>(ctx, p, op, st);
      if (n == 128 && variant != 's')
        return onepass ? launch_resident<Obj, 0, 256, 1, 128, true, 4>(ctx, p, op, st)
                       : launch_resident<Obj, 0, 256, 1, 128, false, 4>(ctx, p, op, st);
    }
  }
  if (n <= 256) {
    // shared-memory policy: 2 or 4 rows of B⁻¹ per thread.  Four rows halve the threads (and so the
    // registers) of a CTA: worth it exactly when that lets MORE problems be resident per SM (measured:
    // n = 96, 100: 2 instead of 1, +23-25 %; with equal residency two rows are 10-15 % faster).  Cheap
    // objectives only — the GEMVs of the others want the threads.  option "smem_rows" = 2 | 4 forces a choice.
    bool four = false;
    if (n >= 64) {
      if (ctx->opt.smem_rows) {
        four = (ctx->opt.smem_rows == 4);
      } else if (Obj::kRegisterPaths) {
        int occ2 = 0, occ4 = 0;
        const int r2 = onepass ? launch_resident<Obj, 0, 1024, 1, 0, true, 2>(ctx, p, op, st, &occ2)
                               : launch_resident<Obj, 0, 1024, 1, 0, false, 2>(ctx, p, op, st, &occ2);
        const int r4 = onepass ? launch_resident<Obj, 0, 1024, 1, 0, true, 4>(ctx, p, op, st, &occ4)
                               : launch_resident<Obj, 0, 1024, 1, 0, false, 4>(ctx, p, op, st, &occ4);
        four = (r2 == QN_OK && r4 == QN_OK && occ4 > occ2);
      }
    }
    if (four)
      return onepass ? launch_resident<Obj, 0, 1024, 1, 0, true, 4>(ctx, p, op, st)
                     : launch_resident<Obj, 0, 1024, 1, 0, false, 4>(ctx, p, op, st);
    return onepass ? launch_resident<Obj, 0, 1024, 1, 0, true>(ctx, p, op, st)
                   : launch_resident<Obj, 0, 1024, 1>(ctx, p, op, st);
  }
  return fail(QN_ERR_UNSUPPORTED_SIZE, "n=%d does not fit the resident path", n);
}

// per-problem scalars of a streamed solve, written on the stream the solve ran on
__global__ void stream_scalars_kernel(double* ell_out, int* iters_out, int* fevals_out, int* status_out,
                                      long long j, double ell, int it, int fe, int status) {
  if (ell_out) ell_out[j] = ell;
  if (iters_out) iters_out[j] = it;
  if (fevals_out) fevals_out[j] = fe;
  if (status_out) status_out[j] = status;
}

// Large n: B⁻¹ in HBM, one problem at a time on the whole GPU (qn_stream.cu).  This path is
// SYNCHRONOUS (the host polls the convergence flag between graph replays): it waits for the
// caller's stream, runs on the context's stream and returns when the batch is finished.
int dispatch_stream(qn_context* ctx, const qn_model* model, qn::BatchParams& p,
                    const qn_backtracking* ls, uint32_t flags, cudaStream_t st) {
  // the objective registry of this path: Rosenbrock (cluster vector kernel, one-pass mode) and the
  // dense Gaussian (CTA-cooperative objective in the single-CTA vector kernels)
  if (model->kind != QN_OBJ_ROSENBROCK && model->kind != QN_OBJ_DENSE_GAUSSIAN)
    return fail(QN_ERR_UNREGISTERED_OBJECTIVE,
                "objective kind %d has no registered device implementation on the streaming path "
                "(n=%d); Rosenbrock and the dense Gaussian have one", model->kind, p.n);
  if (model->kind == QN_OBJ_DENSE_GAUSSIAN && model->n != p.n)
    return fail(QN_ERR_INVALID_ARGUMENT, "dense Gaussian model has n=%d, call has n=%d", model->n, p.n);
  // the graph cannot be captured on the legacy default stream: run on the context's stream,
  // ordered after whatever the caller enqueued on `st`
  cudaStream_t run = ctx->stream;
  if (st != run) QN_CUDA_TRY(cudaStreamSynchronize(st));
  int launches = 0;
  for (long long j = 0; j < p.batch; ++j) {
    double ell;
    int it, fe, status;
    int rc = qn_stream_solve(ctx, p.n, p.theta0 + j * p.theta_stride, ls, p.tol, p.max_iter,
                             p.x_out ? p.x_out + j * p.x_stride : nullptr,
                             p.g_out ? p.g_out + j * p.g_stride : nullptr,
                             p.h_out ? p.h_out + j * p.h_stride : nullptr, p.ldh, &ell, &it, &fe,
                             &status, p.ntrace,
                             p.trace ? p.trace + (size_t)j * p.ntrace * p.n : nullptr, run, nullptr, 0,
                             (flags & QN_FLAG_ONE_PASS) ? 1 : 0, model->kind, model->d_a, model->d_b);
    if (rc) return rc;
    launches += ctx->last_launches;
    stream_scalars_kernel<<<1, 1, 0, run>>>(p.ell_out, p.iters_out, p.fevals_out, p.status_out, j, ell, it, fe, status);
    QN_CUDA_TRY(cudaGetLastError());
  }
  if (st != run) QN_CUDA_TRY(cudaStreamSynchronize(run)); // results are complete when the call returns
  ctx->last_launches = launches;
  return QN_OK;
}

int dispatch_model(qn_context* ctx, const qn_model* model, qn::BatchParams& p,
                   const qn_backtracking* ls, uint32_t flags, cudaStream_t st) {
  // resident up to n = 160 (205 KB of B⁻¹ + vectors in 227 KB of shared memory), HBM beyond
  if ((flags & QN_FLAG_FORCE_STREAM) || p.n > 160) return dispatch_stream(ctx, model, p, ls, flags, st);
  switch (model->kind) {
    case QN_OBJ_ROSENBROCK:
      return dispatch_size<qn::Rosenbrock>(ctx, p, qn::Rosenbrock::Params{}, flags, st);
    case QN_OBJ_DENSE_GAUSSIAN: {
      if (model->n != p.n)
        return fail(QN_ERR_INVALID_ARGUMENT, "dense Gaussian model has n=%d, call has n=%d",
                    model->n, p.n);
      if (!(flags & (QN_FLAG_FORCE_SMEM | QN_FLAG_ONE_PASS))) { // n = 96, 128, 160: leader/helper tensor-memory tiles
        const int rc = qn_launch_gaussian_tilew(ctx, p, model->d_a, model->d_b, st);
        if (rc >= 0) return rc;
      }
      qn::DenseGaussian::Params op{model->d_a, model->d_b, 0};
      // register-pair policy (n = 128): shared memory is free, keep Λ there (128 KB of the 227 KB)
      if (p.n == 128 && !(flags & QN_FLAG_FORCE_SMEM) && ctx->opt.lambda_smem) op.lam_in_smem = 1;
      return dispatch_size<qn::DenseGaussian>(ctx, p, op, flags, st);
    }
    case QN_OBJ_LOGISTIC: {
      if (model->n != p.n)
        return fail(QN_ERR_INVALID_ARGUMENT, "logistic model has n=%d, call has n=%d", model->n,
                    p.n);
      qn::Logistic::Params op{model->m, model->d_a, model->d_b, model->d_c, model->lam};
      return dispatch_size<qn::Logistic>(ctx, p, op, flags, st);
    }
#ifdef QN_USER_OBJECTIVE_HEADER
    case QN_OBJ_USER:
      if (model->n && model->n != p.n)
        return fail(QN_ERR_INVALID_ARGUMENT, "user objective was registered for n=%d, call has n=%d", model->n, p.n);
      return dispatch_size<qn::UserObjective>(ctx, p, qn::UserObjective::make(model->d_a, model->bytes, p.n), flags, st);
#endif
    default:
      return fail(QN_ERR_UNREGISTERED_OBJECTIVE,
                  "objective kind %d has no registered device implementation", model->kind);
  }
}

// option table: key -> field of qn_options (qn_set_option / qn_get_option; env defaults in qn_create)
struct OptKey { const char* key; int qn_options::*field; const char* env; int env_mode; }; // env_mode 0: atoi, 1: presence clears, 2: first char
const OptKey kOptKeys[] = {
  {"tile_tmem", &qn_options::tile_tmem, "QN_TILE_TMEM", 0},
  {"tile_3w", &qn_options::tile_3w, "QN_TILE_3W", 0},
  {"pack", &qn_options::pack, "QN_NO_PACK", 1},
  {"pack_variant", &qn_options::pack_variant, "QN_PACK_VARIANT", 0},
  {"pair_rows", &qn_options::pair_rows, "QN_PAIR_ROWS", 2},
  {"smem_rows", &qn_options::smem_rows, "QN_SMEM_ROWS", 0},
  {"lambda_smem", &qn_options::lambda_smem, "QN_NO_LAMBDA_SMEM", 1},
  {"stream_cluster", &qn_options::stream_cluster, "QN_STREAM_NO_CLUSTER", 1},
  {"stream_tma", &qn_options::stream_tma, "QN_STREAM_TMA", 0},
  {"l2_persist_mb", &qn_options::l2_persist_mb, "QN_L2_PERSIST_MB", 0},
  {"l2_resident", &qn_options::l2_resident, "QN_L2_RESIDENT", 0},
  {"sweep_fine", &qn_options::sweep_fine, "QN_SWEEP_FINE", 0},
  {"sweep_order", &qn_options::sweep_order, "QN_SWEEP_ORDER", 0},
  {"sweep_depth", &qn_options::sweep_depth, "QN_SWEEP_DEPTH", 0},
  {"hint_p1_ld", &qn_options::hint_p1_ld, "QN_HINT_P1_LD", 0},
  {"hint_p2_ld", &qn_options::hint_p2_ld, "QN_HINT_P2_LD", 0},
  {"hint_p2_st", &qn_options::hint_p2_st, "QN_HINT_P2_ST", 0},
  {"hint_one_ld", &qn_options::hint_one_ld, "QN_HINT_ONE_LD", 0},
  {"p1_pad_kb", &qn_options::p1_pad_kb, "QN_P1_PAD_KB", 0},
  {"record_exit", &qn_options::record_exit, "QN_RECORD_EXIT", 0},
  {"lead_parity", &qn_options::lead_parity, "QN_LEAD_PARITY", 0},
  {"lead_warp", &qn_options::lead_warp, "QN_LEAD_WARP", 0},
  {"host_chunks", &qn_options::host_chunks, "QN_HOST_CHUNKS", 0},
};
void options_from_env(qn_options& o) {
  for (const OptKey& k : kOptKeys) {
    const char* v = getenv(k.env);
    if (!v) continue;
    if (k.env_mode == 1) o.*(k.field) = 0;
    else if (k.env_mode == 2) o.*(k.field) = (int)(unsigned char)v[0];
    else o.*(k.field) = atoi(v);
  }
}

int check_common(qn_context* ctx, const qn_model* model, int32_t n, int64_t batch,
                 const double* theta0, const qn_backtracking* ls, int32_t max_iter) {
  if (!ctx) return fail(QN_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (!model) return fail(QN_ERR_UNREGISTERED_OBJECTIVE, "model is NULL (unregistered objective)");
  if (model->ctx != ctx) return fail(QN_ERR_INVALID_ARGUMENT, "model belongs to another context");
  if (n < 1) return fail(QN_ERR_INVALID_ARGUMENT, "n must be >= 1 (got %d)", n);
  if (batch < 0) return fail(QN_ERR_INVALID_ARGUMENT, "batch must be >= 0");
  if (batch > 0 && !theta0) return fail(QN_ERR_INVALID_ARGUMENT, "theta0 is NULL");
  if (!ls) return fail(QN_ERR_INVALID_ARGUMENT, "ls is NULL");
  if (ls->order != 2 && ls->order != 3)
    return fail(QN_ERR_INVALID_ARGUMENT, "BackTracking order must be 2 or 3 (got %d)", ls->order);
  if (max_iter < 0) return fail(QN_ERR_INVALID_ARGUMENT, "max_iter must be >= 0");
  return QN_OK;
}

}  // namespace

// --------------------------------------------------------------- public API
extern "C" {

int32_t qn_version(void) { return QN_VERSION; }
const char* qn_last_error(void) { return g_err; }

void qn_backtracking_default(qn_backtracking* ls, int32_t order) {
  if (!ls) return;
  ls->c1 = 1e-4;
  ls->rho_hi = 0.5;
  ls->rho_lo = 0.1;
  ls->iterations = 1000;
  ls->order = order;
  ls->_pad = 0;
}

int32_t qn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int32_t qn_create(qn_context** out, int32_t device) {
  if (!out) return fail(QN_ERR_INVALID_ARGUMENT, "ctx out-pointer is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(QN_ERR_NO_DEVICE,
                "no usable CUDA device (%s); libqn_cuda has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev)
    return fail(QN_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, ndev);
  QN_CUDA_TRY(cudaSetDevice(device));
  qn_context* ctx = new (std::nothrow) qn_context;
  if (!ctx) return fail(QN_ERR_CUDA, "out of host memory");
  ctx->device = device;
  options_from_env(ctx->opt); // the only place the environment is read
  cudaError_t err = cudaGetDeviceProperties(&ctx->prop, device);
  for (int i = 0; i < qn_context::kRangeStreams && err == cudaSuccess; ++i)
    err = cudaStreamCreateWithFlags(&ctx->range_stream[i], cudaStreamNonBlocking);
  ctx->stream = ctx->range_stream[0];
  if (err == cudaSuccess)
    err = cudaMalloc(&ctx->counter, sizeof(unsigned long long) * qn_context::kCounterSlots);
  if (err == cudaSuccess)
    err = cudaMalloc(&ctx->exit_ns, sizeof(unsigned long long) * (qn_context::kExitSlots + 2));
  if (err != cudaSuccess) {
    delete ctx;
    return fail(QN_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(err));
  }
  if (ctx->prop.major < 10) {
    // The kernels are compiled for sm_100a only; refuse early with a clear message.
    int major = ctx->prop.major, minor = ctx->prop.minor;
    cudaFree(ctx->counter);
    cudaFree(ctx->exit_ns);
    for (cudaStream_t t : ctx->range_stream) cudaStreamDestroy(t);
    delete ctx;
    return fail(QN_ERR_NO_DEVICE, "device is sm_%d%d; libqn_cuda is built for sm_100a only", major,
                minor);
  }
  *out = ctx;
  return QN_OK;
}

int32_t qn_destroy(qn_context* ctx) {
  if (!ctx) return QN_OK;
  cudaSetDevice(ctx->device);
  for (cudaStream_t t : ctx->range_stream) cudaStreamSynchronize(t);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->hstage) cudaFreeHost(ctx->hstage);
  cudaFree(ctx->counter);
  cudaFree(ctx->exit_ns);
  for (cudaStream_t t : ctx->range_stream) cudaStreamDestroy(t);
  delete ctx;
  return QN_OK;
}

int32_t qn_set_option(qn_context* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return fail(QN_ERR_INVALID_ARGUMENT, "NULL argument");
  for (const OptKey& k : kOptKeys)
    if (!strcmp(k.key, key)) {
      std::lock_guard<std::mutex> lock(ctx->mu);
      ctx->opt.*(k.field) = (int)value;
      return QN_OK;
    }
  return fail(QN_ERR_INVALID_ARGUMENT, "unknown option '%s'", key);
}

int32_t qn_get_option(qn_context* ctx, const char* key, int64_t* value) {
  if (!ctx || !key || !value) return fail(QN_ERR_INVALID_ARGUMENT, "NULL argument");
  // read-only facts of the device / the last solve
  if (!strcmp(key, "device_l2_mb")) { *value = ctx->prop.l2CacheSize >> 20; return QN_OK; }
  if (!strcmp(key, "device_l2_persist_max_mb")) { *value = ctx->prop.persistingL2CacheMaxSize >> 20; return QN_OK; }
  if (!strcmp(key, "device_access_window_max_mb")) { *value = ctx->prop.accessPolicyMaxWindowSize >> 20; return QN_OK; }
  if (!strcmp(key, "last_l2_persist_mb")) { *value = ctx->last_l2_persist_bytes >> 20; return QN_OK; }
  for (const OptKey& k : kOptKeys)
    if (!strcmp(k.key, key)) {
      *value = ctx->opt.*(k.field);
      return QN_OK;
    }
  return fail(QN_ERR_INVALID_ARGUMENT, "unknown option '%s'", key);
}

int32_t qn_device_info(qn_context* ctx, char* name, size_t name_len, int32_t* sm_count,
                       int32_t* smem_per_block, int32_t* clock_khz) {
  if (!ctx) return fail(QN_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (name && name_len) {
    strncpy(name, ctx->prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
  if (smem_per_block) *smem_per_block = (int32_t)ctx->prop.sharedMemPerBlockOptin;
  if (clock_khz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    *clock_khz = khz;
  }
  return QN_OK;
}

int32_t qn_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(QN_ERR_INVALID_ARGUMENT, "ptr is NULL");
  QN_CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 1));
  return QN_OK;
}
int32_t qn_host_free(void* ptr) {
  if (ptr) QN_CUDA_TRY(cudaFreeHost(ptr));
  return QN_OK;
}

int32_t qn_state_ld(int32_t n, int32_t native_w) { return qn_host_state_ld(n, native_w); }
int64_t qn_state_len(int32_t n, int32_t ld) { return 7 * (int64_t)ld + (int64_t)n * ld; }

// ---- models
int32_t qn_model_create(qn_context* ctx, int32_t kind, qn_model** out) {
  if (!ctx || !out) return fail(QN_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  if (kind != QN_OBJ_ROSENBROCK)
    return fail(QN_ERR_UNREGISTERED_OBJECTIVE,
                "objective kind %d has no registered parameter-free device implementation "
                "(no CPU fallback; see INTEGRATION.md)", kind);
  qn_model* m = new qn_model;
  m->ctx = ctx;
  m->kind = kind;
  *out = m;
  return QN_OK;
}

int32_t qn_model_dense_gaussian(qn_context* ctx, int32_t n, const double* mu, const double* lambda,
                                qn_model** out) {
  if (!ctx || !out || !mu || !lambda || n < 1) return fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  *out = nullptr;
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  qn_model* m = new qn_model;
  m->ctx = ctx;
  m->kind = QN_OBJ_DENSE_GAUSSIAN;
  m->n = n;
  cudaError_t e = cudaMalloc(&m->d_a, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_b, sizeof(double) * (size_t)n * n);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_a, mu, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(m->d_b, lambda, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    qn_model_free(m);
    return fail(QN_ERR_CUDA, "uploading Gaussian model failed: %s", cudaGetErrorString(e));
  }
  *out = m;
  return QN_OK;
}

int32_t qn_model_logistic(qn_context* ctx, int32_t mrows, int32_t n, const double* X,
                          const double* y, double lam, qn_model** out) {
  if (!ctx || !out || !X || !y || n < 1 || mrows < 1)
    return fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  *out = nullptr;
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  qn_model* m = new qn_model;
  m->ctx = ctx;
  m->kind = QN_OBJ_LOGISTIC;
  m->n = n;
  m->m = mrows;
  m->lam = lam;
  const size_t nx = (size_t)mrows * n;
  std::vector<double> xt(nx);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < mrows; ++i) xt[(size_t)n * i + j] = X[(size_t)mrows * j + i];
  cudaError_t e = cudaMalloc(&m->d_a, sizeof(double) * nx);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_b, sizeof(double) * nx);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_c, sizeof(double) * mrows);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_a, X, sizeof(double) * nx, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_b, xt.data(), sizeof(double) * nx, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_c, y, sizeof(double) * mrows, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    qn_model_free(m);
    return fail(QN_ERR_CUDA, "uploading logistic model failed: %s", cudaGetErrorString(e));
  }
  *out = m;
  return QN_OK;
}

int32_t qn_model_user(qn_context* ctx, int32_t n, const void* params, size_t bytes, qn_model** out) {
  if (!ctx || !out || n < 0 || (bytes > 0 && !params)) return fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  *out = nullptr;
#ifndef QN_USER_OBJECTIVE_HEADER
  return fail(QN_ERR_UNREGISTERED_OBJECTIVE,
              "this libqn_cuda.so was built without a user objective (include/qn_objective.cuh: "
              "QN_USER_OBJECTIVE=my.cuh python build.py); there is no CPU fallback");
#else
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  qn_model* m = new qn_model;
  m->ctx = ctx;
  m->kind = QN_OBJ_USER;
  m->n = n;
  m->bytes = bytes;
  if (bytes > 0) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&m->d_a), bytes);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_a, params, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      qn_model_free(m);
      return fail(QN_ERR_CUDA, "uploading the user objective's parameters failed: %s", cudaGetErrorString(e));
    }
  }
  *out = m;
  return QN_OK;
#endif
}

int32_t qn_model_free(qn_model* m) {
  if (!m) return QN_OK;
  if (m->ctx) cudaSetDevice(m->ctx->device);
  if (m->d_a) cudaFree(m->d_a);
  if (m->d_b) cudaFree(m->d_b);
  if (m->d_c) cudaFree(m->d_c);
  delete m;
  return QN_OK;
}

}  // extern "C"

// ---- optimize!, device pointers; the caller holds ctx->mu
int qn_optimize_device_locked(qn_context* ctx, const qn_model* model, int32_t n, int64_t batch,
                              const double* theta0, int64_t theta_stride, const qn_backtracking* ls,
                              double tol, int32_t max_iter, double* states, int64_t state_stride,
                              int32_t ld, double* x_out, double* g_out, double* ell_out,
                              int32_t* iters_out, int32_t* fevals_out, int32_t* status_out,
                              int32_t ntrace, double* trace_out, uint32_t flags, cudaStream_t st) {
  int rc = check_common(ctx, model, n, batch, theta0, ls, max_iter);
  if (rc) return rc;
  if (theta_stride < n) return fail(QN_ERR_INVALID_ARGUMENT, "theta_stride < n");
  if (states && (ld < n || state_stride < qn_state_len(n, ld)))
    return fail(QN_ERR_INVALID_ARGUMENT, "state block too small: ld=%d stride=%lld for n=%d", ld,
                (long long)state_stride, n);
  if (states && (x_out || g_out))
    return fail(QN_ERR_INVALID_ARGUMENT,
                "device entry: pass either states or compact x_out/g_out, not both");
  if (ntrace < 0 || (ntrace > 0 && !trace_out))
    return fail(QN_ERR_INVALID_ARGUMENT, "ntrace > 0 needs trace_out");
  if (batch == 0) return QN_OK;
  QN_CUDA_TRY(cudaSetDevice(ctx->device));

  qn::BatchParams p{};
  p.n = n;
  p.max_iter = max_iter;
  p.batch = batch;
  p.theta0 = theta0;
  p.theta_stride = theta_stride;
  p.tol = tol;
  p.ls = qn::LsParams{ls->c1, ls->rho_hi, ls->rho_lo, (long long)ls->iterations, ls->order};
  if (states) {
    p.x_out = states;
    p.x_stride = state_stride;
    p.g_out = states + ld;
    p.g_stride = state_stride;
    if (flags & QN_FLAG_WRITE_H) {
      p.h_out = states + 7 * (int64_t)ld;
      p.h_stride = state_stride;
      p.ldh = ld;
    }
  } else {
    p.x_out = x_out;
    p.x_stride = n;
    p.g_out = g_out;
    p.g_stride = n;
  }
  p.ell_out = ell_out;
  p.iters_out = iters_out;
  p.fevals_out = fevals_out;
  p.status_out = status_out;
  p.ntrace = ntrace;
  p.trace = trace_out;
  return dispatch_model(ctx, model, p, ls, flags, st);
}

extern "C" {

int32_t qn_bfgs_optimize_batched_device_f64(qn_context* ctx, const qn_model* model, int32_t n,
                                            int64_t batch, const double* theta0,
                                            int64_t theta_stride, const qn_backtracking* ls,
                                            double tol, int32_t max_iter, double* states,
                                            int64_t state_stride, int32_t ld, double* x_out,
                                            double* g_out, double* ell_out, int32_t* iters_out,
                                            int32_t* fevals_out, int32_t* status_out,
                                            int32_t ntrace, double* trace_out, uint32_t flags,
                                            void* stream) {
  if (!ctx) return fail(QN_ERR_INVALID_ARGUMENT, "ctx is NULL");
  std::lock_guard<std::mutex> lock(ctx->mu);
  // `stream` is used literally: NULL is the CUDA default stream
  return qn_optimize_device_locked(ctx, model, n, batch, theta0, theta_stride, ls, tol, max_iter, states,
                                   state_stride, ld, x_out, g_out, ell_out, iters_out, fevals_out,
                                   status_out, ntrace, trace_out, flags, (cudaStream_t)stream);
}

// ---- optimize!, host pointers (the call a Julia / Python user makes)
int32_t qn_bfgs_optimize_batched_f64(qn_context* ctx, const qn_model* model, int32_t n,
                                     int64_t batch, const double* theta0,
                                     const qn_backtracking* ls, double tol, int32_t max_iter,
                                     double* states, int64_t state_stride, int32_t ld,
                                     double* x_out, double* g_out, double* ell_out,
                                     int32_t* iters_out, int32_t* fevals_out,
                                     int32_t* status_out, int32_t ntrace, double* trace_out,
                                     uint32_t flags) {
  int rc = check_common(ctx, model, n, batch, theta0, ls, max_iter);
  if (rc) return rc;
  if (states && (ld < n || state_stride < qn_state_len(n, ld)))
    return fail(QN_ERR_INVALID_ARGUMENT, "state block too small: ld=%d stride=%lld for n=%d", ld,
                (long long)state_stride, n);
  if (ntrace < 0 || (ntrace > 0 && !trace_out))
    return fail(QN_ERR_INVALID_ARGUMENT, "ntrace > 0 needs trace_out");
  if (batch == 0) return QN_OK;
  QN_CUDA_TRY(cudaSetDevice(ctx->device));

  const bool want_h = states && (flags & QN_FLAG_WRITE_H);
  const size_t nb = (size_t)n * batch;
  // device staging: theta | x | g | ell | iters | fevals | status | trace | H
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_theta = carve(sizeof(double) * nb);
  const size_t o_x = carve(sizeof(double) * nb);
  const size_t o_g = carve(sizeof(double) * nb);
  const size_t o_ell = carve(sizeof(double) * batch);
  const size_t o_it = carve(sizeof(int32_t) * batch);
  const size_t o_fe = carve(sizeof(int32_t) * batch);
  const size_t o_st = carve(sizeof(int32_t) * batch);
  const size_t o_tr = carve(sizeof(double) * nb * (size_t)ntrace);
  // B⁻¹ is staged in the layout of the state blocks (column stride ld), so ONE strided copy returns it
  const size_t hblk = (size_t)n * ld;
  const size_t o_h = carve(want_h ? sizeof(double) * hblk * batch : 0);
  // One lock from staging to the final synchronise: the staging buffer and the streams belong to the
  // context, so two host calls on one context run one after the other.
  std::lock_guard<std::mutex> lock(ctx->mu);
  rc = qn_ensure_stage(ctx, off);
  if (rc) return rc;
  char* base = (char*)ctx->stage;

  // A large batch is cut into column ranges, each on a stream of its own: the host->device copy of
  // range k+1 and the device->host copy of range k-1 run under the kernel of range k, and the CTAs
  // of kernel k+1 take the SM slots kernel k frees while its last problems finish.  Every problem
  // is solved by the same kernel as in one launch (same bits).  All launches are enqueued before
  // the first copy back, so a pageable destination (whose copy blocks the host) costs no overlap.
  const int chunks = qn_host_ranges(ctx, batch, ntrace);
  std::vector<qn::BatchParams> ps((size_t)chunks);
  int launches = 0;
  auto sync_all = [&] {
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < chunks; ++c) {
      cudaError_t ec = cudaStreamSynchronize(ctx->range_stream[c]);
      if (e == cudaSuccess) e = ec;
    }
    return e;
  };
  // everything is enqueued inside one lambda so that an error on any range still waits for the
  // ranges already in flight (they write into the caller's buffers) before the call returns
  auto enqueue = [&]() -> int {
    for (int c = 0; c < chunks; ++c) {
      const int64_t c0 = batch * c / chunks, cnt = batch * (c + 1) / chunks - c0;
      const size_t e0 = (size_t)n * c0, ne = (size_t)n * cnt;
      cudaStream_t st = ctx->range_stream[c];
      QN_CUDA_TRY(cudaMemcpyAsync(base + o_theta + sizeof(double) * e0, theta0 + e0, sizeof(double) * ne,
                                  cudaMemcpyHostToDevice, st));
      if (ntrace > 0) {
        // rows past a problem's last iteration stay as the caller initialised them
        QN_CUDA_TRY(cudaMemcpyAsync(base + o_tr, trace_out, sizeof(double) * nb * ntrace,
                                    cudaMemcpyHostToDevice, st));
      }
      qn::BatchParams& p = ps[c];
      p.n = n;
      p.max_iter = max_iter;
      p.batch = cnt;
      p.theta0 = (const double*)(base + o_theta) + e0;
      p.theta_stride = n;
      p.tol = tol;
      p.ls = qn::LsParams{ls->c1, ls->rho_hi, ls->rho_lo, (long long)ls->iterations, ls->order};
      p.x_out = (double*)(base + o_x) + e0;
      p.x_stride = n;
      p.g_out = (double*)(base + o_g) + e0;
      p.g_stride = n;
      if (want_h) {
        p.h_out = (double*)(base + o_h) + hblk * (size_t)c0;
        p.h_stride = (long long)hblk;
        p.ldh = ld;
        if (ld > n) // the padding rows are never written by the kernels: give the copy defined bytes
          QN_CUDA_TRY(cudaMemsetAsync(p.h_out, 0, sizeof(double) * hblk * cnt, st));
      }
      p.ell_out = (double*)(base + o_ell) + c0;
      p.iters_out = (int*)(base + o_it) + c0;
      p.fevals_out = (int*)(base + o_fe) + c0;
      p.status_out = (int*)(base + o_st) + c0;
      p.ntrace = ntrace;
      p.trace = ntrace ? (double*)(base + o_tr) : nullptr;
      const int drc = dispatch_model(ctx, model, p, ls, flags, st);
      if (drc) return drc;
      launches += ctx->last_launches;
    }
    ctx->last_launches = launches;
    // results back to the caller's buffers, range by range
    for (int c = 0; c < chunks; ++c) {
      const qn::BatchParams& p = ps[c];
      const int64_t c0 = batch * c / chunks, cnt = p.batch;
      const size_t e0 = (size_t)n * c0, ne = (size_t)n * cnt;
      cudaStream_t st = ctx->range_stream[c];
      if (states) {
        double* sblk = states + state_stride * c0;
        // θ* -> offset 0, ∇ -> offset ld of each block (QN.jl:135,138): strided 2-D copies
        QN_CUDA_TRY(cudaMemcpy2DAsync(sblk, sizeof(double) * state_stride, p.x_out, sizeof(double) * n,
                                      sizeof(double) * n, cnt, cudaMemcpyDeviceToHost, st));
        QN_CUDA_TRY(cudaMemcpy2DAsync(sblk + ld, sizeof(double) * state_stride, p.g_out,
                                      sizeof(double) * n, sizeof(double) * n, cnt,
                                      cudaMemcpyDeviceToHost, st));
        if (want_h) // field 8 of every block (QN.jl:142), n·ld contiguous doubles per problem: one strided copy
          QN_CUDA_TRY(cudaMemcpy2DAsync(sblk + 7 * (int64_t)ld, sizeof(double) * state_stride, p.h_out,
                                        sizeof(double) * hblk, sizeof(double) * hblk, cnt,
                                        cudaMemcpyDeviceToHost, st));
      }
      if (x_out) QN_CUDA_TRY(cudaMemcpyAsync(x_out + e0, p.x_out, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
      if (g_out) QN_CUDA_TRY(cudaMemcpyAsync(g_out + e0, p.g_out, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
      if (ell_out) QN_CUDA_TRY(cudaMemcpyAsync(ell_out + c0, p.ell_out, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
      if (iters_out) QN_CUDA_TRY(cudaMemcpyAsync(iters_out + c0, p.iters_out, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, st));
      if (fevals_out) QN_CUDA_TRY(cudaMemcpyAsync(fevals_out + c0, p.fevals_out, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, st));
      if (status_out) QN_CUDA_TRY(cudaMemcpyAsync(status_out + c0, p.status_out, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, st));
      if (ntrace > 0)
        QN_CUDA_TRY(cudaMemcpyAsync(trace_out, base + o_tr, sizeof(double) * nb * ntrace, cudaMemcpyDeviceToHost, st));
    }
    return QN_OK;
  };
  rc = enqueue();
  {
    const cudaError_t e = sync_all();
    if (rc) return rc;
    QN_CUDA_TRY(e);
  }
  return QN_OK;
}

int32_t qn_bfgs_optimize_f64(qn_context* ctx, const qn_model* model, int32_t n, const double* x,
                             const qn_backtracking* ls, double tol, double* state, int32_t ld,
                             double* ell, int32_t* iters, int32_t* fevals, int32_t* status,
                             uint32_t flags) {
  if (!state) return fail(QN_ERR_INVALID_ARGUMENT, "state is NULL");
  return qn_bfgs_optimize_batched_f64(ctx, model, n, 1, x, ls, tol, 10000, state,
                                      qn_state_len(n, ld), ld, nullptr, nullptr, ell, iters, fevals,
                                      status, 0, nullptr, flags);
}

// Tail of the last tile-kernel launch made with option "record_exit" = 1 (after the caller has
// synchronised): span = latest end - earliest start; idle_fraction = Σ_slots (end - exit_slot) /
// (slots · span), the share of the launch's slot-time spent with the work queue empty.
int32_t qn_last_tail_profile(qn_context* ctx, double* span_ms, double* idle_fraction, int32_t* slots) {
  if (!ctx) return fail(QN_ERR_INVALID_ARGUMENT, "ctx is NULL");
  std::lock_guard<std::mutex> lock(ctx->mu);
  const int used = ctx->exit_slots_used;
  if (span_ms) *span_ms = 0.0;
  if (idle_fraction) *idle_fraction = 0.0;
  if (slots) *slots = used;
  if (used <= 0) return QN_OK;
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<unsigned long long> h(used + 2);
  QN_CUDA_TRY(cudaMemcpy(h.data(), ctx->exit_ns, sizeof(unsigned long long) * (used + 2), cudaMemcpyDeviceToHost));
  const unsigned long long t0 = h[0], t1 = h[1];
  if (t1 <= t0) return QN_OK;
  double idle = 0.0;
  int counted = 0;
  for (int k = 0; k < used; ++k)
    if (h[2 + k] >= t0 && h[2 + k] <= t1) { idle += (double)(t1 - h[2 + k]); ++counted; }
  if (span_ms) *span_ms = (double)(t1 - t0) * 1e-6;
  if (idle_fraction && counted) *idle_fraction = idle / ((double)counted * (double)(t1 - t0));
  if (slots) *slots = counted;
  return QN_OK;
}

int32_t qn_last_launch_info(qn_context* ctx, int32_t* path, int32_t* grid, int32_t* block,
                            int32_t* smem_bytes, int32_t* kernel_launches) {
  if (!ctx) return fail(QN_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (path) *path = ctx->last_path;
  if (grid) *grid = ctx->last_grid;
  if (block) *block = ctx->last_block;
  if (smem_bytes) *smem_bytes = ctx->last_smem;
  if (kernel_launches) *kernel_launches = ctx->last_launches;
  return QN_OK;
}

}  // extern "C"
