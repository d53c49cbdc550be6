// qn_launch_small.cu — launchers of the other Rosenbrock-only kernels: several problems per warp
// (n = 8, 16, 32; qn_resident_packed.cuh) and the leader/helper tensor-memory tiles of n = 96, 128, 160
// (qn_resident_tile128.cuh).
#include <cuda_runtime.h>

#include "qn_resident_tile128.cuh"
#include "qn_resident_packed.cuh"
#include "qn_launch.h"

#define fail qn_fail

namespace {

// n = 32·RT Rosenbrock (n = 128: RT = 4, n = 160: RT = 5, n = 96: RT = 3): the leader/helper tile kernel
// with the whole tile in tensor memory (qn_resident_tile128.cuh): one CTA = one problem, a warp per
// rule-C2 column block; two problems per SM at n <= 128, one at n = 160 (480 of the 512 columns).
template <int RT, class WObj = qn::WRosenbrock>
int launch_tilew(qn_context* ctx, qn::BatchParams& p, cudaStream_t st, const typename WObj::Params op = {}) {
  using C = qn::TileW<RT>;
  const size_t smem = sizeof(double) * (size_t)C::kSmemDoubles;
  auto kern = qn::qn_resident_tilew_kernel<RT, WObj>;
  // A cooperative objective reads its parameters (Λ: 8 n² bytes, 128 KB at n = 128) through L1 in every
  // evaluation: leave the SM's 256 KB to the cache except for what the resident CTAs need.
  const int ctas = tmem_residency(ctx, kern, C::THREADS, smem, C::kAlloc);
  const int carve = WObj::kCoop ? (int)std::min<size_t>(100, ((smem + 2048) * ctas * 100 + ctx->prop.sharedMemPerMultiprocessor - 1) / ctx->prop.sharedMemPerMultiprocessor + 5)
                                : (int)cudaSharedmemCarveoutMaxShared;
  QN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  long long grid = (long long)ctas * ctx->prop.multiProcessorCount;
  if (grid > p.batch) grid = p.batch;
  p.counter = qn_next_counter(ctx, st);
  p.lead_div = lead_divisor(ctx);
  p.lead_warp = (ctx->opt.lead_warp >= 0 && ctx->opt.lead_warp < C::THREADS / 32) ? ctx->opt.lead_warp : 0;
  arm_exit_stamps(ctx, p, grid, st);
  kern<<<(unsigned)grid, C::THREADS, smem, st>>>(p, op);
  QN_CUDA_TRY(cudaGetLastError());
  ctx->last_path = 0;
  ctx->last_grid = (int)grid;
  ctx->last_block = C::THREADS;
  ctx->last_smem = (int)smem;
  ctx->last_launches = 1;
  return QN_OK;
}

// n = 8, 16, 32 (Rosenbrock): several problems per warp (qn_resident_packed.cuh)
template <int N, int RPT, int MINB>
int launch_packed(qn_context* ctx, qn::BatchParams& p, cudaStream_t st) {
  using C = qn::PackedCfg<N, RPT>;
  const size_t smem = sizeof(double) * (size_t)C::kSmemDoubles;
  auto kern = qn::qn_resident_packed_kernel<N, RPT, MINB>;
  const int block = C::kWarps * 32;
  int occ = 0;
  QN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
  if (occ < 1) return fail(QN_ERR_UNSUPPORTED_SIZE, "packed kernel does not fit on an SM (n=%d)", N);
  long long grid = (long long)occ * ctx->prop.multiProcessorCount;
  const long long need = (p.batch + C::kWarps * C::kSub - 1) / (C::kWarps * C::kSub);
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  p.counter = qn_next_counter(ctx, st);
  kern<<<(unsigned)grid, block, smem, st>>>(p);
  QN_CUDA_TRY(cudaGetLastError());
  ctx->last_path = 0;
  ctx->last_grid = (int)grid;
  ctx->last_block = block;
  ctx->last_smem = (int)smem;
  ctx->last_launches = 1;
  return QN_OK;
}

}  // namespace

int qn_launch_rosenbrock_small(qn_context* ctx, qn::BatchParams& p, cudaStream_t st, bool onepass) {
  const int n = p.n;
  if (ctx->opt.pack) { // A/B knob: the thread-per-row kernels stay available
    const int rpt = ctx->opt.pack_variant; // tuning knob: rows per lane
    // defaults measured on B200 (2^18 problems): n = 8: 2 rows per lane 8.2e9 it/s (1: 6.7e9);
    // n = 16: 2 rows 3.7e9 (1: 2.8e9, 4: 3.3e9); n = 32: 2 rows 1.14e9 (thread per row: 0.90e9)
    if (n == 8) return rpt == 1 ? launch_packed<8, 1, 8>(ctx, p, st) : rpt == 4 ? launch_packed<8, 4, 4>(ctx, p, st) : launch_packed<8, 2, 6>(ctx, p, st);
    if (n == 16) return rpt == 1 ? launch_packed<16, 1, 6>(ctx, p, st) : rpt == 4 ? launch_packed<16, 4, 2>(ctx, p, st) : launch_packed<16, 2, 4>(ctx, p, st);
    if (n == 32) return launch_packed<32, 2, 2>(ctx, p, st);
  }
  // 'L' (n = 96, 128, 160, the default): leader/helper warps, the whole tile in tensor memory
  const char variant = ctx->opt.pair_rows ? (char)ctx->opt.pair_rows : 'L';
  if (variant == 'L' && !onepass) {
    if (n == 128) return launch_tilew<4>(ctx, p, st);
    if (n == 160) return launch_tilew<5>(ctx, p, st);
    if (n == 96) return launch_tilew<3>(ctx, p, st);
  }
  return -1;
}

// dense Gaussian at n = 96, 128, 160 (BASELINE.json configs[2]: n = 128): the same leader/helper kernel
// with the CTA-cooperative objective (WGaussian)
int qn_launch_gaussian_tilew(qn_context* ctx, qn::BatchParams& p, const double* d_mu, const double* d_lambda, cudaStream_t st) {
  const qn::WGaussian::Params op{d_mu, d_lambda};
  if (ctx->opt.pair_rows != 0 && ctx->opt.pair_rows != 'L') return -1;
  if (p.n == 128) return launch_tilew<4, qn::WGaussian>(ctx, p, st, op);
  if (p.n == 160) return launch_tilew<5, qn::WGaussian>(ctx, p, st, op);
  if (p.n == 96) return launch_tilew<3, qn::WGaussian>(ctx, p, st, op);
  return -1;
}
