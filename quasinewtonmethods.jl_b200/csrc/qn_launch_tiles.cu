// qn_launch_tiles.cu — launchers of the 4×CT tile kernels (33 <= n <= 64, Rosenbrock): B⁻¹ tiles in the
// register file or in tensor memory (qn_resident_tile.cuh).  Its own translation unit: these 18
// instantiations are most of the library's build time.
#include <cuda_runtime.h>

#include "qn_resident_tile.cuh"
#include "qn_resident_tile3w.cuh"
#include "qn_launch.h"

#define fail qn_fail

namespace {

// The register-tile kernel evaluates the objective warp-locally; only Rosenbrock has
// that form (qn_resident_tile.cuh).
template <int N, int MINB, int PF>
int launch_tile(qn_context* ctx, qn::BatchParams& p, cudaStream_t st) {
  const size_t smem = sizeof(double) * (size_t)qn::tile_problem_smem_doubles<N, false>();
  auto kern = qn::qn_resident_tile_kernel<N, MINB, PF>;
  p.lead_div = lead_divisor(ctx);
  int occ = 0;
  QN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 64, smem));
  if (occ < 1) return fail(QN_ERR_UNSUPPORTED_SIZE, "tile kernel does not fit on an SM (n=%d)", N);
  long long grid = (long long)occ * ctx->prop.multiProcessorCount;
  if (grid > p.batch) grid = p.batch;
  p.counter = qn_next_counter(ctx, st);
  arm_exit_stamps(ctx, p, grid, st);
  kern<<<(unsigned)grid, 64, smem, st>>>(p);
  QN_CUDA_TRY(cudaGetLastError());
  ctx->last_path = 0;
  ctx->last_grid = (int)grid;
  ctx->last_block = 64;
  ctx->last_smem = (int)smem;
  ctx->last_launches = 1;
  return QN_OK;
}

// The same kernel with the tile in tensor memory (qn_resident_tile.cuh, TM): CTAs of 128 threads =
// two problems, 128 columns of tensor memory each, 4 CTAs = 8 problems per SM where the register
// file holds 6.  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for a kernel that calls
// tcgen05.alloc (it cannot know the column count); tmem_residency derives the 4 CTAs per SM that
// registers (128 × 128 × 4 = the file), shared memory and 4 × 128 columns allow — measured: the
// throughput grows 1.6e8 → 2.5e8 → 3.0e8 → 3.7e8 iterations/s from 1 to 4 CTAs per SM (n = 60).
template <int N, int PF>
int launch_tile_tmem(qn_context* ctx, qn::BatchParams& p, cudaStream_t st) {
  const size_t smem = 2 * sizeof(double) * (size_t)qn::tile_problem_smem_doubles<N, true>();
  auto kern = qn::qn_resident_tile_tmem_kernel<N, PF>;
  p.lead_div = lead_divisor(ctx);
  QN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  long long grid = (long long)tmem_residency(ctx, kern, 128, smem, 128) * ctx->prop.multiProcessorCount; // 4 per SM on B200
  if (grid > (p.batch + 1) / 2) grid = (p.batch + 1) / 2;
  p.counter = qn_next_counter(ctx, st);
  arm_exit_stamps(ctx, p, 2 * grid, st);
  kern<<<(unsigned)grid, 128, smem, st>>>(p);
  QN_CUDA_TRY(cudaGetLastError());
  ctx->last_path = 0;
  ctx->last_grid = (int)grid;
  ctx->last_block = 128;
  ctx->last_smem = (int)smem;
  ctx->last_launches = 1;
  return QN_OK;
}

// Three warps per problem, speculative pass 1 (qn_resident_tile3w.cuh): CTAs of 384 threads = four problems,
// 256 tensor-memory columns each, two CTAs = 8 problems per SM.
template <int N>
int launch_tile3w(qn_context* ctx, qn::BatchParams& p, cudaStream_t st) {
  const size_t smem = qn::k3wProblems * sizeof(double) * (size_t)qn::Tile3wCfg<N>::kDoubles;
  auto kern = qn::qn_resident_tile3w_kernel<N>;
  QN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  long long grid = (long long)tmem_residency(ctx, kern, qn::k3wThreads, smem, 256) * ctx->prop.multiProcessorCount;
  const long long need = (p.batch + qn::k3wProblems - 1) / qn::k3wProblems;
  if (grid > need) grid = need;
  p.counter = qn_next_counter(ctx, st);
  p.lead_div = lead_divisor(ctx);
  arm_exit_stamps(ctx, p, qn::k3wProblems * grid, st);
  kern<<<(unsigned)grid, qn::k3wThreads, smem, st>>>(p);
  QN_CUDA_TRY(cudaGetLastError());
  ctx->last_path = 0;
  ctx->last_grid = (int)grid;
  ctx->last_block = qn::k3wThreads;
  ctx->last_smem = (int)smem;
  ctx->last_launches = 1;
  return QN_OK;
}

// Where the tile lives: tensor memory (leader/helper pairs, 8 problems per SM) once the batch gives
// every SM its 8 problems — measured on B200 at 2^15 problems it is 19-27 % faster than the
// all-register kernel at every n of the tile range (profiles/r2_lead_parity_sizes.txt); smaller
// batches stay in the register file.  Option "tile_tmem" = 0 / 1 forces either.
inline bool tile_in_tensor_memory(const qn_context* ctx, const int n, const long long batch) {
  (void)n;
  if (ctx->opt.tile_tmem == 0) return false;
  if (ctx->opt.tile_tmem == 1) return true;
  return batch >= 8LL * ctx->prop.multiProcessorCount;
}

}  // namespace

int qn_launch_rosenbrock_tiles(qn_context* ctx, qn::BatchParams& p, cudaStream_t st) {
  switch (p.n) {
#define QN_REG_CASE(N, BLOCK, MINB)
#define QN_TILE_CASE(N, MINB)                                                                       \
  case N:                                                                                            \
    if (tile_in_tensor_memory(ctx, N, p.batch))                                                      \
      return ctx->opt.tile_3w != 0 ? launch_tile3w<N>(ctx, p, st) : launch_tile_tmem<N, 1>(ctx, p, st); \
    /* 6 CTAs per SM (168 registers); 4, 5 and 7 were measured slower */             \
    return launch_tile<N, MINB, 1>(ctx, p, st);
#include "qn_reg_sizes.inc"
#undef QN_REG_CASE
#undef QN_TILE_CASE
    default: return -1;
  }
}
