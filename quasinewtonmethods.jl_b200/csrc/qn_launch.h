// qn_launch.h — helpers shared by the translation units that launch the resident kernels.
#pragma once

#include <algorithm>

#include "qn_device.cuh"
#include "qn_internal.h"

namespace {

// Resident CTAs per SM of a kernel that allocates tensor memory.  cudaOccupancyMaxActiveBlocksPerMulti-
// processor answers 1 for such a kernel (it cannot know the column count), so the residency is derived
// here from the three on-chip limits: registers, shared memory and 512 tensor-memory columns.
template <class K>
int tmem_residency(const qn_context* ctx, K kern, int block, size_t dyn_smem, int columns) {
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); return 1; }
  const int warps = (block + 31) / 32;
  const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;   // allocation granularity: 256 registers per warp
  const int by_regs = regs_per_warp > 0 ? ctx->prop.regsPerMultiprocessor / (regs_per_warp * warps) : 32;
  const size_t per_cta = dyn_smem + fa.sharedSizeBytes + 1024;        // + the 1 KB the system reserves per CTA
  const int by_smem = (int)(ctx->prop.sharedMemPerMultiprocessor / per_cta);
  const int by_tmem = 512 / columns;
  const int by_threads = ctx->prop.maxThreadsPerMultiProcessor / block;
  return std::max(1, std::min(std::min(by_regs, by_smem), std::min(by_tmem, by_threads)));
}

// option "lead_parity" -> BatchParams::lead_div (which warp of a leader/helper pair leads)
inline int lead_divisor(const qn_context* ctx) {
  const int v = ctx->opt.lead_parity;
  return v > 0 ? v : (v < 0 ? ctx->prop.multiProcessorCount : 0);
}

// option "record_exit": arm the exit stamps for a launch with `slots` problem slots
void arm_exit_stamps(qn_context* ctx, qn::BatchParams& p, long long slots, cudaStream_t st) {
  p.exit_ns = nullptr;
  p.exit_cap = 0;
  ctx->exit_slots_used = 0;
  if (!ctx->opt.record_exit) return;
  const int used = (int)std::min<long long>(slots, qn_context::kExitSlots);
  cudaMemsetAsync(ctx->exit_ns, 0xff, sizeof(unsigned long long), st);                 // earliest start
  cudaMemsetAsync(ctx->exit_ns + 1, 0, sizeof(unsigned long long) * (used + 1), st);   // latest end, slots
  p.exit_ns = ctx->exit_ns;
  p.exit_cap = used;
  ctx->exit_slots_used = used;
}

}  // namespace

// The Rosenbrock-only kernels live in their own translation units (they are most of the build time):
// qn_launch_tiles.cu — the 4×CT tile kernels of n = 36..64 (register file / tensor memory);
// qn_launch_small.cu — several problems per warp (n = 8, 16, 32) and the leader/helper
// tensor-memory tiles of n = 96, 128, 160.  Each returns -1 when it has no kernel for the request.
int qn_launch_rosenbrock_tiles(qn_context* ctx, qn::BatchParams& p, cudaStream_t st);
int qn_launch_rosenbrock_small(qn_context* ctx, qn::BatchParams& p, cudaStream_t st, bool onepass);
int qn_launch_gaussian_tilew(qn_context* ctx, qn::BatchParams& p, const double* d_mu, const double* d_lambda, cudaStream_t st);
