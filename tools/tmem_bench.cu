// tmem_bench.cu — tensor-memory load/store bandwidth and latency for plain data (no MMA): would
// TMEM do as a second home for B⁻¹ tiles next to the register file?  
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tmem_bench tools/tmem_bench.cu
//   ./tmem_bench
// Each CTA (4 warps = all 128 lanes) allocates 128 columns and moves 32 columns x 32 lanes x 4 B =
// 4 KB per warp instruction (tcgen05.ld/st .32x32b.x32).  Reports bytes per SM-cycle for 1, 2 and 4
// resident CTAs per SM, for loads, stores, and a load + store round trip, and the latency of one
// dependent load.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define W32(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
               "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), \
               "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), \
               "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])

__device__ __forceinline__ void tmem_ld32(uint32_t (&v)[32], uint32_t addr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : R32(v) : "r"(addr));
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      :: "r"(addr), W32(v));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// mode 0: loads (4 in flight between waits), 1: stores, 2: load -> wait -> store -> wait, 3: one dependent load at a time
__global__ void __launch_bounds__(128) tmem_kernel(const int mode, const int reps, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" :: "l"((unsigned long long)__cvta_generic_to_shared(&base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 32 + i;
#pragma unroll
  for (int c = 0; c < 128; c += 32) tmem_st32(base + c, v);
  wait_st();
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int r = 0; r < reps; ++r) {
      uint32_t a[32], b[32];
      tmem_ld32(a, base + 0); tmem_ld32(b, base + 32);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= a[i] + b[i];
      tmem_ld32(a, base + 64); tmem_ld32(b, base + 96);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= a[i] ^ b[i];
    }
  } else if (mode == 1) {
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int c = 0; c < 128; c += 32) { v[0] += r; tmem_st32(base + c, v); }
      wait_st();
    }
  } else if (mode == 2) {
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        uint32_t a[32];
        tmem_ld32(a, base + c);
        wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] += 1;
        tmem_st32(base + c, a);
      }
      wait_st();
    }
  } else {
    uint32_t off = 0;
    for (int r = 0; r < 4 * reps; ++r) {
      uint32_t a[32];
      tmem_ld32(a, base + (off & 96));
      wait_ld();
      off = a[0] & 0; // dependent address (always 0 extra), keeps the loads serial
      acc ^= a[1];
      off += 32 * (r & 3);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (mode == 2 || mode == 1) { uint32_t a[32]; tmem_ld32(a, base); wait_ld(); acc ^= a[0]; }
  sink[blockIdx.x * 128 + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" :: "r"(base_s));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 8 * sms * 4); cudaMalloc(&d_sink, 4 * sms * 4 * 128);
  const int reps = 2000;
  const char* names[4] = {"load (2 in flight)", "store", "load+store round trip", "dependent load latency"};
  for (int mode = 0; mode < 4; ++mode)
    for (int per = 1; per <= 4; per *= 2) {
      const int grid = sms * per;
      tmem_kernel<<<grid, 128>>>(mode, reps, d_cyc, d_sink); // warm-up
      tmem_kernel<<<grid, 128>>>(mode, reps, d_cyc, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long c[8 * 148]; cudaMemcpy(c, d_cyc, 8 * grid, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < grid; ++i) avg += c[i]; avg /= grid;
      const double instr = 4.0 * reps;            // per warp
      const double bytes = instr * 4 * 4096.0 * per * (mode == 2 ? 2 : 1); // per SM (4 warps per CTA)
      if (mode < 3) printf("%-26s %d CTA/SM (%2d warps): %8.0f cycles, %7.1f B/clk/SM, %6.1f cycles per warp instruction\n",
                           names[mode], per, 4 * per, avg, bytes / avg, avg / instr);
      else printf("%-26s %d CTA/SM: %6.1f cycles per dependent tcgen05.ld.x32 + wait\n", names[mode], per, avg / instr);
    }
  return 0;
}
