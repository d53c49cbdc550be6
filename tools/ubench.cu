// ubench.cu — one-off microbenchmarks that answer layout questions for the resident kernel
// (DESIGN.md §5): cost in SM cycles of broadcast / half-warp-broadcast shared-memory loads of
// 32/64/128 bits, DFMA dependent-issue latency and throughput, SHFL throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// MODE: 0 all lanes same address; 1 one address per half-warp; 2 one address per quarter-warp;
//       3 conflict-free distinct (lane*W)
template <int W, int MODE>
__global__ void lds_pattern(long long* out, int iters, long long* clk) {
  __shared__ __align__(16) unsigned char buf[16384];
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) ((int*)buf)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int sel = (MODE == 0) ? 0 : (MODE == 1) ? (lane >> 4) : (MODE == 2) ? (lane >> 3) : lane;
  unsigned off = sel * ((MODE == 3) ? W : 64);
  long long bits = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const unsigned a = (off + (unsigned)(k * 528 + i * 16)) & (16384 - 16);
      if (W == 4) { bits ^= *(const int*)(buf + a); }
      else if (W == 8) { bits ^= *(const long long*)(buf + a); }
      else { const longlong2 v = *(const longlong2*)(buf + a); bits ^= v.x ^ v.y; }
    }
  }
  long long t1 = clock64();
  if (bits == 0x1234567) out[0] = bits;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int CHAINS>
__global__ void dfma_chain(double* out, int iters, double a, double b, long long* clk) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

__global__ void shfl_tp(double* out, int iters, long long* clk) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x += __shfl_xor_sync(0xffffffffu, x, (k & 15) + 1);
  }
  long long t1 = clock64();
  if (x == 123.456) out[0] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <class F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  long long *out, *clk; double* dout;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&clk, 64)); CK(cudaMalloc(&dout, 64));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, khz);
  const int iters = 2048;
  // full-chip: 8 CTAs x 256 threads per SM -> cycles per warp-load per SM
  const int grid = p.multiProcessorCount * 8, block = 256;
  const double warp_loads_per_sm = 8.0 * (block / 32) * iters * 16.0;
#define RUN_LDS(W, MODE, NAME) { long long h; float ms = timeit([&] { lds_pattern<W, MODE><<<grid, block>>>(out, iters, clk); }); \
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); \
    printf("LDS.%-3d %-28s : %.3f SM-cycles per warp-load (event time %.3f ms)\n", W * 8, NAME, ms * 1e-3 * khz * 1e3 / warp_loads_per_sm, ms); }
  RUN_LDS(4, 0, "broadcast (1 addr/warp)");
  RUN_LDS(8, 0, "broadcast (1 addr/warp)");
  RUN_LDS(16, 0, "broadcast (1 addr/warp)");
  RUN_LDS(8, 1, "1 addr per half-warp");
  RUN_LDS(16, 1, "1 addr per half-warp");
  RUN_LDS(8, 2, "1 addr per quarter-warp");
  RUN_LDS(16, 2, "1 addr per quarter-warp");
  RUN_LDS(4, 3, "distinct, conflict-free");
  RUN_LDS(8, 3, "distinct, conflict-free");
  RUN_LDS(16, 3, "distinct, conflict-free");
  // DFMA: latency (1 warp, 1 chain) and throughput (full chip, 8 chains)
  { long long h; dfma_chain<1><<<1, 32>>>(dout, iters, 0.999, 1e-9, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent-issue latency: %.2f cycles\n", (double)h / (iters * 16.0)); }
  { long long h; dfma_chain<2><<<1, 32>>>(dout, iters, 0.999, 1e-9, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 1 warp, 2 chains: %.2f cycles per DFMA\n", (double)h / (iters * 32.0)); }
  { long long h; dfma_chain<4><<<1, 32>>>(dout, iters, 0.999, 1e-9, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 1 warp, 4 chains: %.2f cycles per DFMA\n", (double)h / (iters * 64.0)); }
  { long long h; dfma_chain<8><<<1, 32>>>(dout, iters, 0.999, 1e-9, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 1 warp, 8 chains: %.2f cycles per DFMA\n", (double)h / (iters * 128.0)); }
  { float ms = timeit([&] { dfma_chain<8><<<grid, block>>>(dout, iters, 0.999, 1e-9, clk); });
    double fmas = (double)grid * block * iters * 16.0 * 8.0;
    printf("DFMA full chip: %.2f TFLOP/s = %.2f DFMA/clk/SM\n", 2 * fmas / (ms * 1e-3) / 1e12, fmas / (ms * 1e-3) / (khz * 1e3) / p.multiProcessorCount); }
  { long long h; shfl_tp<<<1, 32>>>(dout, iters, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("SHFL.64+DADD dependent latency: %.2f cycles per step\n", (double)h / (iters * 16.0)); }
  { float ms = timeit([&] { shfl_tp<<<grid, block>>>(dout, iters, clk); });
    printf("SHFL.64 full chip: %.3f SM-cycles per 64-bit warp-shuffle\n", ms * 1e-3 * khz * 1e3 / (8.0 * (block / 32) * iters * 16.0)); }
  return 0;
}
