// tile_bench.cu — the register-tile kernel (n = 60) alone on a full-size batch: time per launch and a
// hash of every output, so that source variants of qn_resident_tile.cuh can be compared for speed
// AND for identical bits in one gpurun call.  The header is found through -I (point it at the
// variant's directory):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 --expt-relaxed-constexpr \
//        -I quasinewtonmethods.jl_b200/csrc [-DTB_MINB=6 -DTB_PF=1 -DTB_TMEM -DQN_TM_KC=2] -o tile_bench tools/tile_bench.cu
//   ./tile_bench [problems = 65536] [timed launches = 3]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>
#include "qn_resident_tile.cuh"

#ifndef TB_MINB
#define TB_MINB 6
#endif
#ifndef TB_PF
#define TB_PF 1
#endif
#ifndef TB_N
#define TB_N 60
#endif

static unsigned long long fnv(const void* p, size_t bytes, unsigned long long h = 1469598103934665603ULL) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < bytes; ++i) { h ^= c[i]; h *= 1099511628211ULL; }
  return h;
}

// One thread on its own stream: SM cycles and nanoseconds while the benchmark kernel runs -> the
// SM clock the kernel actually ran at (nvidia-smi samples every 100 ms; this is exact).
__global__ void clock_probe(volatile int* stop, unsigned long long* out) {
  unsigned long long t0, t1; long long c0 = clock64(), c1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { c1 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); __nanosleep(2000); } while (!*stop);
  out[0] = (unsigned long long)(c1 - c0); out[1] = t1 - t0;
}

int main(int argc, char** argv) {
  constexpr int n = TB_N;
  const long long nprob = argc > 1 ? atoll(argv[1]) : 65536;
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  std::mt19937_64 rng(20261018);
  std::normal_distribution<double> nd;
  std::vector<double> th((size_t)n * nprob);
  for (auto& v : th) v = nd(rng);
  double *d_th, *d_x, *d_g, *d_ell; int *d_it, *d_fe, *d_st; unsigned long long* d_cnt;
  cudaMalloc(&d_th, th.size() * 8); cudaMalloc(&d_x, th.size() * 8); cudaMalloc(&d_g, th.size() * 8);
  cudaMalloc(&d_ell, nprob * 8); cudaMalloc(&d_it, nprob * 4); cudaMalloc(&d_fe, nprob * 4); cudaMalloc(&d_st, nprob * 4);
  cudaMalloc(&d_cnt, 8);
  cudaMemcpy(d_th, th.data(), th.size() * 8, cudaMemcpyHostToDevice);
  qn::BatchParams p{};
  p.n = n; p.max_iter = 10000; p.batch = nprob; p.theta0 = d_th; p.theta_stride = n; p.tol = 1e-8;
  p.ls = qn::LsParams{1e-4, 0.5, 0.1, 1000, 2};
  p.x_out = d_x; p.x_stride = n; p.g_out = d_g; p.g_stride = n; p.ell_out = d_ell;
  p.iters_out = d_it; p.fevals_out = d_fe; p.status_out = d_st; p.counter = d_cnt;
#ifdef TB_TMEM // the tensor-memory variant: two problems per CTA of 128 threads
  const size_t smem = 2 * sizeof(double) * qn::tile_problem_smem_doubles<n, true>();
  auto kern = qn::qn_resident_tile_tmem_kernel<n, TB_PF>;
  const int block = 128;
#else
  const size_t smem = sizeof(double) * qn::tile_problem_smem_doubles<n, false>();
  auto kern = qn::qn_resident_tile_kernel<n, TB_MINB, TB_PF>;
  const int block = 64;
#endif
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int perSM = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, block, smem);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  p.lead_div = getenv("TB_LEAD_DIV") ? atoi(getenv("TB_LEAD_DIV")) : 0; // -1 -> #SM below
  if (p.lead_div < 0) p.lead_div = sms;
  const int occ_api = perSM;
#ifdef TB_TMEM
  perSM = 4; // the occupancy API answers 1 for a kernel that allocates tensor memory; 4 x 128 columns fit
#endif
  if (getenv("TB_PERSM")) perSM = atoi(getenv("TB_PERSM")); // resident CTAs per SM the grid is sized for
  const int grid = (int)std::min<long long>((nprob + block / 64 - 1) / (block / 64), (long long)perSM * sms);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms;
  int* h_stop; unsigned long long* h_probe; cudaHostAlloc(&h_stop, 4, cudaHostAllocMapped); cudaHostAlloc(&h_probe, 16, cudaHostAllocMapped);
  cudaStream_t ps; cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking);
  double mhz = 0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaMemset(d_cnt, 0, 8);
    cudaDeviceSynchronize();
    *h_stop = 0;
    clock_probe<<<1, 1, 0, ps>>>(h_stop, h_probe);
    cudaEventRecord(e0);
    kern<<<grid, block, smem>>>(p);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    *h_stop = 1;
    cudaError_t e = cudaDeviceSynchronize();
    if (r) mhz = h_probe[1] ? 1e3 * (double)h_probe[0] / (double)h_probe[1] : 0;
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float t; cudaEventElapsedTime(&t, e0, e1);
    if (r) ms.push_back(t);
  }
  std::vector<double> x(th.size()), g(th.size()), ell(nprob);
  std::vector<int> it(nprob), fe(nprob), st(nprob);
  cudaMemcpy(x.data(), d_x, x.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(g.data(), d_g, g.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(ell.data(), d_ell, nprob * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(it.data(), d_it, nprob * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(fe.data(), d_fe, nprob * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(st.data(), d_st, nprob * 4, cudaMemcpyDeviceToHost);
  unsigned long long h = fnv(x.data(), x.size() * 8);
  h = fnv(g.data(), g.size() * 8, h); h = fnv(ell.data(), nprob * 8, h);
  h = fnv(it.data(), nprob * 4, h); h = fnv(fe.data(), nprob * 4, h); h = fnv(st.data(), nprob * 4, h);
  long long iters = 0, conv = 0; for (long long j = 0; j < nprob; ++j) { iters += it[j]; conv += st[j] == 0; }
  std::sort(ms.begin(), ms.end());
  const double med = ms[ms.size() / 2];
  printf("n=%d problems=%lld grid=%d (%d/SM, occupancy API %d) regs=%d smem=%zu: median %.3f ms (min %.3f)  %.4g iterations/s  %.4g solves/s  converged %lld  SM clock under load %.0f MHz  hash %016llx\n",
         n, nprob, grid, perSM, occ_api, fa.numRegs, smem, med, ms[0], iters / (med * 1e-3), nprob / (med * 1e-3), conv, mhz, h);
  return 0;
}
