// phase_timing.cu — per-phase SM-cycle accounting of the register-tile kernel for ONE CTA
// (QN_PHASE_TIMING build of qn_resident_tile.cuh).  Usage: ./phase_timing [problems = resident CTAs / pairs]
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -DQN_PHASE_TIMING -std=c++17 \
//        --expt-relaxed-constexpr -o phase_timing phase_timing.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "../quasinewtonmethods.jl_b200/csrc/qn_resident_tile.cuh"

int main(int argc, char** argv) {
  const int n = 60;
  const int nprob = argc > 1 ? atoi(argv[1]) : 1;
  std::mt19937_64 rng(1234);
  std::normal_distribution<double> nd;
  std::vector<double> th((size_t)n * nprob);
  for (auto& v : th) v = nd(rng);
  double *d_th, *d_x; int* d_it; unsigned long long* d_cnt;
  cudaMalloc(&d_th, th.size() * 8); cudaMalloc(&d_x, th.size() * 8); cudaMalloc(&d_it, nprob * 4); cudaMalloc(&d_cnt, 8);
  cudaMemcpy(d_th, th.data(), th.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(d_cnt, 0, 8);
  qn::BatchParams p{};
  p.n = n; p.max_iter = 10000; p.batch = nprob; p.theta0 = d_th; p.theta_stride = n; p.tol = 1e-8;
  p.ls = qn::LsParams{1e-4, 0.5, 0.1, 1000, 2};
  p.x_out = d_x; p.x_stride = n; p.iters_out = d_it; p.counter = d_cnt; p.lead_div = 0;
#ifdef PT_TMEM // the tensor-memory kernel: CTAs of 128 threads = two problems (build with -DPT_TMEM)
  const size_t smem = 2 * sizeof(double) * qn::tile_problem_smem_doubles<60, true>();
  cudaFuncSetAttribute(qn::qn_resident_tile_tmem_kernel<60, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  qn::qn_resident_tile_tmem_kernel<60, 1><<<(nprob + 1) / 2, 128, smem>>>(p);
#else
  const size_t smem = sizeof(double) * qn::tile_problem_smem_doubles<60, false>();
  qn::qn_resident_tile_kernel<60, 6, 1><<<nprob, 64, smem>>>(p);
#endif
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  unsigned long long clk[16];
  cudaMemcpyFromSymbol(clk, qn::g_phase_clk, sizeof clk);
  std::vector<int> it(nprob);
  cudaMemcpy(it.data(), d_it, nprob * 4, cudaMemcpyDeviceToHost);
  // block 0 solved whichever problems it pulled; with grid == batch it solved about one
  const char* names[9] = {"accept -> loop top", "convergence check", "publish + command", "rho + pass 1 + wait for the partials",
                          "combine + yHy + c1 + publish", "pass 2 + wait for the partials", "combine + first trial + m butterfly",
                          "line search, scalar part", "line search, extra trials"};
  unsigned long long tot = 0; for (int k = 0; k < 9; ++k) tot += clk[k];
  printf("CTAs=%d  iterations of problem 0: %d  total cycles %llu  -> %.0f cycles/iteration\n", nprob, it[0], tot, (double)tot / it[0]);
  for (int k = 0; k < 9; ++k) printf("  %-50s %9llu  %6.1f cyc/iter  %5.1f%%\n", names[k], clk[k], (double)clk[k] / it[0], 100.0 * clk[k] / tot);
  printf("  extra trials: %llu (%.3f per iteration, %.0f cycles each)\n", clk[9], (double)clk[9] / it[0], clk[9] ? (double)clk[8] / clk[9] : 0.0);
  return 0;
}
