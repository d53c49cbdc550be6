// qn_eval.cu — logdensity(obj, θ) / ∂logdensity!(∇, obj, θ) for a batch of points
// (the objective protocol of QN.jl:8-9, evaluated by the registered device
// objectives).  Exists so the device objectives can be tested in isolation.
#include "qn_device.cuh"
#include "qn_internal.h"
#include "qn_objectives.cuh"
#ifdef QN_USER_OBJECTIVE_HEADER
#include QN_USER_OBJECTIVE_HEADER
#endif

namespace qn {

template <class Obj>
__global__ void __launch_bounds__(256, 1)
qn_eval_kernel(int n, long long batch, const double* __restrict__ theta, double* __restrict__ ell,
               double* __restrict__ grad, const typename Obj::Params op) {
  extern __shared__ __align__(16) double smem[];
  const int B = blockDim.x;
  Team tm;
  tm.n = n;
  tm.c = threadIdx.x;
  tm.lane = threadIdx.x & 31;
  tm.red = smem;
  tm.stride = B;
  tm.phase = 0;
  tm.xs = smem + 2 * B;
  tm.nvec = B;
  tm.tpr = 1;
  tm.part = nullptr;
  tm.scratch = smem + 3 * B;
  Obj::prologue(tm, op); // once per CTA (include/qn_objective.cuh)
  for (long long j = blockIdx.x; j < batch; j += gridDim.x) {
    __syncthreads();
    tm.xs[tm.c] = (tm.c < n) ? theta[j * n + tm.c] : 0.0;
    __syncthreads();
    double g = 0.0;
    double l;
    if (grad) l = Obj::template eval<true>(tm, op, g);
    else l = Obj::template eval<false>(tm, op, g);
    if (grad && tm.c < n) grad[j * n + tm.c] = g;
    if (tm.c == 0) ell[j] = l;
  }
}

template <class Obj>
int launch_eval(qn_context* ctx, int n, long long batch, const double* theta, double* ell,
                double* grad, const typename Obj::Params& op, cudaStream_t st) {
  const int B = (n + 31) & ~31;
  const size_t smem = sizeof(double) * (size_t)(3 * B + Obj::scratch_doubles(n, op));
  if (B > 256 || smem > (size_t)ctx->prop.sharedMemPerBlockOptin)
    return qn_fail(QN_ERR_UNSUPPORTED_SIZE, "qn_logdensity_batched_f64: n=%d too large", n);
  auto kern = qn_eval_kernel<Obj>;
  QN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = batch < 4LL * ctx->prop.multiProcessorCount ? batch : 4LL * ctx->prop.multiProcessorCount;
  kern<<<(unsigned)grid, B, smem, st>>>(n, batch, theta, ell, grad, op);
  QN_CUDA_TRY(cudaGetLastError());
  return QN_OK;
}

}  // namespace qn

extern "C" int32_t qn_logdensity_batched_f64(qn_context* ctx, const qn_model* model, int32_t n,
                                             int64_t batch, const double* theta, double* ell_out,
                                             double* grad_out) {
  if (!ctx) return qn_fail(QN_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (!model) return qn_fail(QN_ERR_UNREGISTERED_OBJECTIVE, "model is NULL (unregistered objective)");
  if (model->ctx != ctx) return qn_fail(QN_ERR_INVALID_ARGUMENT, "model belongs to another context");
  if (n < 1 || batch < 0 || !theta || !ell_out) return qn_fail(QN_ERR_INVALID_ARGUMENT, "bad argument");
  if (model->n && model->n != n)
    return qn_fail(QN_ERR_INVALID_ARGUMENT, "model has n=%d, call has n=%d", model->n, n);
  if (batch == 0) return QN_OK;
  QN_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t nb = sizeof(double) * (size_t)n * batch;
  const size_t nbp = (nb + 255) & ~(size_t)255;
  const size_t eb = (sizeof(double) * batch + 255) & ~(size_t)255;
  std::lock_guard<std::mutex> lock(ctx->mu);
  int rc = qn_ensure_stage(ctx, 2 * nbp + eb);
  if (rc) return rc;
  char* base = (char*)ctx->stage;
  double* d_theta = (double*)base;
  double* d_grad = grad_out ? (double*)(base + nbp) : nullptr;
  double* d_ell = (double*)(base + 2 * nbp);
  cudaStream_t st = ctx->stream;
  QN_CUDA_TRY(cudaMemcpyAsync(d_theta, theta, nb, cudaMemcpyHostToDevice, st));
  switch (model->kind) {
    case QN_OBJ_ROSENBROCK:
      rc = qn::launch_eval<qn::Rosenbrock>(ctx, n, batch, d_theta, d_ell, d_grad, qn::Rosenbrock::Params{}, st);
      break;
    case QN_OBJ_DENSE_GAUSSIAN:
      rc = qn::launch_eval<qn::DenseGaussian>(ctx, n, batch, d_theta, d_ell, d_grad,
                                              qn::DenseGaussian::Params{model->d_a, model->d_b}, st);
      break;
    case QN_OBJ_LOGISTIC:
      rc = qn::launch_eval<qn::Logistic>(
          ctx, n, batch, d_theta, d_ell, d_grad,
          qn::Logistic::Params{model->m, model->d_a, model->d_b, model->d_c, model->lam}, st);
      break;
#ifdef QN_USER_OBJECTIVE_HEADER
    case QN_OBJ_USER:
      rc = qn::launch_eval<qn::UserObjective>(ctx, n, batch, d_theta, d_ell, d_grad,
                                              qn::UserObjective::make(model->d_a, model->bytes, n), st);
      break;
#endif
    default:
      rc = qn_fail(QN_ERR_UNREGISTERED_OBJECTIVE,
                   "objective kind %d has no registered device implementation", model->kind);
  }
  if (rc) return rc;
  QN_CUDA_TRY(cudaMemcpyAsync(ell_out, d_ell, sizeof(double) * batch, cudaMemcpyDeviceToHost, st));
  if (grad_out) QN_CUDA_TRY(cudaMemcpyAsync(grad_out, d_grad, nb, cudaMemcpyDeviceToHost, st));
  QN_CUDA_TRY(cudaStreamSynchronize(st));
  return QN_OK;
}
