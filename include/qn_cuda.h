/*
 * qn_cuda.h — C ABI of libqn_cuda.so, the B200 (sm_100a) BFGS engine.
 *
 * Drop-in boundary for the hot path of chriselrod/QuasiNewtonMethods.jl
 * ("QN.jl" = /root/reference/src/QuasiNewtonMethods.jl).  The reference has
 * no FFI of its own: `optimize!` is an ordinary Julia function (QN.jl:237) and
 * objectives are methods on `logdensity` / `∂logdensity!` (QN.jl:8-9).  Each
 * entry point below names the reference interface it stands in for; the
 * Julia-side `ccall` stubs are in INTEGRATION.md and
 * quasinewtonmethods.jl_b200/julia/QuasiNewtonCUDA.jl.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  All matrices column-major.
 *  - Every call returns 0 on success or a QN_ERR_* code; qn_last_error()
 *    returns a thread-local message for the last failure on this thread.
 *  - Host buffers are borrowed for the duration of the call only; the
 *    library never retains a caller pointer (QN.jl:238 `GC.@preserve`).
 *  - There is NO CPU fallback: without a usable CUDA device qn_create fails,
 *    and an objective kind without a device implementation is
 *    QN_ERR_UNREGISTERED_OBJECTIVE (the reference raises MethodError at
 *    QN.jl:254 for an objective without methods).
 *  - Per-problem failure is in-band, as in the reference (QN.jl:291):
 *    ell = NaN plus a status code.
 */
#ifndef QN_CUDA_H
#define QN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QN_VERSION 100 /* 0.1.0 */

/* ---- error codes ---------------------------------------------------- */
enum {
  QN_OK = 0,
  QN_ERR_INVALID_ARGUMENT = 1,
  QN_ERR_CUDA = 2,
  QN_ERR_UNREGISTERED_OBJECTIVE = 3, /* MethodError analogue, QN.jl:254 */
  QN_ERR_UNSUPPORTED_SIZE = 4,
  QN_ERR_NO_DEVICE = 5
};

/* ---- per-problem status (exits of optimize!, QN.jl:255,261,284,291) -- */
enum {
  QN_CONVERGED = 0,         /* maximum(abs, ∇) < tol          QN.jl:257-261 */
  QN_NONFINITE = 1,         /* !isfinite(ℓ₀)                   QN.jl:255     */
  QN_LINESEARCH_FAILED = 2, /* linesearch! returned 0          QN.jl:193,284 */
  QN_MAX_ITER = 3           /* N = 10_000 iterations exhausted QN.jl:250     */
};

/* ---- objective kinds with a registered device implementation -------- */
enum {
  QN_OBJ_ROSENBROCK = 0,     /* test/runtests.jl:4-33 */
  QN_OBJ_DENSE_GAUSSIAN = 1, /* build-defined, DESIGN.md §6 */
  QN_OBJ_LOGISTIC = 2,       /* build-defined, DESIGN.md §6 */
  QN_OBJ_USER = 100          /* the build's plug-in objective, include/qn_objective.cuh */
};

/* ---- flags ------------------------------------------------------------ */
#define QN_FLAG_WRITE_H 0x1u      /* also write final B⁻¹ into the state blocks (field 8) */
#define QN_FLAG_FORCE_SMEM 0x2u   /* use the shared-memory-resident kernel even where the
                                     register-resident one exists (testing/ablation) */
#define QN_FLAG_FORCE_STREAM 0x4u /* use the HBM streaming path regardless of n */
#define QN_FLAG_ONE_PASS 0x8u     /* permission to use the one-pass deferred-update order (16 n²
                                     instead of 24 n² bytes of B⁻¹ traffic per iteration).  Same
                                     mathematics as QN.jl:34-69 but NOT the reference's rounding:
                                     iterates agree with the default mode to ~1e-14 relative over
                                     the first iterations, optima to a few 1e-8.
                                     Honoured by the streaming path and by the shared-memory /
                                     register-pair resident kernels (48 <= n <= 160 outside the
                                     register-tile sizes); the other kernels keep the reference
                                     order, which a caller who sets the flag accepts either way */

/* BackTracking{O}(c₁, ρₕ, ρₗ, iterations): QN.jl:72-80.  order = O ∈ {2,3}. */
typedef struct {
  double c1;          /* 1e-4 */
  double rho_hi;      /* 0.5  */
  double rho_lo;      /* 0.1  */
  int64_t iterations; /* 1000 */
  int32_t order;      /* 2    */
  int32_t _pad;
} qn_backtracking;

/* Defaults of BackTracking() / BackTracking{3}(): QN.jl:79-80 */
void qn_backtracking_default(qn_backtracking* ls, int32_t order);

typedef struct qn_context qn_context; /* one per GPU; owns streams and device scratch */
typedef struct qn_model qn_model;     /* a registered device objective + its parameters */

/* ---- library / context --------------------------------------------- */
int32_t qn_version(void);
const char* qn_last_error(void);
int32_t qn_device_count(void);
int32_t qn_create(qn_context** ctx, int32_t device);
int32_t qn_destroy(qn_context* ctx);
/* name, SM count, shared memory per block (opt-in), clock (kHz) of the context's device */
int32_t qn_device_info(qn_context* ctx, char* name, size_t name_len, int32_t* sm_count,
                       int32_t* smem_per_block, int32_t* clock_khz);

/* Tuning / ablation switches of a context (they replace the QN_* environment variables, which are
 * now read once, in qn_create, as defaults).  Keys: "tile_tmem" (-1 auto | 0 | 1), "pack" (0 | 1),
 * "pack_variant" (0 | 1 | 2 | 4), "pair_rows" (0 auto | 'h' | 't' | '4' | '2' | 's'), "smem_rows" (0 | 2 | 4),
 * "lambda_smem", "stream_cluster", "stream_tma" (-1 auto | 0 | 1), "sweep_depth" (0 | 8 | 16),
 * "sweep_fine" (-1 auto | 0 | 1), "sweep_order" (0 | 1 | 2), "l2_persist_mb", "l2_resident",
 * "hint_p1_ld", "hint_p2_ld", "hint_p2_st", "hint_one_ld", "p1_pad_kb",
 * "tile_3w" (0 | 1), "lead_parity" (0 the same warp of every leader/helper pair leads | d > 0 warp
 * (blockIdx / d) & 1 | -1 d = #SM), "lead_warp" (n = 96, 128, 160 tiles: the leading warp), "host_chunks" (host-pointer entries: column ranges pipelined over
 * the context's streams; -1 auto | 1 off | up to 8), "record_exit" (0 | 1).  None of them changes a
 * result bit.  Unknown key -> QN_ERR_INVALID_ARGUMENT. */
int32_t qn_set_option(qn_context* ctx, const char* key, int64_t value);
int32_t qn_get_option(qn_context* ctx, const char* key, int64_t* value);

/* pinned host memory for callers that want full-speed host<->device copies */
int32_t qn_host_alloc(void** ptr, size_t bytes);
int32_t qn_host_free(void* ptr);

/* ---- BFGSState layout helpers (QN.jl:84-104, 135-142) ---------------- *
 * Block of doubles: x_old, ∇_new, x_new, ∇_old, y, s, B⁻¹y (L each), then
 * B⁻¹ (P×P, column stride L).  optimum = offset 0, gradient = offset L
 * (QN.jl:149-150).  L = align(P, pick_vector_width(P)) for native_w lanes.  */
int32_t qn_state_ld(int32_t n, int32_t native_w);
int64_t qn_state_len(int32_t n, int32_t ld);

/* ---- objective registry (replaces method dispatch on typeof(obj), QN.jl:8-9) */
/* kind-only objectives (Rosenbrock).  Unknown kind -> QN_ERR_UNREGISTERED_OBJECTIVE. */
int32_t qn_model_create(qn_context* ctx, int32_t kind, qn_model** model);
/* ℓ(θ) = -½ (θ-μ)ᵀ Λ (θ-μ);  mu: n, lambda: n×n column-major (host pointers, copied) */
int32_t qn_model_dense_gaussian(qn_context* ctx, int32_t n, const double* mu,
                                const double* lambda, qn_model** model);
/* ℓ(θ) = Σᵢ [yᵢ zᵢ - log(1+e^{zᵢ})] - ½ λ ‖θ‖²,  z = Xθ;  X: m×n column-major */
int32_t qn_model_logistic(qn_context* ctx, int32_t m, int32_t n, const double* X,
                          const double* y, double lam, qn_model** model);
/* The device-objective extension point: the library is built with ONE
 * user-supplied objective (include/qn_objective.cuh; `QN_USER_OBJECTIVE=my.cuh python build.py`),
 * the analogue of adding `logdensity` / `∂logdensity!` methods for a new type (QN.jl:8-9).
 * `params` is an opaque blob of `bytes` bytes that is copied to the device and handed to the
 * objective; n = 0 accepts any dimension.  A library built without a user objective returns
 * QN_ERR_UNREGISTERED_OBJECTIVE. */
int32_t qn_model_user(qn_context* ctx, int32_t n, const void* params, size_t bytes, qn_model** model);
int32_t qn_model_free(qn_model* model);

/* logdensity(obj, θ) and ∂logdensity!(∇, obj, θ) evaluated on the device for a
 * batch of points (QN.jl:8-9 protocol; used to test the device objectives).
 * theta: n×batch host; ell_out: batch; grad_out: n×batch or NULL. */
int32_t qn_logdensity_batched_f64(qn_context* ctx, const qn_model* model, int32_t n,
                                  int64_t batch, const double* theta, double* ell_out,
                                  double* grad_out);

/* ---- optimize! ------------------------------------------------------- *
 * Batched form of optimize!(state, obj, x, ls, tol) (QN.jl:237-292): problem j
 * starts from column j of theta0 (n×batch, read-only, QN.jl:248).
 *
 * Outputs (any may be NULL):
 *   states      batch BFGSState blocks, block j at states + j*state_stride
 *               doubles, column stride ld: θ* is written at offset 0 and ∇ at
 *               offset ld, so the reference's `optimum`/`gradient` accessors
 *               read them unchanged (QN.jl:135,138,149-150); B⁻¹ at offset
 *               7*ld with QN_FLAG_WRITE_H.
 *   x_out,g_out compact n×batch copies of θ* and ∇
 *   ell_out     ℓ(θ*) or NaN                      (return value of optimize!)
 *   iters_out   completed line-search steps
 *   fevals_out  objective evaluations (value-only + value-and-gradient)
 *   status_out  QN_CONVERGED / ...
 *   trace_out   ntrace×n×batch... laid out [batch][ntrace][n]: x_old after each of
 *               the first ntrace iterations (rows past the last iteration untouched)
 * max_iter: the reference hard-codes 10000 (QN.jl:250).
 */
int32_t qn_bfgs_optimize_batched_f64(qn_context* ctx, const qn_model* model, int32_t n,
                                     int64_t batch, const double* theta0,
                                     const qn_backtracking* ls, double tol, int32_t max_iter,
                                     double* states, int64_t state_stride, int32_t ld,
                                     double* x_out, double* g_out, double* ell_out,
                                     int32_t* iters_out, int32_t* fevals_out,
                                     int32_t* status_out, int32_t ntrace, double* trace_out,
                                     uint32_t flags);

/* Same, with every pointer a DEVICE pointer on the context's GPU and the work
 * enqueued on `stream` (a cudaStream_t, used literally: NULL is the CUDA default
 * stream) without synchronising — the PtrBFGSState analogue (QN.jl:117-121): the
 * caller owns the memory.  theta_stride = column stride of theta0 (>= n). */
int32_t qn_bfgs_optimize_batched_device_f64(qn_context* ctx, const qn_model* model, int32_t n,
                                            int64_t batch, const double* theta0,
                                            int64_t theta_stride, const qn_backtracking* ls,
                                            double tol, int32_t max_iter, double* states,
                                            int64_t state_stride, int32_t ld, double* x_out,
                                            double* g_out, double* ell_out, int32_t* iters_out,
                                            int32_t* fevals_out, int32_t* status_out,
                                            int32_t ntrace, double* trace_out, uint32_t flags,
                                            void* stream);

/* Single problem, exactly optimize!(state, obj, x, ls, tol) (QN.jl:237): `state`
 * is one BFGSState block with column stride ld; *ell receives the return value. */
int32_t qn_bfgs_optimize_f64(qn_context* ctx, const qn_model* model, int32_t n,
                             const double* x, const qn_backtracking* ls, double tol,
                             double* state, int32_t ld, double* ell, int32_t* iters,
                             int32_t* fevals, int32_t* status, uint32_t flags);

/* Batch sharded over several GPUs of one box from ONE process (problem j goes
 * to context j mod nctx; no inter-GPU communication).  Host pointers. */
int32_t qn_bfgs_optimize_batched_multi_f64(qn_context* const* ctxs, int32_t nctx,
                                           const qn_model* const* models, int32_t n,
                                           int64_t batch, const double* theta0,
                                           const qn_backtracking* ls, double tol,
                                           int32_t max_iter, double* x_out, double* g_out,
                                           double* ell_out, int32_t* iters_out,
                                           int32_t* fevals_out, int32_t* status_out,
                                           uint32_t flags);

/* Row-sharded large-n mode (BASELINE.json configs[4], "optional row-sharded
 * 8-GPU mode"): ONE problem on up to 8 GPUs of one process.  Context g holds the rows of its
 * 512-row tiles of B⁻¹ (QN.jl:48,61: the unit-stride index c) and a replica of every O(n) vector;
 * the row sums of each sweep are exchanged with peer stores over NVLink and per-shard epoch flags
 * inside the kernels — no host or NCCL call inside an iteration — and every reduction is
 * replicated, so the result is the single-GPU result bit for bit.  Replaces the same reference
 * call as qn_bfgs_optimize_f64: optimize!(state, obj, x, ls, tol), QN.jl:237-292.
 * Host pointers.  h_out (n × ldh, column-major) may be NULL; trace_out is ntrace × n.
 * bench_iters > 0: run exactly that many iterations (first two untimed) and report
 * *ms_per_iter; otherwise iterate to convergence / max_iter. */
int32_t qn_bfgs_optimize_rowsharded_f64(qn_context* const* ctxs, int32_t nctx,
                                        const qn_model* model, int32_t n, const double* theta0,
                                        const qn_backtracking* ls, double tol, int32_t max_iter,
                                        double* x_out, double* g_out, double* h_out, int32_t ldh,
                                        double* ell_out, int32_t* iters_out, int32_t* fevals_out,
                                        int32_t* status_out, int32_t ntrace, double* trace_out,
                                        uint32_t flags, int32_t bench_iters, double* ms_per_iter);

/* How the last optimize call on this context was executed:
 * path 0 = register-resident, 1 = shared-memory-resident, 2 = HBM streaming. */
int32_t qn_last_launch_info(qn_context* ctx, int32_t* path, int32_t* grid, int32_t* block,
                            int32_t* smem_bytes, int32_t* kernel_launches);

/* Tail of the last batched launch made with option "record_exit" = 1 (n = 36..64 tile kernels; call
 * after the results are back): span_ms = first start to last end on the device, idle_fraction =
 * share of the launch's problem-slot time spent after the work queue ran dry, slots = resident
 * problem slots.  Names the limiter of strong scaling (bench.py). */
int32_t qn_last_tail_profile(qn_context* ctx, double* span_ms, double* idle_fraction, int32_t* slots);

/* Streaming (large-n) path benchmark: `iters` BFGS iterations (first two untimed) of an
 * n-dimensional Rosenbrock problem with B⁻¹ in HBM.  gbs = achieved ALGORITHMIC bandwidth,
 * (16 n² + 96 n) bytes per iteration / measured time. */
int32_t qn_stream_bench_f64(qn_context* ctx, int32_t n, int32_t iters, int32_t order,
                            uint32_t flags, double* ms_per_iter, double* gbs);

/* ---- microbenchmarks that define the roofline denominators */
/* sustained fp64 FMA rate, TFLOP/s (2 flops per DFMA), all SMs busy */
int32_t qn_microbench_dfma(qn_context* ctx, double* tflops);
/* sustained shared-memory load bandwidth, GB/s, conflict-free LDS.128 on all SMs */
int32_t qn_microbench_lds(qn_context* ctx, double* gbs);
/* device copy bandwidth (read+write bytes), GB/s */
int32_t qn_microbench_hbm(qn_context* ctx, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* QN_CUDA_H */
