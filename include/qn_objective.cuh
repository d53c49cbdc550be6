// qn_objective.cuh — how to give libqn_cuda.so a device objective of your own.
//
// The reference registers an objective by adding methods to two generic functions,
// `logdensity(obj, θ)` and `∂logdensity!(∇, obj, θ)` (QN.jl:8-9).  The B200 engine runs the whole
// optimize! loop (QN.jl:237-292) inside one kernel, so the objective must be device code that is
// compiled INTO the library: write one header that defines `qn::UserObjective` as below and build
//
//     QN_USER_OBJECTIVE=/abs/path/my_objective.cuh QN_SO_OUT=/abs/path/libqn_cuda_mine.so \
//         python quasinewtonmethods.jl_b200/build.py
//
// then create models with qn_model_user(ctx, n, blob, bytes, &model) (kind QN_OBJ_USER) and call
// the same optimize entry points.  The header is included after csrc/qn_objectives.cuh, so
// `qn::Team`, `qn::chain_rowsum`, `qn::block_chain`, `qn::warp_butterfly` are available; see the
// three shipped objectives there and examples/quartic_objective.cuh for a complete plug-in.
//
//   namespace qn {
//   struct UserObjective {
//     struct Params { ... };                       // trivially copyable; passed to the kernel by value
//     // d_blob: the device copy of the bytes given to qn_model_user (lives as long as the model)
//     static Params make(const void* d_blob, size_t bytes, int n);            // host
//     static constexpr int kKind = QN_OBJ_USER;
//     static constexpr bool kRegisterPaths = false; // true: also instantiate the thread-per-row kernels (n = 8, 16, 32)
//     static constexpr bool kPairRegister = false;  // true: n = 128 keeps B⁻¹ in register tiles
//     static constexpr bool kMemoizeTrials = false; // true (expensive objectives): a line-search trial whose point is bit for
//                                                   // bit the previous trial's re-uses its value instead of evaluating again
//     // doubles of shared scratch (Team::scratch) the objective needs for dimension n
//     __host__ __device__ static int scratch_doubles(int n, const Params&);
//     // once per CTA before its first problem (e.g. stage parameters in shared memory)
//     __device__ static void prologue(Team&, const Params&);
//     // Called by EVERY thread of the CTA after θ was published in team.xs[0..n) and a barrier.
//     // Returns ℓ(θ) in every thread; with GRAD leaves ∇[c] in g for the thread with team.c == c < n.
//     // Must contain at least one __syncthreads() after its last read of team.xs / scratch and before
//     // returning (Team::csum does), because the caller overwrites team.xs right away.
//     template <bool GRAD>
//     __device__ static double eval(Team&, const Params&, double& g);
//   };
//   }
//
// Sums that feed ℓ should go through Team::csum (the canonical order) so results do
// not depend on the launch shape.  Only the resident path (n <= 160) is open to user objectives.
#pragma once
