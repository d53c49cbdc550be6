# runtests_cuda.jl — the reference's own testset (QuasiNewtonMethods.jl test/runtests.jl:35-52) pointed
# at the CUDA engine, plus the batched form.  NEVER EXECUTED HERE: the build image has no Julia
# (SURVEY.md §0, §8(f) row 4).  On a machine with Julia >= 1.6, a B200 and the built library:
#
#     LIBQN_CUDA=/path/to/libqn_cuda.so julia --project=/path/to/QuasiNewtonMethods.jl \
#         -e 'push!(LOAD_PATH, "/path/to/quasinewtonmethods.jl_b200/julia"); include("runtests_cuda.jl")'
#
# The ONLY change against the reference's file is the definition of `Rosenbrock`: the reference
# defines a struct with `logdensity` / `∂logdensity!` methods in Julia (test/runtests.jl:4-33); here
# the name is bound to the objective with the registered device implementation.  `optimize!`,
# `BFGSState`, `optimum`, `gradient` and `BackTracking{3}` are the reference's own.
using QuasiNewtonMethods, StrideArrays
using Test
include(joinpath(@__DIR__, "QuasiNewtonCUDA.jl"))
using .QuasiNewtonCUDA

const Rosenbrock = QuasiNewtonCUDA.Rosenbrock   # <- the one-line swap

@testset "QuasiNewtonMethods.jl on libqn_cuda" begin
  # test/runtests.jl:37-51, verbatim
  for n ∈ 2:24
    @show n
    state = QuasiNewtonMethods.BFGSState{n}(undef)
    x = @StrideArray randn(StaticInt(n))
    # 2nd order
    @test abs(optimize!(state, Rosenbrock(), x)) < 2eps()
    @show QuasiNewtonMethods.optimum(state) .- 1
    @test all(x -> x ≈ 1, QuasiNewtonMethods.optimum(state))
    @test maximum(abs, QuasiNewtonMethods.gradient(state)) < 1e-8
    # 3rd order
    @test abs(optimize!(state, Rosenbrock(), x, QuasiNewtonMethods.BackTracking{3}())) < eps()
    @show QuasiNewtonMethods.optimum(state) .- 1
    @test all(x -> x ≈ 1, QuasiNewtonMethods.optimum(state))
    @test maximum(abs, QuasiNewtonMethods.gradient(state)) < 1e-8
  end
  # README.md:52-75, n = 60
  let n = 60
    state = QuasiNewtonMethods.BFGSState{n}(undef)
    x = @StrideArray randn(StaticInt(n))
    @test abs(optimize!(state, Rosenbrock(), x)) < eps()
    @test QuasiNewtonMethods.optimum(state) ≈ fill(1, n)
    @test maximum(abs, QuasiNewtonMethods.gradient(state)) < 1e-8
  end
end

# Trajectory parity with the REAL reference (closes DESIGN.md §2's "unpinned" gap): the same start
# through the reference's CPU method (the test file's Julia Rosenbrock) and through the GPU.
struct RosenbrockJl end
function QuasiNewtonMethods.logdensity(::RosenbrockJl, θ)
  s = zero(eltype(θ)); N = length(θ) >> 1
  @inbounds @simd for i ∈ 1:N
    s -= 100(θ[i+N] - θ[i]^2)^2 + (1 - θ[i])^2
  end
  isodd(length(θ)) ? -muladd(1 - θ[end], 1 - θ[end], -s) : s
end
function QuasiNewtonMethods.∂logdensity!(∇, ::RosenbrockJl, θ)
  s = zero(eltype(θ)); N = length(θ) >> 1
  @inbounds @simd for i ∈ 1:N
    s -= 100(θ[i+N] - θ[i]^2)^2 + (1 - θ[i])^2
    ∇[i] = 400(θ[i+N] - θ[i]^2) * θ[i] + 2(1 - θ[i])
    ∇[i+N] = 200(θ[i]^2 - θ[i+N])
  end
  if isodd(length(θ))
    δ = 1 - θ[end]; s = -muladd(δ, δ, -s); ∇[end] = 2δ
  end
  s
end

@testset "GPU vs the reference's CPU optimize!" begin
  for n ∈ (2, 7, 24, 60)
    x = @StrideArray randn(StaticInt(n))
    cpu = QuasiNewtonMethods.BFGSState{n}(undef); gpu = QuasiNewtonMethods.BFGSState{n}(undef)
    ℓc = optimize!(cpu, RosenbrockJl(), x)
    ℓg = optimize!(gpu, Rosenbrock(), x)
    @test isfinite(ℓc) == isfinite(ℓg)
    @test maximum(abs, QuasiNewtonMethods.optimum(cpu) .- QuasiNewtonMethods.optimum(gpu)) ≤ 5e-8   # north-star gate is 1e-8; two roundings of one trajectory differ by up to ~3e-8 (SURVEY §7.3-B)
    @test maximum(abs, QuasiNewtonMethods.gradient(gpu)) < 1e-8
  end
end

@testset "batched optimize!(states, obj, Θ)" begin
  n, batch = 60, 4096
  Θ = randn(n, batch)
  states = [QuasiNewtonMethods.BFGSState{n}(undef) for _ in 1:batch]
  r = optimize!(states, Rosenbrock(), Θ)
  @test all(==(0), r.status)
  @test all(s -> maximum(abs, QuasiNewtonMethods.gradient(s)) < 1e-8, states)
  @test all(s -> all(x -> x ≈ 1, QuasiNewtonMethods.optimum(s)), states)
  # an objective without a device implementation never falls back to the CPU silently:
  # a non-DeviceObjective simply is the reference's CPU method; an unknown kind is an error
  @test_throws QuasiNewtonCUDA.QNError QuasiNewtonCUDA.check(ccall((:qn_model_create, QuasiNewtonCUDA.libqn), Int32,
      (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), QuasiNewtonCUDA.default_context().ptr, 77, Ref{Ptr{Cvoid}}(C_NULL)))
end
