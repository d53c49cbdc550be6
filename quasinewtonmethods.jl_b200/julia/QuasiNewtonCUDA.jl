# QuasiNewtonCUDA.jl — Julia-side shim over libqn_cuda.so.
#
# STATUS: NEVER EXECUTED.  No Julia runtime exists in the build image or on the GPU boxes
# (SURVEY.md §0).  It is the binding a maintainer of QuasiNewtonMethods.jl would add; the Python host
# layer (../api.py) is the tested mirror of the same C ABI, call for call.  Every ccall below matches
# a prototype in include/qn_cuda.h.  `julia/runtests_cuda.jl` is the reference's own testset pointed
# at this module, ready for whoever has Julia and a B200 (SURVEY.md §8(f) row 4).
#
# The module ADDS METHODS to the reference's own generic function `QuasiNewtonMethods.optimize!`
# (QuasiNewtonMethods.jl:237), dispatching on the objective's type: an objective that is a
# `DeviceObjective` runs on the GPU, any other objective keeps the reference's CPU method.  So the
# reference's calls work unchanged:
#
#     using QuasiNewtonMethods, QuasiNewtonCUDA
#     state = QuasiNewtonMethods.BFGSState{60}(undef)
#     x = @StrideArray randn(StaticInt(60))                         # or a plain Vector{Float64}
#     optimize!(state, QuasiNewtonCUDA.Rosenbrock(), x)             # README.md:52-59, on the GPU
#     optimize!(state, QuasiNewtonCUDA.Rosenbrock(), x, QuasiNewtonMethods.BackTracking{3}())
#     QuasiNewtonMethods.optimum(state); QuasiNewtonMethods.gradient(state)
#     states = [QuasiNewtonMethods.BFGSState{60}(undef) for _ in 1:65536]
#     optimize!(states, QuasiNewtonCUDA.Rosenbrock(), randn(60, 65536))            # batched, one GPU
#     optimize!(states, QuasiNewtonCUDA.Rosenbrock(), randn(60, 65536);
#               ctxs = [QuasiNewtonCUDA.Context(g) for g in 0:7])                  # batched, 8 GPUs
module QuasiNewtonCUDA

using QuasiNewtonMethods
using QuasiNewtonMethods: AbstractBFGSState, BFGSState, BackTracking
import QuasiNewtonMethods: optimize!, logdensity, ∂logdensity!
using StrideArraysCore: AbstractStrideArray

const libqn = get(ENV, "LIBQN_CUDA", "libqn_cuda.so")

# ---- errors ------------------------------------------------------------------------------
struct QNError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::QNError) = print(io, "libqn_cuda error ", e.code, ": ", e.msg)

const QN_ERR_UNREGISTERED_OBJECTIVE = Int32(3)

last_error() = unsafe_string(ccall((:qn_last_error, libqn), Cstring, ()))
function check(rc::Int32)
    rc == 0 && return nothing
    # QN_ERR_UNREGISTERED_OBJECTIVE is the MethodError analogue (QuasiNewtonMethods.jl:254):
    # an objective without a device implementation is an error, never a CPU fallback.
    throw(QNError(rc, last_error()))
end

# ---- BackTracking{O} -> qn_backtracking (QuasiNewtonMethods.jl:72-80) ----------------------
struct CBackTracking
    c1::Float64
    rho_hi::Float64
    rho_lo::Float64
    iterations::Int64
    order::Int32
    _pad::Int32
end
CBackTracking(ls::BackTracking{O}) where {O} =
    CBackTracking(ls.c₁, ls.ρₕ, ls.ρₗ, ls.iterations, Int32(O), Int32(0))

# ---- context (one per GPU); owns the device-side registrations of objectives -------------------
mutable struct Context
    ptr::Ptr{Cvoid}
    models::IdDict{Any,Ptr{Cvoid}}   # objective (by identity) -> qn_model*, uploaded once
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qn_create, libqn), Int32, (Ref{Ptr{Cvoid}}, Int32), r, device))
        ctx = new(r[], IdDict{Any,Ptr{Cvoid}}())
        finalizer(ctx) do c
            for h in values(c.models)
                ccall((:qn_model_free, libqn), Int32, (Ptr{Cvoid},), h)
            end
            ccall((:qn_destroy, libqn), Int32, (Ptr{Cvoid},), c.ptr)
        end
        ctx
    end
end
const DEFAULT_CONTEXT = Ref{Union{Nothing,Context}}(nothing)
function default_context()
    if DEFAULT_CONTEXT[] === nothing
        DEFAULT_CONTEXT[] = Context(0)
    end
    DEFAULT_CONTEXT[]::Context
end
"qn_set_option: tuning / ablation switches of a context (include/qn_cuda.h lists the keys)."
set_option!(ctx::Context, key::AbstractString, value::Integer) =
    check(ccall((:qn_set_option, libqn), Int32, (Ptr{Cvoid}, Cstring, Int64), ctx.ptr, key, value))

# ---- objective registry: replaces dispatch on typeof(obj) (QuasiNewtonMethods.jl:8-9) --------
"Objectives with a registered CUDA device implementation."
abstract type DeviceObjective end
struct Rosenbrock <: DeviceObjective end                       # test/runtests.jl:4-33, kind 0
struct DenseGaussian <: DeviceObjective                        # kind 1
    μ::Vector{Float64}
    Λ::Matrix{Float64}
end
struct LogisticRegression <: DeviceObjective                   # kind 2
    X::Matrix{Float64}
    y::Vector{Float64}
    λ::Float64
end
"""
The build's plug-in objective (include/qn_objective.cuh, kind QN_OBJ_USER = 100): `params` is the
opaque blob the objective's `make` interprets; needs a library built with QN_USER_OBJECTIVE
(point LIBQN_CUDA at it).  The stock library answers QN_ERR_UNREGISTERED_OBJECTIVE.
"""
struct UserObjective <: DeviceObjective
    params::Vector{Float64}
    n::Int
end

function register(ctx::Context, ::Rosenbrock)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qn_model_create, libqn), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), ctx.ptr, 0, r))
    r[]
end
function register(ctx::Context, m::DenseGaussian)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qn_model_dense_gaussian, libqn), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                ctx.ptr, length(m.μ), m.μ, m.Λ, r))
    r[]
end
function register(ctx::Context, m::LogisticRegression)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qn_model_logistic, libqn), Int32,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
                ctx.ptr, size(m.X, 1), size(m.X, 2), m.X, m.y, m.λ, r))
    r[]
end
function register(ctx::Context, m::UserObjective)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qn_model_user, libqn), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}),
                ctx.ptr, m.n, m.params, sizeof(m.params), r))
    r[]
end
# the handle of an objective on a context: uploaded on first use, cached for the context's life
# (parameter-free singletons like Rosenbrock() share one entry: isbits structs are === equal)
model_handle(ctx::Context, obj::DeviceObjective) = get!(() -> register(ctx, obj), ctx.models, obj)

# ---- the objective protocol itself, evaluated by the device implementation (QN.jl:8-9) ---------
function logdensity(obj::DeviceObjective, θ::AbstractVector{Float64}; ctx::Context=default_context())
    ell = Ref{Float64}(NaN)
    θc = Vector{Float64}(θ)
    check(ccall((:qn_logdensity_batched_f64, libqn), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                ctx.ptr, model_handle(ctx, obj), length(θc), 1, θc, ell, C_NULL))
    ell[]
end
function ∂logdensity!(∇::AbstractVector{Float64}, obj::DeviceObjective, θ::AbstractVector{Float64};
                      ctx::Context=default_context())
    ell = Ref{Float64}(NaN)
    θc = Vector{Float64}(θ); g = Vector{Float64}(undef, length(θc))
    check(ccall((:qn_logdensity_batched_f64, libqn), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                ctx.ptr, model_handle(ctx, obj), length(θc), 1, θc, ell, g))
    copyto!(∇, g)
    ell[]
end

# ---- optimize!(state, obj, x, ls, tol): methods of QuasiNewtonMethods.optimize! (QN.jl:237) ------
# The reference's method is optimize!(state, obj, x::AbstractStrideArray{T}, ls, tol) with `state`
# and `obj` untyped; the methods below are strictly more specific in `obj` (and no less specific in
# `x`), so there is no ambiguity and every other objective type still reaches the CPU method.
function _optimize_device!(state::AbstractBFGSState{P,Float64,L}, obj::DeviceObjective, x,
                           ls::BackTracking, tol, ctx::Context) where {P,L}
    length(x) == P || throw(DimensionMismatch("x has length $(length(x)), state is BFGSState{$P}"))
    h = model_handle(ctx, obj)
    cls = Ref(CBackTracking(ls))
    xc = x isa Vector{Float64} ? x : Vector{Float64}(x)   # contiguous copy; x is read-only (QN.jl:248)
    ell = Ref{Float64}(NaN); iters = Ref{Int32}(0); fevals = Ref{Int32}(0); status = Ref{Int32}(0)
    GC.@preserve state obj xc begin      # QuasiNewtonMethods.jl:238
        check(ccall((:qn_bfgs_optimize_f64, libqn), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{CBackTracking}, Float64,
                     Ptr{Float64}, Int32, Ref{Float64}, Ref{Int32}, Ref{Int32}, Ref{Int32}, UInt32),
                    ctx.ptr, h, P, pointer(xc), cls, Float64(tol), pointer(state), L, ell, iters, fevals, status, 0))
    end
    ell[]   # ℓ at the optimum, or NaN (QuasiNewtonMethods.jl:261,291)
end
optimize!(state::AbstractBFGSState{P,Float64,L}, obj::DeviceObjective, x::AbstractStrideArray{Float64},
          ls::BackTracking=BackTracking(), tol=1e-8; ctx::Context=default_context()) where {P,L} =
    _optimize_device!(state, obj, x, ls, tol, ctx)
optimize!(state::AbstractBFGSState{P,Float64,L}, obj::DeviceObjective, x::AbstractVector{Float64},
          ls::BackTracking=BackTracking(), tol=1e-8; ctx::Context=default_context()) where {P,L} =
    _optimize_device!(state, obj, x, ls, tol, ctx)

# ---- batched optimize!(states, obj, Θ): one problem per column of Θ ---------------------------
# `states` is a vector of BFGSState{P}; θ* lands at offset 0 and ∇ at offset L of each state's block
# (QuasiNewtonMethods.jl:135,138), so `optimum(states[j])` and `gradient(states[j])` work unchanged.
# Mutable-struct states are separate heap objects, so the batch comes back as two contiguous
# matrices that are copied into the states.  `ctxs`: several contexts (one per GPU) shard the batch,
# problem j -> ctxs[j mod G], through qn_bfgs_optimize_batched_multi_f64 (no inter-GPU traffic).
function optimize!(states::AbstractVector{<:BFGSState{P,Float64,L}}, obj::DeviceObjective,
                   Θ::AbstractMatrix{Float64}, ls::BackTracking=BackTracking(), tol=1e-8;
                   ctx::Context=default_context(), ctxs::Union{Nothing,Vector{Context}}=nothing) where {P,L}
    batch = size(Θ, 2)
    (size(Θ, 1) == P && length(states) == batch) || throw(DimensionMismatch("Θ must be $P × $(length(states))"))
    cls = Ref(CBackTracking(ls))
    Θc = Θ isa Matrix{Float64} ? Θ : Matrix{Float64}(Θ)
    xs = Matrix{Float64}(undef, P, batch); gs = Matrix{Float64}(undef, P, batch)
    ell = Vector{Float64}(undef, batch)
    iters = Vector{Int32}(undef, batch); fevals = Vector{Int32}(undef, batch); status = Vector{Int32}(undef, batch)
    if ctxs === nothing || length(ctxs) == 1
        c = ctxs === nothing ? ctx : ctxs[1]
        h = model_handle(c, obj)
        GC.@preserve Θc xs gs ell iters fevals status begin
            check(ccall((:qn_bfgs_optimize_batched_f64, libqn), Int32,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{CBackTracking}, Float64, Int32,
                         Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, UInt32),
                        c.ptr, h, P, batch, Θc, cls, Float64(tol), 10_000,
                        C_NULL, 0, 0, xs, gs, ell, iters, fevals, status, 0, C_NULL, 0))
        end
    else
        cptrs = Ptr{Cvoid}[c.ptr for c in ctxs]
        mptrs = Ptr{Cvoid}[model_handle(c, obj) for c in ctxs]
        GC.@preserve Θc xs gs ell iters fevals status cptrs mptrs begin
            check(ccall((:qn_bfgs_optimize_batched_multi_f64, libqn), Int32,
                        (Ptr{Ptr{Cvoid}}, Int32, Ptr{Ptr{Cvoid}}, Int32, Int64, Ptr{Float64}, Ref{CBackTracking},
                         Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32},
                         Ptr{Int32}, UInt32),
                        cptrs, length(cptrs), mptrs, P, batch, Θc, cls, Float64(tol), 10_000,
                        xs, gs, ell, iters, fevals, status, 0))
        end
    end
    for j in 1:batch
        st = states[j]
        GC.@preserve st begin
            p = pointer(st)
            unsafe_copyto!(p, pointer(xs, (j - 1) * P + 1), P)                 # ref_x_old, offset 0
            unsafe_copyto!(p + L * sizeof(Float64), pointer(gs, (j - 1) * P + 1), P)  # ref_∇_new, offset L
        end
    end
    (ℓ = ell, iterations = iters, evaluations = fevals, status = status)
end

# ---- one LARGE problem on several GPUs (row-sharded B⁻¹, include/qn_cuda.h) --------------------
# optimize!(state, obj, x) for n in the thousands: context g holds the rows of its 512-row tiles of
# B⁻¹; same result as one GPU, bit for bit.  `one_pass = true` permits the deferred-update order
# (QN_FLAG_ONE_PASS).  Returns ℓ (or NaN) like optimize!; θ* and ∇ land in the state block.
const QN_FLAG_ONE_PASS = UInt32(0x8)
function optimize_rowsharded!(state::AbstractBFGSState{P,Float64,L}, obj::DeviceObjective, x::AbstractVector{Float64},
                              ctxs::Vector{Context}, ls::BackTracking=BackTracking(), tol=1e-8;
                              one_pass::Bool=false) where {P,L}
    length(x) == P || throw(DimensionMismatch("x has length $(length(x)), state is BFGSState{$P}"))
    h = model_handle(ctxs[1], obj)
    cls = Ref(CBackTracking(ls))
    ptrs = Ptr{Cvoid}[c.ptr for c in ctxs]
    xc = x isa Vector{Float64} ? x : Vector{Float64}(x)
    xo = Vector{Float64}(undef, P); go = Vector{Float64}(undef, P)
    ell = Ref{Float64}(NaN); iters = Ref{Int32}(0); fevals = Ref{Int32}(0); status = Ref{Int32}(0)
    ms = Ref{Float64}(0.0)
    GC.@preserve state obj xc ptrs xo go begin
        check(ccall((:qn_bfgs_optimize_rowsharded_f64, libqn), Int32,
                    (Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{CBackTracking}, Float64, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{Float64}, Ref{Int32}, Ref{Int32}, Ref{Int32},
                     Int32, Ptr{Float64}, UInt32, Int32, Ref{Float64}),
                    ptrs, length(ptrs), h, P, pointer(xc), cls, Float64(tol), 10_000,
                    xo, go, C_NULL, 0, ell, iters, fevals, status, 0, C_NULL,
                    one_pass ? QN_FLAG_ONE_PASS : UInt32(0), 0, ms))
        p = pointer(state)
        unsafe_copyto!(p, pointer(xo), P)                          # ref_x_old, offset 0
        unsafe_copyto!(p + L * sizeof(Float64), pointer(go), P)    # ref_∇_new, offset L
    end
    ell[]
end

end # module
