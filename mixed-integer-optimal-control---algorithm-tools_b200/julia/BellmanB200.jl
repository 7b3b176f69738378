# BellmanB200.jl -- drop-in replacement of the reference's trust-region subproblem solver.
#
# Load AFTER `include("multi-trust.jl")` (which includes HelpFunctions.jl into Main):
#
#     include("multi-trust.jl")
#     ENV["BELLMAN_B200_LIB"] = "/path/to/libbellman_b200.so"     # optional, see LIB below
#     include("/path/to/BellmanB200.jl")
#     main("fishing")                                               # every example runs unchanged
#
# It overwrites the two methods the reference defines in Main and calls only from TRM
# (multi-trust.jl:110,112,113):
#
#     bellman_TRM!(∇f, u_old, B, β, p, Δt, nu, U, Φ, iterator)      HelpFunctions.jl:20-83
#     eval_u_TRM!(u, u_old, U, Φ, B, nu)                             HelpFunctions.jl:98-124
#
# and forwards them through `ccall` to the C ABI of include/bellman_b200.h.  TRM, TRM_parameters, the
# AbstractObjective API, OptBundle and all examples stay byte-for-byte unchanged.
#
# State: one device plan per TRM run, keyed on objectid(U) (TRM allocates U and Φ once per run,
# multi-trust.jl:71-77, and passes the same objects to every call) and rebuilt whenever β, p, Δt, nu, the
# iterator or the sizes differ from what the plan was created for.  The device keeps its own packed
# tables; the caller's U and Φ are left untouched unless BELLMAN_B200_WRITEBACK=1 (parity/debugging),
# in which case they receive reference-shaped copies.
#
# NOTE: Julia is not available in the build container of this repository, so this file has been
# desk-checked only; tests/ drive exactly the same entry points through Python ctypes
# (mixed-integer-optimal-control---algorithm-tools_b200/api.py mirrors this file call for call).

module BellmanB200

const LIB = get(ENV, "BELLMAN_B200_LIB", joinpath(@__DIR__, "..", "libbellman_b200.so"))
const WRITEBACK = get(ENV, "BELLMAN_B200_WRITEBACK", "0") == "1"
const DEVICE = parse(Cint, get(ENV, "BELLMAN_B200_DEVICE", "0"))

const BB200_OK = Cint(0)
const BB200_ERR_INEXACT = Cint(3)
const BB200_ERR_STALE = Cint(4)

mutable struct Plan
    handle::Ptr{Cvoid}
    n::Int64
    M::Int64
    B::Int64
    K::Int64
    sig::UInt64                 # hash of everything baked into the device plan (see signature())
    u_cache::Matrix{Float64}    # trajectory for budget B produced by the last bellman! (one bb200_solve)
    cache_ok::Bool
end

const PLANS = Dict{UInt64,Plan}()   # objectid(U) => plan
const DEAD = UInt64[]               # keys whose U was garbage collected; freed from plan_for, never from GC context
const DEAD_LOCK = Base.Threads.SpinLock()

last_error() = unsafe_string(ccall((:bb200_last_error, LIB), Cstring, ()))

function check(rc::Cint)
    rc == BB200_OK && return nothing
    msg = last_error()
    # same exception types a user of the reference would see
    rc == BB200_ERR_INEXACT && throw(InexactError(:convert, Int64, msg))     # HelpFunctions.jl:37,57
    rc == BB200_ERR_STALE && throw(BoundsError(msg))                         # stale U cell, HelpFunctions.jl:116
    error("bellman_b200 (code $rc): $msg")
end

function destroy!(p::Plan)
    if p.handle != C_NULL
        ccall((:bb200_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.handle)
        p.handle = C_NULL
    end
    return nothing
end

# Everything that is fixed at plan creation.  The reference's bellman_TRM! is stateless, so a caller may reuse
# U/Φ with another β, p, Δt, nu or iterator (β-continuation, parameter sweeps); the plan is then rebuilt.
signature(n, M, B, β, p, Δt, nu, tuples) = hash((n, M, B, Float64(β), typeof(p), Float64(p), Float64(Δt), nu, tuples))

# Flattens the iterator exactly once per call (the reference re-runs the filtered generator
# K*n times per DP, AdmissibleIterators.jl:26-34) and evaluates the jump-cost table with the
# reference's own expression (HelpFunctions.jl:63-67) so that Julia's `^` decides every bit.
function make_plan(u_old, B, β, p, Δt, nu, tuples, sig)
    M, n = size(u_old)
    K = length(tuples)
    dims = Int64[length(nu[m]) for m = 1:M]
    level_values = Matrix{Int32}(undef, M, K)                    # column k = nu_k  -> C int32[K][M]
    grid_offset = Vector{Int64}(undef, K)
    for (k, l) in enumerate(tuples)
        off = 0; stride = 1
        for m = 1:M
            level_values[m, k] = nu[m][l[m]]
            off += (l[m] - 1) * stride
            stride *= dims[m]
        end
        grid_offset[k] = off
    end
    # jump_cost[j*K + l] (0-based, row-major) == Julia column-major cost[l, j]
    cost = Matrix{Float64}(undef, K, K)
    for (jj, j) in enumerate(tuples), (ll, l) in enumerate(tuples)
        temp_val_2 = 0.
        for m = 1:M
            temp_val_2 += abs(nu[m][j[m]] - nu[m][l[m]])^p      # HelpFunctions.jl:65
        end
        cost[ll, jj] = β * temp_val_2^(1 / p)                    # HelpFunctions.jl:67 (without temp_val_1)
    end
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bb200_plan_create, LIB), Cint,
                (Cint, Int64, Int32, Int32, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64},
                 Float64, Int32, UInt32, Ptr{Ptr{Cvoid}}),
                DEVICE, n, M, K, B, dims, level_values, grid_offset, cost, Float64(Δt), 1, 0, handle))
    plan = Plan(handle[], n, M, B, K, sig, Matrix{Float64}(undef, M, n), false)
    finalizer(destroy!, plan)
    return plan
end

# Frees the plans whose U has been collected.  The finalizer of U only records the key (it may run in the middle
# of any Dict operation); the Dict itself is touched here, from ordinary task context.
function mark_dead(key::UInt64)
    lock(DEAD_LOCK)
    try
        push!(DEAD, key)
    finally
        unlock(DEAD_LOCK)
    end
    return nothing
end

function reap!()
    lock(DEAD_LOCK)
    dead = try
        d = copy(DEAD)
        empty!(DEAD)
        d
    finally
        unlock(DEAD_LOCK)
    end
    for key in dead
        plan = pop!(PLANS, key, nothing)
        plan === nothing || destroy!(plan)
    end
    return nothing
end

function plan_for(U, u_old, B, β, p, Δt, nu, iterator)
    reap!()
    key = objectid(U)
    M, n = size(u_old)
    tuples = collect(iterator)                                   # iteration order == admissible order
    sig = signature(n, M, B, β, p, Δt, nu, tuples)
    plan = get(PLANS, key, nothing)
    if plan === nothing || plan.handle == C_NULL || plan.sig != sig
        fresh = plan === nothing
        fresh || destroy!(plan)
        plan = make_plan(u_old, B, β, p, Δt, nu, tuples, sig)
        PLANS[key] = plan
        # U lives exactly as long as the TRM run: drop the device tables once it is gone
        fresh && finalizer(_ -> mark_dead(key), U)
    end
    return plan
end

# The reference always calls eval_u_TRM!(u, u_old, U, Φ, B, nu) right after bellman_TRM! (multi-trust.jl:112-113),
# so the whole inner iteration -- H2D, DP, selection and backtrack for budget B, D2H -- goes out as ONE bb200_solve
# (a CUDA-graph replay, one synchronisation).  The trajectory is kept in the plan and handed out by the eval_u! that
# follows; smaller radii after a rejected step (multi-trust.jl:109-110) use bb200_select_and_backtrack.
function bellman!(∇f::Matrix{Float64}, u_old::Matrix{Float64}, B, β, p, Δt, nu, U, Φ, iterator)
    plan = plan_for(U, u_old, Int64(B), β, p, Δt, nu, iterator)
    plan.cache_ok = false
    # Julia M x n column-major == C double[n][M]; arrays are GC-rooted for the duration of the ccall
    rc = ccall((:bb200_solve, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
               plan.handle, ∇f, u_old, plan.B, plan.u_cache, C_NULL, C_NULL, C_NULL)
    if rc == BB200_OK
        plan.cache_ok = true
    elseif rc != BB200_ERR_STALE     # no feasible trajectory: the reference fails in eval_u_TRM!, not here
        check(rc)
    end
    if WRITEBACK
        check(ccall((:bb200_export_phi, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), plan.handle, 0, Φ))
        plan.n > 1 && check(ccall((:bb200_export_argmin, LIB), Cint,
                                  (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Int64),
                                  plan.handle, 0, 1, plan.n, U, 0))
    end
    return nothing
end

function eval_u!(u::Matrix{Float64}, u_old, U, Φ, B, nu)
    plan = get(PLANS, objectid(U), nothing)
    plan === nothing && error("eval_u_TRM! called before bellman_TRM! for this U")
    if Int64(B) == plan.B && plan.cache_ok
        copyto!(u, plan.u_cache)
        return nothing
    end
    check(ccall((:bb200_select_and_backtrack, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                plan.handle, Int64(B), u, C_NULL, C_NULL, C_NULL))
    return nothing
end

end # module

# ---- method overwrite in Main: same signatures as HelpFunctions.jl:20 and :98 --------------------
function bellman_TRM!(∇f, u_old, B, β, p, Δt, nu, U, Φ, iterator)
    BellmanB200.bellman!(∇f, u_old, B, β, p, Δt, nu, U, Φ, iterator)
end

function eval_u_TRM!(u, u_old, U, Φ, B, nu)
    BellmanB200.eval_u!(u, u_old, U, Φ, B, nu)
end
