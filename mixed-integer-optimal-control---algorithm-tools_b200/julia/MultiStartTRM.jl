# MultiStartTRM.jl -- S trust-region runs in lock-step on the batched B200 solver (SURVEY 8f N4).
#
# Load after multi-trust.jl and BellmanB200.jl:
#
#     include("multi-trust.jl"); include("BellmanB200.jl"); include("MultiStartTRM.jl")
#     objs = [LVMObj(nt = 1024) for s = 1:64]
#     Js = TRM_multistart(objs, TRM_parameters(β = 1e-4, Δ⁰ = 2, p = Inf))        # x0s default: rand_func per start
#     best = argmin(Js)                                                          # objs[best].x is the best control
#
# What it does.  The reference's TRM (multi-trust.jl:53-170) solves one trust-region subproblem per outer iteration
# (bellman_TRM! + eval_u_TRM!, :112-113) and after a rejected step only re-runs eval_u_TRM! with a halved radius on the
# SAME tables (:109-110).  A sweep over trial radii therefore costs one DP, and the starts are independent: per outer
# iteration all running starts share ONE `bb200_solve_batched` call that returns, for every start, the trajectories of
# the whole halving ladder floor(Δ⁰/2^(k-1)/Δt), k = 1..kmax; each start then walks its own inner loop (:105-159) over
# them.  Every DP result is exactly what the single-start loop computes, so each start's history (log table, final
# control, objective) equals an independent `TRM` run from the same x0.  The executable mirror of this file is
# `TRM_multistart` in oracle/trm_harness.py (tests/test_trm_history.py checks it against S independent runs, on the
# CPU with the oracle as the batched solver and on the GPU through bb200_solve_batched).
#
# NOTE: like BellmanB200.jl this file could not be executed where it was written (no Julia in that image).

module MultiStartB200

using ..BellmanB200: LIB, DEVICE, check, last_error, BB200_OK, BB200_ERR_INEXACT, BB200_ERR_STALE

const MAX_RADII = 16            # kMaxRadii of the library: selections per subproblem and call

# distinct trial budgets in the order the inner loop visits them, and for k = 1..kmax the index of its budget
function radius_ladder(Δ⁰, Δt, kmax)
    radii = Int64[]
    index = Int[]
    Δᵏ = Δ⁰
    for k = 1:kmax
        b = Int64(floor(Δᵏ / Δt))                       # multi-trust.jl:69 (k = 1) and :109 (after halving)
        (isempty(radii) || radii[end] != b) && push!(radii, b)
        push!(index, length(radii))
        Δᵏ = Δᵏ / 2
    end
    return radii, index
end

mutable struct BatchPlan
    handle::Ptr{Cvoid}
    batch::Int
end

function destroy!(p::BatchPlan)
    if p.handle != C_NULL
        ccall((:bb200_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.handle)
        p.handle = C_NULL
    end
    return nothing
end

# Same flattening as BellmanB200.make_plan, with `batch` resident slots; halves the batch while the device is short
# of memory (the argmin table of one slot is (n-1)(B+1)·roundup32(K) bytes).
function make_batch_plan(n, M, B, β, p, Δt, nu, iterator, S)
    tuples = collect(iterator)
    K = length(tuples)
    dims = Int64[length(nu[m]) for m = 1:M]
    level_values = Matrix{Int32}(undef, M, K)
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
    cost = Matrix{Float64}(undef, K, K)
    for (jj, j) in enumerate(tuples), (ll, l) in enumerate(tuples)
        temp_val_2 = 0.
        for m = 1:M
            temp_val_2 += abs(nu[m][j[m]] - nu[m][l[m]])^p      # HelpFunctions.jl:65
        end
        cost[ll, jj] = β * temp_val_2^(1 / p)                    # HelpFunctions.jl:67
    end
    batch = S
    while true
        handle = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:bb200_plan_create, LIB), Cint,
                   (Cint, Int64, Int32, Int32, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64},
                    Float64, Int32, UInt32, Ptr{Ptr{Cvoid}}),
                   DEVICE, n, M, K, B, dims, level_values, grid_offset, cost, Float64(Δt), batch, 0, handle)
        if rc == BB200_OK
            plan = BatchPlan(handle[], batch)
            finalizer(destroy!, plan)
            return plan
        end
        (rc == Cint(6) && batch > 1) || check(rc)                # BB200_ERR_NOMEM: try a smaller wave
        batch = cld(batch, 2)
    end
end

# One batched call: gradients and current controls of the A active starts in, trajectories of `radii` out.
#   df_all, u_old_all :: Array{Float64,3} (M, n, A)  == C double[A][n][M]
#   returns u_all :: Array{Float64,4} (M, n, R, A)   == C double[A][R][n][M],  status :: Matrix{Int32} (R, A)
function solve_batched(plan::BatchPlan, df_all::Array{Float64,3}, u_old_all::Array{Float64,3}, radii::Vector{Int64})
    M, n, A = size(df_all)
    R = length(radii)
    u_all = Array{Float64,4}(undef, M, n, R, A)
    status = Matrix{Int32}(undef, R, A)
    rc = ccall((:bb200_solve_batched, LIB), Cint,
               (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Int64}, Ptr{Int64}, Ptr{Int32}),
               plan.handle, A, df_all, u_old_all, R, radii, u_all, C_NULL, C_NULL, C_NULL, status)
    # per-entry failures (a start whose u_old is not integer valued, a selection without a feasible trajectory) are
    # reported through `status` and raised when -- and only if -- that start reaches that radius, like the reference
    (rc == BB200_OK || rc == BB200_ERR_INEXACT || rc == BB200_ERR_STALE) || check(rc)
    return u_all, status
end

end # module

@doc raw"""
    TRM_multistart(objs, par = TRM_parameters(); x0s = [rand_func(obj) for obj in objs])

Runs `TRM` (multi-trust.jl:53) for every objective in `objs` (same problem class, sizes and `𝓥`; each with its own
start `x0s[s]`) in lock-step on the GPU and returns the vector of final objective values `J + β·TV_p(u, p)`;
`objs[s].x` holds the final control of start `s`, exactly as after `TRM(objs[s], par; x0 = x0s[s])`.
"""
function TRM_multistart(objs::Vector, par::TRM_parameters = TRM_parameters(); x0s = [rand_func(obj) for obj in objs])
    S = length(objs)
    n = objs[1].nt; M = objs[1].nx; Δt = objs[1].tau
    nu = objs[1].𝓥
    @unpack β, Δ⁰, σ, p, kmax, maxiter = par
    B = Int64(floor(Δ⁰ / Δt))                                         # multi-trust.jl:69
    radii, kindex = MultiStartB200.radius_ladder(Δ⁰, Δt, kmax)
    plan = MultiStartB200.make_batch_plan(n, M, B, β, p, Δt, nu, objs[1].iterator, S)

    us = [obj.x for obj in objs]                                      # u aliases obj.x (:65)
    for s = 1:S
        us[s] .= x0s[s]
    end
    u_olds = [copy(u) for u in us]
    J_olds = [eval_f!(obj) for obj in objs]                           # :83
    Js = fill(Inf, S)
    stop = falses(S)
    iter = 1
    while !all(stop) && (iter ≤ maxiter)                              # :92
        act = findall(.!stop)
        A = length(act)
        TV_olds = [TV_p(us[s], p) for s in act]                       # :99
        df_all = Array{Float64,3}(undef, M, n, A)
        uo_all = Array{Float64,3}(undef, M, n, A)
        for (a, s) in enumerate(act)
            eval_df!(objs[s])                                         # :102
            df_all[:, :, a] .= objs[s].df
            uo_all[:, :, a] .= u_olds[s]
        end
        first_radii = radii[1:min(end, MultiStartB200.MAX_RADII)]
        u_all, status = MultiStartB200.solve_batched(plan, df_all, uo_all, first_radii)
        for (a, s) in enumerate(act)
            obj = objs[s]; u = us[s]; u_old = u_olds[s]; ∇f = obj.df
            Δᵏ = Δ⁰; k = 1; ared = 0.; pred = 1.
            TV_old = TV_olds[a]
            lo = 1                                                    # radii[lo : lo+MAX_RADII-1] are at hand
            u_row = @view u_all[:, :, :, a]; st_row = @view status[:, a]
            while (ared < σ * pred) && (k ≤ kmax)                     # :105
                r = kindex[k]
                if !(lo ≤ r < lo + MultiStartB200.MAX_RADII)          # deeper ladder (B ≥ 2^16 only): one more DP
                    lo = r
                    more = radii[lo:min(end, lo + MultiStartB200.MAX_RADII - 1)]
                    u_more, st_more = MultiStartB200.solve_batched(plan, df_all[:, :, a:a], uo_all[:, :, a:a], more)
                    u_row = @view u_more[:, :, :, 1]; st_row = @view st_more[:, 1]
                end
                code = st_row[r - lo + 1]
                code == MultiStartB200.BB200_ERR_INEXACT && throw(InexactError(:convert, Int64, "u_old of start $s"))
                code == MultiStartB200.BB200_ERR_STALE && throw(BoundsError("start $s: no feasible trajectory"))
                u .= @view u_row[:, :, r - lo + 1]                    # what eval_u_TRM!(u, u_old, U, Φ, B_new, nu) yields
                int_val = 0.                                          # :117-121
                for j = 1:n
                    int_val += ∇f[:, j]' * (u_old[:, j] - u[:, j])
                end
                int_val *= Δt
                TV_new = TV_p(u, p)
                J_new = eval_f!(obj)
                pred = int_val + β * (TV_old - TV_new)
                ared = J_olds[s] - J_new + β * (TV_old - TV_new)
                if pred ≤ 0                                           # :129-137
                    Js[s] = J_olds[s]
                    stop[s] = true
                    break
                elseif ared < σ * pred                                # :139-145
                    Δᵏ = Δᵏ / 2
                else                                                  # :147-156
                    u_old .= u
                    J_olds[s] = J_new
                    TV_old = TV_new
                    Js[s] = J_new
                end
                k += 1
            end
        end
        iter += 1
    end
    MultiStartB200.destroy!(plan)
    for obj in objs
        eval_df!(obj)                                                 # :165 (final derivative, for plotting)
    end
    return [Js[s] + β * TV_p(us[s], p) for s = 1:S]
end
