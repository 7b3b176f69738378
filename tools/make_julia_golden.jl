# make_julia_golden.jl -- pins the CPU oracle (and through it the CUDA path) to the REAL reference.
#
# Runs the reference's own, unmodified `bellman_TRM!` and `eval_u_TRM!` (HelpFunctions.jl:20-83, :98-124) and its own
# iterators (julia_opt/AdmissibleIterators.jl) under Julia on every case in tests/golden/julia_in/ (the two known-answer
# tests and all committed golden instances; written by tests/golden/make_julia_inputs.py) and dumps the value table Φ,
# the argmin table U and, for every trial radius, the selected cell and the control trajectory u into
# tests/golden/julia_out/<case>.txt.  tests/test_julia_golden.py then asserts oracle == dump, bit for bit; without the
# dump that test reports "parity unpinned".
#
#     julia --check-bounds=yes tools/make_julia_golden.jl /path/to/mixed-integer-optimal-control---algorithm-tools
#     python -m pytest tests/test_julia_golden.py -q
#
# `--check-bounds=yes` overrides the reference's `@inbounds`: a backtrack that steps onto a cell the DP never wrote
# (U still zero there, SURVEY F10) then raises a BoundsError, which is recorded, instead of reading out of bounds.
#
# Needs only Julia's standard library: HelpFunctions.jl begins with `using StatsBase, Random, DSP, Plots` (start-point
# generation and plotting), so the two functions are evaluated from the reference's own source text instead of
# `include`-ing the whole file.  This script could NOT be executed where it was written (no Julia in that image).
# Written for Julia 1.10 (Manifest.toml:3 of the reference pins julia_version = "1.10.0").

using LinearAlgebra   # norm, HelpFunctions.jl:121

length(ARGS) >= 1 || error("usage: julia --check-bounds=yes tools/make_julia_golden.jl <reference checkout>")
const REFDIR = ARGS[1]
const REPO = normpath(joinpath(@__DIR__, ".."))
const INDIR = joinpath(REPO, "tests", "golden", "julia_in")
const OUTDIR = joinpath(REPO, "tests", "golden", "julia_out")
mkpath(OUTDIR)

include(joinpath(REFDIR, "julia_opt", "AdmissibleIterators.jl"))

# Evaluates `function <name>( ... end` (the closing `end` is the first one in column 1) from the reference's file.
function include_function(path, name)
    lines = readlines(path)
    start = findfirst(l -> startswith(l, "function " * name * "("), lines)
    start === nothing && error("function $name not found in $path")
    stop = findnext(l -> rstrip(l) == "end", lines, start)
    stop === nothing && error("end of function $name not found in $path")
    Base.include_string(Main, join(lines[start:stop], "\n"), path)
    return (start, stop)
end

const HF = joinpath(REFDIR, "HelpFunctions.jl")
const SPAN_BELLMAN = include_function(HF, "bellman_TRM!")
const SPAN_EVAL = include_function(HF, "eval_u_TRM!")

hexf(s) = reinterpret(Float64, parse(UInt64, s; base = 16))
hexvec(s) = Float64[hexf(w) for w in split(s)]
tohex(x::Float64) = string(reinterpret(UInt64, x); base = 16, pad = 16)

function read_case(path)
    d = Dict{String,String}()
    for line in eachline(path)
        isempty(strip(line)) && continue
        kv = split(line, ' '; limit = 2)
        d[String(kv[1])] = length(kv) > 1 ? String(kv[2]) : ""
    end
    return d
end

function run_case(path)
    d = read_case(path)
    name = d["name"]
    M = parse(Int64, d["M"]); n = parse(Int64, d["n"]); B = parse(Int64, d["B"])
    Δt = hexf(d["dt"]); β = hexf(d["beta"])
    p = d["p"] == "Inf" ? Inf : parse(Int64, d["p"])            # Float64 Inf or Int64, as main() / TRM_parameters pass it
    nu = Vector{Int64}[parse.(Int64, split(strip(part))) for part in split(d["nu"], ';')]
    itw = split(d["iterator"])
    make_iterator() = itw[1] == "product" ? product_iterator(nu) :
                      bounded_sum_iterator(nu, parse(Int64, itw[2]), parse(Int64, itw[3]))
    radii = parse.(Int64, split(d["radii"]))
    ∇f = reshape(hexvec(d["df"]), M, n)
    u_old = reshape(hexvec(d["u_old"]), M, n)

    # table allocation exactly as multi-trust.jl:69-77
    dims = zeros(Int64, M + 3)
    dims[1] = M
    dims[2] = B + 1
    dims[3:M+2] .= [length(nu[m]) for m = 1:M]
    dims[M+3] = n - 1
    U = zeros(Int64, dims...)
    Φ = zeros(Float64, dims[2:M+2]..., 2)

    bellman_TRM!(∇f, u_old, B, β, p, Δt, nu, U, Φ, make_iterator())

    open(joinpath(OUTDIR, name * ".txt"), "w") do io
        println(io, "name ", name)
        println(io, "julia_version ", VERSION)
        println(io, "reference_lines bellman_TRM! ", SPAN_BELLMAN[1], "-", SPAN_BELLMAN[2], " eval_u_TRM! ", SPAN_EVAL[1], "-", SPAN_EVAL[2])
        println(io, "Phi_dims ", join(size(Φ), " "))
        println(io, "Phi ", join((tohex(x) for x in vec(Φ)), " "))              # column-major
        println(io, "U_dims ", join(size(U), " "))
        println(io, "U ", join(vec(U), " "))
        for Bn in radii
            u = zeros(Float64, M, n)
            try
                Inds = CartesianIndices(size(Φ)[2:end-1])
                index = argmin(@view Φ[1:Bn+1, Inds, 1])                           # HelpFunctions.jl:106
                eval_u_TRM!(u, u_old, U, Φ, Bn, nu)
                println(io, "radius ", Bn, " ok index ", join(Tuple(index), " "), " phi ", tohex(Φ[index, 1]),
                        " u ", join((tohex(x) for x in vec(u)), " "))
            catch err
                println(io, "radius ", Bn, " error ", typeof(err))
            end
        end
    end
    println("wrote ", name)
end

for f in sort(readdir(INDIR))
    endswith(f, ".txt") && run_case(joinpath(INDIR, f))
end
