# CIAOAlgorithmsCUDA.jl — drop-in Julia front-end over libciao_cuda (B200, sm_100a).
#
# Same solver structs, keyword arguments, `iterate`/`solution`/`iterator` protocol and
# count semantics as kul-optec/CIAOAlgorithms.jl v0.1.1; the bodies of `Base.iterate`
# are `ccall`s into the C ABI of include/ciao_cuda.h instead of Julia loops over
# ProximalOperators calls.  Index draws stay here, through the reference's own RNG
# call sites (rand / StatsBase.sample / randperm on the global RNG), so the sampled
# sequence is the reference's bit for bit.
#
# NOTE: Julia is not installed in the build image, so this file has not been executed
# there; the Python twin (ciaoalgorithms.jl_b200/solvers.py) exercises the same ABI
# calls in the same order and is the tested path.  Field names of ProximalOperators
# objects (A, b, lambda; f, L; y, mu; Q, q; ind, lambda; fs; lb, ub) follow v0.14.
module CIAOAlgorithmsCUDA

using LinearAlgebra, Random, Printf
using ProximalOperators
using StatsBase: sample
using Base.Iterators: take

export solution, SVRG, SAGA, SAG, Finito, Proshi, iterator

const libciao = get(ENV, "LIBCIAO_CUDA", "libciao_cuda")
const Maybe{T} = Union{T,Nothing}
const CIAO_VEC_Z, CIAO_VEC_Z_FULL, CIAO_VEC_W, CIAO_VEC_AV = Cint(0), Cint(1), Cint(2), Cint(3)

struct CiaoError <: Exception
    code::Cint
    msg::String
end
function check(code::Cint)
    code == 0 && return nothing
    throw(CiaoError(code, unsafe_string(ccall((:ciao_last_error, libciao), Cstring, ()))))
end

mutable struct Ctx
    h::Ptr{Cvoid}
    N::Int
    d::Int
    function Ctx(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ciao_create, libciao), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device))
        c = new(r[], 0, 0)
        finalizer(x -> ccall((:ciao_destroy, libciao), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end

# Complex-typed problems (test_lasso.jl:3, T = ComplexF32/ComplexF64) are accepted when their data is real — what the reference's
# tests build (C = rand(R, N, n), :19): complex arithmetic on such data never leaves the real axis, so the engine computes on the
# real parts and the solution comes back in the caller's element type.  Non-zero imaginary parts are refused HERE: the library runs
# genuinely complex data as realified M = 2 blocks with CIAO_REG_NORML1_PAIRS (ciao_set_row_blocks; the Python twin does it,
# operators.pack_F), but this shim's state vectors are not yet laid out as (re, im) pairs.
realdata(a) = eltype(a) <: Complex ?
    (all(iszero, imag.(a)) ? real.(a) : error("complex data with non-zero imaginary parts is outside the engine's scope")) : a

# Multi-GPU hosts (one process per GPU; INTEGRATION.md): declare, before the rows are set, that this process holds the blocks of
# `block` rows number rank, rank + world, … of F (0-based rank).  Static minibatches of a multiple of block·world rows are then
# spread over all GPUs (the ranks' sums meet inside the persistent minibatch kernel).  set_problem! below is the single-GPU path
# (row0 = 0, all of F); a multi-GPU host packs interleaved_rows(…) of F itself and calls ciao_set_rows with N_total = N,
# row0 = rank·block, n_rows = length(interleaved_rows(…)).
set_row_interleave!(c::Ctx, block::Integer, rank::Integer, world::Integer) =
    check(ccall((:ciao_set_row_interleave, libciao), Cint, (Ptr{Cvoid}, Int64, Cint, Cint), c.h, block, rank, world))
interleaved_rows(N::Integer, block::Integer, rank::Integer, world::Integer) = [i for i in 1:N if ((i - 1) ÷ block) % world == rank]

# ---- F / g recognition (replaces dynamic dispatch on F::Array{Tf}, SVRG_basic.jl:2) --------------
function set_problem!(c::Ctx, F, g, N::Int, x0 = nothing)
    F === nothing && (F = fill(ProximalOperators.Zero(), (N,)))   # SVRG.jl:58, SAGA.jl:55, Finito.jl:78, ProShI.jl:54
    f1 = F[1]
    if all(f -> f isa ProximalOperators.Zero, F)         # ∇f_i ≡ 0: least-squares rows with a_i = 0, b_i = 0, λ_i = 0
        x0 === nothing && error("an all-Zero F needs x0 for the dimension")
        d = length(x0)
        A = zeros(Float64, d, N); b = zeros(Float64, N); s = zeros(Float64, N)
        GC.@preserve A b s check(ccall((:ciao_set_rows, libciao), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Float64),
            c.h, 0, N, 0, N, d, A, d, b, s, 0.0))
        c.N, c.d = N, d
    elseif f1 isa ProximalOperators.LeastSquares            # test_lasso.jl:53-54: LeastSquares(A[i:i,:], b[i:i], N); also m×d blocks
        m, d = size(f1.A)
        all(f -> size(f.A) == (m, d) && length(f.b) == m, F) || error("engine covers LeastSquares terms of one common shape m×d")
        A = Matrix{Float64}(undef, d, N * m)             # column-major d×(N·m) == row-major (N·m)×d: component i holds rows (i−1)m+1 … im
        b = Vector{Float64}(undef, N * m); s = Vector{Float64}(undef, N)
        for i = 1:N
            Ai = realdata(F[i].A); bi = realdata(F[i].b)
            for r = 1:m
                copyto!(view(A, :, (i - 1) * m + r), view(Ai, r, :)); b[(i - 1) * m + r] = bi[r]
            end
            s[i] = F[i].lambda
        end
        if m == 1
            GC.@preserve A b s check(ccall((:ciao_set_rows, libciao), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Float64),
                c.h, 0, N, 0, N, d, A, d, b, s, 0.0))
        else                                             # m×d blocks: the general block kernel (csrc/blockseq.cu)
            GC.@preserve A b s check(ccall((:ciao_set_row_blocks, libciao), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Float64),
                c.h, 0, N, m, d, A, d, b, s, 0.0))
        end
        c.N, c.d = N, d
    elseif f1 isa ProximalOperators.Precompose && f1.f isa ProximalOperators.LogisticLoss   # test_logistic_l1.jl:36
        d = size(f1.L, 2)
        A = Matrix{Float64}(undef, d, N); y = Vector{Float64}(undef, N); mu = Vector{Float64}(undef, N)
        for i = 1:N
            copyto!(view(A, :, i), vec(F[i].L)); y[i] = F[i].f.y[1]; mu[i] = F[i].f.mu
        end
        GC.@preserve A y mu check(ccall((:ciao_set_rows, libciao), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Float64),
            c.h, 1, N, 0, N, d, A, d, y, mu, 0.0))
        c.N, c.d = N, d
    elseif f1 isa ProximalOperators.Sum                  # test_sharing.jl:18-22: Sum(Quadratic(diagm(d_i), q), SqrDistL2(IndBox, η))
        quad(f) = first(x for x in f.fs if x isa ProximalOperators.Quadratic)
        dist(f) = first(x for x in f.fs if x isa ProximalOperators.SqrDistL2)
        n = length(quad(f1).q)
        Qd = Matrix{Float64}(undef, n, N); ql = Matrix{Float64}(undef, n, N)
        for i = 1:N
            Q = quad(F[i]).Q
            isdiag(Q) || error("engine covers diagonal Quadratic terms")
            Qd[:, i] .= diag(Q); ql[:, i] .= quad(F[i]).q
        end
        bx = dist(f1).ind
        GC.@preserve Qd ql check(ccall((:ciao_set_blocks, libciao), Cint,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64, Float64),
            c.h, N, n, Qd, n, ql, n, Float64(bx.lb), Float64(bx.ub), Float64(dist(f1).lambda)))
        c.N, c.d = N, n
    else
        error("f_i of type $(typeof(f1)) is outside the engine's scope (no CPU fallback)")
    end
    if g isa ProximalOperators.NormL1
        p = Float64[g.lambda]; GC.@preserve p check(ccall((:ciao_set_reg, libciao), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), c.h, 1, p, 1))
    elseif g isa ProximalOperators.IndBox
        lo = g.lb isa Number ? fill(Float64(g.lb), c.d) : Vector{Float64}(g.lb)
        hi = g.ub isa Number ? fill(Float64(g.ub), c.d) : Vector{Float64}(g.ub)
        p = vcat(lo, hi); GC.@preserve p check(ccall((:ciao_set_reg, libciao), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), c.h, 2, p, length(p)))
    elseif g isa ProximalOperators.Zero
        check(ccall((:ciao_set_reg, libciao), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), c.h, 0, C_NULL, 0))
    else
        error("g of type $(typeof(g)) is outside the engine's scope (Zero, NormL1, IndBox)")
    end
    return c
end

# D2H into the fp64 staging buffer, then into the persistent array typed like x0 (the reference's states are typed by x0:
# `eltype(x) == T`, test_lasso.jl:75); for Float64 problems the two are the same array and nothing is copied.
function getvec!(c::Ctx, which::Cint, out::AbstractVector, buf::Vector{Float64})
    GC.@preserve buf check(ccall((:ciao_get_vec, libciao), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64), c.h, which, buf, length(buf)))
    out === buf || copyto!(out, buf)
    return out
end
# the persistent solution array (eltype of x0) and its fp64 staging buffer
function outvec(x0::AbstractArray)
    out = zeros(eltype(x0), length(x0))          # the caller's element type (Float32, Float64, or complex with real data)
    return out, (out isa Vector{Float64} ? out : zeros(Float64, length(x0)))
end

# ================================ SVRG / SVRG++ (SVRG.jl, SVRG_basic.jl) ================================
struct SVRG{R<:Real}
    γ::Maybe{R}; maxit::Int; verbose::Bool; freq::Int; m::Maybe{Int}; plus::Bool
    function SVRG{R}(; γ::Maybe{R} = nothing, maxit::Int = 10000, verbose::Bool = false, freq::Int = 1000,
                     m::Maybe{Int} = nothing, plus::Bool = false) where {R}
        @assert γ === nothing || γ > 0
        @assert maxit > 0
        @assert freq > 0
        new(γ, maxit, verbose, freq, m, plus)
    end
end
SVRG(::Type{R}; kwargs...) where {R} = SVRG{R}(; kwargs...)
SVRG(; kwargs...) = SVRG(Float64; kwargs...)

struct SVRG_basic_iterable{R,Tx,Tf,Tg}
    F::Tf; g::Tg; x0::Tx; N::Int; L; μ; γ::Maybe{R}; m::Maybe{Int}; plus::Bool
end
mutable struct SVRG_basic_state{R}
    ctx::Ctx; γ::R; m::Int; z_full::Vector; buf::Vector{Float64}; ind::Vector{Int}
end

function Base.iterate(iter::SVRG_basic_iterable{R}) where {R}
    N = iter.N
    m = iter.m === nothing ? N : iter.m
    if iter.γ === nothing
        if iter.plus
            @warn "provide a stepsize γ"; return nothing
        elseif iter.L === nothing || iter.μ === nothing
            @warn "smoothness or convexity parameter absent"; return nothing
        end
        L_M = maximum(iter.L); μ_M = maximum(iter.μ); γ = 1 / (10 * L_M)
        rho = (1 + 4 * L_M * γ^2 * μ_M * (N + 1)) / (μ_M * γ * N * (1 - 4L_M * γ))
        rho >= 1 && @warn "convergence condition violated...provide a stepsize!"
    else
        γ = iter.γ
    end
    c = set_problem!(Ctx(), iter.F, iter.g, N, iter.x0)
    x0 = Vector{Float64}(realdata(iter.x0))
    GC.@preserve x0 check(ccall((:ciao_svrg_init, libciao), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint), c.h, x0, γ, iter.plus))
    state = SVRG_basic_state{R}(c, γ, m, outvec(iter.x0)..., collect(1:N))
    return state, state
end

function Base.iterate(iter::SVRG_basic_iterable{R}, state::SVRG_basic_state{R}) where {R}
    idx = rand(state.ind, state.m)                     # SVRG_basic.jl:73 — the reference's own RNG call
    GC.@preserve idx check(ccall((:ciao_svrg_epoch, libciao), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), state.ctx.h, idx, length(idx)))
    iter.plus && (state.m *= 2)                        # :93
    return state, state
end
solution(state::SVRG_basic_state) = getvec!(state.ctx, CIAO_VEC_Z_FULL, state.z_full, state.buf)   # same array every call (===)

function (solver::SVRG{R})(x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, μ = nothing, N = N) where {R}
    m = solver.m === nothing ? N : solver.m
    maxit = solver.maxit
    if solver.plus && solver.maxit > 25
        maxit = 25
        @warn "exponential number of inner updates...reverted to 25 maximum iterations"
    end
    iter = SVRG_basic_iterable{R,typeof(x0),typeof(F),typeof(g)}(F, g, x0, N, L, μ, solver.γ, m, solver.plus)
    return drive(solver, iter, maxit, s -> s.γ)
end
iterator(solver::SVRG{R}, x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, μ = nothing, N = N) where {R} =
    SVRG_basic_iterable{R,typeof(x0),typeof(F),typeof(g)}(F, g, x0, N, L, μ, solver.γ, solver.m === nothing ? N : solver.m, solver.plus)

# ================================ SAGA / SAG (SAGA.jl, SAGA_basic.jl) ================================
struct SAGA{R<:Real}
    γ::Maybe{R}; maxit::Int; verbose::Bool; freq::Int; SAG_flag::Bool
    function SAGA{R}(; γ::Maybe{R} = nothing, maxit::Int = 10000, verbose::Bool = false, freq::Int = 1000, SAG_flag::Bool = false) where {R}
        @assert γ === nothing || γ > 0
        @assert maxit > 0
        @assert freq > 0
        new(γ, maxit, verbose, freq, SAG_flag)
    end
end
SAGA(::Type{R}; kwargs...) where {R} = SAGA{R}(; kwargs...)
SAGA(; kwargs...) = SAGA(Float64; kwargs...)
SAG(::Type{R}; kwargs...) where {R} = SAGA{R}(; kwargs..., SAG_flag = true)
SAG(; kwargs...) = SAG(Float64; kwargs...)

struct SAGA_basic_iterable{R,Tx,Tf,Tg}
    F::Tf; g::Tg; x0::Tx; N::Int; L; γ::Maybe{R}; SAG::Bool
end
mutable struct SAGA_basic_state{R}
    ctx::Ctx; γ::R; z::Vector; buf::Vector{Float64}; ind::Int
end
function Base.iterate(iter::SAGA_basic_iterable{R}) where {R}
    if iter.γ === nothing
        iter.L === nothing && (@warn "smoothness parameter absent"; return nothing)
        L_M = maximum(iter.L)
        γ = iter.SAG ? 1 / (16 * L_M) : 1 / (3 * L_M)
    else
        γ = iter.γ
    end
    c = set_problem!(Ctx(), iter.F, iter.g, iter.N, iter.x0)
    x0 = Vector{Float64}(realdata(iter.x0))
    GC.@preserve x0 check(ccall((:ciao_saga_init, libciao), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint), c.h, x0, γ, iter.SAG))
    state = SAGA_basic_state{R}(c, γ, outvec(iter.x0)..., 1)
    return state, state
end
# k reference iterations fused into one persistent kernel; the k draws are k calls of rand(1:N) (SAGA_basic.jl:55)
function steps!(iter::SAGA_basic_iterable, state::SAGA_basic_state, k::Int)
    idx = Int64[rand(1:iter.N) for _ = 1:k]
    state.ind = idx[end]
    GC.@preserve idx check(ccall((:ciao_saga_steps, libciao), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), state.ctx.h, idx, k))
    return state
end
Base.iterate(iter::SAGA_basic_iterable{R}, state::SAGA_basic_state{R}) where {R} = (steps!(iter, state, 1); (state, state))
solution(state::SAGA_basic_state) = getvec!(state.ctx, CIAO_VEC_Z, state.z, state.buf)

(solver::SAGA{R})(x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, N = N) where {R} =
    drive(solver, SAGA_basic_iterable{R,typeof(x0),typeof(F),typeof(g)}(F, g, x0, N, L, solver.γ, solver.SAG_flag), solver.maxit, s -> s.γ)
iterator(solver::SAGA{R}, x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, N = N) where {R} =
    SAGA_basic_iterable{R,typeof(x0),typeof(F),typeof(g)}(F, g, x0, N, L, solver.γ, solver.SAG_flag)

# ============================ Finito / MISO / DIAG, LFinito, ProShI ============================
struct Finito{R<:Real}
    γ::Maybe{Union{Array{R},R}}; sweeping::Int8; LFinito::Bool; adaptive::Bool; minibatch::Tuple{Bool,Int}
    maxit::Int; verbose::Bool; freq::Int; α::R; tol::R; tol_b::R
    function Finito{R}(; γ::Maybe{Union{Array{R},R}} = nothing, sweeping = 1, LFinito::Bool = false, adaptive::Bool = false,
                       minibatch::Tuple{Bool,Int} = (false, 1), maxit::Int = 10000, verbose::Bool = false, freq::Int = 10000,
                       α::R = R(0.999), tol::R = R(1e-8), tol_b::R = R(1e-9)) where {R}
        @assert γ === nothing || minimum(γ) > 0
        @assert maxit > 0 && tol > 0 && tol_b > 0 && freq > 0
        new(γ, sweeping, LFinito, adaptive, minibatch, maxit, verbose, freq, α, tol, tol_b)
    end
end
Finito(::Type{R}; kwargs...) where {R} = Finito{R}(; kwargs...)
Finito(; kwargs...) = Finito(Float64; kwargs...)

struct Proshi{R<:Real}
    γ::Maybe{Union{Array{R},R}}; sweeping::Int8; minibatch::Tuple{Bool,Int}; maxit::Int; verbose::Bool; freq::Int; α::R
    function Proshi{R}(; γ::Maybe{Union{Array{R},R}} = nothing, sweeping = 1, minibatch::Tuple{Bool,Int} = (false, 1),
                       maxit::Int = 10000, verbose::Bool = false, freq::Int = 10000, α::R = R(0.999)) where {R}
        @assert γ === nothing || minimum(γ) > 0
        @assert maxit > 0 && freq > 0
        new(γ, sweeping, minibatch, maxit, verbose, freq, α)
    end
end
Proshi(::Type{R}; kwargs...) where {R} = Proshi{R}(; kwargs...)
Proshi(; kwargs...) = Proshi(Float64; kwargs...)

# kind: :finito | :lfinito | :proshi
struct Table_iterable{R,Tx,Tf,Tg}
    kind::Symbol; F::Tf; g::Tg; x0::Tx; N::Int; L; γ; sweeping::Int8; batch::Int; α::R
end
mutable struct Table_state{R}
    kind::Symbol; ctx::Ctx; γ::Vector{R}; hat_γ::R; z::Vector; buf::Vector{Float64}
    s::Vector{<:Vector}; sbuf::Matrix{Float64}    # ProShI: the table x_i as the reference holds it (a vector of vectors) + fp64 staging
    d::Int; idxr::Int; idx::Int; inds::Vector{Int}
end

function stepsizes(iter::Table_iterable{R}) where {R}      # Finito_basic.jl:61-74
    N = iter.N
    if iter.γ === nothing
        iter.L === nothing && (@warn "--> smoothness parameter absent"; return nothing)
        return iter.L isa Real ? fill(iter.α * R(N) / iter.L, N) : R[iter.α * R(N) / iter.L[i] for i = 1:N]
    end
    return iter.γ isa Real ? fill(R(iter.γ), N) : Vector{R}(iter.γ)
end

function Base.iterate(iter::Table_iterable{R}) where {R}
    γ = stepsizes(iter)
    γ === nothing && return nothing
    N = iter.N
    hat_γ = iter.kind == :proshi ? sum(γ) : 1 / sum(1 ./ γ)           # ProShI_basic.jl:82 / Finito_basic.jl:82
    c = set_problem!(Ctx(), iter.F, iter.g, N, iter.x0)
    x0 = Vector{Float64}(realdata(iter.x0))
    γ64 = Vector{Float64}(γ)                                             # the engine computes in fp64 (R may be Float32)
    GC.@preserve x0 γ64 begin
        if iter.kind == :finito
            check(ccall((:ciao_finito_init, libciao), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), c.h, x0, γ64, hat_γ))
        elseif iter.kind == :lfinito
            check(ccall((:ciao_lfinito_init, libciao), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), c.h, x0, γ64, hat_γ))
        else
            check(ccall((:ciao_proshi_init, libciao), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), c.h, x0, γ64, hat_γ))
        end
    end
    d = cld(N, iter.batch)
    Te = real(eltype(iter.x0))
    s = iter.kind == :proshi ? [zeros(Te, length(x0)) for _ = 1:N] : Vector{Te}[]
    sbuf = iter.kind == :proshi ? Matrix{Float64}(undef, length(x0), N) : Matrix{Float64}(undef, 0, 0)
    state = Table_state{R}(iter.kind, c, γ, hat_γ, outvec(iter.x0)..., s, sbuf, d, 1, 0, collect(1:d))
    return state, state
end

batch_rows(iter::Table_iterable, j::Int) = collect(iter.batch*(j-1)+1:min(iter.batch * j, iter.N))   # Finito_basic.jl:52-57

function next_batch!(iter::Table_iterable, state::Table_state)          # Finito_basic.jl:96-108, ProShI_basic.jl:97-109
    if iter.sweeping == 1
        return sample(1:iter.N, iter.batch, replace = false)
    elseif iter.sweeping == 2
        state.idxr = mod(state.idxr, state.d) + 1
    else
        if state.idx == state.d
            state.inds = randperm(state.d); state.idx = 1
        else
            state.idx += 1
        end
        state.idxr = state.inds[state.idx]
    end
    return batch_rows(iter, state.idxr)
end

function steps!(iter::Table_iterable, state::Table_state, k::Int)
    if iter.kind == :lfinito
        for _ = 1:k
            iter.sweeping == 3 && (state.inds = randperm(state.d))       # Finito_LFinito.jl:89
            ord = Vector{Int64}(state.inds)
            GC.@preserve ord check(ccall((:ciao_lfinito_outer, libciao), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64), state.ctx.h, ord, length(ord), iter.batch))
        end
        return state
    end
    idx = Int64[]; ptr = Int64[0]
    for _ = 1:k
        append!(idx, next_batch!(iter, state)); push!(ptr, length(idx))
    end
    GC.@preserve idx ptr begin
        if iter.kind == :finito
            check(ccall((:ciao_finito_steps, libciao), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64), state.ctx.h, idx, ptr, k))
        else
            check(ccall((:ciao_proshi_steps, libciao), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64), state.ctx.h, idx, ptr, k))
        end
    end
    return state
end
Base.iterate(iter::Table_iterable{R}, state::Table_state{R}) where {R} = (steps!(iter, state, 1); (state, state))

function solution(state::Table_state)
    if state.kind == :proshi                                             # ProShI_basic.jl:127-132 — mutates the table on every call
        sbuf = state.sbuf                                                # n×N column-major: column i is x_i
        GC.@preserve sbuf check(ccall((:ciao_proshi_solution, libciao), Cint, (Ptr{Cvoid}, Ptr{Float64}), state.ctx.h, sbuf))
        for i = 1:length(state.s)
            copyto!(state.s[i], view(sbuf, :, i))
        end
        return state.s                                                   # Vector of x_i, as upstream (test_sharing.jl:42-43)
    end
    return getvec!(state.ctx, CIAO_VEC_Z, state.z, state.buf)             # Finito_basic.jl:123, Finito_LFinito.jl:105
end

# ---- adaptive Finito (Finito_adaptive.jl): linesearch on γ_i inside the persistent kernel ------------------------------
struct FINITO_adaptive_iterable{R,Tx,Tf,Tg}
    F::Tf; g::Tg; x0::Tx; N::Int; L; tol::R; tol_b::R; sweeping::Int8; α::R
end
mutable struct FINITO_adaptive_state{R}
    ctx::Ctx; z::Vector; buf::Vector{Float64}; γ::Vector{Float64}; hat_γ::R
    ind::Vector{Int}; idx::Int; idxr::Int            # Finito_adaptive.jl:53-55
end

function refresh!(state::FINITO_adaptive_state)       # γ and hat_γ change on the device during the linesearch
    hg = Ref{Float64}(0); nb = Ref{Int64}(0); γ = state.γ
    GC.@preserve γ check(ccall((:ciao_finito_adaptive_get, libciao), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Int64}), state.ctx.h, γ, C_NULL, C_NULL, hg, nb))
    state.hat_γ = hg[]
    return state
end

function Base.iterate(iter::FINITO_adaptive_iterable{R}) where {R}
    c = set_problem!(Ctx(), iter.F, iter.g, iter.N, iter.x0)
    x0 = Vector{Float64}(realdata(iter.x0))
    # :59-99.  The random restart of the stepsize estimate (:77-83) draws from Julia's global RNG exactly like the reference:
    # the library calls back for every component with ∇f_i(x0+1) == ∇f_i(x0), in ascending order, with t = 1, 2, 4, …
    function perturb(::Ptr{Cvoid}, i::Int64, t::Int64, xeps::Ptr{Float64})::Cint
        println("initial upper bound for L too small")                                               # :78
        unsafe_copyto!(xeps, pointer(x0 .+ rand(t * [-1, 1], size(x0))), length(x0))                 # :79
        return Cint(0)
    end
    cb = @cfunction($perturb, Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}))
    GC.@preserve x0 cb check(ccall((:ciao_finito_adaptive_init_cb, libciao), Cint,
                                   (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Ptr{Cvoid}, Ptr{Cvoid}),
                                   c.h, x0, Float64(iter.α), Float64(iter.tol_b), cb, C_NULL))
    state = FINITO_adaptive_state{R}(c, outvec(iter.x0)..., zeros(iter.N), R(0), collect(1:iter.N), 0, 0)
    return refresh!(state), state
end

function next_index!(iter::FINITO_adaptive_iterable, state::FINITO_adaptive_state)               # :107-119
    if iter.sweeping == 1
        state.idxr = rand(1:iter.N)
    elseif iter.sweeping == 2
        state.idxr = mod(state.idxr, iter.N) + 1
    else
        if state.idx == iter.N
            state.ind = randperm(iter.N); state.idx = 1
        else
            state.idx += 1
        end
        state.idxr = state.ind[state.idx]
    end
    return state.idxr
end

# k steps in one call; returns the number completed (< k ⇔ the reference's `return nothing`, :124-127)
function steps!(iter::FINITO_adaptive_iterable, state::FINITO_adaptive_state, k::Int)
    idx = Int64[next_index!(iter, state) for _ = 1:k]
    done = Ref{Int64}(0)
    GC.@preserve idx check(ccall((:ciao_finito_adaptive_steps, libciao), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Int64}),
                                 state.ctx.h, idx, k, done))
    refresh!(state)
    done[] < k && @warn "parameter `γ` became too small ($(state.γ))"
    return Int(done[])
end
function Base.iterate(iter::FINITO_adaptive_iterable{R}, state::FINITO_adaptive_state{R}) where {R}
    steps!(iter, state, 1) < 1 && return nothing
    return state, state
end
solution(state::FINITO_adaptive_state) = getvec!(state.ctx, CIAO_VEC_Z, state.z, state.buf)       # :162

function table_iterable(solver::Finito{R}, x0, F, g, L, N) where {R}
    if solver.adaptive && !solver.LFinito                                                        # Finito.jl:92-103
        return FINITO_adaptive_iterable{R,typeof(x0),typeof(F),typeof(g)}(F, g, x0, N, L, solver.tol, solver.tol_b, solver.sweeping, solver.α)
    end
    kind = solver.LFinito ? :lfinito : :finito
    Table_iterable{R,typeof(x0),typeof(F),typeof(g)}(kind, F, g, x0, N, L, solver.γ, solver.sweeping, solver.minibatch[2], solver.α)
end
table_iterable(solver::Proshi{R}, x0, F, g, L, N) where {R} =
    Table_iterable{R,typeof(x0),typeof(F),typeof(g)}(:proshi, F, g, x0, N, L, solver.γ, solver.sweeping, solver.minibatch[2], solver.α)

# Finito.jl:66-117, ProShI.jl:42-90 (one method per solver type, as upstream)
(solver::Finito{R})(x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, N = N) where {R} =
    drive(solver, table_iterable(solver, x0, F, g, L, N), solver.maxit, s -> s.hat_γ)
(solver::Proshi{R})(x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, N = N) where {R} =
    drive(solver, table_iterable(solver, x0, F, g, L, N), solver.maxit, s -> s.hat_γ)
iterator(solver::Finito{R}, x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, N = N) where {R} =
    table_iterable(solver, x0, F, g, L, N)
iterator(solver::Proshi{R}, x0::AbstractArray; F = nothing, g = ProximalOperators.Zero(), L = nothing, N = N) where {R} =
    table_iterable(solver, x0, F, g, L, N)

# ---- the driver loop of SVRG.jl:70-83 with the steps between two prints fused into one call ----------
function drive(solver, iter, maxit::Int, field)
    disp(it, state) = @printf "%5d | %.3e  \n" it field(state)
    next = iterate(iter)
    next === nothing && return solution(nothing)          # MethodError, as upstream
    state = next[1]
    it = 1
    fused = !(iter isa SVRG_basic_iterable)
    while it < maxit
        nxt = solver.verbose ? min(maxit, (div(it, solver.freq) + 1) * solver.freq) : maxit
        if iter isa FINITO_adaptive_iterable
            req = nxt - it
            done = steps!(iter, state, req)
            it += done
            done < req && break                             # the iterator ended (`return nothing`)
        elseif fused
            steps!(iter, state, nxt - it); it = nxt
        else
            iterate(iter, state); it += 1
        end
        solver.verbose && mod(it, solver.freq) == 0 && disp(it, state)
    end
    solver.verbose && mod(it, solver.freq) !== 0 && disp(it, state)
    return solution(state), it
end

end # module
