# BossB200.jl -- Julia-side binding of libboss_b200.so for BOSS.jl v0.6.1.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The file is written against
# include/boss_b200.h and BOSS.jl's documented extension interfaces and is syntax-reviewed only; the same
# C ABI is exercised end to end from Python (boss.jl_b200/_lib.py, tests/).
#
# Loading this module after `using BOSS` re-routes the GP hot path to the B200 library by adding more
# specific methods for the reference's own generic functions (no BOSS source changes):
#   BOSS.model_posterior_slice(::GaussianProcess, ::GaussianProcessParams, ::ExperimentData, ::Int)
#       replaces src/models/gaussian_process.jl:133-141 (AbstractGPs.posterior)   -> boss_gp_fit
#   mean / var / mean_and_var(::B200Posterior, x | X)
#       replaces src/models/gaussian_process.jl:143-178                           -> boss_gp_predict
#   BOSS.data_loglike(::GaussianProcess, ::ExperimentData)
#       replaces src/models/gaussian_process.jl:250-280                           -> boss_gp_loglik_batch
#   BOSS.maximize_acquisition(::GridAM | ::SamplingAM, ::BossProblem, ::BossOptions)
#       replaces src/acquisition_maximizers/grid.jl:45-65, sampling.jl:20-57      -> boss_ei_score / boss_ei_score_grid
#   cov / mean_and_cov(::B200Posterior, X)          gaussian_process.jl:163-167,180-184  -> boss_gp_cov
#   BOSS.maximize_acquisition(::SequentialBatchAM, ...)  batch.jl:26-38                   -> boss_gp_append
#   data_loglike_and_grad(model, data, params)      optimization.jl:41,153 (ForwardDiff)  -> boss_gp_loglik_grad_batch
module BossB200

using BOSS
using BOSS: GaussianProcess, GaussianProcessParams, ExperimentData, ModelPosteriorSlice, BossProblem, BossOptions,
            GridAM, SamplingAM, ExpectedImprovement, LinFitness, mean_getindex, best_so_far, get_params, y_dim
using KernelFunctions: SqExponentialKernel, Matern32Kernel, Matern52Kernel
import Statistics: mean, var, cov
import Base: append!
using LinearAlgebra: diag
import StatsBase: mean_and_var, mean_and_cov

const LIB = get(ENV, "BOSS_B200_LIB", joinpath(@__DIR__, "..", "boss.jl_b200", "lib", "libboss_b200.so"))

struct BossB200Error <: Exception
    code::Int
    msg::String
end
last_error() = unsafe_string(ccall((:boss_last_error, LIB), Cstring, ()))
check(rc::Integer) = (rc < 0 && throw(BossB200Error(rc, last_error())); Int(rc))

init(device::Integer=0) = check(ccall((:boss_init, LIB), Cint, (Cint,), device))

kernel_id(::SqExponentialKernel) = 0
kernel_id(::Matern32Kernel) = 1
kernel_id(::Matern52Kernel) = 2
kernel_id(k::BOSS.DiscreteKernel) = kernel_id(k.kernel)
kernel_id(k) = error("BossB200: kernel $(typeof(k)) is outside the accelerated path (use the stock BOSS.jl methods)")
discrete_mask(k::BOSS.DiscreteKernel) = k.dims isa Missing ? nothing : Vector{UInt8}(k.dims)
discrete_mask(k) = nothing

# prior mean evaluated on the host (nothing | constant | closure), gaussian_process.jl:101-103
eval_mean(::Nothing, X::AbstractMatrix) = zeros(size(X, 2))
eval_mean(m::Real, X::AbstractMatrix) = fill(Float64(m), size(X, 2))
eval_mean(m::Function, X::AbstractMatrix) = Float64[m(x) for x in eachcol(X)]

# ---- posterior handle ------------------------------------------------------------------------
mutable struct B200Posterior <: ModelPosteriorSlice{GaussianProcess}
    handle::Ptr{Cvoid}
    mean                         # slice mean: nothing | Real | Function
    loglik::Float64
    function B200Posterior(handle, mean, loglik)
        p = new(handle, mean, loglik)
        finalizer(q -> ccall((:boss_gp_free, LIB), Cvoid, (Ptr{Cvoid},), q.handle), p)
        return p
    end
end

function BOSS.model_posterior_slice(model::GaussianProcess, params::GaussianProcessParams, data::ExperimentData, slice::Int)
    X = Matrix{Float64}(data.X)
    m = mean_getindex(model.mean, slice)
    δ = Vector{Float64}(data.Y[slice, :]) .- eval_mean(m, X)
    λ = Vector{Float64}(params.λ[:, slice])
    mask = discrete_mask(model.kernel)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    ll = Ref{Cdouble}(0.0)
    rc = check(ccall((:boss_gp_fit, LIB), Cint,
        (Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}, Ref{Cdouble}),
        X, size(X, 1), size(X, 2), δ, λ, params.α[slice], params.σ[slice], kernel_id(model.kernel),
        isnothing(mask) ? C_NULL : mask, out, ll))
    rc == 1 && throw(BOSS.LinearAlgebra.PosDefException(1))   # same exception the reference path raises
    return B200Posterior(out[], m, ll[])
end

function mean_and_var(post::B200Posterior, X::AbstractMatrix{<:Real})
    Xs = Matrix{Float64}(X)
    M = size(Xs, 2)
    pm = isnothing(post.mean) ? C_NULL : eval_mean(post.mean, Xs)
    μ = Vector{Float64}(undef, M); σ2 = Vector{Float64}(undef, M); st = Vector{Int32}(undef, M)
    rc = check(ccall((:boss_gp_predict, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
        post.handle, Xs, M, pm, μ, σ2, st))
    if rc == 2   # mirror _clip_var's DomainError (gaussian_process.jl:186-194)
        i = findfirst(!=(0), st)
        throw(DomainError(σ2[i], "The posterior GP predicted variance $(σ2[i]) but only values above -1e-8 are tolerated."))
    end
    return μ, σ2
end
mean_and_var(post::B200Posterior, x::AbstractVector{<:Real}) = first.(mean_and_var(post, hcat(x)))
mean(post::B200Posterior, x) = mean_and_var(post, x)[1]
var(post::B200Posterior, x) = mean_and_var(post, x)[2]

# cov / mean_and_cov(::GaussianProcessPosterior, X)  (gaussian_process.jl:163-167,180-184) -> boss_gp_cov
function mean_and_cov(post::B200Posterior, X::AbstractMatrix{<:Real})
    Xs = Matrix{Float64}(X)
    M = size(Xs, 2)
    pm = isnothing(post.mean) ? C_NULL : eval_mean(post.mean, Xs)
    μ = Vector{Float64}(undef, M); Σ = Matrix{Float64}(undef, M, M)
    rc = check(ccall((:boss_gp_cov, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), post.handle, Xs, M, pm, μ, Σ))
    rc == 2 && throw(DomainError(minimum(diag(Σ)), "The posterior GP predicted a variance below -1e-8."))
    return μ, Σ
end
cov(post::B200Posterior, X::AbstractMatrix{<:Real}) = mean_and_cov(post, X)[2]

# Incremental factor cache: one more data point, same hyper-parameters (O(n^2) instead of a refit).
# Used by the SequentialBatchAM override below in place of `model_posterior(problem)` per speculative point
# (src/acquisition_maximizers/batch.jl:26-38).
function append!(post::B200Posterior, x::AbstractVector{<:Real}, y::Real)
    xv = Vector{Float64}(x)
    δ = Float64(y) - first(eval_mean(post.mean, hcat(xv)))
    ll = Ref{Cdouble}(0.0)
    rc = check(ccall((:boss_gp_append, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ref{Cdouble}),
        post.handle, xv, δ, ll))
    rc == 1 && throw(BOSS.LinearAlgebra.PosDefException(1))
    post.loglik = ll[]
    return post
end

# ---- batched log-likelihood --------------------------------------------------------------------
# data_loglike keeps the reference's closure signature (params -> Real) and additionally accepts a
# vector of params (one library call for the whole batch; used by the batched SamplingMAP below).
function BOSS.data_loglike(model::GaussianProcess, data::ExperimentData)
    X = Matrix{Float64}(data.X)
    d, n = size(X)
    kid = kernel_id(model.kernel)
    mask = discrete_mask(model.kernel)
    ydim = size(data.Y, 1)
    δs = [Vector{Float64}(data.Y[i, :]) .- eval_mean(mean_getindex(model.mean, i), X) for i in 1:ydim]

    function ll_batch(ps::AbstractVector{<:GaussianProcessParams})
        S = length(ps)
        total = zeros(S)
        out = Vector{Float64}(undef, S)
        for i in 1:ydim
            λ = reduce(hcat, [Vector{Float64}(p.λ[:, i]) for p in ps])          # d x S
            α = Float64[p.α[i] for p in ps]; σ = Float64[p.σ[i] for p in ps]
            check(ccall((:boss_gp_loglik_batch, LIB), Cint,
                (Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{UInt8}, Int64, Ptr{Cdouble}),
                X, d, n, δs[i], 0, λ, α, σ, kid, isnothing(mask) ? C_NULL : mask, S, out))
            total .+= out
        end
        return total          # -Inf where K is not positive definite (what safe_data_loglike returns)
    end
    ll_data(p::GaussianProcessParams) = ll_batch([p])[1]
    ll_data(ps::AbstractVector{<:GaussianProcessParams}) = ll_batch(ps)
    return ll_data
end

# Value + gradient w.r.t. [vec(λ); α; σ] per output slice (the vectorizer order, gaussian_process.jl:300-328):
# replaces the ForwardDiff.Dual sweep of OptimizationMAP's gradient algorithms / NUTS
# (src/model_fitters/optimization.jl:41,153).  Returns (ll::Vector (S), grad::Array (d+2, S, y_dim)).
function data_loglike_and_grad(model::GaussianProcess, data::ExperimentData, ps::AbstractVector{<:GaussianProcessParams})
    X = Matrix{Float64}(data.X)
    d, n = size(X)
    kid = kernel_id(model.kernel); mask = discrete_mask(model.kernel)
    ydim = size(data.Y, 1); S = length(ps)
    total = zeros(S); grads = zeros(d + 2, S, ydim)
    out = Vector{Float64}(undef, S); g = Matrix{Float64}(undef, d + 2, S)
    for i in 1:ydim
        δ = Vector{Float64}(data.Y[i, :]) .- eval_mean(mean_getindex(model.mean, i), X)
        λ = reduce(hcat, [Vector{Float64}(p.λ[:, i]) for p in ps])
        α = Float64[p.α[i] for p in ps]; σ = Float64[p.σ[i] for p in ps]
        check(ccall((:boss_gp_loglik_grad_batch, LIB), Cint,
            (Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{UInt8}, Int64,
             Ptr{Cdouble}, Ptr{Cdouble}),
            X, d, n, δ, 0, λ, α, σ, kid, isnothing(mask) ? C_NULL : mask, S, out, g))
        total .+= out
        grads[:, :, i] .= g
    end
    return total, grads
end

# ---- batched acquisition maximisation -------------------------------------------------------------
function score_batch(problem::BossProblem, Xs::Matrix{Float64})
    ei = problem.acquisition::ExpectedImprovement
    ei.fitness isa LinFitness || error("BossB200: only LinFitness is on the accelerated path")
    ps = get_params(problem)
    samples = ps isa AbstractVector ? ps : [ps]
    ydim = y_dim(problem)
    posts = [BOSS.model_posterior_slice(problem.model, p, problem.data, i) for p in samples for i in 1:ydim]
    handles = Ptr{Cvoid}[p.handle for p in posts]
    M = size(Xs, 2)
    means = [p.mean for p in posts[1:ydim]]
    pm = all(isnothing, means) ? C_NULL :
         Matrix{Float64}(reduce(vcat, [eval_mean(m, Xs)' for m in means]))       # ydim x M
    coefs = Vector{Float64}(ei.fitness.coefs)
    b = best_so_far(problem, ei.fitness)
    ymax = Float64[isinf(c) ? Inf : c for c in problem.y_max]
    lb, ub = Vector{Float64}.(problem.domain.bounds)
    cons = isnothing(problem.domain.cons) ? C_NULL : UInt8[all(problem.domain.cons(x) .>= 0.) for x in eachcol(Xs)]
    acq = Vector{Float64}(undef, M); bv = Ref{Cdouble}(0.0); bi = Ref{Int64}(-1)
    GC.@preserve posts check(ccall((:boss_ei_score, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Int64}),
        handles, ydim, length(samples), Xs, M, pm, coefs, isnothing(b) ? C_NULL : Ref(Float64(b)), ymax,
        ei.cons_safe ? lb : C_NULL, ei.cons_safe ? ub : C_NULL, ei.cons_safe ? cons : C_NULL, acq, C_NULL, bv, bi))
    return acq, bv[], Int(bi[]) + 1
end

function BOSS.maximize_acquisition(opt::GridAM, problem::BossProblem, options::BossOptions)
    points = opt.shuffle ? BOSS.shuffle(deepcopy(opt.points)) : opt.points       # grid.jl:47
    Xs = reduce(hcat, points)
    _, val, idx = score_batch(problem, Matrix{Float64}(Xs))
    return points[idx], val
end

# Full product grids without a `cons` filter need no host-side point list at all: the candidates are generated
# on the device from (lo, step, count) (boss_ei_score_grid; grid.jl:30-43 builds the same Iterators.product).
function maximize_grid_on_device(problem::BossProblem, lo::Vector{Float64}, step::Vector{Float64}, count::Vector{Int64})
    ei = problem.acquisition::ExpectedImprovement
    ps = get_params(problem); samples = ps isa AbstractVector ? ps : [ps]
    ydim = y_dim(problem)
    posts = [BOSS.model_posterior_slice(problem.model, p, problem.data, i) for p in samples for i in 1:ydim]
    all(p -> isnothing(p.mean), posts) || error("BossB200: on-device grids need a zero prior mean (closures run on the host)")
    handles = Ptr{Cvoid}[p.handle for p in posts]
    b = best_so_far(problem, ei.fitness)
    lb, ub = Vector{Float64}.(problem.domain.bounds)
    bv = Ref{Cdouble}(0.0); bi = Ref{Int64}(-1); bx = Vector{Float64}(undef, length(lo))
    GC.@preserve posts check(ccall((:boss_ei_score_grid, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble},
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Int64}, Ptr{Cdouble}),
        handles, ydim, length(samples), length(lo), lo, step, count, 0, -1, C_NULL, Vector{Float64}(ei.fitness.coefs),
        isnothing(b) ? C_NULL : Ref(Float64(b)), Float64[isinf(c) ? Inf : c for c in problem.y_max], lb, ub, C_NULL, C_NULL,
        bv, bi, bx))
    return bx, bv[]
end

# SequentialBatchAM (batch.jl:26-38) on the appended factor cache: one fit, then O(n^2) per speculative point.
function BOSS.maximize_acquisition(sb::BOSS.SequentialBatchAM, problem::BossProblem, options::BossOptions)
    problem_ = deepcopy(problem)
    ps = get_params(problem_)
    ps isa AbstractVector && return invoke(BOSS.maximize_acquisition, Tuple{BOSS.SequentialBatchAM, BossProblem, BossOptions},
                                           sb, problem, options)   # BI samples: stock path
    ydim = y_dim(problem_)
    posts = [BOSS.model_posterior_slice(problem_.model, ps, problem_.data, i) for i in 1:ydim]
    xs = Vector{Vector{Float64}}()
    for _ in 1:sb.batch_size
        x, _ = BOSS.maximize_acquisition(sb.am, problem_, options)
        y = [mean(p, x) for p in posts]
        BOSS.augment_dataset!(problem_, x, y)
        foreach(i -> append!(posts[i], x, y[i]), 1:ydim)
        push!(xs, x)
    end
    return reduce(hcat, xs), nothing
end

end # module
