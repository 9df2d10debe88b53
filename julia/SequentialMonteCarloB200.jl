# SequentialMonteCarloB200.jl — drop-in for the particle-filter path of SequentialMonteCarlo.jl
# (charlesknipp/sequential_monte_carlo) over libsmcb200.so (include/smcb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia (SURVEY.md F7).  The file is a
# mechanical mapping of the reference's exported names onto the C ABI; the same mapping is exercised
# from Python (sequential_monte_carlo_b200/*.py) by the test-suite.  Same names, argument order and
# return shapes as the reference:
#   normalize, resample, bootstrap_filter, bootstrap_filter!, log_likelihood      src/particles.jl
#   StateSpaceModel, LinearModel, UnivariateLinearGaussian, LinearGaussian, unobserved_components,
#   UC, UCSV, StochasticVolatility, simulate                                      src/state_space_models.jl
#   SMC, smc², smc²!, density_tempered, expected_parameters                       src/smc_samplers.jl
#   kalman_filter, log_likelihood(y, model)  (scalar and matrix methods)          src/kalman_filter.jl
#   particle_filter, particle_filter!  (guided: affine-Gaussian proposals, docs/SPEC.md §10)   src/particles.jl:28-84
#   MultivariateLinearGaussian, hodrick_prescott  (Kalman filter only)            src/state_space_models.jl:137-202
module SequentialMonteCarloB200

using Distributions, LinearAlgebra, Printf, Statistics

export StateSpaceModel, LinearModel, UnivariateLinearGaussian, LinearGaussian, unobserved_components, UC, UCSV,
       StochasticVolatility, simulate, transition, observation, initial_dist, normalize, resample, bootstrap_filter, bootstrap_filter!, log_likelihood,
       SMC, smc², smc²!, density_tempered, expected_parameters, kalman_filter,
       particle_filter, particle_filter!, AffineGaussianProposal, locally_optimal_proposal,
       MultivariateLinearModel, MultivariateLinearGaussian, hodrick_prescott, state_variances

const LIB = get(ENV, "SMCB200_LIB", joinpath(@__DIR__, "..", "sequential_monte_carlo_b200", "lib", "libsmcb200.so"))
const MULTINOMIAL, STRATIFIED, SYSTEMATIC = Cint(0), Cint(1), Cint(2)
const P_THETA_RESAMPLE, P_MH_PROPOSAL, P_MH_ACCEPT = UInt32(5), UInt32(6), UInt32(7)

# ---------------------------------------------------------------- context
mutable struct Context
    h::Ptr{Cvoid}
    seed::UInt64
    function Context(device::Integer=0, seed::Integer=1998)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:smcb_create, LIB), Cint, (Cint, UInt64, Ref{Ptr{Cvoid}}), device, seed, ref)
        rc == 0 || error(unsafe_string(ccall((:smcb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[], UInt64(seed))
        finalizer(c -> ccall((:smcb_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
    end
end
const DEFAULT = Ref{Union{Nothing,Context}}(nothing)
context() = (DEFAULT[] === nothing && (DEFAULT[] = Context()); DEFAULT[])
check(ctx, rc) = rc == 0 || error(unsafe_string(ccall((:smcb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))
set_rng!(ctx, seed, epoch) = (ctx.seed = seed; check(ctx, ccall((:smcb_set_rng, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt32), ctx.h, seed, epoch)))
# state storage of the single filter from the next bootstrap_filter / log_likelihood on: Float64 (default) or Float32
# (docs/SPEC.md §9: states rounded to binary32 where stored, arithmetic Float64, sorted resamplers only)
set_precision!(ctx, ::Type{Float64}) = check(ctx, ccall((:smcb_set_precision, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, 0))
set_precision!(ctx, ::Type{Float32}) = check(ctx, ccall((:smcb_set_precision, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, 1))

# ---------------------------------------------------------------- models (state_space_models.jl)
abstract type StateSpaceModel end
struct LinearModel <: StateSpaceModel            # :46-58, univariate
    A::Float64; B::Float64; Q::Float64; R::Float64; x0::Float64; σ0::Float64
end
UnivariateLinearGaussian(; A, B, Q, R, x0=0.0, σ0=1.0) = LinearModel(A, B, Q, R, x0, σ0)       # :74-77
LinearGaussian(A, B, Q, R, x0=0.0, σ0=1.0) = LinearModel(A, B, Q, R, x0, σ0)                    # README.md:12-15
unobserved_components(; σε, ση, x0) = LinearModel(1.0, 1.0, σε, ση, x0, σε)                     # :119-128
UC(x0, σε, ση) = unobserved_components(σε=σε, ση=ση, x0=x0)      # order fixed by the example's prior [Normal(3,2), U(0,4), U(0,4)] (:33-37)
struct UCSV <: StateSpaceModel                   # :215-222
    γ::Tuple{Float64,Float64}; x0::Float64; log_σ0::Tuple{Float64,Float64}
end
UCSV(γ::Real, x0::Real, log_σ0::Tuple) = UCSV((Float64(γ), Float64(γ)), Float64(x0), Float64.(log_σ0))
struct StochasticVolatility <: StateSpaceModel   # absent upstream (SURVEY F6)
    μ::Float64; ρ::Float64; σ::Float64
end
StateSpaceModel(spec::StateSpaceModel, dims) = spec                                           # README constructor
kind(::LinearModel) = Cint(0); kind(::StochasticVolatility) = Cint(1); kind(::UCSV) = Cint(2)
statedim(m) = m isa UCSV ? 3 : 1
params8(m::LinearModel) = Float64[m.A, m.B, m.Q, m.R, m.x0, m.σ0, 0, 0]
params8(m::StochasticVolatility) = Float64[m.μ, m.ρ, m.σ, 0, 0, 0, 0, 0]
params8(m::UCSV) = Float64[m.γ[1], m.γ[2], m.x0, m.log_σ0[1], m.log_σ0[2], 0, 0, 0]

# the model methods the reference exports (state_space_models.jl:1): host-side descriptions of what the device functors evaluate
initial_dist(m::LinearModel) = Normal(m.x0, sqrt(m.σ0))                                        # :105-109
transition(m::LinearModel, x::Float64) = Normal(m.A * x, sqrt(m.Q))                            # :87-94
observation(m::LinearModel, x::Float64) = Normal(m.B * x, sqrt(m.R))                           # :96-103
initial_dist(m::StochasticVolatility) = Normal(m.μ, m.σ / sqrt(1 - m.ρ^2))
transition(m::StochasticVolatility, x::Float64) = Normal(m.μ + m.ρ * (x - m.μ), m.σ)
observation(m::StochasticVolatility, x::Float64) = Normal(0.0, exp(0.5 * x))
initial_dist(m::UCSV) = (Normal(m.x0, exp(0.5 * m.log_σ0[1])), Normal(m.log_σ0[1], m.γ[1]), Normal(m.log_σ0[2], m.γ[2]))   # :249-259
transition(m::UCSV, x::Vector{Float64}) = (Normal(x[1], exp(0.5 * x[2])), Normal(x[2], m.γ[1]), Normal(x[3], m.γ[2]))       # :233-242
observation(m::UCSV, x::Vector{Float64}) = Normal(x[1], exp(0.5 * x[3]))                       # :244-247

function simulate(model::StateSpaceModel, T::Int64; seed::Integer=1998)                       # :11-28
    d = statedim(model); x = Matrix{Float64}(undef, T, d); y = Vector{Float64}(undef, T)     # column-major [T,d] == C [d][T]
    ccall((:smcb_simulate, LIB), Cint, (Cint, Ptr{Float64}, Int64, UInt64, Ptr{Float64}, Ptr{Float64}),
          kind(model), params8(model), T, seed, x, y)
    return (d == 1 ? vec(x) : [x[t, :] for t in 1:T]), y
end

# ---------------------------------------------------------------- particles.jl
"device-resident cloud: materialised on `collect` / indexing (SURVEY H6)"
mutable struct Cloud
    ctx::Context; N::Int; d::Int; gen::Int
end
function Base.collect(c::Cloud)
    x = Matrix{Float64}(undef, c.N, c.d)                                                       # [N,d] col-major == C [d][N]
    check(c.ctx, ccall((:smcb_fetch_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.ctx.h, x, C_NULL, C_NULL))
    return c.d == 1 ? vec(x) : [x[i, :] for i in 1:c.N]                                        # Vector{Vector{Float64}} for UCSV (:229-231)
end
function weights(c::Cloud)
    w = Vector{Float64}(undef, c.N)
    check(c.ctx, ccall((:smcb_fetch_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.ctx.h, C_NULL, w, C_NULL))
    return w
end

# quantile(x, p) (README.md:41,51) and quantile(x, weights(w), p), var(x, weights(w)) (examples/inflation_example.jl:44-46)
# of a device-resident cloud, computed on the device: nothing is read back (docs/SPEC.md §8)
function summary(c::Cloud, p::Vector{Float64}; weighted::Bool=true)
    m = Vector{Float64}(undef, c.d); v = similar(m); q = Matrix{Float64}(undef, length(p), c.d)   # [np,d] col-major == C [d][np]
    check(c.ctx, ccall((:smcb_weighted_summary, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       c.ctx.h, p, length(p), weighted, m, v, q))
    return m, v, q
end
Statistics.quantile(c::Cloud, p::Vector{Float64}) = (q = summary(c, p; weighted=false)[3]; c.d == 1 ? vec(q) : q)
Statistics.quantile(c::Cloud, ::Cloud, p::Vector{Float64}) = (q = summary(c, p)[3]; c.d == 1 ? vec(q) : q)   # second argument: the weights handle
Statistics.var(c::Cloud, ::Cloud) = (v = summary(c, Float64[])[2]; c.d == 1 ? v[1] : v)
Statistics.mean(c::Cloud, ::Cloud) = (m = summary(c, Float64[])[1]; c.d == 1 ? m[1] : m)

function normalize(logw::Vector{Float64}; ctx=context())                                       # :5-15
    w = similar(logw); lm = Ref(0.0); es = Ref(0.0)
    check(ctx, ccall((:smcb_normalize, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Float64}, Ptr{Float64}, Ref{Float64}),
                     ctx.h, logw, length(logw), lm, w, es))
    return (lm[], w, es[])
end
const reweight = normalize                                                                     # SURVEY F3

function resample(w::Vector{Float64}, N::Int64=length(w); ctx=context(), resampler=MULTINOMIAL, stream=0, t=0, purpose=3)   # :17-19
    a = Vector{Int64}(undef, length(w))
    check(ctx, ccall((:smcb_resample, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, UInt32, UInt32, UInt32, Ptr{Int64}),
                     ctx.h, w, length(w), resampler, stream, t, purpose, a))
    return a .+ 1                                                                              # 1-based
end

function bootstrap_filter(N::Int64, y::Float64, model::StateSpaceModel; ctx=context())         # :87-105
    lm = Ref(0.0); es = Ref(0.0)
    check(ctx, ccall((:smcb_bootstrap_init, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Float64, UInt32, Ref{Float64}, Ref{Float64}),
                     ctx.h, kind(model), params8(model), N, y, 0, lm, es))
    x = Cloud(ctx, N, statedim(model), 0)
    return x, weights(x), lm[]
end

function bootstrap_filter!(states::Cloud, weights_::Vector{Float64}, y::Float64, model::StateSpaceModel; resampler=MULTINOMIAL)   # :107-129
    lm = Ref(0.0); es = Ref(0.0)
    check(states.ctx, ccall((:smcb_bootstrap_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint, Ref{Float64}, Ref{Float64}),
                            states.ctx.h, params8(model), y, resampler, lm, es))
    return lm[], weights(states), es[]
end

function log_likelihood(N::Int64, y::Vector{Float64}, model::StateSpaceModel; ctx=context(), resampler=MULTINOMIAL)   # :132-147
    z = Ref(0.0)
    check(ctx, ccall((:smcb_log_likelihood, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint, UInt32, Ref{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, kind(model), params8(model), N, y, length(y), resampler, 0, z, C_NULL, C_NULL))
    x = Cloud(ctx, N, statedim(model), 0)
    return x, weights(x), z[]
end

# ---------------------------------------------------------------- kalman_filter.jl
function kalman_filter(model::LinearModel, xt::Float64, Σt::Float64, yt::Float64; ctx=context())   # :29-53
    x = [xt]; s = [Σt]; ll = [0.0]
    check(ctx, ccall((:smcb_kalman_batch_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, params8(model), 1, yt, x, s, ll))
    return x[1], s[1], ll[1]
end
function log_likelihood(y::Vector{Float64}, model::LinearModel; ctx=context(), matched_init=false)   # :55-70
    ll = [0.0]; x = [0.0]; s = [0.0]
    check(ctx, ccall((:smcb_kalman_batch_loglik, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, params8(model), C_NULL, 1, y, length(y), matched_init, ll, x, s))
    return x[1], s[1], ll[1]
end

# matrix methods (:3-27) for multivariate LinearModels with a scalar observation, d <= 4
struct MultivariateLinearModel <: StateSpaceModel                                             # state_space_models.jl:137-154
    A::Matrix{Float64}; B::Matrix{Float64}; Q::Matrix{Float64}; R::Vector{Float64}; x0::Vector{Float64}; σ0::Matrix{Float64}
end
MultivariateLinearGaussian(; A, B, Q, R, X0=zeros(size(A, 1)), Σ0=Matrix(1.0I(size(A, 1)))) =
    MultivariateLinearModel(Float64.(A), Float64.(reshape(B, 1, :)), Float64.(Q), Float64.(vec(R)), Float64.(X0), Float64.(Matrix(Σ0)))
hodrick_prescott(; λ, y, init_cov=1000.0) = MultivariateLinearGaussian(                        # :187-202
    A=[2.0 -1.0; 1.0 0.0], B=[1.0 0.0], Q=[1 / λ 0.0; 0.0 0.0], R=[1.0],
    X0=[3 * y[1] - 2 * y[2], 2 * y[1] - y[2]], Σ0=Matrix(init_cov * I(2)))
# row-major block A, B, Q, R, x0, Σ0 (include/smcb200.h); Julia is column-major, hence the transposes
block(m::MultivariateLinearModel) = vcat(vec(m.A'), vec(m.B), vec(m.Q'), m.R[1:1], m.x0, vec(m.σ0'))
function kalman_filter(model::MultivariateLinearModel, xt::Vector{Float64}, Σt::Matrix{Float64}, yt::Float64; ctx=context())   # :3-27
    d = length(xt); x = copy(xt); s = collect(vec(Σt')); ll = [0.0]
    check(ctx, ccall((:smcb_kalman_mv_batch_step, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, d, block(model), 1, yt, x, s, ll))
    return x, collect(reshape(s, d, d)'), ll[1]
end
function log_likelihood(y::Vector{Float64}, model::MultivariateLinearModel; ctx=context(), matched_init=false)   # :55-70
    d = length(model.x0); ll = [0.0]; x = zeros(d); s = zeros(d * d)
    check(ctx, ccall((:smcb_kalman_mv_batch_loglik, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{UInt8}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, d, block(model), C_NULL, 1, y, length(y), matched_init, ll, x, s))
    return x, collect(reshape(s, d, d)'), ll[1]
end

# ---------------------------------------------------------------- smc_samplers.jl
mutable struct Batch
    h::Ptr{Cvoid}; ctx::Context
    function Batch(ctx, kind, M, N)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx, ccall((:smcb_batch_create, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ref{Ptr{Cvoid}}), ctx.h, kind, M, N, ref))
        b = new(ref[], ctx); finalizer(b -> ccall((:smcb_batch_destroy, LIB), Cint, (Ptr{Cvoid},), b.h), b)
    end
end

mutable struct SMC{SSM,KT}                                                                     # :5-27
    θ::Vector{Vector{Float64}}; ω::Vector{Float64}
    ess::Float64; ess_min::Float64; N::Int64; M::Int64; chain::Int64; logZ::Vector{Float64}
    model::SSM; prior::Sampleable; kernel::KT; acc_threshold::Float64; acc_ratio::Float64
    ctx::Context; cur::Batch; prop::Union{Nothing,Batch}; epoch::UInt32; nres::UInt32; nrej::UInt32; resampler::Cint
end
params(smc, θ) = reduce(hcat, [params8(smc.model(th)) for th in θ])                            # [8,M] col-major == C [M][8]
next_epoch!(smc) = (set_rng!(smc.ctx, smc.ctx.seed, smc.epoch); smc.epoch += 1)

function SMC(N::Int64, M::Int64, model, prior::Sampleable, chain::Int64, ess_threshold::Float64, min_ar::Float64=-1.0;
             ctx=context(), resampler=MULTINOMIAL)                                             # :29-59
    θ = map(m -> collect(rand(prior)), 1:M)
    cur = Batch(ctx, kind(model(θ[1])), M, N)
    SMC(θ, fill(1 / M, M), 1.0 * M, M * ess_threshold, N, M, chain, zeros(M), model, prior, random_walk_kernel, min_ar, 0.0,
        ctx, cur, nothing, UInt32(1), UInt32(0), UInt32(0), resampler)
end
x(smc::SMC) = (a = Array{Float64}(undef, smc.N, statedim(smc.model(smc.θ[1])), smc.M);
               check(smc.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), smc.cur.h, a, C_NULL, C_NULL)); a)
w(smc::SMC) = (a = Matrix{Float64}(undef, smc.N, smc.M);
               check(smc.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), smc.cur.h, C_NULL, a, C_NULL)); a)

# quantile(smc.x[i], weights(smc.w[i]), p) for every θ-particle at once (examples/inflation_example.jl:44,250): [np, d, M]
function state_quantiles(smc::SMC, p::Vector{Float64}; weighted::Bool=true)
    q = Array{Float64}(undef, length(p), statedim(smc.model(smc.θ[1])), smc.M)
    check(smc.ctx, ccall((:smcb_batch_weighted_quantiles, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
                         smc.cur.h, p, length(p), weighted ? 1 : 0, q))
    q
end

# mean(smc.x[i], weights(smc.w[i])), var(smc.x[i], weights(smc.w[i])) for every θ-particle at once (inflation_example.jl:46): [d, M] each
function state_variances(smc::SMC)
    d = statedim(smc.model(smc.θ[1])); mean = Matrix{Float64}(undef, d, smc.M); var = Matrix{Float64}(undef, d, smc.M)
    check(smc.ctx, ccall((:smcb_batch_weighted_moments, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), smc.cur.h, mean, var))
    mean, var
end

# ---------------------------------------------------------------- guided filters (particles.jl:28-84, docs/SPEC.md §10)
# The device evaluates proposals of the family x' ~ Normal(c0 + c1*xp, c2); a proposal is called as proposal(model, y)
# and returns (c0, c1, c2) for the step that assimilates y (the closure over y the reference's signature implies).
struct AffineGaussianProposal
    c0::Float64; c1::Float64; c2::Float64
end
(q::AffineGaussianProposal)(model, y) = (q.c0, q.c1, q.c2)
function locally_optimal_proposal(model::LinearModel, y::Float64)      # p(x' | xp, y) of a univariate LinearModel
    s2 = 1 / (1 / model.Q + model.B^2 / model.R)
    (s2 * model.B * y / model.R, s2 * model.A / model.Q, sqrt(s2))
end
mutable struct GuidedCloud                        # a guided filter is a batch of one θ (N <= 8192)
    b::Batch; N::Int64
end
Base.collect(c::GuidedCloud) = (a = Vector{Float64}(undef, c.N);
    check(c.b.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.b.h, a, C_NULL, C_NULL)); a)
weights(c::GuidedCloud) = (a = Vector{Float64}(undef, c.N);
    check(c.b.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.b.h, C_NULL, a, C_NULL)); a)
particle_filter(N::Int64, y::Float64, model::StateSpaceModel, ::Nothing; ctx=context()) = bootstrap_filter(N, y, model; ctx=ctx)   # :28-51
function particle_filter(N::Int64, y::Float64, model::StateSpaceModel, proposal; ctx=context())
    N > 8192 && return bootstrap_filter(N, y, model; ctx=ctx)           # large clouds continue on the single filter
    b = Batch(ctx, kind(model), 1, N); lm = [0.0]; es = [0.0]
    check(ctx, ccall((:smcb_batch_init, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Float64, UInt32, Ptr{Float64}, Ptr{Float64}),
                     b.h, params8(model), C_NULL, y, 0, lm, es))
    x = GuidedCloud(b, N)
    return x, weights(x), lm[1]
end
particle_filter!(states::Cloud, w::Vector{Float64}, y::Float64, model::StateSpaceModel, ::Nothing; resampler=MULTINOMIAL) =
    bootstrap_filter!(states, w, y, model; resampler=resampler)                                # :55-84 with proposal = nothing
function particle_filter!(states::Cloud, w::Vector{Float64}, y::Float64, model::StateSpaceModel, proposal; resampler=SYSTEMATIC)
    c = Float64[proposal(model, y)...]; lm = Ref(0.0); es = Ref(0.0)     # the large-N single filter: sorted resamplers
    check(states.ctx, ccall((:smcb_guided_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint, Ptr{Float64}, Ref{Float64}, Ref{Float64}),
                            states.ctx.h, params8(model), y, resampler, c, lm, es))
    return lm[], weights(states), es[]
end
function particle_filter!(states::GuidedCloud, w::Vector{Float64}, y::Float64, model::StateSpaceModel, proposal; resampler=MULTINOMIAL)
    c = Float64[proposal(model, y)...]; lm = [0.0]; es = [0.0]
    check(states.b.ctx, ccall((:smcb_batch_step_guided, LIB), Cint,
                              (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                              states.b.h, params8(model), y, resampler, c, lm, es))
    return lm[1], weights(states), es[1]
end

expected_parameters(smc::SMC) = sum(reduce(hcat, smc.θ .* smc.ω), dims=2)                      # :61-65 (properly weighted: SURVEY D6)

function random_walk_kernel(θ::Vector{Vector{Float64}})                                        # :94-101
    Θ = hcat(θ...); dθ = 2.83^2 / size(Θ, 1)
    Σ = norm(cov(Θ')) < 1.e-8 ? 1.e-2I(size(Θ, 1)) : dθ * cov(Θ') + 1.e-10I
    return Matrix(Σ)
end

function resample!(smc::SMC)                                                                   # :74-84
    set_rng!(smc.ctx, smc.ctx.seed, 0)
    a = resample(smc.ω; ctx=smc.ctx, t=smc.nres, purpose=P_THETA_RESAMPLE); smc.nres += 1
    sort!(a)                                                                                   # docs/SPEC.md §5b: exchangeable slots, sorted parents stay on their GPU
    smc.θ = smc.θ[a]; smc.logZ = smc.logZ[a]; smc.ω = fill(1 / smc.M, smc.M)
    check(smc.ctx, ccall((:smcb_batch_gather, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), smc.cur.h, Int32.(a .- 1)))
end

function host_normals(smc, ordinal, k, c, M)
    z = Vector{Float64}(undef, M)
    ccall((:smcb_rng_normals, LIB), Cint, (UInt64, UInt32, UInt32, UInt32, UInt32, UInt32, Int64, Ptr{Float64}),
          smc.ctx.seed, ordinal, k, c, P_MH_PROPOSAL, 0, M, z); z
end
function host_uniforms(smc, ordinal, c, M)
    u = Vector{UInt64}(undef, M)
    ccall((:smcb_rng_uniforms64, LIB), Cint, (UInt64, UInt32, UInt32, UInt32, UInt32, Int64, Ptr{UInt64}),
          smc.ctx.seed, ordinal, 0, c, P_MH_ACCEPT, M, u)
    return Float64.(u .>> 11) .* 2.0^-53
end

function rejuvenate!(smc::SMC, y::Vector{Float64}, ξ::Float64, verbose::Bool)                  # :103-148
    M, d = smc.M, length(smc.θ[1]); acc = falses(M)
    Σ = smc.kernel(smc.θ); scales = 0.5 * reverse(1:smc.chain)
    verbose && @printf("\t[rejuvenating]")
    smc.prop === nothing && (smc.prop = Batch(smc.ctx, kind(smc.model(smc.θ[1])), M, smc.N))
    ordinal = smc.nrej; smc.nrej += 1
    for c in 1:smc.chain
        L = cholesky(Symmetric(scales[c] * Σ)).L
        Z = reduce(hcat, [host_normals(smc, ordinal, k - 1, c - 1, M) for k in 1:d])          # [M,d]
        θp = [smc.θ[m] + L * Z[m, :] for m in 1:M]
        ok = [insupport(smc.prior, θp[m]) for m in 1:M]
        P = params(smc, [ok[m] ? θp[m] : smc.θ[m] for m in 1:M]); zp = Vector{Float64}(undef, M)
        next_epoch!(smc)
        check(smc.ctx, ccall((:smcb_batch_log_likelihood, LIB), Cint,
                             (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Int64, Cint, UInt32, Ptr{Float64}),
                             smc.prop.h, P, UInt8.(ok), y, length(y), smc.resampler, 0, zp))   # ONE launch for all θ (:117-121)
        u = host_uniforms(smc, ordinal, c - 1, M); accept = falses(M)
        for m in 1:M
            ok[m] || continue
            lp = logpdf(smc.prior, θp[m])
            ratio = ξ * (zp[m] - smc.logZ[m]) + lp - logpdf(smc.prior, smc.θ[m])
            if zp[m] + lp > -Inf && log(u[m]) < ratio
                smc.logZ[m] = zp[m]; smc.θ[m] = θp[m]; accept[m] = true; acc[m] = true
            end
        end
        check(smc.ctx, ccall((:smcb_batch_accept, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}), smc.cur.h, smc.prop.h, UInt8.(accept)))
    end
    smc.ω = fill(1 / M, M); smc.acc_ratio = sum(acc) / M
    verbose && @printf("\tacc_rate: %1.5f", smc.acc_ratio)
    return smc
end

function density_tempered(smc::SMC, y::Vector{Float64}, verbose=true)                          # :222-281
    next_epoch!(smc)
    check(smc.ctx, ccall((:smcb_batch_log_likelihood, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Int64, Cint, UInt32, Ptr{Float64}),
                         smc.cur.h, params(smc, smc.θ), C_NULL, y, length(y), smc.resampler, 0, smc.logZ))
    _, smc.ω, smc.ess = reweight(smc.logZ; ctx=smc.ctx)
    ξ = 0.0
    while ξ < 1.0
        resample_flag = true; lower_bound = oldξ = ξ; upper_bound = 2.0
        local newξ, logω
        while upper_bound - lower_bound > 1.e-6
            newξ = (upper_bound + lower_bound) / 2.0
            logω = (newξ - oldξ) * smc.logZ
            _, smc.ω, smc.ess = reweight(logω; ctx=smc.ctx)
            if smc.ess == smc.ess_min; break
            elseif smc.ess < smc.ess_min; upper_bound = newξ
            else lower_bound = newξ end
        end
        if newξ ≥ 1.0
            resample_flag = false; newξ = 1.0; logω = (newξ - oldξ) * smc.logZ
            _, smc.ω, smc.ess = reweight(logω; ctx=smc.ctx)
        end
        ξ = newξ
        verbose && @printf("ξ = %1.5f\tess = %4.3f", ξ, smc.ess)
        if resample_flag
            resample!(smc); rejuvenate!(smc, y, ξ, verbose)
        end
        verbose && print("\n")
    end
end

function smc²(smc::SMC, y::Vector{Float64})                                                    # :288-301
    next_epoch!(smc); lm = Vector{Float64}(undef, smc.M)
    check(smc.ctx, ccall((:smcb_batch_init, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Float64, UInt32, Ptr{Float64}, Ptr{Float64}),
                         smc.cur.h, params(smc, smc.θ), C_NULL, y[1], 0, lm, C_NULL))
    smc.logZ = copy(lm); _, smc.ω, smc.ess = reweight(lm; ctx=smc.ctx)
    return smc
end

function smc²!(smc::SMC, y::Vector{Float64}, t::Int64, verbose::Bool=true)                     # :308-340
    verbose && @printf("t = %4d\tess = %4.3f", t - 1, smc.ess)
    if smc.ess < smc.ess_min
        resample!(smc); rejuvenate!(smc, y[1:(t-1)], 1.0, verbose)
    end
    logω = log.(smc.ω); lm = Vector{Float64}(undef, smc.M)
    check(smc.ctx, ccall((:smcb_batch_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint, Ptr{Float64}, Ptr{Float64}),
                         smc.cur.h, params(smc, smc.θ), y[t], smc.resampler, lm, C_NULL))       # ONE launch (:325-331)
    logω .+= lm; smc.logZ .+= lm
    _, smc.ω, smc.ess = reweight(logω; ctx=smc.ctx)
    verbose && print("\n")
end

end # module
