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
#   SMC, smc², smc²!, density_tempered, expected_parameters (exchange! included;   src/smc_samplers.jl
#     the whole θ level runs on the GPU(s) behind smcb_sampler_*; comm_init! shards θ over one process per GPU)
#   IBIS + smc², smc²!                                                            src/ibis.jl
#   kalman_filter, log_likelihood(y, model)  (scalar and matrix methods)          src/kalman_filter.jl
#   particle_filter, particle_filter!  (guided: affine-Gaussian proposals, docs/SPEC.md §10; UCSV trend move, §10b)   src/particles.jl:28-84
#   MultivariateLinearGaussian, hodrick_prescott  (Kalman filter only)            src/state_space_models.jl:137-202
module SequentialMonteCarloB200

using Distributions, LinearAlgebra, Printf, Statistics

export StateSpaceModel, LinearModel, UnivariateLinearGaussian, LinearGaussian, unobserved_components, UC, UCSV,
       StochasticVolatility, simulate, transition, observation, initial_dist, normalize, resample, bootstrap_filter, bootstrap_filter!, log_likelihood,
       SMC, IBIS, smc², smc²!, density_tempered, expected_parameters, kalman_filter, comm_unique_id, comm_init!,
       particle_filter, particle_filter!, AffineGaussianProposal, UCSVTrendProposal, locally_optimal_proposal,
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
# The sampler lives on the GPU(s): θ, ω, logZ, the log-prior and the parameter blocks of all M θ-particles are device arrays
# behind ONE handle (smcb_sampler, include/smcb200.h); smc², smc²!, density_tempered are one ccall each and the public fields of
# the reference's struct (θ, ω, logZ, ess, N, acc_ratio, x, w  — smc_samplers.jl:5-27) are read back when somebody looks at them.
mutable struct Batch                  # M filters of N particles; also the view of a sampler's live clouds
    h::Ptr{Cvoid}; ctx::Context; owned::Bool
end
function Batch(ctx::Context, kind, M, N)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:smcb_batch_create, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ref{Ptr{Cvoid}}), ctx.h, kind, M, N, ref))
    finalizer(b -> b.owned && ccall((:smcb_batch_destroy, LIB), Cint, (Ptr{Cvoid},), b.h), Batch(ref[], ctx, true))
end

# smcb_sampler_config (include/smcb200.h); isbits, passed by reference
struct SamplerConfig
    kind::Int32; d_theta::Int32; N::Int64; M::Int64; chain::Int32; resampler::Int32; theta_resampler::Int32; reserved::Int32
    ess_threshold::Float64; min_ar::Float64; seed::UInt64
    prior::NTuple{64,Float64}; map_src::NTuple{8,Int32}; map_const::NTuple{8,Float64}
end

# product of univariate priors -> rows (family, p0, p1, lo, hi, c0, c1, 0) of the device prior table (README.md:81-85)
prior_row(d::Normal) = (0.0, d.μ, d.σ, 0.0, 0.0, log(d.σ), 0.0, 0.0)
prior_row(d::LogNormal) = (1.0, d.μ, d.σ, 0.0, 0.0, log(d.σ), 0.0, 0.0)
prior_row(d::Uniform) = (2.0, 0.0, 0.0, d.a, d.b, -log(d.b - d.a), 0.0, 0.0)
prior_row(d::Truncated{<:Normal}) = (3.0, d.untruncated.μ, d.untruncated.σ, d.lower, d.upper, log(d.untruncated.σ), d.logtp, 0.0)
components(p::Product) = p.v
components(p::UnivariateDistribution) = [p]

# model(θ) as a selection map: params[k] = θ[src[k]] or a constant — found by probing the closure on prior draws; every
# model closure of the reference's README / example (lg_mod, uc_mod, ucsv_mod) is of this form
function parameter_map(model, prior)
    Θ = [collect(rand(prior)) for _ in 1:6]; P = [params8(model(θ)) for θ in Θ]
    src = fill(Int32(-1), 8); cst = zeros(8)
    for k in 1:8
        col = [p[k] for p in P]
        j = findfirst(j -> all(Θ[i][j] == col[i] for i in 1:6), 1:length(Θ[1]))
        if j !== nothing; src[k] = j - 1
        elseif all(==(col[1]), col); cst[k] = col[1]
        else error("model(θ) must select components of θ and constants into the model's parameters (device-resident sampler)") end
    end
    return kind(model(Θ[1])), src, cst
end

mutable struct SMC{SSM}                                                                        # :5-27
    h::Ptr{Cvoid}; ctx::Context; model::SSM; prior::Sampleable; M::Int64; chain::Int64; ess_min::Float64; acc_threshold::Float64
    d::Int; y::Vector{Float64}; schedule::Vector{Tuple{Float64,Float64}}; kernel::Function
end
function SMC(N::Int64, M::Int64, model, prior::Sampleable, chain::Int64, ess_threshold::Float64, min_ar::Float64=-1.0;
             ctx=context(), resampler=MULTINOMIAL, theta_resampler=MULTINOMIAL)                # :29-59
    θ0 = [collect(rand(prior)) for _ in 1:M]; d = length(θ0[1])                                # θ = map(m -> rand(prior), 1:M)   :38
    k, src, cst = parameter_map(model, prior)
    rows = vcat([collect(prior_row(c)) for c in components(prior)]..., zeros(8 * (8 - d)))
    cfg = SamplerConfig(k, d, N, M, chain, resampler, theta_resampler, 0, ess_threshold, min_ar, ctx.seed, Tuple(rows), Tuple(src), Tuple(cst))
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:smcb_sampler_create, LIB), Cint, (Ptr{Cvoid}, Ref{SamplerConfig}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                     ctx.h, Ref(cfg), reduce(hcat, θ0), ref))                                   # [d,M] col-major == C [M][d]
    smc = SMC(ref[], ctx, model, prior, M, chain, M * ess_threshold, min_ar, d, Float64[], Tuple{Float64,Float64}[], random_walk_kernel)
    finalizer(s -> ccall((:smcb_sampler_destroy, LIB), Cint, (Ptr{Cvoid},), s.h), smc)
end
function sampler_get(smc::SMC)
    θ = Matrix{Float64}(undef, smc.d, smc.M); ω = Vector{Float64}(undef, smc.M); z = similar(ω)
    ess = Ref(0.0); ar = Ref(0.0); N = Ref{Int64}(0)
    check(smc.ctx, ccall((:smcb_sampler_get, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Float64}, Ref{Int64}),
                         smc.h, θ, ω, z, ess, ar, N))
    return (θ=[θ[:, m] for m in 1:smc.M], ω=ω, logZ=z, ess=ess[], acc_ratio=ar[], N=N[])
end
function clouds(smc::SMC)                                                                      # smc.x, smc.w live here
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(smc.ctx, ccall((:smcb_sampler_clouds, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), smc.h, ref))
    Batch(ref[], smc.ctx, false)
end
function Base.getproperty(smc::SMC, f::Symbol)                                                 # the struct's public fields
    f in (:θ, :ω, :logZ, :ess, :acc_ratio, :N) && return getfield(sampler_get(smc), f)
    f === :x && return fetch_x(clouds(smc), smc)
    f === :w && return fetch_w(clouds(smc), smc)
    return getfield(smc, f)
end
statedim(smc::SMC) = statedim(smc.model(sampler_get(smc).θ[1]))
fetch_x(b::Batch, smc) = (a = Array{Float64}(undef, sampler_get(smc).N, statedim(smc), smc.M);
    check(smc.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), b.h, a, C_NULL, C_NULL)); a)
fetch_w(b::Batch, smc) = (a = Matrix{Float64}(undef, sampler_get(smc).N, smc.M);
    check(smc.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), b.h, C_NULL, a, C_NULL)); a)
function set_data!(smc::SMC, y::Vector{Float64})
    y == getfield(smc, :y) && return
    check(smc.ctx, ccall((:smcb_sampler_set_data, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), smc.h, y, length(y)))
    setfield!(smc, :y, copy(y))
end

# quantile(smc.x[i], weights(smc.w[i]), p) for every θ-particle at once (examples/inflation_example.jl:44,250): [np, d, M]
function state_quantiles(smc::SMC, p::Vector{Float64}; weighted::Bool=true)
    q = Array{Float64}(undef, length(p), statedim(smc), smc.M)
    check(smc.ctx, ccall((:smcb_batch_weighted_quantiles, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
                         clouds(smc).h, p, length(p), weighted ? 1 : 0, q))
    q
end
# mean(smc.x[i], weights(smc.w[i])), var(smc.x[i], weights(smc.w[i])) for every θ-particle at once (inflation_example.jl:46): [d, M] each
function state_variances(smc::SMC)
    d = statedim(smc); mean = Matrix{Float64}(undef, d, smc.M); var = Matrix{Float64}(undef, d, smc.M)
    check(smc.ctx, ccall((:smcb_batch_weighted_moments, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), clouds(smc).h, mean, var))
    mean, var
end

# ---------------------------------------------------------------- multi-GPU: one Julia process per GPU, θ sharded (SURVEY §8e)
# rank 0: id = comm_unique_id(); carry the 128 bytes to the other ranks (MPI.bcast, a shared file, Distributed.jl ...); then
# EVERY rank: comm_init!(ctx, rank, nranks, id).  Samplers created from ctx afterwards shard their θ-particles over the ranks
# (all-gathers and cloud moves run inside the library over NCCL / NVLink); their results do not depend on the number of GPUs.
comm_unique_id() = (id = Vector{UInt8}(undef, 128); ccall((:smcb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || error("NCCL is not available"); id)
comm_init!(ctx::Context, rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    check(ctx, ccall((:smcb_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, rank, nranks, id))

# ---------------------------------------------------------------- guided filters (particles.jl:28-84, docs/SPEC.md §10)
# The device evaluates proposals of the family x' ~ Normal(c0 + c1*xp, c2); a proposal is called as proposal(model, y)
# and returns (c0, c1, c2) for the step that assimilates y (the closure over y the reference's signature implies).
struct AffineGaussianProposal
    c0::Float64; c1::Float64; c2::Float64
end
(q::AffineGaussianProposal)(model, y) = (q.c0, q.c1, q.c2)
# UCSV (docs/SPEC.md §10b): the log-volatilities move by the transition, the trend by the conditionally optimal Gaussian move
# tempered by κ in [0, 1] (0 = bootstrap, 1 = p(x' | x, le, ln', y)); the device takes the triple (κ, 0, 1)
struct UCSVTrendProposal
    κ::Float64
    UCSVTrendProposal(κ=1.0) = (0.0 <= κ <= 1.0 || error("κ must lie in [0, 1]"); new(κ))
end
(q::UCSVTrendProposal)(model, y) = (q.κ, 0.0, 1.0)
function locally_optimal_proposal(model::LinearModel, y::Float64)      # p(x' | xp, y) of a univariate LinearModel
    s2 = 1 / (1 / model.Q + model.B^2 / model.R)
    (s2 * model.B * y / model.R, s2 * model.A / model.Q, sqrt(s2))
end
mutable struct GuidedCloud                        # a guided filter is a batch of one θ (N <= 8192)
    b::Batch; N::Int64; d::Int64                  # d state components (UCSV: 3), fetched as [d][N]
end
Base.collect(c::GuidedCloud) = (a = Vector{Float64}(undef, c.d * c.N);
    check(c.b.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.b.h, a, C_NULL, C_NULL));
    c.d == 1 ? a : [a[(k - 1) * c.N + i] for i in 1:c.N, k in 1:c.d])
weights(c::GuidedCloud) = (a = Vector{Float64}(undef, c.N);
    check(c.b.ctx, ccall((:smcb_batch_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), c.b.h, C_NULL, a, C_NULL)); a)
particle_filter(N::Int64, y::Float64, model::StateSpaceModel, ::Nothing; ctx=context()) = bootstrap_filter(N, y, model; ctx=ctx)   # :28-51
function particle_filter(N::Int64, y::Float64, model::StateSpaceModel, proposal; ctx=context())
    N > 8192 && return bootstrap_filter(N, y, model; ctx=ctx)           # large clouds continue on the single filter
    b = Batch(ctx, kind(model), 1, N); lm = [0.0]; es = [0.0]
    check(ctx, ccall((:smcb_batch_init, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Float64, UInt32, Ptr{Float64}, Ptr{Float64}),
                     b.h, params8(model), C_NULL, y, 0, lm, es))
    x = GuidedCloud(b, N, kind(model) == 2 ? 3 : 1)
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

expected_parameters(smc::SMC) = (g = sampler_get(smc); sum(reduce(hcat, g.θ .* g.ω), dims=2))    # :61-65 (properly weighted: SURVEY D6)

# Σ of the random-walk proposal in the frozen summation order of docs/SPEC.md §11 (what the device sampler uses)  :87-101
function random_walk_kernel(θ::Vector{Vector{Float64}})
    d = length(θ[1]); Σ = Matrix{Float64}(undef, d, d)
    ccall((:smcb_random_walk_sigma, LIB), Cint, (Ptr{Float64}, Int64, Cint, Ptr{Float64}), reduce(hcat, θ), length(θ), d, Σ)
    return Matrix(Σ')
end

# smc²(smc, y): M bootstrap filters at y[1], one launch                                        :288-301
function smc²(smc::SMC, y::Vector{Float64})
    set_data!(smc, y)
    check(smc.ctx, ccall((:smcb_sampler_smc2_init, LIB), Cint, (Ptr{Cvoid},), smc.h))
    return smc
end

# smc²!(smc, y, t): resample! / rejuvenate! / exchange! when ess < ess_min, then M filter steps at y[t] and the reweighting —
# all inside the library (θ-level vectors never leave the GPU); t is Julia's 1-based index                   :308-340
function smc²!(smc::SMC, y::Vector{Float64}, t::Int64, verbose::Bool=true)
    set_data!(smc, y)
    verbose && @printf("t = %4d\tess = %4.3f", t - 1, smc.ess)
    ess = Ref(0.0); rj = Ref{Cint}(0)
    check(smc.ctx, ccall((:smcb_sampler_smc2_step, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Float64}, Ref{Cint}), smc.h, t - 1, ess, rj))
    verbose && rj[] != 0 && @printf("\t[rejuvenating]\tacc_rate: %1.5f", smc.acc_ratio)
    verbose && print("\n")
end

# density_tempered(smc, y): the whole tempering loop (bisection for ξ on the device, resample!, rejuvenate!) in one call   :222-281
function density_tempered(smc::SMC, y::Vector{Float64}, verbose=true)
    set_data!(smc, y)
    sched = Matrix{Float64}(undef, 3, 4096); n = Ref{Cint}(0)
    check(smc.ctx, ccall((:smcb_sampler_density_tempered, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Cint}), smc.h, sched, 4096, n))
    setfield!(smc, :schedule, [(sched[1, i], sched[2, i]) for i in 1:n[]])
    if verbose
        for i in 1:n[]
            @printf("ξ = %1.5f\tess = %4.3f", sched[1, i], sched[2, i])
            sched[3, i] >= 0 && @printf("\t[rejuvenating]\tacc_rate: %1.5f", sched[3, i])
            print("\n")
        end
    end
end

# ---------------------------------------------------------------- ibis.jl: the same θ-level scheme with the Kalman filter as inner filter
# IBIS(M, model, prior, chain, ess_threshold, min_ar): no state particles — x, Σ are the M filtered means / variances.  The M-wide
# Kalman recursions are one device call each (smcb_kalman_batch_*); the θ level (M numbers) runs here with the host-level Philox
# streams and the frozen proposal arithmetic of docs/SPEC.md §11, so a run reproduces the Python host mirror (ibis.py) draw for draw.
mutable struct IBIS{SSM}                                                                       # ibis.jl:3-24
    θ::Vector{Vector{Float64}}; ω::Vector{Float64}; x::Vector{Float64}; Σ::Vector{Float64}
    ess::Float64; ess_min::Float64; M::Int64; chain::Int64; logZ::Vector{Float64}
    model::SSM; prior::Sampleable; kernel::Function; acc_threshold::Float64; acc_ratio::Float64
    ctx::Context; nres::UInt32; nrej::UInt32
end
function IBIS(M::Int64, model, prior::Sampleable, chain::Int64, ess_threshold::Float64, min_ar::Float64=-1.0; ctx=context())   # ibis.jl:26-52
    θ = [collect(rand(prior)) for _ in 1:M]; ms = model.(θ)
    all(m -> m isa LinearModel, ms) || throw(ArgumentError("IBIS needs a univariate LinearModel (its inner filter is kalman_filter)"))
    IBIS(θ, fill(1 / M, M), [m.x0 for m in ms], [m.σ0 for m in ms], 1.0 * M, M * ess_threshold, M, chain, zeros(M), model, prior,
         random_walk_kernel, min_ar, 0.0, ctx, UInt32(0), UInt32(0))
end
pblock(s::IBIS, θ) = reduce(hcat, [params8(s.model(th)) for th in θ])                          # [8,M] col-major == C [M][8]
function kalman_all!(s::IBIS, y::Float64)                                                      # M × kalman_filter(model(θ_m), x_m, Σ_m, y)
    ll = Vector{Float64}(undef, s.M)
    check(s.ctx, ccall((:smcb_kalman_batch_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       s.ctx.h, pblock(s, s.θ), s.M, y, s.x, s.Σ, ll))
    ll
end
function host_normals(seed, ordinal, k, c, M)
    z = Vector{Float64}(undef, M)
    ccall((:smcb_rng_normals, LIB), Cint, (UInt64, UInt32, UInt32, UInt32, UInt32, UInt32, Int64, Ptr{Float64}), seed, ordinal, k, c, P_MH_PROPOSAL, 0, M, z); z
end
function host_uniforms(seed, ordinal, c, M)
    u = Vector{UInt64}(undef, M)
    ccall((:smcb_rng_uniforms64, LIB), Cint, (UInt64, UInt32, UInt32, UInt32, UInt32, Int64, Ptr{UInt64}), seed, ordinal, 0, c, P_MH_ACCEPT, M, u)
    Float64.(u .>> 11) .* 2.0^-53
end
function smc²(s::IBIS, y::Vector{Float64})                                                     # ibis.jl:128-147
    ll = kalman_all!(s, y[1]); s.logZ = copy(ll); _, s.ω, s.ess = reweight(ll; ctx=s.ctx); s
end
function smc²!(s::IBIS, y::Vector{Float64}, t::Int64, verbose::Bool=true)                      # ibis.jl:154-189
    if s.ess < s.ess_min
        set_rng!(s.ctx, s.ctx.seed, 0)
        a = sort!(resample(s.ω; ctx=s.ctx, t=s.nres, purpose=P_THETA_RESAMPLE)); s.nres += 1  # ibis.jl:72-84 (+ docs/SPEC.md §5b)
        s.θ, s.x, s.Σ, s.logZ, s.ω = s.θ[a], s.x[a], s.Σ[a], s.logZ[a], fill(1 / s.M, s.M)
        d = length(s.θ[1]); Σp = s.kernel(s.θ); acc = falses(s.M); ord = s.nrej; s.nrej += 1  # ibis.jl:86-126
        yy = y[1:t-1]; lp = [logpdf(s.prior, th) for th in s.θ]
        for c in 1:s.chain
            L = Matrix{Float64}(undef, d, d); sc = 0.5 * (s.chain - c + 1)
            d == 1 ? (L[1, 1] = sc * Σp[1, 1]) :
                ccall((:smcb_cholesky_lower, LIB), Cint, (Ptr{Float64}, Cint, Float64, Ptr{Float64}), Matrix(Σp'), d, sc, L) == 0 || error("proposal covariance not positive definite")
            d > 1 && (L = Matrix(L'))
            Z = reduce(hcat, [host_normals(s.ctx.seed, ord, k - 1, c - 1, s.M) for k in 1:d])
            θp = [s.θ[m] .+ [sum(Z[m, k] * L[j, k] for k in 1:j) for j in 1:d] for m in 1:s.M]
            ok = [insupport(s.prior, th) for th in θp]
            P = pblock(s, [ok[m] ? θp[m] : s.θ[m] for m in 1:s.M]); zp = fill(-Inf, s.M); xp = similar(s.x); Sp = similar(s.Σ)
            check(s.ctx, ccall((:smcb_kalman_batch_loglik, LIB), Cint,
                               (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                               s.ctx.h, P, UInt8.(ok), s.M, yy, length(yy), 0, zp, xp, Sp))
            u = host_uniforms(s.ctx.seed, ord, c - 1, s.M)
            for m in findall(ok)
                lpp = logpdf(s.prior, θp[m])
                if zp[m] + lpp > -Inf && log(u[m]) < (zp[m] - s.logZ[m]) + (lpp - lp[m])
                    s.logZ[m], s.θ[m], s.x[m], s.Σ[m], lp[m], acc[m] = zp[m], θp[m], xp[m], Sp[m], lpp, true
                end
            end
        end
        s.acc_ratio = sum(acc) / s.M
    end
    ll = kalman_all!(s, y[t]); logω = log.(s.ω) .+ ll; s.logZ .+= ll
    _, s.ω, s.ess = reweight(logω; ctx=s.ctx)
end
expected_parameters(s::IBIS) = sum(reduce(hcat, s.θ .* s.ω), dims=2)

end # module
