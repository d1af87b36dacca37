# CRNNB200.jl — thin Julia shim over libcrnn_b200.so for the DENG-MIT/CRNN scripts.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is kept
# deliberately thin (one `ccall` per C entry point, column-major arrays passed straight
# through) so that it can be checked by reading it against include/crnn_b200.h.
#
# What stays in the script:  p2vec, the optimiser (`update!`), the epoch loop, checkpoints.
# What this replaces:        `solve(prob, alg, u0=u0, p=p, ...)` inside predict_neuralode
#                            (case2/case2.jl:126) and `ForwardDiff.gradient(x -> loss_neuralode(x, i), p)`
#                            (case2/case2.jl:195).
module CRNNB200

using ForwardDiff

const LIB = get(ENV, "CRNN_B200_LIB", "libcrnn_b200.so")

# mirrors of the C structs (include/crnn_b200.h)
struct CModel
    n_state::Int32; n_species::Int32; n_in::Int32; n_reac::Int32; rhs_kind::Int32; n_tab::Int32
    lb::Float64; ub::Float64; gas_R::Float64
    out_scale::Ptr{Float64}; w_in::Ptr{Float64}; w_b::Ptr{Float64}; w_out::Ptr{Float64}
    mw::Ptr{Float64}; tab_t::Ptr{Float64}; tab_T::Ptr{Float64}; tab_P::Ptr{Float64}   # F2 (HyChem) / F5 (Cathode) only
    w_obs::Ptr{Float64}            # observable post-map (heat release, Cathode/src/network.jl:82-91) or C_NULL
    mlp_n_layers::Int32; mlp_act_out::Int32                      # F4 (yeast / QSSA: an MLP supplies the hidden species); 0 otherwise
    mlp_dims::Ptr{Int32}; mlp_in_idx::Ptr{Int32}; mlp_params::Ptr{Float64}; aug_src::Ptr{Int32}; w_J::Ptr{Float64}
end
struct COpts
    alg::Int32; sens_mode::Int32; err_norm_includes_sens::Int32; n_save::Int32; n_obs::Int32
    n_abstol::Int32; n_reltol::Int32; buffers_on_device::Int32
    maxiters::Int64
    t0::Float64; t1::Float64; pred_clamp_lo::Float64; pred_clamp_hi::Float64
    abstol::Ptr{Float64}; reltol::Ptr{Float64}; saveat::Ptr{Float64}; obs_idx::Ptr{Int32}
    qmin::Float64; qmax::Float64; gamma::Float64; beta1::Float64; beta2::Float64
    stream::Ptr{Cvoid}
    qsteady_min::Float64; qsteady_max::Float64
    err_norm_mean_over_state_only::Int32; reserved0::Int32
end

struct CTrainOpts
    p2vec_kind::Int32; optimiser::Int32; batch::Int32; reserved::Int32
    eta::Float64; beta1::Float64; beta2::Float64; eps::Float64; weight_decay::Float64
    expdecay_eta::Float64; expdecay_decay::Float64; expdecay_clip::Float64
    expdecay_step::Int64
    grad_max::Float64
    p2vec_b0::Float64
    n_save_used::Ptr{Int32}        # host [n_steps * batch] or C_NULL: `sample = rand(batchsize:datasize)` per visit (rober_crnn.jl:218)
end

mutable struct Engine
    h::Ptr{Cvoid}
    function Engine(device::Integer=-1)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:crnn_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
        rc == 0 || error("crnn_create failed ($rc): no CUDA device (there is no CPU fallback)")
        e = new(ref[]); finalizer(x -> ccall((:crnn_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), e); e
    end
    "One handle over several GPUs of this process: `Engine([0,1,2,3,4,5,6,7])` (crnn_create_multi)."
    function Engine(devices::AbstractVector{<:Integer})
        ids = Vector{Int32}(devices); ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:crnn_create_multi, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Int32}, Int32), ref, ids, length(ids))
        rc == 0 || error("crnn_create_multi failed ($rc)")
        e = new(ref[]); finalizer(x -> ccall((:crnn_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), e); e
    end
end

"Device-resident `u0_list` / `ode_data_list` (uploaded once; sharded over the GPUs of a multi-device Engine)."
mutable struct Dataset
    d::Ptr{Cvoid}; N::Int
    function Dataset(e::Engine, u0s::Matrix{Float64}, data::Array{Float64,3})   # n_state × N, n_obs × n_save × N
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(e, ccall((:crnn_dataset_create, LIB), Cint,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Int64, Ref{Ptr{Cvoid}}),
              e.h, u0s, data, size(u0s, 1), size(data, 1), size(data, 2), size(u0s, 2), ref))
        ds = new(ref[], size(u0s, 2)); finalizer(x -> ccall((:crnn_dataset_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.d), ds); ds
    end
end
check(e::Engine, rc) = rc == 0 || error(unsafe_string(ccall((:crnn_last_error, LIB), Cstring, (Ptr{Cvoid},), e.h)))

"Problem constants a script defines once (tsteps, tolerances, lb/ub, i_obs, dydt_scale ...)."
Base.@kwdef struct Setup
    rhs_kind::Int32 = 0            # 0: F0 (case1/3/robertson), 1: F1 (case2: Arrhenius row, T as last state), 4: F4 (MLP-augmented inputs, predict),
                                   # 2: F2 (HyChem/crnn_pyrolysis_mass.jl: mass fractions, tabulated T(t), P(t)),
                                   # 3: F5 (Cathode/src/network.jl:68-80: temperature programme T(t), no density map)
    alg::Int32 = 0                 # 0 Tsit5, 1 Rosenbrock23, 2 KenCarp4, 3 AutoTsit5(Rosenbrock23()), 4 TRBDF2, 5 AutoTsit5(TRBDF2())
    sens_mode::Int32 = 1           # 1 forward (ForwardDiff semantics), 2 interpolating adjoint, 3 discrete adjoint
    gas_R::Float64 = 1.98720425864083e-3
    mw::Vector{Float64} = Float64[]        # F2: l_MW
    tab_t::Vector{Float64} = Float64[]     # F2: knots of itpT / itpP
    tab_T::Vector{Float64} = Float64[]
    tab_P::Vector{Float64} = Float64[]
    w_obs::Vector{Float64} = Float64[]     # heat-release weights w_delH (Cathode/src/network.jl:121); empty: observe rows of u
    # F4 (rhs_kind = 4; yeast_glycolysis.jl:128-142, rober_crnn_qssa.jl:111-126): the Flux chain that supplies the hidden input rows
    mlp_dims::Vector{Int32} = Int32[]      # layer widths, e.g. [7, 5, 5, 5, 5]
    mlp_params::Vector{Float64} = Float64[]   # Float64.(Flux.destructure(dudt2)[1])
    mlp_in_idx::Vector{Int32} = Int32[]    # 0-based state rows fed to the MLP
    mlp_act_out::Int32 = 0                 # 0 softplus, 1 exp
    aug_src::Vector{Int32} = Int32[]       # per input row of the CRNN: >= 0 state row, < 0: MLP output -1 - value
    w_J::Vector{Float64} = Float64[]       # additive source term (yeast)
    lb::Float64; ub::Float64
    abstol::Vector{Float64} = [1e-6]; reltol::Vector{Float64} = [1e-3]
    tspan::Tuple{Float64,Float64}; saveat::Vector{Float64}
    obs_idx::Vector{Int32}         # 0-based rows of u (i_obs .- 1)
    pred_clamp::Tuple{Float64,Float64} = (-Inf, Inf)
    out_scale::Union{Nothing,Vector{Float64}} = nothing
    maxiters::Int64 = 100_000
    loss_kind::Int32 = 0           # 0 MAE-scaled, 1 MAE-log (case3)
    yscale::Vector{Float64} = Float64[]    # per observed species (case2.jl:83); used by the Dataset form of loss_grad
end
yscale_of(s::Setup) = isempty(s.yscale) ? ones(length(s.obs_idx)) : s.yscale

function with_structs(f, s::Setup, w_in, w_b, w_out, w_J=s.w_J, mlp_params=s.mlp_params)
    w_in = Matrix{Float64}(w_in); w_b = Vector{Float64}(w_b); w_out = Matrix{Float64}(w_out)
    w_J = Vector{Float64}(w_J); mlp_params = Vector{Float64}(mlp_params)   # F4: the current values when p2vec returns them
    ns, nr = size(w_out); n_in = size(w_in, 1)
    osc = s.out_scale === nothing ? Float64[] : s.out_scale
    GC.@preserve w_in w_b w_out w_J mlp_params osc s begin
        n_state = s.rhs_kind >= 2 ? ns : n_in      # F0/F1: n_in == n_state; F2/F5: n_in = n_species + 2; F4: n_state = n_species
        i2p(v) = isempty(v) ? Ptr{Int32}(C_NULL) : pointer(v)
        f2p(v) = isempty(v) ? Ptr{Float64}(C_NULL) : pointer(v)
        m = CModel(n_state, ns, n_in, nr, s.rhs_kind, length(s.tab_t), s.lb, s.ub, s.gas_R,
                   isempty(osc) ? C_NULL : pointer(osc), pointer(w_in), pointer(w_b), pointer(w_out),
                   f2p(s.mw), f2p(s.tab_t), f2p(s.tab_T), f2p(s.tab_P), f2p(s.w_obs),
                   max(length(s.mlp_dims) - 1, 0), s.mlp_act_out, i2p(s.mlp_dims), i2p(s.mlp_in_idx), f2p(mlp_params),
                   i2p(s.aug_src), f2p(w_J))
        o = COpts(s.alg, s.sens_mode, 1, length(s.saveat), length(s.obs_idx), length(s.abstol), length(s.reltol), 0,
                  s.maxiters, s.tspan[1], s.tspan[2], s.pred_clamp[1], s.pred_clamp[2],
                  pointer(s.abstol), pointer(s.reltol), pointer(s.saveat), pointer(s.obs_idx),
                  0.0, 0.0, 0.0, 0.0, 0.0, C_NULL, 0.0, 0.0, 0, 0)
        f(Ref(m), Ref(o))
    end
end

"""
    predict_neuralode(e, s, p2vec, u0s, p) -> pred[n_obs, n_save, N], n_saved, retcode

Drop-in for the scripts' `predict_neuralode(u0, p)`, batched: `u0s` is n_state × N.
"""
function predict_neuralode(e::Engine, s::Setup, p2vec, u0s::Matrix{Float64}, p; sample=nothing)
    W = p2vec(p)                   # (w_in, w_b, w_out); an F4 p2vec may append (w_J, pnn), else Setup's w_J / mlp_params are used
    N = size(u0s, 2)
    nsu = sample === nothing ? Int32[] : Vector{Int32}(sample)   # n_save_used: tspan = [0, tsteps[sample]] (rober_crnn.jl:125)
    pred = zeros(length(s.obs_idx), length(s.saveat), N)
    n_saved = zeros(Int32, N); ret = zeros(Int32, N)
    with_structs(s, W...) do m, o
        check(e, ccall((:crnn_solve_batch, LIB), Cint,
              (Ptr{Cvoid}, Ref{CModel}, Ref{COpts}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}),
              e.h, m, o, u0s, N, sample === nothing ? C_NULL : nsu, pred, n_saved, ret, C_NULL))
    end
    pred, n_saved, ret
end

"""
    loss_grad(e, s, p2vec, u0s, data, yscale, p) -> (mean loss, mean gradient)

Replaces `ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p)` for a batch of experiments:
`data` is n_obs × n_save × N (a `permutedims` of the scripts' `ode_data_list[i, :, :]`).
The result feeds the unchanged `update!(opt, p, grad)` line.
"""
# rows of the seed matrix: [vec(w_in); w_b; vec(w_out)] and, for an F4 model whose p2vec also returns (w_J, pnn), [...; w_J; pnn] — the
# adjoint sens_modes (2, 3) then return the gradient of the CRNN weights AND the Flux chain (yeast_glycolysis.jl:136-145,246)
flat_weights(p2vec) = q -> vcat(map(vec, p2vec(q))...)

function loss_grad(e::Engine, s::Setup, p2vec, u0s::Matrix{Float64}, data::Array{Float64,3}, yscale::Vector{Float64}, p; sample=nothing)
    nsu = sample === nothing ? Int32[] : Vector{Int32}(sample)   # per-experiment n_save_used (rober_crnn.jl:218)
    W = p2vec(p)                                               # (w_in, w_b, w_out) or, F4, (w_in, w_b, w_out[1:ns, :], w_J, pnn)
    dWdp = Matrix{Float64}(ForwardDiff.jacobian(flat_weights(p2vec), p))      # n_w × np seed matrix
    N = size(u0s, 2); np_ = length(p)
    loss = zeros(N); grad = zeros(np_); n_saved = zeros(Int32, N); ret = zeros(Int32, N)
    with_structs(s, W...) do m, o
        check(e, ccall((:crnn_loss_grad_batch, LIB), Cint,
              (Ptr{Cvoid}, Ref{CModel}, Ref{COpts}, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Int32},
               Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}),
              e.h, m, o, dWdp, np_, u0s, N, sample === nothing ? C_NULL : nsu, data, yscale, s.loss_kind, loss, grad,
              C_NULL, n_saved, ret, C_NULL))
    end
    ok = n_saved .> 0
    sum(loss[ok]) / max(count(ok), 1), grad ./ max(count(ok), 1)
end

"""
    loss_grad(e, s, p2vec, ds, idx, p; sample=nothing) -> (mean loss, mean gradient)

The training-loop form: `ds` is a device-resident `Dataset`, `idx` the 1-based experiments of this step
(`randperm(n_exp_train)[...]`, case2.jl:194; `nothing` = all of them), `sample` the per-experiment random
time truncation of robertson/rober_crnn.jl:218 (`n_save_used`).  Per step only the weights and the seed matrix
travel to the GPU(s) and np + 2 doubles come back.
"""
function loss_grad(e::Engine, s::Setup, p2vec, ds::Dataset, idx, p; sample=nothing)
    W = p2vec(p)
    dWdp = Matrix{Float64}(ForwardDiff.jacobian(flat_weights(p2vec), p))
    np_ = length(p); lsum = zeros(2); grad = zeros(np_)
    ix = idx === nothing ? Int64[] : Vector{Int64}(idx .- 1)
    nsu = sample === nothing ? Int32[] : Vector{Int32}(sample)
    with_structs(s, W...) do m, o
        check(e, ccall((:crnn_loss_grad_indexed, LIB), Cint,
              (Ptr{Cvoid}, Ref{CModel}, Ref{COpts}, Ptr{Float64}, Int32, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int32},
               Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}),
              e.h, m, o, dWdp, np_, ds.d, idx === nothing ? C_NULL : ix, idx === nothing ? ds.N : length(ix),
              sample === nothing ? C_NULL : nsu, yscale_of(s), s.loss_kind, lsum, grad, C_NULL, C_NULL, C_NULL, C_NULL))
    end
    lsum[1] / max(lsum[2], 1.0), grad ./ max(lsum[2], 1.0)
end

"""
    loss_grad_particles(e, s, weights, seeds, u0s, data, yscale; tab_T=nothing) -> (loss[E, P], grad[np, P])

The SVGD loop `for j = 1:size(p)[1] ... ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p_temp)`
(Cathode_NCM333_UQ/src_333/network.jl:222-260) as ONE launch: `weights` is n_w × P (each particle's
`vcat(vec(w_in), w_b, vec(w_out), w_delH)`), `seeds` n_w × np × P (each particle's Jacobian of p2vec), `u0s` n_state × E,
`data` n_obs × n_save × E, `tab_T` n_tab × E per-experiment temperature programmes.
"""
function loss_grad_particles(e::Engine, s::Setup, weights::Matrix{Float64}, seeds::Array{Float64,3}, u0s::Matrix{Float64},
                             data::Array{Float64,3}, yscale::Vector{Float64}; tab_T=nothing)
    P = size(weights, 2); E = size(u0s, 2); np_ = size(seeds, 2)
    loss = zeros(E, P); grad = zeros(np_, P)
    # the model struct only carries dimensions and constants here; every particle's weights come from `weights`
    n_state = size(u0s, 1); n_in = s.rhs_kind >= 2 ? n_state + 2 : n_state
    nreac = div(size(weights, 1), n_in + 1 + n_state + (isempty(s.w_obs) ? 0 : 1))
    with_structs(s, zeros(n_in, nreac), zeros(nreac), zeros(n_state, nreac)) do m, o
        check(e, ccall((:crnn_loss_grad_particles, LIB), Cint,
              (Ptr{Cvoid}, Ref{CModel}, Ref{COpts}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Int32, Ptr{Int32},
               Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}),
              e.h, m, o, weights, seeds, np_, P, u0s, E, C_NULL, data, tab_T === nothing ? C_NULL : tab_T, C_NULL, yscale,
              s.loss_kind, loss, grad, C_NULL, C_NULL, C_NULL))
    end
    loss, grad
end

"""
    train_steps!(e, s, ds, order, p, opt_state; ...) -> (step_loss, step_gnorm)

The epoch loop `for i_exp in randperm(n_exp_train); grad = ForwardDiff.gradient(...); update!(opt, p, grad); end`
(case2/case2.jl:192-198) with every optimiser step ON the device: `order` is the 1-based visiting order (the script's own
`randperm`), `p` and `opt_state` (2np + 4: ADAM m, v, beta powers, ExpDecay eta and count) are updated in place.  The
device runs the script's p2vec itself (`p2vec_kind` 2: case2.jl:91-99, 1: case1.jl:70-78, 3: case3.jl:42-53 with `s.out_scale = dy_std`, 4: rober_crnn.jl:85-96 with
`s.alg = 1`, `s.out_scale = dydt_scale` and `sample` = the per-visit `rand(batchsize:datasize)`), so no weights are passed.
"""
function train_steps!(e::Engine, s::Setup, ds::Dataset, order, p::Vector{Float64}, opt_state::Vector{Float64};
                      ns::Integer, nr::Integer, batch::Integer=1, optimiser::Integer=0, eta=1e-3, beta=(0.9, 0.999), eps=1e-8,
                      weight_decay=0.0, expdecay=(0.0, 1.0, 0, 0.0), grad_max=0.0, p2vec_kind::Integer=2, p2vec_b0=-10.0, sample=nothing)
    ord = Vector{Int64}(order .- 1); n_steps = div(length(ord), batch)
    step_loss = zeros(n_steps); step_gnorm = zeros(n_steps)
    nsu = sample === nothing ? Int32[] : Vector{Int32}(sample)     # one entry per visited experiment
    n_in = s.rhs_kind == 1 ? ns + 1 : ns
    GC.@preserve nsu begin
        t = CTrainOpts(p2vec_kind, optimiser, batch, 0, eta, beta[1], beta[2], eps, weight_decay, expdecay[1], expdecay[2], expdecay[4],
                       expdecay[3], grad_max, p2vec_b0, sample === nothing ? Ptr{Int32}(C_NULL) : pointer(nsu))
        with_structs(s, zeros(n_in, nr), zeros(nr), zeros(ns, nr)) do m, o
            check(e, ccall((:crnn_train_steps, LIB), Cint,
                  (Ptr{Cvoid}, Ref{CModel}, Ref{COpts}, Ref{CTrainOpts}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Int32,
                   Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  e.h, m, o, Ref(t), ds.d, ord, n_steps, yscale_of(s), s.loss_kind, p, opt_state, step_loss, step_gnorm))
        end
    end
    step_loss, step_gnorm
end

"""
    grad_each(e, N, np) -> np × N matrix

Per-experiment gradients d loss_i / d p of the last forward-mode `loss_grad` call - the rows
`rober_crnn_lm.jl:216-218` assembles with `ForwardDiff.jacobian` for its Levenberg-Marquardt step.
"""
function grad_each(e::Engine, N::Integer, np_::Integer)
    g = zeros(np_, N)
    check(e, ccall((:crnn_copy_grad_each, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Cvoid}),
                   e.h, g, N, np_, 0, C_NULL))
    g
end

end # module
