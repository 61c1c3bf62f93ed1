# MMSigB200.jl -- drop-in GPU `fit!` for MultiModalMuSig.jl's MMCTM / CTM / LDA.
#
# The Julia API stays as it is: format_counts_mmctm/ctm/lda, MMCTM(K, α, X), LDA(K, α, η, X),
# fit!(model; tol), model.ϕ / model.props / model.β / model.θ.  Loading this file after
# `using MultiModalMuSig` replaces the two `fit!` methods (reference src/MMCTM.jl:457-494,
# src/LDA.jl:198-224) and MMCTM's `fit_heldout` / `transform` (:554-586, :511-552) by thin wrappers that flatten the model state, `ccall` libmmsig.so
# (include/mmsig.h) and scatter the results back into the nested vectors.  Model construction,
# including the random γ / λ initialisation (src/MMCTM.jl:59-63, src/LDA.jl:36), is untouched.
#
# NOTE: written against the C ABI but NOT executed in the build environment (no Julia there);
# the same entry points are exercised by tests/ through ctypes.
module MMSigB200

using MultiModalMuSig
import MultiModalMuSig: fit!, fit_heldout, transform, MMCTM, IMMCTM, LDA, ILDA, check_convergence

const LIB = get(ENV, "MMSIG_LIB", joinpath(@__DIR__, "..", "multimodalmusig.jl_b200", "libmmsig.so"))

struct MmsigConfig
    device::Int32
    stop_rule::Int32
    profile::Int32
    precision::Int32
    reserved::NTuple{4,Int32}
end

function check(h::Ptr{Cvoid}, rc::Int32)
    if rc != 0
        msg = unsafe_string(ccall((:mmsig_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
        error("libmmsig error $rc: $msg")
    end
end

# precision: 0 = FP64 (default, Julia's Float64), 1 = FP32 tile passes (include/mmsig.h MMSIG_PRECISION_FP32)
function create(; device=0, stop_rule=0, precision=0)
    cfg = Ref(MmsigConfig(Int32(device), Int32(stop_rule), Int32(0), Int32(precision), ntuple(_ -> Int32(0), 4)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mmsig_create, LIB), Int32, (Ref{MmsigConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    check(C_NULL, rc)
    return h[]
end

destroy(h) = ccall((:mmsig_destroy, LIB), Int32, (Ptr{Cvoid},), h)

# X[d][m] (nnz x 2 Int, 1-based terms) -> CSR per modality (int64 rowptr, int32 0-based term, int32 count)
function flatten_counts(X::Vector{Vector{Matrix{Int}}}, M::Int)
    D = length(X)
    rowptr = [zeros(Int64, D + 1) for m in 1:M]
    for m in 1:M, d in 1:D
        rowptr[m][d + 1] = rowptr[m][d] + size(X[d][m], 1)
    end
    term = [Vector{Int32}(undef, rowptr[m][end]) for m in 1:M]
    count = [Vector{Int32}(undef, rowptr[m][end]) for m in 1:M]
    for m in 1:M, d in 1:D
        r = (rowptr[m][d] + 1):rowptr[m][d + 1]
        term[m][r] .= X[d][m][:, 1] .- 1
        count[m][r] .= X[d][m][:, 2]
    end
    return rowptr, term, count
end

# Page-locked ("pinned") host vectors owned by the library (mmsig_host_alloc): copies from / to them run at link
# speed and overlap with kernels, ordinary Julia Vectors are staged by the driver (slower; bench.py prints both).
function pinned_vector(::Type{T}, n::Integer) where T
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:mmsig_host_alloc, LIB), Int32, (UInt64, Ref{Ptr{Cvoid}}), UInt64(max(n, 1) * sizeof(T)), p))
    return unsafe_wrap(Array, Ptr{T}(p[]), n; own=false)
end
host_free(v::Array) = ccall((:mmsig_host_free, LIB), Int32, (Ptr{Cvoid},), pointer(v))

# X[d][m] -> row pointers + 4-byte records term | count << 10 in pinned memory (mmsig_mmctm_fit_host_packed: half the
# host -> device traffic of the counts); `nothing` when a term needs more than 10 bits or a count more than 22
function flatten_counts_packed(X::Vector{Vector{Matrix{Int}}}, M::Int)
    D = length(X)
    rowptr = [pinned_vector(Int64, D + 1) for m in 1:M]
    for m in 1:M
        rowptr[m][1] = 0
        for d in 1:D
            rowptr[m][d + 1] = rowptr[m][d] + size(X[d][m], 1)
        end
    end
    rec = [pinned_vector(UInt32, rowptr[m][end]) for m in 1:M]
    fits = true
    for m in 1:M, d in 1:D
        o = rowptr[m][d]
        for w in 1:size(X[d][m], 1)
            t, c = X[d][m][w, 1] - 1, X[d][m][w, 2]
            fits &= (0 <= t < 1024) & (0 < c < 4194304)
            rec[m][o + w] = UInt32(t & 1023) | (UInt32(c & 4194303) << 10)
        end
    end
    if !fits
        foreach(host_free, rowptr); foreach(host_free, rec)
        return nothing
    end
    return rowptr, rec
end

function flat_rows_pinned(v::Vector{Vector{Float64}})                                   # [d][j] -> D*MK row-major, pinned
    MK = length(v[1])
    out = pinned_vector(Float64, length(v) * MK)
    for d in 1:length(v)
        out[(d - 1) * MK + 1:d * MK] .= v[d]
    end
    return out
end

flat_rows(v::Vector{Vector{Float64}}) = collect(reduce(vcat, v))                         # [d][j] -> D*MK row-major
flat_tables(t::Vector{Vector{Vector{Float64}}}) = collect(reduce(vcat, [reduce(vcat, tm) for tm in t]))   # [m][k][v]

# ---- several GPUs from this process: a group of devices (include/mmsig.h, mmsig_group_*) --------------
function gcheck(g::Ptr{Cvoid}, rc::Int32)
    if rc != 0
        msg = unsafe_string(ccall((:mmsig_group_last_error, LIB), Cstring, (Ptr{Cvoid},), g))
        error("libmmsig error $rc: $msg")
    end
end

function create_group(devices; stop_rule=0, precision=0)
    ids = Int32.(collect(devices))
    cfg = Ref(MmsigConfig(ids[1], Int32(stop_rule), Int32(0), Int32(precision), ntuple(_ -> Int32(0), 4)))
    g = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mmsig_group_create, LIB), Int32, (Ref{MmsigConfig}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), cfg, length(ids), ids, g)
    gcheck(C_NULL, rc)
    return g[]
end

destroy_group(g) = ccall((:mmsig_group_destroy, LIB), Int32, (Ptr{Cvoid},), g)

# flat device results -> the nested vectors of the struct (shapes preserved)
function scatter_state!(model::MMCTM, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props)
    D, M, MK = model.D, model.M, sum(model.K)
    for d in 1:D
        model.λ[d] .= @view λ[(d - 1) * MK + 1:d * MK]
        model.ν[d] .= @view ν[(d - 1) * MK + 1:d * MK]
        model.ζ[d] .= @view ζ[(d - 1) * M + 1:d * M]
        off = 0
        for m in 1:M
            model.props[d][m] = props[(d - 1) * MK + off + 1:(d - 1) * MK + off + model.K[m]]
            off += model.K[m]
        end
    end
    model.μ .= μ
    model.Σ .= transpose(reshape(Σ, MK, MK)); model.invΣ .= transpose(reshape(invΣ, MK, MK))
    o = 0
    for m in 1:M, k in 1:model.K[m]
        r = (o + 1):(o + model.V[m])
        model.γ[m][k] .= @view γ[r]; model.Elnϕ[m][k] .= @view Elnϕ[r]; model.ϕ[m][k] .= @view ϕ[r]
        o += model.V[m]
    end
end

# model.θ[d][m] (K_m x nnz_dm, src/MMCTM.jl:17,52-57): the library never stores θ; it is recomputed from the λ / Elnϕ
# of the last E-step on request, while the handle (or group) is still alive
function materialize_theta!(model::MMCTM, h::Ptr{Cvoid}, grouped::Bool)
    for m in 1:model.M
        nnz = sum(size(model.X[d][m], 1) for d in 1:model.D)
        buf = zeros(model.K[m] * nnz)                                   # [w][k] row-major == K_m x nnz column-major
        if grouped
            gcheck(h, ccall((:mmsig_group_mmctm_get_theta, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), h, m - 1, buf))
        else
            check(h, ccall((:mmsig_mmctm_get_theta, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), h, m - 1, buf))
        end
        o = 0
        for d in 1:model.D
            n = size(model.X[d][m], 1)
            model.θ[d][m] = reshape(buf[o + 1:o + model.K[m] * n], model.K[m], n)
            o += model.K[m] * n
        end
    end
end

# fit!(model; ...) (src/MMCTM.jl:457-494).  Extra keywords: `device` (one GPU), or `devices=0:7` (the samples of this
# fit sharded over several GPUs of this process, bit-identical results), `materialize_θ=true` to fill model.θ.
# `precision=:fp32` selects the optional FP32 mode of the tile passes (include/mmsig.h MMSIG_PRECISION_FP32; default Float64).
function fit!(model::MMCTM; maxiter=100, tol=1e-4, verbose=true, autoα=false, updateΣ=true,
              device=0, devices=nothing, stop_rule=0, materialize_θ=false, precision=:fp64)
    D, M, MK = model.D, model.M, sum(model.K)
    rowptr, term, count = flatten_counts(model.X, M)
    K32, V32 = Int32.(model.K), Int32.(model.V)
    λ, ν = flat_rows(model.λ), flat_rows(model.ν)
    γ = flat_tables(model.γ)
    Σ, invΣ = collect(transpose(model.Σ)), collect(transpose(model.invΣ))               # row-major
    grouped = devices !== nothing && length(devices) > 1
    prec = precision == :fp32 ? 1 : 0
    h = grouped ? create_group(devices; stop_rule=stop_rule, precision=prec) :
                  create(device=(devices === nothing ? device : first(devices)), stop_rule=stop_rule, precision=prec)
    chk = grouped ? gcheck : check
    ll = Vector{Float64}[]
    try
        flags = UInt32((updateΣ ? 1 : 0) | (autoα ? 16 : 0))
        ζ = zeros(D * M); μ = zeros(MK); Elnϕ = similar(γ); ϕ = similar(γ); props = zeros(D * MK)
        elbo = Ref(0.0)
        GC.@preserve rowptr term count begin
            rp = [pointer(r) for r in rowptr]; tp = [pointer(t) for t in term]; cp = [pointer(c) for c in count]
            packed = verbose ? nothing : flatten_counts_packed(model.X, M)
            if !verbose && packed !== nothing
                # one call, 4-byte count records and page-locked buffers: what an end-to-end fit! of a large corpus is bound by
                # is the host -> device link (mmsig_mmctm_fit_host_packed)
                prow, prec = packed
                hist = zeros(M, maxiter); nit = Ref{Int32}(0); conv = Ref{Int32}(0)
                λp, νp = flat_rows_pinned(model.λ), flat_rows_pinned(model.ν)
                ζp = pinned_vector(Float64, D * M); pp = pinned_vector(Float64, D * MK)
                try
                    rpp = [pointer(r) for r in prow]; xp = [pointer(x) for x in prec]
                    if grouped
                        chk(h, ccall((:mmsig_group_mmctm_fit_host_packed, LIB), Int32,
                            (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt32}},
                             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                             Int32, Float64, UInt32, Ptr{Float64}, Ref{Int32}, Ref{Int32},
                             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                            h, D, M, K32, V32, rpp, xp, model.α, γ, λp, νp, model.μ, Σ, invΣ,
                            maxiter, tol, flags, hist, nit, conv, λp, νp, ζp, μ, Σ, invΣ, γ, Elnϕ, ϕ, pp))
                    else
                        chk(h, ccall((:mmsig_mmctm_fit_host_packed, LIB), Int32,
                            (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{UInt32}},
                             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                             Int32, Float64, UInt32, Ptr{Float64}, Ref{Int32}, Ref{Int32},
                             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                            h, D, D, M, K32, V32, rpp, xp, model.α, γ, λp, νp, model.μ, Σ, invΣ,
                            maxiter, tol, flags, hist, nit, conv, λp, νp, ζp, μ, Σ, invΣ, γ, Elnϕ, ϕ, pp))
                    end
                    λ .= λp; ν .= νp; ζ .= ζp; props .= pp
                finally
                    foreach(host_free, prow); foreach(host_free, prec)
                    host_free(λp); host_free(νp); host_free(ζp); host_free(pp)
                end
                ll = [hist[:, i] for i in 1:nit[]]
                model.converged = conv[] != 0
            elseif !verbose
                # one call: counts + state in, the whole loop of src/MMCTM.jl:462-489, state out, with
                # the host<->device copies pipelined behind the E-step (mmsig_mmctm_fit_host)
                hist = zeros(M, maxiter); nit = Ref{Int32}(0); conv = Ref{Int32}(0)
                if grouped
                    chk(h, ccall((:mmsig_group_mmctm_fit_host, LIB), Int32,
                        (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Int32, Float64, UInt32, Ptr{Float64}, Ref{Int32}, Ref{Int32},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        h, D, M, K32, V32, rp, tp, cp, model.α, γ, λ, ν, model.μ, Σ, invΣ,
                        maxiter, tol, flags, hist, nit, conv, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props))
                else
                    chk(h, ccall((:mmsig_mmctm_fit_host, LIB), Int32,
                        (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Int32, Float64, UInt32, Ptr{Float64}, Ref{Int32}, Ref{Int32},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        h, D, D, M, K32, V32, rp, tp, cp, model.α, γ, λ, ν, model.μ, Σ, invΣ,
                        maxiter, tol, flags, hist, nit, conv, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props))
                end
                ll = [hist[:, i] for i in 1:nit[]]
                model.converged = conv[] != 0
            else
                if grouped
                    chk(h, ccall((:mmsig_group_mmctm_set_data, LIB), Int32,
                        (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}}),
                        h, D, M, K32, V32, rp, tp, cp))
                    chk(h, ccall((:mmsig_group_mmctm_set_state, LIB), Int32,
                        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        h, model.α, γ, λ, ν, model.μ, Σ, invΣ))
                else
                    chk(h, ccall((:mmsig_mmctm_set_data, LIB), Int32,
                        (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}}),
                        h, D, D, M, K32, V32, rp, tp, cp))
                    chk(h, ccall((:mmsig_mmctm_set_state, LIB), Int32,
                        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        h, model.α, γ, λ, ν, model.μ, Σ, invΣ))
                end
                llbuf = zeros(M)
                for iter in 1:maxiter                                   # src/MMCTM.jl:462-489
                    if grouped
                        chk(h, ccall((:mmsig_group_mmctm_iterate, LIB), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float64}), h, flags, llbuf))
                    else
                        chk(h, ccall((:mmsig_mmctm_iterate, LIB), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float64}), h, flags, llbuf))
                    end
                    push!(ll, copy(llbuf))
                    println("$iter\tLog-likelihoods: ", join(ll[end], ", "))      # src/MMCTM.jl:482
                    if length(ll) > 10 && check_convergence(ll, tol=tol)
                        model.converged = true
                        break
                    end
                end
                if grouped
                    chk(h, ccall((:mmsig_group_mmctm_get_state, LIB), Int32,
                        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        h, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props))
                else
                    chk(h, ccall((:mmsig_mmctm_get_state, LIB), Int32,
                        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        h, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props))
                end
            end
        end
        if grouped
            chk(h, ccall((:mmsig_group_mmctm_elbo, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}), h, elbo, C_NULL))
        else
            chk(h, ccall((:mmsig_mmctm_elbo, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}), h, elbo, C_NULL))
        end
        scatter_state!(model, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props)
        materialize_θ && materialize_theta!(model, h, grouped)           # otherwise model.θ keeps its previous value
        if autoα && !grouped
            chk(h, ccall((:mmsig_mmctm_get_alpha, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h, model.α))
        elseif autoα
            h0 = ccall((:mmsig_group_member, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), h, 0)      # α is identical on every member
            check(h0, ccall((:mmsig_mmctm_get_alpha, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h0, model.α))
        end
        model.elbo = elbo[]
        model.ll = ll[end]
    finally
        grouped ? destroy_group(h) : destroy(h)
    end
    return ll
end

# "fit many models and pick the best one" (README.md:42; scripts/run_mmctm.jl:77-111): R restarts from the rows of
# γ0s (each a flat [m][k][v] table, e.g. flat_tables(MMCTM(K, α, X).γ)), on one GPU or dealt over `devices`
# (restart r on device r mod n, no communication).  The best restart's state (arg-max ELBO) is loaded into `model`;
# returns (elbo per restart, final log-likelihoods per restart, iterations per restart, index of the best).
function fit_restarts!(model::MMCTM, γ0s::Vector{Vector{Float64}}; maxiter=100, tol=1e-4, updateΣ=true,
                       device=0, devices=nothing, stop_rule=0)
    D, M, MK = model.D, model.M, sum(model.K)
    R = length(γ0s)
    rowptr, term, count = flatten_counts(model.X, M)
    K32, V32 = Int32.(model.K), Int32.(model.V)
    g0 = collect(reduce(vcat, γ0s))
    G = length(γ0s[1])
    grouped = devices !== nothing && length(devices) > 1
    h = grouped ? create_group(devices; stop_rule=stop_rule) :
                  create(device=(devices === nothing ? device : first(devices)), stop_rule=stop_rule)
    chk = grouped ? gcheck : check
    elbos = zeros(R); lls = zeros(M, R); nits = zeros(Int32, R); best = Ref{Int32}(-1)
    try
        flags = UInt32(updateΣ ? 1 : 0)
        λ = zeros(D * MK); ν = zeros(D * MK); ζ = zeros(D * M); μ = zeros(MK); Σ = zeros(MK * MK); invΣ = zeros(MK * MK)
        γ = zeros(G); Elnϕ = zeros(G); ϕ = zeros(G); props = zeros(D * MK)
        GC.@preserve rowptr term count begin
            rp = [pointer(r) for r in rowptr]; tp = [pointer(t) for t in term]; cp = [pointer(c) for c in count]
            if grouped
                chk(h, ccall((:mmsig_group_mmctm_restarts, LIB), Int32,
                    (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}}, Ptr{Float64},
                     Int32, Ptr{Float64}, Int32, Float64, UInt32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ref{Int32}),
                    h, D, M, K32, V32, rp, tp, cp, model.α, R, g0, maxiter, tol, flags, elbos, lls, nits, best))
                chk(h, ccall((:mmsig_group_mmctm_get_state, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props))
            else
                chk(h, ccall((:mmsig_mmctm_set_data, LIB), Int32,
                    (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}}),
                    h, D, D, M, K32, V32, rp, tp, cp))
                chk(h, ccall((:mmsig_mmctm_set_state, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h, model.α, g0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
                chk(h, ccall((:mmsig_mmctm_restarts, LIB), Int32,
                    (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Float64, UInt32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ref{Int32}),
                    h, R, g0, maxiter, tol, flags, elbos, lls, nits, best))
                chk(h, ccall((:mmsig_mmctm_get_state, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props))
            end
        end
        scatter_state!(model, λ, ν, ζ, μ, Σ, invΣ, γ, Elnϕ, ϕ, props)
        model.elbo = elbos[best[] + 1]
        model.ll = lls[:, best[] + 1]
    finally
        grouped ? destroy_group(h) : destroy(h)
    end
    return elbos, [lls[:, r] for r in 1:R], Int.(nits), Int(best[]) + 1
end

# ---- fit_heldout / transform (src/MMCTM.jl:554-586, :511-552): the same E-step with the topics and
# (unless fit_gaussian) the Gaussian prior frozen, selected by flags of mmsig_mmctm_iterate:
#   MMSIG_FLAG_UPDATE_SIGMA = 1, FREEZE_TOPICS = 2, FREEZE_MU = 4, UNSMOOTHED = 8
function frozen_loop!(newmodel::MMCTM, ϕsrc, flags::UInt32; maxiter, tol, verbose, device=0, stop_rule=0)
    D, M, MK = newmodel.D, newmodel.M, sum(newmodel.K)
    rowptr, term, count = flatten_counts(newmodel.X, M)
    K32, V32 = Int32.(newmodel.K), Int32.(newmodel.V)
    λ, ν = flat_rows(newmodel.λ), flat_rows(newmodel.ν)
    γ, ϕ = flat_tables(newmodel.γ), flat_tables(ϕsrc)
    Σ, invΣ = collect(transpose(newmodel.Σ)), collect(transpose(newmodel.invΣ))
    h = create(device=device, stop_rule=stop_rule)
    ll = Vector{Float64}[]
    try
        GC.@preserve rowptr term count begin
            rp = [pointer(r) for r in rowptr]; tp = [pointer(t) for t in term]; cp = [pointer(c) for c in count]
            check(h, ccall((:mmsig_mmctm_set_data, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}}),
                h, D, D, M, K32, V32, rp, tp, cp))
        end
        check(h, ccall((:mmsig_mmctm_set_state, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, newmodel.α, γ, λ, ν, newmodel.μ, Σ, invΣ))
        check(h, ccall((:mmsig_mmctm_set_phi, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h, ϕ))   # newmodel.ϕ = deepcopy(model.ϕ)
        llbuf = zeros(M)
        for iter in 1:maxiter
            check(h, ccall((:mmsig_mmctm_iterate, LIB), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float64}), h, flags, llbuf))
            push!(ll, copy(llbuf))
            verbose && println("$iter\tLog-likelihoods: ", join(ll[end], ", "))
            if length(ll) > 10 && check_convergence(ll, tol=tol)
                newmodel.converged = true
                break
            end
        end
        ζ = zeros(D * M); μ = zeros(MK); props = zeros(D * MK)
        check(h, ccall((:mmsig_mmctm_get_state, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, λ, ν, ζ, μ, Σ, invΣ, C_NULL, C_NULL, C_NULL, props))
        for d in 1:D
            newmodel.λ[d] .= @view λ[(d - 1) * MK + 1:d * MK]
            newmodel.ν[d] .= @view ν[(d - 1) * MK + 1:d * MK]
            newmodel.ζ[d] .= @view ζ[(d - 1) * M + 1:d * M]
            off = 0
            for m in 1:M
                newmodel.props[d][m] = props[(d - 1) * MK + off + 1:(d - 1) * MK + off + newmodel.K[m]]
                off += newmodel.K[m]
            end
        end
        newmodel.μ .= μ
        newmodel.Σ .= transpose(reshape(Σ, MK, MK)); newmodel.invΣ .= transpose(reshape(invΣ, MK, MK))
        newmodel.ll = ll[end]
    finally
        destroy(h)
    end
    return newmodel
end

function fit_heldout(Xheldout::Vector{Vector{Matrix{Int}}}, model::MMCTM; maxiter=100, verbose=false, device=0)
    heldout_model = MMCTM(model.K, model.α, Xheldout)                       # src/MMCTM.jl:557-563
    heldout_model.μ .= model.μ; heldout_model.Σ .= model.Σ; heldout_model.invΣ .= model.invΣ
    heldout_model.γ = deepcopy(model.γ); heldout_model.Elnϕ = deepcopy(model.Elnϕ); heldout_model.ϕ = deepcopy(model.ϕ)
    return frozen_loop!(heldout_model, model.ϕ, UInt32(2 | 4); maxiter=maxiter, tol=1e-4, verbose=verbose, device=device)
end

function transform(model::MMCTM, X::Vector{Vector{Matrix{Int}}}; maxiter=1000, tol=1e4, fit_gaussian=false,
                   verbose=false, device=0)
    newmodel = MMCTM(model.K, model.α, X)                                   # src/MMCTM.jl:514-520 (invΣ stays I, as there)
    newmodel.ϕ = deepcopy(model.ϕ)
    if !fit_gaussian
        newmodel.μ = deepcopy(model.μ); newmodel.Σ = deepcopy(model.Σ)
    end
    flags = UInt32(2 | 8 | (fit_gaussian ? 1 : 4))
    return frozen_loop!(newmodel, model.ϕ, flags; maxiter=maxiter, tol=tol, verbose=verbose, device=device)
end

# ---- fit!(::IMMCTM) (src/IMMCTM.jl:525-545): the MMCTM's device path over composite tables; the
# feature tables travel flat in the index order model.γ[m][k][i][j], model.α[m][i], and
# model.features[m] goes over as a V x I row-major block of 0-based values.
flat_ftables(t) = collect(reduce(vcat, [reduce(vcat, [reduce(vcat, tk) for tk in tm]) for tm in t]))   # [m][k][i][j]

function fit!(model::IMMCTM; maxiter=100, tol=1e-4, verbose=true, autoα=false, device=0, stop_rule=0)
    D, M, MK = model.D, model.M, sum(model.K)
    rowptr, term, count = flatten_counts(model.X, M)
    K32, V32, I32 = Int32.(model.K), Int32.(model.V), Int32.(model.I)
    feats = [Int32.(collect(transpose(model.features[m] .- 1))) for m in 1:M]      # I x V column-major = V x I row-major
    λ, ν = flat_rows(model.λ), flat_rows(model.ν)
    γ, α = flat_ftables(model.γ), collect(reduce(vcat, model.α))
    Σ, invΣ = collect(transpose(model.Σ)), collect(transpose(model.invΣ))
    h = create(device=device, stop_rule=stop_rule)
    ll = Vector{Float64}[]
    try
        GC.@preserve rowptr term count feats begin
            rp = [pointer(r) for r in rowptr]; tp = [pointer(t) for t in term]; cp = [pointer(c) for c in count]
            fp = [pointer(f) for f in feats]
            check(h, ccall((:mmsig_mmctm_set_data, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Int64}}, Ptr{Ptr{Int32}}, Ptr{Ptr{Int32}}),
                h, D, D, M, K32, V32, rp, tp, cp))
            check(h, ccall((:mmsig_immctm_set_features, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Ptr{Int32}}), h, I32, fp))
        end
        check(h, ccall((:mmsig_immctm_set_state, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, α, γ, λ, ν, model.μ, Σ, invΣ))
        flags = UInt32(1 | (autoα ? 16 : 0))                    # the IMMCTM's fit! always updates Σ (:533)
        llbuf = zeros(M)
        for iter in 1:maxiter
            check(h, ccall((:mmsig_mmctm_iterate, LIB), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float64}), h, flags, llbuf))
            push!(ll, copy(llbuf))
            verbose && println("$iter\tLog-likelihoods: ", join(ll[end], ", "))
            if length(ll) > 10 && check_convergence(ll, tol=tol)
                model.converged = true
                break
            end
        end
        elbo = Ref(0.0)
        check(h, ccall((:mmsig_mmctm_elbo, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}), h, elbo, C_NULL))
        ζ = zeros(D * M); μ = zeros(MK); Elnϕ = similar(γ)
        check(h, ccall((:mmsig_mmctm_get_state, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, λ, ν, ζ, μ, Σ, invΣ, C_NULL, C_NULL, C_NULL, C_NULL))
        check(h, ccall((:mmsig_immctm_get_tables, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h, γ, Elnϕ, α))
        for d in 1:D
            model.λ[d] .= @view λ[(d - 1) * MK + 1:d * MK]
            model.ν[d] .= @view ν[(d - 1) * MK + 1:d * MK]
            model.ζ[d] .= @view ζ[(d - 1) * M + 1:d * M]
        end
        model.μ .= μ
        model.Σ .= transpose(reshape(Σ, MK, MK)); model.invΣ .= transpose(reshape(invΣ, MK, MK))
        o = 0; a = 0
        for m in 1:M
            for k in 1:model.K[m], i in 1:model.I[m]
                r = (o + 1):(o + model.J[m][i])
                model.γ[m][k][i] .= @view γ[r]; model.Elnϕ[m][k][i] .= @view Elnϕ[r]
                o += model.J[m][i]
            end
            model.α[m] .= @view α[(a + 1):(a + model.I[m])]
            a += model.I[m]
        end
        model.elbo = elbo[]
        model.ll = ll[end]
    finally
        destroy(h)
    end
    return ll
end

function fit!(model::LDA; maxiter=1000, tol=1e-4, verbose=true, device=0, precision=:fp64)
    D, K, V = model.D, model.K, model.V
    rowptr = zeros(Int64, D + 1)
    for d in 1:D rowptr[d + 1] = rowptr[d] + size(model.X[d], 1) end
    term = Vector{Int32}(undef, rowptr[end]); count = Vector{Int32}(undef, rowptr[end])
    for d in 1:D
        r = (rowptr[d] + 1):rowptr[d + 1]
        term[r] .= model.X[d][:, 1] .- 1; count[r] .= model.X[d][:, 2]
    end
    λ = Float64.(vec(model.λ))            # V x K column-major == [k][v]
    h = create(device=device, precision=(precision == :fp32 ? 1 : 0))
    ll = Float64[]
    try
        check(h, ccall((:mmsig_lda_set_data, LIB), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}), h, D, D, K, V, rowptr, term, count))
        check(h, ccall((:mmsig_lda_set_state, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Ptr{Float64}, Ptr{Float64}),
            h, model.α, model.η, λ, C_NULL))
        l = Ref(0.0)
        for iter in 1:maxiter                                   # src/LDA.jl:201-219
            check(h, ccall((:mmsig_lda_iterate, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}), h, l))
            push!(ll, l[])
            verbose && println("$iter\tLog-likelihood: ", ll[end])
            if length(ll) > 10 && check_convergence(ll, tol=tol)
                model.converged = true
                break
            end
        end
        elbo = Ref(0.0)
        check(h, ccall((:mmsig_lda_elbo, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}), h, elbo, C_NULL))
        Elnβ = similar(λ); β = similar(λ); γ = zeros(K * D); Elnθ = zeros(K * D); θ = zeros(K * D)
        check(h, ccall((:mmsig_lda_get_state, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, λ, Elnβ, β, γ, Elnθ, θ))
        model.λ .= reshape(λ, V, K); model.Elnβ .= reshape(Elnβ, V, K); model.β .= reshape(β, V, K)
        model.γ .= reshape(γ, K, D); model.Elnθ .= reshape(Elnθ, K, D); model.θ .= reshape(θ, K, D)
        model.elbo = elbo[]
        model.ll = ll[end]
    finally
        destroy(h)
    end
    return ll
end

# ---- fit!(::ILDA) (src/ILDA.jl:233-259): the LDA's device path over composite tables; the M-step runs over
# the feature tables λ_i[j, k] (mmsig_ilda_*).  Tables travel flat as [k][i][j].
function fit!(model::ILDA; maxiter=1000, tol=1e-4, verbose=true, device=0)
    D, K, I, J = model.D, model.K, model.I, model.J
    V = size(model.features, 1)
    rowptr = zeros(Int64, D + 1)
    for d in 1:D rowptr[d + 1] = rowptr[d] + size(model.X[d], 1) end
    term = Vector{Int32}(undef, rowptr[end]); count = Vector{Int32}(undef, rowptr[end])
    for d in 1:D
        r = (rowptr[d] + 1):rowptr[d + 1]
        term[r] .= model.X[d][:, 1] .- 1; count[r] .= model.X[d][:, 2]
    end
    feat = Int32.(vec(permutedims(model.features .- 1)))       # V x I row-major, 0-based
    flat(t) = Float64[t[i][j, k] for k in 1:K for i in 1:I for j in 1:J[i]]
    function unflat!(t, x)
        o = 0
        for k in 1:K, i in 1:I, j in 1:J[i]
            t[i][j, k] = x[o += 1]
        end
    end
    λ = flat(model.λ)
    h = create(device=device)
    ll = Float64[]
    try
        check(h, ccall((:mmsig_lda_set_data, LIB), Int32,
            (Ptr{Cvoid}, Int64, Int64, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}), h, D, D, K, V, rowptr, term, count))
        check(h, ccall((:mmsig_ilda_set_features, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), h, I, feat))
        check(h, ccall((:mmsig_ilda_set_state, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, model.α, model.η, λ, C_NULL))
        l = Ref(0.0)
        for iter in 1:maxiter
            check(h, ccall((:mmsig_lda_iterate, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}), h, l))
            push!(ll, l[])
            verbose && println("$iter\tLog-likelihood: ", ll[end])
            if length(ll) > 10 && check_convergence(ll, tol=tol)
                model.converged = true
                break
            end
        end
        elbo = Ref(0.0)
        check(h, ccall((:mmsig_lda_elbo, LIB), Int32, (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}), h, elbo, C_NULL))
        Elnβ = similar(λ); γ = zeros(K * D); Elnθ = zeros(K * D); θ = zeros(K * D)
        check(h, ccall((:mmsig_ilda_get_tables, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, λ, Elnβ))
        check(h, ccall((:mmsig_lda_get_state, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, C_NULL, C_NULL, C_NULL, γ, Elnθ, θ))
        unflat!(model.λ, λ); unflat!(model.Elnβ, Elnβ)
        model.β = [model.λ[i] ./ sum(model.λ[i], dims=1) for i in 1:I]      # update_β!, src/ILDA.jl:127-129
        model.γ .= reshape(γ, K, D); model.Elnθ .= reshape(Elnθ, K, D)
        model.θ = reshape(θ, K, D)
        model.elbo = elbo[]
        model.ll = ll[end]
    finally
        destroy(h)
    end
    return ll
end

end # module
