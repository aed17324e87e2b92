# FeatureExtractionB200.jl -- thin Julia wrapper (modern Julia syntax) over the C ABI of
# include/fxb200.h.  It keeps the names and positional order of the reference's solvers
# (trthatcher/FeatureExtraction.jl: src/kpca.jl:49,64; src/klda.jl:57,71; export list
# src/FeatureExtraction.jl:7-16) and replaces everything under them with one ccall each.
#
# NOTE: Julia is not available in the build image, so this file has never been executed there;
# every call below is exercised through the identical C ABI by the Python ctypes binding
# (featureextraction.jl_b200/api.py) and the GPU parity tests.
module FeatureExtractionB200

export GaussianKernel, PolynomialKernel, SigmoidKernel, kpca, transform_kpca, klda, transform_klda,
       pca, transform_pca, lda, transform_lda, KPCAModel, fit_kpca, comm_unique_id,
       distributed_context, use_context!, peer_access

const libfx = get(ENV, "FXB200_LIB", "libfxb200.so")

# Own kernel structs (MLKernels 0.1 cannot be loaded by a modern Julia); constructor argument
# order as in test/test_kpca.jl:49.
abstract type Kernel end
struct GaussianKernel <: Kernel; alpha::Float64; end                       # exp(-alpha |x-y|^2)
struct PolynomialKernel <: Kernel; a::Float64; c::Float64; d::Float64; end  # (a x'y + c)^d
struct SigmoidKernel <: Kernel; a::Float64; c::Float64; end                 # tanh(a x'y + c)
kernel_id(::GaussianKernel) = Cint(0); kparams(k::GaussianKernel) = Float64[k.alpha, 0, 0]
kernel_id(::PolynomialKernel) = Cint(1); kparams(k::PolynomialKernel) = Float64[k.a, k.c, k.d]
kernel_id(::SigmoidKernel) = Cint(2); kparams(k::SigmoidKernel) = Float64[k.a, k.c, 0]

mutable struct Context
    h::Ptr{Cvoid}
    Context(h::Ptr{Cvoid}) = new(h)                     # adopt an existing handle (distributed_context)
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:fx_init, libfx), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r), "fx_init")
        c = new(r[]); finalizer(x -> ccall((:fx_shutdown, libfx), Cvoid, (Ptr{Cvoid},), x.h), c); c
    end
end

# Multi-GPU from ONE Julia process (the normal way): the library drives the listed devices itself
# (one host thread per GPU, NCCL and NVLink peer access set up inside); `kpca(X, κ; k=...)` stays a
# single call.   use_context!(multi_context([0, 1, 2, 3, 4, 5, 6, 7]))
function multi_context(devices::AbstractVector{<:Integer})
    devs = Cint.(collect(devices))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fx_init_multi, libfx), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), devs, length(devs), r),
          "fx_init_multi")
    c = Context(r[])
    finalizer(x -> ccall((:fx_shutdown, libfx), Cvoid, (Ptr{Cvoid},), x.h), c)
    c
end

# Multi-GPU alternative: one Julia process per GPU (Distributed / MPI.jl).  Rank 0 draws the communicator id,
# the host program ships the 128 bytes to every rank, every rank opens its context and installs it
# with `use_context!`; the same `kpca(X, κ; k=...)` call on every rank then returns the full (V, Λ).
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:fx_comm_unique_id, libfx), Cint, (Ptr{Cvoid},), id), "fx_comm_unique_id")
    id
end
function distributed_context(device::Integer, rank::Integer, world::Integer, id::Vector{UInt8})
    length(id) == 128 || throw(ArgumentError("the communicator id has 128 bytes"))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fx_init_distributed, libfx), Cint, (Cint, Cint, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                device, rank, world, id, r), "fx_init_distributed")
    c = Context(r[])
    finalizer(x -> ccall((:fx_shutdown, libfx), Cvoid, (Ptr{Cvoid},), x.h), c)
    c
end
use_context!(c::Context) = (default_ctx[] = c)
# true if the last collective fit shared the symmetric blocks of K between GPUs over NVLink
peer_access(c::Context) = ccall((:fx_ctx_peer_access, libfx), Cint, (Ptr{Cvoid},), c.h) != 0
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
ctx() = (default_ctx[] === nothing && (default_ctx[] = Context(0)); default_ctx[])

# info < 0 -> ArgumentError / DimensionMismatch (common.jl:87-89), info > 0 -> ErrorException
# (the analogue of LAPACKException, common.jl:90-92)
function check(info::Cint, what)
    info == 0 && return
    msg = unsafe_string(ccall((:fx_last_error, libfx), Cstring, ()))
    info < 0 && occursin("DimensionMismatch", msg) && throw(DimensionMismatch("$what: $msg"))
    info < 0 && throw(ArgumentError("$what: invalid argument #$(-info): $msg"))
    error("$what: fxb200 error $info: $msg")
end

mutable struct KPCAModel
    h::Ptr{Cvoid}; k::Int
    KPCAModel(h, k) = (m = new(h, k); finalizer(x -> ccall((:fx_model_free, libfx), Cvoid, (Ptr{Cvoid},), x.h), m); m)
end

"""
    kpca(X, κ, α=0, ϵ=0; k=size(X,1), center=true) -> (V, Λ)

Same contract as the reference `kpca` (src/kpca.jl:49-57); `k` and `center` are additions.
"""
function kpca(X::Matrix{Float64}, κ::Kernel, α::Float64 = 0.0, ϵ::Float64 = 0.0;
              k::Integer = size(X, 1), center::Bool = true)
    V, Λ, _ = fit_kpca(X, κ, α, ϵ; k = k, center = center, keep_model = false)
    (V, Λ)
end

function fit_kpca(X::Matrix{Float64}, κ::Kernel, α = 0.0, ϵ = 0.0; k = size(X, 1), center = true,
                  keep_model = true)
    n, p = size(X)
    V = Matrix{Float64}(undef, n, k); Λ = Vector{Float64}(undef, k)
    mh = Ref{Ptr{Cvoid}}(C_NULL)
    info = ccall((:fx_kpca_fit_f64, libfx), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Ptr{Float64}, Cint, Cint,
                  Float64, Float64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Ptr{Cvoid}},
                  Ptr{Cvoid}),
                 ctx().h, X, n, p, stride(X, 2), kernel_id(κ), kparams(κ), k, center, α, ϵ, C_NULL,
                 V, n, Λ, keep_model ? mh : C_NULL, C_NULL)
    check(info, "kpca")
    (V, Λ, keep_model ? KPCAModel(mh[], k) : nothing)
end

"Reference signature `transform_kpca(X, κ, W, Z)` (src/kpca.jl:64): stateless, rebuilds the model."
function transform_kpca(X::Matrix{Float64}, κ::Kernel, W::Matrix{Float64}, Z::Matrix{Float64};
                        center::Bool = true)
    size(W, 1) == size(X, 1) || throw(DimensionMismatch("W must have as many rows as X"))
    mh = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:fx_model_create_f64, libfx), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Cint,
                 Int64, Cint, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                ctx().h, X, size(X, 1), size(X, 2), stride(X, 2), kernel_id(κ), kparams(κ), W,
                size(W, 2), stride(W, 2), center, mh, C_NULL), "model_create")
    transform_kpca(KPCAModel(mh[], size(W, 2)), Z)
end

function transform_kpca(m::KPCAModel, Z::Matrix{Float64})
    out = Matrix{Float64}(undef, size(Z, 1), m.k)
    check(ccall((:fx_kpca_transform_f64, libfx), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Cvoid}),
                ctx().h, m.h, Z, size(Z, 1), stride(Z, 2), out, size(out, 1), C_NULL), "transform_kpca")
    out
end

"""
    klda(X, κ, y, k, frequency, α_b=0, ϵ_b=0, α_w=0, ϵ_w=0; ncomp=k-1) -> (A, Λ)

The LDA pipeline of src/klda.jl:34-61 applied to K = κ(X,X) with the pencil (Σ_b, Σ_w)
(`klda!` is exported upstream but never defined).  Labels are 1-based (klda.jl:8).
"""
function klda(X::Matrix{Float64}, κ::Kernel, y::Vector{<:Integer}, k::Integer, frequency::Vector{Float64},
              α_b = 0.0, ϵ_b = 0.0, α_w = 0.0, ϵ_w = 0.0; ncomp::Integer = k - 1)
    n, p = size(X)
    length(y) == n || throw(DimensionMismatch("X and y must have the same number of rows."))
    A = Matrix{Float64}(undef, n, ncomp); Λ = Vector{Float64}(undef, ncomp)
    y64 = convert(Vector{Int64}, y)
    check(ccall((:fx_klda_fit_f64, libfx), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Ptr{Float64}, Ptr{Int64}, Cint,
                 Ptr{Float64}, Float64, Float64, Float64, Float64, Cint, Ptr{Cvoid}, Ptr{Float64}, Int64,
                 Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx().h, X, n, p, stride(X, 2), kernel_id(κ), kparams(κ), y64, k, frequency, α_b, ϵ_b,
                α_w, ϵ_w, ncomp, C_NULL, A, n, Λ, C_NULL, C_NULL), "klda")
    (A, Λ)
end

"`pca(X, α, ϵ)` (src/kpca.jl:17-31) on the device kernels (p <= 512)."
function pca(X::Matrix{Float64}, α::Float64 = 0.0, ϵ::Float64 = 0.0)
    n, p = size(X)
    V = Matrix{Float64}(undef, p, p); Λ = Vector{Float64}(undef, p)
    check(ccall((:fx_pca_f64, libfx), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Float64, Float64, Ptr{Float64}, Int64,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
                ctx().h, X, n, p, stride(X, 2), α, ϵ, V, p, Λ, C_NULL, C_NULL), "pca")
    (V, Λ)
end

"`transform_pca(W, Z)` (src/kpca.jl:37) and `transform_lda(W, Z)` (src/klda.jl:65): Z * W."
function transform_pca(W::Matrix{Float64}, Z::Matrix{Float64})
    size(Z, 2) == size(W, 1) || throw(DimensionMismatch("Z and W do not conform"))
    out = Matrix{Float64}(undef, size(Z, 1), size(W, 2))
    check(ccall((:fx_project_f64, libfx), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64,
                 Ptr{Cvoid}),
                ctx().h, W, size(W, 1), size(W, 2), stride(W, 2), Z, size(Z, 1), stride(Z, 2), out,
                size(out, 1), C_NULL), "transform_pca")
    out
end
transform_lda(W::Matrix{Float64}, Z::Matrix{Float64}) = transform_pca(W, Z)

"`lda!(X, y, k, frequency, α_b, ϵ_b, α_w, ϵ_w)` (src/klda.jl:57-61), out of place, pencil (Σ_b, Σ_w)."
function lda(X::Matrix{Float64}, y::Vector{<:Integer}, k::Integer, frequency::Vector{Float64},
             α_b = 0.0, ϵ_b = 0.0, α_w = 0.0, ϵ_w = 0.0)
    n, p = size(X)
    length(y) == n || throw(DimensionMismatch("X and y must have the same number of rows."))
    V = Matrix{Float64}(undef, p, p); Λ = Vector{Float64}(undef, p)
    check(ccall((:fx_lda_f64, libfx), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Int64}, Cint, Ptr{Float64}, Float64, Float64,
                 Float64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Cvoid}),
                ctx().h, X, n, p, stride(X, 2), convert(Vector{Int64}, y), k, frequency, α_b, ϵ_b, α_w, ϵ_w,
                V, p, Λ, C_NULL), "lda")
    (V, Λ)
end

"Reference signature `transform_klda(X, κ, W, Z)` (src/klda.jl:71): κ(Z,X) * W, no centring."
transform_klda(X::Matrix{Float64}, κ::Kernel, W::Matrix{Float64}, Z::Matrix{Float64}) =
    transform_kpca(X, κ, W, Z; center = false)

end # module
