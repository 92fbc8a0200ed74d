# FEASTSolverB200.jl -- Julia shim over libfeast_cuda.so (include/feast_cuda.h).
#
# Keeps the reference's entry points, argument order, keyword names, defaults, in-place
# mutation of X and return tuples (src/feast.jl:3-156, src/nlfeast.jl:2-84, src/contour.jl),
# and replaces the inner blocks of the drivers by one `ccall` each.  The m0 x m0 reduced
# eigenproblem / SVD stays on LinearAlgebra (host LAPACK), exactly as upstream.
#
# Julia is not installed in the build environment, so this file is mechanical and has not been
# executed there; the Python twin (feastsolver_jl_b200/feast.py) is the tested client of the
# same ABI and the two are kept line-for-line parallel.
module FEASTSolverB200

using LinearAlgebra
using SparseArrays

export feast!, gen_feast!, dual_gen_feast!, nlfeast!, ifeast!, contour_estimate_eig
export in_contour, circular_contour_trapezoidal, circular_contour_gauss,
       rectangular_contour_gauss, rectangular_contour_trapezoidal, rational_func
export Contour, CircularContour, RectangularContour, CustomContour

const libfeast = get(ENV, "LIBFEAST_CUDA", "libfeast_cuda.so")

# ---------------------------------------------------------------- contour types (contour.jl:1-24)
abstract type Contour end
struct CircularContour <: Contour
    c::Number; r::Real; nodes::AbstractArray; weights::AbstractArray
end
struct RectangularContour <: Contour
    bottom_left::Complex; top_right::Complex; nodes::AbstractArray; weights::AbstractArray
    RectangularContour(bl, tr, n, w) = (real(bl) < real(tr) && imag(bl) < imag(tr)) ? new(bl, tr, n, w) : error("Invalid corners")
end
struct CustomContour <: Contour
    nodes::AbstractArray; weights::AbstractArray
end
Base.length(contour::Contour) = 1

function _ctor(sym, args, N, msg)
    z = Vector{ComplexF64}(undef, N); w = Vector{ComplexF64}(undef, N)
    rc = sym(args..., N, z, w)
    rc == -1 && error("Invalid corners")
    rc != 0 && error(msg)
    z, w
end
function circular_contour_trapezoidal(c, r, N=16)
    z, w = _ctor((a...) -> ccall((:feast_contour_circular_trapezoidal, libfeast), Cint,
                 (ComplexF64, Cdouble, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), a...), (ComplexF64(c), Float64(r)), N, "bad N")
    CircularContour(c, r, z, w)
end
function circular_contour_gauss(c, r, N=16)
    z, w = _ctor((a...) -> ccall((:feast_contour_circular_gauss, libfeast), Cint,
                 (ComplexF64, Cdouble, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), a...), (ComplexF64(c), Float64(r)), N,
                 "Number of nodes must be multiple of 2")
    CircularContour(c, r, z, w)
end
function rectangular_contour_gauss(bottom_left, top_right, N=16)
    z, w = _ctor((a...) -> ccall((:feast_contour_rectangular_gauss, libfeast), Cint,
                 (ComplexF64, ComplexF64, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), a...),
                 (ComplexF64(bottom_left), ComplexF64(top_right)), N, "Number of nodes must be multiple of 4")
    RectangularContour(complex(bottom_left), complex(top_right), z, w)
end
function rectangular_contour_trapezoidal(bottom_left, top_right, N=16)
    z, w = _ctor((a...) -> ccall((:feast_contour_rectangular_trapezoidal, libfeast), Cint,
                 (ComplexF64, ComplexF64, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), a...),
                 (ComplexF64(bottom_left), ComplexF64(top_right)), N, "Number of nodes must be multiple of 4")
    RectangularContour(complex(bottom_left), complex(top_right), z, w)
end
in_contour(λ, c::Number, r::Real) = abs.(λ .- c) .<= r
in_contour(λ, contour::CircularContour) = abs.(λ .- contour.c) .<= contour.r
in_contour(λ, contour::RectangularContour) =
    (real.(contour.bottom_left) .< real.(λ) .< real.(contour.top_right)) .& (imag.(contour.bottom_left) .< imag.(λ) .< imag.(contour.top_right))
function rational_func(z, contour)
    S = 0.0 + 0.0im
    for i = 1:size(contour.nodes, 1)
        S += contour.weights[i] / (contour.nodes[i] - z)
    end
    S
end

# ---------------------------------------------------------------- context handle
mutable struct FeastCtx
    h::Ptr{Cvoid}
    function FeastCtx(device::Integer=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:feast_ctx_create, libfeast), Cint, (Ref{Ptr{Cvoid}}, Cint), out, device)
        rc != 0 && error(unsafe_string(ccall((:feast_last_error, libfeast), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(out[])
        finalizer(c -> (c.h != C_NULL && ccall((:feast_ctx_destroy, libfeast), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
        ctx
    end
end
function _ck(ctx::FeastCtx, rc::Integer; allow=(0,))
    rc in allow && return rc
    error(unsafe_string(ccall((:feast_last_error, libfeast), Cstring, (Ptr{Cvoid},), ctx.h)))
end

# operator upload: dense Matrix (any eltype), SparseMatrixCSC (1-based Int64), Diagonal, UniformScaling
function _set_operator!(ctx, slot, M::AbstractMatrix, N)
    if M isa SparseMatrixCSC
        Tv = eltype(M) <: Complex ? ComplexF64 : Float64
        nz = convert(Vector{Tv}, M.nzval); cp = convert(Vector{Int64}, M.colptr); rv = convert(Vector{Int64}, M.rowval)
        _ck(ctx, ccall((:feast_set_csc, libfeast), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint),
                       ctx.h, slot, N, cp, rv, nz, Tv == ComplexF64, 1))
    elseif M isa Diagonal
        _set_operator!(ctx, slot, sparse(M), N)
    else
        Tv = eltype(M) <: Complex ? ComplexF64 : Float64
        D = convert(Matrix{Tv}, M)                      # integer A (runtests.jl:16) is converted here
        _ck(ctx, ccall((:feast_set_dense, libfeast), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Int64, Cint),
                       ctx.h, slot, N, D, stride(D, 2), Tv == ComplexF64))
    end
end
_set_operator!(ctx, slot, ::UniformScaling, N) =
    _ck(ctx, ccall((:feast_set_identity, libfeast), Cint, (Ptr{Cvoid}, Cint, Int64), ctx.h, slot, N))

function _check_plugins(factorizer, left_divider, mixed_prec)
    (factorizer !== lu || left_divider !== ldiv!) &&
        error("custom factorizer/left_divider callbacks cannot run inside libfeast_cuda (no CPU fallback)")
    mixed_prec && error("mixed_prec=true is not implemented in this build")
end

# ---------------------------------------------------------------- linear drivers
function _linear!(X, A, B, contour, iter, ϵ, debug, store, generalized; solver_kind=0, inner_tol=1e-8, unfiltered=false)
    N, m₀ = size(X)
    size(A, 1) != size(A, 2) && error("Incorrect dimensions of A, must be square")   # feast.jl:13
    size(A, 1) != N && error("Incorrect dimensions of X, must match A")                 # feast.jl:15
    ctx = FeastCtx()
    if generalized && (issparse(A) != issparse(B)) && !(B isa UniformScaling)
        A = Matrix(A); B = Matrix(B)                     # the library wants all-dense or all-sparse
    end
    _set_operator!(ctx, 0, A, N)
    generalized && _set_operator!(ctx, 1, B, N)
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, generalized ? 1 : 0, generalized ? 2 : 1))
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, length(z), z, w))
    _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, solver_kind, 0, inner_tol, 4000, store))
    Xc = convert(Matrix{ComplexF64}, X)
    _ck(ctx, ccall((:feast_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64), ctx.h, N, m₀, Xc, N))
    Λ, res = zeros(ComplexF64, m₀), zeros(m₀)
    Aq, Bq = zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀)
    for nit = 0:iter
        _ck(ctx, ccall((:feast_project, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                       ctx.h, Aq, generalized ? Bq : C_NULL))                       # feast.jl:41-43 / 117-121
        F = generalized ? eigen!(Aq, Bq) : eigen!(Aq)                                  # feast.jl:45 / 122 (host LAPACK)
        Λ .= F.values
        Xq = convert(Matrix{ComplexF64}, F.vectors)
        _ck(ctx, ccall((:feast_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}),
                       ctx.h, Xq, Λ, res))                                             # feast.jl:48-50 / 125-127
        contour_nonempty = reduce(|, in_contour(Λ, contour))
        if contour_nonempty && maximum(res[in_contour(Λ, contour)]) < ϵ
            debug && println("converged in $nit iteration")
            break
        end
        if nit < iter                                                                  # feast.jl:57-71
            _ck(ctx, ccall((:feast_contour_apply, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint, Ptr{Cvoid}),
                           ctx.h, Λ, 0, C_NULL); allow=(0, 2000))
        end
    end
    _ck(ctx, ccall((:feast_get_X, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64), ctx.h, Xc, N))
    X .= Xc
    finalize(ctx)
    unfiltered && return Λ, X, res
    inside = in_contour(Λ, contour)
    !reduce(|, inside) && println("no eigenvalues found in contour!")
    Λ[inside], X[:, inside], res[inside]
end

function feast!(X::AbstractMatrix, A::AbstractMatrix; nodes::Integer=8, iter::Integer=10, c=complex(0.0, 0.0), r=1.0, ϵ=1e-12,
                debug=false, store=false, mixed_prec=false, factorizer=lu, left_divider=ldiv!)
    feast!(X, A, circular_contour_trapezoidal(c, r, nodes); iter=iter, debug=debug, ϵ=ϵ, store=store,
           mixed_prec=mixed_prec, factorizer=factorizer, left_divider=left_divider)
end
function feast!(X::AbstractMatrix, A::AbstractMatrix, contour::Contour; iter::Integer=10, ϵ=1e-12, debug=false, store=false,
                mixed_prec=false, factorizer=lu, left_divider=ldiv!)
    _check_plugins(factorizer, left_divider, mixed_prec)
    _linear!(X, A, I, contour, iter, ϵ, debug, store, false)
end
function gen_feast!(X::AbstractMatrix, A::AbstractMatrix, B::AbstractMatrix; nodes::Integer=8, iter::Integer=10,
                    c=complex(0.0, 0.0), r=1.0, debug=false, store=false, ϵ=1e-12, factorizer=lu, left_divider=ldiv!)
    gen_feast!(X, A, B, circular_contour_trapezoidal(c, r, nodes); iter=iter, debug=debug, ϵ=ϵ, factorizer=factorizer, left_divider=ldiv!)
end
function gen_feast!(X::AbstractMatrix, A::AbstractMatrix, B::AbstractMatrix, contour::Contour; iter::Integer=10, debug=false,
                    store=false, ϵ=1e-12, factorizer=lu, left_divider=ldiv!)
    _check_plugins(factorizer, left_divider, false)
    _linear!(X, A, B, contour, iter, ϵ, debug, store, true)
end

# ---------------------------------------------------------------- two-sided driver (src/feast.jl:158-257)
function dual_gen_feast!(Xr::AbstractMatrix, Xl::AbstractMatrix, A::AbstractMatrix, B; nodes::Integer=8, iter::Integer=10,
                         c=complex(0.0, 0.0), r=1.0, debug=false, store=false, ϵ=1e-12, factorizer=lu, left_divider=ldiv!)
    dual_gen_feast!(Xr, Xl, A, B, circular_contour_trapezoidal(c, r, nodes); iter=iter, debug=debug, ϵ=ϵ, factorizer=factorizer, left_divider=ldiv!)
end
function dual_gen_feast!(Xr::AbstractMatrix, Xl::AbstractMatrix, A::AbstractMatrix, B, contour::Contour; iter::Integer=10,
                         debug=false, store=false, ϵ=1e-12, factorizer=lu, left_divider=ldiv!)
    _check_plugins(factorizer, left_divider, false)
    N, m₀ = size(Xl)
    size(A, 1) != size(A, 2) && error("Incorrect dimensions of A, must be square")
    size(A, 1) != N && error("Incorrect dimensions of X, must match A")
    ctx = FeastCtx()
    _set_operator!(ctx, 0, A, N); _set_operator!(ctx, 1, B, N)
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, 1, 2))
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, length(z), z, w))
    _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, 0, 0, 1e-8, 4000, store))
    Xrc, Xlc = convert(Matrix{ComplexF64}, Xr), convert(Matrix{ComplexF64}, Xl)
    _ck(ctx, ccall((:feast_dual_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64, Ptr{ComplexF64}, Int64),
                   ctx.h, N, m₀, Xrc, N, Xlc, N))
    Λ, resr = zeros(ComplexF64, m₀), zeros(m₀)
    G, Aq, Bq = zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀)
    for nit = 0:iter
        _ck(ctx, ccall((:feast_dual_project, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), ctx.h, G))          # feast.jl:199
        S = svd!(copy(G))
        sc = Diagonal(1.0 ./ sqrt.(S.S))                       # intended elementwise scaling (feast.jl:200-201,205)
        Mr = convert(Matrix{ComplexF64}, S.V * sc); Ml = convert(Matrix{ComplexF64}, S.U * sc)
        _ck(ctx, ccall((:feast_dual_rotate, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                       ctx.h, Mr, Ml, Aq, Bq))
        F = eigen(Aq, Bq); Λ .= F.values                                                                        # feast.jl:206
        Fl = eigen(Matrix(Aq'), Matrix(Bq'))                                                                     # feast.jl:210
        p = [argmin(abs.(Fl.values .- conj(l))) for l in Λ]
        Xqr = convert(Matrix{ComplexF64}, F.vectors); Xql = convert(Matrix{ComplexF64}, Fl.vectors[:, p])
        _ck(ctx, ccall((:feast_dual_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}),
                       ctx.h, Xqr, Xql, Λ, resr))
        contour_nonempty = reduce(|, in_contour(Λ, contour))
        if contour_nonempty && maximum(resr[in_contour(Λ, contour)]) < ϵ
            break
        end
        if nit < iter
            _ck(ctx, ccall((:feast_dual_contour_apply, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Cvoid}), ctx.h, Λ, C_NULL); allow=(0, 2000))
        end
    end
    _ck(ctx, ccall((:feast_dual_get, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Ptr{ComplexF64}, Int64), ctx.h, Xrc, N, Xlc, N))
    Xr .= Xrc; Xl .= Xlc
    finalize(ctx)
    inside = in_contour(Λ, contour)
    !reduce(|, inside) && println("no eigenvalues found in contour!")
    Λ[inside], Xr[:, inside], Xl[:, inside], resr[inside]
end

# ---------------------------------------------------------------- stochastic count estimate (src/stochastic.jl:2-33)
function contour_estimate_eig(A::AbstractMatrix, contour::Contour, B=I; samples::Integer=min(100, size(A, 1)), ϵ=1e-12,
                              debug=false, mixed_prec=false, factorizer=lu, left_divider=ldiv!)
    _check_plugins(factorizer, left_divider, mixed_prec)
    N = size(A, 1)
    X = randn(ComplexF64, N, samples)
    ctx = FeastCtx()
    _set_operator!(ctx, 0, A, N)
    B isa UniformScaling || _set_operator!(ctx, 1, B, N)
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, B isa UniformScaling ? 0 : 1, B isa UniformScaling ? 1 : 2))
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, length(z), z, w))
    _ck(ctx, ccall((:feast_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64), ctx.h, N, samples, X, N))
    est = Ref{Cdouble}(0.0)
    _ck(ctx, ccall((:feast_estimate_count, libfeast), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ptr{Cvoid}), ctx.h, est, C_NULL); allow=(0, 2000))
    finalize(ctx)
    est[]
end

# ifeast!(A, X0, nodes, iter; c, r, debug, ϵ)  (src/feast_experimental.jl:1-60): inexact inner solves, exactly `iter`
# passes, all m0 pairs returned; on the device the filter is applied in residual-inverse-iteration form by the Krylov path.
function ifeast!(A::AbstractMatrix, X₀::AbstractMatrix, nodes::Integer, iter::Integer;
                 c=complex(0.0, 0.0), r=1.0, debug=false, ϵ=0.05)
    issparse(A) || error("ifeast! on the B200 path needs a sparse A (Krylov inner solves)")
    X = convert(Matrix{ComplexF64}, deepcopy(X₀))
    _linear!(X, A, I, circular_contour_trapezoidal(c, r, nodes), iter, -1.0, debug, false, false;
             solver_kind=2, inner_tol=sqrt(eps(Float64)), unfiltered=true)
end

# ---------------------------------------------------------------- nonlinear driver (added method: coefficients)
function nlfeast!(T::AbstractVector{<:AbstractMatrix}, X::AbstractMatrix{ComplexF64}, nodes::Integer, iter::Integer;
                  c=complex(0.0, 0.0), r=1.0, debug=false, ϵ=10e-12, store=true, spurious=1e-5, factorizer=lu, left_divider=ldiv!)
    _check_plugins(factorizer, left_divider, false)
    N, m₀ = size(X)
    ctx = FeastCtx()
    for (i, Ai) in enumerate(T)
        _set_operator!(ctx, i - 1, Ai, N)
    end
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, 2, length(T)))
    contour = circular_contour_trapezoidal(c, r, nodes)                                 # nlfeast.jl:8
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, nodes, z, w))
    _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, 0, 0, 1e-8, 4000, store))
    _ck(ctx, ccall((:feast_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64), ctx.h, N, m₀, X, N))
    _ck(ctx, ccall((:feast_orthonormalize_X, libfeast), Cint, (Ptr{Cvoid},), ctx.h)) # nlfeast.jl:12-13
    Λ, res = zeros(ComplexF64, m₀), Array{Float64}(undef, m₀)
    Rf, G1 = zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀)
    for nit = 0:iter
        _ck(ctx, ccall((:feast_contour_apply, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint, Ptr{Cvoid}),
                       ctx.h, Λ, nit == 0, C_NULL); allow=(0, 2000))                   # nlfeast.jl:36-61
        _ck(ctx, ccall((:feast_beyn_reduce, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, Rf, G1))
        S = svd!(copy(Rf))                                                             # m0 x m0 part of utils.jl:70
        Am = (S.U' * G1) * S.V * Diagonal(1 ./ S.S)                                    # utils.jl:71-73
        F = eigen!(Am)                                                                 # utils.jl:74
        Λ .= F.values
        Xq = convert(Matrix{ComplexF64}, S.U * F.vectors)                              # utils.jl:75
        _ck(ctx, ccall((:feast_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}),
                       ctx.h, Xq, Λ, res))                                             # nlfeast.jl:66-67
        res_inside = res[in_contour.(Λ, c, r)]
        if size(res_inside, 1) > 0 && maximum(res_inside) < ϵ
            break
        end
        if nit > 1 && sum(res_inside .< spurious) > 0 && maximum(res_inside[res_inside .< spurious]) < ϵ
            break
        end
    end
    _ck(ctx, ccall((:feast_get_X, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64), ctx.h, X, N))
    finalize(ctx)
    Λ, X, res
end

# ---------------------------------------------------------------- nonlinear driver, reference signature (closure)
# nlfeast!(T::Function, X, nodes, iter; ...)  (src/nlfeast.jl:2-84): T is opaque, so it is evaluated HERE -- once per
# contour node (first pass only when store=true keeps the factorisations) and once per Ritz value for the residuals
# (src/utils.jl:107,154 evaluate T(λ_j) m0 times as well) -- and each sample is uploaded into slot 0 of a
# FEAST_PROBLEM_SAMPLED (= 3) problem; solves, accumulation, Beyn reduction and residual columns run on the device.
function _set_sample!(ctx, M::AbstractMatrix, N)
    if issparse(M)
        S = SparseMatrixCSC{eltype(M) <: Complex ? ComplexF64 : Float64, Int64}(M)
        _ck(ctx, ccall((:feast_set_sample_csc, libfeast), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint),
                       ctx.h, N, S.colptr, S.rowval, S.nzval, eltype(S) <: Complex, 1))
    else
        D = convert(Matrix{eltype(M) <: Complex ? ComplexF64 : Float64}, M)
        _ck(ctx, ccall((:feast_set_sample_dense, libfeast), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint),
                       ctx.h, N, D, N, eltype(D) <: Complex))
    end
end
function nlfeast!(T::Function, X::AbstractMatrix{ComplexF64}, nodes::Integer, iter::Integer;
                  c=complex(0.0, 0.0), r=1.0, debug=false, ϵ=10e-12, store=true, spurious=1e-5, factorizer=lu, left_divider=ldiv!)
    _check_plugins(factorizer, left_divider, false)
    N, m₀ = size(X)
    ctx = FeastCtx()
    contour = circular_contour_trapezoidal(c, r, nodes)                                 # nlfeast.jl:8
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _set_operator!(ctx, 0, T(z[1]), N)
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, 3, 1))
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, nodes, z, w))
    _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, 0, 0, 1e-8, 4000, store))
    _ck(ctx, ccall((:feast_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64), ctx.h, N, m₀, X, N))
    _ck(ctx, ccall((:feast_orthonormalize_X, libfeast), Cint, (Ptr{Cvoid},), ctx.h)) # nlfeast.jl:12-13
    Λ, res = zeros(ComplexF64, m₀), Array{Float64}(undef, m₀)
    Rf, G1 = zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀)
    for nit = 0:iter
        for k = 1:nodes                                                                # nlfeast.jl:36-61, node by node
            if ccall((:feast_node_needs_sample, libfeast), Cint, (Ptr{Cvoid}, Cint), ctx.h, k - 1) != 0
                _set_sample!(ctx, T(z[k]), N)
            end
            phase = (k == 1 ? 1 : 0) | (k == nodes ? 2 : 0)
            _ck(ctx, ccall((:feast_contour_node, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Cint, Cint, Ptr{Cvoid}),
                           ctx.h, k - 1, Λ, nit == 0, phase, C_NULL); allow=(0, 2000))
        end
        _ck(ctx, ccall((:feast_beyn_reduce, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, Rf, G1))
        S = svd!(copy(Rf))                                                             # m0 x m0 part of utils.jl:70
        Am = (S.U' * G1) * S.V * Diagonal(1 ./ S.S)                                    # utils.jl:71-73
        F = eigen!(Am)                                                                 # utils.jl:74
        Λ .= F.values
        Xq = convert(Matrix{ComplexF64}, S.U * F.vectors)                              # utils.jl:75
        _ck(ctx, ccall((:feast_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}),
                       ctx.h, Xq, Λ, res))                                             # X = Q Xq, normalize!
        for j = 1:m₀                                                                   # update_R! / residuals, utils.jl:104-109,151-157
            Tj = T(Λ[j])
            _set_sample!(ctx, Tj, N)
            rj = Ref{Cdouble}(0.0)
            _ck(ctx, ccall((:feast_sampled_residual, libfeast), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ref{Cdouble}), ctx.h, j - 1, norm(Tj), rj))
            res[j] = rj[]
        end
        res_inside = res[in_contour.(Λ, c, r)]
        if size(res_inside, 1) > 0 && maximum(res_inside) < ϵ
            break
        end
        if nit > 1 && sum(res_inside .< spurious) > 0 && maximum(res_inside[res_inside .< spurious]) < ϵ
            break
        end
    end
    _ck(ctx, ccall((:feast_get_X, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64), ctx.h, X, N))
    finalize(ctx)
    Λ, X, res
end

# ---------------------------------------------------------------- fine-grained plugin path
# Works with the UNMODIFIED reference drivers:  feast!(X, A; factorizer=B200Factorizer(ctx), left_divider=b200_ldiv!)
# (src/utils.jl:173-179: F = factorizer(C); left_divider(Y, F, X); finalize!(F)).
mutable struct B200Factor
    ctx::FeastCtx
    h::Ptr{Cvoid}
end
function b200_factorizer(C::AbstractMatrix)
    ctx = FeastCtx()
    N = size(C, 1)
    _set_operator!(ctx, 0, C, N)
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, 0, 1))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    coef = ComplexF64[1.0, 0.0]                                                        # C itself (no shift)
    _ck(ctx, ccall((:feast_factorize, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint, Ref{Ptr{Cvoid}}), ctx.h, coef, 2, out))
    B200Factor(ctx, out[])
end
function b200_ldiv!(Y::AbstractMatrix, F::B200Factor, X::AbstractMatrix)
    N, m = size(X)
    Xc = convert(Matrix{ComplexF64}, X); Yc = Matrix{ComplexF64}(undef, N, m)
    _ck(F.ctx, ccall((:feast_solve, libfeast), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64, Ptr{ComplexF64}, Int64, Cint),
                     F.ctx.h, F.h, N, m, Xc, N, Yc, N, 0); allow=(0, 2000))
    Y .= Yc
end
b200_finalize!(F::B200Factor) = (ccall((:feast_factor_free, libfeast), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), F.ctx.h, F.h); finalize(F.ctx))

end # module
