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

export feast!, gen_feast!, dual_gen_feast!, nlfeast!, nlfeast_it!, ifeast!, contour_estimate_eig, beyn, block_SS!, nlfeast_moments!
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

# nlfeast_it!(T, X, nodes, iter; c, r, debug, ϵ)  (src/nlfeast.jl:87-171): nlfeast with INEXACT inner solves -- relative
# tolerance 1e-3 in the first contour pass (:106), 1e-8 afterwards (:139) -- stopping when max(res[inside]) < ϵ (:164).
# Krylov inner solves over the tiled SpMM; the warm start `Tinv` of upstream (nodes x N x m0 of storage) is not kept: in
# residual-inverse-iteration form the right-hand side shrinks with the outer iteration, which plays the same role.
function nlfeast_it!(T::AbstractVector{<:AbstractMatrix}, X::AbstractMatrix{ComplexF64}, nodes::Integer, iter::Integer;
                     c=complex(0.0, 0.0), r=1.0, debug=false, ϵ=0.05)
    N, m₀ = size(X)
    ctx = FeastCtx()
    for (i, Ai) in enumerate(T)
        _set_operator!(ctx, i - 1, Ai, N)
    end
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, 2, length(T)))
    contour = circular_contour_trapezoidal(c, r, nodes)
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, nodes, z, w))
    _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, 2, 0, 1e-3, 4000, false))
    _ck(ctx, ccall((:feast_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64), ctx.h, N, m₀, X, N))
    _ck(ctx, ccall((:feast_orthonormalize_X, libfeast), Cint, (Ptr{Cvoid},), ctx.h))
    Λ, res = zeros(ComplexF64, m₀), Array{Float64}(undef, m₀)
    Rf, G1 = zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀)
    for nit = 0:iter
        nit == 1 && _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, 2, 0, 1e-8, 4000, false))
        _ck(ctx, ccall((:feast_contour_apply, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint, Ptr{Cvoid}),
                       ctx.h, Λ, nit == 0, C_NULL); allow=(0, 2000))
        _ck(ctx, ccall((:feast_beyn_reduce, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, Rf, G1))
        S = svd!(copy(Rf)); F = eigen!((S.U' * G1) * S.V * Diagonal(1 ./ S.S))
        Λ .= F.values
        Xq = convert(Matrix{ComplexF64}, S.U * F.vectors)
        _ck(ctx, ccall((:feast_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}), ctx.h, Xq, Λ, res))
        res_inside = res[in_contour.(Λ, c, r)]
        if nit >= 1 && size(res_inside, 1) > 0 && maximum(res_inside) < ϵ                 # nlfeast.jl:164
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

# ---------------------------------------------------------------- one-shot contour solvers and higher moments
# beyn / block_SS! / nlfeast_moments! (src/beyn.jl:2-94, src/nlfeast.jl:173-318) over the device moment accumulators
# S_p = sum_k w_k z_k^p (...)  (feast_set_moments: a p-loop in the accumulate kernel).  `T` is the coefficient list
# (device assembly); the closure form goes through the sampled-operator entries exactly as in nlfeast!(T::Function, ...).
function _nep_ctx(T::AbstractVector{<:AbstractMatrix}, X, nodes, c, r, store)
    N, m₀ = size(X)
    ctx = FeastCtx()
    for (i, Ai) in enumerate(T)
        _set_operator!(ctx, i - 1, Ai, N)
    end
    _ck(ctx, ccall((:feast_set_problem, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint), ctx.h, 2, length(T)))
    contour = circular_contour_trapezoidal(c, r, nodes)
    z = convert(Vector{ComplexF64}, contour.nodes); w = convert(Vector{ComplexF64}, contour.weights)
    _ck(ctx, ccall((:feast_set_contour, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, nodes, z, w))
    _ck(ctx, ccall((:feast_set_solver, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Cint), ctx.h, 0, 0, 1e-8, 4000, store))
    Xc = convert(Matrix{ComplexF64}, X)
    _ck(ctx, ccall((:feast_set_subspace, libfeast), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{ComplexF64}, Int64), ctx.h, N, m₀, Xc, N))
    ctx
end
_moments!(ctx, k) = _ck(ctx, ccall((:feast_set_moments, libfeast), Cint, (Ptr{Cvoid}, Cint), ctx.h, k))
_pass!(ctx, Λ, first) = _ck(ctx, ccall((:feast_contour_apply, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint, Ptr{Cvoid}),
                                       ctx.h, Λ, first, C_NULL); allow=(0, 2000))
function _gram(ctx, a, b, m₀)
    G = zeros(ComplexF64, m₀, m₀)
    _ck(ctx, ccall((:feast_block_gram, libfeast), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}), ctx.h, a, b, G))
    G
end
function _combine_residual!(ctx, W, λ, N, m₀)          # X = [S_0 .. S_{k-1}] W ; normalise ; relative residuals
    _ck(ctx, ccall((:feast_moment_combine, libfeast), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Int64), ctx.h, size(W, 1) ÷ m₀, W, size(W, 1)))
    res = Array{Float64}(undef, m₀)
    _ck(ctx, ccall((:feast_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}), ctx.h, C_NULL, λ, res))
    res
end
function _getX(ctx, N, m₀)
    X = Matrix{ComplexF64}(undef, N, m₀)
    _ck(ctx, ccall((:feast_get_X, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64), ctx.h, X, N))
    X
end

function beyn(T::AbstractVector{<:AbstractMatrix}, A::AbstractMatrix, X::AbstractMatrix, nodes::Integer; c=complex(0.0, 0.0), r=1.0)
    N, m₀ = size(X)
    size(A, 1) != size(A, 2) && error("Incorrect dimensions of A, must be square")
    size(A, 1) != N && error("Incorrect dimensions of X₀, must match A")
    ctx = _nep_ctx(T, X, nodes, c, r, false)
    _moments!(ctx, 2)
    _pass!(ctx, zeros(ComplexF64, m₀), 1)                                               # beyn.jl:16-21
    Rf, G1 = zeros(ComplexF64, m₀, m₀), zeros(ComplexF64, m₀, m₀)
    _ck(ctx, ccall((:feast_beyn_reduce, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx.h, Rf, G1))
    S = svd!(copy(Rf)); F = eigen!((S.U' * G1) * S.V * Diagonal(1 ./ S.S))              # beyn.jl:22-24
    Λ = convert(Vector{ComplexF64}, F.values); Xq = convert(Matrix{ComplexF64}, S.U * F.vectors)
    res = Array{Float64}(undef, m₀); fro = Array{Float64}(undef, m₀)
    _ck(ctx, ccall((:feast_recover_residual, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}), ctx.h, Xq, Λ, res))
    _ck(ctx, ccall((:feast_last_fro, libfeast), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.h, fro))
    Xn = _getX(ctx, N, m₀); finalize(ctx)
    res .*= fro                                                                         # absolute residual, beyn.jl:28
    p = sortperm(res)
    Λ[p], Xn[:, p], res[p]
end

function block_SS!(T::AbstractVector{<:AbstractMatrix}, X::AbstractMatrix{ComplexF64}, nodes=2^4, moments=2;
                   c=complex(0.0, 0.0), r=1.0, debug=false, Y=rand(ComplexF64, size(X)...))
    N, m₀ = size(X); K = moments * m₀
    ctx = _nep_ctx(T, X, nodes, c, r, false)
    _ck(ctx, ccall((:feast_orthonormalize_X, libfeast), Cint, (Ptr{Cvoid},), ctx.h))    # beyn.jl:41
    _moments!(ctx, 2 * moments + 1)
    _pass!(ctx, zeros(ComplexF64, m₀), 1)                                               # beyn.jl:50-56
    _ck(ctx, ccall((:feast_set_X, libfeast), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Int64), ctx.h, convert(Matrix{ComplexF64}, Y), N))
    G = [_gram(ctx, -1, p, m₀) for p = 0:2*moments]                                     # Y' * S_p
    Q₀, Q₁ = zeros(ComplexF64, m₀ * moments, K), zeros(ComplexF64, m₀ * moments, K)
    for i = 1:moments, j = 1:moments
        Q₀[(i-1)*m₀+1:i*m₀, (j-1)*m₀+1:j*m₀] .= G[i+j]                                  # S_{i+j-1}, beyn.jl:66
        Q₁[(i-1)*m₀+1:i*m₀, (j-1)*m₀+1:j*m₀] .= G[i+j+1]                                # S_{i+j},   beyn.jl:67
    end
    V = svd(Q₀); n = min(count(V.S / V.S[1] .> 1e-13), K)                               # beyn.jl:77-78
    Λ, Xq = eigen!(V.U[:, 1:n]' * Q₁ * V.V[:, 1:n], V.U[:, 1:n]' * Q₀ * V.V[:, 1:n])     # beyn.jl:80-83
    W = V.V[:, 1:n] * Xq
    Xn = Matrix{ComplexF64}(undef, N, n); res = zeros(n)
    for c0 = 1:m₀:n                                                                     # the device block is m0 wide
        c1 = min(n, c0 + m₀ - 1); k = c1 - c0 + 1
        Wc = repeat(W[:, c0:c0], 1, m₀); Wc[:, 1:k] .= W[:, c0:c1]
        λc = fill(ComplexF64(Λ[c0]), m₀); λc[1:k] .= Λ[c0:c1]
        rc = _combine_residual!(ctx, convert(Matrix{ComplexF64}, Wc), λc, N, m₀)        # beyn.jl:87-93
        Xn[:, c0:c1] .= _getX(ctx, N, m₀)[:, 1:k]; res[c0:c1] .= rc[1:k]
    end
    finalize(ctx)
    Λ, Xn, res
end

# nlfeast_moments!: the tall SVD of the block-Hankel matrix is formed from m0 x m0 Gram blocks S_a' S_b (see DESIGN.md);
# directions below 1e-7 of the largest singular value are dropped.
function nlfeast_moments!(T::AbstractVector{<:AbstractMatrix}, X::AbstractMatrix{ComplexF64}, nodes::Integer, iter::Integer;
                          c=complex(0.0, 0.0), r=1.0, debug=false, ϵ=10e-12, moments=2, store=true, spurious=1e-5)
    N, m₀ = size(X); M = moments; K = M * m₀
    ctx = _nep_ctx(T, X, nodes, c, r, store)
    _moments!(ctx, 2M)
    Λ = ComplexF64[]; res = Float64[]; W = zeros(ComplexF64, K, 0); λx = zeros(ComplexF64, m₀)
    function reduce!()
        g(a, b) = a <= b ? _gram(ctx, a, b, m₀) : Matrix(_gram(ctx, b, a, m₀)')
        G0, G01 = zeros(ComplexF64, K, K), zeros(ComplexF64, K, K)
        for j = 0:M-1, jp = 0:M-1
            G0[j*m₀+1:(j+1)*m₀, jp*m₀+1:(jp+1)*m₀] .= sum(g(i + j, i + jp) for i = 0:M-1)
            G01[j*m₀+1:(j+1)*m₀, jp*m₀+1:(jp+1)*m₀] .= sum(g(i + j, i + jp + 1) for i = 0:M-1)
        end
        E = eigen(Hermitian((G0 + G0') / 2)); ev = reverse(E.values); Vv = E.vectors[:, end:-1:1]
        keep = ev .> 1e-14 * ev[1]; sv = sqrt.(ev[keep]); Vv = Vv[:, keep]
        F = eigen!(Diagonal(1 ./ sv) * (Vv' * G01 * Vv) * Diagonal(1 ./ sv))             # nlfeast.jl:222-225
        Wn = Vv * Diagonal(1 ./ sv) * F.vectors; λ = convert(Vector{ComplexF64}, F.values); nk = length(λ)
        rs = zeros(nk)
        for c0 = 1:m₀:nk
            c1 = min(nk, c0 + m₀ - 1); k = c1 - c0 + 1
            Wc = repeat(Wn[:, c0:c0], 1, m₀); Wc[:, 1:k] .= Wn[:, c0:c1]
            λc = fill(λ[c0], m₀); λc[1:k] .= λ[c0:c1]
            rs[c0:c1] .= _combine_residual!(ctx, convert(Matrix{ComplexF64}, Wc), λc, N, m₀)[1:k]
        end
        p = sortperm(rs); Λ = λ[p]; res = rs[p]; W = Wn[:, p]                            # utils.jl:125-133
        nb = min(m₀, nk); Wc = repeat(W[:, 1:1], 1, m₀); Wc[:, 1:nb] .= W[:, 1:nb]
        λx = fill(Λ[1], m₀); λx[1:nb] .= Λ[1:nb]
        _combine_residual!(ctx, convert(Matrix{ComplexF64}, Wc), λx, N, m₀)             # X = Y[:, 1:m0] and its R
    end
    _pass!(ctx, zeros(ComplexF64, m₀), 1); reduce!()                                    # nlfeast.jl:195-236
    for nit = 1:iter
        _pass!(ctx, λx, 0); reduce!()                                                   # nlfeast.jl:255-283
        nb = min(m₀, length(Λ))
        res_inside = res[1:nb][in_contour.(Λ[1:nb], c, r)]
        if size(res_inside, 1) > 0 && maximum(res_inside) < ϵ
            break
        end
        if nit > 1 && sum(res_inside .< spurious) > 0 && maximum(res_inside[res_inside .< spurious]) < ϵ
            break
        end
    end
    X .= _getX(ctx, N, m₀)
    Yall = Matrix{ComplexF64}(undef, N, length(Λ))
    for c0 = 1:m₀:length(Λ)
        c1 = min(length(Λ), c0 + m₀ - 1); k = c1 - c0 + 1
        Wc = repeat(W[:, c0:c0], 1, m₀); Wc[:, 1:k] .= W[:, c0:c1]
        λc = fill(Λ[c0], m₀); λc[1:k] .= Λ[c0:c1]
        _combine_residual!(ctx, convert(Matrix{ComplexF64}, Wc), λc, N, m₀)
        Yall[:, c0:c1] .= _getX(ctx, N, m₀)[:, 1:k]
    end
    finalize(ctx)
    Λ, Yall, res
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
