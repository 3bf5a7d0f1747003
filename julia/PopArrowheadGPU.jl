# PopArrowheadGPU.jl -- Julia-side bindings of libpop_arrowhead.so (include/pop_arrowhead.h).
#
# Drop-in for the arrowhead hot path of PiecewiseOrthogonalPolynomials.jl: the user-visible calls
#   M = C'C, Δ = weaklaplacian(C), M[KR,KR], reversecholesky(Symmetric(S)).L/.U, F \ X, X / F, A*X, X*A
# keep their signatures; the methods below re-bind the reference's own dispatch points for Float64:
#   materialize!(::MatMulVecAdd/MatMulMatAdd{<:ArrowheadLayouts})      src/arrowhead.jl:191,214,237
#   reversecholesky_layout!(::ArrowheadLayouts, A, ::Type{UpperTriangular})  src/arrowhead.jl:276
#   materialize!(::MatLdivVec{<:TriangularLayout{...,ArrowheadLayout}})  src/arrowhead.jl:425-486
# plus a NEW MatLdivMat-style method so a matrix right-hand side goes down in one call (the reference
# falls back to a column loop), the ADI / heat drivers of examples/ as single calls, and `DevicePanel`,
# a matrix that stays in device memory between calls.
# NOT EXECUTED in the build container (no Julia there); it is the binding a maintainer adds, kept
# declarative.  `include` it after `using PiecewiseOrthogonalPolynomials`.
module PopArrowheadGPU

using LinearAlgebra, BandedMatrices, ArrayLayouts, MatrixFactorizations
using PiecewiseOrthogonalPolynomials
using PiecewiseOrthogonalPolynomials.BBBArrowheadMatrices: BBBArrowheadMatrix, ArrowheadLayouts, ArrowheadLayout
import ArrayLayouts: materialize!, MatMulVecAdd, MatMulMatAdd, MatLdivVec, TriangularLayout
import MatrixFactorizations: reversecholesky_layout!

const libpop = get(ENV, "POP_ARROWHEAD_LIB", "libpop_arrowhead.so")

struct PopDesc            # mirrors `pop_desc`
    m::Int32; nh::Int32; q::Int32; nB::Int32; nC::Int32; lD::Int32; uD::Int32; uniform::Int32
end

const POP_LEFT, POP_RIGHT = Cint(0), Cint(1)
const POP_GENERAL, POP_SYM_UPPER = Cint(0), Cint(1)
const POP_UPPER, POP_LOWER = Cint(0), Cint(1)

function check(rc::Integer, what)
    rc == 0 && return
    msg = unsafe_string(ccall((:pop_last_error, libpop), Cstring, ()))
    rc > 0 && throw(PosDefException(rc))                        # block factor: non-positive pivot
    rc == -2 && throw(DimensionMismatch("$what: $msg"))
    error("$what: $msg (code $rc)")
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
    if CTX[] == C_NULL
        check(ccall((:pop_create, libpop), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cvoid}), CTX, 0, C_NULL), "pop_create")
    end
    CTX[]
end

# ---- packing: anything with the fields A, B, C, D of a BBBArrowheadMatrix -> flat arrays in the ABI layout ----
# Works on the matrix itself and on its Symmetric / Triangular views: their `getproperty` (src/arrowhead.jl:154-167,
# :489-521) already presents the blocks the view stands for (Symmetric: C = adjoint.(B), D = symmetric.(D); Upper: C = ()).
# ABI (include/pop_arrowhead.h): B[b][t][k] = entry (hat k+t-1, bubble (b,k)) with 0-based k and hat index, i.e. with
# Julia's 1-based k the hat of slot t ∈ 0:2 is i = k + t - 1 FOR BOTH BASES (ContinuousPolynomial{1}: rows k, k+1 ->
# slots 1, 2; DirichletPolynomial: rows k-1, k -> slots 0, 1); C likewise transposed; D[lD+d][j][k] = D_k[j, j+d].
function pack(A)
    nh, m = size(A.A, 1), length(A.D)
    q = size(A.D[1], 1); lD, uD = bandwidths(A.D[1])
    Ad = Float64[A.A[i, i] for i in 1:nh]
    Au = Float64[A.A[i, i+1] for i in 1:nh-1]; Al = Float64[A.A[i+1, i] for i in 1:nh-1]
    slots(Bs, tr) = begin
        out = zeros(m, 3, length(Bs))                            # memory order [k][t][b] = B[b][t][k] of the ABI
        for (b, Bb) in enumerate(Bs), k in 1:m, t in 0:2
            i = k + t - 1                                        # hat (1-based) of slot t
            1 <= i <= nh && (out[k, t+1, b] = tr ? Bb[k, i] : Bb[i, k])
        end
        out
    end
    Bp, Cp = slots(A.B, false), slots(A.C, true)
    Dp = zeros(m, q, lD + uD + 1)
    for k in 1:m, j in 1:q, d in -lD:uD
        1 <= j + d <= q && (Dp[k, j, lD+d+1] = A.D[k][j, j+d])
    end
    PopDesc(m, nh, q, length(A.B), length(A.C), lD, uD, 0), Ad, Au, Al, Bp, Cp, Dp
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(p) h = new(p); finalizer(x -> ccall((:pop_arrowhead_free, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.ptr), h) end
end

"Upload a BBBArrowheadMatrix or one of its Symmetric / Triangular views as a general packed arrowhead."
function upload(A)
    d, Ad, Au, Al, Bp, Cp, Dp = pack(A)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Ad Au Al Bp Cp Dp check(ccall((:pop_arrowhead_upload, libpop), Cint,
        (Ptr{Cvoid}, Ref{PopDesc}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
        ctx(), d, Ad, Au, Al, Bp, Cp, Dp, out), "pop_arrowhead_upload")
    Handle(out[])
end
# Symmetric(P): upload the parent once and let the library read its upper data (POP_SYM_UPPER), as the reference's
# getproperty does -- half the bytes of the general form
upload_sym(S::Symmetric{<:Any,<:BBBArrowheadMatrix}) = (upload(parent(S)), POP_SYM_UPPER)
upload_sym(A) = (upload(A), POP_GENERAL)

# ---- device-resident panels ------------------------------------------------------------------------------
"An n x r Float64 matrix in device memory (column-major, even leading dimension).  `DevicePanel(X)` uploads a host
matrix, `Array(P)` / `copyto!(X, P)` bring it back; ldiv!, rdiv!, mul!, adi_poisson and heat_steps! accept it, so a
chain of operations moves data over PCIe twice instead of twice per call."
mutable struct DevicePanel
    ptr::Ptr{Float64}; ld::Int64; rows::Int64; cols::Int64
    function DevicePanel(rows::Integer, cols::Integer)
        p, ld = Ref{Ptr{Float64}}(C_NULL), Ref{Int64}(0)
        check(ccall((:pop_panel_alloc, libpop), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Float64}}, Ref{Int64}), ctx(), rows, cols, p, ld), "DevicePanel")
        P = new(p[], ld[], rows, cols)
        finalizer(x -> ccall((:pop_panel_free, libpop), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx(), x.ptr), P)
    end
end
Base.size(P::DevicePanel) = (Int(P.rows), Int(P.cols))
function Base.copyto!(P::DevicePanel, X::StridedMatrix{Float64})
    size(X) == size(P) || throw(DimensionMismatch("DevicePanel is $(size(P)), matrix is $(size(X))"))
    GC.@preserve X check(ccall((:pop_panel_upload, libpop), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64),
        ctx(), X, stride(X, 2), P.ptr, P.ld, P.rows, P.cols), "upload")
    P
end
function Base.copyto!(X::StridedMatrix{Float64}, P::DevicePanel)
    size(X) == size(P) || throw(DimensionMismatch("DevicePanel is $(size(P)), matrix is $(size(X))"))
    GC.@preserve X check(ccall((:pop_panel_download, libpop), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64),
        ctx(), P.ptr, P.ld, X, stride(X, 2), P.rows, P.cols), "download")
    X
end
DevicePanel(X::StridedMatrix{Float64}) = copyto!(DevicePanel(size(X)...), X)
Base.Array(P::DevicePanel) = copyto!(Matrix{Float64}(undef, size(P)...), P)

# ---- products: src/arrowhead.jl:191-258 --------------------------------------------------------------
function materialize!(M::MatMulMatAdd{<:ArrowheadLayouts,<:AbstractColumnMajor,<:AbstractColumnMajor,Float64})
    α, A, X, β, Y = M.α, M.A, M.B, M.β, M.C
    h, sym = upload_sym(A)
    GC.@preserve X Y check(ccall((:pop_mul_host, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}, Int64, Int64),
        ctx(), h.ptr, POP_LEFT, sym, 0, α, X, stride(X, 2), β, Y, stride(Y, 2), size(X, 2)), "mul!")
    Y
end
function materialize!(M::MatMulVecAdd{<:ArrowheadLayouts,<:AbstractStridedLayout,<:AbstractStridedLayout,Float64})
    materialize!(MulAdd(M.α, M.A, reshape(M.B, :, 1), M.β, reshape(M.C, :, 1))); M.C
end
function materialize!(M::MatMulMatAdd{<:AbstractColumnMajor,<:ArrowheadLayouts,<:AbstractColumnMajor,Float64})   # X*A
    α, X, A, β, Y = M.α, M.A, M.B, M.β, M.C
    h, sym = upload_sym(A)
    GC.@preserve X Y check(ccall((:pop_mul_host, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}, Int64, Int64),
        ctx(), h.ptr, POP_RIGHT, sym, 0, α, X, stride(X, 2), β, Y, stride(Y, 2), size(X, 1)), "mul!")
    Y
end
"Y ← α A X + β Y (side = POP_LEFT) or α X A + β Y (POP_RIGHT) on device-resident panels; `h` from `upload_sym(A)`."
function mul_panel!(Y::DevicePanel, h::Handle, sym::Cint, X::DevicePanel, side::Cint = POP_LEFT, α = 1.0, β = 0.0)
    check(ccall((:pop_mul, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}, Int64, Int64),
        ctx(), h.ptr, side, sym, 0, α, X.ptr, X.ld, β, Y.ptr, Y.ld, side == POP_LEFT ? X.cols : X.rows), "mul!")
    Y
end

# ---- reverse Cholesky: src/arrowhead.jl:276-292 -----------------------------------------------------
# factor on the device, copy the O(n) factor back into A's own storage so .U, getindex and U*U' ≈ S
# (test/test_arrowhead.jl:145) keep working; the device handle is cached for the solves.  BBBArrowheadMatrix is an
# immutable struct, so the cache is keyed by the first of its fields that has an identity (the Vector of D blocks after
# `reversecholcopy`, src/arrowhead.jl:266-272); a matrix with no such field is simply not cached.
const FACTORS = WeakKeyDict{Any,Handle}()
function cachekey(A::BBBArrowheadMatrix)
    for f in (A.D, A.A, parent(A.A))
        ismutable(f) && return f
    end
    nothing
end
function reversecholesky_layout!(::ArrowheadLayouts, A::BBBArrowheadMatrix{Float64}, ::Type{UpperTriangular})
    h = upload(A)
    check(ccall((:pop_revchol_inplace, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), h.ptr), "reversecholesky!")
    d, Ad, Au, Al, Bp, Cp, Dp = pack(A)
    GC.@preserve Ad Au Bp Dp check(ccall((:pop_arrowhead_download, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx(), h.ptr, 1, Ad, Au, C_NULL, Bp, C_NULL, Dp), "factor download")
    unpack_upper!(A, Ad, Au, Bp, Dp)                             # writes A.A, A.B[b], A.D[k] upper data
    k = cachekey(A)
    k === nothing || (FACTORS[k] = h)
    UpperTriangular(A), 0
end

"Device handle of the factor behind `F`: the cached one, or -- for a factor computed elsewhere, e.g. by the reference's
own CPU code -- an upload of `F.U` marked as a factor (pop_arrowhead_as_factor; it is NOT factored again)."
function factor_handle(F::ReverseCholesky{Float64})
    A = parent(F.U)::BBBArrowheadMatrix
    k = cachekey(A)
    k !== nothing && haskey(FACTORS, k) && return FACTORS[k]
    h = upload(UpperTriangular(A))                               # upper data only: C = ()
    check(ccall((:pop_arrowhead_as_factor, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), h.ptr), "factor upload")
    k === nothing || (FACTORS[k] = h)
    h
end

# ---- F \ X and X / F: one call for the whole panel (test/test_arrowhead.jl:130, adi_poisson.jl:54-55) --
# (MatrixFactorizations stores either the matrix itself or its UpperTriangular view as `factors`, depending on the version:
# both are bound; `F.U` is an UpperTriangular of the BBBArrowheadMatrix in either case)
const RevCholArrowhead = Union{ReverseCholesky{Float64,<:BBBArrowheadMatrix},
                               ReverseCholesky{Float64,<:UpperTriangular{Float64,<:BBBArrowheadMatrix}}}
function LinearAlgebra.ldiv!(F::RevCholArrowhead, X::StridedVecOrMat{Float64})
    h = factor_handle(F)
    GC.@preserve X check(ccall((:pop_solve_host, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), h.ptr, POP_LEFT, X, X isa AbstractVector ? length(X) : stride(X, 2), size(X, 2), 0), "ldiv!")
    X
end
function LinearAlgebra.rdiv!(X::StridedMatrix{Float64}, F::RevCholArrowhead)
    h = factor_handle(F)
    GC.@preserve X check(ccall((:pop_solve_host, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), h.ptr, POP_RIGHT, X, stride(X, 2), size(X, 1), 0), "rdiv!")
    X
end
# `F \ X` and `X / F` allocate their result in the reference; bound directly (instead of copy + ldiv!) the library reads X and
# writes the new matrix, so the two directions of the host link work on different buffers
function Base.:\(F::RevCholArrowhead, X::StridedMatrix{Float64})
    Y = similar(X)
    GC.@preserve X Y check(ccall((:pop_solve_host_to, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), factor_handle(F).ptr, POP_LEFT, X, stride(X, 2), Y, stride(Y, 2), size(X, 2), 0), "\\")
    Y
end
function Base.:/(X::StridedMatrix{Float64}, F::RevCholArrowhead)
    Y = similar(X)
    GC.@preserve X Y check(ccall((:pop_solve_host_to, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), factor_handle(F).ptr, POP_RIGHT, X, stride(X, 2), Y, stride(Y, 2), size(X, 1), 0), "/")
    Y
end
function LinearAlgebra.ldiv!(F::RevCholArrowhead, X::DevicePanel)
    check(ccall((:pop_solve, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), factor_handle(F).ptr, POP_LEFT, X.ptr, X.ld, X.cols, 0), "ldiv!")
    X
end
function LinearAlgebra.rdiv!(X::DevicePanel, F::RevCholArrowhead)
    check(ccall((:pop_solve, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), factor_handle(F).ptr, POP_RIGHT, X.ptr, X.ld, X.rows, 0), "rdiv!")
    X
end

# ---- triangular solves: src/arrowhead.jl:425-486 ------------------------------------------------------
for (UPLO, UNIT, uplo, unit) in (('U', 'N', POP_UPPER, 0), ('U', 'U', POP_UPPER, 1), ('L', 'N', POP_LOWER, 0), ('L', 'U', POP_LOWER, 1))
    @eval function materialize!(M::MatLdivVec{<:TriangularLayout{$UPLO,$UNIT,ArrowheadLayout},<:AbstractStridedLayout})
        T, b = M.A, M.B
        h = upload(parent(T))
        GC.@preserve b check(ccall((:pop_trsm_host, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Float64}, Int64, Int64),
            ctx(), h.ptr, POP_LEFT, $uplo, 0, $unit, b, length(b), 1), "ldiv!")
        b
    end
end

# unpack_upper!: inverse of `pack` for the upper data of a factor (diagonal and super-diagonal of A, the
# hat-bubble couplings B, the upper bands of every D_k); the lower parts are not touched (UpperTriangular view).
function unpack_upper!(A::BBBArrowheadMatrix{Float64}, Ad, Au, Bp, Dp)
    nh, m = size(A.A, 1), length(A.D)
    q = size(A.D[1], 1); lD, uD = bandwidths(A.D[1])
    for i in 1:nh
        A.A[i, i] = Ad[i]
        i < nh && (A.A[i, i+1] = Au[i])
    end
    for (b, Bb) in enumerate(A.B), k in 1:m, t in 0:2              # Bp[k, t+1, b] = entry (hat k+t-1, bubble (b,k))
        i = k + t - 1
        l, u = bandwidths(Bb)
        1 <= i <= nh && -l <= k - i <= u && (Bb[i, k] = Bp[k, t+1, b])     # inside the band of B only (the other slots are zero)
    end
    for k in 1:m, j in 1:q, d in 0:uD
        j + d <= q && (A.D[k][j, j+d] = Dp[k, j, lD+d+1])
    end
    A
end

# ---- the drivers of examples/ as single calls -----------------------------------------------------------
"ADI(A = M, B = -M, C = K, F, a, b, c, d; tolerance) of examples/adi_poisson.jl:42-59 on the device: solves K X M + M X K = F.
`K`, `M`: Symmetric arrowheads (e.g. `-weaklaplacian(Q)[KR,KR]`, `grammatrix(Q)[KR,KR]`); F: host matrix (one copy in, the
whole iteration on the device, one copy out) or a DevicePanel (nothing crosses PCIe).  (a, b) enclose the spectrum of
K⁻¹M as in :91-94; c, d default to -b, -a (:107)."
function adi_poisson(K, M, F::StridedMatrix{Float64}, a, b, c = -b, d = -a; tolerance = 1e-15)
    hK, hM = upload(parent(K)), upload(parent(M))
    X = similar(F)
    GC.@preserve F X check(ccall((:pop_adi_poisson_host, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint),
        ctx(), hK.ptr, hM.ptr, F, stride(F, 2), X, stride(X, 2), a, b, c, d, tolerance, 0), "ADI")
    X
end
function adi_poisson(K, M, F::DevicePanel, a, b, c = -b, d = -a; tolerance = 1e-15)
    hK, hM = upload(parent(K)), upload(parent(M))
    X = DevicePanel(F.rows, F.cols)
    check(ccall((:pop_adi_poisson, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cint),
        ctx(), hK.ptr, hM.ptr, F.ptr, F.ld, X.ptr, X.ld, a, b, c, d, tolerance, 0), "ADI")
    X
end

"`nsteps` of the theta scheme (M + θ dt K) u' = (M - (1-θ) dt K) u with one reused factorisation on a DevicePanel:
θ = 1 is the backward-Euler loop of examples/heat.jl:58-63, θ = 1/2 the trapezoid loop of :41-45; dims = 2 steps an n x n
field in both directions."
function heat_steps!(U::DevicePanel, K, M, dt, nsteps; theta = 1.0, dims = 1)
    hK, hM = upload(parent(K)), upload(parent(M))
    check(ccall((:pop_heat_steps_theta, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Int64, Int64),
                ctx(), hK.ptr, hM.ptr, dt, theta, nsteps, dims, U.ptr, U.ld, U.cols), "heat")
    U
end
heat_steps!(u::StridedVecOrMat{Float64}, K, M, dt, nsteps; kw...) =
    (P = DevicePanel(reshape(u, size(u, 1), :)); heat_steps!(P, K, M, dt, nsteps; kw...); copyto!(reshape(u, size(u, 1), :), P); u)

# ---- callers' side of the ADI example (examples/adi_poisson.jl:60-76,105,111) ------------------------------------
"analysis_2D (examples/adi_poisson.jl:60-76): `T = plan * Matrix(I, p, p)` from `plan_grid_transform(legendre(0..dx), (p,p))`."
function analysis_2D_gpu(V::DevicePanel, F::DevicePanel, n::Integer, p::Integer, T::Matrix{Float64})
    GC.@preserve T check(ccall((:pop_legendre_analysis_2d, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
                ctx(), n, p, T, V.ptr, V.ld, F.ptr, F.ld), "analysis_2D")
    F
end

basis_id(::ContinuousPolynomial{1}) = Cint(0)
basis_id(::DirichletPolynomial) = Cint(1)

"(Q'P)[KR,KR] X (side = POP_LEFT) or X (Q'P)[KR,KR]' (POP_RIGHT), examples/adi_poisson.jl:105; N = length(KR)."
function weakform_apply_gpu(Q, N::Integer, side::Cint, X::DevicePanel, Y::DevicePanel)
    r = Q.points
    m = length(r) - 1
    hk = r isa AbstractRange ? Float64[] : collect(Float64, diff(r))     # rooted for the duration of the call
    GC.@preserve hk check(ccall((:pop_weakform_apply, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64),
                ctx(), basis_id(Q), m, N, Float64(r[2] - r[1]), isempty(hk) ? Ptr{Float64}(C_NULL) : pointer(hk), side, X.ptr, X.ld, Y.ptr, Y.ld,
                side == POP_LEFT ? X.cols : X.rows), "weak form")
    Y
end

"Legendre coefficients -> C1 / Dirichlet coefficients (src/continuouspolynomial.jl:84-146); throws like the reference on a jump."
function legendre_to_c1_gpu(Q, N::Integer, side::Cint, X::DevicePanel, Y::DevicePanel; scale = 1.0)
    m = length(Q.points) - 1
    jump = Ref{Cdouble}(0.0)
    check(ccall((:pop_legendre_to_c1, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Ref{Cdouble}),
                ctx(), basis_id(Q), m, N, side, X.ptr, X.ld, Y.ptr, Y.ld, side == POP_LEFT ? X.cols : X.rows, jump), "transform")
    jump[] ≤ 1000 * N * eps() * scale || throw(ArgumentError("Discontinuity in data on order of $(jump[])."))
    Y
end

"grammatrix / weaklaplacian on a non-uniform grid straight on the device (one D block per element)."
function assemble_mesh(Q, N::Integer)
    hk = collect(Float64, diff(Q.points))
    hm, hl = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve hk check(ccall((:pop_assemble_mesh, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}),
                ctx(), basis_id(Q), length(hk), N, hk, hm, hl), "assemble")
    Handle(hm[]), Handle(hl[])
end

end # module
