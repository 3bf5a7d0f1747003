# PopArrowheadGPU.jl -- Julia-side bindings of libpop_arrowhead.so (include/pop_arrowhead.h).
#
# Drop-in for the arrowhead hot path of PiecewiseOrthogonalPolynomials.jl: the user-visible calls
#   M = C'C, Δ = weaklaplacian(C), M[KR,KR], reversecholesky(Symmetric(S)).L/.U, F \ X, X / F, A*X, X*A
# keep their signatures; the methods below re-bind the reference's own dispatch points for Float64:
#   materialize!(::MatMulVecAdd/MatMulMatAdd{<:ArrowheadLayouts})      src/arrowhead.jl:191,214,237
#   reversecholesky_layout!(::ArrowheadLayouts, A, ::Type{UpperTriangular})  src/arrowhead.jl:276
#   materialize!(::MatLdivVec{<:TriangularLayout{...,ArrowheadLayout}})  src/arrowhead.jl:425-486
# plus a NEW MatLdivMat method so a matrix right-hand side goes down in one call (the reference
# falls back to a column loop).  NOT EXECUTED in the build container (no Julia there); it is the
# binding a maintainer adds, kept declarative.  `include` it after `using PiecewiseOrthogonalPolynomials`.
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

# ---- packing: BBBArrowheadMatrix -> flat arrays in the ABI layout -----------------------------------
# B[b][t][k] = entry (hat k+t-1, bubble (b,k)), C likewise transposed, D[lD+d][j][k] = D_k[j, j+d]
function pack(A::BBBArrowheadMatrix{Float64})
    nh, m = size(A.A, 1), length(A.D)
    q = size(A.D[1], 1); lD, uD = bandwidths(A.D[1])
    Ad = [A.A[i, i] for i in 1:nh]
    Au = [A.A[i, i+1] for i in 1:nh-1]; Al = [A.A[i+1, i] for i in 1:nh-1]
    slots(Bs, tr) = begin
        out = zeros(3, m, length(Bs))                            # column-major: [t][k][b] reversed below
        for (b, Bb) in enumerate(Bs), k in 1:m, t in 0:2
            i = k + t - 1 - (nh == m + 1 ? 0 : 1) + 1            # hat index (1-based) of slot t
            1 <= i <= nh && (out[t+1, k, b] = tr ? Bb[k, i] : Bb[i, k])
        end
        permutedims(out, (2, 1, 3))                              # -> [k][t][b] in memory = B[b][t][k]
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

function upload(A::BBBArrowheadMatrix{Float64})
    d, Ad, Au, Al, Bp, Cp, Dp = pack(A)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Ad Au Al Bp Cp Dp check(ccall((:pop_arrowhead_upload, libpop), Cint,
        (Ptr{Cvoid}, Ref{PopDesc}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
        ctx(), d, Ad, Au, Al, Bp, Cp, Dp, out), "pop_arrowhead_upload")
    Handle(out[])
end

# ---- products: src/arrowhead.jl:191-258 --------------------------------------------------------------
function materialize!(M::MatMulMatAdd{<:ArrowheadLayouts,<:AbstractColumnMajor,<:AbstractColumnMajor,Float64})
    α, A, X, β, Y = M.α, M.A, M.B, M.β, M.C
    h = upload(A)
    GC.@preserve X Y check(ccall((:pop_mul_host, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}, Int64, Int64),
        ctx(), h.ptr, POP_LEFT, POP_GENERAL, 0, α, X, stride(X, 2), β, Y, stride(Y, 2), size(X, 2)), "mul!")
    Y
end
function materialize!(M::MatMulVecAdd{<:ArrowheadLayouts,<:AbstractStridedLayout,<:AbstractStridedLayout,Float64})
    materialize!(MulAdd(M.α, M.A, reshape(M.B, :, 1), M.β, reshape(M.C, :, 1))); M.C
end
function materialize!(M::MatMulMatAdd{<:AbstractColumnMajor,<:ArrowheadLayouts,<:AbstractColumnMajor,Float64})   # X*A
    α, X, A, β, Y = M.α, M.A, M.B, M.β, M.C
    h = upload(A)
    GC.@preserve X Y check(ccall((:pop_mul_host, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}, Int64, Int64),
        ctx(), h.ptr, POP_RIGHT, POP_GENERAL, 0, α, X, stride(X, 2), β, Y, stride(Y, 2), size(X, 1)), "mul!")
    Y
end

# ---- reverse Cholesky: src/arrowhead.jl:276-292 -----------------------------------------------------
# factor on the device, copy the O(n) factor back into A's own storage so .U, getindex and U*U' ≈ S
# (test/test_arrowhead.jl:145) keep working; the device handle is cached for the solves.
const FACTORS = WeakKeyDict{Any,Handle}()
function reversecholesky_layout!(::ArrowheadLayouts, A::BBBArrowheadMatrix{Float64}, ::Type{UpperTriangular})
    h = upload(A)
    check(ccall((:pop_revchol_inplace, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), h.ptr), "reversecholesky!")
    d, Ad, Au, Al, Bp, Cp, Dp = pack(A)
    GC.@preserve Ad Au Bp Dp check(ccall((:pop_arrowhead_download, libpop), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx(), h.ptr, 1, Ad, Au, C_NULL, Bp, C_NULL, Dp), "factor download")
    unpack_upper!(A, Ad, Au, Bp, Dp)                             # writes A.A, A.B[b], A.D[k] upper data
    FACTORS[A] = h
    UpperTriangular(A), 0
end

# ---- F \ X and X / F: one call for the whole panel (test/test_arrowhead.jl:130, adi_poisson.jl:54-55) --
function LinearAlgebra.ldiv!(F::ReverseCholesky{Float64,<:UpperTriangular{Float64,<:BBBArrowheadMatrix}}, X::StridedVecOrMat{Float64})
    h = get!(() -> (u = upload(parent(F.U)); ccall((:pop_revchol_inplace, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), u.ptr); u), FACTORS, parent(F.U))
    GC.@preserve X check(ccall((:pop_solve_host, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), h.ptr, POP_LEFT, X, stride(X, 2), size(X, 2), 0), "ldiv!")
    X
end
function LinearAlgebra.rdiv!(X::StridedMatrix{Float64}, F::ReverseCholesky{Float64,<:UpperTriangular{Float64,<:BBBArrowheadMatrix}})
    h = FACTORS[parent(F.U)]
    GC.@preserve X check(ccall((:pop_solve_host, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Int64, Cint),
        ctx(), h.ptr, POP_RIGHT, X, stride(X, 2), size(X, 1), 0), "rdiv!")
    X
end

# ---- triangular solves: src/arrowhead.jl:425-486 ------------------------------------------------------
for (UPLO, UNIT, Tri, uplo, unit) in (('U', 'N', UpperTriangular, POP_UPPER, 0), ('U', 'U', UnitUpperTriangular, POP_UPPER, 1),
                                      ('L', 'N', LowerTriangular, POP_LOWER, 0), ('L', 'U', UnitLowerTriangular, POP_LOWER, 1))
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
    for (b, Bb) in enumerate(A.B), k in 1:m, t in 0:2              # Bp[k, t+1, b] = entry (hat of slot t, bubble (b,k))
        i = k + t - 1 - (nh == m + 1 ? 0 : 1) + 1
        1 <= i <= nh && (Bb[i, k] = Bp[k, t+1, b])
    end
    for k in 1:m, j in 1:q, d in 0:uD
        j + d <= q && (A.D[k][j, j+d] = Dp[k, j, lD+d+1])
    end
    A
end


# ---- callers' side of the ADI example (examples/adi_poisson.jl:60-76,105,111; examples/heat.jl:26-47) -------------
# Device-resident panels are `CuArray`s or raw device pointers obtained elsewhere; the signatures below take pointers
# and leading dimensions exactly like the header.

"analysis_2D (examples/adi_poisson.jl:60-76): `T = plan * Matrix(I, p, p)` from `plan_grid_transform(legendre(0..dx), (p,p))`."
analysis_2D_gpu(dV::Ptr{Float64}, ldv, dF::Ptr{Float64}, ldf, n::Integer, p::Integer, T::Matrix{Float64}) =
    check(ccall((:pop_legendre_analysis_2d, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
                ctx(), n, p, T, dV, ldv, dF, ldf), "analysis_2D")

basis_id(::ContinuousPolynomial{1}) = Cint(0)
basis_id(::DirichletPolynomial) = Cint(1)

"(Q'P)[KR,KR] X (side = POP_LEFT) or X (Q'P)[KR,KR]' (POP_RIGHT), examples/adi_poisson.jl:105; N = length(KR)."
function weakform_apply_gpu(Q, N::Integer, side::Cint, dX::Ptr{Float64}, ldx, dY::Ptr{Float64}, ldy, cnt)
    r = Q.points
    m = length(r) - 1
    hk = r isa AbstractRange ? C_NULL : pointer(collect(Float64, diff(r)))
    check(ccall((:pop_weakform_apply, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64),
                ctx(), basis_id(Q), m, N, Float64(r[2] - r[1]), hk, side, dX, ldx, dY, ldy, cnt), "weak form")
end

"Legendre coefficients -> C1 / Dirichlet coefficients (src/continuouspolynomial.jl:84-146); throws like the reference on a jump."
function legendre_to_c1_gpu(Q, N::Integer, side::Cint, dX::Ptr{Float64}, ldx, dY::Ptr{Float64}, ldy, cnt; scale = 1.0)
    m = length(Q.points) - 1
    jump = Ref{Cdouble}(0.0)
    check(ccall((:pop_legendre_to_c1, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Ref{Cdouble}),
                ctx(), basis_id(Q), m, N, side, dX, ldx, dY, ldy, cnt, jump), "transform")
    jump[] ≤ 1000 * N * eps() * scale || throw(ArgumentError("Discontinuity in data on order of $(jump[])."))
end

"theta scheme with one reused factorisation: theta = 1/2 is the trapezoid loop of examples/heat.jl:41-45."
heat_steps_theta!(hK, hM, dt, theta, nsteps, dims, dU::Ptr{Float64}, ldu, nrhs) =
    check(ccall((:pop_heat_steps_theta, libpop), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Int64, Int64),
                ctx(), hK, hM, dt, theta, nsteps, dims, dU, ldu, nrhs), "heat")

"grammatrix / weaklaplacian on a non-uniform grid straight on the device (one D block per element)."
function assemble_mesh(Q, N::Integer)
    hk = collect(Float64, diff(Q.points))
    hm, hl = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pop_assemble_mesh, libpop), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}),
                ctx(), basis_id(Q), length(hk), N, hk, hm, hl), "assemble")
    hm[], hl[]
end

end # module
