# Times the TRUE reference (PiecewiseOrthogonalPolynomials.jl, unmodified) on the host cores for the configurations of
# BASELINE.json, with the metric of bench.py (unknown*RHS/s for the arrowhead solve, s/solve for the ADI).
# NOT executed in the build container (no Julia, no network there: SURVEY.md 8c); run it where Julia >= 1.10 and the
# package are installed:   julia -t auto baseline/ref_julia.jl [config]      (config = 1, 2 (default), 3, 4)
# It prints one JSON line per configuration, shaped like `bench.py --impl reference`.
using PiecewiseOrthogonalPolynomials, MatrixFactorizations, LinearAlgebra, BlockArrays, Random
using PiecewiseOrthogonalPolynomials: weaklaplacian, grammatrix
import Base.Threads: nthreads

json(d) = "{" * join(("\"$k\": " * (v isa AbstractString ? "\"$v\"" : string(v)) for (k, v) in d), ", ") * "}"

function helmholtz_solve(m, N, nrhs; reps = 3)
    # config 1 (README, m=3, N=10, 1 RHS) and config 2 (m=1024, N=40, 4096 RHS): F = reversecholesky(Symmetric(-Δ+M)); F \ X
    C = ContinuousPolynomial{1}(range(-1, 1; length = m + 1))
    KR = Block.(Base.OneTo(N))
    S = (-weaklaplacian(C) + grammatrix(C))[KR, KR]
    F = reversecholesky(Symmetric(parent(S)))
    n = size(S, 1)
    Random.seed!(0)
    X = randn(n, nrhs)
    t = minimum(1:reps) do _
        Y = copy(X)
        @elapsed (Y .= F \ Y)                      # the reference solves column by column (src/arrowhead.jl:425-486)
    end
    n, t
end

function adi_poisson(m, N)
    # config 4: examples/adi_poisson.jl:42-59 with the Dirichlet basis; include the example's ADI / ADI_shifts first
    include(joinpath(pkgdir(PiecewiseOrthogonalPolynomials), "examples", "adi_poisson.jl"))
    error("examples/adi_poisson.jl runs its own problem on include; adapt r, p there (r = range(-1,1,$(m+1)), p = $N) and time `ADI`")
end

cfg = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 2
if cfg == 1 || cfg == 2
    m, N, nrhs = cfg == 1 ? (3, 10, 1) : (1024, 40, 4096)
    n, t = helmholtz_solve(m, N, nrhs)
    println(json(["impl" => "reference", "kind" => "reference", "metric" => "arrowhead_solve_unknowns_rhs_per_s", "value" => n * nrhs / t,
                  "unit" => "unknown*RHS/s", "cores" => nthreads(), "n" => n, "nrhs" => nrhs, "s_per_step" => t]))
elseif cfg == 3
    # config 3: Dirichlet m=256, N=128, 64 shifts x 1024 RHS: factor + solve per shift
    Q = DirichletPolynomial(range(-1, 1; length = 257)); KR = Block.(Base.OneTo(128))
    K = parent((-weaklaplacian(Q))[KR, KR]); M = parent(grammatrix(Q)[KR, KR])
    n = size(K, 1); X = randn(n, 1024)
    t = @elapsed for σ in exp10.(range(log10(π^2 / 4), 9; length = 64))
        F = reversecholesky(Symmetric(K + σ * M)); F \ X
    end
    println(json(["impl" => "reference", "kind" => "reference", "metric" => "shifted_batch_unknowns_rhs_per_s", "value" => n * 1024 * 64 / t,
                  "unit" => "unknown*RHS/s", "cores" => nthreads(), "n" => n, "s_total" => t]))
else
    adi_poisson(128, 64)
end
