"""B200-native arrowhead hot path of PiecewiseOrthogonalPolynomials.jl behind the reference's API.

All arithmetic runs in libpop_arrowhead.so (hand-written CUDA for sm_100a, C ABI in
include/pop_arrowhead.h).  There is no CPU fallback."""
from ._lib import (Context, DimensionMismatch, PopError, PosDefException, StructureError, build, default_context,
                   SOLVE_AUTO, SOLVE_FUSED, SOLVE_FUSED_SMEM, SOLVE_STAGED, LIB_PATH, EXPORTED)
from .arrowhead import (BBBArrowheadMatrix, LowerTriangular, ReverseCholesky, Symmetric, UnitLowerTriangular,
                        UnitUpperTriangular, UpperTriangular, ldiv, mul, rdiv, reversecholesky, reversecholesky_inplace, as_reversecholesky,
                        reversecholesky_shifted)
from .bases import Block, BlockRange, ContinuousPolynomial, DirichletPolynomial, LazyArrowhead, grammatrix, weaklaplacian
from .drivers import (DistMatrix, ElementDecomposition, adi_poisson_dd, dd_columns, DistTranspose, adi_poisson_p2p, heat_steps_p2p, adi_num_iterations, adi_poisson, adi_poisson_distributed, adi_shifts, heat_steps,
                      shard_bounds, spectrum_bounds, eigvals)
from .rhs import (DiscontinuityError, analysis_2d, cell_grid, legendre_grid_transform, legendre_to_c1, transform, transform_2d,
                  weakform_apply, weakform_rhs, c1_to_legendre, synthesis_2d, itransform, itransform_2d, legendre_synthesis_matrix)
from ._lib import LEFT, RIGHT
