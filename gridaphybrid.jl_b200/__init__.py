"""gridaphybrid_b200: B200-native (sm_100a) implementation of GridapHybrid.jl's per-cell hybridisation
hot path (static condensation -> skeleton assembly -> backward recovery) behind the reference's
Map / operator interface.  The arithmetic lives in libgridaphybrid_b200.so (csrc/, C ABI in
include/ghb.h); this package is the host-side mirror of the reference interface.

Export list mirrors /root/reference/src/GridapHybrid.jl:15-32 for the symbols on the path.
"""
from ._lib import GhbError, build, lib  # noqa: F401
from .context import BlockPlan, Context  # noqa: F401
from .blocks import ArrayBlock, CondensedCells, MatrixBlock, PackedCells, VectorBlock  # noqa: F401
from .maps import (BackwardStaticCondensationMap, RestrictArrayBlockMap, Scalar2ArrayBlockMap,  # noqa: F401
                   StaticCondensationMap, SumFacetsMap, compute_bulk_to_skeleton_l2_projection_dofs, default_context,
                   lazy_map, set_default_context)
from .skeleton import CartesianSkeleton, FacetFESpace, MultiFieldFacetFESpace  # noqa: F401
from .assembly import (SparseMatrixAssembler, SparseMatrixCSC, SparseMatrixCSR, assemble_matrix_and_vector,  # noqa: F401
                       assemble_matrix_and_vector_csr, attach_dirichlet, condense_and_assemble)
from .families import AffineCells, AffineRecordFamily, cartesian_coefficients  # noqa: F401
from .operators import (AffineFEOperator, HybridAffineFEOperator, HybridFEOperator,  # noqa: F401
                        hybrid_backslash_solve, solve_skeleton)
