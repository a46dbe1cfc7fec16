# GridapHybridB200.jl -- thin Julia glue that routes the three array-level sites of GridapHybrid's
# hybridisation path to libgridaphybrid_b200.so (C ABI: include/ghb.h).
#
# NOT EXECUTABLE IN THE BUILD CONTAINER (no Julia, no Gridap there): this is the binding a maintainer adds
# on the reference side; it is exercised through the identical C entry points from Python in tests/.
#
# Sites replaced (reference file:line):
#   lazy_map(StaticCondensationMap(bf,sf), t)            src/HybridAffineFEOperators.jl:338
#   assemble_matrix_and_vector(assem, data)              src/HybridAffineFEOperators.jl:46, src/HybridLinearSolvers.jl:43-44
#   lazy_map(BackwardStaticCondensationMap(bf,sf),t,lhk) src/HybridAffineFEOperators.jl:117-118 (+ assemble_vector :149)
module GridapHybridB200

using Gridap, GridapHybrid, SparseArrays
using Gridap.Fields: ArrayBlock, MatrixBlock, VectorBlock

const lib = get(ENV, "GHB_LIB", "libgridaphybrid_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function ctx()
  if CTX[] == C_NULL
    rc = ccall((:ghb_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), 0, CTX)
    rc == 0 || error("ghb_create failed ($rc): no CUDA device -- there is no CPU fallback")
  end
  CTX[]
end
check(rc) = rc == 0 || error(unsafe_string(ccall((:ghb_last_error, lib), Cstring, (Ptr{Cvoid},), ctx())))

# ---- packing: lazy cell array of (MatrixBlock, VectorBlock) -> packed records -------------------------
struct PackedCells
  A::Matrix{Float64}      # lenA x ncells (column = one cell record)
  b::Matrix{Float64}      # lenb x ncells
  ndofs::Vector{Int32}
  touched::Matrix{UInt8}
end

function pack(t::AbstractArray)
  cache = array_cache(t)
  A1, b1 = getindex!(cache, t, 1)
  nf = length(b1.array)
  brs, _ = GridapHybrid._compute_brs_bcs(A1)            # src/StaticCondensationMap.jl:72-84
  touched = UInt8.(A1.touched)
  lenA = sum(brs[i] * brs[j] for j in 1:nf, i in 1:nf if A1.touched[i, j])
  A = Matrix{Float64}(undef, lenA, length(t)); b = Matrix{Float64}(undef, sum(brs), length(t))
  for c in 1:length(t)
    Ac, bc = getindex!(cache, t, c)                     # evaluates the lazy integration of cell c (unchanged Gridap)
    o = 0
    for j in 1:nf, i in 1:nf                            # block-column-major, each block column-major
      if Ac.touched[i, j]
        n = brs[i] * brs[j]; copyto!(A, (c - 1) * lenA + o + 1, Ac.array[i, j], 1, n); o += n
      end
    end
    o = 0
    for i in 1:nf
      copyto!(b, (c - 1) * size(b, 1) + o + 1, bc.array[i], 1, brs[i]); o += brs[i]
    end
  end
  PackedCells(A, b, Int32.(brs), touched)
end

function plan(k::StaticCondensationMap, p::PackedCells)
  id = Ref{Cint}(-1)
  check(ccall((:ghb_plan_blocks, lib), Cint,
              (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{UInt8}, Cint, Ptr{Int32}, Cint, Ptr{Int32}, Ref{Cint}),
              ctx(), length(p.ndofs), p.ndofs, p.touched, length(k.interior_fields), Int32.(k.interior_fields),
              length(k.boundary_fields), Int32.(k.boundary_fields), id))
  q = zeros(Int64, 4); check(ccall((:ghb_plan_query, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), ctx(), id[], q))
  id[], q   # q = n_i, n_b, lenA, lenb
end

# ---- site 1: lazy_map(StaticCondensationMap, t) ------------------------------------------------------------
struct CondensedCells <: AbstractVector{Tuple{Matrix{Float64},Vector{Float64}}}
  S::Array{Float64,3}     # n_b x n_b x ncells
  g::Matrix{Float64}      # n_b x ncells
  packed::PackedCells
  planid::Cint
end
Base.size(a::CondensedCells) = (size(a.g, 2),)
Base.getindex(a::CondensedCells, c::Integer) = (a.S[:, :, c], a.g[:, c])   # same element type as the reference's lazy array

function Gridap.Arrays.lazy_map(k::StaticCondensationMap, t::AbstractArray)
  p = pack(t); id, q = plan(k, p); nb = q[2]; n = size(p.A, 2)
  S = Array{Float64}(undef, nb, nb, n); g = Matrix{Float64}(undef, nb, n); info = Vector{Int32}(undef, n)
  check(ccall((:ghb_condense_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cint),
              ctx(), id, n, p.A, p.b, S, g, info, 0))
  Gridap.Helpers.@check all(==(0), info)                # src/StaticCondensationMap.jl:180
  CondensedCells(S, g, p, id)
end

# ---- site 2: assemble_matrix_and_vector on condensed data ----------------------------------------------------
# `cell_ids`: n_b x ncells Int64 from get_cell_dof_ids(M, dK) (RestrictFacetDoFsToSkeleton, :388-439);
# `dirichlet_values`: get_dirichlet_dof_values(M) for the lift of _attach_dirichlet (:41-42), or nothing.
function assemble_condensed(a::CondensedCells, cell_ids::Matrix{Int64}, nfree::Integer, dirichlet_values)
  nnz = Ref{Int64}(0); nb, n = size(cell_ids)
  check(ccall((:ghb_assemble_symbolic, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Int64, Ref{Int64}),
              ctx(), n, nb, cell_ids, nfree, nnz))
  colptr = Vector{Int64}(undef, nfree + 1); rowval = Vector{Int64}(undef, nnz[])
  check(ccall((:ghb_assemble_pattern, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), ctx(), colptr, rowval))
  nzval = Vector{Float64}(undef, nnz[]); rhs = Vector{Float64}(undef, nfree)
  # dirichlet_values must be a device pointer in the C ABI; a production glue keeps it in a CuArray
  dv = dirichlet_values === nothing ? C_NULL : pointer(dirichlet_values)
  check(ccall((:ghb_assemble_numeric_f64, lib), Cint,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              ctx(), a.S, a.g, dv, nzval, rhs))
  SparseMatrixCSC(nfree, nfree, colptr, rowval, nzval), rhs   # identical layout to sparse(I,J,V,m,n)
end

# ---- site 3: lazy_map(BackwardStaticCondensationMap, t, lhk) + assemble_vector ------------------------------
function backsub(k::BackwardStaticCondensationMap, p::PackedCells, lam_free, lam_dirichlet, cell_ids::Matrix{Int64})
  id, q = plan(k.static_condensation, p); ni = q[1]; n = size(p.A, 2)
  u = Matrix{Float64}(undef, ni, n); info = Vector{Int32}(undef, n)
  check(ccall((:ghb_backsub_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int32}),
              ctx(), id, n, p.A, p.b, lam_free, lam_dirichlet, cell_ids, u, info))
  x = Vector{Float64}(undef, ni * n + length(lam_free))
  check(ccall((:ghb_scatter_free_dof_values, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
              ctx(), id, n, u, lam_free, length(lam_free), x))
  x                                                     # free dof values of the full space (:134-149)
end

# ---- next rows (SURVEY 8f) -----------------------------------------------------------------------------------
# f-3: lazy_map(compute_bulk_to_skeleton_l2_projection_dofs, A_array, B_array) of test/P_m.jl:17 on the batch:
# A [n, n, nbatch], B [n, m, nbatch] (Julia arrays are column-major per system already)
function l2_projection_dofs(A::Array{Float64,3}, B::Array{Float64,3})
  n, m, nb = size(B, 1), size(B, 2), size(B, 3)
  X = similar(B); info = Vector{Int32}(undef, nb)
  check(ccall((:ghb_l2_projection_dofs_f64, lib), Cint,
              (Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
              ctx(), nb, n, m, A, B, X, info))
  @assert all(==(0), info)                              # `A\B` would have thrown SingularException
  X
end

# f-1: element records of an affine family (Cartesian mesh, cell-wise constant coefficients): evaluate the lazy cell
# array `t` of _add_static_condensation's input on the `rep` representative cells only, solve for the tables, expand
# on the device.  `coef_rep` [ntab, ntab] and `coef` [ntab, ncells] are the coefficient vectors (1, boundary flags of
# the low-side facets, cell origin, material extras) of the representative / of all cells.
function expand_records(k::StaticCondensationMap, t::AbstractArray, rep::Vector{Int}, coef_rep::Matrix{Float64},
                        coef::Matrix{Float64})
  pr = pack(t[rep]); id, q = plan(k, pr)
  TA = (coef_rep' \ pr.A')'                             # [lenA, ntab]: records are columns on the Julia side
  Tb = (coef_rep' \ pr.b')'
  ncells = size(coef, 2); ntab = size(coef, 1)
  A = Matrix{Float64}(undef, size(pr.A, 1), ncells); b = Matrix{Float64}(undef, size(pr.b, 1), ncells)
  check(ccall((:ghb_expand_records_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              ctx(), id, ncells, ntab, Matrix(TA), Matrix(Tb), coef, A, b))   # a production glue keeps A, b in CuArrays
  PackedCells(A, b, pr.ndofs, pr.touched)
end

# f-4: CSR hand-off.  The pattern is structurally symmetric: rowptr/colval are the colptr/rowval of
# ghb_assemble_pattern; ghb_assemble_numeric_csr_f64(ctx, S, g, dv, nzval, rhs) fills the values in row-major order
# (S must be a device array; it is transposed in place) -> SparseMatrixCSR{1}(m, n, rowptr, colval, nzval).

end # module
