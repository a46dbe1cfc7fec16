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
# The condensed array is LAZY like the reference's: nothing is condensed until it is consumed.  The assembler (site 2)
# consumes it through ghb_condense_assemble_f64, which streams the packed records (pageable Julia Arrays are staged
# through the library's pinned double buffer) and never materialises S_K, g_K on the host.  Indexing a cell -- the
# evaluate! contract of the reference, src/StaticCondensationMap.jl:195 -- condenses the batch once into DEVICE buffers
# owned by this object (ghb_device_alloc) and copies that one cell back.
mutable struct CondensedCells <: AbstractVector{Tuple{Matrix{Float64},Vector{Float64}}}
  packed::PackedCells
  planid::Cint
  nb::Int
  dS::Ptr{Float64}        # device [n_b*n_b, ncells], C_NULL until materialised
  dg::Ptr{Float64}        # device [n_b, ncells]
end
Base.size(a::CondensedCells) = (size(a.packed.A, 2),)

function materialize!(a::CondensedCells)
  a.dS != C_NULL && return a
  n = length(a); pS = Ref{Ptr{Cvoid}}(C_NULL); pg = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ghb_device_alloc, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx(), 8 * a.nb^2 * n, pS))
  check(ccall((:ghb_device_alloc, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx(), 8 * a.nb * n, pg))
  info = Vector{Int32}(undef, n)
  check(ccall((:ghb_condense_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cint),
              ctx(), a.planid, n, a.packed.A, a.packed.b, pS[], pg[], info, 0))
  Gridap.Helpers.@check all(==(0), info)                # src/StaticCondensationMap.jl:180
  a.dS = pS[]; a.dg = pg[]
  finalizer(a) do x
    ccall((:ghb_device_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.dS)
    ccall((:ghb_device_free, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.dg)
  end
  a
end

function Base.getindex(a::CondensedCells, c::Integer)   # same element type as the reference's lazy array
  materialize!(a); nb = a.nb
  S = Matrix{Float64}(undef, nb, nb); g = Vector{Float64}(undef, nb)
  check(ccall((:ghb_copy, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx(), S, a.dS + 8 * nb^2 * (c - 1), 8 * nb^2))
  check(ccall((:ghb_copy, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx(), g, a.dg + 8 * nb * (c - 1), 8 * nb))
  (S, g)
end

function Gridap.Arrays.lazy_map(k::StaticCondensationMap, t::AbstractArray)
  p = pack(t); id, q = plan(k, p)
  CondensedCells(p, id, q[2], C_NULL, C_NULL)
end

# ---- site 2: assemble_matrix_and_vector on condensed data ----------------------------------------------------
# `cell_ids`: n_b x ncells Int64 from get_cell_dof_ids(M, dK) (RestrictFacetDoFsToSkeleton, :388-439);
# `dirichlet_values`: get_dirichlet_dof_values(M) for the lift of _attach_dirichlet (:41-42), or nothing (Newton path,
# src/HybridLinearSolvers.jl:37-41).  The symbolic pattern is a handle of the context: one per assembler, reused
# across Newton iterations (`pattern` is returned by the first call and passed back afterwards).
struct SkeletonPattern
  id::Cint
  colptr::Vector{Int64}
  rowval::Vector{Int64}
end

function symbolic(cell_ids::Matrix{Int64}, nfree::Integer)
  nnz = Ref{Int64}(0); nb, n = size(cell_ids)
  check(ccall((:ghb_assemble_symbolic, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Int64, Ref{Int64}),
              ctx(), n, nb, cell_ids, nfree, nnz))
  id = ccall((:ghb_assemble_current, lib), Cint, (Ptr{Cvoid},), ctx())
  colptr = Vector{Int64}(undef, nfree + 1); rowval = Vector{Int64}(undef, nnz[])
  check(ccall((:ghb_assemble_pattern, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), ctx(), colptr, rowval))
  SkeletonPattern(id, colptr, rowval)
end

function assemble_condensed(a::CondensedCells, cell_ids::Matrix{Int64}, nfree::Integer, dirichlet_values;
                            pattern::SkeletonPattern = symbolic(cell_ids, nfree))
  check(ccall((:ghb_assemble_select, lib), Cint, (Ptr{Cvoid}, Cint), ctx(), pattern.id))
  nzval = Vector{Float64}(undef, length(pattern.rowval)); rhs = Vector{Float64}(undef, nfree)
  dv = dirichlet_values === nothing ? C_NULL : pointer(dirichlet_values)        # host vector: its length travels with it
  ndv = dirichlet_values === nothing ? 0 : length(dirichlet_values)
  if a.dS == C_NULL
    n = length(a); info = Vector{Int32}(undef, n)
    GC.@preserve dirichlet_values check(ccall((:ghb_condense_assemble_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
              ctx(), a.planid, n, a.packed.A, a.packed.b, dv, ndv, nzval, rhs, info))
    Gridap.Helpers.@check all(==(0), info)
  else                                                  # somebody indexed the cells: S_K, g_K already live on the device
    GC.@preserve dirichlet_values check(ccall((:ghb_assemble_numeric_f64, lib), Cint,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
              ctx(), a.dS, a.dg, dv, ndv, nzval, rhs))
  end
  SparseMatrixCSC(nfree, nfree, pattern.colptr, pattern.rowval, nzval), rhs     # identical layout to sparse(I,J,V,m,n)
end

# ---- site 3: lazy_map(BackwardStaticCondensationMap, t, lhk) + assemble_vector ------------------------------
# lam_free = get_free_dof_values(lh), lam_dirichlet = get_dirichlet_dof_values(M): plain host Vectors
function backsub(k::BackwardStaticCondensationMap, p::PackedCells, lam_free::Vector{Float64},
                 lam_dirichlet::Vector{Float64}, cell_ids::Matrix{Int64})
  id, q = plan(k.static_condensation, p); ni = q[1]; n = size(p.A, 2)
  u = Matrix{Float64}(undef, ni, n); info = Vector{Int32}(undef, n)
  check(ccall((:ghb_backsub_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Int64},
               Ptr{Float64}, Ptr{Int32}),
              ctx(), id, n, p.A, p.b, lam_free, length(lam_free), lam_dirichlet, length(lam_dirichlet), cell_ids, u, info))
  Gridap.Helpers.@check all(==(0), info)
  x = Vector{Float64}(undef, ni * n + length(lam_free))
  check(ccall((:ghb_scatter_free_dof_values, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
              ctx(), id, n, u, lam_free, length(lam_free), x))
  x                                                     # free dof values of the full space (:134-149)
end

# ---- multi-GPU (one Julia process per GPU, MPI.jl for the plumbing): the two exchanges through the C ABI ------
function comm_init(comm)                                # comm::MPI.Comm
  id = zeros(UInt8, 128)
  MPI.Comm_rank(comm) == 0 && check(ccall((:ghb_comm_unique_id, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx(), id))
  MPI.Bcast!(id, 0, comm)
  check(ccall((:ghb_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx(), MPI.Comm_size(comm), MPI.Comm_rank(comm), id))
end
# collective 1: ghb_pack_cut_plane_f64 -> ghb_exchange_cut_plane_f64 -> ghb_assemble_numeric_slab_f64 (device buffers from
# ghb_device_alloc); collective 2: ghb_allgather_lambda_f64(ctx, owned, counts, all) before backsub (:113-118).

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

# f-1 without materialising the records: the condensation kernel forms A_K = sum_t coef[t,K] TA[:,t] in its loader
# (bit-identical to expand_records + the resident path); returns the assembled system like assemble_condensed
function assemble_affine(k::StaticCondensationMap, t::AbstractArray, rep::Vector{Int}, coef_rep::Matrix{Float64},
                         coef::Matrix{Float64}, cell_ids::Matrix{Int64}, nfree, dirichlet_values;
                         pattern = symbolic(cell_ids, nfree))
  pr = pack(t[rep]); id, q = plan(k, pr)
  TA = Matrix((coef_rep' \ pr.A')'); Tb = Matrix((coef_rep' \ pr.b')')
  ncells = size(coef, 2); ntab = size(coef, 1)
  check(ccall((:ghb_assemble_select, lib), Cint, (Ptr{Cvoid}, Cint), ctx(), pattern.id))
  nzval = Vector{Float64}(undef, length(pattern.rowval)); rhs = Vector{Float64}(undef, nfree); info = Vector{Int32}(undef, ncells)
  check(ccall((:ghb_condense_assemble_affine_f64, lib), Cint,
              (Ptr{Cvoid}, Cint, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64},
               Ptr{Float64}, Ptr{Int32}),
              ctx(), id, ncells, ntab, TA, Tb, coef, dirichlet_values, length(dirichlet_values), nzval, rhs, info))
  Gridap.Helpers.@check all(==(0), info)
  SparseMatrixCSC(nfree, nfree, pattern.colptr, pattern.rowval, nzval), rhs
end

# f-4: CSR hand-off.  The pattern is structurally symmetric: rowptr/colval are the colptr/rowval of
# ghb_assemble_pattern; ghb_assemble_numeric_csr_f64(ctx, S, g, dv, ndv, nzval, rhs) fills the values in row-major order
# (S must be a device array; it is transposed in place) -> SparseMatrixCSR{1}(m, n, rowptr, colval, nzval).

end # module
