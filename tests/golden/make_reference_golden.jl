# make_reference_golden.jl -- run the REAL reference (GridapHybrid.jl + the Gridap it pins) on its own 2x2 test cases and dump
# what tests/test_reference_golden.py compares this repo against: per-cell packed records (A_K, b_K), condensed blocks
# (S_K, g_K), the cell dof ids of the skeleton space, the assembled SparseMatrixCSC (colptr / rowval / nzval) + rhs, the
# skeleton solution and the recovered interior dof values u_K.
#
# NOT RUN IN THE BUILD CONTAINER (no Julia there).  A maintainer with Julia runs, from the reference checkout:
#     julia --project=. /path/to/repo/tests/golden/make_reference_golden.jl /path/to/repo/tests/golden/reference
# (instantiates Manifest.toml: Gridap 0.18.2 @ exploring_hybridization), commits the three JSON files it writes, and the
# parity of this repo is then pinned by the running reference instead of by the restatement in oracle/ ("parity unpinned"
# in DESIGN.md section 5 becomes "pinned").
#
# The cases are the reference's own: test/DarcyHDGTests.jl:27-142, test/DarcyRTHTests.jl:24-83,
# test/MultiFieldLagrangeMultipliersTests.jl:24-84.  The steps below are the body of the HybridAffineFEOperator constructor
# (src/HybridAffineFEOperators.jl:10-50) and of solve! (:67-100), unrolled so that the intermediates can be written out.
using Gridap, GridapHybrid, SparseArrays
using Gridap.Arrays, Gridap.Fields, Gridap.FESpaces, Gridap.CellData, Gridap.Geometry

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "reference")
mkpath(outdir)

# ---- minimal JSON writer (no package beyond the reference's Manifest) ---------------------------------------------
jnum(x::Integer) = string(x)
jnum(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : "null"          # repr round-trips Float64 exactly
jarr(v) = "[" * join((x isa AbstractArray ? jarr(x) : jnum(x) for x in v), ",") * "]"
function jwrite(path, d::Vector{Pair{String,Any}})
  open(path, "w") do io
    print(io, "{")
    for (k, (name, val)) in enumerate(d)
      k > 1 && print(io, ",")
      print(io, "\"", name, "\":", val isa AbstractString ? "\"" * val * "\"" : (val isa AbstractArray ? jarr(val) : jnum(val)))
    end
    print(io, "}")
  end
end

# packed record of one cell: touched blocks in block-column-major order, each block column-major (include/ghb.h)
function pack_cell(A::ArrayBlock, b::ArrayBlock)
  nf = length(b.array)
  rec = Float64[]
  for j in 1:nf, i in 1:nf
    A.touched[i, j] && append!(rec, vec(A.array[i, j]))
  end
  bv = Float64[]
  for i in 1:nf
    append!(bv, b.array[i])
  end
  rec, bv
end

function dump_case(name, weakform, X, Y, bulk, skel)
  u = get_trial_fe_basis(X); v = get_fe_basis(Y)
  biform, liform = weakform(u, v)
  obiform, oliform = GridapHybrid._merge_bulk_and_skeleton_contributions(biform, liform)
  matvec, mat, vec_ = Gridap.FESpaces._pair_contribution_when_possible(obiform, oliform)
  Γ = first(keys(matvec.dict))
  t = matvec.dict[Γ]                                     # lazy cell array of (A_K, b_K): input of StaticCondensationMap
  ncells = length(t)
  A1, b1 = t[1]
  brs, _ = GridapHybrid._compute_brs_bcs(A1)
  recsA = Vector{Vector{Float64}}(); recsb = Vector{Vector{Float64}}()
  for c in 1:ncells
    Ac, bc = t[c]
    r, bv = pack_cell(Ac, bc)
    push!(recsA, r); push!(recsb, bv)
  end
  # site 1: StaticCondensationMap (src/StaticCondensationMap.jl:152-196)
  k = StaticCondensationMap(bulk, skel)
  cond = lazy_map(k, t)
  S = Vector{Vector{Float64}}(); g = Vector{Vector{Float64}}()
  for c in 1:ncells
    Sc, gc = cond[c]
    push!(S, copy(vec(Sc))); push!(g, copy(gc))          # the Map returns views into its cache: copy
  end
  # site 2: the constructor's tail (src/HybridAffineFEOperators.jl:31-46)
  matvec2 = GridapHybrid._add_static_condensation(matvec, bulk, skel)
  M, L = GridapHybrid._setup_fe_spaces_skeleton_system(X, Y, skel)
  if length(skel) != 1
    matvec2 = GridapHybrid._block_skeleton_system_contributions(matvec2, L)
  end
  assem = SparseMatrixAssembler(M, L)
  uhd = zero(M)
  matvec3, mat3 = Gridap.FESpaces._attach_dirichlet(matvec2, mat, uhd)
  data = Gridap.FESpaces._collect_cell_matrix_and_vector(M, L, matvec3, mat3, vec_)
  Asp, rhs = assemble_matrix_and_vector(assem, data)
  ids = get_cell_dof_ids(M, Γ)                           # RestrictFacetDoFsToSkeleton (:388-439)
  cell_ids = Vector{Vector{Int}}()
  for c in 1:ncells
    idc = ids[c]
    push!(cell_ids, idc isa ArrayBlock ? reduce(vcat, [collect(idc.array[f]) for f in 1:length(idc.array)]) : collect(idc))
  end
  dvals = length(skel) == 1 ? collect(get_dirichlet_dof_values(M)) :
          reduce(vcat, [collect(get_dirichlet_dof_values(M[f])) for f in 1:length(skel)])
  # skeleton solve + site 3 (src/HybridAffineFEOperators.jl:67-150)
  op = HybridAffineFEOperator(weakform, X, Y, bulk, skel)
  lh = solve(op.skeleton_op)
  lam_free = collect(get_free_dof_values(lh))
  lhk = get_cell_dof_values(lh, Γ)
  mback = BackwardStaticCondensationMap(bulk, skel)
  uk = lazy_map(mback, t, lhk)
  ucells = Vector{Vector{Float64}}()
  for c in 1:ncells
    blk = uk[c]
    push!(ucells, reduce(vcat, [collect(blk.array[f]) for f in 1:length(bulk)]))
  end
  xh = solve(op)
  jwrite(joinpath(outdir, name * ".json"), Pair{String,Any}[
    "case" => name, "gridap" => string(pkgversion(Gridap)), "ncells" => ncells,
    "ndofs" => collect(brs), "touched" => [[Int(A1.touched[i, j]) for j in 1:size(A1.touched, 2)] for i in 1:size(A1.touched, 1)],
    "interior" => bulk, "boundary" => skel,
    "A" => recsA, "b" => recsb, "S" => S, "g" => g, "cell_ids" => cell_ids, "dirichlet_values" => dvals,
    "nrows" => size(Asp, 1), "colptr" => Asp.colptr, "rowval" => Asp.rowval, "nzval" => Asp.nzval, "rhs" => rhs,
    "lambda_free" => lam_free, "u" => ucells, "free_dof_values" => collect(get_free_dof_values(xh))])
  println("wrote ", joinpath(outdir, name * ".json"))
end

# ---- the reference's own problems (2x2 Cartesian quads) -------------------------------------------------------------
uex(x) = VectorValue(1 + x[1], 1 + x[2])
Gridap.divergence(::typeof(uex)) = (x) -> 2
pex(x) = -3.14
∇pex(x) = zero(x)
Gridap.∇(::typeof(pex)) = ∇pex
fex(x) = uex(x) + ∇pex(x)

model = CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
D = num_cell_dims(model)
Ω = Triangulation(ReferenceFE{D}, model)
Γ = Triangulation(ReferenceFE{D - 1}, model)
∂K = GridapHybrid.Skeleton(model)

# 1. Darcy HDG order 1 (test/DarcyHDGTests.jl:27-142)
let order = 1
  V = TestFESpace(Ω, ReferenceFE(lagrangian, VectorValue{D,Float64}, order; space=:P); conformity=:L2)
  Q = TestFESpace(Ω, ReferenceFE(lagrangian, Float64, order - 1; space=:P); conformity=:L2)
  M = TestFESpace(Γ, ReferenceFE(lagrangian, Float64, order; space=:P); conformity=:L2, dirichlet_tags=collect(5:8))
  Y = MultiFieldFESpace([V, Q, M])
  X = MultiFieldFESpace([TrialFESpace(V), TrialFESpace(Q), TrialFESpace(M, pex)])
  τ = 1.0
  degree = 2 * (order + 1)
  dΩ = Measure(Ω, degree); d∂K = Measure(∂K, degree)
  n = get_cell_normal_vector(∂K); nₒ = get_cell_owner_normal_vector(∂K)
  a((uh, ph, lh), (vh, qh, mh)) = ∫(vh ⋅ uh - (∇ ⋅ vh) * ph - ∇(qh) ⋅ uh)dΩ + ∫((vh ⋅ n) * lh)d∂K + ∫(qh * (uh ⋅ n))d∂K +
                                  ∫(τ * qh * ph * (n ⋅ nₒ))d∂K - ∫(τ * qh * lh * (n ⋅ nₒ))d∂K + ∫(mh * (uh ⋅ n))d∂K +
                                  ∫(τ * mh * ph * (n ⋅ nₒ))d∂K - ∫(τ * mh * lh * (n ⋅ nₒ))d∂K
  l((vh, qh, mh)) = ∫(vh ⋅ fex + qh * (∇ ⋅ uex)) * dΩ
  dump_case("darcy_hdg_k1_2x2", (u, v) -> (a(u, v), l(v)), X, Y, [1, 2], [3])
end

# 2. Darcy RT-H order 0 (test/DarcyRTHTests.jl:24-83)
let order = 0
  V = TestFESpace(Ω, ReferenceFE(raviart_thomas, Float64, order); conformity=:L2)
  Q = TestFESpace(Ω, ReferenceFE(lagrangian, Float64, order); conformity=:L2)
  M = TestFESpace(Γ, ReferenceFE(lagrangian, Float64, order); conformity=:L2, dirichlet_tags=collect(5:8))
  Y = MultiFieldFESpace([V, Q, M])
  X = MultiFieldFESpace([TrialFESpace(V), TrialFESpace(Q), TrialFESpace(M, pex)])
  degree = 2 * (order + 1)
  dΩ = Measure(Ω, degree); d∂K = Measure(∂K, degree)
  n = get_cell_normal_vector(∂K)
  a((uh, ph, lh), (vh, qh, mh)) = ∫(vh ⋅ uh - (∇ ⋅ vh) * ph + qh * (∇ ⋅ uh))dΩ + ∫((vh ⋅ n) * lh)d∂K + ∫(mh * (uh ⋅ n))d∂K
  l((vh, qh, mh)) = ∫(vh ⋅ fex + qh * (∇ ⋅ uex)) * dΩ
  dump_case("darcy_rth_k0_2x2", (u, v) -> (a(u, v), l(v)), X, Y, [1, 2], [3])
end

# 3. two Lagrange-multiplier fields (test/MultiFieldLagrangeMultipliersTests.jl:24-84): Scalar2ArrayBlockMap path
let order = 0
  ru = ReferenceFE(raviart_thomas, Float64, order); rp = ReferenceFE(lagrangian, Float64, order)
  V1 = TestFESpace(Ω, ru; conformity=:L2); V2 = TestFESpace(Ω, ru; conformity=:L2)
  Q1 = TestFESpace(Ω, rp; conformity=:L2); Q2 = TestFESpace(Ω, rp; conformity=:L2)
  M1 = TestFESpace(Γ, rp; conformity=:L2, dirichlet_tags=collect(5:8)); M2 = TestFESpace(Γ, rp; conformity=:L2, dirichlet_tags=collect(5:8))
  Y = MultiFieldFESpace([V1, V2, Q1, Q2, M1, M2])
  X = MultiFieldFESpace([TrialFESpace(V1), TrialFESpace(V2), TrialFESpace(Q1), TrialFESpace(Q2), TrialFESpace(M1, pex), TrialFESpace(M2, pex)])
  degree = 2 * (order + 1)
  dΩ = Measure(Ω, degree); d∂K = Measure(∂K, degree)
  n = get_cell_normal_vector(∂K)
  ablock((uh, ph, lh), (vh, qh, mh)) = ∫(vh ⋅ uh - (∇ ⋅ vh) * ph + qh * (∇ ⋅ uh))dΩ + ∫((vh ⋅ n) * lh)d∂K + ∫(mh * (uh ⋅ n))d∂K
  lblock((vh, qh, mh)) = ∫(vh ⋅ fex + qh * (∇ ⋅ uex)) * dΩ
  a((u1, u2, p1, p2, l1, l2), (v1, v2, q1, q2, m1, m2)) = ablock((u1, p1, l1), (v1, q1, m1)) + ablock((u2, p2, l2), (v2, q2, m2))
  l((v1, v2, q1, q2, m1, m2)) = lblock((v1, q1, m1)) + lblock((v2, q2, m2))
  dump_case("multifield_2skel_2x2", (u, v) -> (a(u, v), l(v)), X, Y, collect(1:4), collect(5:6))
end
