/* ghb.h -- C ABI of libgridaphybrid_b200.so: B200 (sm_100a) implementation of GridapHybrid.jl's
 * per-cell hybridisation hot path (static condensation -> skeleton assembly -> backward recovery).
 *
 * The reference has no FFI; its "plugin API" is Gridap's Map protocol (return_cache / evaluate! /
 * lazy_map).  These entry points are what a thin Julia glue binds with `ccall` at the three
 * array-level sites of the reference (INTEGRATION.md shows the glue):
 *
 *   lazy_map(StaticCondensationMap(bf,sf), t)          src/HybridAffineFEOperators.jl:338  -> ghb_condense_f64
 *   assemble_matrix_and_vector(assem, data)            src/HybridAffineFEOperators.jl:46,
 *                                                      src/HybridLinearSolvers.jl:43-44   -> ghb_assemble_symbolic,
 *                                                                                            ghb_assemble_numeric_f64,
 *                                                                                            ghb_condense_assemble_f64
 *   lazy_map(BackwardStaticCondensationMap(bf,sf),t,l) src/HybridAffineFEOperators.jl:117-118 -> ghb_backsub_f64
 *   assemble_vector(assem, ...)                        src/HybridAffineFEOperators.jl:149  -> ghb_scatter_free_dof_values
 *
 * Conventions (Julia side): all ids are Int64 and 1-based; a dof id < 0 is a Dirichlet dof
 * (-k = k-th Dirichlet value); dense matrices are column-major; all floating data is FP64.
 *
 * Pointers: every array argument may be a device pointer or a host pointer (pageable or pinned);
 * the library asks cudaPointerGetAttributes.  The caller owns every array; the library keeps no
 * caller pointer past the call (the symbolic phase copies what it caches).
 * Errors: every function returns GHB_OK (0) or a negative GHB_E*; ghb_last_error(ctx) holds text.
 * Per-cell factorisation failures are NOT errors: they are reported in info[] with LAPACK dgetrf
 * semantics (info[c] = k > 0: U(k,k) is exactly zero), mirroring `@check info==0`
 * (src/StaticCondensationMap.jl:180); outputs of such a cell are NaN.
 * Threading: one caller thread per ctx (the reference is serial); distinct ctxs are independent.
 * There is no CPU fallback: without a CUDA device ghb_create fails with GHB_ENODEVICE.
 */
#ifndef GHB_H
#define GHB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GHB_OK 0
#define GHB_EINVAL (-1)     /* bad argument (message in ghb_last_error) */
#define GHB_ECUDA (-2)      /* CUDA runtime error */
#define GHB_ENODEVICE (-3)  /* no usable CUDA device: there is no CPU fallback */
#define GHB_ENOMEM (-4)
#define GHB_EUNSUPPORTED (-5) /* valid input the library does not handle (e.g. a dof shared by >2 cells) */
#define GHB_ESTATE (-6)     /* call order violated (e.g. numeric before symbolic) */

typedef struct ghb_ctx ghb_ctx; /* opaque: device, stream, plans, cached symbolic pattern, scratch */

/* ---- context ------------------------------------------------------------------------------ */
int ghb_create(int device_id, ghb_ctx** out);
void ghb_destroy(ghb_ctx* ctx);
const char* ghb_last_error(const ghb_ctx* ctx);
/* Streams: a new ctx owns a non-blocking stream.  ghb_set_stream makes it launch on a caller-provided
 * cudaStream_t instead (e.g. torch's current stream; NULL = the legacy default stream).  Calls whose
 * outputs are device pointers return without waiting (stream-ordered); outputs to host pointers are
 * complete on return. */
int ghb_set_stream(ghb_ctx* ctx, void* cuda_stream);
int ghb_synchronize(ghb_ctx* ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
int64_t ghb_launch_count(const ghb_ctx* ctx);
/* Debugging / A-B knobs.  ghb_create reads them ONCE from the environment (GHB_FORCE_GENERIC, GHB_CW, GHB_DMMA_LL,
 * GHB_FACTORS_GENERIC, GHB_MAX_CTAS_PER_SM, GHB_LL_CTAS, GHB_WARP_ONE_CELL, GHB_WARP_TWO_ROWS, GHB_DEBUG,
 * GHB_STREAM_CHUNK_BYTES); this call changes one by its lower-case name without the prefix ("cw", "stream_chunk_bytes",
 * ...).  Kernel-choice knobs act on plans created afterwards; nothing on the launch path reads the environment. */
int ghb_set_option(ghb_ctx* ctx, const char* name, int64_t value);
/* Device buffers for hosts without a CUDA array library of their own: the Julia glue keeps S_K, g_K between the
 * condensation site and the assembly site on the device instead of round-tripping them through host memory.
 * ghb_copy moves bytes between any two of host / device (cudaMemcpyDefault) and is complete on return. */
int ghb_device_alloc(ghb_ctx* ctx, int64_t bytes, void** out);
int ghb_device_free(ghb_ctx* ctx, void* ptr);
int ghb_copy(ghb_ctx* ctx, void* dst, const void* src, int64_t bytes);
/* Page-lock caller memory in place (cudaHostRegister / cudaHostUnregister) for hosts without a CUDA library of their own:
 * a Julia `Array` registered once is copied from at the PCIe rate by every later call (ghb_condense_assemble_f64 streams
 * pinned records directly, pageable ones through a staging buffer at the host-memcpy rate).  The caller unregisters before
 * the array is freed or moved. */
int ghb_host_register(ghb_ctx* ctx, void* ptr, int64_t bytes);
int ghb_host_unregister(ghb_ctx* ctx, void* ptr);
/* Name of the condensation kernel variant chosen for a plan: "dmma_34_36", "dmma_33_12", "dmma_56_16" (FP64 DMMA,
   interior rows in shared memory, boundary rows in registers), "large_dmma" (64 < n_i <= 128, streamed), "warp_7_8", "warp_16_8" (register-resident),
   "cw_34_36", "cw_33_12", "cw_40_36", "cw_21_16" (one warp per cell, FP64 DMMA, csrc/condense_cw.cu), "generic" (any plan). */
const char* ghb_plan_kernel_name(ghb_ctx* ctx, int plan_id);

/* ---- block plan: mirrors StaticCondensationMap{IFT,BFT} + the touched mask ------------------
 * (src/StaticCondensationMap.jl:3-34 struct + preconditions, :72-84 block sizes)
 * nfields        number of fields of the cell-wise block system
 * ndofs[f]       per-cell dofs of field f (for a skeleton field: nlfacets * ndofs per facet)
 * touched        nfields x nfields, column-major, 1 = block present in the packed record
 * interior[]     1-based field ids condensed out (bulk), boundary[] kept (skeleton); together a
 *                disjoint cover of 1:nfields, else GHB_EINVAL.  A touched mask leaving a block row of
 *                the interior without any touched block is rejected.
 * Packed record of one cell: the touched blocks in block-column-major order (j outer, i inner), each
 * block column-major ndofs[i] x ndofs[j]  => lenA doubles;  b: all field vectors concatenated => lenb.
 */
int ghb_plan_blocks(ghb_ctx* ctx, int nfields, const int32_t* ndofs, const uint8_t* touched, int n_int,
                    const int32_t* interior, int n_bnd, const int32_t* boundary, int* plan_id);
/* out[0..3] = n_i, n_b, lenA, lenb */
int ghb_plan_query(ghb_ctx* ctx, int plan_id, int64_t out[4]);

/* ---- (a3) static condensation: (A_K, b_K) -> (S_K, g_K) for a batch of cells -----------------
 * replaces evaluate!(cache, ::StaticCondensationMap, A, b)   src/StaticCondensationMap.jl:152-196
 * A [ncells][lenA], b [ncells][lenb]  packed records;  S [ncells][n_b*n_b] col-major; g [ncells][n_b];
 * info [ncells] (may be NULL).  keep_factors != 0 additionally stores X = A11^-1 A12 and y = A11^-1 b1
 * inside the ctx for ghb_backsub_f64 (factor reuse); 0 follows the reference (recompute).
 */
int ghb_condense_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b, double* S,
                     double* g, int32_t* info, int keep_factors);
/* Generation of the stored factors (bumped by every keep_factors condensation; -1: none stored): a caller that holds
 * several condensed batches checks it before ghb_backsub_f64(A = b = NULL) to know the factors are still its own. */
int64_t ghb_factors_generation(const ghb_ctx* ctx);

/* ---- (a6) id glue: facet dofs -> cell boundary ids ------------------------------------------
 * replaces RestrictFacetDoFsToSkeleton / restrict_facet_dof_ids_to_cell_boundary
 * (src/HybridAffineFEOperators.jl:388-439): out[c][lf*ndofs_f + d] = facet_data[cell_wise_facets[c][lf]-1][d]
 */
int ghb_restrict_facet_dofs_i64(ghb_ctx* ctx, int64_t ncells, int nlfacets, int ndofs_f,
                                const int64_t* cell_wise_facets, const int64_t* facet_data, int64_t* out);

/* ---- (f-1) element records of an affine family, generated on the device -----------------------
 * On Cartesian / affine meshes with cell-wise constant coefficients the element matrices the reference integrates
 * cell by cell (lazy_map(::IntegrationMap, ...) + SumFacetsMap + Densify, src/GridapAPIExtensions.jl:442-451; weak form
 * test/DarcyHDGTests.jl:125-135) are linear combinations of a few reference records:
 *     A_K = sum_t coef[K][t] * TA[t],   b_K = sum_t coef[K][t] * Tb[t]      (t < ntab <= 16)
 * e.g. Darcy HDG on a Cartesian mesh: TA[0] the interior cell, one table per axis for the owner-normal sign flip of a
 * boundary low-side facet, one per axis for the x-dependence of the load, one for the permeability; the glue gets the
 * tables from the reference's own integration on ntab representative cells (no FE code here) and the coefficients
 * from the cell index.  TA [ntab][lenA], Tb [ntab][lenb], coef [ncells][ntab]; A, b: packed records as for
 * ghb_condense_f64.  The records never cross PCIe; HBM-write bound. */
int ghb_expand_records_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA, const double* Tb,
                           const double* coef, double* A, double* b);

/* The same family condensed WITHOUT materialising the records: the condensation kernel forms A_K, b_K in its loader (a
 * batch of cells per CTA: every table element is fetched once per batch; the records live in per-warp scratch that stays
 * in L2), so the 8(lenA+lenb) bytes per cell never reach HBM -- the element-matrix generation of the reference's
 * lazy_map chain (src/GridapAPIExtensions.jl:442-451 -> src/HybridAffineFEOperators.jl:338) fused into the loader.
 * Results are bit-identical to ghb_expand_records_f64 followed by ghb_condense_f64 / ghb_condense_assemble_f64 (same
 * summation order).  Every plan with a cell-warp kernel takes this path (ghb_plan_kernel_name "cw_<n_i>_<n_b>" or
 * "cw_pad_<class>": n_i <= 64, n_b <= 40; odd record lengths go through zero-padded table rows); the other plans, and
 * table counts whose staging does not fit the image of a small class, expand chunks of records into a device temporary
 * instead (same results).  TA, Tb, coef may be host pointers (staged).  keep_factors as in ghb_condense_f64:
 * X = A11^-1 [A12 | b1] stays in the ctx and ghb_backsub_f64 with A = b = NULL recovers the bulk unknowns from it.
 * ghb_condense_assemble_affine_f64 needs the selected symbolic pattern like
 * ghb_condense_assemble_f64. */
int ghb_condense_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA, const double* Tb,
                            const double* coef, double* S, double* g, int32_t* info, int keep_factors);
int ghb_condense_assemble_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA,
                                     const double* Tb, const double* coef, const double* dirichlet_vals,
                                     int64_t ndirichlet, double* nzval, double* rhs, int32_t* info);
/* ... and the backward map (a11, src/BackwardStaticCondensationMap.jl:84-99) of the same family: u_K = A11^-1 (b1 -
 * A12 lambda_K) with the records formed in the loader; arguments as ghb_backsub_f64.  With
 * ghb_condense_assemble_affine_f64 the whole solve runs without the 8(lenA+lenb) bytes per cell ever existing in HBM.
 * Bit-identical to ghb_expand_records_f64 + ghb_backsub_f64. */
int ghb_backsub_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA, const double* Tb,
                           const double* coef, const double* lambda_free, int64_t nlambda_free,
                           const double* lambda_dirichlet, int64_t nlambda_dirichlet, const int64_t* cell_ids, double* u,
                           int32_t* info);

/* ---- (f-3) bulk -> skeleton L2 projection dofs --------------------------------------------------
 * replaces compute_bulk_to_skeleton_l2_projection_dofs (src/GridapAPIExtensions.jl:453-500; called per (cell, local facet)
 * by the elasticity / Hencky forms through test/P_m.jl:4-23): X = A \ B with A [nbatch][n*n] the facet mass matrices of
 * the skeleton space (column-major) and B [nbatch][n*nrhs] the moments of the bulk basis (nrhs = 1: a FE function).
 * Like Julia's `\` on a square dense matrix: LU-type elimination with partial pivoting; info[s] = k+1 if the k-th pivot
 * column of system s is exactly zero (then X of that system is NaN); info may be NULL.  n <= 128 (one warp per system
 * up to 32, one CTA per system above). */
int ghb_l2_projection_dofs_f64(ghb_ctx* ctx, int64_t nbatch, int n, int nrhs, const double* A, const double* B, double* X,
                               int32_t* info);

/* ---- (a9) in-cell sum over local facets -------------------------------------------------------
 * replaces SumFacetsMap.evaluate! (src/SumFacetsMap.jl:19-30, wired in at src/GridapAPIExtensions.jl:442-451) on the
 * batch: in [ncells][nlfacets][len] = the dK contributions of every local facet already laid out on the cell record
 * (facet-field blocks are disjoint, so their "sum" is placement); out [ncells][len] = in[:,0,:] + in[:,1,:] + ...
 * added left to right like the reference. */
int ghb_sum_facets_f64(ghb_ctx* ctx, int64_t ncells, int nlfacets, int64_t len, const double* in, double* out);

/* ---- (a8) assembly --------------------------------------------------------------------------
 * replaces assemble_matrix_and_vector(SparseMatrixAssembler(M,L), data)
 * (src/HybridAffineFEOperators.jl:38,46) producing SparseMatrixCSC{Float64,Int64} + Vector{Float64}.
 * Symbolic: builds and caches the CSC pattern of  sparse(I,J,V,nrows,nrows)  and the gather map.
 * cell_ids [ncells][n_b] 1-based, <=0 = Dirichlet (not assembled).  A dof may belong to at most 2
 * cells (facet dofs; else GHB_EUNSUPPORTED) and may not repeat inside a cell.
 */
int ghb_assemble_symbolic(ghb_ctx* ctx, int64_t ncells, int n_b, const int64_t* cell_ids, int64_t nrows,
                          int64_t* nnz_out);
/* Symbolic patterns are handles.  Every ghb_assemble_symbolic / ghb_assemble_symbolic_slab creates a new pattern and
 * selects it; ghb_assemble_current returns its handle (>= 0).  An assembler that shares the ctx with others (two
 * operators or meshes, a slab assembler next to a global one, a Newton loop) selects its own pattern before
 * ghb_assemble_pattern / ghb_assemble_numeric* / ghb_condense_assemble_f64; ghb_assemble_release frees one. */
int ghb_assemble_current(const ghb_ctx* ctx);
int ghb_assemble_select(ghb_ctx* ctx, int pattern_id);
int ghb_assemble_release(ghb_ctx* ctx, int pattern_id);
/* Copies the cached pattern out: colptr [nrows+1], rowval [nnz]; 1-based Int64 (Julia CSC). */
int ghb_assemble_pattern(ghb_ctx* ctx, int64_t* colptr, int64_t* rowval);
/* Numeric: nzval [nnz], rhs [nrows].  dirichlet_vals [ndirichlet] (may be NULL) applies the lift of
 * _attach_dirichlet (src/HybridAffineFEOperators.jl:41-42): g_K <- g_K - S_K * vals_K on cells that
 * have a Dirichlet dof, vals_K[l] = dirichlet_vals[-id-1] for id<0, else 0.  Host or device pointer. */
int ghb_assemble_numeric_f64(ghb_ctx* ctx, const double* S, const double* g, const double* dirichlet_vals,
                             int64_t ndirichlet, double* nzval, double* rhs);

/* ---- (f-4) CSR hand-off (SparseMatricesCSR / device sparse solvers instead of the CSC round trip to UMFPACK,
 * src/HybridAffineFEOperators.jl:74, src/HybridLinearSolvers.jl:45) ---------------------------
 * Every cell contributes a full n_b x n_b block, so the pattern of the skeleton matrix is structurally symmetric:
 * the CSR arrays rowptr / colval of A are the colptr / rowval that ghb_assemble_pattern returns (column indices
 * ascending within a row), and the CSR values of A are the CSC values of A^T = sum of the transposed cell blocks.
 * This call computes rhs from S (Dirichlet lift included), transposes every cell block of S IN PLACE (S is left
 * transposed: call it with a scratch copy, or transpose back by a second call's side effect) and gathers nzval in
 * CSR order.  S must be a device pointer; single-GPU patterns only (no ghost cells). */
int ghb_assemble_numeric_csr_f64(ghb_ctx* ctx, double* S, const double* g, const double* dirichlet_vals,
                                 int64_t ndirichlet, double* nzval, double* rhs);
/* ---- (e) multi-GPU slabs: cells are partitioned in contiguous slabs, one process per GPU --------
 * A slab owns the dofs of the facets first touched by its cells = a contiguous range of
 * global columns [col_begin, col_end) (1-based, end exclusive).  Its cell list is its own cells followed
 * by `nghost` ghost cells (the bottom layer of the slab above), whose local facet 0 lies on the cut
 * plane.  Only the leading `ghost_ncols` columns of a ghost cell's S_K (and entries of g_K) exist on
 * this rank: they arrive through ghb_pack_cut_plane_f64 on the sender + an NCCL send/recv.
 * The pattern holds the owned columns only: colptr [ncols_owned+1], rowval = GLOBAL row ids. */
int ghb_assemble_symbolic_slab(ghb_ctx* ctx, int64_t ncells_local, int64_t nghost, int ghost_ncols, int n_b,
                               const int64_t* cell_ids /* [(ncells_local+nghost)][n_b] */, int64_t nrows_global,
                               int64_t col_begin, int64_t col_end, int64_t* nnz_out);
/* Sender side of the cut-plane exchange: for `ncut` consecutive cells starting at S/g/cell_ids, packs
 * out[c] = { S_K[:, 0:ncols] (n_b*ncols, col-major), g_K[0:ncols] - (S_K*vals_K)[0:ncols] }. */
int ghb_pack_cut_plane_f64(ghb_ctx* ctx, int64_t ncut, int n_b, int ncols, const double* S, const double* g,
                           const int64_t* cell_ids, const double* dirichlet_vals, double* out);
/* nzval [nnz owned], rhs [ncols_owned]; `ghost` = the received packed buffer [nghost][n_b*ghost_ncols+ghost_ncols]. */
int ghb_assemble_numeric_slab_f64(ghb_ctx* ctx, const double* S, const double* g, const double* ghost,
                                  const double* dirichlet_vals, double* nzval, double* rhs);

/* Fused slab path (device pointers, plans with a cell-warp kernel): ghb_condense_scatter_slab_f64 condenses the local cells
 * and adds S_K straight into the zeroed nzval (every entry has at most two contributions, so the floating-point atomics
 * are order-independent and bit-equal to the gather); S_K is stored only for the cells with a Dirichlet dof and for
 * the first keep_cut cells (the layer whose cut-plane columns go down), g_K for all; zero_nzval = 0 if the caller has
 * zeroed nzval itself.  After ghb_pack_cut_plane_f64 and
 * the exchange, ghb_assemble_finish_slab_f64 adds the ghost cells' contributions and gathers the rhs. */
int ghb_condense_scatter_slab_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b, double* S,
                                  double* g, int32_t* info, double* nzval, int64_t keep_cut, int zero_nzval);
/* the same first half with the records of an affine family formed in the loader (see ghb_condense_affine_f64): a slab of
 * a multi-GPU run goes from coefficient vectors to its share of nzval without the records existing; device pointers. */
int ghb_condense_scatter_slab_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA,
                                         const double* Tb, const double* coef, double* S, double* g, int32_t* info,
                                         double* nzval, int64_t keep_cut, int zero_nzval);
int ghb_assemble_finish_slab_f64(ghb_ctx* ctx, const double* S, const double* g, const double* ghost,
                                 const double* dirichlet_vals, double* nzval, double* rhs);

/* Condensation + numeric assembly in one call (S_K, g_K never seen by the caller; they live in a device scratch
 * buffer between the two kernels).  Host records are streamed in chunks (option stream_chunk_bytes, 256 MB): the H2D
 * copy of chunk k+1 overlaps the condensation of chunk k, and the columns whose cells are all condensed are assembled
 * and copied back while later records are still arriving.  Pinned caller memory is copied from directly; PAGEABLE
 * caller memory (a Julia Array, a numpy array) is staged through two pinned buffers owned by the ctx, filled by host
 * threads, so that the copies stay asynchronous.  With DEVICE records and a cell-warp plan the two kernels are one
 * (option fused_assembly, default on): S_K goes from the accumulators into nzval, see the slab variant above. */
int ghb_condense_assemble_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b,
                              const double* dirichlet_vals, int64_t ndirichlet, double* nzval, double* rhs,
                              int32_t* info);

/* ---- (a11)+(a12) backward static condensation -----------------------------------------------
 * replaces evaluate!(cache, ::BackwardStaticCondensationMap, A, b, x)
 * (src/BackwardStaticCondensationMap.jl:61-102) fed by get_cell_dof_values(lh, dK)
 * (src/HybridAffineFEOperators.jl:113): lambda_K[l] = lambda_free[id-1] (id>0) or lambda_dirichlet[-id-1].
 * u [ncells][n_i] = interior fields concatenated in `interior` order.  If the last condense call used
 * keep_factors on the same cells, A and b may be NULL and the stored X, y are used (info[] then repeats the info of
 * that condensation).  lambda_free [nlambda_free] and lambda_dirichlet [nlambda_dirichlet] (may be NULL) are host or
 * device pointers.
 */
int ghb_backsub_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b,
                    const double* lambda_free, int64_t nlambda_free, const double* lambda_dirichlet,
                    int64_t nlambda_dirichlet, const int64_t* cell_ids, double* u, int32_t* info);
/* (a12) full-space free-dof vector (src/HybridAffineFEOperators.jl:134-149):
 * x = [bulk field 1 cell-major | bulk field 2 | ... | lambda_free].  x has sum(n_i)*ncells + nlambda entries. */
int ghb_scatter_free_dof_values(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* u,
                                const double* lambda_free, int64_t nlambda, double* x);

/* ---- (e) the two NCCL exchanges behind the C ABI (a rank-per-GPU host in any language; csrc/comm.cu) ---------------
 * NCCL is loaded at run time (libnccl.so.2); without it these return GHB_EUNSUPPORTED and everything else works.
 * ghb_comm_unique_id: rank 0 fills id128 (128 bytes = ncclUniqueId) and ships it to the other ranks with whatever
 * the host has (MPI, a store, a file); ghb_comm_init: every rank, on its ctx's device (ncclCommInitRank).
 * ghb_exchange_cut_plane_f64: collective 1 -- rank r > 0 sends `nsend` doubles (its ghb_pack_cut_plane_f64 buffer) to
 * r-1, rank r < P-1 receives `nrecv` doubles from r+1 (its ghost buffer); one grouped ncclSend/ncclRecv on the ctx's
 * stream.  ghb_allgather_lambda_f64: collective 2 before the backward step (src/HybridAffineFEOperators.jl:113-118) --
 * counts[r] = owned lambda entries of rank r (host array [P]); `all` [sum counts] receives the global vector, `owned`
 * = this rank's range (NULL: already in place inside `all`). */
int ghb_comm_unique_id(ghb_ctx* ctx, void* id128);
int ghb_comm_init(ghb_ctx* ctx, int nranks, int rank, const void* id128);
int ghb_comm_destroy(ghb_ctx* ctx);
int ghb_exchange_cut_plane_f64(ghb_ctx* ctx, const double* send_down, int64_t nsend, double* recv_from_up,
                               int64_t nrecv);
int ghb_allgather_lambda_f64(ghb_ctx* ctx, const double* owned, const int64_t* counts, double* all);

/* ---- synthetic workload (bench / tests): counter-based Philox4x32-10, bit-identical to oracle ---
 * Fills records of cells [cell_start, cell_start+ncells) (oracle.synth_cell_records). */
int ghb_synth_fill_f64(ghb_ctx* ctx, int plan_id, int64_t cell_start, int64_t ncells, uint64_t seed, double* A,
                       double* b);
/* Cartesian mesh glue for the synthetic workload (closed form of Gridap's first-touch facet numbering,
 * SURVEY A1-A3): writes cell_wise_facets [ncells][2*D] for cells [cell_start, cell_start+ncells) of a
 * dims[0] x ... x dims[D-1] mesh (x fastest), 1-based Int64. */
int ghb_cartesian_cell_wise_facets(ghb_ctx* ctx, int D, const int64_t* dims, int64_t cell_start, int64_t ncells,
                                   int64_t* cell_wise_facets);

#ifdef __cplusplus
}
#endif
#endif /* GHB_H */
