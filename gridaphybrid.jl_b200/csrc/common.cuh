// common.cuh -- context, block plan and launch helpers shared by all kernels of libgridaphybrid_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/ghb.h"

namespace ghb {

// Device view of a block plan (mirrors StaticCondensationMap + touched mask,
// /root/reference/src/StaticCondensationMap.jl:3-34,72-84).
struct PlanDev {
  int n_i, n_b, n;      // interior / boundary / total dofs per cell
  int lenA, lenb;       // doubles per packed record
  const int32_t* emap;  // [n*(n+1)] condensed-order (interior rows first) col-major element -> record
                        // offset; -1 = structural zero; column n maps into the b record
};

// Debugging / A-B knobs.  Read from the environment ONCE per context (ghb_create), changed by ghb_set_option, and
// snapshotted into every plan when it is created: nothing on the launch path calls getenv.
struct Options {
  int force_generic = 0;      // GHB_FORCE_GENERIC   plans use the generic kernels only
  int cw = 1;                 // GHB_CW              one-warp-per-cell DMMA kernel for the shapes it is instantiated for
  int dmma_ll = 1;            // GHB_DMMA_LL         left-looking (1) or right-looking (0) 4-warps-per-cell DMMA kernel
  int factors_generic = 0;    // GHB_FACTORS_GENERIC keep_factors through the generic kernel
  int max_ctas_per_sm = 0;    // GHB_MAX_CTAS_PER_SM cap on resident CTAs (occupancy sweeps), 0 = none
  int ll_ctas = 0;            // GHB_LL_CTAS         register budget of the (34,36) left-looking instantiation
  int warp_one_cell = 0;      // GHB_WARP_ONE_CELL   small-cell kernels: one cell per warp
  int warp_two_rows = 0;      // GHB_WARP_TWO_ROWS   (16,8): two rows per lane
  int debug = 0;              // GHB_DEBUG           print launch geometry
  int cw_q4 = 1;              // GHB_CW_Q4           interleaved row tiles / 256-bit accesses for the n_b = 36 shapes (A/B knob)
  int cw_back = 1;            // GHB_CW_BACK         backward map of cell-warp plans on the cell-warp kernel (0: dmma / generic)
  int fused_assembly = 1;     // GHB_FUSED_ASSEMBLY  condensation kernels scatter S_K into nzval themselves (cell-warp plans)
  int64_t stream_chunk_bytes = (int64_t)256 << 20;   // GHB_STREAM_CHUNK_BYTES: chunk of the host-record streaming path
};

// One-time (per kernel instantiation) shared-memory opt-in and occupancy query; `per_sm` < 0: not done yet.
struct KernelSetup { int per_sm = -1; };

struct Plan {
  Options opt;                // snapshot of the context's knobs at plan creation
  int nfields = 0;
  std::vector<int32_t> ndofs, interior, boundary;
  std::vector<uint8_t> touched;        // col-major nfields x nfields
  std::vector<int64_t> block_offset;   // [i + nfields*j] -> offset in record or -1
  std::vector<int32_t> field_offset_b; // offset of field f inside the b record
  std::vector<int32_t> row_field, row_local;  // condensed row -> (field index 0-based, local dof)
  int n_i = 0, n_b = 0, n = 0, lenA = 0, lenb = 0;
  int32_t* d_emap = nullptr;
  int32_t* d_colbase = nullptr;   // tables of the DMMA kernel's on-the-fly re-layout (condense_dmma.cu)
  uint8_t* d_rowf = nullptr;
  uint8_t* d_rowl = nullptr;
  int32_t* d_xoff = nullptr;      // left-looking DMMA kernel: bottom-block record offsets in fragment order
  bool use_dmma = false;
  bool use_cw = false;            // one-warp-per-cell DMMA kernel (condense_cw.cu)
  int cw_pf[6] = {0, 0, 0, 0, 0, 0};   // record ranges (offset, length) of A12 / A21 / A22 for its L2 prefetches
  unsigned char* d_cw = nullptr;  // its tables in one block: loader | rowA12 | colA21 | rowb (byte offsets cw_off)
  int cw_nld = 0;
  size_t cw_off[5] = {0, 0, 0, 0, 0};
  bool cw_q4 = false;             // the plan's blocks allow the 256-bit (interleaved row tile) instantiation
  int cw_pad = 0;                 // 0: tuned instantiation of the exact shape; else the padded n_i class of the generic one
  bool use_warp = false;          // register-resident warp kernels for small cells (condense_warp.cu)
  bool use_large = false;         // streamed large-cell kernel, 64 < n_i <= 128 (condense_large.cu)
  bool all_touched = false;
  const char* kernel_name = "generic";
  PlanDev dev() const { return PlanDev{n_i, n_b, n, lenA, lenb, d_emap}; }
};

// Cached symbolic phase of the assembler (pattern of sparse(I,J,V) + gather map).
struct AsmState {
  bool valid = false;
  int64_t ncells = 0, nrows = 0, nnz = 0;   // ncells = local + ghost; nrows = owned columns
  int64_t ncells_local = 0, nghost = 0, nrows_global = 0, col0 = 0;
  int n_b = 0, ghost_ncols = 0;
  int64_t* d_ids = nullptr;      // [ncells][n_b] copy of cell ids
  int64_t* d_occ = nullptr;      // [nrows][2]  cell*n_b + l of the (<=2) occurrences, cell ascending; -1 none
  uint8_t* d_sorted = nullptr;   // [ncells][n_b] local indices of positive ids sorted by id
  uint8_t* d_npos = nullptr;     // [ncells] number of positive ids
  uint8_t* d_celldir = nullptr;  // [ncells] 1 if the cell has a Dirichlet dof
  int64_t* d_colptr = nullptr;   // [nrows+1] 1-based
  int64_t* d_rowval = nullptr;   // [nnz] 1-based
  uint8_t* d_src = nullptr;      // [nnz][2] local row index inside occurrence 0 / 1, 255 = none
  // scatter map of the fused condensation + assembly (built on first use, asm_scatter_prepare)
  int64_t* d_colpos = nullptr;   // [ncells][n_b] 0-based nzval offset of the column of local dof lj, -1: not an owned column
  uint8_t* d_rowrank = nullptr;  // [ncells][n_b (lj)][n_b (li)] rank of row ids[li] inside that column, 255: not assembled
  uint8_t* d_keepS = nullptr;    // [ncells_local] 1: the fused kernel must also store S_K (Dirichlet lift / cut-plane pack)
  int64_t keep_cut = -1;         // leading cells whose S_K is kept for ghb_pack_cut_plane_f64 (d_keepS was built for it)
  // streaming (host records): columns complete after each chunk of cells, cached per chunk size
  int64_t ready_chunk = 0;
  std::vector<int64_t> ready_J;  // [nchunks] number of complete columns once chunks 0..k are condensed
  std::vector<int64_t> ready_p;  // [nchunks] 0-based nzval offset of column ready_J[k]
};

struct Factors {
  int plan_id = -1;
  int64_t ncells = 0;
  double* d_X = nullptr;  // [ncells][n_i*(n_b+1)] col-major n_i x (n_b+1): A11^-1 [A12 | b1]
  int32_t* d_info = nullptr;   // [ncells] info[] of the condensation that produced X (inside the same allocation)
  uint64_t generation = 0;     // bumped by every keep_factors condensation
  size_t bytes = 0;
};

}  // namespace ghb

struct ghb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t copy_stream = nullptr;  // H2D staging of host-pointer calls
  cudaStream_t d2h_stream = nullptr;   // results of the streaming path go back while records still come in
  int sm_count = 0;
  size_t smem_optin = 0;
  int64_t launches = 0;
  std::string err;
  ghb::Options opt;
  std::vector<ghb::Plan*> plans;
  ghb::AsmState as;                     // the selected symbolic pattern (ghb_assemble_select)
  int as_id = -1;                       // its handle, -1: none
  std::vector<ghb::AsmState> as_store;  // the other patterns, indexed by handle
  ghb::Factors fac;
  // pinned double buffer of the host-record streaming path (pageable caller memory is staged through it)
  void* pinned[2] = {nullptr, nullptr};
  size_t pinned_bytes = 0;
  // NCCL communicator of the multi-GPU exchanges (comm.cu; ncclComm_t, loaded at run time)
  void* comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  // scratch records of the kernels that generate their element records themselves (condense_cw_gen.cu)
  double* gen_scratch = nullptr;
  size_t gen_scratch_bytes = 0;
  uint16_t* gen_need = nullptr;         // chunk list of the GEN + BACK kernels
  double* gen_tab = nullptr;            // padded copy of the tables (odd record lengths / unaligned caller tables)
  size_t gen_tab_bytes = 0;
};

namespace ghb {

int fail(ghb_ctx* ctx, int code, const std::string& msg);

#define GHB_CUDA(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ghb::fail(ctx, GHB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

#define GHB_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != GHB_OK) return _rc; \
  } while (0)

// After a kernel launch: count it and surface launch errors.
#define GHB_LAUNCHED(ctx)                                                                        \
  do {                                                                                           \
    (ctx)->launches++;                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess)                                                                       \
      return ghb::fail(ctx, GHB_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)

bool is_device_ptr(const void* p);

// raises the dynamic shared-memory limit of a kernel only when a launch needs more than any launch before it
#define GHB_SMEM_OPTIN(ctx, kern, bytes)                                                                      \
  do {                                                                                                        \
    static size_t _cur = 0;                                                                                   \
    if ((size_t)(bytes) > _cur) {                                                                             \
      GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      _cur = (size_t)(bytes);                                                                                 \
    }                                                                                                         \
  } while (0)

// cudaFuncSetAttribute + occupancy once per kernel instantiation; returns the resident CTAs per SM to launch with
template <typename K>
int kernel_setup(ghb_ctx* ctx, const Options& opt, K kern, int threads, size_t smem, int carveout, KernelSetup& ks,
                 const char* name, int* per_sm_out) {
  if (ks.per_sm < 0) {
    GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (carveout) GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout == 1 ? 100 : carveout));
    int per_sm = 0;
    GHB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) return fail(ctx, GHB_ECUDA, std::string(name) + " does not fit on an SM");
    if (opt.debug) fprintf(stderr, "%s: %d CTAs/SM x %d threads, %zu B smem\n", name, per_sm, threads, smem);
    ks.per_sm = per_sm;
  }
  *per_sm_out = opt.max_ctas_per_sm > 0 ? std::max(1, std::min(ks.per_sm, opt.max_ctas_per_sm)) : ks.per_sm;
  return GHB_OK;
}

// RAII device view of a caller array that may live on host or device.
template <typename T>
struct Arg {
  ghb_ctx* ctx;
  T* host = nullptr;   // non-null if the caller pointer is a host pointer
  T* dev = nullptr;
  size_t count = 0;
  bool out = false, owned = false;
  int rc = GHB_OK;
  Arg(ghb_ctx* c, const T* p, size_t n, bool in, bool out_) : ctx(c), count(n), out(out_) {
    if (p == nullptr || n == 0) { dev = const_cast<T*>(p); return; }
    if (is_device_ptr(p)) { dev = const_cast<T*>(p); return; }
    host = const_cast<T*>(p);
    cudaError_t e = cudaMallocAsync((void**)&dev, n * sizeof(T), c->stream);
    if (e != cudaSuccess) { rc = fail(c, GHB_ENOMEM, std::string("cudaMallocAsync: ") + cudaGetErrorString(e)); dev = nullptr; return; }
    owned = true;
    if (in) {
      e = cudaMemcpyAsync(dev, host, n * sizeof(T), cudaMemcpyHostToDevice, c->stream);
      if (e != cudaSuccess) rc = fail(c, GHB_ECUDA, std::string("H2D: ") + cudaGetErrorString(e));
    }
  }
  // copy results back (if host) -- call explicitly so errors can be returned
  int finish() {
    if (owned && out && rc == GHB_OK) {
      cudaError_t e = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) return fail(ctx, GHB_ECUDA, std::string("D2H: ") + cudaGetErrorString(e));
      e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) return fail(ctx, GHB_ECUDA, std::string("sync: ") + cudaGetErrorString(e));
    }
    return GHB_OK;
  }
  ~Arg() {
    if (owned && dev) cudaFreeAsync(dev, ctx->stream);
  }
  Arg(const Arg&) = delete;
  Arg& operator=(const Arg&) = delete;
};

// kernels / launchers implemented in the other translation units
int launch_condense_generic(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                            double* S, double* g, int32_t* info, double* X);
const char* warp_kernel_name(const Plan& p);
int launch_condense_warp(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                         double* g, int32_t* info);
int launch_backsub_dmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                        const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info);
int launch_backsub_warp(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                        const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info);
bool dmma_supported(const Plan& p);
bool large_supported(const ghb_ctx* ctx, const Plan& p);
int launch_condense_large(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                          double* g, int32_t* info, double* X);
int launch_backsub_large(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                         const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info);
int dmma_prepare(ghb_ctx* ctx, Plan& p);
bool cw_supported(const Plan& p);
bool cw_pad_supported(const Plan& p);
int cw_pad_class_of(const Plan& p);
const char* cw_kernel_name(const Plan& p);
int cw_prepare(ghb_ctx* ctx, Plan& p);
int launch_condense_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                       double* g, int32_t* info, double* X);
int launch_condense_dmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                         double* g, int32_t* info, double* X = nullptr);
int launch_backsub_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, const double* lam_free,
                      const double* lam_dir, const int64_t* ids, double* u, int32_t* info);
int launch_backsub(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, const double* lam_free,
                   const double* lam_dir, const int64_t* ids, double* u, int32_t* info);
int launch_backsub_cw_gen(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                          const double* coef, const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                          int32_t* info);
struct ScatterArgs;
// records of an affine family generated inside the condensation kernel (sc != NULL: fused assembly as well)
bool cw_gen_supported(const Plan& p, int ntab);
int launch_condense_cw_gen(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                           const double* coef, double* S, double* g, int32_t* info, double* X, const ScatterArgs* sc);
// dispatch: tuned kernel when the plan has one and no factors are requested, else the generic kernel
int launch_condense(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                    double* g, int32_t* info, double* X);
int launch_backsub_generic(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                           const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                           int32_t* info);
int launch_backsub_factors(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* X, const double* lam_free,
                           const double* lam_dir, const int64_t* ids, double* u);
int asm_symbolic(ghb_ctx* ctx, int64_t ncells_local, int64_t nghost, int ghost_ncols, int n_b, const int64_t* d_ids,
                 int64_t nrows_global, int64_t col0, int64_t ncols);
enum { ASM_MATRIX = 1, ASM_RHS = 2 };   // which: phases of the numeric assembly
int asm_numeric_range(ghb_ctx* ctx, const double* S, const double* g, const double* ghost, const double* dvals,
                      double* nzval, double* rhs, int64_t j0, int64_t j1, int which = ASM_MATRIX | ASM_RHS);
int asm_ready_columns(ghb_ctx* ctx, int64_t chunk, int nchunks);
// fused path: scatter map of the selected pattern (lazily), flags of the cells whose S_K must be kept; ghost scatter
int asm_scatter_prepare(ghb_ctx* ctx, int64_t keep_cut);
int asm_scatter_ghosts(ghb_ctx* ctx, const double* ghost, double* nzval);
struct ScatterArgs {             // epilogue of the condensation kernel in fused mode
  double* nzval;
  const int64_t* colpos;
  const uint8_t* rowrank;
  const uint8_t* keepS;
};
int launch_condense_cw_scatter(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                               double* g, int32_t* info, const ScatterArgs& sc);
int asm_numeric(ghb_ctx* ctx, const double* S, const double* g, const double* ghost, const double* dvals,
                double* nzval, double* rhs);
int asm_pack_cut_plane(ghb_ctx* ctx, int64_t ncut, int n_b, int ncols, const double* S, const double* g,
                       const int64_t* ids, const double* dvals, double* out);
void asm_free(ghb_ctx* ctx);
void comm_free(ghb_ctx* ctx);
int launch_restrict_facet_dofs(ghb_ctx* ctx, int64_t ncells, int nlf, int nf, const int64_t* cwf,
                               const int64_t* fdata, int64_t* out);
int launch_sum_facets(ghb_ctx* ctx, int64_t ncells, int nlf, int64_t len, const double* in, double* out);
int launch_transpose_blocks(ghb_ctx* ctx, int64_t ncells, int n, double* S);
int launch_batched_solve(ghb_ctx* ctx, int64_t nbatch, int n, int m, const double* A, const double* B, double* X,
                         int32_t* info);
int launch_expand_records(ghb_ctx* ctx, int64_t ncells, int len, int ntab, const double* T, const double* coef,
                          double* out);
int launch_scatter_free(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* u, const double* lam,
                        int64_t nlam, double* x);
int launch_synth_fill(ghb_ctx* ctx, const Plan& p, int64_t cell_start, int64_t ncells, uint64_t seed, double* A,
                      double* b);
int launch_cartesian_facets(ghb_ctx* ctx, int D, const int64_t* dims, int64_t cell_start, int64_t ncells,
                            int64_t* out);

}  // namespace ghb
