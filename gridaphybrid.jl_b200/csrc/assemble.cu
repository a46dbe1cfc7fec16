// assemble.cu -- skeleton assembly: symbolic phase (CSC pattern of Julia's sparse(I,J,V,m,n)) and the
// atomic-free, owner-computes numeric phase.
//
// Reference semantics (Gridap SparseMatrixAssembler; call sites
// /root/reference/src/HybridAffineFEOperators.jl:38,46 and src/HybridLinearSolvers.jl:43-44):
//   cells ascending, `for lj, for li`: push (ids[li], ids[lj], S[li,lj]) when both ids > 0;
//   b[ids[li]] += g[li]; then sparse(I,J,V,m,n): columns ascending, rows ascending inside a column,
//   duplicates summed, stored zeros kept.
// Facet dofs belong to at most two cells, so column j of the result is the sorted union of the positive
// ids of the (<=2) cells that contain j.  The symbolic phase stores, per dof, its <=2 occurrences
// (cell ascending) and, per cell, the order of its positive ids; a column is then a two-pointer merge.
// The numeric phase *gathers*: the thread that owns a stored entry adds its <=2 contributions in cell
// order -- no atomics, bitwise reproducible, and identical to the reference's left-to-right sum.
#include <algorithm>

#include "common.cuh"

namespace ghb {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

// ---- exclusive scan of int32 counts into int64 offsets (three small kernels) ------------------
__global__ void scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
  __shared__ int64_t red[kScanThreads / 32];
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t s = 0;
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + threadIdx.x + (int64_t)k * kScanThreads;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void scan_tile_offsets(int64_t* tile_sums, int64_t ntiles, int64_t* total) {
  // single block: sequential chunks of blockDim.x with a block-wide scan each
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < ntiles; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    int64_t v = i < ntiles ? tile_sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      int64_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < ntiles) tile_sums[i] = carry + buf[threadIdx.x] - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += buf[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// out[i] = one_based + exclusive prefix; also writes out[n] = one_based + total
__global__ void scan_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_offsets,
                           int64_t* __restrict__ out, int64_t one_based) {
  __shared__ int64_t wsum[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  // warp inclusive scan of thread sums
  int64_t inc = s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int64_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  int64_t run = tile_offsets[blockIdx.x] + woff + inc - s + one_based;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
    if (i == n - 1) out[n] = run;
  }
}

// ---- symbolic kernels --------------------------------------------------------------------------
// occurrences of every positive dof: at most two cells, kept cell-ascending; also per-cell checks
// Columns [col0, col0+ncols) (0-based dof index) are owned by this slab; ids outside are rows only.
// Ghost cells (index >= ncells_local) may touch an owned column only through their leading ghost_ncols
// local dofs (the cut-plane facet block whose S columns were exchanged).
__global__ void occ_count_kernel(int64_t ncells, int64_t ncells_local, int ghost_ncols, int n_b,
                                 const int64_t* __restrict__ ids, int64_t nrows, int64_t col0, int64_t ncols,
                                 int32_t* __restrict__ cnt, int* __restrict__ err) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncells * n_b) return;
  int64_t id = ids[t];
  if (id > 0) {
    if (id > nrows) { atomicExch(err, 1); return; }
    int64_t k = id - 1 - col0;
    if (k < 0 || k >= ncols) return;
    int64_t cell = t / n_b;
    if (cell >= ncells_local && (int)(t - cell * n_b) >= ghost_ncols) { atomicExch(err, 4); return; }
    int old = atomicAdd(&cnt[k], 1);
    if (old >= 2) atomicExch(err, 2);
  }
}

__global__ void occ_fill_kernel(int64_t ncells, int n_b, const int64_t* __restrict__ ids, int64_t col0,
                                int64_t ncols, unsigned long long* __restrict__ occ) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncells * n_b) return;
  int64_t id = ids[t] - col0;
  if (id > 0 && id <= ncols) {
    // two slots per dof, initialised to ~0; smaller (cell*n_b+l) ends in slot 0: deterministic
    unsigned long long v = (unsigned long long)t;
    unsigned long long prev = atomicMin(&occ[2 * (id - 1)], v);
    if (prev != ~0ull) {
      unsigned long long larger = prev > v ? prev : v;
      atomicMin(&occ[2 * (id - 1) + 1], larger);
    }
  }
}

// per cell: order of positive ids (rank by counting), Dirichlet flag, duplicate check
__global__ void cell_sort_kernel(int64_t ncells, int n_b, const int64_t* __restrict__ ids,
                                 uint8_t* __restrict__ sorted, uint8_t* __restrict__ npos,
                                 uint8_t* __restrict__ celldir, int* __restrict__ err) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncells * n_b) return;
  int64_t cell = t / n_b;
  int li = (int)(t - cell * n_b);
  const int64_t* cid = ids + cell * n_b;
  int64_t id = cid[li];
  int rank = 0, np = 0, dir = 0;
  for (int l = 0; l < n_b; ++l) {
    int64_t o = cid[l];
    if (o > 0) {
      ++np;
      if (o < id) ++rank;
      if (o == id && l != li) atomicExch(err, 3);
    } else if (o < 0) {
      dir = 1;
    }
  }
  if (id > 0) sorted[cell * n_b + rank] = (uint8_t)li;
  if (li == 0) { npos[cell] = (uint8_t)np; celldir[cell] = (uint8_t)dir; }
}

// Column merge.  FILL=false: count distinct rows of column j.  FILL=true: write rowval and the gather map.
template <bool FILL>
__global__ void column_merge_kernel(int64_t nrows, int n_b, const int64_t* __restrict__ ids,
                                    const unsigned long long* __restrict__ occ, const uint8_t* __restrict__ sorted,
                                    const uint8_t* __restrict__ npos, int32_t* __restrict__ colcnt,
                                    const int64_t* __restrict__ colptr, int64_t* __restrict__ rowval,
                                    uint8_t* __restrict__ src) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nrows) return;
  unsigned long long o0 = occ[2 * j], o1 = occ[2 * j + 1];
  if (o0 == ~0ull) { if (!FILL) colcnt[j] = 0; return; }
  int64_t c0 = (int64_t)(o0 / n_b);
  const int64_t* id0 = ids + c0 * n_b;
  const uint8_t* s0 = sorted + c0 * n_b;
  int n0 = npos[c0];
  int64_t p = FILL ? colptr[j] - 1 : 0;
  if (o1 == ~0ull) {
    if constexpr (!FILL) {
      colcnt[j] = n0;
    } else {
      for (int a = 0; a < n0; ++a) {
        rowval[p] = id0[s0[a]];
        src[2 * p] = s0[a]; src[2 * p + 1] = 255;
        ++p;
      }
    }
    return;
  }
  int64_t c1 = (int64_t)(o1 / n_b);
  const int64_t* id1 = ids + c1 * n_b;
  const uint8_t* s1 = sorted + c1 * n_b;
  int n1 = npos[c1];
  int a = 0, bq = 0, cnt = 0;
  while (a < n0 || bq < n1) {
    int64_t va = a < n0 ? id0[s0[a]] : INT64_MAX;
    int64_t vb = bq < n1 ? id1[s1[bq]] : INT64_MAX;
    int64_t v = va < vb ? va : vb;
    if (FILL) {
      rowval[p] = v;
      src[2 * p] = va == v ? s0[a] : 255;
      src[2 * p + 1] = vb == v ? s1[bq] : 255;
      ++p;
    }
    if (va == v) ++a;
    if (vb == v) ++bq;
    ++cnt;
  }
  if (!FILL) colcnt[j] = cnt;
}

// ---- numeric kernels ---------------------------------------------------------------------------
// one warp per column: lanes own the stored entries of the column
struct GhostView {
  int64_t ncells_local;   // cells >= ncells_local are ghosts
  const double* G;        // [nghost][stride]: leading ghost_ncols columns of S (n_b x ghost_ncols), then g[0:ghost_ncols]
  int64_t stride;
};

__device__ __forceinline__ const double* s_column(const double* S, const GhostView& gv, int64_t c, int l, int n_b) {
  return c < gv.ncells_local ? S + c * (int64_t)n_b * n_b + (int64_t)l * n_b
                             : gv.G + (c - gv.ncells_local) * gv.stride + (int64_t)l * n_b;
}

#ifndef GHB_GATHER_MINB
#define GHB_GATHER_MINB 8
#endif
#ifndef GHB_GATHER_UNR
#define GHB_GATHER_UNR 5
#endif

// (cell, local dof) of an occurrence code o = cell*n_b + l without a 64-bit division: M = ceil(2^64 / n_b),
// exact for o < 2^56 and n_b <= 255 (the error term o*(M*n_b - 2^64) stays below 2^64)
__device__ __forceinline__ void occ_decode(unsigned long long o, unsigned long long M, int n_b, int64_t& c, int& l) {
  const unsigned long long q = n_b == 1 ? o : __umul64hi(o, M);
  c = (int64_t)q;
  l = (int)(o - q * (unsigned long long)n_b);
}

// GL lanes per column: a facet column of C3 has <= 66 entries, so 16 lanes waste less than 32 and double the
// number of independent columns in flight (17 ms -> 12 ms at 128^3).  The kernel is a chain of dependent DRAM
// accesses per column (colptr/occ -> gather map -> S -> store): the metadata of the group's next column is fetched
// while the current one is processed, the gather-map bytes of UNR entries are read before their S values, and the
// kernel is compiled for 8 resident blocks per SM (tools/time_assembly.py: 13.8 -> 11.7 ms at 128^3).  Loading the
// two S columns whole and permuting them through shared memory measured no better (13.9 ms).
template <int GL>
__global__ void __launch_bounds__(256, GHB_GATHER_MINB) gather_nzval_kernel(int64_t j0, int64_t nrows, int n_b, const int64_t* __restrict__ colptr,
                                                           const unsigned long long* __restrict__ occ,
                                                           const uint8_t* __restrict__ src,
                                                           const double* __restrict__ S, GhostView gv,
                                                           double* __restrict__ nzval) {
  constexpr int UNR = GHB_GATHER_UNR;
  const int gl = threadIdx.x % GL;
  const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / GL;
  int64_t j = j0 + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GL;     // columns [j0, nrows)
  if (j >= nrows) return;
  const unsigned long long M = n_b > 1 ? 0xffffffffffffffffull / (unsigned)n_b + 1ull : 0ull;
  const unsigned short* __restrict__ src16 = reinterpret_cast<const unsigned short*>(src);   // (a, b) byte pairs
  int64_t p0 = colptr[j] - 1, p1 = colptr[j + 1] - 1;
  unsigned long long o0 = occ[2 * j], o1 = occ[2 * j + 1];
  while (true) {
    const int64_t jn = j + ngrp;
    int64_t np0 = 0, np1 = 0;
    unsigned long long no0 = ~0ull, no1 = ~0ull;
    if (jn < nrows) { np0 = colptr[jn] - 1; np1 = colptr[jn + 1] - 1; no0 = occ[2 * jn]; no1 = occ[2 * jn + 1]; }
    if (p0 != p1) {
      int64_t c0, c1 = 0;
      int l0, l1 = 0;
      occ_decode(o0, M, n_b, c0, l0);
      const double* col0 = s_column(S, gv, c0, l0, n_b);
      const double* col1 = col0;
      if (o1 != ~0ull) { occ_decode(o1, M, n_b, c1, l1); col1 = s_column(S, gv, c1, l1, n_b); }
      const int len = (int)(p1 - p0);                      // entries of this column
      const unsigned short* sp = src16 + p0;
      double* nz = nzval + p0;
      for (int pb = gl; pb < len; pb += UNR * GL) {
        unsigned sc[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) sc[q] = pb + q * GL < len ? sp[pb + q * GL] : 0xffffu;
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
          const unsigned a = sc[q] & 0xffu, bq = sc[q] >> 8;
          const double va = a != 255u ? col0[a] : 0.0;
          const double vb = bq != 255u ? col1[bq] : 0.0;
          // cell-ascending order, as the reference's COO sum
          if (pb + q * GL < len) nz[pb + q * GL] = a != 255u ? (bq != 255u ? va + vb : va) : vb;
        }
      }
    }
    if (jn >= nrows) break;
    j = jn; p0 = np0; p1 = np1; o0 = no0; o1 = no1;
  }
}

// rhs gather with the Dirichlet lift g_K - S_K*vals_K (AttachDirichletMap)
// (ghost contributions arrive already lifted by the sender, see pack_cut_plane_kernel)
__global__ void gather_rhs_kernel(int64_t i0, int64_t nrows, int n_b, const unsigned long long* __restrict__ occ,
                                  const int64_t* __restrict__ ids, const uint8_t* __restrict__ celldir,
                                  const double* __restrict__ S, const double* __restrict__ g, GhostView gv,
                                  int ghost_ncols, const double* __restrict__ dvals, double* __restrict__ rhs) {
  int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           // rows [i0, nrows)
  if (i >= nrows) return;
  double acc = 0.0;
  bool first = true;
  for (int q = 0; q < 2; ++q) {
    unsigned long long o = occ[2 * i + q];
    if (o == ~0ull) break;
    int64_t c = (int64_t)(o / n_b);
    int l = (int)(o - (unsigned long long)c * n_b);
    double v;
    if (c >= gv.ncells_local) {
      v = gv.G[(c - gv.ncells_local) * gv.stride + (int64_t)n_b * ghost_ncols + l];
    } else {
      v = g[c * n_b + l];
    }
    if (c < gv.ncells_local && dvals && celldir[c]) {
      const int64_t* cid = ids + c * n_b;
      const double* Sc = S + c * (int64_t)n_b * n_b;
      for (int lj = 0; lj < n_b; ++lj) {
        int64_t id = cid[lj];
        if (id < 0) v = fma(-dvals[-id - 1], Sc[l + (int64_t)lj * n_b], v);
      }
    }
    acc = first ? v : acc + v;
    first = false;
  }
  rhs[i] = acc;
}

// Cut-plane export of a slab's bottom-layer cells to the slab below (collective 1): the
// leading `ncols` columns of S_K (the dofs of local facet 0, which the lower slab owns) and the matching
// entries of g_K with the Dirichlet lift already applied.
__global__ void pack_cut_plane_kernel(int64_t ncut, int n_b, int ncols, const double* __restrict__ S,
                                      const double* __restrict__ g, const int64_t* __restrict__ ids,
                                      const double* __restrict__ dvals, double* __restrict__ out) {
  const int64_t stride = (int64_t)n_b * ncols + ncols;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncut * stride) return;
  int64_t c = t / stride;
  int r = (int)(t - c * stride);
  const double* Sc = S + c * (int64_t)n_b * n_b;
  if (r < n_b * ncols) { out[t] = Sc[r]; return; }
  int l = r - n_b * ncols;
  double v = g[c * n_b + l];
  if (dvals) {
    const int64_t* cid = ids + c * n_b;
    for (int lj = 0; lj < n_b; ++lj) {
      int64_t id = cid[lj];
      if (id < 0) v = fma(-dvals[-id - 1], Sc[l + (int64_t)lj * n_b], v);
    }
  }
  out[t] = v;
}

int exclusive_scan(ghb_ctx* ctx, const int32_t* d_in, int64_t n, int64_t* d_out, int64_t one_based, int64_t* h_total) {
  int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  int64_t* d_tiles = nullptr;
  GHB_CUDA(ctx, cudaMallocAsync((void**)&d_tiles, (ntiles + 1) * sizeof(int64_t), ctx->stream));
  scan_tile_sums<<<(unsigned)ntiles, kScanThreads, 0, ctx->stream>>>(d_in, n, d_tiles);
  GHB_LAUNCHED(ctx);
  scan_tile_offsets<<<1, 1024, 0, ctx->stream>>>(d_tiles, ntiles, d_tiles + ntiles);
  GHB_LAUNCHED(ctx);
  scan_apply<<<(unsigned)ntiles, kScanThreads, 0, ctx->stream>>>(d_in, n, d_tiles, d_out, one_based);
  GHB_LAUNCHED(ctx);
  if (h_total) {
    GHB_CUDA(ctx, cudaMemcpyAsync(h_total, d_tiles + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  cudaFreeAsync(d_tiles, ctx->stream);
  return GHB_OK;
}

}  // namespace

void asm_free(ghb_ctx* ctx) {
  AsmState& as = ctx->as;
  cudaFree(as.d_ids); cudaFree(as.d_occ); cudaFree(as.d_sorted); cudaFree(as.d_npos); cudaFree(as.d_celldir);
  cudaFree(as.d_colptr); cudaFree(as.d_rowval); cudaFree(as.d_src);
  cudaFree(as.d_colpos); cudaFree(as.d_rowrank); cudaFree(as.d_keepS);
  as = AsmState();
}

int asm_symbolic(ghb_ctx* ctx, int64_t ncells_local, int64_t nghost, int ghost_ncols, int n_b, const int64_t* d_ids,
                 int64_t nrows_global, int64_t col0, int64_t ncols) {
  AsmState& as = ctx->as;
  const int64_t ncells = ncells_local + nghost;
  const int64_t nrows = ncols;   // owned columns (== rows of the local rhs)
  as.ncells = ncells; as.ncells_local = ncells_local; as.nghost = nghost; as.ghost_ncols = ghost_ncols;
  as.n_b = n_b; as.nrows = nrows; as.nrows_global = nrows_global; as.col0 = col0;
  as.ready_chunk = 0; as.ready_J.clear(); as.ready_p.clear();
  const int64_t nent = ncells * n_b;
  const unsigned eb = (unsigned)((nent + 255) / 256), rb = (unsigned)((nrows + 255) / 256);
  int32_t* d_cnt = nullptr;
  int* d_err = nullptr;
  GHB_CUDA(ctx, cudaMallocAsync((void**)&d_cnt, nrows * sizeof(int32_t), ctx->stream));
  GHB_CUDA(ctx, cudaMallocAsync((void**)&d_err, sizeof(int), ctx->stream));
  GHB_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, nrows * sizeof(int32_t), ctx->stream));
  GHB_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_occ, 2 * nrows * sizeof(int64_t)));
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_sorted, nent));
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_npos, ncells));
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_celldir, ncells));
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_colptr, (nrows + 1) * sizeof(int64_t)));
  GHB_CUDA(ctx, cudaMemsetAsync(as.d_occ, 0xff, 2 * nrows * sizeof(int64_t), ctx->stream));
  occ_count_kernel<<<eb, 256, 0, ctx->stream>>>(ncells, ncells_local, ghost_ncols, n_b, d_ids, nrows_global, col0, ncols, d_cnt, d_err);
  GHB_LAUNCHED(ctx);
  cell_sort_kernel<<<eb, 256, 0, ctx->stream>>>(ncells, n_b, d_ids, as.d_sorted, as.d_npos, as.d_celldir, d_err);
  GHB_LAUNCHED(ctx);
  int h_err = 0;
  GHB_CUDA(ctx, cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_err) {
    cudaFreeAsync(d_cnt, ctx->stream); cudaFreeAsync(d_err, ctx->stream);
    if (h_err == 1) return fail(ctx, GHB_EINVAL, "ghb_assemble_symbolic: a cell id exceeds nrows");
    if (h_err == 2) return fail(ctx, GHB_EUNSUPPORTED, "ghb_assemble_symbolic: a dof belongs to more than 2 cells (only facet dofs are supported)");
    if (h_err == 4) return fail(ctx, GHB_EINVAL, "ghb_assemble_symbolic_slab: a ghost cell touches an owned column outside its leading ghost_ncols dofs");
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_assemble_symbolic: a dof id repeats inside one cell");
  }
  occ_fill_kernel<<<eb, 256, 0, ctx->stream>>>(ncells, n_b, d_ids, col0, ncols, (unsigned long long*)as.d_occ);
  GHB_LAUNCHED(ctx);
  // pass 1: column counts (reuse d_cnt), scan -> colptr (1-based)
  column_merge_kernel<false><<<rb, 256, 0, ctx->stream>>>(nrows, n_b, d_ids, (const unsigned long long*)as.d_occ,
                                                          as.d_sorted, as.d_npos, d_cnt, nullptr, nullptr, nullptr);
  GHB_LAUNCHED(ctx);
  int64_t nnz = 0;
  GHB_TRY(exclusive_scan(ctx, d_cnt, nrows, as.d_colptr, 1, &nnz));
  as.nnz = nnz;
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_rowval, std::max<int64_t>(nnz, 1) * sizeof(int64_t)));
  GHB_CUDA(ctx, cudaMalloc((void**)&as.d_src, std::max<int64_t>(nnz, 1) * 2));
  column_merge_kernel<true><<<rb, 256, 0, ctx->stream>>>(nrows, n_b, d_ids, (const unsigned long long*)as.d_occ,
                                                         as.d_sorted, as.d_npos, nullptr, as.d_colptr, as.d_rowval, as.d_src);
  GHB_LAUNCHED(ctx);
  cudaFreeAsync(d_cnt, ctx->stream); cudaFreeAsync(d_err, ctx->stream);
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  as.valid = true;
  return GHB_OK;
}

// ---- fused condensation + assembly: scatter map ---------------------------------------------------------------
// Every stored entry of the skeleton matrix has at most TWO contributions (a facet has two cells), so adding them with
// floating-point atomics into a zeroed nzval is bit-reproducible and equal to the reference's COO sum: 0 + a + b and
// 0 + b + a are the same number.  The condensation kernel therefore scatters S_K straight from its accumulators
// (red.global.add.f64) instead of storing it for a second kernel to gather.  One thread per (cell, local column): the same
// two-pointer merge of the two cells' sorted ids as column_merge_kernel, recording where THIS cell's rows land.
__global__ void scatter_map_kernel(int64_t ncells, int n_b, const int64_t* __restrict__ ids,
                                   const unsigned long long* __restrict__ occ, const uint8_t* __restrict__ sorted,
                                   const uint8_t* __restrict__ npos, const int64_t* __restrict__ colptr, int64_t col0,
                                   int64_t ncols, int64_t* __restrict__ colpos, uint8_t* __restrict__ rowrank) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncells * n_b) return;
  const int64_t c = t / n_b;
  uint8_t* rr = rowrank + t * n_b;
  for (int l = 0; l < n_b; ++l) rr[l] = 255;
  const int64_t j = ids[t] - col0 - 1;                     // owned column index (0-based) or outside
  if (ids[t] <= 0 || j < 0 || j >= ncols) { colpos[t] = -1; return; }
  colpos[t] = colptr[j] - 1;
  const unsigned long long o0 = occ[2 * j], o1 = occ[2 * j + 1];
  const int64_t c0 = (int64_t)(o0 / n_b);
  const int64_t* id0 = ids + c0 * n_b;
  const uint8_t* s0 = sorted + c0 * n_b;
  const int n0 = npos[c0];
  if (o1 == ~0ull) {
    for (int a = 0; a < n0; ++a) rr[s0[a]] = (uint8_t)a;     // c0 == c: the only occurrence, plain stores
    return;
  }
  const int64_t c1 = (int64_t)(o1 / n_b);
  const int64_t* id1 = ids + c1 * n_b;
  const uint8_t* s1 = sorted + c1 * n_b;
  const int n1 = npos[c1];
  int a = 0, bq = 0, pos = 0;
  while (a < n0 || bq < n1) {
    const int64_t va = a < n0 ? id0[s0[a]] : INT64_MAX;
    const int64_t vb = bq < n1 ? id1[s1[bq]] : INT64_MAX;
    const int64_t v = va < vb ? va : vb;
    // bit 7: both cells have this row (the facet block they share): the only entries that need an atomic add; every
    // other entry of the column has ONE contribution and is stored
    const unsigned sh = (va == v && vb == v) ? 0x80u : 0u;
    if (va == v) { if (c0 == c) rr[s0[a]] = (uint8_t)(pos | sh); ++a; }
    if (vb == v) { if (c1 == c) rr[s1[bq]] = (uint8_t)(pos | sh); ++bq; }
    ++pos;
  }
}

__global__ void keep_flags_kernel(int64_t ncells_local, int64_t keep_cut, const uint8_t* __restrict__ celldir,
                                  uint8_t* __restrict__ keep) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncells_local) keep[c] = (uint8_t)((celldir[c] != 0) || c < keep_cut);
}

// contributions of the ghost cells (the packed cut-plane columns received from the slab above): one thread per entry
__global__ void scatter_ghost_kernel(int64_t nghost, int64_t ncells_local, int n_b, int ncols, const double* __restrict__ G,
                                     int64_t stride, const int64_t* __restrict__ colpos, const uint8_t* __restrict__ rowrank,
                                     double* __restrict__ nzval) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)n_b * ncols;
  if (t >= nghost * per) return;
  const int64_t gc = t / per;
  const int e = (int)(t - gc * per);
  const int lj = e / n_b, li = e - lj * n_b;
  const int64_t c = ncells_local + gc;
  const int64_t cp = colpos[c * n_b + lj];
  const uint8_t rk = rowrank[(c * n_b + lj) * n_b + li];
  if (cp >= 0 && rk != 255) atomicAdd(nzval + cp + (rk & 0x7f), G[gc * stride + e]);
}

int asm_scatter_prepare(ghb_ctx* ctx, int64_t keep_cut) {
  AsmState& as = ctx->as;
  if (as.n_b > 63) return fail(ctx, GHB_EUNSUPPORTED, "fused assembly: n_b > 63 (ranks are 7 bits)");
  if (!as.d_colpos) {
    const int64_t nent = as.ncells * as.n_b;
    GHB_CUDA(ctx, cudaMalloc((void**)&as.d_colpos, nent * sizeof(int64_t)));
    GHB_CUDA(ctx, cudaMalloc((void**)&as.d_rowrank, nent * as.n_b));
    GHB_CUDA(ctx, cudaMalloc((void**)&as.d_keepS, std::max<int64_t>(as.ncells_local, 1)));
    scatter_map_kernel<<<(unsigned)((nent + 127) / 128), 128, 0, ctx->stream>>>(
        as.ncells, as.n_b, as.d_ids, (const unsigned long long*)as.d_occ, as.d_sorted, as.d_npos, as.d_colptr, as.col0, as.nrows,
        as.d_colpos, as.d_rowrank);
    GHB_LAUNCHED(ctx);
    as.keep_cut = -1;
  }
  if (as.keep_cut != keep_cut) {
    keep_flags_kernel<<<(unsigned)((as.ncells_local + 255) / 256), 256, 0, ctx->stream>>>(as.ncells_local, keep_cut, as.d_celldir, as.d_keepS);
    GHB_LAUNCHED(ctx);
    as.keep_cut = keep_cut;
  }
  return GHB_OK;
}

int asm_scatter_ghosts(ghb_ctx* ctx, const double* ghost, double* nzval) {
  const AsmState& as = ctx->as;
  if (as.nghost == 0 || as.ghost_ncols == 0) return GHB_OK;
  const int64_t tot = as.nghost * (int64_t)as.n_b * as.ghost_ncols;
  scatter_ghost_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(
      as.nghost, as.ncells_local, as.n_b, as.ghost_ncols, ghost, (int64_t)as.n_b * as.ghost_ncols + as.ghost_ncols, as.d_colpos,
      as.d_rowrank, nzval);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int asm_pack_cut_plane(ghb_ctx* ctx, int64_t ncut, int n_b, int ncols, const double* S, const double* g,
                       const int64_t* ids, const double* dvals, double* out) {
  int64_t tot = ncut * ((int64_t)n_b * ncols + ncols);
  if (tot == 0) return GHB_OK;
  pack_cut_plane_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ncut, n_b, ncols, S, g, ids, dvals, out);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

// numeric assembly of the owned columns [j0, j1) (and the matching rhs entries)
int asm_numeric_range(ghb_ctx* ctx, const double* S, const double* g, const double* ghost, const double* dvals,
                      double* nzval, double* rhs, int64_t j0, int64_t j1, int which) {
  const AsmState& as = ctx->as;
  if (j1 <= j0) return GHB_OK;
  GhostView gv{as.ncells_local, ghost, (int64_t)as.n_b * as.ghost_ncols + as.ghost_ncols};
  if (which & ASM_MATRIX) {
    int64_t blocks = std::min<int64_t>((j1 - j0 + 15) / 16, (int64_t)ctx->sm_count * 16);
    gather_nzval_kernel<16><<<(unsigned)blocks, 256, 0, ctx->stream>>>(j0, j1, as.n_b, as.d_colptr,
                                                                   (const unsigned long long*)as.d_occ, as.d_src, S, gv, nzval);
    GHB_LAUNCHED(ctx);
  }
  if (which & ASM_RHS) {
    gather_rhs_kernel<<<(unsigned)((j1 - j0 + 255) / 256), 256, 0, ctx->stream>>>(
        j0, j1, as.n_b, (const unsigned long long*)as.d_occ, as.d_ids, as.d_celldir, S, g, gv, as.ghost_ncols, dvals, rhs);
    GHB_LAUNCHED(ctx);
  }
  return GHB_OK;
}

int asm_numeric(ghb_ctx* ctx, const double* S, const double* g, const double* ghost, const double* dvals,
                double* nzval, double* rhs) {
  return asm_numeric_range(ctx, S, g, ghost, dvals, nzval, rhs, 0, ctx->as.nrows);
}

// ---- streaming support (ghb_condense_assemble_f64 with host records): when the cells [0, (k+1)*chunk) have been
// condensed, the columns [0, J_k) are complete, J_k = first column with an occurrence in a later cell.  One kernel
// finds, per chunk, the first column that needs it; a suffix minimum on the host turns that into J_k.
__global__ void first_column_of_chunk_kernel(int64_t nrows, int n_b, int64_t chunk, int nchunks,
                                             const unsigned long long* __restrict__ occ,
                                             unsigned long long* __restrict__ first) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nrows) return;
  const unsigned long long o1 = occ[2 * j + 1], o0 = occ[2 * j];
  const unsigned long long o = o1 != ~0ull ? o1 : o0;          // occurrences are stored cell-ascending
  if (o == ~0ull) return;                                      // dof without a cell: an empty column
  const int64_t k = (int64_t)(o / (unsigned long long)n_b) / chunk;
  if (k < nchunks) atomicMin(first + k, (unsigned long long)j);
}

__global__ void pick_colptr_kernel(int n, const int64_t* __restrict__ J, const int64_t* __restrict__ colptr,
                                   int64_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = colptr[J[i]] - 1;
}

int asm_ready_columns(ghb_ctx* ctx, int64_t chunk, int nchunks) {
  AsmState& as = ctx->as;
  if (as.ready_chunk == chunk && (int)as.ready_J.size() == nchunks) return GHB_OK;
  unsigned long long* d_first = nullptr;
  GHB_CUDA(ctx, cudaMallocAsync((void**)&d_first, (size_t)nchunks * 8, ctx->stream));
  GHB_CUDA(ctx, cudaMemsetAsync(d_first, 0xff, (size_t)nchunks * 8, ctx->stream));
  first_column_of_chunk_kernel<<<(unsigned)((as.nrows + 255) / 256), 256, 0, ctx->stream>>>(
      as.nrows, as.n_b, chunk, nchunks, (const unsigned long long*)as.d_occ, d_first);
  GHB_LAUNCHED(ctx);
  std::vector<unsigned long long> first(nchunks);
  GHB_CUDA(ctx, cudaMemcpyAsync(first.data(), d_first, (size_t)nchunks * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  as.ready_J.assign(nchunks, as.nrows);
  unsigned long long m = (unsigned long long)as.nrows;
  for (int k = nchunks - 1; k >= 0; --k) {       // J_k = first column needing a chunk > k
    as.ready_J[k] = (int64_t)m;
    m = std::min(m, first[k]);
  }
  as.ready_J[nchunks - 1] = as.nrows;
  // nzval offsets of the boundaries
  int64_t *d_J = (int64_t*)d_first, *d_p = nullptr;
  GHB_CUDA(ctx, cudaMallocAsync((void**)&d_p, (size_t)nchunks * 8, ctx->stream));
  GHB_CUDA(ctx, cudaMemcpyAsync(d_J, as.ready_J.data(), (size_t)nchunks * 8, cudaMemcpyHostToDevice, ctx->stream));
  pick_colptr_kernel<<<(nchunks + 127) / 128, 128, 0, ctx->stream>>>(nchunks, d_J, as.d_colptr, d_p);
  GHB_LAUNCHED(ctx);
  as.ready_p.assign(nchunks, 0);
  GHB_CUDA(ctx, cudaMemcpyAsync(as.ready_p.data(), d_p, (size_t)nchunks * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFreeAsync(d_first, ctx->stream);
  cudaFreeAsync(d_p, ctx->stream);
  as.ready_chunk = chunk;
  return GHB_OK;
}

}  // namespace ghb
