// glue.cu -- index glue around the hot path and the synthetic workload generator.
//   * RestrictFacetDoFsToSkeleton gather       (/root/reference/src/HybridAffineFEOperators.jl:405-434)
//   * full-space free-dof scatter              (/root/reference/src/HybridAffineFEOperators.jl:134-149)
//   * Cartesian cell_wise_facets, closed form of Gridap's first-touch numbering
//   * Philox4x32-10 synthetic records, bit-identical to oracle/oracle.py::synth_cell_records
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace ghb {

namespace {

__global__ void restrict_facet_dofs_kernel(int64_t ncells, int nlf, int nf, const int64_t* __restrict__ cwf,
                                           const int64_t* __restrict__ fdata, int64_t* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per = (int64_t)nlf * nf;
  if (t >= ncells * per) return;
  int64_t cell = t / per;
  int r = (int)(t - cell * per);
  int lf = r / nf, d = r - lf * nf;
  int64_t f = cwf[cell * nlf + lf];
  out[t] = fdata[(f - 1) * nf + d];
}

// x = [field 1 of all cells | field 2 of all cells | ... | lambda_free]; u is [ncells][n_i] with the
// interior fields concatenated per cell.
__global__ void scatter_free_kernel(int64_t ncells, int n_i, int nint, const int32_t* __restrict__ fsize,
                                    const int32_t* __restrict__ foff, const double* __restrict__ u,
                                    const double* __restrict__ lam, int64_t nlam, double* __restrict__ x) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nu = ncells * n_i;
  if (t < nu) {
    int64_t cell = t / n_i;
    int r = (int)(t - cell * n_i);
    int f = 0;
    while (f + 1 < nint && r >= foff[f + 1]) ++f;
    int l = r - foff[f];
    x[(int64_t)foff[f] * ncells + cell * fsize[f] + l] = u[t];
  } else if (t < nu + nlam) {
    x[t] = lam[t - nu];
  }
}

// SumFacetsMap (/root/reference/src/SumFacetsMap.jl:19-30) on the batch: out[c][e] = in[c][0][e] + in[c][1][e] + ...
// summed left to right like the reference's chained BroadcastingFieldOpMap(+).  16-byte vector accesses when len is even.
__global__ void sum_facets_kernel(int64_t ncells, int nlf, int64_t len, const double* __restrict__ in,
                                  double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if ((len & 1) == 0) {
    const int64_t len2 = len >> 1, tot = ncells * len2;
    const double2* in2 = reinterpret_cast<const double2*>(in);
    double2* out2 = reinterpret_cast<double2*>(out);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += stride) {
      const int64_t c = t / len2, e = t - c * len2;
      const double2* p = in2 + c * nlf * len2 + e;
      double2 acc = p[0];
      for (int f = 1; f < nlf; ++f) { const double2 v = p[f * len2]; acc.x += v.x; acc.y += v.y; }
      out2[t] = acc;
    }
  } else {
    const int64_t tot = ncells * len;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += stride) {
      const int64_t c = t / len, e = t - c * len;
      const double* p = in + c * nlf * len + e;
      double acc = p[0];
      for (int f = 1; f < nlf; ++f) acc += p[f * len];
      out[t] = acc;
    }
  }
}

// ---- Cartesian first-touch facet numbering ------------------------------------------------------
struct Mesh {
  int D;
  int64_t dims[3], stride[4];
};

__device__ __forceinline__ int64_t count_low(const Mesh& m, int64_t c, int a) {
  // number of cells c' < c whose index along axis a is 0
  int64_t hi = c / m.stride[a + 1], rem = c - hi * m.stride[a + 1];
  return hi * m.stride[a] + (rem < m.stride[a] ? rem : m.stride[a]);
}

// id (1-based) of the facet (axis, side) of cell c, for side==1 or a low facet on the mesh boundary
__device__ int64_t own_facet_id(const Mesh& m, int64_t c, int axis, int side) {
  int64_t run = (int64_t)m.D * c;
  for (int a = 0; a < m.D; ++a) run += count_low(m, c, a);
  for (int a = m.D - 1; a >= 0; --a) {
    int64_t ia = (c / m.stride[a]) % m.dims[a];
    if (ia == 0) {
      ++run;
      if (a == axis && side == 0) return run;
    }
    ++run;
    if (a == axis && side == 1) return run;
  }
  return -1;
}

__global__ void cartesian_facets_kernel(Mesh m, int64_t cell_start, int64_t ncells, int64_t* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nlf = 2 * m.D;
  if (t >= ncells * nlf) return;
  int64_t c = cell_start + t / nlf;
  int lf = (int)(t % nlf);
  int axis = m.D - 1 - lf / 2, side = lf & 1;
  int64_t ia = (c / m.stride[axis]) % m.dims[axis];
  int64_t id = (side == 1 || ia == 0) ? own_facet_id(m, c, axis, side) : own_facet_id(m, c - m.stride[axis], axis, 1);
  out[t] = id;
}

// ---- Philox4x32-10 --------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ double uniform_pm1(uint64_t cell, uint32_t entry, uint32_t stream, uint64_t seed) {
  uint32_t c0 = entry, c1 = stream, c2 = (uint32_t)cell, c3 = (uint32_t)(cell >> 32);
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint64_t k = ((uint64_t)c0 << 21) | (c1 >> 11);
  return (double)k * 0x1.0p-52 - 1.0;
}

__global__ void synth_fill_kernel(int64_t cell_start, int64_t ncells, int len, uint32_t stream, uint64_t seed,
                                  double* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tot = ncells * len;
  for (; t < tot; t += stride) {
    int64_t cell = t / len;
    uint32_t e = (uint32_t)(t - cell * len);
    out[t] = uniform_pm1((uint64_t)(cell_start + cell), e, stream, seed);
  }
}

__global__ void synth_diag_kernel(PlanDev p, int64_t cell_start, int64_t ncells, uint64_t seed, double d,
                                  double* __restrict__ A) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncells * p.n_i) return;
  int64_t cell = t / p.n_i;
  int r = (int)(t - cell * p.n_i);
  uint64_t gc = (uint64_t)(cell_start + cell);
  uint32_t c0 = 0, c1 = 3, c2 = (uint32_t)gc, c3 = (uint32_t)(gc >> 32);
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  int s = (int)(c0 % (uint32_t)p.n_i);
  int c = (r + s) % p.n_i;
  int off = p.emap[r + (size_t)p.n * c];
  if (off >= 0) A[cell * p.lenA + off] += d;
}

}  // namespace

int launch_restrict_facet_dofs(ghb_ctx* ctx, int64_t ncells, int nlf, int nf, const int64_t* cwf,
                               const int64_t* fdata, int64_t* out) {
  int64_t tot = ncells * nlf * nf;
  restrict_facet_dofs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ncells, nlf, nf, cwf, fdata, out);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int launch_sum_facets(ghb_ctx* ctx, int64_t ncells, int nlf, int64_t len, const double* in, double* out) {
  const int64_t work = ncells * ((len & 1) ? len : len / 2);
  const int64_t blocks = std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 32);
  if (blocks > 0) {
    sum_facets_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ncells, nlf, len, in, out);
    GHB_LAUNCHED(ctx);
  }
  return GHB_OK;
}

// Records of an affine family: out[c][e] = sum_t coef[c][t] * T[t][e].  A thread owns one record element e,
// keeps its ntab table values in registers and walks a slice of cells: per cell ntab FMAs (coefficients are uniform
// loads) and one coalesced 8-byte store -- HBM-write bound (8 len bytes per cell, tables read once per block).
template <int MAXT>
__global__ void __launch_bounds__(256) expand_records_kernel(int64_t ncells, int len, int ntab, const double* __restrict__ T,
                                                             const double* __restrict__ coef, double* __restrict__ out,
                                                             int64_t cells_per_block) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  double t[MAXT];
#pragma unroll
  for (int q = 0; q < MAXT; ++q) t[q] = q < ntab ? T[(size_t)q * len + e] : 0.0;
  const int64_t c0 = (int64_t)blockIdx.y * cells_per_block;
  const int64_t c1 = c0 + cells_per_block < ncells ? c0 + cells_per_block : ncells;
  int64_t c = c0;
  for (; c + 4 <= c1; c += 4) {                    // four cells per step: independent FMA chains, four stores in flight
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < MAXT; ++q) {
      if (q < ntab) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(__ldg(coef + (c + u) * ntab + q), t[q], acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) __stcs(out + (c + u) * (int64_t)len + e, acc[u]);   // written once, read by another kernel
  }
  for (; c < c1; ++c) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < MAXT; ++q)
      if (q < ntab) acc = fma(__ldg(coef + c * ntab + q), t[q], acc);
    __stcs(out + c * (int64_t)len + e, acc);
  }
}

int launch_expand_records(ghb_ctx* ctx, int64_t ncells, int len, int ntab, const double* T, const double* coef,
                          double* out) {
  if (ncells <= 0 || len <= 0) return GHB_OK;
  const unsigned gx = (unsigned)((len + 255) / 256);
  // enough blocks to fill the machine, long enough cell slices to amortise the table loads
  int64_t slices = std::max<int64_t>(1, std::min<int64_t>(ncells, ((int64_t)ctx->sm_count * 16 + gx - 1) / gx));
  slices = std::min<int64_t>(slices, 65535);
  const int64_t per = (ncells + slices - 1) / slices;
  dim3 grid(gx, (unsigned)((ncells + per - 1) / per));
  if (ntab <= 8) expand_records_kernel<8><<<grid, 256, 0, ctx->stream>>>(ncells, len, ntab, T, coef, out, per);
  else expand_records_kernel<16><<<grid, 256, 0, ctx->stream>>>(ncells, len, ntab, T, coef, out, per);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

// In-place transpose of every cell's n x n column-major block (CSR hand-off, see ghb_assemble_numeric_csr_f64): one
// CTA per cell at a time, the block staged in shared memory (n x (n+1), conflict-free both ways), coalesced in and out.
__global__ void transpose_blocks_kernel(int64_t ncells, int n, double* __restrict__ S) {
  extern __shared__ double tile[];
  const int nn = n * n;
  for (int64_t c = blockIdx.x; c < ncells; c += gridDim.x) {
    double* Sc = S + c * (int64_t)nn;
    for (int e = threadIdx.x; e < nn; e += blockDim.x) { const int j = e / n, i = e - j * n; tile[i * (n + 1) + j] = Sc[e]; }
    __syncthreads();
    for (int e = threadIdx.x; e < nn; e += blockDim.x) { const int j = e / n, i = e - j * n; Sc[e] = tile[j * (n + 1) + i]; }
    __syncthreads();
  }
}
// blocks too large for shared memory: swap the pairs directly (uncoalesced on one side)
__global__ void transpose_blocks_swap_kernel(int64_t ncells, int n, double* __restrict__ S) {
  const int64_t nn = (int64_t)n * n;
  for (int64_t c = blockIdx.x; c < ncells; c += gridDim.x) {
    double* Sc = S + c * nn;
    for (int64_t e = threadIdx.x; e < nn; e += blockDim.x) {
      const int j = (int)(e / n), i = (int)(e - (int64_t)j * n);
      if (i < j) { const double a = Sc[e], b = Sc[j + (int64_t)i * n]; Sc[e] = b; Sc[j + (int64_t)i * n] = a; }
    }
  }
}

int launch_transpose_blocks(ghb_ctx* ctx, int64_t ncells, int n, double* S) {
  if (ncells <= 0 || n <= 1) return GHB_OK;
  const size_t smem = (size_t)n * (n + 1) * sizeof(double);
  const int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * 8);
  if (smem <= ctx->smem_optin) {
    GHB_SMEM_OPTIN(ctx, transpose_blocks_kernel, smem);
    transpose_blocks_kernel<<<(unsigned)grid, 256, smem, ctx->stream>>>(ncells, n, S);
  } else {
    transpose_blocks_swap_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(ncells, n, S);
  }
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int launch_scatter_free(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* u, const double* lam,
                        int64_t nlam, double* x) {
  const int nint = (int)p.interior.size();
  std::vector<int32_t> h(2 * (nint + 1), 0);
  int o = 0;
  for (int k = 0; k < nint; ++k) { h[k] = p.ndofs[p.interior[k] - 1]; h[nint + 1 + k] = o; o += h[k]; }
  h[nint + 1 + nint] = o;
  int32_t* d = nullptr;
  GHB_CUDA(ctx, cudaMallocAsync((void**)&d, h.size() * sizeof(int32_t), ctx->stream));
  GHB_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // h is a stack temporary
  int64_t tot = ncells * p.n_i + nlam;
  if (tot > 0) {
    scatter_free_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ncells, p.n_i, nint, d, d + nint + 1, u, lam, nlam, x);
    GHB_LAUNCHED(ctx);
  }
  cudaFreeAsync(d, ctx->stream);
  return GHB_OK;
}

int launch_synth_fill(ghb_ctx* ctx, const Plan& p, int64_t cell_start, int64_t ncells, uint64_t seed, double* A,
                      double* b) {
  const int64_t blocks = (int64_t)ctx->sm_count * 16;
  synth_fill_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(cell_start, ncells, p.lenA, 1u, seed, A);
  GHB_LAUNCHED(ctx);
  synth_fill_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(cell_start, ncells, p.lenb, 2u, seed, b);
  GHB_LAUNCHED(ctx);
  if (p.n_i > 0) {
    const double d = std::sqrt((double)p.n_i) + 1.0;
    int64_t tot = ncells * p.n_i;
    synth_diag_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(p.dev(), cell_start, ncells, seed, d, A);
    GHB_LAUNCHED(ctx);
  }
  return GHB_OK;
}

int launch_cartesian_facets(ghb_ctx* ctx, int D, const int64_t* dims, int64_t cell_start, int64_t ncells,
                            int64_t* out) {
  Mesh m;
  m.D = D;
  m.stride[0] = 1;
  for (int d = 0; d < 3; ++d) {
    m.dims[d] = d < D ? dims[d] : 1;
    m.stride[d + 1] = m.stride[d] * m.dims[d];
  }
  int64_t tot = ncells * 2 * D;
  cartesian_facets_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(m, cell_start, ncells, out);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

}  // namespace ghb
