// condense_generic.cu -- generic (any block plan) static condensation and backward static condensation.
//
// One CTA per cell.  The cell's augmented system  W = [A11 A12 b1; A21 A22 b2]  (condensed order,
// column-major, n x (n+1)) is gathered from the packed record into shared memory (global scratch when
// it does not fit) and the first n_i columns are eliminated by Gaussian elimination with partial
// pivoting restricted to the interior rows -- the same pivot rule as dgetrf
// (/root/reference/src/StaticCondensationMap.jl:179, first max |.| like idamax).  What is left in the
// lower-right block is S = A22 - A21 A11^-1 A12 and g = b2 - A21 A11^-1 b1 (:183-192).
// Every thread owns whole columns, so the row swap of a step needs no barrier; warp 0 owns the pivot
// column.  This is the portable fallback for arbitrary shapes; the tuned kernels live in
// condense_dmma.cu.
#include <algorithm>

#include "common.cuh"

namespace ghb {

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ void warp_argmax(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

// Forward elimination of columns [0, n_i) of W (n rows, ncols columns, leading dim ld).
// Pivot rows restricted to [k, n_i).  Returns LAPACK info (0 ok, k+1 if the k-th pivot is exactly 0).
__device__ int eliminate(double* W, int ld, int n, int n_i, int ncols, int* s_piv, int* s_info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 0; k < n_i; ++k) {
    if (warp == 0) {
      double* col = W + (size_t)k * ld;
      double best = -1.0;
      int bi = n_i;
      for (int i = k + lane; i < n_i; i += 32) {
        double v = fabs(col[i]);
        if (v > best) { best = v; bi = i; }  // ascending i per lane: strict '>' keeps the first max
      }
      warp_argmax(best, bi);
      double pv = col[bi];
      __syncwarp();
      if (pv == 0.0) {
        if (lane == 0) *s_info = k + 1;
      } else {
        if (lane == 0) {
          *s_piv = bi;
          double t = col[k]; col[k] = pv; col[bi] = t;
        }
        __syncwarp();
        double r = 1.0 / pv;  // dgetf2 scales by the reciprocal
        for (int i = k + 1 + lane; i < n; i += 32) col[i] *= r;
      }
    }
    __syncthreads();
    if (*s_info) return *s_info;
    const int piv = *s_piv;
    const double* l = W + (size_t)k * ld;
    for (int j = k + 1 + tid; j < ncols; j += blockDim.x) {
      double* col = W + (size_t)j * ld;
      double u = col[piv];
      col[piv] = col[k];
      col[k] = u;
      for (int i = k + 1; i < n; ++i) col[i] = fma(-l[i], u, col[i]);
    }
    __syncthreads();
  }
  return 0;
}

// Back substitution with the upper triangle left in W[0:n_i, 0:n_i] on columns [c0, c1): thread per column.
__device__ void back_substitute(double* W, int ld, int n_i, int c0, int c1) {
  for (int j = c0 + threadIdx.x; j < c1; j += blockDim.x) {
    double* col = W + (size_t)j * ld;
    for (int k = n_i - 1; k >= 0; --k) {
      const double* uk = W + (size_t)k * ld;
      double x = col[k] / uk[k];
      col[k] = x;
      for (int i = 0; i < k; ++i) col[i] = fma(-uk[i], x, col[i]);
    }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(kThreads) condense_generic_kernel(PlanDev p, int64_t ncells,
                                                                    const double* __restrict__ A,
                                                                    const double* __restrict__ b,
                                                                    double* __restrict__ S, double* __restrict__ g,
                                                                    int32_t* __restrict__ info, double* __restrict__ X,
                                                                    double* __restrict__ scratch, int ld) {
  extern __shared__ double sm[];
  __shared__ int s_piv, s_info;
  const int n = p.n, n_i = p.n_i, n_b = p.n_b;
  double* W = SMEM ? sm : scratch + (size_t)blockIdx.x * ld * (n + 1);
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const double* Arec = A + cell * p.lenA;
    const double* brec = b + cell * p.lenb;
    for (int idx = threadIdx.x; idx < n * (n + 1); idx += blockDim.x) {
      int j = idx / n, i = idx - j * n;
      int off = p.emap[idx];
      W[i + (size_t)j * ld] = j < n ? (off >= 0 ? Arec[off] : 0.0) : brec[off];
    }
    if (threadIdx.x == 0) { s_info = 0; s_piv = 0; }
    __syncthreads();
    int inf = eliminate(W, ld, n, n_i, n + 1, &s_piv, &s_info);
    double* Sc = S + cell * (int64_t)n_b * n_b;
    double* gc = g + cell * (int64_t)n_b;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int idx = threadIdx.x; idx < n_b * (n_b + 1); idx += blockDim.x) {
      int j = idx / n_b, i = idx - j * n_b;
      double v = inf ? qnan : W[(n_i + i) + (size_t)(n_i + j) * ld];
      if (j < n_b) Sc[idx] = v; else gc[i] = v;
    }
    if (info && threadIdx.x == 0) info[cell] = inf;
    if (X) {  // keep_factors: X = A11^-1 [A12 | b1]
      if (!inf) back_substitute(W, ld, n_i, n_i, n + 1);
      __syncthreads();
      double* Xc = X + cell * (int64_t)n_i * (n_b + 1);
      for (int idx = threadIdx.x; idx < n_i * (n_b + 1); idx += blockDim.x) {
        int j = idx / n_i, i = idx - j * n_i;
        Xc[idx] = inf ? qnan : W[i + (size_t)(n_i + j) * ld];
      }
    }
    __syncthreads();
  }
}

// Backward static condensation (/root/reference/src/BackwardStaticCondensationMap.jl:84-99):
// r = b1 - A12*lambda_K ; LU = P*A11 (recomputed, as the reference does) ; u = A11^-1 r.
template <bool SMEM>
__global__ void __launch_bounds__(kThreads) backsub_generic_kernel(PlanDev p, int64_t ncells,
                                                                   const double* __restrict__ A,
                                                                   const double* __restrict__ b,
                                                                   const double* __restrict__ lam_free,
                                                                   const double* __restrict__ lam_dir,
                                                                   const int64_t* __restrict__ ids,
                                                                   double* __restrict__ u, int32_t* __restrict__ info,
                                                                   double* __restrict__ scratch, int ld) {
  extern __shared__ double sm[];
  __shared__ int s_piv, s_info;
  const int n = p.n, n_i = p.n_i, n_b = p.n_b;
  // layout: W = [A11 | r] (n_i x (n_i+1)), then lambda_K (n_b)
  double* W = SMEM ? sm : scratch + (size_t)blockIdx.x * ((size_t)ld * (n_i + 1) + n_b);
  double* lam = W + (size_t)ld * (n_i + 1);
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const double* Arec = A + cell * p.lenA;
    const double* brec = b + cell * p.lenb;
    for (int l = threadIdx.x; l < n_b; l += blockDim.x) {
      int64_t id = ids[cell * n_b + l];
      lam[l] = id > 0 ? lam_free[id - 1] : (id < 0 && lam_dir ? lam_dir[-id - 1] : 0.0);
    }
    for (int idx = threadIdx.x; idx < n_i * n_i; idx += blockDim.x) {
      int j = idx / n_i, i = idx - j * n_i;
      int off = p.emap[i + (size_t)n * j];
      W[i + (size_t)j * ld] = off >= 0 ? Arec[off] : 0.0;
    }
    if (threadIdx.x == 0) { s_info = 0; s_piv = 0; }
    __syncthreads();
    // gemv!('N',-1,A12,x,1,b1): column-oriented axpy, ascending columns
    for (int i = threadIdx.x; i < n_i; i += blockDim.x) {
      double r = brec[p.emap[i + (size_t)n * n]];
      for (int j = 0; j < n_b; ++j) {
        int off = p.emap[i + (size_t)n * (n_i + j)];
        if (off >= 0) r = fma(-Arec[off], lam[j], r);
      }
      W[i + (size_t)n_i * ld] = r;
    }
    __syncthreads();
    int inf = eliminate(W, ld, n_i, n_i, n_i + 1, &s_piv, &s_info);
    if (!inf) back_substitute(W, ld, n_i, n_i, n_i + 1);
    __syncthreads();
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = threadIdx.x; i < n_i; i += blockDim.x) u[cell * (int64_t)n_i + i] = inf ? qnan : W[i + (size_t)n_i * ld];
    if (info && threadIdx.x == 0) info[cell] = inf;
    __syncthreads();
  }
}

// u = y - X*lambda_K with stored factors X = A11^-1 [A12 | b1]: one thread per (cell, interior row),
// consecutive threads on consecutive rows, so every load of a column of X is a coalesced run of n_i doubles and no lane
// idles on the ragged rows beyond a multiple of 32 (one warp per cell measured 178 M cells/s on (34,36), this 1.5-3x).
// ids and lambda of a cell are re-read by its n_i threads from L1.  HBM-bound: 8 n_i (n_b + 2) + 16 n_b bytes per cell.
__global__ void __launch_bounds__(256) backsub_factors_kernel(int n_i, int n_b, int64_t ncells,
                                                              const double* __restrict__ X,
                                                              const double* __restrict__ lam_free,
                                                              const double* __restrict__ lam_dir,
                                                              const int64_t* __restrict__ ids,
                                                              double* __restrict__ u) {
  const int64_t total = ncells * n_i;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = t / n_i;
    const int i = (int)(t - cell * n_i);
    const double* Xc = X + cell * (int64_t)n_i * (n_b + 1) + i;
    const int64_t* idc = ids + cell * n_b;
    double r = Xc[(size_t)n_b * n_i];
    int j = 0;
    for (; j + 4 <= n_b; j += 4) {                 // four independent id -> lambda chains in flight
      int64_t id[4];
      double x[4], l[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { id[q] = __ldg(idc + j + q); x[q] = Xc[(size_t)(j + q) * n_i]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        l[q] = id[q] > 0 ? __ldg(lam_free + id[q] - 1) : (id[q] < 0 && lam_dir ? __ldg(lam_dir - id[q] - 1) : 0.0);
#pragma unroll
      for (int q = 0; q < 4; ++q) r = fma(-x[q], l[q], r);   // ascending columns, as gemv
    }
    for (; j < n_b; ++j) {
      const int64_t id = __ldg(idc + j);
      const double lj = id > 0 ? __ldg(lam_free + id - 1) : (id < 0 && lam_dir ? __ldg(lam_dir - id - 1) : 0.0);
      r = fma(-Xc[(size_t)j * n_i], lj, r);
    }
    u[t] = r;
  }
}

}  // namespace

static int pick_ld(int n) { return n | 1; }  // odd leading dimension: column owners hit distinct banks

int launch_condense_generic(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                            double* S, double* g, int32_t* info, double* X) {
  const int ld = pick_ld(p.n);
  const size_t wbytes = (size_t)ld * (p.n + 1) * sizeof(double);
  const bool smem = wbytes + 1024 <= ctx->smem_optin;
  int per_sm = smem ? (int)std::max<size_t>(1, std::min<size_t>(16, (ctx->smem_optin + 1024) / (wbytes + 1024))) : 8;
  int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * per_sm);
  double* scratch = nullptr;
  if (smem) {
    GHB_SMEM_OPTIN(ctx, condense_generic_kernel<true>, wbytes);
    condense_generic_kernel<true><<<(unsigned)grid, kThreads, wbytes, ctx->stream>>>(p.dev(), ncells, A, b, S, g, info, X, nullptr, ld);
  } else {
    GHB_CUDA(ctx, cudaMallocAsync((void**)&scratch, wbytes * grid, ctx->stream));
    condense_generic_kernel<false><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(p.dev(), ncells, A, b, S, g, info, X, scratch, ld);
  }
  GHB_LAUNCHED(ctx);
  if (scratch) cudaFreeAsync(scratch, ctx->stream);
  return GHB_OK;
}

int launch_backsub_generic(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                           const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                           int32_t* info) {
  const int ld = pick_ld(p.n_i);
  const size_t wbytes = ((size_t)ld * (p.n_i + 1) + p.n_b) * sizeof(double);
  const bool smem = wbytes + 1024 <= ctx->smem_optin;
  int per_sm = smem ? (int)std::max<size_t>(1, std::min<size_t>(16, (ctx->smem_optin + 1024) / (wbytes + 1024))) : 8;
  int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * per_sm);
  double* scratch = nullptr;
  if (smem) {
    GHB_SMEM_OPTIN(ctx, backsub_generic_kernel<true>, wbytes);
    backsub_generic_kernel<true><<<(unsigned)grid, kThreads, wbytes, ctx->stream>>>(p.dev(), ncells, A, b, lam_free, lam_dir, ids, u, info, nullptr, ld);
  } else {
    GHB_CUDA(ctx, cudaMallocAsync((void**)&scratch, wbytes * grid, ctx->stream));
    backsub_generic_kernel<false><<<(unsigned)grid, kThreads, 0, ctx->stream>>>(p.dev(), ncells, A, b, lam_free, lam_dir, ids, u, info, scratch, ld);
  }
  GHB_LAUNCHED(ctx);
  if (scratch) cudaFreeAsync(scratch, ctx->stream);
  return GHB_OK;
}

int launch_backsub_factors(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* X, const double* lam_free,
                           const double* lam_dir, const int64_t* ids, double* u) {
  int64_t blocks = std::min<int64_t>((ncells * p.n_i + 255) / 256, (int64_t)ctx->sm_count * 8);
  if (blocks < 1) return GHB_OK;
  backsub_factors_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p.n_i, p.n_b, ncells, X, lam_free, lam_dir, ids, u);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

}  // namespace ghb
