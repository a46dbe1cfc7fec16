// batched_solve.cu -- bulk -> skeleton L2 projection dofs: X = A \ B for a batch of small square systems.
// Replaces compute_bulk_to_skeleton_l2_projection_dofs (/root/reference/src/GridapAPIExtensions.jl:453-500: `A\B` per
// (cell, local facet) with A the facet mass matrix of the skeleton space, n x n, and B the n x m moments of the bulk
// basis -- or a vector -- used by the elasticity / Hencky forms through test/P_m.jl:4-23).  Julia's `\` on a square dense
// matrix is an LU with partial pivoting (dgetrf + dgetrs); here:
//   * one warp per system, one row per lane (n <= 32), the row of A in registers;
//   * Gauss-Jordan elimination with partial pivoting among the rows not chosen yet (exact first-maximum rule, like
//     idamax), implicit pivoting: rows never move, the lane chosen at step k ends up holding solution component k;
//     a lane keeps its multipliers in place of its row, the pivot lane the reciprocal of the pivot;
//   * the right-hand sides are streamed in chunks of 8 columns through the recorded elimination (one 64-bit shuffle per
//     column and step), so m is unbounded and A is read once.
// info[s] = k+1 if the k-th pivot column is exactly zero (dgetrf semantics), X of that system is NaN.
// HBM-bound in principle (8 (n^2 + 2 n m) bytes per system); the elimination is shuffle-bound like the other warp kernels.
#include <algorithm>

#include "common.cuh"

namespace ghb {

namespace {

template <int NMAX>
__global__ void __launch_bounds__(256) batched_solve_kernel(int64_t nbatch, int n, int m, const double* __restrict__ A,
                                                            const double* __restrict__ B, double* __restrict__ X,
                                                            int32_t* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool valid = lane < n;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int64_t s = warp; s < nbatch; s += nwarps) {
    const double* As = A + s * (int64_t)n * n;
    double a[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) a[j] = (valid && j < n) ? As[lane + (int64_t)n * j] : 0.0;
    int ch = -1;          // step at which this lane's row was chosen as the pivot row
    int myq = 0;          // lane k remembers the pivot lane of step k
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      if (k < n) {
        // pivot search: largest |a[k]| among the rows still in play, lowest lane on ties (idamax)
        const double av = fabs(a[k]);
        double v = (valid && ch < 0) ? (av == av ? av : 0.0) : -1.0;   // NaN (after a zero pivot) counts as 0: all lanes must agree
        int q = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, v, off);
          const int oq = __shfl_xor_sync(0xffffffffu, q, off);
          if (ov > v || (ov == v && oq < q)) { v = ov; q = oq; }
        }
        if (v == 0.0 && bad == 0) bad = k + 1;               // exactly singular: first zero pivot column
        const double piv = __shfl_sync(0xffffffffu, a[k], q);
        const double rinv = 1.0 / piv;
        const bool me = lane == q;
        if (me) ch = k;
        if (lane == k) myq = q;
        const double mult = a[k];                             // this row's multiplier of step k
#pragma unroll
        for (int j = k + 1; j < NMAX; ++j) {
          if (j < n) {
            const double pj = __shfl_sync(0xffffffffu, a[j], q) * rinv;   // scaled pivot row
            a[j] = me ? pj : fma(-mult, pj, a[j]);
          }
        }
        a[k] = me ? rinv : mult;
      }
    }
    if (info && lane == 0) info[s] = bad;
    // ---- right-hand sides, 8 columns at a time, through the recorded elimination
    const double* Bs = B + s * (int64_t)n * m;
    double* Xs = X + s * (int64_t)n * m;
    for (int c0 = 0; c0 < m; c0 += 8) {
      double x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (valid && c0 + j < m) ? Bs[lane + (int64_t)n * (c0 + j)] : 0.0;
#pragma unroll
      for (int k = 0; k < NMAX; ++k) {
        if (k < n) {
          const int q = __shfl_sync(0xffffffffu, myq, k);
          const double rinv = __shfl_sync(0xffffffffu, a[k], q);
          const bool me = lane == q;
          const double mult = a[k];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double pj = __shfl_sync(0xffffffffu, x[j], q) * rinv;
            x[j] = me ? pj : fma(-mult, pj, x[j]);
          }
        }
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < m) Xs[ch + (int64_t)n * (c0 + j)] = bad ? qnan : x[j];
      }
    }
  }
}

}  // namespace

int launch_batched_solve(ghb_ctx* ctx, int64_t nbatch, int n, int m, const double* A, const double* B, double* X,
                         int32_t* info) {
  if (nbatch <= 0 || m <= 0) return GHB_OK;
  const int64_t blocks = std::min<int64_t>((nbatch + 7) / 8, (int64_t)ctx->sm_count * 8);
  if (n <= 8) batched_solve_kernel<8><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nbatch, n, m, A, B, X, info);
  else if (n <= 16) batched_solve_kernel<16><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nbatch, n, m, A, B, X, info);
  else if (n <= 24) batched_solve_kernel<24><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nbatch, n, m, A, B, X, info);
  else batched_solve_kernel<32><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nbatch, n, m, A, B, X, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

}  // namespace ghb
