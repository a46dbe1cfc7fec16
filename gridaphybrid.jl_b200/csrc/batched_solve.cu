// batched_solve.cu -- bulk -> skeleton L2 projection dofs: X = A \ B for a batch of small square systems.
// Replaces compute_bulk_to_skeleton_l2_projection_dofs (/root/reference/src/GridapAPIExtensions.jl:453-500: `A\B` per
// (cell, local facet) with A the facet mass matrix of the skeleton space, n x n, and B the n x m moments of the bulk
// basis -- or a vector -- used by the elasticity / Hencky forms through test/P_m.jl:4-23).  Julia's `\` on a square dense
// matrix is an LU with partial pivoting (dgetrf + dgetrs); here, one warp per system (n <= 32):
//   * factorisation with one row per lane in registers: partial pivoting among the rows not chosen yet (exact
//     first-maximum rule, like idamax), implicit pivoting (rows never move); the factors are written to the warp's
//     shared-memory slot in pivoted order (row of step k -> row k), L below the diagonal, U above, 1/u_kk on it;
//   * right-hand sides with one COLUMN per lane (32 columns per pass): the column is gathered through the pivot order
//     into registers, forward and backward substitution run with static register indices and the factors as uniform
//     (broadcast) shared-memory loads -- n^2 FMAs per column and no shuffles (the first version streamed the columns
//     through the recorded elimination with one shuffle per column and step: 36 M systems/s at (18, 60));
//   * nrhs = 1 (a FE function) keeps the row-per-lane form: Gauss-Jordan on the augmented column.
// info[s] = k+1 if the k-th pivot column is exactly zero (dgetrf semantics), X of that system is NaN.
// HBM-bound in principle: 8 (n^2 + 2 n m) bytes per system.
#include <algorithm>

#include "common.cuh"

namespace ghb {

namespace {

// The elimination steps as compile-time recursions: every register-array index is a constant, whatever the compiler's
// unrolling heuristics decide (with plain `#pragma unroll` loops the NMAX >= 16 instantiations kept a[] in local memory).
template <int K, int NMAX>
struct FactorSteps {
  static __device__ __forceinline__ void run(double (&a)[NMAX], int n, int lane, bool valid, int& ch, int& myq, int& bad) {
    if (K < n) {
      // pivot search: largest |a[K]| among the rows still in play, lowest lane on ties (idamax): lexicographic maximum
      // of the (high, low) words of |a|, two REDUX and a ballot; NaN (after a zero pivot) counts as 0 so that all lanes agree
      const unsigned long long bits = (unsigned long long)__double_as_longlong(a[K]) & 0x7fffffffffffffffull;
      const bool play = valid && ch < 0;
      const bool num = a[K] == a[K];
      const unsigned hi = (play && num) ? (unsigned)(bits >> 32) : 0u, lo = (play && num) ? (unsigned)bits : 0u;
      const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned mlo = __reduce_max_sync(0xffffffffu, (play && hi == mhi) ? lo : 0u);
      const unsigned win = __ballot_sync(0xffffffffu, play && hi == mhi && lo == mlo);
      const int q = __ffs(win) - 1;
      if ((mhi | mlo) == 0u && bad == 0) bad = K + 1;         // exactly singular: first zero pivot column
      const double piv = __shfl_sync(0xffffffffu, a[K], q);
      const double rinv = 1.0 / piv;
      const bool me = lane == q;
      if (me) ch = K;
      if (lane == K) myq = q;
      const double mult = a[K];                                // this row's multiplier of step K
#pragma unroll
      for (int j = K + 1; j < NMAX; ++j) {
        if (j < n) {
          const double pj = __shfl_sync(0xffffffffu, a[j], q) * rinv;   // scaled pivot row
          a[j] = me ? pj : fma(-mult, pj, a[j]);
        }
      }
      a[K] = me ? rinv : mult;
    }
    FactorSteps<K + 1, NMAX>::run(a, n, lane, valid, ch, myq, bad);
  }
};
template <int NMAX>
struct FactorSteps<NMAX, NMAX> {
  static __device__ __forceinline__ void run(double (&)[NMAX], int, int, bool, int&, int&, int&) {}
};

template <int K, int NMAX>
struct VecSteps {   // one right-hand side, row per lane, through the recorded elimination
  static __device__ __forceinline__ void run(const double (&a)[NMAX], int n, int lane, int myq, double& x) {
    if (K < n) {
      const int q = __shfl_sync(0xffffffffu, myq, K);
      const double rinv = __shfl_sync(0xffffffffu, a[K], q);
      const double pj = __shfl_sync(0xffffffffu, x, q) * rinv;
      x = lane == q ? pj : fma(-a[K], pj, x);
    }
    VecSteps<K + 1, NMAX>::run(a, n, lane, myq, x);
  }
};
template <int NMAX>
struct VecSteps<NMAX, NMAX> {
  static __device__ __forceinline__ void run(const double (&)[NMAX], int, int, int, double&) {}
};

template <int K, int NMAX>
struct ColSteps {   // one right-hand-side column per lane: Gauss-Jordan step K with the factors as uniform loads
  static __device__ __forceinline__ void run(double (&x)[NMAX], const double* __restrict__ F, int n) {
    if (K < n) {
      const double pk = x[K] * F[K * (NMAX + 1) + K];
#pragma unroll
      for (int i = 0; i < NMAX; ++i)
        if (i != K && i < n) x[i] = fma(-F[i * (NMAX + 1) + K], pk, x[i]);
      x[K] = pk;
    }
    ColSteps<K + 1, NMAX>::run(x, F, n);
  }
};
template <int NMAX>
struct ColSteps<NMAX, NMAX> {
  static __device__ __forceinline__ void run(double (&)[NMAX], const double*, int) {}
};

template <int NMAX>
__global__ void __launch_bounds__(256) batched_solve_kernel(int64_t nbatch, int n, int m, const double* __restrict__ A,
                                                            const double* __restrict__ B, double* __restrict__ X,
                                                            int32_t* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool valid = lane < n;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  extern __shared__ double fac[];                // per warp: factors [NMAX][NMAX+1] + pivot order
  for (int64_t s = warp; s < nbatch; s += nwarps) {
    const double* As = A + s * (int64_t)n * n;
    double a[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) a[j] = (valid && j < n) ? As[lane + (int64_t)n * j] : 0.0;
    int ch = -1;          // step at which this lane's row was chosen as the pivot row
    int myq = 0;          // lane k remembers the pivot lane of step k
    int bad = 0;
    FactorSteps<0, NMAX>::run(a, n, lane, valid, ch, myq, bad);
    if (info && lane == 0) info[s] = bad;
    const double* Bs = B + s * (int64_t)n * m;
    double* Xs = X + s * (int64_t)n * m;
    if (m == 1) {
      // one right-hand side: row per lane through the recorded elimination (Gauss-Jordan, see the loop above)
      double x = valid ? Bs[lane] : 0.0;
      VecSteps<0, NMAX>::run(a, n, lane, myq, x);
      if (valid) Xs[ch] = bad ? qnan : x;
      continue;
    }
    // ---- factors to shared memory in pivoted order: F[k][j] (row-major, leading dimension NMAX + 1)
    double* F = fac + (threadIdx.x >> 5) * (NMAX * (NMAX + 1) + NMAX);
    int* perm = reinterpret_cast<int*>(F + NMAX * (NMAX + 1));          // perm[k] = source row of pivot k
    __syncwarp();
    if (valid) {
#pragma unroll
      for (int j = 0; j < NMAX; ++j)
        if (j < n) F[ch * (NMAX + 1) + j] = a[j];
      perm[ch] = lane;
    }
    __syncwarp();
    for (int c0 = 0; c0 < m; c0 += 32) {
      const int c = c0 + lane;
      const bool cv = c < m;
      double x[NMAX];
#pragma unroll
      for (int i = 0; i < NMAX; ++i) x[i] = (cv && i < n) ? Bs[perm[i < n ? i : 0] + (int64_t)n * c] : 0.0;
      // the elimination recorded in F is Gauss-Jordan: step k scales row k by 1/u_kk and clears column k in every other
      // row; F[i][k] (i != k) is the multiplier of row i at step k, F[k][k] the reciprocal of the pivot
      ColSteps<0, NMAX>::run(x, F, n);
      if (cv) {
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
          if (i < n) Xs[i + (int64_t)n * c] = bad ? qnan : x[i];
      }
    }
  }
}

}  // namespace

template <int NMAX>
static int launch_bs(ghb_ctx* ctx, int64_t nbatch, int n, int m, const double* A, const double* B, double* X, int32_t* info) {
  constexpr int WARPS = 8;
  const size_t smem = (size_t)WARPS * (NMAX * (NMAX + 1) + NMAX) * sizeof(double);
  auto kern = batched_solve_kernel<NMAX>;
  GHB_SMEM_OPTIN(ctx, kern, smem);
  const int64_t blocks = std::min<int64_t>((nbatch + WARPS - 1) / WARPS, (int64_t)ctx->sm_count * 8);
  kern<<<(unsigned)blocks, 32 * WARPS, smem, ctx->stream>>>(nbatch, n, m, A, B, X, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int launch_batched_solve(ghb_ctx* ctx, int64_t nbatch, int n, int m, const double* A, const double* B, double* X,
                         int32_t* info) {
  if (nbatch <= 0 || m <= 0) return GHB_OK;
  if (n <= 8) return launch_bs<8>(ctx, nbatch, n, m, A, B, X, info);
  if (n <= 16) return launch_bs<16>(ctx, nbatch, n, m, A, B, X, info);
  if (n <= 24) return launch_bs<24>(ctx, nbatch, n, m, A, B, X, info);
  return launch_bs<32>(ctx, nbatch, n, m, A, B, X, info);
}

}  // namespace ghb
