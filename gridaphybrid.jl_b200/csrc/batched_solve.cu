// batched_solve.cu -- bulk -> skeleton L2 projection dofs: X = A \ B for a batch of small square systems.
// Replaces compute_bulk_to_skeleton_l2_projection_dofs (/root/reference/src/GridapAPIExtensions.jl:453-500: `A\B` per
// (cell, local facet) with A the facet mass matrix of the skeleton space, n x n, and B the n x m moments of the bulk
// basis -- or a vector -- used by the elasticity / Hencky forms through test/P_m.jl:4-23).  Julia's `\` on a square dense
// matrix is an LU with partial pivoting (dgetrf + dgetrs); here, one warp per system for n <= 32 (one CTA per system in
// shared memory for 32 < n <= 128, see batched_solve_big_kernel):
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

// ---- 32 < n <= 128: one CTA (4 warps) per system, LU with partial pivoting in shared memory ------------------------
// (facet spaces of vector-valued unknowns at high order, e.g. 3 x 15 = 45 dofs per facet for k = 4 in 3-D).  The matrix
// lives in shared memory with leading dimension n + 1; per step: warp 0 finds the pivot (first maximum of |a| in the
// current row order, like idamax), the rows are swapped physically (dlaswp) together with the row-index vector, every
// thread owns one row of the trailing update.  The right-hand sides go through in chunks of 32 columns held in shared
// memory: dgetrs's forward and backward substitution, one (row group, column) per thread.
__global__ void __launch_bounds__(128) batched_solve_big_kernel(int64_t nbatch, int n, int m, const double* __restrict__ A,
                                                                const double* __restrict__ B, double* __restrict__ X,
                                                                int32_t* __restrict__ info) {
  extern __shared__ double sm[];
  const int ld = n + 1;
  double* F = sm;                         // [n][ld] row-major: L below the diagonal (unit), U on and above
  double* Xc = F + (size_t)n * ld;        // [n][33]  chunk of right-hand sides
  int* perm = reinterpret_cast<int*>(Xc + (size_t)n * 33);   // [n] source row of every position
  __shared__ int s_piv, s_bad;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int64_t s = blockIdx.x; s < nbatch; s += gridDim.x) {
    const double* As = A + s * (int64_t)n * n;
    for (int e = tid; e < n * n; e += 128) { const int j = e / n, i = e - j * n; F[i * ld + j] = As[e]; }   // column-major in
    for (int i = tid; i < n; i += 128) perm[i] = i;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
      if (wid == 0) {
        // idamax over positions k..n-1: largest |a|, lowest position on ties; NaN counts as 0 (after a zero pivot)
        unsigned long long best = 0ull; int bi = n;
        for (int i = k + lane; i < n; i += 32) {
          const double v = F[i * ld + k];
          const unsigned long long bits = v == v ? ((unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull) : 0ull;
          if (bits > best || (bits == best && i < bi)) { best = bits; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
          s_piv = bi;
          if (best == 0ull && s_bad == 0) s_bad = k + 1;      // dgetrf's info: first exactly zero pivot
        }
      }
      __syncthreads();
      const int p = s_piv;
      if (p != k) {                                           // dlaswp: whole rows trade places
        for (int j = tid; j < n; j += 128) { const double t0 = F[k * ld + j]; F[k * ld + j] = F[p * ld + j]; F[p * ld + j] = t0; }
        if (tid == 0) { const int t0 = perm[k]; perm[k] = perm[p]; perm[p] = t0; }
        __syncthreads();
      }
      const double rinv = 1.0 / F[k * ld + k];                // dgetf2 scales the column by the reciprocal of the pivot
      for (int i = k + 1 + tid; i < n; i += 128) {
        const double l = F[i * ld + k] * rinv;
        F[i * ld + k] = l;
        for (int j = k + 1; j < n; ++j) F[i * ld + j] = fma(-l, F[k * ld + j], F[i * ld + j]);
      }
      __syncthreads();
    }
    const int bad = s_bad;
    if (info && tid == 0) info[s] = bad;
    const double* Bs = B + s * (int64_t)n * m;
    double* Xs = X + s * (int64_t)n * m;
    for (int c0 = 0; c0 < m; c0 += 32) {
      const int c = c0 + lane, nc = min(32, m - c0);
      for (int i = wid; i < n; i += 4) Xc[i * 33 + lane] = lane < nc ? Bs[perm[i] + (int64_t)n * c] : 0.0;   // P B
      __syncthreads();
      for (int k = 0; k < n; ++k) {                           // L y = P B (unit lower)
        const double xk = Xc[k * 33 + lane];
        for (int i = k + 1 + wid; i < n; i += 4) Xc[i * 33 + lane] = fma(-F[i * ld + k], xk, Xc[i * 33 + lane]);
        __syncthreads();
      }
      for (int k = n - 1; k >= 0; --k) {                      // U x = y
        if (wid == 0) Xc[k * 33 + lane] = Xc[k * 33 + lane] / F[k * ld + k];
        __syncthreads();
        const double xk = Xc[k * 33 + lane];
        for (int i = wid; i < k; i += 4) Xc[i * 33 + lane] = fma(-F[i * ld + k], xk, Xc[i * 33 + lane]);
        __syncthreads();
      }
      for (int i = wid; i < n; i += 4)
        if (lane < nc) Xs[i + (int64_t)n * c] = bad ? qnan : Xc[i * 33 + lane];
      __syncthreads();
    }
  }
}

}  // namespace

static int launch_bs_big(ghb_ctx* ctx, int64_t nbatch, int n, int m, const double* A, const double* B, double* X, int32_t* info) {
  const size_t smem = ((size_t)n * (n + 1) + (size_t)n * 33) * sizeof(double) + (size_t)n * sizeof(int);
  auto kern = batched_solve_big_kernel;
  GHB_SMEM_OPTIN(ctx, kern, smem);
  const int per_sm = std::max(1, (int)(200000 / (smem + 1024)));
  const int64_t blocks = std::min<int64_t>(nbatch, (int64_t)ctx->sm_count * std::min(per_sm, 8));
  kern<<<(unsigned)blocks, 128, smem, ctx->stream>>>(nbatch, n, m, A, B, X, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

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
  if (n <= 32) return launch_bs<32>(ctx, nbatch, n, m, A, B, X, info);
  return launch_bs_big(ctx, nbatch, n, m, A, B, X, info);
}

}  // namespace ghb
