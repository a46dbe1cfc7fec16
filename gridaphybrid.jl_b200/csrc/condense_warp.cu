// condense_warp.cu -- static condensation and backward static condensation for SMALL cells (n = n_i + n_b <= 32:
// Darcy HDG k=1 on quads (7,8), RT-H k=1 on quads (16,8)): one warp per cell, one row of the augmented system per
// lane, the whole cell in registers, no shared memory.  These configurations are HBM-bound (AI 0.85 - 1.8 flop/B), so
// what matters is coalesced record loads and enough cells in flight; the elimination itself is the FMA-warp-tile
// path of BASELINE.json's north_star ("plain FMA warp tiles otherwise").
//
// Replaces evaluate!(cache, ::StaticCondensationMap, A, b) (/root/reference/src/StaticCondensationMap.jl:152-196) and
// evaluate!(cache, ::BackwardStaticCondensationMap, A, b, x) (src/BackwardStaticCondensationMap.jl:61-102).
// Pivoting: partial pivoting over the interior rows with implicit row exchange (a lane keeps its row); the pivot is
// the row of largest magnitude to 2^-15 relative (one REDUX.MAX on a packed key), exact-zero columns give LAPACK's info.
#include <algorithm>

#include "common.cuh"

// The warp kernels are compiled for 8 resident blocks per SM for n <= 16 and 6 above (register caps 64 / 80): measured
// best of 4 / 6 / 8 on a B200.  Staging the next record in shared memory with cp.async (double buffered per warp) was
// measured too and did not pay (817 / 340 vs 932 / 364 M cells/s): these kernels are bound by the elimination chain,
// not by the exposed load latency.

namespace ghb {

namespace {

__device__ __forceinline__ double fast_rcp_w(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// pivot search among `cand` lanes on value v: returns the lane (lowest among the maxima at key resolution) and
// sets zero=true when every candidate is below 2^-1017 (reported as an exact zero pivot)
__device__ __forceinline__ int pivot_lane(double v, bool cand, int lane, bool& zero) {
  const unsigned h = (unsigned)(__double_as_longlong(v) >> 32) & 0x7fffffffu;
  const unsigned key = cand ? (((h >> 5) << 5) | (unsigned)(31 - lane)) : 0u;
  const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
  zero = (kmax >> 5) == 0u;
  return 31 - (int)(kmax & 31u);
}

// One warp per cell.  Lane r < N holds row r of W = [A11 A12 b1; A21 A22 b2] (condensed order).
template <int NI, int NB>
__global__ void __launch_bounds__(128, (NI + NB <= 16 ? 8 : 6)) condense_warp_kernel(PlanDev p, int64_t ncells, const double* __restrict__ A,
                                                            const double* __restrict__ b, double* __restrict__ S,
                                                            double* __restrict__ g, int32_t* __restrict__ info) {
  constexpr int N = NI + NB;
  static_assert(N <= 32, "one row per lane");
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool rowok = lane < N;
  const int32_t* em = p.emap + (rowok ? lane : 0);   // record offsets of this lane's row (L1-resident table)
  for (int64_t cell = warp; cell < ncells; cell += nwarps) {
    const double* Arec = A + cell * p.lenA;
    const double* brec = b + cell * p.lenb;
    double a[N + 1];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int o = rowok ? em[N * j] : -1;
      a[j] = o >= 0 ? Arec[o] : 0.0;
    }
    a[N] = rowok ? brec[em[N * N]] : 0.0;
    bool chosen = false;
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const bool cand = lane < NI && !chosen;
      const double rc = fast_rcp_w(a[k]);
      bool zero;
      const int q = pivot_lane(a[k], cand, lane, zero);
      bad = (bad == 0 && zero) ? k + 1 : bad;
      const double rinv = __shfl_sync(0xffffffffu, rc, q);
      const bool me = lane == q;
      const bool upd = rowok && !chosen && !me;       // interior rows still in play and every boundary row
      chosen = chosen || me;
      const double nl = upd ? -(a[k] * rinv) : 0.0;
#pragma unroll
      for (int j = k + 1; j <= N; ++j) {
        const double pj = __shfl_sync(0xffffffffu, a[j], q);
        a[j] = fma(nl, pj, a[j]);
      }
    }
    // boundary rows now hold S (columns NI..N-1) and g (column N)
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (lane >= NI && lane < N) {
      double* Sc = S + cell * (int64_t)NB * NB + (lane - NI);
#pragma unroll
      for (int j = 0; j < NB; ++j) Sc[(int64_t)NB * j] = bad ? qnan : a[NI + j];
      g[cell * (int64_t)NB + (lane - NI)] = bad ? qnan : a[N];
    }
    if (info && lane == 0) info[cell] = bad;
  }
}

// Two cells per warp: each half-warp holds one cell, R rows per lane (n <= 16 R).  The elimination is bound by the
// shuffle pipe (one 64-bit broadcast per remaining column and pivot: 0.7 shuffles per cycle per SM in the one-cell
// kernels); with two cells per instruction every broadcast serves both, so the shuffle count per cell halves.  The
// pivot search is a butterfly inside the half-warp (xor 8, 4, 2, 1 never leave it).
template <int NI, int NB, int R>
__global__ void __launch_bounds__(128, (R == 1 ? 8 : 3))
condense_warp2_kernel(PlanDev p, int64_t ncells, const double* __restrict__ A, const double* __restrict__ b,
                      double* __restrict__ S, double* __restrict__ g, int32_t* __restrict__ info) {
  constexpr int N = NI + NB;
  static_assert(N <= 16 * R && R <= 2, "R rows per lane of a half-warp");
  const int lane = threadIdx.x & 31, hl = lane & 15, hbase = lane & 16;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pair = warp; 2 * pair < ncells; pair += nwarps) {
    const int64_t cell = 2 * pair + (lane >> 4);
    const bool cellok = cell < ncells;                 // odd cell count: the upper half idles on the last pair
    const double* Arec = A + (cellok ? cell : 0) * p.lenA;
    const double* brec = b + (cellok ? cell : 0) * p.lenb;
    double a[R][N + 1];
    bool live[R], chosen[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = hl + 16 * r;
      live[r] = row < N && cellok;
      chosen[r] = false;
      const int32_t* em = p.emap + (row < N ? row : 0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int o = live[r] ? em[N * j] : -1;
        a[r][j] = o >= 0 ? Arec[o] : 0.0;
      }
      a[r][N] = live[r] ? brec[em[N * N]] : 0.0;
    }
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      // key = |a| (exponent + 15 mantissa bits) << 5 | (31 - row): largest magnitude, lowest row, per half-warp
      unsigned key = 0u;
      double rc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = hl + 16 * r;
        const bool cand = row < NI && !chosen[r];
        const unsigned h = (unsigned)(__double_as_longlong(a[r][k]) >> 32) & 0x7fffffffu;
        const unsigned kr = cand ? (((h >> 5) << 5) | (unsigned)(31 - row)) : 0u;
        key = kr > key ? kr : key;
        rc[r] = fast_rcp_w(a[r][k]);                   // speculative: overlaps the search
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const unsigned other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      const bool zero = (key >> 5) == 0u;              // every candidate below 2^-1017: reported as a zero pivot
      bad = (bad == 0 && zero) ? k + 1 : bad;
      const int prow = 31 - (int)(key & 31u);
      const bool from2 = R > 1 && prow >= 16;          // which register set of the pivot lane (uniform per half)
      const int q = hbase | (prow & 15);
      const double rinv = __shfl_sync(0xffffffffu, from2 ? rc[R - 1] : rc[0], q);
      double nl[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool me = lane == q && (r == 1) == from2;
        const bool upd = live[r] && !chosen[r] && !me;
        chosen[r] = chosen[r] || me;
        nl[r] = upd ? -(a[r][k] * rinv) : 0.0;
      }
#pragma unroll
      for (int j = k + 1; j <= N; ++j) {
        const double pj = __shfl_sync(0xffffffffu, from2 ? a[R - 1][j] : a[0][j], q);
#pragma unroll
        for (int r = 0; r < R; ++r) a[r][j] = fma(nl[r], pj, a[r][j]);
      }
    }
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = hl + 16 * r;
      if (cellok && row >= NI && row < N) {
        double* Sc = S + cell * (int64_t)NB * NB + (row - NI);
#pragma unroll
        for (int j = 0; j < NB; ++j) Sc[(int64_t)NB * j] = bad ? qnan : a[r][NI + j];
        g[cell * (int64_t)NB + (row - NI)] = bad ? qnan : a[r][N];
      }
    }
    if (info && cellok && hl == 0) info[cell] = bad;
  }
}

// Backward map: lane r < NI holds row r of [A11 | b1 - A12*lambda_K]; Gauss-Jordan with partial pivoting, so the
// lane whose row was chosen at step k ends up holding u[k].
template <int NI, int NB>
__global__ void __launch_bounds__(128, (NI + NB <= 16 ? 8 : 6)) backsub_warp_kernel(PlanDev p, int64_t ncells, const double* __restrict__ A,
                                                           const double* __restrict__ b,
                                                           const double* __restrict__ lam_free,
                                                           const double* __restrict__ lam_dir,
                                                           const int64_t* __restrict__ ids, double* __restrict__ u,
                                                           int32_t* __restrict__ info) {
  constexpr int N = NI + NB;
  static_assert(NI <= 32, "one interior row per lane");
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool rowok = lane < NI;
  const int32_t* em = p.emap + (rowok ? lane : 0);
  for (int64_t cell = warp; cell < ncells; cell += nwarps) {
    const double* Arec = A + cell * p.lenA;
    const double* brec = b + cell * p.lenb;
    double a[NI + 1];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int o = rowok ? em[N * j] : -1;
      a[j] = o >= 0 ? Arec[o] : 0.0;
    }
    double r = rowok ? brec[em[N * N]] : 0.0;
    // r -= A12 * lambda_K  (gemv!('N',-1,A12,x,1,b1), ascending columns)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int64_t id = ids[cell * NB + j];        // uniform across the warp
      const double lj = id > 0 ? lam_free[id - 1] : (id < 0 && lam_dir ? lam_dir[-id - 1] : 0.0);
      const int o = rowok ? em[N * (NI + j)] : -1;
      if (o >= 0) r = fma(-Arec[o], lj, r);
    }
    a[NI] = r;
    bool chosen = false;
    int step = -1, bad = 0;
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const bool cand = rowok && !chosen;
      const double rc = fast_rcp_w(a[k]);
      bool zero;
      const int q = pivot_lane(a[k], cand, lane, zero);
      bad = (bad == 0 && zero) ? k + 1 : bad;
      const double rinv = __shfl_sync(0xffffffffu, rc, q);
      const bool me = lane == q;
      if (me) { chosen = true; step = k; }
      // Gauss-Jordan: normalise the pivot row, eliminate column k from every other row
      const double scale = me ? rinv : 1.0;
      const double nl = (rowok && !me) ? -a[k] : 0.0;
#pragma unroll
      for (int j = k + 1; j <= NI; ++j) {
        const double pj = __shfl_sync(0xffffffffu, a[j], q) * rinv;
        a[j] = me ? a[j] * scale : fma(nl, pj, a[j]);
      }
    }
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (rowok && step >= 0) u[cell * (int64_t)NI + step] = bad ? qnan : a[NI];
    if (info && lane == 0) info[cell] = bad;
  }
}

// Backward map, two cells per warp (n_i <= 16): see condense_warp2_kernel.
template <int NI, int NB>
__global__ void __launch_bounds__(128, (NI <= 8 ? 8 : 5)) backsub_warp2_kernel(PlanDev p, int64_t ncells, const double* __restrict__ A,
                                                                const double* __restrict__ b,
                                                                const double* __restrict__ lam_free,
                                                                const double* __restrict__ lam_dir,
                                                                const int64_t* __restrict__ ids, double* __restrict__ u,
                                                                int32_t* __restrict__ info) {
  constexpr int N = NI + NB;
  static_assert(NI <= 16, "one interior row per lane of a half-warp");
  const int lane = threadIdx.x & 31, hl = lane & 15, hbase = lane & 16;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool rowok = hl < NI;
  const int32_t* em = p.emap + (rowok ? hl : 0);
  for (int64_t pair = warp; 2 * pair < ncells; pair += nwarps) {
    const int64_t cell = 2 * pair + (lane >> 4);
    const bool cellok = cell < ncells;
    const bool live = rowok && cellok;
    const int64_t cc = cellok ? cell : 0;
    const double* Arec = A + cc * p.lenA;
    const double* brec = b + cc * p.lenb;
    double a[NI + 1];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int o = live ? em[N * j] : -1;
      a[j] = o >= 0 ? Arec[o] : 0.0;
    }
    double r = live ? brec[em[N * N]] : 0.0;
    // r -= A12 * lambda_K  (gemv!('N',-1,A12,x,1,b1), ascending columns); ids are uniform inside a half-warp
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int64_t id = ids[cc * NB + j];
      const double lj = id > 0 ? lam_free[id - 1] : (id < 0 && lam_dir ? lam_dir[-id - 1] : 0.0);
      const int o = live ? em[N * (NI + j)] : -1;
      if (o >= 0) r = fma(-Arec[o], lj, r);
    }
    a[NI] = r;
    bool chosen = false;
    int step = -1, bad = 0;
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const bool cand = rowok && !chosen;
      const double rc = fast_rcp_w(a[k]);
      const unsigned h = (unsigned)(__double_as_longlong(a[k]) >> 32) & 0x7fffffffu;
      unsigned key = cand ? (((h >> 4) << 4) | (unsigned)(15 - hl)) : 0u;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const unsigned other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      const bool zero = (key >> 4) == 0u;
      bad = (bad == 0 && zero) ? k + 1 : bad;
      const int q = hbase | (15 - (int)(key & 15u));
      const double rinv = __shfl_sync(0xffffffffu, rc, q);
      const bool me = lane == q;
      if (me) { chosen = true; step = k; }
      // Gauss-Jordan: normalise the pivot row, eliminate column k from every other row
      const double scale = me ? rinv : 1.0;
      const double nl = (rowok && !me) ? -a[k] : 0.0;
#pragma unroll
      for (int j = k + 1; j <= NI; ++j) {
        const double pj = __shfl_sync(0xffffffffu, a[j], q) * rinv;
        a[j] = me ? a[j] * scale : fma(nl, pj, a[j]);
      }
    }
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (live && step >= 0) u[cell * (int64_t)NI + step] = bad ? qnan : a[NI];
    if (info && cellok && hl == 0) info[cell] = bad;
  }
}

template <int NI, int NB>
int launch_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S, double* g,
              int32_t* info) {
  const int64_t blocks = std::min<int64_t>((ncells + 3) / 4, (int64_t)ctx->sm_count * 32);
  condense_warp_kernel<NI, NB><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p.dev(), ncells, A, b, S, g, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

template <int NI, int NB>
int launch_bw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, const double* lf,
              const double* ld, const int64_t* ids, double* u, int32_t* info) {
  const int64_t blocks = std::min<int64_t>((ncells + 3) / 4, (int64_t)ctx->sm_count * 32);
  backsub_warp_kernel<NI, NB><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p.dev(), ncells, A, b, lf, ld, ids, u, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

}  // namespace

// shapes with a register-resident warp kernel (BASELINE.json C1 and C2 k=1)
const char* warp_kernel_name(const Plan& p) {
  if (p.n_i == 7 && p.n_b == 8) return "warp_7_8";
  if (p.n_i == 16 && p.n_b == 8) return "warp_16_8";
  return nullptr;
}

int launch_condense_warp(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                         double* g, int32_t* info) {
  // (7,8): two cells per warp, 932 -> 1537 M cells/s.  (16,8) with two rows per lane (R = 2, 168 registers, 3 blocks
  // per SM) measured 302 vs 364 M cells/s for the one-cell kernel, so it stays on the latter (GHB_WARP_TWO_ROWS=1 selects
  // the R = 2 variant for experiments).
  if (p.n_i == 7 && p.n_b == 8 && !p.opt.warp_one_cell) {
    const int64_t blocks = std::min<int64_t>((ncells + 7) / 8, (int64_t)ctx->sm_count * 32);
    condense_warp2_kernel<7, 8, 1><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p.dev(), ncells, A, b, S, g, info);
    GHB_LAUNCHED(ctx);
    return GHB_OK;
  }
  if (p.n_i == 16 && p.n_b == 8 && p.opt.warp_two_rows) {
    const int64_t blocks = std::min<int64_t>((ncells + 7) / 8, (int64_t)ctx->sm_count * 32);
    condense_warp2_kernel<16, 8, 2><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p.dev(), ncells, A, b, S, g, info);
    GHB_LAUNCHED(ctx);
    return GHB_OK;
  }
  if (p.n_i == 7 && p.n_b == 8) return launch_cw<7, 8>(ctx, p, ncells, A, b, S, g, info);
  if (p.n_i == 16 && p.n_b == 8) return launch_cw<16, 8>(ctx, p, ncells, A, b, S, g, info);
  return fail(ctx, GHB_EUNSUPPORTED, "no warp kernel for this shape");
}

int launch_backsub_warp(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                        const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info) {
  if (!p.opt.warp_one_cell) {
    const int64_t blocks = std::min<int64_t>((ncells + 7) / 8, (int64_t)ctx->sm_count * 32);
    if (p.n_i == 7 && p.n_b == 8)
      backsub_warp2_kernel<7, 8><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p.dev(), ncells, A, b, lam_free, lam_dir, ids, u, info);
    else
      backsub_warp2_kernel<16, 8><<<(unsigned)blocks, 128, 0, ctx->stream>>>(p.dev(), ncells, A, b, lam_free, lam_dir, ids, u, info);
    GHB_LAUNCHED(ctx);
    return GHB_OK;
  }
  if (p.n_i == 7 && p.n_b == 8) return launch_bw<7, 8>(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  if (p.n_i == 16 && p.n_b == 8) return launch_bw<16, 8>(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  return fail(ctx, GHB_EUNSUPPORTED, "no warp kernel for this shape");
}

}  // namespace ghb
