// condense_large.cu -- static condensation and backward static condensation for LARGE cells (64 < n_i <= 128, any
// boundary size): BASELINE.json configs 4 and 5, elasticity HDG k=2 on hexes (n_i, n_b) = (120, 108) and Hencky HDG
// k=1 (106, 72).  Such a record (n^2 + n doubles: 418 kB / 255 kB) does not fit in shared memory, so only A11 lives
// there and the rest is streamed:
//
//   1. A11 -> shared memory (leading dimension = 4 mod 8, in-tile column permutation pc(): conflict-free fragments);
//   2. blocked right-looking LU with partial pivoting, panel width 8: warp 0 factorises the panel (one row per lane,
//      up to four register sets, implicit pivoting, one REDUX.MAX per pivot), then all 8 warps update the column tiles
//      to the right with FP64 DMMA (unit-lower solve as inv(L_pp) * G, trailing update C -= L * U);
//      inv(L_pp), inv(U_pp) and the row moves of every panel are kept (L is stored LINPACK-style: the
//      multipliers of a panel stay in the row order of its own time);
//   3. [A12 | b1] is streamed in chunks of 8 column tiles; every warp owns one
//      column tile and runs, without any block-wide barrier, the forward solve (replaying each panel's row moves),
//      the backward solve (both blocked, the
//      diagonal blocks through their inverses) and the Schur update S(:, tile) = A22(:, tile) - A21 * X(:, tile) with
//      A21 / A22 read straight from the record as DMMA fragments (64-byte runs of the packed columns).
//
// Replaces evaluate!(cache, ::StaticCondensationMap, A, b) (/root/reference/src/StaticCondensationMap.jl:152-196:
// getrf!, getrs!, gemm!, getrs!, gemv!) and evaluate!(cache, ::BackwardStaticCondensationMap, A, b, x)
// (src/BackwardStaticCondensationMap.jl:61-102: gemv!, getrf!, getrs!) for these shapes; X = A11^-1 [A12 | b1] is a
// by-product, so keep_factors costs only its store.  Pivot rule and info[] semantics as in condense_dmma.cu.
// FP64-bound shapes (AI 12-14 flop/B): roofline = 37.1 TFLOP/s / F_cond.  16 warps per CTA (128 registers): the column-tile
// updates of the factorisation and the Schur phase use all of them (8 -> 16 warps: 1.06 -> 1.14 and 1.29 -> 1.47 M cells/s).
#include <algorithm>

#include "common.cuh"

namespace ghb {

namespace {

// optional phase timing (-DGHB_LTRACE): thread 0 of CTA 0 prints clock64() deltas of its first cell
#ifdef GHB_LTRACE
__device__ long long g_ltrace[256];
__device__ const char* g_lname[256];
#define LTRACE(name) do { if (blockIdx.x == 0 && threadIdx.x == 0 && cell == 0 && t_n < 256) { g_ltrace[t_n] = clock64(); g_lname[t_n] = name; ++t_n; } } while (0)
#else
#define LTRACE(name) do { } while (0)
#endif

#ifndef GHB_LARGE_THREADS
#define GHB_LARGE_THREADS 512
#endif
constexpr int kLT = GHB_LARGE_THREADS;       // threads per CTA, one CTA per SM
constexpr int kWarps = kLT / 32;
constexpr int kCT = 8;         // column tiles of [A12 | b1] per chunk: one per warp of the solve phase (the other warps join for the Schur phase)
// (the panel warp holds up to 4 rows per lane: n_i <= 128)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// in-tile column permutation of the shared-memory images (see condense_dmma.cu)
__device__ __host__ __forceinline__ constexpr int pc(int c) { return c ^ ((c >> 2) & 1); }

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

__device__ __forceinline__ void cp_async8_z(void* smem_dst, const void* gsrc, bool ok) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "r"(ok ? 8 : 0));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// one-instruction L2 prefetch of a contiguous global range (16-byte granules)
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, size_t bytes) {
  const unsigned long long a0 = ((unsigned long long)gptr + 15ull) & ~15ull;
  const unsigned long long a1 = ((unsigned long long)gptr + bytes) & ~15ull;
  if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}

struct LCtl {
  int ndisp;
  int psrc[8];     // position (before this panel's permutation) of the k-th pivot row
  int dsrc[8];     // displaced rows: old position inside the diagonal block ...
  int ddst[8];     // ... and the vacated position they move to
  int vac[8];
  int pad[3];
  double rinv[8];
};

struct LargeTables {
  const int32_t* colbase;  // [(n+1)*nf]: record offset of (first row of field f, column c) or -1; c == n: offset in b
  const uint8_t* rowf;     // [n] field slot of condensed row r
  const uint8_t* rowl;     // [n] row inside its field
  int nf;
};

// ---- panel factorisation by one warp: rows [c0, ni) of the 8 columns [c0, c0+8) -------------------
template <int kMaxSets>
__device__ __forceinline__ void panel_factor_l(double* __restrict__ Wa, const int ld, const int ni, const int c0,
                                               const int npiv, LCtl* ctl, int* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const int nrows = ni - c0;
  double a[kMaxSets][8];
  int ch[kMaxSets];
  bool v[kMaxSets];
  double* base = Wa + c0 + lane + ld * c0;
#pragma unroll
  for (int s = 0; s < kMaxSets; ++s) {
    v[s] = lane + 32 * s < nrows;
    ch[s] = -1;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[s][j] = v[s] ? base[32 * s + ld * pc(j)] : 0.0;
  }
  double myrinv = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < npiv) {
      // pivot search: key = |a| (exponent + 14 mantissa bits) << 7 | (127 - row): largest magnitude, lowest row
      unsigned key = 0u;
#pragma unroll
      for (int s = 0; s < kMaxSets; ++s) {
        const bool cand = v[s] && ch[s] < 0;
        const unsigned h = (unsigned)(__double_as_longlong(a[s][k]) >> 32) & 0x7fffffffu;
        const unsigned ks = cand ? (((h >> 6) << 7) | (unsigned)(127 - (32 * s + lane))) : 0u;
        key = ks > key ? ks : key;
      }
      unsigned kmax = __reduce_max_sync(0xffffffffu, key);
      if ((kmax >> 7) == 0u) {
        // every candidate below 2^-1016: decide exactly (zero column => LAPACK info = k+1), go on with the first row
        unsigned any = 0u;
        int row = -1;
#pragma unroll
        for (int s = 0; s < kMaxSets; ++s) {
          const bool cand = v[s] && ch[s] < 0;
          const unsigned long long e = cand ? ((unsigned long long)__double_as_longlong(a[s][k]) & 0x7fffffffffffffffull) : 0ull;
          any |= (unsigned)(e >> 32) | (unsigned)e;
          const unsigned bal = __ballot_sync(0xffffffffu, cand);
          if (row < 0 && bal) row = 32 * s + __ffs(bal) - 1;
        }
        if (__reduce_max_sync(0xffffffffu, any) == 0u) {
          if (lane == 0 && *info == 0) *info = c0 + k + 1;
        }
        kmax = (unsigned)(127 - row);
      }
      const int prow = 127 - (int)(kmax & 127u);
      const int sp = prow >> 5, q = prow & 31;
      double pk = a[0][k];
#pragma unroll
      for (int s = 1; s < kMaxSets; ++s) pk = sp == s ? a[s][k] : pk;
      const double rinv = fast_rcp(__shfl_sync(0xffffffffu, pk, q));
      if (lane == k) myrinv = rinv;
      double nl[kMaxSets];
#pragma unroll
      for (int s = 0; s < kMaxSets; ++s) {
        const bool me = sp == s && lane == q;
        const bool upd = v[s] && ch[s] < 0 && !me;
        if (me) ch[s] = k;
        const double l = a[s][k] * rinv;               // dgetf2: scale by the reciprocal
        a[s][k] = upd ? l : a[s][k];
        nl[s] = upd ? -l : 0.0;                        // rows out of play: a + 0*p = a
      }
#pragma unroll
      for (int j = k + 1; j < 8; ++j) {
        double pj = a[0][j];
#pragma unroll
        for (int s = 1; s < kMaxSets; ++s) pj = sp == s ? a[s][j] : pj;
        pj = __shfl_sync(0xffffffffu, pj, q);
#pragma unroll
        for (int s = 0; s < kMaxSets; ++s) a[s][j] = fma(nl[s], pj, a[s][j]);
      }
    }
  }
  // new positions: pivot rows to the top of the block, rows displaced from it into the vacated slots
  const bool disp = v[0] && lane < npiv && ch[0] < 0;
  const unsigned mdisp = __ballot_sync(0xffffffffu, disp);
  const unsigned ltm = (1u << lane) - 1u;
  int nvac = 0;
#pragma unroll
  for (int s = 0; s < kMaxSets; ++s) {
    const bool vc = v[s] && (32 * s + lane >= npiv) && ch[s] >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, vc);
    if (vc) ctl->vac[nvac + __popc(m & ltm)] = c0 + 32 * s + lane;
    nvac += __popc(m);
  }
  __syncwarp();
  int np[kMaxSets];
#pragma unroll
  for (int s = 0; s < kMaxSets; ++s) {
    np[s] = c0 + 32 * s + lane;
    if (ch[s] >= 0) np[s] = c0 + ch[s];
    else if (s == 0 && disp) np[s] = ctl->vac[__popc(mdisp & ltm)];
    if (ch[s] >= 0) ctl->psrc[ch[s]] = c0 + 32 * s + lane;
  }
  if (disp) {
    const int t = __popc(mdisp & ltm);
    ctl->dsrc[t] = c0 + lane;
    ctl->ddst[t] = np[0];
  }
  if (lane == 0) ctl->ndisp = __popc(mdisp);
  __syncwarp();
  double* wb = Wa + ld * c0;
#pragma unroll
  for (int s = 0; s < kMaxSets; ++s) {
    if (v[s]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) wb[np[s] + ld * pc(j)] = a[s][j];
    }
  }
  if (lane < 8) ctl->rinv[lane] = myrinv;
}

// inverses of the two triangles of a factorised 8x8 diagonal block (lanes 0-7 own one column each; rows/columns
// >= npiv are identity for L and zero for U^-1); storage (row + 8*col)
__device__ __forceinline__ void invert_unit_lower_l(const double* __restrict__ D, const int ld, const int npiv,
                                                    double* __restrict__ Linv) {
  const int n = threadIdx.x & 31;
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (i == n) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m < i; ++m) {
      const double lim = i < npiv ? D[i + ld * pc(m)] : 0.0;
      if (m & 1) s1 = fma(-lim, x[m], s1); else s0 = fma(-lim, x[m], s0);
    }
    x[i] = (i > n) ? (s0 + s1) : x[i];
  }
  if (n < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) Linv[i + 8 * n] = x[i];
  }
}

__device__ __forceinline__ void invert_upper_l(const double* __restrict__ D, const int ld, const int npiv,
                                               const double* __restrict__ rinv, double* __restrict__ Dinv) {
  const int n = threadIdx.x & 31;
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (i == n && n < npiv) ? rinv[i] : 0.0;
#pragma unroll
  for (int i = 6; i >= 0; --i) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = i + 1; m < 8; ++m) {
      const double uim = m < npiv ? D[i + ld * pc(m)] : 0.0;
      if (m & 1) s1 = fma(uim, x[m], s1); else s0 = fma(uim, x[m], s0);
    }
    x[i] = (i < n && n < npiv) ? -(s0 + s1) * rinv[i] : x[i];
  }
  if (n < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) Dinv[i + 8 * n] = x[i];
  }
}

// One warp, one 8-column tile `T` (leading dimension ld, rows < ni): rows of block p <- inv(L_pp) * (pivot rows),
// displaced rows moved, then rows of the blocks I > p: T_I -= L_Ip * T_p.  Used for the column tiles of A11 during
// the factorisation and, replaying the stored row moves of every panel, for the forward solve of a streamed chunk.
__device__ __forceinline__ void tile_lower_step(double* __restrict__ T, const double* __restrict__ Wa, const int ld,
                                                const int ni, const int p, const int NT, const LCtl* __restrict__ ctl,
                                                const double* __restrict__ Linv) {
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int ce0 = pc(2 * tig), ce1 = pc(2 * tig + 1), ka0 = tig, ka1 = pc(4 + tig), nb = pc(gid);
  const int c0 = 8 * p;
  const int npiv = (ni - c0) < 8 ? (ni - c0) : 8;
  double* colg = T + ld * nb;
  double* colt = T + ld * tig;
  int ps0 = tig < npiv ? c0 + tig : -1, ps1 = 4 + tig < npiv ? c0 + 4 + tig : -1, dsr = -1, dds = -1;
  if (ctl) {
    ps0 = tig < npiv ? ctl->psrc[tig] : -1;
    ps1 = 4 + tig < npiv ? ctl->psrc[4 + tig] : -1;
    const int nd = ctl->ndisp;
    dsr = gid < nd ? ctl->dsrc[gid] : -1;
    dds = gid < nd ? ctl->ddst[gid] : -1;
  }
  const double li0 = Linv[gid + 8 * tig], li1 = Linv[gid + 8 * (4 + tig)];
  const double g0 = ps0 >= 0 ? colg[ps0] : 0.0;
  const double g1 = ps1 >= 0 ? colg[ps1] : 0.0;
  const double dv0 = dsr >= 0 ? colt[dsr] : 0.0;
  const double dv1 = dsr >= 0 ? colt[dsr + 4 * ld] : 0.0;
  __syncwarp();
  double u0 = 0.0, u1 = 0.0;
  dmma(u0, u1, li0, g0);
  dmma(u0, u1, li1, g1);
  double* cc0 = T + gid + ld * ce0;
  double* cc1 = T + gid + ld * ce1;
  if (c0 + gid < ni) { cc0[c0] = u0; cc1[c0] = u1; }
  if (dds >= 0) { colt[dds] = dv0; colt[dds + 4 * ld] = dv1; }
  __syncwarp();
  if (p + 1 < NT) {
    const double bf0 = -colg[c0 + tig], bf1 = -colg[c0 + 4 + tig];
    // four row tiles per step: independent accumulators, k-step outermost (a warp issues in order)
#pragma unroll 1
    for (int I = p + 1; I < NT; I += 4) {
      double a0[4], a1[4], d0[4], d1[4];
      bool vv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = 8 * (I + q) + gid;
        vv[q] = I + q < NT && r < ni;
        a0[q] = vv[q] ? Wa[r + ld * (c0 + ka0)] : 0.0;
        a1[q] = vv[q] ? Wa[r + ld * (c0 + ka1)] : 0.0;
        d0[q] = vv[q] ? cc0[8 * (I + q)] : 0.0;
        d1[q] = vv[q] ? cc1[8 * (I + q)] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) dmma(d0[q], d1[q], a0[q], bf0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dmma(d0[q], d1[q], a1[q], bf1);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (vv[q]) { cc0[8 * (I + q)] = d0[q]; cc1[8 * (I + q)] = d1[q]; }
    }
  }
  __syncwarp();
}

// One warp, one 8-column tile: block p <- inv(U_pp) * T_p, then the blocks I < p: T_I -= U_Ip * T_p (backward solve).
__device__ __forceinline__ void tile_upper_step(double* __restrict__ T, const double* __restrict__ Wa, const int ld,
                                                const int ni, const int p, const double* __restrict__ Dinv) {
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int ce0 = pc(2 * tig), ce1 = pc(2 * tig + 1), ka0 = tig, ka1 = pc(4 + tig), nb = pc(gid);
  const int c0 = 8 * p;
  double* colg = T + ld * nb;
  const double di0 = Dinv[gid + 8 * tig], di1 = Dinv[gid + 8 * (4 + tig)];
  const double g0 = c0 + tig < ni ? colg[c0 + tig] : 0.0;
  const double g1 = c0 + 4 + tig < ni ? colg[c0 + 4 + tig] : 0.0;
  __syncwarp();
  double x0 = 0.0, x1 = 0.0;
  dmma(x0, x1, di0, g0);
  dmma(x0, x1, di1, g1);
  double* cc0 = T + gid + ld * ce0;
  double* cc1 = T + gid + ld * ce1;
  if (c0 + gid < ni) { cc0[c0] = x0; cc1[c0] = x1; }
  __syncwarp();
  if (p > 0) {
    const double bf0 = c0 + tig < ni ? -colg[c0 + tig] : 0.0;
    const double bf1 = c0 + 4 + tig < ni ? -colg[c0 + 4 + tig] : 0.0;
#pragma unroll 1
    for (int I = 0; I < p; I += 4) {
      double a0[4], a1[4], d0[4], d1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = 8 * (I + q) + gid;
        const bool vq = I + q < p;
        a0[q] = vq ? Wa[r + ld * (c0 + ka0)] : 0.0;
        a1[q] = vq ? Wa[r + ld * (c0 + ka1)] : 0.0;
        d0[q] = vq ? cc0[8 * (I + q)] : 0.0;
        d1[q] = vq ? cc1[8 * (I + q)] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) dmma(d0[q], d1[q], a0[q], bf0);
#pragma unroll
      for (int q = 0; q < 4; ++q) dmma(d0[q], d1[q], a1[q], bf1);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (I + q < p) { cc0[8 * (I + q)] = d0[q]; cc1[8 * (I + q)] = d1[q]; }
    }
  }
  __syncwarp();
}

// MODE 0: static condensation (S, g; X = A11^-1 [A12 | b1] if requested).  MODE 1: backward map (u).
template <int MODE>
__global__ void __launch_bounds__(kLT, 1)
condense_large_kernel(PlanDev pl, LargeTables tb, int ld, int64_t ncells, const double* __restrict__ A,
                      const double* __restrict__ b, double* __restrict__ S, double* __restrict__ g,
                      int32_t* __restrict__ info, double* __restrict__ Xout, const double* __restrict__ lam_free,
                      const double* __restrict__ lam_dir, const int64_t* __restrict__ ids, double* __restrict__ uout) {
  const int ni = pl.n_i, nbd = pl.n_b, n = pl.n, nf = tb.nf;
  const int NT = (ni + 7) / 8;                       // row / column tiles of A11 = panels
  extern __shared__ __align__(16) double smem[];
  double* Wa = smem;                                 // [NT*8][ld]   A11 -> L\U
  double* Xc = Wa + (size_t)NT * 8 * ld;             // [kCT*8][ld]  chunk of [A12 | b1] -> X
  double* LinvAll = Xc + (size_t)kCT * 8 * ld;       // [NT][64]
  double* DinvAll = LinvAll + (size_t)NT * 64;       // [NT][64]
  double* s_lam = DinvAll + (size_t)NT * 64;         // [n_b] (backward map)
  LCtl* ctlAll = reinterpret_cast<LCtl*>(s_lam + ((nbd + 1) & ~1));   // [NT] row moves + pivot reciprocals per panel
  int* s_info = reinterpret_cast<int*>(ctlAll + NT);
  int* s_colbase = s_info + 4;                       // [(n+1)*nf]
  unsigned short* s_rowinfo = reinterpret_cast<unsigned short*>(s_colbase + (n + 1) * nf);   // [n] (field << 8 | local)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int ce0 = pc(2 * tig), ce1 = pc(2 * tig + 1), nb = pc(gid);

  for (int i = tid; i < (n + 1) * nf; i += kLT) s_colbase[i] = tb.colbase[i];
  for (int i = tid; i < n; i += kLT) s_rowinfo[i] = (unsigned short)((tb.rowf[i] << 8) | tb.rowl[i]);
  for (int i = tid; i < NT * 8 * ld; i += kLT) Wa[i] = 0.0;
  for (int i = tid; i < kCT * 8 * ld; i += kLT) Xc[i] = 0.0;
  __syncthreads();
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int ngrp = kLT / ni > 0 ? kLT / ni : 1;        // column groups of the A11 loader (n_i <= 128 < kLT)
  const int ni4 = ni;

  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const double* Arec = A + cell * pl.lenA;
    const double* brec = b + cell * pl.lenb;
#ifdef GHB_LTRACE
    int t_n = 0;
    LTRACE("start");
#endif
    // ------------------------------------------------------------------ A11 -> shared memory
    // the whole record is pulled into L2 now (one instruction per 64 kB piece): the chunk loads and the Schur phase,
    // tens of microseconds later, then find it there instead of paying DRAM latency on dependent loads
    if (tid == 32) {
      const size_t bytes = (size_t)pl.lenA * 8;
      for (size_t o = 0; o < bytes; o += 65536) l2_prefetch_bulk((const char*)Arec + o, bytes - o < 65536 ? bytes - o : 65536);
    }
    {
      // thread = (row, column group): asynchronous 8-byte copies, all in flight at once
      const int r = tid % ni4, grp = tid / ni4;       // ni4 = rows rounded so that kLT/ni4 groups exist
      if (r < ni && grp < ngrp) {
        const int ri = s_rowinfo[r];
        const double* src = Arec + (ri & 0xff);
        const int* cb = s_colbase + (ri >> 8);
        double* dst = Wa + r;
#pragma unroll 4
        for (int c = grp; c < ni; c += ngrp) {
          const int off = cb[c * nf];
          cp_async8_z(dst + ld * pc(c), src + (off >= 0 ? off : 0), off >= 0);
        }
      }
      cp_async_commit_wait_all();
    }
    if (tid == 0) *s_info = 0;
    if (MODE == 1) {
      for (int j = tid; j < nbd; j += kLT) {
        const int64_t id = ids[cell * nbd + j];
        s_lam[j] = id > 0 ? lam_free[id - 1] : (id < 0 && lam_dir ? lam_dir[-id - 1] : 0.0);
      }
    }
    __syncthreads();
    LTRACE("load A11");
    // ------------------------------------------------------------------ LU of A11
    // Bulk-synchronous: warp 0 factorises panel p and inverts its diagonal block, then all warps apply it to the column
    // tiles to the right.  (A one-panel look-ahead -- warp 0 updating tile p+1 and factorising panel p+1 while warps
    // 1-7 update the other tiles -- measured slower, 0.68 vs 0.71 M cells/s on (120,108): the panel chain slows down
    // under the shared-memory traffic of the update warps by more than the hidden update time.)
    // (A panel spread over one warp per 32 rows, keys and pivot row exchanged through shared memory with two named
    // barriers per pivot, also measured slower: 6.7-7.0 k cycles per panel against 6.6 / 5.2 / 4.2 / 3.8 k for the
    // single-warp panel on 4 / 3 / 2 / 1 register sets -- a barrier round trip costs more than the extra sets.)
    auto factor_panel = [&](int p) {
      const int c0 = 8 * p;
      const int npiv = (ni - c0) < 8 ? (ni - c0) : 8;
      LCtl* ctl = ctlAll + p;
      const int nsets = (ni - c0 + 31) >> 5;              // rows still in play / 32
      if (nsets >= 4) panel_factor_l<4>(Wa, ld, ni, c0, npiv, ctl, s_info);
      else if (nsets == 3) panel_factor_l<3>(Wa, ld, ni, c0, npiv, ctl, s_info);
      else if (nsets == 2) panel_factor_l<2>(Wa, ld, ni, c0, npiv, ctl, s_info);
      else panel_factor_l<1>(Wa, ld, ni, c0, npiv, ctl, s_info);
      __syncwarp();
      invert_unit_lower_l(Wa + c0 + ld * c0, ld, npiv, LinvAll + 64 * p);
      invert_upper_l(Wa + c0 + ld * c0, ld, npiv, ctl->rinv, DinvAll + 64 * p);
    };
#pragma unroll 1
    for (int p = 0; p < NT; ++p) {
      if (warp == 0) factor_panel(p);
      __syncthreads();
      LTRACE("  panel + inverses");
      for (int J = p + 1 + warp; J < NT; J += kWarps)
        tile_lower_step(Wa + ld * 8 * J, Wa, ld, ni, p, NT, ctlAll + p, LinvAll + 64 * p);
      __syncthreads();
      LTRACE("  column tiles");
    }
    const bool failed = *s_info != 0;
    // ------------------------------------------------------------------ [A12 | b1] in chunks of kCT column tiles
    const int ncolsR = MODE == 0 ? nbd + 1 : 1;                  // right-hand side columns
    const int ntilesR = (ncolsR + 7) / 8;
#pragma unroll 1
    for (int t0 = 0; t0 < ntilesR; t0 += kCT) {
      const int nt = ntilesR - t0 < kCT ? ntilesR - t0 : kCT;    // tiles of this chunk
      // Xc[row][j] = record(interior row, column n_i + 8*t0 + j)
      if (MODE == 0) {
        const int r = tid % ni4, grp = tid / ni4;
        if (r < ni && grp < ngrp) {
          const int ri = s_rowinfo[r];
          const int* cb = s_colbase + (ri >> 8);
          double* dst = Xc + r;
#pragma unroll 4
          for (int j = grp; j < nt * 8; j += ngrp) {
            const int c = 8 * t0 + j;                            // column of [A12 | b1]
            const int off = c <= nbd ? cb[(ni + c) * nf] : -1;
            const double* src = (c < nbd ? Arec : brec) + (ri & 0xff) + (off >= 0 ? off : 0);
            cp_async8_z(dst + ld * (8 * (j >> 3) + pc(j & 7)), src, off >= 0);
          }
        }
        cp_async_commit_wait_all();
      } else {
        // r = b1 - A12 * lambda_K (gemv!('N', -1, A12, x, 1, b1), ascending columns); the other columns are zero
        for (int idx = tid; idx < 8 * ld; idx += kLT) Xc[idx] = 0.0;
        __syncthreads();
        for (int pos = tid; pos < ni; pos += kLT) {
          const int ri = s_rowinfo[pos];
          const int f = ri >> 8, lr = ri & 0xff;
          double r = brec[s_colbase[n * nf + f] + lr];
          for (int j = 0; j < nbd; ++j) {
            const int off = s_colbase[(ni + j) * nf + f];
            if (off >= 0) r = fma(-Arec[off + lr], s_lam[j], r);
          }
          Xc[pos + ld * pc(0)] = r;
        }
      }
      __syncthreads();
      LTRACE("chunk load");
      if (warp < nt) {
        double* T = Xc + ld * 8 * warp;
        // forward solve L y = P c (the row moves of every panel replayed), then backward solve U x = y
#pragma unroll 1
        for (int p = 0; p < NT; ++p) tile_lower_step(T, Wa, ld, ni, p, NT, ctlAll + p, LinvAll + 64 * p);
        LTRACE("  forward solve (warp 0)");
#pragma unroll 1
        for (int p = NT - 1; p >= 0; --p) tile_upper_step(T, Wa, ld, ni, p, DinvAll + 64 * p);
        LTRACE("  backward solve (warp 0)");
        const int cbase = 8 * (t0 + warp);                       // first column of [A12 | b1] in this tile
        if (MODE == 0) {
          if (Xout) {      // keep_factors: X = A11^-1 [A12 | b1], n_i x (n_b + 1) column-major
            double* Xg = Xout + cell * (int64_t)ni * (nbd + 1);
            for (int idx = lane; idx < 8 * ni; idx += 32) {
              const int j = idx / ni, i = idx - j * ni;
              if (cbase + j <= nbd) Xg[i + (int64_t)ni * (cbase + j)] = failed ? qnan : T[i + ld * pc(j)];
            }
          }
        } else if (warp == 0) {
          double* uc = uout + cell * (int64_t)ni;
          for (int i = lane; i < ni; i += 32) uc[i] = failed ? qnan : T[i + ld * pc(0)];
        }
      }
      __syncthreads();
      LTRACE("solves done (all warps)");
      if (MODE == 0) {
        // Schur update of the chunk, S(:, chunk) = A22(:, chunk) - A21 * X: a warp owns row tiles of the boundary
        // block (I = warp mod 8), reads its A21 row tile once from the record (A fragments: 8 consecutive rows x 4
        // columns per load, 64-byte runs; the loads of the next k-tile are issued before the DMMAs of the current
        // one) and sweeps the column tiles of the chunk with X as B fragments from shared memory.
        const int NBT = (nbd + 7) / 8;
        double* Sc = S + cell * (int64_t)nbd * nbd;
        double* gc = g + cell * (int64_t)nbd;
#pragma unroll 1
        for (int I = warp; I < NBT; I += kWarps) {
          const int r = 8 * I + gid;
          const bool rv = r < nbd;
          const int ri = rv ? s_rowinfo[ni + r] : 0;
          const int f = ri >> 8;
          const double* Arow = Arec + (ri & 0xff);
          const double* brow = brec + (ri & 0xff);
          double acc[kCT][2];
#pragma unroll
          for (int j = 0; j < kCT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 8 * (t0 + j) + 2 * tig + e;          // column of [A22 | b2]
              double val = 0.0;
              if (rv && j < nt && c <= nbd) {
                const int off = s_colbase[(ni + c) * nf + f];
                if (off >= 0) val = (c < nbd ? Arow : brow)[off];
              }
              acc[j][e] = val;
            }
          }
          const int* cbf = s_colbase + f;
          auto load_a = [&](int k, double& a0, double& a1) {
            const int kc0 = 8 * k + tig, kc1 = kc0 + 4;
            const int o0 = (rv && kc0 < ni) ? cbf[kc0 * nf] : -1;
            const int o1 = (rv && kc1 < ni) ? cbf[kc1 * nf] : -1;
            const double v0 = Arow[o0 >= 0 ? o0 : 0], v1 = Arow[o1 >= 0 ? o1 : 0];
            a0 = o0 >= 0 ? v0 : 0.0;
            a1 = o1 >= 0 ? v1 : 0.0;
          };
          // A fragments are fetched two k-tiles ahead into a ring of three register pairs; the k loop is unrolled by
          // three so that the ring index is static (a register move of a pending load would wait for it)
          double af0[3], af1[3];
          load_a(0, af0[0], af1[0]);
          if (NT > 1) load_a(1, af0[1], af1[1]);
          const double* xb = Xc + ld * nb;
          const int ld8 = 8 * ld;
#pragma unroll 1
          for (int k3 = 0; k3 < NT; k3 += 3) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              const int k = k3 + u;
              if (k < NT) {
                if (k + 2 < NT) load_a(k + 2, af0[(u + 2) % 3], af1[(u + 2) % 3]);
                const int kr0 = 8 * k + tig, kr1 = kr0 + 4;      // X rows of this lane's B fragments
                // all B fragments first, then one DMMA per column tile and k-half: eight independent accumulators
                // in flight (tiles beyond the chunk multiply zeros into accumulators that are never stored)
                // (rows >= n_i of Xc are zero for the whole kernel and tiles >= nt hold finite leftovers, so the loads
                // need no predicate)
                double bf0[kCT], bf1[kCT];
#pragma unroll
                for (int j = 0; j < kCT; ++j) {
                  bf0[j] = -xb[kr0 + ld8 * j];
                  bf1[j] = -xb[kr1 + ld8 * j];
                }
#pragma unroll
                for (int j = 0; j < kCT; ++j) dmma(acc[j][0], acc[j][1], af0[u], bf0[j]);
#pragma unroll
                for (int j = 0; j < kCT; ++j) dmma(acc[j][0], acc[j][1], af1[u], bf1[j]);
              }
            }
          }
          if (rv) {
#pragma unroll
            for (int j = 0; j < kCT; ++j) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = 8 * (t0 + j) + 2 * tig + e;
                const double val = failed ? qnan : acc[j][e];
                if (j < nt) {
                  if (c < nbd) Sc[r + (int64_t)nbd * c] = val;
                  else if (c == nbd) gc[r] = val;
                }
              }
            }
          }
        }
        __syncthreads();
        LTRACE("Schur update");
      }
    }
    if (info && tid == 0) info[cell] = *s_info;
    __syncthreads();
#ifdef GHB_LTRACE
    if (blockIdx.x == 0 && tid == 0 && cell == 0)
      for (int i = 1; i < t_n; ++i) printf("%-28s %8lld\n", g_lname[i], g_ltrace[i] - g_ltrace[i - 1]);
#endif
  }
}

int ld_for_l(int x) { int v = ((x + 3) / 8) * 8 + 4; return v >= x ? v : v + 8; }

size_t large_smem_bytes(const Plan& p, int ld) {
  const int NT = (p.n_i + 7) / 8;
  size_t d = (size_t)NT * 8 * ld + (size_t)kCT * 8 * ld + 2 * (size_t)NT * 64 + ((p.n_b + 1) & ~1);
  return d * 8 + (size_t)NT * sizeof(LCtl) + 16 + (size_t)(p.n + 1) * p.nfields * 4 + 2 * (size_t)p.n + 32;
}

}  // namespace

bool large_supported(const ghb_ctx* ctx, const Plan& p) {
  if (p.n_i <= 64 || p.n_i > 128 || p.n_b > 255 || p.nfields > 255) return false;
  for (int r = 0; r < p.n; ++r) if (p.row_local[r] > 255) return false;
  return large_smem_bytes(p, ld_for_l(p.n_i)) <= ctx->smem_optin;
}

template <int MODE>
static int launch_large(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                        double* g, int32_t* info, double* X, const double* lam_free, const double* lam_dir,
                        const int64_t* ids, double* u) {
  const int ld = ld_for_l(p.n_i);
  const size_t smem = large_smem_bytes(p, ld);
  auto kern = condense_large_kernel<MODE>;
  GHB_SMEM_OPTIN(ctx, kern, smem);
  LargeTables tb{p.d_colbase, p.d_rowf, p.d_rowl, p.nfields};
  const int64_t grid = std::min<int64_t>(ncells, ctx->sm_count);
  kern<<<(unsigned)grid, kLT, smem, ctx->stream>>>(p.dev(), tb, ld, ncells, A, b, S, g, info, X, lam_free, lam_dir, ids, u);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int launch_condense_large(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                          double* g, int32_t* info, double* X) {
  return launch_large<0>(ctx, p, ncells, A, b, S, g, info, X, nullptr, nullptr, nullptr, nullptr);
}

int launch_backsub_large(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                         const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info) {
  return launch_large<1>(ctx, p, ncells, A, b, nullptr, nullptr, info, nullptr, lam_free, lam_dir, ids, u);
}

}  // namespace ghb
