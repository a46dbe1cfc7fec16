// condense_dmma.cu -- tuned static condensation and backward static condensation for mid-size cells on FP64 DMMA
// tensor cores (mma.sync.m8n8k4.f64; measured full-rate on B200: 37.1 TFLOP/s, see profiles/r01_ubench_fp64.txt).
// Instantiated for 3-D HDG k=2 (n_i, n_b) = (34, 36) and RT-H k=2 / k=3 on quads (33, 12), (56, 16).
//
// Two condensation kernels share the top block; condense_dmma_ll_kernel (left-looking bottom block in registers, 8 cells
// per SM, further down) is the one the library launches, condense_dmma_kernel (right-looking bottom block in shared
// memory, 5 cells per SM) is kept for A/B runs (GHB_DMMA_LL=0).
//
// condense_dmma_kernel: one CTA (4 warps) per cell, 5 CTAs per SM for (34,36).  Replaces
// evaluate!(cache, ::StaticCondensationMap, A, b) (/root/reference/src/StaticCondensationMap.jl:152-196):
//   * the packed record is re-laid out on the fly by branch-free cp.async (16 bytes where the block heights allow it)
//     into two dense column-major images in shared memory: Wt = [A11 A12 b1] (n_i rows) and Bt = [A21 A22 b2]
//     (n_b rows; for (33,12) and (56,16) A22/b2 go straight into the S accumulators instead, Cfg::DIRECT_S);
//   * top block (getrf! :179 + the L-solve half of getrs! :183,:189): blocked right-looking LU of Wt with
//     partial pivoting, panel width 8, with look-ahead: warp 0 only factorises panels (one row per lane,
//     pivot found with one REDUX.MAX on a packed magnitude|row key), warps 1-3 own the column tiles:
//     row gather + unit-lower solve (DMMA with the inverted diagonal block) + trailing update (DMMA);
//     the next panel's column tile is done first and handed back to warp 0 through a named barrier;
//   * bottom block (the U-solve half of getrs!, gemm! :186, gemv! :192 re-associated as
//     S = A22 - (A21 U^-1)(L^-1 P A12)): warps 1-3 own row tiles of Bt; per panel
//     L21 = X * inv(U_pp) and S -= L21 * U12 (DMMA, S accumulators in registers), overlapped with the
//     factorisation of the following panels.
// backsub_dmma_kernel: evaluate!(cache, ::BackwardStaticCondensationMap, A, b, x)
// (src/BackwardStaticCondensationMap.jl:61-102) on the same panel warp / column-tile owners, image [A11 | b1 - A12 x].
// Pivoting: the pivot of a column is an entry whose magnitude equals the column maximum to 2^-15 relative
// (lowest row among those); S does not depend on the pivot order, only rounding does (parity bar 1e-11).
#include <algorithm>

#include "common.cuh"

// resident CTAs per SM the condensation kernels are compiled for (tuning knobs; see profiles/)
#ifndef GHB_MINB34
#define GHB_MINB34 5
#endif
#ifndef GHB_L2PREFETCH
#define GHB_L2PREFETCH 1
#endif
#ifndef GHB_MINB33
#define GHB_MINB33 7
#endif

namespace ghb {

namespace {

// Optional timeline trace (compile with -DGHB_TRACE, see tools/trace_condense.cu): CTA 0 records clock64() at the
// phase boundaries of its first cells into a global buffer.  No effect on the product build.
#ifdef GHB_TRACE
__device__ long long g_trace[64 * 4 * 64];
#define TRACE(ev)                                                                                      \
  do {                                                                                                 \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_cell < 64)                                 \
      g_trace[(trace_cell * 4 + (threadIdx.x >> 5)) * 64 + (ev)] = clock64();                          \
  } while (0)
#else
#define TRACE(ev) do { } while (0)
#endif

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc));
}
// zero-filling variants: `ok == false` writes zeros without reading (src-size 0), so the loader has no branch
__device__ __forceinline__ void cp_async16_z(void* smem_dst, const void* gsrc, bool ok) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gsrc), "r"(ok ? 16 : 0));
}
__device__ __forceinline__ void cp_async8_z(void* smem_dst, const void* gsrc, bool ok) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(ok ? 8 : 0));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ double neg(double x) { return -x; }

// one-instruction L2 prefetch of a contiguous global range (UBLKPF): issued by one thread for the record the CTA
// will process next, so that its cp.async loads hit L2 instead of paying the DRAM latency at the head of the cell
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, unsigned bytes) {
  // the instruction wants a 16-byte aligned address and size: shrink the range to whole granules (the cut-off
  // head/tail, at most 15 bytes each, is fetched by the ordinary loads)
  const unsigned long long a0 = ((unsigned long long)gptr + 15ull) & ~15ull;
  const unsigned long long a1 = ((unsigned long long)gptr + bytes) & ~15ull;
  if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}

// Shared-memory column permutation inside every 8-column tile (logical columns 4<->5 and 6<->7 swapped):
// with a leading dimension of 36 doubles it makes the A, B *and* C fragment access patterns of
// mma.m8n8k4.f64 bank-conflict free at the same time (C fragments touch columns 2t, 2t+1; A/B touch t, t+4).
__device__ __host__ __forceinline__ constexpr int pc(int c) { return c ^ ((c >> 2) & 1); }

// tables of a plan for the on-the-fly re-layout (built on the host, tiny)
struct DmmaTables {
  const int32_t* colbase;  // [(n+1)*nf]: record offset of (first row of field f, column c) or -1; c==n: offset in b
  const uint8_t* rowf;     // [n]: field slot of condensed row r
  const uint8_t* rowl;     // [n]: row inside its field
  int nf;
  const int32_t* xoff;     // left-looking kernel: [BT][CT][2][32] record offset of the accumulator-fragment element of
                           // (row tile, column tile, e, lane); -1 = zero; <= -2: offset -2-v in the b record
};

// smallest leading dimension >= x with LD = 4 (mod 8): 2*LD = 8 or 24 (mod 32) words, the condition for
// conflict-free A/B/C fragment accesses (together with the in-tile column permutation pc())
constexpr int ld_for(int x) { return ((x + 3) / 8) * 8 + 4 >= x ? ((x + 3) / 8) * 8 + 4 : ((x + 3) / 8) * 8 + 12; }

template <int NI, int NB>
struct Cfg {
  static constexpr int N = NI + NB;
  static constexpr int NC = N + 1;             // + rhs column
  static constexpr int RT = (NI + 7) / 8;      // row tiles of the top block
  static constexpr int BT = (NB + 7) / 8;      // row tiles of the bottom block
  static constexpr int CT = (NC + 7) / 8;      // column tiles
  static constexpr int NP = (NI + 7) / 8;      // panels
  static constexpr int LDW = ld_for(NI);       // leading dimensions of the two images
  static constexpr int LDB = ld_for(NB);
  static constexpr int NCP = CT * 8;
  static_assert(NI <= LDW && NB <= LDB && LDW % 8 == 4 && LDB % 8 == 4, "bad leading dimension");
  static_assert(NI <= 64, "the panel warp holds at most two rows per lane");
  static constexpr int WT_DOUBLES = NCP * LDW;
  // DIRECT_S: A22 and b2 go from the record straight into the S accumulators (registers) and Bt keeps only the
  // A21 tiles; pays off where the smaller footprint buys resident CTAs ((33,12): 5 -> 7 per SM, +16%), loses for
  // (34,36) where registers cap the occupancy anyway (profiles/r01_condense_dmma_timeline.md)
  static constexpr bool DIRECT_S = NI != 34;
  static constexpr int BT_DOUBLES = (DIRECT_S ? RT * 8 : NCP) * LDB;
  static size_t smem_bytes(int nf) {   // arrays + 2 panel control blocks + info + re-layout tables
    return (size_t)(WT_DOUBLES + BT_DOUBLES) * 8 + 2 * 1232 + 16 + (size_t)(N + 1) * nf * 4 + 2 * N + 2 * NB + 16;
  }
};

// per-panel control block written by the panel warp (double buffered by panel parity)
struct PanelCtl {
  int ndisp;             // rows displaced out of the diagonal block
  int psrc[8];           // physical row (before this panel's permutation) of the k-th pivot row
  int dsrc[8];           // displaced rows: old position (inside the diagonal block) ...
  int ddst[8];           // ... and the vacated position they move to
  int vac[8];            // scratch: vacated positions by rank
  int pad[3];
  double rinv[8];        // reciprocals of the pivots
  double Linv[64];       // inverse of the unit-lower diagonal block, A operand: Linv[i + 8*k]
  double Dinv[64];       // inverse of the upper diagonal block (npiv x npiv, zero padded), B operand: Dinv[k + 8*n]
};

static_assert(sizeof(PanelCtl) == 1232, "smem_bytes() assumes this");
// named barriers: ids are immediates (a register id makes ptxas reserve all 16 and caps occupancy)
enum { BAR_PANEL = 1, BAR_COL = 3, BAR_UDONE = 5, BAR_UW = 7 };   // + panel parity
enum { BAR_UW_BACK = 6 };   // the backward kernel uses BAR_UDONE once (id 5): 8 barriers -> 8 CTAs per SM
template <int ID, int COUNT>
__device__ __forceinline__ void bar_sync_i() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT>
__device__ __forceinline__ void bar_arrive_i() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT>
__device__ __forceinline__ void bar_sync(int parity) { if (parity) bar_sync_i<ID + 1, COUNT>(); else bar_sync_i<ID, COUNT>(); }
template <int ID, int COUNT>
__device__ __forceinline__ void bar_arrive(int parity) { if (parity) bar_arrive_i<ID + 1, COUNT>(); else bar_arrive_i<ID, COUNT>(); }

// fast reciprocal: MUFU.RCP64H seed + two Newton steps (branch free; ~1 ulp)
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// ---- panel factorisation: one warp, one row per lane, implicit pivoting --------------------------
// Factorises columns [c0, c0+npiv) of Wt over rows [c0, NI) and carries the other columns of the 8-wide
// column tile through the eliminations.  Rows are not exchanged while factorising: a lane keeps its row
// and remembers at which step it was chosen.  On write-back the k-th pivot row goes to position c0+k and
// the rows it displaces from the diagonal block go to the vacated positions (any consistent row order is
// a valid row-permuted LU; S does not depend on it).  Publishes that permutation, inv(L_pp) and inv(U_pp).
template <int NI, int LDW, bool TWO>
__device__ __forceinline__ void panel_factor(double* __restrict__ Wt, const int c0, const int npiv, PanelCtl* ctl,
                                             int* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const int nrows = NI - c0;
  const bool v1 = lane < nrows;
  const bool v2 = TWO && (lane + 32 < nrows);
  double a[8], a2[8];
  int ch1 = -1, ch2 = -1;                       // step at which this lane's row became the pivot row
  double* base = Wt + c0 + lane + LDW * c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = v1 ? base[LDW * pc(j)] : 0.0;
    a2[j] = v2 ? base[32 + LDW * pc(j)] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < npiv) {
      // ---- pivot search: one REDUX.MAX over key = |a| (exponent + 15 mantissa bits) << 6 | (63 - row)
      const bool c1 = v1 && ch1 < 0;
      const bool c2 = v2 && ch2 < 0;
      const unsigned h1 = (unsigned)(__double_as_longlong(a[k]) >> 32) & 0x7fffffffu;
      const unsigned h2 = (unsigned)(__double_as_longlong(a2[k]) >> 32) & 0x7fffffffu;
      unsigned key = c1 ? (((h1 >> 5) << 6) | (unsigned)(63 - lane)) : 0u;
      if (TWO) {
        const unsigned key2 = c2 ? (((h2 >> 5) << 6) | (unsigned)(31 - lane)) : 0u;
        key = key > key2 ? key : key2;
      }
      const double rc1 = fast_rcp(a[k]);        // speculative reciprocal, overlaps the reduction
      const double rc2 = TWO ? fast_rcp(a2[k]) : 0.0;
      unsigned kmax = __reduce_max_sync(0xffffffffu, key);
      if ((kmax >> 6) == 0u) {
        // all candidates below 2^-1017: decide exactly (zero column => LAPACK info = k+1)
        const unsigned long long e1 = c1 ? ((unsigned long long)__double_as_longlong(a[k]) & 0x7fffffffffffffffull) : 0ull;
        const unsigned long long e2 = c2 ? ((unsigned long long)__double_as_longlong(a2[k]) & 0x7fffffffffffffffull) : 0ull;
        const unsigned lo1 = __reduce_max_sync(0xffffffffu, (unsigned)(e1 >> 32) | (unsigned)(e2 >> 32));
        const unsigned lo2 = __reduce_max_sync(0xffffffffu, (unsigned)e1 | (unsigned)e2);
        if ((lo1 | lo2) == 0u) {
          if (lane == 0 && *info == 0) *info = c0 + k + 1;
        }
        // keep going with the first candidate row (results of a failed cell are overwritten with NaN)
        const unsigned bb1 = __ballot_sync(0xffffffffu, c1);
        const unsigned bb2 = __ballot_sync(0xffffffffu, c2);
        const int row = bb1 ? (__ffs(bb1) - 1) : (32 + __ffs(bb2) - 1);
        kmax = (unsigned)(63 - row);
      }
      const int prow = 63 - (int)(kmax & 63u);  // row of the pivot inside the panel (0..63)
      const bool from2 = TWO && prow >= 32;
      const int q = prow & 31;
      const double rinv = __shfl_sync(0xffffffffu, from2 ? rc2 : rc1, q);
      if (lane == 0) ctl->rinv[k] = rinv;     // published with the rest of the control block
      const bool me1 = !from2 && lane == q, me2 = from2 && lane == q;
      if (me1) ch1 = k;
      if (me2) ch2 = k;
      // ---- multipliers and rank-1 update of the rows still in play
      const bool u1 = c1 && !me1, u2 = TWO && c2 && !me2;
      const double l1 = a[k] * rinv, l2 = a2[k] * rinv;   // dgetf2: scale by the reciprocal
      a[k] = u1 ? l1 : a[k];
      const double m1 = u1 ? l1 : 0.0;                    // rows out of play: a - 0*p = a
      double m2 = 0.0;
      if (TWO) { a2[k] = u2 ? l2 : a2[k]; m2 = u2 ? l2 : 0.0; }
#pragma unroll
      for (int j = k + 1; j < 8; ++j) {
        const double pj = __shfl_sync(0xffffffffu, from2 ? a2[j] : a[j], q);   // pivot row entry, to everyone
        a[j] = fma(-m1, pj, a[j]);                        // (the negation is an operand modifier of DFMA)
        if (TWO) a2[j] = fma(-m2, pj, a2[j]);
      }
    }
  }
  // ---- new positions: pivot rows first, displaced rows into the vacated slots
  const bool disp = v1 && lane < npiv && ch1 < 0;          // rows of the diagonal block not chosen
  const bool vc1 = v1 && lane >= npiv && ch1 >= 0;
  const bool vc2 = v2 && ch2 >= 0;
  const unsigned mdisp = __ballot_sync(0xffffffffu, disp);
  const unsigned mv1 = __ballot_sync(0xffffffffu, vc1);
  const unsigned mv2 = TWO ? __ballot_sync(0xffffffffu, vc2) : 0u;
  const unsigned lt = (1u << lane) - 1u;
  if (vc1) ctl->vac[__popc(mv1 & lt)] = c0 + lane;
  if (vc2) ctl->vac[__popc(mv1) + __popc(mv2 & lt)] = c0 + 32 + lane;
  __syncwarp();
  int np1 = c0 + lane, np2 = c0 + 32 + lane;
  if (ch1 >= 0) np1 = c0 + ch1;
  else if (disp) np1 = ctl->vac[__popc(mdisp & lt)];
  if (ch2 >= 0) np2 = c0 + ch2;
  if (ch1 >= 0) ctl->psrc[ch1] = c0 + lane;
  if (ch2 >= 0) ctl->psrc[ch2] = c0 + 32 + lane;
  if (disp) {
    const int t = __popc(mdisp & lt);
    ctl->dsrc[t] = c0 + lane;
    ctl->ddst[t] = np1;
  }
  if (lane == 0) ctl->ndisp = __popc(mdisp);
  double* wb = Wt + LDW * c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (v1) wb[np1 + LDW * pc(j)] = a[j];
    if (v2) wb[np2 + LDW * pc(j)] = a2[j];
  }
  if (lane >= npiv && lane < 8) ctl->rinv[lane] = 0.0;   // beyond a partial panel
}

// ---- inverses of the 8x8 diagonal block of a factorised panel, by the (otherwise idle) update warps:
// lanes 0-7 each own one column and run the same branch-free substitution; entries of L\U are uniform
// (broadcast) shared-memory loads.  Rows/columns >= npiv are treated as identity (L) / zero (U^-1).
template <int LDW>
__device__ __forceinline__ void invert_unit_lower(const double* __restrict__ D, const int npiv, double* __restrict__ Linv) {
  const int n = threadIdx.x & 31;              // column (only lanes 0-7 store)
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (i == n) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m < i; ++m) {
      const double lim = i < npiv ? D[i + LDW * pc(m)] : 0.0;
      if (m & 1) s1 = fma(-lim, x[m], s1); else s0 = fma(-lim, x[m], s0);
    }
    x[i] = (i > n) ? (s0 + s1) : x[i];
  }
  if (n < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) Linv[i + 8 * n] = x[i];
  }
}

template <int LDW>
__device__ __forceinline__ void invert_upper(const double* __restrict__ D, const int npiv, const double* __restrict__ rinv,
                                             double* __restrict__ Dinv) {
  const int n = threadIdx.x & 31;              // column (only lanes 0-7 store)
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (i == n && n < npiv) ? rinv[i] : 0.0;
#pragma unroll
  for (int i = 6; i >= 0; --i) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = i + 1; m < 8; ++m) {
      const double uim = m < npiv ? D[i + LDW * pc(m)] : 0.0;
      if (m & 1) s1 = fma(uim, x[m], s1); else s0 = fma(uim, x[m], s0);
    }
    x[i] = (i < n && n < npiv) ? -(s0 + s1) * rinv[i] : x[i];
  }
  if (n < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) Dinv[i + 8 * n] = x[i];
  }
}

// RPC = rows per cp.async: 2 (16 bytes) when every vertical pair of the condensed matrix is contiguous and aligned in
// the packed record (all block heights even), else 1 (8 bytes).
template <int NI, int NB, int RPC>
__global__ void __launch_bounds__(128, (NI > 40 ? 4 : (NI == 34 ? GHB_MINB34 : GHB_MINB33)))
condense_dmma_kernel(DmmaTables tb, int lenA, int lenb, int64_t ncells, const double* __restrict__ A,
                     const double* __restrict__ b, double* __restrict__ S, double* __restrict__ g,
                     int32_t* __restrict__ info) {
  using C = Cfg<NI, NB>;
  constexpr int N = C::N, LDW = C::LDW, LDB = C::LDB, RT = C::RT, BT = C::BT, CT = C::CT, NP = C::NP;
  constexpr int SJ0 = NI / 8;                   // first column tile holding S columns
  constexpr int NSJ = CT - SJ0;                 // S column tiles
  constexpr int MAXROWS = (BT + 2) / 3;         // bottom row tiles per update warp
  extern __shared__ __align__(16) double smem[];
  double* Wt = smem;
  double* Bt = Wt + C::WT_DOUBLES;
  PanelCtl* ctl2 = reinterpret_cast<PanelCtl*>(Bt + C::BT_DOUBLES);   // [2]
  int* s_info = reinterpret_cast<int*>(ctl2 + 2);
  int* s_colbase = s_info + 4;                                         // [(N+1)*nf]
  unsigned short* s_rowinfo = reinterpret_cast<unsigned short*>(s_colbase + (N + 1) * tb.nf);  // [N/RPC]
  unsigned short* s_rowinfo2 = s_rowinfo + N / RPC;                                            // [NB] boundary rows
  static_assert(RPC == 1 || (NI % 2 == 0 && NB % 2 == 0), "16-byte cp.async needs even block heights");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;    // fragment coordinates
  // memory columns (inside a tile) of this lane's fragment elements, see pc()
  const int ce0 = pc(2 * tig), ce1 = pc(2 * tig + 1);   // C fragment
  const int ka0 = tig, ka1 = pc(4 + tig);               // A fragment (k-steps 0, 1)
  const int nb = pc(gid);                               // B fragment column

  // padding never written by the loader: zero it once (rows NI.. of Wt, rows NB.. of Bt, column NC)
  for (int i = tid; i < C::WT_DOUBLES; i += 128) Wt[i] = 0.0;
  for (int i = tid; i < C::BT_DOUBLES; i += 128) Bt[i] = 0.0;
  for (int i = tid; i < (N + 1) * tb.nf; i += 128) s_colbase[i] = tb.colbase[i];
  for (int i = tid; i < N / RPC; i += 128) s_rowinfo[i] = (unsigned short)((tb.rowf[RPC * i] << 8) | tb.rowl[RPC * i]);
  for (int i = tid; i < NB; i += 128) s_rowinfo2[i] = (unsigned short)((tb.rowf[NI + i] << 8) | tb.rowl[NI + i]);
  __syncthreads();

  // loader: a thread owns one copy unit (RPC rows; interior rows -> Wt, boundary rows -> Bt) and walks the columns
  constexpr int HP = N / RPC;                   // copy units per column
  constexpr int LG = 128 / HP;                  // column groups
  static_assert(LG >= 1, "cell too tall for the loader");
  constexpr int LB = 6;                         // loader batch: look-ups in flight per thread
  const int l_grp = tid / HP, l_rp = tid - l_grp * HP;
  const bool l_on = l_grp < LG;
  const int l_ri = l_on ? s_rowinfo[l_rp] : 0;
  const int l_f = l_ri >> 8, l_lr = l_ri & 0xff;
  const bool l_top = RPC * l_rp < NI;
  double* const l_dst0 = l_top ? Wt + RPC * l_rp : Bt + (RPC * l_rp - NI);
  const int l_ld = l_top ? LDW : LDB;

  int trace_cell = 0; (void)trace_cell;
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x, ++trace_cell) {
    // ------------------------------------------------------------------ load + re-layout
    TRACE(0);
    if (l_on) {
      const double* Arec = A + cell * lenA + l_lr;
      const double* brec = b + cell * lenb + l_lr;
      const int* cb = s_colbase + l_f;
      const int cend = (l_top || !C::DIRECT_S) ? N : 8 * SJ0;   // DIRECT_S: only the A21 tiles of boundary rows live in Bt
      // batches of LB columns: all table look-ups first, then the copies, so that the shared-memory latency of the
      // look-up is paid once per batch instead of once per copy (the copies are volatile asm and do not reorder)
#pragma unroll 1
      for (int cb0 = l_grp; cb0 < cend; cb0 += LB * LG) {
        int off[LB];
#pragma unroll
        for (int q = 0; q < LB; ++q) { const int c = cb0 + q * LG; off[q] = c < cend ? cb[c * tb.nf] : -2; }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
          const int c = cb0 + q * LG;
          double* dst = l_dst0 + l_ld * pc(c);
          const double* src = Arec + (off[q] >= 0 ? off[q] : 0);   // untouched block: zero fill, nothing read
          if (off[q] != -2) { if (RPC == 2) cp_async16_z(dst, src, off[q] >= 0); else cp_async8_z(dst, src, off[q] >= 0); }
        }
      }
      if ((l_top || !C::DIRECT_S) && l_grp == N % LG) {                                              // rhs column
        if (RPC == 2) cp_async16(l_dst0 + l_ld * pc(N), brec + cb[N * tb.nf]);
        else cp_async8(l_dst0 + l_ld * pc(N), brec + cb[N * tb.nf]);
      }
    }
    // S accumulators of the owned bottom row tiles (column tiles SJ0..CT-1) as C fragments, straight from the
    // record: 8 consecutive rows x 4 columns per load instruction, i.e. full 64-byte runs of the packed columns
    double acc[MAXROWS][NSJ][2];
    if (C::DIRECT_S && warp != 0) {
      const double* Arec = A + cell * lenA;
      const double* brec = b + cell * lenb;
#pragma unroll
      for (int ri = 0; ri < MAXROWS; ++ri) {
        const int I = (warp - 1) * MAXROWS + ri;
        const int r = 8 * I + gid;
        const bool rv = I < BT && r < NB;
        const int ri_ = rv ? s_rowinfo2[r] : 0;
        const int f = ri_ >> 8, lr = ri_ & 0xff;
#pragma unroll
        for (int js = 0; js < NSJ; ++js) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * (SJ0 + js) + 2 * tig + e;
            const int off = (rv && c <= N) ? s_colbase[c * tb.nf + f] : -1;
            acc[ri][js][e] = off >= 0 ? (c < N ? Arec : brec)[off + lr] : 0.0;
          }
        }
      }
    }
    TRACE(1);
    if (tid == 0) *s_info = 0;
#if GHB_L2PREFETCH
    if (tid == 32 && cell + gridDim.x < ncells) {
      l2_prefetch_bulk(A + (cell + gridDim.x) * lenA, (unsigned)(lenA * 8));
      l2_prefetch_bulk(b + (cell + gridDim.x) * lenb, (unsigned)(lenb * 8));
    }
#endif
    cp_async_commit_wait_all();
    __syncthreads();
    TRACE(2);

    if (warp == 0) {
      // ================================================================ panel warp
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
        if (p > 0) bar_sync<BAR_COL, 64>(p & 1);             // column tile p is up to date
        TRACE(4 + 6 * p);
        PanelCtl* ctl = ctl2 + (p & 1);
        if ((NI - c0) > 32) panel_factor<NI, LDW, true>(Wt, c0, npiv, ctl, s_info);
        else panel_factor<NI, LDW, false>(Wt, c0, npiv, ctl, s_info);
        TRACE(5 + 6 * p);
        bar_arrive<BAR_PANEL, 128>(p & 1);
        // inv(U_pp) for the bottom block, computed while the update warps prepare the next column tile.
        // (inv(L_pp) stays with the tile owner here: with both inverses the panel warp becomes as critical as the
        // update warps, 39.2 -> 38.4 M cells/s; the backward kernel, whose update warps are light, moves it here)
        invert_upper<LDW>(Wt + c0 + LDW * c0, npiv, ctl->rinv, ctl->Dinv);
        TRACE(6 + 6 * p);
        bar_arrive<BAR_UDONE, 128>(p & 1);
      }
    } else {
      // ================================================================ update warps
      const int uw = warp - 1;                    // 0..2
      if (!C::DIRECT_S) {
#pragma unroll
        for (int ri = 0; ri < MAXROWS; ++ri) {
          const int I = uw * MAXROWS + ri;
          const int r = 8 * I + gid;
          const bool rv = I < BT && r < NB;
#pragma unroll
          for (int js = 0; js < NSJ; ++js) {
            const double* cp0 = Bt + r + LDB * 8 * (SJ0 + js);
            acc[ri][js][0] = rv ? cp0[LDB * ce0] : 0.0;
            acc[ri][js][1] = rv ? cp0[LDB * ce1] : 0.0;
          }
        }
      }
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
        PanelCtl* ctl = ctl2 + (p & 1);
        bar_sync<BAR_PANEL, 128>(p & 1);
        TRACE(4 + 6 * p);
        // ---- the owner of the next panel's column tile inverts L_pp (on the critical path)
        if (uw == (p + 1) % 3) invert_unit_lower<LDW>(Wt + c0 + LDW * c0, npiv, ctl->Linv);
        // ---- owned column tiles J > p (J = uw mod 3); the next panel's tile first
        const int nd = ctl->ndisp;
        const int ps0 = tig < npiv ? ctl->psrc[tig] : -1;
        const int ps1 = 4 + tig < npiv ? ctl->psrc[4 + tig] : -1;
        const int dsr = gid < nd ? ctl->dsrc[gid] : -1;
        const int dds = gid < nd ? ctl->ddst[gid] : -1;
        bar_sync<BAR_UW, 96>(p & 1);
        TRACE(5 + 6 * p);
        const double li0 = ctl->Linv[gid + 8 * tig], li1 = ctl->Linv[gid + 8 * (4 + tig)];
        int Jfirst = p + 1 + (uw + 3 - (p + 1) % 3) % 3;   // first owned tile > p
#pragma unroll 1
        for (int J = Jfirst; J < CT; J += 3) {
          double* colg = Wt + LDW * (8 * J + nb);      // B-fragment column of this lane
          double* colt = Wt + LDW * (8 * J + tig);     // displaced rows: lane (t=gid) moves columns tig, tig+4
          const double g0 = ps0 >= 0 ? colg[ps0] : 0.0;
          const double g1 = ps1 >= 0 ? colg[ps1] : 0.0;
          const double dv0 = dsr >= 0 ? colt[dsr] : 0.0;
          const double dv1 = dsr >= 0 ? colt[dsr + 4 * LDW] : 0.0;
          __syncwarp();
          double u0 = 0.0, u1 = 0.0;              // U12 tile = inv(L_pp) * gathered rows
          dmma(u0, u1, li0, g0);
          dmma(u0, u1, li1, g1);
          double* cc0 = Wt + gid + LDW * (8 * J + ce0);      // C-fragment columns of this lane in tile J
          double* cc1 = Wt + gid + LDW * (8 * J + ce1);
          if (c0 + gid < NI) { cc0[c0] = u0; cc1[c0] = u1; }
          if (dds >= 0) { colt[dds] = dv0; colt[dds + 4 * LDW] = dv1; }
          __syncwarp();
          if (p + 1 < RT) {
            const double bf0 = neg(colg[c0 + tig]);
            const double bf1 = neg(colg[c0 + 4 + tig]);
            // two row tiles per step with independent accumulators, DMMAs interleaved (the asm statements keep
            // their order): a single accumulator serialises the whole column tile on the 26-cycle DMMA latency.
            // The panel's multipliers (A fragments) are re-read per tile: holding them in registers across the
            // tiles cost more in register pressure than the loads (36.4 vs 38.7 M cells/s).
            {
              int I = p + 1;
#pragma unroll 1
              for (; I + 1 < RT; I += 2) {
                const int rA = 8 * I + gid, rB = rA + 8;
                const bool vA = rA < NI, vB = rB < NI;
                const double aA0 = vA ? Wt[rA + LDW * (c0 + ka0)] : 0.0, aA1 = vA ? Wt[rA + LDW * (c0 + ka1)] : 0.0;
                const double aB0 = vB ? Wt[rB + LDW * (c0 + ka0)] : 0.0, aB1 = vB ? Wt[rB + LDW * (c0 + ka1)] : 0.0;
                double dA0 = vA ? cc0[8 * I] : 0.0, dA1 = vA ? cc1[8 * I] : 0.0;
                double dB0 = vB ? cc0[8 * I + 8] : 0.0, dB1 = vB ? cc1[8 * I + 8] : 0.0;
                dmma(dA0, dA1, aA0, bf0);
                dmma(dB0, dB1, aB0, bf0);
                dmma(dA0, dA1, aA1, bf1);
                dmma(dB0, dB1, aB1, bf1);
                if (vA) { cc0[8 * I] = dA0; cc1[8 * I] = dA1; }
                if (vB) { cc0[8 * I + 8] = dB0; cc1[8 * I + 8] = dB1; }
              }
              if (I < RT) {
                const int r = 8 * I + gid;
                const bool rv = r < NI;
                const double a0 = rv ? Wt[r + LDW * (c0 + ka0)] : 0.0;
                const double a1 = rv ? Wt[r + LDW * (c0 + ka1)] : 0.0;
                double d0 = rv ? cc0[8 * I] : 0.0, d1 = rv ? cc1[8 * I] : 0.0;
                dmma(d0, d1, a0, bf0);
                dmma(d0, d1, a1, bf1);
                if (rv) { cc0[8 * I] = d0; cc1[8 * I] = d1; }
              }
            }
          }
          if (J == p + 1 && p + 1 < NP) bar_arrive<BAR_COL, 64>((p + 1) & 1);
        }
        // ---- bottom block, owned row tiles: needs every U[p][J] and Dinv_p (from the panel warp)
        TRACE(6 + 6 * p);
        bar_sync<BAR_UDONE, 128>(p & 1);
        TRACE(7 + 6 * p);
        const double dj0 = ctl->Dinv[tig + 8 * gid], dj1 = ctl->Dinv[4 + tig + 8 * gid];
        const double* ub0 = Wt + c0 + tig + LDW * nb;        // B fragment of U[p][J]: ub0[LDW*8*J], ub0[4 + LDW*8*J]
        if (npiv == 8) {
          // (every DMMA sequence below is written k-step outermost so that consecutive instructions belong to
          // independent accumulators: a warp issues in order, and the asm statements keep their order)
          double la[MAXROWS][2];
          {
            // L = X * Dinv with X = Bt tile (I,p) (already updated right-looking): A fragments straight from smem
            double xa[MAXROWS][2], l[MAXROWS][2];
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) {
              const int r = 8 * (uw * MAXROWS + ri) + gid;
              const bool rv = uw * MAXROWS + ri < BT && r < NB;
              xa[ri][0] = rv ? Bt[r + LDB * (c0 + ka0)] : 0.0;
              xa[ri][1] = rv ? Bt[r + LDB * (c0 + ka1)] : 0.0;
              l[ri][0] = 0.0; l[ri][1] = 0.0;
            }
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) dmma(l[ri][0], l[ri][1], xa[ri][0], dj0);
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) dmma(l[ri][0], l[ri][1], xa[ri][1], dj1);
            __syncwarp();
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) {
              const int r = 8 * (uw * MAXROWS + ri) + gid;
              const bool rv = uw * MAXROWS + ri < BT && r < NB;
              double* bp = Bt + r + LDB * c0;
              if (rv) { bp[LDB * ce0] = l[ri][0]; bp[LDB * ce1] = l[ri][1]; }   // park L to read it back as A fragments
            }
            __syncwarp();
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) {
              const int r = 8 * (uw * MAXROWS + ri) + gid;
              const bool rv = uw * MAXROWS + ri < BT && r < NB;
              const double* bp = Bt + r + LDB * c0;
              la[ri][0] = rv ? bp[LDB * ka0] : 0.0;
              la[ri][1] = rv ? bp[LDB * ka1] : 0.0;
            }
          }
          // A21 part still in shared memory: tiles p < J < SJ0
#pragma unroll 1
          for (int J = p + 1; J < SJ0; ++J) {
            const double bf0 = neg(ub0[LDW * 8 * J]), bf1 = neg(ub0[4 + LDW * 8 * J]);
            double d[MAXROWS][2];
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) {
              const int r = 8 * (uw * MAXROWS + ri) + gid;
              const bool rv = uw * MAXROWS + ri < BT && r < NB;
              const double* cp0 = Bt + r + LDB * 8 * J;
              d[ri][0] = rv ? cp0[LDB * ce0] : 0.0; d[ri][1] = rv ? cp0[LDB * ce1] : 0.0;
            }
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) dmma(d[ri][0], d[ri][1], la[ri][0], bf0);
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) dmma(d[ri][0], d[ri][1], la[ri][1], bf1);
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri) {
              const int r = 8 * (uw * MAXROWS + ri) + gid;
              const bool rv = uw * MAXROWS + ri < BT && r < NB;
              double* cp0 = Bt + r + LDB * 8 * J;
              if (rv) { cp0[LDB * ce0] = d[ri][0]; cp0[LDB * ce1] = d[ri][1]; }
            }
          }
          // S part in registers
          // (rows beyond the block have la = 0: their DMMAs leave the unused accumulators alone)
#pragma unroll
          for (int js = 0; js < NSJ; js += 2) {
            constexpr int W = 2;
            double bf[W][2];
#pragma unroll
            for (int q = 0; q < W; ++q) {
              const int J = SJ0 + (js + q < NSJ ? js + q : js);
              bf[q][0] = neg(ub0[LDW * 8 * J]); bf[q][1] = neg(ub0[4 + LDW * 8 * J]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
              for (int q = 0; q < W; ++q)
#pragma unroll
                for (int ri = 0; ri < MAXROWS; ++ri)
                  if (js + q < NSJ) dmma(acc[ri][js + q < NSJ ? js + q : js][0], acc[ri][js + q < NSJ ? js + q : js][1], la[ri][k], bf[q][k]);
          }
        } else {
          // partial last panel (npiv < 8): its column tile is the first S tile, held in registers.
          // X (C fragment) -> A fragment through the (dead) Bt tile; L = X * Dinv (zero outside npiv cols)
          double lq[MAXROWS];
#pragma unroll
          for (int ri = 0; ri < MAXROWS; ++ri) {
            const int I = uw * MAXROWS + ri;
            const int r = 8 * I + gid;
            const bool rv = I < BT && r < NB;
            double* bp = Bt + r + LDB * c0;
            __syncwarp();
            if (rv) { bp[LDB * ce0] = acc[ri][0][0]; bp[LDB * ce1] = acc[ri][0][1]; }
            __syncwarp();
            const double xa0 = rv ? bp[LDB * ka0] : 0.0;
            double l0 = 0.0, l1 = 0.0;
            dmma(l0, l1, xa0, dj0);
            __syncwarp();
            if (rv) { bp[LDB * ce0] = l0; bp[LDB * ce1] = l1; }
            __syncwarp();
            lq[ri] = rv ? bp[LDB * ka0] : 0.0;
          }
          // trailing columns of the panel's own tile: B = U_pp rows < npiv, columns >= npiv
          const double ub = (tig < npiv && gid >= npiv) ? neg(Wt[c0 + tig + LDW * (c0 + nb)]) : 0.0;
#pragma unroll
          for (int ri = 0; ri < MAXROWS; ++ri)
            if (uw * MAXROWS + ri < BT) dmma(acc[ri][0][0], acc[ri][0][1], lq[ri], ub);
#pragma unroll
          for (int js = 1; js < NSJ; ++js) {
            const double bf0 = neg(ub0[LDW * 8 * (SJ0 + js)]);
#pragma unroll
            for (int ri = 0; ri < MAXROWS; ++ri)
              if (uw * MAXROWS + ri < BT) dmma(acc[ri][js][0], acc[ri][js][1], lq[ri], bf0);
          }
        }
        TRACE(8 + 6 * p);
      }
      TRACE(40);
      // ---- store S, g
      const bool failed = *s_info != 0;
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      double* Sc = S + cell * (int64_t)NB * NB;
      double* gc = g + cell * (int64_t)NB;
#pragma unroll
      for (int ri = 0; ri < MAXROWS; ++ri) {
        const int I = uw * MAXROWS + ri;
        const int r = 8 * I + gid;
        if (I < BT && r < NB) {
#pragma unroll
          for (int js = 0; js < NSJ; ++js) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 8 * (SJ0 + js) + 2 * tig + e;
              const double v = failed ? qnan : acc[ri][js][e];
              if (c >= NI) {
                if (c < N) Sc[r + (int64_t)NB * (c - NI)] = v;
                else if (c == N) gc[r] = v;
              }
            }
          }
        }
      }
    }
    TRACE(41);
    __syncthreads();
    TRACE(42);
    if (info && tid == 0) info[cell] = *s_info;
  }
}

// ---- condensation with a left-looking, register-resident bottom block -----------------------------
// Same top block as condense_dmma_kernel (panel warp + column-tile owners, look-ahead), but only Wt = [A11 A12 b1]
// lives in shared memory: 26.8 KB per cell for (34,36) instead of 45 KB, and no S accumulators are carried through
// the top block (64 registers), so 8 cells are resident per SM instead of 5.  The bottom block [A21 A22 b2] never
// touches shared memory: after the top block is final every warp takes row tiles of 8 boundary rows, reads them from
// the record as accumulator fragments (64-byte runs of the packed columns; the record was pulled into L2 by the bulk
// prefetch) and runs, panel by panel,
//     Y_p = -X_p inv(U_pp),   X_J += Y_p U_pJ  (J > p)
// entirely in registers (S = A22 - (A21 U^-1)(L^-1 P A12), the same re-association as above).  The accumulator
// fragment of Y_p is fed back as the A operand without a shared-memory round trip: accumulator column n of a tile
// stands for logical column sg(n), so that k-step s of lane tig is logical column sg(2 tig + s) and the B fragments of
// U_pJ are read from rows c0 + sg(2 tig + s) -- with sg = (0,2,1,3,6,4,7,5) these loads are bank-conflict free for
// leading dimensions = 4 (mod 8) under the pc() column permutation (tools/emulate_bottom.py checks indices and banks).
__device__ __forceinline__ int sg8(int n) { return (int)((0x57463120u >> (4 * n)) & 7u); }

template <int NI, int NB>
struct LLCfg {
  using C = Cfg<NI, NB>;
  static size_t smem_bytes(int nf) {   // Wt + inv(U_pp) of every panel + 2 panel control blocks + info + tables
    return (size_t)(C::WT_DOUBLES + C::NP * 64) * 8 + 2 * sizeof(PanelCtl) + 16 + (size_t)(C::N + 1) * nf * 4 +
           2 * NI + 16;
  }
};
enum { BAR_UW_LL = 5 };   // + panel parity: ids 0..6 -> 7 barriers per CTA, 8 CTAs per SM

template <int NI, int NB, int RPC, int MINB, bool KEEPX>
__global__ void __launch_bounds__(128, MINB)
condense_dmma_ll_kernel(DmmaTables tb, int lenA, int lenb, int64_t ncells, const double* __restrict__ A,
                        const double* __restrict__ b, double* __restrict__ S, double* __restrict__ g,
                        int32_t* __restrict__ info, double* __restrict__ X) {
  using C = Cfg<NI, NB>;
  constexpr int N = C::N, LDW = C::LDW, RT = C::RT, BT = C::BT, CT = C::CT, NP = C::NP;
  constexpr int SJ0 = NI / 8;                   // first column tile holding S columns
  extern __shared__ __align__(16) double smem[];
  double* Wt = smem;
  double* s_dinv = Wt + C::WT_DOUBLES;                                  // [NP][64]: inv(U_pp), B operand k + 8 n
  PanelCtl* ctl2 = reinterpret_cast<PanelCtl*>(s_dinv + NP * 64);      // [2]
  int* s_info = reinterpret_cast<int*>(ctl2 + 2);
  int* s_colbase = s_info + 4;                                         // [(N+1)*nf]
  unsigned short* s_rowinfo = reinterpret_cast<unsigned short*>(s_colbase + (N + 1) * tb.nf);  // [NI/RPC] interior rows
  static_assert(RPC == 1 || NI % 2 == 0, "16-byte cp.async needs an even interior height");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;    // fragment coordinates
  const int ce0 = pc(2 * tig), ce1 = pc(2 * tig + 1);   // top block: C fragment
  const int ka0 = tig, ka1 = pc(4 + tig);               // top block: A fragment (k-steps 0, 1)
  const int nb = pc(gid);                               // top block: B fragment column

  for (int i = tid; i < C::WT_DOUBLES; i += 128) Wt[i] = 0.0;   // padding never written by the loader
  for (int i = tid; i < (N + 1) * tb.nf; i += 128) s_colbase[i] = tb.colbase[i];
  constexpr int HP = NI / RPC;                  // copy units per column (interior rows only)
  constexpr int LG = 128 / HP;                  // column groups
  static_assert(LG >= 1, "cell too tall for the loader");
  constexpr int LB = 6;                         // loader batch: look-ups in flight per thread
  for (int i = tid; i < HP; i += 128) s_rowinfo[i] = (unsigned short)((tb.rowf[RPC * i] << 8) | tb.rowl[RPC * i]);
  __syncthreads();
  const int l_grp = tid / HP, l_rp = tid - l_grp * HP;
  const bool l_on = l_grp < LG;
  const int l_ri = l_on ? s_rowinfo[l_rp] : 0;
  const int l_f = l_ri >> 8, l_lr = l_ri & 0xff;
  double* const l_dst0 = Wt + RPC * l_rp;

  int trace_cell = 0; (void)trace_cell;
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x, ++trace_cell) {
    // ------------------------------------------------------------------ load + re-layout of the interior rows
    TRACE(0);
    if (l_on) {
      const double* Arec = A + cell * lenA + l_lr;
      const double* brec = b + cell * lenb + l_lr;
      const int* cb = s_colbase + l_f;
#pragma unroll 1
      for (int cb0 = l_grp; cb0 < N; cb0 += LB * LG) {
        int off[LB];
#pragma unroll
        for (int q = 0; q < LB; ++q) { const int c = cb0 + q * LG; off[q] = c < N ? cb[c * tb.nf] : -2; }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
          const int c = cb0 + q * LG;
          double* dst = l_dst0 + LDW * pc(c);
          const double* src = Arec + (off[q] >= 0 ? off[q] : 0);   // untouched block: zero fill, nothing read
          if (off[q] != -2) { if (RPC == 2) cp_async16_z(dst, src, off[q] >= 0); else cp_async8_z(dst, src, off[q] >= 0); }
        }
      }
      if (l_grp == N % LG) {                                                                         // rhs column
        if (RPC == 2) cp_async16(l_dst0 + LDW * pc(N), brec + cb[N * tb.nf]);
        else cp_async8(l_dst0 + LDW * pc(N), brec + cb[N * tb.nf]);
      }
    }
    if (tid == 0) *s_info = 0;
#if GHB_L2PREFETCH
    // the boundary rows of this record are read (from L2) only after the top block, a whole cell latency from now: one
    // bulk prefetch of the record.  (One cell *ahead*, as in the right-looking kernel, keeps 2 x 8 x 148 records in
    // flight, more than L2 holds: measured 1.7x the algorithmic DRAM reads.)
    if (tid == 32) l2_prefetch_bulk(A + cell * lenA, (unsigned)(lenA * 8));
#endif
    TRACE(1);
    cp_async_commit_wait_all();
    __syncthreads();
    TRACE(2);

    if (warp == 0) {
      // ================================================================ panel warp
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
        PanelCtl* ctl = ctl2 + (p & 1);
        if (p > 0) bar_sync<BAR_COL, 64>(p & 1);             // column tile p is up to date
        TRACE(4 + 6 * p);
        if ((NI - c0) > 32) panel_factor<NI, LDW, true>(Wt, c0, npiv, ctl, s_info);
        else panel_factor<NI, LDW, false>(Wt, c0, npiv, ctl, s_info);
        bar_arrive<BAR_PANEL, 128>(p & 1);
        TRACE(5 + 6 * p);
        __syncwarp();
        invert_upper<LDW>(Wt + c0 + LDW * c0, npiv, ctl->rinv, s_dinv + 64 * p);   // read by the bottom block
        TRACE(6 + 6 * p);
      }
    } else {
      // ================================================================ update warps: column tiles J > p
      const int uw = warp - 1;                    // 0..2
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
        PanelCtl* ctl = ctl2 + (p & 1);
        bar_sync<BAR_PANEL, 128>(p & 1);
        TRACE(4 + 6 * p);
        // ---- the owner of the next panel's column tile inverts L_pp (on the critical path)
        if (uw == (p + 1) % 3) invert_unit_lower<LDW>(Wt + c0 + LDW * c0, npiv, ctl->Linv);
        const int nd = ctl->ndisp;
        const int ps0 = tig < npiv ? ctl->psrc[tig] : -1;
        const int ps1 = 4 + tig < npiv ? ctl->psrc[4 + tig] : -1;
        const int dsr = gid < nd ? ctl->dsrc[gid] : -1;
        const int dds = gid < nd ? ctl->ddst[gid] : -1;
        bar_sync<BAR_UW_LL, 96>(p & 1);
        TRACE(5 + 6 * p);
        const double li0 = ctl->Linv[gid + 8 * tig], li1 = ctl->Linv[gid + 8 * (4 + tig)];
        const int Jfirst = p + 1 + (uw + 3 - (p + 1) % 3) % 3;   // first owned tile > p (J mod 3 == uw)
#pragma unroll 1
        for (int J = Jfirst; J < CT; J += 3) {
          double* colg = Wt + LDW * (8 * J + nb);      // B-fragment column of this lane
          double* colt = Wt + LDW * (8 * J + tig);     // displaced rows: lane (t=gid) moves columns tig, tig+4
          const double g0 = ps0 >= 0 ? colg[ps0] : 0.0;
          const double g1 = ps1 >= 0 ? colg[ps1] : 0.0;
          const double dv0 = dsr >= 0 ? colt[dsr] : 0.0;
          const double dv1 = dsr >= 0 ? colt[dsr + 4 * LDW] : 0.0;
          __syncwarp();
          double u0 = 0.0, u1 = 0.0;              // U12 tile = inv(L_pp) * gathered rows
          dmma(u0, u1, li0, g0);
          dmma(u0, u1, li1, g1);
          double* cc0 = Wt + gid + LDW * (8 * J + ce0);
          double* cc1 = Wt + gid + LDW * (8 * J + ce1);
          if (c0 + gid < NI) { cc0[c0] = u0; cc1[c0] = u1; }
          if (dds >= 0) { colt[dds] = dv0; colt[dds + 4 * LDW] = dv1; }
          __syncwarp();
          if (p + 1 < RT) {
            const double bf0 = neg(colg[c0 + tig]);
            const double bf1 = neg(colg[c0 + 4 + tig]);
            int I = p + 1;
#pragma unroll 1
            for (; I + 1 < RT; I += 2) {           // two row tiles per step: independent accumulators in flight
              const int rA = 8 * I + gid, rB = rA + 8;
              const bool vA = rA < NI, vB = rB < NI;
              const double aA0 = vA ? Wt[rA + LDW * (c0 + ka0)] : 0.0, aA1 = vA ? Wt[rA + LDW * (c0 + ka1)] : 0.0;
              const double aB0 = vB ? Wt[rB + LDW * (c0 + ka0)] : 0.0, aB1 = vB ? Wt[rB + LDW * (c0 + ka1)] : 0.0;
              double dA0 = vA ? cc0[8 * I] : 0.0, dA1 = vA ? cc1[8 * I] : 0.0;
              double dB0 = vB ? cc0[8 * I + 8] : 0.0, dB1 = vB ? cc1[8 * I + 8] : 0.0;
              dmma(dA0, dA1, aA0, bf0);
              dmma(dB0, dB1, aB0, bf0);
              dmma(dA0, dA1, aA1, bf1);
              dmma(dB0, dB1, aB1, bf1);
              if (vA) { cc0[8 * I] = dA0; cc1[8 * I] = dA1; }
              if (vB) { cc0[8 * I + 8] = dB0; cc1[8 * I + 8] = dB1; }
            }
            if (I < RT) {
              const int r = 8 * I + gid;
              const bool rv = r < NI;
              const double a0 = rv ? Wt[r + LDW * (c0 + ka0)] : 0.0;
              const double a1 = rv ? Wt[r + LDW * (c0 + ka1)] : 0.0;
              double d0 = rv ? cc0[8 * I] : 0.0, d1 = rv ? cc1[8 * I] : 0.0;
              dmma(d0, d1, a0, bf0);
              dmma(d0, d1, a1, bf1);
              if (rv) { cc0[8 * I] = d0; cc1[8 * I] = d1; }
            }
          }
          if (J == p + 1 && p + 1 < NP) bar_arrive<BAR_COL, 64>((p + 1) & 1);   // hand the next panel's tile back
          if (J == Jfirst) TRACE(6 + 6 * p);
        }
        TRACE(7 + 6 * p);
      }
    }
    // ================================================================ bottom block: row tiles, registers only
    TRACE(40);
    __syncthreads();                               // U, L^-1 P [A12 b1] and every inv(U_pp) are final
    TRACE(41);
    {
      const int r0s = sg8(2 * tig), r1s = sg8(2 * tig + 1);   // logical column of accumulator element e / row of k-step s
      const int sgid = sg8(gid);                              // logical column of output column n = gid
      const double* ubase = Wt + LDW * pc(sgid);              // B fragments of U_pJ: ubase[c0 + r_s + LDW*8*J]
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      const double* Arec = A + cell * lenA;
      const double* brec = b + cell * lenb;
      double* Sc = S + cell * (int64_t)NB * NB;
      double* gc = g + cell * (int64_t)NB;
#pragma unroll 1
      for (int I = warp; I < BT; I += 4) {
        const int r = 8 * I + gid;
        const bool rv = r < NB;
        // accumulator fragments straight from the record: the offsets come from a per-plan table laid out per lane
        // (one coalesced look-up per element, immediate offsets), all look-ups first
        const int32_t* xo = tb.xoff + (size_t)I * (CT * 64) + lane;
        int xoffs[CT][2];
#pragma unroll
        for (int J = 0; J < CT; ++J) {
          xoffs[J][0] = __ldg(xo + J * 64);
          xoffs[J][1] = __ldg(xo + J * 64 + 32);
        }
        TRACE(I < 4 ? 42 : 46);
        double x[CT][2];
#pragma unroll
        for (int J = 0; J < CT; ++J) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int o = xoffs[J][e];
            if (J == N / 8) x[J][e] = o >= 0 ? Arec[o] : (o < -1 ? brec[-2 - o] : 0.0);   // the tile of the rhs column
            else x[J][e] = o >= 0 ? Arec[o] : 0.0;
          }
        }
        const bool failed = *s_info != 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int c0 = 8 * p;
          const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
          const double* dv = s_dinv + 64 * p + 8 * sgid;      // zero padded beyond npiv
          double y0 = 0.0, y1 = 0.0;                          // Y_p = -X_p inv(U_pp)
          dmma(y0, y1, x[p][0], neg(dv[r0s]));
          if (npiv > 2) dmma(y0, y1, x[p][1], neg(dv[r1s]));  // k-step 1 stands for rows 2..5
          if (npiv < 8) {
            // partial last panel: the rest of its own tile (columns >= npiv) is updated with U_pp rows < npiv
            const double q0 = (r0s < npiv && sgid >= npiv) ? ubase[c0 + r0s + LDW * c0] : 0.0;
            dmma(x[p][0], x[p][1], y0, q0);
            if (npiv > 2) {
              const double q1 = (r1s < npiv && sgid >= npiv) ? ubase[c0 + r1s + LDW * c0] : 0.0;
              dmma(x[p][0], x[p][1], y1, q1);
            }
          }
          // k-step outermost: consecutive DMMAs belong to independent accumulators
#pragma unroll
          for (int J = p + 1; J < CT; ++J) {
            const double b0 = (npiv == 8 || r0s < npiv) ? ubase[c0 + r0s + LDW * 8 * J] : 0.0;
            dmma(x[J][0], x[J][1], y0, b0);
          }
          if (npiv > 2) {
#pragma unroll
            for (int J = p + 1; J < CT; ++J) {
              const double b1 = (npiv == 8 || r1s < npiv) ? ubase[c0 + r1s + LDW * 8 * J] : 0.0;
              dmma(x[J][0], x[J][1], y1, b1);
            }
          }
        }
        TRACE(I < 4 ? 44 : 48);
        if (rv) {
#pragma unroll
          for (int J = SJ0; J < CT; ++J) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 8 * J + (e ? r1s : r0s);
              const double v = failed ? qnan : x[J][e];
              if (c >= NI) {
                if (c < N) Sc[r + (int64_t)NB * (c - NI)] = v;
                else if (c == N) gc[r] = v;
              }
            }
          }
        }
      }
    }
    TRACE(50);
    if (KEEPX) {
      // ---- keep_factors: X = A11^-1 [A12 | b1] = U^-1 (L^-1 P [A12 b1]), in place in Wt.  The column tiles
      // of the right-hand sides go over the warps; blocked back substitution from the last panel up:
      // X_p = inv(U_pp) W_p, then W_I -= U_Ip X_p for the row tiles above.  Stores are masked to the right-hand-side
      // columns (the tile that straddles column NI also holds columns of U).
      __syncthreads();                             // the bottom block has read L^-1 P [A12 b1]
#pragma unroll 1
      for (int J = SJ0 + warp; J < CT; J += 4) {
        double* colg = Wt + LDW * (8 * J + nb);
        double* cc0 = Wt + gid + LDW * (8 * J + ce0);
        double* cc1 = Wt + gid + LDW * (8 * J + ce1);
        const bool w0 = 8 * J + 2 * tig >= NI, w1 = 8 * J + 2 * tig + 1 >= NI;
#pragma unroll 1
        for (int p = NP - 1; p >= 0; --p) {
          const int c0 = 8 * p;
          const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
          const double* dv = s_dinv + 64 * p;
          const double a0 = dv[gid + 8 * tig], a1 = dv[gid + 8 * (4 + tig)];     // A[i][k] = inv(U_pp)[i][k]
          const double t0 = tig < npiv ? colg[c0 + tig] : 0.0;                    // B[k][n] = W[c0 + k][n]
          const double t1 = 4 + tig < npiv ? colg[c0 + 4 + tig] : 0.0;
          double x0 = 0.0, x1 = 0.0;
          dmma(x0, x1, a0, t0);
          dmma(x0, x1, a1, t1);
          __syncwarp();
          if (gid < npiv) {
            if (w0) cc0[c0] = x0;
            if (w1) cc1[c0] = x1;
          }
          __syncwarp();
          if (p > 0) {
            const double bf0 = tig < npiv ? neg(colg[c0 + tig]) : 0.0;            // -X_p as B fragments
            const double bf1 = 4 + tig < npiv ? neg(colg[c0 + 4 + tig]) : 0.0;
#pragma unroll 1
            for (int I = 0; I < p; ++I) {
              const int r = 8 * I + gid;
              const double u0 = tig < npiv ? Wt[r + LDW * (c0 + ka0)] : 0.0;      // A[i][k] = U[r][c0 + k]
              const double u1 = 4 + tig < npiv ? Wt[r + LDW * (c0 + ka1)] : 0.0;
              double d0 = cc0[8 * I], d1 = cc1[8 * I];
              dmma(d0, d1, u0, bf0);
              dmma(d0, d1, u1, bf1);
              if (w0) cc0[8 * I] = d0;
              if (w1) cc1[8 * I] = d1;
            }
            __syncwarp();
          }
        }
      }
      __syncthreads();
      const bool failedx = *s_info != 0;
      double* Xc = X + cell * (int64_t)(NI * (NB + 1));
      for (int e = tid; e < NI * (NB + 1); e += 128) {
        const int c = e / NI, i = e - c * NI, lc = NI + c;
        Xc[e] = failedx ? __longlong_as_double(0x7ff8000000000000LL) : Wt[i + LDW * (8 * (lc >> 3) + pc(lc & 7))];
      }
    }
    __syncthreads();
    TRACE(51);
    if (info && tid == 0) info[cell] = *s_info;
  }
}

// ---- backward static condensation on the same machinery ------------------------------------------
// u_K = A11^-1 (b1 - A12 lambda_K)   (/root/reference/src/BackwardStaticCondensationMap.jl:84-99; the LU is
// recomputed from the record, as the reference does).  Image W = [A11 | r] (n_i rows, n_i+1 columns): the same
// panel warp and column-tile owners as the condensation kernel factorise it (there is no bottom block), then
// warp 0 back-substitutes with U column by column.  Only A11, A12 and b1 are read from the record:
// B_back = 8(n_i^2 + n_i n_b + 2 n_i) + 16 n_b bytes per cell (ids and gathered lambda included).
template <int NI, int NB>
struct BackCfg {
  using C = Cfg<NI, NB>;
  static constexpr int CTB = (NI + 1 + 7) / 8;          // column tiles of [A11 | r]
  static constexpr int NBP = (NB + 1) & ~1;
  static size_t smem_bytes(int nf) {
    return (size_t)(CTB * 8 * C::LDW + NBP + C::NP * 8) * 8 + 2 * sizeof(PanelCtl) + 16 +
           (size_t)(C::N + 1) * nf * 4 + 2 * NI + 16;
  }
};

template <int NI, int NB, int RPC>
__global__ void __launch_bounds__(128, 8)
backsub_dmma_kernel(DmmaTables tb, int lenA, int lenb, int64_t ncells, const double* __restrict__ A,
                    const double* __restrict__ b, const double* __restrict__ lam_free,
                    const double* __restrict__ lam_dir, const int64_t* __restrict__ ids, double* __restrict__ u,
                    int32_t* __restrict__ info) {
  using C = Cfg<NI, NB>;
  using BC = BackCfg<NI, NB>;
  constexpr int N = C::N, LDW = C::LDW, RT = C::RT, NP = C::NP, CTB = BC::CTB;
  static_assert(RPC == 1 || NI % 2 == 0, "16-byte cp.async needs an even interior height");
  extern __shared__ __align__(16) double smem[];
  double* Wt = smem;                                               // [CTB*8][LDW]
  double* s_lam = Wt + CTB * 8 * LDW;                              // [NB]
  double* s_rinv = s_lam + BC::NBP;                                // [NP*8]
  PanelCtl* ctl2 = reinterpret_cast<PanelCtl*>(s_rinv + NP * 8);   // [2]
  int* s_info = reinterpret_cast<int*>(ctl2 + 2);
  int* s_colbase = s_info + 4;                                     // [(N+1)*nf]
  unsigned short* s_rowinfo = reinterpret_cast<unsigned short*>(s_colbase + (N + 1) * tb.nf);  // [NI/RPC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int ce0 = pc(2 * tig), ce1 = pc(2 * tig + 1);
  const int ka0 = tig, ka1 = pc(4 + tig);
  const int nb = pc(gid);
  constexpr int HP = (NI + RPC - 1) / RPC;      // copy units per column of A11
  constexpr int LG = 128 / HP;                  // column groups
  static_assert(LG >= 1, "cell too tall for the loader");
  constexpr int LB = 6;                         // loader batch: look-ups in flight per thread
  for (int i = tid; i < CTB * 8 * LDW; i += 128) Wt[i] = 0.0;
  for (int i = tid; i < (N + 1) * tb.nf; i += 128) s_colbase[i] = tb.colbase[i];
  for (int i = tid; i < HP; i += 128) s_rowinfo[i] = (unsigned short)((tb.rowf[RPC * i] << 8) | tb.rowl[RPC * i]);
  __syncthreads();
  const int l_grp = tid / HP, l_rp = tid - l_grp * HP;
  const bool l_on = l_grp < LG;
  const int l_ri = l_on ? s_rowinfo[l_rp] : 0;
  const int l_f = l_ri >> 8, l_lr = l_ri & 0xff;
  // the thread that forms r for interior row tid
  const int r_f = tid < NI ? tb.rowf[tid] : 0, r_lr = tid < NI ? tb.rowl[tid] : 0;

  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const double* Arec = A + cell * lenA;
    const double* brec = b + cell * lenb;
    if (l_on) {
      const int* cb = s_colbase + l_f;
#pragma unroll 1
      for (int cb0 = l_grp; cb0 < NI; cb0 += LB * LG) {      // look-ups first, then the copies (see the forward kernel)
        int off[LB];
#pragma unroll
        for (int q = 0; q < LB; ++q) { const int c = cb0 + q * LG; off[q] = c < NI ? cb[c * tb.nf] : -2; }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
          const int c = cb0 + q * LG;
          double* dst = Wt + RPC * l_rp + LDW * pc(c);
          const double* src = Arec + l_lr + (off[q] >= 0 ? off[q] : 0);
          if (off[q] != -2) { if (RPC == 2) cp_async16_z(dst, src, off[q] >= 0); else cp_async8_z(dst, src, off[q] >= 0); }
        }
      }
    }
    // lambda_K through the cell ids (get_cell_dof_values, src/HybridAffineFEOperators.jl:113)
    if (tid < NB) {
      const int64_t id = ids[cell * NB + tid];
      s_lam[tid] = id > 0 ? lam_free[id - 1] : (id < 0 && lam_dir ? lam_dir[-id - 1] : 0.0);
    }
    if (tid == 0) *s_info = 0;
#if GHB_L2PREFETCH
    if (tid == 64 && cell + gridDim.x < ncells) {   // A11 and A12 lead the record; b is small
      l2_prefetch_bulk(A + (cell + gridDim.x) * lenA, (unsigned)(lenA * 8));
      l2_prefetch_bulk(b + (cell + gridDim.x) * lenb, (unsigned)(lenb * 8));
    }
#endif
    __syncthreads();
    // r = b1 - A12 * lambda_K, ascending columns (gemv!('N',-1,A12,x,1,b1)): one interior row per thread, A12 read
    // straight from the record (consecutive threads read consecutive rows of a column)
    if (tid < NI) {
      double r = brec[s_colbase[N * tb.nf + r_f] + r_lr];
#pragma unroll 12
      for (int j = 0; j < NB; ++j) {
        const int off = s_colbase[(NI + j) * tb.nf + r_f];
        const double a = Arec[(off >= 0 ? off : 0) + r_lr];
        r = fma(off >= 0 ? -a : 0.0, s_lam[j], r);
      }
      Wt[tid + LDW * pc(NI)] = r;
    }
    cp_async_commit_wait_all();
    __syncthreads();

    if (warp == 0) {
      // ================================================================ panel warp
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
        if (p > 0) bar_sync<BAR_COL, 64>(p & 1);
        PanelCtl* ctl = ctl2 + (p & 1);
        if ((NI - c0) > 32) panel_factor<NI, LDW, true>(Wt, c0, npiv, ctl, s_info);
        else panel_factor<NI, LDW, false>(Wt, c0, npiv, ctl, s_info);
        __syncwarp();
        if (lane < 8) s_rinv[c0 + lane] = ctl->rinv[lane];
        bar_arrive<BAR_PANEL, 128>(p & 1);
        invert_unit_lower<LDW>(Wt + c0 + LDW * c0, npiv, ctl->Linv);
        bar_arrive<BAR_UW_BACK, 128>(p & 1);
      }
      // ---- every column tile is final: back substitution U x = y, rows lane and lane+32 in registers
      bar_sync<BAR_UDONE, 128>(0);
      const double* ycol = Wt + LDW * pc(NI);
      double y1 = lane < NI ? ycol[lane] : 0.0;
      double y2 = lane + 32 < NI ? ycol[lane + 32] : 0.0;
#pragma unroll 1
      for (int k = NI - 1; k >= 0; --k) {
        const double* uk = Wt + LDW * pc(k);
        const double u1k = uk[lane];
        const double u2k = lane + 32 < NI ? uk[lane + 32] : 0.0;
        const double yk = __shfl_sync(0xffffffffu, k >= 32 ? y2 : y1, k & 31);
        const double xk = yk * s_rinv[k];
        if (lane == (k & 31)) { if (k >= 32) y2 = xk; else y1 = xk; }
        if (lane < k) y1 = fma(-u1k, xk, y1);
        if (NI > 32 && lane + 32 < k) y2 = fma(-u2k, xk, y2);
      }
      const bool failed = *s_info != 0;
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      double* uc = u + cell * (int64_t)NI;
      if (lane < NI) uc[lane] = failed ? qnan : y1;
      if (lane + 32 < NI) uc[lane + 32] = failed ? qnan : y2;
    } else {
      // ================================================================ update warps: column tiles J > p
      const int uw = warp - 1;
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int c0 = 8 * p;
        const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
        PanelCtl* ctl = ctl2 + (p & 1);
        bar_sync<BAR_PANEL, 128>(p & 1);
        const int nd = ctl->ndisp;
        const int ps0 = tig < npiv ? ctl->psrc[tig] : -1;
        const int ps1 = 4 + tig < npiv ? ctl->psrc[4 + tig] : -1;
        const int dsr = gid < nd ? ctl->dsrc[gid] : -1;
        const int dds = gid < nd ? ctl->ddst[gid] : -1;
        bar_sync<BAR_UW_BACK, 128>(p & 1);          // inv(L_pp) from the panel warp
        const double li0 = ctl->Linv[gid + 8 * tig], li1 = ctl->Linv[gid + 8 * (4 + tig)];
        const int Jfirst = p + 1 + (uw + 3 - (p + 1) % 3) % 3;
#pragma unroll 1
        for (int J = Jfirst; J < CTB; J += 3) {
          double* colg = Wt + LDW * (8 * J + nb);
          double* colt = Wt + LDW * (8 * J + tig);
          const double g0 = ps0 >= 0 ? colg[ps0] : 0.0;
          const double g1 = ps1 >= 0 ? colg[ps1] : 0.0;
          const double dv0 = dsr >= 0 ? colt[dsr] : 0.0;
          const double dv1 = dsr >= 0 ? colt[dsr + 4 * LDW] : 0.0;
          __syncwarp();
          double u0 = 0.0, u1 = 0.0;              // U12 tile = inv(L_pp) * gathered rows
          dmma(u0, u1, li0, g0);
          dmma(u0, u1, li1, g1);
          double* cc0 = Wt + gid + LDW * (8 * J + ce0);
          double* cc1 = Wt + gid + LDW * (8 * J + ce1);
          if (c0 + gid < NI) { cc0[c0] = u0; cc1[c0] = u1; }
          if (dds >= 0) { colt[dds] = dv0; colt[dds + 4 * LDW] = dv1; }
          __syncwarp();
          if (p + 1 < RT) {
            const double bf0 = neg(colg[c0 + tig]);
            const double bf1 = neg(colg[c0 + 4 + tig]);
#pragma unroll 1
            for (int I = p + 1; I < RT; ++I) {
              const int r = 8 * I + gid;
              const bool rv = r < NI;
              const double a0 = rv ? Wt[r + LDW * (c0 + ka0)] : 0.0;
              const double a1 = rv ? Wt[r + LDW * (c0 + ka1)] : 0.0;
              double d0 = rv ? cc0[8 * I] : 0.0, d1 = rv ? cc1[8 * I] : 0.0;
              dmma(d0, d1, a0, bf0);
              dmma(d0, d1, a1, bf1);
              if (rv) { cc0[8 * I] = d0; cc1[8 * I] = d1; }
            }
          }
          if (J == p + 1 && p + 1 < NP) bar_arrive<BAR_COL, 64>((p + 1) & 1);
        }
      }
      bar_arrive<BAR_UDONE, 128>(0);
    }
    __syncthreads();
    if (info && tid == 0) info[cell] = *s_info;
  }
}

}  // namespace

// host side -----------------------------------------------------------------------------------------
// every vertical pair (2q, 2q+1) of the condensed matrix contiguous and 16-byte aligned in the packed record?
static bool pairs_aligned(const Plan& p) {
  if (p.lenA % 2 || p.lenb % 2 || p.n % 2 || p.n_i % 2) return false;
  for (int r = 0; r + 1 < p.n; r += 2)
    if (p.row_field[r] != p.row_field[r + 1] || p.row_local[r] + 1 != p.row_local[r + 1] || p.row_local[r] % 2) return false;
  for (int f = 0; f < p.nfields; ++f) {
    if (p.ndofs[f] % 2 || p.field_offset_b[f] % 2) return false;
    for (int q = 0; q < p.nfields; ++q)
      if (p.block_offset[f + p.nfields * q] >= 0 && p.block_offset[f + p.nfields * q] % 2) return false;
  }
  return true;
}

// shapes with a DMMA instantiation: 3-D HDG k=2 (34,36) and RT-H k=2, k=3 on quads (33,12), (56,16)
bool dmma_supported(const Plan& p) {
  if (p.nfields > 8) return false;
  for (int r = 0; r < p.n; ++r) if (p.row_local[r] > 255) return false;
  if (p.n_i == 34 && p.n_b == 36) return pairs_aligned(p);
  if (p.n_i == 56 && p.n_b == 16) return pairs_aligned(p);
  if (p.n_i == 33 && p.n_b == 12) return true;            // 8-byte copies
  return false;
}

int dmma_prepare(ghb_ctx* ctx, Plan& p) {
  // colbase[(c)*nf + fslot]: fslot enumerates fields in original order (0-based field id)
  const int n = p.n, nf = p.nfields;
  std::vector<int32_t> colbase((size_t)(n + 1) * nf, -1);
  for (int c = 0; c < n; ++c) {
    const int fc = p.row_field[c], lc = p.row_local[c];
    for (int f = 0; f < nf; ++f) {
      int64_t bo = p.block_offset[f + nf * fc];
      if (bo >= 0) colbase[(size_t)c * nf + f] = (int32_t)(bo + (int64_t)lc * p.ndofs[f]);
    }
  }
  for (int f = 0; f < nf; ++f) colbase[(size_t)n * nf + f] = p.field_offset_b[f];
  std::vector<uint8_t> rowf(n), rowl(n);
  for (int r = 0; r < n; ++r) { rowf[r] = (uint8_t)p.row_field[r]; rowl[r] = (uint8_t)p.row_local[r]; }
  GHB_CUDA(ctx, cudaMalloc((void**)&p.d_colbase, colbase.size() * sizeof(int32_t)));
  GHB_CUDA(ctx, cudaMalloc((void**)&p.d_rowf, 2 * n));
  p.d_rowl = p.d_rowf + n;
  GHB_CUDA(ctx, cudaMemcpy(p.d_colbase, colbase.data(), colbase.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  GHB_CUDA(ctx, cudaMemcpy(p.d_rowf, rowf.data(), n, cudaMemcpyHostToDevice));
  GHB_CUDA(ctx, cudaMemcpy(p.d_rowl, rowl.data(), n, cudaMemcpyHostToDevice));
  {
    // left-looking kernel: record offsets of the bottom block [A21 A22 b2] in accumulator-fragment order
    const int ni = p.n_i, nbd = p.n_b, BT = (nbd + 7) / 8, CT = (n + 1 + 7) / 8;
    static const int sg[8] = {0, 2, 1, 3, 6, 4, 7, 5};
    std::vector<int32_t> xoff((size_t)BT * CT * 64, -1);
    for (int I = 0; I < BT; ++I)
      for (int J = 0; J < CT; ++J)
        for (int e = 0; e < 2; ++e)
          for (int lane = 0; lane < 32; ++lane) {
            const int r = 8 * I + (lane >> 2), c = 8 * J + sg[2 * (lane & 3) + e];
            if (r >= nbd || c > n) continue;
            const int f = p.row_field[ni + r], lr = p.row_local[ni + r];
            const int32_t base = colbase[(size_t)c * nf + f];
            if (base < 0) continue;
            xoff[((size_t)(I * CT + J) * 2 + e) * 32 + lane] = c < n ? base + lr : -2 - (base + lr);
          }
    GHB_CUDA(ctx, cudaMalloc((void**)&p.d_xoff, xoff.size() * sizeof(int32_t)));
    GHB_CUDA(ctx, cudaMemcpy(p.d_xoff, xoff.data(), xoff.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  p.kernel_name = p.n_i == 34 ? "dmma_34_36" : (p.n_i == 33 ? "dmma_33_12" : "dmma_56_16");
  return GHB_OK;
}

template <int NI, int NB, int RPC>
static int launch_dmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                       double* g, int32_t* info) {
  using C = Cfg<NI, NB>;
  auto kern = condense_dmma_kernel<NI, NB, RPC>;
  const size_t smem = C::smem_bytes(p.nfields);
  static KernelSetup ks;
  int per_sm = 0;
  GHB_TRY(kernel_setup(ctx, p.opt, kern, 128, smem, true, ks, "condense_dmma_kernel", &per_sm));
  DmmaTables tb{p.d_colbase, p.d_rowf, p.d_rowl, p.nfields, p.d_xoff};
  int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * per_sm);
  kern<<<(unsigned)grid, 128, smem, ctx->stream>>>(tb, p.lenA, p.lenb, ncells, A, b, S, g, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

// resident CTAs per SM the left-looking kernels are compiled for (64 / 72 / 80 / 96 registers at 8 / 7 / 6 / 5)
#ifndef GHB_LL_MINB34
#define GHB_LL_MINB34 8
#endif
#ifndef GHB_LL_MINB33
#define GHB_LL_MINB33 8
#endif
#ifndef GHB_LL_MINB56
#define GHB_LL_MINB56 5
#endif

template <int NI, int NB, int RPC, int MINB, bool KEEPX = false>
static int launch_dmma_ll(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                          double* g, int32_t* info, double* X = nullptr) {
  auto kern = condense_dmma_ll_kernel<NI, NB, RPC, MINB, KEEPX>;
  const size_t smem = LLCfg<NI, NB>::smem_bytes(p.nfields);
  static KernelSetup ks;
  int per_sm = 0;
  GHB_TRY(kernel_setup(ctx, p.opt, kern, 128, smem, true, ks, "condense_dmma_ll_kernel", &per_sm));
  DmmaTables tb{p.d_colbase, p.d_rowf, p.d_rowl, p.nfields, p.d_xoff};
  int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * per_sm);
  kern<<<(unsigned)grid, 128, smem, ctx->stream>>>(tb, p.lenA, p.lenb, ncells, A, b, S, g, info, X);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

// kernel selection of the DMMA shapes: the left-looking kernel is faster on all three (profiles/r01_condense_ll.md);
// GHB_DMMA_LL=0 forces the right-looking kernel (A/B runs, tools/ab_ll.py)
static bool use_ll(const Plan& p) {
  return p.opt.dmma_ll != 0;
}

template <int NI, int NB, int RPC>
static int launch_bdmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                        const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info) {
  auto kern = backsub_dmma_kernel<NI, NB, RPC>;
  const size_t smem = BackCfg<NI, NB>::smem_bytes(p.nfields);
  static KernelSetup ks;
  int per_sm = 0;
  GHB_TRY(kernel_setup(ctx, p.opt, kern, 128, smem, true, ks, "backsub_dmma_kernel", &per_sm));
  DmmaTables tb{p.d_colbase, p.d_rowf, p.d_rowl, p.nfields, p.d_xoff};
  int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * per_sm);
  kern<<<(unsigned)grid, 128, smem, ctx->stream>>>(tb, p.lenA, p.lenb, ncells, A, b, lam_free, lam_dir, ids, u, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int launch_backsub_dmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b,
                        const double* lam_free, const double* lam_dir, const int64_t* ids, double* u, int32_t* info) {
  if (p.n_i == 34) return launch_bdmma<34, 36, 2>(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  if (p.n_i == 56) return launch_bdmma<56, 16, 2>(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  return launch_bdmma<33, 12, 1>(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
}

int launch_condense_dmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                         double* g, int32_t* info, double* X) {
  if (X) {   // keep_factors: the left-looking kernel with the back substitution X = U^-1 (L^-1 P [A12 b1]) appended
    if (p.n_i == 34) return launch_dmma_ll<34, 36, 2, GHB_LL_MINB34, true>(ctx, p, ncells, A, b, S, g, info, X);
    if (p.n_i == 56) return launch_dmma_ll<56, 16, 2, GHB_LL_MINB56, true>(ctx, p, ncells, A, b, S, g, info, X);
    return launch_dmma_ll<33, 12, 1, GHB_LL_MINB33, true>(ctx, p, ncells, A, b, S, g, info, X);
  }
  if (use_ll(p)) {
    if (p.n_i == 34) {
      const int m = p.opt.ll_ctas > 0 ? p.opt.ll_ctas : GHB_LL_MINB34;   // tuning knob: register budget of this instantiation
      if (m == 6) return launch_dmma_ll<34, 36, 2, 6>(ctx, p, ncells, A, b, S, g, info);
      if (m == 7) return launch_dmma_ll<34, 36, 2, 7>(ctx, p, ncells, A, b, S, g, info);
      return launch_dmma_ll<34, 36, 2, 8>(ctx, p, ncells, A, b, S, g, info);
    }
    if (p.n_i == 56) return launch_dmma_ll<56, 16, 2, GHB_LL_MINB56>(ctx, p, ncells, A, b, S, g, info);
    return launch_dmma_ll<33, 12, 1, GHB_LL_MINB33>(ctx, p, ncells, A, b, S, g, info);
  }
  if (p.n_i == 34) return launch_dmma<34, 36, 2>(ctx, p, ncells, A, b, S, g, info);
  if (p.n_i == 56) return launch_dmma<56, 16, 2>(ctx, p, ncells, A, b, S, g, info);
  return launch_dmma<33, 12, 1>(ctx, p, ncells, A, b, S, g, info);
}

}  // namespace ghb
