// condense_dmma.cu -- tuned static condensation for mid-size cells (3-D HDG k=2: n_i=34, n_b=36) on
// FP64 DMMA tensor cores (mma.sync.m8n8k4.f64; measured full-rate on B200: 37.1 TFLOP/s, see
// profiles/r01_ubench_fp64.txt).
//
// One CTA (4 warps) per cell, 5 CTAs per SM.  Replaces evaluate!(cache, ::StaticCondensationMap, A, b)
// (/root/reference/src/StaticCondensationMap.jl:152-196):
//   * the packed record is re-laid out on the fly by 16-byte cp.async into two dense column-major
//     images in shared memory: Wt = [A11 A12 b1] (n_i rows) and Bt = [A21 A22 b2] (n_b rows);
//   * phase 1 (getrf! :179 + the L-solve half of getrs! :183,:189): blocked right-looking LU of Wt with
//     partial pivoting, panel width 8.  The panel is factorised by one warp with one row per lane
//     (pivot = first max |.| like idamax, found with two REDUX.MAX on the 64-bit magnitude), row
//     interchanges and the unit-lower solve run one column per lane, the trailing update is DMMA;
//   * phase 2 (the U-solve half of getrs!, gemm! :186, gemv! :192 re-associated as
//     S = A22 - (A21 U^-1)(L^-1 P A12)): left-looking sweep over the column tiles of Bt with the
//     accumulators in registers: L21 = (A21 - L21 U) * inv(U_pp) and S = A22 - L21 * U12, all DMMA.
// The pivot sequence is LAPACK's; only the association/rounding of the updates differs (parity bar 1e-11).
#include <algorithm>

#include "common.cuh"

namespace ghb {

namespace {

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ double neg(double x) { return -x; }

// tables of a plan for the on-the-fly re-layout (built on the host, tiny)
struct DmmaTables {
  const int32_t* colbase;  // [(n+1)*nf]: record offset of (first row of field f, column c) or -1; c==n: offset in b
  const uint8_t* rowf;     // [n]: field slot of condensed row r
  const uint8_t* rowl;     // [n]: row inside its field
  int nf;
};

template <int NI, int NB>
struct Cfg {
  static constexpr int N = NI + NB;
  static constexpr int NC = N + 1;             // + rhs column
  static constexpr int RT = (NI + 7) / 8;      // row tiles of the top block
  static constexpr int BT = (NB + 7) / 8;      // row tiles of the bottom block
  static constexpr int CT = (NC + 7) / 8;      // column tiles
  static constexpr int NP = (NI + 7) / 8;      // panels
  static constexpr int LDW = 36;               // leading dims: 2*LD = 8 (mod 32) words -> conflict-free fragments
  static constexpr int LDB = 36;
  static constexpr int NCP = CT * 8;
  static_assert(NI <= LDW && NB <= LDB, "leading dimension too small");
  static_assert(NI % 2 == 0 && NB % 2 == 0, "16-byte cp.async needs even block heights");
  static constexpr int WT_DOUBLES = NCP * LDW;
  static constexpr int BT_DOUBLES = NCP * LDB;
  static size_t smem_bytes(int nf) {   // arrays + Dinv + rinv + control block + re-layout tables
    return (size_t)(WT_DOUBLES + BT_DOUBLES + NP * 64 + 8) * 8 + 512 + (size_t)(N + 1) * nf * 4 + N + 64;
  }
};

struct Shared {   // sizeof <= 512
  // small control block placed after the big arrays
  int info;
  int ndisp;             // displaced rows of the current panel
  int psrc[8];           // physical row (before this panel's permutation) of the k-th pivot row
  int dsrc[8];           // displaced rows: old position (inside the diagonal block) ...
  int ddst[8];           // ... and the vacated position they move to
  int vac[40];           // scratch: vacated positions by rank
};

// ---- panel factorisation: one warp, one row per lane, implicit pivoting --------------------------
// Factorises columns [c0, c0+npiv) of Wt over rows [c0, NI) and carries the other columns of the 8-wide
// column tile through the eliminations.  Rows are not exchanged while factorising: a lane keeps its row
// and remembers at which step it was chosen.  On write-back the k-th pivot row goes to position c0+k and
// the rows it displaces from the diagonal block go to the vacated positions (any consistent row order is
// a valid row-permuted LU; S does not depend on it).  Publishes that permutation for the trailing columns.
template <int NI, int LDW, bool TWO>
__device__ __forceinline__ void panel_factor(double* __restrict__ Wt, const int c0, const int npiv, Shared* sh,
                                             double* __restrict__ rinv_out) {
  const int lane = threadIdx.x & 31;
  const int nrows = NI - c0;
  const bool v1 = lane < nrows;
  const bool v2 = TWO && (lane + 32 < nrows);
  double a[8], a2[8];
  int ch1 = -1, ch2 = -1;                       // step at which this lane's row became the pivot row
  double* base = Wt + c0 + lane + LDW * c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = v1 ? base[LDW * j] : 0.0;
    a2[j] = v2 ? base[32 + LDW * j] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < npiv) {
      // ---- pivot search: max |a[k]| over the rows not chosen yet (lowest lane wins ties)
      const bool c1 = v1 && ch1 < 0;
      const bool c2 = v2 && ch2 < 0;
      const unsigned long long key1 = c1 ? ((unsigned long long)__double_as_longlong(a[k]) & 0x7fffffffffffffffull) : 0ull;
      const unsigned long long key2 = c2 ? ((unsigned long long)__double_as_longlong(a2[k]) & 0x7fffffffffffffffull) : 0ull;
      const double rc1 = __drcp_rn(a[k]);       // speculative reciprocal, overlaps the reduction
      const double rc2 = TWO ? __drcp_rn(a2[k]) : 0.0;
      const unsigned long long km = TWO ? (key1 > key2 ? key1 : key2) : key1;
      const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(km >> 32));
      const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(km >> 32) == hi ? (unsigned)km : 0u);
      const unsigned long long kmax = ((unsigned long long)hi << 32) | lo;
      if (kmax == 0ull) {                       // exact zero pivot column: LAPACK info = k+1 (uniform branch)
        if (lane == 0 && sh->info == 0) sh->info = c0 + k + 1;
        break;
      }
      const unsigned b1 = __ballot_sync(0xffffffffu, c1 && key1 == kmax);
      const unsigned b2 = TWO ? __ballot_sync(0xffffffffu, c2 && key2 == kmax) : 0u;
      const bool from2 = TWO && (b1 == 0u);
      const int q = from2 ? (__ffs(b2) - 1) : (__ffs(b1) - 1);
      const double rinv = __shfl_sync(0xffffffffu, from2 ? rc2 : rc1, q);
      if (lane == 0) rinv_out[k] = rinv;
      const bool me1 = !from2 && lane == q, me2 = from2 && lane == q;
      if (me1) ch1 = k;
      if (me2) ch2 = k;
      // ---- multipliers and rank-1 update of the rows still in play
      const bool u1 = c1 && !me1, u2 = TWO && c2 && !me2;
      const double l1 = a[k] * rinv, l2 = a2[k] * rinv;   // dgetf2: scale by the reciprocal
      if (u1) a[k] = l1;
      if (u2) a2[k] = l2;
#pragma unroll
      for (int j = k + 1; j < 8; ++j) {
        const double pj = __shfl_sync(0xffffffffu, from2 ? a2[j] : a[j], q);   // pivot row entry, to everyone
        if (u1) a[j] = fma(-l1, pj, a[j]);
        if (u2) a2[j] = fma(-l2, pj, a2[j]);
      }
    }
  }
  // ---- new positions: pivot rows first, displaced rows into the vacated slots
  const bool disp = v1 && lane < 8 && ch1 < 0 && lane < npiv + 0 * nrows;   // rows of the diagonal block not chosen
  const bool vc1 = v1 && lane >= npiv && ch1 >= 0;
  const bool vc2 = v2 && ch2 >= 0;
  const unsigned mdisp = __ballot_sync(0xffffffffu, disp);
  const unsigned mv1 = __ballot_sync(0xffffffffu, vc1);
  const unsigned mv2 = __ballot_sync(0xffffffffu, vc2);
  const unsigned lt = (1u << lane) - 1u;
  if (vc1) sh->vac[__popc(mv1 & lt)] = c0 + lane;
  if (vc2) sh->vac[__popc(mv1) + __popc(mv2 & lt)] = c0 + 32 + lane;
  __syncwarp();
  int np1 = c0 + lane, np2 = c0 + 32 + lane;
  if (ch1 >= 0) np1 = c0 + ch1;
  else if (disp) np1 = sh->vac[__popc(mdisp & lt)];
  if (ch2 >= 0) np2 = c0 + ch2;
  if (ch1 >= 0) sh->psrc[ch1] = c0 + lane;
  if (ch2 >= 0) sh->psrc[ch2] = c0 + 32 + lane;
  if (disp) {
    const int t = __popc(mdisp & lt);
    sh->dsrc[t] = c0 + lane;
    sh->ddst[t] = np1;
  }
  if (lane == 0) sh->ndisp = __popc(mdisp);
  double* wb = Wt + LDW * c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (v1) wb[np1 + LDW * j] = a[j];
    if (v2) wb[np2 + LDW * j] = a2[j];
  }
}

// inverse of the npiv x npiv upper-triangular diagonal block, stored as the 8x8 B-operand
// Dinv[k + 8*n] (column-major), zero outside the npiv block.  Lane n < 8 computes column n.
template <int LDW>
__device__ __forceinline__ void invert_upper(const double* __restrict__ Wt, const int c0, const int npiv,
                                             const double* __restrict__ rinv, double* __restrict__ Dinv) {
  const int n = threadIdx.x & 31;
  if (n >= 8) return;
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.0;
  const double* U = Wt + c0 + LDW * c0;
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (i < npiv && i <= n && n < npiv) {
      double s = (i == n) ? 1.0 : 0.0;
#pragma unroll
      for (int m = i + 1; m < 8; ++m)
        if (m <= n) s = fma(-U[i + LDW * m], x[m], s);
      x[i] = s * rinv[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) Dinv[i + 8 * n] = x[i];
}

template <int NI, int NB>
__global__ void __launch_bounds__(128, 5)
condense_dmma_kernel(DmmaTables tb, int lenA, int lenb, int64_t ncells, const double* __restrict__ A,
                     const double* __restrict__ b, double* __restrict__ S, double* __restrict__ g,
                     int32_t* __restrict__ info) {
  using C = Cfg<NI, NB>;
  constexpr int N = C::N, NC = C::NC, LDW = C::LDW, LDB = C::LDB, RT = C::RT, BT = C::BT, CT = C::CT, NP = C::NP;
  constexpr int NF = 8;                          // max fields handled by the smem tables
  extern __shared__ __align__(16) double smem[];
  double* Wt = smem;
  double* Bt = Wt + C::WT_DOUBLES;
  double* Dinv = Bt + C::BT_DOUBLES;            // [NP][64]
  double* rinv = Dinv + NP * 64;                // [8]
  Shared* sh = reinterpret_cast<Shared*>(rinv + 8);
  int* s_colbase = reinterpret_cast<int*>(sh + 1);           // [(N+1)*nf]
  unsigned short* s_rowinfo = reinterpret_cast<unsigned short*>(s_colbase + (N + 1) * tb.nf);  // [N/2]: f<<8 | local row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;    // fragment coordinates
  (void)NF;

  // padding never written by the loader: zero it once (rows NI.. of Wt, rows NB.. of Bt, column NC)
  for (int i = tid; i < C::WT_DOUBLES; i += 128) Wt[i] = 0.0;
  for (int i = tid; i < C::BT_DOUBLES; i += 128) Bt[i] = 0.0;
  for (int i = tid; i < (N + 1) * tb.nf; i += 128) s_colbase[i] = tb.colbase[i];
  for (int i = tid; i < N / 2; i += 128) s_rowinfo[i] = (unsigned short)((tb.rowf[2 * i] << 8) | tb.rowl[2 * i]);
  __syncthreads();

  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    // ------------------------------------------------------------------ load + re-layout
    {
      const double* Arec = A + cell * lenA;
      const double* brec = b + cell * lenb;
      constexpr int HP = N / 2;  // row pairs per column
      // idx = tid + 128*m  ->  (c, rp) advanced incrementally (no division)
      int c = tid / HP, rp = tid - c * HP;
      constexpr int DC = 128 / HP, DR = 128 - DC * HP;
      for (; c < NC;) {
        const int ri = s_rowinfo[rp];
        const int off = s_colbase[c * tb.nf + (ri >> 8)];
        const int r = 2 * rp;
        double* dst = r < NI ? Wt + r + LDW * c : Bt + (r - NI) + LDB * c;
        if (off >= 0) {
          cp_async16(dst, (c < N ? Arec : brec) + off + (ri & 0xff));
        } else {
          dst[0] = 0.0; dst[1] = 0.0;
        }
        c += DC; rp += DR;
        if (rp >= HP) { rp -= HP; ++c; }
      }
      if (tid == 0) sh->info = 0;
      cp_async_commit_wait_all();
      __syncthreads();
    }

    // ------------------------------------------------------------------ phase 1: blocked LU of Wt
    bool ok = true;
#pragma unroll 1
    for (int p = 0; p < NP; ++p) {
      const int c0 = 8 * p;
      const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
      if (warp == 0) {
        if ((NI - c0) > 32) panel_factor<NI, LDW, true>(Wt, c0, npiv, sh, rinv);
        else panel_factor<NI, LDW, false>(Wt, c0, npiv, sh, rinv);
      }
      __syncthreads();
      if (sh->info != 0) { ok = false; break; }
      // ---- warp 3 inverts the diagonal block; warps 0-2: one trailing column per thread: gather the pivot
      //      rows, move the displaced rows, unit-lower solve
      if (warp == 3) {
        invert_upper<LDW>(Wt, c0, npiv, rinv, Dinv + p * 64);
      } else {
        const int c = c0 + 8 + tid;             // trailing column of this thread
        if (c < NC) {
          double* col = Wt + LDW * c;
          double u[8], dv[8];
          const int nd = sh->ndisp;
#pragma unroll
          for (int k = 0; k < 8; ++k) u[k] = k < npiv ? col[sh->psrc[k]] : 0.0;
#pragma unroll
          for (int t = 0; t < 8; ++t) dv[t] = t < nd ? col[sh->dsrc[t]] : 0.0;
          const double* Lp = Wt + c0 + LDW * c0;
#pragma unroll
          for (int j = 0; j < 7; ++j) {
#pragma unroll
            for (int i = j + 1; i < 8; ++i)
              if (i < npiv) u[i] = fma(-Lp[i + LDW * j], u[j], u[i]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < npiv) col[c0 + k] = u[k];
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < nd) col[sh->ddst[t]] = dv[t];
        }
      }
      __syncthreads();
      // ---- trailing update of the top block: C[I][J] -= L[I][p] * U[p][J],  I > p, J > p
      if (p + 1 < RT) {
        for (int J = p + 1 + warp; J < CT; J += 4) {
          const double bf0 = neg(Wt[c0 + tig + LDW * (8 * J + gid)]);
          const double bf1 = neg(Wt[c0 + 4 + tig + LDW * (8 * J + gid)]);
          for (int I = p + 1; I < RT; ++I) {
            const int r = 8 * I + gid;
            const bool rv = r < NI;
            const double a0 = rv ? Wt[r + LDW * (c0 + tig)] : 0.0;
            const double a1 = rv ? Wt[r + LDW * (c0 + 4 + tig)] : 0.0;
            double* cp0 = Wt + r + LDW * (8 * J + 2 * tig);
            double d0 = rv ? cp0[0] : 0.0, d1 = rv ? cp0[LDW] : 0.0;
            dmma(d0, d1, a0, bf0);
            dmma(d0, d1, a1, bf1);
            if (rv) { cp0[0] = d0; cp0[LDW] = d1; }
          }
        }
        __syncthreads();
      }
    }

    // ------------------------------------------------------------------ phase 2: bottom block
    double* Sc = S + cell * (int64_t)NB * NB;
    double* gc = g + cell * (int64_t)NB;
    if (ok) {
      for (int I = warp; I < BT; I += 4) {
        const int r = 8 * I + gid;               // row inside Bt
        const bool rv = r < NB;
        double la[2 * NP];                        // A fragments of L21[I][K], K < NP (two k-steps each)
#pragma unroll
        for (int J = 0; J < CT; ++J) {
          double* cp0 = Bt + r + LDB * (8 * J + 2 * tig);
          double x0 = rv ? cp0[0] : 0.0, x1 = rv ? cp0[LDB] : 0.0;
#pragma unroll
          for (int K = 0; K < NP; ++K) {
            if (K < J) {
              dmma(x0, x1, la[2 * K], neg(Wt[8 * K + tig + LDW * (8 * J + gid)]));
              if (8 * K + 4 < NI) dmma(x0, x1, la[2 * K + 1], neg(Wt[8 * K + 4 + tig + LDW * (8 * J + gid)]));
            }
          }
          if (J < NP) {
            // L = X * Dinv_J  (C fragment -> A fragment through shared memory: the tile is private to the warp)
            const int npv = (NI - 8 * J) < 8 ? (NI - 8 * J) : 8;
            if (rv) { cp0[0] = x0; cp0[LDB] = x1; }
            __syncwarp();
            double xa0 = rv ? Bt[r + LDB * (8 * J + tig)] : 0.0;
            double xa1 = rv ? Bt[r + LDB * (8 * J + 4 + tig)] : 0.0;
            double l0 = 0.0, l1 = 0.0;
            const double* Dj = Dinv + J * 64;
            dmma(l0, l1, xa0, Dj[tig + 8 * gid]);
            if (npv > 4) dmma(l0, l1, xa1, Dj[4 + tig + 8 * gid]);
            if (npv < 8) {
              // partial panel: columns >= npv of this tile are trailing columns: X -= L * U_pp[:, npv..]
              // (L is zero outside its first npv columns because Dinv is)
              __syncwarp();
              if (rv) { cp0[0] = l0; cp0[LDB] = l1; }   // park L to read it back as an A fragment
              __syncwarp();
              double lt0 = rv ? Bt[r + LDB * (8 * J + tig)] : 0.0;
              // B operand: rows < npv of the U tile, zero for columns < npv (those hold L\U of the panel)
              const int kk = tig, nn = gid;
              double ub = (kk < npv && nn >= npv) ? neg(Wt[8 * J + kk + LDW * (8 * J + nn)]) : 0.0;
              dmma(x0, x1, lt0, ub);
              la[2 * J] = lt0;
              la[2 * J + 1] = 0.0;
              // final S values of this tile's trailing columns
              __syncwarp();
              if (rv) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int c = 8 * J + 2 * tig + e;
                  const double v = e ? x1 : x0;
                  if (c >= NI) {
                    if (c < N) Sc[r + (int64_t)NB * (c - NI)] = v;
                    else if (c == N) gc[r] = v;
                  }
                }
              }
            } else {
              __syncwarp();
              if (rv) { cp0[0] = l0; cp0[LDB] = l1; }
              __syncwarp();
              la[2 * J] = rv ? Bt[r + LDB * (8 * J + tig)] : 0.0;
              la[2 * J + 1] = rv ? Bt[r + LDB * (8 * J + 4 + tig)] : 0.0;
            }
          } else if (rv) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 8 * J + 2 * tig + e;
              const double v = e ? x1 : x0;
              if (c < N) Sc[r + (int64_t)NB * (c - NI)] = v;
              else if (c == N) gc[r] = v;
            }
          }
        }
      }
    } else {
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = tid; i < NB * NB; i += 128) Sc[i] = qnan;
      for (int i = tid; i < NB; i += 128) gc[i] = qnan;
    }
    if (info && tid == 0) info[cell] = sh->info;
    __syncthreads();
  }
}

}  // namespace

// host side -----------------------------------------------------------------------------------------
bool dmma_supported(const Plan& p) {
  if (!(p.n_i == 34 && p.n_b == 36)) return false;
  if (p.lenA % 2 || p.lenb % 2) return false;
  // every vertical pair (2q, 2q+1) of the condensed matrix must be contiguous and 16-byte aligned in the record
  const int n = p.n;
  for (int r = 0; r < n; r += 2) {
    if (p.row_field[r] != p.row_field[r + 1] || p.row_local[r] + 1 != p.row_local[r + 1] || p.row_local[r] % 2) return false;
  }
  for (int f = 0; f < p.nfields; ++f) {
    if (p.ndofs[f] % 2) return false;
    if (p.field_offset_b[f] % 2) return false;
    for (int q = 0; q < p.nfields; ++q)
      if (p.block_offset[f + p.nfields * q] >= 0 && p.block_offset[f + p.nfields * q] % 2) return false;
  }
  return p.nfields <= 8;
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
  p.kernel_name = "dmma_34_36";
  return GHB_OK;
}

int launch_condense_dmma(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                         double* g, int32_t* info) {
  using C = Cfg<34, 36>;
  auto kern = condense_dmma_kernel<34, 36>;
  const size_t smem = C::smem_bytes(p.nfields);
  GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  DmmaTables tb{p.d_colbase, p.d_rowf, p.d_rowl, p.nfields};
  int64_t grid = std::min<int64_t>(ncells, (int64_t)ctx->sm_count * 5);
  kern<<<(unsigned)grid, 128, smem, ctx->stream>>>(tb, p.lenA, p.lenb, ncells, A, b, S, g, info);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

}  // namespace ghb
