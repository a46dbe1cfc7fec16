// condense_cw.cu -- static condensation with ONE WARP PER CELL and no barriers ("cell-warp" kernel), FP64 DMMA.
// Replaces evaluate!(cache, ::StaticCondensationMap, A, b) (/root/reference/src/StaticCondensationMap.jl:152-196)
// for mid-size cells (n_i <= 64 interior dofs, one skeleton field).
//
// Why: the 4-warps-per-cell kernels of condense_dmma.cu spend most of their time in named barriers between the panel
// warp and the column-tile owners (profiles/r01_condense_dmma_summary.md: 7.7 barrier stalls per issue, 12.5 k
// warp-instructions per cell).  Here every warp owns a whole cell: there is nothing to synchronise, 12-16 independent
// cells are resident per SM, and the work is arranged so that almost every instruction is a DMMA or the load that feeds it
// (~4 k warp-instructions per cell).  tools/emulate_cellwarp.py is the lane-level executable specification.
//
// Layout idea: every register tile is held TRANSPOSED, tile[g][c] = W[row slot c][column g].  The D fragment of
// mma.m8n8k4 (lane (g,t) holds D[g][2t], D[g][2t+1]) is then directly the A operand of the next product (k-step e of lane
// t stands for column 2t+e), so triangular solves and the Schur update chain in registers; only the B operands come from
// shared memory (L, U, the inverted diagonal blocks) or straight from the record in L2 (A21).
//
//   load     A11 -> shared memory, row-major by ORIGINAL row, leading dimension 40, 16-byte chunks XOR-swizzled with
//            (row>>1)&3 (conflict-free row-per-lane and B-fragment accesses).  Rows never move afterwards:
//            perm[slot] = original row at that pivot position.
//   LU       getrf! (:179): left-looking by column tiles (updates of a tile chain in registers), panel factorisation
//            with one row per lane (implicit partial pivoting, LAPACK's first-maximum rule), inverses of the 8x8 diagonal
//            blocks by a 16-lane substitution (lanes 0-7: inv(L_pp) columns, lanes 8-15: inv(U_pp) columns on a reversed
//            copy).  Multipliers and off-diagonal U tiles are stored negated (DMMA only adds).
//   per 8 columns J of [A12 b1] (getrs! :183,:189, gemm! :186, gemv! :192):
//            Z = L^-1 P A12_J (record -> registers), X = U^-1 Z, S_J = A22_J - A21 X; X is A11^-1 [A12 b1], so
//            keep_factors is one extra store.
#include <algorithm>

#include "common.cuh"

#ifndef GHB_CW_WPC
#define GHB_CW_WPC 4          // warps (= cells in flight) per CTA
#endif
#ifndef GHB_CW_MINB
#define GHB_CW_MINB 3         // CTAs per SM the kernel is compiled for
#endif
#ifndef GHB_CW_EXACT
#define GHB_CW_EXACT 1        // 1: LAPACK's pivot (first exact maximum); 0: maximum to 2^-15 relative
#endif
#ifndef GHB_CW_PREFETCH
#define GHB_CW_PREFETCH 2     // 0: none, 1: whole record at the head of the cell, 2: A12 at the head, A21/A22 after the LU
#endif

namespace ghb {

namespace {

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8_z(unsigned smem_dst, const void* gsrc, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(ok ? 8 : 0));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void l2_prefetch(const void* gptr, unsigned bytes) {
  const unsigned long long a0 = ((unsigned long long)gptr + 15ull) & ~15ull;
  const unsigned long long a1 = ((unsigned long long)gptr + bytes) & ~15ull;
  if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds128(unsigned a, double& v0, double& v1) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts128(unsigned a, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ double flip(double x) {   // -x on the integer pipe (keeps the FP64 pipe for DMMA/DFMA)
  return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}
__device__ __forceinline__ double fast_rcp(double x) {   // MUFU.RCP64H seed + two Newton steps (~1 ulp)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// tables of a plan (built on the host by cw_prepare)
struct CwTables {
  const int32_t* colbase;   // [(n+1)*nf]: record offset of (first row of field f, column c) or -1; c == n: offsets in b
  const uint8_t* rowf;      // [n] field of condensed row r
  const uint8_t* rowl;      // [n] row inside its field
  int nf;
  int fb;                   // the boundary field
  int pf12_off, pf12_len;   // record ranges (doubles) for the L2 prefetches: A12, A21, A22
  int pf21_off, pf21_len;
  int pf22_off, pf22_len;
};

constexpr int CW_LDL = 40;                  // doubles per image row
constexpr unsigned CW_ROWB = CW_LDL * 8;    // bytes per image row

template <int NI, int NB>
struct CwCfg {
  static constexpr int N = NI + NB;
  static constexpr int NC = NB + 1;            // right-hand-side columns: A12 | b1
  static constexpr int RT = (NI + 7) / 8;      // tiles over the interior dofs (panels)
  static constexpr int CTB = (NC + 7) / 8;     // column tiles of [A12 b1]
  static constexpr int BTM = (NB + 7) / 8;     // row tiles over the boundary dofs
  static constexpr int DUMMY = NI;             // the all-zero image row
  static_assert(NI <= CW_LDL && NI <= 64, "image row too short");
  // per-warp shared memory (bytes)
  static constexpr unsigned OFF_INVL = (NI + 1) * CW_ROWB;
  static constexpr unsigned OFF_INVU = OFF_INVL + RT * 512;
  static constexpr unsigned OFF_SCALEU = OFF_INVU + RT * 512;   // [8] -1/pivot, reversed (substitution of inv(U))
  static constexpr unsigned OFF_PERM = OFF_SCALEU + 64;         // [8*RT] bytes
  static constexpr unsigned OFF_INFO = OFF_PERM + ((8 * RT + 15) / 16) * 16;
  static constexpr unsigned WARP_BYTES = OFF_INFO + 16;
  static size_t smem_bytes(int wpc, int nf) {   // + CTA-shared: ones[8], colbase, rowinfo
    return (size_t)wpc * WARP_BYTES + 64 + (size_t)(N + 1) * nf * 4 + ((2 * N + 15) / 16) * 16;
  }
};

// byte offset of image element (row, col) inside the warp's W image
__device__ __forceinline__ unsigned w_off(unsigned row, unsigned col) {
  return row * CW_ROWB + ((((col >> 1) ^ ((row >> 1) & 3u))) << 4) + ((col & 1u) << 3);
}

// ---- panel factorisation: one row per lane, implicit pivoting --------------------------------------
// a[j]: columns c0..c0+7 of the lane's row (a2: second register set, rows 32.. of the first panels).  act: the lane's row
// is still in play.  Returns ch (step at which the row became the pivot row, -1 otherwise).  On return rows still in play
// hold the NEGATED multipliers; pivot row k holds negated multipliers in columns < k and its U row in columns >= k.
template <bool TWO>
__device__ __forceinline__ void panel_factor(double (&a)[8], double (&a2)[8], const bool act1, const bool act2,
                                             const int row1, const int row2, const int npiv, const int c0, int& ch1,
                                             int& ch2, const unsigned scaleu_addr, const unsigned info_addr) {
  const int lane = threadIdx.x & 31;
  ch1 = -1; ch2 = -1;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < npiv) {
      const bool c1 = act1 && ch1 < 0;
      const bool c2 = TWO && act2 && ch2 < 0;
      const unsigned h1 = (unsigned)__double2hiint(a[k]) & 0x7fffffffu;
      const unsigned h2 = TWO ? ((unsigned)__double2hiint(a2[k]) & 0x7fffffffu) : 0u;
      // key = |a| (exponent + 15 mantissa bits) << 6 | preference (set 1 before set 2, low lanes first)
      const unsigned k1 = c1 ? (((h1 >> 5) << 6) | (unsigned)(63 - lane)) : 0u;
      const unsigned k2 = c2 ? (((h2 >> 5) << 6) | (unsigned)(31 - lane)) : 0u;
      const unsigned key = TWO ? (k1 > k2 ? k1 : k2) : k1;
      const double rc1 = fast_rcp(a[k]);          // speculative reciprocal, overlaps the reduction
      const double rc2 = TWO ? fast_rcp(a2[k]) : 0.0;
      unsigned kmax = __reduce_max_sync(0xffffffffu, key);
      int q = 31 - (int)(kmax & 31u);
      bool from2 = TWO && ((kmax & 32u) == 0u);
      bool slow = (kmax >> 6) == 0u;              // all candidates below 2^-1017 (or none): decide exactly
      if (GHB_CW_EXACT) {
        // more than one candidate within 2^-15 of the maximum: resolve with the full magnitude and LAPACK's
        // first-maximum (lowest original row) rule
        const unsigned t1 = __ballot_sync(0xffffffffu, c1 && (k1 >> 6) == (kmax >> 6));
        const unsigned t2 = TWO ? __ballot_sync(0xffffffffu, c2 && (k2 >> 6) == (kmax >> 6)) : 0u;
        slow = slow || (__popc(t1) + __popc(t2) > 1);
      }
      if (slow) {
        const unsigned long long m1 = c1 ? ((unsigned long long)__double_as_longlong(a[k]) & 0x7fffffffffffffffull) : 0ull;
        const unsigned long long m2 = c2 ? ((unsigned long long)__double_as_longlong(a2[k]) & 0x7fffffffffffffffull) : 0ull;
        const unsigned long long mm = m1 > m2 ? m1 : m2;
        const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(mm >> 32));
        const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(mm >> 32) == hi ? (unsigned)mm : 0u);
        const unsigned long long best = ((unsigned long long)hi << 32) | lo;
        if (best == 0ull) {                       // zero column: dgetrf's info = first zero pivot (1-based)
          if (lane == 0) {
            int cur;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(cur) : "r"(info_addr));
            if (cur == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(info_addr), "r"(c0 + k + 1) : "memory");
          }
        }
        // lowest original row among the exact maxima (for a zero column: the first candidate row)
        const unsigned e1 = (c1 && m1 == best) ? (((unsigned)row1 << 6) | 32u | (unsigned)lane) : 0xffffffffu;
        const unsigned e2 = (c2 && m2 == best) ? (((unsigned)row2 << 6) | (unsigned)lane) : 0xffffffffu;
        const unsigned emin = __reduce_min_sync(0xffffffffu, e1 < e2 ? e1 : e2);
        q = (int)(emin & 31u);
        from2 = TWO && ((emin & 32u) == 0u);
      }
      const double rinv = __shfl_sync(0xffffffffu, from2 ? rc2 : rc1, q);
      if (lane == 0) sts64(scaleu_addr + 8u * (unsigned)(7 - k), -rinv);   // -1/pivot, reversed order (inv(U) substitution)
      const bool me1 = !from2 && lane == q, me2 = from2 && lane == q;
      if (me1) ch1 = k;
      if (me2) ch2 = k;
      const bool u1 = c1 && !me1, u2 = c2 && !me2;
      const double l1 = a[k] * rinv;              // dgetf2: scale by the reciprocal of the pivot
      const double m1v = u1 ? l1 : 0.0;           // rows out of play: a - 0*p = a
      a[k] = u1 ? flip(l1) : a[k];
      double m2v = 0.0;
      if (TWO) {
        const double l2 = a2[k] * rinv;
        m2v = u2 ? l2 : 0.0;
        a2[k] = u2 ? flip(l2) : a2[k];
      }
#pragma unroll
      for (int j = k + 1; j < 8; ++j) {
        const double pj = __shfl_sync(0xffffffffu, from2 ? a2[j] : a[j], q);   // pivot row entry, to everyone
        a[j] = fma(-m1v, pj, a[j]);
        if (TWO) a2[j] = fma(-m2v, pj, a2[j]);
      }
    }
  }
}

template <int NI, int NB, int WPC, int MINB, bool KEEPX, bool AL16>
__global__ void __launch_bounds__(32 * WPC, MINB)
condense_cw_kernel(CwTables tb, int lenA, int lenb, int64_t ncells, const double* __restrict__ A,
                   const double* __restrict__ b, double* __restrict__ S, double* __restrict__ gout,
                   int32_t* __restrict__ info, double* __restrict__ X) {
  using C = CwCfg<NI, NB>;
  constexpr int N = C::N, NC = C::NC, RT = C::RT, CTB = C::CTB, BTM = C::BTM, DUMMY = C::DUMMY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  unsigned char* wsp = smem_raw + (size_t)warp * C::WARP_BYTES;
  const unsigned ws = (unsigned)__cvta_generic_to_shared(wsp);           // the warp's image (shared-window address)
  const unsigned a_invL = ws + C::OFF_INVL, a_invU = ws + C::OFF_INVU, a_scaleU = ws + C::OFF_SCALEU;
  const unsigned a_perm = ws + C::OFF_PERM, a_info = ws + C::OFF_INFO;
  unsigned char* shp = smem_raw + (size_t)WPC * C::WARP_BYTES;           // CTA-shared tables
  const unsigned a_ones = (unsigned)__cvta_generic_to_shared(shp);
  int* s_colbase = reinterpret_cast<int*>(shp + 64);                     // [(N+1)*nf]
  unsigned short* s_rowinfo = reinterpret_cast<unsigned short*>(s_colbase + (N + 1) * tb.nf);   // [N]: field<<8 | local row
  const int nf = tb.nf;

  // one-time: zero the warp's region (pad columns, the dummy row and the inverse tiles stay zero where never written)
  for (unsigned i = lane; i < C::WARP_BYTES / 8; i += 32) reinterpret_cast<double*>(wsp)[i] = 0.0;
  for (int i = threadIdx.x; i < 8; i += 32 * WPC) reinterpret_cast<double*>(shp)[i] = 1.0;
  for (int i = threadIdx.x; i < (N + 1) * nf; i += 32 * WPC) s_colbase[i] = tb.colbase[i];
  for (int i = threadIdx.x; i < N; i += 32 * WPC) s_rowinfo[i] = (unsigned short)((tb.rowf[i] << 8) | tb.rowl[i]);
  __syncthreads();

  // loader: lane <-> interior row (lane, and 32 + lane)
  const int lr1 = lane < NI ? lane : 0, lr2 = (NI > 32 && lane + 32 < NI) ? lane + 32 : 0;
  const bool lv1 = lane < NI, lv2 = NI > 32 && lane + 32 < NI;
  const int li1 = s_rowinfo[lr1], li2 = s_rowinfo[lr2];
  const unsigned ldst1 = ws + (unsigned)lr1 * CW_ROWB, ldst2 = ws + (unsigned)lr2 * CW_ROWB;
  const unsigned lsw1 = ((unsigned)lr1 >> 1) & 3u, lsw2 = ((unsigned)lr2 >> 1) & 3u;
  // S phase: record offsets of the A21 fragments of this lane, element (boundary row 8m + g, interior column 8p+2t+e)
  // = cbk[p][e] + 8m (one boundary field: its rows are contiguous in every packed column)
  const int lb0 = tb.fb;

  const int64_t wstride = (int64_t)gridDim.x * WPC;
  for (int64_t cell = (int64_t)blockIdx.x * WPC + warp; cell < ncells; cell += wstride) {
    const double* Arec = A + cell * lenA;
    const double* brec = b + cell * lenb;
    // ------------------------------------------------------------------ load A11 (transposing 8-byte cp.async)
    {
      const double* s1 = Arec + (li1 & 0xff);
      const double* s2 = Arec + (li2 & 0xff);
      const int* cb1 = s_colbase + (li1 >> 8);
      const int* cb2 = s_colbase + (li2 >> 8);
#pragma unroll 2
      for (int cc = 0; cc < (NI + 1) / 2; ++cc) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = 2 * cc + h;
          if (c < NI) {
            const int o1 = cb1[c * nf];
            if (lv1) cp_async8_z(ldst1 + ((((unsigned)cc ^ lsw1)) << 4) + 8u * h, s1 + (o1 >= 0 ? o1 : 0), o1 >= 0);
            if (NI > 32) {
              const int o2 = cb2[c * nf];
              if (lv2) cp_async8_z(ldst2 + ((((unsigned)cc ^ lsw2)) << 4) + 8u * h, s2 + (o2 >= 0 ? o2 : 0), o2 >= 0);
            }
          }
        }
      }
      if (GHB_CW_PREFETCH == 1) { if (lane == 0) l2_prefetch(Arec, (unsigned)lenA * 8u); }
      if (GHB_CW_PREFETCH == 2) { if (lane == 0) l2_prefetch(Arec + tb.pf12_off, (unsigned)tb.pf12_len * 8u); }
    }
    // perm = identity, info = 0
    {
      unsigned char* pp = wsp + C::OFF_PERM;
      pp[lane] = (unsigned char)(lane < NI ? lane : DUMMY);
      if (lane + 32 < 8 * RT) pp[lane + 32] = (unsigned char)(lane + 32 < NI ? lane + 32 : DUMMY);
      if (lane == 0) *reinterpret_cast<int*>(wsp + C::OFF_INFO) = 0;
    }
    cp_async_wait_all();
    __syncwarp();

    // ------------------------------------------------------------------ LU of A11, left-looking by column tiles
    int myrow = lane < NI ? lane : -1;                       // row of register set 1 (-1: lane retired)
    int myrow2 = (NI > 32 && lane + 32 < NI) ? lane + 32 : -1;
    bool two = NI > 32;
#pragma unroll 1
    for (int R = 0; R < RT; ++R) {
      const int c0 = 8 * R;
      const int npiv = (NI - c0) < 8 ? (NI - c0) : 8;
      if (R > 0) {
        // ---- bring column tile R up to date with the panels q < R (registers only), store it back
        unsigned ea[RT][2];                                  // image addresses of this lane's tile elements
        double T[RT][2];
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          const unsigned pr = lds_u16(a_perm + (unsigned)(8 * j) + 2u * (unsigned)t);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const unsigned row = (e ? (pr >> 8) : pr) & 0xffu;
            ea[j][e] = ws + w_off(row, (unsigned)(c0 + g));
          }
        }
#pragma unroll
        for (int j = 0; j < RT; ++j) { T[j][0] = lds64(ea[j][0]); T[j][1] = lds64(ea[j][1]); }
        unsigned rb[RT];                                     // B-fragment rows: slot 8i + g, chunk t (swizzled)
#pragma unroll
        for (int i = 1; i < RT; ++i) {
          const unsigned row = lds_u8(a_perm + (unsigned)(8 * i + g));
          rb[i] = ws + row * CW_ROWB + ((((unsigned)t ^ ((row >> 1) & 3u))) << 4);
        }
#pragma unroll
        for (int q = 0; q < RT - 1; ++q) {
          if (q < R) {
            double b0, b1;
            lds128(a_invL + 512u * q + 64u * g + 16u * t, b0, b1);
            double u0 = 0.0, u1 = 0.0;                       // U_qR^T = T_q inv(L_qq)^T
            dmma(u0, u1, T[q][0], b0);
            dmma(u0, u1, T[q][1], b1);
            T[q][0] = u0; T[q][1] = u1;
#pragma unroll
            for (int i = q + 1; i < RT; ++i) {               // T_i -= U_qR^T L_iq^T   (multipliers are stored negated)
              double l0, l1;
              lds128(rb[i] + 64u * q, l0, l1);
              dmma(T[i][0], T[i][1], u0, l0);
              dmma(T[i][0], T[i][1], u1, l1);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          const bool isU = j < R;                            // rows of earlier pivots: final U entries, stored negated
          sts64(ea[j][0], isU ? flip(T[j][0]) : T[j][0]);
          sts64(ea[j][1], isU ? flip(T[j][1]) : T[j][1]);
        }
        __syncwarp();
      }
      // ---- panel: one row per lane
      double a[8], a2[8];
      const bool act1 = myrow >= 0, act2 = two && myrow2 >= 0;
      const unsigned r1 = act1 ? (unsigned)myrow : (unsigned)DUMMY, r2 = act2 ? (unsigned)myrow2 : (unsigned)DUMMY;
      const unsigned pa1 = ws + r1 * CW_ROWB + 64u * R, pa2 = ws + r2 * CW_ROWB + 64u * R;
      const unsigned sw1 = (r1 >> 1) & 3u, sw2 = (r2 >> 1) & 3u;
#pragma unroll
      for (int c = 0; c < 4; ++c) lds128(pa1 + (((unsigned)c ^ sw1) << 4), a[2 * c], a[2 * c + 1]);
      if (two) {
#pragma unroll
        for (int c = 0; c < 4; ++c) lds128(pa2 + (((unsigned)c ^ sw2) << 4), a2[2 * c], a2[2 * c + 1]);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) a2[c] = 0.0;
      }
      int ch1, ch2;
      if (two) panel_factor<true>(a, a2, act1, act2, myrow, myrow2, npiv, c0, ch1, ch2, a_scaleU, a_info);
      else panel_factor<false>(a, a2, act1, false, myrow, 0, npiv, c0, ch1, ch2, a_scaleU, a_info);
      // write back (rows in play only: the others hold U entries of this column tile)
      if (act1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(pa1 + (((unsigned)c ^ sw1) << 4), a[2 * c], a[2 * c + 1]);
      }
      if (two && act2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(pa2 + (((unsigned)c ^ sw2) << 4), a2[2 * c], a2[2 * c + 1]);
      }
      // the 8x8 diagonal block in pivot order, twice: natural (inv(L) lanes) and reversed in both directions (inv(U) lanes);
      // staged in the tiles that receive the inverses
      {
        const unsigned stL = a_invL + 512u * R, stU = a_invU + 512u * R;
        if (npiv < 8) {                       // partial last panel: rows >= npiv of the stage are zero
          sts128(stL + 16u * lane, 0.0, 0.0);
          sts128(stU + 16u * lane, 0.0, 0.0);
          __syncwarp();
        }
        if (ch1 >= 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sts128(stL + 64u * ch1 + 16u * c, a[2 * c], a[2 * c + 1]);
            sts128(stU + 64u * (7 - ch1) + 16u * (3 - c), a[2 * c + 1], a[2 * c]);
          }
        }
        if (two && ch2 >= 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sts128(stL + 64u * ch2 + 16u * c, a2[2 * c], a2[2 * c + 1]);
            sts128(stU + 64u * (7 - ch2) + 16u * (3 - c), a2[2 * c + 1], a2[2 * c]);
          }
        }
      }
      // ---- perm: the pivots of this panel, then the rows still in play (set 1 in lane order, then set 2)
      {
        const bool rest1 = act1 && ch1 < 0, rest2 = act2 && ch2 < 0;
        const unsigned mr1 = __ballot_sync(0xffffffffu, rest1);
        const unsigned lt = (1u << lane) - 1u;
        unsigned char* pp = wsp + C::OFF_PERM + c0;
        if (ch1 >= 0) pp[ch1] = (unsigned char)myrow;
        else if (rest1) pp[npiv + __popc(mr1 & lt)] = (unsigned char)myrow;
        if (two) {
          const unsigned mr2 = __ballot_sync(0xffffffffu, rest2);
          if (ch2 >= 0) pp[ch2] = (unsigned char)myrow2;
          else if (rest2) pp[npiv + __popc(mr1) + __popc(mr2 & lt)] = (unsigned char)myrow2;
        }
        if (ch1 >= 0) myrow = -1;
        if (ch2 >= 0) myrow2 = -1;
        if (two) {
          // compaction: rows of the second register set move into retired lanes
          const unsigned fr = __ballot_sync(0xffffffffu, myrow < 0);
          const unsigned m2 = __ballot_sync(0xffffffffu, myrow2 >= 0);
          const int nmove = min(__popc(fr), __popc(m2));
          const int idx = __popc(fr & lt);
          const int src = (int)__fns(m2, 0, idx + 1);              // lane holding the idx-th remaining row of set 2
          const int v = __shfl_sync(0xffffffffu, myrow2, src & 31);
          if (((fr >> lane) & 1u) && idx < nmove) myrow = v;
          if (myrow2 >= 0 && __popc(m2 & lt) < nmove) myrow2 = -1;
          two = __any_sync(0xffffffffu, myrow2 >= 0);
        }
      }
      __syncwarp();
      // ---- inverses of the diagonal block: lanes 0-7 columns of inv(L_pp), lanes 8-15 columns of inv(U_pp) (reversed)
      {
        const int cidx = lane & 7, grp = (lane >> 3) & 1;
        const unsigned st = (grp ? a_invU : a_invL) + 512u * R;
        const unsigned sc = grp ? a_scaleU : a_ones;
        const int ncol = grp ? 7 - cidx : cidx;             // column of the inverse this lane computes
        const bool cv = ncol < npiv;
        double z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0.0;
        {
          const double v = lds64(sc + 8u * cidx);           // grp 0: 1; grp 1: -1/u_nn
          const double zi = grp ? flip(v) : v;
#pragma unroll
          for (int i = 0; i < 8; ++i) if (i == cidx) z[i] = cv ? zi : 0.0;
        }
#pragma unroll
        for (int s = 1; s < 8; ++s) {
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int j = 0; j < (s + 1) / 2; ++j) {
            double v0, v1;
            lds128(st + 64u * s + 16u * j, v0, v1);
            d0 = fma(v0, z[2 * j], d0);
            d1 = fma(v1, z[2 * j + 1], d1);
          }
          const double scl = lds64(sc + 8u * s);            // grp 0: 1; grp 1: -1/u_ii of the row of this step
          const bool rowok = (grp ? 7 - s : s) < npiv;      // rows beyond a partial block
          if (s > cidx) z[s] = (cv && rowok) ? (d0 + d1) * scl : 0.0;
        }
        __syncwarp();                                        // every lane has read its stage rows
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // tile[row][col]: grp 0: row i, col cidx; grp 1: row 7-i, col 7-cidx
            const unsigned o = grp ? (64u * (7 - i) + 8u * (7 - cidx)) : (64u * i + 8u * cidx);
            sts64(st + o, z[i]);
          }
        }
      }
      __syncwarp();
    }

    // ------------------------------------------------------------------ phase B: column tiles of [A12 b1]
    if (GHB_CW_PREFETCH == 2 && lane == 0) {
      l2_prefetch(Arec + tb.pf21_off, (unsigned)tb.pf21_len * 8u);
      if (tb.pf22_len > 0) l2_prefetch(Arec + tb.pf22_off, (unsigned)tb.pf22_len * 8u);
    }
    int failed;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(failed) : "r"(a_info));
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // per-slot record info of this lane's tile rows (slot 8j + 2t + e): field<<8 | local row, 0xffff for the dummy row
    unsigned sinfo[RT];
    unsigned rb[RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const unsigned pr = lds_u16(a_perm + (unsigned)(8 * j) + 2u * (unsigned)t);
      const unsigned ra = pr & 0xffu, rbq = (pr >> 8) & 0xffu;
      const unsigned ia = ra < (unsigned)NI ? s_rowinfo[ra] : 0xffffu;
      const unsigned ib = rbq < (unsigned)NI ? s_rowinfo[rbq] : 0xffffu;
      sinfo[j] = ia | (ib << 16);
      const unsigned row = lds_u8(a_perm + (unsigned)(8 * j + g));
      rb[j] = ws + row * CW_ROWB + ((((unsigned)t ^ ((row >> 1) & 3u))) << 4);
    }
    // A21 fragment offsets (see above): k = 8p + 2t + e
    int cbk[RT][2];
#pragma unroll
    for (int p = 0; p < RT; ++p)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * p + 2 * t + e;
        const int v = k < NI ? s_colbase[k * nf + lb0] : -1;
        cbk[p][e] = v >= 0 ? v + g : -1;
      }
    const int a22base = s_colbase[NI * nf + lb0];            // record offset of A22(0,0) or -1 (untouched: RT-H)
    const int b2base = s_colbase[N * nf + lb0];
    double* Sc = S + cell * (int64_t)NB * NB;
    double* gc = gout + cell * (int64_t)NB;

#pragma unroll 1
    for (int J = 0; J < CTB; ++J) {
      const int col = 8 * J + g;                             // column of [A12 b1] held by this lane's fragments
      const bool cA = col < NB, cB = col == NB;              // A12 column / the right-hand side / padding
      // ---- T = (P A12_J)^T straight from the record
      double T[RT][2];
      {
        const int* cbp = s_colbase + (size_t)(cA ? NI + col : N) * nf;
        const double* base = cB ? brec : Arec;
#pragma unroll
        for (int j = 0; j < RT; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const unsigned si = (sinfo[j] >> (16 * e)) & 0xffffu;
            const int o = (si != 0xffffu && (cA || cB)) ? cbp[si >> 8] : -1;
            T[j][e] = o >= 0 ? __ldg(base + o + (si & 0xffu)) : 0.0;
          }
      }
      // ---- Z = L^-1 P A12_J
#pragma unroll
      for (int q = 0; q < RT; ++q) {
        double b0, b1;
        lds128(a_invL + 512u * q + 64u * g + 16u * t, b0, b1);
        double z0 = 0.0, z1 = 0.0;
        dmma(z0, z1, T[q][0], b0);
        dmma(z0, z1, T[q][1], b1);
        T[q][0] = z0; T[q][1] = z1;
#pragma unroll
        for (int i = q + 1; i < RT; ++i) {
          double l0, l1;
          lds128(rb[i] + 64u * q, l0, l1);
          dmma(T[i][0], T[i][1], z0, l0);
          dmma(T[i][0], T[i][1], z1, l1);
        }
      }
      // ---- X = U^-1 Z
#pragma unroll
      for (int q = RT - 1; q >= 0; --q) {
        double b0, b1;
        lds128(a_invU + 512u * q + 64u * g + 16u * t, b0, b1);
        double x0 = 0.0, x1 = 0.0;
        dmma(x0, x1, T[q][0], b0);
        dmma(x0, x1, T[q][1], b1);
        T[q][0] = x0; T[q][1] = x1;
#pragma unroll
        for (int p = q - 1; p >= 0; --p) {                   // off-diagonal U tiles are stored negated
          double u0, u1;
          lds128(rb[p] + 64u * q, u0, u1);
          dmma(T[p][0], T[p][1], x0, u0);
          dmma(T[p][0], T[p][1], x1, u1);
        }
      }
      if (KEEPX) {
        // X = A11^-1 [A12 | b1], col-major n_i x (n_b+1) per cell
        if (cA || cB) {
          double* Xc = X + cell * (int64_t)(NI * NC) + (int64_t)col * NI;
#pragma unroll
          for (int p = 0; p < RT; ++p) {
            const int k = 8 * p + 2 * t;
            if (k < NI) Xc[k] = failed ? qnan : T[p][0];
            if (k + 1 < NI) Xc[k + 1] = failed ? qnan : T[p][1];
          }
        }
      }
      // ---- S_J = A22_J - A21 X_J  (transposed tiles: acc[m][e] = S[8m + 2t + e][col])
      double acc[BTM][2];
#pragma unroll
      for (int m = 0; m < BTM; ++m) {
        const int r = 8 * m + 2 * t;
        acc[m][0] = 0.0; acc[m][1] = 0.0;
        if (cA) {
          if (a22base >= 0) {
            const double* src = Arec + a22base + col * NB + r;
            if (AL16) {
              if (r < NB) { const double2 v = __ldg(reinterpret_cast<const double2*>(src)); acc[m][0] = v.x; acc[m][1] = v.y; }
            } else {
              if (r < NB) acc[m][0] = __ldg(src);
              if (r + 1 < NB) acc[m][1] = __ldg(src + 1);
            }
          }
        } else if (cB) {
          if (r < NB) acc[m][0] = __ldg(brec + b2base + r);
          if (r + 1 < NB) acc[m][1] = __ldg(brec + b2base + r + 1);
        }
      }
#pragma unroll
      for (int p = 0; p < RT; ++p) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (8 * p + e < NI) {                              // k-step with at least one real column
            const double xa = flip(T[p][e]);
            const int o = cbk[p][e];
            double bf[BTM];
#pragma unroll
            for (int m = 0; m < BTM; ++m)
              bf[m] = (o >= 0 && (NB % 8 == 0 || m + 1 < BTM || 8 * m + g < NB)) ? __ldg(Arec + o + 8 * m) : 0.0;
#pragma unroll
            for (int m = 0; m < BTM; ++m) dmma(acc[m][0], acc[m][1], xa, bf[m]);
          }
        }
      }
      // ---- store
      if (cA || cB) {
        double* dst = cA ? Sc + (int64_t)col * NB : gc;
#pragma unroll
        for (int m = 0; m < BTM; ++m) {
          const int r = 8 * m + 2 * t;
          const double v0 = failed ? qnan : acc[m][0], v1 = failed ? qnan : acc[m][1];
          if (AL16) {
            if (r < NB) *reinterpret_cast<double2*>(dst + r) = make_double2(v0, v1);
          } else {
            if (r < NB) dst[r] = v0;
            if (r + 1 < NB) dst[r + 1] = v1;
          }
        }
      }
    }
    if (info && lane == 0) info[cell] = failed;
    __syncwarp();
  }
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------
// Plans the kernel is instantiated for: one boundary field, n_i <= 40 (rows 32.. move into retired lanes after the first
// panel), interior fields first in the condensed order (always true: Plan orders interior rows first).
static bool cw_shape(int ni, int nb) {
  return (ni == 34 && nb == 36) || (ni == 33 && nb == 12) || (ni == 40 && nb == 36) || (ni == 21 && nb == 16);
}

bool cw_supported(const Plan& p) {
  if (p.boundary.size() != 1 || p.nfields > 8) return false;
  for (int r = 0; r < p.n; ++r) if (p.row_local[r] > 254) return false;
  return cw_shape(p.n_i, p.n_b);
}

static bool cw_al16(const Plan& p) {   // 16-byte aligned pairs of boundary rows in A22 / S
  const int fb = p.boundary[0] - 1;
  const int64_t bo = p.block_offset[fb + p.nfields * fb];
  return p.n_b % 2 == 0 && p.lenA % 2 == 0 && (bo < 0 || bo % 2 == 0);
}

int cw_prepare(ghb_ctx* ctx, Plan& p) {
  const int n = p.n, nf = p.nfields;
  if (!p.d_colbase) {
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
  }
  // bounding record ranges of the A12 / A21 / A22 blocks (L2 prefetches)
  auto range = [&](bool rows_int, bool cols_int, int& off, int& len) {
    int64_t lo = -1, hi = -1;
    for (int fj = 0; fj < nf; ++fj)
      for (int fi = 0; fi < nf; ++fi) {
        const bool ri = std::find(p.interior.begin(), p.interior.end(), fi + 1) != p.interior.end();
        const bool ci = std::find(p.interior.begin(), p.interior.end(), fj + 1) != p.interior.end();
        const int64_t bo = p.block_offset[fi + nf * fj];
        if (ri != rows_int || ci != cols_int || bo < 0) continue;
        const int64_t e = bo + (int64_t)p.ndofs[fi] * p.ndofs[fj];
        lo = lo < 0 ? bo : std::min(lo, bo);
        hi = std::max(hi, e);
      }
    off = lo < 0 ? 0 : (int)lo;
    len = lo < 0 ? 0 : (int)(hi - lo);
  };
  range(true, false, p.cw_pf[0], p.cw_pf[1]);
  range(false, true, p.cw_pf[2], p.cw_pf[3]);
  range(false, false, p.cw_pf[4], p.cw_pf[5]);
  return GHB_OK;
}

template <int NI, int NB, bool KEEPX, bool AL16>
static int launch_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S, double* g,
                     int32_t* info, double* X) {
  constexpr int WPC = GHB_CW_WPC, MINB = GHB_CW_MINB;
  auto kern = condense_cw_kernel<NI, NB, WPC, MINB, KEEPX, AL16>;
  const size_t smem = CwCfg<NI, NB>::smem_bytes(WPC, p.nfields);
  static int per_sm_cached = -1;          // per instantiation: attributes and occupancy are set once
  if (per_sm_cached < 0) {
    GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GHB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    GHB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * WPC, smem));
    if (per_sm < 1) return fail(ctx, GHB_ECUDA, "condense_cw_kernel does not fit on an SM");
    if (const char* cap = getenv("GHB_MAX_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(cap)));   // profiling knob
    if (getenv("GHB_DEBUG")) fprintf(stderr, "condense_cw<%d,%d>: %d CTAs/SM x %d warps, %zu B smem\n", NI, NB, per_sm, WPC, smem);
    per_sm_cached = per_sm;
  }
  const int fb = p.boundary[0] - 1;
  CwTables tb{p.d_colbase, p.d_rowf, p.d_rowl, p.nfields, fb,
              p.cw_pf[0], p.cw_pf[1], p.cw_pf[2], p.cw_pf[3], p.cw_pf[4], p.cw_pf[5]};
  const int64_t want = (ncells + WPC - 1) / WPC;
  const int64_t grid = std::min<int64_t>(want, (int64_t)ctx->sm_count * per_sm_cached);
  kern<<<(unsigned)grid, 32 * WPC, smem, ctx->stream>>>(tb, p.lenA, p.lenb, ncells, A, b, S, g, info, X);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

template <int NI, int NB>
static int launch_cw_shape(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                           double* g, int32_t* info, double* X) {
  const bool al = cw_al16(p);
  if (X) {
    if (al) return launch_cw<NI, NB, true, true>(ctx, p, ncells, A, b, S, g, info, X);
    return launch_cw<NI, NB, true, false>(ctx, p, ncells, A, b, S, g, info, X);
  }
  if (al) return launch_cw<NI, NB, false, true>(ctx, p, ncells, A, b, S, g, info, X);
  return launch_cw<NI, NB, false, false>(ctx, p, ncells, A, b, S, g, info, X);
}

int launch_condense_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                       double* g, int32_t* info, double* X) {
  if (p.n_i == 34 && p.n_b == 36) return launch_cw_shape<34, 36>(ctx, p, ncells, A, b, S, g, info, X);
  if (p.n_i == 33 && p.n_b == 12) return launch_cw_shape<33, 12>(ctx, p, ncells, A, b, S, g, info, X);
  if (p.n_i == 40 && p.n_b == 36) return launch_cw_shape<40, 36>(ctx, p, ncells, A, b, S, g, info, X);
  if (p.n_i == 21 && p.n_b == 16) return launch_cw_shape<21, 16>(ctx, p, ncells, A, b, S, g, info, X);
  return fail(ctx, GHB_EUNSUPPORTED, "condense_cw: shape not instantiated");
}

}  // namespace ghb
