// condense_cw_kernel.cuh -- static condensation with ONE WARP PER CELL and no barriers ("cell-warp" kernel), FP64 DMMA.
// Replaces evaluate!(cache, ::StaticCondensationMap, A, b) (/root/reference/src/StaticCondensationMap.jl:152-196)
// for mid-size cells (n_i <= 64 interior dofs, one skeleton field).
//
// Why: the 4-warps-per-cell kernels of condense_dmma.cu spend most of their time in named barriers between the panel
// warp and the column-tile owners (profiles/r01_condense_dmma_summary.md: 7.7 barrier stalls per issue, 12.5 k
// warp-instructions per cell).  Here every warp owns a whole cell: nothing to synchronise, 12-16 independent cells per SM.
// The first version of this kernel (profiles/r02_cw_summary.md) executed 9.9 k warp-instructions per cell, two thirds of them
// integer address arithmetic; this version takes every address from tables built once per plan (loader), once per CTA
// (A21 fragments) or once per cell (pivoted row addresses), ~5 k warp-instructions per cell.
//
// Layout idea: every register tile is held TRANSPOSED, tile[g][c] = W[row slot c][column g].  The D fragment of
// mma.m8n8k4 (lane (g,t) holds D[g][2t], D[g][2t+1]) is then directly the A operand of the next product, so triangular
// solves and the Schur update chain in registers; only the B operands come from shared memory (L, U, the inverted
// diagonal blocks) or straight from the record in L2 (A21).
//
// Index convention ("mu order"): inside every group of 8 pivots / interior columns, the element with logical index k
// (pivot order, column order) is kept at PHYSICAL position mu(k) = 2 (k & 3) + (k >> 2) -- in the image columns, in the
// pivot position table and in the inverse tiles.  D register e of lane (g,t) (hardware column 2t + e) then stands for
// logical index 4e + t, i.e. k-step e of the next product contracts the logical indices 4e .. 4e+3: a partial last
// panel (n_i = 34: two pivots) needs one k-step instead of two in every product it takes part in.
//
//   load     A11 -> shared memory by a per-plan table of (record offset, image offset) pairs, 8-byte cp.async in record
//            order (coalesced); image row-major by ORIGINAL row, 16-byte chunks XOR-swizzled with (row >> 1) & 3.
//            Rows never move: the position table prow[] holds, per pivot position, the image address of its row.
//   LU       getrf! (:179): left-looking by column tiles (updates of a tile chain in registers), panel factorisation
//            with one row per lane (implicit partial pivoting; exact ties are broken like dgetf2/idamax: lowest position
//            in LAPACK's swapped row order, tracked per row), inverses of the 8x8 diagonal blocks by a 16-lane
//            substitution (lanes 0-7: columns of inv(L_pp), lanes 8-15: columns of inv(U_pp) on a reversed copy).
//            Multipliers and off-diagonal U tiles are stored negated (DMMA only adds).
//   per 8 columns J of [A12 b1] (getrs! :183,:189, gemm! :186, gemv! :192):
//            Z = L^-1 P A12_J (record -> registers), X = U^-1 Z, S_J = A22_J - A21 X; X is A11^-1 [A12 b1], so
//            keep_factors is one extra store.
#pragma once
#include <algorithm>
#include <type_traits>

#include "common.cuh"

#ifndef GHB_CW_WPC
#define GHB_CW_WPC 4          // warps (= cells in flight) per CTA
#endif
#ifndef GHB_CW_MINB
#define GHB_CW_MINB 4         // CTAs per SM the kernel is compiled for
#endif
#ifndef GHB_CW_NJ
#define GHB_CW_NJ 1           // column tiles of [A12 b1] per pass of phase B (2: B fragments shared by two DMMA chains)
#endif
#ifndef GHB_CW_CARVEOUT
#define GHB_CW_CARVEOUT 100   // preferred shared-memory carve-out (%): below 100 leaves L1 for the A21 / A12 reads
#endif
#ifndef GHB_CW_EXACT
#define GHB_CW_EXACT 1        // 1: LAPACK's pivot (first exact maximum in swapped order); 0: maximum to 2^-15 relative
#endif
#ifndef GHB_CW_PREFETCH
#define GHB_CW_PREFETCH 2     // 0: none, 1: whole record at the head of the cell, 2: A12 at the head, A21/A22 after the LU,
                              // 3: the warp's next record when the LU of this one is done
#endif

namespace ghb {

namespace {

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8_z(unsigned smem_dst, const void* gsrc, unsigned nbytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(nbytes));
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
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds_2u32(unsigned a, unsigned& v0, unsigned& v1) {
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a));
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts128(unsigned a, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ double flip(double x) {   // -x on the integer pipe (keeps the FP64 pipe for DMMA/DFMA)
  return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}
__device__ __forceinline__ double flip_if(double x, unsigned mask) {   // mask = 0x80000000 or 0
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}
__device__ __forceinline__ double neg_rcp(double x) {   // -1/x: MUFU.RCP64H seed + two Newton steps (~1 ulp)
  const double y = flip(x);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
  double e = fma(-y, r, 1.0);
  r = fma(r, e, r);
  e = fma(-y, r, 1.0);
  return fma(r, e, r);
}
// ---- TMA bulk copies into shared memory, completion on an mbarrier (GEN kernels: table chunks) ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__host__ __device__ constexpr int mu(int k) { return 2 * (k & 3) + (k >> 2); }   // logical index -> physical position

// tables of a plan (built on the host by cw_prepare)
struct CwArgs {
  const uint2* ldtab;       // [nld] loader: (byte offset in the record or 0xffffffff = zero, byte offset in the image)
  const uint32_t* rowA12;   // [n_i + 1] per interior row: record offset of A12(row, 0) | stride << 16; 0xffffffff: zero row
  const uint16_t* rowb;     // [n_i + 1] per interior row: offset in the b record
  const int32_t* colA21;    // [8 RT] per interior column: record offset of A21(0, col) or -1
  const int32_t* colbase;   // PAD kernels: [(n+1)*nf] record offset of (first row of field f, condensed column c) or -1;
                            // c == n: offsets in the b record
  const uint16_t* rowinfo;  // PAD kernels: [n] field << 8 | local row of condensed row r
  int n_i, n_b, nf;         // PAD kernels: the plan's real sizes (<= the padded NI, NB of the instantiation)
  int nld;
  int a22base, b2base;      // record offsets of A22(0,0) (-1: untouched) and of b2 in the b record
  int al16;                 // pairs of boundary rows are 16-byte aligned in A22, b2, S and g
  int pf12_off, pf12_len;   // record ranges (doubles) for the L2 prefetches: A12, A21, A22
  int pf21_off, pf21_len;
  int pf22_off, pf22_len;
  int lenA, lenb;
  int64_t ncells;
  const double* A;
  const double* b;
  double* S;
  double* g;
  int32_t* info;
  double* X;
  // fused assembly (SCAT kernels): S_K is added into the zeroed nzval through the scatter map of the selected pattern
  double* nzval;
  const int64_t* colpos;    // [ncells][n_b] 0-based nzval offset of the column of local dof lj, -1: not assembled here
  const uint8_t* rowrank;   // [ncells][n_b][n_b] rank of row li inside the column of lj, 255: not assembled
  const uint8_t* keepS;     // [ncells] 1: also store S_K (Dirichlet lift, cut-plane pack); NULL: never
  // records of an affine family generated in the loader (GEN kernels): A_K = sum_t coef[K][t] TA[t]
  const double* TA;         // [ntab][lenAp] rows padded to an even length (16-byte aligned rows for the TMA copies)
  const double* Tb;         // [ntab][lenbp]
  int lenAp, lenbp;
  const double* coef;       // [ncells][ntab]
  double* scratch;          // [gridDim.x * WPC][slot] the record of the cell a warp is working on (stays in L2)
  int64_t slot;             // doubles per scratch record (lenAp + lenbp rounded up to 128 bytes); b_K sits at offset lenAp
  int ntab;
  // backward map (BACK kernels, BackwardStaticCondensationMap): u_K = A11^-1 (b1 - A12 lambda_K)
  const double* lam_free;   // free skeleton dof values
  const double* lam_dir;    // Dirichlet skeleton dof values (may be NULL: zeros)
  const int64_t* ids;       // [ncells][n_b] 1-based skeleton dof ids, < 0: Dirichlet
  double* u;                // [ncells][n_i] interior dof values, condensed order
  const uint16_t* gen_need; // chunks to generate (GEN + BACK: only those that hold A11, A12 or b1), NULL: all
  int gen_nneed;
  int gen_E;                // table elements per staged chunk (a multiple of 16; 2 buffers x ntab x (gen_E + 4) doubles fit an image)
};

// record loads: the records of a GEN kernel are rewritten in place by the CTA itself -- no non-coherent loads
template <bool GEN>
__device__ __forceinline__ double ldrec(const double* p) { return GEN ? *p : __ldg(p); }
template <bool GEN>
__device__ __forceinline__ double2 ldrec2(const double2* p) { return GEN ? *p : __ldg(p); }
// 256-bit accesses (sm_100: LDG.E.256 / STG.E.256), 32-byte aligned addresses
template <bool GEN>
__device__ __forceinline__ void ldrec4(const double* p, double& a, double& b, double& c, double& d) {
  if (GEN) asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
  else asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <bool GEN>
__device__ __forceinline__ void strec4(double* p, double a, double b, double c, double d) {
  if (GEN) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
  else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int NI, int NB>
struct CwCfg {
  static constexpr int NC = NB + 1;            // right-hand-side columns: A12 | b1
  static constexpr int RT = (NI + 7) / 8;      // tiles over the interior dofs (panels)
  static constexpr int CTB = (NC + 7) / 8;     // column tiles of [A12 b1]
  static constexpr int BTM = (NB + 7) / 8;     // row tiles over the boundary dofs
  static constexpr int NPL = NI - 8 * (RT - 1);   // pivots of the last panel
  static constexpr int DUMMY = NI;             // the all-zero image row
  static constexpr unsigned ROWB = 64u * (RT | 1);   // bytes per image row: an odd number of 64-byte tiles
  static_assert(NI <= 64, "two register sets hold at most 64 rows");
  // per-warp shared memory (bytes)
  static constexpr unsigned OFF_INVL = (NI + 1) * ROWB;
  static constexpr unsigned OFF_SCALEU = OFF_INVL + RT * 512;   // [8] -1/pivot, reversed (substitution of inv(U))
  static constexpr unsigned OFF_PROW = OFF_SCALEU + 64;         // [8*RT] u32: image address info of the row at a position
  static constexpr unsigned OFF_INFO = OFF_PROW + 32 * RT;
  static constexpr unsigned WARP_BYTES = OFF_INFO + 16;
  // CTA-shared: ones[8], rowA12[NI+1] (u32), rowb[NI+1] (u16)
  static constexpr unsigned SH_ROWA12 = 64;
  static constexpr unsigned SH_ROWB = SH_ROWA12 + 4 * ((NI + 1 + 3) & ~3);
  static constexpr unsigned SH_BYTES = SH_ROWB + 2 * ((NI + 1 + 7) & ~7);
  // PAD kernels add: colbase[(NI+NB+1) * 8 fields] (i32), rowinfo[NI+NB] (u16)
  static constexpr unsigned SH_COLBASE = SH_BYTES;
  static constexpr unsigned SH_ROWINFO = SH_COLBASE + 4 * (NI + NB + 1) * 8;
  static constexpr unsigned SH_BYTES_PAD = SH_ROWINFO + 2 * ((NI + NB + 7) & ~7);
  // GEN kernels add: two mbarriers per warp (table staging rings) and 16 zero bytes
  static constexpr unsigned SH_BAR = (SH_BYTES + 15u) & ~15u;
  static constexpr unsigned SH_BAR_PAD = (SH_BYTES_PAD + 15u) & ~15u;
  static constexpr int MAXTAB = 16;
  static size_t smem_bytes(int wpc, bool pad = false, bool gen = false) {
    return (size_t)wpc * WARP_BYTES + (gen ? (pad ? SH_BAR_PAD : SH_BAR) + 16u * wpc + 16u : (pad ? SH_BYTES_PAD : SH_BYTES));
  }
  // position-table entry of image row r: byte offset of the row | swizzle bits (4,5) | r << 16
  __host__ __device__ static constexpr unsigned enc(unsigned r) { return r * ROWB | ((r & 6u) << 3) | (r << 16); }
};

// ---- exact pivot choice (rare path): LAPACK's idamax = first maximum of |a| in the current (swapped) row order ----------
// v*: candidate values, c*: candidate flags, pos*: positions of the rows in LAPACK's order.  Returns lane | from2 << 5, and
// records dgetrf's info (first zero pivot, 1-based) when the whole column is zero.
__device__ __noinline__ unsigned pivot_exact(double v1, double v2, bool c1, bool c2, int pos1, int pos2, unsigned info_addr,
                                             int col1) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long m1 = c1 ? ((unsigned long long)__double_as_longlong(v1) & 0x7fffffffffffffffull) : 0ull;
  const unsigned long long m2 = c2 ? ((unsigned long long)__double_as_longlong(v2) & 0x7fffffffffffffffull) : 0ull;
  const unsigned long long mm = m1 > m2 ? m1 : m2;
  const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(mm >> 32));
  const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(mm >> 32) == hi ? (unsigned)mm : 0u);
  const unsigned long long best = ((unsigned long long)hi << 32) | lo;
  if (best == 0ull && lane == 0) {              // zero column: dgetrf's info = first zero pivot (1-based)
    int cur;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(cur) : "r"(info_addr));
    if (cur == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(info_addr), "r"(col1) : "memory");
  }
  // lowest position among the exact maxima (for a zero column: the first candidate)
  const unsigned e1 = (c1 && m1 == best) ? (((unsigned)pos1 << 6) | lane) : 0xffffffffu;
  const unsigned e2 = (c2 && m2 == best) ? (((unsigned)pos2 << 6) | 32u | lane) : 0xffffffffu;
  const unsigned emin = __reduce_min_sync(0xffffffffu, e1 < e2 ? e1 : e2);
  return emin & 63u;
}

// ---- panel factorisation: one row per lane, implicit pivoting --------------------------------------
// a[k]: logical columns c0..c0+7 of the lane's row (a2: second register set, rows 32.. while more than 32 rows are in
// play).  act: the lane's row is still in play.  ch: step at which the row became the pivot row (-1 otherwise).  On return
// rows still in play hold the NEGATED multipliers.  The pivot row of step k is final when it is chosen (negated multipliers
// in columns < k, its U row in columns >= k): its lane writes it to row mu(k) of the stage tile (the diagonal block in pivot
// order, input of the inversion) together with -1/pivot, and every lane reads the entries it needs back from there -- one
// 16-byte broadcast load per column pair instead of two shuffles and two register moves per column.
// pos: position of the row in LAPACK's swapped order (tie-breaking only).
template <bool TWO>
__device__ __forceinline__ void panel_factor(double (&a)[8], double (&a2)[8], const bool act1, const bool act2, int& pos1,
                                             int& pos2, const int npiv, const int c0, int& ch1, int& ch2,
                                             const unsigned stage, const unsigned scaleu_addr, const unsigned info_addr) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned pref = 31u - lane;
  ch1 = -1; ch2 = -1;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < npiv) {
      const bool c1 = act1 && ch1 < 0;
      const bool c2 = TWO && act2 && ch2 < 0;
      // key = |a| (exponent + 15 mantissa bits) << 6 | set 1 before set 2 | low lanes first
      const unsigned h1 = ((unsigned)__double2hiint(a[k]) << 1) & 0xffffffc0u;
      const unsigned k1 = c1 ? (h1 | 32u | pref) : 0u;
      unsigned key = k1, k2 = 0u;
      if (TWO) {
        const unsigned h2 = ((unsigned)__double2hiint(a2[k]) << 1) & 0xffffffc0u;
        k2 = c2 ? (h2 | pref) : 0u;
        key = k1 > k2 ? k1 : k2;
      }
      const double nr1 = neg_rcp(a[k]);             // speculative -1/pivot, overlaps the reduction
      const double nr2 = TWO ? neg_rcp(a2[k]) : 0.0;
      const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
      unsigned q = 31u - (kmax & 31u);
      bool from2 = TWO && ((kmax & 32u) == 0u);
      bool slow = kmax < 64u;                       // all candidates below 2^-1017 (or none): decide exactly
      if (GHB_CW_EXACT) {
        // more than one candidate within 2^-15 of the maximum: resolve with the full magnitude
        unsigned tt = __ballot_sync(0xffffffffu, (k1 ^ kmax) < 64u);
        slow = slow || (tt & (tt - 1u)) != 0u;
        if (TWO) {
          const unsigned t2 = __ballot_sync(0xffffffffu, (k2 ^ kmax) < 64u);
          slow = slow || (t2 & (t2 - 1u)) != 0u || (tt != 0u && t2 != 0u);
        }
      }
      if (slow) {
        const unsigned r = pivot_exact(a[k], TWO ? a2[k] : 0.0, c1, c2, pos1, pos2, info_addr, c0 + k + 1);
        q = r & 31u;
        from2 = TWO && (r & 32u) != 0u;
      }
      const bool me1 = !from2 && lane == q, me2 = from2 && lane == q;
      const unsigned srow = stage + 64u * (unsigned)mu(k);
      const unsigned sscl = scaleu_addr + 8u * (unsigned)(7 - k);   // -1/pivot, reversed order (inv(U) substitution)
      if (me1) {
        ch1 = k;
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(srow + 16u * c, a[c], a[4 + c]);
        sts64(sscl, nr1);
      }
      if (TWO && me2) {
        ch2 = k;
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(srow + 16u * c, a2[c], a2[4 + c]);
        sts64(sscl, nr2);
      }
      __syncwarp();
      const double nrinv = lds64(sscl);
      if (GHB_CW_EXACT) {
        // dlaswp: the row at position c0+k trades places with the pivot row
        const int pq = __shfl_sync(0xffffffffu, from2 ? pos2 : pos1, q);
        if (pos1 == c0 + k) pos1 = pq;
        if (TWO && pos2 == c0 + k) pos2 = pq;
      }
      const bool u1 = c1 && !me1, u2 = c2 && !me2;
      double m1v = 0.0, m2v = 0.0;                  // negated multipliers (dgetf2 scales by the reciprocal of the pivot)
      if (u1) { m1v = a[k] * nrinv; a[k] = m1v; }
      if (TWO && u2) { m2v = a2[k] * nrinv; a2[k] = m2v; }
      double pr[8];                                 // pivot row, logical columns > k
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c > k || 4 + c > k) lds128(srow + 16u * c, pr[c], pr[4 + c]);
#pragma unroll
      for (int j = k + 1; j < 8; ++j) {
        a[j] = fma(m1v, pr[j], a[j]);
        if (TWO) a2[j] = fma(m2v, pr[j], a2[j]);
      }
    }
  }
}

// PAD: the shape-generic mode.  The instantiation is for a PADDED shape (NI a multiple of 8, NB = 40); the plan's real
// n_i <= NI, n_b <= NB arrive at run time.  The image is completed with an identity on the pad diagonal (pad rows are
// never pivots of real columns, pad columns have their own row as the only candidate: LAPACK's choices on the real part
// are unchanged), everything outside the real blocks reads as zero, and all record offsets come from the plan's
// (column, field) table -- any number of interior and skeleton fields, any touched mask.
// GEN: the records never exist in HBM.  The WPC warps of a CTA work in lockstep on a batch of WPC cells: all threads
// generate the WPC records of an affine family together (every table element is fetched once per batch and combined with
// the WPC coefficient vectors), each into the warp's private scratch record -- rewritten for every cell, so it lives in
// L2 -- and after one barrier every warp condenses its own cell from its scratch record exactly as from a resident one.
// Q4 (NB = 36 shapes whose A21 / A22 blocks, records and outputs are 32-byte aligned): the boundary row tiles 0..3 are
// INTERLEAVED -- tile m holds the rows 4n + m, n = 0..7 -- so that a lane's A21 fragments of the four tiles are four
// consecutive rows of one column (one 256-bit load instead of four 64-bit ones: 100 instead of 250 A21 loads per cell),
// and its D fragments of the four tiles are the eight consecutive rows 8t .. 8t+7 of a column of S (two 256-bit stores /
// A22 loads instead of four 128-bit ones).  Tile 4 keeps the rows 32 + n.
template <int NI, int NB, int WPC, int MINB, bool KEEPX, bool SPARSE, bool PAD, bool SCAT, bool GEN = false, bool BACK = false,
          bool Q4 = false>
__global__ void __launch_bounds__(32 * WPC, MINB) condense_cw_kernel(const CwArgs ar) {
  using C = CwCfg<NI, NB>;
  constexpr int RT = C::RT, NPL = C::NPL, DUMMY = C::DUMMY;
  constexpr int BTM = C::BTM;
  static_assert(!PAD || (SPARSE && NI % 8 == 0), "PAD kernels are instantiated for padded shapes");
  static_assert(!Q4 || (!PAD && !BACK && C::BTM == 5 && NB % 4 == 0 && NB < 40), "Q4: four interleaved row tiles + a short fifth");
  static_assert(!BACK || (!KEEPX && !SCAT), "BACK kernels: the backward map (resident records, or GEN: records formed in the loader)");
  static_assert(!GEN || WPC <= 8, "GEN kernels: at most 8 cells per batch (rows of a DMMA tile)");
  const int nir = PAD ? ar.n_i : NI, nbr = PAD ? ar.n_b : NB;       // real sizes
  const int NC = nbr + 1;                                           // right-hand-side columns: A12 | b1
  const int CTB = PAD ? (NC + 7) / 8 : C::CTB;
  const int btm = PAD ? (nbr + 7) / 8 : BTM;
  constexpr unsigned ROWB = C::ROWB;
  constexpr bool HAS2 = NI > 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  unsigned char* wsp = smem_raw + (size_t)warp * C::WARP_BYTES;
  const unsigned ws = (unsigned)__cvta_generic_to_shared(wsp);           // the warp's image (shared-window address)
  const unsigned a_invL = ws + C::OFF_INVL, a_scaleU = ws + C::OFF_SCALEU;
  const unsigned a_prow = ws + C::OFF_PROW, a_info = ws + C::OFF_INFO;
  unsigned char* shp = smem_raw + (size_t)WPC * C::WARP_BYTES;           // CTA-shared tables
  const unsigned a_sh = (unsigned)__cvta_generic_to_shared(shp);
  const unsigned a_ones = a_sh, a_rowA12 = a_sh + C::SH_ROWA12, a_rowb = a_sh + C::SH_ROWB;

  // one-time: zero the warp's region (pad columns, the dummy row and the inverse tiles stay zero where never written)
  for (unsigned i = lane; i < C::WARP_BYTES / 8; i += 32) reinterpret_cast<double*>(wsp)[i] = 0.0;
  for (int i = threadIdx.x; i < 8; i += 32 * WPC) reinterpret_cast<double*>(shp)[i] = 1.0;
  if (GEN && threadIdx.x == 0) {
    const unsigned bars = (unsigned)__cvta_generic_to_shared(shp) + (PAD ? C::SH_BAR_PAD : C::SH_BAR);
    for (int w = 0; w < 2 * WPC; ++w) mbar_init(bars + 8u * w, 1u);
    sts128(bars + 16u * WPC, 0.0, 0.0);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i <= NI; i += 32 * WPC) {
    reinterpret_cast<uint32_t*>(shp + C::SH_ROWA12)[i] = ar.rowA12[i];
    reinterpret_cast<uint16_t*>(shp + C::SH_ROWB)[i] = ar.rowb[i];
  }
  const int nf = ar.nf, ntot = nir + nbr;
  const unsigned a_colbase = a_sh + C::SH_COLBASE, a_rowinfo = a_sh + C::SH_ROWINFO;
  if (PAD) {
    for (int i = threadIdx.x; i < (ntot + 1) * nf; i += 32 * WPC) reinterpret_cast<int32_t*>(shp + C::SH_COLBASE)[i] = ar.colbase[i];
    for (int i = threadIdx.x; i < ntot; i += 32 * WPC) reinterpret_cast<uint16_t*>(shp + C::SH_ROWINFO)[i] = ar.rowinfo[i];
  }
  __syncthreads();

  // lane constants
  const unsigned G8 = 8u * (unsigned)mu(g);                  // image column offset of logical column g of a tile
  const unsigned T16 = 16u * (unsigned)t;
  const bool vl0 = t < NPL, vl1 = 4 + t < NPL;               // logical indices t, 4+t exist in the last panel
  const bool vb = !PAD && ((NB % 8 == 0) || (8 * (BTM - 1) + g < NB));   // boundary row 8(BTM-1)+g exists
  // PAD: field << 8 | local of this lane's boundary rows (8m + g: A21 fragments; 8m + 2t + e: A22 / b2), 0xffff: no such row
  unsigned rinfo21[PAD ? BTM : 1], rinfo22[PAD ? BTM : 1][2];
  if (PAD) {
#pragma unroll
    for (int m = 0; m < BTM; ++m) {
      const int r = 8 * m + g;
      rinfo21[m] = r < nbr ? reinterpret_cast<uint16_t*>(shp + C::SH_ROWINFO)[nir + r] : 0xffffu;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r2 = 8 * m + 2 * t + e;
        rinfo22[m][e] = r2 < nbr ? reinterpret_cast<uint16_t*>(shp + C::SH_ROWINFO)[nir + r2] : 0xffffu;
      }
    }
  }
  const unsigned encA = C::enc(lane < NI ? lane : DUMMY);
  const unsigned encB = C::enc((HAS2 && lane + 32 < NI) ? lane + 32 : DUMMY);
  // A21 fragments: record offset of element (boundary row 8m + g, interior column 8p + 4e + t) = cbk[p][e] + 8m
  int cbk[RT][2];
#pragma unroll
  for (int p = 0; p < RT; ++p)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 8 * p + 4 * e + t;
      const int v = PAD ? (k < nir ? k * nf : -1) : ar.colA21[k];      // PAD: row of the (column, field) table
      cbk[p][e] = PAD ? v : (v >= 0 ? v + g : -1);
    }
  const int lenA = ar.lenA, lenb = ar.lenb;
  const bool al16 = ar.al16 != 0;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  const int64_t wstride = (int64_t)gridDim.x * WPC;
  const int ntab = ar.ntab;
  double* const slot0 = GEN ? ar.scratch + (int64_t)blockIdx.x * WPC * ar.slot : nullptr;   // scratch records of this CTA
  // Records of a batch: out[w][e] = sum_t coef[w][t] T[t][e] for the WPC cells of the batch is a small GEMM and runs on
  // DMMA: rows = cells of the batch (8 rows; WPC of them used), k = tables (4 per step, zero-padded), columns = 8 record
  // elements per tile.  The element range is cut into chunks of gen_E elements that the warps take round-robin; a warp
  // brings the ntab table rows of its chunk in with TMA bulk copies (one lane per table issues) -- a two-deep staging ring laid over
  // its own (dead) image, the copies of its next chunk fly while it combines the current one, no CTA barrier inside the
  // phase -- holds the A operand (coefficients) in registers for the whole batch, and sends the D fragments (two
  // consecutive elements of one cell per lane) to the scratch records of all WPC cells with 16-byte stores.  DMMA
  // accumulates in ascending k from C: the sum runs in table order from 0.0 like expand_records_kernel (glue.cu),
  // zero-padded steps add +0 -- the records are bit-identical to its.
  const unsigned a_bar = a_sh + (PAD ? C::SH_BAR_PAD : C::SH_BAR) + 16u * (unsigned)warp;   // the warp's two staging barriers
  const unsigned a_zero = a_sh + (PAD ? C::SH_BAR_PAD : C::SH_BAR) + 16u * WPC;             // eight zero bytes (B fragments of zero-padded tables)
  unsigned gphase = 0u;                                      // parity of the two staging barriers
  auto gen_records = [&](const int64_t base) {
    const int E = ar.gen_E;                                  // elements per chunk, a multiple of 16
    const int RS = E + 4;                                    // row stride of the ring = 4 (mod 16): the B-fragment loads of a half-warp (4 tables x 4 elements) hit 16 distinct banks
    const int lenAp = ar.lenAp, lenbp = ar.lenbp;            // generated lengths: a padding element may follow the record
    const int nchA = (lenAp + E - 1) / E;
    const int nch = ar.gen_need ? ar.gen_nneed : nchA + (lenbp + E - 1) / E;   // entries of the chunk list
    const int KS = (ntab + 3) >> 2;
    double ca[4];                                            // A fragments: coef[cell g][4s + t]
#pragma unroll
    for (int s = 0; s < 4; ++s)
      ca[s] = (g < WPC && base + g < ar.ncells && 4 * s + t < ntab) ? __ldg(ar.coef + (base + g) * ntab + 4 * s + t) : 0.0;
    if (lane < WPC && base + wstride + lane < ar.ncells)     // the coefficients of the next batch: DRAM -> L2 now
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ar.coef + (base + wstride + lane) * ntab));
    auto issue = [&](const int kl, const int buf) {
      const int k = ar.gen_need ? (int)__ldg(ar.gen_need + kl) : kl;      // chunk of the record behind list entry kl
      const bool isA = k < nchA;
      const int len = isA ? lenAp : lenbp;
      const int e0 = (isA ? k : k - nchA) * E, cnt = min(E, len - e0);
      const double* T = (isA ? ar.TA : ar.Tb) + e0;
      const unsigned bar = a_bar + 8u * (unsigned)buf;
      if (lane == 0) mbar_expect_tx(bar, (unsigned)(ntab * cnt) * 8u);
      __syncwarp();
      if (lane < ntab) {                                     // lane tq brings in the row of table tq
        fence_proxy_async();                                 // the image was written through the generic proxy
        bulk_g2s(ws + (unsigned)((buf * ntab + lane) * RS) * 8u, T + (size_t)lane * len, (unsigned)cnt * 8u, bar);
      }
    };
    // the warp's own image is free as soon as its own cell is done: its first two chunks fly while it waits for the others
#ifdef GHB_CW_GEN_KO
    if (base == (int64_t)blockIdx.x * WPC || GHB_CW_GEN_KO == 2) {
#else
    {
#endif
      if (warp < nch) issue(warp, 0);
      if (warp + WPC < nch) issue(warp + WPC, 1);
    }
    __syncthreads();          // every warp is done with its previous cell: the scratch records are free
    double* const dlane = slot0 + (g < WPC ? g : 0) * ar.slot + 2 * t;   // this lane's D fragments: cell g, elements 2t, 2t+1
#ifdef GHB_CW_GEN_KO       // knock-out experiments (timing only, wrong results): 1: generate the first batch only,
                           // 2: no scratch stores after the first batch, 3: no TMA / DMMA after the first batch, stores only
    const bool ko_first = base == (int64_t)blockIdx.x * WPC;
    const bool ko_skip = GHB_CW_GEN_KO == 1 && !ko_first;
    const bool ko_nostore = GHB_CW_GEN_KO == 2 && !ko_first;
    const bool ko_notma = GHB_CW_GEN_KO == 3 && !ko_first;
#else
    const bool ko_skip = false, ko_nostore = false, ko_notma = false;
#endif
    for (int kl = warp, buf = 0; kl < nch && !ko_skip; kl += WPC, buf ^= 1) {
      const int k = ar.gen_need ? (int)__ldg(ar.gen_need + kl) : kl;
      const bool isA = k < nchA;
      const int e0 = (isA ? k : k - nchA) * E, cnt = min(E, (isA ? lenAp : lenbp) - e0);
      double* const dst = dlane + (isA ? 0 : lenAp) + e0;
      // B fragment of k-step s: T[4s + t][e0 + 8 tile + g]; the lanes of a zero-padded table read a zero with stride 0
      unsigned sbs[4], sst[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const bool tv = 4 * s + t < ntab;
        sbs[s] = tv ? ws + (unsigned)((buf * ntab + 4 * s + t) * RS + g) * 8u : a_zero;
        sst[s] = tv ? 64u : 0u;
      }
      if (!ko_notma) {
        mbar_wait(a_bar + 8u * (unsigned)buf, (gphase >> buf) & 1u);
        gphase ^= 1u << buf;
      }
      const int nfull = cnt >> 3;                            // whole tiles; a last partial tile goes the predicated way
#ifndef GHB_CW_GEN_TU
#define GHB_CW_GEN_TU 4
#endif
      constexpr int TU = GHB_CW_GEN_TU;                      // independent tiles in flight
      int tl0 = 0;
      for (; tl0 + TU <= nfull; tl0 += TU) {
        double d[TU][2];
#pragma unroll
        for (int u = 0; u < TU; ++u) { d[u][0] = 0.0; d[u][1] = 0.0; }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s >= KS) break;
          double bv[TU];
#pragma unroll
          for (int u = 0; u < TU; ++u) bv[u] = ko_notma ? 1.0 : lds64(sbs[s] + (unsigned)(tl0 + u) * sst[s]);
          if (!ko_notma) {
#pragma unroll
            for (int u = 0; u < TU; ++u) dmma(d[u][0], d[u][1], ca[s], bv[u]);
          } else {
#pragma unroll
            for (int u = 0; u < TU; ++u) d[u][0] += bv[u] * ca[s];
          }
        }
        if (g < WPC && !ko_nostore) {
          double* dp = dst + 8 * tl0;
#pragma unroll
          for (int u = 0; u < TU; ++u) {
#if defined(GHB_CW_GEN_STORE) && GHB_CW_GEN_STORE == 1      // A/B: L2-only stores (no L1 write-through lookup)
            __stcg(reinterpret_cast<double2*>(dp + 8 * u), make_double2(d[u][0], d[u][1]));
#elif defined(GHB_CW_GEN_STORE) && GHB_CW_GEN_STORE == 2    // A/B: write-through hint
            __stwt(reinterpret_cast<double2*>(dp + 8 * u), make_double2(d[u][0], d[u][1]));
#else
            *reinterpret_cast<double2*>(dp + 8 * u) = make_double2(d[u][0], d[u][1]);
#endif
          }
        }
      }
      for (; 8 * tl0 < cnt; ++tl0) {                         // the last tiles of the chunk, one at a time
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s >= KS) break;
          const double bv = 8 * tl0 + g < cnt ? lds64(sbs[s] + (unsigned)tl0 * sst[s]) : 0.0;
          dmma(d0, d1, ca[s], bv);
        }
        if (g < WPC && 8 * tl0 + 2 * t < cnt) *reinterpret_cast<double2*>(dst + 8 * tl0) = make_double2(d0, d1);
      }
      __syncwarp();                                          // the warp is done with this buffer
      if (kl + 2 * WPC < nch && !ko_notma) issue(kl + 2 * WPC, buf);
    }
    __syncthreads();          // the WPC records of the batch are complete
    // the staging ring overwrote the image: restore the zeros the cell code relies on (pad columns, dummy row, inverse tiles)
    for (unsigned o = 16u * lane; o < C::WARP_BYTES; o += 512u) sts128(ws + o, 0.0, 0.0);
    __syncwarp();
  };
  for (int64_t cell = (int64_t)blockIdx.x * WPC + warp; GEN ? (cell - warp < ar.ncells) : (cell < ar.ncells); cell += wstride) {
    if (GEN) {
      gen_records(cell - warp);   // starts and ends with a barrier: the WPC records of the batch are complete
      if (cell >= ar.ncells) continue;
    }
    const double* Arec = GEN ? slot0 + warp * ar.slot : ar.A + cell * lenA;
    const double* brec = GEN ? Arec + ar.lenAp : ar.b + cell * lenb;
    // ------------------------------------------------------------------ load A11 (table-driven 8-byte cp.async)
    // the table entries of a batch are fetched before its copies are issued (the registers are free at this point)
    {
      const char* Ab = reinterpret_cast<const char*>(Arec);
      constexpr int NLD = (NI * NI + 31) / 32, LB = 20;
#pragma unroll
      for (int i0 = 0; i0 < NLD; i0 += LB) {
        uint2 e[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i)
          if (i0 + i < NLD) e[i] = __ldg(ar.ldtab + 32 * (i0 + i) + lane);
#pragma unroll
        for (int i = 0; i < LB; ++i)
          if (i0 + i < NLD) {
            if (SPARSE) {
              const bool z = e[i].x == 0xffffffffu;
              cp_async8_z(ws + e[i].y, Ab + (z ? 0u : e[i].x), z ? 0u : 8u);
            } else {
              cp_async8_z(ws + e[i].y, Ab + e[i].x, 8u);
            }
          }
      }
      if (!GEN && GHB_CW_PREFETCH == 1) { if (lane == 0) l2_prefetch(Arec, (unsigned)lenA * 8u); }
      if (!GEN && GHB_CW_PREFETCH == 2) { if (lane == 0) l2_prefetch(Arec + ar.pf12_off, (unsigned)ar.pf12_len * 8u); }
    }
    // positions = identity, info = 0
    if (8 * RT >= 32 || lane < 8 * RT) sts_u32(a_prow + 4u * lane, encA);
    if (lane + 32 < 8 * RT) sts_u32(a_prow + 4u * (lane + 32), encB);
    if (lane == 0) sts_u32(a_info, 0u);
    cp_async_wait_all();
    if (PAD && nir + lane < NI) {                            // identity on the pad diagonal
      const unsigned r = (unsigned)(nir + lane);
      sts64(ws + r * ROWB + ((8u * (8u * (r >> 3) + (unsigned)mu(r & 7))) ^ ((r & 6u) << 3)), 1.0);
    }
    __syncwarp();

    // ------------------------------------------------------------------ LU of A11, left-looking by column tiles
    int myrow = lane < NI ? lane : -1;                       // row of register set 1 (-1: lane retired)
    int myrow2 = (HAS2 && lane + 32 < NI) ? lane + 32 : -1;
    int pos1 = lane, pos2 = lane + 32;
    unsigned enc1 = encA, enc2 = encB;
    bool two = HAS2;
#pragma unroll 1
    for (int R = 0; R < RT; ++R) {
      const int c0 = 8 * R;
      const int npiv = (R == RT - 1) ? NPL : 8;
      if (R > 0) {
        // ---- bring column tile R up to date with the panels q < R (registers only), store it back
        unsigned ea[RT][2];                                  // image addresses of this lane's tile elements
        double T[RT][2];
        const unsigned wsc = ws + 64u * (unsigned)R;
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          unsigned w0, w1;
          lds_2u32(a_prow + 32u * j + 8u * (unsigned)t, w0, w1);
          ea[j][0] = wsc + ((w0 & 0xffffu) ^ G8);
          ea[j][1] = wsc + ((w1 & 0xffffu) ^ G8);
        }
#pragma unroll
        for (int j = 0; j < RT; ++j) { T[j][0] = lds64(ea[j][0]); T[j][1] = lds64(ea[j][1]); }
        unsigned rb[RT];                                     // B-fragment rows: position 8i + g, chunk t (swizzled)
#pragma unroll
        for (int i = 1; i < RT; ++i) rb[i] = ws + ((lds_u32(a_prow + 32u * i + 4u * (unsigned)g) & 0xffffu) ^ T16);
#pragma unroll
        for (int q = 0; q < RT - 1; ++q) {
          if (q < R) {
            double b0, b1;
            lds128(a_invL + 512u * q + 64u * g + T16, b0, b1);
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
          const unsigned fm = j < R ? 0x80000000u : 0u;      // rows of earlier pivots: final U entries, stored negated
          sts64(ea[j][0], flip_if(T[j][0], fm));
          sts64(ea[j][1], flip_if(T[j][1], fm));
        }
        __syncwarp();
      }
      // ---- panel: one row per lane; registers in logical column order (image chunk c holds logical columns c, 4 + c)
      double a[8], a2[8];
      const bool act1 = myrow >= 0, act2 = HAS2 && two && myrow2 >= 0;
      const unsigned r1 = act1 ? (unsigned)myrow : (unsigned)DUMMY, r2 = act2 ? (unsigned)myrow2 : (unsigned)DUMMY;
      const unsigned pa1 = ws + r1 * ROWB + 64u * R, pa2 = ws + r2 * ROWB + 64u * R;
      const unsigned sw1 = (r1 >> 1) & 3u, sw2 = (r2 >> 1) & 3u;
#pragma unroll
      for (int c = 0; c < 4; ++c) lds128(pa1 + (((unsigned)c ^ sw1) << 4), a[c], a[4 + c]);
      if (HAS2 && two) {
#pragma unroll
        for (int c = 0; c < 4; ++c) lds128(pa2 + (((unsigned)c ^ sw2) << 4), a2[c], a2[4 + c]);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) a2[c] = 0.0;
      }
      const unsigned stL = a_invL + 512u * R;   // stage tile of this panel, then inv(L_pp)
      if (npiv < 8) {                         // partial last panel: unused stage rows, scales and positions are zero / dummy
        sts128(stL + 16u * lane, 0.0, 0.0);
        if (lane < 8) {
          sts64(a_scaleU + 8u * lane, 0.0);
          sts_u32(a_prow + 32u * R + 4u * lane, C::enc(DUMMY));
        }
        __syncwarp();
      }
      int ch1, ch2;
      if (HAS2 && two) panel_factor<true>(a, a2, act1, act2, pos1, pos2, npiv, c0, ch1, ch2, stL, a_scaleU, a_info);
      else panel_factor<false>(a, a2, act1, false, pos1, pos2, npiv, c0, ch1, ch2, stL, a_scaleU, a_info);
      // write back the rows still in play (negated multipliers); the pivot rows are in the stage tile
      if (act1 && ch1 < 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(pa1 + (((unsigned)c ^ sw1) << 4), a[c], a[4 + c]);
      }
      if (HAS2 && act2 && ch2 < 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(pa2 + (((unsigned)c ^ sw2) << 4), a2[c], a2[4 + c]);
      }
      // ---- positions: the pivots of this panel at mu(step), then the rows still in play (set 1 in lane order, then set 2)
      {
        const bool rest1 = act1 && ch1 < 0, rest2 = act2 && ch2 < 0;
        const unsigned mr1 = __ballot_sync(0xffffffffu, rest1);
        const unsigned lt = (1u << lane) - 1u;
        const unsigned pp = a_prow + 32u * R;
        if (ch1 >= 0) sts_u32(pp + 4u * (unsigned)(2 * (ch1 & 3) + (ch1 >> 2)), enc1);
        else if (rest1) sts_u32(pp + 32u + 4u * __popc(mr1 & lt), enc1);
        if (ch1 >= 0) myrow = -1;
        if (HAS2 && two) {
          const unsigned mr2 = __ballot_sync(0xffffffffu, rest2);
          if (ch2 >= 0) sts_u32(pp + 4u * (unsigned)(2 * (ch2 & 3) + (ch2 >> 2)), enc2);
          else if (rest2) sts_u32(pp + 32u + 4u * (__popc(mr1) + __popc(mr2 & lt)), enc2);
          if (ch2 >= 0) myrow2 = -1;
          // compaction: rows of the second register set move into retired lanes
          const unsigned fr = __ballot_sync(0xffffffffu, myrow < 0);
          const unsigned m2 = __ballot_sync(0xffffffffu, myrow2 >= 0);
          const int nmove = min(__popc(fr), __popc(m2));
          const int idx = __popc(fr & lt);
          const int src = (int)__fns(m2, 0, idx + 1) & 31;            // lane holding the idx-th remaining row of set 2
          const int v = __shfl_sync(0xffffffffu, myrow2, src);
          const int vp = __shfl_sync(0xffffffffu, pos2, src);
          const unsigned ve = __shfl_sync(0xffffffffu, enc2, src);
          if (((fr >> lane) & 1u) && idx < nmove) { myrow = v; pos1 = vp; enc1 = ve; }
          if (myrow2 >= 0 && __popc(m2 & lt) < nmove) myrow2 = -1;
          two = __any_sync(0xffffffffu, myrow2 >= 0);
        }
      }
      __syncwarp();
      // ---- inverses of the diagonal block: lanes 0-7 columns of inv(L_pp), lanes 8-15 columns of inv(U_pp).  The U lanes
      //      run the same recurrence on the block reversed in both directions (step i is row 7 - i, register m is column
      //      7 - m; mu(7 - k) = 7 - mu(k), so their stage offsets are 504 minus those of the L lanes).  inv(L_pp) replaces
      //      the stage tile; inv(U_pp) goes to the (dead) diagonal block of the pivot rows in the image.
      {
        const int cidx = lane & 7, grp = (lane >> 3) & 1;
        const unsigned sb = grp ? stL + 504u : stL;
        const int sg = grp ? -1 : 1;
        const unsigned sc = grp ? a_scaleU : a_ones;
        const double sgn = grp ? -1.0 : 1.0;
        double z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double d0 = (i == cidx) ? sgn : 0.0, d1 = 0.0;
#pragma unroll
          for (int m = 0; m < i; ++m) {
            unsigned ad;                                     // one IMAD (the compiler would emit an add and a predicated add)
            asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(ad) : "r"(sg), "r"(64 * mu(i) + 8 * mu(m)), "r"(sb));
            const double v = lds64(ad);
            if (m & 1) d1 = fma(v, z[m], d1); else d0 = fma(v, z[m], d0);
          }
          z[i] = (d0 + d1) * lds64(sc + 8u * i);             // grp 0: 1; grp 1: -1/u_ss of the row of this step
        }
        __syncwarp();                                        // every lane has read its stage rows
        const unsigned mc = 8u * (unsigned)(2 * (cidx & 3) + (cidx >> 2));
        if (lane < 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) sts64(stL + 64u * mu(i) + mc, z[i]);
        } else if (lane < 16) {
          // logical (row 7 - i, column 7 - cidx): image row of pivot position mu(7 - i), physical column 7 - mu(cidx)
          const unsigned colx = 56u - mc;
          const unsigned wsc = ws + 64u * (unsigned)R;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const unsigned w = lds_u32(a_prow + 32u * R + 4u * (unsigned)(7 - mu(i)));
            sts64(wsc + ((w & 0xffffu) ^ colx), z[i]);
          }
        }
      }
      __syncwarp();
    }

    // ------------------------------------------------------------------ phase B: column tiles of [A12 b1]
    if (!GEN && !BACK && GHB_CW_PREFETCH == 2 && lane == 0) {
      l2_prefetch(Arec + ar.pf21_off, (unsigned)ar.pf21_len * 8u);
      if (ar.pf22_len > 0) l2_prefetch(Arec + ar.pf22_off, (unsigned)ar.pf22_len * 8u);
    }
    if (!GEN && GHB_CW_PREFETCH == 3 && lane == 0 && cell + wstride < ar.ncells)   // the whole next record of this warp
      l2_prefetch(Arec + wstride * lenA, (unsigned)lenA * 8u);
    const int failed = (int)lds_u32(a_info);
    // per position of this lane's tile rows (8j + 2t + e): record offset of A12(row, column g) and 8 x its column stride;
    // B-fragment rows of the L / U tiles
    int o[RT][2];
    int st8[RT][2];
    unsigned rb[RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      unsigned w0, w1;
      lds_2u32(a_prow + 32u * j + 8u * (unsigned)t, w0, w1);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const unsigned w = e ? w1 : w0;
        if (PAD) {                                           // field << 8 | local of the row at this position, -1: pad row
          const unsigned row = w >> 16;
          unsigned short si = 0xffffu;
          if (row < (unsigned)nir) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(si) : "r"(a_rowinfo + (row << 1)));
          o[j][e] = si == 0xffffu ? -1 : (int)si;
          st8[j][e] = 0;
          continue;
        }
        const unsigned ra = lds_u32(a_rowA12 + ((w >> 16) << 2));
        const bool ok = !(SPARSE && ra == 0xffffffffu) && !(j == RT - 1 && !(e ? vl1 : vl0));
        const int strd = (int)(ra >> 16);
        o[j][e] = ok ? (int)(ra & 0xffffu) + g * strd : -1;
        st8[j][e] = ok ? 8 * strd : 0;
      }
      rb[j] = ws + ((lds_u32(a_prow + 32u * j + 4u * (unsigned)g) & 0xffffu) ^ T16);
    }
    double* Sc = ar.S + cell * (int64_t)nbr * nbr;
    double* gc = ar.g + cell * (int64_t)nbr;
    const bool keepS = !SCAT || (ar.keepS != nullptr && ar.keepS[cell] != 0);   // fused mode: S_K only where it is needed

    // One pass handles NJ column tiles at a time: every B fragment (L / U tiles from shared memory, A21 from L2) is loaded
    // once and used for NJ independent DMMA chains.
    auto pass = [&](auto njc, const int J0) {
      constexpr int NJ = decltype(njc)::value;
      int col[NJ];                                           // column of [A12 b1] held by this lane's fragments, per tile
      bool cA[NJ];                                           // A12 column (else: the right-hand side, or padding)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) { col[jj] = 8 * (J0 + jj) + g; cA[jj] = col[jj] < nbr; }
      const double* tbase = Arec;
      if (PAD) {
        if (!cA[0]) tbase = brec;
      } else if (NJ == 1 && J0 == CTB - 1 && !cA[0]) {
        // the right-hand side b1 (padding columns repeat it; they are never stored); the last tile is always a single pass
        tbase = brec;
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          unsigned w0, w1;
          lds_2u32(a_prow + 32u * j + 8u * (unsigned)t, w0, w1);
          unsigned short h0, h1;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h0) : "r"(a_rowb + ((w0 >> 16) << 1)));
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h1) : "r"(a_rowb + ((w1 >> 16) << 1)));
          o[j][0] = (j == RT - 1 && !vl0) ? -1 : (int)h0;
          o[j][1] = (j == RT - 1 && !vl1) ? -1 : (int)h1;
        }
      }
      // ---- T = (P A12_J)^T straight from the record
      double T[NJ][RT][2];
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
        for (int j = 0; j < RT; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (j == RT - 1 && 4 * e >= NPL) { T[jj][j][e] = 0.0; continue; }
            if (PAD) {
              // (column, field) table: condensed column n_i + col of A12, or the right-hand side (column n)
              T[jj][j][e] = 0.0;
              if (o[j][e] >= 0 && col[jj] <= nbr) {
                const int ccol = cA[jj] ? nir + col[jj] : ntot;
                const int base = (int)lds_u32(a_colbase + 4u * (unsigned)(ccol * nf + (o[j][e] >> 8)));
                if (base >= 0) T[jj][j][e] = ldrec<GEN>(tbase + base + (o[j][e] & 0xff));
              }
              continue;
            }
            const int of = o[j][e] + jj * st8[j][e];
            if (SPARSE || j == RT - 1) T[jj][j][e] = o[j][e] >= 0 ? ldrec<GEN>(tbase + of) : 0.0;
            else T[jj][j][e] = ldrec<GEN>(tbase + of);
          }
#pragma unroll
      for (int j = 0; j < RT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) o[j][e] += NJ * st8[j][e];
      // ---- Z = L^-1 P A12_J
#pragma unroll
      for (int q = 0; q < RT; ++q) {
        double b0, b1;
        lds128(a_invL + 512u * q + 64u * g + T16, b0, b1);
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          double z0 = 0.0, z1 = 0.0;
          dmma(z0, z1, T[jj][q][0], b0);
          if (q < RT - 1 || NPL > 4) dmma(z0, z1, T[jj][q][1], b1);
          T[jj][q][0] = z0; T[jj][q][1] = z1;
        }
#pragma unroll
        for (int i = q + 1; i < RT; ++i) {
          double l0, l1;
          lds128(rb[i] + 64u * q, l0, l1);
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) {
            dmma(T[jj][i][0], T[jj][i][1], T[jj][q][0], l0);
            dmma(T[jj][i][0], T[jj][i][1], T[jj][q][1], l1);
          }
        }
      }
      // ---- X = U^-1 Z
#pragma unroll
      for (int q = RT - 1; q >= 0; --q) {
        double b0, b1;
        lds128(rb[q] + 64u * q, b0, b1);                     // inv(U_qq) sits in the diagonal block of the image
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          double x0 = 0.0, x1 = 0.0;
          dmma(x0, x1, T[jj][q][0], b0);
          if (q < RT - 1 || NPL > 4) dmma(x0, x1, T[jj][q][1], b1);
          T[jj][q][0] = x0; T[jj][q][1] = x1;
        }
#pragma unroll
        for (int p = q - 1; p >= 0; --p) {                   // off-diagonal U tiles are stored negated
          double u0, u1;
          lds128(rb[p] + 64u * q, u0, u1);
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) {
            dmma(T[jj][p][0], T[jj][p][1], T[jj][q][0], u0);
            if (q < RT - 1 || NPL > 4) dmma(T[jj][p][0], T[jj][p][1], T[jj][q][1], u1);
          }
        }
      }
      if (KEEPX) {
        // X = A11^-1 [A12 | b1], col-major n_i x (n_b+1) per cell
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj)
          if (col[jj] < NC) {
            double* Xc = ar.X + cell * (int64_t)(nir * NC) + (int64_t)col[jj] * nir;
#pragma unroll
            for (int p = 0; p < RT; ++p)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int k = 8 * p + 4 * e + t;
                if (8 * p + 4 * e < NI && k < nir && (p < RT - 1 || (e ? vl1 : vl0))) Xc[k] = failed ? qnan : T[jj][p][e];
              }
          }
      }
      // ---- S_J = A22_J - A21 X_J  (transposed tiles: acc[m][e] = S[8m + 2t + e][col])
      double acc[NJ][BTM][2];
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        const double* ini = cA[jj] ? ((SPARSE && ar.a22base < 0) ? nullptr : Arec + ar.a22base + col[jj] * NB) : brec + ar.b2base;
#pragma unroll
        for (int m = 0; m < BTM; ++m) {
          const int r = 8 * m + 2 * t;
          acc[jj][m][0] = 0.0; acc[jj][m][1] = 0.0;
          if (PAD) {
            if (m < btm && col[jj] <= nbr) {
              const int ccol = cA[jj] ? nir + col[jj] : ntot;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const unsigned ri = rinfo22[m][e];
                if (ri != 0xffffu) {
                  const int base = (int)lds_u32(a_colbase + 4u * (unsigned)(ccol * nf + (int)(ri >> 8)));
                  if (base >= 0) acc[jj][m][e] = ldrec<GEN>(tbase + base + (int)(ri & 0xffu));
                }
              }
            }
            continue;
          }
          if (Q4 && m < 4) continue;                         // tiles 0..3: interleaved rows, loaded below
          if (!SPARSE || ini != nullptr) {
            if (al16) {
              if (NB % 8 == 0 || r < NB) { const double2 v = ldrec2<GEN>(reinterpret_cast<const double2*>(ini + r)); acc[jj][m][0] = v.x; acc[jj][m][1] = v.y; }
            } else {
              if (NB % 8 == 0 || r < NB) acc[jj][m][0] = ldrec<GEN>(ini + r);
              if (NB % 8 == 0 || r + 1 < NB) acc[jj][m][1] = ldrec<GEN>(ini + r + 1);
            }
          }
        }
      }
      if (Q4) {
        // acc[m][e] = S[8t + 4e + m][col], m < 4: rows 8t .. 8t+7 of the column
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const double* ini = cA[jj] ? ((SPARSE && ar.a22base < 0) ? nullptr : Arec + ar.a22base + col[jj] * NB) : brec + ar.b2base;
          if (!SPARSE || ini != nullptr) {
            if (cA[jj]) {                                    // A22 column: 32-byte aligned
              ldrec4<GEN>(ini + 8 * t, acc[jj][0][0], acc[jj][1][0], acc[jj][2][0], acc[jj][3][0]);
              ldrec4<GEN>(ini + 8 * t + 4, acc[jj][0][1], acc[jj][1][1], acc[jj][2][1], acc[jj][3][1]);
            } else {                                         // b2 (the lanes of the right-hand-side column): 8-byte aligned
#pragma unroll
              for (int m = 0; m < 4; ++m) { acc[jj][m][0] = ldrec<GEN>(ini + 8 * t + m); acc[jj][m][1] = ldrec<GEN>(ini + 8 * t + 4 + m); }
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < RT; ++p) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (8 * p + 4 * e < NI) {                          // k-step with at least one real column
            const int of = cbk[p][e];
            const bool okc = (!SPARSE || of >= 0) && (p < RT - 1 || (e ? vl1 : vl0));
            const double* src = Arec + (okc ? of : 0);
            double bf[BTM];
            if (Q4) {                                        // rows 4g .. 4g+3 of the column in one load; tile 4: row 32 + g
              bf[0] = 0.0; bf[1] = 0.0; bf[2] = 0.0; bf[3] = 0.0;
              if (!(SPARSE || p == RT - 1) || okc) ldrec4<GEN>(src + 3 * g, bf[0], bf[1], bf[2], bf[3]);
              bf[4] = (okc && vb) ? ldrec<GEN>(src + 32) : 0.0;
            }
#pragma unroll
            for (int m = 0; m < BTM; ++m) {
              if (Q4) continue;
              if (PAD) {
                bf[m] = 0.0;
                if (of >= 0 && rinfo21[m] != 0xffffu) {
                  const int base = (int)lds_u32(a_colbase + 4u * (unsigned)(of + (int)(rinfo21[m] >> 8)));
                  if (base >= 0) bf[m] = ldrec<GEN>(Arec + base + (int)(rinfo21[m] & 0xffu));
                }
                continue;
              }
              if (SPARSE || p == RT - 1 || (m == BTM - 1 && NB % 8 != 0))
                bf[m] = (okc && (m < BTM - 1 || vb)) ? ldrec<GEN>(src + 8 * m) : 0.0;
              else
                bf[m] = ldrec<GEN>(src + 8 * m);
            }
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
              const double xa = flip(T[jj][p][e]);
#pragma unroll
              for (int m = 0; m < BTM; ++m)
                if (!PAD || m < btm) dmma(acc[jj][m][0], acc[jj][m][1], xa, bf[m]);
            }
          }
        }
      }
      // ---- fused assembly: add this column of S_K to its place in the CSC values (at most two cells contribute to an
      //      entry of a zeroed nzval, so the floating-point atomics are order-independent and bit-reproducible)
      if (SCAT) {
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj)
          if (cA[jj]) {
            const int64_t ce = cell * nbr + col[jj];
            const int64_t cpos = __ldg(ar.colpos + ce);
            if (cpos >= 0) {
              const uint8_t* rr = ar.rowrank + ce * nbr;
              double* nz = ar.nzval + cpos;
              if (Q4) {                                      // rows 8t + 4e + m: four ranks per 32-bit load
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const unsigned rk4 = __ldg(reinterpret_cast<const unsigned*>(rr + 8 * t + 4 * e));
#pragma unroll
                  for (int m = 0; m < 4; ++m) {
                    const unsigned rk = (rk4 >> (8 * m)) & 0xffu;
                    if (rk != 255u) atomicAdd(nz + (rk & 0x7fu), failed ? qnan : acc[jj][m][e]);
                  }
                }
              }
#pragma unroll
              for (int m = 0; m < BTM; ++m) {
                if (Q4 && m < 4) continue;
                const int r = 8 * m + 2 * t;
                if (r < nbr) {
                  unsigned rk0, rk1 = 255u;
                  if ((nbr & 1) == 0) { const unsigned v = __ldg(reinterpret_cast<const unsigned short*>(rr + r)); rk0 = v & 0xffu; rk1 = v >> 8; }
                  else { rk0 = __ldg(rr + r); if (r + 1 < nbr) rk1 = __ldg(rr + r + 1); }
                  // bit 7 of a rank marks the entries shared with the neighbour cell.  Storing the others instead of adding
                  // them measured SLOWER (48.2 vs 46.6 ms at 128^3: divergent store / atomic paths), so everything is added
                  const double v0 = failed ? qnan : acc[jj][m][0], v1 = failed ? qnan : acc[jj][m][1];
                  if (rk0 != 255u) atomicAdd(nz + (rk0 & 0x7fu), v0);
                  if (rk1 != 255u) atomicAdd(nz + (rk1 & 0x7fu), v1);
                }
              }
            }
          }
      }
      // ---- store
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj)
        if (col[jj] < NC && (keepS || !cA[jj])) {
          double* dst = cA[jj] ? Sc + (int64_t)col[jj] * nbr : gc;
          if (Q4) {                                          // rows 8t .. 8t+3 and 8t+4 .. 8t+7: 32-byte aligned in S and g
            if (failed) {
              strec4<GEN>(dst + 8 * t, qnan, qnan, qnan, qnan);
              strec4<GEN>(dst + 8 * t + 4, qnan, qnan, qnan, qnan);
            } else {
              strec4<GEN>(dst + 8 * t, acc[jj][0][0], acc[jj][1][0], acc[jj][2][0], acc[jj][3][0]);
              strec4<GEN>(dst + 8 * t + 4, acc[jj][0][1], acc[jj][1][1], acc[jj][2][1], acc[jj][3][1]);
            }
          }
#pragma unroll
          for (int m = 0; m < BTM; ++m) {
            if (Q4 && m < 4) continue;
            const int r = 8 * m + 2 * t;
            double v0 = acc[jj][m][0], v1 = acc[jj][m][1];
            if (failed) { v0 = qnan; v1 = qnan; }
            if (PAD) {
              if (r < nbr) dst[r] = v0;
              if (r + 1 < nbr) dst[r + 1] = v1;
              continue;
            }
            if (GEN) {       // streaming stores: S_K must not push the scratch records out of L2
              if (al16) {
                if (NB % 8 == 0 || r < NB) __stcs(reinterpret_cast<double2*>(dst + r), make_double2(v0, v1));
              } else {
                if (NB % 8 == 0 || r < NB) __stcs(dst + r, v0);
                if (NB % 8 == 0 || r + 1 < NB) __stcs(dst + r + 1, v1);
              }
              continue;
            }
            if (al16) {
              if (NB % 8 == 0 || r < NB) *reinterpret_cast<double2*>(dst + r) = make_double2(v0, v1);
            } else {
              if (NB % 8 == 0 || r < NB) dst[r] = v0;
              if (NB % 8 == 0 || r + 1 < NB) dst[r + 1] = v1;
            }
          }
        }
    };
    if (BACK) {
      // ---- backward map (/root/reference/src/BackwardStaticCondensationMap.jl:84-99): r = b1 - A12 lambda_K (gemv!),
      //      u_K = A11^-1 r (getrs! on the factors just computed).  The lane holds column g of every A12 tile: it
      //      accumulates A12(row, 8J + g) lambda(8J + g) over the tiles J for its rows, the eight lanes of equal t then
      //      add up their columns; every column of the right-hand-side tile carries the same r (only column 0 is stored).
      double racc[RT][2];
#pragma unroll
      for (int j = 0; j < RT; ++j) { racc[j][0] = 0.0; racc[j][1] = 0.0; }
      const int ctba = (nbr + 7) >> 3;
#pragma unroll 1
      for (int J = 0; J < ctba; ++J) {
        const int col = 8 * J + g;
        const bool cAl = col < nbr;
        double lamv = 0.0;
        if (cAl) {
          const int64_t id = __ldg(ar.ids + cell * nbr + col);
          lamv = id > 0 ? __ldg(ar.lam_free + (id - 1)) : ((id < 0 && ar.lam_dir) ? __ldg(ar.lam_dir + (-id - 1)) : 0.0);
        }
#pragma unroll
        for (int j = 0; j < RT; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (j == RT - 1 && 4 * e >= NPL) continue;
            double v = 0.0;
            if (PAD) {
              if (o[j][e] >= 0 && cAl) {
                const int base = (int)lds_u32(a_colbase + 4u * (unsigned)((nir + col) * nf + (o[j][e] >> 8)));
                if (base >= 0) v = ldrec<GEN>(Arec + base + (o[j][e] & 0xff));
              }
            } else {
              if (o[j][e] >= 0 && cAl) v = ldrec<GEN>(Arec + o[j][e]);
              o[j][e] += st8[j][e];
            }
            racc[j][e] = fma(-v, lamv, racc[j][e]);
          }
      }
      double T[RT][2];
#pragma unroll
      for (int j = 0; j < RT; ++j) {
        unsigned w0, w1;
        lds_2u32(a_prow + 32u * j + 8u * (unsigned)t, w0, w1);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double r = racc[j][e];
          r += __shfl_xor_sync(0xffffffffu, r, 4);
          r += __shfl_xor_sync(0xffffffffu, r, 8);
          r += __shfl_xor_sync(0xffffffffu, r, 16);
          const unsigned row = (e ? w1 : w0) >> 16;
          double bv = 0.0;
          if (!(j == RT - 1 && 4 * e >= NPL) && row < (unsigned)nir && !(j == RT - 1 && !(e ? vl1 : vl0))) {
            if (PAD) {
              unsigned short si;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(si) : "r"(a_rowinfo + (row << 1)));
              const int base = (int)lds_u32(a_colbase + 4u * (unsigned)(ntot * nf + (int)(si >> 8)));
              bv = ldrec<GEN>(brec + base + (int)(si & 0xffu));
            } else {
              unsigned short h;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(a_rowb + (row << 1)));
              bv = ldrec<GEN>(brec + (int)h);
            }
            T[j][e] = bv + r;
          } else {
            T[j][e] = 0.0;
          }
        }
      }
      // Z = L^-1 P r
#pragma unroll
      for (int q = 0; q < RT; ++q) {
        double b0, b1;
        lds128(a_invL + 512u * q + 64u * g + T16, b0, b1);
        double z0 = 0.0, z1 = 0.0;
        dmma(z0, z1, T[q][0], b0);
        if (q < RT - 1 || NPL > 4) dmma(z0, z1, T[q][1], b1);
        T[q][0] = z0; T[q][1] = z1;
#pragma unroll
        for (int i = q + 1; i < RT; ++i) {
          double l0, l1;
          lds128(rb[i] + 64u * q, l0, l1);
          dmma(T[i][0], T[i][1], T[q][0], l0);
          dmma(T[i][0], T[i][1], T[q][1], l1);
        }
      }
      // u = U^-1 Z
#pragma unroll
      for (int q = RT - 1; q >= 0; --q) {
        double b0, b1;
        lds128(rb[q] + 64u * q, b0, b1);
        double x0 = 0.0, x1 = 0.0;
        dmma(x0, x1, T[q][0], b0);
        if (q < RT - 1 || NPL > 4) dmma(x0, x1, T[q][1], b1);
        T[q][0] = x0; T[q][1] = x1;
#pragma unroll
        for (int p = q - 1; p >= 0; --p) {
          double u0, u1;
          lds128(rb[p] + 64u * q, u0, u1);
          dmma(T[p][0], T[p][1], T[q][0], u0);
          if (q < RT - 1 || NPL > 4) dmma(T[p][0], T[p][1], T[q][1], u1);
        }
      }
      if (g == 0) {
        double* uc = ar.u + cell * (int64_t)nir;
#pragma unroll
        for (int p = 0; p < RT; ++p)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 8 * p + 4 * e + t;
            if (8 * p + 4 * e < NI && k < nir && (p < RT - 1 || (e ? vl1 : vl0))) uc[k] = failed ? qnan : T[p][e];
          }
      }
    } else {
      int J = 0;
      if (GHB_CW_NJ == 2 && !PAD) {
#pragma unroll 1
        for (; J + 2 < CTB; J += 2) pass(std::integral_constant<int, 2>{}, J);   // the last tile is a single pass
      }
#pragma unroll 1
      for (; J < CTB; ++J) pass(std::integral_constant<int, 1>{}, J);
    }
    if (ar.info && lane == 0) ar.info[cell] = failed;
    if (failed) {
      // a singular cell leaves NaN / Inf in the pad columns and the dummy row: restore the zeros the next cell relies on
      __syncwarp();
      for (unsigned i = lane; i < C::WARP_BYTES / 8; i += 32) reinterpret_cast<double*>(wsp)[i] = 0.0;
    }
    __syncwarp();
  }
}

// the plan-dependent part of the kernel arguments (shared with condense_cw_gen.cu)
inline void cw_fill_args(const Plan& p, CwArgs& ar) {
  const int nf = p.nfields, fb = p.boundary[0] - 1;
  ar.ldtab = reinterpret_cast<const uint2*>(p.d_cw);
  ar.rowA12 = reinterpret_cast<const uint32_t*>(p.d_cw + p.cw_off[0]);
  ar.colA21 = reinterpret_cast<const int32_t*>(p.d_cw + p.cw_off[1]);
  ar.colbase = reinterpret_cast<const int32_t*>(p.d_cw + p.cw_off[2]);
  ar.rowb = reinterpret_cast<const uint16_t*>(p.d_cw + p.cw_off[3]);
  ar.rowinfo = reinterpret_cast<const uint16_t*>(p.d_cw + p.cw_off[4]);
  ar.n_i = p.n_i; ar.n_b = p.n_b; ar.nf = nf;
  ar.nld = p.cw_nld;
  ar.a22base = (int)p.block_offset[fb + nf * fb];
  ar.b2base = p.field_offset_b[fb];
  ar.al16 = (p.n_b % 2 == 0 && p.lenA % 2 == 0 && p.lenb % 2 == 0 && ar.b2base % 2 == 0 &&
             (ar.a22base < 0 || ar.a22base % 2 == 0)) ? 1 : 0;
  ar.pf12_off = p.cw_pf[0]; ar.pf12_len = p.cw_pf[1];
  ar.pf21_off = p.cw_pf[2]; ar.pf21_len = p.cw_pf[3];
  ar.pf22_off = p.cw_pf[4]; ar.pf22_len = p.cw_pf[5];
  ar.lenA = p.lenA; ar.lenb = p.lenb;
  ar.TA = nullptr; ar.Tb = nullptr; ar.coef = nullptr; ar.scratch = nullptr; ar.slot = 0; ar.ntab = 0;
  ar.lenAp = 0; ar.lenbp = 0; ar.gen_E = 0; ar.gen_need = nullptr; ar.gen_nneed = 0;
  ar.lam_free = nullptr; ar.lam_dir = nullptr; ar.ids = nullptr; ar.u = nullptr;
}

}  // namespace

}  // namespace ghb

// Shapes the kernel is instantiated for (condense_cw.cu; condense_cw_gen.cu: the same shapes with generated records)
#ifdef GHB_CW_MORE_SHAPES   // experiments: shapes that have other tuned kernels
#define GHB_CW_SHAPES(X) X(34, 36) X(33, 12) X(40, 36) X(21, 16) X(56, 16) X(16, 8)
#else
#define GHB_CW_SHAPES(X) X(34, 36) X(33, 12) X(40, 36) X(21, 16)
#endif
// padded classes of the shape-generic (PAD) kernels: n_i <= NI, n_b <= 40
#define GHB_CW_PAD_NB 40
#define GHB_CW_PAD_CLASSES(X) X(16) X(24) X(32) X(40) X(48) X(56) X(64)
