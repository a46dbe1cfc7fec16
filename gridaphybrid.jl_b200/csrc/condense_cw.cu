// condense_cw.cu -- host side of the cell-warp kernel (condense_cw_kernel.cuh): plan tables, instantiations, launch.
#include "condense_cw_kernel.cuh"

namespace ghb {

// ---- host side ----------------------------------------------------------------------------------------
// Shapes the kernel is instantiated for: one boundary field, n_i <= 64, interior fields first in the condensed order
// (always true: Plan orders interior rows first).

static bool cw_shape(int ni, int nb) {
#define X(a, b) if (ni == a && nb == b) return true;
  GHB_CW_SHAPES(X)
#undef X
  return false;
}


static int cw_pad_class(int ni) {
#define X(a) if (ni <= a) return a;
  GHB_CW_PAD_CLASSES(X)
#undef X
  return 0;
}

// tuned instantiation: exact shape, one skeleton field
bool cw_supported(const Plan& p) {
  if (p.boundary.size() != 1 || p.lenA >= 65536 || p.lenb >= 65536) return false;
  for (int f = 0; f < p.nfields; ++f) if (p.ndofs[f] > 255) return false;
  return cw_shape(p.n_i, p.n_b);
}

// shape-generic instantiation: any fields / touched mask with n_i <= 64, n_b <= 40
int cw_pad_class_of(const Plan& p) { return cw_pad_class(p.n_i); }

bool cw_pad_supported(const Plan& p) {
  if (p.nfields > 8 || p.n_b > GHB_CW_PAD_NB || cw_pad_class(p.n_i) == 0 || p.lenb >= 65536) return false;
  for (int f = 0; f < p.nfields; ++f) if (p.ndofs[f] > 255) return false;
  return true;
}

const char* cw_kernel_name(const Plan& p) {
  if (!p.cw_pad) {
#define X(a, b) if (p.n_i == a && p.n_b == b) return "cw_" #a "_" #b;
    GHB_CW_SHAPES(X)
#undef X
  }
#define X(a) if (p.cw_pad == a) return "cw_pad_" #a;
  GHB_CW_PAD_CLASSES(X)
#undef X
  return "cw";
}

template <int NI, int NB>
static void cw_tables(const Plan& p, std::vector<uint2>& ld, std::vector<uint32_t>& rowA12, std::vector<uint16_t>& rowb,
                      std::vector<int32_t>& colA21) {
  using C = CwCfg<NI, NB>;
  const int nf = p.nfields, fb = p.boundary[0] - 1;
  // loader: interior element (r, c) -> image row r, tile c / 8, physical column mu(c % 8), chunk-swizzled
  for (int c = 0; c < NI; ++c)
    for (int r = 0; r < NI; ++r) {
      const bool real = r < p.n_i && c < p.n_i;              // PAD classes: everything outside the real block is zero
      const int64_t bo = real ? p.block_offset[p.row_field[r] + nf * p.row_field[c]] : -1;
      const uint32_t src = bo < 0 ? 0xffffffffu : (uint32_t)(8 * (bo + (int64_t)p.row_local[c] * p.ndofs[p.row_field[r]] + p.row_local[r]));
      const uint32_t colb = 8u * (uint32_t)(8 * (c / 8) + mu(c % 8));
      const uint32_t dst = (uint32_t)r * C::ROWB + (colb ^ (((uint32_t)r & 6u) << 3));
      ld.push_back(make_uint2(src, dst));
    }
  std::stable_sort(ld.begin(), ld.end(), [](const uint2& x, const uint2& y) { return x.x < y.x; });
  // padding of the last batch (< 32 entries, each with a slot of its own so that no two copies of a batch write the same
  // address): record element 0 (fully touched plans: their copies carry no zero-fill predicate) or zeros into the stage
  // tile of panel 0, which the first panel rewrites completely before anything reads it
  static_assert(NI >= 8, "a full first panel");
  for (uint32_t i = 0; ld.size() % 32; ++i)
    ld.push_back(make_uint2(p.all_touched ? 0u : 0xffffffffu, C::OFF_INVL + 8u * i));
  rowA12.assign(NI + 1, 0xffffffffu);
  rowb.assign(NI + 1, 0);
  for (int r = 0; r < NI && r < p.n_i; ++r) {
    const int fr = p.row_field[r];
    const int64_t bo = p.block_offset[fr + nf * fb];
    if (bo >= 0 && !p.cw_pad) rowA12[r] = (uint32_t)(bo + p.row_local[r]) | ((uint32_t)p.ndofs[fr] << 16);
    rowb[r] = (uint16_t)(p.field_offset_b[fr] + p.row_local[r]);
  }
  colA21.assign(8 * C::RT, -1);
  for (int c = 0; c < NI && c < p.n_i && !p.cw_pad; ++c) {
    const int64_t bo = p.block_offset[fb + nf * p.row_field[c]];
    if (bo >= 0) colA21[c] = (int32_t)(bo + (int64_t)p.row_local[c] * NB);
  }
}

int cw_prepare(ghb_ctx* ctx, Plan& p) {
  if (!p.d_cw) {
    std::vector<uint2> ld;
    std::vector<uint32_t> rowA12;
    std::vector<uint16_t> rowb;
    std::vector<int32_t> colA21;
    if (!p.cw_pad) {
#define X(a, b) if (p.n_i == a && p.n_b == b) cw_tables<a, b>(p, ld, rowA12, rowb, colA21);
      GHB_CW_SHAPES(X)
#undef X
    } else {
#define X(a) if (p.cw_pad == a) cw_tables<a, GHB_CW_PAD_NB>(p, ld, rowA12, rowb, colA21);
      GHB_CW_PAD_CLASSES(X)
#undef X
    }
    // PAD kernels: (condensed column, field) -> record offset of the field's first row in that column, and the
    // (field, local row) of every condensed row
    const int n = p.n, nf = p.nfields;
    std::vector<int32_t> colbase((size_t)(n + 1) * nf, -1);
    std::vector<uint16_t> rowinfo((size_t)n + 1, 0);
    for (int c = 0; c < n; ++c) {
      const int fc = p.row_field[c], lc = p.row_local[c];
      for (int f = 0; f < nf; ++f) {
        const int64_t bo = p.block_offset[f + nf * fc];
        if (bo >= 0) colbase[(size_t)c * nf + f] = (int32_t)(bo + (int64_t)lc * p.ndofs[f]);
      }
      rowinfo[c] = (uint16_t)((p.row_field[c] << 8) | p.row_local[c]);
    }
    for (int f = 0; f < nf; ++f) colbase[(size_t)n * nf + f] = p.field_offset_b[f];
    // one device block: ldtab | rowA12 | colA21 | colbase | rowb | rowinfo
    const size_t b0 = ld.size() * sizeof(uint2), b1 = rowA12.size() * 4, b2 = colA21.size() * 4, b3 = colbase.size() * 4,
                 b4 = ((rowb.size() * 2 + 3) & ~(size_t)3), b5 = rowinfo.size() * 2;
    GHB_CUDA(ctx, cudaMalloc((void**)&p.d_cw, b0 + b1 + b2 + b3 + b4 + b5));
    GHB_CUDA(ctx, cudaMemcpy(p.d_cw, ld.data(), b0, cudaMemcpyHostToDevice));
    GHB_CUDA(ctx, cudaMemcpy(p.d_cw + b0, rowA12.data(), b1, cudaMemcpyHostToDevice));
    GHB_CUDA(ctx, cudaMemcpy(p.d_cw + b0 + b1, colA21.data(), b2, cudaMemcpyHostToDevice));
    GHB_CUDA(ctx, cudaMemcpy(p.d_cw + b0 + b1 + b2, colbase.data(), b3, cudaMemcpyHostToDevice));
    GHB_CUDA(ctx, cudaMemcpy(p.d_cw + b0 + b1 + b2 + b3, rowb.data(), rowb.size() * 2, cudaMemcpyHostToDevice));
    GHB_CUDA(ctx, cudaMemcpy(p.d_cw + b0 + b1 + b2 + b3 + b4, rowinfo.data(), b5, cudaMemcpyHostToDevice));
    p.cw_nld = (int)ld.size();
    p.cw_off[0] = b0; p.cw_off[1] = b0 + b1; p.cw_off[2] = b0 + b1 + b2; p.cw_off[3] = b0 + b1 + b2 + b3;
    p.cw_off[4] = b0 + b1 + b2 + b3 + b4;
  }
  // Q4 kernels (256-bit accesses): every A21 / A22 block, the records and the outputs must be 32-byte aligned
  {
    const int nfq = p.nfields, fbq = p.boundary.empty() ? 0 : p.boundary[0] - 1;
    bool ok = !p.cw_pad && p.boundary.size() == 1 && p.n_b % 4 == 0 && p.lenA % 4 == 0;
    for (int f = 0; f < nfq && ok; ++f) {
      const int64_t bo = p.block_offset[fbq + nfq * f];      // (lambda, f): A21 blocks and A22
      if (bo >= 0 && bo % 4 != 0) ok = false;
    }
    p.cw_q4 = ok;
  }
  // bounding record ranges of the A12 / A21 / A22 blocks (L2 prefetches)
  const int nf = p.nfields;
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

template <int NI, int NB, bool KEEPX, bool SPARSE, bool PAD = false, bool SCAT = false, bool Q4 = false>
static int launch_cw(ghb_ctx* ctx, const Plan& p, CwArgs& ar) {
  constexpr int WPC = GHB_CW_WPC;
  // PAD classes: as many CTAs per SM as their shared memory allows (the register cap follows)
  constexpr int fit = (int)(233472u / (WPC * CwCfg<NI, NB>::WARP_BYTES + CwCfg<NI, NB>::SH_BYTES_PAD + 1024u));
  constexpr int MINB = PAD ? (fit < 1 ? 1 : (fit > 4 ? 4 : fit)) : GHB_CW_MINB;
  auto kern = condense_cw_kernel<NI, NB, WPC, MINB, KEEPX, SPARSE, PAD, SCAT, false, false, Q4>;
  const size_t smem = CwCfg<NI, NB>::smem_bytes(WPC, PAD);
  static KernelSetup ks;
  int per_sm = 0;
  GHB_TRY(kernel_setup(ctx, p.opt, kern, 32 * WPC, smem, GHB_CW_CARVEOUT, ks, "condense_cw_kernel", &per_sm));
  const int64_t want = (ar.ncells + WPC - 1) / WPC;
  const int64_t grid = std::min<int64_t>(want, (int64_t)ctx->sm_count * per_sm);
  kern<<<(unsigned)grid, 32 * WPC, smem, ctx->stream>>>(ar);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

template <int NI, int NB>
static int launch_cw_shape(ghb_ctx* ctx, const Plan& p, CwArgs& ar) {
#ifdef GHB_CW_Q4_BUILD   // measured slower (50.0 vs 51.6 M cells/s on C3, profiles/r02_cw_summary.md): not built by default
  if constexpr (NB == 36) {        // interleaved row tiles with 256-bit accesses where the plan's blocks are 32-byte aligned
    if (p.cw_q4 && p.opt.cw_q4) {
      if (ar.nzval) {
        if (p.all_touched) return launch_cw<NI, NB, false, false, false, true, true>(ctx, p, ar);
        return launch_cw<NI, NB, false, true, false, true, true>(ctx, p, ar);
      }
      if (ar.X) {
        if (p.all_touched) return launch_cw<NI, NB, true, false, false, false, true>(ctx, p, ar);
        return launch_cw<NI, NB, true, true, false, false, true>(ctx, p, ar);
      }
      if (p.all_touched) return launch_cw<NI, NB, false, false, false, false, true>(ctx, p, ar);
      return launch_cw<NI, NB, false, true, false, false, true>(ctx, p, ar);
    }
  }
#endif
  if (ar.nzval) {
    if (p.all_touched) return launch_cw<NI, NB, false, false, false, true>(ctx, p, ar);
    return launch_cw<NI, NB, false, true, false, true>(ctx, p, ar);
  }
  if (ar.X) {
    if (p.all_touched) return launch_cw<NI, NB, true, false>(ctx, p, ar);
    return launch_cw<NI, NB, true, true>(ctx, p, ar);
  }
  if (p.all_touched) return launch_cw<NI, NB, false, false>(ctx, p, ar);
  return launch_cw<NI, NB, false, true>(ctx, p, ar);
}

template <int NI>
static int launch_cw_pad(ghb_ctx* ctx, const Plan& p, CwArgs& ar) {
  if (ar.nzval) return launch_cw<NI, GHB_CW_PAD_NB, false, true, true, true>(ctx, p, ar);
  if (ar.X) return launch_cw<NI, GHB_CW_PAD_NB, true, true, true>(ctx, p, ar);
  return launch_cw<NI, GHB_CW_PAD_NB, false, true, true>(ctx, p, ar);
}

static int launch_condense_cw_impl(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                                   double* g, int32_t* info, double* X, const ScatterArgs* sc) {
  CwArgs ar;
  cw_fill_args(p, ar);
  ar.nzval = sc ? sc->nzval : nullptr;
  ar.colpos = sc ? sc->colpos : nullptr;
  ar.rowrank = sc ? sc->rowrank : nullptr;
  ar.keepS = sc ? sc->keepS : nullptr;
  ar.ncells = ncells;
  ar.A = A; ar.b = b; ar.S = S; ar.g = g; ar.info = info; ar.X = X;
  if (p.cw_pad) {
#define X(a) if (p.cw_pad == a) return launch_cw_pad<a>(ctx, p, ar);
    GHB_CW_PAD_CLASSES(X)
#undef X
  } else {
#define X(a, b) if (p.n_i == a && p.n_b == b) return launch_cw_shape<a, b>(ctx, p, ar);
    GHB_CW_SHAPES(X)
#undef X
  }
  return fail(ctx, GHB_EUNSUPPORTED, "condense_cw: shape not instantiated");
}

int launch_condense_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                       double* g, int32_t* info, double* X) {
  return launch_condense_cw_impl(ctx, p, ncells, A, b, S, g, info, X, nullptr);
}

int launch_condense_cw_scatter(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                               double* g, int32_t* info, const ScatterArgs& sc) {
  return launch_condense_cw_impl(ctx, p, ncells, A, b, S, g, info, nullptr, &sc);
}

}  // namespace ghb
