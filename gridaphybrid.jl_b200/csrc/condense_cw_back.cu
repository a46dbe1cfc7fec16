// condense_cw_back.cu -- BackwardStaticCondensationMap on the cell-warp kernel (BACK instantiations of
// condense_cw_kernel.cuh): u_K = A11^-1 (b1 - A12 lambda_K) with the LU recomputed per cell, as the reference does
// (/root/reference/src/BackwardStaticCondensationMap.jl:84-99).  One warp per cell, no barriers; serves every plan with a
// cell-warp kernel -- the tuned shapes and the shape-generic (PAD) classes, which had only the generic backward kernel.
#include "condense_cw_kernel.cuh"

namespace ghb {

template <int NI, int NB, bool SPARSE, bool PAD>
static int launch_cw_back(ghb_ctx* ctx, const Plan& p, CwArgs& ar) {
  constexpr int WPC = GHB_CW_WPC;
  constexpr int fit = (int)(233472u / (WPC * CwCfg<NI, NB>::WARP_BYTES + CwCfg<NI, NB>::SH_BYTES_PAD + 1024u));
  constexpr int MINB = PAD ? (fit < 1 ? 1 : (fit > 4 ? 4 : fit)) : GHB_CW_MINB;
  auto kern = condense_cw_kernel<NI, NB, WPC, MINB, false, SPARSE, PAD, false, false, true>;
  const size_t smem = CwCfg<NI, NB>::smem_bytes(WPC, PAD);
  static KernelSetup ks;
  int per_sm = 0;
  GHB_TRY(kernel_setup(ctx, p.opt, kern, 32 * WPC, smem, GHB_CW_CARVEOUT, ks, "condense_cw_kernel<BACK>", &per_sm));
  const int64_t want = (ar.ncells + WPC - 1) / WPC;
  const int64_t grid = std::min<int64_t>(want, (int64_t)ctx->sm_count * per_sm);
  kern<<<(unsigned)grid, 32 * WPC, smem, ctx->stream>>>(ar);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

int launch_backsub_cw(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, const double* lam_free,
                      const double* lam_dir, const int64_t* ids, double* u, int32_t* info) {
  if (!p.use_cw) return fail(ctx, GHB_EUNSUPPORTED, "backsub_cw: the plan has no cell-warp kernel");
  CwArgs ar;
  cw_fill_args(p, ar);
  ar.nzval = nullptr; ar.colpos = nullptr; ar.rowrank = nullptr; ar.keepS = nullptr;
  ar.ncells = ncells;
  ar.A = A; ar.b = b; ar.S = nullptr; ar.g = nullptr; ar.info = info; ar.X = nullptr;
  ar.lam_free = lam_free; ar.lam_dir = lam_dir; ar.ids = ids; ar.u = u;
  if (p.cw_pad) {
#define X(a) if (p.cw_pad == a) return launch_cw_back<a, GHB_CW_PAD_NB, true, true>(ctx, p, ar);
    GHB_CW_PAD_CLASSES(X)
#undef X
  } else {
#define X(a, b) \
  if (p.n_i == a && p.n_b == b) \
    return p.all_touched ? launch_cw_back<a, b, false, false>(ctx, p, ar) : launch_cw_back<a, b, true, false>(ctx, p, ar);
    GHB_CW_SHAPES(X)
#undef X
  }
  return fail(ctx, GHB_EUNSUPPORTED, "backsub_cw: shape not instantiated");
}

}  // namespace ghb
