// condense_cw_gen_pad.cu -- GEN instantiations (records of an affine family formed in the loader, condense_cw_gen.cuh) of
// the shape-generic classes of the cell-warp kernel: any plan with n_i <= 64, n_b <= 40.  Own translation unit (compile time).
#include "condense_cw_gen.cuh"

namespace ghb {

bool cw_gen_pad_fits(const Plan& p, int ntab) {
#define X(a) if (p.cw_pad == a) return gen_fits<a, GHB_CW_PAD_NB>(ntab);
  GHB_CW_PAD_CLASSES(X)
#undef X
  return false;
}

int launch_condense_cw_gen_pad(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                               const double* coef, double* S, double* g, int32_t* info, double* X, const ScatterArgs* sc) {
  CwArgs ar;
  GHB_TRY(gen_args_condense(ctx, p, ncells, ntab, TA, Tb, coef, S, g, info, X, sc, ar));
#define X(a)                                                                                                            \
  if (p.cw_pad == a) {                                                                                                  \
    if (ar.nzval) return launch_cw_gen<a, GHB_CW_PAD_NB, true, true, false, false, true>(ctx, p, ar);                   \
    if (ar.X) return launch_cw_gen<a, GHB_CW_PAD_NB, true, false, false, false, true, true>(ctx, p, ar);                \
    return launch_cw_gen<a, GHB_CW_PAD_NB, true, false, false, false, true>(ctx, p, ar);                                \
  }
  GHB_CW_PAD_CLASSES(X)
#undef X
  return fail(ctx, GHB_EUNSUPPORTED, "condense_cw<GEN>: class not instantiated");
}

int launch_backsub_cw_gen_pad(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                              const double* coef, const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                              int32_t* info) {
  CwArgs ar;
  GHB_TRY(gen_args_backsub(ctx, p, ncells, ntab, TA, Tb, coef, lam_free, lam_dir, ids, u, info, ar));
#define X(a) if (p.cw_pad == a) return launch_cw_gen<a, GHB_CW_PAD_NB, true, false, false, true, true>(ctx, p, ar);
  GHB_CW_PAD_CLASSES(X)
#undef X
  return fail(ctx, GHB_EUNSUPPORTED, "backsub_cw<GEN>: class not instantiated");
}

}  // namespace ghb
