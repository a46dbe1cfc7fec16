// condense_cw_gen.cu -- the cell-warp kernel with the element records of an affine family generated in its loader
// (/root/reference/src/GridapAPIExtensions.jl:442-451 and src/SumFacetsMap.jl:19-30 produce the cell
// matrices the reference condenses; here A_K = sum_t coef[K][t] TA[t] is formed per batch of cells inside the
// condensation kernel and never written to HBM).  Own translation units: the GEN instantiations of the tuned shapes here,
// those of the shape-generic classes in condense_cw_gen_pad.cu, in parallel with the resident-record ones of condense_cw.cu.
#include "condense_cw_gen.cuh"

namespace ghb {

bool cw_gen_supported(const Plan& p, int ntab) {
  if (!p.use_cw || ntab < 1 || ntab > 16) return false;
  if (p.cw_pad) return cw_gen_pad_fits(p, ntab);
#define X(a, b) if (p.n_i == a && p.n_b == b) return gen_fits<a, b>(ntab);
  GHB_CW_SHAPES(X)
#undef X
  return false;
}

template <int NI, int NB>
static int launch_cw_gen_shape(ghb_ctx* ctx, const Plan& p, CwArgs& ar) {
#ifdef GHB_CW_Q4_BUILD   // measured slower (profiles/r02_cw_summary.md): not built by default
  if constexpr (NB == 36) {
    // 256-bit accesses: the scratch record keeps the block offsets of the plan and is 128-byte aligned
    if (p.cw_q4 && p.opt.cw_q4 && ar.lenAp == p.lenA && !ar.X) {
      if (ar.nzval) {
        if (p.all_touched) return launch_cw_gen<NI, NB, false, true, true>(ctx, p, ar);
        return launch_cw_gen<NI, NB, true, true, true>(ctx, p, ar);
      }
      if (p.all_touched) return launch_cw_gen<NI, NB, false, false, true>(ctx, p, ar);
      return launch_cw_gen<NI, NB, true, false, true>(ctx, p, ar);
    }
  }
#endif
  if (ar.nzval) {
    if (p.all_touched) return launch_cw_gen<NI, NB, false, true>(ctx, p, ar);
    return launch_cw_gen<NI, NB, true, true>(ctx, p, ar);
  }
  if (ar.X) {       // keep_factors: X = A11^-1 [A12 | b1] is one extra store of the cell code
    if (p.all_touched) return launch_cw_gen<NI, NB, false, false, false, false, false, true>(ctx, p, ar);
    return launch_cw_gen<NI, NB, true, false, false, false, false, true>(ctx, p, ar);
  }
  if (p.all_touched) return launch_cw_gen<NI, NB, false, false>(ctx, p, ar);
  return launch_cw_gen<NI, NB, true, false>(ctx, p, ar);
}

int launch_condense_cw_gen(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                           const double* coef, double* S, double* g, int32_t* info, double* X, const ScatterArgs* sc) {
  if (ntab < 1 || ntab > 16) return fail(ctx, GHB_EINVAL, "condense_cw<GEN>: need 1 <= ntab <= 16");
  if (!cw_gen_supported(p, ntab)) return fail(ctx, GHB_EUNSUPPORTED, "condense_cw<GEN>: plan without a cell-warp kernel that can stage the tables");
  if (X && sc) return fail(ctx, GHB_EINVAL, "condense_cw<GEN>: stored factors and the fused scatter exclude each other");
  if (p.cw_pad) return launch_condense_cw_gen_pad(ctx, p, ncells, ntab, TA, Tb, coef, S, g, info, X, sc);
  CwArgs ar;
  GHB_TRY(gen_args_condense(ctx, p, ncells, ntab, TA, Tb, coef, S, g, info, X, sc, ar));
#define X(a, b) if (p.n_i == a && p.n_b == b) return launch_cw_gen_shape<a, b>(ctx, p, ar);
  GHB_CW_SHAPES(X)
#undef X
  return fail(ctx, GHB_EUNSUPPORTED, "condense_cw<GEN>: shape not instantiated");
}

// backward map with the records formed in the loader: u_K = A11^-1 (b1 - A12 lambda_K) straight from the coefficient
// vectors -- with ghb_condense_assemble_affine_f64 the whole solve runs without the records ever existing in HBM
int launch_backsub_cw_gen(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                          const double* coef, const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                          int32_t* info) {
  if (ntab < 1 || ntab > 16) return fail(ctx, GHB_EINVAL, "backsub_cw<GEN>: need 1 <= ntab <= 16");
  if (!cw_gen_supported(p, ntab)) return fail(ctx, GHB_EUNSUPPORTED, "backsub_cw<GEN>: plan without a cell-warp kernel that can stage the tables");
  if (p.cw_pad) return launch_backsub_cw_gen_pad(ctx, p, ncells, ntab, TA, Tb, coef, lam_free, lam_dir, ids, u, info);
  CwArgs ar;
  GHB_TRY(gen_args_backsub(ctx, p, ncells, ntab, TA, Tb, coef, lam_free, lam_dir, ids, u, info, ar));
#define X(a, b) \
  if (p.n_i == a && p.n_b == b) \
    return p.all_touched ? launch_cw_gen<a, b, false, false, false, true>(ctx, p, ar) : launch_cw_gen<a, b, true, false, false, true>(ctx, p, ar);
  GHB_CW_SHAPES(X)
#undef X
  return fail(ctx, GHB_EUNSUPPORTED, "backsub_cw<GEN>: shape not instantiated");
}

}  // namespace ghb
