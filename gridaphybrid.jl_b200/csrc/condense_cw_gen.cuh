// condense_cw_gen.cuh -- launch machinery shared by the GEN translation units (condense_cw_gen.cu: tuned shapes,
// condense_cw_gen_pad.cu: shape-generic classes): scratch records, chunk length, chunk list of the backward map, padded tables.
#pragma once
#include <vector>

#include "condense_cw_kernel.cuh"

#ifndef GHB_CW_GEN_WPC
#define GHB_CW_GEN_WPC 8      // cells per batch = warps per CTA of the GEN kernels (rows of the DMMA tile that forms the records)
#endif
#ifndef GHB_CW_GEN_MINB
#define GHB_CW_GEN_MINB (16 / GHB_CW_GEN_WPC)
#endif

namespace ghb {

namespace {

// a cell-warp plan whose image can stage two chunks of ntab table rows (E >= 16 elements)
template <int NI, int NB>
static bool gen_fits(int ntab) { return (size_t)CwCfg<NI, NB>::WARP_BYTES / (2 * (size_t)ntab * 8) >= 20; }

template <int NI, int NB, bool SPARSE, bool SCAT, bool Q4 = false, bool BACK = false, bool PAD = false, bool KEEPX = false>
static int launch_cw_gen(ghb_ctx* ctx, const Plan& p, CwArgs& ar) {
  // tuned shapes: 2 CTAs of 8 warps; shape-generic classes: 8 warps per CTA where eight images fit the SM, else 4
  constexpr unsigned per8 = 8u * CwCfg<NI, NB>::WARP_BYTES + CwCfg<NI, NB>::SH_BAR_PAD + 16u * 8u + 16u + 1024u;
  constexpr unsigned per4 = 4u * CwCfg<NI, NB>::WARP_BYTES + CwCfg<NI, NB>::SH_BAR_PAD + 16u * 4u + 16u + 1024u;
  constexpr int WPC = !PAD ? GHB_CW_GEN_WPC : (per8 <= 233472u ? 8 : 4);
  constexpr int fitp = (int)(233472u / (WPC == 8 ? per8 : per4));
  constexpr int MINB = !PAD ? GHB_CW_GEN_MINB : (fitp < 1 ? 1 : (fitp > 16 / WPC ? 16 / WPC : fitp));
  auto kern = condense_cw_kernel<NI, NB, WPC, MINB, KEEPX, SPARSE, PAD, SCAT, true, BACK, Q4>;
  const size_t smem = CwCfg<NI, NB>::smem_bytes(WPC, PAD, true);
  static KernelSetup ks;
  int per_sm = 0;
  GHB_TRY(kernel_setup(ctx, p.opt, kern, 32 * WPC, smem, GHB_CW_CARVEOUT, ks, "condense_cw_kernel<GEN>", &per_sm));
  const int64_t want = (ar.ncells + WPC - 1) / WPC;
  const int64_t grid = std::min<int64_t>(want, (int64_t)ctx->sm_count * per_sm);
  // one scratch record per resident warp, rewritten for every cell (L2-resident: 148 SMs x 16 warps x 39.8 kB = 94 MB on C3)
  ar.slot = ((int64_t)ar.lenAp + ar.lenbp + 15) & ~(int64_t)15;
  const size_t need = (size_t)grid * WPC * ar.slot * sizeof(double);
  if (ctx->gen_scratch_bytes < need) {
    if (ctx->gen_scratch) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->gen_scratch); }
    ctx->gen_scratch = nullptr; ctx->gen_scratch_bytes = 0;
    if (cudaMalloc((void**)&ctx->gen_scratch, need) != cudaSuccess) { cudaGetLastError(); return fail(ctx, GHB_ENOMEM, "condense_cw<GEN>: scratch records"); }
    ctx->gen_scratch_bytes = need;
  }
  ar.scratch = ctx->gen_scratch;
  // chunk of table elements a warp stages per TMA round: two buffers of ntab rows of E + 4 doubles inside its image
  const size_t cap = std::min<size_t>(1028, (size_t)CwCfg<NI, NB>::WARP_BYTES / (2 * (size_t)ar.ntab * 8));
  if (cap < 20) return fail(ctx, GHB_EUNSUPPORTED, "condense_cw<GEN>: image too small to stage the tables");
  ar.gen_E = cap >= 36 ? (int)((cap - 4) & ~(size_t)31) : 16;   // whole groups of four 8-element tiles where the image allows
  if (BACK) {
    // the backward map reads A11, A12 and b1 only: the chunks of the record that hold nothing else are not generated
    const int E = ar.gen_E, nf = p.nfields;
    const int nchA = (ar.lenAp + E - 1) / E, nchb = (ar.lenbp + E - 1) / E;
    std::vector<uint16_t> need;
    auto interior = [&](int f) { return std::find(p.interior.begin(), p.interior.end(), f + 1) != p.interior.end(); };
    for (int k = 0; k < nchA; ++k) {
      const int64_t lo = (int64_t)k * E, hi = std::min<int64_t>(lo + E, p.lenA);
      bool hit = false;
      for (int fj = 0; fj < nf && !hit; ++fj)
        for (int fi = 0; fi < nf && !hit; ++fi) {
          const int64_t bo = p.block_offset[fi + nf * fj];
          if (bo < 0 || !interior(fi)) continue;                  // block rows of interior fields: A11 and A12
          hit = bo < hi && bo + (int64_t)p.ndofs[fi] * p.ndofs[fj] > lo;
        }
      if (hit) need.push_back((uint16_t)k);
    }
    for (int k = 0; k < nchb; ++k) {
      const int lo = k * E, hi = std::min(lo + E, p.lenb);
      bool hit = false;
      for (int f = 0; f < nf && !hit; ++f)
        hit = interior(f) && p.field_offset_b[f] < hi && p.field_offset_b[f] + p.ndofs[f] > lo;
      if (hit) need.push_back((uint16_t)(nchA + k));
    }
    if (!ctx->gen_need) GHB_CUDA(ctx, cudaMalloc((void**)&ctx->gen_need, 4096 * sizeof(uint16_t)));
    if (need.size() > 4096) return fail(ctx, GHB_EUNSUPPORTED, "backsub_cw<GEN>: chunk list too long");
    GHB_CUDA(ctx, cudaMemcpyAsync(ctx->gen_need, need.data(), need.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    ar.gen_need = ctx->gen_need; ar.gen_nneed = (int)need.size();
  }
  kern<<<(unsigned)grid, 32 * WPC, smem, ctx->stream>>>(ar);
  GHB_LAUNCHED(ctx);
  return GHB_OK;
}

// tables for the TMA copies: 16-byte aligned rows (odd record lengths or unaligned caller tables go through a zero-padded
// copy owned by the context: ntab rows of lenAp = lenA + (lenA & 1) doubles)
inline int gen_tables(ghb_ctx* ctx, const Plan& p, int ntab, const double* TA, const double* Tb, CwArgs& ar) {
  ar.lenAp = p.lenA + (p.lenA & 1); ar.lenbp = p.lenb + (p.lenb & 1);
  if (ar.lenAp != p.lenA || ar.lenbp != p.lenb || (((uintptr_t)TA | (uintptr_t)Tb) & 15)) {
    const size_t need = (size_t)ntab * (ar.lenAp + ar.lenbp) * sizeof(double);
    if (ctx->gen_tab_bytes < need) {
      if (ctx->gen_tab) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->gen_tab); }
      ctx->gen_tab = nullptr; ctx->gen_tab_bytes = 0;
      if (cudaMalloc((void**)&ctx->gen_tab, need) != cudaSuccess) { cudaGetLastError(); return fail(ctx, GHB_ENOMEM, "condense_cw<GEN>: padded tables"); }
      ctx->gen_tab_bytes = need;
    }
    double* pA = ctx->gen_tab;
    double* pb = pA + (size_t)ntab * ar.lenAp;
    GHB_CUDA(ctx, cudaMemsetAsync(ctx->gen_tab, 0, need, ctx->stream));
    GHB_CUDA(ctx, cudaMemcpy2DAsync(pA, (size_t)ar.lenAp * 8, TA, (size_t)p.lenA * 8, (size_t)p.lenA * 8, ntab, cudaMemcpyDeviceToDevice, ctx->stream));
    GHB_CUDA(ctx, cudaMemcpy2DAsync(pb, (size_t)ar.lenbp * 8, Tb, (size_t)p.lenb * 8, (size_t)p.lenb * 8, ntab, cudaMemcpyDeviceToDevice, ctx->stream));
    ar.TA = pA; ar.Tb = pb;
  } else {
    ar.TA = TA; ar.Tb = Tb;
  }
  return GHB_OK;
}

// kernel arguments of a GEN condensation / backward map (the launch fills scratch, chunk length and chunk list)
inline int gen_args_condense(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                             const double* coef, double* S, double* g, int32_t* info, double* X, const ScatterArgs* sc, CwArgs& ar) {
  cw_fill_args(p, ar);
  ar.nzval = sc ? sc->nzval : nullptr;
  ar.colpos = sc ? sc->colpos : nullptr;
  ar.rowrank = sc ? sc->rowrank : nullptr;
  ar.keepS = sc ? sc->keepS : nullptr;
  ar.ncells = ncells;
  ar.A = nullptr; ar.b = nullptr; ar.S = S; ar.g = g; ar.info = info; ar.X = X;
  ar.coef = coef; ar.ntab = ntab;
  return gen_tables(ctx, p, ntab, TA, Tb, ar);
}

inline int gen_args_backsub(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                            const double* coef, const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                            int32_t* info, CwArgs& ar) {
  cw_fill_args(p, ar);
  ar.nzval = nullptr; ar.colpos = nullptr; ar.rowrank = nullptr; ar.keepS = nullptr;
  ar.ncells = ncells;
  ar.A = nullptr; ar.b = nullptr; ar.S = nullptr; ar.g = nullptr; ar.info = info; ar.X = nullptr;
  ar.coef = coef; ar.ntab = ntab;
  ar.lam_free = lam_free; ar.lam_dir = lam_dir; ar.ids = ids; ar.u = u;
  return gen_tables(ctx, p, ntab, TA, Tb, ar);
}

}  // namespace

// the shape-generic classes (condense_cw_gen_pad.cu)
bool cw_gen_pad_fits(const Plan& p, int ntab);
int launch_condense_cw_gen_pad(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                               const double* coef, double* S, double* g, int32_t* info, double* X, const ScatterArgs* sc);
int launch_backsub_cw_gen_pad(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                              const double* coef, const double* lam_free, const double* lam_dir, const int64_t* ids, double* u,
                              int32_t* info);

}  // namespace ghb
