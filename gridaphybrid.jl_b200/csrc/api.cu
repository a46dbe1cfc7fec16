// api.cu -- C ABI entry points (include/ghb.h): context, block plans, argument staging, dispatch.
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "common.cuh"

namespace ghb {

int fail(ghb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// true for cudaHostAlloc'ed / cudaHostRegister'ed memory; false for pageable host memory
static bool is_pinned_host_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// device temporaries / events of one call: released on every exit path (stream-ordered free; events after a sync of the
// side streams so that no copy into a caller buffer is still in flight when an error is returned)
struct CallTmp {
  ghb_ctx* ctx;
  std::vector<void*> dev;
  std::vector<cudaEvent_t> ev;
  bool side_streams = false;
  explicit CallTmp(ghb_ctx* c) : ctx(c) {}
  cudaError_t alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e == cudaSuccess) dev.push_back(*p);
    return e;
  }
  cudaError_t event(cudaEvent_t* e) {
    cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    if (r == cudaSuccess) ev.push_back(*e);
    return r;
  }
  ~CallTmp() {
    if (side_streams) {
      if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
      if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
    }
    for (void* p : dev) cudaFreeAsync(p, ctx->stream);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};

// pageable -> pinned staging copy on a few host threads (one memcpy stream does not reach the PCIe rate)
static void host_copy_parallel(void* dst, const void* src, size_t bytes) {
  unsigned nt = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));   // 16 threads measured no better (host copy bandwidth)
  if (bytes < ((size_t)8 << 20)) nt = 1;
  if (nt == 1) { memcpy(dst, src, bytes); return; }
  std::vector<std::thread> th;
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  for (unsigned i = 0; i < nt; ++i) {
    const size_t o = std::min(bytes, (size_t)i * per), n = std::min(per, bytes - o);
    if (n) th.emplace_back([=] { memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, n); });
  }
  for (auto& t : th) t.join();
}

int launch_condense(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, double* S,
                    double* g, int32_t* info, double* X) {
  if (X && ctx->opt.factors_generic) return launch_condense_generic(ctx, p, ncells, A, b, S, g, info, X);   // A/B knob
  if (p.use_cw) return launch_condense_cw(ctx, p, ncells, A, b, S, g, info, X);
  if (p.use_dmma) return launch_condense_dmma(ctx, p, ncells, A, b, S, g, info, X);
  if (p.use_warp && X == nullptr) return launch_condense_warp(ctx, p, ncells, A, b, S, g, info);
  if (p.use_large) return launch_condense_large(ctx, p, ncells, A, b, S, g, info, X);
  return launch_condense_generic(ctx, p, ncells, A, b, S, g, info, X);
}

// backward map with the LU recomputed (as the reference does): kernel by plan
int launch_backsub(ghb_ctx* ctx, const Plan& p, int64_t ncells, const double* A, const double* b, const double* lam_free,
                   const double* lam_dir, const int64_t* ids, double* u, int32_t* info) {
  if (p.use_cw && ctx->opt.cw_back) return launch_backsub_cw(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  if (p.use_dmma) return launch_backsub_dmma(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  if (p.use_large) return launch_backsub_large(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  if (p.use_warp) return launch_backsub_warp(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
  return launch_backsub_generic(ctx, p, ncells, A, b, lam_free, lam_dir, ids, u, info);
}

// factor storage of a keep_factors condensation: X = A11^-1 [A12 | b1] plus the info[] of that condensation (ghb_backsub_f64
// with A = b = NULL reports it: a singular cell has NaN factors); `generation` lets a caller check the factors are its own
static int factor_storage(ghb_ctx* ctx, const Plan& p, int plan_id, int64_t ncells, double** X, int32_t** finfo) {
  const size_t need = (size_t)ncells * p.n_i * (p.n_b + 1) * sizeof(double) + (size_t)ncells * sizeof(int32_t);
  if (ctx->fac.bytes < need) {
    if (ctx->fac.d_X) cudaFree(ctx->fac.d_X);
    ctx->fac.d_X = nullptr; ctx->fac.bytes = 0;
    if (cudaMalloc((void**)&ctx->fac.d_X, need) != cudaSuccess) { cudaGetLastError(); return fail(ctx, GHB_ENOMEM, "factor storage"); }
    ctx->fac.bytes = need;
  }
  ctx->fac.plan_id = plan_id; ctx->fac.ncells = ncells; ctx->fac.generation++;
  ctx->fac.d_info = reinterpret_cast<int32_t*>(ctx->fac.d_X + (size_t)ncells * p.n_i * (p.n_b + 1));
  *X = ctx->fac.d_X;
  *finfo = ctx->fac.d_info;
  return GHB_OK;
}

static int64_t env_i64(const char* name, int64_t dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoll(e) : dflt;
}

// the only place the environment is read: once per context
static void options_from_env(Options& o) {
  o.force_generic = (int)env_i64("GHB_FORCE_GENERIC", o.force_generic);
  o.cw = (int)env_i64("GHB_CW", o.cw);
  o.dmma_ll = (int)env_i64("GHB_DMMA_LL", o.dmma_ll);
  o.factors_generic = (int)env_i64("GHB_FACTORS_GENERIC", o.factors_generic);
  o.max_ctas_per_sm = (int)env_i64("GHB_MAX_CTAS_PER_SM", o.max_ctas_per_sm);
  o.ll_ctas = (int)env_i64("GHB_LL_CTAS", o.ll_ctas);
  o.warp_one_cell = (int)env_i64("GHB_WARP_ONE_CELL", o.warp_one_cell);
  o.warp_two_rows = (int)env_i64("GHB_WARP_TWO_ROWS", o.warp_two_rows);
  o.debug = (int)env_i64("GHB_DEBUG", o.debug);
  o.fused_assembly = (int)env_i64("GHB_FUSED_ASSEMBLY", o.fused_assembly);
  o.cw_back = (int)env_i64("GHB_CW_BACK", o.cw_back);
  o.cw_q4 = (int)env_i64("GHB_CW_Q4", o.cw_q4);
  o.stream_chunk_bytes = std::max<int64_t>(1, env_i64("GHB_STREAM_CHUNK_BYTES", o.stream_chunk_bytes));
}

// ---- symbolic patterns are handles: ctx->as is the SELECTED pattern, the others wait in ctx->as_store -------------------
static void asm_stash(ghb_ctx* ctx) {
  if (ctx->as_id >= 0) { ctx->as_store[ctx->as_id] = ctx->as; ctx->as = AsmState(); ctx->as_id = -1; }
}
static int asm_new(ghb_ctx* ctx) {
  asm_stash(ctx);
  ctx->as_store.emplace_back();
  ctx->as_id = (int)ctx->as_store.size() - 1;
  ctx->as = AsmState();
  return ctx->as_id;
}

static Plan* get_plan(ghb_ctx* ctx, int id) {
  if (!ctx || id < 0 || id >= (int)ctx->plans.size()) return nullptr;
  return ctx->plans[id];
}

}  // namespace ghb

using namespace ghb;

extern "C" {

int ghb_create(int device_id, ghb_ctx** out) {
  if (!out) return GHB_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return GHB_ENODEVICE; }
  if (device_id < 0 || device_id >= ndev) return GHB_EINVAL;
  ghb_ctx* ctx = new (std::nothrow) ghb_ctx();
  if (!ctx) return GHB_ENOMEM;
  ctx->device = device_id;
  if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return GHB_ECUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete ctx; return GHB_ECUDA; }
  options_from_env(ctx->opt);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GHB_ECUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GHB_ECUDA; }
  // keep freed async allocations cached in the pool (avoid re-mapping per call)
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return GHB_OK;
}

void ghb_destroy(ghb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (Plan* p : ctx->plans) {
    if (p && p->d_emap) cudaFree(p->d_emap);
    if (p && p->d_colbase) cudaFree(p->d_colbase);
    if (p && p->d_rowf) cudaFree(p->d_rowf);
    if (p && p->d_xoff) cudaFree(p->d_xoff);
    if (p && p->d_cw) cudaFree(p->d_cw);
    delete p;
  }
  asm_free(ctx);
  for (AsmState& st : ctx->as_store) { ctx->as = st; asm_free(ctx); }
  if (ctx->fac.d_X) cudaFree(ctx->fac.d_X);
  for (int i = 0; i < 2; ++i)
    if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
  if (ctx->comm) comm_free(ctx);
  if (ctx->gen_scratch) cudaFree(ctx->gen_scratch);
  if (ctx->gen_tab) cudaFree(ctx->gen_tab);
  if (ctx->gen_need) cudaFree(ctx->gen_need);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  delete ctx;
}

const char* ghb_last_error(const ghb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int ghb_set_option(ghb_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return GHB_EINVAL;
  Options& o = ctx->opt;
  const std::string n(name);
  if (n == "force_generic") o.force_generic = (int)value;
  else if (n == "cw") o.cw = (int)value;
  else if (n == "dmma_ll") o.dmma_ll = (int)value;
  else if (n == "factors_generic") o.factors_generic = (int)value;
  else if (n == "max_ctas_per_sm") o.max_ctas_per_sm = (int)value;
  else if (n == "ll_ctas") o.ll_ctas = (int)value;
  else if (n == "warp_one_cell") o.warp_one_cell = (int)value;
  else if (n == "warp_two_rows") o.warp_two_rows = (int)value;
  else if (n == "debug") o.debug = (int)value;
  else if (n == "fused_assembly") o.fused_assembly = (int)value;
  else if (n == "cw_back") o.cw_back = (int)value;
  else if (n == "cw_q4") o.cw_q4 = (int)value;
  else if (n == "stream_chunk_bytes") o.stream_chunk_bytes = std::max<int64_t>(1, value);
  else return fail(ctx, GHB_EINVAL, "ghb_set_option: unknown option " + n);
  for (Plan* p : ctx->plans)          // launch-time knobs follow; the kernel choice of existing plans does not change
    if (p) { p->opt.max_ctas_per_sm = o.max_ctas_per_sm; p->opt.debug = o.debug; p->opt.dmma_ll = o.dmma_ll; p->opt.ll_ctas = o.ll_ctas; p->opt.cw_q4 = o.cw_q4; }
  return GHB_OK;
}

int ghb_set_stream(ghb_ctx* ctx, void* s) {
  if (!ctx) return GHB_EINVAL;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  // NULL is a valid handle: the legacy default stream (what torch uses unless told otherwise)
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return GHB_OK;
}

/* device buffers for hosts without a CUDA array library of their own (the Julia glue keeps S_K, g_K and the CSC values
   on the device between the call sites instead of round-tripping them through host memory) */
int ghb_device_alloc(ghb_ctx* ctx, int64_t bytes, void** out) {
  if (!ctx || !out || bytes < 0) return GHB_EINVAL;
  *out = nullptr;
  if (bytes == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  if (cudaMalloc(out, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, GHB_ENOMEM, "ghb_device_alloc"); }
  return GHB_OK;
}

int ghb_device_free(ghb_ctx* ctx, void* ptr) {
  if (!ctx) return GHB_EINVAL;
  if (!ptr) return GHB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  GHB_CUDA(ctx, cudaFree(ptr));
  return GHB_OK;
}

int ghb_copy(ghb_ctx* ctx, void* dst, const void* src, int64_t bytes) {
  if (!ctx || bytes < 0 || (bytes > 0 && (!dst || !src))) return GHB_EINVAL;
  if (bytes == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  GHB_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, ctx->stream));
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GHB_OK;
}

int ghb_host_register(ghb_ctx* ctx, void* ptr, int64_t bytes) {
  if (!ctx || !ptr || bytes <= 0) return ctx ? fail(ctx, GHB_EINVAL, "ghb_host_register: bad argument") : GHB_EINVAL;
  cudaSetDevice(ctx->device);
  const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();                        // not sticky: do not let the next launch check trip over it
    return fail(ctx, GHB_ECUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  }
  return GHB_OK;
}

int ghb_host_unregister(ghb_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return ctx ? fail(ctx, GHB_EINVAL, "ghb_host_unregister: bad argument") : GHB_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, GHB_ECUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
  }
  return GHB_OK;
}

int ghb_synchronize(ghb_ctx* ctx) {
  if (!ctx) return GHB_EINVAL;
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GHB_OK;
}

int64_t ghb_launch_count(const ghb_ctx* ctx) { return ctx ? ctx->launches : -1; }

const char* ghb_plan_kernel_name(ghb_ctx* ctx, int plan_id) {
  Plan* p = get_plan(ctx, plan_id);
  return p ? p->kernel_name : "";
}

int ghb_plan_blocks(ghb_ctx* ctx, int nfields, const int32_t* ndofs, const uint8_t* touched, int n_int,
                    const int32_t* interior, int n_bnd, const int32_t* boundary, int* plan_id) {
  if (!ctx) return GHB_EINVAL;
  if (!ndofs || !touched || !plan_id || (n_int > 0 && !interior) || (n_bnd > 0 && !boundary))
    return fail(ctx, GHB_EINVAL, "ghb_plan_blocks: null argument");
  if (nfields <= 0 || n_int < 0 || n_bnd <= 0 || n_int + n_bnd != nfields)
    return fail(ctx, GHB_EINVAL, "ghb_plan_blocks: interior+boundary must cover 1:nfields (need >=1 boundary field)");
  // _check_preconditions (StaticCondensationMap.jl:16-34): disjoint cover of 1:nfields
  std::vector<int> seen(nfields, 0);
  for (int k = 0; k < n_int + n_bnd; ++k) {
    int f = k < n_int ? interior[k] : boundary[k - n_int];
    if (f < 1 || f > nfields || seen[f - 1]) return fail(ctx, GHB_EINVAL, "ghb_plan_blocks: fields are not a disjoint cover of 1:nfields");
    seen[f - 1] = 1;
  }
  for (int f = 0; f < nfields; ++f)
    if (ndofs[f] <= 0) return fail(ctx, GHB_EINVAL, "ghb_plan_blocks: ndofs must be positive");
  // a block row (or column) without any touched block has no defined size in the reference
  for (int f = 0; f < nfields; ++f) {
    bool r = false, c = false;
    for (int q = 0; q < nfields; ++q) { r |= touched[f + nfields * q] != 0; c |= touched[q + nfields * f] != 0; }
    if (!r || !c) return fail(ctx, GHB_EINVAL, "ghb_plan_blocks: a field has no touched block in its block row/column");
  }
  Plan* p = new (std::nothrow) Plan();
  if (!p) return fail(ctx, GHB_ENOMEM, "plan alloc");
  p->nfields = nfields;
  p->ndofs.assign(ndofs, ndofs + nfields);
  p->touched.assign(touched, touched + nfields * nfields);
  p->interior.assign(interior, interior + n_int);
  p->boundary.assign(boundary, boundary + n_bnd);
  p->block_offset.assign((size_t)nfields * nfields, -1);
  int64_t off = 0;
  p->all_touched = true;
  for (int j = 0; j < nfields; ++j)
    for (int i = 0; i < nfields; ++i) {
      if (touched[i + nfields * j]) {
        p->block_offset[i + nfields * j] = off;
        off += (int64_t)ndofs[i] * ndofs[j];
      } else {
        p->all_touched = false;
      }
    }
  p->lenA = (int)off;
  p->field_offset_b.resize(nfields);
  int bo = 0;
  for (int f = 0; f < nfields; ++f) { p->field_offset_b[f] = bo; bo += ndofs[f]; }
  p->lenb = bo;
  for (int k = 0; k < n_int + n_bnd; ++k) {
    int f = (k < n_int ? interior[k] : boundary[k - n_int]) - 1;
    for (int l = 0; l < ndofs[f]; ++l) { p->row_field.push_back(f); p->row_local.push_back(l); }
    (k < n_int ? p->n_i : p->n_b) += ndofs[f];
  }
  p->n = p->n_i + p->n_b;
  if (p->n_b > 255) { delete p; return fail(ctx, GHB_EUNSUPPORTED, "ghb_plan_blocks: n_b > 255 not supported"); }
  // element map in condensed order
  const int n = p->n;
  std::vector<int32_t> emap((size_t)n * (n + 1));
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      int fi = p->row_field[i], fj = p->row_field[j];
      int64_t bofs = p->block_offset[fi + nfields * fj];
      emap[i + (size_t)n * j] = bofs < 0 ? -1 : (int32_t)(bofs + p->row_local[i] + (int64_t)p->row_local[j] * ndofs[fi]);
    }
  for (int i = 0; i < n; ++i) emap[i + (size_t)n * n] = p->field_offset_b[p->row_field[i]] + p->row_local[i];
  cudaSetDevice(ctx->device);
  if (cudaMalloc((void**)&p->d_emap, emap.size() * sizeof(int32_t)) != cudaSuccess) { delete p; return fail(ctx, GHB_ENOMEM, "emap alloc"); }
  cudaMemcpy(p->d_emap, emap.data(), emap.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  p->opt = ctx->opt;
  const bool force = ctx->opt.force_generic != 0;
  if (dmma_supported(*p) && !force) {
    int rc = dmma_prepare(ctx, *p);
    if (rc != GHB_OK) { cudaFree(p->d_emap); delete p; return rc; }
    p->use_dmma = true;
  } else if (warp_kernel_name(*p) && !force) {
    p->use_warp = true;
    p->kernel_name = warp_kernel_name(*p);
  } else if (large_supported(ctx, *p) && !force) {
    int rc = dmma_prepare(ctx, *p);      // same re-layout tables
    if (rc != GHB_OK) { cudaFree(p->d_emap); delete p; return rc; }
    p->use_large = true;
    p->kernel_name = "large_dmma";
  }
  // the one-warp-per-cell kernel takes the condensation of the shapes it is instantiated for (option cw = 0: A/B runs
  // against the 4-warps-per-cell kernels)
  if (cw_supported(*p) && !force && ctx->opt.cw) {
    int rc = cw_prepare(ctx, *p);
    if (rc != GHB_OK) { cudaFree(p->d_emap); delete p; return rc; }
    p->use_cw = true;
    p->kernel_name = cw_kernel_name(*p);
  } else if (!p->use_dmma && !p->use_warp && !p->use_large && cw_pad_supported(*p) && !force && ctx->opt.cw) {
    // no tuned kernel for this shape: the shape-generic instantiation of the same kernel (padded n_i class, any fields)
    p->cw_pad = cw_pad_class_of(*p);
    int rc = cw_prepare(ctx, *p);
    if (rc != GHB_OK) { cudaFree(p->d_emap); delete p; return rc; }
    p->use_cw = true;
    p->kernel_name = cw_kernel_name(*p);
  }
  ctx->plans.push_back(p);
  *plan_id = (int)ctx->plans.size() - 1;
  return GHB_OK;
}

int ghb_plan_query(ghb_ctx* ctx, int plan_id, int64_t out[4]) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p || !out) return fail(ctx, GHB_EINVAL, "ghb_plan_query: bad plan id");
  out[0] = p->n_i; out[1] = p->n_b; out[2] = p->lenA; out[3] = p->lenb;
  return GHB_OK;
}

int ghb_condense_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b, double* S,
                     double* g, int32_t* info, int keep_factors) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_condense_f64: bad plan id");
  if (ncells < 0) return fail(ctx, GHB_EINVAL, "ghb_condense_f64: ncells < 0");
  if (ncells == 0) return GHB_OK;
  if (!A || !b || !S || !g) return fail(ctx, GHB_EINVAL, "ghb_condense_f64: null array");
  cudaSetDevice(ctx->device);
  double* X = nullptr;
  int32_t* finfo = nullptr;
  if (keep_factors) GHB_TRY(factor_storage(ctx, *p, plan_id, ncells, &X, &finfo));
  else ctx->fac.plan_id = -1;
  Arg<double> dA(ctx, A, (size_t)ncells * p->lenA, true, false); GHB_TRY(dA.rc);
  Arg<double> db(ctx, b, (size_t)ncells * p->lenb, true, false); GHB_TRY(db.rc);
  Arg<double> dS(ctx, S, (size_t)ncells * p->n_b * p->n_b, false, true); GHB_TRY(dS.rc);
  Arg<double> dg(ctx, g, (size_t)ncells * p->n_b, false, true); GHB_TRY(dg.rc);
  Arg<int32_t> di(ctx, info, info ? (size_t)ncells : 0, false, true); GHB_TRY(di.rc);
  GHB_TRY(launch_condense(ctx, *p, ncells, dA.dev, db.dev, dS.dev, dg.dev, finfo ? finfo : di.dev, X));
  if (finfo && di.dev)
    GHB_CUDA(ctx, cudaMemcpyAsync(di.dev, finfo, (size_t)ncells * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  GHB_TRY(dS.finish()); GHB_TRY(dg.finish()); GHB_TRY(di.finish());
  return GHB_OK;
}

int64_t ghb_factors_generation(const ghb_ctx* ctx) { return (ctx && ctx->fac.plan_id >= 0) ? (int64_t)ctx->fac.generation : -1; }

int ghb_restrict_facet_dofs_i64(ghb_ctx* ctx, int64_t ncells, int nlfacets, int ndofs_f,
                                const int64_t* cell_wise_facets, const int64_t* facet_data, int64_t* out) {
  if (!ctx) return GHB_EINVAL;
  if (ncells < 0 || nlfacets <= 0 || ndofs_f <= 0 || !cell_wise_facets || !facet_data || !out)
    return fail(ctx, GHB_EINVAL, "ghb_restrict_facet_dofs_i64: bad argument");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  if (!is_device_ptr(facet_data) || !is_device_ptr(cell_wise_facets) || !is_device_ptr(out))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_restrict_facet_dofs_i64: device pointers required (facet table size is not passed)");
  return launch_restrict_facet_dofs(ctx, ncells, nlfacets, ndofs_f, cell_wise_facets, facet_data, out);
}

int ghb_sum_facets_f64(ghb_ctx* ctx, int64_t ncells, int nlfacets, int64_t len, const double* in, double* out) {
  if (!ctx) return GHB_EINVAL;
  if (ncells < 0 || nlfacets <= 0 || len <= 0 || !in || !out) return fail(ctx, GHB_EINVAL, "ghb_sum_facets_f64: bad argument");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  Arg<double> din(ctx, in, (size_t)ncells * nlfacets * len, true, false); GHB_TRY(din.rc);
  Arg<double> dout(ctx, out, (size_t)ncells * len, false, true); GHB_TRY(dout.rc);
  GHB_TRY(launch_sum_facets(ctx, ncells, nlfacets, len, din.dev, dout.dev));
  GHB_TRY(dout.finish());
  return GHB_OK;
}

int ghb_l2_projection_dofs_f64(ghb_ctx* ctx, int64_t nbatch, int n, int nrhs, const double* A, const double* B, double* X,
                               int32_t* info) {
  if (!ctx) return GHB_EINVAL;
  if (nbatch < 0 || n < 1 || nrhs < 1 || !A || !B || !X) return fail(ctx, GHB_EINVAL, "ghb_l2_projection_dofs_f64: bad argument");
  if (n > 128) return fail(ctx, GHB_EUNSUPPORTED, "ghb_l2_projection_dofs_f64: n > 128");
  if (nbatch == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  Arg<double> dA(ctx, A, (size_t)nbatch * n * n, true, false); GHB_TRY(dA.rc);
  Arg<double> dB(ctx, B, (size_t)nbatch * n * nrhs, true, false); GHB_TRY(dB.rc);
  Arg<double> dX(ctx, X, (size_t)nbatch * n * nrhs, false, true); GHB_TRY(dX.rc);
  Arg<int32_t> di(ctx, info, info ? (size_t)nbatch : 0, false, true); GHB_TRY(di.rc);
  GHB_TRY(launch_batched_solve(ctx, nbatch, n, nrhs, dA.dev, dB.dev, dX.dev, di.dev));
  GHB_TRY(dX.finish()); GHB_TRY(di.finish());
  return GHB_OK;
}

int ghb_expand_records_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA, const double* Tb,
                           const double* coef, double* A, double* b) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_expand_records_f64: bad plan id");
  if (ncells < 0 || ntab < 1 || ntab > 16 || !TA || !Tb || !coef || !A || !b)
    return fail(ctx, GHB_EINVAL, "ghb_expand_records_f64: bad argument (need 1 <= ntab <= 16, non-null arrays)");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  Arg<double> dTA(ctx, TA, (size_t)ntab * p->lenA, true, false); GHB_TRY(dTA.rc);
  Arg<double> dTb(ctx, Tb, (size_t)ntab * p->lenb, true, false); GHB_TRY(dTb.rc);
  Arg<double> dc(ctx, coef, (size_t)ncells * ntab, true, false); GHB_TRY(dc.rc);
  Arg<double> dA(ctx, A, (size_t)ncells * p->lenA, false, true); GHB_TRY(dA.rc);
  Arg<double> db(ctx, b, (size_t)ncells * p->lenb, false, true); GHB_TRY(db.rc);
  GHB_TRY(launch_expand_records(ctx, ncells, p->lenA, ntab, dTA.dev, dc.dev, dA.dev));
  GHB_TRY(launch_expand_records(ctx, ncells, p->lenb, ntab, dTb.dev, dc.dev, db.dev));
  GHB_TRY(dA.finish()); GHB_TRY(db.finish());
  return GHB_OK;
}

// records of an affine family for plans without a GEN kernel: chunks of records expanded into a device temporary
static int condense_affine_chunked(ghb_ctx* ctx, const Plan& p, int64_t ncells, int ntab, const double* TA, const double* Tb,
                                   const double* coef, double* S, double* g, int32_t* info, CallTmp& tmp) {
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncells, ((int64_t)256 << 20) / ((p.lenA + p.lenb) * 8)));
  double *tA = nullptr, *tb = nullptr;
  GHB_CUDA(ctx, tmp.alloc((void**)&tA, (size_t)chunk * p.lenA * 8));
  GHB_CUDA(ctx, tmp.alloc((void**)&tb, (size_t)chunk * p.lenb * 8));
  for (int64_t c0 = 0; c0 < ncells; c0 += chunk) {
    const int64_t nc = std::min(chunk, ncells - c0);
    GHB_TRY(launch_expand_records(ctx, nc, p.lenA, ntab, TA, coef + c0 * ntab, tA));
    GHB_TRY(launch_expand_records(ctx, nc, p.lenb, ntab, Tb, coef + c0 * ntab, tb));
    GHB_TRY(launch_condense(ctx, p, nc, tA, tb, S + c0 * p.n_b * p.n_b, g + c0 * p.n_b, info ? info + c0 : nullptr, nullptr));
  }
  return GHB_OK;
}

int ghb_condense_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA, const double* Tb,
                            const double* coef, double* S, double* g, int32_t* info, int keep_factors) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_condense_affine_f64: bad plan id");
  if (ncells < 0 || ntab < 1 || ntab > 16 || !TA || !Tb || !coef || !S || !g)
    return fail(ctx, GHB_EINVAL, "ghb_condense_affine_f64: bad argument (need 1 <= ntab <= 16, non-null arrays)");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  double* X = nullptr;
  int32_t* finfo = nullptr;
  if (keep_factors) GHB_TRY(factor_storage(ctx, *p, plan_id, ncells, &X, &finfo));
  else ctx->fac.plan_id = -1;
  Arg<double> dTA(ctx, TA, (size_t)ntab * p->lenA, true, false); GHB_TRY(dTA.rc);
  Arg<double> dTb(ctx, Tb, (size_t)ntab * p->lenb, true, false); GHB_TRY(dTb.rc);
  Arg<double> dc(ctx, coef, (size_t)ncells * ntab, true, false); GHB_TRY(dc.rc);
  Arg<double> dS(ctx, S, (size_t)ncells * p->n_b * p->n_b, false, true); GHB_TRY(dS.rc);
  Arg<double> dg(ctx, g, (size_t)ncells * p->n_b, false, true); GHB_TRY(dg.rc);
  Arg<int32_t> di(ctx, info, info ? (size_t)ncells : 0, false, true); GHB_TRY(di.rc);
  CallTmp tmp(ctx);
  const bool gen = cw_gen_supported(*p, ntab);
  if (gen) {
    GHB_TRY(launch_condense_cw_gen(ctx, *p, ncells, ntab, dTA.dev, dTb.dev, dc.dev, dS.dev, dg.dev, finfo ? finfo : di.dev, X, nullptr));
  } else {
    // plans without a GEN kernel: chunks of records expanded into a device temporary
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncells, ((int64_t)256 << 20) / ((p->lenA + p->lenb) * 8)));
    double *tA = nullptr, *tb = nullptr;
    GHB_CUDA(ctx, tmp.alloc((void**)&tA, (size_t)chunk * p->lenA * 8));
    GHB_CUDA(ctx, tmp.alloc((void**)&tb, (size_t)chunk * p->lenb * 8));
    int32_t* ip = finfo ? finfo : di.dev;
    for (int64_t c0 = 0; c0 < ncells; c0 += chunk) {
      const int64_t nc = std::min(chunk, ncells - c0);
      GHB_TRY(launch_expand_records(ctx, nc, p->lenA, ntab, dTA.dev, dc.dev + c0 * ntab, tA));
      GHB_TRY(launch_expand_records(ctx, nc, p->lenb, ntab, dTb.dev, dc.dev + c0 * ntab, tb));
      GHB_TRY(launch_condense(ctx, *p, nc, tA, tb, dS.dev + c0 * p->n_b * p->n_b, dg.dev + c0 * p->n_b, ip ? ip + c0 : nullptr,
                              X ? X + c0 * (int64_t)p->n_i * (p->n_b + 1) : nullptr));
    }
  }
  if (finfo && di.dev)
    GHB_CUDA(ctx, cudaMemcpyAsync(di.dev, finfo, (size_t)ncells * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  GHB_TRY(dS.finish()); GHB_TRY(dg.finish()); GHB_TRY(di.finish());
  return GHB_OK;
}

int ghb_condense_assemble_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA,
                                     const double* Tb, const double* coef, const double* dirichlet_vals_in,
                                     int64_t ndirichlet, double* nzval, double* rhs, int32_t* info) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_affine_f64: bad plan id");
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_condense_assemble_affine_f64: call ghb_assemble_symbolic first");
  const AsmState& as = ctx->as;
  if (as.nghost) return fail(ctx, GHB_ESTATE, "ghb_condense_assemble_affine_f64: the cached pattern has ghost cells (slab mode)");
  if (ncells != as.ncells || p->n_b != as.n_b)
    return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_affine_f64: ncells / n_b differ from the symbolic phase");
  if (ntab < 1 || ntab > 16 || !TA || !Tb || !coef || !nzval || !rhs || ndirichlet < 0)
    return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_affine_f64: bad argument (need 1 <= ntab <= 16, non-null arrays)");
  cudaSetDevice(ctx->device);
  Arg<double> dd(ctx, dirichlet_vals_in, dirichlet_vals_in ? (size_t)ndirichlet : 0, true, false); GHB_TRY(dd.rc);
  Arg<double> dTA(ctx, TA, (size_t)ntab * p->lenA, true, false); GHB_TRY(dTA.rc);
  Arg<double> dTb(ctx, Tb, (size_t)ntab * p->lenb, true, false); GHB_TRY(dTb.rc);
  Arg<double> dc(ctx, coef, (size_t)ncells * ntab, true, false); GHB_TRY(dc.rc);
  Arg<double> dz(ctx, nzval, (size_t)as.nnz, false, true); GHB_TRY(dz.rc);
  Arg<double> dr(ctx, rhs, (size_t)as.nrows, false, true); GHB_TRY(dr.rc);
  Arg<int32_t> di(ctx, info, info ? (size_t)ncells : 0, false, true); GHB_TRY(di.rc);
  CallTmp tmp(ctx);
  double *dS = nullptr, *dg = nullptr;
  const bool gen = cw_gen_supported(*p, ntab);
  // the fused kernel stores S_K only for the cells whose Dirichlet lift needs it: without Dirichlet values S is never touched
  const bool needS = !(gen && ctx->opt.fused_assembly) || dd.dev != nullptr;
  GHB_CUDA(ctx, tmp.alloc((void**)&dS, needS ? (size_t)ncells * p->n_b * p->n_b * sizeof(double) : 64));
  GHB_CUDA(ctx, tmp.alloc((void**)&dg, (size_t)ncells * p->n_b * sizeof(double)));
  if (gen && ctx->opt.fused_assembly) {
    // one kernel from coefficients to CSC values: records generated in the loader, S_K scattered into the zeroed nzval
    GHB_TRY(asm_scatter_prepare(ctx, 0));
    GHB_CUDA(ctx, cudaMemsetAsync(dz.dev, 0, (size_t)as.nnz * sizeof(double), ctx->stream));
    ScatterArgs sc{dz.dev, as.d_colpos, as.d_rowrank, dd.dev ? as.d_keepS : nullptr};
    GHB_TRY(launch_condense_cw_gen(ctx, *p, ncells, ntab, dTA.dev, dTb.dev, dc.dev, dS, dg, di.dev, nullptr, &sc));
    GHB_TRY(asm_numeric_range(ctx, dS, dg, nullptr, dd.dev, dz.dev, dr.dev, 0, as.nrows, ASM_RHS));
  } else {
    if (gen)
      GHB_TRY(launch_condense_cw_gen(ctx, *p, ncells, ntab, dTA.dev, dTb.dev, dc.dev, dS, dg, di.dev, nullptr, nullptr));
    else
      GHB_TRY(condense_affine_chunked(ctx, *p, ncells, ntab, dTA.dev, dTb.dev, dc.dev, dS, dg, di.dev, tmp));
    GHB_TRY(asm_numeric(ctx, dS, dg, nullptr, dd.dev, dz.dev, dr.dev));
  }
  GHB_TRY(dz.finish()); GHB_TRY(dr.finish()); GHB_TRY(di.finish());
  return GHB_OK;
}

int ghb_assemble_symbolic(ghb_ctx* ctx, int64_t ncells, int n_b, const int64_t* cell_ids, int64_t nrows,
                          int64_t* nnz_out) {
  if (!ctx) return GHB_EINVAL;
  if (ncells <= 0 || n_b <= 0 || n_b > 255 || !cell_ids || nrows <= 0)
    return fail(ctx, GHB_EINVAL, "ghb_assemble_symbolic: bad argument (need ncells>0, 0<n_b<=255, nrows>0)");
  cudaSetDevice(ctx->device);
  asm_new(ctx);
  size_t cnt = (size_t)ncells * n_b;
  GHB_CUDA(ctx, cudaMalloc((void**)&ctx->as.d_ids, cnt * sizeof(int64_t)));
  GHB_CUDA(ctx, cudaMemcpyAsync(ctx->as.d_ids, cell_ids, cnt * sizeof(int64_t),
                                is_device_ptr(cell_ids) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  int rc = asm_symbolic(ctx, ncells, 0, 0, n_b, ctx->as.d_ids, nrows, 0, nrows);
  if (rc != GHB_OK) { asm_free(ctx); return rc; }
  if (nnz_out) *nnz_out = ctx->as.nnz;
  return GHB_OK;
}

int ghb_assemble_symbolic_slab(ghb_ctx* ctx, int64_t ncells_local, int64_t nghost, int ghost_ncols, int n_b,
                               const int64_t* cell_ids, int64_t nrows_global, int64_t col_begin, int64_t col_end,
                               int64_t* nnz_out) {
  if (!ctx) return GHB_EINVAL;
  if (ncells_local <= 0 || nghost < 0 || n_b <= 0 || n_b > 255 || !cell_ids || nrows_global <= 0 || col_begin < 1 ||
      col_end < col_begin || col_end > nrows_global + 1 || ghost_ncols < 0 || ghost_ncols > n_b)
    return fail(ctx, GHB_EINVAL, "ghb_assemble_symbolic_slab: bad argument");
  if (col_end == col_begin) return fail(ctx, GHB_EINVAL, "ghb_assemble_symbolic_slab: empty owned column range");
  cudaSetDevice(ctx->device);
  asm_new(ctx);
  size_t cnt = (size_t)(ncells_local + nghost) * n_b;
  GHB_CUDA(ctx, cudaMalloc((void**)&ctx->as.d_ids, cnt * sizeof(int64_t)));
  GHB_CUDA(ctx, cudaMemcpyAsync(ctx->as.d_ids, cell_ids, cnt * sizeof(int64_t),
                                is_device_ptr(cell_ids) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  int rc = asm_symbolic(ctx, ncells_local, nghost, ghost_ncols, n_b, ctx->as.d_ids, nrows_global, col_begin - 1,
                        col_end - col_begin);
  if (rc != GHB_OK) { asm_free(ctx); return rc; }
  if (nnz_out) *nnz_out = ctx->as.nnz;
  return GHB_OK;
}

int ghb_pack_cut_plane_f64(ghb_ctx* ctx, int64_t ncut, int n_b, int ncols, const double* S, const double* g,
                           const int64_t* cell_ids, const double* dirichlet_vals, double* out) {
  if (!ctx) return GHB_EINVAL;
  if (ncut < 0 || n_b <= 0 || ncols <= 0 || ncols > n_b || !S || !g || !out || (dirichlet_vals && !cell_ids))
    return fail(ctx, GHB_EINVAL, "ghb_pack_cut_plane_f64: bad argument");
  cudaSetDevice(ctx->device);
  if (!is_device_ptr(S) || !is_device_ptr(g) || !is_device_ptr(out) || (cell_ids && !is_device_ptr(cell_ids)) ||
      (dirichlet_vals && !is_device_ptr(dirichlet_vals)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_pack_cut_plane_f64: device pointers required (feeds NCCL)");
  return asm_pack_cut_plane(ctx, ncut, n_b, ncols, S, g, cell_ids, dirichlet_vals, out);
}

int ghb_assemble_numeric_slab_f64(ghb_ctx* ctx, const double* S, const double* g, const double* ghost,
                                  const double* dirichlet_vals, double* nzval, double* rhs) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_assemble_numeric_slab_f64: call ghb_assemble_symbolic_slab first");
  if (!S || !g || !nzval || !rhs || (ctx->as.nghost > 0 && !ghost))
    return fail(ctx, GHB_EINVAL, "ghb_assemble_numeric_slab_f64: null array");
  cudaSetDevice(ctx->device);
  if (!is_device_ptr(S) || !is_device_ptr(g) || !is_device_ptr(nzval) || !is_device_ptr(rhs) ||
      (ghost && !is_device_ptr(ghost)) || (dirichlet_vals && !is_device_ptr(dirichlet_vals)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_assemble_numeric_slab_f64: device pointers required");
  return asm_numeric(ctx, S, g, ghost, dirichlet_vals, nzval, rhs);
}

int ghb_assemble_current(const ghb_ctx* ctx) { return ctx ? ctx->as_id : -1; }

int ghb_assemble_select(ghb_ctx* ctx, int pattern_id) {
  if (!ctx) return GHB_EINVAL;
  if (pattern_id < 0 || pattern_id >= (int)ctx->as_store.size()) return fail(ctx, GHB_EINVAL, "ghb_assemble_select: bad pattern id");
  if (pattern_id == ctx->as_id) return GHB_OK;
  if (!ctx->as_store[pattern_id].valid) return fail(ctx, GHB_ESTATE, "ghb_assemble_select: the pattern was released or its symbolic phase failed");
  asm_stash(ctx);
  ctx->as = ctx->as_store[pattern_id];
  ctx->as_store[pattern_id] = AsmState();
  ctx->as_id = pattern_id;
  return GHB_OK;
}

int ghb_assemble_release(ghb_ctx* ctx, int pattern_id) {
  if (!ctx) return GHB_EINVAL;
  if (pattern_id < 0 || pattern_id >= (int)ctx->as_store.size()) return fail(ctx, GHB_EINVAL, "ghb_assemble_release: bad pattern id");
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (pattern_id == ctx->as_id) { asm_free(ctx); ctx->as_id = -1; return GHB_OK; }
  AsmState keep = ctx->as;
  ctx->as = ctx->as_store[pattern_id];
  asm_free(ctx);
  ctx->as_store[pattern_id] = AsmState();
  ctx->as = keep;
  return GHB_OK;
}

/* fused condensation + assembly of a slab (device pointers): condenses the local cells and scatters S_K into the zeroed
   nzval; S_K is stored for the cells with a Dirichlet dof and for the first keep_cut cells (the layer whose cut-plane
   columns ghb_pack_cut_plane_f64 sends down), g_K for all. */
int ghb_condense_scatter_slab_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b, double* S,
                                  double* g, int32_t* info, double* nzval, int64_t keep_cut, int zero_nzval) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_condense_scatter_slab_f64: bad plan id");
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_condense_scatter_slab_f64: call ghb_assemble_symbolic_slab first");
  const AsmState& as = ctx->as;
  if (ncells != as.ncells_local || p->n_b != as.n_b || keep_cut < 0)
    return fail(ctx, GHB_EINVAL, "ghb_condense_scatter_slab_f64: ncells / n_b differ from the symbolic phase");
  if (!p->use_cw) return fail(ctx, GHB_EUNSUPPORTED, "ghb_condense_scatter_slab_f64: the plan has no cell-warp kernel");
  if (!A || !b || !S || !g || !nzval) return fail(ctx, GHB_EINVAL, "ghb_condense_scatter_slab_f64: null array");
  cudaSetDevice(ctx->device);
  if (!is_device_ptr(A) || !is_device_ptr(b) || !is_device_ptr(S) || !is_device_ptr(g) || !is_device_ptr(nzval) ||
      (info && !is_device_ptr(info)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_condense_scatter_slab_f64: device pointers required");
  GHB_TRY(asm_scatter_prepare(ctx, keep_cut));
  if (zero_nzval) GHB_CUDA(ctx, cudaMemsetAsync(nzval, 0, (size_t)as.nnz * sizeof(double), ctx->stream));
  ScatterArgs sc{nzval, as.d_colpos, as.d_rowrank, as.d_keepS};
  return launch_condense_cw_scatter(ctx, *p, ncells, A, b, S, g, info, sc);
}

int ghb_condense_scatter_slab_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA,
                                         const double* Tb, const double* coef, double* S, double* g, int32_t* info,
                                         double* nzval, int64_t keep_cut, int zero_nzval) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_condense_scatter_slab_affine_f64: bad plan id");
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_condense_scatter_slab_affine_f64: call ghb_assemble_symbolic_slab first");
  const AsmState& as = ctx->as;
  if (ncells != as.ncells_local || p->n_b != as.n_b || keep_cut < 0)
    return fail(ctx, GHB_EINVAL, "ghb_condense_scatter_slab_affine_f64: ncells / n_b differ from the symbolic phase");
  if (ntab < 1 || ntab > 16 || !TA || !Tb || !coef || !S || !g || !nzval)
    return fail(ctx, GHB_EINVAL, "ghb_condense_scatter_slab_affine_f64: bad argument (need 1 <= ntab <= 16, non-null arrays)");
  if (!cw_gen_supported(*p, ntab))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_condense_scatter_slab_affine_f64: the plan has no cell-warp kernel that can stage the tables");
  cudaSetDevice(ctx->device);
  if (!is_device_ptr(TA) || !is_device_ptr(Tb) || !is_device_ptr(coef) || !is_device_ptr(S) || !is_device_ptr(g) ||
      !is_device_ptr(nzval) || (info && !is_device_ptr(info)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_condense_scatter_slab_affine_f64: device pointers required");
  GHB_TRY(asm_scatter_prepare(ctx, keep_cut));
  if (zero_nzval) GHB_CUDA(ctx, cudaMemsetAsync(nzval, 0, (size_t)as.nnz * sizeof(double), ctx->stream));
  ScatterArgs sc{nzval, as.d_colpos, as.d_rowrank, as.d_keepS};
  return launch_condense_cw_gen(ctx, *p, ncells, ntab, TA, Tb, coef, S, g, info, nullptr, &sc);
}

/* second half: contributions of the ghost cells (the packed buffer received from the slab above) and the rhs gather */
int ghb_assemble_finish_slab_f64(ghb_ctx* ctx, const double* S, const double* g, const double* ghost,
                                 const double* dirichlet_vals, double* nzval, double* rhs) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->as.valid || !ctx->as.d_colpos) return fail(ctx, GHB_ESTATE, "ghb_assemble_finish_slab_f64: call ghb_condense_scatter_slab_f64 first");
  if (!S || !g || !nzval || !rhs || (ctx->as.nghost > 0 && !ghost)) return fail(ctx, GHB_EINVAL, "ghb_assemble_finish_slab_f64: null array");
  cudaSetDevice(ctx->device);
  if (!is_device_ptr(S) || !is_device_ptr(g) || !is_device_ptr(nzval) || !is_device_ptr(rhs) ||
      (ghost && !is_device_ptr(ghost)) || (dirichlet_vals && !is_device_ptr(dirichlet_vals)))
    return fail(ctx, GHB_EUNSUPPORTED, "ghb_assemble_finish_slab_f64: device pointers required");
  GHB_TRY(asm_scatter_ghosts(ctx, ghost, nzval));
  return asm_numeric_range(ctx, S, g, ghost, dirichlet_vals, nzval, rhs, 0, ctx->as.nrows, ASM_RHS);
}

int ghb_assemble_pattern(ghb_ctx* ctx, int64_t* colptr, int64_t* rowval) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_assemble_pattern: no symbolic phase cached");
  cudaSetDevice(ctx->device);
  if (colptr)
    GHB_CUDA(ctx, cudaMemcpyAsync(colptr, ctx->as.d_colptr, (ctx->as.nrows + 1) * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
  if (rowval)
    GHB_CUDA(ctx, cudaMemcpyAsync(rowval, ctx->as.d_rowval, ctx->as.nnz * sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
  GHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GHB_OK;
}

int ghb_assemble_numeric_f64(ghb_ctx* ctx, const double* S, const double* g, const double* dirichlet_vals,
                             int64_t ndirichlet, double* nzval, double* rhs) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_assemble_numeric_f64: call ghb_assemble_symbolic first");
  if (ctx->as.nghost) return fail(ctx, GHB_ESTATE, "ghb_assemble_numeric_f64: the cached pattern has ghost cells; use ghb_assemble_numeric_slab_f64");
  if (!S || !g || !nzval || !rhs) return fail(ctx, GHB_EINVAL, "ghb_assemble_numeric_f64: null array");
  cudaSetDevice(ctx->device);
  const AsmState& as = ctx->as;
  if (ndirichlet < 0) return fail(ctx, GHB_EINVAL, "ghb_assemble_numeric_f64: ndirichlet < 0");
  Arg<double> dd(ctx, dirichlet_vals, dirichlet_vals ? (size_t)ndirichlet : 0, true, false); GHB_TRY(dd.rc);
  Arg<double> dS(ctx, S, (size_t)as.ncells * as.n_b * as.n_b, true, false); GHB_TRY(dS.rc);
  Arg<double> dg(ctx, g, (size_t)as.ncells * as.n_b, true, false); GHB_TRY(dg.rc);
  Arg<double> dz(ctx, nzval, (size_t)as.nnz, false, true); GHB_TRY(dz.rc);
  Arg<double> dr(ctx, rhs, (size_t)as.nrows, false, true); GHB_TRY(dr.rc);
  GHB_TRY(asm_numeric(ctx, dS.dev, dg.dev, nullptr, dd.dev, dz.dev, dr.dev));
  GHB_TRY(dz.finish()); GHB_TRY(dr.finish());
  return GHB_OK;
}

int ghb_assemble_numeric_csr_f64(ghb_ctx* ctx, double* S, const double* g, const double* dirichlet_vals,
                                 int64_t ndirichlet, double* nzval, double* rhs) {
  if (!ctx) return GHB_EINVAL;
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_assemble_numeric_csr_f64: call ghb_assemble_symbolic first");
  if (ctx->as.nghost) return fail(ctx, GHB_EUNSUPPORTED, "ghb_assemble_numeric_csr_f64: slab patterns (ghost cells) are not supported");
  if (!S || !g || !nzval || !rhs) return fail(ctx, GHB_EINVAL, "ghb_assemble_numeric_csr_f64: null array");
  const AsmState& as = ctx->as;
  if (!is_device_ptr(S)) return fail(ctx, GHB_EUNSUPPORTED, "ghb_assemble_numeric_csr_f64: S is transposed in place and must be a device pointer");
  if (ndirichlet < 0) return fail(ctx, GHB_EINVAL, "ghb_assemble_numeric_csr_f64: ndirichlet < 0");
  cudaSetDevice(ctx->device);
  Arg<double> dd(ctx, dirichlet_vals, dirichlet_vals ? (size_t)ndirichlet : 0, true, false); GHB_TRY(dd.rc);
  Arg<double> dg(ctx, g, (size_t)as.ncells * as.n_b, true, false); GHB_TRY(dg.rc);
  Arg<double> dz(ctx, nzval, (size_t)as.nnz, false, true); GHB_TRY(dz.rc);
  Arg<double> dr(ctx, rhs, (size_t)as.nrows, false, true); GHB_TRY(dr.rc);
  // the rhs (with the Dirichlet lift g_K - S_K vals_K) needs S_K itself, the row-major values its transpose
  GHB_TRY(asm_numeric_range(ctx, S, dg.dev, nullptr, dd.dev, dz.dev, dr.dev, 0, as.nrows, ASM_RHS));
  GHB_TRY(launch_transpose_blocks(ctx, as.ncells, as.n_b, S));
  GHB_TRY(asm_numeric_range(ctx, S, dg.dev, nullptr, dd.dev, dz.dev, dr.dev, 0, as.nrows, ASM_MATRIX));
  GHB_TRY(dz.finish()); GHB_TRY(dr.finish());
  return GHB_OK;
}

int ghb_condense_assemble_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b,
                              const double* dirichlet_vals_in, int64_t ndirichlet, double* nzval, double* rhs,
                              int32_t* info) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_f64: bad plan id");
  if (!ctx->as.valid) return fail(ctx, GHB_ESTATE, "ghb_condense_assemble_f64: call ghb_assemble_symbolic first");
  const AsmState& as = ctx->as;
  if (as.nghost) return fail(ctx, GHB_ESTATE, "ghb_condense_assemble_f64: the cached pattern has ghost cells (slab mode)");
  if (ncells != as.ncells || p->n_b != as.n_b)
    return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_f64: ncells / n_b differ from the symbolic phase");
  if (!A || !b || !nzval || !rhs) return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_f64: null array");
  if (ndirichlet < 0) return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_f64: ndirichlet < 0");
  cudaSetDevice(ctx->device);
  Arg<double> dd(ctx, dirichlet_vals_in, dirichlet_vals_in ? (size_t)ndirichlet : 0, true, false); GHB_TRY(dd.rc);
  const double* dirichlet_vals = dd.dev;
  const bool hostA = !is_device_ptr(A), hostb = !is_device_ptr(b);
  if (hostA != hostb) return fail(ctx, GHB_EINVAL, "ghb_condense_assemble_f64: A and b must both be host or both device");
  double *dS = nullptr, *dg = nullptr;
  CallTmp tmp(ctx);
  // the fused kernel stores S_K only for the cells whose Dirichlet lift needs it: without Dirichlet values S is never touched
  const bool needS = !(!hostA && p->use_cw && ctx->opt.fused_assembly) || dirichlet_vals != nullptr;
  GHB_CUDA(ctx, tmp.alloc((void**)&dS, needS ? (size_t)ncells * p->n_b * p->n_b * sizeof(double) : 64));
  GHB_CUDA(ctx, tmp.alloc((void**)&dg, (size_t)ncells * p->n_b * sizeof(double)));
  Arg<int32_t> di(ctx, info, info ? (size_t)ncells : 0, false, true); GHB_TRY(di.rc);
  int rc = GHB_OK;
  if (!hostA && p->use_cw && ctx->opt.fused_assembly) {
    // fused: the condensation kernel adds S_K into the zeroed nzval itself (scatter map of the pattern, built on first
    // use); S_K is stored only for the cells with a Dirichlet dof (the lift of the rhs needs it)
    Arg<double> dz(ctx, nzval, (size_t)as.nnz, false, true); rc = dz.rc;
    Arg<double> dr(ctx, rhs, (size_t)as.nrows, false, true); if (rc == GHB_OK) rc = dr.rc;
    if (rc == GHB_OK) rc = asm_scatter_prepare(ctx, 0);
    if (rc == GHB_OK && cudaMemsetAsync(dz.dev, 0, (size_t)as.nnz * sizeof(double), ctx->stream) != cudaSuccess)
      rc = fail(ctx, GHB_ECUDA, "ghb_condense_assemble_f64: memset of nzval");
    if (rc == GHB_OK) {
      ScatterArgs sc{dz.dev, as.d_colpos, as.d_rowrank, dirichlet_vals ? as.d_keepS : nullptr};
      rc = launch_condense_cw_scatter(ctx, *p, ncells, A, b, dS, dg, di.dev, sc);
    }
    if (rc == GHB_OK) rc = asm_numeric_range(ctx, dS, dg, nullptr, dirichlet_vals, dz.dev, dr.dev, 0, as.nrows, ASM_RHS);
    if (rc == GHB_OK) rc = dz.finish();
    if (rc == GHB_OK) rc = dr.finish();
  } else if (!hostA) {
    rc = launch_condense(ctx, *p, ncells, A, b, dS, dg, di.dev, nullptr);
    if (rc == GHB_OK) {
      Arg<double> dz(ctx, nzval, (size_t)as.nnz, false, true); rc = dz.rc;
      Arg<double> dr(ctx, rhs, (size_t)as.nrows, false, true); if (rc == GHB_OK) rc = dr.rc;
      if (rc == GHB_OK) rc = asm_numeric(ctx, dS, dg, nullptr, dirichlet_vals, dz.dev, dr.dev);
      if (rc == GHB_OK) rc = dz.finish();
      if (rc == GHB_OK) rc = dr.finish();
    }
  } else {
    // Host records are streamed through two device chunk buffers: the H2D copy of chunk k+1 overlaps the
    // condensation of chunk k.  After chunk k the columns whose cells have all been condensed are assembled and, if
    // nzval/rhs are host arrays, copied back on a third stream, so that the D2H traffic overlaps the H2D traffic.
    const int64_t chunk_bytes = ctx->opt.stream_chunk_bytes;   // option stream_chunk_bytes (tests force many chunks)
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncells, chunk_bytes / ((p->lenA + p->lenb) * 8)));
    const int nchunks = (int)((ncells + chunk - 1) / chunk);
    rc = asm_ready_columns(ctx, chunk, nchunks);
    Arg<double> dz(ctx, nzval, (size_t)as.nnz, false, false); if (rc == GHB_OK) rc = dz.rc;   // copied back piecewise
    Arg<double> dr(ctx, rhs, (size_t)as.nrows, false, false); if (rc == GHB_OK) rc = dr.rc;
    if (rc == GHB_OK && (dz.host || dr.host) && !ctx->d2h_stream &&
        cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess)
      rc = fail(ctx, GHB_ECUDA, "ghb_condense_assemble_f64: cannot create the D2H stream");
    double* dA[2] = {nullptr, nullptr};
    double* db[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, k_done[2] = {nullptr, nullptr}, g_done = nullptr;
    // Pageable caller memory (a Julia Array, a numpy array) would make every cudaMemcpyAsync a synchronous staged copy:
    // such records go through two pinned staging buffers owned by the context, filled by host threads while the
    // previous chunk is on its way to the device.
    const bool pageable = !is_pinned_host_ptr(A) || !is_pinned_host_ptr(b);
    const size_t stage_bytes = (size_t)chunk * (p->lenA + p->lenb) * 8;
    if (rc == GHB_OK && pageable && ctx->pinned_bytes < stage_bytes) {
      for (int i = 0; i < 2; ++i) {
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        ctx->pinned[i] = nullptr;
      }
      ctx->pinned_bytes = 0;
      for (int i = 0; i < 2 && rc == GHB_OK; ++i)
        if (cudaHostAlloc(&ctx->pinned[i], stage_bytes, cudaHostAllocDefault) != cudaSuccess) {
          cudaGetLastError();
          rc = fail(ctx, GHB_ENOMEM, "ghb_condense_assemble_f64: pinned staging buffers");
        }
      if (rc == GHB_OK) ctx->pinned_bytes = stage_bytes;
    }
    tmp.side_streams = true;
    if (rc == GHB_OK) {
      for (int i = 0; i < 2; ++i) {
        GHB_CUDA(ctx, tmp.alloc((void**)&dA[i], (size_t)chunk * p->lenA * 8));
        GHB_CUDA(ctx, tmp.alloc((void**)&db[i], (size_t)chunk * p->lenb * 8));
        GHB_CUDA(ctx, tmp.event(&h2d_done[i]));
        GHB_CUDA(ctx, tmp.event(&k_done[i]));
      }
      GHB_CUDA(ctx, tmp.event(&g_done));
      GHB_CUDA(ctx, cudaEventRecord(k_done[0], ctx->stream));
      GHB_CUDA(ctx, cudaEventRecord(k_done[1], ctx->stream));
    }
    int64_t c0 = 0, jprev = 0, pprev = 0;
    for (int it = 0; c0 < ncells && rc == GHB_OK; ++it, c0 += chunk) {
      int s = it & 1;
      int64_t nc = std::min(chunk, ncells - c0);
      GHB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, k_done[s], 0));
      const double* srcA = A + c0 * p->lenA;
      const double* srcb = b + c0 * p->lenb;
      if (pageable) {
        GHB_CUDA(ctx, cudaEventSynchronize(h2d_done[s]));   // the copy that last read this staging buffer is done
        double* stA = static_cast<double*>(ctx->pinned[s]);
        double* stb = stA + (size_t)chunk * p->lenA;
        host_copy_parallel(stA, srcA, (size_t)nc * p->lenA * 8);
        host_copy_parallel(stb, srcb, (size_t)nc * p->lenb * 8);
        srcA = stA; srcb = stb;
      }
      GHB_CUDA(ctx, cudaMemcpyAsync(dA[s], srcA, (size_t)nc * p->lenA * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      GHB_CUDA(ctx, cudaMemcpyAsync(db[s], srcb, (size_t)nc * p->lenb * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      GHB_CUDA(ctx, cudaEventRecord(h2d_done[s], ctx->copy_stream));
      GHB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, h2d_done[s], 0));
      rc = launch_condense(ctx, *p, nc, dA[s], db[s], dS + c0 * p->n_b * p->n_b, dg + c0 * p->n_b,
                           di.dev ? di.dev + c0 : nullptr, nullptr);
      GHB_CUDA(ctx, cudaEventRecord(k_done[s], ctx->stream));
      // columns completed by this chunk
      const int64_t jnow = as.ready_J[it], pnow = it + 1 < nchunks ? as.ready_p[it] : as.nnz;
      if (rc == GHB_OK && jnow > jprev) {
        rc = asm_numeric_range(ctx, dS, dg, nullptr, dirichlet_vals, dz.dev, dr.dev, jprev, jnow);
        if (rc == GHB_OK && (dz.host || dr.host)) {
          GHB_CUDA(ctx, cudaEventRecord(g_done, ctx->stream));
          GHB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, g_done, 0));
          if (dz.host && pnow > pprev)
            GHB_CUDA(ctx, cudaMemcpyAsync(dz.host + pprev, dz.dev + pprev, (size_t)(pnow - pprev) * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream));
          if (dr.host)
            GHB_CUDA(ctx, cudaMemcpyAsync(dr.host + jprev, dr.dev + jprev, (size_t)(jnow - jprev) * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
        jprev = jnow; pprev = pnow;
      }
    }
    if (ctx->d2h_stream && (dz.host || dr.host)) {
      cudaError_t e = cudaStreamSynchronize(ctx->d2h_stream);
      if (e != cudaSuccess && rc == GHB_OK) rc = fail(ctx, GHB_ECUDA, std::string("D2H: ") + cudaGetErrorString(e));
    }
  }
  if (rc == GHB_OK) rc = di.finish();
  return rc;   // ~CallTmp releases dS, dg, the chunk buffers and the events on this and on every early-return path
}

int ghb_backsub_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* A, const double* b,
                    const double* lambda_free_in, int64_t nlambda_free, const double* lambda_dirichlet_in,
                    int64_t nlambda_dirichlet, const int64_t* cell_ids, double* u, int32_t* info) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_backsub_f64: bad plan id");
  if (ncells < 0) return fail(ctx, GHB_EINVAL, "ghb_backsub_f64: ncells < 0");
  if (ncells == 0) return GHB_OK;
  if (!cell_ids || !u || !lambda_free_in) return fail(ctx, GHB_EINVAL, "ghb_backsub_f64: null array");
  if (nlambda_free < 0 || nlambda_dirichlet < 0) return fail(ctx, GHB_EINVAL, "ghb_backsub_f64: negative length");
  cudaSetDevice(ctx->device);
  Arg<double> dlf(ctx, lambda_free_in, (size_t)nlambda_free, true, false); GHB_TRY(dlf.rc);
  Arg<double> dld(ctx, lambda_dirichlet_in, lambda_dirichlet_in ? (size_t)nlambda_dirichlet : 0, true, false); GHB_TRY(dld.rc);
  const double* lambda_free = dlf.dev;
  const double* lambda_dirichlet = dld.dev;
  Arg<int64_t> dids(ctx, cell_ids, (size_t)ncells * p->n_b, true, false); GHB_TRY(dids.rc);
  Arg<double> du(ctx, u, (size_t)ncells * p->n_i, false, true); GHB_TRY(du.rc);
  Arg<int32_t> di(ctx, info, info ? (size_t)ncells : 0, false, true); GHB_TRY(di.rc);
  if (!A && !b) {
    if (ctx->fac.plan_id != plan_id || ctx->fac.ncells != ncells)
      return fail(ctx, GHB_ESTATE, "ghb_backsub_f64: A,b are NULL but no matching keep_factors condensation is stored");
    GHB_TRY(launch_backsub_factors(ctx, *p, ncells, ctx->fac.d_X, lambda_free, lambda_dirichlet, dids.dev, du.dev));
    if (di.dev)   // the info of the condensation that produced the factors (a singular cell has NaN factors)
      GHB_CUDA(ctx, cudaMemcpyAsync(di.dev, ctx->fac.d_info, ncells * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    if (!A || !b) return fail(ctx, GHB_EINVAL, "ghb_backsub_f64: A and b must both be given or both NULL");
    Arg<double> dA(ctx, A, (size_t)ncells * p->lenA, true, false); GHB_TRY(dA.rc);
    Arg<double> db(ctx, b, (size_t)ncells * p->lenb, true, false); GHB_TRY(db.rc);
    GHB_TRY(launch_backsub(ctx, *p, ncells, dA.dev, db.dev, lambda_free, lambda_dirichlet, dids.dev, du.dev, di.dev));
    GHB_TRY(du.finish()); GHB_TRY(di.finish());
    return GHB_OK;
  }
  GHB_TRY(du.finish()); GHB_TRY(di.finish());
  return GHB_OK;
}

int ghb_backsub_affine_f64(ghb_ctx* ctx, int plan_id, int64_t ncells, int ntab, const double* TA, const double* Tb,
                           const double* coef, const double* lambda_free_in, int64_t nlambda_free,
                           const double* lambda_dirichlet_in, int64_t nlambda_dirichlet, const int64_t* cell_ids, double* u,
                           int32_t* info) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_backsub_affine_f64: bad plan id");
  if (ncells < 0 || ntab < 1 || ntab > 16 || !TA || !Tb || !coef || !cell_ids || !u || !lambda_free_in ||
      nlambda_free < 0 || nlambda_dirichlet < 0)
    return fail(ctx, GHB_EINVAL, "ghb_backsub_affine_f64: bad argument (need 1 <= ntab <= 16, non-null arrays)");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  Arg<double> dTA(ctx, TA, (size_t)ntab * p->lenA, true, false); GHB_TRY(dTA.rc);
  Arg<double> dTb(ctx, Tb, (size_t)ntab * p->lenb, true, false); GHB_TRY(dTb.rc);
  Arg<double> dc(ctx, coef, (size_t)ncells * ntab, true, false); GHB_TRY(dc.rc);
  Arg<double> dlf(ctx, lambda_free_in, (size_t)nlambda_free, true, false); GHB_TRY(dlf.rc);
  Arg<double> dld(ctx, lambda_dirichlet_in, lambda_dirichlet_in ? (size_t)nlambda_dirichlet : 0, true, false); GHB_TRY(dld.rc);
  Arg<int64_t> dids(ctx, cell_ids, (size_t)ncells * p->n_b, true, false); GHB_TRY(dids.rc);
  Arg<double> du(ctx, u, (size_t)ncells * p->n_i, false, true); GHB_TRY(du.rc);
  Arg<int32_t> di(ctx, info, info ? (size_t)ncells : 0, false, true); GHB_TRY(di.rc);
  if (cw_gen_supported(*p, ntab) && ctx->opt.cw_back) {
    GHB_TRY(launch_backsub_cw_gen(ctx, *p, ncells, ntab, dTA.dev, dTb.dev, dc.dev, dlf.dev, dld.dev, dids.dev, du.dev, di.dev));
  } else {
    // plans without a GEN kernel: chunks of records expanded into a device temporary
    CallTmp tmp(ctx);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ncells, ((int64_t)256 << 20) / ((p->lenA + p->lenb) * 8)));
    double *tA = nullptr, *tb = nullptr;
    GHB_CUDA(ctx, tmp.alloc((void**)&tA, (size_t)chunk * p->lenA * 8));
    GHB_CUDA(ctx, tmp.alloc((void**)&tb, (size_t)chunk * p->lenb * 8));
    for (int64_t c0 = 0; c0 < ncells; c0 += chunk) {
      const int64_t nc = std::min(chunk, ncells - c0);
      GHB_TRY(launch_expand_records(ctx, nc, p->lenA, ntab, dTA.dev, dc.dev + c0 * ntab, tA));
      GHB_TRY(launch_expand_records(ctx, nc, p->lenb, ntab, dTb.dev, dc.dev + c0 * ntab, tb));
      GHB_TRY(launch_backsub(ctx, *p, nc, tA, tb, dlf.dev, dld.dev, dids.dev + c0 * p->n_b, du.dev + c0 * p->n_i,
                             di.dev ? di.dev + c0 : nullptr));
    }
  }
  GHB_TRY(du.finish()); GHB_TRY(di.finish());
  return GHB_OK;
}

int ghb_scatter_free_dof_values(ghb_ctx* ctx, int plan_id, int64_t ncells, const double* u,
                                const double* lambda_free, int64_t nlambda, double* x) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_scatter_free_dof_values: bad plan id");
  if (ncells < 0 || nlambda < 0 || !u || !x || (nlambda > 0 && !lambda_free))
    return fail(ctx, GHB_EINVAL, "ghb_scatter_free_dof_values: bad argument");
  // the layout below is the reference's MultiFieldFESpace order (trial field order) only when the bulk fields come
  // first and ascending -- what every reference test uses (I = 1:nI); anything else would be silently permuted
  for (size_t k = 0; k < p->interior.size(); ++k)
    if (p->interior[k] != (int)k + 1)
      return fail(ctx, GHB_EUNSUPPORTED, "ghb_scatter_free_dof_values: interior fields must be 1..nI in ascending order");
  cudaSetDevice(ctx->device);
  Arg<double> du(ctx, u, (size_t)ncells * p->n_i, true, false); GHB_TRY(du.rc);
  Arg<double> dl(ctx, lambda_free, (size_t)nlambda, true, false); GHB_TRY(dl.rc);
  Arg<double> dx(ctx, x, (size_t)ncells * p->n_i + nlambda, false, true); GHB_TRY(dx.rc);
  GHB_TRY(launch_scatter_free(ctx, *p, ncells, du.dev, dl.dev, nlambda, dx.dev));
  GHB_TRY(dx.finish());
  return GHB_OK;
}

int ghb_synth_fill_f64(ghb_ctx* ctx, int plan_id, int64_t cell_start, int64_t ncells, uint64_t seed, double* A,
                       double* b) {
  Plan* p = get_plan(ctx, plan_id);
  if (!p) return fail(ctx, GHB_EINVAL, "ghb_synth_fill_f64: bad plan id");
  if (ncells < 0 || cell_start < 0 || !A || !b) return fail(ctx, GHB_EINVAL, "ghb_synth_fill_f64: bad argument");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  Arg<double> dA(ctx, A, (size_t)ncells * p->lenA, false, true); GHB_TRY(dA.rc);
  Arg<double> db(ctx, b, (size_t)ncells * p->lenb, false, true); GHB_TRY(db.rc);
  GHB_TRY(launch_synth_fill(ctx, *p, cell_start, ncells, seed, dA.dev, db.dev));
  GHB_TRY(dA.finish()); GHB_TRY(db.finish());
  return GHB_OK;
}

int ghb_cartesian_cell_wise_facets(ghb_ctx* ctx, int D, const int64_t* dims, int64_t cell_start, int64_t ncells,
                                   int64_t* cell_wise_facets) {
  if (!ctx) return GHB_EINVAL;
  if ((D != 2 && D != 3) || !dims || ncells < 0 || cell_start < 0 || !cell_wise_facets)
    return fail(ctx, GHB_EINVAL, "ghb_cartesian_cell_wise_facets: bad argument (D must be 2 or 3)");
  int64_t tot = 1;
  for (int d = 0; d < D; ++d) { if (dims[d] <= 0) return fail(ctx, GHB_EINVAL, "dims must be positive"); tot *= dims[d]; }
  if (cell_start + ncells > tot) return fail(ctx, GHB_EINVAL, "cell range exceeds the mesh");
  if (ncells == 0) return GHB_OK;
  cudaSetDevice(ctx->device);
  Arg<int64_t> dout(ctx, cell_wise_facets, (size_t)ncells * 2 * D, false, true); GHB_TRY(dout.rc);
  GHB_TRY(launch_cartesian_facets(ctx, D, dims, cell_start, ncells, dout.dev));
  GHB_TRY(dout.finish());
  return GHB_OK;
}

}  // extern "C"
