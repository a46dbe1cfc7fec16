// Timeline trace of condense_dmma_ll_kernel<34,36> (CTA 0, cells 2-4): clock64() at phase boundaries.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/trace_ll tools/trace_ll.cu
//   GHB_MAX_CTAS_PER_SM=1 tools/_bin/trace_ll 65536      (unloaded)   /   tools/_bin/trace_ll 262144   (8 CTAs per SM)
// Events.  Panel warp (warp 0): 0 cell start, 1 loader issued, 2 interior rows in shared memory; per panel p: 4+6p column
// tile p ready (panel starts), 5+6p panel published, 6+6p inv(U_pp) done.  Update warps: 4+6p panel p published (arrival),
// 5+6p inv(L_pp) ready, 6+6p first owned column tile done, 7+6p owned column tiles done.  All warps: 40 top block done
// (arrival at the barrier), 41 bottom block starts, 42 / 46 row tile (first / second) starts, 44 / 48 its panels done,
// 50 stores issued, 51 cell end.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define GHB_TRACE 1
#include "../gridaphybrid.jl_b200/csrc/condense_dmma.cu"
namespace ghb { int fail(ghb_ctx*, int code, const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); return code; } }
int main(int argc, char** argv) {
  using namespace ghb;
  const int64_t ncells = argc > 1 ? atoll(argv[1]) : 65536;
  setenv("GHB_DMMA_LL", "1", 1);
  ghb_ctx ctx; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); ctx.sm_count = pr.multiProcessorCount; ctx.stream = 0;
  Plan p; p.nfields = 3; p.ndofs = {30, 4, 36}; p.interior = {1, 2}; p.boundary = {3}; p.touched.assign(9, 1);
  p.block_offset.assign(9, -1); int64_t off = 0;
  for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) { p.block_offset[i + 3 * j] = off; off += p.ndofs[i] * p.ndofs[j]; }
  p.lenA = (int)off; p.field_offset_b = {0, 30, 34}; p.lenb = 70; p.n_i = 34; p.n_b = 36; p.n = 70;
  for (int f = 0; f < 3; ++f) for (int l = 0; l < p.ndofs[f]; ++l) { p.row_field.push_back(f); p.row_local.push_back(l); }
  dmma_prepare(&ctx, p);
  double *A, *b, *S, *g; int* info;
  cudaMalloc(&A, ncells * p.lenA * 8); cudaMalloc(&b, ncells * 70 * 8); cudaMalloc(&S, ncells * 1296 * 8); cudaMalloc(&g, ncells * 36 * 8); cudaMalloc(&info, ncells * 4);
  std::vector<double> h((size_t)ncells * p.lenA);
  srand(1); for (auto& x : h) x = rand() / (double)RAND_MAX - 0.5;
  for (int64_t c = 0; c < ncells; ++c) for (int i = 0; i < 30; ++i) h[c * p.lenA + i * 31] += 7.0;
  cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice); cudaMemset(b, 0, ncells * 70 * 8);
  for (int r = 0; r < 2; ++r) launch_condense_dmma(&ctx, p, ncells, A, b, S, g, info);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<long long> t(64 * 4 * 64);
  cudaMemcpyFromSymbol(t.data(), g_trace, t.size() * 8);
  for (int cell = 2; cell < 5; ++cell) {
    long long t0 = t[(cell * 4 + 0) * 64 + 0];
    for (int w = 0; w < 4; ++w) {
      printf("cell %d warp %d:", cell, w);
      for (int e = 0; e < 52; ++e) { long long v = t[(cell * 4 + w) * 64 + e]; if (v) printf(" e%d=%lld", e, v - t0); }
      printf("\n");
    }
  }
  return 0;
}
