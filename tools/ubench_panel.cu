// Unloaded cost of the serial pieces of condense_dmma_kernel<34,36>: one warp alone on an SM runs the panel
// factorisation, the two 8x8 inverses and the back substitution loop on a random image; cycles by clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/ubench_panel tools/ubench_panel.cu
#include <cstdio>
#include <cstdlib>
#include "../gridaphybrid.jl_b200/csrc/condense_dmma.cu"

using namespace ghb;
namespace ghb { int fail(ghb_ctx*, int code, const std::string&) { return code; } }

__global__ void k_panel(const double* __restrict__ src, long long* __restrict__ out, double* __restrict__ sink) {
  constexpr int NI = 34, LDW = 36;
  __shared__ double Wt[72 * LDW];
  __shared__ PanelCtl ctl;
  __shared__ int info;
  const int lane = threadIdx.x;
  for (int i = lane; i < 72 * LDW; i += 32) Wt[i] = src[i];
  if (lane == 0) info = 0;
  __syncwarp();
  long long t[12];
  for (int rep = 0; rep < 2; ++rep) {      // second repetition: instruction cache warm
    for (int i = lane; i < 72 * LDW; i += 32) Wt[i] = src[i];
    __syncwarp();
    t[0] = clock64();
    panel_factor<NI, LDW, true>(Wt, 0, 8, &ctl, &info);
    __syncwarp();
    t[1] = clock64();
    panel_factor<NI, LDW, false>(Wt, 8, 8, &ctl, &info);
    __syncwarp();
    t[2] = clock64();
    panel_factor<NI, LDW, false>(Wt, 16, 8, &ctl, &info);
    __syncwarp();
    t[3] = clock64();
    panel_factor<NI, LDW, false>(Wt, 32, 2, &ctl, &info);
    __syncwarp();
    t[4] = clock64();
    invert_unit_lower<LDW>(Wt + 8 + LDW * 8, 8, ctl.Linv);
    __syncwarp();
    t[5] = clock64();
    invert_upper<LDW>(Wt + 8 + LDW * 8, 8, ctl.rinv, ctl.Dinv);
    __syncwarp();
    t[6] = clock64();
  }
  if (lane == 0) for (int i = 0; i < 7; ++i) out[i] = t[i];
  sink[lane] = Wt[lane] + ctl.Linv[lane] + ctl.Dinv[lane];
}

int main() {
  const int n = 72 * 36;
  double* h = (double*)malloc(n * sizeof(double));
  srand(1);
  for (int i = 0; i < n; ++i) h[i] = (double)rand() / RAND_MAX - 0.5;
  double *d, *sink; long long* out;
  cudaMalloc(&d, n * sizeof(double)); cudaMalloc(&sink, 32 * sizeof(double)); cudaMalloc(&out, 16 * sizeof(long long));
  cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice);
  k_panel<<<1, 32>>>(d, out, sink);
  long long t[16];
  cudaMemcpy(t, out, 7 * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  const char* names[] = {"panel 0 (34 rows, two register sets, 8 pivots)", "panel 1 (26 rows, 8 pivots)",
                         "panel 2 (18 rows, 8 pivots)", "panel 4 (2 rows, 2 pivots)", "invert_unit_lower 8x8",
                         "invert_upper 8x8"};
  for (int i = 0; i < 6; ++i) printf("%-50s %6lld cycles\n", names[i], t[i + 1] - t[i]);
  return 0;
}
