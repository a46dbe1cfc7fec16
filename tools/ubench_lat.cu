// Latency microbenchmark for the primitives on the panel warp's critical chain (single warp, dependent chains).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int iters, double seed) {
  const int lane = threadIdx.x & 31;
  long long t0, t1;
  // 1. REDUX.MAX dependent chain
  unsigned v = lane * 7919u + 13u;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __reduce_max_sync(0xffffffffu, v ^ (unsigned)(lane + i));
  t1 = clock64();
  if (lane == 0) out[0] = (t1 - t0); out[8] = v;
  // 2. SHFL dependent chain
  unsigned s = lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) s = __shfl_sync(0xffffffffu, s + 1u, (s + i) & 31);
  t1 = clock64();
  if (lane == 0) out[1] = (t1 - t0); out[9] = s;
  // 3. DFMA dependent chain
  double x = seed + lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64();
  if (lane == 0) out[2] = (t1 - t0); out[10] = (long long)x;
  // 4. rcp.approx + 2 Newton
  double r = seed + 1.5;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r));
    double e = fma(-r, y, 1.0); y = fma(y, e, y); e = fma(-r, y, 1.0); y = fma(y, e, y);
    r = y + 1.25;
  }
  t1 = clock64();
  if (lane == 0) out[3] = (t1 - t0); out[11] = (long long)(r * 1e6);
  // 5. ballot + ffs chain
  unsigned bq = lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { unsigned m = __ballot_sync(0xffffffffu, ((bq + i) & 3) == 0); bq = __ffs(m) + lane; }
  t1 = clock64();
  if (lane == 0) out[4] = (t1 - t0); out[12] = bq;
  // 6. one synthetic pivot step chain: key -> REDUX -> decode -> SHFL(double) -> DMUL -> DFMA -> key
  double a = seed * (lane + 1), piv = 1.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    unsigned h = (unsigned)(__double_as_longlong(a) >> 32) & 0x7fffffffu;
    unsigned key = ((h >> 5) << 6) | (unsigned)(63 - lane);
    unsigned km = __reduce_max_sync(0xffffffffu, key);
    int q = 31 - (int)(km & 31u);
    double rv = __shfl_sync(0xffffffffu, piv, q);
    double l = a * rv;
    a = fma(-l, 0.5, a) + 1e-3;
  }
  t1 = clock64();
  if (lane == 0) out[5] = (t1 - t0); out[13] = (long long)(a * 1e6);
  // 7. LDS dependent chain (pointer chasing in shared memory)
  __shared__ int sm[64];
  sm[lane] = (lane * 5 + 1) & 31; sm[lane + 32] = lane; __syncwarp();
  int pidx = lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) pidx = sm[pidx];
  t1 = clock64();
  if (lane == 0) out[6] = (t1 - t0); out[14] = pidx;
  // 8. DMMA dependent chain
  double c0 = lane, c1 = 1.0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(1.0000001), "d"(0.999));
  t1 = clock64();
  if (lane == 0) out[7] = (t1 - t0); out[15] = (long long)(c0 + c1);
}
int main() {
  long long* d; cudaMalloc(&d, 16 * 8);
  const int iters = 4096;
  k<<<1, 32>>>(d, iters, 1.37); cudaDeviceSynchronize();
  k<<<1, 32>>>(d, iters, 1.37); cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[8] = {"REDUX.MAX", "SHFL.IDX", "DFMA", "rcp.approx+2Newton+add", "ballot+ffs", "pivot-step chain", "LDS", "DMMA"};
  for (int i = 0; i < 8; ++i) printf("%-24s %7.1f cycles/iter\n", names[i], (double)h[i] / iters);
  return 0;
}
