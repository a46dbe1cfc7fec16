// Microbenchmark: FP64 pipe peak on B200 -- DFMA vs DMMA (mma.sync f64 shapes).
// Decides whether trailing updates go on DMMA or on plain FMA warp tiles (kernel choice).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs
template <int ILP>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs
template <int ILP>
__global__ void k_dmma1684(double* out, int iters, double a, double b) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs
template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA issue mixed with an equal number of independent DFMA per thread: do they share the pipe?
template <int ILP>
__global__ void k_mix(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; f[i] = i + 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
      f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32; int blocks = sms * 2;
    if (warps == 32) blocks = sms;  // 1024 threads/block, 1 block/SM => 32 warps/SM; else 2 blocks/SM
    double nthreads = (double)threads * blocks;
    {
      constexpr int ILP = 8;
      float ms = time_it([&] { k_dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk %2d  DFMA ilp8       : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * nthreads * ILP * iters / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      float ms = time_it([&] { k_dmma884<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk %2d  DMMA m8n8k4 ilp4: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * (nthreads / 32) * 256 * ILP * iters / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      float ms = time_it([&] { k_dmma1684<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk %2d  DMMA m16n8k4    : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * (nthreads / 32) * 512 * ILP * iters / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      float ms = time_it([&] { k_dmma1688<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk %2d  DMMA m16n8k8    : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * (nthreads / 32) * 1024 * ILP * iters / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      float ms = time_it([&] { k_dmma16816<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk %2d  DMMA m16n8k16   : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, 2.0 * (nthreads / 32) * 2048 * ILP * iters / ms * 1e-9);
    }
    {
      constexpr int ILP = 4;
      float ms = time_it([&] { k_mix<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk %2d  MIX 884+DFMA    : %8.3f ms  %7.2f TFLOP/s (sum)\n", warps, ms, 2.0 * (nthreads / 32) * (256 + 32) * ILP * iters / ms * 1e-9);
    }
  }
  // latency: single warp, dependent chain
  {
    float ms = time_it([&] { k_dmma884<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4 dependent chain: %.1f ns/op\n", ms * 1e6 / iters);
    ms = time_it([&] { k_dfma<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA dependent chain: %.1f ns/op\n", ms * 1e6 / iters);
    ms = time_it([&] { k_dmma1688<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA m16n8k8 dependent chain: %.1f ns/op\n", ms * 1e6 / iters);
  }
  cudaFree(out);
  return 0;
}
