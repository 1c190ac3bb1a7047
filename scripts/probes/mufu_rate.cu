// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/mufu_rate scripts/probes/mufu_rate.cu   (measured on B200: 16 tanh / ex2 per clock per SM)
// MUFU throughput probe: tanh.approx.f32 vs ex2.approx.f32 vs rcp.approx.f32 per SM per clock (B200).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d));
    } else if (OP == 1) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else {
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  const char* names[3] = {"tanh.approx.f32", "ex2.approx.ftz.f32", "rcp.approx.ftz.f32"};
  for (int op = 0; op < 3; ++op)
    for (int threads : {128, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, threads>>>(out, iters, cyc);
        if (op == 1) k<1><<<148, threads>>>(out, iters, cyc);
        if (op == 2) k<2><<<148, threads>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      printf("%-20s %4d threads/SM: %.2f ops/clk/SM\n", names[op], threads, 4.0 * iters * threads / (double)*cyc);
    }
  return 0;
}
