// fp64 / fp32 instruction throughput per SM on this GPU: 148 CTAs x 256 threads (8 warps per SM, the wide solver's shape),
// ILP 8 independent chains per thread.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_rate.cu -o fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, double seed) {
  double a[8];
  float f[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; f[i] = (float)a[i]; }
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma(a[i], m, c);
      if (OP == 1) a[i] = __dadd_rn(a[i], c);
      if (OP == 2) a[i] = __dmul_rn(a[i], m);
      if (OP == 3) a[i] = a[i] < a[(i + 1) & 7] ? a[i] + c : a[(i + 1) & 7];  // compare + select + add
      if (OP == 4) f[i] = fmaf(f[i], 1.0000001f, 1e-9f);
      if (OP == 5) a[i] = sqrt(a[i] + 2.0);
      if (OP == 6) a[i] = a[i] / (m + a[(i + 1) & 7]);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + f[i];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, double* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<148, 256>>>(out, iters, 1.0);
  cudaEventRecord(e0);
  k<OP><<<148, 256>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr_per_sm = 8.0 * 8 * iters;  // warps x chains x iterations
  printf("%-28s %8.3f ms  %7.2f cycles per warp-instruction per SM (1.965 GHz)  = %6.1f lanes/clk/SM\n", name, ms,
         ms * 1e-3 * 1.965e9 / warp_instr_per_sm, 32.0 / (ms * 1e-3 * 1.965e9 / warp_instr_per_sm));
}
int main() {
  double* out; cudaMalloc(&out, 148 * 256 * 8);
  run<0>("DFMA", out, 20000); run<1>("DADD", out, 20000); run<2>("DMUL", out, 20000); run<3>("DSETP+SEL+DADD", out, 20000);
  run<4>("FFMA", out, 20000); run<5>("sqrt(double)+DADD", out, 2000); run<6>("double division", out, 2000);
  return 0;
}
