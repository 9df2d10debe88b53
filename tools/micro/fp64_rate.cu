// Microbenchmark: FP64 pipe rates on this GPU (DFMA / DADD / DMUL, F2I.F64, I2F.F64, DDIV, DSQRT, MUFU.RCP64H)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3 + a, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a); }
    if (OP == 1) { x0 = x0 + a; x1 = x1 + a; x2 = x2 + a; x3 = x3 + a; x4 = x4 + a; x5 = x5 + a; x6 = x6 + a; x7 = x7 + a; }
    if (OP == 2) { acc += (int)x0 + (int)x1 + (int)x2 + (int)x3 + (int)x4 + (int)x5 + (int)x6 + (int)x7; x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a; }
    if (OP == 3) { x0 = x0 / b; x1 = x1 / b; x2 = x2 / b; x3 = x3 / b; x4 = x4 / b; x5 = x5 / b; x6 = x6 / b; x7 = x7 / b; }
    if (OP == 4) { x0 = sqrt(x0) + a; x1 = sqrt(x1) + a; x2 = sqrt(x2) + a; x3 = sqrt(x3) + a; x4 = sqrt(x4) + a; x5 = sqrt(x5) + a; x6 = sqrt(x6) + a; x7 = sqrt(x7) + a; }
    if (OP == 5) { acc += (long long)(unsigned long long)x0 + (long long)(unsigned long long)x1 + (long long)(unsigned long long)x2 + (long long)(unsigned long long)x3; x0 += a; x1 += a; x2 += a; x3 += a; }
    if (OP == 6) { float f0 = (float)x0, f1 = (float)x1, f2 = (float)x2, f3 = (float)x3; f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f); x0 = f0; x1 = f1; x2 = f2; x3 = f3; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (double)acc;
}
template <int OP>
void run(const char* name, double ops_per_iter) {
  double* out;
  int blocks = 148 * 8, threads = 256, iters = 4096;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  k<OP><<<blocks, threads>>>(out, 16, 1e-9, 0.999999);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, iters, 1e-9, 0.999999);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double total = (double)blocks * threads * iters * ops_per_iter;
  printf("%-28s %8.3f ms  %8.2f Gop/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ms, total / ms / 1e6, total / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("DFMA", 8); run<1>("DADD", 8); run<2>("F2I.F64->S32 (+DADD)", 8); run<3>("DDIV", 8); run<4>("DSQRT(+DADD)", 8);
  run<5>("F2I.F64->U64 (+DADD)", 4); run<6>("F64->F32 FFMA F32->F64", 4);
  return 0;
}
