// micro-benchmark: issue throughput of scalar FFMA/FMUL/FADD vs packed FFMA2/FMUL2/FADD2 on sm_100a
// (8 independent chains per thread, 1024 threads per SM x 148 SMs).  Prints warp-instructions / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float a, float nz, int iters)
{
  float2 v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = make_float2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f + j);
  const float2 A = make_float2(a, a), NZ = make_float2(nz, nz);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE == 0) { v[j].x = __fmaf_rn(v[j].x, a, nz); v[j].y = __fmaf_rn(v[j].y, a, nz); }   // 2 FFMA
      if (MODE == 1) { v[j] = __ffma2_rn(v[j], A, NZ); }                                       // 1 FFMA2
      if (MODE == 2) { v[j].x = __fadd_rn(v[j].x, a); v[j].y = __fadd_rn(v[j].y, a); }
      if (MODE == 3) { v[j] = __fadd2_rn(v[j], A); }
      if (MODE == 4) { v[j].x = __fmul_rn(v[j].x, a); v[j].y = __fmul_rn(v[j].y, a); }
      if (MODE == 5) { v[j] = __fmul2_rn(v[j], A); }
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += v[j].x + v[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_iter)
{
  float* out; cudaMalloc(&out, 148 * 4 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  k<MODE><<<148 * 2, 1024>>>(out, 1.0000001f, -0.f, 100);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, 1.0000001f, -0.f, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double winst = 148.0 * 2 * 32 * (double)iters * per_iter;   // warp instructions
  printf("%-8s %8.3f ms  %.3e warp-inst/s  = %.2f per SM per clk @1.965GHz  (%.1f Gflop-lanes/s/SM)\n", name, ms,
         winst / (ms * 1e-3), winst / (ms * 1e-3) / 148 / 1.965e9, 0.);
  cudaFree(out);
}
int main()
{
  run<0>("FFMA", 16); run<1>("FFMA2", 8); run<2>("FADD", 16); run<3>("FADD2", 8); run<4>("FMUL", 16); run<5>("FMUL2", 8);
  return 0;
}
