// Single-warp-per-sub-partition latency / issue-rate probe for the float64 pipe on sm_100a (not product code).
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int CHAINS, int OP>
__global__ void k(double* out, long long* clk, int iters, double b, double c) {
  double a[CHAINS];
  float f[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x * 1e-3 + i; f[i] = (float)a[i]; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) a[i] = fma(a[i], b, c);            // DFMA
      if (OP == 1) a[i] = a[i] + c;                   // DADD
      if (OP == 2) f[i] = fmaf(f[i], (float)b, (float)c);  // FFMA
      if (OP == 3) f[i] = ex2f(f[i]);                 // MUFU
      if (OP == 4) { f[i] = (float)a[i]; a[i] += (double)f[i] * 1e-30 + c; }   // F2F.F32.F64 + F2F.F64.F32 + DFMA
      if (OP == 5) a[i] = fmax(a[i] * b, c);          // DMUL + DMNMX
    }
  }
  const long long t1 = clock64();
  double s = 0; float sf = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { s += a[i]; sf += f[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sf;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int CHAINS, int OP>
void run(const char* name, int threads) {
  double* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 8);
  const int iters = 4096;
  k<CHAINS, OP><<<148, threads>>>(out, clk, iters, 1.0000001, 1e-9);
  k<CHAINS, OP><<<148, threads>>>(out, clk, iters, 1.0000001, 1e-9);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-34s threads/SM %4d chains %d: %7.2f cycles per op-per-warp (%.2f per iteration)\n", name, threads, CHAINS, (double)h / iters / CHAINS, (double)h / iters);
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<1, 0>("DFMA dependent", 128); run<2, 0>("DFMA", 128); run<4, 0>("DFMA", 128); run<8, 0>("DFMA", 128); run<8, 0>("DFMA", 512);
  run<1, 1>("DADD dependent", 128); run<8, 1>("DADD", 128);
  run<1, 2>("FFMA dependent", 128); run<8, 2>("FFMA", 128);
  run<1, 3>("MUFU.EX2 dependent", 128); run<8, 3>("MUFU.EX2", 128);
  run<1, 4>("F2F.32.64 + F2F.64.32 + DFMA dep", 128); run<8, 4>("F2F pair + DFMA", 128);
  run<1, 5>("DMUL + fmax dependent", 128); run<8, 5>("DMUL + fmax", 128);
  return 0;
}
