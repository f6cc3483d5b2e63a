// FP64 pipe microbenchmarks on B200: dependent-issue latency and per-SMSP throughput vs ILP/warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rcpchain(double* out, long long* cyc, int iters) {
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps, const char* tag) {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  int iters = 4096;
  chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per = double(c) / iters;  // cycles per ILP-group per warp
  printf("%s ILP=%d warps=%d (per SMSP %.1f): %.2f cycles per iteration -> %.3f DFMA/cycle/SMSP\n", tag, ILP, warps,
         warps / 4.0, per, ILP * (warps / 4.0) / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(1, "dep-latency"); run<2>(1, "ilp"); run<4>(1, "ilp"); run<8>(1, "ilp");
  run<1>(4, "tlp"); run<1>(8, "tlp"); run<1>(16, "tlp"); run<1>(32, "tlp");
  run<2>(16, "mix"); run<2>(24, "mix"); run<4>(16, "mix"); run<2>(32, "mix");
  double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  rcpchain<<<1, 32>>>(out, cyc, 4096); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H dependent latency: %.2f cycles\n", double(c) / 4096);
  return 0;
}
