// Does a half-rate FP64 instruction block the SMSP's issue port for its second cycle?
// Per thread: ND independent DFMA chains, NF independent FFMA chains, NX MUFU chains, interleaved in one
// unrolled loop body; 16 warps per SM (4 per SMSP), one CTA.  If the issue port is shared 1:1 the time
// per iteration is 2*ND + NF (+NX); if FP64 only occupies its own pipe it is max(2*ND, ND + NF + NX).
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NF, int NX>
__global__ void mix(double* out, long long* cyc, int iters, double a, double b, float af, float bf) {
  double x[ND > 0 ? ND : 1];
  float y[NF > 0 ? NF : 1];
  float z[NX > 0 ? NX : 1];
#pragma unroll
  for (int i = 0; i < ND; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NF; ++i) y[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < NX; ++i) z[i] = 1.5f + threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < ND) x[i] = fma(x[i], a, b);
      if (i < NF) y[i] = fmaf(y[i], af, bf);
      if (i + 8 < NF) y[i + 8] = fmaf(y[i + 8], af, bf);
      if (i < NX) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(z[i]) : "f"(z[i]));
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ND; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += y[i];
#pragma unroll
  for (int i = 0; i < NX; ++i) s += z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ND, int NF, int NX>
void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int r = 0; r < 2; ++r) { mix<ND, NF, NX><<<1, warps * 32>>>(out, cyc, iters, 0.999999, 1e-7, 0.999f, 1e-3f); cudaDeviceSynchronize(); }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = double(c) / iters / (warps / 4.0);  // SMSP cycles per (warp, iteration)
  printf("DFMA %d FFMA %2d MUFU %d warps %2d: %6.2f cyc/iter/warp | port-shared %2d, own-pipe %2d, xu-bound %2d\n", ND, NF, NX,
         warps, per, 2 * ND + NF + NX, (2 * ND > ND + NF + NX ? 2 * ND : ND + NF + NX), 8 * NX);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<4, 0, 0>(16); run<0, 8, 0>(16); run<4, 4, 0>(16); run<4, 8, 0>(16); run<4, 12, 0>(16); run<4, 16, 0>(16);
  run<8, 8, 0>(16); run<8, 16, 0>(16);
  run<0, 0, 2>(16); run<4, 0, 1>(16); run<4, 4, 1>(16); run<4, 8, 1>(16); run<4, 8, 2>(16);
  run<4, 4, 0>(8); run<4, 8, 0>(8);
  return 0;
}
