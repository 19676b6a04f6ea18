// Micro-benchmark (round 2): ex2.approx.ftz.f32 throughput per SM as a function of warps per scheduler and of the FMA-pipe
// instructions interleaved per exponential.  16 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench_mufu tools/ubench/ubench_mufu.cu
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int NFMA>
__global__ void k(int iters, float a, float b, float* out, long long* cyc) {
  float x[16], y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = 0.001f * (threadIdx.x + i); y[i] = 0.5f + 0.01f * i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      x[i] = ex2f(x[i] * 0.25f - 1.0f * (NFMA < 0));     // one FMUL feeding the MUFU (kept tiny so values stay bounded)
#pragma unroll
      for (int f = 0; f < (NFMA > 0 ? NFMA : 0); ++f) y[i] = fmaf(y[i], a, b);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NFMA>
void run(int threads) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<NFMA><<<sms, threads>>>(10, 0.999f, 0.001f, out, cyc);
  k<NFMA><<<sms, threads>>>(iters, 0.999f, 0.001f, out, cyc);
  cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double ex2 = (double)iters * 16 * threads;
  printf("warps/scheduler %d, FMA-pipe instr per ex2 %d (+1 FMUL): %6.2f ex2 / clk / SM   (%6.2f issue slots / clk / scheduler)\n", threads / 128,
         NFMA > 0 ? NFMA : 0, ex2 / c, ex2 * ((NFMA > 0 ? NFMA : 0) + 2) / 32.0 / 4.0 / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>(threads); run<1>(threads); run<2>(threads); run<3>(threads); run<4>(threads); run<6>(threads);
  }
  return 0;
}
