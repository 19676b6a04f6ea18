// Micro-benchmark (round 2): what bounds the S = 4096 attention softmax loop on B200?
//   1. tcgen05.ld throughput (32x32b.x32 = 4 KB per warp instruction) with 4 / 8 warps per SM, 1 / 2 CTAs per SM
//   2. MUFU.EX2 throughput with 4 / 8 / 16 warps per SM
//   3. both together + the FMA-pipe work of the softmax fast path (fmax, ffma, fadd, pack)
//   4. a polynomial exp2 on the FMA pipe (candidate for off-loading a fraction of the exponentials)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench_softmax_pipes tools/ubench/ubench_softmax_pipes.cu
// Prints cycles per 128x128 score tile per SM for each variant (the attention kernel's unit; MUFU bound = 1024).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
// exp2 on the FMA pipe: round-to-nearest split with the magic-number trick, degree-3 minimax on [-0.5, 0.5], exponent by
// integer add.  Relative error ~1e-4 (below the bf16 rounding of P, 4e-3).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);           // in [-0.5, 0.5]
  float p = fmaf(f, 0.0558011f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

enum { V_LDTM = 1, V_MUFU = 2, V_FMA = 4, V_POLY25 = 8, V_STTM = 16, V_NOMAX = 32, V_NOSUM = 64 };

// one "tile" for a warp = 32 rows x 128 columns (its lane quarter of a 128x128 score tile): 4 loads of 32 columns
template <int V>
__global__ void __launch_bounds__(128) bench(int iters, float scale, float mref, float* out, long long* cycles) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_slot + ((uint32_t)(warp * 32) << 16);
  float acc = 0.f, mx = -1e30f;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.001f * (float)(threadIdx.x + i));
  if (V & V_LDTM) {                                  // initialise the columns that are read
    for (int c = 0; c < 8; ++c) { tmem_st16(tbase + c * 16, r); }
    tmem_st_wait();
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (V & V_LDTM) { tmem_ld32(tbase + (uint32_t)((c & 3) * 32), r); tmem_ld_wait(); }
      uint32_t pw[16];
      float l8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, m8[8] = {-1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f, -1e30f};
      float pv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float sv = __uint_as_float(r[i]);
        if (!(V & V_LDTM)) sv += acc * 1e-30f;       // keep a dependence on the loop so nothing is hoisted
        if ((V & V_FMA) && !(V & V_NOMAX)) m8[i & 7] = fmaxf(m8[i & 7], sv);
        const float a = (V & V_FMA) ? fmaf(sv, scale, -mref) : sv;
        float e;
        if ((V & V_POLY25) && (i & 3) == 3) e = ex2_poly(a);
        else if (V & V_MUFU) e = ex2f(a);
        else e = a;
        pv[i] = e;
        if (!(V & V_NOSUM)) l8[i & 7] += e;
      }
      if (V & V_FMA) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pw[i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
        if (V & V_STTM) tmem_st16(tbase + 128 + (uint32_t)(c * 16), pw);
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += __uint_as_float(pw[i] & 0x3f800000u) * 1e-30f;
        }
      }
      if (V & V_NOSUM) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) acc += pv[i] * 1e-30f;
      }
      acc += ((l8[0] + l8[1]) + (l8[2] + l8[3])) + ((l8[4] + l8[5]) + (l8[6] + l8[7]));
      mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7]))));
    }
    if (V & V_STTM) tmem_st_wait();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + mx;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(256) : "memory");
}

template <int V>
void run(const char* name, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 2 * 128);
  cudaMalloc(&cyc, sizeof(long long));
  const int iters = 2000;
  bench<V><<<sms * ctas_per_sm, 128>>>(10, 1.1f, 3.0f, out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<V><<<sms * ctas_per_sm, 128>>>(iters, 1.1f, 3.0f, out, cyc);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long c = 0; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  // per CTA one iteration = one 128x128 tile (4 warps x 32 rows x 128 columns); per SM: ctas_per_sm tiles in parallel
  printf("%-58s CTAs/SM %d: %8.1f cycles per 128x128 tile per SM (%.3f ms, %s)\n", name, ctas_per_sm,
         (double)c / iters / ctas_per_sm, ms, cudaGetErrorString(err));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int n = 1; n <= 2; ++n) {
    run<V_LDTM>("tcgen05.ld only (64 KB per tile)", n);
    run<V_MUFU>("MUFU.EX2 only (16384 per tile) + sum", n);
    run<V_MUFU | V_NOSUM>("MUFU.EX2 only, no sum", n);
    run<V_LDTM | V_MUFU>("tcgen05.ld + MUFU + sum", n);
    run<V_LDTM | V_MUFU | V_FMA>("ld + fmax + ffma + MUFU + sum + pack (fast path)", n);
    run<V_LDTM | V_MUFU | V_FMA | V_STTM>("fast path + tcgen05.st of P", n);
    run<V_LDTM | V_MUFU | V_FMA | V_STTM | V_NOMAX>("fast path + st, no max", n);
    run<V_LDTM | V_MUFU | V_FMA | V_STTM | V_NOMAX | V_NOSUM>("fast path + st, no max, no sum", n);
    run<V_LDTM | V_MUFU | V_FMA | V_STTM | V_POLY25>("fast path + st, 1/4 of the exponentials by polynomial", n);
    run<V_LDTM | V_MUFU | V_FMA | V_STTM | V_POLY25 | V_NOMAX | V_NOSUM>("fast path + st, 1/4 polynomial, no max, no sum", n);
    run<V_FMA | V_POLY25 | V_MUFU>("no ld: fmax + ffma + 3/4 MUFU + 1/4 polynomial + sum + pack", n);
  }
  return 0;
}
