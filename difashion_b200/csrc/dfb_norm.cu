// GroupNorm (+SiLU) and LayerNorm producing the bf16 A operands of the following GEMM / conv.
// Statistics are fp32 (diffusers keeps norms in fp32 under autocast).  Inputs are
// the fp32 residual stream (NHWC); the skip-concat of the up blocks is read from its two sources
// directly (never materialised in fp32) — groups may straddle the concat seam (C=1920, 960).
#include "dfb_host.h"
#include "../../include/dfb200.h"

namespace dfb {

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics, deterministic (no atomics, fixed reduction order -> bitwise reproducible):
//   stage 1: grid = (pixel chunks, B); each thread owns fixed float4 channel-vectors (coalesced
//            across the warp) and walks its share of the chunk's pixels; per-thread partials go to
//            shared memory and one thread per group folds them in thread order; the block writes
//            partial[b][chunk][g] = (sum, sumsq).
//   stage 2: one warp per (b, g) folds the chunks in order -> stats[b][g] = (mean, rstd).
// Channels-per-group is even, so a float4 spans <= 2 groups.
// ---------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_VEC_PER_THREAD = 4;   // supports C <= 4096
constexpr int GN_MAX_CHUNKS = 256;         // partial sums per image (workspace sizing)
constexpr int GN_BATCH = 8;                // loads in flight per thread in the apply kernel

__global__ void __launch_bounds__(GN_THREADS)
groupnorm_stats_kernel(const float* __restrict__ src0, int c0, int ld0, const float* __restrict__ src1, int c1, int ld1,
                       int hw, int groups, int pix_per_block, float* __restrict__ partial) {
  __shared__ float s_part[GN_THREADS * GN_MAX_VEC_PER_THREAD][4];   // (sum_a, sq_a, sum_b, sq_b) per thread-vector
  __shared__ short s_ga[GN_THREADS * GN_MAX_VEC_PER_THREAD], s_gb[GN_THREADS * GN_MAX_VEC_PER_THREAD];
  const int b = blockIdx.y;
  const int C = c0 + c1;
  const int cg = C / groups;
  const int nvec = C >> 2;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  // thread -> (channel vector, pixel lane): with few channels several pixel lanes share a block
  const int vlanes = nvec < GN_THREADS ? nvec : GN_THREADS;
  const int plane_cnt = GN_THREADS / vlanes;            // >= 1
  const int tv = threadIdx.x % vlanes;
  const int tp = threadIdx.x / vlanes;
  int slot = threadIdx.x;
#pragma unroll
  for (int it = 0; it < GN_MAX_VEC_PER_THREAD; ++it, slot += GN_THREADS) {
    const int v = tv + it * GN_THREADS;
    float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
    int ga = -1, gb = -1;
    if (v < nvec && tp < plane_cnt) {
      const int c = v << 2;
      const float* base;
      int ld, cc;
      if (c < c0) { base = src0; ld = ld0; cc = c; } else { base = src1; ld = ld1; cc = c - c0; }
      base += (size_t)b * hw * ld + cc;
      for (int px = p_begin + tp; px < p_end; px += plane_cnt) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(base + (size_t)px * ld));
        sa += x.x + x.y; qa += x.x * x.x + x.y * x.y;
        sb += x.z + x.w; qb += x.z * x.z + x.w * x.w;
      }
      ga = c / cg; gb = (c + 2) / cg;
    }
    s_part[slot][0] = sa; s_part[slot][1] = qa; s_part[slot][2] = sb; s_part[slot][3] = qb;
    s_ga[slot] = (short)ga; s_gb[slot] = (short)gb;
  }
  __syncthreads();
  // fixed-order fold: warp w handles groups w, w+8, ...; lanes stride the slots, then a shuffle tree
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = GN_THREADS * ((nvec + GN_THREADS - 1) / GN_THREADS);
  for (int g = warp; g < groups; g += GN_THREADS / 32) {
    float s = 0.f, q = 0.f;
    for (int i = lane; i < nslots; i += 32) {
      if (s_ga[i] == g) { s += s_part[i][0]; q += s_part[i][1]; }
      if (s_gb[i] == g) { s += s_part[i][2]; q += s_part[i][3]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      float* dst = partial + (((size_t)b * gridDim.x + blockIdx.x) * groups + g) * 2;
      dst[0] = s; dst[1] = q;
    }
  }
}

__global__ void groupnorm_finalize_kernel(const float* __restrict__ partial, int chunks, int groups, int B, float inv_n,
                                          float eps, float* __restrict__ stats) {
  // one warp per (b, g)
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = w / groups, g = w - b * groups;
  if (b >= B) return;
  float s = 0.f, q = 0.f;
  for (int c = lane; c < chunks; c += 32) {
    const float* src = partial + (((size_t)b * chunks + c) * groups + g) * 2;
    s += src[0]; q += src[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    stats[((size_t)b * groups + g) * 2 + 0] = mean;
    stats[((size_t)b * groups + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// Statistics from the producers' epilogue partials (dfb_gemm_params.gn_partial: [M/32][C/2][2] per source):
// one warp per (b, g) folds the (row block, channel pair) partials of its group in a fixed order.
__global__ void groupnorm_finalize_partials_kernel(const float* __restrict__ p0, int c0, const float* __restrict__ p1, int c1,
                                                   int hw, int groups, int B, float eps, float* __restrict__ stats) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = w / groups, g = w - b * groups;
  if (b >= B) return;
  const int C = c0 + c1;
  const int cg = C / groups;
  const int ppg = cg >> 1;                 // channel pairs per group (cg is even)
  const int pair0 = g * ppg;
  const int rbs = hw >> 5;                 // 32-row blocks per image
  const int h0 = c0 >> 1, h1 = c1 >> 1;
  float s = 0.f, q = 0.f;
  for (int idx = lane; idx < rbs * ppg; idx += 32) {
    const int rb = idx / ppg;
    const int pr = pair0 + (idx - rb * ppg);
    const float* src = pr < h0 ? p0 + ((size_t)(b * rbs + rb) * h0 + pr) * 2
                               : p1 + ((size_t)(b * rbs + rb) * h1 + (pr - h0)) * 2;
    const float2 v = __ldg(reinterpret_cast<const float2*>(src));
    s += v.x; q += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const float inv_n = 1.f / ((float)cg * (float)hw);
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    stats[((size_t)b * groups + g) * 2 + 0] = mean;
    stats[((size_t)b * groups + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// y = (x - mean) * rstd * gamma + beta (optionally SiLU) -> bf16 [B, HW, C] (ld_out), optional raw
// bf16 copy of x (A operand of the 1x1 shortcut folded into conv2).  Same thread mapping as the
// statistics kernel: a thread owns fixed channel vectors, so gamma/beta/mean/rstd are folded into a
// per-thread (scale, shift) once and the pixel loop is load -> FMA -> SiLU -> store (no index math).
template <typename TOut>
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_apply_kernel(const float* __restrict__ src0, int c0, int ld0, const float* __restrict__ src1, int c1, int ld1,
                       int hw, int groups, int pix_per_block, const float* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                       TOut* __restrict__ out, int ld_out, TOut* __restrict__ raw_out, int ld_raw) {
  const int b = blockIdx.y;
  const int C = c0 + c1;
  const int cg = C / groups;
  const int nvec = C >> 2;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  const int vlanes = nvec < GN_THREADS ? nvec : GN_THREADS;
  const int plane_cnt = GN_THREADS / vlanes;
  const int tv = threadIdx.x % vlanes;
  const int tp = threadIdx.x / vlanes;
  if (tp >= plane_cnt) return;
  const float* st = stats + (size_t)b * groups * 2;
  for (int v = tv; v < nvec; v += GN_THREADS) {
    const int c = v << 2;
    const float* base;
    int ld, cc;
    if (c < c0) { base = src0; ld = ld0; cc = c; } else { base = src1; ld = ld1; cc = c - c0; }
    base += (size_t)b * hw * ld + cc;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
    const int ga = c / cg, gb = (c + 2) / cg;
    const float ma = st[ga * 2], ra = st[ga * 2 + 1];
    const float mb = st[gb * 2], rb = st[gb * 2 + 1];
    const float s0 = ra * g.x, s1 = ra * g.y, s2 = rb * g.z, s3 = rb * g.w;
    const float h0 = bt.x - ma * s0, h1 = bt.y - ma * s1, h2 = bt.z - mb * s2, h3 = bt.w - mb * s3;
    TOut* o = out + (size_t)b * hw * ld_out + c;
    TOut* ro = raw_out ? raw_out + (size_t)b * hw * ld_raw + c : nullptr;
    auto emit = [&](int px, const float4& x) {
      float y0 = fmaf(x.x, s0, h0), y1 = fmaf(x.y, s1, h1), y2 = fmaf(x.z, s2, h2), y3 = fmaf(x.w, s3, h3);
      if (silu) {
        if constexpr (sizeof(TOut) == 4) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }   // fp32 path: IEEE division
        else { y0 = silu_fast_f(y0); y1 = silu_fast_f(y1); y2 = silu_fast_f(y2); y3 = silu_fast_f(y3); }
      }
      store4(o + (size_t)px * ld_out, y0, y1, y2, y3);
      if (ro) store4(ro + (size_t)px * ld_raw, x.x, x.y, x.z, x.w);
    };
    // GN_BATCH independent 16-byte loads in flight per thread before any of them is consumed (the compiler does
    // not hoist loads over the stores of an unrolled loop: ncu showed one exposed DRAM round trip per pixel)
    int px = p_begin + tp;
    for (; px + (GN_BATCH - 1) * plane_cnt < p_end; px += GN_BATCH * plane_cnt) {
      float4 xb[GN_BATCH];
#pragma unroll
      for (int k = 0; k < GN_BATCH; ++k) xb[k] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(px + k * plane_cnt) * ld));
#pragma unroll
      for (int k = 0; k < GN_BATCH; ++k) emit(px + k * plane_cnt, xb[k]);
    }
    for (; px < p_end; px += plane_cnt) emit(px, __ldg(reinterpret_cast<const float4*>(base + (size_t)px * ld)));
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim: one warp per row, row held in registers (C <= 2048), two-pass
// (mean, then centred variance) in fp32; bf16 output.
// ---------------------------------------------------------------------------------------------
template <int LN_MAX_VEC, typename TOut>   // float4 per lane -> C <= 128 * LN_MAX_VEC
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, TOut* __restrict__ out, int ld_out, int rows, int C) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 2;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const float* xr = x + (size_t)row * ld_x;
    float4 v[LN_MAX_VEC];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int j = lane + k * 32;
      if (j < nvec) {
        v[k] = __ldg(reinterpret_cast<const float4*>(xr) + j);
        s += v[k].x + v[k].y + v[k].z + v[k].w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int j = lane + k * 32;
      if (j < nvec) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    TOut* orow = out + (size_t)row * ld_out;
#pragma unroll
    for (int k = 0; k < LN_MAX_VEC; ++k) {
      const int j = lane + k * 32;
      if (j < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + j);
        const float y0 = (v[k].x - mean) * rstd * g.x + bt.x;
        const float y1 = (v[k].y - mean) * rstd * g.y + bt.y;
        const float y2 = (v[k].z - mean) * rstd * g.z + bt.z;
        const float y3 = (v[k].w - mean) * rstd * g.w + bt.w;
        store4(orow + (j << 2), y0, y1, y2, y3);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row softmax of a materialised score matrix: out[r, :] = softmax(scale * in[r, :]).  Used by the VAE decoder's
// single-head d = 512 attention (one layer, 4096 tokens per image), which runs as GEMMs around this kernel
// because its head dimension exceeds the flash kernel's TMEM budget.  One block per row, row held in
// registers (cols <= 256 * 4 * SM_MAX_VEC), fp32 math.
// ---------------------------------------------------------------------------------------------
constexpr int SM_MAX_VEC = 8;
template <typename TOut>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ in, int in_ld, float scale,
                                                           TOut* __restrict__ out, int out_ld, int rows, int cols) {
  __shared__ float red[8];
  const int nvec = cols >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(in + (size_t)row * in_ld);
    float4 v[SM_MAX_VEC];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < SM_MAX_VEC; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < nvec) {
        v[k] = __ldg(src + j);
        mx = fmaxf(fmaxf(mx, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    const float m2 = mx * scale;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < SM_MAX_VEC; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < nvec) {
        v[k].x = expf(fmaf(v[k].x, scale, -m2)); v[k].y = expf(fmaf(v[k].y, scale, -m2));
        v[k].z = expf(fmaf(v[k].z, scale, -m2)); v[k].w = expf(fmaf(v[k].w, scale, -m2));
        sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    const float inv = 1.f / sum;
    TOut* orow = out + (size_t)row * out_ld;
#pragma unroll
    for (int k = 0; k < SM_MAX_VEC; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < nvec) store4(orow + (j << 2), v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
    }
  }
}

}  // namespace dfb

using namespace dfb;

static int launch_gn_apply(const float* src0, int c0, int ld0, const float* src1, int c1, int ld1, int B, int hw, int groups,
                           const float* stats, const float* gamma, const float* beta, int silu, void* out, int out_dtype,
                           int ld_out, void* raw_out, int ld_raw, cudaStream_t stream) {
  // ~16 CTAs per SM over (B x pixel chunks); DFB_GN_CTAS_PER_SM overrides (tuning hook, read once)
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    const char* e = getenv("DFB_GN_CTAS_PER_SM");
    ctas_per_sm = e ? atoi(e) : 16;
    if (ctas_per_sm < 1) ctas_per_sm = 16;
  }
  int ach = (num_sms() * ctas_per_sm + B - 1) / B;
  if (ach > hw) ach = hw;
  if (ach < 1) ach = 1;
  const int appb = (hw + ach - 1) / ach;
  ach = (hw + appb - 1) / appb;
  if (out_dtype == DFB_DTYPE_F32)
    groupnorm_apply_kernel<float><<<dim3(ach, B), GN_THREADS, 0, stream>>>(src0, c0, ld0, src1, c1, ld1, hw, groups, appb, stats,
                                                                           gamma, beta, silu, (float*)out, ld_out, (float*)raw_out, ld_raw);
  else
    groupnorm_apply_kernel<__nv_bfloat16><<<dim3(ach, B), GN_THREADS, 0, stream>>>(src0, c0, ld0, src1, c1, ld1, hw, groups, appb, stats,
                                                                                   gamma, beta, silu, (__nv_bfloat16*)out, ld_out,
                                                                                   (__nv_bfloat16*)raw_out, ld_raw);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

extern "C" {

size_t dfb_groupnorm_ws_floats(int B, int groups) {
  return (size_t)B * groups * 2 * (1 + GN_MAX_CHUNKS);
}

int dfb_groupnorm(const float* src0, int c0, int ld0, const float* src1, int c1, int ld1, int B, int hw, int groups,
                  float eps, const float* gamma, const float* beta, int silu, float* stats_ws, void* out, int out_dtype,
                  int ld_out, void* raw_out, int ld_raw, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DFB_REQUIRE(src0 && gamma && beta && stats_ws && out, "dfb_groupnorm: null buffer");
  DFB_REQUIRE(c1 == 0 || src1 != nullptr, "dfb_groupnorm: second source missing");
  const int Cch = c0 + c1;
  DFB_REQUIRE(B > 0 && hw > 0 && groups > 0 && groups <= 64 && Cch % groups == 0, "dfb_groupnorm: bad sizes");
  DFB_REQUIRE((Cch / groups) % 2 == 0 && c0 % 4 == 0 && c1 % 4 == 0, "dfb_groupnorm: channels/group must be even, sources multiple of 4");
  DFB_REQUIRE(Cch / 4 <= GN_THREADS * GN_MAX_VEC_PER_THREAD, "dfb_groupnorm: too many channels");
  DFB_REQUIRE(ld0 % 4 == 0 && ld1 % 4 == 0 && ld_out % 4 == 0 && ld_raw % 4 == 0, "dfb_groupnorm: pitches must be multiples of 4");
  // Pixels per block are a function of the IMAGE size only (32 .. hw / GN_MAX_CHUNKS): the order in which an image's
  // statistics are summed — hence every bit of its output — must not depend on how many images share the launch, or the
  // row chunks of a batch would not reproduce the unsplit batch.  (It used to be sized to fill the machine from B: measured
  // on B200, a 48-row batch of 4x4 images then differed from the same rows run 16 at a time in the last bits of the
  // statistics, which flips bf16 roundings downstream.)
  int ppb = (hw + GN_MAX_CHUNKS - 1) / GN_MAX_CHUNKS;
  if (ppb < 32) ppb = 32;
  const int chunks = (hw + ppb - 1) / ppb;
  float* stats = stats_ws;                                   // [B, groups, 2] (mean, rstd)
  float* partial = stats_ws + (size_t)B * groups * 2;        // [B, chunks, groups, 2]
  groupnorm_stats_kernel<<<dim3(chunks, B), GN_THREADS, 0, stream>>>(src0, c0, ld0, src1, c1, ld1, hw, groups, ppb, partial);
  DFB_CHECK_CUDA(cudaGetLastError());
  {
    const int warps = B * groups;
    const int threads = 128;
    const int blocks = (warps * 32 + threads - 1) / threads;
    const float inv_n = 1.f / ((float)(Cch / groups) * (float)hw);
    groupnorm_finalize_kernel<<<blocks, threads, 0, stream>>>(partial, chunks, groups, B, inv_n, eps, stats);
    DFB_CHECK_CUDA(cudaGetLastError());
  }
  return launch_gn_apply(src0, c0, ld0, src1, c1, ld1, B, hw, groups, stats, gamma, beta, silu, out, out_dtype, ld_out, raw_out,
                         ld_raw, stream);
}

int dfb_groupnorm_fused(const float* src0, int c0, int ld0, const float* partial0, const float* src1, int c1, int ld1,
                        const float* partial1, int B, int hw, int groups, float eps, const float* gamma,
                        const float* beta, int silu, float* stats_ws, void* out, int out_dtype, int ld_out, void* raw_out,
                        int ld_raw, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DFB_REQUIRE(src0 && partial0 && gamma && beta && stats_ws && out, "dfb_groupnorm_fused: null buffer");
  DFB_REQUIRE(c1 == 0 || (src1 != nullptr && partial1 != nullptr), "dfb_groupnorm_fused: second source missing");
  const int Cch = c0 + c1;
  DFB_REQUIRE(B > 0 && hw > 0 && hw % 32 == 0 && groups > 0 && groups <= 64 && Cch % groups == 0, "dfb_groupnorm_fused: bad sizes");
  DFB_REQUIRE((Cch / groups) % 2 == 0 && c0 % 4 == 0 && c1 % 4 == 0, "dfb_groupnorm_fused: channels/group must be even, sources multiple of 4");
  DFB_REQUIRE(Cch / 4 <= GN_THREADS * GN_MAX_VEC_PER_THREAD, "dfb_groupnorm_fused: too many channels");
  DFB_REQUIRE(ld0 % 4 == 0 && ld1 % 4 == 0 && ld_out % 4 == 0 && ld_raw % 4 == 0, "dfb_groupnorm_fused: pitches must be multiples of 4");
  {
    const int warps = B * groups;
    const int threads = 128;
    const int blocks = (warps * 32 + threads - 1) / threads;
    groupnorm_finalize_partials_kernel<<<blocks, threads, 0, stream>>>(partial0, c0, partial1, c1, hw, groups, B, eps, stats_ws);
    DFB_CHECK_CUDA(cudaGetLastError());
  }
  return launch_gn_apply(src0, c0, ld0, src1, c1, ld1, B, hw, groups, stats_ws, gamma, beta, silu, out, out_dtype, ld_out, raw_out,
                         ld_raw, stream);
}

int dfb_layernorm(const float* x, int ld_x, const float* gamma, const float* beta, float eps, void* out, int out_dtype,
                  int ld_out, int rows, int Cch, void* stream) {
  DFB_REQUIRE(x && gamma && beta && out, "dfb_layernorm: null buffer");
  DFB_REQUIRE(rows > 0 && Cch > 0 && Cch % 4 == 0 && Cch <= 2048, "dfb_layernorm: C must be a multiple of 4, <= 2048");
  DFB_REQUIRE(ld_x % 4 == 0 && ld_out % 4 == 0, "dfb_layernorm: pitches must be multiples of 4");
  long long blocks = ((long long)rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
#define DFB_LN_LAUNCH(T)                                                                                                   \
  do {                                                                                                                      \
    T* o = (T*)out;                                                                                                         \
    if (Cch <= 384) layernorm_kernel<3, T><<<(int)blocks, 256, 0, st>>>(x, ld_x, gamma, beta, eps, o, ld_out, rows, Cch);     \
    else if (Cch <= 640) layernorm_kernel<5, T><<<(int)blocks, 256, 0, st>>>(x, ld_x, gamma, beta, eps, o, ld_out, rows, Cch); \
    else if (Cch <= 1280) layernorm_kernel<10, T><<<(int)blocks, 256, 0, st>>>(x, ld_x, gamma, beta, eps, o, ld_out, rows, Cch); \
    else layernorm_kernel<16, T><<<(int)blocks, 256, 0, st>>>(x, ld_x, gamma, beta, eps, o, ld_out, rows, Cch);               \
  } while (0)
  if (out_dtype == DFB_DTYPE_F32) DFB_LN_LAUNCH(float);
  else DFB_LN_LAUNCH(__nv_bfloat16);
#undef DFB_LN_LAUNCH
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_softmax_rows(const float* in, int in_ld, float scale, void* out, int out_dtype, int out_ld, int rows, int cols,
                     void* stream) {
  DFB_REQUIRE(in && out && rows > 0 && cols > 0, "dfb_softmax_rows: bad args");
  DFB_REQUIRE(cols % 4 == 0 && cols <= 256 * 4 * SM_MAX_VEC && in_ld % 4 == 0 && out_ld % 4 == 0,
              "dfb_softmax_rows: cols must be a multiple of 4, at most 8192; pitches multiples of 4");
  int blocks = rows < num_sms() * 8 ? rows : num_sms() * 8;
  if (out_dtype == DFB_DTYPE_F32)
    softmax_rows_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, in_ld, scale, (float*)out, out_ld, rows, cols);
  else
    softmax_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, in_ld, scale, (__nv_bfloat16*)out, out_ld, rows, cols);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

}  // extern "C"
