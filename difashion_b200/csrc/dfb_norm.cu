// GroupNorm (+SiLU) and LayerNorm producing the bf16 A operands of the following GEMM / conv.
// Statistics are fp32 (diffusers keeps norms in fp32 under autocast).  Inputs are
// the fp32 residual stream (NHWC); the skip-concat of the up blocks is read from its two sources
// directly (never materialised in fp32) — groups may straddle the concat seam (C=1920, 960).
#include "dfb_host.h"
#include "../../include/dfb200.h"

namespace dfb {

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[b][g] = (sum, sumsq) accumulated with fp32 atomics.
// grid = (pixel chunks, B); each thread owns fixed float4 channel-vectors (coalesced across the
// warp) and walks the chunk's pixels.  Channels-per-group is even, so a float4 spans <= 2 groups.
// ---------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_VEC_PER_THREAD = 4;   // supports C <= 4096

__global__ void __launch_bounds__(GN_THREADS)
groupnorm_stats_kernel(const float* __restrict__ src0, int c0, int ld0, const float* __restrict__ src1, int c1, int ld1,
                       int hw, int groups, int pix_per_block, float* __restrict__ stats) {
  __shared__ float s_sum[64], s_sq[64];
  const int b = blockIdx.y;
  const int C = c0 + c1;
  const int cg = C / groups;
  const int nvec = C >> 2;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  if (threadIdx.x < groups) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
  __syncthreads();
  // thread -> (channel vector, pixel lane): with few channels several pixel lanes share a block
  const int vlanes = nvec < GN_THREADS ? nvec : GN_THREADS;
  const int plane_cnt = GN_THREADS / vlanes;            // >= 1
  const int tv = threadIdx.x % vlanes;
  const int tp = threadIdx.x / vlanes;
  for (int v = tv; v < nvec && tp < plane_cnt; v += GN_THREADS) {
    const int c = v << 2;
    const float* base;
    int ld, cc;
    if (c < c0) { base = src0; ld = ld0; cc = c; } else { base = src1; ld = ld1; cc = c - c0; }
    base += (size_t)b * hw * ld + cc;
    float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
    for (int px = p_begin + tp; px < p_end; px += plane_cnt) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(base + (size_t)px * ld));
      sa += x.x + x.y; qa += x.x * x.x + x.y * x.y;
      sb += x.z + x.w; qb += x.z * x.z + x.w * x.w;
    }
    const int ga = c / cg, gb = (c + 2) / cg;
    if (ga == gb) {
      atomicAdd(&s_sum[ga], sa + sb); atomicAdd(&s_sq[ga], qa + qb);
    } else {
      atomicAdd(&s_sum[ga], sa); atomicAdd(&s_sq[ga], qa);
      atomicAdd(&s_sum[gb], sb); atomicAdd(&s_sq[gb], qb);
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    atomicAdd(&stats[((size_t)b * groups + threadIdx.x) * 2 + 0], s_sum[threadIdx.x]);
    atomicAdd(&stats[((size_t)b * groups + threadIdx.x) * 2 + 1], s_sq[threadIdx.x]);
  }
}

// y = (x - mean) * rstd * gamma + beta (optionally SiLU) -> bf16 [B, HW, C] (ld_out), optional raw
// bf16 copy of x (A operand of the 1x1 shortcut folded into conv2).
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_apply_kernel(const float* __restrict__ src0, int c0, int ld0, const float* __restrict__ src1, int c1, int ld1,
                       int hw, int groups, float eps, const float* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, int silu, __nv_bfloat16* __restrict__ out, int ld_out,
                       __nv_bfloat16* __restrict__ raw_out, int ld_raw, int B) {
  const int C = c0 + c1;
  const int cg = C / groups;
  const int nvec = C >> 2;
  const float inv_n = 1.f / ((float)cg * (float)hw);
  const long long total = (long long)B * hw * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const long long r = i / nvec;          // b * hw + px
    const int b = (int)(r / hw);
    const int c = v << 2;
    const float* base;
    int ld, cc;
    if (c < c0) { base = src0; ld = ld0; cc = c; } else { base = src1; ld = ld1; cc = c - c0; }
    const float4 x = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * ld + cc));
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
    const int ga = c / cg, gb = (c + 2) / cg;
    const float* st = stats + (size_t)b * groups * 2;
    const float ma = st[ga * 2] * inv_n;
    const float ra = rsqrtf(fmaxf(st[ga * 2 + 1] * inv_n - ma * ma, 0.f) + eps);
    float mb = ma, rb = ra;
    if (gb != ga) {
      mb = st[gb * 2] * inv_n;
      rb = rsqrtf(fmaxf(st[gb * 2 + 1] * inv_n - mb * mb, 0.f) + eps);
    }
    float y0 = (x.x - ma) * ra * g.x + bt.x;
    float y1 = (x.y - ma) * ra * g.y + bt.y;
    float y2 = (x.z - mb) * rb * g.z + bt.z;
    float y3 = (x.w - mb) * rb * g.w + bt.w;
    if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); y2 = silu_f(y2); y3 = silu_f(y3); }
    *reinterpret_cast<uint2*>(out + (size_t)r * ld_out + c) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    if (raw_out)
      *reinterpret_cast<uint2*>(raw_out + (size_t)r * ld_raw + c) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim: one warp per row, row held in registers (C <= 2048), two-pass
// (mean, then centred variance) in fp32; bf16 output.
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAX_VEC = 16;   // float4 per lane -> C <= 2048

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, __nv_bfloat16* __restrict__ out, int ld_out, int rows, int C) {
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
    __nv_bfloat16* orow = out + (size_t)row * ld_out;
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
        *reinterpret_cast<uint2*>(orow + (j << 2)) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
      }
    }
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_groupnorm(const float* src0, int c0, int ld0, const float* src1, int c1, int ld1, int B, int hw, int groups,
                  float eps, const float* gamma, const float* beta, int silu, float* stats_ws, void* out_bf16, int ld_out,
                  void* raw_out_bf16, int ld_raw, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DFB_REQUIRE(src0 && gamma && beta && stats_ws && out_bf16, "dfb_groupnorm: null buffer");
  DFB_REQUIRE(c1 == 0 || src1 != nullptr, "dfb_groupnorm: second source missing");
  const int Cch = c0 + c1;
  DFB_REQUIRE(B > 0 && hw > 0 && groups > 0 && groups <= 64 && Cch % groups == 0, "dfb_groupnorm: bad sizes");
  DFB_REQUIRE((Cch / groups) % 2 == 0 && c0 % 4 == 0 && c1 % 4 == 0, "dfb_groupnorm: channels/group must be even, sources multiple of 4");
  DFB_REQUIRE(Cch / 4 <= GN_THREADS * GN_MAX_VEC_PER_THREAD, "dfb_groupnorm: too many channels");
  DFB_REQUIRE(ld0 % 4 == 0 && ld1 % 4 == 0 && ld_out % 4 == 0 && ld_raw % 4 == 0, "dfb_groupnorm: pitches must be multiples of 4");
  DFB_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(float) * 2 * groups * B, stream));
  // enough blocks to fill the machine: ~4 waves of CTAs over (B x pixel chunks)
  int chunks = (num_sms() * 4 + B - 1) / B;
  if (chunks > hw) chunks = hw;
  if (chunks < 1) chunks = 1;
  const int ppb = (hw + chunks - 1) / chunks;
  chunks = (hw + ppb - 1) / ppb;
  groupnorm_stats_kernel<<<dim3(chunks, B), GN_THREADS, 0, stream>>>(src0, c0, ld0, src1, c1, ld1, hw, groups, ppb, stats_ws);
  DFB_CHECK_CUDA(cudaGetLastError());
  const long long total = (long long)B * hw * (Cch / 4);
  long long blocks = (total + GN_THREADS - 1) / GN_THREADS;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  groupnorm_apply_kernel<<<(int)blocks, GN_THREADS, 0, stream>>>(src0, c0, ld0, src1, c1, ld1, hw, groups, eps, stats_ws, gamma,
                                                                beta, silu, (__nv_bfloat16*)out_bf16, ld_out,
                                                                (__nv_bfloat16*)raw_out_bf16, ld_raw, B);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_layernorm(const float* x, int ld_x, const float* gamma, const float* beta, float eps, void* out_bf16, int ld_out,
                  int rows, int Cch, void* stream) {
  DFB_REQUIRE(x && gamma && beta && out_bf16, "dfb_layernorm: null buffer");
  DFB_REQUIRE(rows > 0 && Cch > 0 && Cch % 4 == 0 && Cch <= LN_MAX_VEC * 128, "dfb_layernorm: C must be a multiple of 4, <= 2048");
  DFB_REQUIRE(ld_x % 4 == 0 && ld_out % 4 == 0, "dfb_layernorm: pitches must be multiples of 4");
  long long blocks = ((long long)rows + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  layernorm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ld_x, gamma, beta, eps, (__nv_bfloat16*)out_bf16, ld_out, rows, Cch);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

}  // extern "C"
