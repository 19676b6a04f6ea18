// HBM-streaming kernels of the denoising step: CFG combine + scheduler update, mutual-condition
// gather/blend, layout conversions, nearest-2x upsample, space-to-depth, timestep embedding.
// All are coalesced, 128-bit vectorised grid-stride kernels; none has data reuse, so there is no
// shared-memory staging (roofline = HBM bandwidth).
#include "dfb_host.h"
#include "../../include/dfb200.h"

namespace dfb {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

static inline int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// CFG combine + scheduler update (difashion.py:525-569).
//   e0 = sum_b w[b] * eps_b          (eps: NHWC fp32 [nb*N, HW, 4], branch-major like the reference's cat)
//   x_out = cx * x_src + ck[0]*e0 + ck[1]*h1 + ck[2]*h2 + ck[3]*h3 + cn * noise      (NCHW fp32)
//   eps_out (optional) = e0          (NCHW fp32, PLMS history)
// One thread = 4 consecutive pixels of one item: float4 loads from every stream.
// ---------------------------------------------------------------------------------------------
struct CfgStepParams {
  const float* eps;
  int nb;
  float w[4];
  const float* x_src;
  float cx;
  float ck[4];
  const float* hist[3];
  const float* noise;
  float cn;
  float* x_out;
  float* eps_out;
  int n_items, hw;
  int eps_nchw;
};

__global__ void __launch_bounds__(256) cfg_step_kernel(const CfgStepParams p) {
  const int quads_per_item = p.hw >> 2;
  const long long total = (long long)p.n_items * quads_per_item;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / quads_per_item);
    const int px = (int)(i - (long long)n * quads_per_item) << 2;
    float e[4][4];  // [pixel][channel]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) e[a][c] = 0.f;
    for (int b = 0; b < p.nb; ++b) {
      const float wb = p.w[b];
      if (p.eps_nchw) {
        const float* src = p.eps + (((size_t)b * p.n_items + n) * 4) * p.hw + px;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v = ldg4(src + (size_t)c * p.hw);
          e[0][c] += wb * v.x; e[1][c] += wb * v.y; e[2][c] += wb * v.z; e[3][c] += wb * v.w;
        }
      } else {
        const float* src = p.eps + (((size_t)b * p.n_items + n) * p.hw + px) * 4;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 v = ldg4(src + a * 4);
          e[a][0] += wb * v.x; e[a][1] += wb * v.y; e[a][2] += wb * v.z; e[a][3] += wb * v.w;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const size_t o = ((size_t)n * 4 + c) * p.hw + px;
      const float4 xs = ldg4(p.x_src + o);
      float r[4] = {p.cx * xs.x, p.cx * xs.y, p.cx * xs.z, p.cx * xs.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) r[a] += p.ck[0] * e[a][c];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (p.hist[k]) {
          const float4 h = ldg4(p.hist[k] + o);
          r[0] += p.ck[k + 1] * h.x; r[1] += p.ck[k + 1] * h.y; r[2] += p.ck[k + 1] * h.z; r[3] += p.ck[k + 1] * h.w;
        }
      }
      if (p.noise) {
        const float4 z = ldg4(p.noise + o);
        r[0] += p.cn * z.x; r[1] += p.cn * z.y; r[2] += p.cn * z.z; r[3] += p.cn * z.w;
      }
      *reinterpret_cast<float4*>(p.x_out + o) = make_float4(r[0], r[1], r[2], r[3]);
      if (p.eps_out) *reinterpret_cast<float4*>(p.eps_out + o) = make_float4(e[0][c], e[1][c], e[2][c], e[3][c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Mutual-condition gather-sum (difashion.py:475-488): for generated item n, sum the latents of
// the other slots of its outfit: idx >= 0 -> given item row of all_latents, idx < 0 -> generated
// sibling row (-idx-1) of prev_latents, sentinel INT_MIN -> skip.  Output bf16 [N, D] (A operand
// of the MutualEncoder's first Linear).
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) mutual_gather_sum_kernel(const float* __restrict__ all_latents,
                                                                const float* __restrict__ prev_latents,
                                                                const int* __restrict__ idx, int n_items,
                                                                int n_src, int d, TOut* __restrict__ out) {
  const int dq = d >> 2;
  const long long total = (long long)n_items * dq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / dq);
    const int q = (int)(i - (long long)n * dq) << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n_src; ++s) {
      const int j = __ldg(idx + n * n_src + s);
      if (j == INT_MIN) continue;
      const float* src = j >= 0 ? all_latents + (size_t)j * d : prev_latents + (size_t)(-j - 1) * d;
      const float4 v = ldg4(src + q);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    store4(out + (size_t)n * d + q, acc.x, acc.y, acc.z, acc.w);
  }
}

// ---------------------------------------------------------------------------------------------
// Mutual blend + history concat + CFG branch expansion (difashion.py:458-459, 494-515, 388-390):
//   out[b, n, px, 0:4] = (1-eta) * x[n, :, px] + eta * (use_m[b] ? m[n, :, px] : null[:, px])
//   out[b, n, px, 4:8] =            use_h[b] ? hist[n, :, px] : null[:, px]
// inputs NCHW fp32, output NHWC bf16 [nb*N, HW, 8] — the UNet's conv_in A operand.  The x4 latent
// expansion is never materialised in fp32.
// ---------------------------------------------------------------------------------------------
struct BlendParams {
  const float* x;
  const float* m;
  const float* hist;
  const float* null_latent;
  float eta;
  int nb;
  int use_m[4];
  int use_h[4];
  int n_items, hw;
  void* out;
};

template <typename TOut>
__global__ void __launch_bounds__(256) mutual_blend_kernel(const BlendParams p) {
  const long long total = (long long)p.n_items * p.hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / p.hw);
    const int px = (int)(i - (long long)n * p.hw);
    float x[4], m[4], h[4], z[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const size_t o = ((size_t)n * 4 + c) * p.hw + px;
      x[c] = __ldg(p.x + o);
      z[c] = __ldg(p.null_latent + (size_t)c * p.hw + px);
      m[c] = p.m ? __ldg(p.m + o) : z[c];
      h[c] = p.hist ? __ldg(p.hist + o) : z[c];
    }
    // Both candidates of the blended latent are computed ONCE per pixel, with explicitly rounded operations, and the
    // branch loop only selects: branches with equal flags then hold equal bits by construction.  (Written inside the loop
    // as `(1 - eta) * x + eta * m`, nvcc unrolled the loop by two and contracted the expression into an FMA differently
    // in the unrolled body and in the remainder iteration — FFMA(m, eta, x * (1 - eta)) against FFMA(x, 1 - eta, m * eta) —
    // so with an ODD branch count the last branch differed from an identical earlier one in the last fp32 bit, which now
    // and then flips a bf16 rounding: the shared-CFG-prefix mismatch of round 1's driver run.)  The reference computes
    // two products and a sum (difashion.py:513), which is what the intrinsics spell out.
    const float w_x = 1.f - p.eta;
    float vm[4], vz[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float xs = __fmul_rn(w_x, x[c]);
      vm[c] = __fadd_rn(xs, __fmul_rn(p.eta, m[c]));
      vz[c] = __fadd_rn(xs, __fmul_rn(p.eta, z[c]));
    }
    for (int b = 0; b < p.nb; ++b) {
      float v[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[c] = p.use_m[b] ? vm[c] : vz[c];
        v[4 + c] = p.use_h[b] ? h[c] : z[c];
      }
      const size_t o = (((size_t)b * p.n_items + n) * p.hw + px) * 8;
      TOut* op = reinterpret_cast<TOut*>(p.out) + o;
      store4(op, v[0], v[1], v[2], v[3]);
      store4(op + 4, v[4], v[5], v[6], v[7]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversions at the diffusers-API boundary
// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TIn* __restrict__ in, TOut* __restrict__ out,
                                                           int B, int C, int HW) {
  // small C (8): one thread per pixel reads C strided-but-coalesced-across-threads values
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int px = (int)(i - (long long)b * HW);
    for (int c = 0; c < C; ++c)
      store1(out + ((size_t)b * HW + px) * C + c, (float)in[((size_t)b * C + c) * HW + px]);
  }
}

template <typename TOut>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, TOut* __restrict__ out, int B, int C, int HW) {
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int px = (int)(i - (long long)b * HW);
    for (int c = 0; c < C; ++c)
      out[((size_t)b * C + c) * HW + px] = (TOut)in[((size_t)b * HW + px) * C + c];
  }
}

// fp32/bf16 [rows, cols] -> bf16 [rows_out >= rows ... ] with per-batch row padding:
// in [B, S, D] -> out [B, S_pad, D] (rows >= S zero).  Used for encoder_hidden_states (S=77 -> 80).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) pad_cast_rows_kernel(const TIn* __restrict__ in, TOut* __restrict__ out,
                                                            int B, int S, int S_pad, int D) {
  const long long total = (long long)B * S_pad * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int dcol = (int)(i % D);
    const long long r = i / D;
    const int s = (int)(r % S_pad);
    const int b = (int)(r / S_pad);
    store1(out + i, s < S ? (float)in[((size_t)b * S + s) * D + dcol] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 -> MMA operand (bf16, or fp32 on the verification path), 4 elements per thread: the low-resolution input of the
// fused upsample-phase convolutions (dfb_gemm up2x)
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ in, TOut* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    store4(out + 4 * i, v.x, v.y, v.z, v.w);
  }
}

// ---------------------------------------------------------------------------------------------
// nearest-2x upsample (Upsample2D, App. A.2 item 7): fp32 NHWC [B,H,W,C] -> bf16 NHWC [B,2H,2W,C]
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ in, TOut* __restrict__ out,
                                                         int B, int H, int W, int C) {
  const int cq = C >> 2;
  const long long total = (long long)B * H * W * cq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cq) << 2;
    long long r = i / cq;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    const float4 v = ldg4(in + (((size_t)b * H + h) * W + w) * C + c);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
        store4(out + (((size_t)b * 2 * H + 2 * h + dy) * 2 * W + 2 * w + dx) * C + c, v.x, v.y, v.z, v.w);
  }
}

// ---------------------------------------------------------------------------------------------
// space-to-depth for the stride-2 Downsample2D conv: fp32 NHWC [B,H,W,C] -> bf16 [B,H/2,W/2,4C],
// plane p = (h&1)*2 + (w&1) stored at channel offset p*C.  The stride-2 3x3 conv then becomes a
// stride-1 9-tap gather over these planes (tap table built on the host).
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) space_to_depth_kernel(const float* __restrict__ in, TOut* __restrict__ out,
                                                             int B, int H, int W, int C) {
  const int cq = C >> 2;
  const long long total = (long long)B * H * W * cq;
  const int H2 = H >> 1, W2 = W >> 1;
  // four independent 16-byte loads in flight per thread (one per grid stride): with a single one the kernel sat at 3.9 TB/s of
  // tensor bytes (ncu launch list, 0.28 ms for the 64x64x320 tensor), exposed DRAM latency per iteration
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
    float4 v[U];
    size_t dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      dst[u] = (size_t)-1;
      if (i < total) {
        const int c = (int)(i % cq) << 2;
        long long r = i / cq;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int b = (int)(r / H);
        v[u] = ldg4(in + (((size_t)b * H + h) * W + w) * C + c);
        const int plane = (h & 1) * 2 + (w & 1);
        dst[u] = ((((size_t)b * H2 + (h >> 1)) * W2 + (w >> 1)) * 4 + plane) * C + c;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] != (size_t)-1) store4(out + dst[u], v[u].x, v[u].y, v[u].z, v[u].w);
  }
}

// ---------------------------------------------------------------------------------------------
// sinusoidal timestep projection (Timesteps, App. A.2 item 2): fp32 math, [cos | sin], bf16 out
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void timestep_embedding_kernel(const float* __restrict__ t, TOut* __restrict__ out, int B, int dim,
                                          int flip_sin_to_cos, float freq_shift) {
  const int half = dim >> 1;
  const int total = B * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / half, k = i - b * half;
    const float freq = expf(-logf(10000.f) * (float)k / ((float)half - freq_shift));
    const float arg = t[b] * freq;
    const float s = sinf(arg), c = cosf(arg);
    TOut* o = out + (size_t)b * dim;
    if (flip_sin_to_cos) { store1(o + k, c); store1(o + half + k, s); }
    else { store1(o + k, s); store1(o + half + k, c); }
  }
}

// ---------------------------------------------------------------------------------------------
// CLIPTextEmbeddings: token-table gather + position embedding (float4 per thread, fp32 residual stream out)
// ---------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                                    float* __restrict__ out, int rows, int S, int D4, int vocab) {
  const long long total = (long long)rows * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / D4), c = (int)(i - (long long)r * D4);
    int id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);       // ids are validated on the host; never read out of the table
    const float4 a = __ldg(reinterpret_cast<const float4*>(tok) + (size_t)id * D4 + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos) + (size_t)(r % S) * D4 + c);
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// ---------------------------------------------------------------------------------------------
// VaeImageProcessor.postprocess (denormalize + uint8): fp32 NHWC [pixels, in_c] -> uint8 [pixels, 3]
// ---------------------------------------------------------------------------------------------
__global__ void image_to_uint8_kernel(const float* __restrict__ in, int in_c, uint8_t* __restrict__ out, long long pixels) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const float* src = in + i * in_c;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = src[c] * 0.5f + 0.5f;                        // (image / 2 + 0.5).clamp(0, 1)
      v = fminf(fmaxf(v, 0.f), 1.f);
      out[i * 3 + c] = (uint8_t)rintf(v * 255.f);            // (x * 255).round().astype("uint8"): half to even
    }
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_cfg_step(const float* eps, int eps_nchw, int nb, const float* w, const float* x_src, float cx, const float* ck,
                 const float* hist1, const float* hist2, const float* hist3, const float* noise, float cn,
                 float* x_out, float* eps_out, int n_items, int hw, void* stream) {
  DFB_REQUIRE(eps && w && x_src && ck && x_out, "dfb_cfg_step: null buffer");
  DFB_REQUIRE(nb >= 1 && nb <= 4, "dfb_cfg_step: 1..4 guidance branches");
  DFB_REQUIRE(n_items > 0 && hw > 0 && hw % 4 == 0, "dfb_cfg_step: hw must be a positive multiple of 4");
  DFB_REQUIRE(((uintptr_t)eps | (uintptr_t)x_src | (uintptr_t)x_out) % 16 == 0, "dfb_cfg_step: buffers must be 16B aligned");
  CfgStepParams p;
  memset(&p, 0, sizeof(p));
  p.eps = eps; p.nb = nb;
  for (int b = 0; b < nb; ++b) p.w[b] = w[b];
  p.x_src = x_src; p.cx = cx;
  for (int k = 0; k < 4; ++k) p.ck[k] = ck[k];
  p.hist[0] = hist1; p.hist[1] = hist2; p.hist[2] = hist3;
  p.noise = noise; p.cn = cn; p.x_out = x_out; p.eps_out = eps_out;
  p.n_items = n_items; p.hw = hw; p.eps_nchw = eps_nchw ? 1 : 0;
  const long long work = (long long)n_items * (hw / 4);
  cfg_step_kernel<<<grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(p);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_mutual_gather_sum(const float* all_latents, const float* prev_latents, const int32_t* idx, int n_items,
                          int n_src, int d, void* out, int out_dtype, void* stream) {
  DFB_REQUIRE(prev_latents && idx && out, "dfb_mutual_gather_sum: null buffer");
  DFB_REQUIRE(n_items > 0 && n_src >= 0 && d > 0 && d % 4 == 0, "dfb_mutual_gather_sum: bad sizes");
  const long long work = (long long)n_items * (d / 4);
  const int g = grid_for(work, 256);
  if (out_dtype == DFB_DTYPE_F32)
    mutual_gather_sum_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(all_latents, prev_latents, idx, n_items, n_src, d, (float*)out);
  else
    mutual_gather_sum_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(all_latents, prev_latents, idx, n_items, n_src, d, (__nv_bfloat16*)out);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_mutual_blend(const float* x, const float* m, const float* hist, const float* null_latent, float eta, int nb,
                     const int32_t* use_m, const int32_t* use_h, int n_items, int hw, void* out, int out_dtype, void* stream) {
  DFB_REQUIRE(x && null_latent && out && use_m && use_h, "dfb_mutual_blend: null buffer");
  DFB_REQUIRE(nb >= 1 && nb <= 4 && n_items > 0 && hw > 0, "dfb_mutual_blend: bad sizes");
  DFB_REQUIRE(((uintptr_t)out) % 16 == 0, "dfb_mutual_blend: output must be 16B aligned");
  BlendParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.m = m; p.hist = hist; p.null_latent = null_latent; p.eta = eta; p.nb = nb;
  for (int b = 0; b < nb; ++b) { p.use_m[b] = use_m[b]; p.use_h[b] = use_h[b]; }
  p.n_items = n_items; p.hw = hw; p.out = out;
  const int g = grid_for((long long)n_items * hw, 256);
  if (out_dtype == DFB_DTYPE_F32) mutual_blend_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(p);
  else mutual_blend_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(p);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int B, int Cch, int HW, void* stream) {
  DFB_REQUIRE(in && out && B > 0 && Cch > 0 && HW > 0, "dfb_nchw_to_nhwc: bad args");
  const int g = grid_for((long long)B * HW, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == DFB_DTYPE_F32 && out_dtype == DFB_DTYPE_F32)
    nchw_to_nhwc_kernel<float, float><<<g, 256, 0, st>>>((const float*)in, (float*)out, B, Cch, HW);
  else if (in_dtype == DFB_DTYPE_F32)
    nchw_to_nhwc_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, B, Cch, HW);
  else if (out_dtype == DFB_DTYPE_F32)
    nchw_to_nhwc_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, B, Cch, HW);
  else
    nchw_to_nhwc_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, Cch, HW);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_nhwc_to_nchw(const float* in, void* out, int out_dtype, int B, int Cch, int HW, void* stream) {
  DFB_REQUIRE(in && out && B > 0 && Cch > 0 && HW > 0, "dfb_nhwc_to_nchw: bad args");
  const int g = grid_for((long long)B * HW, 256);
  if (out_dtype == DFB_DTYPE_F32)
    nhwc_to_nchw_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(in, (float*)out, B, Cch, HW);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, B, Cch, HW);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_pad_cast_rows(const void* in, int in_dtype, void* out, int out_dtype, int B, int S, int S_pad, int D, void* stream) {
  DFB_REQUIRE(in && out && B > 0 && S > 0 && S_pad >= S && D > 0, "dfb_pad_cast_rows: bad args");
  const int g = grid_for((long long)B * S_pad * D, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == DFB_DTYPE_F32 && out_dtype == DFB_DTYPE_F32)
    pad_cast_rows_kernel<float, float><<<g, 256, 0, st>>>((const float*)in, (float*)out, B, S, S_pad, D);
  else if (in_dtype == DFB_DTYPE_F32)
    pad_cast_rows_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, B, S, S_pad, D);
  else if (out_dtype == DFB_DTYPE_F32)
    pad_cast_rows_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, B, S, S_pad, D);
  else
    pad_cast_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, S, S_pad, D);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_cast_f32(const float* in, void* out, int out_dtype, long long n, void* stream) {
  DFB_REQUIRE(in && out && n > 0 && n % 4 == 0, "dfb_cast_f32: bad args (n must be a positive multiple of 4)");
  DFB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "dfb_cast_f32: buffers must be 16-byte aligned");
  const int g = grid_for(n / 4, 256);
  if (out_dtype == DFB_DTYPE_F32) cast_rows_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(in, (float*)out, n / 4);
  else cast_rows_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n / 4);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_upsample2x(const float* in, void* out, int out_dtype, int B, int H, int W, int Cch, void* stream) {
  DFB_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && Cch > 0 && Cch % 4 == 0, "dfb_upsample2x: bad args");
  const int g = grid_for((long long)B * H * W * (Cch / 4), 256);
  if (out_dtype == DFB_DTYPE_F32) upsample2x_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(in, (float*)out, B, H, W, Cch);
  else upsample2x_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, B, H, W, Cch);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_space_to_depth(const float* in, void* out, int out_dtype, int B, int H, int W, int Cch, void* stream) {
  DFB_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && Cch % 4 == 0, "dfb_space_to_depth: bad args");
  const int g = grid_for((long long)B * H * W * (Cch / 4), 256);
  if (out_dtype == DFB_DTYPE_F32) space_to_depth_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(in, (float*)out, B, H, W, Cch);
  else space_to_depth_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, B, H, W, Cch);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_timestep_embedding(const float* t, void* out, int out_dtype, int B, int dim, int flip_sin_to_cos, float freq_shift,
                           void* stream) {
  DFB_REQUIRE(t && out && B > 0 && dim > 0 && dim % 2 == 0, "dfb_timestep_embedding: bad args");
  const int total = B * dim / 2;
  if (out_dtype == DFB_DTYPE_F32)
    timestep_embedding_kernel<float><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, (float*)out, B, dim, flip_sin_to_cos, freq_shift);
  else
    timestep_embedding_kernel<__nv_bfloat16><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, (__nv_bfloat16*)out, B, dim, flip_sin_to_cos, freq_shift);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_embed_tokens(const int32_t* ids, const float* token_table, const float* position_table, float* out, int B, int S,
                     int D, int vocab, void* stream) {
  DFB_REQUIRE(ids && token_table && position_table && out, "dfb_embed_tokens: null buffer");
  DFB_REQUIRE(B > 0 && S > 0 && D > 0 && D % 4 == 0 && vocab > 0, "dfb_embed_tokens: bad shape (D must be a multiple of 4)");
  DFB_REQUIRE((((uintptr_t)token_table | (uintptr_t)position_table | (uintptr_t)out) & 15) == 0, "dfb_embed_tokens: buffers must be 16B aligned");
  embed_tokens_kernel<<<grid_for((long long)B * S * (D / 4), 256), 256, 0, (cudaStream_t)stream>>>(ids, token_table, position_table, out,
                                                                                                 B * S, S, D / 4, vocab);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_image_to_uint8(const float* in, int in_c, uint8_t* out, long long pixels, void* stream) {
  DFB_REQUIRE(in && out && in_c >= 3 && pixels > 0, "dfb_image_to_uint8: bad args");
  image_to_uint8_kernel<<<grid_for(pixels, 256), 256, 0, (cudaStream_t)stream>>>(in, in_c, out, pixels);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

}  // extern "C"
