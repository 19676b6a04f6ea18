// fp32 verification path: the same operators as dfb_gemm / dfb_attention with fp32 operands on the CUDA cores
// (FFMA), for the reference's fp32 parity bar (per-step noise-prediction rel-L2 <= 1e-4; BASELINE.json
// north_star).  The tensor-core path rounds MMA operands to bf16 (8e-3 rel-L2 per UNet call); tf32 tensor-core
// operands would still cost ~1e-3, so this path keeps every operand in fp32.  It is NOT the throughput path:
// plain 64x64x16 register-tiled SGEMM with the implicit-GEMM (segment, tap, channel) addressing of dfb_gemm, and
// a thread-per-query online-softmax attention.  Same parameter structs, same packed-weight layout (fp32
// storage), so the host orchestration is shared with the bf16 path.
#include "dfb_host.h"
#include "../../include/dfb200.h"

namespace dfb {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

struct SgemmParams {
  const float* a[2];
  int a_ld[2], a_c[2], ntaps[2];
  int tap_dh[18], tap_dw[18], tap_coff[18];
  int nseg, conv, B, H, W, M, N;
  const float* w;
  int w_ld;
  const float* bias;
  const float* rowbias;
  int rowbias_ld, rows_per_batch;
  const float* residual;
  int res_ld;
  float* out;
  int out_ld, act;
};

__global__ void __launch_bounds__(SG_THREADS) sgemm_kernel(const SgemmParams p) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;               // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  // loader mapping: row (A: output pixel, W: output channel) = tid / 4, 4 consecutive k = (tid % 4) * 4
  const int lrow = tid >> 2, lk = (tid & 3) << 2;
  const int am = m0 + lrow;
  int ab = 0, ah = 0, aw = 0;
  if (p.conv && am < p.M) {
    ab = am / (p.H * p.W);
    const int r = am - ab * p.H * p.W;
    ah = r / p.W;
    aw = r - ah * p.W;
  }
  const int wn = n0 + lrow;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int kbase = 0;                                         // k index of (segment, tap, channel 0) in the packed weights
  for (int s = 0; s < p.nseg; ++s) {
    const float* A = p.a[s];
    const int ac = p.a_c[s], acp = (ac + 63) / 64 * 64, ald = p.a_ld[s];
    for (int t = 0; t < p.ntaps[s]; ++t, kbase += acp) {
      const int ti = s * 9 + t;
      const float* arow = nullptr;                       // null: this pixel's tap falls outside the image (zero padding)
      if (am < p.M) {
        if (p.conv) {
          const int hh = ah + p.tap_dh[ti], ww = aw + p.tap_dw[ti];
          if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
            arow = A + ((size_t)(ab * p.H + hh) * p.W + ww) * ald + p.tap_coff[ti];
        } else {
          arow = A + (size_t)am * ald + p.tap_coff[ti];
        }
      }
      for (int c0 = 0; c0 < ac; c0 += SG_BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + lk + e;
          if (arow != nullptr && c < ac) av[e] = __ldg(arow + c);
          if (wn < p.N && c < ac) wv[e] = __ldg(p.w + (size_t)wn * p.w_ld + kbase + c);
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) { As[lk + e][lrow] = av[e]; Ws[lk + e][lrow] = wv[e]; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
          const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
          const float a_[4] = {a4.x, a4.y, a4.z, a4.w}, w_[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], w_[j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float x = acc[i][j];
      if (p.bias) x += __ldg(p.bias + n);
      if (p.rowbias) x += __ldg(p.rowbias + (size_t)(m / p.rows_per_batch) * p.rowbias_ld + n);
      if (p.act == DFB_ACT_SILU) x = silu_f(x);
      else if (p.act == DFB_ACT_LEAKY_RELU) x = x > 0.f ? x : 0.01f * x;
      else if (p.act == DFB_ACT_TANH) x = tanhf(x);
      else if (p.act == DFB_ACT_QUICK_GELU) x = x / (1.0f + expf(-1.702f * x));
      else if (p.act == DFB_ACT_GELU) x = gelu_erf_f(x);
      if (p.residual) x += p.residual[(size_t)m * p.res_ld + n];
      p.out[(size_t)m * p.out_ld + n] = x;
    }
  }
}

// GEGLU on the packed (16 value | 16 gate) column groups of a fp32 [M, N] projection (bias already added):
// out[m, g*16 + j] = in[m, g*32 + j] * gelu_erf(in[m, g*32 + 16 + j])
__global__ void __launch_bounds__(256) geglu_f32_kernel(const float* __restrict__ in, int in_ld, float* __restrict__ out, int out_ld,
                                                        int M, int n_out) {
  const long long total = (long long)M * n_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / n_out), o = (int)(i - (long long)m * n_out);
    const int g = o >> 4, j = o & 15;
    const float* r = in + (size_t)m * in_ld + g * 32;
    out[(size_t)m * out_ld + o] = r[j] * gelu_erf_f(r[16 + j]);
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 attention: block = 128 queries of one (batch, head), thread = query.  Q transposed in shared memory
// ([d][query]: conflict-free), K/V tiles of 32 keys in shared memory (broadcast reads), O in registers.
// ---------------------------------------------------------------------------------------------
constexpr int AF_Q = 128, AF_KV = 32;

template <int DMAX>
__global__ void __launch_bounds__(AF_Q) attn_f32_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                       const float* __restrict__ v, float* __restrict__ out, int q_ld, int k_ld,
                                                       int v_ld, int out_ld, int q_col0, int k_col0, int v_col0, int out_col0,
                                                       int Sq, int Skv, int dp, float scale, int causal) {
  extern __shared__ float sm[];
  float* qs = sm;                         // [dp][AF_Q]
  float* ks = qs + (size_t)dp * AF_Q;     // [AF_KV][dp]
  float* vs = ks + (size_t)AF_KV * dp;    // [AF_KV][dp]
  const int tid = threadIdx.x;
  const int head = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * AF_Q + tid;
  const float* qb = q + (size_t)b * Sq * q_ld + q_col0 + head * dp;
  const float* kb = k + (size_t)b * Skv * k_ld + k_col0 + head * dp;
  const float* vb = v + (size_t)b * Skv * v_ld + v_col0 + head * dp;
  for (int i = tid; i < AF_Q * dp; i += AF_Q) {
    const int r = i / dp, c = i - r * dp;
    const int qq = blockIdx.x * AF_Q + r;
    qs[(size_t)c * AF_Q + r] = qq < Sq ? qb[(size_t)qq * q_ld + c] : 0.f;
  }
  float o[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) o[d] = 0.f;
  float mrun = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < Skv; j0 += AF_KV) {
    __syncthreads();
    for (int i = tid; i < AF_KV * dp; i += AF_Q) {
      const int r = i / dp, c = i - r * dp;
      const bool ok = j0 + r < Skv;
      ks[i] = ok ? kb[(size_t)(j0 + r) * k_ld + c] : 0.f;
      vs[i] = ok ? vb[(size_t)(j0 + r) * v_ld + c] : 0.f;
    }
    __syncthreads();
    int nk = min(AF_KV, Skv - j0);
    if (causal) nk = min(nk, qi - j0 + 1);          // key j0 + j visible only when j0 + j <= qi
    for (int j = 0; j < nk; ++j) {
      float sdot = 0.f;
      const float* kr = ks + (size_t)j * dp;
      for (int d = 0; d < dp; ++d) sdot = fmaf(qs[(size_t)d * AF_Q + tid], kr[d], sdot);
      sdot *= scale;
      const float mnew = fmaxf(mrun, sdot);
      const float alpha = expf(mrun - mnew);         // exp(-inf) = 0 on the first key
      const float pj = expf(sdot - mnew);
      l = l * alpha + pj;
      const float* vr = vs + (size_t)j * dp;
#pragma unroll
      for (int d = 0; d < DMAX; ++d)
        if (d < dp) o[d] = fmaf(o[d], alpha, pj * vr[d]);
      mrun = mnew;
    }
  }
  if (qi < Sq) {
    const float inv = 1.f / l;
    float* orow = out + (size_t)b * Sq * out_ld + (size_t)qi * out_ld + out_col0 + head * dp;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (d < dp) orow[d] = o[d] * inv;
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_gemm_f32(const dfb_gemm_params* q, void* stream) {
  DFB_REQUIRE(q != nullptr, "dfb_gemm_f32: null params");
  DFB_REQUIRE(q->nseg == 1 || q->nseg == 2, "dfb_gemm_f32: nseg must be 1 or 2");
  DFB_REQUIRE(q->M > 0 && q->N > 0, "dfb_gemm_f32: empty problem");
  DFB_REQUIRE(q->w != nullptr && q->out != nullptr && q->a[0] != nullptr, "dfb_gemm_f32: null buffer");
  DFB_REQUIRE(q->out_dtype == DFB_DTYPE_F32 && (q->residual == nullptr || q->res_dtype == DFB_DTYPE_F32),
              "dfb_gemm_f32: output and residual must be fp32");
  DFB_REQUIRE(!q->geglu && q->gn_partial == nullptr, "dfb_gemm_f32: GEGLU pairing is dfb_geglu_f32; no fused GroupNorm statistics");
  SgemmParams p;
  memset(&p, 0, sizeof(p));
  int kp_total = 0;
  for (int s = 0; s < q->nseg; ++s) {
    DFB_REQUIRE(q->a[s] != nullptr && q->a_c[s] > 0 && q->ntaps[s] >= 1 && q->ntaps[s] <= 9, "dfb_gemm_f32: bad A segment");
    p.a[s] = (const float*)q->a[s];
    p.a_ld[s] = q->a_ld[s];
    p.a_c[s] = q->a_c[s];
    p.ntaps[s] = q->ntaps[s];
    for (int t = 0; t < q->ntaps[s]; ++t) {
      p.tap_dh[s * 9 + t] = q->tap_dh[s][t];
      p.tap_dw[s * 9 + t] = q->tap_dw[s][t];
      p.tap_coff[s * 9 + t] = q->tap_coff[s][t];
      DFB_REQUIRE(q->tap_coff[s][t] >= 0 && q->tap_coff[s][t] + q->a_c[s] <= q->a_ld[s], "dfb_gemm_f32: channel extent exceeds the row pitch");
      if (!q->conv) DFB_REQUIRE(q->tap_dh[s][t] == 0 && q->tap_dw[s][t] == 0, "dfb_gemm_f32: shifts need conv addressing");
    }
    kp_total += q->ntaps[s] * ((q->a_c[s] + 63) / 64 * 64);
  }
  DFB_REQUIRE(q->act >= DFB_ACT_NONE && q->act <= DFB_ACT_GELU, "dfb_gemm_f32: unknown activation");
  DFB_REQUIRE(q->up2x == 0, "dfb_gemm_f32: the fused upsample phases exist on the tcgen05 path only (upsample, then convolve)");
  DFB_REQUIRE(kp_total <= q->w_ld, "dfb_gemm_f32: packed weight K extent smaller than the A operand implies");
  if (q->conv) DFB_REQUIRE(q->B > 0 && q->H > 0 && q->W > 0 && (long long)q->B * q->H * q->W == q->M, "dfb_gemm_f32: conv geometry does not match M");
  p.nseg = q->nseg; p.conv = q->conv ? 1 : 0; p.B = q->B; p.H = q->H; p.W = q->W; p.M = q->M; p.N = q->N;
  p.w = (const float*)q->w; p.w_ld = q->w_ld;
  p.bias = q->bias; p.rowbias = q->rowbias; p.rowbias_ld = q->rowbias_ld;
  p.rows_per_batch = q->rows_per_batch > 0 ? q->rows_per_batch : 1;
  p.residual = (const float*)q->residual; p.res_ld = q->res_ld;
  p.out = (float*)q->out; p.out_ld = q->out_ld; p.act = q->act;
  dim3 grid((q->N + SG_BN - 1) / SG_BN, (q->M + SG_BM - 1) / SG_BM);
  DFB_REQUIRE(grid.y <= 65535u, "dfb_gemm_f32: M too large (verification path: at most 4M rows per call)");
  sgemm_kernel<<<grid, SG_THREADS, 0, (cudaStream_t)stream>>>(p);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_geglu_f32(const float* in, int in_ld, float* out, int out_ld, int M, int N, void* stream) {
  DFB_REQUIRE(in && out && M > 0 && N > 0 && N % 32 == 0, "dfb_geglu_f32: N must be a positive multiple of 32");
  const long long total = (long long)M * (N / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  geglu_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, in_ld, out, out_ld, M, N / 2);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

int dfb_attention_f32(const dfb_attn_params* a, void* stream) {
  DFB_REQUIRE(a && a->q && a->k && a->v && a->out, "dfb_attention_f32: null buffer");
  DFB_REQUIRE(a->B > 0 && a->heads > 0 && a->Sq > 0 && a->Skv > 0, "dfb_attention_f32: empty problem");
  DFB_REQUIRE(a->dp > 0 && a->dp <= 160, "dfb_attention_f32: head dim must be in [1,160]");
  DFB_REQUIRE(a->B <= 65535 && a->heads <= 65535, "dfb_attention_f32: grid limits");
  const size_t smem = ((size_t)a->dp * AF_Q + 2 * (size_t)AF_KV * a->dp) * sizeof(float);
  dim3 grid((a->Sq + AF_Q - 1) / AF_Q, a->heads, a->B);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set[64] = {false};
  int dev = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[dev] = true;
  }
#define DFB_AF_LAUNCH(D)                                                                                                        \
  attn_f32_kernel<D><<<grid, AF_Q, smem, st>>>((const float*)a->q, (const float*)a->k, (const float*)a->v, (float*)a->out,     \
                                                a->q_ld, a->k_ld, a->v_ld, a->out_ld, a->q_col0, a->k_col0, a->v_col0,        \
                                                a->out_col0, a->Sq, a->Skv, a->dp, a->scale, a->causal)
  if (a->dp <= 32) DFB_AF_LAUNCH(32);
  else if (a->dp <= 64) DFB_AF_LAUNCH(64);
  else if (a->dp <= 96) DFB_AF_LAUNCH(96);
  else DFB_AF_LAUNCH(160);
#undef DFB_AF_LAUNCH
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}

}  // extern "C"
