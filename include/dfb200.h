/* dfb200 — C ABI of the B200-native DiFashion denoising-step kernels (libdfb200.so).
 *
 * The reference (YiyanXu/DiFashion) has no FFI layer: the hot path is the diffusers 0.18.2 Python
 * object API consumed at DiFashion/models/difashion.py:456-577.  Each entry point below names the
 * reference call site (or the diffusers op behind it) that it replaces.  Conventions:
 *   - every function returns 0 (DFB_OK) or a negative error code; dfb_strerror / dfb_last_error
 *     give text.  Nothing here synchronises the device or allocates device memory: the CALLER
 *     owns every buffer and passes the CUDA stream to enqueue on (cudaStream_t as void*).
 *   - activations are NHWC ("pixels x channels") row-major; bf16 = raw uint16 storage.
 *   - thread-safe for distinct streams / devices (one process per GPU is the intended model).
 */
#ifndef DFB200_H_
#define DFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB200_ABI_VERSION 6

#define DFB_OK 0
#define DFB_ERR_INVALID (-1)
#define DFB_ERR_CUDA (-2)
#define DFB_ERR_UNSUPPORTED (-3)
#define DFB_ERR_NO_DRIVER (-4)

#define DFB_DTYPE_BF16 0
#define DFB_DTYPE_F32 1

#define DFB_ACT_NONE 0
#define DFB_ACT_SILU 1
#define DFB_ACT_LEAKY_RELU 2 /* slope 0.01 (nn.LeakyReLU default, MutualEncoder) */
#define DFB_ACT_TANH 3
#define DFB_ACT_QUICK_GELU 4 /* x * sigmoid(1.702 x): CLIPTextModel mlp (hidden_act "quick_gelu") */
#define DFB_ACT_GELU 5       /* exact erf GELU: CLIPTextModel mlp with hidden_act "gelu" (SD-2's OpenCLIP text encoder) */

const char* dfb_strerror(int rc);
const char* dfb_last_error(void);
int dfb_abi_version(void);
int dfb_num_sms(void);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM convolution (tcgen05 + TMEM + TMA).
 *
 * Replaces: nn.Conv2d 3x3 / 1x1 and nn.Linear inside diffusers UNet2DConditionModel.forward
 * (reference call DiFashion/models/difashion.py:518-523): ResnetBlock2D.conv1/conv2/
 * conv_shortcut, Downsample2D.conv, Upsample2D.conv, Transformer2DModel.proj_in/proj_out,
 * Attention.to_q/to_k/to_v/to_out, GEGLU.proj, FeedForward.net[2], time_emb_proj, conv_in/out;
 * and MutualEncoder.mlp Linear layers (difashion.py:31-37).
 *
 *   out[m, n] = epilogue( sum_seg sum_tap sum_c  A_seg[pixel(m) + tap, c] * Wt[n, k(seg,tap,c)] )
 *
 * A operand: up to two segments (bf16).  conv == 0: A_seg is a row-major [M, a_c] matrix with
 * row pitch a_ld elements.  conv == 1: A_seg is NHWC [B, H, W, a_c-slice] with pixel pitch a_ld
 * elements; each tap reads the input shifted by (dh, dw) (zero outside the image) at channel
 * offset coff.  Weights Wt: bf16 [N, Kp] row-major ("K-major"), Kp = sum over segments of
 * ntaps * ceil64(a_c), k index ordered (segment, tap, channel).
 * Epilogue (fp32): + bias[n] + rowbias[(m / rows_per_batch) * rowbias_ld + n]
 *                  + residual[m * res_ld + n]; optional GEGLU pairing;
 * result stored as bf16 or fp32 at out[m * out_ld + n].
 * ------------------------------------------------------------------------------------------ */
typedef struct dfb_gemm_params {
  const void* a[2];      /* bf16 A segments (a[1] may be NULL when nseg == 1)            */
  int32_t a_ld[2];       /* elements between consecutive rows / pixels                     */
  int32_t a_c[2];        /* channels (K extent) of the segment; reads beyond are zero      */
  int32_t ntaps[2];      /* taps per segment (1 for a plain GEMM)                          */
  int32_t tap_dh[2][9];  /* per tap: row shift, column shift, channel offset (elements)    */
  int32_t tap_dw[2][9];
  int32_t tap_coff[2][9];
  int32_t nseg;
  int32_t conv;          /* 0 = plain [M,K]; 1 = NHWC shifted-window addressing            */
  int32_t B, H, W;       /* geometry when conv == 1 (M must equal B*H*W)                   */
  int32_t M, N;
  const void* w;         /* bf16 [N, w_ld]                                                 */
  int32_t w_ld;          /* = Kp                                                           */
  const float* bias;     /* [N] or NULL (for GEGLU: packed in the interleaved order)       */
  const float* rowbias;  /* [M / rows_per_batch, rowbias_ld] or NULL                       */
  int32_t rowbias_ld;
  int32_t rows_per_batch;
  const void* residual;  /* [M, res_ld] or NULL                                            */
  int32_t res_ld;
  int32_t res_dtype;     /* DFB_DTYPE_*                                                    */
  void* out;             /* [M, out_ld]                                                    */
  int32_t out_ld;
  int32_t out_dtype;     /* DFB_DTYPE_*                                                    */
  int32_t geglu;         /* 1: columns come in (16 value | 16 gate) groups; N_out = N / 2  */
  int32_t act;           /* DFB_ACT_* applied after bias/rowbias, before the residual      */
  int32_t block_n;       /* N tile (multiple of 32, <= 256); 0 = choose automatically      */
  float* gn_partial;     /* NULL, or fp32 [M/32, N/2, 2]: per (32-row block, channel pair) sum and sum of
                            squares of the fp32 output, emitted from the epilogue for a following GroupNorm
                            (dfb_groupnorm_fused); needs fp32 output, M % 32 == 0, N % 4 == 0            */
  int32_t cta_group;     /* 0 = automatic; 1 = one CTA per 128-row tile; 2 = CTA pair (cluster of 2 on one TPC,
                            tcgen05.mma.cta_group::2 over 256 rows, each CTA feeding half of the B tile)  */
  int32_t up2x;          /* 0 = off.  1 + 2a + b (a, b in {0, 1}): this conv computes phase (a, b) of a nearest-2x upsample
                            followed by a 3x3 convolution (diffusers Upsample2D) on the LOW-resolution input: output
                            pixel (2i + a, 2j + b) depends on a 2x2 input window only, with the 3x3 weights summed per
                            window tap (4 phase GEMMs of 4 taps instead of one 9-tap GEMM on 4x the pixels: 16/36 of the
                            MACs).  Needs conv == 1; row m = (b*H + i)*W + j is stored at row
                            (2*(m / W) + a) * 2W + 2*(m % W) + b of `out` ([B, 2H, 2W, N]); gn_partial (if given) is the
                            buffer of the whole [B*2H*2W, N] output and receives this phase's blocks at
                            [img * 4*HW/32 + phase * HW/32 + block].  tcgen05 path only.                      */
} dfb_gemm_params;

int dfb_gemm(const dfb_gemm_params* p, void* stream);

/* fp32 verification path (BASELINE.json north_star: per-step noise-prediction rel-L2 <= 1e-4 in fp32): the same
 * operator with fp32 A segments, fp32 packed weights (same [N, Kp] layout), fp32 residual / output, on the CUDA
 * cores (FFMA).  Same struct; geglu must be 0 (pair the packed columns with dfb_geglu_f32 afterwards:
 * out[m, g*16+j] = in[m, g*32+j] * gelu_erf(in[m, g*32+16+j])), gn_partial must be NULL, block_n is ignored.
 * Not a throughput path. */
int dfb_gemm_f32(const dfb_gemm_params* p, void* stream);
int dfb_geglu_f32(const float* in, int in_ld, float* out, int out_ld, int M, int N, void* stream);

/* ------------------------------------------------------------------------------------------
 * Flash attention forward (tcgen05 QK^T and PV, online softmax, TMA-fed).
 *
 * Replaces: the attention processor call `attn.processor(attn, hidden_states,
 * encoder_hidden_states)` of every diffusers Attention layer (attn1 self / attn2 cross) in
 * UNet2DConditionModel.forward — xformers.memory_efficient_attention in the reference
 * (DiFashion/models/difashion.py:109-118) — softmax(Q K^T * scale) V per (batch, head).
 *
 * q: bf16 [B, Sq, q_ld], k: [B, Skv, k_ld], v: [B, Skv, v_ld], out: [B, Sq, out_ld]; head h
 * occupies columns [col0 + h*dp, col0 + (h+1)*dp) of its operand, dp = head dim padded to a
 * multiple of 16 with zero columns.  Skv needs no padding in memory (TMA zero-fills, the
 * kernel masks columns >= Skv).
 * ------------------------------------------------------------------------------------------ */
typedef struct dfb_attn_params {
  const void* q; const void* k; const void* v; void* out;
  int32_t q_ld, k_ld, v_ld, out_ld;
  int32_t q_col0, k_col0, v_col0, out_col0;
  int32_t B, heads, Sq, Skv, dp;
  float scale;            /* softmax scale (true head_dim ** -0.5)                          */
  int32_t block_kv;       /* KV tile (multiple of 16, <= 128); 0 = automatic                */
  int32_t dbg_v_lbo, dbg_v_sbo; /* 0 = default; test hooks for the V descriptor strides     */
  int32_t dbg_flags;      /* 0 = default; tuning hooks: bit1 one CTA per SM, bit3 single-buffer kernel, bit4 P via smem, bit6 no short-KV kernel, bits 8-11 query tiles per CTA of the short-KV kernel, bit12 round-1 double-buffered kernel instead of attn_fwd_sa_kernel, bits 13-14 share of the exponentials on the FMA pipe (1 none, 2 = 2/16, 3 = 4/16), bit15 no 8-softmax-warp kernel, bits 16-19 timing perturbation of the 8-softmax-warp kernel for the robustness tests (slow MMA issuer / TMA producer / odd- / even-tile softmax warps) */
  void* dbg_timeline;     /* NULL, or device buffer of >= 4096 int64: clock64 stamps of CTA (0,0,0) (tuning)  */
  int32_t causal;         /* 1: key j is visible to query i only when j <= i (CLIPTextModel's causal mask,
                             DiFashion/models/difashion.py:339-353); 0 everywhere in the UNet              */
  int32_t ones_col;       /* 0: none.  c + 1: the caller promises V[:, head*dp + c] == 1.0 for every key row and head (a padding
                             column of the padded head, written by the q|k|v projection's bias): the long-sequence
                             self-attention kernel then takes the softmax denominator from O[:, c] — accumulated by the P V
                             MMA from the same bf16 probabilities as the numerator — and out[:, head*dp + c] comes back as 1.
                             Kernels that keep their own running sum ignore it.                                  */
  void* workspace;        /* NULL, or caller-owned device scratch of >= dfb_attention_ws_bytes(B, heads, Sq) bytes (no
                             initialisation needed): lets the self-attention path with ones_col use the 8-softmax-warp kernel,
                             which flags tiles whose scores overflow its static reference maximum there; a second launch on
                             the same stream recomputes flagged tiles exactly.                                     */
} dfb_attn_params;

size_t dfb_attention_ws_bytes(int B, int heads, int Sq);

int dfb_attention(const dfb_attn_params* p, void* stream);
/* fp32 verification path: q/k/v/out fp32, same layout conventions (dp <= 160), expf softmax on the CUDA cores. */
int dfb_attention_f32(const dfb_attn_params* p, void* stream);

/* ABI self-check for language bindings (ctypes / cgo / JNI mirrors of the structs above). */
size_t dfb_sizeof_gemm_params(void);
size_t dfb_sizeof_attn_params(void);

/* ------------------------------------------------------------------------------------------
 * Norm kernels producing the GEMM operands from the fp32 residual stream (NHWC).  Every operand-producing
 * kernel below takes `out_dtype`: DFB_DTYPE_BF16 for the tensor-core path, DFB_DTYPE_F32 for the fp32
 * verification path (dfb_gemm_f32 / dfb_attention_f32).
 * dfb_groupnorm replaces nn.GroupNorm(32, C)(+SiLU) of ResnetBlock2D.norm1/norm2,
 * Transformer2DModel.norm and conv_norm_out; reads the up-block skip concat from its two sources
 * (torch.cat([h, skip], 1) is never materialised); optional raw copy of the input in the output dtype (the A
 * operand of conv_shortcut).  stats_ws: fp32 scratch of dfb_groupnorm_ws_floats(B, groups) elements.
 * Deterministic (fixed-order reductions, no atomics).
 * dfb_layernorm replaces BasicTransformerBlock.norm1/2/3.
 * ------------------------------------------------------------------------------------------ */
size_t dfb_groupnorm_ws_floats(int B, int groups);
int dfb_groupnorm(const float* src0, int c0, int ld0, const float* src1, int c1, int ld1, int B, int hw,
                  int groups, float eps, const float* gamma, const float* beta, int silu, float* stats_ws,
                  void* out, int out_dtype, int ld_out, void* raw_out, int ld_raw, void* stream);
/* GroupNorm whose statistics come from the producers' epilogues (dfb_gemm_params.gn_partial) instead of an
 * extra pass over the input: finalize (fixed-order fold of the partials) + apply.  hw % 32 == 0. */
int dfb_groupnorm_fused(const float* src0, int c0, int ld0, const float* partial0, const float* src1, int c1, int ld1,
                        const float* partial1, int B, int hw, int groups, float eps, const float* gamma,
                        const float* beta, int silu, float* stats_ws, void* out, int out_dtype, int ld_out, void* raw_out,
                        int ld_raw, void* stream);
int dfb_layernorm(const float* x, int ld_x, const float* gamma, const float* beta, float eps, void* out, int out_dtype,
                  int ld_out, int rows, int C, void* stream);

/* Row softmax of a materialised score matrix: out[r, :] = softmax(scale * in[r, :]) (fp32 in, bf16/fp32 out).
 * Replaces the softmax of the VAE decoder's single-head d=512 mid-block attention (AutoencoderKL.decode,
 * DiFashion/models/difashion.py:579), which runs as dfb_gemm launches around this kernel. */
int dfb_softmax_rows(const float* in, int in_ld, float scale, void* out, int out_dtype, int out_ld, int rows, int cols,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-streaming kernels of the loop body (DiFashion/models/difashion.py:456-577).
 * ------------------------------------------------------------------------------------------ */
/* CFG combine (difashion.py:525-566) fused with the scheduler update (:569; diffusers
 * DDIMScheduler.step / PNDMScheduler.step_plms):
 *   e0 = sum_b w[b] * eps[b]   (eps: fp32, branch-major [nb*N, ...]; NHWC [hw,4] from the UNet kernels, or
 *                               NCHW [4,hw] when eps_nchw != 0 — the public scheduler.step() path)
 *   x_out = cx*x_src + ck[0]*e0 + ck[1]*hist1 + ck[2]*hist2 + ck[3]*hist3 + cn*noise   (fp32 NCHW [N,4,hw])
 *   eps_out (optional) = e0.   w: host float[nb]; ck: host float[4].                           */
int dfb_cfg_step(const float* eps, int eps_nchw, int nb, const float* w, const float* x_src, float cx, const float* ck,
                 const float* hist1, const float* hist2, const float* hist3, const float* noise, float cn,
                 float* x_out, float* eps_out, int n_items, int hw, void* stream);
/* Mutual-condition neighbour sum (difashion.py:475-488): out[n] = sum_s src(idx[n, s]); idx >= 0 ->
 * all_latents row, idx < 0 -> prev_latents row (-idx-1), INT32_MIN -> skip.  out [N, d]. */
int dfb_mutual_gather_sum(const float* all_latents, const float* prev_latents, const int32_t* idx, int n_items,
                          int n_src, int d, void* out, int out_dtype, void* stream);
/* Mutual blend + history concat + CFG branch expansion (difashion.py:458-459, :494-515, :388-390):
 * writes the UNet input NHWC [nb*N, hw, 8].  use_m / use_h: host int32[nb].              */
int dfb_mutual_blend(const float* x, const float* m, const float* hist, const float* null_latent, float eta,
                     int nb, const int32_t* use_m, const int32_t* use_h, int n_items, int hw, void* out, int out_dtype,
                     void* stream);
/* Layout conversions at the diffusers API boundary (NCHW tensors in, NHWC inside). */
int dfb_nchw_to_nhwc(const void* in, int in_dtype, void* out, int out_dtype, int B, int C, int HW, void* stream);
int dfb_nhwc_to_nchw(const float* in, void* out, int out_dtype, int B, int C, int HW, void* stream);
int dfb_pad_cast_rows(const void* in, int in_dtype, void* out, int out_dtype, int B, int S, int S_pad, int D, void* stream);
/* fp32 [n] -> MMA operand dtype (n % 4 == 0, 16-byte aligned): the low-resolution input of the fused upsample phases. */
int dfb_cast_f32(const float* in, void* out, int out_dtype, long long n, void* stream);
/* Upsample2D nearest-2x (fp32 NHWC -> NHWC operand) and the space-to-depth feeding Downsample2D's
 * stride-2 conv (fp32 NHWC [B,H,W,C] -> [B,H/2,W/2,4C]). */
int dfb_upsample2x(const float* in, void* out, int out_dtype, int B, int H, int W, int C, void* stream);
int dfb_space_to_depth(const float* in, void* out, int out_dtype, int B, int H, int W, int C, void* stream);
/* CLIPTextEmbeddings (transformers CLIPTextModel, called at DiFashion/models/difashion.py:339-341, :352):
 * out[b*S + s, :] = token_table[ids[b*S + s], :] + position_table[s, :]   (fp32 [B*S, D]; ids int32 in [0, vocab)). */
int dfb_embed_tokens(const int32_t* ids, const float* token_table, const float* position_table, float* out, int B, int S,
                     int D, int vocab, void* stream);
/* VaeImageProcessor.postprocess(image, output_type="pil"/"np") after AutoencoderKL.decode (difashion.py:579-592):
 * out[p, c] = uint8(round(clamp(in[p * in_c + c] / 2 + 0.5, 0, 1) * 255)), c < 3 (round half to even, as numpy).
 * in: fp32 NHWC with in_c >= 3 channels per pixel; out: uint8 [pixels, 3] (HWC RGB, what PIL / np.save consume). */
int dfb_image_to_uint8(const float* in, int in_c, uint8_t* out, long long pixels, void* stream);
/* diffusers Timesteps (get_timestep_embedding): t fp32 [B] -> [B, dim] = [cos | sin]. */
int dfb_timestep_embedding(const float* t, void* out, int out_dtype, int B, int dim, int flip_sin_to_cos, float freq_shift,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFB200_H_ */
