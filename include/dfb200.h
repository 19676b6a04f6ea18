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

#define DFB200_ABI_VERSION 1

#define DFB_OK 0
#define DFB_ERR_INVALID (-1)
#define DFB_ERR_CUDA (-2)
#define DFB_ERR_UNSUPPORTED (-3)
#define DFB_ERR_NO_DRIVER (-4)

#define DFB_DTYPE_BF16 0
#define DFB_DTYPE_F32 1

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
  int32_t block_n;       /* N tile (multiple of 32, <= 256); 0 = choose automatically      */
} dfb_gemm_params;

int dfb_gemm(const dfb_gemm_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFB200_H_ */
