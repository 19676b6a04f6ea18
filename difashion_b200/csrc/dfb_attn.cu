// Flash-style attention forward on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces diffusers' Attention processors (XFormersAttnProcessor / AttnProcessor2_0) that run
// inside UNet2DConditionModel.forward (reference call DiFashion/models/difashion.py:518-523,
// xformers enabled at :109-118): softmax(Q K^T * scale) V per (batch row, head).
//
// One CTA = one 128-row query tile of one (batch, head).  192 threads:
//   warp 0     : TMA producer (Q once; K/V tiles through a 2-stage ring)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5 : softmax (thread <-> query row <-> TMEM lane): S from TMEM, online max/sum in the
//                exp2 domain with lazy rescaling, P (bf16) -> shared memory, O rescale in TMEM,
//                final O / l -> bf16 global store.
// Head dim is padded to a multiple of 16 (d=40 -> 48) by the projection weights' packing; every
// operand is staged as 16-element (32-byte) column chunks with the 32-byte swizzle:
//   Q, K : K-major  A / B operands of S = Q K^T        (N = block_kv, K = dp)
//   P    : K-major  A operand of O += P V              (K = block_kv), written by the softmax warps
//   V    : MN-major B operand ([kv rows][16 d-cols] chunks; LBO = chunk stride, SBO = 8 kv rows)
// S occupies TMEM columns [0, block_kv), O columns [block_kv, block_kv + dp).
// Several CTAs are resident per SM so one CTA's MMAs overlap another's exponentials.
#include "dfb_host.h"
#include "../../include/dfb200.h"

namespace dfb {

constexpr int ATT_BLOCK_Q = 128;
#ifndef DFB_ATTN_POLY_ONES_DEFAULT
#define DFB_ATTN_POLY_ONES_DEFAULT 0      // until measured: every exponential on the MUFU
#endif
#ifndef DFB_ATTN_POLY_DEFAULT
#define DFB_ATTN_POLY_DEFAULT 0
#endif
#ifndef DFB_ATTN_SA8_TILES_DEFAULT
#define DFB_ATTN_SA8_TILES_DEFAULT 1      // measured on B200 (B = 64, S = 4096, d = 40): column split 2.905 ms, tile split 2.783 ms
#endif
#ifndef DFB_ATTN_SA8_POLY_DEFAULT
#define DFB_ATTN_SA8_POLY_DEFAULT 2      // measured on B200: 2.903 (none) / 2.864 (2 of 16) / 3.017 (4) / 3.463 ms (8) at B = 64, S = 4096, d = 40
#endif
constexpr int ATT_THREADS = 192;
constexpr int ATT_STAGES = 2;

struct AttnMaps {
  CUtensorMap q, k, v;
};

struct AttnKernelParams {
  int Sq, Skv, dp, block_kv, n_kv_tiles;
  int q_col0, k_col0, v_col0;     // column of head 0 in each operand (elements)
  float scale_log2;
  __nv_bfloat16* out;
  int out_ld, out_col0;
  long long out_batch_stride;     // elements
  uint32_t tmem_cols;
  uint32_t v_lbo, v_sbo;          // MN-major V descriptor strides (bytes)
  long long* timeline;            // tuning hook: clock64 stamps of the first softmax thread of CTA (0,0,0)
  int kv_stages;                  // K/V ring depth of the double-buffered kernel (2 or 3)
  int q_tiles;                    // query tiles per CTA of the short-KV kernel
  int causal;                     // key j visible to query i only when j <= i (generic single-buffer kernel only)
  int l_col;                      // attn_fwd_sa_kernel<ONES>: column of O that accumulates the softmax denominator (V holds 1.0 there)
  int s_ring;                     // attn_fwd_sa_kernel: score / probability buffers in tensor memory (2 or 3)
  int tl_second;                  // attn_fwd_sa_kernel tuning hook: linear index of the second CTA that writes stamps
  int* redo_flags;                // attn_fwd_sa8_kernel -> attn_fwd_sa_kernel<.., REDO>: one flag per CTA (caller's workspace)
  int dbg_delay;                  // test hooks of attn_fwd_sa8_kernel (dbg_flags bits 16-19): bit0 slow MMA issuer, bit1 slow TMA producer, bit2 / bit3 the softmax warps of the odd / even tiles lag
  int n_q_tiles;                  // query tiles per (batch, head) = grid.x of the per-tile kernels (the REDO kernel scans that many flags)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (Measured and rejected: moving 3/16 of the exponentials to an FMA-pipe polynomial.  With two CTAs per SM the
// softmax warps then become issue-bound and the kernel gets 13 % slower — profiles/r01_attention_experiments.md.)
// KV_STATIC > 0: block_kv == KV_STATIC, the score row is held in registers (one TMEM pass);
// KV_STATIC == 0: any block_kv (multiple of 16), two TMEM passes (max, then exponentials).
template <int KV_STATIC>
__global__ void __launch_bounds__(ATT_THREADS, KV_STATIC == 0 ? 1 : 2)
attn_fwd_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int dchunks = p.dp >> 4;
  const uint32_t q_bytes = (uint32_t)dchunks * ATT_BLOCK_Q * 32u;
  const uint32_t kv_chunk_bytes = (uint32_t)p.block_kv * 32u;
  const uint32_t kv_tile_bytes = (uint32_t)dchunks * kv_chunk_bytes;       // K (or V) tile
  const uint32_t p_bytes = (uint32_t)(p.block_kv >> 4) * ATT_BLOCK_Q * 32u;
  const uint32_t sQ = smem_base;
  const uint32_t sP = sQ + q_bytes;
  const uint32_t sKV = sP + p_bytes;                                       // stage s: K then V
  const uint32_t bar_base = sKV + ATT_STAGES * 2 * kv_tile_bytes;
  const uint32_t q_full = bar_base;
  const uint32_t s_full = bar_base + 8;
  const uint32_t p_full = bar_base + 16;
  const uint32_t o_done = bar_base + 24;
  auto kv_full = [&](int s) { return bar_base + 32u + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 32u + 8u * (ATT_STAGES + s); };
  const uint32_t tmem_ptr_smem = bar_base + 32u + 8u * (2 * ATT_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    for (int s = 0; s < ATT_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + (uint32_t)p.block_kv;

  if (warp == 0) {
    // ---------------- TMA producer (warp-uniform loop, one elected lane issues) ----------------
    if (elect_one()) {
      mbar_expect_tx(q_full, q_bytes);
      for (int c = 0; c < dchunks; ++c)
        tma_load_3d(&maps.q, sQ + (uint32_t)c * ATT_BLOCK_Q * 32u, q_full, p.q_col0 + head * p.dp + c * 16,
                    qt * ATT_BLOCK_Q, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % ATT_STAGES;
      const uint32_t ph = (uint32_t)(j / ATT_STAGES) & 1u;
      mbar_wait(kv_empty(st), ph ^ 1u);
      if (elect_one()) {
        const uint32_t sK = sKV + (uint32_t)st * 2 * kv_tile_bytes;
        const uint32_t sV = sK + kv_tile_bytes;
        mbar_expect_tx(kv_full(st), 2 * kv_tile_bytes);
        for (int c = 0; c < dchunks; ++c)
          tma_load_3d(&maps.k, sK + (uint32_t)c * kv_chunk_bytes, kv_full(st), p.k_col0 + head * p.dp + c * 16,
                      j * p.block_kv, b);
        for (int c = 0; c < dchunks; ++c)
          tma_load_3d(&maps.v, sV + (uint32_t)c * kv_chunk_bytes, kv_full(st), p.v_col0 + head * p.dp + c * 16,
                      j * p.block_kv, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // Whole warp runs the warp-uniform loop (descriptors stay in uniform registers); one elected lane
    // issues tcgen05.mma / tcgen05.commit.
    const uint32_t idesc_qk = make_idesc_f16(ATT_BLOCK_Q, (uint32_t)p.block_kv, true, 0, 0);
    const uint32_t idesc_pv = make_idesc_f16(ATT_BLOCK_Q, (uint32_t)p.dp, true, 0, 1);
    const uint64_t desc_q0 = make_smem_desc(sQ, 16, 256, SWZ_32B);
    const uint64_t desc_p0 = make_smem_desc(sP, 16, 256, SWZ_32B);
    const uint64_t desc_k0 = make_smem_desc(sKV, 16, 256, SWZ_32B);
    const uint64_t desc_v0 = make_smem_desc(sKV + kv_tile_bytes, p.v_lbo, p.v_sbo, SWZ_32B);
    const uint32_t stage_step = (2 * kv_tile_bytes) >> 4;
    const uint32_t kchunk_step = kv_chunk_bytes >> 4;
    const int pv_steps = p.block_kv >> 4;
    auto issue_qk = [&](int st) {
      if (elect_one()) {
        const uint64_t dk = desc_k0 + (uint64_t)((uint32_t)st * stage_step);
        for (int c = 0; c < dchunks; ++c)
          umma_f16_ss(tmem_S, desc_q0 + (uint64_t)(c * (ATT_BLOCK_Q * 32 / 16)), dk + (uint64_t)((uint32_t)c * kchunk_step),
                      idesc_qk, c != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % ATT_STAGES;
      mbar_wait(p_full, (uint32_t)j & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc_v0 + (uint64_t)((uint32_t)st * stage_step);
        for (int k = 0; k < pv_steps; ++k)
          umma_f16_ss(tmem_O, desc_p0 + (uint64_t)(k * (ATT_BLOCK_Q * 32 / 16)), dv + (uint64_t)(k * (512 / 16)), idesc_pv,
                      (j | k) != 0);
        umma_commit(o_done);
        umma_commit(kv_empty(st));
      }
      __syncwarp();
      if (j + 1 < n_tiles) {
        const int st2 = (j + 1) % ATT_STAGES;
        mbar_wait(kv_full(st2), (uint32_t)((j + 1) / ATT_STAGES) & 1u);
        tc_fence_after();
        issue_qk(st2);
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue warps ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int q_row = qt * ATT_BLOCK_Q + row;
    float m_ref = -INFINITY, l = 0.f;
    const int kv_chunks = p.block_kv >> 4;
    const uint32_t p_row = sP + (uint32_t)row * 32u;
    const uint32_t flip = (uint32_t)((row >> 2) & 1) << 4;     // 32B-swizzle: 16B halves swap on rows 4..7 of 8
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int kv_valid = min(p.block_kv, p.Skv - j * p.block_kv);
      if constexpr (KV_STATIC > 0) {
        // ---- score row in registers.  Fast path: exponentials are computed OPTIMISTICALLY against the
        // running reference max m_ref while the TMEM load of the next 32-column chunk is in flight
        // (TMEM->RF bandwidth and the MUFU are both ~1 tile-time resources, so they must overlap);
        // the tile max is tracked on the side and, when it exceeds m_ref by more than 2^8 (always on
        // the first tile, rarely afterwards), the tile is redone from the registers with a new
        // reference (careful path: O and l rescaled). ----
        uint32_t sreg[KV_STATIC];
        if (j > 0) {
          mbar_wait(o_done, (uint32_t)(j - 1) & 1u);   // PV_{j-1} retired: P buffer free, O stable
          tc_fence_after();
        }
        bool careful = (kv_valid != KV_STATIC) || (j == 0);
        float mx = -INFINITY;
        if (!careful) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
          tmem_ld_32x32b_x32(tmem_S + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&sreg[0]));
#pragma unroll
          for (int c = 0; c < KV_STATIC / 32; ++c) {
            tmem_ld_wait();                                  // chunk c has landed
            if (c + 1 < KV_STATIC / 32)
              tmem_ld_32x32b_x32(tmem_S + lane_addr + (uint32_t)((c + 1) * 32),
                                 *reinterpret_cast<uint32_t(*)[32]>(&sreg[(c + 1) * 32]));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float pv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __uint_as_float(sreg[c * 32 + h * 16 + i]);
                m4[i & 3] = fmaxf(m4[i & 3], sv);
                pv[i] = ex2f(fmaf(sv, p.scale_log2, -m_ref));
                l4[i & 3] += pv[i];
              }
              const uint32_t dst = p_row + (uint32_t)(c * 2 + h) * ATT_BLOCK_Q * 32u;
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (0u ^ flip)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (16u ^ flip)), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
          careful = __any_sync(0xffffffffu, mx > m_ref + 8.0f);
          if (!careful) l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        } else {
#pragma unroll
          for (int c = 0; c < KV_STATIC / 32; ++c)
            tmem_ld_32x32b_x32(tmem_S + lane_addr + (uint32_t)(c * 32), *reinterpret_cast<uint32_t(*)[32]>(&sreg[c * 32]));
          tmem_ld_wait();
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < KV_STATIC; ++i)
            if (i < kv_valid) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sreg[i]));
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
        }
        if (careful) {
          const bool need = mx > m_ref + 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? mx : m_ref;
            const float alpha = ex2f(m_ref - m_new);     // m_ref = -inf on the first tile -> 0
            if (j > 0) {
              for (int c = 0; c < dchunks; ++c) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                tmem_st_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
              }
              tmem_st_wait();
            }
            l *= alpha;
            m_ref = m_new;
          }
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < KV_STATIC / 16; ++c) {
            float pv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float e = ex2f(fmaf(__uint_as_float(sreg[c * 16 + i]), p.scale_log2, -m_ref));
              pv[i] = (c * 16 + i < kv_valid) ? e : 0.f;
              l4[i & 3] += pv[i];
            }
            const uint32_t dst = p_row + (uint32_t)c * ATT_BLOCK_Q * 32u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (0u ^ flip)), "r"(pack_bf16x2(pv[0], pv[1])),
                         "r"(pack_bf16x2(pv[2], pv[3])), "r"(pack_bf16x2(pv[4], pv[5])), "r"(pack_bf16x2(pv[6], pv[7]))
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (16u ^ flip)), "r"(pack_bf16x2(pv[8], pv[9])),
                         "r"(pack_bf16x2(pv[10], pv[11])), "r"(pack_bf16x2(pv[12], pv[13])), "r"(pack_bf16x2(pv[14], pv[15]))
                         : "memory");
          }
          l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        }
      } else {
      // causal (CLIP text encoder): this row sees keys j*block_kv + i <= q_row only
      const int kv_lim = p.causal ? min(kv_valid, q_row - j * p.block_kv + 1) : kv_valid;
      float mx = -INFINITY;
      for (int c = 0; c < kv_chunks; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_addr + (uint32_t)(c * 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c * 16 + i < kv_lim) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      mx *= p.scale_log2;
      const bool need = mx > m_ref + 8.0f;
      const bool need_any = __any_sync(0xffffffffu, need);
      if (j > 0) {
        mbar_wait(o_done, (uint32_t)(j - 1) & 1u);   // PV_{j-1} retired: P buffer free, O stable
        tc_fence_after();
      }
      if (need_any) {
        const float m_new = need ? mx : m_ref;
        const float alpha = ex2f(m_ref - m_new);     // m_ref = -inf on the first tile -> 0
        if (j > 0) {
          for (int c = 0; c < dchunks; ++c) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
          }
          tmem_st_wait();
        }
        l *= alpha;
        m_ref = m_new;
      }
      for (int c = 0; c < kv_chunks; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_addr + (uint32_t)(c * 16), r);
        tmem_ld_wait();
        float pv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e = ex2f(__uint_as_float(r[i]) * p.scale_log2 - m_ref);
          pv[i] = (c * 16 + i < kv_lim) ? e : 0.f;
          l += pv[i];
        }
        const uint32_t dst = p_row + (uint32_t)c * ATT_BLOCK_Q * 32u;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (0u ^ flip)), "r"(pack_bf16x2(pv[0], pv[1])),
                     "r"(pack_bf16x2(pv[2], pv[3])), "r"(pack_bf16x2(pv[4], pv[5])), "r"(pack_bf16x2(pv[6], pv[7]))
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (16u ^ flip)), "r"(pack_bf16x2(pv[8], pv[9])),
                     "r"(pack_bf16x2(pv[10], pv[11])), "r"(pack_bf16x2(pv[12], pv[13])), "r"(pack_bf16x2(pv[14], pv[15]))
                     : "memory");
      }
      }
      fence_proxy_async_smem();     // generic-proxy P writes -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 ----
    mbar_wait(o_done, (uint32_t)(n_tiles - 1) & 1u);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)q_row * p.out_ld + p.out_col0 + head * p.dp;
    for (int c = 0; c < dchunks; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
      tmem_ld_wait();
      if (q_row < p.Sq) {
        uint4 a, bq;
        a.x = pack_bf16x2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
        a.y = pack_bf16x2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
        a.z = pack_bf16x2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
        a.w = pack_bf16x2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        bq.x = pack_bf16x2(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l);
        bq.y = pack_bf16x2(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l);
        bq.z = pack_bf16x2(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l);
        bq.w = pack_bf16x2(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 16) = a;
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = bq;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// =================================================================================================
// Double-buffered variant (block_kv == KV, KV = 64 or 128): the main self-attention kernel.
// Two score buffers S[0|1] in TMEM and two probability buffers P[0|1] in shared memory decouple the
// softmax warps from the tensor core: while the softmax warps work on tile j the MMA warp runs
// O += P_{j-1} V_{j-1} and S[(j+1)&1] = Q K_{j+1}^T, so neither the MMA execution (~850 cycles per
// 128x128 tile: the A operands come from shared memory) nor the mbarrier hand-offs sit on the
// softmax critical path (measured on B200 with in-kernel clock stamps, see profiles/).
//   MMA warp     : QK_0, QK_1; for j: wait p_full[j&1] -> PV_j (commit o_done[j&1], kv_empty) -> QK_{j+2}
//   softmax warps: for j: wait s_full[j&1] (+ o_done[j&1] of tile j-2: P buffer free) -> exponentials
//                  against the running reference max -> P[j&1] -> arrive p_full[j&1]
// K/V tiles flow through a 3-stage TMA ring.
// =================================================================================================
// PT: the probabilities go back into tensor memory (over the score buffer they came from) and O += P V reads
// its A operand from TMEM (tcgen05.mma TS form): no P round trip through shared memory, no proxy fence, and the
// PV MMA no longer pays the 4 KB-per-instruction shared-memory A fetch.
template <int KV, bool PT>
__global__ void __launch_bounds__(ATT_THREADS, KV == 64 ? 2 : 1)
attn_fwd_db_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnKernelParams p) {
  constexpr int NST_MAX = 3;
  const int NST = p.kv_stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int dchunks = p.dp >> 4;
  const uint32_t q_bytes = (uint32_t)dchunks * ATT_BLOCK_Q * 32u;
  constexpr uint32_t kv_chunk_bytes = KV * 32u;
  const uint32_t kv_tile_bytes = (uint32_t)dchunks * kv_chunk_bytes;
  constexpr uint32_t p_bytes = (KV / 16) * ATT_BLOCK_Q * 32u;
  const uint32_t sQ = smem_base;
  const uint32_t sP = sQ + q_bytes;                                        // P[0], P[1]
  const uint32_t sKV = sP + (PT ? 0u : 2 * p_bytes);                        // stage s: K then V
  const uint32_t bar_base = sKV + NST * 2 * kv_tile_bytes;
  const uint32_t q_full = bar_base;
  auto s_full = [&](int i) { return bar_base + 8u + 8u * i; };
  auto p_full = [&](int i) { return bar_base + 24u + 8u * i; };
  auto o_done = [&](int i) { return bar_base + 40u + 8u * i; };
  auto kv_full = [&](int s) { return bar_base + 56u + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 56u + 8u * (NST_MAX + s); };
  const uint32_t tmem_ptr_smem = bar_base + 56u + 8u * (2 * NST_MAX);

  // Roles: warps 0..3 softmax (TMEM lane quarter = warp), warp 4 TMA producer, warp 5 MMA issuer.  The issue
  // arbiter favours the highest warp id of a scheduler, so the two single-lane "control" warps must sit above
  // the (issue-heavy) softmax warps or their MMA / TMA issue gets starved.
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;
  constexpr int W_TMA = 4, W_MMA = 5;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 128);
      mbar_init(o_done(i), 1);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == W_MMA) tmem_alloc(tmem_ptr_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tmem_O = tmem_base + 2u * KV;
  const bool tl_cta = p.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

  if (warp == W_TMA) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {
      mbar_expect_tx(q_full, q_bytes);
      for (int c = 0; c < dchunks; ++c)
        tma_load_3d(&maps.q, sQ + (uint32_t)c * ATT_BLOCK_Q * 32u, q_full, p.q_col0 + head * p.dp + c * 16,
                    qt * ATT_BLOCK_Q, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(kv_empty(st), ph ^ 1u);
      if (elect_one()) {
        const uint32_t sK = sKV + (uint32_t)st * 2 * kv_tile_bytes;
        const uint32_t sV = sK + kv_tile_bytes;
        mbar_expect_tx(kv_full(st), 2 * kv_tile_bytes);
        for (int c = 0; c < dchunks; ++c)
          tma_load_3d(&maps.k, sK + (uint32_t)c * kv_chunk_bytes, kv_full(st), p.k_col0 + head * p.dp + c * 16, j * KV, b);
        for (int c = 0; c < dchunks; ++c)
          tma_load_3d(&maps.v, sV + (uint32_t)c * kv_chunk_bytes, kv_full(st), p.v_col0 + head * p.dp + c * 16, j * KV, b);
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1u; }
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer (warp-uniform, one elected lane issues) ----------------
    const uint32_t idesc_qk = make_idesc_f16(ATT_BLOCK_Q, (uint32_t)KV, true, 0, 0);
    const uint32_t idesc_pv = make_idesc_f16(ATT_BLOCK_Q, (uint32_t)p.dp, true, 0, 1);
    const uint64_t desc_q0 = make_smem_desc(sQ, 16, 256, SWZ_32B);
    const uint64_t desc_p0 = make_smem_desc(sP, 16, 256, SWZ_32B);
    const uint64_t desc_k0 = make_smem_desc(sKV, 16, 256, SWZ_32B);
    const uint64_t desc_v0 = make_smem_desc(sKV + kv_tile_bytes, KV * 32u, 256, SWZ_32B);
    const uint32_t stage_step = (2 * kv_tile_bytes) >> 4;
    auto issue_qk = [&](int jj) {                 // S[jj & 1] = Q K_jj^T
      const int st = jj % NST;
      mbar_wait(kv_full(st), (uint32_t)(jj / NST) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = desc_k0 + (uint64_t)((uint32_t)st * stage_step);
        const uint32_t tS = tmem_base + (uint32_t)(jj & 1) * KV;
        for (int c = 0; c < dchunks; ++c)
          umma_f16_ss(tS, desc_q0 + (uint64_t)(c * (ATT_BLOCK_Q * 32 / 16)), dk + (uint64_t)(c * (int)(kv_chunk_bytes >> 4)),
                      idesc_qk, c != 0);
        umma_commit(s_full(jj & 1));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    if (n_tiles > 1) issue_qk(1);
    for (int j = 0; j < n_tiles; ++j) {
      const int bi = j & 1;
      const int st = j % NST;
      mbar_wait(p_full(bi), (uint32_t)(j >> 1) & 1u);      // P[bi] written, S[bi] consumed
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc_v0 + (uint64_t)((uint32_t)st * stage_step);
        const uint64_t dp = desc_p0 + (uint64_t)((uint32_t)bi * (p_bytes >> 4));
#pragma unroll
        for (int k = 0; k < KV / 16; ++k) {
          if constexpr (PT)
            umma_f16_ts(tmem_O, tmem_base + (uint32_t)bi * KV + (uint32_t)(8 * k), dv + (uint64_t)(k * (512 / 16)), idesc_pv, (j | k) != 0);
          else
            umma_f16_ss(tmem_O, dp + (uint64_t)(k * (ATT_BLOCK_Q * 32 / 16)), dv + (uint64_t)(k * (512 / 16)), idesc_pv, (j | k) != 0);
        }
        umma_commit(o_done(bi));
        umma_commit(kv_empty(st));
      }
      __syncwarp();
      if (j + 2 < n_tiles) issue_qk(j + 2);
    }
  } else {
    // ---------------- softmax / correction / epilogue warps ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int q_row = qt * ATT_BLOCK_Q + row;
    float m_ref = -INFINITY, l = 0.f;
    const uint32_t flip = (uint32_t)((row >> 2) & 1) << 4;     // 32B-swizzle: 16B halves swap on rows 4..7 of 8
    const bool tls = tl_cta && warp == 0 && lane == 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int bi = j & 1;
      const uint32_t tS = tmem_base + (uint32_t)bi * KV + lane_addr;
      const uint32_t p_row = sP + (uint32_t)bi * p_bytes + (uint32_t)row * 32u;
      if (tls && j < 60) p.timeline[j * 8 + 0] = clock64();
      mbar_wait(s_full(bi), (uint32_t)(j >> 1) & 1u);
      if (!PT && j >= 2) mbar_wait(o_done(bi), (uint32_t)((j >> 1) - 1) & 1u);     // PV_{j-2} retired: P[bi] is free
      tc_fence_after();
      if (tls && j < 60) p.timeline[j * 8 + 1] = clock64();
      const int kv_valid = min(KV, p.Skv - j * KV);
      uint32_t sreg[KV];
      uint32_t pw[PT ? KV / 2 : 1];            // packed bf16 probabilities (TMEM path)
      bool careful = (kv_valid != KV) || (j == 0);
      float mx = -INFINITY;
      if (!careful) {
        // optimistic pass: exponentials against the running reference max, TMEM load of the next
        // 32-column chunk in flight meanwhile; 8 independent max / sum chains for ILP
        float m8[8], l8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { m8[i] = -INFINITY; l8[i] = 0.f; }
        tmem_ld_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sreg[0]));
#pragma unroll
        for (int c = 0; c < KV / 32; ++c) {
          tmem_ld_wait();
          if (c + 1 < KV / 32)
            tmem_ld_32x32b_x32(tS + (uint32_t)((c + 1) * 32), *reinterpret_cast<uint32_t(*)[32]>(&sreg[(c + 1) * 32]));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float pv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float sv = __uint_as_float(sreg[c * 32 + h * 16 + i]);
              m8[i & 7] = fmaxf(m8[i & 7], sv);
              pv[i] = ex2f(fmaf(sv, p.scale_log2, -m_ref));
              l8[i & 7] += pv[i];
            }
            if constexpr (PT) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pw[(c * 2 + h) * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
            } else {
              const uint32_t dst = p_row + (uint32_t)(c * 2 + h) * ATT_BLOCK_Q * 32u;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (0u ^ flip)), "r"(pack_bf16x2(pv[0], pv[1])),
                           "r"(pack_bf16x2(pv[2], pv[3])), "r"(pack_bf16x2(pv[4], pv[5])), "r"(pack_bf16x2(pv[6], pv[7])) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (16u ^ flip)), "r"(pack_bf16x2(pv[8], pv[9])),
                           "r"(pack_bf16x2(pv[10], pv[11])), "r"(pack_bf16x2(pv[12], pv[13])), "r"(pack_bf16x2(pv[14], pv[15])) : "memory");
            }
          }
        }
        const float tsum = ((l8[0] + l8[1]) + (l8[2] + l8[3])) + ((l8[4] + l8[5]) + (l8[6] + l8[7]));
        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7]))) * p.scale_log2;
        careful = __any_sync(0xffffffffu, mx > m_ref + 8.0f);
        if (!careful) l += tsum;
      } else {
#pragma unroll
        for (int c = 0; c < KV / 32; ++c)
          tmem_ld_32x32b_x32(tS + (uint32_t)(c * 32), *reinterpret_cast<uint32_t(*)[32]>(&sreg[c * 32]));
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < KV; ++i)
          if (i < kv_valid) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sreg[i]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      }
      if (careful) {
        // new reference max: rescale O (needs every earlier PV retired) and l, redo the tile from registers
        const bool need = mx > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_ref;
          const float alpha = ex2f(m_ref - m_new);     // m_ref = -inf on the first tile -> 0
          if (j > 0) {
            mbar_wait(o_done((j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
            tc_fence_after();
            for (int c = 0; c < dchunks; ++c) {
              uint32_t r[16];
              tmem_ld_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
            }
            tmem_st_wait();
          }
          l *= alpha;
          m_ref = m_new;
        }
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < KV / 16; ++c) {
          float pv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e = ex2f(fmaf(__uint_as_float(sreg[c * 16 + i]), p.scale_log2, -m_ref));
            pv[i] = (c * 16 + i < kv_valid) ? e : 0.f;
            l4[i & 3] += pv[i];
          }
          if constexpr (PT) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pw[c * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
          } else {
            const uint32_t dst = p_row + (uint32_t)c * ATT_BLOCK_Q * 32u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (0u ^ flip)), "r"(pack_bf16x2(pv[0], pv[1])),
                         "r"(pack_bf16x2(pv[2], pv[3])), "r"(pack_bf16x2(pv[4], pv[5])), "r"(pack_bf16x2(pv[6], pv[7])) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (16u ^ flip)), "r"(pack_bf16x2(pv[8], pv[9])),
                         "r"(pack_bf16x2(pv[10], pv[11])), "r"(pack_bf16x2(pv[12], pv[13])), "r"(pack_bf16x2(pv[14], pv[15])) : "memory");
          }
        }
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      }
      if (tls && j < 60) p.timeline[j * 8 + 2] = clock64();
      if constexpr (PT) {
        // P (bf16, two per column) over the first KV/2 columns of the score buffer it was computed from
#pragma unroll
        for (int c = 0; c < KV / 64; ++c)
          tmem_st_32x32b_x32(tS + (uint32_t)(c * 32), *reinterpret_cast<uint32_t(*)[32]>(&pw[c * 32]));
        tmem_st_wait();
      } else {
        fence_proxy_async_smem();   // generic-proxy P writes -> visible to the tensor core's async proxy
      }
      tc_fence_before();
      mbar_arrive(p_full(bi));
      if (tls && j < 60) p.timeline[j * 8 + 3] = clock64();
    }
    // ---- epilogue: O / l -> bf16 ----
    mbar_wait(o_done((n_tiles - 1) & 1), (uint32_t)((n_tiles - 1) >> 1) & 1u);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)q_row * p.out_ld + p.out_col0 + head * p.dp;
    for (int c = 0; c < dchunks; ++c) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem_O + lane_addr + (uint32_t)(c * 16), r);
      tmem_ld_wait();
      if (q_row < p.Sq) {
        uint4 a, bq;
        a.x = pack_bf16x2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
        a.y = pack_bf16x2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
        a.z = pack_bf16x2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
        a.w = pack_bf16x2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        bq.x = pack_bf16x2(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l);
        bq.y = pack_bf16x2(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l);
        bq.z = pack_bf16x2(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l);
        bq.w = pack_bf16x2(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 16) = a;
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = bq;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// =================================================================================================
// Short-KV kernel: the cross-attention layers (S_kv = 77 / 85 text tokens: ONE key tile, block_kv = S_kv rounded up
// to 16).  The generic single-buffer kernel above is latency-bound here (ncu: tensor 7 %, DRAM 13 %): every CTA pays
// TMEM allocation, barrier init, descriptor fetches and the K/V load for 128 queries' worth of work.  This kernel
// keeps K/V resident and streams `q_tiles` consecutive 128-row query tiles of one (batch, head) through a
// software pipeline:
//   TMA warp     : K/V once; Q tiles through a 2-slot ring
//   MMA warp     : S[i&1] = Q_i K^T (issued one tile ahead);  O[i&1] = P_i V   (P from tensor memory, TS form)
//   softmax warps: EIGHT — warps w and w + 4 share TMEM lane quarter w & 3 and alternate query tiles (warp h owns slot h of the
//                  S / O double buffers: tiles i = h, h + 2, ...), so each scheduler interleaves four of these latency-bound
//                  chains (two CTAs per SM) instead of two: 0.527 -> 0.458 ms at 256 rows x 64x64, 0.261 -> 0.229 at 32x32
//                  (profiles/r02_cross_attention_eight_warps.log).  Tile i: the epilogue of the warp's previous tile i - 2
//                  (O[h] / l -> bf16; its P V retired a tile ago), then two passes over S[h] in TMEM (max, then
//                  exponentials; P over the score columns).  Same arithmetic per row as the four-warp form: same bits.
//                  (Two 16-column TMEM loads per wait instead of one: slower, 0.582 ms, like the one-pass form below.)
// TMEM columns: S[0], S[1] (block_kv each), O[0], O[1] (dp each).
// =================================================================================================
// (Holding the 80-column score row in registers — one TMEM pass instead of two — measured 4 % slower: 0.284 vs 0.272 ms
// at B=128; the two-pass form is kept.)
constexpr int ATT_SHORT_THREADS = 320;      // 8 softmax warps + TMA producer + MMA issuer
__global__ void __launch_bounds__(ATT_SHORT_THREADS, 2)
attn_short_kv_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int dchunks = p.dp >> 4;
  const int bkv = p.block_kv;
  const uint32_t q_bytes = (uint32_t)dchunks * ATT_BLOCK_Q * 32u;
  const uint32_t kv_chunk_bytes = (uint32_t)bkv * 32u;
  const uint32_t kv_tile_bytes = (uint32_t)dchunks * kv_chunk_bytes;
  const uint32_t sQ = smem_base;                          // Q ring: 2 slots
  const uint32_t sK = sQ + 2 * q_bytes;
  const uint32_t sV = sK + kv_tile_bytes;
  const uint32_t bar_base = sV + kv_tile_bytes;
  const uint32_t kv_full = bar_base;
  auto q_full = [&](int i) { return bar_base + 8u + 8u * i; };
  auto q_empty = [&](int i) { return bar_base + 24u + 8u * i; };
  auto s_full = [&](int i) { return bar_base + 40u + 8u * i; };
  auto p_full = [&](int i) { return bar_base + 56u + 8u * i; };
  auto o_full = [&](int i) { return bar_base + 72u + 8u * i; };
  auto o_empty = [&](int i) { return bar_base + 88u + 8u * i; };
  const uint32_t tmem_ptr_smem = bar_base + 104u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int n_qtiles = (p.Sq + ATT_BLOCK_Q - 1) / ATT_BLOCK_Q;
  const int qt0 = blockIdx.x * p.q_tiles;
  const int nt = min(p.q_tiles, n_qtiles - qt0);          // query tiles of this CTA (>= 1)
  constexpr int W_TMA = 8, W_MMA = 9;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full(i), 1);
      mbar_init(q_empty(i), 1);
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 128);
      mbar_init(o_full(i), 1);
      mbar_init(o_empty(i), 128);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == W_MMA) tmem_alloc(tmem_ptr_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tmem_O0 = tmem_base + 2u * (uint32_t)bkv;

  if (warp == W_TMA) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * kv_tile_bytes);
      for (int c = 0; c < dchunks; ++c)
        tma_load_3d(&maps.k, sK + (uint32_t)c * kv_chunk_bytes, kv_full, p.k_col0 + head * p.dp + c * 16, 0, b);
      for (int c = 0; c < dchunks; ++c)
        tma_load_3d(&maps.v, sV + (uint32_t)c * kv_chunk_bytes, kv_full, p.v_col0 + head * p.dp + c * 16, 0, b);
    }
    __syncwarp();
    for (int i = 0; i < nt; ++i) {
      const int sl = i & 1;
      mbar_wait(q_empty(sl), ((uint32_t)(i >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(q_full(sl), q_bytes);
        for (int c = 0; c < dchunks; ++c)
          tma_load_3d(&maps.q, sQ + (uint32_t)sl * q_bytes + (uint32_t)c * ATT_BLOCK_Q * 32u, q_full(sl),
                      p.q_col0 + head * p.dp + c * 16, (qt0 + i) * ATT_BLOCK_Q, b);
      }
      __syncwarp();
    }
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc_qk = make_idesc_f16(ATT_BLOCK_Q, (uint32_t)bkv, true, 0, 0);
    const uint32_t idesc_pv = make_idesc_f16(ATT_BLOCK_Q, (uint32_t)p.dp, true, 0, 1);
    const uint64_t desc_q0 = make_smem_desc(sQ, 16, 256, SWZ_32B);
    const uint64_t desc_k0 = make_smem_desc(sK, 16, 256, SWZ_32B);
    const uint64_t desc_v0 = make_smem_desc(sV, kv_chunk_bytes, 256, SWZ_32B);
    const int pv_steps = bkv >> 4;
    auto issue_qk = [&](int i) {                  // S[i & 1] = Q_i K^T
      const int sl = i & 1;
      mbar_wait(q_full(sl), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq = desc_q0 + (uint64_t)((uint32_t)sl * (q_bytes >> 4));
        for (int c = 0; c < dchunks; ++c)
          umma_f16_ss(tmem_base + (uint32_t)(sl * bkv), dq + (uint64_t)(c * (ATT_BLOCK_Q * 32 / 16)),
                      desc_k0 + (uint64_t)((uint32_t)c * (kv_chunk_bytes >> 4)), idesc_qk, c != 0);
        umma_commit(s_full(sl));
        umma_commit(q_empty(sl));
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    issue_qk(0);
    for (int i = 0; i < nt; ++i) {
      const int sl = i & 1;
      if (i + 1 < nt) issue_qk(i + 1);            // S[(i+1)&1] was last read by PV_{i-1}, issued earlier (in-order pipe)
      mbar_wait(p_full(sl), (uint32_t)(i >> 1) & 1u);
      if (i >= 2) mbar_wait(o_empty(sl), (uint32_t)((i >> 1) - 1) & 1u);   // epilogue of tile i-2 has drained O[sl]
      tc_fence_after();
      if (elect_one()) {
        for (int k = 0; k < pv_steps; ++k)
          umma_f16_ts(tmem_O0 + (uint32_t)(sl * p.dp), tmem_base + (uint32_t)(sl * bkv) + (uint32_t)(8 * k),
                      desc_v0 + (uint64_t)(k * (512 / 16)), idesc_pv, k != 0);
        umma_commit(o_full(sl));
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax + epilogue warps ----------------
    const int quarter = warp & 3;
    const int half = warp >> 2;                   // this warp's slot of the S / O double buffers: tiles half, half + 2, ...
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int kv_chunks = bkv >> 4;
    float l_prev = 1.f;
    auto epilogue = [&](int i, float l) {         // O[i & 1] / l -> bf16 rows of query tile qt0 + i
      const int sl = i & 1;
      mbar_wait(o_full(sl), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int q_row = (qt0 + i) * ATT_BLOCK_Q + row;
      __nv_bfloat16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)q_row * p.out_ld + p.out_col0 + head * p.dp;
      for (int c = 0; c < dchunks; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_O0 + (uint32_t)(sl * p.dp) + lane_addr + (uint32_t)(c * 16), r);
        tmem_ld_wait();
        if (q_row < p.Sq) {
          uint32_t w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = pack_bf16x2(__uint_as_float(r[2 * k]) * inv_l, __uint_as_float(r[2 * k + 1]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc_fence_before();
      mbar_arrive(o_empty(sl));
    };
    int i_last = -1;
    for (int i = half; i < nt; i += 2) {
      const int sl = i & 1;
      const uint32_t tS = tmem_base + (uint32_t)(sl * bkv) + lane_addr;
      if (i >= 2) epilogue(i - 2, l_prev);        // this warp's previous tile: frees O[sl] for P V of tile i
      i_last = i;
      mbar_wait(s_full(sl), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      float l = 0.f;
      {
        // only the last chunk can hold columns >= Skv: the others run without the per-element mask
        const int nfull = p.Skv >> 4;
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int c = 0; c < kv_chunks; ++c) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tS + (uint32_t)(c * 16), r);
          tmem_ld_wait();
          if (c < nfull) {
#pragma unroll
            for (int k = 0; k < 16; ++k) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (c * 16 + k < p.Skv) m4[k & 3] = fmaxf(m4[k & 3], __uint_as_float(r[k]));
          }
        }
        const float m2 = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < kv_chunks; ++c) {
          uint32_t r[16], w[8];
          tmem_ld_32x32b_x16(tS + (uint32_t)(c * 16), r);
          tmem_ld_wait();
          float pv[16];
          if (c < nfull) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              pv[k] = ex2f(fmaf(__uint_as_float(r[k]), p.scale_log2, -m2));
              l4[k & 3] += pv[k];
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float e = ex2f(fmaf(__uint_as_float(r[k]), p.scale_log2, -m2));
              pv[k] = (c * 16 + k < p.Skv) ? e : 0.f;
              l4[k & 3] += pv[k];
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = pack_bf16x2(pv[2 * k], pv[2 * k + 1]);
          // P chunk c (16 bf16 = 8 columns) over score columns [8c, 8c+8): already consumed (8c + 8 <= 16c for c >= 1,
          // and chunk 0 sits in registers)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tS + (uint32_t)(c * 8)),
                       "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(sl));
      l_prev = l;
    }
    if (i_last >= 0) epilogue(i_last, l_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace dfb

#include "dfb_attn_sa.cuh"
#include "dfb_attn_sa8.cuh"

using namespace dfb;

// Share of the exponentials attn_fwd_sa_kernel computes on the FMA pipe (of every 16).  DFB_ATTN_POLY = 0 | 2 | 4 overrides
// the built-in choice (measured on B200, profiles/r02_attention_sa_*.log).
static int attn_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int attn_poly_default(bool ones) {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("DFB_ATTN_POLY");
    env = e ? atoi(e) : -1;
    if (env != -1 && env != 0 && env != 2 && env != 4) env = -1;
  }
  if (env >= 0) return env;
  return ones ? DFB_ATTN_POLY_ONES_DEFAULT : DFB_ATTN_POLY_DEFAULT;
}

extern "C" size_t dfb_attention_ws_bytes(int B, int heads, int Sq) {
  return (size_t)4 * (size_t)B * (size_t)heads * (size_t)((Sq + ATT_BLOCK_Q - 1) / ATT_BLOCK_Q);
}

extern "C" int dfb_attention(const dfb_attn_params* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DFB_REQUIRE(a && a->q && a->k && a->v && a->out, "dfb_attention: null buffer");
  DFB_REQUIRE(a->B > 0 && a->heads > 0 && a->Sq > 0 && a->Skv > 0, "dfb_attention: empty problem");
  DFB_REQUIRE(a->dp % 16 == 0 && a->dp >= 16 && a->dp <= 256, "dfb_attention: padded head dim must be a multiple of 16 in [16,256]");
  DFB_REQUIRE(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0 && a->out_ld % 8 == 0, "dfb_attention: row pitches must be multiples of 8");
  DFB_REQUIRE(a->q_col0 % 8 == 0 && a->k_col0 % 8 == 0 && a->v_col0 % 8 == 0 && a->out_col0 % 8 == 0, "dfb_attention: column offsets must be multiples of 8");
  DFB_REQUIRE(((uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v | (uintptr_t)a->out) % 16 == 0, "dfb_attention: buffers must be 16B aligned");

  AttnKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  int bkv = a->block_kv;
  if (bkv <= 0) {
    const int skv16 = (a->Skv + 15) / 16 * 16;
    const int pref = a->dp <= 48 ? 128 : 64;
    bkv = skv16 < pref ? skv16 : pref;
    if (skv16 <= 128 && a->dp <= 80 && skv16 > pref) bkv = skv16;   // short KV (cross-attention): one tile
  }
  DFB_REQUIRE(bkv % 16 == 0 && bkv >= 16 && bkv <= 128, "dfb_attention: block_kv must be a multiple of 16 in [16,128]");
  // kernel family: the double-buffered kernel for multi-tile KV (self-attention), the single-buffer kernel for
  // one-tile / odd-sized KV (cross-attention, tiny shapes).  dbg_flags bit3 forces the single-buffer family.
  DFB_REQUIRE(a->causal == 0 || a->Sq == a->Skv, "dfb_attention: causal needs Sq == Skv");
  const bool causal = a->causal != 0;                              // generic single-buffer kernel only
  bool use_db = !causal && !(a->dbg_flags & 8) && (bkv == 64 || bkv == 128) && a->Skv > bkv;
  if (use_db && a->block_kv <= 0) bkv = 64;                       // default tile of the double-buffered kernel
  if (use_db && bkv == 128 && 2 * 128 + a->dp > 512) use_db = false;
  kp.Sq = a->Sq; kp.Skv = a->Skv; kp.dp = a->dp; kp.block_kv = bkv;
  kp.n_kv_tiles = (a->Skv + bkv - 1) / bkv;
  kp.q_col0 = a->q_col0; kp.k_col0 = a->k_col0; kp.v_col0 = a->v_col0;
  kp.scale_log2 = a->scale * 1.4426950408889634f;
  kp.out = (__nv_bfloat16*)a->out;
  kp.out_ld = a->out_ld; kp.out_col0 = a->out_col0;
  kp.out_batch_stride = (long long)a->Sq * a->out_ld;
  // short-KV kernel (one key tile, K/V resident, query tiles streamed): cross-attention.  dbg_flags bit6 disables it.
  const bool use_short = !causal && !use_db && a->Skv <= bkv && 2 * (bkv + a->dp) <= 512 && a->Sq > ATT_BLOCK_Q && (a->dbg_flags & (8 | 64)) == 0;
  // attn_fwd_sa_kernel (dfb_attn_sa.cuh): the double-buffered kernel with P in tensor memory, reworked for the long
  // self-attention layers — softmax denominator from the P V MMA (ones column of V, `ones_col`), part of the exponentials
  // on the FMA pipe, TMEM loads prefetched across tiles.  dbg_flags bit12 keeps the round-1 kernel (A/B runs);
  // bits 13-14 choose the polynomial share: 1 -> none, 2 -> 2 of 16, 3 -> 4 of 16, 0 -> default (DFB_ATTN_POLY or built-in).
  // (DFB_ATTN_R01=1: round 1's double-buffered kernel everywhere — the A/B partner in profiles/r02_attention_step_ab.json)
  const bool use_sa = use_db && bkv == 64 && (a->dbg_flags & (16 | 4096)) == 0 && attn_env_int("DFB_ATTN_R01", 0) == 0;
  const bool sa_ones = use_sa && a->ones_col > 0;
  DFB_REQUIRE(a->ones_col >= 0 && a->ones_col <= a->dp, "dfb_attention: ones_col must be 0 (none) or 1 + a column of the padded head");
  int sa_poly = 0;
  if (use_sa) {
    const int sel = (a->dbg_flags >> 13) & 3;
    sa_poly = sel == 1 ? 0 : sel == 2 ? 2 : sel == 3 ? 4 : attn_poly_default(sa_ones);
  }
  kp.l_col = sa_ones ? a->ones_col - 1 : 0;
  // attn_fwd_sa8_kernel (dfb_attn_sa8.cuh): eight softmax warps per CTA, static reference maximum, overflow flagged into the
  // caller's workspace and redone by attn_fwd_sa_kernel<.., REDO>.  Needs the ones column, room for three score buffers
  // (dp <= 64) and the workspace; dbg_flags bit15 / DFB_ATTN_SA8=0 keep the 4-warp kernel.
  const bool use_sa8 = sa_ones && 3 * bkv + a->dp <= 256 && a->workspace != nullptr && (a->dbg_flags & 32768) == 0 &&
                       attn_env_int("DFB_ATTN_SA8", 1) != 0;
  kp.redo_flags = (int*)a->workspace;
  kp.n_q_tiles = (a->Sq + ATT_BLOCK_Q - 1) / ATT_BLOCK_Q;
  kp.dbg_delay = (a->dbg_flags >> 16) & 15;
  // three score buffers where they still leave room for two CTAs per SM (3 * 64 + dp <= 256 columns: dp <= 64)
  kp.s_ring = (use_sa && 3 * bkv + a->dp <= 256 && (use_sa8 || attn_env_int("DFB_ATTN_S_RING", 3) >= 3)) ? 3 : 2;
  uint32_t need_cols = use_short ? (uint32_t)(2 * (bkv + a->dp)) : use_sa ? (uint32_t)(kp.s_ring * bkv + a->dp)
                                                                         : (uint32_t)((use_db ? 2 : 1) * bkv + a->dp), cols = 32;
  while (cols < need_cols) cols <<= 1;
  DFB_REQUIRE(cols <= 512, "dfb_attention: block_kv + dp exceeds TMEM");
  kp.tmem_cols = cols;
  kp.v_lbo = a->dbg_v_lbo > 0 ? (uint32_t)a->dbg_v_lbo : (uint32_t)bkv * 32u;
  kp.v_sbo = a->dbg_v_sbo > 0 ? (uint32_t)a->dbg_v_sbo : 256u;
  kp.timeline = (long long*)a->dbg_timeline;
  kp.tl_second = num_sms();
  kp.causal = causal ? 1 : 0;

  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    uint64_t dims[3] = {(uint64_t)a->q_ld, (uint64_t)a->Sq, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->q_ld * 2, (uint64_t)a->q_ld * 2 * a->Sq};
    uint32_t box[3] = {16, ATT_BLOCK_Q, 1};
    int rc = make_tmap(&maps.q, a->q, 2, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc != DFB_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->k_ld, (uint64_t)a->Skv, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->k_ld * 2, (uint64_t)a->k_ld * 2 * a->Skv};
    uint32_t box[3] = {16, (uint32_t)bkv, 1};
    int rc = make_tmap(&maps.k, a->k, 2, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc != DFB_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->v_ld, (uint64_t)a->Skv, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->v_ld * 2, (uint64_t)a->v_ld * 2 * a->Skv};
    uint32_t box[3] = {16, (uint32_t)bkv, 1};
    int rc = make_tmap(&maps.v, a->v, 2, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc != DFB_OK) return rc;
  }

  const int dch = a->dp / 16;
  const bool p_tmem = use_db && (a->dbg_flags & 16) == 0;     // P through tensor memory (TS MMA); dbg bit4: smem P instead
  size_t smem;
  if (use_db) {
    // 3 K/V stages when two CTAs still fit per SM with them, else 2
    const size_t fixed = 1024 + (size_t)dch * ATT_BLOCK_Q * 32 + (p_tmem ? 0 : 2 * (size_t)(bkv / 16) * ATT_BLOCK_Q * 32) + 256;
    const size_t stage = (size_t)2 * dch * bkv * 32;
    kp.kv_stages = (fixed + 3 * stage + 1024 <= (size_t)113 * 1024 || fixed + 2 * stage + 1024 > (size_t)113 * 1024) ? 3 : 2;
    if (use_sa) {
      // deeper ring: the TMA of tile j + NST - 1 is issued when PV_{j-1} retires, i.e. NST - 2 tile periods before QK needs it
      int want = attn_env_int("DFB_ATTN_KV_STAGES", 6);
      if (want > 8) want = 8;
      while (want > kp.kv_stages && fixed + (size_t)want * stage + 1024 > (size_t)113 * 1024) --want;
      if (want > kp.kv_stages) kp.kv_stages = want;
    }
    smem = fixed + (size_t)kp.kv_stages * stage + (use_sa8 ? 1536 : 0);
  } else {
    smem = 1024 + (size_t)dch * ATT_BLOCK_Q * 32 + (size_t)(bkv / 16) * ATT_BLOCK_Q * 32 + (size_t)ATT_STAGES * 2 * dch * bkv * 32 + 128;
  }
  const int n_qtiles = (a->Sq + ATT_BLOCK_Q - 1) / ATT_BLOCK_Q;
  if (use_short) {
    smem = 1024 + 2 * (size_t)dch * ATT_BLOCK_Q * 32 + (size_t)2 * dch * bkv * 32 + 256;
    // query tiles per CTA: amortise the per-CTA set-up, but keep >= ~4 CTAs per SM slot for load balance
    int qtiles = 8;
    while (qtiles > 1 && (long long)((n_qtiles + qtiles - 1) / qtiles) * a->heads * a->B < 8LL * num_sms()) qtiles >>= 1;
    if ((a->dbg_flags >> 8) & 0xF) qtiles = (a->dbg_flags >> 8) & 0xF;      // test hook: forced tile count per CTA
    kp.q_tiles = qtiles;
  }
  if ((a->dbg_flags & 2) && smem < 120 * 1024) smem = 120 * 1024;      // tuning hook: force one CTA per SM
  DFB_REQUIRE(smem <= 227 * 1024, "dfb_attention: tile configuration exceeds shared memory");
  static bool attr_set[64] = {false};
  int dev = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    const int mx = 227 * 1024;
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_db_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_db_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_db_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_db_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_short_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    DFB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_sa8_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    attr_set[dev] = true;
  }
  dim3 grid((a->Sq + ATT_BLOCK_Q - 1) / ATT_BLOCK_Q, a->heads, a->B);
  if (use_short) {
    dim3 gshort((n_qtiles + kp.q_tiles - 1) / kp.q_tiles, a->heads, a->B);
    attn_short_kv_kernel<<<gshort, ATT_SHORT_THREADS, smem, stream>>>(maps, kp);
  } else if (use_sa8) {
    // share of the exponentials on the FMA pipe: dbg_flags bits 13-14 (1 none, 2 = 2/16, 3 = 4/16), DFB_ATTN_SA8_POLY = 0|2|4|8, or built-in
    const bool tiles8 = attn_env_int("DFB_ATTN_SA8_TILES", DFB_ATTN_SA8_TILES_DEFAULT) != 0;
    // (the polynomial share pays only with the column split: 2.867 vs 2.905 ms; with the tile split 2.821 vs 2.783)
    int poly8 = attn_env_int("DFB_ATTN_SA8_POLY", tiles8 ? 0 : DFB_ATTN_SA8_POLY_DEFAULT);
    const int sel8 = (a->dbg_flags >> 13) & 3;
    if (sel8) poly8 = sel8 == 1 ? 0 : sel8 == 2 ? 2 : 4;
    // DFB_ATTN_SA8_TILES = 1 (default): the warps of a lane quarter alternate TILES instead of splitting every tile's columns
    if (tiles8) {
      if (poly8 == 4) attn_fwd_sa8_kernel<4, true><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
      else if (poly8 == 2) attn_fwd_sa8_kernel<2, true><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
      else attn_fwd_sa8_kernel<0, true><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
    } else if (poly8 == 8) attn_fwd_sa8_kernel<8><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
    else if (poly8 == 4) attn_fwd_sa8_kernel<4><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
    else if (poly8 == 2) attn_fwd_sa8_kernel<2><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
    else attn_fwd_sa8_kernel<0><<<grid, ATT_SA8_THREADS, smem, stream>>>(maps, kp);
    DFB_CHECK_CUDA(cudaGetLastError());
    // flagged tiles only (normally none): one CTA per (batch, head) scans that head's flags
    attn_fwd_sa_kernel<true, 0, true><<<dim3(1, grid.y, grid.z), ATT_THREADS, smem, stream>>>(maps, kp);
  } else if (use_sa) {
#define DFB_SA_LAUNCH(O_, P_) attn_fwd_sa_kernel<O_, P_><<<grid, ATT_THREADS, smem, stream>>>(maps, kp)
    if (sa_ones) { if (sa_poly == 4) DFB_SA_LAUNCH(true, 4); else if (sa_poly == 2) DFB_SA_LAUNCH(true, 2); else DFB_SA_LAUNCH(true, 0); }
    else { if (sa_poly == 4) DFB_SA_LAUNCH(false, 4); else if (sa_poly == 2) DFB_SA_LAUNCH(false, 2); else DFB_SA_LAUNCH(false, 0); }
#undef DFB_SA_LAUNCH
  }
  else if (use_db && bkv == 64 && p_tmem)
    attn_fwd_db_kernel<64, true><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  else if (use_db && bkv == 64)
    attn_fwd_db_kernel<64, false><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  else if (use_db && p_tmem)
    attn_fwd_db_kernel<128, true><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  else if (use_db)
    attn_fwd_db_kernel<128, false><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  else if (bkv == 128 && !causal)
    attn_fwd_kernel<128><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  else if (bkv == 64 && !causal)
    attn_fwd_kernel<64><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  else
    attn_fwd_kernel<0><<<grid, ATT_THREADS, smem, stream>>>(maps, kp);
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}
