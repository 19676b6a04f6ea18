// attn_fwd_sa8_kernel — the long-sequence self-attention kernel with EIGHT softmax warps per CTA (two CTAs per SM: four
// softmax warps per scheduler).  Included by dfb_attn.cu.
//
// Why (round-2 measurements, profiles/r02_attention_*.log): on sm_100a an ex2 occupies its warp's issue slot for ~8 cycles
// (16 lanes / clk / SM), and inside ONE warp that time does not overlap the warp's own FMA-pipe work nor its per-tile
// fixed latencies (barrier wait, TMEM load, P store, fence, arrive): a softmax warp alone needs 717 cycles of "math" for 64
// columns (512 of them MUFU) plus ~340 of latency.  Overlap only happens ACROSS warps.  With one softmax warp per scheduler
// per CTA (attn_fwd_sa_kernel, two CTAs per SM) a scheduler has two warps to interleave and reaches 81 % of the MUFU bound
// (1260 cycles per 128 x 128 score tile against 1024).  Here each 128-row tile is worked by eight warps — warp w and w + 4
// share TMEM lane quarter w & 3 and take columns [0, 32) and [32, 64) of every 64-column score tile — so a scheduler has
// four softmax warps (32 exponentials per thread per tile, ~100 registers).
//
// What makes the split cheap:
//   * ONES (required): the softmax denominator is accumulated by the P V MMA in O[:, l_col] (V holds 1.0 there), so the two
//     halves of a row need no sum exchange.
//   * STATIC reference maximum: m_ref of a row is the maximum of its first score tile (exchanged once between the two warps
//     through shared memory) and is never moved afterwards — P is bf16 and O fp32, both with 8 exponent bits, so
//     probabilities up to 2^100 relative to m_ref lose nothing (relative precision is what matters; the normalisation by
//     O[:, l_col] cancels the common factor).  No per-tile agreement between the two warps, no O rescale in the loop.
//   * Overflow (a score more than 100 log2-units above the first tile's maximum — never seen on attention logits) is
//     DETECTED, not handled: the CTA raises its flag in `redo_flags` and the host launches attn_fwd_sa_kernel<.., REDO> right
//     behind, whose CTAs exit at once unless flagged (exact lazy-rescale path).  The flag is written for every CTA, so the
//     buffer needs no initialisation.
//   * P goes back into tensor memory over the score columns its OWN warp loaded: columns [0, 16) (keys 0..31) and
//     [32, 48) (keys 32..63) of the score buffer; the TS-form P V MMA takes its four K = 16 steps from there.
#pragma once

namespace dfb {

constexpr int ATT_SA8_THREADS = 320;      // 8 softmax warps + TMA producer + MMA issuer

// POLY: of every 16 exponentials of the full-tile path, POLY are computed on the FMA pipe (ex2_poly3, dfb_attn_sa.cuh).  With four
// softmax warps per scheduler the loop is MUFU-queue-bound (ncu: mio_throttle is its top stall), which is the regime where
// moving work to the idle FMA pipe can pay (it did not with one or two warps per scheduler).
// TILES: how the two warps of a lane quarter divide the work.  false: the two 32-column halves of EVERY score tile (lock-step:
// both wait for the same S, both arrive on the same P).  true: ALTERNATE TILES — warp h takes all 64 columns (two 32-column
// passes) of tiles j = h (mod 2).  The two warps of a quarter sit on the same scheduler (TMEM lane quarter = warp id mod 4 =
// scheduler), so in the column split a scheduler sees only two independent streams (one per CTA) of two lock-stepped warps each;
// with the tile split it sees four streams that are a tile apart by construction, with half as many barrier operations per score.
// The static reference maximum is what allows it: consecutive tiles of a row are handled by different warps with no hand-over.
// (Three warps per lane quarter, one per buffer of the score ring, measured no faster: 2.85 vs 2.78 ms — profiles/r02_attention_sa8_tile_split.log.)
template <int POLY, bool TILES = false>
__global__ void __launch_bounds__(ATT_SA8_THREADS, 2)
attn_fwd_sa8_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnKernelParams p) {
  constexpr int KV = ATT_SA_KV;
  constexpr int NST_MAX = 8;
  constexpr int NS = 3;                    // score / probability ring (3 * 64 + dp <= 256 columns)
  const int NST = p.kv_stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int dchunks = p.dp >> 4;
  const uint32_t q_bytes = (uint32_t)dchunks * ATT_BLOCK_Q * 32u;
  constexpr uint32_t kv_chunk_bytes = KV * 32u;
  const uint32_t kv_tile_bytes = (uint32_t)dchunks * kv_chunk_bytes;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + q_bytes;                                        // stage s: K then V
  const uint32_t bar_base = sKV + NST * 2 * kv_tile_bytes;
  const uint32_t q_full = bar_base;
  auto s_full = [&](int i) { return bar_base + 8u + 8u * i; };
  auto p_full = [&](int i) { return bar_base + 8u + 8u * (NS + i); };
  auto o_done = [&](int i) { return bar_base + 8u + 8u * (2 * NS + i); };
  auto kv_full = [&](int s) { return bar_base + 8u + 8u * (3 * NS + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u + 8u * (3 * NS + NST_MAX + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u + 8u * (3 * NS + 2 * NST_MAX);
  const uint32_t flag_smem = tmem_ptr_smem + 4u;
  const uint32_t mx_smem = tmem_ptr_smem + 8u;                               // [128 rows][2 halves] float: first-tile maxima

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;
  constexpr int NW = 2;                    // softmax warps per lane quarter
  constexpr int W_TMA = 4 * NW, W_MMA = 4 * NW + 1;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), TILES ? 128 : 256);
      mbar_init(o_done(i), 1);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag_smem), "r"(0u) : "memory");
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == W_MMA) tmem_alloc(tmem_ptr_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tmem_O = tmem_base + (uint32_t)NS * KV;

  if (warp == W_TMA) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {
      mbar_expect_tx(q_full, q_bytes);
      for (int c = 0; c < dchunks; ++c)
        tma_load_3d(&maps.q, sQ + (uint32_t)c * ATT_BLOCK_Q * 32u, q_full, p.q_col0 + head * p.dp + c * 16, qt * ATT_BLOCK_Q, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(kv_empty(st), ph ^ 1u);
      if (p.dbg_delay & 2) __nanosleep(3000);             // test hook: a slow TMA producer
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
    const uint64_t desc_k0 = make_smem_desc(sKV, 16, 256, SWZ_32B);
    const uint64_t desc_v0 = make_smem_desc(sKV + kv_tile_bytes, KV * 32u, 256, SWZ_32B);
    const uint32_t stage_step = (2 * kv_tile_bytes) >> 4;
    int qk_st = 0, qk_sb = 0;
    uint32_t qk_ph = 0;
    auto issue_qk = [&]() {
      mbar_wait(kv_full(qk_st), qk_ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = desc_k0 + (uint64_t)((uint32_t)qk_st * stage_step);
        const uint32_t tS = tmem_base + (uint32_t)qk_sb * KV;
        for (int c = 0; c < dchunks; ++c)
          umma_f16_ss(tS, desc_q0 + (uint64_t)(c * (ATT_BLOCK_Q * 32 / 16)), dk + (uint64_t)(c * (int)(kv_chunk_bytes >> 4)), idesc_qk, c != 0);
        umma_commit(s_full(qk_sb));
      }
      __syncwarp();
      if (++qk_st == NST) { qk_st = 0; qk_ph ^= 1u; }
      if (++qk_sb == NS) qk_sb = 0;
    };
    mbar_wait(q_full, 0);
    for (int i = 0; i < NS && i < n_tiles; ++i) issue_qk();
    int st = 0, sb = 0;
    uint32_t sph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(p_full(sb), sph);                         // both halves of P written over S[sb]
      if (p.dbg_delay & 1) __nanosleep(4000);             // test hook: a slow MMA issuer
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc_v0 + (uint64_t)((uint32_t)st * stage_step);
        const uint32_t tP = tmem_base + (uint32_t)sb * KV;
#pragma unroll
        for (int k = 0; k < KV / 16; ++k)                  // keys 16k .. 16k+15: P columns 8k (k < 2) or 32 + 8(k - 2); tile split: 8k
          umma_f16_ts(tmem_O, tP + (uint32_t)((TILES || k < 2) ? 8 * k : 32 + 8 * (k - 2)), dv + (uint64_t)(k * (512 / 16)), idesc_pv, (j | k) != 0);
        umma_commit(o_done(sb));
        umma_commit(kv_empty(st));
      }
      __syncwarp();
      if (j + NS < n_tiles) issue_qk();
      if (++st == NST) st = 0;
      if (++sb == NS) { sb = 0; sph ^= 1u; }
    }
  } else {
    // ---------------- softmax warps: (lane quarter, column half) ----------------
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int q_row = qt * ATT_BLOCK_Q + row;
    const uint32_t col0 = (uint32_t)half * 32u;           // score columns [col0, col0 + 32); P goes to [col0, col0 + 16)
    float m_ref = 0.f;
    bool overflow = false;
    if constexpr (TILES) {
      // ---- tile split: warp `half` handles tiles half, half + 2, ...; 32 columns at a time (registers) ----
      {
        // reference maximum = maximum of the row's FIRST tile, found by warp 0 of the quarter, read by warp 1
        if (half == 0) {
          mbar_wait(s_full(0), 0);
          tc_fence_after();
          float mh = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t sreg[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + (uint32_t)(c * 32), sreg);
            tmem_ld_wait();
            const int kv_valid = min(32, p.Skv - c * 32);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < kv_valid) mh = fmaxf(mh, __uint_as_float(sreg[i]));
          }
          m_ref = mh * p.scale_log2;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(mx_smem + (uint32_t)row * 4u), "f"(m_ref) : "memory");
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NW) : "memory");     // the warps of this lane quarter
        if (half != 0) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m_ref) : "r"(mx_smem + (uint32_t)row * 4u) : "memory");
      }
      int sb = half;                     // score buffer and barrier phase of tile j = half, half + NW, ...
      uint32_t sph = 0;
      for (int j = half; j < n_tiles; j += NW) {
        const uint32_t tS = tmem_base + (uint32_t)sb * KV + lane_addr;
        mbar_wait(s_full(sb), sph);
        if ((p.dbg_delay & 4) && half == 1) __nanosleep(5000);      // test hook: the softmax warps of the odd tiles lag
        if ((p.dbg_delay & 8) && half == 0) __nanosleep(5000);      //            ... of the even tiles
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sreg[32];
          tmem_ld_32x32b_x32(tS + (uint32_t)(c * 32), sreg);
          tmem_ld_wait();
          const int kv_valid = min(32, p.Skv - j * KV - c * 32);
          uint32_t pw[16];
          if (kv_valid == 32) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float pv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __uint_as_float(sreg[h * 16 + i]);
                m4[i & 3] = fmaxf(m4[i & 3], sv);
                const float x = fmaf(sv, p.scale_log2, -m_ref);
                const bool on_fma = POLY > 0 && ((i + 1) % (16 / (POLY > 0 ? POLY : 1))) == 0;
                pv[i] = on_fma ? ex2_poly3(x) : ex2f(x);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) pw[h * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
            }
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
            overflow |= mx > m_ref + 100.0f;
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float pv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float x = fmaf(__uint_as_float(sreg[h * 16 + i]), p.scale_log2, -m_ref);
                const bool ok = h * 16 + i < kv_valid;
                overflow |= ok && x > 100.0f;
                pv[i] = ok ? ex2f(x) : 0.f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) pw[h * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
            }
          }
          // P of keys 32c .. 32c+31 -> columns [16c, 16c + 16) of the score buffer (score columns < 32(c + 1) are consumed)
          tmem_st_32x32b_x16(tS + (uint32_t)(c * 16), pw);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(sb));
        sb += NW;
        if (sb >= NS) { sb -= NS; sph ^= 1u; }
      }
    } else {
    int sb = 0;
    uint32_t sph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t tS = tmem_base + (uint32_t)sb * KV + lane_addr + col0;
      mbar_wait(s_full(sb), sph);
      if ((p.dbg_delay & 4) && half == 1) __nanosleep(5000);
      if ((p.dbg_delay & 8) && half == 0) __nanosleep(5000);
      tc_fence_after();
      uint32_t sreg[32];
      tmem_ld_32x32b_x32(tS, sreg);
      tmem_ld_wait();
      const int kv_valid = min(32, p.Skv - j * KV - (int)col0);       // < 32 only in a ragged last tile (may be <= 0)
      if (j == 0) {
        // reference maximum of the row = maximum of its first tile (both halves; tile 0 always has a valid key in half 0)
        float mh = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < kv_valid) mh = fmaxf(mh, __uint_as_float(sreg[i]));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(mx_smem + (uint32_t)(row * 2 + half) * 4u), "f"(mh) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");     // the two warps of this lane quarter
        float mo;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mo) : "r"(mx_smem + (uint32_t)(row * 2 + (half ^ 1)) * 4u) : "memory");
        m_ref = fmaxf(mh, mo) * p.scale_log2;
      }
      uint32_t pw[16];
      if (kv_valid == 32) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float pv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float sv = __uint_as_float(sreg[h * 16 + i]);
            m4[i & 3] = fmaxf(m4[i & 3], sv);
            const float x = fmaf(sv, p.scale_log2, -m_ref);
            const bool on_fma = POLY > 0 && ((i + 1) % (16 / (POLY > 0 ? POLY : 1))) == 0;
            pv[i] = on_fma ? ex2_poly3(x) : ex2f(x);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) pw[h * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
        overflow |= mx > m_ref + 100.0f;
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float pv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = fmaf(__uint_as_float(sreg[h * 16 + i]), p.scale_log2, -m_ref);
            const bool ok = h * 16 + i < kv_valid;
            overflow |= ok && x > 100.0f;
            pv[i] = ok ? ex2f(x) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) pw[h * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
        }
      }
      tmem_st_32x32b_x16(tS, pw);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(sb));
      if (++sb == NS) { sb = 0; sph ^= 1u; }
    }
    }
    if (__any_sync(0xffffffffu, overflow) && lane == 0)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag_smem), "r"(1u) : "memory");
    // ---- epilogue: O / O[:, l_col] -> bf16; the two halves split the 16-column output chunks ----
    const int sb_last = (n_tiles - 1) % NS;
    const uint32_t ph_last = (uint32_t)((n_tiles - 1) / NS) & 1u;
    if constexpr (TILES) {
      // A parity wait is only right when the barrier is exactly one phase behind.  The warp that does NOT own the last tile
      // has consumed S(n - 2) only, which says nothing about P V(n - 4) — the previous phase of o_done(sb_last) — having
      // retired: with a slow MMA issuer the wait below passed one phase early and O was read with the last tiles missing
      // (found with compute-sanitizer's timing, tools/attn_race_probe.py; reproduced by the dbg_delay hook in
      // tests/test_attn_gpu.py).  S(n - 1) complete => Q K^T(n - 1) retired => P V(n - 4), issued before it, retired.
#ifndef DFB_SA8_NO_EPILOGUE_FIX      // (A/B builds only: shows that the robustness test catches the missing wait)
      mbar_wait(s_full(sb_last), ph_last);
#endif
    }
    mbar_wait(o_done(sb_last), ph_last);
    tc_fence_after();
    float l;
    {
      uint32_t r1[16];
      tmem_ld_32x32b_x16(tmem_O + lane_addr + (uint32_t)(p.l_col & ~15), r1);
      tmem_ld_wait();
      l = __uint_as_float(r1[0]);
#pragma unroll
      for (int i = 1; i < 16; ++i)
        if (i == (p.l_col & 15)) l = __uint_as_float(r1[i]);
    }
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.out + (size_t)b * p.out_batch_stride + (size_t)q_row * p.out_ld + p.out_col0 + head * p.dp;
    for (int c = half; c < dchunks; c += NW) {
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
  if (threadIdx.x == 0) {
    // one flag per CTA, written unconditionally (the buffer needs no initialisation): 1 = a score overflowed the static
    // reference maximum, attn_fwd_sa_kernel<.., REDO> recomputes this tile
    uint32_t f;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(f) : "r"(flag_smem) : "memory");
    const int lin_block = (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    p.redo_flags[lin_block] = (int)f;
  }
}

}  // namespace dfb
