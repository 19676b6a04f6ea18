// Self-attention kernel for the long-sequence, small-head layers of the UNet (S = 4096, d = 40 -> 48: five of the
// sixteen self-attention layers and 71 of the 84 ms attention takes in a 256-row step).  Included by dfb_attn.cu.
//
// Same pipeline as attn_fwd_db_kernel<64, true> (TMA K/V ring -> S = Q K^T into two TMEM score buffers -> softmax warps
// -> P back into TMEM over its score buffer -> TS-form O += P V), reworked around what round 2's measurements say bounds
// the softmax loop (tools/ubench/ubench_softmax_pipes.cu, profiles/r02_ubench_softmax_pipes.log):
//   * the MUFU does 16 ex2 / clk / SM -> 1024 cycles per 128 x 128 score tile; tcgen05.ld is NOT a bound (>= 200 B/clk/SM,
//     not the 64 the B300 notes suggest); the fast path costs 4.5 issue slots per score (fmax, ffma, ex2, fadd, half a
//     pack), i.e. 576 issue cycles per scheduler per tile — so neither pipe is full, the loop loses its time in the
//     per-tile fixed latencies (wait for S, first TMEM load, P store, fence, arrive: ~380 of 1411 cycles) that the two
//     co-resident CTAs take in convoy.
//   ONES: V carries 1.0 in a padding column of every head (written by the q|k|v projection's bias, attention.py), so the
//     P V MMA accumulates the softmax denominator in O[:, l_col] — from the same bf16-rounded probabilities the numerator
//     uses.  One FADD per score less (3.5 issue slots), no separate l to rescale.
//   POLY: POLY of every 16 exponentials are computed on the FMA pipe (Cody-Waite split by the magic-number add, degree-3
//     minimax polynomial, exponent by integer add; max relative error 7.5e-5, 1/50 of P's bf16 rounding) instead of the
//     MUFU — only worth it once ONES has freed the issue slots (round 1 measured the same idea 13 % slower without it).
//   Prefetch across tiles: the first TMEM load of tile j + 1 is issued BEFORE the P store / fence / arrive of tile j, so
//     the two fixed latencies overlap each other instead of adding up.
#pragma once

namespace dfb {

// 2^x on the FMA / ALU pipes.  x <= ~8 by the kernel's rescale rule; x < -125 flushes to 0 like ex2.approx.ftz.
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                    // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);              // x - round(x) in [-0.5, 0.5]
  float p = fmaf(0.0551716685295105f, f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

constexpr int ATT_SA_KV = 64;

// One 128-query tile `qt` of (batch blockIdx.z, head blockIdx.y).  RERUN: the CTA may run several tiles one after the other
// (the REDO kernel below): tensor memory is allocated without giving up the allocation permit and the mbarriers are
// invalidated at the end so that the next tile can initialise them again.
template <bool ONES, int POLY, bool RERUN>
__device__ __forceinline__ void attn_sa_tile(const AttnMaps& maps, const AttnKernelParams& p, const int qt) {
  constexpr int KV = ATT_SA_KV;
  constexpr int NST_MAX = 8;               // K/V ring depth (TMA issued that many tiles ahead)
  constexpr int NS_MAX = 3;                // score / probability ring in tensor memory
  const int NST = p.kv_stages;
  const int NS = p.s_ring;                 // 3 when 3 * 64 + dp <= 256 columns (two CTAs per SM), else 2
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
  auto p_full = [&](int i) { return bar_base + 8u + 8u * (NS_MAX + i); };
  auto o_done = [&](int i) { return bar_base + 8u + 8u * (2 * NS_MAX + i); };
  auto kv_full = [&](int s) { return bar_base + 8u + 8u * (3 * NS_MAX + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u + 8u * (3 * NS_MAX + NST_MAX + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u + 8u * (3 * NS_MAX + 2 * NST_MAX);

  // warps 0..3 softmax (TMEM lane quarter = warp), warp 4 TMA producer, warp 5 MMA issuer (control warps at the highest
  // warp ids: the issue arbiter favours them)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;
  constexpr int W_TMA = 4, W_MMA = 5;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) {
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
  if (warp == W_MMA) {
    if constexpr (RERUN) tmem_alloc_keep_permit(tmem_ptr_smem, p.tmem_cols);
    else tmem_alloc(tmem_ptr_smem, p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tmem_O = tmem_base + (uint32_t)NS * KV;
  // tuning hook: stamps of two CTAs — linear block 0 and block `tl_second` (= number of SMs: with two CTAs per SM the block
  // scheduler places it next to block 0) — 2048 int64 each, the SM id in the last slot
  const int lin_block = (int)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
  const int tl_slot = (RERUN || p.timeline == nullptr) ? -1 : (lin_block == 0 ? 0 : (lin_block == p.tl_second ? 1 : -1));
  const bool tl_cta = tl_slot >= 0;
  long long* const tl = tl_cta ? p.timeline + 2048 * tl_slot : nullptr;

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
    // ring positions kept incrementally (no division in the loop): QK runs NS tiles ahead of PV
    int qk_st = 0, qk_sb = 0;
    uint32_t qk_ph = 0;
    auto issue_qk = [&]() {                       // S[next score buffer] = Q K^T of the next K/V stage
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
      mbar_wait(p_full(sb), sph);                         // P written over S[sb]
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc_v0 + (uint64_t)((uint32_t)st * stage_step);
#pragma unroll
        for (int k = 0; k < KV / 16; ++k)
          umma_f16_ts(tmem_O, tmem_base + (uint32_t)sb * KV + (uint32_t)(8 * k), dv + (uint64_t)(k * (512 / 16)), idesc_pv, (j | k) != 0);
        umma_commit(o_done(sb));
        umma_commit(kv_empty(st));
      }
      __syncwarp();
      if (j + NS < n_tiles) issue_qk();                   // reuses score buffer sb: ordered after PV_j on the tensor pipe
      if (++st == NST) st = 0;
      if (++sb == NS) { sb = 0; sph ^= 1u; }
    }
  } else {
    // ---------------- softmax / correction / epilogue warps ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int q_row = qt * ATT_BLOCK_Q + row;
    float m_ref = -INFINITY, l = 0.f;
    const bool tls = tl_cta && warp == 0 && lane == 0;
    if (tls) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tl[2047] = (long long)smid;
    }
    uint32_t sreg[KV];
    uint32_t pw[KV / 2];                       // packed bf16 probabilities

    // one 16-score block against the running reference max -> 8 packed words (+ max / sum chains)
    auto block16 = [&](const uint32_t* s16, uint32_t* w8, float (&m8)[8], float (&l8)[8]) {
      float pv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float sv = __uint_as_float(s16[i]);
        m8[i & 7] = fmaxf(m8[i & 7], sv);
        const float x = fmaf(sv, p.scale_log2, -m_ref);
        // POLY of 16 on the FMA pipe, spread evenly (i = 3, 7, 11, 15 for POLY = 4; 7, 15 for POLY = 2)
        const bool on_fma = POLY > 0 && ((i + 1) % (16 / (POLY > 0 ? POLY : 1))) == 0;
        pv[i] = on_fma ? ex2_poly3(x) : ex2f(x);
        if constexpr (!ONES) l8[i & 7] += pv[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) w8[i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
    };

    // first score chunk of tile 0
    mbar_wait(s_full(0), 0);
    tc_fence_after();
    tmem_ld_32x32b_x32(tmem_base + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&sreg[0]));

    int sb = 0;                 // score buffer of tile j, its barrier phase, and the same for tile j - 1 (rescale wait)
    uint32_t sph = 0;
    int sb_prev = 0;
    uint32_t sph_prev = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t tS = tmem_base + (uint32_t)sb * KV + lane_addr;
      int sb_next = sb + 1;
      uint32_t sph_next = sph;
      if (sb_next == NS) { sb_next = 0; sph_next ^= 1u; }
      if (tls && j < 64) tl[j * 8 + 0] = clock64();
      const int kv_valid = min(KV, p.Skv - j * KV);
      bool careful = (kv_valid != KV) || (j == 0);
      float mx = -INFINITY;
      // chunk 0 (columns 0..31) is in flight since the end of the previous tile; chunk 1 follows
      tmem_ld_wait();
      tmem_ld_32x32b_x32(tS + 32u, *reinterpret_cast<uint32_t(*)[32]>(&sreg[32]));
      if (tls && j < 64) tl[j * 8 + 1] = clock64();
      if (!careful) {
        float m8[8], l8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { m8[i] = -INFINITY; l8[i] = 0.f; }
        block16(&sreg[0], &pw[0], m8, l8);
        block16(&sreg[16], &pw[8], m8, l8);
        tmem_ld_wait();
        block16(&sreg[32], &pw[16], m8, l8);
        block16(&sreg[48], &pw[24], m8, l8);
        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7]))) * p.scale_log2;
        careful = __any_sync(0xffffffffu, mx > m_ref + 8.0f);
        if constexpr (!ONES) {
          if (!careful) l += ((l8[0] + l8[1]) + (l8[2] + l8[3])) + ((l8[4] + l8[5]) + (l8[6] + l8[7]));
        }
      } else {
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < KV; ++i)
          if (i < kv_valid) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sreg[i]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      }
      if (careful) {
        // new reference max: rescale O (needs every earlier PV retired; with ONES the denominator column is part of O),
        // redo the tile from the registers
        const bool need = mx > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_ref;
          const float alpha = ex2f(m_ref - m_new);     // m_ref = -inf on the first tile -> 0
          if (j > 0) {
            mbar_wait(o_done(sb_prev), sph_prev);
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
#pragma unroll
          for (int i = 0; i < 8; ++i) pw[c * 8 + i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
        }
        if constexpr (!ONES) l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      }
      if (tls && j < 64) tl[j * 8 + 2] = clock64();
      // the next tile's first chunk goes in flight BEFORE this tile's P store / fence / arrive (sreg[0..31] is dead)
      if (j + 1 < n_tiles) {
        mbar_wait(s_full(sb_next), sph_next);
        tc_fence_after();
        if (tls && j < 64) tl[j * 8 + 4] = clock64();
        tmem_ld_32x32b_x32(tmem_base + (uint32_t)sb_next * KV + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&sreg[0]));
      }
      // P (bf16, two per column) over the first KV/2 columns of the score buffer it was computed from
      tmem_st_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&pw[0]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(sb));
      if (tls && j < 64) tl[j * 8 + 3] = clock64();
      sb_prev = sb; sph_prev = sph;
      sb = sb_next; sph = sph_next;
    }
    // ---- epilogue: O / l -> bf16 ----
    mbar_wait(o_done(sb_prev), sph_prev);
    tc_fence_after();
    if constexpr (ONES) {
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
  if constexpr (RERUN) {
    if (warp == W_TMA && lane == 0) {
      mbar_inval(q_full);
      for (int i = 0; i < NS; ++i) { mbar_inval(s_full(i)); mbar_inval(p_full(i)); mbar_inval(o_done(i)); }
      for (int st = 0; st < NST; ++st) { mbar_inval(kv_full(st)); mbar_inval(kv_empty(st)); }
    }
    __syncthreads();
  }
}

// REDO: the exact fallback behind attn_fwd_sa8_kernel (dfb_attn_sa8.cuh) — launched right behind it on a grid of ONE CTA
// per (batch, head) (a CTA per query tile, exiting at once unless flagged, cost 0.3 - 0.4 ms per launch at 49 k - 65 k CTAs:
// profiles/r02_ncu_launches_step_v2.csv).  The CTA reads the flags of its `n_q_tiles` query tiles 32 at a time (every warp
// loads the same words, so the ballot is CTA-uniform) and recomputes the flagged tiles one after the other — normally none.
template <bool ONES, int POLY, bool REDO = false>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_sa_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnKernelParams p) {
  if constexpr (REDO) {
    const int* flags = p.redo_flags + (size_t)(blockIdx.z * gridDim.y + blockIdx.y) * (size_t)p.n_q_tiles;
    const int lane = threadIdx.x & 31;
    for (int q0 = 0; q0 < p.n_q_tiles; q0 += 32) {
      const int f = q0 + lane < p.n_q_tiles ? flags[q0 + lane] : 0;
      uint32_t mask = __ballot_sync(0xffffffffu, f != 0);
      while (mask) {
        const int qt = q0 + __ffs((int)mask) - 1;
        mask &= mask - 1u;
        attn_sa_tile<ONES, POLY, true>(maps, p, qt);
      }
    }
  } else {
    attn_sa_tile<ONES, POLY, false>(maps, p, (int)blockIdx.x);
  }
}

}  // namespace dfb
