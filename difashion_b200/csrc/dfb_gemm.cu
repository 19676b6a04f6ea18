// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   warp EW     : TMA producer A (tile 128x64 bf16 via 2-D or 4-D tiled tensor map, OOB = zero
//                 gives the 3x3 'same' padding for free)
//   warp EW+1   : TMA producer B (tile block_n x 64)
//   warp EW+2   : TMEM allocator + single-lane tcgen05.mma issuer (M=128, N=block_n, K=16)
//   warps 0..EW-1: epilogue (tcgen05.ld -> registers -> swizzled smem transpose -> row-coalesced
//                 bias / time-embedding row bias / activation / residual / GEGLU -> bf16 or fp32
//                 global stores touching full 128-byte rows)
//
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) between TMA and MMA, and a 2-deep TMEM
// accumulator ring (tmem_full/tmem_empty) between MMA and epilogue, so the epilogue of tile i
// overlaps the main loop of tile i+1.  One CTA per SM, grid = min(#tiles, #SMs).
//
// Replaces (see include/dfb200.h): every nn.Conv2d / nn.Linear of diffusers'
// UNet2DConditionModel.forward as called from DiFashion/models/difashion.py:518-523.
#include "dfb_host.h"
#include "../../include/dfb200.h"

namespace dfb {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;                     // one 128-byte swizzle atom of bf16
constexpr int GEMM_MAX_STAGES = 8;                   // ring depth is chosen per launch from the stage size (see dfb_gemm)
constexpr int GEMM_MAX_BLOCK_N = 256;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;          // 16 KB
constexpr int GEMM_B_BYTES = GEMM_MAX_BLOCK_N * GEMM_BLOCK_K * 2;      // 32 KB (max)
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;          // 48 KB
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_EPI_STAGE_BYTES = 4096;                              // per-warp 32x32 fp32 transpose tile
constexpr int GEMM_THREADS = 96 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_SMEM_BYTES = 4 * GEMM_STAGE_BYTES + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;   // 227 KB budget
constexpr int GEMM_TMEM_COLS = 512;

struct GemmMaps {
  CUtensorMap a[2];
  CUtensorMap b;
  CUtensorMap c;     // bf16 output, box = one epilogue chunk (32 rows x 32 columns, or x 16 GEGLU outputs): TMA-store epilogue
};

struct GemmKernelParams {
  int M, N, block_n, n_tiles_m, n_tiles_n;
  int stages, stage_bytes;                    // smem ring: depth and stride (A 16 KB + B block_n x 128 B per stage)
  int n_sub, acc_ring;                        // MMA column sub-tiles per work item (block_n = n_sub * sub_n <= 512) and
                                              // accumulator ring depth in tensor memory (2 when 2 * block_n <= 512, else 1)
  int conv, TH, TB, tiles_per_img, w_tiles;   // w_tiles > 1: images wider than 128 pixels, one tile = 128 pixels of a row
  int nseg;
  int seg_ntaps[2];
  int seg_ncblk[2];
  int tap_dh[18], tap_dw[18], tap_coff[18];
  const float* bias;
  const float* rowbias;
  int rowbias_ld, rows_per_batch;
  const void* residual;
  int res_ld, res_fp32;
  void* out;
  int out_ld, out_fp32, geglu, vec_ok, act;
  float* gn_partial;   // optional GroupNorm partial statistics of the fp32 output
  int up_w, up_a, up_b, up_blk;   // up_w > 0: phase (up_a, up_b) of a fused nearest-2x upsample + 3x3 conv; rows are scattered into
                                  // the [B, 2H, 2W, N] output (up_w = W of the input, up_blk = H*W/32 partial blocks per image)
  int tma_out;                    // bf16 output written with cp.async.bulk.tensor stores through maps.c (modes 0 and GEGLU)
};

// TMA store of one staged chunk (shared -> global, bulk async-group completion); issued by one lane
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {      // at most N of this thread's bulk groups still READING shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Epilogue specialisations (compile-time, so the hot epilogue loop carries no runtime flag tests and the
// kernel image stays small enough for the instruction cache); EPI_GENERIC keeps every option at run time.
enum : int { EPI_OUT_F32 = 1, EPI_RES_F32 = 2, EPI_ROWBIAS = 4, EPI_GEGLU = 8, EPI_GENERIC = 16 };

// EW = number of epilogue warps (8, or 16 for the instruction-heavy GEGLU epilogue): EW/4 warps share a
// TMEM lane quarter and take 32-column chunks round-robin.
// CTA2: CTA-pair mode (cluster of 2, tcgen05.mma.cta_group::2): one work item is a 256 x block_n tile; CTA `rank` owns
// M tile 2*pm + rank (its 128 rows of A, its 128 accumulator rows, its epilogue) and loads half of the B tile; the
// leader (rank 0) issues the MMAs for both.  Per SM the tensor core then reads A (16 KB) + HALF a B tile per k-block
// from shared memory instead of A + a whole one, which is what bounds the 1-CTA kernel at block_n <= 224.
template <int MODE, int EW = GEMM_EPI_WARPS, bool CTA2 = false>
__global__ void __launch_bounds__(96 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmKernelParams p) {
  constexpr bool kGeneric = (MODE & EPI_GENERIC) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int STAGES = p.stages;
  const uint32_t STAGE_BYTES = (uint32_t)p.stage_bytes;
  const uint32_t epi_base = smem_base + (uint32_t)STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES;
  // barrier layout (8 bytes each): full[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * GEMM_MAX_STAGES + 4);

  // Roles: warps 0..EW-1 epilogue (TMEM lane quarter = warp & 3), warp EW TMA producer, warp EW+1 MMA issuer.
  // The issue arbiter favours the highest warp id of a scheduler, so the two single-lane control warps sit
  // above the issue-heavy epilogue warps (otherwise their TMA / MMA issue is starved).
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Two TMA producer warps (A tiles / B tiles): ncu showed ONE producer warp spending ~65 % of its time on the ~64
  // dependent uniform-datapath instructions of a k-block (address arithmetic, R2UR moves, elect, two UTMALDG), i.e.
  // ~400 cycles per k-block against 320 cycles of MMA at block_n = 160 — the narrow tiles were producer-issue-bound.
  constexpr int W_TMA = EW, W_TMA_B = EW + 1, W_MMA = EW + 2;
  // work items: tiles (1-CTA) or tile pairs (CTA2), strided over the CTAs / clusters of the persistent grid
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int n_units = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int unit0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_items = (CTA2 ? (p.n_tiles_m + 1) / 2 : p.n_tiles_m) * p.n_tiles_n;
  auto item_mt = [&](int item) { const int q = item / p.n_tiles_n; return CTA2 ? 2 * q + (int)rank : q; };
  auto item_nt = [&](int item) { return item % p.n_tiles_n; };
  const int nk = p.seg_ntaps[0] * p.seg_ncblk[0] + (p.nseg > 1 ? p.seg_ntaps[1] * p.seg_ncblk[1] : 0);

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    if (p.nseg > 1) tma_prefetch_desc(&maps.a[1]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);          // one arrive.expect_tx per producer warp (A, B)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CTA2 ? 2 * EW : EW);   // one arrive per epilogue warp (of both CTAs, on the leader's barrier)
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == W_MMA) {
    if constexpr (CTA2) tmem_alloc_pair(tmem_ptr_smem, GEMM_TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, GEMM_TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();     // barriers initialised and tensor memory allocated in BOTH CTAs
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == W_TMA) {
    // ===================== TMA producer, A tiles: ONE lane runs the loop =====================
    // (ncu: with the warp-uniform loop + elect_one per k-block this warp spent most of its time resolving the ~9
    // branches / reconvergence points of an iteration, ~250-400 cycles per k-block against 320 cycles of MMA at
    // block_n = 160; a single-lane loop has no elect, no BSSY/BSYNC, no __syncwarp.)
    // CTA2: each CTA loads its own A tile; the bytes of both are counted on the leader's full barrier.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (CTA2 ? 2u : 1u) * GEMM_A_BYTES;
      const uint32_t fbar0 = CTA2 ? mapa_shared(full_bar(0), 0) : full_bar(0);      // (the leader's) full barriers
      for (int item = unit0; item < num_items; item += n_units) {
        const int mt = item_mt(item);
        int b0 = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          if (p.TB == 1) {
            b0 = mt / p.tiles_per_img;
            const int r = mt - b0 * p.tiles_per_img;
            if (p.w_tiles > 1) {
              h0 = r / p.w_tiles;
              w0 = (r - h0 * p.w_tiles) * GEMM_BLOCK_M;
            } else {
              h0 = r * p.TH;
            }
          } else {
            b0 = mt * p.TB;
          }
        }
        const int row0 = mt * GEMM_BLOCK_M;
        for (int s = 0; s < p.nseg; ++s) {
          const void* tm = &maps.a[s];
          const int ncb = p.seg_ncblk[s];
          for (int tap = 0; tap < p.seg_ntaps[s]; ++tap) {
            const int ti = s * 9 + tap;
            const int cw = w0 + p.tap_dw[ti], ch = h0 + p.tap_dh[ti];
            int cc = p.tap_coff[ti];
            for (int cb = 0; cb < ncb; ++cb, cc += GEMM_BLOCK_K) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t sa = smem_base + (uint32_t)stage * STAGE_BYTES;
              const uint32_t fb = fbar0 + 8u * (uint32_t)stage;
              if (!CTA2 || rank == 0) mbar_expect_tx(full_bar(stage), tx_bytes);
              if constexpr (CTA2) {
                if (p.conv) tma_load_4d_pair(tm, sa, fb, cc, cw, ch, b0);
                else tma_load_2d_pair(tm, sa, fb, cc, row0);
              } else {
                if (p.conv) tma_load_4d(tm, sa, fb, cc, cw, ch, b0);
                else tma_load_2d(tm, sa, fb, cc, row0);
              }
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == W_TMA_B) {
    // ===================== TMA producer, B tiles (CTA2: this CTA's half of the tile); one lane =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // per MMA sub-tile (sub_n columns) this CTA holds sub_n rows of B, or its half of them in a CTA pair
      const uint32_t sub_n = (uint32_t)p.block_n / (uint32_t)p.n_sub;
      const uint32_t b_rows = CTA2 ? sub_n >> 1 : sub_n;
      const uint32_t tx_bytes = (CTA2 ? 2u : 1u) * (uint32_t)p.n_sub * b_rows * GEMM_BLOCK_K * 2;
      const uint32_t fbar0 = CTA2 ? mapa_shared(full_bar(0), 0) : full_bar(0);
      for (int item = unit0; item < num_items; item += n_units) {
        const int n0 = item_nt(item) * p.block_n + (int)(rank * b_rows);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sb = smem_base + (uint32_t)stage * STAGE_BYTES + GEMM_A_BYTES;
          if (!CTA2 || rank == 0) mbar_expect_tx(full_bar(stage), tx_bytes);
          for (int sub = 0; sub < p.n_sub; ++sub) {
            const uint32_t dst = sb + (uint32_t)sub * b_rows * (GEMM_BLOCK_K * 2);
            if constexpr (CTA2) tma_load_2d_pair(&maps.b, dst, fbar0 + 8u * (uint32_t)stage, kb * GEMM_BLOCK_K, n0 + sub * (int)sub_n);
            else tma_load_2d(&maps.b, dst, fbar0 + 8u * (uint32_t)stage, kb * GEMM_BLOCK_K, n0 + sub * (int)sub_n);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == W_MMA && (!CTA2 || rank == 0)) {
    // ===================== MMA issuer (CTA2: the leader CTA only) =====================
    // The whole warp runs the (warp-uniform) loop so that descriptors live in uniform registers; one
    // elected lane issues tcgen05.mma / tcgen05.commit.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t sub_n = (uint32_t)p.block_n / (uint32_t)p.n_sub;
    const uint32_t idesc = make_idesc_f16(CTA2 ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M, sub_n, true);
    const uint64_t desc_a0 = make_smem_desc(smem_base, 16, 1024, SWZ_128B);
    const uint32_t sub_b_step = ((CTA2 ? sub_n >> 1 : sub_n) * (GEMM_BLOCK_K * 2)) >> 4;     // descriptor units between B sub-tiles
    const bool ring2 = p.acc_ring == 2;
    int it = 0;
    for (int item = unit0; item < num_items; item += n_units, ++it) {
      const int as = ring2 ? (it & 1) : 0;
      const uint32_t aphase = ring2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * GEMM_MAX_BLOCK_N;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t da = desc_a0 + (uint64_t)(((uint32_t)stage * STAGE_BYTES) >> 4);
        const uint64_t db = da + (uint64_t)(GEMM_A_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the
            // (addr >> 4) start-address field
            if constexpr (CTA2) umma_f16_ss_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            else umma_f16_ss(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            if (p.n_sub == 2) {       // second column sub-tile of a wide (257..512) tile: same A tile, next B rows, next TMEM columns
              if constexpr (CTA2) umma_f16_ss_pair(tmem_d + sub_n, da + (uint64_t)(2 * k), db + (uint64_t)(sub_b_step + 2 * k), idesc, (kb | k) != 0);
              else umma_f16_ss(tmem_d + sub_n, da + (uint64_t)(2 * k), db + (uint64_t)(sub_b_step + 2 * k), idesc, (kb | k) != 0);
            }
          }
          if constexpr (CTA2) {
            umma_commit_pair(empty_bar(stage));                     // frees the smem slot of BOTH CTAs
            if (kb == nk - 1) umma_commit_pair(tfull_bar(as));      // accumulators ready for both epilogues
          } else {
            umma_commit(empty_bar(stage));       // frees the smem slot when these MMAs retire
            if (kb == nk - 1) umma_commit(tfull_bar(as));   // accumulator ready for the epilogue
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp < EW) {
    // ===================== epilogue warps (0..EW-1) =====================
    // Two warps per TMEM lane quarter; each takes every other 32-column chunk.  Accumulators go
    // TMEM -> registers (thread = row) -> a per-warp swizzled smem tile -> registers in a
    // row-coalesced layout (8 lanes x 16 B per row), where bias / time-embedding row bias /
    // activation / residual are applied and global memory is touched with full 128-byte rows.
    const int ew = warp;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // which share of the chunks (0 .. EW/4-1)
    const uint32_t stg = epi_base + (uint32_t)ew * (GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES / EW);
    // run-time flags in generic mode, compile-time constants otherwise
    const bool f_geglu = kGeneric ? (p.geglu != 0) : ((MODE & EPI_GEGLU) != 0);
    const bool f_out32 = kGeneric ? (p.out_fp32 != 0) : ((MODE & EPI_OUT_F32) != 0);
    const bool f_res = kGeneric ? (p.residual != nullptr) : ((MODE & EPI_RES_F32) != 0);
    const bool f_res32 = kGeneric ? (p.res_fp32 != 0) : true;
    const bool f_rowbias = kGeneric ? (p.rowbias != nullptr) : ((MODE & EPI_ROWBIAS) != 0);
    const bool f_vec = kGeneric ? (p.vec_ok != 0) : true;       // specialised modes require aligned, N % 4 == 0
    const int f_act = kGeneric ? p.act : 0;
    // Residual reads are software-pipelined one chunk ahead (across tile boundaries too): small-K GEMMs with an
    // fp32 residual are HBM-latency-bound in the epilogue, so every warp keeps a second 4 KB of loads in flight
    // while it transposes / stores the current chunk.
    constexpr bool kPipeRes = !kGeneric && (MODE & EPI_RES_F32) != 0;
    constexpr int CSTEP = EW / 4;
    float4 resn[8];
    int pf_item = -1, pf_c = -1;
    auto load_res = [&](int mt_, int nt_, int c_, float4 (&dst)[8]) {
      const int r0 = mt_ * GEMM_BLOCK_M + quarter * 32 + (lane >> 3);
      const int col_ = nt_ * p.block_n + c_ * 32 + 4 * (lane & 7);
      const bool ok = col_ + 4 <= p.N;
      const float* rp = reinterpret_cast<const float*>(p.residual) + (size_t)r0 * p.res_ld + col_;
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        dst[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && r0 + itr * 4 < p.M) dst[itr] = __ldg(reinterpret_cast<const float4*>(rp + (size_t)(itr * 4) * p.res_ld));
      }
    };
    int it = 0;
    int tma_chunk = 0;                          // TMA-store epilogue: staging tile parity of this warp
    const uint32_t tempty_leader0 = CTA2 ? mapa_shared(tempty_bar(0), 0) : 0u;     // the leader's tmem_empty barriers
    for (int item = unit0; item < num_items; item += n_units, ++it) {
      const int mt = item_mt(item);
      const int nt = item_nt(item);
      const int as = p.acc_ring == 2 ? (it & 1) : 0;
      const uint32_t aphase = p.acc_ring == 2 ? ((it >> 1) & 1) : (it & 1);
      const int row0 = mt * GEMM_BLOCK_M + quarter * 32;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)as * GEMM_MAX_BLOCK_N;
      const bool full_rows = row0 + 32 <= p.M;                 // warp-uniform: no row masking needed
      bool waited = false;
      // the warps of a lane quarter rotate their first chunk from tile to tile, so an odd chunk count
      // (block_n = 160: 5 chunks over 2 warps) balances over consecutive tiles
      for (int c = (half + it) % CSTEP; c < p.block_n / 32; c += CSTEP) {
        const int n0 = nt * p.block_n + c * 32;
        if (n0 >= p.N) break;                  // warp-uniform
        if (f_geglu) {
          if (!waited) { mbar_wait(tfull_bar(as), aphase); tc_fence_after(); waited = true; }
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          // row layout: columns [0,16) = values, [16,32) = gates of output columns n0/2 .. n0/2+15
          float o[16];
          if (!kGeneric && p.bias) {     // specialised mode: bias is 16-byte aligned, N % 32 == 0 -> uniform float4 loads
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bv = __ldg(bp + j4), bg = __ldg(bp + 4 + j4);
              o[4 * j4 + 0] = (__uint_as_float(r[4 * j4 + 0]) + bv.x) * gelu_sigmoid_f(__uint_as_float(r[16 + 4 * j4 + 0]) + bg.x);
              o[4 * j4 + 1] = (__uint_as_float(r[4 * j4 + 1]) + bv.y) * gelu_sigmoid_f(__uint_as_float(r[16 + 4 * j4 + 1]) + bg.y);
              o[4 * j4 + 2] = (__uint_as_float(r[4 * j4 + 2]) + bv.z) * gelu_sigmoid_f(__uint_as_float(r[16 + 4 * j4 + 2]) + bg.z);
              o[4 * j4 + 3] = (__uint_as_float(r[4 * j4 + 3]) + bv.w) * gelu_sigmoid_f(__uint_as_float(r[16 + 4 * j4 + 3]) + bg.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + n0 + j) : 0.f);
              const float g = __uint_as_float(r[16 + j]) + (p.bias ? __ldg(p.bias + n0 + 16 + j) : 0.f);
              o[j] = a * gelu_sigmoid_f(g);
            }
          }
          if constexpr (!kGeneric) {
            if (p.tma_out) {
              // TMA-store epilogue: this row's 16 bf16 outputs (32 bytes) go into a [32 rows][32 B] tile with the 32-byte swizzle
              // (conflict-free 16-byte stores), one lane issues the bulk tensor store; two tiles per warp alternate, so the store
              // of chunk i drains while chunk i + 1 is computed.  Half the shared-memory traffic of the transposing path and no
              // LDS / STG in the warp's instruction stream; rows >= M and columns >= N/2 are clipped by the tensor map.
              const uint32_t dst0 = stg + (uint32_t)(tma_chunk & 1) * 1024u;
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
              const uint32_t rowb = dst0 + (uint32_t)lane * 32u;
              const uint32_t sw = (uint32_t)((lane >> 2) & 1) << 4;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (0u ^ sw)), "r"(pack_bf16x2(o[0], o[1])), "r"(pack_bf16x2(o[2], o[3])),
                           "r"(pack_bf16x2(o[4], o[5])), "r"(pack_bf16x2(o[6], o[7])) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (16u ^ sw)), "r"(pack_bf16x2(o[8], o[9])), "r"(pack_bf16x2(o[10], o[11])),
                           "r"(pack_bf16x2(o[12], o[13])), "r"(pack_bf16x2(o[14], o[15])) : "memory");
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) tma_store_2d(&maps.c, dst0, n0 >> 1, row0);
              ++tma_chunk;
              continue;
            }
          }
          // staging tile [32 rows][64 B], 64-byte swizzle
#pragma unroll
          for (int u2 = 0; u2 < 4; ++u2) {
            const uint32_t dst = stg + (uint32_t)lane * 64u + (uint32_t)((u2 ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o[4 * u2]), "f"(o[4 * u2 + 1]),
                         "f"(o[4 * u2 + 2]), "f"(o[4 * u2 + 3]) : "memory");
          }
          __syncwarp();
          const int ug = lane & 3;
          const int colo = (n0 >> 1) + 4 * ug;
#pragma unroll
          for (int itr = 0; itr < 4; ++itr) {
            const int rr = itr * 8 + (lane >> 2);
            const int m = row0 + rr;
            float4 x;
            const uint32_t src = stg + (uint32_t)rr * 64u + (uint32_t)((ug ^ ((rr >> 1) & 3)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(src));
            if (full_rows || m < p.M) {
              const size_t off = (size_t)m * p.out_ld + colo;
              if (f_out32) {
                float* dst = reinterpret_cast<float*>(p.out) + off;
                if (f_vec) *reinterpret_cast<float4*>(dst) = x;
                else { dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w; }
              } else {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
                if (f_vec) *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
                else { dst[0] = __float2bfloat16(x.x); dst[1] = __float2bfloat16(x.y); dst[2] = __float2bfloat16(x.z); dst[3] = __float2bfloat16(x.w); }
              }
            }
          }
          __syncwarp();
          continue;
        }
        if constexpr (MODE == 0) {
          if (p.tma_out) {
            // TMA-store epilogue for the bf16-output, bias-only GEMMs (q|k|v, q, K/V projections): thread = accumulator row, its
            // 32 columns (64 bytes) staged with the 64-byte swizzle, one bulk tensor store per chunk (see the GEGLU path)
            float4 bv[8];
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              bv[j4] = (p.bias && n0 + 4 * j4 < p.N) ? __ldg(bp + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (!waited) { mbar_wait(tfull_bar(as), aphase); tc_fence_after(); waited = true; }
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            const uint32_t dst0 = stg + (uint32_t)(tma_chunk & 1) * 2048u;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            const uint32_t rowb = dst0 + (uint32_t)lane * 64u;
            const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 b0 = bv[2 * q4], b1 = bv[2 * q4 + 1];
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (((uint32_t)q4 ^ sw) << 4)),
                           "r"(pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) + b0.x, __uint_as_float(r[8 * q4 + 1]) + b0.y)),
                           "r"(pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) + b0.z, __uint_as_float(r[8 * q4 + 3]) + b0.w)),
                           "r"(pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) + b1.x, __uint_as_float(r[8 * q4 + 5]) + b1.y)),
                           "r"(pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) + b1.z, __uint_as_float(r[8 * q4 + 7]) + b1.w)) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(&maps.c, dst0, n0, row0);
            ++tma_chunk;
            continue;
          }
        }
        if constexpr (!kGeneric) {
          // ---- specialised epilogue (aligned, N % 4 == 0, no activation): straight-line code.  lane -> 4 columns
          //      (u) x 8 rows (rsub + 4 * itr); a lane has either all 4 of its columns or none. ----
          const int u = lane & 7, rsub = lane >> 3;
          const int col = n0 + 4 * u;
          const bool lane_ok = col < p.N;
          float4 resv[8];
          if constexpr (kPipeRes) {
            if (pf_item == item && pf_c == c) {
#pragma unroll
              for (int itr = 0; itr < 8; ++itr) resv[itr] = resn[itr];
            } else {
              load_res(mt, nt, c, resv);
            }
            // this warp's next chunk: same tile, or its first chunk of the CTA's next tile
            int nitem = item, nc = c + CSTEP;
            if (nc >= p.block_n / 32 || nt * p.block_n + nc * 32 >= p.N) {
              nitem = item + n_units;
              nc = (half + it + 1) % CSTEP;
            }
            pf_item = -1;
            if (nitem < num_items) {
              const int nmt = item_mt(nitem), nnt = item_nt(nitem);
              if (nc < p.block_n / 32 && nnt * p.block_n + nc * 32 < p.N) {
                load_res(nmt, nnt, nc, resn);
                pf_item = nitem;
                pf_c = nc;
              }
            }
          }
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (lane_ok) {
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            if constexpr ((MODE & EPI_ROWBIAS) != 0) {
              // rows_per_batch % 32 == 0 (checked on the host): the warp's 32 rows share one batch row
              const int rbrow = min(row0, p.M - 1) / p.rows_per_batch;
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)rbrow * p.rowbias_ld + col));
              b4.x += t.x; b4.y += t.y; b4.z += t.z; b4.w += t.w;
            }
          }
          if (!waited) { mbar_wait(tfull_bar(as), aphase); tc_fence_after(); waited = true; }
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          // staging tile [32 rows][128 B], 128-byte swizzle
#pragma unroll
          for (int u2 = 0; u2 < 8; ++u2) {
            const uint32_t dst = stg + (uint32_t)lane * 128u + (uint32_t)((u2 ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * u2]), "r"(r[4 * u2 + 1]),
                         "r"(r[4 * u2 + 2]), "r"(r[4 * u2 + 3]) : "memory");
          }
          __syncwarp();
          float4 x[8];
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int rr = itr * 4 + rsub;
            const uint32_t src = stg + (uint32_t)rr * 128u + (uint32_t)((u ^ (rr & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[itr].x), "=f"(x[itr].y), "=f"(x[itr].z), "=f"(x[itr].w) : "r"(src));
          }
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            x[itr].x += b4.x; x[itr].y += b4.y; x[itr].z += b4.z; x[itr].w += b4.w;
            if constexpr ((MODE & EPI_RES_F32) != 0) {
              x[itr].x += resv[itr].x; x[itr].y += resv[itr].y; x[itr].z += resv[itr].z; x[itr].w += resv[itr].w;
            }
          }
          const bool f_gn = (MODE & EPI_OUT_F32) != 0 && p.gn_partial != nullptr;
          if (f_gn) {
            // (sum, sumsq) of this lane's two channel pairs over the warp's 32 rows (gn_partial needs M % 32 == 0,
            // so every row is valid): fold the 4 row groups, lanes 0..7 write 128 contiguous bytes per chunk
            float gs0 = 0.f, gq0 = 0.f, gs1 = 0.f, gq1 = 0.f;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              gs0 += x[itr].x + x[itr].y; gq0 += x[itr].x * x[itr].x + x[itr].y * x[itr].y;
              gs1 += x[itr].z + x[itr].w; gq1 += x[itr].z * x[itr].z + x[itr].w * x[itr].w;
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              gs0 += __shfl_xor_sync(0xffffffffu, gs0, o); gq0 += __shfl_xor_sync(0xffffffffu, gq0, o);
              gs1 += __shfl_xor_sync(0xffffffffu, gs1, o); gq1 += __shfl_xor_sync(0xffffffffu, gq1, o);
            }
            if (lane < 8 && lane_ok && row0 < p.M) {
              size_t blk = (size_t)(row0 >> 5);
              if constexpr (MODE == EPI_OUT_F32) {
                if (p.up_w > 0)      // upsample phase: this phase's blocks of image `img` sit at [img * 4 * up_blk + phase * up_blk, ...)
                  blk = (blk / (size_t)p.up_blk) * (size_t)(4 * p.up_blk) + (size_t)((2 * p.up_a + p.up_b) * p.up_blk) + blk % (size_t)p.up_blk;
              }
              *reinterpret_cast<float4*>(p.gn_partial + (blk * (size_t)(p.N >> 1) + (size_t)(col >> 1)) * 2) = make_float4(gs0, gq0, gs1, gq1);
            }
          }
          const int mrow = row0 + rsub;
          if constexpr ((MODE & EPI_OUT_F32) != 0) {
            if constexpr (MODE == EPI_OUT_F32) {
              if (p.up_w > 0) {
                // fused upsample phase (a, b): row (img * H + i) * W + j of the low-resolution grid -> pixel (2i + a, 2j + b) of
                // the [B, 2H, 2W, N] output (W is a power of two: conv tiles need 128 % W == 0)
                const int sh = 31 - __clz(p.up_w);
#pragma unroll
                for (int itr = 0; itr < 8; ++itr) {
                  const int m = mrow + itr * 4;
                  if (lane_ok && (full_rows || m < p.M)) {
                    const size_t orow = ((size_t)(2 * (m >> sh) + p.up_a) << (sh + 1)) + (size_t)(2 * (m & (p.up_w - 1)) + p.up_b);
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.out_ld + col) = x[itr];
                  }
                }
                __syncwarp();
                continue;
              }
            }
            float* op = reinterpret_cast<float*>(p.out) + (size_t)mrow * p.out_ld + col;
            const size_t ostep = (size_t)4 * p.out_ld;
            if (full_rows) {
              if (lane_ok) {
#pragma unroll
                for (int itr = 0; itr < 8; ++itr) *reinterpret_cast<float4*>(op + itr * ostep) = x[itr];
              }
            } else {
#pragma unroll
              for (int itr = 0; itr < 8; ++itr)
                if (lane_ok && mrow + itr * 4 < p.M) *reinterpret_cast<float4*>(op + itr * ostep) = x[itr];
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)mrow * p.out_ld + col;
            const size_t ostep = (size_t)4 * p.out_ld;
            if (full_rows) {
              if (lane_ok) {
#pragma unroll
                for (int itr = 0; itr < 8; ++itr)
                  *reinterpret_cast<uint2*>(op + itr * ostep) = make_uint2(pack_bf16x2(x[itr].x, x[itr].y), pack_bf16x2(x[itr].z, x[itr].w));
              }
            } else {
#pragma unroll
              for (int itr = 0; itr < 8; ++itr)
                if (lane_ok && mrow + itr * 4 < p.M)
                  *reinterpret_cast<uint2*>(op + itr * ostep) = make_uint2(pack_bf16x2(x[itr].x, x[itr].y), pack_bf16x2(x[itr].z, x[itr].w));
            }
          }
          __syncwarp();
          continue;
        }
        // ---- prefetch everything this chunk reads from global memory (row-coalesced layout: lane ->
        //      4 columns, 8 row groups) BEFORE waiting for the accumulator, so HBM latency is hidden ----
        const int u = lane & 7;
        const int col = n0 + 4 * u;
        const int nval = p.N - col;             // > 0: valid columns among this lane's 4
        const bool vec = f_vec && nval >= 4;
        float4 resv[8];
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        float rb4[4] = {0.f, 0.f, 0.f, 0.f};
        bool rb_uniform = false;
        if (f_res && vec) {
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int m = row0 + itr * 4 + (lane >> 3);
            resv[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (full_rows || m < p.M) {
              const size_t roff = (size_t)m * p.res_ld + col;
              if (f_res32) {
                resv[itr] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + roff));
              } else {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff));
                resv[itr] = make_float4(bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y));
              }
            }
          }
        }
        if (p.bias && nval > 0) {
          if (vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nval) b4[e] = __ldg(p.bias + col + e);
          }
        }
        if (f_rowbias && nval > 0) {
          const int mlast = min(row0 + 31, p.M - 1);
          rb_uniform = row0 < p.M && (row0 / p.rows_per_batch) == (mlast / p.rows_per_batch);   // warp-uniform
          if (rb_uniform) {
            const float* rb = p.rowbias + (size_t)(row0 / p.rows_per_batch) * p.rowbias_ld + col;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nval) rb4[e] = __ldg(rb + e);       // one load per chunk; ADDED AFTER the bias, like the per-row path below
          }
        }
        const bool f_gn = (kGeneric || (MODE & EPI_OUT_F32)) && p.gn_partial != nullptr;
        float gs0 = 0.f, gq0 = 0.f, gs1 = 0.f, gq1 = 0.f;
        if (!waited) { mbar_wait(tfull_bar(as), aphase); tc_fence_after(); waited = true; }
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        // staging tile [32 rows][128 B], 128-byte swizzle
#pragma unroll
        for (int u2 = 0; u2 < 8; ++u2) {
          const uint32_t dst = stg + (uint32_t)lane * 128u + (uint32_t)((u2 ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * u2]), "r"(r[4 * u2 + 1]),
                       "r"(r[4 * u2 + 2]), "r"(r[4 * u2 + 3]) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + (lane >> 3);
          const int m = row0 + rr;
          float x[4];
          const uint32_t src = stg + (uint32_t)rr * 128u + (uint32_t)((u ^ (rr & 7)) << 4);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "r"(src));
          if ((!full_rows && m >= p.M) || nval <= 0) continue;
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] += b4[e];
          // (acc + bias) + rowbias in BOTH paths: folding the row bias into the column bias first rounds differently, and
          // whether a 32-row chunk lies in one batch row depends on the row's position in the batch (images of 16 or 4
          // pixels: the last chunk of an odd batch) — found with tools/shared_prefix_diag2.py on B200, round 2
          if (f_rowbias && rb_uniform) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] += rb4[e];
          }
          if (f_rowbias && !rb_uniform) {
            const float* rb = p.rowbias + (size_t)(m / p.rows_per_batch) * p.rowbias_ld + col;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nval) x[e] += __ldg(rb + e);
          }
          if (kGeneric && f_act) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (f_act == DFB_ACT_SILU) x[e] = silu_f(x[e]);
              else if (f_act == DFB_ACT_LEAKY_RELU) x[e] = x[e] > 0.f ? x[e] : 0.01f * x[e];
              else if (f_act == DFB_ACT_QUICK_GELU) x[e] = x[e] / (1.0f + __expf(-1.702f * x[e]));
              else if (f_act == DFB_ACT_GELU) x[e] = gelu_fast_f(x[e]);
              else x[e] = tanhf(x[e]);
            }
          }
          if (f_res) {
            if (vec) {
              x[0] += resv[itr].x; x[1] += resv[itr].y; x[2] += resv[itr].z; x[3] += resv[itr].w;
            } else {
              const size_t roff = (size_t)m * p.res_ld + col;
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < nval)
                  x[e] += f_res32 ? reinterpret_cast<const float*>(p.residual)[roff + e]
                                  : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[roff + e]);
            }
          }
          if (f_gn) {
            gs0 += x[0] + x[1]; gq0 += x[0] * x[0] + x[1] * x[1];
            gs1 += x[2] + x[3]; gq1 += x[2] * x[2] + x[3] * x[3];
          }
          // fused upsample phase: row (b*H + i)*W + j of the low-resolution grid -> pixel (2i + a, 2j + b) of the output
          const size_t orow = p.up_w > 0 ? (size_t)(2 * (m / p.up_w) + p.up_a) * (size_t)(2 * p.up_w) + (size_t)(2 * (m % p.up_w) + p.up_b)
                                         : (size_t)m;
          const size_t off = orow * p.out_ld + col;
          if (f_out32) {
            float* dst = reinterpret_cast<float*>(p.out) + off;
            if (vec) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
            else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < nval) dst[e] = x[e];
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
            if (vec) *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]));
            else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < nval) dst[e] = __float2bfloat16(x[e]);
            }
          }
        }
        if (f_gn) {
          // fold the 4 row groups (lanes with equal column unit), then lanes 0..7 write (sum, sumsq) of their two
          // channel pairs over this warp's 32 rows: partial[row_block][pair][2], 128 contiguous bytes per chunk
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            gs0 += __shfl_xor_sync(0xffffffffu, gs0, o); gq0 += __shfl_xor_sync(0xffffffffu, gq0, o);
            gs1 += __shfl_xor_sync(0xffffffffu, gs1, o); gq1 += __shfl_xor_sync(0xffffffffu, gq1, o);
          }
          if (lane < 8 && nval >= 4 && row0 < p.M) {
            size_t blk = (size_t)(row0 >> 5);
            if (p.up_w > 0)      // this phase's blocks of image `img` sit at [img * 4 * up_blk + phase * up_blk, ...)
              blk = (blk / (size_t)p.up_blk) * (size_t)(4 * p.up_blk) + (size_t)((2 * p.up_a + p.up_b) * p.up_blk) + blk % (size_t)p.up_blk;
            *reinterpret_cast<float4*>(p.gn_partial + (blk * (size_t)(p.N >> 1) + (size_t)(col >> 1)) * 2) =
                make_float4(gs0, gq0, gs1, gq1);
          }
        }
        __syncwarp();
      }
      if (!waited) { mbar_wait(tfull_bar(as), aphase); tc_fence_after(); }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) mbar_arrive_cluster(tempty_leader0 + 8u * (uint32_t)as);
        else mbar_arrive(tempty_bar(as));
      }
    }
    // TMA-store epilogue: wait for this lane's bulk stores to COMPLETE (not only to have read shared memory) before the CTA
    // retires: the next kernel on the stream reads what they write
    if (p.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();     // the peer's shared / tensor memory stays alive until the leader's MMAs retired
  else __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_pair(tmem_base, GEMM_TMEM_COLS);
    else tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

// DFB_GEMM_TMA_STORE=0 disables the TMA-store epilogue (read once)
static bool tma_store_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_GEMM_TMA_STORE");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// tuning hook: DFB_GEMM_STAGES=<n> caps the smem ring depth (read once)
static int q_stages_override() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_GEMM_STAGES");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// tuning hook: DFB_GEMM_BN="640:224,320:160" overrides the automatic N tile for the listed N (read once)
static int bn_override(int N) {
  static int keys[8], vals[8], n = -1;
  if (n < 0) {
    n = 0;
    const char* e = getenv("DFB_GEMM_BN");
    while (e && *e && n < 8) {
      int k = 0, v = 0;
      if (sscanf(e, "%d:%d", &k, &v) == 2) { keys[n] = k; vals[n] = v; ++n; }
      e = strchr(e, ',');
      if (e) ++e;
    }
  }
  for (int i = 0; i < n; ++i)
    if (keys[i] == N) return vals[i];
  return 0;
}

// tuning hook: DFB_GEMM_CTA_GROUP=1|2 forces the 1-CTA / CTA-pair kernel wherever the caller leaves it automatic
static int cta_group_override() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_GEMM_CTA_GROUP");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
  }
  return v;
}

static int q_stages_pair_override() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_GEMM_STAGES_PAIR");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static bool wide_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_GEMM_WIDE");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

static int choose_block_n(int N) {
  if (bn_override(N) > 0) return bn_override(N);
  // cost of one row of output tiles ~ n_tiles * (block_n + 48): executed MMA columns plus a per-tile charge for
  // re-reading the A tile (L2 -> smem) and the epilogue hand-off.  Measured on B200 (profiles/r01_gemm_tile_tuning.md):
  // N = 640 runs 1 % faster per UNet step as 3 x 224 (5 % zero columns) than as 4 x 160; every other N of the UNet
  // keeps its zero-waste tile.
  static const int cands[] = {256, 224, 192, 160, 128, 96, 64, 32};
  int best = 32, best_cost = 1 << 30;
  for (int bn : cands) {
    const int cost = ((N + bn - 1) / bn) * (bn + 48);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_gemm(const dfb_gemm_params* q, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DFB_REQUIRE(q != nullptr, "dfb_gemm: null params");
  DFB_REQUIRE(q->nseg == 1 || q->nseg == 2, "dfb_gemm: nseg must be 1 or 2");
  DFB_REQUIRE(q->M > 0 && q->N > 0, "dfb_gemm: empty problem");
  DFB_REQUIRE(q->act >= DFB_ACT_NONE && q->act <= DFB_ACT_GELU, "dfb_gemm: unknown activation");
  DFB_REQUIRE(q->w != nullptr && q->out != nullptr && q->a[0] != nullptr, "dfb_gemm: null buffer");
  DFB_REQUIRE((reinterpret_cast<uintptr_t>(q->w) & 15) == 0 && (q->w_ld % 8) == 0, "dfb_gemm: weights must be 16B aligned");

  GemmKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));

  kp.M = q->M;
  kp.N = q->N;
  int bn = q->block_n > 0 ? q->block_n : choose_block_n(q->N);
  DFB_REQUIRE(q->cta_group >= 0 && q->cta_group <= 2, "dfb_gemm: cta_group must be 0 (auto), 1 or 2");
  DFB_REQUIRE(bn % 32 == 0 && bn >= 32 && (bn <= GEMM_MAX_BLOCK_N || (bn <= 2 * GEMM_MAX_BLOCK_N && bn % 64 == 0)),
              "dfb_gemm: block_n must be a multiple of 32 in [32,256], or a multiple of 64 in (256,512] (two MMA sub-tiles)");
  kp.conv = q->conv ? 1 : 0;
  kp.nseg = q->nseg;

  uint32_t boxA[4];
  if (kp.conv) {
    const int B = q->B, H = q->H, W = q->W;
    DFB_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)B * H * W == q->M, "dfb_gemm: conv geometry does not match M");
    kp.w_tiles = 1;
    if (W > 128) {
      // wide images (VAE decoder: 256 / 512 pixels): one tile = 128 consecutive pixels of one row
      DFB_REQUIRE(W % 128 == 0, "dfb_gemm: conv width above 128 must be a multiple of 128");
      kp.TH = 1;
      kp.TB = 1;
      kp.w_tiles = W / 128;
      kp.tiles_per_img = H * kp.w_tiles;
      kp.n_tiles_m = B * kp.tiles_per_img;
      boxA[0] = GEMM_BLOCK_K; boxA[1] = 128; boxA[2] = 1; boxA[3] = 1;
    } else {
      DFB_REQUIRE((128 % W) == 0, "dfb_gemm: conv width must divide 128");
      int TH = 128 / W;
      if (TH > H) TH = H;
      DFB_REQUIRE(H % TH == 0 && (128 % (W * TH)) == 0, "dfb_gemm: conv height incompatible with 128-pixel tiles");
      const int TB = 128 / (W * TH);
      kp.TH = TH;
      kp.TB = TB;
      kp.tiles_per_img = H / TH;
      kp.n_tiles_m = TB == 1 ? B * kp.tiles_per_img : (B + TB - 1) / TB;
      boxA[0] = GEMM_BLOCK_K; boxA[1] = (uint32_t)W; boxA[2] = (uint32_t)TH; boxA[3] = (uint32_t)TB;
    }
  } else {
    kp.n_tiles_m = (q->M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    boxA[0] = GEMM_BLOCK_K; boxA[1] = GEMM_BLOCK_M;
  }

  // CTA-pair mode (cluster of 2, cta_group::2 MMA): every problem with at least two waves of 256-row work items.
  // Measured per shape on B200 (profiles/r01_gemm_tile_tuning.md): the N = 640 / 1280 convs go from 0.99-1.43 to
  // 1.36-1.60 PFLOP/s; the HBM-bound short-K GEMMs are neutral once the remote tmem_empty arrive is .relaxed.
  // q->cta_group / DFB_GEMM_CTA_GROUP force either kernel.
  long long k_total = 0;
  for (int s = 0; s < q->nseg; ++s) k_total += (long long)q->ntaps[s] * q->a_c[s];
  // Wide tile for 256 < N <= 512 (the UNet's N = 320 convs): ONE tile of all N columns as two MMA sub-tiles that share
  // the A tile in shared memory (half the A loads and producer work per flop of the 2 x 160 tiling).  Its accumulators
  // take 2 * N / 2 ... N <= 512 columns, i.e. the ring is one deep and the epilogue no longer overlaps the next main
  // loop — worth it only for deep problems (K >= 2048: the epilogue is ~5 % of the main loop).  DFB_GEMM_WIDE=0 disables.
  if (q->block_n <= 0 && wide_enabled() && q->N > GEMM_MAX_BLOCK_N && q->N <= 2 * GEMM_MAX_BLOCK_N && q->N % 64 == 0 &&
      k_total >= 2048 && (kp.n_tiles_m + 1) / 2 >= num_sms())
    bn = q->N;
  kp.block_n = bn;
  kp.n_sub = bn > GEMM_MAX_BLOCK_N ? 2 : 1;
  kp.acc_ring = 2 * bn <= GEMM_TMEM_COLS ? 2 : 1;
  kp.n_tiles_n = (q->N + bn - 1) / bn;
  const int n_pair_items = ((kp.n_tiles_m + 1) / 2) * kp.n_tiles_n;
  int cta_group = q->cta_group > 0 ? q->cta_group : cta_group_override();
  if (cta_group == 0) cta_group = (kp.n_tiles_m >= 2 && n_pair_items >= num_sms()) ? 2 : 1;
  const bool pair = cta_group == 2;
  // smem ring: stage = A tile (16 KB) + this CTA's B tile (block_n x 128 B, half of that in a CTA pair; a multiple of
  // 1 KB because block_n % 32 == 0, which keeps every operand 1024-byte aligned for the 128-byte swizzle).
  kp.stage_bytes = GEMM_A_BYTES + (pair ? bn / 2 : bn) * GEMM_BLOCK_K * 2;
  {
    const int budget = GEMM_SMEM_BYTES - GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES - 1024 - 256;
    int st = budget / kp.stage_bytes;
    // 1-CTA kernel: depth 4 (deeper rings fit for narrow tiles but measured no better, profiles/r01_gemm_tile_tuning.md).
    // CTA pair: the half-size B tiles leave room for 5 (block_n 256) to 8 stages, and with two producer warps the deeper
    // ring pays (268-270 vs 273 ms per step at depth 8 vs 4).  DFB_GEMM_STAGES / DFB_GEMM_STAGES_PAIR override the caps.
    const int cap = pair ? (q_stages_pair_override() > 0 ? q_stages_pair_override() : 8) : (q_stages_override() > 0 ? q_stages_override() : 4);
    st = st < cap ? st : cap;
    kp.stages = st > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : (st < 2 ? 2 : st);
  }

  int kp_total = 0;
  for (int s = 0; s < q->nseg; ++s) {
    DFB_REQUIRE(q->a[s] != nullptr, "dfb_gemm: null A segment");
    DFB_REQUIRE(q->ntaps[s] >= 1 && q->ntaps[s] <= 9, "dfb_gemm: 1..9 taps per segment");
    DFB_REQUIRE(q->a_c[s] > 0, "dfb_gemm: empty A segment");
    DFB_REQUIRE((reinterpret_cast<uintptr_t>(q->a[s]) & 15) == 0 && (q->a_ld[s] % 8) == 0, "dfb_gemm: A must be 16B aligned with a_ld % 8 == 0");
    kp.seg_ntaps[s] = q->ntaps[s];
    kp.seg_ncblk[s] = (q->a_c[s] + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    int max_coff = 0;
    for (int t = 0; t < q->ntaps[s]; ++t) {
      kp.tap_dh[s * 9 + t] = q->tap_dh[s][t];
      kp.tap_dw[s * 9 + t] = q->tap_dw[s][t];
      kp.tap_coff[s * 9 + t] = q->tap_coff[s][t];
      DFB_REQUIRE(q->tap_coff[s][t] >= 0, "dfb_gemm: negative channel offset");
      if (q->tap_coff[s][t] > max_coff) max_coff = q->tap_coff[s][t];
      if (!kp.conv) DFB_REQUIRE(q->tap_dh[s][t] == 0 && q->tap_dw[s][t] == 0, "dfb_gemm: shifts need conv addressing");
    }
    DFB_REQUIRE(max_coff == 0 || (q->a_c[s] % GEMM_BLOCK_K) == 0, "dfb_gemm: channel offsets need a_c % 64 == 0");
    const uint64_t cext = (uint64_t)max_coff + (uint64_t)q->a_c[s];
    DFB_REQUIRE(cext <= (uint64_t)q->a_ld[s], "dfb_gemm: channel extent exceeds the row pitch");
    kp_total += q->ntaps[s] * kp.seg_ncblk[s] * GEMM_BLOCK_K;
    int rc;
    if (kp.conv) {
      uint64_t dims[4] = {cext, (uint64_t)q->W, (uint64_t)q->H, (uint64_t)q->B};
      uint64_t str[3] = {(uint64_t)q->a_ld[s] * 2, (uint64_t)q->a_ld[s] * 2 * q->W, (uint64_t)q->a_ld[s] * 2 * q->W * q->H};
      rc = make_tmap(&maps.a[s], q->a[s], 2, 4, dims, str, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      uint64_t dims[2] = {cext, (uint64_t)q->M};
      uint64_t str[1] = {(uint64_t)q->a_ld[s] * 2};
      rc = make_tmap(&maps.a[s], q->a[s], 2, 2, dims, str, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc != DFB_OK) return rc;
  }
  DFB_REQUIRE(kp_total <= q->w_ld, "dfb_gemm: packed weight K extent smaller than the A operand implies");
  {
    uint64_t dims[2] = {(uint64_t)q->w_ld, (uint64_t)q->N};
    uint64_t str[1] = {(uint64_t)q->w_ld * 2};
    const int sub_n = bn / kp.n_sub;
    uint32_t box[2] = {GEMM_BLOCK_K, (uint32_t)(pair ? sub_n / 2 : sub_n)};  // per MMA sub-tile; CTA pair: each CTA loads half of it
    int rc = make_tmap(&maps.b, q->w, 2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != DFB_OK) return rc;
  }

  kp.bias = q->bias;
  kp.rowbias = q->rowbias;
  kp.rowbias_ld = q->rowbias_ld;
  kp.rows_per_batch = q->rows_per_batch > 0 ? q->rows_per_batch : 1;
  kp.residual = q->residual;
  kp.res_ld = q->res_ld;
  kp.res_fp32 = q->res_dtype == DFB_DTYPE_F32;
  kp.out = q->out;
  kp.out_ld = q->out_ld;
  kp.out_fp32 = q->out_dtype == DFB_DTYPE_F32;
  kp.geglu = q->geglu ? 1 : 0;
  kp.act = q->act;
  kp.gn_partial = q->gn_partial;
  kp.up_w = 0; kp.up_a = 0; kp.up_b = 0; kp.up_blk = 1;
  if (q->up2x != 0) {
    DFB_REQUIRE(q->up2x >= 1 && q->up2x <= 4, "dfb_gemm: up2x must be 0 or 1 + 2a + b");
    DFB_REQUIRE(q->conv && !q->geglu && q->residual == nullptr && q->rowbias == nullptr,
                "dfb_gemm: up2x needs conv geometry and takes no residual / rowbias / GEGLU");
    DFB_REQUIRE(!q->gn_partial || (q->H * q->W) % 32 == 0, "dfb_gemm: up2x with gn_partial needs H*W % 32 == 0");
    kp.up_w = q->W;
    kp.up_a = (q->up2x - 1) >> 1;
    kp.up_b = (q->up2x - 1) & 1;
    kp.up_blk = (q->H * q->W) / 32 > 0 ? (q->H * q->W) / 32 : 1;
  }
  if (q->gn_partial)
    DFB_REQUIRE(q->out_dtype == DFB_DTYPE_F32 && !q->geglu && q->M % 32 == 0 && q->N % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(q->gn_partial) & 15) == 0,
                "dfb_gemm: gn_partial needs fp32 output, M % 32 == 0, N % 4 == 0, 16B-aligned buffer");
  if (kp.geglu) DFB_REQUIRE(q->N % 32 == 0 && q->residual == nullptr, "dfb_gemm: GEGLU needs N % 32 == 0 and no residual");
  const int out_elem = kp.out_fp32 ? 4 : 2;
  bool vec_ok = ((reinterpret_cast<uintptr_t>(q->out) & 15) == 0) && (((size_t)q->out_ld * out_elem) % 16 == 0);
  if (q->residual) {
    const int res_elem = kp.res_fp32 ? 4 : 2;
    vec_ok = vec_ok && ((reinterpret_cast<uintptr_t>(q->residual) & 15) == 0) && (((size_t)q->res_ld * res_elem) % 16 == 0);
  }
  kp.vec_ok = vec_ok ? 1 : 0;

  // epilogue specialisation
  int mode = EPI_GENERIC;
  // (an upsample phase takes the specialised path in its one shape: fp32 output, bias only, optional GroupNorm partials)
  const bool up_ok = q->up2x == 0 || (kp.out_fp32 && q->residual == nullptr && q->rowbias == nullptr && !kp.geglu && (q->W & (q->W - 1)) == 0);
  const bool simple = kp.vec_ok && (q->N % 4) == 0 && kp.act == 0 && up_ok && (q->residual == nullptr || kp.res_fp32) &&
                      (q->bias == nullptr || (reinterpret_cast<uintptr_t>(q->bias) & 15) == 0) &&
                      (q->rowbias == nullptr || (kp.rows_per_batch % 32 == 0 && (q->rowbias_ld % 4) == 0 &&
                                                 (reinterpret_cast<uintptr_t>(q->rowbias) & 15) == 0));
  if (simple) {
    if (kp.geglu) {
      if (!kp.out_fp32) mode = EPI_GEGLU;
    } else {
      mode = (kp.out_fp32 ? EPI_OUT_F32 : 0) | (q->residual ? EPI_RES_F32 : 0) | (q->rowbias ? EPI_ROWBIAS : 0);
      // instantiated combinations only; anything else runs the generic kernel
      if (!(mode == 0 || mode == EPI_RES_F32 || mode == EPI_OUT_F32 || mode == (EPI_OUT_F32 | EPI_RES_F32) ||
            mode == (EPI_OUT_F32 | EPI_ROWBIAS)))
        mode = EPI_GENERIC;
    }
  }
  // TMA-store epilogue (cp.async.bulk.tensor shared -> global) for the bf16-output GEMMs: bias-only (mode 0: q|k|v, q, K/V
  // projections) and GEGLU.  DFB_GEMM_TMA_STORE=0 keeps the transposing st.global epilogue.
  kp.tma_out = 0;
  if ((mode == 0 || mode == EPI_GEGLU) && !kp.out_fp32 && q->up2x == 0 && tma_store_enabled()) {
    const int n_out = kp.geglu ? q->N / 2 : q->N;
    uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)q->M};
    uint64_t str[1] = {(uint64_t)q->out_ld * 2};
    uint32_t box[2] = {kp.geglu ? 16u : 32u, 32u};
    int rc = make_tmap(&maps.c, q->out, 2, 2, dims, str, box, kp.geglu ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != DFB_OK) return rc;
    kp.tma_out = 1;
  }
  const int num_tiles = kp.n_tiles_m * kp.n_tiles_n;
  int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  if (pair) {
    const int clusters = n_pair_items < num_sms() / 2 ? n_pair_items : num_sms() / 2;
    grid = 2 * clusters;
  }
  static bool attr_done[64] = {false};
  int dev = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = 2;
  lattr[0].val.clusterDim.y = 1;
  lattr[0].val.clusterDim.z = 1;
  lc.gridDim = dim3((unsigned)grid, 1, 1);
  lc.dynamicSmemBytes = GEMM_SMEM_BYTES;
  lc.stream = stream;
  lc.attrs = lattr;
  lc.numAttrs = 1;
#define DFB_GEMM_CASE(M_)                                                                                        \
  case M_: {                                                                                                     \
    constexpr int EW_ = ((M_) == EPI_GEGLU) ? 16 : GEMM_EPI_WARPS;                                                 \
    if (first) {                                                                                                 \
      DFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<(M_), EW_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES)); \
      DFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<(M_), EW_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));  \
    }                                                                                                            \
    if (launch && !pair) gemm_tcgen05_kernel<(M_), EW_, false><<<grid, 96 + 32 * EW_, GEMM_SMEM_BYTES, stream>>>(maps, kp); \
    if (launch && pair) {                                                                                        \
      lc.blockDim = dim3(96 + 32 * EW_, 1, 1);                                                                   \
      DFB_CHECK_CUDA(cudaLaunchKernelEx(&lc, gemm_tcgen05_kernel<(M_), EW_, true>, maps, kp));                   \
    }                                                                                                            \
  } break;
  const bool need_attr = dev >= 0 && dev < 64 && !attr_done[dev];
  for (int pass = need_attr ? 0 : 1; pass < 2; ++pass) {
    const bool first = pass == 0, launch = pass == 1;
    const int modes[7] = {0, EPI_RES_F32, EPI_OUT_F32, EPI_OUT_F32 | EPI_RES_F32, EPI_OUT_F32 | EPI_ROWBIAS, EPI_GEGLU, EPI_GENERIC};
    for (int i = 0; i < (first ? 7 : 1); ++i) {
      switch (first ? modes[i] : mode) {
        DFB_GEMM_CASE(0)
        DFB_GEMM_CASE(EPI_RES_F32)
        DFB_GEMM_CASE(EPI_OUT_F32)
        DFB_GEMM_CASE(EPI_OUT_F32 | EPI_RES_F32)
        DFB_GEMM_CASE(EPI_OUT_F32 | EPI_ROWBIAS)
        DFB_GEMM_CASE(EPI_GEGLU)
        DFB_GEMM_CASE(EPI_GENERIC)
        default:
          set_last_error_msg("dfb_gemm: internal: epilogue mode not instantiated");
          return DFB_ERR_INVALID;
      }
    }
  }
#undef DFB_GEMM_CASE
  if (need_attr) attr_done[dev] = true;
  DFB_CHECK_CUDA(cudaGetLastError());
  return DFB_OK;
}
