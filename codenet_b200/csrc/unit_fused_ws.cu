// A stride-1 ShuffleNetV2 unit with 116 / 122-channel halves as ONE kernel, WARP-SPECIALISED: the same arithmetic and the same
// phase bodies as unit_fused.cu (see there for the design), but the phases of CONSECUTIVE tiles run side by side on different
// warps of one CTA per SM instead of one after the other on all of them:
//
//   control warp (one elected lane)   TMA load of the branch tile A1(t)  ->  G1(t): tcgen05 pw1 into TMEM acc1
//                                     ... -> G2(t): tcgen05 pw3 into TMEM acc2[t & 1] as soon as the stencil of tile t is done
//   E1 group, 8 warps                 E1(t): acc1 -> requantised int8 `mid[t & 1]` in shared memory
//   S group, 8 warps                  S(t): depthwise stencil mid[t & 1] -> A tile a2[t & 1] of G2
//   E2 group, 8 warps                 E2(t): acc2[t & 1] + pass-through bytes -> interleaved output in the staging segments -> TMA
//                                     store; requests the pass-through tile of tile t + 2
//
// so E1(t + 2), S(t + 1) and E2(t) run at the same time: 24 warps with independent instruction streams instead of 16 that move from
// phase to phase together (a warp of these latency-bound phases issues one instruction per ~9 clocks whatever else runs), nobody waits
// for the tensor core with nothing else runnable, and there is no CTA-wide barrier on the per-tile path: hand-overs are mbarriers
// (acc1 drained, mid full / free, A tile full, acc2 drained, MMA done, TMA landed) and one named barrier inside the E2 group around
// its TMA store.  In unit_fused.cu two CTAs per SM reach 35 % issue utilisation on this stage.
//
// Shared memory (one CTA per SM): A1 23 KB, mid 2 x 23 KB, A tile 2 x 16 KB, staging 32 KB, pass-through tile 2 x 16 KB, both weight
// matrices 32 KB, constants 4 KB.  TMEM: acc1 256 columns (two M = 128 blocks), acc2 2 x 128 columns.
#include "unit_fused.cuh"
#include <algorithm>

#define UW_THREADS 800                         // 8 E1 warps + 8 stencil warps + 8 E2 warps + the control warp
#define UW_GROUP 256
#define UW_HP 128
#define UW_TW 16
#define UW_TH 8
#define UW_IW (UW_TW + 2)
#define UW_IH (UW_TH + 2)
#define UW_PIX1 (UW_IW * UW_IH)                // 180 rows of the first GEMM
#define UW_BLK1 64                             // A row at which its second M = 128 block starts
#define UW_A1_BYTES (((UW_PIX1 * 128) + 1023) & ~1023)
#define UW_SMEM_LIMIT (227 * 1024)

struct UwParams {
  int H, W, tiles_x, tiles_y; unsigned ntiles;
  int txs, tys;
  uint32_t off_mid, off_a2, off_stg, off_pass, off_w1, off_w3, off_kc1, off_kc3, off_bar;
  const int8_t* w1; const int8_t* w3;
  const int4* kc1; const int4* kc3;
  const uint32_t* wpk; const int4* ki;
  uint32_t pad_word;
  int8_t* dump_c1; int8_t* dump_d2;
};

__device__ __forceinline__ void uw_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

template <int PG>
__global__ void __launch_bounds__(UW_THREADS, 1)
unit_fused_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmP,
                     const __grid_constant__ CUtensorMap tmO, const UwParams p) {
  constexpr int HP = UW_HP;
  pdl_launch_dependents();
  extern __shared__ uint8_t uw_smem_raw[];
  const uint32_t sbase = smem_u32(uw_smem_raw) + ((1024u - (smem_u32(uw_smem_raw) & 1023u)) & 1023u);
  const uint32_t s_a1 = sbase, s_mid0 = sbase + p.off_mid, s_a2 = sbase + p.off_a2, s_stg = sbase + p.off_stg, s_pass = sbase + p.off_pass;
  const uint32_t s_w1 = sbase + p.off_w1, s_w3 = sbase + p.off_w3, s_kc1 = sbase + p.off_kc1, s_kc3 = sbase + p.off_kc3;
  const uint32_t mid_stride = (uint32_t)((UW_PIX1 * HP + 1023) & ~1023);
  // barriers (8 bytes each): 0 a1_full, 1..2 pass_full[2], 3 g1_done, 4..5 g2_done[2], 6 acc1_free, 7..8 mid_full[2], 9..10 mid_free[2],
  // 11..12 a2_full[2], 13..14 acc2_free[2]; then the TMEM slot
  const uint32_t bars = sbase + p.off_bar;
  const uint32_t B_A1 = bars, B_PASS = bars + 8, B_G1 = bars + 24, B_G2 = bars + 32, B_ACC1 = bars + 48, B_MIDF = bars + 56, B_MIDE = bars + 72,
                 B_A2 = bars + 88, B_ACC2 = bars + 104, tmem_slot = bars + 120;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto tile_coords = [&](unsigned tile, int& tx, int& ty, int& b) {
    if (p.txs >= 0) {
      tx = (int)(tile & (unsigned)(p.tiles_x - 1)); tile >>= p.txs;
      ty = (int)(tile & (unsigned)(p.tiles_y - 1)); b = (int)(tile >> p.tys);
    } else {
      tx = (int)(tile % (unsigned)p.tiles_x); tile /= (unsigned)p.tiles_x;
      ty = (int)(tile % (unsigned)p.tiles_y); b = (int)(tile / (unsigned)p.tiles_y);
    }
  };
  if (tid == 0) {
    mbar_init(B_A1, 1); mbar_init(B_PASS, 1); mbar_init(B_PASS + 8, 1); mbar_init(B_G1, 1); mbar_init(B_G2, 1); mbar_init(B_G2 + 8, 1);
    mbar_init(B_ACC1, UW_GROUP); mbar_init(B_MIDF, UW_GROUP); mbar_init(B_MIDF + 8, UW_GROUP);
    mbar_init(B_MIDE, UW_GROUP); mbar_init(B_MIDE + 8, UW_GROUP); mbar_init(B_A2, UW_GROUP); mbar_init(B_A2 + 8, UW_GROUP);
    mbar_init(B_ACC2, UW_GROUP); mbar_init(B_ACC2 + 8, UW_GROUP);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
  }
  if (warp == 24) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 256 * 8; i += UW_THREADS) {
    const bool second = i >= 128 * 8;
    const int j = second ? i - 128 * 8 : i, n = j >> 3, c = j & 7;
    const uint4 v = __ldg((const uint4*)((second ? p.w3 : p.w1) + (size_t)n * 128) + c);
    sts_u128((second ? s_w3 : s_w1) + (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) << 4), v.x, v.y, v.z, v.w);
  }
  for (int i = tid; i < 256; i += UW_THREADS) {
    const bool second = i >= 128;
    const int4 v = __ldg(second ? p.kc3 + (i - 128) : p.kc1 + i);
    sts_u128(second ? s_kc3 + 16u * (uint32_t)(i - 128) : s_kc1 + 16u * (uint32_t)i, (uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                  // everything above is constant; activations need the previous grid

  if (warp == 24) {
    // ===================== control: TMA loads of A1, both GEMMs =====================
    if (uf_elect()) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t d_a1 = make_smem_desc(s_a1), d_a1b = make_smem_desc(s_a1 + UW_BLK1 * 128u), d_w1 = make_smem_desc(s_w1);
      const uint64_t d_a2[2] = {make_smem_desc(s_a2), make_smem_desc(s_a2 + 16384u)}, d_w3 = make_smem_desc(s_w3);
      auto load_a1 = [&](unsigned tile) {
        int tx, ty, b; tile_coords(tile, tx, ty, b);
        mbar_expect_tx(B_A1, UW_PIX1 * 128u);
        tma_load_4d(s_a1, &tmA, HP, tx * UW_TW - 1, ty * UW_TH - 1, b, B_A1);
      };
      auto issue_g1 = [&](uint32_t it) {       // tile `it` of this CTA: A1 landed, acc1 drained by E1(it - 1)
        uf_wait(B_A1, it & 1u);
        if (it > 0) uf_wait(B_ACC1, (it - 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int k = 0; k < HP / 32; ++k)
            umma_i8(tmem_base + (uint32_t)(blk * 128), (blk ? d_a1b : d_a1) + (uint64_t)(2 * k), d_w1 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit(B_G1);
      };
      unsigned tile = blockIdx.x;
      if (tile < p.ntiles) { load_a1(tile); issue_g1(0); }
      uint32_t it = 0;
      for (; tile < p.ntiles; tile += gridDim.x, ++it) {
        const bool has_next = (unsigned long long)tile + gridDim.x < p.ntiles;
        uf_wait(B_G1, it & 1u);                // G1(it) done: A1 has been consumed
        if (has_next) { load_a1(tile + gridDim.x); issue_g1(it + 1); }
        // G2(it): the stencil has filled the A tile a2[it & 1]; the accumulator slot was drained by E2(it - 2)
        uf_wait(B_A2 + 8u * (it & 1u), (it >> 1) & 1u);
        if (it >= 2) uf_wait(B_ACC2 + 8u * (it & 1u), ((it - 2) >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HP / 32; ++k)
          umma_i8(tmem_base + 256u + 128u * (it & 1u), d_a2[it & 1u] + (uint64_t)(2 * k), d_w3 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit(B_G2 + 8u * (it & 1u));
      }
    }
  } else if (warp < 8) {
    // ===================== E1 group =====================
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t pad = p.pad_word;
    uint32_t it = 0;
    for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      int tx, ty, b; tile_coords(tile, tx, ty, b);
      const uint32_t slot = it & 1u, s_mid = s_mid0 + slot * mid_stride;
      uf_wait(B_G1, it & 1u);
      if (it >= 2) uf_wait(B_MIDE + 8u * slot, ((it - 2) >> 1) & 1u);       // the stencil of tile it - 2 has read this `mid` buffer
      tc_fence_after();
      uint32_t taddr[2], mpix[2], mswz[2]; bool valid[2], inside[2]; int8_t* dump[2];
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int row = blk * UW_BLK1 + q * 32 + lane;
        const int r = (row * 3641) >> 16, c = row - r * UW_IW;
        valid[blk] = blk == 0 || (row >= 128 && row < UW_PIX1);
        inside[blk] = (unsigned)(ty * UW_TH - 1 + r) < (unsigned)p.H && (unsigned)(tx * UW_TW - 1 + c) < (unsigned)p.W;
        taddr[blk] = tmem_base + (uint32_t)(blk * 128) + ((uint32_t)(q * 32) << 16);
        mpix[blk] = s_mid + (uint32_t)row * (uint32_t)HP; mswz[blk] = uf_mid_swz<HP>((uint32_t)c) << 4;
        dump[blk] = (p.dump_c1 && valid[blk] && r >= 1 && r <= UW_TH && c >= 1 && c <= UW_TW)
                        ? p.dump_c1 + (((size_t)b * p.H + (ty * UW_TH - 1 + r)) * p.W + (tx * UW_TW - 1 + c)) * HP : nullptr;
      }
      if (q < 2) {
        const uint32_t t1[1] = {taddr[0]}, m1[1] = {mpix[0]}, s1[1] = {mswz[0]}; const bool v1[1] = {true}, i1[1] = {inside[0]};
        int8_t* const d1[1] = {dump[0]};
        uf_e1<HP, true, 1>(t1, m1, s1, v1, i1, hf, s_kc1, -128, pad, d1);
      } else {
        uf_e1<HP, true, 2>(taddr, mpix, mswz, valid, inside, hf, s_kc1, -128, pad, dump);
      }
      tc_fence_before();
      uw_arrive(B_ACC1);                       // acc1 drained: G1 of the next tile may run
      uw_arrive(B_MIDF + 8u * slot);           // mid[slot] complete
    }
  } else if (warp < 16) {
    // ===================== stencil group: S(it): mid[it & 1] -> a2[it & 1] =====================
    const int gt = tid - UW_GROUP;
    constexpr int CW = HP / 4;                 // 32 channel words x 8 pixel pairs = the group's 256 threads, 8 rows each
    const int cw = gt % CW, pg = gt / CW;
    uint32_t Wt[4][6]; int2 km[4]; long long kb[4];
    {
      const uint4* wv = (const uint4*)(p.wpk + (size_t)cw * 24);
      uint32_t flat[24];
#pragma unroll
      for (int i = 0; i < 6; ++i) { const uint4 v = __ldg(wv + i); flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w; }
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 6; ++i) Wt[c][i] = flat[c * 6 + i];
#pragma unroll
      for (int c = 0; c < 4; ++c) { km[c] = __ldg((const int2*)(p.ki + cw * 4 + c)); kb[c] = __ldg((const long long*)(p.ki + cw * 4 + c) + 1); }
    }
    uint32_t mo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t c = (uint32_t)(2 * pg + j);
      mo[j] = uf_mid_off<HP>(c, c, (uint32_t)(cw >> 2)) + (uint32_t)((cw & 3) * 4);
    }
    uint32_t ao[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t m = (uint32_t)(2 * pg + j);
      ao[j] = s_a2 + m * 128u + ((((uint32_t)cw >> 2) ^ (m & 7u)) << 4) + (uint32_t)((cw & 3) * 4);
    }
    uint32_t it = 0;
    for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      int tx, ty, b; tile_coords(tile, tx, ty, b);
      const uint32_t slot = it & 1u, s_mid = s_mid0 + slot * mid_stride, a_off = slot * 16384u;
      uf_wait(B_MIDF + 8u * slot, (it >> 1) & 1u);
      if (it >= 2) uf_wait(B_G2 + 8u * slot, ((it - 2) >> 1) & 1u);       // G2(it - 2) has consumed a2[slot]
      auto read_row = [&](int mr, uint32_t (&T)[4]) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = lds_u32(s_mid + mo[j] + (uint32_t)(mr * UW_IW * HP));
        transpose4x4(w[0], w[1], w[2], w[3], T[0], T[1], T[2], T[3]);
      };
      uint32_t Tm[4], Tc[4], Tp[4];
      read_row(0, Tm);
      read_row(1, Tc);
#pragma unroll
      for (int r = 0; r < UW_TH; ++r) {
        read_row(r + 2, Tp);
        int a0[4], a1[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a0[c] = dp4a_ss(Tp[c], Wt[c][4], dp4a_ss(Tc[c], Wt[c][2], dp4a_ss(Tm[c], Wt[c][0], 0)));
          a1[c] = dp4a_ss(Tp[c], Wt[c][5], dp4a_ss(Tc[c], Wt[c][3], dp4a_ss(Tm[c], Wt[c][1], 0)));
        }
        const uint32_t o0 = uf_rq_word<true>(a0, km, kb), o1 = uf_rq_word<true>(a1, km, kb);
        sts_u32(a_off + ao[0] + (uint32_t)(r * UW_TW * 128), o0);
        sts_u32(a_off + ao[1] + (uint32_t)(r * UW_TW * 128), o1);
        if (p.dump_d2) {
          uint32_t* d = (uint32_t*)(p.dump_d2 + (((size_t)b * p.H + (ty * UW_TH + r)) * p.W + (tx * UW_TW + 2 * pg)) * HP) + cw;
          d[0] = o0; d[HP / 4] = o1;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { Tm[c] = Tc[c]; Tc[c] = Tp[c]; }
      }
      fence_async_smem();                      // A tile (generic proxy) -> tensor core (async proxy)
      uw_arrive(B_A2 + 8u * slot);
      uw_arrive(B_MIDE + 8u * slot);
    }
  } else {
    // ===================== E2 group: acc2[it & 1] + pass[it & 1] -> staging -> TMA store =====================
    const int gw = warp - 16;
    const int q = gw & 3, hf = gw >> 2;
    const bool leader = gw == 0 && uf_elect(); // issues the group's TMA traffic (store, pass-through loads) and waits for it
    auto load_pass = [&](unsigned tile, uint32_t slot) {
      int tx, ty, b; tile_coords(tile, tx, ty, b);
      mbar_expect_tx(B_PASS + 8u * slot, 128u * 128u);
      tma_load_4d(s_pass + slot * 16384u, &tmP, 0, tx * UW_TW, ty * UW_TH, b, B_PASS + 8u * slot);
    };
    if (leader) {
      if (blockIdx.x < p.ntiles) load_pass(blockIdx.x, 0u);
      if ((unsigned long long)blockIdx.x + gridDim.x < p.ntiles) load_pass(blockIdx.x + gridDim.x, 1u);
    }
    uint32_t it = 0;
    for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      int tx, ty, b; tile_coords(tile, tx, ty, b);
      const uint32_t slot = it & 1u;
      uf_wait(B_G2 + 8u * slot, (it >> 1) & 1u);
      uf_wait(B_PASS + 8u * slot, (it >> 1) & 1u);
      tc_fence_after();
      if (leader) tma_store_wait_read0();      // the previous store has read the staging segments
      named_bar_sync(1, UW_GROUP);
      {
        const int m = q * 32 + lane;
        const uint32_t taddr = tmem_base + 256u + 128u * slot + ((uint32_t)(q * 32) << 16);
        const uint32_t prow = s_pass + slot * 16384u + (uint32_t)m * 128u, srow = s_stg + (uint32_t)m * 128u, x7 = (uint32_t)(m & 7);
        if (hf == 0) uf_e2_fast<HP, PG, 0>(taddr, prow, srow, x7 << 4, s_kc3);
        else uf_e2_fast<HP, PG, 1>(taddr, prow, srow, x7 << 4, s_kc3);
      }
      tc_fence_before();
      uw_arrive(B_ACC2 + 8u * slot);           // acc2[slot] drained
      fence_async_smem();                      // staging -> TMA store
      named_bar_sync(1, UW_GROUP);
      if (leader) {
        tma_store_4d(&tmO, 0, tx * UW_TW, ty * UW_TH, b, s_stg);
        tma_store_4d(&tmO, 128, tx * UW_TW, ty * UW_TH, b, s_stg + 16384u);
        tma_store_commit();
        if ((unsigned long long)tile + 2ull * gridDim.x < p.ntiles) load_pass(tile + 2u * gridDim.x, slot);   // pass[slot] has been consumed
      }
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 24) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool unit_fused_ws_ok(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, int x_pitch, int mid_pitch, int out_pitch, int H, int W) {
  if (mid_pitch != UW_HP || !unit_fused_ok(pw1, dw, pw3, x_pitch, mid_pitch, out_pitch, H, W)) return false;
  if (!pw1.sh0 || !pw3.sh0 || !dw.sh0 || pw1.rq.lo > -128 || pw3.rq.lo > -128) return false;
  if (pw3.il_hp != UW_HP || (pw3.il_pg != 58 && pw3.il_pg != 61) || pw1.BN != 128 || pw3.BN != 128) return false;
  return (g_cdn_debug_flags & (1u << 29)) == 0;                         // bit 29: the barrier-phased kernel of unit_fused.cu (A/B)
}

int unit_fused_ws_launch(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, const int8_t* x, int8_t* out,
                         int batch, int H, int W, int zx_mid, int8_t* dump_c1, int8_t* dump_d2, cudaStream_t st) {
  const int HP = UW_HP;
  CDN_CHECK(unit_fused_ws_ok(pw1, dw, pw3, 2 * HP, HP, 2 * HP, H, W), CDN_ERR_INVALID, "unit_fused_ws: layer triple not eligible");
  UwParams p; memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.tiles_x = W / UW_TW; p.tiles_y = H / UW_TH;
  const long long ntiles = (long long)batch * p.tiles_x * p.tiles_y;
  if (ntiles == 0) return 0;
  CDN_CHECK(ntiles < (1ll << 31) - 4 * 160, CDN_ERR_INVALID, "unit_fused_ws: tensor too large for 32-bit indexing");
  p.ntiles = (unsigned)ntiles;
  p.txs = p.tys = -1;
  if (!(p.tiles_x & (p.tiles_x - 1)) && !(p.tiles_y & (p.tiles_y - 1))) {
    p.txs = 0; while ((1 << p.txs) < p.tiles_x) ++p.txs;
    p.tys = 0; while ((1 << p.tys) < p.tiles_y) ++p.tys;
  }
  p.w1 = pw1.w; p.w3 = pw3.w; p.kc1 = (const int4*)pw1.kc; p.kc3 = (const int4*)pw3.kc;
  p.wpk = dw.wpk1; p.ki = (const int4*)dw.ki;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx_mid) * 0x01010101u;
  p.dump_c1 = dump_c1; p.dump_d2 = dump_d2;
  uint32_t o = UW_A1_BYTES;
  p.off_mid = o; o += 2u * (uint32_t)((UW_PIX1 * HP + 1023) & ~1023);
  p.off_a2 = o; o += 2u * 16384u;
  p.off_stg = o; o += 32768u;
  p.off_pass = o; o += 2u * 16384u;
  p.off_w1 = o; o += 16384u;
  p.off_w3 = o; o += 16384u;
  p.off_kc1 = o; o += 2048u;
  p.off_kc3 = o; o += 2048u;
  p.off_bar = o; o += 128u;
  const size_t smem = (size_t)o + 1024;
  CDN_CHECK(smem <= UW_SMEM_LIMIT, CDN_ERR_INVALID, "unit_fused_ws: %zu bytes of shared memory", smem);
  CUtensorMap tmA, tmP, tmO;
  if (int r = make_tmap_nhwc_swz(&tmA, x, (uint64_t)(2 * HP), (uint64_t)W, (uint64_t)H, (uint64_t)batch, 128, UW_IW, UW_IH)) return r;
  if (int r = make_tmap_nhwc_swz(&tmP, x, (uint64_t)(2 * HP), (uint64_t)W, (uint64_t)H, (uint64_t)batch, 128, UW_TW, UW_TH)) return r;
  if (int r = make_tmap_nhwc_swz(&tmO, out, (uint64_t)(2 * HP), (uint64_t)W, (uint64_t)H, (uint64_t)batch, 128, UW_TW, UW_TH)) return r;
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, UwParams) = pw3.il_pg == 58 ? unit_fused_ws_kernel<58> : unit_fused_ws_kernel<61>;
  static bool attr_set[2][64] = {};
  if (cdn_first_on_device(attr_set[pw3.il_pg == 58 ? 0 : 1])) {
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UW_SMEM_LIMIT));
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const unsigned blocks = (unsigned)std::min<long long>(ntiles, (long long)cdn_num_sms());
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(UW_THREADS); cfg.stream = st; cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CDN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmP, tmO, p));
  CDN_LAUNCH_CHECK("unit_fused_ws_kernel");
  return 0;
}
