// Branch 2 of the FIRST stride-2 ShuffleNetV2 unit (layer1.0 of CoDeNet1x: QuantBaseNode.forward with stride 2,
// quant_modules.py:878-907 / shufflenetv2_dcn.py:57-114) as ONE kernel: conv 1x1 (24 -> 58) + BN + ReLU + QuantAct at FULL
// resolution -> depthwise 3x3 stride 2 + BN + QuantAct -> conv 1x1 + BN + ReLU + QuantAct -> cat with branch 1 + channel_shuffle.
//
// As three launches this is the most expensive group of the network (0.35 of 3.0 ms at batch 256): the first conv writes a
// [B,128,128,64] tensor (268 MB) that the depthwise conv reads back.  Here that tensor lives in shared memory one tile at a time.
//
// Same structure as unit_fused.cu (see there); what differs:
//   * a tile of 8 x 16 OUTPUT pixels needs 17 x 33 input pixels (rows 2y-1 .. 2y+15): 561 rows of the first GEMM = five M = 128
//     blocks.  K = 32 bytes per pixel (24 channels): the input tile arrives by ONE 4-D TMA load with the 32-byte swizzle and is the
//     A operand as it lands (18 KB); the weights use the same 32-byte K-major layout
//   * TMEM holds four blocks at a time (4 x 64 columns, two CTAs per SM): block 4 (rows 512..560) is issued into block 0's columns
//     as soon as those are drained and completes behind the epilogue of blocks 1..3
//   * the stencil is the stride-2 one of dw_tma.cu: a thread owns one channel word of one output column and walks the 17 rows
//   * the pass-through operand of the interleaving epilogue is branch 1's output (its own tensor), fetched by TMA like x1 there
#include "unit_fused.cuh"
#include <algorithm>

#define US_THREADS 256
#define US_TW 16
#define US_TH 8
#define US_IW (2 * US_TW + 1)                 // 33
#define US_IH (2 * US_TH + 1)                 // 17
#define US_PIX (US_IW * US_IH)                // 561 rows of the first GEMM
#define US_MW 34                              // pixels per row of the `mid` tile (even: a row advances by a multiple of 128 bytes)
#define US_HP 64
#define US_A1_BYTES (640 * 32)                // five blocks of 128 rows x 32 bytes (the TMA writes 561 rows)
#define US_MID_BYTES ((US_IH * US_MW * US_HP + 1023) & ~1023)
#define US_SMEM_LIMIT (227 * 1024)
extern unsigned long long* g_pw_dbg;

struct UsParams {
  int Hin, Win, Ho, Wo, tiles_x, tiles_y; unsigned ntiles;
  int txs, tys;
  uint32_t off_mid, off_a2, off_pass, off_w1, off_w3, off_kc1, off_kc3, off_bar;
  const int8_t* w1; const int8_t* w3;        // [64][128] integer weights (K padded to 128; the first conv uses 32 bytes of a row)
  const int4* kc1; const int4* kc3;
  const uint32_t* wpk; const int4* ki;       // depthwise conv, stride-2 packing [channel][3]
  uint32_t pad_word;
  int8_t* dump_c1; int8_t* dump_d2;
};

// K-major operand with 32-byte rows and the 32-byte swizzle: 8-row groups of 256 bytes
__device__ __forceinline__ uint64_t us_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                    // SWIZZLE_32B
  return d;
}

template <int PG>
__global__ void __launch_bounds__(US_THREADS, 2)
unit_s2_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmP,
                     const __grid_constant__ CUtensorMap tmO, const UsParams p) {
  constexpr int HP = US_HP;
  pdl_launch_dependents();
  extern __shared__ uint8_t us_smem_raw[];
  const uint32_t sbase = smem_u32(us_smem_raw) + ((1024u - (smem_u32(us_smem_raw) & 1023u)) & 1023u);
  const uint32_t s_a1 = sbase, s_mid = sbase + p.off_mid, s_a2 = sbase + p.off_a2, s_pass = sbase + p.off_pass;
  const uint32_t s_w1 = sbase + p.off_w1, s_w3 = sbase + p.off_w3, s_kc1 = sbase + p.off_kc1, s_kc3 = sbase + p.off_kc3;
  const uint32_t bar_a = sbase + p.off_bar, bar_p = bar_a + 8, bar_m = bar_a + 16, tmem_slot = bar_a + 24;
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
  auto load_a1 = [&](unsigned tile) {
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    mbar_expect_tx(bar_a, US_PIX * 32u);
    tma_load_4d(s_a1, &tmA, 0, 2 * tx * US_TW - 1, 2 * ty * US_TH - 1, b, bar_a);
  };
  auto load_pass = [&](unsigned tile) {
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    mbar_expect_tx(bar_p, 128u * 128u);
    tma_load_4d(s_pass, &tmP, 0, tx * US_TW, ty * US_TH, b, bar_p);
  };

  if (tid == 0) {
    mbar_init(bar_a, 1); mbar_init(bar_p, 1); mbar_init(bar_m, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // first conv's weights: 64 rows x 32 bytes, 32-byte swizzle (16-byte chunk ^= bit 2 of the row); second conv's: 128-byte rows
  for (int i = tid; i < 64 * 2; i += US_THREADS) {
    const int n = i >> 1, c = i & 1;
    const uint4 v = __ldg((const uint4*)(p.w1 + (size_t)n * 128) + c);
    sts_u128(s_w1 + (uint32_t)n * 32u + (uint32_t)((c ^ ((n >> 2) & 1)) << 4), v.x, v.y, v.z, v.w);
  }
  for (int i = tid; i < 64 * 8; i += US_THREADS) {
    const int n = i >> 3, c = i & 7;
    const uint4 v = __ldg((const uint4*)(p.w3 + (size_t)n * 128) + c);
    sts_u128(s_w3 + (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) << 4), v.x, v.y, v.z, v.w);
  }
  for (int i = tid; i < 128; i += US_THREADS) {
    const bool second = i >= 64;
    const int4 v = __ldg(second ? p.kc3 + (i - 64) : p.kc1 + i);
    sts_u128(second ? s_kc3 + 16u * (uint32_t)(i - 64) : s_kc1 + 16u * (uint32_t)i, (uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w);
  }
  // stencil role: channel word cw of output column ox
  const int cw = tid & 15, ox = tid >> 4;
  uint32_t Wt[4][3]; int2 km[4]; long long kb[4];
  {
    const uint4* wv = (const uint4*)(p.wpk + (size_t)cw * 12);
    uint32_t flat[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const uint4 v = __ldg(wv + i); flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w; }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 3; ++i) Wt[c][i] = flat[c * 3 + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) { km[c] = __ldg((const int2*)(p.ki + cw * 4 + c)); kb[c] = __ldg((const long long*)(p.ki + cw * 4 + c) + 1); }
  }
  uint32_t mo[3];                              // `mid` offsets of this thread's three input pixels in tile row 0
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const uint32_t c = (uint32_t)(2 * ox + j);
    mo[j] = s_mid + uf_mid_off<HP>(c, c, (uint32_t)(cw >> 2)) + (uint32_t)((cw & 3) * 4);
  }
  const uint32_t ao = s_a2 + (uint32_t)ox * 128u + ((((uint32_t)cw >> 2) ^ ((uint32_t)ox & 7u)) << 4) + (uint32_t)((cw & 3) * 4);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int q = warp & 3, hf = warp >> 2;
  const uint32_t pad = p.pad_word;
  const uint64_t d_a1 = us_desc_sw32(s_a1), d_w1 = us_desc_sw32(s_w1), d_a2 = make_smem_desc(s_a2), d_w3 = make_smem_desc(s_w3);

  pdl_wait();
  // warp 0's elected lane issues the MMAs, warp 1's every TMA load / store (see unit_fused.cu)
  if (warp == 1 && blockIdx.x < p.ntiles) { if (uf_elect()) { load_a1(blockIdx.x); load_pass(blockIdx.x); } }
  uint32_t it = 0, mph = 0;
  for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    const bool has_next = (unsigned long long)tile + gridDim.x < p.ntiles;
    const int iy0 = 2 * ty * US_TH - 1, ix0 = 2 * tx * US_TW - 1;
    // E1 of one block or of two blocks sharing the per-column constants: 128 rows of the first GEMM each -> int8 `mid` pixels
    auto e1_setup = [&](int blk, uint32_t tcol, uint32_t& taddr, uint32_t& mpix, uint32_t& mswz, bool& valid, bool& inside, int8_t*& dump) {
      const int row = blk * 128 + q * 32 + lane;
      const int r = (row * 1986) >> 16, c = row - r * US_IW;   // row / 33 for row < 640
      valid = row < US_PIX;
      inside = (unsigned)(iy0 + r) < (unsigned)p.Hin && (unsigned)(ix0 + c) < (unsigned)p.Win;
      taddr = tmem_base + tcol + ((uint32_t)(q * 32) << 16);
      mpix = s_mid + (uint32_t)(r * US_MW + c) * (uint32_t)HP; mswz = uf_mid_swz<HP>((uint32_t)c) << 4;
      dump = (p.dump_c1 && valid && inside && r >= 1 && c >= 1) ? p.dump_c1 + (((size_t)b * p.Hin + (iy0 + r)) * p.Win + (ix0 + c)) * HP : nullptr;
    };
    auto e1_pair = [&](int blk0, uint32_t tcol0) {
      uint32_t taddr[2], mpix[2], mswz[2]; bool valid[2], inside[2]; int8_t* dump[2];
      e1_setup(blk0, tcol0, taddr[0], mpix[0], mswz[0], valid[0], inside[0], dump[0]);
      e1_setup(blk0 + 1, tcol0 + 64u, taddr[1], mpix[1], mswz[1], valid[1], inside[1], dump[1]);
      uf_e1<HP, true, 2>(taddr, mpix, mswz, valid, inside, hf, s_kc1, -128, pad, dump);
    };
    auto e1_block = [&](int blk, uint32_t tcol) {
      uint32_t taddr[1], mpix[1], mswz[1]; bool valid[1], inside[1]; int8_t* dump[1];
      e1_setup(blk, tcol, taddr[0], mpix[0], mswz[0], valid[0], inside[0], dump[0]);
      uf_e1<HP, true, 1>(taddr, mpix, mswz, valid, inside, hf, s_kc1, -128, pad, dump);
    };
    // ---- G1, blocks 0..3 -> TMEM columns 0 / 64 / 128 / 192 ------------------------------------------------------------------
    if (warp == 0) { if (uf_elect()) {
      uf_wait(bar_a, it & 1u);
      tc_fence_after();
#pragma unroll
      for (int blk = 0; blk < 4; ++blk)
        umma_i8(tmem_base + (uint32_t)(blk * 64), d_a1 + (uint64_t)(blk * 256), d_w1, idesc, 0u);   // 128 rows x 32 B = 4096 B = 256 units
      umma_commit(bar_m);
      uf_wait(bar_m, mph);
    } }
    mph ^= 1u;
    __syncthreads();
    tc_fence_after();
    e1_pair(0, 0u);
    tc_fence_before();
    __syncthreads();                           // block 0's columns are drained: block 4 goes there, behind the epilogue of blocks 2 and 3
    if (warp == 0) { if (uf_elect()) {
      tc_fence_after();
      umma_i8(tmem_base, d_a1 + (uint64_t)(4 * 256), d_w1, idesc, 0u);
      umma_commit(bar_m);
    } }
    e1_pair(2, 128u);
    if (warp == 0) { if (uf_elect()) uf_wait(bar_m, mph); }
    mph ^= 1u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && has_next) { if (uf_elect()) load_a1(tile + gridDim.x); }   // the input tile has been consumed by the tensor core
    if (q < 2) e1_block(4, 0u);                // rows 512..560 live in lane quarters 0 and 1
    if (warp == 1) { if (uf_elect()) tma_store_wait_read0(); }   // the previous tile's store has read the staging (= A2) buffer: the stencil may write A2
    tc_fence_before();
    __syncthreads();                           // `mid` complete, all accumulators drained
    // ---- S: depthwise 3x3 stride 2 over `mid` -> A tile of the second GEMM (row = oy*16 + ox) ---------------------------------
    {
      auto read_row = [&](int mr, uint32_t (&T)[4]) {
        uint32_t w[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) w[j] = lds_u32(mo[j] + (uint32_t)(mr * US_MW * HP));
        transpose4x4(w[0], w[1], w[2], pad, T[0], T[1], T[2], T[3]);
      };
      uint32_t Tm[4], Tc[4], Tp[4];
      read_row(0, Tm);
#pragma unroll
      for (int r = 0; r < US_TH; ++r) {
        read_row(2 * r + 1, Tc);
        read_row(2 * r + 2, Tp);
        int a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = dp4a_ss(Tp[c], Wt[c][2], dp4a_ss(Tc[c], Wt[c][1], dp4a_ss(Tm[c], Wt[c][0], 0)));
        const uint32_t o = uf_rq_word<true>(a, km, kb);
        sts_u32(ao + (uint32_t)(r * US_TW * 128), o);
        if (p.dump_d2) {
          uint32_t* d = (uint32_t*)(p.dump_d2 + (((size_t)b * p.Ho + (ty * US_TH + r)) * p.Wo + (tx * US_TW + ox)) * HP) + cw;
          *d = o;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) Tm[c] = Tp[c];
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- G2: [128 x 64] x [64 x 64] -> TMEM columns 64.. --------------------------------------------------------------------------
    if (warp == 0) { if (uf_elect()) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HP / 32; ++k)
        umma_i8(tmem_base + 64u, d_a2 + (uint64_t)(2 * k), d_w3 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
      umma_commit(bar_m);
      uf_wait(bar_p, it & 1u);
      uf_wait(bar_m, mph);
    } }
    mph ^= 1u;
    __syncthreads();
    tc_fence_after();
    // ---- E2 ------------------------------------------------------------------------------------------------------------------------
    {
      const int m = q * 32 + lane;
      const uint32_t taddr = tmem_base + 64u + ((uint32_t)(q * 32) << 16);
      const uint32_t prow = s_pass + (uint32_t)m * 128u, srow = s_a2 + (uint32_t)m * 128u, x7 = (uint32_t)(m & 7);
      if (hf == 0) uf_e2_fast<HP, PG, 0>(taddr, prow, srow, x7 << 4, s_kc3);
      else uf_e2_fast<HP, PG, 1>(taddr, prow, srow, x7 << 4, s_kc3);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { if (uf_elect()) {
      tma_store_4d(&tmO, 0, tx * US_TW, ty * US_TH, b, s_a2);
      tma_store_commit();
      if (has_next) load_pass(tile + gridDim.x);
    } }
  }
  if (warp == 1) { if (uf_elect()) tma_store_wait_all(); }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool unit_s2_fused_ok(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, int x_pitch, int mid_pitch, int pass_pitch,
                      int out_pitch, int H, int W) {
  if (mid_pitch != US_HP || x_pitch != 32 || out_pitch != 2 * US_HP || pass_pitch < 64 || pass_pitch > 128) return false;
  if ((H | W) & 1 || (H / 2) % US_TH || (W / 2) % US_TW) return false;
  if (!pw1.use_int || !pw3.use_int || !dw.use_int || !pw1.sh0 || !pw3.sh0 || !dw.sh0) return false;
  if (!dw.ki || !dw.wpk2 || !pw1.kc || !pw3.kc) return false;
  if (pw1.rq.lo > -128 || pw3.rq.lo > -128 || dw.rq.lo > -128) return false;
  if (pw1.n_f32 || pw3.n_f32 || pw1.n_tiles != 1 || pw3.n_tiles != 1 || pw1.Kp != 128 || pw3.Kp != 128) return false;
  if (pw1.k_off != 0 || pw1.K > 32 || pw1.BN != US_HP || pw1.has_pass) return false;
  if (pw3.k_off != 0 || pw3.K != US_HP || pw3.BN != 64 || !pw3.has_pass || pw3.il_pg != 29 || pw3.il_hp != US_HP) return false;
  if (dw.cw_total * 4 != US_HP) return false;
  return !(g_cdn_debug_flags & (1u << 23));                             // bit 23: never fuse the stride-2 unit (A/B)
}

// x: [B][H][W][32] (the unit's input), pass: [B][H/2][W/2][pass_pitch] (branch 1's output), out: [B][H/2][W/2][128]
int unit_s2_fused_launch(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, const int8_t* x, const int8_t* pass, int pass_pitch,
                         int8_t* out, int batch, int H, int W, int zx_mid, int8_t* dump_c1, int8_t* dump_d2, cudaStream_t st) {
  CDN_CHECK(unit_s2_fused_ok(pw1, dw, pw3, 32, US_HP, pass_pitch, 2 * US_HP, H, W), CDN_ERR_INVALID, "unit_s2_fused: layer triple not eligible");
  UsParams p; memset(&p, 0, sizeof(p));
  p.Hin = H; p.Win = W; p.Ho = H / 2; p.Wo = W / 2; p.tiles_x = p.Wo / US_TW; p.tiles_y = p.Ho / US_TH;
  const long long ntiles = (long long)batch * p.tiles_x * p.tiles_y;
  if (ntiles == 0) return 0;
  CDN_CHECK(ntiles < (1ll << 31) - 4 * 160, CDN_ERR_INVALID, "unit_s2_fused: tensor too large for 32-bit indexing");
  p.ntiles = (unsigned)ntiles;
  p.txs = p.tys = -1;
  if (!(p.tiles_x & (p.tiles_x - 1)) && !(p.tiles_y & (p.tiles_y - 1))) {
    p.txs = 0; while ((1 << p.txs) < p.tiles_x) ++p.txs;
    p.tys = 0; while ((1 << p.tys) < p.tiles_y) ++p.tys;
  }
  p.w1 = pw1.w; p.w3 = pw3.w; p.kc1 = (const int4*)pw1.kc; p.kc3 = (const int4*)pw3.kc;
  p.wpk = dw.wpk2; p.ki = (const int4*)dw.ki;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx_mid) * 0x01010101u;
  p.dump_c1 = dump_c1; p.dump_d2 = dump_d2;
  uint32_t o = US_A1_BYTES;
  p.off_mid = o; o += US_MID_BYTES;
  p.off_a2 = o; o += 16384u;
  p.off_pass = o; o += 16384u;
  p.off_w1 = o; o += 2048u;
  p.off_w3 = o; o += 8192u;
  p.off_kc1 = o; o += 1024u;
  p.off_kc3 = o; o += 1024u;
  p.off_bar = o; o += 64u;
  const size_t smem = (size_t)o + 1024;
  CDN_CHECK(smem <= US_SMEM_LIMIT, CDN_ERR_INVALID, "unit_s2_fused: %zu bytes of shared memory", smem);
  CUtensorMap tmA, tmP, tmO;
  if (int r = make_tmap_nhwc_swz(&tmA, x, 32, (uint64_t)W, (uint64_t)H, (uint64_t)batch, 32, US_IW, US_IH)) return r;
  if (int r = make_tmap_nhwc_swz(&tmP, pass, (uint64_t)pass_pitch, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)batch, 128, US_TW, US_TH)) return r;
  if (int r = make_tmap_nhwc_swz(&tmO, out, 2 * US_HP, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)batch, 128, US_TW, US_TH)) return r;
  auto kern = unit_s2_fused_kernel<29>;
  static bool attr_set[64] = {};
  if (cdn_first_on_device(attr_set)) {
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, US_SMEM_LIMIT));
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const int per_sm = std::max(1, std::min<int>(2, (int)(US_SMEM_LIMIT / (smem + 1024))));
  const unsigned blocks = (unsigned)std::min<long long>(ntiles, (long long)cdn_num_sms() * per_sm);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(US_THREADS); cfg.stream = st; cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CDN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmP, tmO, p));
  CDN_LAUNCH_CHECK("unit_s2_fused_kernel");
  return 0;
}
