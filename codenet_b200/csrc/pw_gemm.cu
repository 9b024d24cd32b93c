// 1x1 convolutions as int8 tcgen05 GEMMs (sm_100a): D[pixel][n] = sum_k A[pixel][k] * W[n][k], s8 x s8 -> s32 in TMEM.
//
// These GEMMs are HBM-bound (K, N = 24..1024, M = millions of pixels): the kernel is built around the byte stream,
// not around the tensor pipe.
//   * persistent, warp-specialised CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
//     warps 2..9 = epilogue (two warps per TMEM lane quarter, splitting the column chunks between them)
//   * weights stay RESIDENT in shared memory for the whole kernel whenever [N x K] fits (all but two layers); the
//     activation tiles (128 pixels x 128 B of K) stream through a ring of up to 8 TMA stages, so several tiles are
//     in flight per SM; accumulators are double-buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps
//     the loads and MMAs of tiles i+1..
//   * epilogue (the issue-bound part): tcgen05.ld -> branch-free fp32 requantisation with a rounding-boundary guard
//     (exact fp64 re-evaluation of a 16-column group only when some lane is within eps of a boundary, ~1e-4 of the
//     elements) -> bytes placed by the layer's chunk table (plain / interleaved with the pass-through half = split +
//     cat + channel_shuffle) into a swizzled staging segment -> one TMA store per 128-byte output segment, staging
//     double/triple buffered.  Per-column constants live in shared memory as one 16-byte record.
//     Head convs instead write fp32 NCHW planes (fp64 epilogue, 24 columns).
//   * a SIMT kernel with identical semantics exists for bring-up and as an on-device cross-check
//     (cdn_set_debug_flags bit 0); it is not a fallback: nothing selects it automatically.
#include "layers.cuh"
#include "tc_ptx.cuh"
#include <algorithm>

#define PW_BM 128
#define PW_BK 128
#define PW_MAX_BN 256
#ifndef PW_EPI_WARPS
#define PW_EPI_WARPS 16
#endif
#ifndef PW_EPI_GROUPS
#define PW_EPI_GROUPS 2                      // independent epilogue groups working on alternate tiles (antiphase)
#endif
#define PW_GROUP_THREADS (32 * PW_EPI_WARPS / PW_EPI_GROUPS)
#define PW_EPI_PARTS (PW_EPI_WARPS / PW_EPI_GROUPS / 4)
#define PW_EPI_THREADS (32 * PW_EPI_WARPS)
#define PW_STORE_WARP0 (2 + PW_EPI_WARPS)     // one TMA-store warp per epilogue group follows the epilogue warps
#define PW_THREADS (64 + PW_EPI_THREADS + 32 * PW_EPI_GROUPS)
#define PW_MAX_STAGES 8
#define PW_SMEM_LIMIT (227 * 1024)

struct PwSeg { int seg, cb, ce, pad; };       // output segment (128-byte column of the out tensor) and its chunk range

struct PwParams {
  int num_k_blocks, k_off, BN, n_tiles, stages, has_pass, pass_segs, pass_bufs, nbuf, resident, n_chunks, n_segs;
  int groups;                                // epilogue groups in use: 2, or 1 when shared memory is too tight for two
  long long m_tiles, pixels;
  const cdn_pw_chunk* chunks;                // sorted by (N tile, output segment)
  const PwSeg* segs; const int* tile_seg;    // tile_seg[n_tiles + 1]: first PwSeg of every N tile
  const float4* kc;                          // per GEMM column {Mh, Bh, bits(acc_bias + MAGIC_I), 0}
  const double* M; const double* B; const int32_t* acc_bias;
  float lo_f, thr;
  int use_int, lo_i;                         // integer requantisation: kc holds one RqInt per column (acc_bias folded in)
  unsigned dbg;                              // experiments only (cdn_set_debug_flags bits 2..4)
  unsigned long long* dbg_cyc;               // [16] cycle accumulators of block 0 (bit 4)
  // fp32 head output
  int n_f32, ppi; float* out_f32; const double* Mf; const double* bf;
  // SIMT cross-check path
  const int8_t* in; int in_pitch; const int8_t* pass; int pass_pitch; int8_t* out; int out_pitch;
  const int8_t* w; int Kp, N;
};

#define DBG_T0() long long t0__ = (p.dbg & 16u) ? clock64() : 0
#define DBG_ACC(slot) do { if ((p.dbg & 16u) && blockIdx.x == 0 && lane == 0) { long long t1__ = clock64(); atomicAdd(p.dbg_cyc + (slot), (unsigned long long)(t1__ - t0__)); t0__ = t1__; } } while (0)

// ---------------------------------------------------------------------------------------------------------
// epilogue math
// ---------------------------------------------------------------------------------------------------------
// NCOL accumulators -> NCOL "magic" floats whose low byte is the int8 result.  `kc` points at the constants of the
// first column, stored per PAIR of columns (c even): kc[c] = {Mh[c], Mh[c+1], Bh[c], Bh[c+1]},
// kc[c+1] = {bits(acc_bias[c] + MAGIC_I), bits(acc_bias[c+1] + MAGIC_I), -, -}: one LDS.128 + one LDS.64 feed a
// packed FFMA2 directly.  Returns true when some column of this lane came within eps of a rounding boundary.
template <int NCOL>
__device__ __forceinline__ bool requant_cols_fast(const uint32_t (&acc)[16], const float4* __restrict__ kc, float lo_f,
                                                  float thr, uint32_t (&rb)[16]) {
  RqGuard g; rq_guard_init(g);
#pragma unroll
  for (int i = 0; i < NCOL; i += 2) {
    const float4 mb = kc[i];
    const float2 ab = *reinterpret_cast<const float2*>(kc + i + 1);
    rq_fast2((int)acc[i] + __float_as_int(ab.x), (int)acc[i + 1] + __float_as_int(ab.y), make_float2(mb.x, mb.y),
             make_float2(mb.z, mb.w), lo_f, g, rb[i], rb[i + 1]);
  }
  return rq_group_bad(g, thr);
}
__device__ __forceinline__ int kc_acc_bias(const float4* __restrict__ kc, int n) {     // kc = table base, n = column
  const float4 r = kc[(n & ~1) + 1];
  return __float_as_int((n & 1) ? r.y : r.x) - CDN_MAGIC_I;
}

// exact fp64 re-evaluation of the same NCOL columns (rare); kc = table base here
template <int NCOL>
__device__ __forceinline__ void requant_cols_exact(const uint32_t (&acc)[16], const float4* __restrict__ kc, int col,
                                                   const double* __restrict__ Md, const double* __restrict__ Bd, float lo_f,
                                                   uint32_t (&rb)[16]) {
#pragma unroll
  for (int i = 0; i < NCOL; ++i)
    rb[i] = rq_exact((int)acc[i] + kc_acc_bias(kc, col + i), __ldg(Md + col + i), __ldg(Bd + col + i), lo_f);
}

// keep the first nb bytes of a 16-byte vector, zero the rest
__device__ __forceinline__ uint32_t mask_word(uint32_t w, int rem) {
  return rem >= 4 ? w : (rem <= 0 ? 0u : (w & (0xffffffffu >> (8 * (4 - rem)))));
}
__device__ __forceinline__ uint4 mask_tail(uint4 o, int nb) {
  return make_uint4(mask_word(o.x, nb), mask_word(o.y, nb - 4), mask_word(o.z, nb - 8), mask_word(o.w, nb - 12));
}

// 16 output bytes of one chunk.  acc: raw accumulators of the chunk's columns; pass_lo/hi: 8 pass-through bytes.
__device__ __forceinline__ uint4 chunk_bytes(const cdn_pw_chunk& ck, const uint32_t (&acc)[16], const float4* __restrict__ kc,
                                             const double* __restrict__ Md, const double* __restrict__ Bd, float lo_f, float thr,
                                             uint32_t pass_lo, uint32_t pass_hi) {
  uint32_t q[16];
  uint4 o;
  if (ck.pass_off < 0) {
    if (ck.count == 0) return make_uint4(0u, 0u, 0u, 0u);
    const bool bad = requant_cols_fast<16>(acc, kc + ck.col, lo_f, thr, q);
    if (__any_sync(__activemask(), bad)) { if (bad) requant_cols_exact<16>(acc, kc, ck.col, Md, Bd, lo_f, q); }
    o.x = pack4_lowbytes(q[0], q[1], q[2], q[3]);   o.y = pack4_lowbytes(q[4], q[5], q[6], q[7]);
    o.z = pack4_lowbytes(q[8], q[9], q[10], q[11]); o.w = pack4_lowbytes(q[12], q[13], q[14], q[15]);
    if (ck.count < 16) o = mask_tail(o, ck.count);   // pad bytes of the pixel stay zero
  } else {
    const bool bad = requant_cols_fast<8>(acc, kc + ck.col, lo_f, thr, q);
    if (__any_sync(__activemask(), bad)) { if (bad) requant_cols_exact<8>(acc, kc, ck.col, Md, Bd, lo_f, q); }
    const uint32_t n_lo = pack4_lowbytes(q[0], q[1], q[2], q[3]), n_hi = pack4_lowbytes(q[4], q[5], q[6], q[7]);
    // out[2i] = pass[i], out[2i+1] = new[i]
    o.x = __byte_perm(pass_lo, n_lo, 0x5140); o.y = __byte_perm(pass_lo, n_lo, 0x7362);
    o.z = __byte_perm(pass_hi, n_hi, 0x5140); o.w = __byte_perm(pass_hi, n_hi, 0x7362);
    if (ck.count < 8) o = mask_tail(o, 2 * ck.count);
  }
  return o;
}

// The same 16 output bytes with the integer requantisation (RqInt per column, raw accumulators): no guard, no slow path.
// NS: every column of the layer requantises with shift 0 (PwDevice::sh0): the SHF after the IMAD.HI is dropped
template <bool NS> __device__ __forceinline__ int rq_int_sel(int v, const int4& r) {
  if (NS) { const long long x = (long long)v * r.x + (long long)(((unsigned long long)(uint32_t)r.w << 32) | (uint32_t)r.z); return (int)(x >> 32); }
  return rq_int(v, r);
}
template <bool LO, bool NS = false>
__device__ __forceinline__ uint4 chunk_bytes_int(const cdn_pw_chunk& ck, const uint32_t (&acc)[16], const int4* __restrict__ kc,
                                                 int lo, uint32_t pass_lo, uint32_t pass_hi) {
  const int4* __restrict__ k = kc + ck.col;
  int q[16];
  uint4 o;
  if (ck.pass_off < 0) {
    if (ck.count == 0) return make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = rq_int_sel<NS>((int)acc[i], k[i]); if (LO) q[i] = max(q[i], lo); }
    o.x = pack_sat4(q[0], q[1], q[2], q[3]);   o.y = pack_sat4(q[4], q[5], q[6], q[7]);
    asm volatile("" ::: "memory");             // keeps the second half's 8 LDS.128 (32 registers) from being hoisted too
#pragma unroll
    for (int i = 8; i < 16; ++i) { q[i] = rq_int_sel<NS>((int)acc[i], k[i]); if (LO) q[i] = max(q[i], lo); }
    o.z = pack_sat4(q[8], q[9], q[10], q[11]); o.w = pack_sat4(q[12], q[13], q[14], q[15]);
    if (ck.count < 16) o = mask_tail(o, ck.count);   // pad bytes of the pixel stay zero
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = rq_int_sel<NS>((int)acc[i], k[i]); if (LO) q[i] = max(q[i], lo); }
    const uint32_t n_lo = pack_sat4(q[0], q[1], q[2], q[3]), n_hi = pack_sat4(q[4], q[5], q[6], q[7]);
    o.x = __byte_perm(pass_lo, n_lo, 0x5140); o.y = __byte_perm(pass_lo, n_lo, 0x7362);
    o.z = __byte_perm(pass_hi, n_hi, 0x5140); o.w = __byte_perm(pass_hi, n_hi, 0x7362);
    if (ck.count < 8) o = mask_tail(o, 2 * ck.count);
  }
  return o;
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------------------
// Shared memory carve (host mirrors this in pw_device_build):
//   [B resident: num_k_blocks x Ntot x 128 B]  (resident only)
//   [ring: stages x (A 16 KB [+ B block BN x 128 B rounded to 1 KB when streamed])]
//   [pass: 2 x pass_segs x 16 KB] [staging: nbuf x 16 KB] [kc: Ntot x 16 B] [chunks] [segs] [tile_seg] [barriers]
// RQ selects the epilogue at compile time, so the hot loop holds one variant only: 0 = guarded fp32 requantisation,
// 1 = integer requantisation, 2 = integer with an explicit lower clamp (lo > -128), 3 = fp32 NCHW head planes,
// 4 = integer requantisation with shift 0 in every column (no SHF after the IMAD.HI).
template <int RQ>
__global__ void __launch_bounds__(PW_THREADS, 1)
pw_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmO,
                  const PwParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  const int Ntot = p.BN * p.n_tiles;
  const uint32_t a_bytes = PW_BM * PW_BK, b_blk = ((uint32_t)p.BN * PW_BK + 1023) & ~1023u;
  const uint32_t stage_bytes = a_bytes + (p.resident ? 0u : b_blk);
  uint8_t* s_B = smem;
  uint8_t* s_ring = s_B + (p.resident ? (size_t)p.num_k_blocks * Ntot * PW_BK : 0);
  uint8_t* s_pass = s_ring + (size_t)p.stages * stage_bytes;
  uint8_t* s_out = s_pass + (size_t)p.pass_bufs * p.pass_segs * 16384;
  float4* s_kc = (float4*)(s_out + (size_t)p.groups * p.nbuf * 16384);
  cdn_pw_chunk* s_chunks = (cdn_pw_chunk*)(s_kc + Ntot);
  PwSeg* s_segs = (PwSeg*)(s_chunks + ((p.n_chunks + 1) & ~1));
  int* s_tile_seg = (int*)(s_segs + p.n_segs);
  uint64_t* bars = (uint64_t*)(s_tile_seg + ((p.n_tiles + 1 + 3) & ~3));
  // barrier slots: full[8], empty[8], tmem_full[4], tmem_empty[4], pass_full[2], pass_empty[2], b_full, tmem_ptr
  const uint32_t bar_base = smem_u32(bars);
  auto FULL = [&](int s) { return bar_base + 8u * s; };
  auto EMPTY = [&](int s) { return bar_base + 8u * (8 + s); };
  auto TFULL = [&](int s) { return bar_base + 8u * (16 + s); };
  auto TEMPTY = [&](int s) { return bar_base + 8u * (20 + s); };
  auto PFULL = [&](int s) { return bar_base + 8u * (42 + s); };
  auto PEMPTY = [&](int s) { return bar_base + 8u * (46 + s); };
  const uint32_t BFULL = bar_base + 8u * 28;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(bars + 29);
  auto SFULL = [&](int g_, int b_) { return bar_base + 8u * (30 + g_ * 3 + b_); };     // staging buffer written by the group
  auto SEMPTY = [&](int g_, int b_) { return bar_base + 8u * (36 + g_ * 3 + b_); };    // ... drained by its TMA store
  // accumulator stages: each epilogue group owns one (BN > 128) or two (BN <= 128) of the 512 TMEM columns' stages, so
  // the MMAs of a group's next tile run while it still drains the previous one
  const int nacc = p.BN <= 128 ? 4 : 2;
  const uint32_t acc_stride = p.BN <= 128 ? 128u : 256u;
  // nacc and pass_bufs are powers of two and n_tiles is 1 for most layers: no integer division on the per-tile paths
  const uint32_t acc_mask = (uint32_t)nacc - 1u, acc_shift = nacc == 4 ? 2u : 1u;
  const uint32_t pb_mask = (uint32_t)p.pass_bufs - 1u, pb_shift = p.pass_bufs == 4 ? 2u : (p.pass_bufs == 2 ? 1u : 0u);
  auto split_tile = [&](unsigned tile, unsigned& mt, int& nt) {
    if (p.n_tiles == 1) { mt = tile; nt = 0; } else { mt = tile / (unsigned)p.n_tiles; nt = (int)(tile - mt * (unsigned)p.n_tiles); }
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(TFULL(s), 1); mbar_init(TEMPTY(s), PW_GROUP_THREADS); }
    for (int s = 0; s < 4; ++s) { mbar_init(PFULL(s), 1); mbar_init(PEMPTY(s), PW_GROUP_THREADS); }
    for (int g_ = 0; g_ < PW_EPI_GROUPS; ++g_)
      for (int b_ = 0; b_ < 3; ++b_) { mbar_init(SFULL(g_, b_), PW_GROUP_THREADS); mbar_init(SEMPTY(g_, b_), 1); }
    mbar_init(BFULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
  }
  if (warp == 1) {                           // TMEM: the whole 512 columns (1 CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // per-column constants and the chunk / segment tables -> shared memory
  for (int i = threadIdx.x; i < Ntot; i += PW_THREADS) s_kc[i] = p.kc[i];
  for (int i = threadIdx.x; i < p.n_chunks; i += PW_THREADS) s_chunks[i] = p.chunks[i];
  for (int i = threadIdx.x; i < p.n_segs; i += PW_THREADS) s_segs[i] = p.segs[i];
  for (int i = threadIdx.x; i <= p.n_tiles; i += PW_THREADS) s_tile_seg[i] = p.tile_seg[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const unsigned total_tiles = (unsigned)(p.m_tiles * p.n_tiles);   // host guarantees < 2^31
  const long long t_start = (p.dbg & 16u) ? clock64() : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.resident) {
        mbar_expect_tx(BFULL, (uint32_t)p.num_k_blocks * (uint32_t)Ntot * PW_BK);
        for (int kb = 0; kb < p.num_k_blocks; ++kb)
          for (int nt = 0; nt < p.n_tiles; ++nt)
            tma_load_2d(smem_u32(s_B + ((size_t)kb * Ntot + (size_t)nt * p.BN) * PW_BK), &tmB, kb * PW_BK, nt * p.BN, BFULL);
      }
      pdl_wait();                              // weights above are constants; activations need the previous grid
      // Two independent cursors -- activation (and streamed weight) blocks into the ring, pass-through tiles into
      // their own buffers -- advanced by polling, so a full pass buffer never stops the activation prefetch.
      int stage = 0; uint32_t phase = 0;
      unsigned a_tile = blockIdx.x; int a_kb = 0;
      unsigned p_tile = p.has_pass ? blockIdx.x : total_tiles; uint32_t p_it = 0;
      uint32_t idle = 0;
      while (a_tile < total_tiles || p_tile < total_tiles) {
        bool progressed = false;
        if (p_tile < total_tiles) {
          const int pb = (int)(p_it & pb_mask); const uint32_t pph = (p_it >> pb_shift) & 1;
          if (mbar_test(PEMPTY(pb), pph ^ 1)) {
            unsigned mt; int nt_; split_tile(p_tile, mt, nt_);
            mbar_expect_tx(PFULL(pb), (uint32_t)p.pass_segs * 16384u);
            for (int s = 0; s < p.pass_segs; ++s)
              tma_load_2d(smem_u32(s_pass + ((size_t)pb * p.pass_segs + s) * 16384), &tmP, s * 128, (int)(mt * PW_BM), PFULL(pb));
            p_tile += gridDim.x; ++p_it; progressed = true;
          }
        }
        // without pass-through tiles there is nothing to interleave: block on the ring slot (hardware-suspended wait)
        if (a_tile < total_tiles && !p.has_pass) { DBG_T0(); mbar_wait(EMPTY(stage), phase ^ 1); DBG_ACC(0); }
        if (a_tile < total_tiles && (!p.has_pass || mbar_test(EMPTY(stage), phase ^ 1))) {
          unsigned mt; int nt; split_tile(a_tile, mt, nt);
          uint8_t* sa = s_ring + (size_t)stage * stage_bytes;
          if (p.resident) {
            mbar_expect_tx(FULL(stage), a_bytes);
            tma_load_2d(smem_u32(sa), &tmA, p.k_off + a_kb * PW_BK, (int)(mt * PW_BM), FULL(stage));
          } else {
            mbar_expect_tx(FULL(stage), a_bytes + (uint32_t)p.BN * PW_BK);
            tma_load_2d(smem_u32(sa), &tmA, p.k_off + a_kb * PW_BK, (int)(mt * PW_BM), FULL(stage));
            tma_load_2d(smem_u32(sa + a_bytes), &tmB, a_kb * PW_BK, nt * p.BN, FULL(stage));
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (++a_kb == p.num_k_blocks) { a_kb = 0; a_tile += gridDim.x; }
          progressed = true;
        }
        if (!progressed) {
          if (++idle > PW_SPIN_LIMIT) { printf("cdn pw_gemm: producer timeout (block %d)\n", blockIdx.x); __trap(); }
        } else idle = 0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: S32 accumulate, A/B signed int8, K-major both, N = BN, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(PW_BM >> 4) << 24);
      if (p.resident) { mbar_wait(BFULL, 0); tc_fence_after(); }
      int stage = 0; uint32_t phase = 0; uint32_t it = 0;
      for (unsigned tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        unsigned mt_; int nt; split_tile(tile, mt_, nt);
        const int as = (int)(it & acc_mask);                       // accumulator stages rotate over the CTA's tiles
        const uint32_t aphase = (it >> acc_shift) & 1;
        { DBG_T0(); mbar_wait(TEMPTY(as), aphase ^ 1); DBG_ACC(2); }
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)as * acc_stride;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          { DBG_T0(); mbar_wait(FULL(stage), phase); DBG_ACC(3); }
          tc_fence_after();
          const uint32_t sa = smem_u32(s_ring + (size_t)stage * stage_bytes);
          const uint32_t sb = p.resident ? smem_u32(s_B + ((size_t)kb * Ntot + (size_t)nt * p.BN) * PW_BK) : sa + a_bytes;
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < PW_BK / 32; ++k)
            umma_i8(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(EMPTY(stage));           // frees the smem stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(TFULL(as));                // accumulator complete
      }
    }
  } else if (warp >= PW_STORE_WARP0) {
    // ===================== TMA-store warps (one per epilogue group) =====================
    // The hand-off staging buffer -> global memory is off the epilogue's critical path: the group signals SFULL when a
    // 128-byte output segment is complete, this warp issues the store, waits until the TMA engine has read the
    // buffer and hands it back through SEMPTY.
    const int grp = warp - PW_STORE_WARP0;
    pdl_wait();
    if (lane == 0 && RQ != 3 && grp < p.groups) {
      uint32_t g = 0; int sbuf = 0; uint32_t sph = 0;          // staging buffer cursor (index, phase)
      for (unsigned tile = blockIdx.x + grp * gridDim.x; tile < total_tiles; tile += p.groups * gridDim.x) {
        unsigned mt; int nt; split_tile(tile, mt, nt);
        const int sg0 = s_tile_seg[nt], sg1 = s_tile_seg[nt + 1];
        for (int sg = sg0; sg < sg1; ++sg, ++g) {
          const int b_ = sbuf; const uint32_t ph = sph;
          if (++sbuf == p.nbuf) { sbuf = 0; sph ^= 1; }
          DBG_T0();
          mbar_wait(SFULL(grp, b_), ph);
          if (grp == 0) DBG_ACC(9);
          tma_store_2d(&tmO, s_segs[sg].seg * 128, (int)(mt * PW_BM), smem_u32(s_out + ((size_t)grp * p.nbuf + b_) * 16384));
          tma_store_commit();
          tma_store_wait_read0();
          if (grp == 0) DBG_ACC(10);
          mbar_arrive(SEMPTY(grp, b_));
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    // Two independent groups of 8 warps: group g owns TMEM accumulator stage g, pass buffer g and its own staging
    // buffers, and handles every other tile of this CTA, so one group's latency phases (barrier waits, TMEM /
    // shared-memory loads, the TMA store hand-off) overlap the other group's arithmetic.
    pdl_wait();                                // the fp32 head path stores to global memory directly
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int grp = (warp - 2) / (PW_EPI_WARPS / PW_EPI_GROUPS);
    const int half = ((warp - 2) >> 2) % PW_EPI_PARTS;   // which of the PW_EPI_PARTS warps of the quarter (inside the group)
    const int row = q * 32 + lane;             // row inside the tile
    const int gt = (threadIdx.x - 64) % PW_GROUP_THREADS;
    uint8_t* const s_out_g = s_out + (size_t)grp * p.nbuf * 16384;
    uint32_t it2 = 0, g = 0;                   // tiles / output segments done by this group
    int sbuf = 0; uint32_t sph = 0;            // staging buffer cursor (index, phase)
    for (unsigned tile = blockIdx.x + grp * gridDim.x; tile < total_tiles && grp < p.groups; tile += p.groups * gridDim.x, ++it2) {
      unsigned mt; int nt; split_tile(tile, mt, nt);
      const uint32_t itg = it2 * (uint32_t)p.groups + grp;                   // CTA-wide tile counter (as producer / MMA count)
      const int as = (int)(itg & acc_mask);
      const uint32_t aphase = (itg >> acc_shift) & 1;
      DBG_T0();
      mbar_wait(TFULL(as), aphase);
      if (warp == 2) DBG_ACC(4);
      tc_fence_after();
      const int pb = (int)(itg & pb_mask);
      if (p.has_pass) mbar_wait(PFULL(pb), (itg >> pb_shift) & 1);
      if (warp == 2) DBG_ACC(5);
      const uint32_t tacc = tmem_base + (uint32_t)as * acc_stride + ((uint32_t)(q * 32) << 16);
      if (RQ == 3) {
        // fp32 NCHW planes: out[img][n][pix] = fl32(fl64(acc*Mf[n]) + bf[n])
        const unsigned pix = mt * PW_BM + row;
        const unsigned img = pix / (unsigned)p.ppi; const int pi = (int)(pix - img * (unsigned)p.ppi);
        for (int c0 = half * 16; c0 < p.BN; c0 += 16 * PW_EPI_PARTS) {
          uint32_t acc[16];
          tmem_ld16(tacc + (uint32_t)c0, acc);
          tmem_ld_wait();
          if (pix < p.pixels) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = nt * p.BN + c0 + i;
              if (n < p.n_f32) {
                const int a = (int)acc[i] + kc_acc_bias(s_kc, n);
                const double y = __dadd_rn(__dmul_rn((double)a, __ldg(p.Mf + n)), __ldg(p.bf + n));
                p.out_f32[((size_t)img * p.n_f32 + n) * p.ppi + pi] = (float)y;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(TEMPTY(as));
      } else {
        const uint8_t* pass_row = s_pass + (size_t)pb * p.pass_segs * 16384 + row * 128;
        const int sg0 = s_tile_seg[nt], sg1 = s_tile_seg[nt + 1];
        for (int sg = sg0; sg < sg1; ++sg, ++g) {
          const PwSeg S = s_segs[sg];
          const int sb = sbuf;
          mbar_wait(SEMPTY(grp, sb), sph ^ 1);                                     // the store that last used this buffer has drained
          if (++sbuf == p.nbuf) { sbuf = 0; sph ^= 1; }
          if (warp == 2) DBG_ACC(6);
          uint8_t* stg = s_out_g + (size_t)sb * 16384 + row * 128;
          for (int c = S.cb + half; c < S.ce; c += PW_EPI_PARTS) {
            const cdn_pw_chunk ck = s_chunks[c];
            uint32_t acc[16];
            uint32_t pass_lo = 0, pass_hi = 0;
            if (ck.count > 0) {
              const uint32_t ta = tacc + (uint32_t)(ck.col - nt * p.BN);
              if (ck.pass_off < 0) tmem_ld16(ta, acc); else tmem_ld8(ta, acc);
            }
            if (ck.pass_off >= 0) {
              // 8 pass-through bytes at an arbitrary byte offset: two aligned 8-byte words (each inside one 16-byte
              // swizzle unit) and a funnel shift
              const int o8 = ck.pass_off & ~7, b = ck.pass_off & 7, o9 = o8 + 8;
              const uint2 pa = *(const uint2*)(pass_row + (size_t)(o8 >> 7) * 16384 + (((((o8 & 127) >> 4) ^ (row & 7)) << 4) | (o8 & 8)));
              const uint2 pb = *(const uint2*)(pass_row + (size_t)(o9 >> 7) * 16384 + (((((o9 & 127) >> 4) ^ (row & 7)) << 4) | (o9 & 8)));
              uint32_t w0 = pa.x, w1 = pa.y, w2 = pb.x;
              if (b >= 4) { w0 = pa.y; w1 = pb.x; w2 = pb.y; }
              const int sh = 8 * (b & 3);
              pass_lo = __funnelshift_r(w0, w1, sh); pass_hi = __funnelshift_r(w1, w2, sh);
            }
            tmem_ld_wait();
            uint4 o;
            if (RQ == 4) o = chunk_bytes_int<false, true>(ck, acc, (const int4*)s_kc, p.lo_i, pass_lo, pass_hi);
            else if (RQ == 1) o = chunk_bytes_int<false>(ck, acc, (const int4*)s_kc, p.lo_i, pass_lo, pass_hi);
            else if (RQ == 2) o = chunk_bytes_int<true>(ck, acc, (const int4*)s_kc, p.lo_i, pass_lo, pass_hi);
            else if (p.dbg & 4u) o = make_uint4(acc[0], acc[1], pass_lo, pass_hi);     // experiment: no requant math
            else o = chunk_bytes(ck, acc, s_kc, p.M, p.B, p.lo_f, p.thr, pass_lo, pass_hi);
            const int j = (ck.dst_off & 127) >> 4;
            *(uint4*)(stg + ((j ^ (row & 7)) << 4)) = o;
          }
          if (warp == 2) DBG_ACC(7);
          if (sg + 1 == sg1) {                 // last TMEM / pass read of this tile is behind us
            tc_fence_before();
            mbar_arrive(TEMPTY(as));
            if (p.has_pass) mbar_arrive(PEMPTY(pb));
          }
          fence_async_smem();
          mbar_arrive(SFULL(grp, sb));
          if (warp == 2) DBG_ACC(8);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if ((p.dbg & 16u) && blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(p.dbg_cyc + 11, (unsigned long long)(clock64() - t_start)); atomicAdd(p.dbg_cyc + 12, 1ull); }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------------------
// SIMT cross-check kernel: one thread per (pixel, chunk) [int8 output] or (pixel, 16 columns) [fp32 output]
// ---------------------------------------------------------------------------------------------------------
__global__ void pw_gemm_simt_kernel(const PwParams p, int total_chunks) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = p.n_f32 > 0 ? (p.N / 16) : total_chunks;
  if (idx >= p.pixels * per_row) return;
  const long long pix = idx / per_row; const int ci = (int)(idx % per_row);
  cdn_pw_chunk ck;
  if (p.n_f32 > 0) { ck.col = (int16_t)(ci * 16); ck.count = 16; ck.pass_off = -1; ck.dst_off = 0; }
  else ck = p.chunks[ci];
  uint32_t acc[16];
  const int8_t* a = p.in + (size_t)pix * p.in_pitch + p.k_off;
  const int ncol = ck.pass_off < 0 ? 16 : 8;
  for (int i = 0; i < 16; ++i) {
    int s = 0;
    if (i < ncol && ck.count > 0) {
      const int8_t* w = p.w + (size_t)(ck.col + i) * p.Kp;
      for (int k = 0; k < p.Kp; ++k) {
        int kk = p.k_off + k;
        int av = (kk < p.in_pitch) ? (int)a[k] : 0;       // TMA zero-fills beyond the pixel
        s += av * (int)w[k];
      }
    }
    acc[i] = (uint32_t)s;
  }
  if (p.n_f32 > 0) {
    const long long img = pix / p.ppi; const int pi = (int)(pix - img * p.ppi);
    for (int i = 0; i < 16; ++i) {
      int n = ck.col + i;
      if (n < p.n_f32) {
        double y = __dadd_rn(__dmul_rn((double)((int)acc[i] + p.acc_bias[n]), p.Mf[n]), p.bf[n]);
        p.out_f32[((size_t)img * p.n_f32 + n) * p.ppi + pi] = (float)y;
      }
    }
    return;
  }
  uint32_t pass_lo = 0, pass_hi = 0;
  if (ck.pass_off >= 0)
    for (int i = 0; i < 8; ++i) {
      uint32_t byte = (ck.pass_off + i < p.pass_pitch) ? (uint8_t)p.pass[(size_t)pix * p.pass_pitch + ck.pass_off + i] : 0u;
      if (i < 4) pass_lo |= byte << (8 * i); else pass_hi |= byte << (8 * (i - 4));
    }
  uint4 o;
  if (p.use_int) {
    o = chunk_bytes_int<true>(ck, acc, (const int4*)p.kc, p.lo_i, pass_lo, pass_hi);
  } else o = chunk_bytes(ck, acc, p.kc, p.M, p.B, p.lo_f, p.thr, pass_lo, pass_hi);
  if (ck.dst_off + 16 <= p.out_pitch) *(uint4*)(p.out + (size_t)pix * p.out_pitch + ck.dst_off) = o;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// 2D uint8 tensor [rows][cols] with row pitch `pitch` bytes; box = [box_rows][128 bytes], 128-byte swizzle.
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  CDN_CHECK(enc != nullptr, CDN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CDN_CHECK(((uintptr_t)base & 15) == 0 && pitch % 16 == 0, CDN_ERR_INVALID, "TMA: base/pitch must be 16-byte aligned");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CDN_CHECK(r == CUDA_SUCCESS, CDN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu pitch=%llu)",
            (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch);
  return 0;
}

// 4-D uint8 tensor [batch][H][W][pitch bytes] (NHWC); box = [1][box_h][box_w][pitch], no swizzle, zero fill outside.
int make_tmap_nhwc(CUtensorMap* m, const void* base, uint64_t pitch, uint64_t W, uint64_t H, uint64_t batch, uint32_t box_w, uint32_t box_h) {
  PFN_encodeTiled enc = get_encode();
  CDN_CHECK(enc != nullptr, CDN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CDN_CHECK(((uintptr_t)base & 15) == 0 && pitch % 16 == 0 && pitch <= 256 && box_w <= 256 && box_h <= 256, CDN_ERR_INVALID,
            "TMA: base/pitch must be 16-byte aligned, box dimensions <= 256");
  cuuint64_t dims[4] = {pitch, W, H, batch};
  cuuint64_t strides[3] = {pitch, W * pitch, H * W * pitch};
  cuuint32_t box[4] = {(cuuint32_t)pitch, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CDN_CHECK(r == CUDA_SUCCESS, CDN_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC) failed with CUresult %d (pitch=%llu W=%llu H=%llu batch=%llu)",
            (int)r, (unsigned long long)pitch, (unsigned long long)W, (unsigned long long)H, (unsigned long long)batch);
  return 0;
}

// NHWC int8 tensor of any pixel pitch, box = {box_c channel bytes (<= 256), box_w, box_h, 1}: a channel SLICE of a row band
int make_tmap_nhwc_box(CUtensorMap* m, const void* base, uint64_t pitch, uint64_t W, uint64_t H, uint64_t batch,
                       uint32_t box_c, uint32_t box_w, uint32_t box_h) {
  PFN_encodeTiled enc = get_encode();
  CDN_CHECK(enc != nullptr, CDN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CDN_CHECK(((uintptr_t)base & 15) == 0 && pitch % 16 == 0 && box_c % 16 == 0 && box_c <= 256 && box_w <= 256 && box_h <= 256 &&
            box_c >= 16 && box_w >= 1 && box_h >= 1, CDN_ERR_INVALID, "TMA: base/pitch/slice must be 16-byte aligned, box dimensions <= 256");
  cuuint64_t dims[4] = {pitch, W, H, batch};
  cuuint64_t strides[3] = {pitch, W * pitch, H * W * pitch};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CDN_CHECK(r == CUDA_SUCCESS, CDN_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC box) failed with CUresult %d (pitch=%llu W=%llu H=%llu batch=%llu box=%u,%u,%u)",
            (int)r, (unsigned long long)pitch, (unsigned long long)W, (unsigned long long)H, (unsigned long long)batch, box_c, box_w, box_h);
  return 0;
}

void pw_device_free(PwDevice& d) {
  cudaFree(d.w); cudaFree(d.chunks); cudaFree(d.segs); cudaFree(d.tile_seg); cudaFree(d.kc); cudaFree(d.Mf); cudaFree(d.bf);
  dev_requant_free(d.rq);
  d = PwDevice();
}

int pw_device_build(PwDevice& d, const cdn_pw_desc* desc, int pass_pitch) {
  CDN_CHECK(desc && desc->wq, CDN_ERR_INVALID, "pw: null descriptor");
  CDN_CHECK(desc->K > 0 && desc->N > 0 && desc->N % 16 == 0 && desc->N <= 4096, CDN_ERR_INVALID,
            "pw: N=%d must be a positive multiple of 16 (<= 4096), K=%d > 0", desc->N, desc->K);
  d.K = desc->K; d.k_off = desc->k_off; d.N = desc->N; d.n_f32 = desc->n_f32;
  d.Kp = (desc->K + PW_BK - 1) / PW_BK * PW_BK;
  d.num_k_blocks = d.Kp / PW_BK;
  // N tiling: one tile when N <= 256, else tiles of 256 columns (so tiles own whole 128-byte output segments)
  d.n_tiles = (d.N + PW_MAX_BN - 1) / PW_MAX_BN;
  d.BN = d.n_tiles == 1 ? d.N : PW_MAX_BN;
  const int Np = d.BN * d.n_tiles;
  // weights, zero padded to [Np][Kp]; acc_bias = zx * sum_k w.  The epilogue converts the accumulator through the
  // 1.5*2^23 magic constant, which needs |acc + acc_bias| < 2^22: with a = q + zx in [-255, 255] that is a bound on
  // sum_k |w|, checked per column.
  std::vector<int8_t> w((size_t)Np * d.Kp, 0);
  std::vector<int32_t> ab(Np, 0);
  for (int n = 0; n < d.N; ++n) {
    long long sum = 0, asum = 0;
    for (int k = 0; k < d.K; ++k) { int8_t v = desc->wq[(size_t)n * d.K + k]; w[(size_t)n * d.Kp + k] = v; sum += v; asum += v < 0 ? -v : v; }
    CDN_CHECK(asum * 255 < (1ll << 22), CDN_ERR_INVALID, "pw: column %d: sum|w| = %lld makes the accumulator exceed 2^22 (K=%d)", n, asum, d.K);
    ab[n] = (int32_t)(desc->zx * sum);
  }
  if (dev_upload(&d.w, w.data(), w.size())) return CDN_ERR_CUDA;
  std::vector<float4> kc(Np);
  double min_thr = 0.5;
  auto fill_kc = [&](const DevRequant&, const cdn_requant* rq) {      // pair layout, see requant_cols_fast
    for (int n = 0; n < Np; n += 2) {
      RqFast f[2] = {{0.f, 0.f, 0.5f}, {0.f, 0.f, 0.5f}};
      float z[2];
      for (int j = 0; j < 2; ++j) {
        if (rq && n + j < rq->n) { f[j] = rq_fast_from(rq->M[n + j], rq->B[n + j]); min_thr = std::min(min_thr, (double)f[j].thr); }
        int32_t bits = ab[n + j] + CDN_MAGIC_I_HOST;
        memcpy(&z[j], &bits, 4);
      }
      kc[n] = make_float4(f[0].Mh, f[1].Mh, f[0].Bh, f[1].Bh);
      kc[n + 1] = make_float4(z[0], z[1], 0.f, 0.f);
    }
  };
  std::vector<cdn_pw_chunk> ch;
  std::vector<PwSeg> segs;
  std::vector<int> tile_seg(d.n_tiles + 1, 0);
  if (d.n_f32 > 0) {
    CDN_CHECK(desc->Mf && desc->bf && d.n_f32 <= d.N, CDN_ERR_INVALID, "pw: fp32 output needs Mf/bf and n_f32 <= N");
    std::vector<double> Mf(Np, 0.0), bf(Np, 0.0), one(Np, 1.0), zero(Np, 0.0);
    for (int n = 0; n < d.n_f32; ++n) { Mf[n] = desc->Mf[n]; bf[n] = desc->bf[n]; }
    if (dev_upload(&d.Mf, Mf.data(), Np)) return CDN_ERR_CUDA;
    if (dev_upload(&d.bf, bf.data(), Np)) return CDN_ERR_CUDA;
    cdn_requant dummy{one.data(), zero.data(), -128, Np};
    if (int r = dev_requant_upload(d.rq, &dummy, ab.data(), Np)) return r;
    fill_kc(d.rq, nullptr);
    d.has_pass = 0; d.n_chunks = 0; d.n_segs = 0; d.nbuf = 0; d.pass_segs = 0;
    ch.push_back(cdn_pw_chunk{0, 0, -1, 0});      // keep the device arrays non-empty
    segs.push_back(PwSeg{0, 0, 0, 0});
  } else {
    CDN_CHECK(desc->rq.n == d.N && desc->rq.M && desc->rq.B, CDN_ERR_INVALID, "pw: requant constants must have n == N");
    CDN_CHECK(desc->chunks && desc->n_chunks > 0, CDN_ERR_INVALID, "pw: int8 output needs a chunk table");
    if (int r = dev_requant_upload(d.rq, &desc->rq, ab.data(), Np)) return r;
    fill_kc(d.rq, &desc->rq);
    // integer requantisation (common.cuh, RqInt): one exact fixed-point pair per column over the accumulator range the
    // column can reach (|q + zx| <= A, so |v| <= A * sum|w|), re-based to the raw accumulator sum_k w*q of the MMA.
    // If any column has none the layer keeps the guarded fp32 table built above.
    if (!(g_cdn_debug_flags & 128u)) {         // bit 7: force the guarded fp32 epilogue (A/B measurements)
      const long long A = std::max(std::abs((long long)desc->zx - 128), std::abs((long long)desc->zx + 127));
      std::vector<RqInt> ki(Np, RqInt{0, 0, 0});
      bool ok = true;
      for (int n = 0; n < d.N && ok; ++n) {
        long long asum = 0;
        for (int k = 0; k < d.K; ++k) { const int v = desc->wq[(size_t)n * d.K + k]; asum += v < 0 ? -v : v; }
        ok = rq_int_solve(desc->rq.M[n], desc->rq.B[n], d.rq.lo, -A * asum, A * asum, &ki[n]) &&
             rq_int_rebase(&ki[n], ab[n], 128 * asum);
      }
      if (ok) {
        static_assert(sizeof(RqInt) == sizeof(float4), "RqInt must be one 16-byte record");
        memcpy(kc.data(), ki.data(), (size_t)Np * sizeof(RqInt));
        d.use_int = 1;
        d.sh0 = 1;
        for (int n = 0; n < Np; ++n) if (ki[n].sh != 0) d.sh0 = 0;
      }
    }
    // the canonical interleave of a ShuffleNetV2 unit (plan.py emit_pw): group g = 0, 1 of PG channels, chunk j of a group takes
    // columns g*Gp + 8j, pass-through bytes g*PG + 8j and lands at byte g*Hp + 16j (unit_fused.cu has variants with PG fixed)
    {
      int pg = 0, hp = 0;
      // PG = half of the interleaved columns; Hp = 16 bytes per chunk of a group
      int n_il = 0; bool all_il = true;
      for (int i = 0; i < desc->n_chunks; ++i) { const cdn_pw_chunk& c = desc->chunks[i]; if (c.count > 0) { n_il += c.count; if (c.pass_off < 0) all_il = false; } }
      if (all_il && n_il > 0 && n_il % 2 == 0 && desc->n_chunks % 2 == 0) {
        pg = n_il / 2; hp = 16 * (desc->n_chunks / 2);
        const int Gp = (pg + 7) & ~7;
        bool ok = true;
        for (int g = 0; g < 2 && ok; ++g)
          for (int j = 0; j < hp / 16 && ok; ++j) {
            const int cnt = std::max(0, std::min(8, pg - 8 * j)), dst = g * hp + 16 * j;
            bool found = false;
            for (int i = 0; i < desc->n_chunks && !found; ++i) {
              const cdn_pw_chunk& c = desc->chunks[i];
              if (c.dst_off != dst) continue;
              found = cnt == 0 ? c.count == 0 : (c.count == cnt && c.col == g * Gp + 8 * j && c.pass_off == g * pg + 8 * j);
            }
            ok = found;
          }
        d.il_pg = ok ? pg : 0; d.il_hp = ok ? hp : 0;
      }
    }
    // chunks grouped by N tile, then by 128-byte output segment; every tile owns whole segments
    d.has_pass = 0;
    int seg_cursor = 0;
    for (int t = 0; t < d.n_tiles; ++t) {
      tile_seg[t] = (int)segs.size();
      std::vector<cdn_pw_chunk> mine;
      int smin = 1 << 30, smax = -1;
      for (int i = 0; i < desc->n_chunks; ++i) {
        cdn_pw_chunk c = desc->chunks[i];
        if (c.count == 0 || c.col / d.BN != t) continue;
        const int width = c.pass_off < 0 ? 16 : 8;
        CDN_CHECK(c.col % 8 == 0 && c.col % d.BN + width <= d.BN && c.col + c.count <= d.N && c.count <= width,
                  CDN_ERR_INVALID, "pw: chunk col=%d count=%d crosses the tile/N boundary", c.col, c.count);
        CDN_CHECK(c.dst_off % 16 == 0 && c.dst_off >= 0, CDN_ERR_INVALID, "pw: chunk dst_off must be a multiple of 16");
        if (c.pass_off >= 0) {
          d.has_pass = 1;
          CDN_CHECK(c.pass_off + c.count <= pass_pitch, CDN_ERR_INVALID, "pw: pass offset %d beyond the pass pixel", c.pass_off);
        }
        mine.push_back(c);
        smin = std::min(smin, c.dst_off >> 7); smax = std::max(smax, c.dst_off >> 7);
      }
      CDN_CHECK(smax >= 0, CDN_ERR_INVALID, "pw: N tile %d has no output chunk", t);
      for (int i = 0; i < desc->n_chunks; ++i) {   // zero-fill chunks that fall into this tile's segment range
        cdn_pw_chunk c = desc->chunks[i];
        if (c.count == 0 && (c.dst_off >> 7) >= smin && (c.dst_off >> 7) <= smax) { c.pass_off = -1; c.col = (int16_t)(t * d.BN); mine.push_back(c); }
      }
      CDN_CHECK(smin >= seg_cursor, CDN_ERR_INVALID, "pw: output segments of N tiles overlap (tile %d)", t);
      seg_cursor = smax + 1;
      std::stable_sort(mine.begin(), mine.end(), [](const cdn_pw_chunk& a, const cdn_pw_chunk& b) { return a.dst_off < b.dst_off; });
      for (size_t i = 0; i < mine.size(); ++i) {
        const int sg = mine[i].dst_off >> 7;
        if (segs.size() == (size_t)tile_seg[t] || segs.back().seg != sg) segs.push_back(PwSeg{sg, (int)ch.size(), (int)ch.size(), 0});
        ch.push_back(mine[i]);
        segs.back().ce = (int)ch.size();
      }
    }
    tile_seg[d.n_tiles] = (int)segs.size();
    d.n_chunks = (int)ch.size(); d.n_segs = (int)segs.size();
  }
  d.thr = (float)min_thr;
  if (dev_upload(&d.chunks, ch.data(), ch.size())) return CDN_ERR_CUDA;
  if (dev_upload((PwSeg**)&d.segs, segs.data(), segs.size())) return CDN_ERR_CUDA;
  if (dev_upload(&d.tile_seg, tile_seg.data(), tile_seg.size())) return CDN_ERR_CUDA;
  if (dev_upload((float4**)&d.kc, kc.data(), kc.size())) return CDN_ERR_CUDA;
  // shared memory plan (mirrors the carve in the kernel)
  int pass_need = 0;
  if (d.has_pass) for (int i = 0; i < desc->n_chunks; ++i) if (desc->chunks[i].pass_off >= 0) pass_need = std::max(pass_need, (desc->chunks[i].pass_off & ~7) + 16);
  d.pass_segs = d.has_pass ? (pass_need + 127) / 128 : 0;
  const size_t b_blk = ((size_t)d.BN * PW_BK + 1023) & ~(size_t)1023;
  const size_t b_all = (size_t)d.num_k_blocks * Np * PW_BK;
  const size_t tables = (size_t)Np * 16 + (size_t)((d.n_chunks + 1) & ~1) * 8 + (size_t)d.n_segs * 16 + (size_t)((d.n_tiles + 1 + 3) & ~3) * 4 + 56 * 8;
  auto plan = [&](bool resident, int groups, int pass_bufs, int nbuf, int& stages) {
    const size_t fixed = 1024 + (resident ? b_all : 0) + (size_t)pass_bufs * d.pass_segs * 16384 + (size_t)groups * nbuf * 16384 + tables;
    const size_t stage_bytes = 16384 + (resident ? 0 : b_blk);
    if (fixed + 2 * stage_bytes > PW_SMEM_LIMIT) { stages = 0; return (size_t)0; }
    stages = (int)std::min<size_t>(PW_MAX_STAGES, (PW_SMEM_LIMIT - fixed) / stage_bytes);
    return fixed + (size_t)stages * stage_bytes;
  };
  // preference: resident weights, double-buffered pass tile, and at least two tiles' worth of activation stages;
  // relax one requirement at a time until the layer fits
  int stages = 0; size_t bytes = 0; bool found = false;
  const int nbuf_hi = d.n_f32 > 0 ? 0 : ((g_cdn_debug_flags & 32u) ? 2 : 3), nbuf_lo = d.n_f32 > 0 ? 0 : 2;
  for (int groups = PW_EPI_GROUPS; groups >= 1 && !found; --groups)   // one epilogue group only when nothing else fits
    for (int want = 2; want >= 0 && !found; --want)          // tiles in flight we insist on (0: anything that runs)
      for (int resident = 1; resident >= 0 && !found; --resident)
        // every epilogue group needs its own pass buffer(s): with a shared one a group could be two mbarrier phases behind
        for (int pass_bufs = d.has_pass ? 2 * groups : groups; pass_bufs >= groups && !found; pass_bufs -= groups)
          for (int nbuf = nbuf_hi; nbuf >= nbuf_lo && !found; --nbuf) {
            int st = 0; const size_t by = plan(resident != 0, groups, pass_bufs, nbuf, st);
            const int need = std::max(2, std::min(want * d.num_k_blocks, want == 2 ? 8 : 3));
            if (st >= need) { found = true; stages = st; bytes = by; d.resident = resident; d.pass_bufs = pass_bufs; d.nbuf = nbuf; d.groups = groups; }
          }
  CDN_CHECK(stages >= 2, CDN_ERR_INVALID, "pw: layer does not fit in shared memory (K=%d N=%d BN=%d pass segs=%d)", d.K, d.N, d.BN, d.pass_segs);
  d.stages = stages;
  d.smem_bytes = bytes;
  if (int r = make_tmap_2d(&d.tmB, d.w, (uint64_t)d.Kp, (uint64_t)Np, (uint64_t)d.Kp, (uint32_t)d.BN)) return r;
  return 0;
}

unsigned long long* g_pw_dbg = nullptr;
static int g_pw_launch_index = 0;          // experiments: flags bits 8..15 select ONE launch (1-based) to instrument
extern "C" int cdn_debug_read_cycles(unsigned long long* out16, int reset) {
  if (!g_pw_dbg) { CDN_CUDA(cudaMalloc(&g_pw_dbg, 16 * 8)); CDN_CUDA(cudaMemset(g_pw_dbg, 0, 16 * 8)); }
  CDN_CUDA(cudaDeviceSynchronize());
  if (out16) CDN_CUDA(cudaMemcpy(out16, g_pw_dbg, 16 * 8, cudaMemcpyDeviceToHost));
  if (reset) { CDN_CUDA(cudaMemset(g_pw_dbg, 0, 16 * 8)); g_pw_launch_index = 0; }
  return 0;
}

int pw_init_attrs() {
  static bool attr_set[64] = {};
  if (cdn_first_on_device(attr_set)) {
    CDN_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_LIMIT));
    CDN_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_LIMIT));
    CDN_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_LIMIT));
    CDN_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_LIMIT));
    CDN_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_LIMIT));
  }
  return 0;
}

int pw_launch(const PwDevice& d, const int8_t* in, int in_pitch, long long pixels, const int8_t* pass, int pass_pitch,
              int8_t* out, int out_pitch, float* out_f32, int ppi, const CUtensorMap* tmA, const CUtensorMap* tmP,
              const CUtensorMap* tmO, cudaStream_t st) {
  if (pixels == 0) return 0;
  CDN_CHECK(in_pitch % 16 == 0 && d.k_off + d.K <= in_pitch, CDN_ERR_INVALID, "pw: input pitch %d too small for k_off+K=%d", in_pitch, d.k_off + d.K);
  CDN_CHECK(!d.has_pass || pass, CDN_ERR_INVALID, "pw: interleaving chunks need a pass-through tensor");
  CDN_CHECK(d.n_f32 > 0 ? (out_f32 != nullptr && ppi > 0) : (out != nullptr), CDN_ERR_INVALID, "pw: missing output pointer");
  PwParams p; memset(&p, 0, sizeof(p));
  p.num_k_blocks = d.num_k_blocks; p.k_off = d.k_off; p.BN = d.BN; p.n_tiles = d.n_tiles; p.stages = d.stages;
  p.groups = d.groups; p.has_pass = d.has_pass; p.pass_segs = d.pass_segs; p.pass_bufs = d.pass_bufs; p.nbuf = d.nbuf; p.resident = d.resident;
  p.n_chunks = d.n_chunks; p.n_segs = d.n_segs;
  p.m_tiles = (pixels + PW_BM - 1) / PW_BM; p.pixels = pixels;
  p.chunks = d.chunks; p.segs = (const PwSeg*)d.segs; p.tile_seg = d.tile_seg; p.kc = (const float4*)d.kc;
  p.M = d.rq.M; p.B = d.rq.B; p.acc_bias = d.rq.acc_bias;
  p.lo_f = (float)d.rq.lo; p.thr = d.thr; p.use_int = d.use_int; p.lo_i = d.rq.lo; p.dbg = g_cdn_debug_flags; p.dbg_cyc = g_pw_dbg; if (!g_pw_dbg) p.dbg &= ~16u;
  { const int sel = (int)((g_cdn_debug_flags >> 8) & 0xff); ++g_pw_launch_index; if (sel && sel != g_pw_launch_index) p.dbg &= ~16u; }
  p.n_f32 = d.n_f32; p.ppi = ppi; p.out_f32 = out_f32; p.Mf = d.Mf; p.bf = d.bf;
  p.in = in; p.in_pitch = in_pitch; p.pass = pass; p.pass_pitch = pass_pitch; p.out = out; p.out_pitch = out_pitch;
  p.w = d.w; p.Kp = d.Kp; p.N = d.BN * d.n_tiles;
  if (g_cdn_debug_flags & 1u) {
    const int per_row = d.n_f32 > 0 ? p.N / 16 : d.n_chunks;
    long long total = pixels * per_row;
    pw_gemm_simt_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p, d.n_chunks);
    CDN_LAUNCH_CHECK("pw_gemm_simt_kernel");
    return 0;
  }
  CUtensorMap a, pm, o;
  if (tmA) a = *tmA; else if (int r = make_tmap_2d(&a, in, (uint64_t)in_pitch, (uint64_t)pixels, (uint64_t)in_pitch, PW_BM)) return r;
  if (d.has_pass) { if (tmP) pm = *tmP; else if (int r = make_tmap_2d(&pm, pass, (uint64_t)pass_pitch, (uint64_t)pixels, (uint64_t)pass_pitch, PW_BM)) return r; }
  else pm = a;
  if (d.n_f32 == 0) { if (tmO) o = *tmO; else if (int r = make_tmap_2d(&o, out, (uint64_t)out_pitch, (uint64_t)pixels, (uint64_t)out_pitch, PW_BM)) return r; }
  else o = a;
  if (int r = pw_init_attrs()) return r;
  long long tiles = p.m_tiles * p.n_tiles;
  CDN_CHECK(tiles < (1ll << 31) && pixels < (1ll << 31), CDN_ERR_INVALID, "pw: too many pixels for 32-bit tile indexing");
  int grid = (int)std::min<long long>(tiles, cdn_num_sms());
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PW_THREADS); cfg.dynamicSmemBytes = d.smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;   // bit 6: disable PDL (experiments)
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (d.n_f32 > 0) CDN_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<3>, a, d.tmB, pm, o, p));
  else if (!d.use_int) CDN_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<0>, a, d.tmB, pm, o, p));
  else if (d.rq.lo > -128) CDN_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<2>, a, d.tmB, pm, o, p));
  else if (d.sh0 && !(g_cdn_debug_flags & (1u << 22))) CDN_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<4>, a, d.tmB, pm, o, p));   // bit 22: keep the shift (A/B)
  else CDN_CUDA(cudaLaunchKernelEx(&cfg, pw_gemm_tc_kernel<1>, a, d.tmB, pm, o, p));
  CDN_LAUNCH_CHECK("pw_gemm_tc_kernel");
  return 0;
}
