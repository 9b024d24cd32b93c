// 1x1 convolutions as int8 tcgen05 GEMMs (sm_100a): D[pixel][n] = sum_k A[pixel][k] * W[n][k], s8 x s8 -> s32 in TMEM.
//
//   * persistent, warp-specialised CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
//     warps 2..5 = epilogue (each owns the 32 TMEM lanes of its quarter)
//   * A (activations, NHWC rows = pixels) and W (weights, K-major) arrive by TMA as [rows x 128 B] boxes with the
//     128-byte swizzle, 128 bytes of K per stage; UMMA 128 x BN x 32, four per stage; accumulators double-buffered
//     in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1
//   * epilogue: tcgen05.ld -> exact requantisation (common.cuh) -> bytes are placed by the layer's chunk table
//     (plain / interleaved with a pass-through tensor = split + cat + channel_shuffle) into a swizzled staging tile
//     -> TMA store.  Head convs instead write fp32 NCHW planes.
//   * a SIMT dp4a kernel with identical semantics exists for bring-up and as an on-device cross-check
//     (cdn_set_debug_flags bit 0); it is not a fallback: nothing selects it automatically.
#include "layers.cuh"
#include <algorithm>

#define PW_BM 128
#define PW_BK 128
#define PW_MAX_BN 256
#define PW_THREADS 192
#define PW_SPIN_LIMIT (1u << 26)

struct PwParams {
  int num_k_blocks, k_off, BN, n_tiles, stages, has_pass, pass_segs, max_segs;
  long long m_tiles, pixels;
  const cdn_pw_chunk* chunks; const int* chunk_begin;  // chunk_begin[n_tiles + 1]
  const int* seg_begin;                                // [2*t] first 128-byte output segment of N tile t, [2*t+1] count
  const float* Mh; const float* Bh; const float* thr; const double* M; const double* B; const int32_t* acc_bias;
  float lo_f;
  // fp32 head output
  int n_f32, ppi; float* out_f32; const double* Mf; const double* bf;
  // SIMT cross-check path
  const int8_t* in; int in_pitch; const int8_t* pass; int pass_pitch; int8_t* out; int out_pitch;
  const int8_t* w; int Kp, N;
};

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > PW_SPIN_LIMIT) { printf("cdn pw_gemm: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), canonical 1
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// shared epilogue math: 16 (or 8 interleaved) output bytes of one chunk from accumulators
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rq_col(const PwParams& p, int acc, int n) {
  return requant_bits(acc + __ldg(p.acc_bias + n), __ldg(p.Mh + n), __ldg(p.Bh + n), __ldg(p.thr + n), p.lo_f, p.M, p.B, n);
}

// acc[0..count) -> out words; passb: 8 pass-through bytes packed in two words (interleave mode)
__device__ __forceinline__ uint4 chunk_bytes(const PwParams& p, const cdn_pw_chunk& ck, const uint32_t (&acc)[16],
                                             uint32_t pass_lo, uint32_t pass_hi) {
  uint32_t q[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] = (i < ck.count) ? rq_col(p, (int)acc[i], ck.col + i) : 0u;
  uint4 o;
  if (ck.pass_off < 0) {
    o.x = pack4_lowbytes(q[0], q[1], q[2], q[3]);   o.y = pack4_lowbytes(q[4], q[5], q[6], q[7]);
    o.z = pack4_lowbytes(q[8], q[9], q[10], q[11]); o.w = pack4_lowbytes(q[12], q[13], q[14], q[15]);
  } else {
    uint32_t n_lo = pack4_lowbytes(q[0], q[1], q[2], q[3]), n_hi = pack4_lowbytes(q[4], q[5], q[6], q[7]);
    // out[2i] = pass[i], out[2i+1] = new[i]
    o.x = __byte_perm(pass_lo, n_lo, 0x5140); o.y = __byte_perm(pass_lo, n_lo, 0x7362);
    o.z = __byte_perm(pass_hi, n_hi, 0x5140); o.w = __byte_perm(pass_hi, n_hi, 0x7362);
    // zero the tail beyond 2*count bytes
    int nb = 2 * ck.count;
    uint32_t* ow = &o.x;
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      int rem = nb - 4 * wd;
      ow[wd] = rem >= 4 ? ow[wd] : (rem <= 0 ? 0u : (ow[wd] & (0xffffffffu >> (8 * (4 - rem)))));
    }
  }
  return o;
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PW_THREADS, 1)
pw_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmO,
                  const PwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16 KB | B BN*128)] [pass: pass_segs x 16 KB] [out staging: max_segs x 16 KB] [barriers]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = PW_BM * PW_BK, b_bytes = (uint32_t)p.BN * PW_BK;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  uint8_t* s_pass = smem + (size_t)p.stages * stage_bytes;
  uint8_t* s_out = s_pass + (size_t)p.pass_segs * 16384;
  uint64_t* bars = (uint64_t*)(s_out + (size_t)p.max_segs * 16384);
  // barrier slots: full[8], empty[8], tmem_full[2], tmem_empty[2], pass_full, pass_empty, tmem_ptr
  const uint32_t bar_base = smem_u32(bars);
  auto FULL = [&](int s) { return bar_base + 8u * s; };
  auto EMPTY = [&](int s) { return bar_base + 8u * (8 + s); };
  auto TFULL = [&](int s) { return bar_base + 8u * (16 + s); };
  auto TEMPTY = [&](int s) { return bar_base + 8u * (18 + s); };
  const uint32_t PFULL = bar_base + 8u * 20, PEMPTY = bar_base + 8u * 21;
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(TFULL(s), 1); mbar_init(TEMPTY(s), 128); }
    mbar_init(PFULL, 1); mbar_init(PEMPTY, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {                           // TMEM: the whole 512 columns (1 CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, pphase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long mt = tile / p.n_tiles; const int nt = (int)(tile % p.n_tiles);
        if (p.has_pass) {
          mbar_wait(PEMPTY, pphase ^ 1);
          mbar_expect_tx(PFULL, (uint32_t)p.pass_segs * 16384u);
          for (int s = 0; s < p.pass_segs; ++s)
            tma_load_2d(smem_u32(s_pass + (size_t)s * 16384), &tmP, s * 128, (int)(mt * PW_BM), PFULL);
          pphase ^= 1;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(EMPTY(stage), phase ^ 1);
          mbar_expect_tx(FULL(stage), a_bytes + b_bytes);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          tma_load_2d(smem_u32(sa), &tmA, p.k_off + kb * PW_BK, (int)(mt * PW_BM), FULL(stage));
          tma_load_2d(smem_u32(sa + a_bytes), &tmB, kb * PW_BK, nt * p.BN, FULL(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: S32 accumulate, A/B signed int8, K-major both, N = BN, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(PW_BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(TEMPTY(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * PW_MAX_BN);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(FULL(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < PW_BK / 32; ++k)
            umma_i8(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(EMPTY(stage));           // frees the smem stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(TFULL(as));                // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;             // row inside the tile
    const int et = threadIdx.x - 64;           // 0..127
    int as = 0; uint32_t aphase = 0, pphase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const long long mt = tile / p.n_tiles; const int nt = (int)(tile % p.n_tiles);
      mbar_wait(TFULL(as), aphase);
      tc_fence_after();
      if (p.has_pass) mbar_wait(PFULL, pphase);
      const uint32_t tacc = tmem_base + (uint32_t)(as * PW_MAX_BN) + ((uint32_t)(q * 32) << 16);
      const int c_begin = p.chunk_begin[nt], c_end = p.chunk_begin[nt + 1];
      const int seg0 = p.seg_begin[2 * nt], nseg = p.seg_begin[2 * nt + 1];
      if (p.n_f32 > 0) {
        // fp32 NCHW planes: out[img][n][pix] = acc*Mf[n] + bf[n]
        const long long pix = mt * PW_BM + row;
        const long long img = pix / p.ppi; const int pi = (int)(pix - img * p.ppi);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          uint32_t acc[16];
          tmem_ld16(tacc + (uint32_t)c0, acc);
          tmem_ld_wait();
          if (pix < p.pixels) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              int n = nt * p.BN + c0 + i;
              if (n < p.n_f32) {
                double y = __dadd_rn(__dmul_rn((double)((int)acc[i] + __ldg(p.acc_bias + n)), __ldg(p.Mf + n)), __ldg(p.bf + n));
                p.out_f32[((size_t)img * p.n_f32 + n) * p.ppi + pi] = (float)y;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(TEMPTY(as));
      } else {
        // the previous tile's TMA stores must have finished reading the staging buffer
        if (et == 0) tma_store_wait_read();
        named_bar_sync(1, 128);
        for (int c = c_begin; c < c_end; ++c) {
          const cdn_pw_chunk ck = p.chunks[c];
          uint32_t acc[16];
          uint32_t pass_lo = 0, pass_hi = 0;
          if (ck.count > 0) {
            const uint32_t ta = tacc + (uint32_t)(ck.col - nt * p.BN);
            if (ck.pass_off < 0) tmem_ld16(ta, acc); else tmem_ld8(ta, acc);
          }
          if (ck.pass_off >= 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              int off = ck.pass_off + i;
              uint32_t byte = 0;
              if (i < ck.count) {
                const uint8_t* seg = s_pass + (size_t)(off >> 7) * 16384 + row * 128;
                byte = seg[(((off & 127) >> 4) ^ (row & 7)) << 4 | (off & 15)];
              }
              if (i < 4) pass_lo |= byte << (8 * i); else pass_hi |= byte << (8 * (i - 4));
            }
          }
          tmem_ld_wait();
          uint4 o = chunk_bytes(p, ck, acc, pass_lo, pass_hi);
          const int seg = (ck.dst_off >> 7) - seg0, j = (ck.dst_off & 127) >> 4;
          *(uint4*)(s_out + (size_t)seg * 16384 + row * 128 + ((j ^ (row & 7)) << 4)) = o;
        }
        tc_fence_before();
        mbar_arrive(TEMPTY(as));
        if (p.has_pass) mbar_arrive(PEMPTY);
        fence_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
          for (int s = 0; s < nseg; ++s)
            tma_store_2d(&tmO, (seg0 + s) * 128, (int)(mt * PW_BM), smem_u32(s_out + (size_t)s * 16384));
          tma_store_commit();
        }
      }
      if (p.has_pass) pphase ^= 1;
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
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
  for (int i = 0; i < 16; ++i) {
    int s = 0;
    if (i < ck.count) {
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
      uint32_t byte = (i < ck.count) ? (uint8_t)p.pass[(size_t)pix * p.pass_pitch + ck.pass_off + i] : 0u;
      if (i < 4) pass_lo |= byte << (8 * i); else pass_hi |= byte << (8 * (i - 4));
    }
  uint4 o = chunk_bytes(p, ck, acc, pass_lo, pass_hi);
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

void pw_device_free(PwDevice& d) {
  cudaFree(d.w); cudaFree(d.chunks); cudaFree(d.chunk_begin); cudaFree(d.seg_begin); cudaFree(d.Mf); cudaFree(d.bf);
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
  CDN_CHECK((long long)d.K * 8 * 383 < (1 << 24), CDN_ERR_INVALID, "pw: K=%d too large for exact fp32 conversion of the accumulator", d.K);
  const int Np = d.BN * d.n_tiles;
  // weights, zero padded to [Np][Kp]; acc_bias = zx * sum_k w
  std::vector<int8_t> w((size_t)Np * d.Kp, 0);
  std::vector<int32_t> ab(Np, 0);
  for (int n = 0; n < d.N; ++n) {
    int sum = 0;
    for (int k = 0; k < d.K; ++k) { int8_t v = desc->wq[(size_t)n * d.K + k]; w[(size_t)n * d.Kp + k] = v; sum += v; }
    ab[n] = desc->zx * sum;
  }
  if (dev_upload(&d.w, w.data(), w.size())) return CDN_ERR_CUDA;
  if (d.n_f32 > 0) {
    CDN_CHECK(desc->Mf && desc->bf && d.n_f32 <= d.N, CDN_ERR_INVALID, "pw: fp32 output needs Mf/bf and n_f32 <= N");
    std::vector<double> Mf(Np, 0.0), bf(Np, 0.0), one(Np, 1.0), zero(Np, 0.0);
    for (int n = 0; n < d.n_f32; ++n) { Mf[n] = desc->Mf[n]; bf[n] = desc->bf[n]; }
    if (dev_upload(&d.Mf, Mf.data(), Np)) return CDN_ERR_CUDA;
    if (dev_upload(&d.bf, bf.data(), Np)) return CDN_ERR_CUDA;
    cdn_requant dummy{one.data(), zero.data(), -128, Np};
    if (int r = dev_requant_upload(d.rq, &dummy, ab.data(), Np)) return r;
    d.max_segs = 0; d.has_pass = 0; d.n_chunks = 0;
    std::vector<int> zb(2 * d.n_tiles + 2, 0);
    if (dev_upload(&d.chunk_begin, zb.data(), zb.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.seg_begin, zb.data(), zb.size())) return CDN_ERR_CUDA;
  } else {
    CDN_CHECK(desc->rq.n == d.N && desc->rq.M && desc->rq.B, CDN_ERR_INVALID, "pw: requant constants must have n == N");
    CDN_CHECK(desc->chunks && desc->n_chunks > 0, CDN_ERR_INVALID, "pw: int8 output needs a chunk table");
    if (int r = dev_requant_upload(d.rq, &desc->rq, ab.data(), Np)) return r;
    // sort chunks by N tile, validate, derive the 128-byte output segments each tile owns
    std::vector<cdn_pw_chunk> ch;
    std::vector<int> cb(d.n_tiles + 1, 0), sb(2 * d.n_tiles, 0);
    d.has_pass = 0; d.max_segs = 0;
    int seg_cursor = 0;
    for (int t = 0; t < d.n_tiles; ++t) {
      cb[t] = (int)ch.size();
      int smin = 1 << 30, smax = -1;
      for (int i = 0; i < desc->n_chunks; ++i) {
        cdn_pw_chunk c = desc->chunks[i];
        int tile = c.count > 0 ? c.col / d.BN : -1;
        if (c.count == 0) {                  // zero-fill chunk: give it to the tile owning its segment (decided below)
          continue;
        }
        if (tile != t) continue;
        CDN_CHECK(c.col % 8 == 0 && c.col + c.count <= (t + 1) * d.BN && c.col + c.count <= d.N,
                  CDN_ERR_INVALID, "pw: chunk col=%d count=%d crosses the tile/N boundary", c.col, c.count);
        CDN_CHECK(c.count <= (c.pass_off < 0 ? 16 : 8) && c.dst_off % 16 == 0, CDN_ERR_INVALID, "pw: bad chunk");
        if (c.pass_off >= 0) { d.has_pass = 1; CDN_CHECK(c.pass_off + c.count <= pass_pitch, CDN_ERR_INVALID, "pw: pass offset beyond pass pitch"); }
        ch.push_back(c);
        smin = std::min(smin, c.dst_off >> 7); smax = std::max(smax, c.dst_off >> 7);
      }
      CDN_CHECK(smax >= 0, CDN_ERR_INVALID, "pw: N tile %d has no output chunk", t);
      // zero-fill chunks that fall into this tile's segment range
      for (int i = 0; i < desc->n_chunks; ++i) {
        cdn_pw_chunk c = desc->chunks[i];
        if (c.count == 0 && (c.dst_off >> 7) >= smin && (c.dst_off >> 7) <= smax) { c.pass_off = -1; c.col = (int16_t)(t * d.BN); ch.push_back(c); }
      }
      CDN_CHECK(smin >= seg_cursor, CDN_ERR_INVALID, "pw: output segments of N tiles overlap (tile %d)", t);
      seg_cursor = smax + 1;
      sb[2 * t] = smin; sb[2 * t + 1] = smax - smin + 1;
      d.max_segs = std::max(d.max_segs, smax - smin + 1);
    }
    cb[d.n_tiles] = (int)ch.size();
    // a TMEM 16-column load must stay inside the accumulator stage
    for (auto& c : ch) CDN_CHECK((c.col % d.BN) + (c.pass_off < 0 ? 16 : 8) <= PW_MAX_BN, CDN_ERR_INVALID, "pw: chunk column overflow");
    d.n_chunks = (int)ch.size();
    if (dev_upload(&d.chunks, ch.data(), ch.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.chunk_begin, cb.data(), cb.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.seg_begin, sb.data(), sb.size())) return CDN_ERR_CUDA;
  }
  if (int r = make_tmap_2d(&d.tmB, d.w, (uint64_t)d.Kp, (uint64_t)Np, (uint64_t)d.Kp, (uint32_t)d.BN)) return r;
  // shared memory budget
  const size_t stage_bytes = PW_BM * PW_BK + (((size_t)d.BN * PW_BK + 1023) & ~(size_t)1023);
  int pass_need = 0;
  if (d.has_pass) for (int i = 0; i < desc->n_chunks; ++i) if (desc->chunks[i].pass_off >= 0) pass_need = std::max(pass_need, desc->chunks[i].pass_off + desc->chunks[i].count);
  const int pass_segs = d.has_pass ? (pass_need + 127) / 128 : 0;
  d.pass_segs = pass_segs;
  const size_t fixed = (size_t)(pass_segs + d.max_segs) * 16384 + 256 + 1024;
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages > d.num_k_blocks * 2 && stages > 2) stages = std::max(2, d.num_k_blocks * 2);
  CDN_CHECK(stages >= 2, CDN_ERR_INVALID, "pw: layer does not fit in shared memory (BN=%d, out segs=%d, pass segs=%d)", d.BN, d.max_segs, pass_segs);
  d.stages = stages;
  d.smem_bytes = (size_t)stages * stage_bytes + fixed;
  return 0;
}

int pw_init_attrs() {
  static bool attr_set = false;
  if (!attr_set) {
    CDN_CUDA(cudaFuncSetAttribute(pw_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
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
  p.has_pass = d.has_pass; p.pass_segs = d.pass_segs; p.max_segs = d.max_segs;
  p.m_tiles = (pixels + PW_BM - 1) / PW_BM; p.pixels = pixels;
  p.chunks = d.chunks; p.chunk_begin = d.chunk_begin; p.seg_begin = d.seg_begin;
  p.Mh = d.rq.Mh; p.Bh = d.rq.Bh; p.thr = d.rq.thr; p.M = d.rq.M; p.B = d.rq.B; p.acc_bias = d.rq.acc_bias;
  p.lo_f = (float)d.rq.lo;
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
  int grid = (int)std::min<long long>(tiles, cdn_num_sms());
  pw_gemm_tc_kernel<<<grid, PW_THREADS, d.smem_bytes, st>>>(a, d.tmB, pm, o, p);
  CDN_LAUNCH_CHECK("pw_gemm_tc_kernel");
  return 0;
}
