// 1x1 convolution of the float model on the tensor cores: fp32 in, fp32 out, every product formed from a 3-way TF32 split.
//
// Replaces the reference's nn.Conv2d(k=1) [+ folded BatchNorm] [+ ReLU] of the float PoseShuffleNetV2
// (lib/models/networks/shufflenetv2_dcn.py:63-99 BaseNode branches, :209-216 conv5, :244-271 heads, :286-300 deconv
// conv_channel) on channel slices of NCHW tensors, like cdn_pw_slice_f32 (f32_net.cu) whose SIMT kernel stays the
// fallback for pixel counts that are not a multiple of 256.
//
// Arithmetic.  x = xh + xl, w = wh + wl with xh / wh the value rounded to TF32 (11 significant bits, cvt.rna) and xl / wl the
// exact fp32 remainder (read by the tensor core with 11 significant bits).  Each k-step issues three tcgen05.mma kind::tf32
// into one fp32 TMEM accumulator: xl*wh + xh*wl + xh*wh.  xh*wh is exact in fp32 (22-bit product), the cross terms carry a
// relative error <= 2^-11 on a term that is <= 2^-11 of the product, the dropped xl*wl is <= 2^-22 of it: every product
// is within ~2^-21 of exact (fp32 FMA chains of the reference round the running sum to 2^-24 at every step instead).
//
// Shape.  D[256 pixels][BN channels] per tile = X^T W^T: the activation tile is the A operand in MN-major form.  32-bit
// MN-major operands have one legal layout (SWIZZLE_128B with a 32-byte atom, UMMA layout type 1) -- exactly what the TMA
// swizzle mode 128B_ATOM_32B writes for a box {32 pixels, 16 channels, 8 pixel groups} of the NCHW tensor, so no transposition
// is ever executed.  The packed weights (cdn_pw_tf32x3_pack: 128-byte rows of 16 channels hi + the same 16 lo) are the K-major
// B operand (SWIZZLE_128B).  Warp roles: 0 TMA producer, 1-2 MMA issuers (one per 128-pixel block), 4-7 split the raw
// activation tile in place (hi) and into its sibling buffer (lo), 8-15 epilogue (lane = pixel, 184 registers by setmaxnreg).
// Accumulation.  The tensor core adds into its fp32 accumulator with truncation; over a long K that is a systematic shrink
// (measured 7e-9 * K relative).  So it only ever sums ONE 16-channel stage: the cross terms first (small), then the two exact
// hi*hi products; the epilogue warps move every such chunk sum into register totals with round-to-nearest fp32 adds
// (2 or 4 chunk accumulators in TMEM rotate, so the tensor core runs ahead), and bias / ReLU / the stores happen once per
// tile through a shared-memory slab (rounds of 16 channels, 16-byte coalesced stores).  Measured error of one layer: 1.0-1.3e-7 relative L2, independent of K.
// Tried and dropped: a deeper raw-activation ring (6 x 16 KB in flight) with double-buffered operand stages -- the 128 x 128
// layers did not move (they are not bound by bytes in flight), the compute-heavy ones lost 5-15 %.
// Bound.  Per 16-channel stage the shared memory moves 176 KB (tensor-core operand reads 96, TMA writes 32, split 48) = 1375
// clocks at 128 B/clock against 768 clocks of tensor time: the compute-heavy layers run at 50-55 % of the kind::tf32 rate
// (x3 instruction slots), the 128 x 128 layers at 0.4-0.5 of HBM.
#include "tc_ptx.cuh"
#include "layers.cuh"
#include <cuda.h>
#include <algorithm>

#define PT_M 256                       // pixels per tile = two M = 128 accumulators
#define PT_KC 16                       // channels per pipeline stage (two K = 8 steps)
#define PT_STAGES 4
#define PT_XBYTES (PT_M * PT_KC * 4)   // 16 KB: raw / hi activation tile of a stage; the same again for lo
#define PT_WMAX (128 * 128)             // 16 KB: weight tile of a stage at BN = 128 (rows of 16 hi + 16 lo floats)
#define PT_STAGE_BYTES (2 * PT_XBYTES + PT_WMAX)
#define PT_THREADS 512                 // 16 warps: TMA, 2 x MMA, 1 idle | 4 split | 8 epilogue
#define PT_EPI_WARPS 8

struct PtParams {
  const float* bias;                   // [Co] or null
  float* out;
  int C, Co, BN, NT, per;              // input channels, output channels, columns per N tile (multiple of 16), N tiles, channels per N tile
  int out_ctotal, out_coff, out_cstride, relu;
  int ppi, tiles_per_img;              // pixels per image, ppi / 256
  unsigned total_tiles;                // batch * tiles_per_img * NT
  int num_k;                           // ceil(C / 16)
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A: MN-major 32-bit operands have ONE legal shared-memory layout, SWIZZLE_128B_BASE32B (cute::UMMA::LayoutType 1, what the TMA
// swizzle mode 128B_ATOM_32B writes): 128-byte rows = 32 pixels of one channel, rows 128 bytes apart, the four 32-byte chunks of
// a row XORed with (row % 4); an atom is 4 channels (512 bytes).  A K = 8 instruction reads two atoms `SBO` = 512 bytes apart;
// the next 32 pixels (next atom along M) are `lbo` bytes further.
__device__ __forceinline__ uint64_t pt_desc_a(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
  return d;
}
// B: K-major, SWIZZLE_128B.  128-byte rows = [16 channels hi | 16 channels lo] of one output channel, 8-row groups 1 KB apart;
// a K = 8 step is 32 bytes inside the row (hi: 0 / 32, lo: 64 / 96).
__device__ __forceinline__ uint64_t pt_desc_b(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
               "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
// mbarrier wait without a printf call in the loop (a call site would force the 128 live totals of the epilogue through the stack)
__device__ __forceinline__ void pt_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 21)) __trap();
  }
}
// round to TF32 (10 explicit mantissa bits), nearest with ties away from zero = cvt.rna.tf32.f32 for finite inputs, in two
// integer instructions (the cvt itself expands to ~9 with its NaN handling)
__device__ __forceinline__ uint32_t pt_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ bool pt_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int PT_CHUNK>
__global__ void __launch_bounds__(PT_THREADS, 1)
pw_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const PtParams p) {
  extern __shared__ __align__(1024) uint8_t pt_smem[];
  __shared__ __align__(8) unsigned long long s_bar[3 * PT_STAGES + 8];
  __shared__ uint32_t s_tmem;
  uint8_t* ring = (uint8_t*)(((uintptr_t)pt_smem + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define RAW(s) smem_u32(&s_bar[(s)])
#define FULL(s) smem_u32(&s_bar[PT_STAGES + (s)])
#define EMPTY(s) smem_u32(&s_bar[2 * PT_STAGES + (s)])
#define CFULL(b) smem_u32(&s_bar[3 * PT_STAGES + (b)])
#define CEMPTY(b) smem_u32(&s_bar[3 * PT_STAGES + 4 + (b)])
  if (threadIdx.x == 0) {
    for (int s = 0; s < PT_STAGES; ++s) { mbar_init(RAW(s), 1); mbar_init(FULL(s), 128); mbar_init(EMPTY(s), 2); }
    for (int b = 0; b < 4; ++b) { mbar_init(CFULL(b), 2); mbar_init(CEMPTY(b), PT_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint32_t w_bytes = (uint32_t)p.BN * 128;                     // one weight tile: BN rows of (16 hi + 16 lo) floats
  const uint32_t stage_u32 = smem_u32(ring) + PT_STAGES * PT_STAGE_BYTES;   // output staging slab (two 16 KB halves) behind the ring
  const int nch = (p.num_k + PT_CHUNK - 1) / PT_CHUNK;              // accumulation chunks per tile
  // chunk accumulators: two pixel blocks x bnp columns each; 512 TMEM columns hold 2 of them at BN > 64 and 4 at BN <= 64 (the
  // memory-bound layers: the tensor core then runs a whole short tile ahead while the epilogue stores the previous one)
  const bool small = p.BN <= 64;
  const uint32_t bnp = small ? 64u : 128u;
  const uint32_t nbuf_mask = small ? 3u : 1u, nbuf_shift = small ? 2u : 1u;
  auto split_tile = [&](unsigned t, int& b, int& px0, int& nt) {
    nt = (int)(t % (unsigned)p.NT);
    const unsigned pt = t / (unsigned)p.NT;
    b = (int)(pt / (unsigned)p.tiles_per_img);
    px0 = (int)(pt - (unsigned)b * p.tiles_per_img) * PT_M;
  };

  if (warp < 8) {
    // warpgroups 0 (TMA, two MMA issuers, one idle warp) and 1 (split) hand their registers to the two epilogue warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (unsigned t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          int b, px0, nt; split_tile(t, b, px0, nt);
          for (int kc = 0; kc < p.num_k; ++kc) {
            pt_wait(EMPTY(stage), phase ^ 1);
            const uint32_t sx = smem_u32(ring + (size_t)stage * PT_STAGE_BYTES);
            mbar_expect_tx(RAW(stage), (uint32_t)PT_XBYTES + w_bytes);
            // one box = 32 pixels x 16 channels x 8 pixel groups: 128-byte rows, channel rows 128 bytes apart, pixel groups 2 KB apart
            tma_load_4d(sx, &tmX, 0, kc * PT_KC, px0 >> 5, b, RAW(stage));
            tma_load_2d(sx + 2 * PT_XBYTES, &tmW, kc * 32, nt * p.BN, RAW(stage));   // BN rows of 128 bytes: [hi 16 channels | lo 16 channels]
            if (++stage == PT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp <= 2) {
      // ===================== MMA issuers: warp 1 -> pixel block 0, warp 2 -> pixel block 1 =====================
      // One thread cannot feed the tensor core here: twelve K = 8 instructions per 16-channel stage cost ~14 issue slots each on
      // the uniform datapath (measured: 1750 clocks per stage against 768 of tensor time), so each pixel block has its own issuer
      // and every descriptor is a per-stage base plus a compile-time constant (start-address field, 16-byte units, never carries).
      const int mb = warp - 1;
      const bool leader = pt_elect();
      // fp32 accumulate, A / B = TF32, A MN-major, B K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_base = pt_desc_a(smem_u32(ring) + mb * 8192, 2048), b_base = pt_desc_b(smem_u32(ring) + 2 * PT_XBYTES);
      const uint32_t tcol = tmem_base + (uint32_t)mb * bnp;
      int stage = 0; uint32_t phase = 0; uint32_t g = 0;
      for (unsigned t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        for (int c = 0; c < nch; ++c, ++g) {
          const int nst = min(PT_CHUNK, p.num_k - c * PT_CHUNK);
          const int cb = (int)(g & nbuf_mask); const uint32_t cph = (g >> nbuf_shift) & 1;
          pt_wait(CEMPTY(cb), cph ^ 1);                               // the epilogue has drained this chunk accumulator
          const uint32_t tacc = tcol + (uint32_t)cb * 2u * bnp;
          // the cross terms first, while the fresh accumulator is still small: their additions then round at the cross terms'
          // own magnitude; the exact hi*hi products follow as nst * 2 additions per accumulator
          {
            int st = stage; uint32_t ph = phase;
            for (int s = 0; s < nst; ++s) {
              pt_wait(FULL(st), ph);
              tc_fence_after();
              const uint64_t so = (uint64_t)((uint32_t)st * (PT_STAGE_BYTES >> 4));
              const uint64_t ah = a_base + so, bh = b_base + so;
              if (leader) {
                umma_tf32(tacc, ah + (PT_XBYTES >> 4), bh, idesc, s != 0 ? 1u : 0u);         // x_lo * w_hi, channels 0..7
                umma_tf32(tacc, ah, bh + 4, idesc, 1u);                                      // x_hi * w_lo
                umma_tf32(tacc, ah + ((PT_XBYTES + 1024) >> 4), bh + 2, idesc, 1u);          // channels 8..15
                umma_tf32(tacc, ah + (1024 >> 4), bh + 6, idesc, 1u);
              }
              if (++st == PT_STAGES) { st = 0; ph ^= 1; }
            }
          }
          for (int s = 0; s < nst; ++s) {
            const uint64_t so = (uint64_t)((uint32_t)stage * (PT_STAGE_BYTES >> 4));
            const uint64_t ah = a_base + so, bh = b_base + so;
            if (leader) {
              umma_tf32(tacc, ah, bh, idesc, 1u);
              umma_tf32(tacc, ah + (1024 >> 4), bh + 2, idesc, 1u);
              umma_commit(EMPTY(stage));                              // (one of two arrivals) frees the stage when everything issued so far has retired
            }
            if (++stage == PT_STAGES) { stage = 0; phase ^= 1; }
          }
          if (leader) umma_commit(CFULL(cb));
          __syncwarp();
        }
      }
    } else if (warp >= 4) {
      // ===================== split warps: raw fp32 -> (hi in place, lo beside), elementwise, layout-agnostic =====================
      const int tid = threadIdx.x - 128;
      int stage = 0; uint32_t phase = 0;
      for (unsigned t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        for (int kc = 0; kc < p.num_k; ++kc) {
          pt_wait(RAW(stage), phase);
          const uint32_t sx = smem_u32(ring + (size_t)stage * PT_STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < PT_XBYTES / 16 / 128; ++i) {
            const uint32_t a = sx + (uint32_t)(i * 128 + tid) * 16;
            const uint4 v = lds_u128(a);
            uint4 h, l;
            // (The tensor core reads only the top 19 bits of an fp32 word, so the raw value could serve as a truncated "hi" without
            // this write-back: measured 2 % faster and 10 % less accurate than the rounded split; not used.)
            h.x = pt_tf32(__uint_as_float(v.x)); h.y = pt_tf32(__uint_as_float(v.y));
            h.z = pt_tf32(__uint_as_float(v.z)); h.w = pt_tf32(__uint_as_float(v.w));
            l.x = pt_tf32(__uint_as_float(v.x) - __uint_as_float(h.x)); l.y = pt_tf32(__uint_as_float(v.y) - __uint_as_float(h.y));
            l.z = pt_tf32(__uint_as_float(v.z) - __uint_as_float(h.z)); l.w = pt_tf32(__uint_as_float(v.w) - __uint_as_float(h.w));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h.x), "r"(h.y), "r"(h.z), "r"(h.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + PT_XBYTES), "r"(l.x), "r"(l.y), "r"(l.z), "r"(l.w) : "memory");
          }
          fence_async_smem();                    // generic-proxy writes -> visible to the tensor core's async proxy
          mbar_arrive(FULL(stage));
          if (++stage == PT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: lane = pixel; chunk accumulators summed in registers with fp32 round-to-nearest =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    const int q = warp & 3;                                           // TMEM lane quarter this warp may read
    const int ch = (warp - 8) >> 2;
    // BN > 64: this warp owns columns ch * 64 .. + 64 of both pixel blocks; BN <= 64: all (<= 64) columns of pixel block ch
    const int mb0 = small ? ch : 0, nmb = small ? 1 : 2, col0 = small ? 0 : ch * 64;
    uint32_t g = 0;
    for (unsigned t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int b, px0, nt; split_tile(t, b, px0, nt);
      float2 tot[2][32];                                            // column pairs: the flush adds two columns per FADD2
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int j = 0; j < 32; ++j) tot[mb][j] = make_float2(0.f, 0.f);
      for (int c = 0; c < nch; ++c, ++g) {
        const int cb = (int)(g & nbuf_mask); const uint32_t cph = (g >> nbuf_shift) & 1;
        pt_wait(CFULL(cb), cph);
        tc_fence_after();
#pragma unroll
        for (int mbi = 0; mbi < 2; ++mbi) {
          if (mbi < nmb) {
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb * 2u * bnp + (uint32_t)(mb0 + mbi) * bnp + (uint32_t)col0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (col0 + i * 32 < p.BN) {                             // columns past BN (a multiple of 16) hold stale data that is never stored
                uint32_t r0[32];
                tmem_ld32(tbase + (uint32_t)(i * 32), r0);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  tot[mbi][i * 16 + j] = cdn_fadd2(tot[mbi][i * 16 + j], make_float2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1])));
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(CEMPTY(cb));
      }
      // End of tile: the totals leave through a staging slab in shared memory (conflict-free 128-byte rows per warp), from which
      // all eight epilogue warps store 16 bytes per lane (512 contiguous bytes of one channel plane per warp) after adding the
      // bias / applying ReLU.  (Storing straight from the registers -- 128 predicated 4-byte stores per lane, each with its own
      // address -- cost 6 us per tile; handing the slab's 1 KB rows to the bulk-copy engine instead of LDS + STG.128 was measured
      // 1.3-1.4x slower; 61 instructions of 64-bit address arithmetic per stored float4 in the first LDS + STG loop cost 15 %.)
      const int n0 = nt * p.per;
      const int nreal = min(p.per, p.Co - n0);                      // real output channels of this N tile
      const size_t plane = (size_t)p.out_cstride * p.ppi;
      float* obase = p.out + ((size_t)b * p.out_ctotal + p.out_coff + (size_t)n0 * p.out_cstride) * p.ppi + px0;
      const int et = threadIdx.x - 256;                             // 0..255 among the epilogue warps
      const int srow0 = et >> 6, f4 = et & 63;                      // store phase: this thread's float4 of slab rows srow0 + 4 i
      // Rounds of 16 channels through two 16 KB slab halves (one barrier per round: when a thread passes the barrier of round
      // r + 1 every thread has finished storing round r, so round r + 2 may refill that half).  BN > 64: a round takes 8 columns
      // from each column half (all eight warps fill it); BN <= 64: 16 consecutive columns of both pixel blocks.
      named_bar_sync(1, 256);                                       // the previous tile's last round has been stored (it may have used half 0)
#define PT_STORE_ROUND(R, CHAN_OF_ROW)                                                                                    \
      {                                                                                                                   \
        named_bar_sync(1, 256);                                                                                           \
        const uint32_t half_u32 = stage_u32 + (uint32_t)(((R) & 1) * 16384);                                              \
        _Pragma("unroll")                                                                                                 \
        for (int i = 0; i < 4; ++i) {                                                                                     \
          const int row = srow0 + 4 * i, chn = CHAN_OF_ROW;                                                               \
          if (chn < nreal) {                                                                                              \
            float4 v = lds_f4(half_u32 + (uint32_t)((row * 256 + f4 * 4) * 4));                                           \
            const float bv = p.bias ? __ldg(p.bias + n0 + chn) : 0.f;                                                     \
            v.x += bv; v.y += bv; v.z += bv; v.w += bv;                                                                   \
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }   \
            *(float4*)(obase + (size_t)chn * plane + f4 * 4) = v;                                                         \
          }                                                                                                               \
        }                                                                                                                 \
      }
      if (small) {
        const uint32_t sp = stage_u32 + (uint32_t)((mb0 * 128 + q * 32 + lane) * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r * 16 < nreal) {                                       // uniform over the CTA
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 t2 = tot[0][(16 * r + j) >> 1];
              sts_f32(sp + (uint32_t)((r & 1) * 16384 + j * 1024), (j & 1) ? t2.y : t2.x);
            }
            PT_STORE_ROUND(r, 16 * r + row)
          }
        }
      } else {
        const uint32_t sp = stage_u32 + (uint32_t)((ch * 8 * 256 + q * 32 + lane) * 4);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (r * 8 < nreal) {                                        // uniform over the CTA (the low column half is never empty then)
#pragma unroll
            for (int mbi = 0; mbi < 2; ++mbi)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 t2 = tot[mbi][(8 * r + j) >> 1];
                sts_f32(sp + (uint32_t)((r & 1) * 16384 + j * 1024 + mbi * 512), (j & 1) ? t2.y : t2.x);
              }
            PT_STORE_ROUND(r, (row < 8 ? 8 * r + row : 56 + 8 * r + row))
          }
        }
      }
#undef PT_STORE_ROUND
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
#undef RAW
#undef FULL
#undef EMPTY
#undef CFULL
#undef CEMPTY
}

// ---- weight packing ----------------------------------------------------------------------------------------------------
static inline int pt_nt(int Co) { return (Co + 127) / 128; }
static inline int pt_bn(int Co) { const int nt = pt_nt(Co); return (((Co + nt - 1) / nt) + 15) & ~15; }
static inline int pt_kpad(int C) { return (C + PT_KC - 1) / PT_KC * PT_KC; }

extern "C" size_t cdn_pw_tf32x3_packed_floats(int Co, int C) {
  if (Co < 1 || C < 1) return 0;
  return (size_t)pt_nt(Co) * pt_bn(Co) * pt_kpad(C) * 2;
}

// packed[row][kc][0..15] = hi, [16..31] = lo of channels kc*16 .. +16 of the row's output channel (zero padding everywhere else):
// one 128-byte TMA row per (output channel, 16 input channels)
__global__ void pw_tf32x3_pack_kernel(const float* __restrict__ w, int Co, int C, int rows, int kpad, int bn, float* __restrict__ packed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * kpad * 2) return;
  const int r = (int)(i / (2 * kpad)), col = (int)(i - (long long)r * 2 * kpad);
  const int kc = col >> 5, within = col & 31, k = kc * 16 + (within & 15);
  // packed row r = N tile r / bn, column r % bn; output channel = tile * per_tile + column with per_tile = ceil(Co / NT)
  const int nt = (Co + 127) / 128, per = (Co + nt - 1) / nt;
  const int tile = r / bn, cl = r - tile * bn;
  const int co = tile * per + cl;
  float v = 0.f;
  if (cl < per && co < Co && k < C) v = w[(size_t)co * C + k];
  const float h = __uint_as_float(pt_tf32(v));
  packed[i] = (within < 16) ? h : __uint_as_float(pt_tf32(v - h));
}

extern "C" int cdn_pw_tf32x3_pack(const float* d_w, int Co, int C, float* d_packed, cdn_stream_t stream) {
  CDN_CHECK(d_w && d_packed && Co >= 1 && C >= 1, CDN_ERR_INVALID, "pw_tf32x3_pack: bad arguments");
  const int rows = pt_nt(Co) * pt_bn(Co), kpad = pt_kpad(C);
  const long long n = (long long)rows * kpad * 2;
  pw_tf32x3_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_w, Co, C, rows, kpad, pt_bn(Co), d_packed);
  CDN_LAUNCH_CHECK("pw_tf32x3_pack_kernel");
  return 0;
}

typedef CUresult (*PFN_encodeTiled_pt)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled_pt pt_get_encode() {
  static PFN_encodeTiled_pt fn = nullptr;
  if (!fn) {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled_pt)ptr;
  }
  return fn;
}

extern "C" int cdn_pw_slice_tf32x3(const float* input, int in_ctotal, int in_coff, int C, const float* d_wpacked,
                                   const float* bias, float* output, int out_ctotal, int out_coff, int out_cstride, int Co, int relu,
                                   int B, int pixels_per_image, cdn_stream_t stream) {
  CDN_CHECK(input && d_wpacked && output && C >= 1 && Co >= 1 && in_coff >= 0 && in_coff + C <= in_ctotal && out_cstride >= 1 &&
            out_coff >= 0 && out_coff + (long long)(Co - 1) * out_cstride < out_ctotal, CDN_ERR_INVALID, "pw_slice_tf32x3: channel slice out of range");
  CDN_CHECK(pixels_per_image >= PT_M && pixels_per_image % PT_M == 0, CDN_ERR_INVALID,
            "pw_slice_tf32x3: pixels per image (%d) must be a multiple of %d", pixels_per_image, PT_M);
  CDN_CHECK((((uintptr_t)input | (uintptr_t)output | (uintptr_t)d_wpacked) & 15) == 0, CDN_ERR_INVALID,
            "pw_slice_tf32x3: tensors must be 16-byte aligned");
  CDN_CHECK(B >= 0 && B <= 65535, CDN_ERR_INVALID, "pw_slice_tf32x3: batch out of range");
  if (B == 0) return 0;
  PFN_encodeTiled_pt enc = pt_get_encode();
  CDN_CHECK(enc != nullptr, CDN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  PtParams p;
  p.bias = bias; p.out = output; p.C = C; p.Co = Co; p.NT = pt_nt(Co); p.BN = pt_bn(Co);
  p.out_ctotal = out_ctotal; p.out_coff = out_coff; p.out_cstride = out_cstride; p.relu = relu;
  p.ppi = pixels_per_image; p.tiles_per_img = pixels_per_image / PT_M;
  const unsigned long long tiles = (unsigned long long)B * p.tiles_per_img * p.NT;
  CDN_CHECK(tiles < (1ull << 31), CDN_ERR_INVALID, "pw_slice_tf32x3: too many tiles");
  p.total_tiles = (unsigned)tiles; p.num_k = pt_kpad(C) / PT_KC;
  p.per = (Co + p.NT - 1) / p.NT;              // N tile nt holds channels [nt * per, nt * per + per) in its first `per` columns
  CUtensorMap tmX, tmW;
  {
    // pixels as (32, ppi / 32): dims = {32 pixels, C channels, ppi / 32 pixel groups, B}; the box order (pixels, channels, groups) is
    // the shared-memory layout the A descriptors expect
    cuuint64_t dims[4] = {32, (cuuint64_t)C, (cuuint64_t)pixels_per_image / 32, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)pixels_per_image * 4, 128, (cuuint64_t)in_ctotal * pixels_per_image * 4};
    cuuint32_t box[4] = {32, PT_KC, PT_M / 32, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)(input + (size_t)in_coff * pixels_per_image), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CDN_CHECK(r == CUDA_SUCCESS, CDN_ERR_CUDA, "pw_slice_tf32x3: activation tensor map failed with CUresult %d", (int)r);
  }
  {
    const int kpad = pt_kpad(C);
    cuuint64_t dims[2] = {(cuuint64_t)kpad * 2, (cuuint64_t)p.NT * p.BN};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 8};
    cuuint32_t box[2] = {32, (cuuint32_t)p.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)d_wpacked, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CDN_CHECK(r == CUDA_SUCCESS, CDN_ERR_CUDA, "pw_slice_tf32x3: weight tensor map failed with CUresult %d", (int)r);
  }
  const size_t smem = (size_t)PT_STAGES * PT_STAGE_BYTES + 32768 + 1024;   // ring + output staging slab + alignment slack
  static bool attr_set[64] = {};
  if (cdn_first_on_device(attr_set)) {
    CDN_CUDA(cudaFuncSetAttribute(pw_tf32x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CDN_CUDA(cudaFuncSetAttribute(pw_tf32x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<unsigned long long>(tiles, (unsigned long long)sms);
  // template argument = pipeline stages (of 16 channels) accumulated in the tensor core before the sum moves to registers
  if (g_cdn_debug_flags & (1u << 31)) pw_tf32x3_kernel<2><<<grid, PT_THREADS, smem, (cudaStream_t)stream>>>(tmX, tmW, p);
  else pw_tf32x3_kernel<1><<<grid, PT_THREADS, smem, (cudaStream_t)stream>>>(tmX, tmW, p);
  CDN_LAUNCH_CHECK("pw_tf32x3_kernel");
  return 0;
}
