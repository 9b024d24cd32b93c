// The detection heads' tail as ONE kernel: depthwise 3x3 through the virtual x2 upsample (QuantDepthwiseNode's dw conv +
// BN + ReLU + QuantAct, quant_modules.py:1061-1071) as the A-tile PRODUCER of the block-diagonal 1x1 output conv
// (Quant_Conv2d with fp32 bias), which runs on tcgen05 and writes the fp32 NCHW head planes.  The int8 depthwise output
// ([B,128,128,192] = 805 MB per step at config c) is never written to or re-read from HBM; the arithmetic is that of
// dw3x3_v2_kernel<1,1,true> (dw.cu) followed by pw_gemm_tc_kernel<3>'s epilogue (pw_gemm.cu), bit for bit.
//
// A CTA walks over tiles of 4 x 8 STORED pixels = 8 x 16 output pixels = the 128 rows of one UMMA:
//   * thread (cw, pc) owns one channel word (4 channels) of stored column pc and slides down the 8 stored rows (window
//     of three byte-transposed rows in registers, 2 dp4a per output and channel, integer requantisation), writing its
//     2x2 outputs per stored row as int8 words into the shared-memory A tile in the K-major 128B-swizzled UMMA layout
//     (row m = oy*8 + ox of the tile, K = channel)
//   * one thread issues K/32 tcgen05.mma (kind::i8, M=128, N = padded head planes) against the weight matrix resident in
//     shared memory; the accumulator lives in TMEM
//   * warps 0..3 read their TMEM lane quarter, apply fl32(fl64(acc * Mf) + bf) and store 32-byte runs of the fp32 planes
//   * the stored input of a tile plus its halo (6 x 10 pixels) arrives by ONE 4-D TMA load per tile, two tiles ahead, into
//     a double buffer: the stencil reads shared memory only, so no thread ever waits for DRAM (the first version loaded
//     rows with LDG one row ahead and ran at 0.94 ms, latency-bound; the separate kernels took 0.40 + 0.28 ms)
// Three CTAs per SM overlap each other's phases.
#include "layers.cuh"
#include "tc_ptx.cuh"
#include <algorithm>

#define HF_MAX_THREADS 192
#define HF_CTAS 3                            // resident CTAs per SM (shared memory: ~68 KB each)
#ifndef HF_RQ
#define HF_RQ rq_int_hi
#endif
#define HF_SMEM_MAX (96 * 1024)              // N = 128 head planes: 1024 + 32768 + 32768 + 23040 + 4096 + 64
#define HF_TW 4                              // stored columns per tile
#define HF_TH 8                              // stored rows per tile

struct HfParams {
  int in_pitch;                              // stored input [B][Hs][Ws][pitch] (bytes), read through the tensor map
  int in_bytes;                              // bytes of one TMA box: (HF_TH + 2) x (HF_TW + 2) pixels x pitch
  int Hs, Ws, tiles_x, tiles_y;
  unsigned ntiles;
  int cw_total, nthreads;                    // channel words (K/4); cw_total * HF_TW worker threads
  uint32_t pad_word;
  const uint32_t* wpk;                       // [channel][8] upsample-folded packed tap weights
  const int4* ki; int lo_i;                  // integer requantisation of the depthwise conv
  const int8_t* w; int Kp, NB, ksteps;       // 1x1 weights [NB][Kp], NB = padded N (multiple of 16), K/32 UMMA steps
  int tmem_cols;
  const int32_t* acc_bias; const double* Mf; const double* bf; int n_f32;
  float* out; int ppi, Wout;                 // fp32 planes [B][n_f32][ppi]
};

// LO = 2: no lower clamp AND shift 0 in every channel (DwDevice::sh0): no SHF after the IMAD.HI
template <int LO>
__device__ __forceinline__ uint32_t hf_rq_word(const int (&acc)[4], const int2 (&km)[4], const long long (&kb)[4], int lo) {
  int q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { q[c] = LO == 2 ? rq_int_hi0(acc[c], km[c].x, kb[c]) : HF_RQ(acc[c], km[c].x, km[c].y, kb[c]); if (LO == 1) q[c] = max(q[c], lo); }
  return pack_sat4(q[0], q[1], q[2], q[3]);
}
// one stored row of the staged tile -> the three pixels x-1, x, x+1 of this thread's channel word; TMA zero-fills pixels
// outside the image, the layer needs REAL zero there (q = -zx)
__device__ __forceinline__ void hf_row(uint32_t rowp, int pitch, bool yok, const bool (&cok)[3], uint32_t pad, uint32_t (&T)[4]) {
  uint32_t w[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { const uint32_t v = lds_u32(rowp + (uint32_t)(j * pitch)); w[j] = (yok && cok[j]) ? v : pad; }
  transpose4x4(w[0], w[1], w[2], pad, T[0], T[1], T[2], T[3]);
}

template <int LO>
__global__ void __launch_bounds__(HF_MAX_THREADS, HF_CTAS) heads_fused_kernel(const __grid_constant__ CUtensorMap tmI, const HfParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t hf_smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array, so every pointer below stays in the shared address space
  // (through a uintptr_t round trip the compiler emitted generic LD.E / ST.E with 64-bit address arithmetic)
  uint8_t* smem = hf_smem_raw + ((1024u - (smem_u32(hf_smem_raw) & 1023u)) & 1023u);
  uint8_t* s_A = smem;                                       // 2 K blocks x [128 rows][128 B], 128B swizzle
  uint8_t* s_B = s_A + 2 * 16384;                            // 2 K blocks x [NB rows][128 B]
  uint8_t* s_in = s_B + 2 * (size_t)p.NB * 128;              // 2 x [HF_TH + 2][HF_TW + 2][pitch]: staged input tiles
  uint8_t* s_colp = s_in + 2 * (size_t)p.in_bytes;           // [NB] 32-byte records {double Mf, double bf, int acc_bias, -}
  const uint32_t s_col = smem_u32(s_colp);
  uint64_t* s_bar = (uint64_t*)(s_colp + (size_t)p.NB * 32); // [0] MMA done, [1..2] input buffer full
  volatile uint32_t* tmem_slot = (volatile uint32_t*)(s_bar + 3);
  const uint32_t bar = smem_u32(s_bar), ibar = bar + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool worker = tid < p.nthreads;
  const int cw = tid % p.cw_total, pc = tid / p.cw_total;

  auto tile_coords = [&](unsigned tile, int& tx, int& ty, int& b) {
    tx = (int)(tile % (unsigned)p.tiles_x); tile /= (unsigned)p.tiles_x;
    ty = (int)(tile % (unsigned)p.tiles_y); b = (int)(tile / (unsigned)p.tiles_y);
  };
  auto load_tile = [&](unsigned tile, int buf) {               // one thread: the tile's stored pixels + halo, zero-filled outside
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    mbar_expect_tx(ibar + 8u * buf, (uint32_t)p.in_bytes);
    tma_load_4d(smem_u32(s_in + (size_t)buf * p.in_bytes), &tmI, 0, tx * HF_TW - 1, ty * HF_TH - 1, b, ibar + 8u * buf);
  };
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(ibar, 1); mbar_init(ibar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmI) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // 1x1 weights -> shared memory in the UMMA K-major 128B-swizzled layout (16-byte units), epilogue constants
  {
    const int upr = p.Kp >> 4;                               // 16-byte units per weight row
    for (int i = tid; i < p.NB * upr; i += blockDim.x) {
      const int n = i / upr, c = i - n * upr, kb = c >> 3, cc = c & 7;
      const uint4 v = __ldg((const uint4*)(p.w + (size_t)n * p.Kp) + c);
      *(uint4*)(s_B + (size_t)kb * p.NB * 128 + n * 128 + ((cc ^ (n & 7)) << 4)) = v;
    }
    for (int i = tid; i < p.NB; i += blockDim.x) {
      *(double*)(s_colp + i * 32) = p.Mf[i]; *(double*)(s_colp + i * 32 + 8) = p.bf[i]; *(int*)(s_colp + i * 32 + 16) = p.acc_bias[i];
    }
  }
  // per-thread depthwise constants: 4 channels x 8 packed weight words, 4 RqInt records
  uint32_t W[4][8]; int2 km[4]; long long kb[4];   // RqInt as {Mi, sh} and a 64-bit Bi (loaded as one 64-bit value)
  if (worker) {
    const uint4* wv = (const uint4*)(p.wpk + (size_t)cw * 32);
    uint32_t flat[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const uint4 v = __ldg(wv + i); flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w; }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) W[c][i] = flat[c * 8 + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) { km[c] = __ldg((const int2*)(p.ki + cw * 4 + c)); kb[c] = __ldg((const long long*)(p.ki + cw * 4 + c) + 1); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // instruction descriptor: S32 accumulate, A/B signed int8, K-major both, N = NB, M = 128
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // byte offset of this thread's word inside an A row: K block, 16-byte unit (before the swizzle), byte in the unit
  const int kbyte = cw * 4;
  const uint32_t a_unit = (uint32_t)((kbyte & 127) >> 4);
  const uint32_t a_base = smem_u32(s_A) + (uint32_t)(kbyte >> 7) * 16384u + (uint32_t)(kbyte & 15);
  pdl_wait();                                // everything above is constant; activations need the previous grid
  if (tid == 0) {
    if (blockIdx.x < p.ntiles) load_tile(blockIdx.x, 0);
    if (blockIdx.x + gridDim.x < p.ntiles) load_tile(blockIdx.x + gridDim.x, 1);
  }
  uint32_t phase = 0, it = 0;
  const int row_bytes = (HF_TW + 2) * p.in_pitch;
  for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    const int buf = (int)(it & 1u);
    mbar_wait(ibar + 8u * buf, (it >> 1) & 1u);
    if (worker) {
      const int xg = tx * HF_TW + pc, y0 = ty * HF_TH;
      bool cok[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) cok[j] = (unsigned)(xg - 1 + j) < (unsigned)p.Ws;
      // staged row r holds stored row y0 - 1 + r; this thread reads pixels pc .. pc + 2 of it (stored columns xg-1 .. xg+1)
      uint32_t rowp = smem_u32(s_in) + (uint32_t)(buf * p.in_bytes + pc * p.in_pitch + kbyte);
      uint32_t Tm[4], Tc[4], Tp[4];
      hf_row(rowp, p.in_pitch, y0 >= 1, cok, p.pad_word, Tm); rowp += row_bytes;
      hf_row(rowp, p.in_pitch, true, cok, p.pad_word, Tc); rowp += row_bytes;
      for (int r = 0; r < HF_TH; ++r) {
        hf_row(rowp, p.in_pitch, y0 + r + 1 < p.Hs, cok, p.pad_word, Tp); rowp += row_bytes;
#pragma unroll
        for (int yp = 0; yp < 2; ++yp) {
          int a0[4], a1[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t ta = yp ? Tc[c] : Tm[c], tb = yp ? Tp[c] : Tc[c];
            a0[c] = dp4a_ss(tb, W[c][4 * yp + 2], dp4a_ss(ta, W[c][4 * yp + 0], 0));
            a1[c] = dp4a_ss(tb, W[c][4 * yp + 3], dp4a_ss(ta, W[c][4 * yp + 1], 0));
          }
          const uint32_t o0 = hf_rq_word<LO>(a0, km, kb, p.lo_i), o1 = hf_rq_word<LO>(a1, km, kb, p.lo_i);
          const uint32_t m0 = (uint32_t)((2 * r + yp) * 8 + 2 * pc);                            // A row of the left output pixel
          sts_u32(a_base + m0 * 128u + ((a_unit ^ (m0 & 7u)) << 4), o0);
          sts_u32(a_base + (m0 + 1u) * 128u + ((a_unit ^ ((m0 + 1u) & 7u)) << 4), o1);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { Tm[c] = Tc[c]; Tc[c] = Tp[c]; }
      }
    }
    fence_async_smem();                      // generic-proxy writes of the A tile (and, first time, B) -> async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      // this tile's input buffer has been consumed by every thread: refill it with the tile after the next
      if ((unsigned long long)tile + 2ull * gridDim.x < p.ntiles) load_tile(tile + 2u * gridDim.x, buf);
      tc_fence_after();
      for (int s = 0; s < p.ksteps; ++s) {
        const int kb = s >> 2, k = s & 3;
        const uint64_t adesc = make_smem_desc(smem_u32(s_A + (size_t)kb * 16384)) + (uint64_t)(2 * k);
        const uint64_t bdesc = make_smem_desc(smem_u32(s_B + (size_t)kb * p.NB * 128)) + (uint64_t)(2 * k);
        umma_i8(tmem_base, adesc, bdesc, idesc, s != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;       // every thread: the A tile may be overwritten, the accumulator is complete
    tc_fence_after();
    if (warp < 4) {
      const int m = tid;                                          // TMEM lane = tile row
      const int oy = m >> 3, ox = m & 7;
      float* o = p.out + (size_t)b * p.n_f32 * p.ppi + (size_t)(ty * (2 * HF_TH) + oy) * p.Wout + tx * (2 * HF_TW) + ox;
      const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int nf = p.n_f32, ppi = p.ppi;
      for (int c0 = 0; c0 < nf; c0 += 8) {
        uint32_t acc[16];
        tmem_ld8(tacc + (uint32_t)c0, acc);
        tmem_ld_wait();
        // branch-free: the constant records are padded to NB >= c0 + 8 columns, only the store is predicated
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 k = lds_u128(s_col + (uint32_t)(c0 + i) * 32u);                 // {Mf, bf}
          const int a = (int)acc[i] + (int)lds_u32(s_col + (uint32_t)(c0 + i) * 32u + 16u);
          const double yv = __dadd_rn(__dmul_rn((double)a, __hiloint2double((int)k.y, (int)k.x)), __hiloint2double((int)k.w, (int)k.z));
          if (c0 + i < nf) *o = (float)yv;
          o += ppi;
        }
      }
    }
    tc_fence_before();                       // the next tile's MMAs overwrite the accumulator after the next barrier
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
}

bool heads_fused_ok(const DwDevice& dw, const PwDevice& pw, int in_pitch, int mid_pitch, int Hs, int Ws) {
  return dw.use_int && dw.u_ok && dw.ki && dw.wpku && pw.n_f32 > 0 && pw.n_tiles == 1 && pw.N <= 128 && pw.k_off == 0 &&
         pw.K == mid_pitch && in_pitch == mid_pitch && dw.cw_total * 4 == mid_pitch && mid_pitch % 32 == 0 &&
         mid_pitch >= 128 && mid_pitch <= HF_MAX_THREADS && Hs % HF_TH == 0 && Ws % HF_TW == 0;
}

int heads_fused_launch(const DwDevice& dw, const PwDevice& pw, const int8_t* in, int in_pitch, int batch, int Hs, int Ws,
                       int zx, float* out_f32, cudaStream_t st) {
  CDN_CHECK(heads_fused_ok(dw, pw, in_pitch, pw.K, Hs, Ws), CDN_ERR_INVALID, "heads_fused: layer pair not eligible");
  HfParams p; memset(&p, 0, sizeof(p));
  p.in_pitch = in_pitch; p.in_bytes = (HF_TH + 2) * (HF_TW + 2) * in_pitch;
  p.Hs = Hs; p.Ws = Ws; p.tiles_x = Ws / HF_TW; p.tiles_y = Hs / HF_TH;
  const long long ntiles = (long long)batch * p.tiles_x * p.tiles_y;
  CDN_CHECK(ntiles < (1ll << 31) - 2 * 160 * HF_CTAS, CDN_ERR_INVALID, "heads_fused: tensor too large for 32-bit indexing");
  CDN_CHECK(p.in_bytes % 128 == 0, CDN_ERR_INVALID, "heads_fused: staged tile must be a multiple of 128 bytes");
  CUtensorMap tmI;
  if (int r = make_tmap_nhwc(&tmI, in, (uint64_t)in_pitch, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)batch, HF_TW + 2, HF_TH + 2)) return r;
  if (ntiles == 0) return 0;
  p.ntiles = (unsigned)ntiles;
  p.cw_total = dw.cw_total; p.nthreads = dw.cw_total * HF_TW;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx) * 0x01010101u;
  p.wpk = dw.wpku; p.ki = (const int4*)dw.ki; p.lo_i = dw.rq.lo;
  p.w = pw.w; p.Kp = pw.Kp; p.NB = pw.BN; p.ksteps = pw.K / 32;
  p.tmem_cols = 32; while (p.tmem_cols < p.NB) p.tmem_cols <<= 1;
  p.acc_bias = pw.rq.acc_bias; p.Mf = pw.Mf; p.bf = pw.bf; p.n_f32 = pw.n_f32;
  p.out = out_f32; p.ppi = 4 * Hs * Ws; p.Wout = 2 * Ws;
  const int threads = std::max(128, (p.nthreads + 31) / 32 * 32);
  const size_t smem = 1024 + 2 * 16384 + 2 * (size_t)p.NB * 128 + 2 * (size_t)p.in_bytes + (size_t)p.NB * 32 + 64;
  CDN_CHECK(smem <= HF_SMEM_MAX, CDN_ERR_INVALID, "heads_fused: %zu bytes of shared memory", smem);
  const bool lo_on = p.lo_i > -128;
  const int var = lo_on ? 1 : ((dw.sh0 && !(g_cdn_debug_flags & (1u << 22))) ? 2 : 0);
  auto kern = var == 1 ? heads_fused_kernel<1> : (var == 2 ? heads_fused_kernel<2> : heads_fused_kernel<0>);
  static bool attr_set[3][64] = {};
  if (cdn_first_on_device(attr_set[var])) {
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HF_SMEM_MAX));
    // the kernel lives in shared memory (input by TMA, A tile, weights) and has no use for L1: without this the driver's
    // carve-out left room for two CTAs per SM only (0.67 instead of 0.43 ms)
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  // persistent grid: the CTAs that are resident at this shared-memory size (227 KB per SM, 1 KB reserved per CTA): 3 per SM
  // for the 20-class heads, 2 for 80 classes.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor answered 2 for the 66 KB
  // configuration that ncu shows running 3 per SM, which cost a third of the throughput; registers allow 3 in every case.)
  const int per_sm = std::max(1, std::min<int>(HF_CTAS, (int)((227 * 1024) / (smem + 1024))));
  const unsigned blocks = (unsigned)std::min<long long>(ntiles, (long long)cdn_num_sms() * per_sm);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.stream = st; cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CDN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmI, p));
  CDN_LAUNCH_CHECK("heads_fused_kernel");
  return 0;
}
