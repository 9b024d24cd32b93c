// Depthwise 3x3, stride 1, with the input tile staged by TMA (int8 NHWC, integer requantisation).
//
// dw3x3_v2_kernel (dw.cu) loads every input row with LDG one iteration ahead and waits for L2 / DRAM on its first use
// (long-scoreboard stalls).  Here a persistent CTA of 128 threads walks over tiles of TW x TH output pixels
// (TW * pitch = 1024 bytes); the tile's input pixels plus a one-pixel halo arrive by ONE 4-D cp.async.bulk.tensor two tiles
// ahead into a double buffer, and the stencil -- the same window / transpose / dp4a / IMAD.HI code as dw3x3_v2_kernel<1,0,true>,
// so the same bits -- reads shared memory only.  TMA zero-fills pixels outside the image; the layer needs REAL zero there
// (q = -zx), substituted when the row is read.
#include "layers.cuh"
#include "tc_ptx.cuh"
#include <algorithm>

#define DT_THREADS 128
#define DT_CTAS 6

struct DwTParams {
  int pitch, in_bytes, buf_stride;           // input pixel pitch (bytes); bytes of one staged tile; in_bytes rounded up to 128
  int in_w;                                  // staged tile width in pixels
  int H, W, TW, TH, tiles_x, tiles_y;        // OUTPUT size and tile
  int Hin, Win;                              // input size
  unsigned ntiles;
  int cw_total;
  uint32_t pad_word;
  const uint32_t* wpk;                       // [channel][6] packed tap weights (left / right pixel per kernel row)
  const int4* ki; int lo_i;
  uint32_t* out; int out_pitch_w;
};

// NS: every channel requantises with shift 0 (DwDevice::sh0) -- no SHF after the IMAD.HI
template <int S, bool LO, bool NS>
__global__ void __launch_bounds__(DT_THREADS, DT_CTAS) dw3x3_tma_kernel(const __grid_constant__ CUtensorMap tmI, const DwTParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t dt_smem_raw[];
  uint8_t* smem = dt_smem_raw + ((128u - (smem_u32(dt_smem_raw) & 127u)) & 127u);
  const uint32_t s_in = smem_u32(smem);
  const uint32_t ibar = s_in + 2u * (uint32_t)p.buf_stride;
  const int tid = threadIdx.x;
  const int cw = tid % p.cw_total, pg = tid / p.cw_total;      // channel word, pixel pair of the tile row
  auto tile_coords = [&](unsigned tile, int& tx, int& ty, int& b) {
    tx = (int)(tile % (unsigned)p.tiles_x); tile /= (unsigned)p.tiles_x;
    ty = (int)(tile % (unsigned)p.tiles_y); b = (int)(tile / (unsigned)p.tiles_y);
  };
  auto load_tile = [&](unsigned tile, int buf) {
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    mbar_expect_tx(ibar + 8u * buf, (uint32_t)p.in_bytes);
    tma_load_4d(s_in + (uint32_t)(buf * p.buf_stride), &tmI, 0, S * tx * p.TW - 1, S * ty * p.TH - 1, b, ibar + 8u * buf);
  };
  if (tid == 0) {
    mbar_init(ibar, 1); mbar_init(ibar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmI) : "memory");
  }
  constexpr int NW = S == 2 ? 3 : 6;
  uint32_t Wt[4][NW]; int2 km[4]; long long kb[4];
  {
    const uint4* wv = (const uint4*)(p.wpk + (size_t)cw * 4 * NW);
    uint32_t flat[4 * NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) { const uint4 v = __ldg(wv + i); flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w; }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < NW; ++i) Wt[c][i] = flat[c * NW + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) { km[c] = __ldg((const int2*)(p.ki + cw * 4 + c)); kb[c] = __ldg((const long long*)(p.ki + cw * 4 + c) + 1); }
  }
  __syncthreads();
  pdl_wait();
  if (tid == 0) {
    if (blockIdx.x < p.ntiles) load_tile(blockIdx.x, 0);
    if (blockIdx.x + gridDim.x < p.ntiles) load_tile(blockIdx.x + gridDim.x, 1);
  }
  const int row_bytes = p.in_w * p.pitch;
  uint32_t it = 0;
  for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    int tx, ty, b; tile_coords(tile, tx, ty, b);
    const int buf = (int)(it & 1u);
    mbar_wait(ibar + 8u * buf, (it >> 1) & 1u);
    // stride 1: thread = pixel pair (2pg, 2pg+1) of the tile row, staged columns 2pg .. 2pg+3 = image columns xg-1 .. xg+2
    // stride 2: thread = output pixel pg, staged columns 2pg .. 2pg+2 = image columns 2xg-1 .. 2xg+1 (the 4th is unused)
    const int xl = S == 2 ? pg : 2 * pg, xg = tx * p.TW + xl, y0 = ty * p.TH;
    const int xi = S == 2 ? 2 * xg - 1 : xg - 1;                 // image column of the first staged pixel this thread reads
    bool cok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cok[j] = (S == 2 && j == 3) ? false : (unsigned)(xi + j) < (unsigned)p.Win;
    uint32_t rowp = s_in + (uint32_t)(buf * p.buf_stride + 2 * pg * p.pitch + cw * 4);
    auto read_row = [&](bool yok, uint32_t (&T)[4]) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (S == 2 && j == 3) { w[j] = p.pad_word; continue; }
        const uint32_t v = lds_u32(rowp + (uint32_t)(j * p.pitch)); w[j] = (yok && cok[j]) ? v : p.pad_word;
      }
      transpose4x4(w[0], w[1], w[2], w[3], T[0], T[1], T[2], T[3]);
      rowp += (uint32_t)row_bytes;
    };
    uint32_t Tm[4], Tc[4], Tp[4];
    uint32_t* o = p.out + ((size_t)((size_t)b * p.H + y0) * p.W + xg) * p.out_pitch_w + cw;
    if (S == 1) {
      read_row(y0 >= 1, Tm);
      read_row(true, Tc);
      for (int r = 0; r < p.TH; ++r) {
        read_row(y0 + r + 1 < p.Hin, Tp);
        int a0[4], a1[4], q0[4], q1[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a0[c] = dp4a_ss(Tp[c], Wt[c][4 % NW], dp4a_ss(Tc[c], Wt[c][2], dp4a_ss(Tm[c], Wt[c][0], 0)));
          a1[c] = dp4a_ss(Tp[c], Wt[c][5 % NW], dp4a_ss(Tc[c], Wt[c][3 % NW], dp4a_ss(Tm[c], Wt[c][1], 0)));
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          q0[c] = NS ? rq_int_hi0(a0[c], km[c].x, kb[c]) : rq_int_hi(a0[c], km[c].x, km[c].y, kb[c]);
          q1[c] = NS ? rq_int_hi0(a1[c], km[c].x, kb[c]) : rq_int_hi(a1[c], km[c].x, km[c].y, kb[c]);
          if (LO) { q0[c] = max(q0[c], p.lo_i); q1[c] = max(q1[c], p.lo_i); }
        }
        o[0] = pack_sat4(q0[0], q0[1], q0[2], q0[3]);
        o[p.out_pitch_w] = pack_sat4(q1[0], q1[1], q1[2], q1[3]);
        o += (size_t)p.W * p.out_pitch_w;
#pragma unroll
        for (int c = 0; c < 4; ++c) { Tm[c] = Tc[c]; Tc[c] = Tp[c]; }
      }
    } else {
      read_row(y0 >= 1, Tm);                                     // image row 2*y0 - 1
      for (int r = 0; r < p.TH; ++r) {
        read_row(true, Tc);                                      // 2(y0+r)
        read_row(2 * (y0 + r) + 1 < p.Hin, Tp);                  // 2(y0+r) + 1
        int q0[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int a = dp4a_ss(Tp[c], Wt[c][2], dp4a_ss(Tc[c], Wt[c][1], dp4a_ss(Tm[c], Wt[c][0], 0)));
          q0[c] = NS ? rq_int_hi0(a, km[c].x, kb[c]) : rq_int_hi(a, km[c].x, km[c].y, kb[c]);
          if (LO) q0[c] = max(q0[c], p.lo_i);
        }
        o[0] = pack_sat4(q0[0], q0[1], q0[2], q0[3]);
        o += (size_t)p.W * p.out_pitch_w;
#pragma unroll
        for (int c = 0; c < 4; ++c) Tm[c] = Tp[c];
      }
    }
    __syncthreads();                         // every thread is done with this buffer: refill it with the tile after the next
    if (tid == 0 && (unsigned long long)tile + 2ull * gridDim.x < p.ntiles) load_tile(tile + 2u * gridDim.x, buf);
  }
}

static int dw_tma_tw(int pitch, int stride) { return (stride == 2 ? 512 : 1024) / pitch; }   // 128 threads = channel words x pixels (pairs)

bool dw_tma_ok(const DwDevice& d, int in_pitch, int out_pitch, int H, int W, int in_shift, int stride) {
  if (!d.use_int || !d.ki || in_shift != 0 || (stride != 1 && stride != 2) || (g_cdn_debug_flags & 512u)) return false;   // bit 9: LDG kernels (A/B)
  if (in_pitch != 64 && in_pitch != 128 && in_pitch != 256 && !(stride == 2 && in_pitch == 32)) return false;
  if (d.cw_total * 4 != in_pitch || out_pitch < in_pitch) return false;
  if (stride == 2 && ((H | W) & 1)) return false;
  if ((stride == 2 && (g_cdn_debug_flags & 1024u))) return false;                                                        // bit 10: stride 2 on the LDG kernel
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  return Wo % dw_tma_tw(in_pitch, stride) == 0 && Ho % 8 == 0;
}

int dw_tma_launch(const DwDevice& d, const int8_t* in, int in_pitch, int8_t* out, int out_pitch, int batch, int H, int W,
                  int stride, int zx, cudaStream_t st) {
  DwTParams p; memset(&p, 0, sizeof(p));
  p.pitch = in_pitch; p.Hin = H; p.Win = W; p.H = (H - 1) / stride + 1; p.W = (W - 1) / stride + 1;
  p.TW = dw_tma_tw(in_pitch, stride); p.TH = 8;
  p.in_w = stride == 2 ? 2 * p.TW + 1 : p.TW + 2;
  const int in_h = stride == 2 ? 2 * p.TH + 1 : p.TH + 2;
  p.in_bytes = in_h * p.in_w * in_pitch; p.buf_stride = (p.in_bytes + 127) / 128 * 128;
  p.tiles_x = p.W / p.TW; p.tiles_y = p.H / p.TH;
  const long long ntiles = (long long)batch * p.tiles_x * p.tiles_y;
  if (ntiles == 0) return 0;
  CDN_CHECK(ntiles < (1ll << 31) - 2 * 148 * DT_CTAS, CDN_ERR_INVALID, "dw (TMA): tensor too large for 32-bit indexing");
  p.ntiles = (unsigned)ntiles;
  p.cw_total = d.cw_total;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx) * 0x01010101u;
  p.wpk = stride == 2 ? d.wpk2 : d.wpk1; p.ki = (const int4*)d.ki; p.lo_i = d.rq.lo;
  p.out = (uint32_t*)out; p.out_pitch_w = out_pitch / 4;
  CUtensorMap tmI;
  if (int r = make_tmap_nhwc(&tmI, in, (uint64_t)in_pitch, (uint64_t)W, (uint64_t)H, (uint64_t)batch, p.in_w, in_h)) return r;
  const size_t smem = 128 + 2 * (size_t)p.buf_stride + 32;
  CDN_CHECK(smem <= 64 * 1024, CDN_ERR_INVALID, "dw (TMA): %zu bytes of shared memory", smem);
  const bool lo_on = p.lo_i > -128;
  const bool ns = d.sh0 && !lo_on && !(g_cdn_debug_flags & (1u << 22));       // bit 22: keep the shift (A/B)
  void (*kern)(CUtensorMap, DwTParams) = stride == 2 ? (lo_on ? dw3x3_tma_kernel<2, true, false> : (ns ? dw3x3_tma_kernel<2, false, true> : dw3x3_tma_kernel<2, false, false>))
                                                     : (lo_on ? dw3x3_tma_kernel<1, true, false> : (ns ? dw3x3_tma_kernel<1, false, true> : dw3x3_tma_kernel<1, false, false>));
  static bool attr_set[6][64] = {};
  const int ai = (stride == 2 ? 3 : 0) + (lo_on ? 1 : (ns ? 2 : 0));
  if (cdn_first_on_device(attr_set[ai])) {
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  // persistent grid = the CTAs that are resident at this shared-memory size (227 KB per SM, 1 KB reserved per CTA; registers
  // allow DT_CTAS)
  const int per_sm = std::max(1, std::min<int>(DT_CTAS, (int)((227 * 1024) / (smem + 1024))));
  const unsigned blocks = (unsigned)std::min<long long>(ntiles, (long long)cdn_num_sms() * per_sm);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(DT_THREADS); cfg.stream = st; cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CDN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmI, p));
  CDN_LAUNCH_CHECK("dw3x3_tma_kernel");
  return 0;
}
