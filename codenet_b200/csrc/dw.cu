// Depthwise 3x3 and the fused co-designed deformable depthwise convolution (int8 NHWC, CUDA cores: these are
// gathers / stencils bound by HBM and instruction issue, not contractions -- see DESIGN.md).
//
// Thread mapping: one lane owns 4 consecutive channels (one 32-bit word of the NHWC pixel) and keeps their 9 tap
// weights, requantisation constants and scale-conv weights in registers; a warp covers 128 channels of one pixel
// (or 32/LPP pixels when the pixel is narrower than 128 bytes), so every tap is one fully coalesced 128-byte
// request.  The 3x3 MAC runs on dp4a: two 4x4 byte transposes turn (tap,channel) words into (channel,tap) words
// (8 PRMT on the ALU pipe) followed by 4 dp4a on the FMA pipe, so both issue pipes stay busy.
// Out-of-image taps contribute REAL zero, i.e. the grid value -zx (a = q + zx = 0); zx*sum(w) is added back
// exactly through acc_bias.
#include "layers.cuh"

struct DwParams {
  const uint32_t* in; uint32_t* out;
  int in_pitch_w, out_pitch_w;               // pitches in 32-bit words
  int Hs, Ws;                                // stored input size
  int Hin, Win;                              // logical input size (after virtual x2 upsample)
  int Hout, Wout;
  int shift, stride;
  int cw_total;                              // channel words = Cp/4 (of the narrower of in/out pitch)
  int G, lpp;                                // channel groups of 32 words; lanes per pixel
  long long total;                           // batch*Hout*Wout
  uint32_t pad_word;                         // (-zx) x4
  const uint32_t* wA; const uint32_t* wB; const uint32_t* wC;   // [cw_total*4] packed tap weights per channel
  const float* Mh; const float* Bh; const float* thr; const double* M; const double* B; const int32_t* acc_bias;
  float lo_f;
  // deformable part
  const uint32_t* ws;                        // [cw_total] packed scale-conv weights
  long long acc_s_bias;                      // zx * sum(ws)
  double Ms, bs, ss, zs, u_lo, u_hi;
  float* sval;
};

struct LaneConsts {
  uint32_t wA[4], wB[4], wC[4];
  float Mh[4], Bh[4], thr[4];
  int ab[4];
};

__device__ __forceinline__ void load_lane_consts(const DwParams& p, int cw, bool active, LaneConsts& k) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int ch = cw * 4 + c;
    k.wA[c] = active ? p.wA[ch] : 0u; k.wB[c] = active ? p.wB[ch] : 0u; k.wC[c] = active ? p.wC[ch] : 0u;
    k.Mh[c] = active ? p.Mh[ch] : 0.f; k.Bh[c] = active ? p.Bh[ch] : 0.f; k.thr[c] = active ? p.thr[ch] : 1.f;
    k.ab[c] = active ? p.acc_bias[ch] : 0;
  }
}

// 9 tap words (4 channels each) -> requantised word
__device__ __forceinline__ uint32_t mac9_requant(const uint32_t (&x)[9], const LaneConsts& k, const DwParams& p, int cw) {
  uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
  transpose4x4(x[0], x[1], x[2], x[3], a0, a1, a2, a3);
  transpose4x4(x[4], x[5], x[6], x[7], b0, b1, b2, b3);
  int acc0 = dp4a_ss(a0, k.wA[0], k.ab[0]); acc0 = dp4a_ss(b0, k.wB[0], acc0); acc0 = dp4a_ss(x[8], k.wC[0], acc0);
  int acc1 = dp4a_ss(a1, k.wA[1], k.ab[1]); acc1 = dp4a_ss(b1, k.wB[1], acc1); acc1 = dp4a_ss(x[8], k.wC[1], acc1);
  int acc2 = dp4a_ss(a2, k.wA[2], k.ab[2]); acc2 = dp4a_ss(b2, k.wB[2], acc2); acc2 = dp4a_ss(x[8], k.wC[2], acc2);
  int acc3 = dp4a_ss(a3, k.wA[3], k.ab[3]); acc3 = dp4a_ss(b3, k.wB[3], acc3); acc3 = dp4a_ss(x[8], k.wC[3], acc3);
  uint32_t r0 = requant_bits(acc0, k.Mh[0], k.Bh[0], k.thr[0], p.lo_f, p.M, p.B, cw * 4 + 0);
  uint32_t r1 = requant_bits(acc1, k.Mh[1], k.Bh[1], k.thr[1], p.lo_f, p.M, p.B, cw * 4 + 1);
  uint32_t r2 = requant_bits(acc2, k.Mh[2], k.Bh[2], k.thr[2], p.lo_f, p.M, p.B, cw * 4 + 2);
  uint32_t r3 = requant_bits(acc3, k.Mh[3], k.Bh[3], k.thr[3], p.lo_f, p.M, p.B, cw * 4 + 3);
  return pack4_lowbytes(r0, r1, r2, r3);
}

// ---------------------------------------------------------------------------------------------------------
// plain depthwise 3x3, stride 1/2, optional virtual x2 upsample of the input
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw3x3_kernel(DwParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int ppw = 32 / p.lpp, sub = lane / p.lpp, cl = lane % p.lpp;
  const long long pblocks = (p.total + ppw - 1) / ppw;
  const long long items = pblocks * p.G;       // (pixel block, channel group) pairs, group fastest
  LaneConsts k; int cur_g = -1;
  for (long long it = (long long)blockIdx.x * nw + warp; it < items; it += (long long)gridDim.x * nw) {
    const int g = (int)(it % p.G);
    const int cw = g * 32 + cl;
    const bool active = cw < p.cw_total;
    if (g != cur_g) { load_lane_consts(p, cw, active, k); cur_g = g; }
    long long pix = (it / p.G) * ppw + sub;
    if (pix >= p.total || !active) continue;
    int wo = (int)(pix % p.Wout); long long t = pix / p.Wout; int ho = (int)(t % p.Hout); long long b = t / p.Hout;
    const uint32_t* img = p.in + (size_t)b * p.Hs * p.Ws * p.in_pitch_w + cw;
    uint32_t x[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int y = ho * p.stride - 1 + i;
      bool yok = (unsigned)y < (unsigned)p.Hin;
      const uint32_t* row = img + (size_t)(y >> p.shift) * p.Ws * p.in_pitch_w;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int xx = wo * p.stride - 1 + j;
        bool ok = yok && (unsigned)xx < (unsigned)p.Win;
        x[i * 3 + j] = ok ? __ldg(row + (size_t)(xx >> p.shift) * p.in_pitch_w) : p.pad_word;
      }
    }
    p.out[(size_t)pix * p.out_pitch_w + cw] = mac9_requant(x, k, p, cw);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fused deformable depthwise conv.  MODE 0: integer offsets; MODE 1: bilinear (fp64, follows the oracle's
// operation order exactly: dcn_deform_conv_cuda_kernel.cu:83-114,210-227 restated on exact integers).
// ---------------------------------------------------------------------------------------------------------
template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT) deform_dw_kernel(DwParams p) {
  extern __shared__ int s_part[];            // [2][pixels_per_iter][G] partial scale dots (G > 1 only)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int g = warp % p.G, slot = warp / p.G, nslots = nw / p.G;
  const int ppw = 32 / p.lpp, sub = lane / p.lpp, cl = lane % p.lpp;
  const int cw = g * 32 + cl;
  const bool active = cw < p.cw_total;
  LaneConsts k; load_lane_consts(p, cw, active, k);
  const uint32_t wsw = active ? p.ws[cw] : 0u;
  const int per_iter = nslots * ppw;
  int parity = 0;
  for (long long base = (long long)blockIdx.x * per_iter; base < p.total; base += (long long)gridDim.x * per_iter) {
    const int pslot = slot * ppw + sub;
    long long pix = base + pslot;
    const bool pv = pix < p.total;
    long long pp = pv ? pix : p.total - 1;
    int w = (int)(pp % p.Wout); long long t = pp / p.Wout; int h = (int)(t % p.Hout); long long b = t / p.Hout;
    const uint32_t* img = p.in + (size_t)b * p.Hs * p.Ws * p.in_pitch_w + cw;
    uint32_t xc = active ? __ldg(img + ((size_t)(h >> p.shift) * p.Ws + (w >> p.shift)) * p.in_pitch_w) : 0u;
    int part = dp4a_ss(xc, wsw, 0);
    for (int o = p.lpp >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (p.G > 1) {
      int* sp = s_part + parity * per_iter * p.G;
      if (cl == 0) sp[pslot * p.G + g] = part;
      __syncthreads();
      part = 0;
      for (int gg = 0; gg < p.G; ++gg) part += sp[pslot * p.G + gg];
      parity ^= 1;                            // next iteration writes the other buffer: one barrier per iteration
    }
    // offset scalar (all lanes of the pixel compute it redundantly; fp64, mul/add kept separate as in the oracle)
    double u = __dadd_rn(__dmul_rn((double)((long long)part + p.acc_s_bias), p.Ms), p.bs);
    u = fmin(fmax(u, p.u_lo), p.u_hi);
    double qs = rint(__dsub_rn(__dmul_rn(p.ss, u), p.zs));
    double s = __ddiv_rn(__dadd_rn(qs, p.zs), p.ss);
    if (MODE == 0) s = rint(s);
    if (p.sval != nullptr && pv && g == 0 && cl == 0) p.sval[pix] = (float)s;
    if (!pv || !active) continue;
    if (MODE == 0) {
      const int si = (int)s;
      uint32_t x[9];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        int y = h + (i - 1) * si;
        bool yok = (unsigned)y < (unsigned)p.Hin;
        const uint32_t* row = img + (size_t)(y >> p.shift) * p.Ws * p.in_pitch_w;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          int xx = w + (j - 1) * si;
          bool ok = yok && (unsigned)xx < (unsigned)p.Win;
          x[i * 3 + j] = (i == 1 && j == 1) ? xc : (ok ? __ldg(row + (size_t)(xx >> p.shift) * p.in_pitch_w) : p.pad_word);
        }
      }
      p.out[(size_t)pix * p.out_pitch_w + cw] = mac9_requant(x, k, p, cw);
    } else {
      // bilinear: real values a = q + zx, zero outside the image
      const int zx = -(int)(int8_t)(p.pad_word & 0xff);
      const double d = __dsub_rn(s, 1.0);
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      // unpack per-channel tap weights from the packed registers
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double h_im = __dadd_rn((double)(h - 1 + i), (double)(i - 1) * d);
        double hl_d = floor(h_im);
        double lh = __dsub_rn(h_im, hl_d), hh = __dsub_rn(1.0, lh);
        int hl = (int)hl_d;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int tap = i * 3 + j;
          double w_im = __dadd_rn((double)(w - 1 + j), (double)(j - 1) * d);
          bool inside = h_im > -1.0 && w_im > -1.0 && h_im < (double)p.Hin && w_im < (double)p.Win;
          double wl_d = floor(w_im);
          double lw = __dsub_rn(w_im, wl_d), hw = __dsub_rn(1.0, lw);
          int wl = (int)wl_d;
          double bw[4] = {__dmul_rn(hh, hw), __dmul_rn(hh, lw), __dmul_rn(lh, hw), __dmul_rn(lh, lw)};
          double val[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int cnr = 0; cnr < 4; ++cnr) {
            int yy = hl + (cnr >> 1), xx = wl + (cnr & 1);
            bool ok = inside && yy >= 0 && yy <= p.Hin - 1 && xx >= 0 && xx <= p.Win - 1;
            uint32_t word = 0; 
            if (ok) word = __ldg(img + ((size_t)(yy >> p.shift) * p.Ws + (xx >> p.shift)) * p.in_pitch_w);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              double v = ok ? (double)((int)(int8_t)((word >> (8 * c)) & 0xff) + zx) : 0.0;
              double term = __dmul_rn(bw[cnr], v);
              val[c] = (cnr == 0) ? term : __dadd_rn(val[c], term);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t wword = tap < 4 ? k.wA[c] : (tap < 8 ? k.wB[c] : k.wC[c]);
            int sh = tap < 8 ? 8 * (tap & 3) : 8 * c;
            double wq = (double)(int)(int8_t)((wword >> sh) & 0xff);
            acc[c] = __dadd_rn(acc[c], __dmul_rn(wq, val[c]));
          }
        }
      }
      uint32_t r[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int ch = cw * 4 + c;
        double td = __dadd_rn(__dmul_rn(acc[c], __ldg(p.M + ch)), __ldg(p.B + ch));
        td = fmin(fmax(rint(td), (double)p.lo_f), 127.0);
        r[c] = (uint32_t)((int)td & 0xff);
      }
      p.out[(size_t)pix * p.out_pitch_w + cw] = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
int dw_device_build(DwDevice& d, const int8_t* wq, const int8_t* ws, int C, int Cp, int zx, const cdn_requant* rq) {
  CDN_CHECK(rq && rq->n == C && rq->M && rq->B, CDN_ERR_INVALID, "dw: requant constants must have n == C");
  CDN_CHECK(Cp % 32 == 0 && Cp >= C, CDN_ERR_INVALID, "dw: pitch %d must be a multiple of 32 and >= C=%d", Cp, C);
  CDN_CHECK(-zx >= -128 && -zx <= 127, CDN_ERR_INVALID,
            "dw: input zero point %d puts real zero outside the int8 grid (range does not contain 0)", zx);
  std::vector<uint32_t> wA(Cp, 0), wB(Cp, 0), wC(Cp, 0), wsw(Cp / 4, 0);
  std::vector<int32_t> ab(Cp, 0);
  for (int c = 0; c < C; ++c) {
    const int8_t* w = wq + c * 9;
    int sum = 0;
    for (int t = 0; t < 9; ++t) sum += w[t];
    for (int t = 0; t < 4; ++t) { wA[c] |= (uint32_t)(uint8_t)w[t] << (8 * t); wB[c] |= (uint32_t)(uint8_t)w[4 + t] << (8 * t); }
    wC[c] = (uint32_t)(uint8_t)w[8] << (8 * (c & 3));
    ab[c] = zx * sum;
  }
  d.acc_s_bias = 0;
  if (ws) for (int c = 0; c < C; ++c) { wsw[c / 4] |= (uint32_t)(uint8_t)ws[c] << (8 * (c & 3)); d.acc_s_bias += (long long)zx * ws[c]; }
  d.cw_total = Cp / 4;
  if (dev_upload(&d.wA, wA.data(), Cp)) return CDN_ERR_CUDA;
  if (dev_upload(&d.wB, wB.data(), Cp)) return CDN_ERR_CUDA;
  if (dev_upload(&d.wC, wC.data(), Cp)) return CDN_ERR_CUDA;
  if (dev_upload(&d.ws, wsw.data(), Cp / 4)) return CDN_ERR_CUDA;
  return dev_requant_upload(d.rq, rq, ab.data(), Cp);
}

void dw_device_free(DwDevice& d) {
  cudaFree(d.wA); cudaFree(d.wB); cudaFree(d.wC); cudaFree(d.ws); dev_requant_free(d.rq);
  d = DwDevice();
}

static void fill_common(DwParams& p, const DwDevice& d, const int8_t* in, int in_pitch, int8_t* out, int out_pitch,
                        int batch, int H, int W, int in_shift, int stride, int zx) {
  p.in = (const uint32_t*)in; p.out = (uint32_t*)out;
  p.in_pitch_w = in_pitch / 4; p.out_pitch_w = out_pitch / 4;
  p.Hin = H; p.Win = W; p.shift = in_shift; p.Hs = H >> in_shift; p.Ws = W >> in_shift; p.stride = stride;
  p.Hout = (H - 1) / stride + 1; p.Wout = (W - 1) / stride + 1;
  p.cw_total = d.cw_total;
  p.G = (d.cw_total + 31) / 32;
  int lpp = 32; while (lpp / 2 >= d.cw_total && lpp > 1) lpp /= 2;
  p.lpp = lpp;
  p.total = (long long)batch * p.Hout * p.Wout;
  uint32_t pb = (uint32_t)(uint8_t)(int8_t)(-zx);
  p.pad_word = pb * 0x01010101u;
  p.wA = d.wA; p.wB = d.wB; p.wC = d.wC; p.ws = d.ws;
  p.Mh = d.rq.Mh; p.Bh = d.rq.Bh; p.thr = d.rq.thr; p.M = d.rq.M; p.B = d.rq.B; p.acc_bias = d.rq.acc_bias;
  p.lo_f = (float)d.rq.lo;
  p.acc_s_bias = d.acc_s_bias;
  p.sval = nullptr;
}

int dw_launch(const DwDevice& d, const int8_t* in, int in_pitch, int8_t* out, int out_pitch, int batch, int H, int W,
              int in_shift, int stride, int zx, cudaStream_t st) {
  CDN_CHECK(stride == 1 || stride == 2, CDN_ERR_INVALID, "dw: stride must be 1 or 2");
  CDN_CHECK(in_shift == 0 || (in_shift == 1 && H % 2 == 0 && W % 2 == 0), CDN_ERR_INVALID, "dw: bad in_shift");
  CDN_CHECK(in_pitch >= d.cw_total * 4 && out_pitch >= d.cw_total * 4, CDN_ERR_INVALID, "dw: pitch smaller than channels");
  DwParams p; memset(&p, 0, sizeof(p));
  fill_common(p, d, in, in_pitch, out, out_pitch, batch, H, W, in_shift, stride, zx);
  if (p.total == 0) return 0;
  const int nw = 8;
  long long items = ((p.total + (32 / p.lpp) - 1) / (32 / p.lpp)) * p.G;
  long long blocks = (items + nw - 1) / nw;
  long long cap = (long long)cdn_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dw3x3_kernel<<<(unsigned)blocks, nw * 32, 0, st>>>(p);
  CDN_LAUNCH_CHECK("dw3x3_kernel");
  return 0;
}

int deform_launch(const DwDevice& d, const cdn_deform_scale* sc, const int8_t* in, int in_pitch, int8_t* out,
                  int out_pitch, int batch, int H, int W, int in_shift, int zx, float* sval, cudaStream_t st) {
  CDN_CHECK(sc && (sc->mode == 0 || sc->mode == 1), CDN_ERR_INVALID, "deform: mode must be 0 (round) or 1 (bilinear)");
  CDN_CHECK(sc->bound >= 1 && sc->bound <= 64, CDN_ERR_INVALID, "deform: offset bound %d out of range", sc->bound);
  CDN_CHECK(in_shift == 0 || (in_shift == 1 && H % 2 == 0 && W % 2 == 0), CDN_ERR_INVALID, "deform: bad in_shift");
  CDN_CHECK(in_pitch >= d.cw_total * 4 && out_pitch >= d.cw_total * 4, CDN_ERR_INVALID, "deform: pitch smaller than channels");
  DwParams p; memset(&p, 0, sizeof(p));
  fill_common(p, d, in, in_pitch, out, out_pitch, batch, H, W, in_shift, 1, zx);
  p.Ms = sc->Ms; p.bs = sc->bs; p.ss = sc->ss; p.zs = sc->zs;
  p.u_lo = (double)(-sc->bound + 1); p.u_hi = (double)sc->bound;
  p.sval = sval;
  if (p.total == 0) return 0;
  CDN_CHECK(p.G <= 32, CDN_ERR_INVALID, "deform: more than 4096 channels not supported");
  int nw = 8; if (p.G > 8) nw = p.G; else nw = (8 / p.G) * p.G;
  int per_iter = (nw / p.G) * (32 / p.lpp);
  long long blocks = (p.total + per_iter - 1) / per_iter;
  long long cap = (long long)cdn_num_sms() * (nw > 16 ? 2 : 8);
  if (blocks > cap) blocks = cap;
  size_t smem = p.G > 1 ? (size_t)2 * per_iter * p.G * sizeof(int) : 0;
  if (nw <= 8) {
    if (sc->mode == 0) deform_dw_kernel<0, 256><<<(unsigned)blocks, nw * 32, smem, st>>>(p);
    else deform_dw_kernel<1, 256><<<(unsigned)blocks, nw * 32, smem, st>>>(p);
  } else {
    if (sc->mode == 0) deform_dw_kernel<0, 1024><<<(unsigned)blocks, nw * 32, smem, st>>>(p);
    else deform_dw_kernel<1, 1024><<<(unsigned)blocks, nw * 32, smem, st>>>(p);
  }
  CDN_LAUNCH_CHECK("deform_dw_kernel");
  return 0;
}
