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
#include <algorithm>

#include "dw_params.cuh"

// ---------------------------------------------------------------------------------------------------------
// plain depthwise 3x3 (v2): sliding window down a column strip, horizontal taps packed for dp4a
//
// A thread owns one 32-bit channel word (4 channels) of a PAIR of horizontally adjacent output pixels and walks
// down R output rows.  Per input row it loads the 4 pixels x0-1 .. x0+2 (coalesced: lanes run along channels),
// transposes them once (8 PRMT) into per-channel words T[c] = (p(x0-1), p(x0), p(x0+1), p(x0+2)), and keeps the last
// three rows' T in registers.  An output is then 3 dp4a per channel: weights (w0,w1,w2,0) for the left pixel and
// (0,w0,w1,w2) for the right one -- 4.5 issue slots per output byte instead of 15 for tap-by-tap code.
//   STRIDE 2: one output pixel per thread, taps (2x-1, 2x, 2x+1), two new input rows per output row.
//   SHIFT 1 (virtual nearest x2 upsample of the input, stride 1): the 2x2 outputs that share a stored pixel
//   neighbourhood are produced together; rows/columns that read the same stored pixel twice have their weights
//   pre-added on the host, so an output costs 2 dp4a per channel.
// Accumulators start at acc_bias + MAGIC_I and are requantised with the lean guarded sequence of common.cuh.
// ---------------------------------------------------------------------------------------------------------
template <int STRIDE, int SHIFT> struct DwV2 { static constexpr int NW = STRIDE == 2 ? 3 : (SHIFT ? 8 : 6); };

struct DwV2Params {
  const uint32_t* in; uint32_t* out;
  int in_pitch_w, out_pitch_w;
  int Hs, Ws, Hout, Wout;                    // stored input size, output size
  int cw_total, PG, R, nstrips;              // channel words, pixel groups per row, rows per strip, strips per image
  long long nthreads;
  uint32_t pad_word;
  const uint32_t* wpk;                       // [channel][NW] packed tap weights
  const float2* mb; const int* abm;          // per channel (Mh, Bh), acc_bias + MAGIC_I
  const double* M; const double* B;
  float lo_f, thr;
  const int4* ki; int lo_i;                  // integer requantisation (RqInt per channel, acc_bias folded in)
};

struct DwLane { float Mh[4], Bh[4]; int abm[4]; };
// integer requantisation of 4 channels of one pixel (raw accumulators) -> packed word
template <bool LO>
__device__ __forceinline__ uint32_t dw_rq_word_int(const int (&acc)[4], const int4 (&r)[4], int lo) {
  int q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { q[c] = rq_int(acc[c], r[c]); if (LO) q[c] = max(q[c], lo); }
  return pack_sat4(q[0], q[1], q[2], q[3]);
}

// requantise 4 channels of one pixel: accumulators (already magic-biased) -> packed word (two packed f32x2 pairs)
__device__ __forceinline__ uint32_t dw_rq_word(const int (&acc)[4], const DwLane& k, float lo_f, RqGuard& g) {
  uint32_t r0, r1, r2, r3;
  rq_fast2(acc[0], acc[1], make_float2(k.Mh[0], k.Mh[1]), make_float2(k.Bh[0], k.Bh[1]), lo_f, g, r0, r1);
  rq_fast2(acc[2], acc[3], make_float2(k.Mh[2], k.Mh[3]), make_float2(k.Bh[2], k.Bh[3]), lo_f, g, r2, r3);
  return pack4_lowbytes(r0, r1, r2, r3);
}
__device__ __noinline__ uint32_t dw_rq_word_exact(int a0, int a1, int a2, int a3, const double* __restrict__ M,
                                                  const double* __restrict__ B, int ch, float lo_f) {
  const uint32_t r0 = rq_exact(a0 - CDN_MAGIC_I, __ldg(M + ch), __ldg(B + ch), lo_f);
  const uint32_t r1 = rq_exact(a1 - CDN_MAGIC_I, __ldg(M + ch + 1), __ldg(B + ch + 1), lo_f);
  const uint32_t r2 = rq_exact(a2 - CDN_MAGIC_I, __ldg(M + ch + 2), __ldg(B + ch + 2), lo_f);
  const uint32_t r3 = rq_exact(a3 - CDN_MAGIC_I, __ldg(M + ch + 3), __ldg(B + ch + 3), lo_f);
  return pack4_lowbytes(r0, r1, r2, r3);
}

// one stored row -> raw pixel words (4 channels each) of 4 (or 3) horizontally adjacent stored pixels; xo[j] = word
// offset of pixel j inside the row, or -1 when that pixel is outside the image (or unused).  Fetch and transpose are
// split so that the loads of the NEXT row are issued a whole loop iteration before they are consumed (these kernels
// were load-latency bound: 60% of the stall samples sat on the first PRMT after the row's LDGs).
__device__ __forceinline__ void dw_fetch_row(const uint32_t* __restrict__ row, bool yok, const int (&xo)[4], uint32_t pad,
                                             uint32_t (&w)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = (yok && xo[j] >= 0) ? __ldg(row + xo[j]) : pad;
}
__device__ __forceinline__ void dw_transpose_row(const uint32_t (&w)[4], uint32_t (&T)[4]) {
  transpose4x4(w[0], w[1], w[2], w[3], T[0], T[1], T[2], T[3]);
}

template <int STRIDE, int SHIFT, bool INT>
__global__ void __launch_bounds__(128) dw3x3_v2_kernel(const DwV2Params p) {
  pdl_launch_dependents();
  constexpr int NW = DwV2<STRIDE, SHIFT>::NW;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;        // host guarantees nthreads < 2^31
  if (idx >= (unsigned)p.nthreads) return;
  const int cw = (int)(idx % (unsigned)p.cw_total); unsigned t = idx / (unsigned)p.cw_total;
  const int pg = (int)(t % (unsigned)p.PG); t /= (unsigned)p.PG;
  const int strip = (int)(t % (unsigned)p.nstrips); const int b = (int)(t / (unsigned)p.nstrips);
  // per-lane constants: 4 channels x NW packed weight words are contiguous (16-byte aligned)
  uint32_t W[4][NW]; DwLane k; int4 ki[4];
  {
    const uint4* wv = (const uint4*)(p.wpk + (size_t)cw * 4 * NW);
    uint32_t flat[4 * NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) { const uint4 v = __ldg(wv + i); flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w; }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < NW; ++i) W[c][i] = flat[c * NW + i];
    if (INT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { ki[c] = __ldg(p.ki + cw * 4 + c); k.abm[c] = 0; }
    } else {
      const float4 m0 = __ldg((const float4*)(p.mb + cw * 4)), m1 = __ldg((const float4*)(p.mb + cw * 4) + 1);
      k.Mh[0] = m0.x; k.Bh[0] = m0.y; k.Mh[1] = m0.z; k.Bh[1] = m0.w; k.Mh[2] = m1.x; k.Bh[2] = m1.y; k.Mh[3] = m1.z; k.Bh[3] = m1.w;
      const int4 av = __ldg((const int4*)(p.abm + cw * 4));
      k.abm[0] = av.x; k.abm[1] = av.y; k.abm[2] = av.z; k.abm[3] = av.w;
    }
  }
  const bool lo_on = p.lo_i > -128;
  pdl_wait();                                  // constants above do not depend on the previous grid; activations do
  const int rs_in = p.Ws * p.in_pitch_w;                                 // words per stored input row
  const uint32_t* img = p.in + (size_t)b * p.Hs * rs_in + cw;
  uint32_t* outb = p.out + (size_t)b * p.Hout * p.Wout * p.out_pitch_w + cw;
  const int ch0 = cw * 4;
  const int xs0 = (STRIDE == 2 ? 2 * pg : (SHIFT ? pg : 2 * pg)) - 1;     // first stored pixel of the window
  int xo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xx = xs0 + j;
    xo[j] = ((STRIDE == 2 || SHIFT) && j == 3) || (unsigned)xx >= (unsigned)p.Ws ? -1 : xx * p.in_pitch_w;
  }

  if (STRIDE == 1) {
    // SHIFT 0: output pixels x0, x0+1 of rows y0..y1-1.  SHIFT 1: stored column pg, stored rows y0..y1-1 -> output
    // rows 2r, 2r+1 and columns 2pg, 2pg+1.  Both walk stored rows r with the window (r-1, r, r+1).
    const int rows = SHIFT ? p.Hs : p.Hout;
    const int y0 = strip * p.R, y1 = min(y0 + p.R, rows);
    uint32_t Tm[4], Tc[4], Tp[4], raw[4];
    const uint32_t* row = img + (y0 - 1) * rs_in;                          // may point before the image; guarded by yok
    dw_fetch_row(row, y0 >= 1, xo, p.pad_word, raw); dw_transpose_row(raw, Tm); row += rs_in;
    dw_fetch_row(row, true, xo, p.pad_word, raw); dw_transpose_row(raw, Tc); row += rs_in;
    dw_fetch_row(row, y0 + 1 < p.Hs, xo, p.pad_word, raw); row += rs_in;  // row y0+1, consumed in the first iteration
    const int x0 = 2 * pg;
    uint32_t* o = outb + ((SHIFT ? 2 * y0 : y0) * p.Wout + x0) * p.out_pitch_w;
    const bool two = x0 + 1 < p.Wout;
    for (int y = y0; y < y1; ++y) {
      dw_transpose_row(raw, Tp);
      dw_fetch_row(row, y + 2 < p.Hs && y + 1 < y1, xo, p.pad_word, raw); row += rs_in;   // prefetch for the next iteration
#pragma unroll
      for (int yp = 0; yp < (SHIFT ? 2 : 1); ++yp) {
        int a0[4], a1[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (SHIFT) {
            const uint32_t ta = yp ? Tc[c] : Tm[c], tb = yp ? Tp[c] : Tc[c];
            a0[c] = dp4a_ss(tb, W[c][(4 * yp + 2) % NW], dp4a_ss(ta, W[c][(4 * yp + 0) % NW], k.abm[c]));
            a1[c] = dp4a_ss(tb, W[c][(4 * yp + 3) % NW], dp4a_ss(ta, W[c][(4 * yp + 1) % NW], k.abm[c]));
          } else {
            a0[c] = dp4a_ss(Tp[c], W[c][4 % NW], dp4a_ss(Tc[c], W[c][2], dp4a_ss(Tm[c], W[c][0], k.abm[c])));
            a1[c] = dp4a_ss(Tp[c], W[c][5 % NW], dp4a_ss(Tc[c], W[c][3 % NW], dp4a_ss(Tm[c], W[c][1], k.abm[c])));
          }
        }
        uint32_t o0, o1;
        if (INT) {
          if (lo_on) { o0 = dw_rq_word_int<true>(a0, ki, p.lo_i); o1 = dw_rq_word_int<true>(a1, ki, p.lo_i); }
          else { o0 = dw_rq_word_int<false>(a0, ki, p.lo_i); o1 = dw_rq_word_int<false>(a1, ki, p.lo_i); }
        } else {
          RqGuard g; rq_guard_init(g);
          o0 = dw_rq_word(a0, k, p.lo_f, g); o1 = dw_rq_word(a1, k, p.lo_f, g);
          if (rq_group_bad(g, p.thr)) {
            o0 = dw_rq_word_exact(a0[0], a0[1], a0[2], a0[3], p.M, p.B, ch0, p.lo_f);
            o1 = dw_rq_word_exact(a1[0], a1[1], a1[2], a1[3], p.M, p.B, ch0, p.lo_f);
          }
        }
        o[0] = o0;
        if (SHIFT || two) o[p.out_pitch_w] = o1;
        o += p.Wout * p.out_pitch_w;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { Tm[c] = Tc[c]; Tc[c] = Tp[c]; }
    }
  } else {
    const int xo_ = pg, y0 = strip * p.R, y1 = min(y0 + p.R, p.Hout);
    uint32_t Tm[4], Tc[4], Tp[4], rawc[4], rawp[4];
    const uint32_t* row = img + (2 * y0 - 1) * rs_in;
    dw_fetch_row(row, y0 >= 1, xo, p.pad_word, rawc); dw_transpose_row(rawc, Tm); row += rs_in;
    dw_fetch_row(row, true, xo, p.pad_word, rawc); row += rs_in;
    dw_fetch_row(row, 2 * y0 + 1 < p.Hs, xo, p.pad_word, rawp); row += rs_in;
    uint32_t* o = outb + (y0 * p.Wout + xo_) * p.out_pitch_w;
    for (int y = y0; y < y1; ++y) {
      dw_transpose_row(rawc, Tc); dw_transpose_row(rawp, Tp);
      const bool more = y + 1 < y1;
      dw_fetch_row(row, more, xo, p.pad_word, rawc); row += rs_in;                            // rows 2(y+1), 2(y+1)+1
      dw_fetch_row(row, more && 2 * y + 3 < p.Hs, xo, p.pad_word, rawp); row += rs_in;
      int a0[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        a0[c] = dp4a_ss(Tp[c], W[c][2], dp4a_ss(Tc[c], W[c][1], dp4a_ss(Tm[c], W[c][0], k.abm[c])));
      uint32_t o0;
      if (INT) o0 = lo_on ? dw_rq_word_int<true>(a0, ki, p.lo_i) : dw_rq_word_int<false>(a0, ki, p.lo_i);
      else {
        RqGuard g; rq_guard_init(g);
        o0 = dw_rq_word(a0, k, p.lo_f, g);
        if (rq_group_bad(g, p.thr)) o0 = dw_rq_word_exact(a0[0], a0[1], a0[2], a0[3], p.M, p.B, ch0, p.lo_f);
      }
      o[0] = o0;
      o += p.Wout * p.out_pitch_w;
#pragma unroll
      for (int c = 0; c < 4; ++c) Tm[c] = Tp[c];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// fused co-designed deformable depthwise conv (v2).  One CTA works on tiles of DEF_NP consecutive output pixels:
//   phase A  offset scalar: every warp reduces the C->1 scale conv of its pixels (dp4a + shuffle tree), then ONE
//            lane per pixel runs the fp64 Hardtanh / QuantAct / (round) chain, so that scalar code is issued once
//            per 8 pixels instead of once per pixel and channel group; s goes to shared memory
//   phase C  gather + 3x3 depthwise MAC + requantisation: a warp owns a fixed group of 128 channels (constants in
//            registers) and walks over the tile's pixels.  MODE 0 (integer offsets): 9 coalesced 128-byte taps,
//            two byte transposes, 12 dp4a, lean guarded requantisation.  MODE 1 (bilinear): fp64, following the
//            oracle's operation order exactly (dcn_deform_conv_cuda_kernel.cu:83-114,210-227 on exact integers).
// ---------------------------------------------------------------------------------------------------------
#define DEF_NP 64
#define DEF_QCAP 1024

template <int MODE, bool INT>
__global__ void __launch_bounds__(256, MODE == 0 ? 4 : 3) deform_dw_v2_kernel(const DwParams p) {
  pdl_launch_dependents();
  __shared__ double s_s[DEF_NP];
  __shared__ uint32_t s_hw[DEF_NP];
  __shared__ int s_b[DEF_NP];                  // word offset of the pixel's image
  __shared__ int s_si[DEF_NP];                 // integer offset scalar (MODE 0)
  __shared__ uint32_t s_q[MODE == 1 ? DEF_QCAP : 1]; __shared__ int s_qn;   // MODE 1: words awaiting exact evaluation
  __shared__ int s_hl[MODE == 1 ? 2 * DEF_NP : 1], s_wl[MODE == 1 ? 2 * DEF_NP : 1];      // MODE 1: floor of the outer tap rows / columns
  __shared__ float s_lh[MODE == 1 ? 2 * DEF_NP : 1], s_lw[MODE == 1 ? 2 * DEF_NP : 1];   //         and their fractional parts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rs_in = p.Ws * p.in_pitch_w;
  // channel-group assignment of this warp in phase C
  const int Gw = p.G >= 8 ? 8 : (p.G >= 4 ? 4 : (p.G >= 2 ? 2 : 1));
  const int wg = warp % Gw, nslot = 8 / Gw;
  // narrow layers (fewer than 32 channel words): a warp covers ppw = 32 / lpp pixels at once instead of idling lanes
  const int lpp = p.lpp, ppw = 32 / lpp, sub = lane / lpp, cl = lane % lpp;
  const int wslot = (warp / Gw) * ppw + sub, jstep = nslot * ppw;
  const long long ntiles = (p.total + DEF_NP - 1) / DEF_NP;
  if (threadIdx.x == 0) s_qn = 0;              // published by the first tile's __syncthreads
  pdl_wait();
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * DEF_NP;
    // ---------------- phase A: s for the tile's pixels ----------------
    {
      int mine = 0; uint32_t my_hw = 0; int my_b = 0;
      // pixel coordinates advance incrementally (one division per warp and tile)
      const unsigned pix0 = (unsigned)base + warp;
      int w = (int)(pix0 % (unsigned)p.Wout); unsigned t0 = pix0 / (unsigned)p.Wout;
      int h = (int)(t0 % (unsigned)p.Hout); int b = (int)(t0 / (unsigned)p.Hout);
#pragma unroll 1
      for (int i = 0; i < DEF_NP / 8; ++i) {
        const long long pix = base + warp + 8 * i;
        int part = 0;
        if (pix < p.total) {
          const uint32_t* c = p.in + (size_t)b * p.Hs * rs_in + (h >> p.shift) * rs_in + (w >> p.shift) * p.in_pitch_w;
          int cw = lane;
          for (; cw + 96 < p.cw_total; cw += 128) {            // 4 independent loads in flight
            const uint32_t x0 = __ldg(c + cw), x1 = __ldg(c + cw + 32), x2 = __ldg(c + cw + 64), x3 = __ldg(c + cw + 96);
            part = dp4a_ss(x0, __ldg(p.ws + cw), part); part = dp4a_ss(x1, __ldg(p.ws + cw + 32), part);
            part = dp4a_ss(x2, __ldg(p.ws + cw + 64), part); part = dp4a_ss(x3, __ldg(p.ws + cw + 96), part);
          }
          for (; cw < p.cw_total; cw += 32) part = dp4a_ss(__ldg(c + cw), __ldg(p.ws + cw), part);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == i) { mine = part; my_hw = ((uint32_t)h << 16) | (uint32_t)w; my_b = b; }
        w += 8;
        while (w >= p.Wout) { w -= p.Wout; if (++h == p.Hout) { h = 0; ++b; } }
      }
      if (lane < DEF_NP / 8) {
        // fp64, mul/add kept separate as in the oracle
        double u = __dadd_rn(__dmul_rn((double)((long long)mine + p.acc_s_bias), p.Ms), p.bs);
        u = fmin(fmax(u, p.u_lo), p.u_hi);
        const double qs = rint(__dsub_rn(__dmul_rn(p.ss, u), p.zs));
        double s = __ddiv_rn(__dadd_rn(qs, p.zs), p.ss);
        if (MODE == 0) s = rint(s);
        s_s[warp + 8 * lane] = s;
        s_hw[warp + 8 * lane] = my_hw; s_b[warp + 8 * lane] = my_b * p.Hs * rs_in; s_si[warp + 8 * lane] = (int)s;
        if (MODE == 1) {
          // sample rows / columns of the outer taps (i, j = 0 and 2; the centre row / column is h, w exactly):
          // h_im = (h - 1 + i) + (i - 1)(s - 1) in fp64 as the oracle forms it, its floor and fractional part
          const int jp = warp + 8 * lane, hh_ = (int)(my_hw >> 16), ww_ = (int)(my_hw & 0xffffu);
          const double dd = __dsub_rn(s, 1.0);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double sg = e ? 1.0 : -1.0;
            const double him = __dadd_rn((double)(hh_ - 1 + 2 * e), sg * dd), wim = __dadd_rn((double)(ww_ - 1 + 2 * e), sg * dd);
            const double hf = floor(him), wf = floor(wim);
            s_hl[2 * jp + e] = (int)hf; s_lh[2 * jp + e] = (float)__dsub_rn(him, hf);
            s_wl[2 * jp + e] = (int)wf; s_lw[2 * jp + e] = (float)__dsub_rn(wim, wf);
          }
        }
        const long long pix = base + warp + 8 * lane;
        if (p.sval != nullptr && pix < p.total) p.sval[pix] = (float)s;
      }
    }
    __syncthreads();
    // ---------------- phase C: gather, MAC, requantise ----------------
    for (int g = wg; g < p.G; g += Gw) {
      const int cw = g * 32 + cl;
      const bool active = cw < p.cw_total;
      LaneConsts k; load_lane_consts(p, cw, active, k);
      float wf[MODE == 1 ? 4 : 1][MODE == 1 ? 9 : 1];                     // MODE 1: tap weights as floats
      if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t wword = tap < 4 ? k.wA[c] : (tap < 8 ? k.wB[c] : k.wC[c]);
            const int sh = tap < 8 ? 8 * (tap & 3) : 8 * c;
            wf[MODE == 1 ? c : 0][MODE == 1 ? tap : 0] = (float)(int)(int8_t)((wword >> sh) & 0xff);
          }
      }
      float Mh[4], Bh[4]; int abm[4]; int4 ki[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (INT) { ki[c] = active ? __ldg(p.ki + cw * 4 + c) : make_int4(0, 0, 0, 0); abm[c] = 0; }
        else {
          const float2 mb = active ? __ldg(p.mb + cw * 4 + c) : make_float2(0.f, 0.f);
          Mh[c] = mb.x; Bh[c] = mb.y; abm[c] = active ? __ldg(p.abm + cw * 4 + c) : CDN_MAGIC_I;
        }
      }
      if (MODE == 0) {
        // Software pipeline, unrolled by two with ping-pong tap registers: the 9 taps of the NEXT pixel are in flight
        // while this pixel is transposed, MAC-ed and requantised.  All addressing is unsigned 32-bit off one base
        // pointer (one IMAD.WIDE per load; the 64-bit version spent half of its issue slots on address arithmetic).
        const uint32_t* const base_in = p.in + cw;
        uint32_t* const base_out = p.out + (size_t)base * p.out_pitch_w + cw;
        const int nvalid = active ? (int)min((long long)DEF_NP, p.total - base) : 0;
        auto fetch = [&](int j, uint32_t (&x)[9]) {
          const uint32_t hw = s_hw[j];
          const int w = (int)(hw & 0xffffu), h = (int)(hw >> 16), si = s_si[j];
          const unsigned ob = (unsigned)s_b[j];
          unsigned ro[3], co[3]; bool yok[3], xok[3];
#pragma unroll
          for (int i = 0; i < 3; i += 2) {
            const int y = h + (i - 1) * si, xx = w + (i - 1) * si;
            yok[i] = (unsigned)y < (unsigned)p.Hin; xok[i] = (unsigned)xx < (unsigned)p.Win;
            ro[i] = ob + (unsigned)((min(max(y, 0), p.Hin - 1) >> p.shift) * rs_in);
            co[i] = (unsigned)((min(max(xx, 0), p.Win - 1) >> p.shift) * p.in_pitch_w);
          }
          yok[1] = xok[1] = true; ro[1] = ob + (unsigned)((h >> p.shift) * rs_in); co[1] = (unsigned)((w >> p.shift) * p.in_pitch_w);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) x[i * 3 + jj] = __ldg(base_in + (ro[i] + co[jj]));
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) if (!(yok[i] && xok[jj])) x[i * 3 + jj] = p.pad_word;
        };
        auto compute = [&](int j, const uint32_t (&x)[9]) {
          uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
          transpose4x4(x[0], x[1], x[2], x[3], a0, a1, a2, a3);
          transpose4x4(x[4], x[5], x[6], x[7], b0, b1, b2, b3);
          int acc[4];
          acc[0] = dp4a_ss(x[8], k.wC[0], dp4a_ss(b0, k.wB[0], dp4a_ss(a0, k.wA[0], abm[0])));
          acc[1] = dp4a_ss(x[8], k.wC[1], dp4a_ss(b1, k.wB[1], dp4a_ss(a1, k.wA[1], abm[1])));
          acc[2] = dp4a_ss(x[8], k.wC[2], dp4a_ss(b2, k.wB[2], dp4a_ss(a2, k.wA[2], abm[2])));
          acc[3] = dp4a_ss(x[8], k.wC[3], dp4a_ss(b3, k.wB[3], dp4a_ss(a3, k.wA[3], abm[3])));
          uint32_t o;
          if (INT) o = p.lo_i > -128 ? dw_rq_word_int<true>(acc, ki, p.lo_i) : dw_rq_word_int<false>(acc, ki, p.lo_i);
          else {
            RqGuard gd; rq_guard_init(gd);
            uint32_t r0, r1, r2, r3;
            rq_fast2(acc[0], acc[1], make_float2(Mh[0], Mh[1]), make_float2(Bh[0], Bh[1]), p.lo_f, gd, r0, r1);
            rq_fast2(acc[2], acc[3], make_float2(Mh[2], Mh[3]), make_float2(Bh[2], Bh[3]), p.lo_f, gd, r2, r3);
            o = pack4_lowbytes(r0, r1, r2, r3);
            if (rq_group_bad(gd, p.thr_layer)) o = dw_rq_word_exact(acc[0], acc[1], acc[2], acc[3], p.M, p.B, cw * 4, p.lo_f);
          }
          base_out[(unsigned)(j * p.out_pitch_w)] = o;
        };
        uint32_t xa[9], xb[9];
        int j = wslot;
        if (j < nvalid) fetch(j, xa);
#pragma unroll 1
        for (; j < nvalid; j += 2 * jstep) {
          const int j1 = j + jstep, j2 = j + 2 * jstep;
          if (j1 < nvalid) fetch(j1, xb);
          compute(j, xa);
          if (j2 < nvalid) fetch(j2, xa);
          if (j1 < nvalid) compute(j1, xb);
        }
      } else {
#pragma unroll 1
      for (int j = wslot; j < DEF_NP; j += jstep) {
        const long long pix = base + j;
        if (pix >= p.total || !active) continue;
        const uint32_t hw = s_hw[j];
        const int w = (int)(hw & 0xffffu), h = (int)(hw >> 16);
        const uint32_t* img = p.in + s_b[j] + cw;
        const double s = s_s[j];
        // ---- fp32 fast path: the 36 corners of the 9 taps lie on a 5 x 5 lattice (rows hl0, hl0+1, h, hl2, hl2+1 and the
        // same for columns); blend separably (columns, then the row weight), guard the rounding, fall back to fp64
        uint32_t oword; bool queued = false;
        {
          const int zx = -(int)(int8_t)(p.pad_word & 0xff);
          const float unb = 8388608.0f + 128.0f - (float)zx;                 // as_float(0x4B000000 | (q ^ 0x80)) - unb = q + zx
          const int hl0 = s_hl[2 * j], hl2 = s_hl[2 * j + 1], wl0 = s_wl[2 * j], wl2 = s_wl[2 * j + 1];
          const float lh0 = s_lh[2 * j], lh2 = s_lh[2 * j + 1], lw0 = s_lw[2 * j], lw2 = s_lw[2 * j + 1];
          const int R[5] = {hl0, hl0 + 1, h, hl2, hl2 + 1}, Cc[5] = {wl0, wl0 + 1, w, wl2, wl2 + 1};
          const float rho[5] = {1.0f - lh0, lh0, 1.0f, 1.0f - lh2, lh2};
          const float cwt[4] = {1.0f - lw0, lw0, 1.0f - lw2, lw2};
          unsigned co[5]; bool cv[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            cv[q] = (unsigned)Cc[q] < (unsigned)p.Win;
            co[q] = (unsigned)((min(max(Cc[q], 0), p.Win - 1) >> p.shift) * p.in_pitch_w);
          }
          // Two channels per FMA-pipe instruction (Blackwell's packed f32x2 add / mul / fma): per lane the operations and
          // their order are exactly those of the scalar form (mul, then fma ...), so the error bound above is unchanged.
          float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          const float2 nunb = make_float2(-unb, -unb);
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const bool rv = (unsigned)R[r] < (unsigned)p.Hin;
            const unsigned ro = (unsigned)((min(max(R[r], 0), p.Hin - 1) >> p.shift) * rs_in);
            uint32_t xw[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) xw[q] = __ldg(img + (ro + co[q]));
#pragma unroll
            for (int q = 0; q < 5; ++q) xw[q] = ((rv && cv[q]) ? xw[q] : p.pad_word) ^ 0x80808080u;
            const int ti = r < 2 ? 0 : (r == 2 ? 1 : 2);                     // tap row fed by this lattice row
            const float2 rho2 = make_float2(rho[r], rho[r]);
#pragma unroll
            for (int cp = 0; cp < 2; ++cp) {
              const int c = 2 * cp;
              float2 a[5];
#pragma unroll
              for (int q = 0; q < 5; ++q)
                a[q] = cdn_fadd2(make_float2(__uint_as_float(__byte_perm(xw[q], 0x4B000000u, 0x7650 + c)),
                                             __uint_as_float(__byte_perm(xw[q], 0x4B000000u, 0x7650 + c + 1))), nunb);
              const float2 u0 = cdn_ffma2(make_float2(cwt[1], cwt[1]), a[1], cdn_fmul2(make_float2(cwt[0], cwt[0]), a[0]));
              const float2 u2 = cdn_ffma2(make_float2(cwt[3], cwt[3]), a[4], cdn_fmul2(make_float2(cwt[2], cwt[2]), a[3]));
              const float2 w0 = make_float2(wf[c][ti * 3], wf[c + 1][ti * 3]), w1 = make_float2(wf[c][ti * 3 + 1], wf[c + 1][ti * 3 + 1]),
                           w2 = make_float2(wf[c][ti * 3 + 2], wf[c + 1][ti * 3 + 2]);
              const float2 t = cdn_ffma2(w2, u2, cdn_ffma2(w1, a[2], cdn_fmul2(w0, u0)));
              acc2[cp] = cdn_ffma2(rho2, t, acc2[cp]);
            }
          }
          const float acc[4] = {acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y};
          RqGuard gd; rq_guard_init(gd);
          uint32_t rr[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float t = fmaf(acc[c], Mh[c], Bh[c]);
            t = fmaxf(t, p.lo_f);
            gd.tmax = fmaxf(gd.tmax, t);
            const float r_ = __fadd_rn(t, CDN_MAGIC_F), kk = __fadd_rn(r_, -CDN_MAGIC_F);
            gd.d0 = fmaxf(gd.d0, fabsf(__fadd_rn(t, -kk)));
            rr[c] = __float_as_uint(r_);
          }
          oword = pack4_lowbytes(rr[0], rr[1], rr[2], rr[3]);
          if (rq_group_bad(gd, p.thr_bil)) {
            // Rare (~1 % of the words) but ~30x the work: doing it here would stall the whole warp for one lane, so
            // the word is queued and all 256 threads drain the queue densely at the end of the tile.
            const int pos = atomicAdd(&s_qn, 1);
            if (pos < DEF_QCAP) { s_q[pos] = ((uint32_t)j << 16) | (uint32_t)cw; queued = true; }
            else oword = deform_bilinear_exact_word(p, img, h, w, s, cw);
          }
        }
        if (!queued) p.out[(size_t)pix * p.out_pitch_w + cw] = oword;
      }
      }
    }
    if (MODE == 1) {                           // drain the exact-evaluation queue of this tile
      __syncthreads();
      const int nq = min(s_qn, DEF_QCAP);
      for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        const uint32_t e = s_q[i];
        const int j = (int)(e >> 16), cw = (int)(e & 0xffffu);
        const uint32_t hw = s_hw[j];
        p.out[(size_t)(base + j) * p.out_pitch_w + cw] =
            deform_bilinear_exact_word(p, p.in + s_b[j] + cw, (int)(hw >> 16), (int)(hw & 0xffffu), s_s[j], cw);
      }
    }
    __syncthreads();                         // s_s is rewritten by the next tile
    if (MODE == 1 && threadIdx.x == 0) s_qn = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// integer-offset layer, v3.  Everything that is the same for all channels of a pixel is done ONCE per pixel:
//   phase A  the C->1 scale conv is reduced per pixel by a warp (dp4a + shuffle tree, scale weights in registers); one
//            lane per pixel then maps the integer dot product to the integer offset scalar s by counting host-computed
//            thresholds (s is a monotone step function of the dot product: Hardtanh, QuantAct and both roundings are
//            monotone; the host evaluates the fp64 chain of dcn_deform_conv.py:295-330 / quant_modules.py:648-653 at
//            the step positions, so no fp64 runs on the device) and writes the 9 clamped tap offsets (32-bit word
//            offsets from the tensor base) and a 9-bit out-of-image mask to shared memory.
//   phase C  a warp owns 128 channels: per pixel 3 LDS.128 (tap offsets) + 1 LDS (mask), 9 coalesced 128-byte loads,
//            (border pixels only: 9 selects of the pad word), two byte transposes, 12 dp4a, requantisation, one store.
//            v2 recomputed the tap coordinates per channel group and pixel (~230 instructions per pixel and warp with
//            register spills); this loop is ~80.
// ---------------------------------------------------------------------------------------------------------
#ifndef DEF3_CTAS
#define DEF3_CTAS 3
#endif
#define DEF3_MAX_NP 256
// p.np = pixels per tile (64, 128 or 256; a multiple of 8): wide layers (many channel groups per pixel) take small tiles
// for load balance, narrow ones large tiles so that the per-tile work (constants, barriers, the per-pixel scalar code that
// runs on np/8 lanes of every warp) is amortised over 32 pixels per warp.
// RQ: 0 = guarded fp32 requantisation, 1 = integer, 2 = integer with an explicit lower clamp (lo > -128)
template <int RQ>
__global__ void __launch_bounds__(256, DEF3_CTAS) deform_int_v3_kernel(const DwParams p) {
  constexpr bool INT = RQ != 0;
  pdl_launch_dependents();
  // per pixel: the addresses of channel word 0 of the 9 taps (out-of-image taps point at the layer's pad pixel, which
  // holds q = -zx, i.e. real zero: no select and no mask in the gather loop), padded to 80 bytes
  __shared__ __align__(16) unsigned long long s_ptr[DEF3_MAX_NP][10];
  __shared__ int s_thr[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rs_in = p.Ws * p.in_pitch_w;
  const int Gw = p.G >= 8 ? 8 : (p.G >= 4 ? 4 : (p.G >= 2 ? 2 : 1));
  const int wg = warp % Gw, nslot = 8 / Gw;
  const int lpp = p.lpp, ppw = 32 / lpp, sub = lane / lpp, cl = lane % lpp;
  const int wslot = (warp / Gw) * ppw + sub, jstep = nslot * ppw;
  const int np = p.np;
  const long long ntiles = (p.total + np - 1) / np;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_thr[i] = i < p.s_n ? __ldg(p.s_thr + i) : 0x7fffffff;
  // phase C constants: a warp serves ONE channel group whenever G <= 8 (every shipped layer); they then stay in
  // registers for the whole kernel
  uint32_t wA[4], wB[4], wC[4];
  int Mi[4], sh[4]; long long Bi[4]; float Mh[4], Bh[4]; int abm[4];
  auto load_consts = [&](int cw, bool active) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ch = cw * 4 + c;
      wA[c] = active ? __ldg(p.wA + ch) : 0u; wB[c] = active ? __ldg(p.wB + ch) : 0u; wC[c] = active ? __ldg(p.wC + ch) : 0u;
      if (INT) {
        // Unconditional loads (inactive lanes read the last channel word's constants and never store): a select on these
        // values hides the sign extensions from the compiler, which then expands the 32 x 32 + 64 multiply-add of
        // rq_int into a generic 64 x 64-bit product (six instructions instead of one IMAD.HI).
        const int chv = min(cw, p.cw_total - 1) * 4 + c;
        const int2 ms = __ldg(reinterpret_cast<const int2*>(p.ki + chv));
        Mi[c] = ms.x; sh[c] = ms.y; Bi[c] = __ldg(reinterpret_cast<const long long*>(p.ki + chv) + 1); abm[c] = 0;
      } else {
        const float2 mb = active ? __ldg(p.mb + ch) : make_float2(0.f, 0.f);
        Mh[c] = mb.x; Bh[c] = mb.y; abm[c] = active ? __ldg(p.abm + ch) : CDN_MAGIC_I;
      }
    }
  };
  const bool single = p.G <= Gw;
  if (single) load_consts(wg * 32 + cl, wg * 32 + cl < p.cw_total);
  uint32_t so = smem_addr_u32(&s_ptr[0][0]);
  asm volatile("" : "+r"(so));                 // opaque: otherwise the shared-window base is rematerialised in every iteration
  pdl_wait();
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * np;
    __syncthreads();                           // previous tile's phase C is done with s_ptr (and s_thr is visible)
    // ---------------- phase A ----------------
    {
      // Lane i of a warp OWNS pixel j = warp + 8 i of the tile (coordinates, offset scalar, tap offsets).  The dot
      // products are formed 8 pixels at a time so that 8 independent loads are in flight per lane (the one-pixel-at-a-time
      // version spent a quarter of the kernel's stall samples on its first dp4a), and reduced with a value-halving
      // butterfly: 4 + 2 + 1 + 1 + 1 shuffles for 8 pixels instead of 8 x 5.
      const int per_warp = np >> 3;
      int my_h = 0, my_w = 0; unsigned my_ob = 0, coff = 0;
      if (lane < per_warp) {
        const long long pix = base + warp + 8 * lane;
        if (pix < p.total) {
          const unsigned pp = (unsigned)pix;
          my_w = (int)(pp % (unsigned)p.Wout); const unsigned t0 = pp / (unsigned)p.Wout;
          my_h = (int)(t0 % (unsigned)p.Hout);
          my_ob = (t0 / (unsigned)p.Hout) * (unsigned)(p.Hs * rs_in);
          coff = my_ob + (unsigned)((my_h >> p.shift) * rs_in + (my_w >> p.shift) * p.in_pitch_w);
        }
      }
      uint32_t wsr[8];                         // scale-conv weights of this lane's channel words (C <= 1024 in registers)
#pragma unroll
      for (int i = 0; i < 8; ++i) wsr[i] = lane + 32 * i < p.cw_total ? __ldg(p.ws + lane + 32 * i) : 0u;
      int mine = 0;
      const uint32_t* const in_l = p.in + lane;
#pragma unroll 1
      for (int bt = 0; bt < (per_warp >> 3); ++bt) {
        unsigned ob[8]; int part[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) { ob[b] = __shfl_sync(0xffffffffu, coff, 8 * bt + b); part[b] = 0; }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (32 * k < p.cw_total) {
            const bool on = lane + 32 * k < p.cw_total;
#pragma unroll
            for (int b = 0; b < 8; ++b) part[b] = dp4a_ss(on ? __ldg(word_ptr(in_l, ob[b] + 32u * k)) : 0u, wsr[k], part[b]);
          }
        for (int cw = lane + 256; cw < p.cw_total; cw += 32) {
          const uint32_t wv = __ldg(p.ws + cw);
#pragma unroll
          for (int b = 0; b < 8; ++b) part[b] = dp4a_ss(__ldg(word_ptr(p.in, ob[b] + (unsigned)cw)), wv, part[b]);
        }
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        int u[4], t[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = (b4 ? part[i + 4] : part[i]) + __shfl_xor_sync(0xffffffffu, b4 ? part[i] : part[i + 4], 16);
#pragma unroll
        for (int i = 0; i < 2; ++i) t[i] = (b3 ? u[i + 2] : u[i]) + __shfl_xor_sync(0xffffffffu, b3 ? u[i] : u[i + 2], 8);
        int r = (b2 ? t[1] : t[0]) + __shfl_xor_sync(0xffffffffu, b2 ? t[0] : t[1], 4);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        // lanes 4m .. 4m+3 now hold the dot product of the batch's pixel m, m = 4 b4 + 2 b3 + b2; hand it to its owner
        const int got = __shfl_sync(0xffffffffu, r, 4 * (lane & 7));
        if ((lane >> 3) == bt) mine = got;
      }
      if (lane < per_warp) {
        // s = s_lo + number of thresholds <= dot product (binary search over the ascending table, padded with INT_MAX)
        int cnt = 0;
#pragma unroll
        for (int step = 64; step > 0; step >>= 1) if (mine >= s_thr[cnt + step - 1]) cnt += step;
        const int si = p.s_lo + cnt;
        const int j = warp + 8 * lane;
        const unsigned ob = my_ob;
        unsigned ro[3], co[3]; unsigned ybad = 0, xbad = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int y = my_h + (i - 1) * si, xx = my_w + (i - 1) * si;
          if ((unsigned)y >= (unsigned)p.Hin) ybad |= 1u << i;
          if ((unsigned)xx >= (unsigned)p.Win) xbad |= 1u << i;
          ro[i] = ob + (unsigned)((min(max(y, 0), p.Hin - 1) >> p.shift) * rs_in);
          co[i] = (unsigned)((min(max(xx, 0), p.Win - 1) >> p.shift) * p.in_pitch_w);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            const bool bad = ((ybad >> i) | (xbad >> jj)) & 1u;
            s_ptr[j][i * 3 + jj] = bad ? (unsigned long long)p.pad_px : (unsigned long long)(p.in + (ro[i] + co[jj]));
          }
        const long long pix = base + j;
        if (p.sval != nullptr && pix < p.total) p.sval[pix] = (float)si;
      }
    }
    __syncthreads();
    // ---------------- phase C ----------------
    const int nvalid = (int)min((long long)np, p.total - base);
    for (int g = wg; g < p.G; g += Gw) {
      const int cw = g * 32 + cl;
      const bool active = cw < p.cw_total;
      if (!single) load_consts(cw, active);
      uint32_t* base_out = p.out + (size_t)base * p.out_pitch_w + cw;
      asm volatile("" : "+l"(base_out));       // keep it a 64-bit base: the store address is one IMAD.WIDE.U32 away
      const int nv = active ? nvalid : 0;
      const uint32_t cwu = (uint32_t)cw;
      auto fetch = [&](int j, uint32_t (&x)[9]) {
        const uint32_t sa = so + (uint32_t)j * 80u;
        const uint4 q0 = lds_u128(sa), q1 = lds_u128(sa + 16), q2 = lds_u128(sa + 32), q3 = lds_u128(sa + 48);
        const uint2 q4 = lds_u64(sa + 64);
        x[0] = __ldg(word_ptr64(q0.x, q0.y, cwu)); x[1] = __ldg(word_ptr64(q0.z, q0.w, cwu)); x[2] = __ldg(word_ptr64(q1.x, q1.y, cwu));
        x[3] = __ldg(word_ptr64(q1.z, q1.w, cwu)); x[4] = __ldg(word_ptr64(q2.x, q2.y, cwu)); x[5] = __ldg(word_ptr64(q2.z, q2.w, cwu));
        x[6] = __ldg(word_ptr64(q3.x, q3.y, cwu)); x[7] = __ldg(word_ptr64(q3.z, q3.w, cwu)); x[8] = __ldg(word_ptr64(q4.x, q4.y, cwu));
      };
      auto compute = [&](int j, const uint32_t (&x)[9]) {
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        transpose4x4(x[0], x[1], x[2], x[3], a0, a1, a2, a3);
        transpose4x4(x[4], x[5], x[6], x[7], b0, b1, b2, b3);
        int acc[4];
        acc[0] = dp4a_ss(x[8], wC[0], dp4a_ss(b0, wB[0], dp4a_ss(a0, wA[0], abm[0])));
        acc[1] = dp4a_ss(x[8], wC[1], dp4a_ss(b1, wB[1], dp4a_ss(a1, wA[1], abm[1])));
        acc[2] = dp4a_ss(x[8], wC[2], dp4a_ss(b2, wB[2], dp4a_ss(a2, wA[2], abm[2])));
        acc[3] = dp4a_ss(x[8], wC[3], dp4a_ss(b3, wB[3], dp4a_ss(a3, wA[3], abm[3])));
        uint32_t o;
        if (INT) {
          int q[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) { q[c] = rq_int(acc[c], Mi[c], sh[c], Bi[c]); if (RQ == 2) q[c] = max(q[c], p.lo_i); }
          o = pack_sat4(q[0], q[1], q[2], q[3]);
        } else {
          RqGuard gd; rq_guard_init(gd);
          uint32_t r0, r1, r2, r3;
          rq_fast2(acc[0], acc[1], make_float2(Mh[0], Mh[1]), make_float2(Bh[0], Bh[1]), p.lo_f, gd, r0, r1);
          rq_fast2(acc[2], acc[3], make_float2(Mh[2], Mh[3]), make_float2(Bh[2], Bh[3]), p.lo_f, gd, r2, r3);
          o = pack4_lowbytes(r0, r1, r2, r3);
          if (rq_group_bad(gd, p.thr_layer)) o = dw_rq_word_exact(acc[0], acc[1], acc[2], acc[3], p.M, p.B, cw * 4, p.lo_f);
        }
        *word_ptr(base_out, (uint32_t)j * (uint32_t)p.out_pitch_w) = o;
      };
      uint32_t xa[9], xb[9];
      int j = wslot;
      if (j < nv) fetch(j, xa);
#pragma unroll 1
      for (; j < nv; j += 2 * jstep) {
        const int j1 = j + jstep, j2 = j + 2 * jstep;
        if (j1 < nv) fetch(j1, xb);
        compute(j, xa);
        if (j2 < nv) fetch(j2, xa);
        if (j1 < nv) compute(j1, xb);
      }
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
  // v2 kernels: tap weights packed for horizontal dp4a (see dw3x3_v2_kernel), per-channel fast constants
  {
    auto pk = [](int b0, int b1, int b2, int b3) {
      return (uint32_t)(uint8_t)(int8_t)b0 | ((uint32_t)(uint8_t)(int8_t)b1 << 8) | ((uint32_t)(uint8_t)(int8_t)b2 << 16) |
             ((uint32_t)(uint8_t)(int8_t)b3 << 24);
    };
    std::vector<uint32_t> w1((size_t)Cp * 6, 0), w2((size_t)Cp * 3, 0), wu((size_t)Cp * 8, 0);
    std::vector<float2> mb(Cp, make_float2(0.f, 0.f));
    std::vector<int32_t> abm(Cp, CDN_MAGIC_I_HOST);
    double min_thr = 0.5; bool u_ok = true;
    for (int c = 0; c < C; ++c) {
      const int8_t* w = wq + c * 9;
      for (int r = 0; r < 3; ++r) {
        w1[(size_t)c * 6 + 2 * r] = pk(w[3 * r], w[3 * r + 1], w[3 * r + 2], 0);
        w1[(size_t)c * 6 + 2 * r + 1] = pk(0, w[3 * r], w[3 * r + 1], w[3 * r + 2]);
        w2[(size_t)c * 3 + r] = pk(w[3 * r], w[3 * r + 1], w[3 * r + 2], 0);
      }
      // upsample-folded: [ypar][term] row vectors, each -> px0 (v0, v1+v2, 0, 0) and px1 (0, v0+v1, v2, 0)
      int rows[2][2][3];
      for (int j = 0; j < 3; ++j) {
        rows[0][0][j] = w[j];            rows[0][1][j] = w[3 + j] + w[6 + j];
        rows[1][0][j] = w[j] + w[3 + j]; rows[1][1][j] = w[6 + j];
      }
      for (int yp = 0; yp < 2; ++yp)
        for (int tm = 0; tm < 2; ++tm) {
          const int* v = rows[yp][tm];
          const int a[2][3] = {{v[0], v[1] + v[2], 0}, {0, v[0] + v[1], v[2]}};
          for (int px = 0; px < 2; ++px) {
            for (int j = 0; j < 3; ++j) if (a[px][j] < -128 || a[px][j] > 127) u_ok = false;
            wu[(size_t)c * 8 + 4 * yp + 2 * tm + px] = pk(a[px][0], a[px][1], a[px][2], 0);
          }
        }
      RqFast f = rq_fast_from(rq->M[c], rq->B[c]);
      mb[c] = make_float2(f.Mh, f.Bh);
      abm[c] = ab[c] + CDN_MAGIC_I_HOST;
      min_thr = std::min(min_thr, (double)f.thr);
    }
    d.thr = (float)min_thr; d.u_ok = u_ok ? 1 : 0;
    // integer requantisation (common.cuh, RqInt): exact over |v| <= A * sum|w|, A = max |q + zx|; acc_bias folded in so the
    // dp4a chains start at zero.  The upsample-folded kernel pre-adds weights, which can only shrink the reachable range.
    d.use_int = 0;
    if (!(g_cdn_debug_flags & 128u)) {
      const long long A = std::max(std::abs((long long)zx - 128), std::abs((long long)zx + 127));
      std::vector<RqInt> ki(Cp, RqInt{0, 0, 0});
      bool ok = true;
      for (int c = 0; c < C && ok; ++c) {
        long long asum = 0;
        for (int t = 0; t < 9; ++t) { const int v = wq[c * 9 + t]; asum += v < 0 ? -v : v; }
        ok = rq_int_solve(rq->M[c], rq->B[c], std::max(-128, std::min(127, rq->lo)), -A * asum, A * asum, &ki[c]) &&
             rq_int_rebase(&ki[c], ab[c], 128 * asum);
      }
      if (ok) {
        if (dev_upload((RqInt**)&d.ki, ki.data(), ki.size())) return CDN_ERR_CUDA;
        d.use_int = 1; d.sh0 = 1;
        for (const RqInt& k : ki) if (k.sh != 0) d.sh0 = 0;
      }
    }
    // fp32 bilinear fast path (deform MODE 1), u = 2^-24, A = 255 >= |q + zx|, S = sum_ij |w_ij| of the channel:
    //   column blend u0 = fl(c1*a1 + fl(c0*a0)), c0 + c1 = 1, weights off by <= u each:        |err| <= 4 A u
    //   t_r = 3 fma over the tap columns:                     |err| <= S_i A (4 + 3) u  (S_i = sum_j |w_ij|)
    //   acc += rho_r * t_r, rho off by <= 2u, the rho of a tap row sum to 1, 5 fma roundings:  |err| <= S A (7 + 2 + 5) u
    // so |acc_fp32 - acc| <= 14 S A u; 16 is used.  Mapped through M and added to the requantisation bound.
    double max_eps = 0.0;
    for (int c = 0; c < C; ++c) {
      double asum = 0; for (int t = 0; t < 9; ++t) asum += fabs((double)wq[c * 9 + t]);
      const double eps_acc = asum * 255.0 * 16.0 * ldexp(1.0, -24);
      RqFast f = rq_fast_from(rq->M[c], rq->B[c]);
      max_eps = std::max(max_eps, eps_acc * fabs(rq->M[c]) + (0.5 - (double)f.thr));
    }
    d.thr_bil = (float)(0.5 - std::min(max_eps, 0.49));
    if (dev_upload(&d.wpk1, w1.data(), w1.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.wpk2, w2.data(), w2.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.wpku, wu.data(), wu.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.mb, mb.data(), mb.size())) return CDN_ERR_CUDA;
    if (dev_upload(&d.abm, abm.data(), abm.size())) return CDN_ERR_CUDA;
  }
  if (dev_upload(&d.wA, wA.data(), Cp)) return CDN_ERR_CUDA;
  if (dev_upload(&d.wB, wB.data(), Cp)) return CDN_ERR_CUDA;
  if (dev_upload(&d.wC, wC.data(), Cp)) return CDN_ERR_CUDA;
  if (dev_upload(&d.ws, wsw.data(), Cp / 4)) return CDN_ERR_CUDA;
  return dev_requant_upload(d.rq, rq, ab.data(), Cp);
}

// Integer-offset mode: the offset scalar as a step function of the scale conv's integer dot product.
//   v = sum_c ws[c]*(q_c + zx);  u = clamp(fl(fl(v*Ms) + bs), -bound+1, bound);  qs = rint(fl(fl(ss*u) - zs));
//   s = rint(fl(fl(qs + zs) / ss))          (dcn_deform_conv.py:295-330, quant_modules.py:648-653, DeformConvWithOffsetRound)
// Every step is monotone non-decreasing in v (Ms, ss > 0), so s(v) = s_lo + #{k : v >= T_k}; T_k is found by bisection
// with exactly these fp64 operations.  Stored relative to the raw dot product sum_c ws[c]*q_c the kernel accumulates.
static inline double deform_s_of_v(long long v, const cdn_deform_scale* sc) {
  volatile double a = (double)v * sc->Ms; volatile double u = a + sc->bs;
  double uc = fmin(fmax((double)u, (double)(-sc->bound + 1)), (double)sc->bound);
  volatile double m = sc->ss * uc; volatile double d = m - sc->zs;
  const double qs = nearbyint((double)d);
  volatile double n = qs + sc->zs; volatile double s = n / sc->ss;
  return nearbyint((double)s);
}
int deform_scale_build(DwDevice& d, const cdn_deform_scale* sc, const int8_t* ws, int C, int zx) {
  d.s_mode0_ok = 0;
  if (sc->mode != 0) return 0;
  CDN_CHECK(sc->Ms > 0.0 && sc->ss > 0.0 && std::isfinite(sc->bs) && std::isfinite(sc->zs), CDN_ERR_INVALID,
            "deform: scale quantiser must have positive scales (Ms=%g ss=%g)", sc->Ms, sc->ss);
  long long asum = 0, sum = 0;
  for (int c = 0; c < C; ++c) { asum += ws[c] < 0 ? -ws[c] : ws[c]; sum += ws[c]; }
  const long long A = std::max(std::abs((long long)zx - 128), std::abs((long long)zx + 127));
  const long long vmin = -A * asum, vmax = A * asum, bias = (long long)zx * sum;
  CDN_CHECK(vmax - bias < (1ll << 31) - 2 && vmin - bias > -(1ll << 31) + 2, CDN_ERR_INVALID, "deform: scale conv accumulator exceeds 32 bits");
  const int s_min = (int)deform_s_of_v(vmin, sc), s_max = (int)deform_s_of_v(vmax, sc);
  CDN_CHECK(s_max - s_min <= 127 && s_min >= -sc->bound - 1 && s_max <= sc->bound + 1, CDN_ERR_INVALID,
            "deform: offset scalar range [%d, %d] inconsistent with bound %d", s_min, s_max, sc->bound);
  std::vector<int32_t> thr;
  for (int L = s_min + 1; L <= s_max; ++L) {               // smallest v with s(v) >= L
    long long lo_v = vmin, hi_v = vmax;                    // s(lo_v) < L <= s(hi_v)
    while (hi_v - lo_v > 1) {
      const long long mid = lo_v + (hi_v - lo_v) / 2;
      if (deform_s_of_v(mid, sc) >= (double)L) hi_v = mid; else lo_v = mid;
    }
    thr.push_back((int32_t)(hi_v - bias));
  }
  if (d.s_thr) { cudaFree(d.s_thr); d.s_thr = nullptr; }
  if (d.pad_px) { cudaFree(d.pad_px); d.pad_px = nullptr; }
  {                                                        // the pad pixel out-of-image taps are redirected to
    std::vector<uint32_t> pad((size_t)d.cw_total, (uint32_t)(uint8_t)(int8_t)(-zx) * 0x01010101u);
    if (dev_upload(&d.pad_px, pad.data(), pad.size())) return CDN_ERR_CUDA;
  }
  if (dev_upload(&d.s_thr, thr.data(), thr.size())) return CDN_ERR_CUDA;
  d.s_n = (int)thr.size(); d.s_lo = s_min; d.s_mode0_ok = 1;
  return 0;
}

void dw_device_free(DwDevice& d) {
  cudaFree(d.wA); cudaFree(d.wB); cudaFree(d.wC); cudaFree(d.ws); dev_requant_free(d.rq);
  cudaFree(d.wpk1); cudaFree(d.wpk2); cudaFree(d.wpku); cudaFree(d.mb); cudaFree(d.abm); cudaFree(d.ki); cudaFree(d.s_thr); cudaFree(d.pad_px);
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
  p.mb = d.mb; p.abm = d.abm; p.thr_layer = d.thr; p.thr_bil = d.thr_bil;
  p.ki = (const int4*)d.ki; p.lo_i = d.rq.lo;
}

int dw_launch(const DwDevice& d, const int8_t* in, int in_pitch, int8_t* out, int out_pitch, int batch, int H, int W,
              int in_shift, int stride, int zx, cudaStream_t st) {
  CDN_CHECK(stride == 1 || stride == 2, CDN_ERR_INVALID, "dw: stride must be 1 or 2");
  CDN_CHECK(in_shift == 0 || (in_shift == 1 && stride == 1 && H % 2 == 0 && W % 2 == 0), CDN_ERR_INVALID, "dw: bad in_shift");
  CDN_CHECK(in_pitch >= d.cw_total * 4 && out_pitch >= d.cw_total * 4, CDN_ERR_INVALID, "dw: pitch smaller than channels");
  CDN_CHECK(!in_shift || d.u_ok, CDN_ERR_INVALID, "dw: weights too large for the upsample-folded kernel (sums exceed int8)");
  if (dw_tma_ok(d, in_pitch, out_pitch, H, W, in_shift, stride)) return dw_tma_launch(d, in, in_pitch, out, out_pitch, batch, H, W, stride, zx, st);
  DwV2Params p; memset(&p, 0, sizeof(p));
  p.in = (const uint32_t*)in; p.out = (uint32_t*)out;
  p.in_pitch_w = in_pitch / 4; p.out_pitch_w = out_pitch / 4;
  p.Hs = H >> in_shift; p.Ws = W >> in_shift;
  p.Hout = (H - 1) / stride + 1; p.Wout = (W - 1) / stride + 1;
  p.cw_total = d.cw_total;
  const int rows = in_shift ? p.Hs : p.Hout;            // rows a strip walks over
  p.R = stride == 2 ? 8 : 16;
  if (rows < p.R) p.R = rows;
  p.nstrips = (rows + p.R - 1) / p.R;
  p.PG = in_shift ? p.Ws : (stride == 2 ? p.Wout : (p.Wout + 1) / 2);
  p.nthreads = (long long)batch * p.nstrips * p.PG * p.cw_total;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx) * 0x01010101u;
  p.wpk = in_shift ? d.wpku : (stride == 2 ? d.wpk2 : d.wpk1);
  p.mb = d.mb; p.abm = d.abm; p.M = d.rq.M; p.B = d.rq.B;
  p.lo_f = (float)d.rq.lo; p.thr = d.thr;
  p.ki = (const int4*)d.ki; p.lo_i = d.rq.lo;
  if (p.nthreads == 0) return 0;
  CDN_CHECK(p.nthreads < (1ll << 31) && (long long)p.Hs * p.Ws * p.in_pitch_w < (1ll << 31) &&
            (long long)p.Hout * p.Wout * p.out_pitch_w < (1ll << 31), CDN_ERR_INVALID, "dw: tensor too large for 32-bit indexing");
  const unsigned blocks = (unsigned)((p.nthreads + 127) / 128);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(128); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (d.use_int) {
    if (in_shift) CDN_CUDA(cudaLaunchKernelEx(&cfg, dw3x3_v2_kernel<1, 1, true>, p));
    else if (stride == 2) CDN_CUDA(cudaLaunchKernelEx(&cfg, dw3x3_v2_kernel<2, 0, true>, p));
    else CDN_CUDA(cudaLaunchKernelEx(&cfg, dw3x3_v2_kernel<1, 0, true>, p));
  } else {
    if (in_shift) CDN_CUDA(cudaLaunchKernelEx(&cfg, dw3x3_v2_kernel<1, 1, false>, p));
    else if (stride == 2) CDN_CUDA(cudaLaunchKernelEx(&cfg, dw3x3_v2_kernel<2, 0, false>, p));
    else CDN_CUDA(cudaLaunchKernelEx(&cfg, dw3x3_v2_kernel<1, 0, false>, p));
  }
  CDN_LAUNCH_CHECK("dw3x3_v2_kernel");
  return 0;
}

int deform_launch(const DwDevice& d, const cdn_deform_scale* sc, const int8_t* in, int in_pitch, int8_t* out,
                  int out_pitch, int batch, int H, int W, int in_shift, int zx, float* sval, cudaStream_t st) {
  CDN_CHECK(sc && (sc->mode == 0 || sc->mode == 1), CDN_ERR_INVALID, "deform: mode must be 0 (round) or 1 (bilinear)");
  CDN_CHECK(sc->bound >= 1 && sc->bound <= 64, CDN_ERR_INVALID, "deform: offset bound %d out of range", sc->bound);
  CDN_CHECK(in_shift == 0 || (in_shift == 1 && H % 2 == 0 && W % 2 == 0), CDN_ERR_INVALID, "deform: bad in_shift");
  CDN_CHECK(in_pitch >= d.cw_total * 4 && out_pitch >= d.cw_total * 4, CDN_ERR_INVALID, "deform: pitch smaller than channels");
  if ((long long)batch * H * W == 0) return 0;
  DwParams p; memset(&p, 0, sizeof(p));
  fill_common(p, d, in, in_pitch, out, out_pitch, batch, H, W, in_shift, 1, zx);
  p.Ms = sc->Ms; p.bs = sc->bs; p.ss = sc->ss; p.zs = sc->zs;
  p.u_lo = (double)(-sc->bound + 1); p.u_hi = (double)sc->bound;
  p.sval = sval;
  if (p.total == 0) return 0;
  CDN_CHECK((long long)batch * p.Hs * p.Ws * p.in_pitch_w < (1ll << 31) && p.total < (1ll << 31) && p.Hout < 65536 && p.Wout < 65536,
            CDN_ERR_INVALID, "deform: tensor too large for 32-bit indexing");
  if (deform_tile_ok(d, sc, in_pitch, out_pitch, batch, H, W, in_shift))
    return deform_tile_launch(d, sc, in, in_pitch, out, out_pitch, batch, H, W, in_shift, zx, sval, &p, st);
  // v3 tile size: at least ~3 tiles per CTA slot, and wide layers (8 channel groups per pixel) keep small tiles
  p.np = DEF_NP;
  if (sc->mode == 0) {
    const long long slots = (long long)cdn_num_sms() * DEF3_CTAS;
    const int want = p.G >= 8 ? 64 : (p.G >= 4 ? 128 : DEF3_MAX_NP);
    p.np = want;
    while (p.np > 64 && (p.total + p.np - 1) / p.np < 3 * slots) p.np >>= 1;
  }
  const long long ntiles = (p.total + p.np - 1) / p.np;
  const long long cap = (long long)cdn_num_sms() * (sc->mode == 0 ? DEF3_CTAS : 8);
  const unsigned blocks = (unsigned)std::min(ntiles, cap);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (sc->mode == 0) {
    CDN_CHECK(d.s_thr != nullptr && d.s_mode0_ok, CDN_ERR_STATE, "deform: integer-offset thresholds were not built (deform_scale_build)");
    p.s_thr = d.s_thr; p.s_n = d.s_n; p.s_lo = d.s_lo;
    p.pad_px = d.pad_px;
    if (!d.use_int) CDN_CUDA(cudaLaunchKernelEx(&cfg, deform_int_v3_kernel<0>, p));
    else if (d.rq.lo > -128) CDN_CUDA(cudaLaunchKernelEx(&cfg, deform_int_v3_kernel<2>, p));
    else CDN_CUDA(cudaLaunchKernelEx(&cfg, deform_int_v3_kernel<1>, p));
  } else CDN_CUDA(cudaLaunchKernelEx(&cfg, deform_dw_v2_kernel<1, false>, p));
  CDN_LAUNCH_CHECK("deform_dw_v2_kernel");
  return 0;
}
