// Parameter block and the exact (fp64) bilinear evaluation shared by the deformable kernels of dw.cu and deform_tile.cu.
#pragma once
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
  const float2* mb; const int* abm; float thr_layer;   // lean requantisation constants (v2 kernels)
  float thr_bil;                             // guard of the fp32 bilinear fast path (0.5 - eps, host-derived bound)
  const int4* ki; int lo_i;                  // integer requantisation (RqInt per channel, acc_bias folded in)
  int np; const uint32_t* pad_px;            // v3: pixels per tile; a pixel of pad words (q = -zx in every channel)
  const int* s_thr; int s_n, s_lo;           // integer offsets (v3): s = s_lo + #{k : acc_s >= s_thr[k]}, thresholds ascending
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

// Exact bilinear evaluation of one output word (4 channels): fp64 on exact integers in the oracle's operation order
// (dcn_deform_conv_cuda_kernel.cu:83-114,210-227).  Used for every element by nothing any more: it is the fallback of the
// guarded fp32 fast path below (and the definition of the result).
static __device__ __noinline__ uint32_t deform_bilinear_exact_word(const DwParams& p, const uint32_t* __restrict__ img, int h, int w,
                                                            double s, int cw) {
  LaneConsts k; load_lane_consts(p, cw, true, k);        // rare path: fetch its own constants, keep the fast loop lean
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
  return r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
}

