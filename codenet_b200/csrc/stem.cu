// Stem: fp32 NCHW image -> 3x3 conv (8-bit integer weights) -> ReLU -> QuantAct -> int8 NHWC [-> MaxPool(3,2,1)].
// Restates layer0 of the quantised graph (portable_quantizer/quantization_utils/quantize_model.py:26-35).
// The image is raw fp32, so the accumulation runs in fp64: every product (fp32 value x int8) is exact and the sum
// follows the oracle's (ci, i, j) order, which makes the accumulator bit-identical to the CPU restatement.
#include "layers.cuh"
#include <algorithm>

#define STEM_MAXC 32

struct StemParams {
  const float* img; const uint8_t* img_u8; const float* lut; int8_t* out;   // img_u8 != nullptr: uint8 HWC input + LUT
  int H, W, Ho, Wo, stride, C, out_pitch;
  long long total;
  const double* w;                           // [C][27] as doubles
  const double* M; const double* B;
  double lo;
  // the same constants by value: kernel parameters live in the constant bank, so a fully unrolled DFMA takes its
  // weight operand straight from c[0x0][imm] (the shared-memory version issued one LDS per DFMA and was LSU-bound)
  double cw[STEM_MAXC * 27]; double cM[STEM_MAXC], cB[STEM_MAXC];
  // guarded fp32 fast path (stem_fast_kernel): weights as fp32 pairs (taps 2j, 2j+1; 27 padded to 28), the requantisation line in
  // fp32, and the per-channel guard  thr = thr0 - ga * sum|x|  (see the kernel)
  alignas(8) float wf[STEM_MAXC * 28]; alignas(8) float Mf[STEM_MAXC]; alignas(8) float Bf[STEM_MAXC]; alignas(8) float ga[STEM_MAXC]; alignas(8) float thr0[STEM_MAXC];
};

// U8 = true: the image arrives as uint8 [B][H][W][3] (what cv2 hands to the reference's pre_process,
// lib/detectors/base_detector.py:48-76) and is normalised through a 3x256 look-up table holding
// fl32((u/255. - mean[c]) / std[c]) exactly as numpy evaluates it (:66), so the result is bit-identical to feeding the
// pre-normalised fp32 image -- with a quarter of the host-to-device bytes.
template <bool U8>
__global__ void __launch_bounds__(128) stem_kernel(const __grid_constant__ StemParams p) {
  pdl_launch_dependents();
  __shared__ float slut[U8 ? 768 : 1];
  if (U8) for (int i = threadIdx.x; i < 768; i += blockDim.x) slut[i] = p.lut[i];
  __syncthreads();
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= p.total) return;
  int wo = (int)(pix % p.Wo); long long t = pix / p.Wo; int ho = (int)(t % p.Ho); long long b = t / p.Ho;
  double x[27];
  if (U8) {
    const uint8_t* im = p.img_u8 + (size_t)b * p.H * p.W * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int y = ho * p.stride - 1 + i;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int xx = wo * p.stride - 1 + j;
        bool ok = (unsigned)y < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
        const uint8_t* px = im + ((size_t)y * p.W + xx) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) x[ci * 9 + i * 3 + j] = ok ? (double)slut[ci * 256 + __ldg(px + ci)] : 0.0;
      }
    }
  } else {
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = p.img + ((size_t)b * 3 + ci) * p.H * p.W;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        int y = ho * p.stride - 1 + i;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          int xx = wo * p.stride - 1 + j;
          bool ok = (unsigned)y < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
          x[ci * 9 + i * 3 + j] = ok ? (double)__ldg(plane + (size_t)y * p.W + xx) : 0.0;
        }
      }
    }
  }
  const int lo_i = (int)p.lo;
  uint32_t words[STEM_MAXC / 4];
#pragma unroll
  for (int i = 0; i < STEM_MAXC / 4; ++i) words[i] = 0;
#pragma unroll
  for (int c = 0; c < STEM_MAXC; ++c) {
    if (c < p.C) {                                   // uniform
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 27; ++k) acc = fma(p.cw[c * 27 + k], x[k], acc);   // products exact => fma == mul,add
      const double td = __dadd_rn(__dmul_rn(acc, p.cM[c]), p.cB[c]);
      // round-half-even + saturating conversion in one instruction, clamp on the integer side (same result as
      // clamp(rint(td)) for every td; the fp64 min/max/rint of the first version were three slow-pipe operations)
      const int q = min(max(__double2int_rn(td), lo_i), 127);
      words[c >> 2] |= (uint32_t)(q & 0xff) << (8 * (c & 3));
    }
  }
  uint4* dst = (uint4*)(p.out + (size_t)pix * p.out_pitch);
  dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
  dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
}


// ---- guarded fp32 fast path ------------------------------------------------------------------------------------------------
// The definition of the result is the fp64 chain of stem_kernel.  Here the 27-term dot product runs in fp32 -- two interleaved
// fma chains over the even / odd taps as packed FFMA2 (14 instead of 27 issue slots, on a pipe with four times the fp64 rate) -- and
// the result is used only when it provably rounds like the fp64 value:
//   acc' = fp32 result, acc = fp64 chain, T = sum_k |w_k x_k| <= Wmax * S with S = sum_k |x_k|:   |acc' - acc| <= 30 u T   (28 roundings,
//   u = 2^-24, plus the fp64 chain's own 27 * 2^-53 T);   t' = fmaf(acc', M32, B32) against t = fl64(fl64(acc M) + B):
//   |t' - t| <= 30 u T M (1 + u) + T u M + u |B| + u |t'| + 2^-51 |t|  <=  34 u M Wmax S + 1.05 u (|B| + 300)  =: ga * S + gb   for |t'| <= 300.
// With r = rint(t') the fp64 value rounds to the same integer whenever |t' - r| < 0.5 - (ga S + gb); |t'| > 300 saturates either way
// (ga S + gb is far below 100 there or the test fails and the channel is re-evaluated).  Channels that fail the test -- a few per
// thousand -- are re-evaluated by the fp64 chain itself, so every output equals stem_kernel's bit for bit (NaN / Inf inputs fail the
// comparison and take the fp64 path too).
__device__ __forceinline__ unsigned long long stem_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
// VS = 4 / 2: the conv stride, known at compile time -- a thread fetches the VS pixels  VS*wo .. VS*wo + VS-1  of a row with ONE
// aligned vector load (lanes of a warp are consecutive output columns: fully coalesced, where 27 scalar loads per thread touched
// four times the sectors they used) and takes the left tap, column VS*wo - 1, from its neighbour lane's vector by shuffle (the first
// lane of a warp, or of an image row, loads that one pixel itself / pads).  VS = 0: scalar loads for any stride.
template <bool U8, int VS>
__global__ void __launch_bounds__(128) stem_fast_kernel(const __grid_constant__ StemParams p) {
  pdl_launch_dependents();
  __shared__ float slut[U8 ? 768 : 1];
  if (U8) for (int i = threadIdx.x; i < 768; i += blockDim.x) slut[i] = p.lut[i];
  __syncthreads();
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < p.total;
  if (VS == 0 && !live) return;
  if (!live) pix = p.total - 1;                      // vector variant: every lane takes part in the shuffles
  int wo = (int)(pix % p.Wo); long long t = pix / p.Wo; int ho = (int)(t % p.Ho); long long b = t / p.Ho;
  float x[28];
  x[27] = 0.f;
  if (VS > 0) {
    const int lane = threadIdx.x & 31;
    const bool from_left = lane > 0 && wo > 0;       // the neighbour lane holds column VS*wo - 1 of the same row
    if (U8) {
      // uint8 HWC: the VS = 4 pixels are 12 bytes = three aligned words; the left tap is the last pixel of the neighbour's third word
      const uint8_t* im = p.img_u8 + (size_t)b * p.H * p.W * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int y = ho * VS - 1 + i;
        const bool yok = (unsigned)y < (unsigned)p.H;
        const uint32_t* row = (const uint32_t*)(im + ((size_t)(yok ? y : 0) * p.W + (size_t)VS * wo) * 3);
        const uint32_t w0 = __ldg(row), w1 = __ldg(row + 1), w2 = __ldg(row + 2);
        uint32_t left = __shfl_up_sync(0xffffffffu, w2 >> 8, 1);
        if (!from_left) left = wo > 0 ? ((uint32_t)__ldg((const uint8_t*)row - 3) | ((uint32_t)__ldg((const uint8_t*)row - 2) << 8) | ((uint32_t)__ldg((const uint8_t*)row - 1) << 16)) : 0u;
        const uint32_t p0 = left, p1 = w0 & 0xffffffu, p2 = (w0 >> 24) | ((w1 & 0xffffu) << 8);
        const bool ok0 = yok && wo > 0;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          x[ci * 9 + i * 3 + 0] = ok0 ? slut[ci * 256 + ((p0 >> (8 * ci)) & 0xffu)] : 0.f;
          x[ci * 9 + i * 3 + 1] = yok ? slut[ci * 256 + ((p1 >> (8 * ci)) & 0xffu)] : 0.f;
          x[ci * 9 + i * 3 + 2] = yok ? slut[ci * 256 + ((p2 >> (8 * ci)) & 0xffu)] : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* plane = p.img + ((size_t)b * 3 + ci) * p.H * p.W;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int y = ho * VS - 1 + i;
          const bool yok = (unsigned)y < (unsigned)p.H;
          const float* row = plane + (size_t)(yok ? y : 0) * p.W + (size_t)VS * wo;
          float v0, v1, vl;
          if (VS == 4) { const float4 v = __ldg((const float4*)row); v0 = v.x; v1 = v.y; vl = v.w; }
          else { const float2 v = __ldg((const float2*)row); v0 = v.x; v1 = v.y; vl = v.y; }
          float left = __shfl_up_sync(0xffffffffu, vl, 1);
          if (!from_left) left = wo > 0 ? __ldg(row - 1) : 0.f;
          x[ci * 9 + i * 3 + 0] = (yok && wo > 0) ? left : 0.f;
          x[ci * 9 + i * 3 + 1] = yok ? v0 : 0.f;
          x[ci * 9 + i * 3 + 2] = yok ? v1 : 0.f;
        }
      }
    }
    if (!live) return;
  } else if (U8) {
    const uint8_t* im = p.img_u8 + (size_t)b * p.H * p.W * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      int y = ho * p.stride - 1 + i;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int xx = wo * p.stride - 1 + j;
        bool ok = (unsigned)y < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
        const uint8_t* px = im + ((size_t)y * p.W + xx) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) x[ci * 9 + i * 3 + j] = ok ? slut[ci * 256 + __ldg(px + ci)] : 0.f;
      }
    }
  } else {
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = p.img + ((size_t)b * 3 + ci) * p.H * p.W;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        int y = ho * p.stride - 1 + i;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          int xx = wo * p.stride - 1 + j;
          bool ok = (unsigned)y < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
          x[ci * 9 + i * 3 + j] = ok ? __ldg(plane + (size_t)y * p.W + xx) : 0.f;
        }
      }
    }
  }
  float S = 0.f;
#pragma unroll
  for (int k = 0; k < 27; ++k) S += fabsf(x[k]);
  S *= 1.0001f;                                      // the fp32 sum of 27 terms is below the true sum by at most 27 u
  unsigned long long xp[14];
#pragma unroll
  for (int j = 0; j < 14; ++j) xp[j] = ((unsigned long long)__float_as_uint(x[2 * j + 1]) << 32) | __float_as_uint(x[2 * j]);
  const int lo_i = (int)p.lo;
  const float lo_f = (float)lo_i;
  uint32_t words[STEM_MAXC / 4];
#pragma unroll
  for (int i = 0; i < STEM_MAXC / 4; ++i) words[i] = 0;
  uint32_t redo = 0;
  // two channels at a time: the requantisation line, the rounding and the guard as packed fp32 operations.  The value is clamped to
  // [lo, 127] BEFORE rounding (inside the clamp the distance to the rounding boundary is what it was; outside it the result is the
  // clamp value for t' and for the fp64 t alike, as long as the guard holds -- and a guard <= 0 fails the test), so the low byte of
  // the magic-biased float is the int8 result and no integer min / max is needed.
  const unsigned long long S2 = ((unsigned long long)__float_as_uint(S) << 32) | __float_as_uint(S);
  const unsigned long long MAG2 = ((unsigned long long)__float_as_uint(CDN_MAGIC_F) << 32) | __float_as_uint(CDN_MAGIC_F);
  const unsigned long long NMAG2 = ((unsigned long long)__float_as_uint(-CDN_MAGIC_F) << 32) | __float_as_uint(-CDN_MAGIC_F);
  auto f2lo = [](unsigned long long v) { return __uint_as_float((uint32_t)v); };
  auto f2hi = [](unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); };
  auto pk = [](float lo, float hi) { return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo); };
  auto add2 = [](unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; };
#pragma unroll
  for (int c = 0; c < STEM_MAXC; c += 4) {
    if (c < p.C) {                                   // uniform
      uint32_t rb[4];
#pragma unroll
      for (int h = 0; h < 4; h += 2) {
        float acc[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          unsigned long long a2 = 0ull;
#pragma unroll
          for (int j = 0; j < 14; ++j) a2 = stem_ffma2(xp[j], *reinterpret_cast<const unsigned long long*>(&p.wf[(c + h + e) * 28 + 2 * j]), a2);
          acc[e] = __fadd_rn(f2lo(a2), f2hi(a2));
        }
        const unsigned long long M2 = *reinterpret_cast<const unsigned long long*>(&p.Mf[c + h]);
        const unsigned long long B2 = *reinterpret_cast<const unsigned long long*>(&p.Bf[c + h]);
        const unsigned long long G2 = *reinterpret_cast<const unsigned long long*>(&p.ga[c + h]);     // holds -ga
        const unsigned long long T2 = *reinterpret_cast<const unsigned long long*>(&p.thr0[c + h]);
        const unsigned long long t2 = stem_ffma2(pk(acc[0], acc[1]), M2, B2);
        const float t0 = fminf(fmaxf(f2lo(t2), lo_f), 127.f), t1 = fminf(fmaxf(f2hi(t2), lo_f), 127.f);
        const unsigned long long tc = pk(t0, t1);
        const unsigned long long rm = add2(tc, MAG2);            // rint(t) + 1.5 * 2^23 per half
        const unsigned long long r = add2(rm, NMAG2);
        const unsigned long long d = stem_ffma2(r, pk(-1.f, -1.f), tc);   // t - r, exact
        const unsigned long long thr = stem_ffma2(G2, S2, T2);
        // a non-finite pixel makes S (and with it thr) NaN or -Inf: the comparison fails and the channel takes the fp64 path
        if (!(fabsf(f2lo(d)) < f2lo(thr))) redo |= 1u << (c + h);
        if (!(fabsf(f2hi(d)) < f2hi(thr))) redo |= 1u << (c + h + 1);
        rb[h] = (uint32_t)rm; rb[h + 1] = (uint32_t)(rm >> 32);
      }
      words[c >> 2] = pack4_lowbytes(rb[0], rb[1], rb[2], rb[3]);
    }
  }
  if (p.C & 3) { words[p.C >> 2] &= (1u << (8 * (p.C & 3))) - 1u; redo &= (1u << p.C) - 1u; }   // bytes of channels >= C stay zero
  while (redo) {                                     // rare: the fp64 chain of stem_kernel for the channels that were too close to call
    const int c = __ffs(redo) - 1;
    redo &= redo - 1;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 27; ++k) acc = fma(p.cw[c * 27 + k], (double)x[k], acc);
    const double td = __dadd_rn(__dmul_rn(acc, p.cM[c]), p.cB[c]);
    const int q = min(max(__double2int_rn(td), lo_i), 127);
    const int sh = 8 * (c & 3);
#pragma unroll
    for (int i = 0; i < STEM_MAXC / 4; ++i)
      if (i == (c >> 2)) words[i] = (words[i] & ~(0xffu << sh)) | ((uint32_t)(q & 0xff) << sh);
  }
  uint4* dst = (uint4*)(p.out + (size_t)pix * p.out_pitch);
  dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
  dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
}

// MaxPool2d(3, 2, 1) on the int8 grid (monotone, so it commutes with the quantiser; padding = -inf)
__global__ void maxpool3s2_kernel(const uint32_t* in, uint32_t* out, int H, int W, int Ho, int Wo, int pitch_w,
                                  long long total_words) {
  pdl_launch_dependents();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_words) return;
  int cw = (int)(idx % pitch_w); long long pix = idx / pitch_w;
  int wo = (int)(pix % Wo); long long t = pix / Wo; int ho = (int)(t % Ho); long long b = t / Ho;
  uint32_t m = 0x80808080u;
  for (int i = 0; i < 3; ++i) {
    int y = 2 * ho - 1 + i; if ((unsigned)y >= (unsigned)H) continue;
    for (int j = 0; j < 3; ++j) {
      int x = 2 * wo - 1 + j; if ((unsigned)x >= (unsigned)W) continue;
      m = __vmaxs4(m, __ldg(in + (((size_t)b * H + y) * W + x) * pitch_w + cw));
    }
  }
  out[idx] = m;
}


int stem_device_build(StemDevice& d, const int8_t* wq, int C, const cdn_requant* rq) {
  CDN_CHECK(C > 0 && C <= STEM_MAXC, CDN_ERR_INVALID, "stem: C=%d must be in 1..%d", C, STEM_MAXC);
  CDN_CHECK(rq && rq->n == C && rq->M && rq->B, CDN_ERR_INVALID, "stem: requant constants must have n == C");
  std::vector<double> w(C * 27);
  for (int i = 0; i < C * 27; ++i) w[i] = (double)wq[i];
  if (dev_upload(&d.w, w.data(), w.size())) return CDN_ERR_CUDA;
  if (dev_upload(&d.M, rq->M, C)) return CDN_ERR_CUDA;
  if (dev_upload(&d.B, rq->B, C)) return CDN_ERR_CUDA;
  d.C = C; d.lo = (double)rq->lo;
  d.hw = w; d.hM.assign(rq->M, rq->M + C); d.hB.assign(rq->B, rq->B + C);
  return 0;
}
void stem_device_free(StemDevice& d) { cudaFree(d.w); cudaFree(d.M); cudaFree(d.B); d = StemDevice(); }

// tmp: scratch for the un-pooled map when pool != 0 (batch*Ho*Wo*out_pitch bytes), else unused.
int stem_launch(const StemDevice& d, const float* img, int batch, int H, int W, int stride, int pool,
                int8_t* out, int out_pitch, int8_t* tmp, cudaStream_t st, const uint8_t* img_u8, const float* lut) {
  CDN_CHECK(out_pitch == 32, CDN_ERR_INVALID, "stem: out pitch must be 32");
  CDN_CHECK(stride >= 1 && stride <= 4, CDN_ERR_INVALID, "stem: bad stride");
  StemParams p;
  p.img = img; p.img_u8 = img_u8; p.lut = lut; p.H = H; p.W = W; p.stride = stride; p.C = d.C; p.out_pitch = out_pitch;
  p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
  p.total = (long long)batch * p.Ho * p.Wo;
  p.w = d.w; p.M = d.M; p.B = d.B; p.lo = d.lo;
  memset(p.cw, 0, sizeof(p.cw)); memset(p.cM, 0, sizeof(p.cM)); memset(p.cB, 0, sizeof(p.cB));
  memcpy(p.cw, d.hw.data(), d.hw.size() * sizeof(double));
  memcpy(p.cM, d.hM.data(), d.hM.size() * sizeof(double)); memcpy(p.cB, d.hB.data(), d.hB.size() * sizeof(double));
  p.out = pool ? tmp : out;
  CDN_CHECK(!pool || tmp, CDN_ERR_INVALID, "stem: pooling needs a scratch buffer");
  if (p.total == 0) return 0;
  CDN_CHECK(img_u8 == nullptr || lut != nullptr, CDN_ERR_INVALID, "stem: uint8 input needs the normalisation table");
  // guarded fp32 fast path: constants of the requantisation line in fp32 and the per-channel guard (derivation at the kernel)
  memset(p.wf, 0, sizeof(p.wf));
  const double u = ldexp(1.0, -24);
  bool fast = !(g_cdn_debug_flags & (1u << 26));                         // bit 26: the all-fp64 kernel (A/B, cross-check)
  for (int c = 0; c < d.C; ++c) {
    double wmax = 0.0;
    for (int k = 0; k < 27; ++k) { p.wf[c * 28 + k] = (float)d.hw[c * 27 + k]; wmax = std::max(wmax, fabs(d.hw[c * 27 + k])); }
    const double M = d.hM[c], B = d.hB[c];
    p.Mf[c] = (float)M; p.Bf[c] = (float)B;
    const double ga = 34.0 * u * fabs(M) * wmax, gb = 1.05 * u * (fabs(B) + 300.0);
    p.ga[c] = -(float)(ga * 1.0001); p.thr0[c] = (float)((0.5 - gb) * 0.9999);      // ga is stored negated: thr = fma(-ga, S, thr0)
    if (!std::isfinite(M) || !std::isfinite(B) || !(0.5 - gb > 0.01)) fast = false;
  }
  for (int c = d.C; c < STEM_MAXC; ++c) { p.Mf[c] = 0.f; p.Bf[c] = 0.f; p.ga[c] = 0.f; p.thr0[c] = 0.5f; }
  const unsigned grid = (unsigned)((p.total + 127) / 128);
  // vector loads need the taps 1, 2 inside the image for every output column (W a multiple of the stride) and aligned rows
  const bool vec = !(g_cdn_debug_flags & (1u << 27)) && W % stride == 0 && W % 4 == 0 &&                  // bit 27: scalar loads (A/B)
                   (img_u8 ? (stride == 4 && ((uintptr_t)img_u8 & 3) == 0) : ((stride == 4 || stride == 2) && ((uintptr_t)img & 15) == 0));
  if (fast) {
    if (img_u8) { if (vec) stem_fast_kernel<true, 4><<<grid, 128, 0, st>>>(p); else stem_fast_kernel<true, 0><<<grid, 128, 0, st>>>(p); }
    else if (vec && stride == 4) stem_fast_kernel<false, 4><<<grid, 128, 0, st>>>(p);
    else if (vec) stem_fast_kernel<false, 2><<<grid, 128, 0, st>>>(p);
    else stem_fast_kernel<false, 0><<<grid, 128, 0, st>>>(p);
  }
  else if (img_u8) stem_kernel<true><<<grid, 128, 0, st>>>(p);
  else stem_kernel<false><<<grid, 128, 0, st>>>(p);
  CDN_LAUNCH_CHECK("stem_kernel");
  if (pool) {
    int Hp = (p.Ho - 1) / 2 + 1, Wp = (p.Wo - 1) / 2 + 1;
    long long words = (long long)batch * Hp * Wp * (out_pitch / 4);
    maxpool3s2_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>((const uint32_t*)tmp, (uint32_t*)out, p.Ho, p.Wo,
                                                                        Hp, Wp, out_pitch / 4, words);
    CDN_LAUNCH_CHECK("maxpool3s2_kernel");
  }
  return 0;
}


// ---- module-level helpers (compat.QuantAct / MaxPool on an int8 grid) --------------------------------------------------
// QuantAct.forward on a real-valued tensor (quant_modules.py:202-225 with frozen range; quant_utils.py:31-39):
// q = rint(fl64(fl64(s * x) - z)) saturated to int8, fp32 NCHW in, int8 NHWC out (channels beyond C zero-filled).
__global__ void quantize_f32_i8_kernel(const float* __restrict__ in, int8_t* __restrict__ out, int C, int pitch, long long hw,
                                       long long total, double s, double z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (pixel, channel slot)
  if (i >= total) return;
  const int c = (int)(i % pitch); const long long pix = i / pitch;
  const long long b = pix / hw, r = pix - b * hw;
  int q = 0;
  if (c < C) {
    const double t = __dsub_rn(__dmul_rn(s, (double)__ldg(in + ((size_t)b * C + c) * hw + r)), z);
    q = (int)fmin(fmax(rint(t), -128.0), 127.0);
  }
  out[i] = (int8_t)q;
}
extern "C" int cdn_quantize_f32_i8(const float* d_in, int batch, int C, int H, int W, double scale, double zero, int8_t* d_out,
                                   int out_pitch, cdn_stream_t stream) {
  CDN_CHECK(d_in && d_out && batch >= 0 && C >= 1 && H >= 1 && W >= 1 && out_pitch >= C, CDN_ERR_INVALID, "quantize: bad arguments");
  const long long total = (long long)batch * H * W * out_pitch;
  if (total == 0) return 0;
  quantize_f32_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, C, out_pitch, (long long)H * W,
                                                                                             total, scale, zero);
  CDN_LAUNCH_CHECK("quantize_f32_i8_kernel");
  return 0;
}
// nn.MaxPool2d(3, 2, 1) on an int8 NHWC grid (the stem's pool when the graph is run module by module)
extern "C" int cdn_maxpool3s2_i8(const int8_t* d_in, int batch, int H, int W, int pitch, int8_t* d_out, cdn_stream_t stream) {
  CDN_CHECK(d_in && d_out && batch >= 0 && H >= 1 && W >= 1 && pitch % 4 == 0, CDN_ERR_INVALID, "maxpool: bad arguments");
  const int Hp = (H - 1) / 2 + 1, Wp = (W - 1) / 2 + 1;
  const long long words = (long long)batch * Hp * Wp * (pitch / 4);
  if (words == 0) return 0;
  maxpool3s2_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)d_in, (uint32_t*)d_out, H, W, Hp, Wp,
                                                                                        pitch / 4, words);
  CDN_LAUNCH_CHECK("maxpool3s2_kernel");
  return 0;
}
