// Stem: fp32 NCHW image -> 3x3 conv (8-bit integer weights) -> ReLU -> QuantAct -> int8 NHWC [-> MaxPool(3,2,1)].
// Restates layer0 of the quantised graph (portable_quantizer/quantization_utils/quantize_model.py:26-35).
// The image is raw fp32, so the accumulation runs in fp64: every product (fp32 value x int8) is exact and the sum
// follows the oracle's (ci, i, j) order, which makes the accumulator bit-identical to the CPU restatement.
#include "layers.cuh"

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
  if (img_u8) stem_kernel<true><<<(unsigned)((p.total + 127) / 128), 128, 0, st>>>(p);
  else stem_kernel<false><<<(unsigned)((p.total + 127) / 128), 128, 0, st>>>(p);
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
