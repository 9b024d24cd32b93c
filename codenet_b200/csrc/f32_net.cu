// fp32 NCHW building blocks of the float CoDeNet model (BASELINE config 5 / SURVEY.md 8(a) a1, a10-a12): BN is folded on
// the host, so every layer is conv + bias (+ ReLU).  These are plain SIMT kernels (coalesced along W, fp32 FMA
// accumulation -- the 1e-4 contract of the float path rules TF32 out); the W4A8 path is the optimised one.
#include "common.cuh"
#include <algorithm>

// ---- dense 3x3 conv for the stem (Ci = 3): one thread per output pixel computes every output channel --------------
#define C3_MAXCO 32
__global__ void __launch_bounds__(128) conv3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ out, int Ci, int Co, int H, int W, int Ho, int Wo, int stride,
                                                          int relu, long long total) {
  extern __shared__ float sw[];                 // [Co][Ci][9] + [Co]
  for (int i = threadIdx.x; i < Co * Ci * 9; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Co; i += blockDim.x) sw[Co * Ci * 9 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo); const long long t = idx / Wo; const int ho = (int)(t % Ho); const long long b = t / Ho;
  float acc[C3_MAXCO];
#pragma unroll
  for (int c = 0; c < C3_MAXCO; ++c) acc[c] = c < Co ? sw[Co * Ci * 9 + c] : 0.f;
  for (int ci = 0; ci < Ci; ++ci) {
    const float* plane = in + ((size_t)b * Ci + ci) * H * W;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int y = ho * stride - 1 + i;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int x = wo * stride - 1 + j;
        const float v = ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) ? __ldg(plane + (size_t)y * W + x) : 0.f;
#pragma unroll
        for (int c = 0; c < C3_MAXCO; ++c) if (c < Co) acc[c] = fmaf(sw[(c * Ci + ci) * 9 + i * 3 + j], v, acc[c]);
      }
    }
  }
  for (int c = 0; c < Co; ++c) {
    float v = acc[c];
    if (relu) v = fmaxf(v, 0.f);
    out[(((size_t)b * Co + c) * Ho + ho) * Wo + wo] = v;
  }
}

// the stride-4 stem (Ci = 3, W % 4 == 0): the three taps of a row are columns 4 wo - 1 .. 4 wo + 1, i.e. one aligned 16-byte load
// per lane (fully coalesced where the scalar kernel reads every fourth float) plus the left halo; weights transposed to
// [ci][tap][co] in shared memory so that one broadcast LDS.128 feeds four output channels.  Same FMA order per output.
__global__ void __launch_bounds__(128) conv3x3_s4_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ out, int Ci, int Co, int H, int W, int Ho, int Wo, int relu,
                                                             unsigned total) {
  extern __shared__ __align__(16) float sw[];   // [Ci][9][C3_MAXCO] + [C3_MAXCO]
  for (int i = threadIdx.x; i < Ci * 9 * C3_MAXCO; i += blockDim.x) {
    const int co = i % C3_MAXCO, t = i / C3_MAXCO;                  // t = ci * 9 + tap
    sw[i] = co < Co ? w[(size_t)co * Ci * 9 + t] : 0.f;
  }
  for (int i = threadIdx.x; i < C3_MAXCO; i += blockDim.x) sw[Ci * 9 * C3_MAXCO + i] = (bias && i < Co) ? bias[i] : 0.f;
  __syncthreads();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const unsigned wo = idx % (unsigned)Wo; const unsigned t = idx / (unsigned)Wo; const unsigned ho = t % (unsigned)Ho; const unsigned b = t / (unsigned)Ho;
  float acc[C3_MAXCO];
#pragma unroll
  for (int c = 0; c < C3_MAXCO; ++c) acc[c] = sw[Ci * 9 * C3_MAXCO + c];
  for (int ci = 0; ci < Ci; ++ci) {
    const float* plane = in + ((size_t)b * Ci + ci) * H * W;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int y = (int)ho * 4 - 1 + i;
      const bool rv = (unsigned)y < (unsigned)H;
      const float* rp = plane + (size_t)(rv ? y : 0) * W + 4 * wo;
      const float4 q = __ldg((const float4*)rp);
      const float v[3] = {(rv && wo > 0) ? __ldg(rp - 1) : 0.f, rv ? q.x : 0.f, rv ? q.y : 0.f};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4* wp = (const float4*)(sw + ((ci * 9 + i * 3 + j) * C3_MAXCO));
#pragma unroll
        for (int c4 = 0; c4 < C3_MAXCO / 4; ++c4) {
          const float4 ww = wp[c4];
          acc[4 * c4] = fmaf(ww.x, v[j], acc[4 * c4]); acc[4 * c4 + 1] = fmaf(ww.y, v[j], acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(ww.z, v[j], acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(ww.w, v[j], acc[4 * c4 + 3]);
        }
      }
    }
  }
  float* op = out + ((size_t)b * Co * Ho + ho) * Wo + wo;
#pragma unroll
  for (int c = 0; c < C3_MAXCO; ++c)
    if (c < Co) op[(size_t)c * Ho * Wo] = relu ? fmaxf(acc[c], 0.f) : acc[c];
}

extern "C" int cdn_conv3x3_f32(const float* input, const float* weight, const float* bias, float* output, int B, int Ci, int Co,
                               int H, int W, int stride, int relu, cdn_stream_t stream) {
  CDN_CHECK(input && weight && output && Ci >= 1 && Co >= 1 && Co <= C3_MAXCO && stride >= 1, CDN_ERR_INVALID,
            "conv3x3_f32: bad arguments (Co <= %d)", C3_MAXCO);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = (long long)B * Ho * Wo;
  if (total == 0) return 0;
  if (stride == 4 && W % 4 == 0 && ((uintptr_t)input & 15) == 0 && total < (1ll << 32) - 128 && (size_t)(Ci * 9 + 1) * C3_MAXCO * 4 <= 48 * 1024) {
    conv3x3_s4_f32_kernel<<<(unsigned)((total + 127) / 128), 128, (size_t)(Ci * 9 + 1) * C3_MAXCO * sizeof(float), (cudaStream_t)stream>>>(
        input, weight, bias, output, Ci, Co, H, W, Ho, Wo, relu, (unsigned)total);
    CDN_LAUNCH_CHECK("conv3x3_s4_f32_kernel");
    return 0;
  }
  const size_t smem = ((size_t)Co * Ci * 9 + Co) * sizeof(float);
  conv3x3_f32_kernel<<<(unsigned)((total + 127) / 128), 128, smem, (cudaStream_t)stream>>>(input, weight, bias, output, Ci, Co, H, W, Ho, Wo,
                                                                                         stride, relu, total);
  CDN_LAUNCH_CHECK("conv3x3_f32_kernel");
  return 0;
}

// ---- depthwise 3x3, pad 1, stride 1/2 -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ out, int C, int H, int W, int Ho, int Wo, int stride, int relu,
                                                        long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo); long long t = idx / Wo; const int ho = (int)(t % Ho); t /= Ho; const int c = (int)(t % C);
  const float* plane = in + (size_t)t * H * W;                      // t = b*C + c
  const float* wk = w + c * 9;
  float acc = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int y = ho * stride - 1 + i;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int x = wo * stride - 1 + j;
      if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) acc = fmaf(__ldg(wk + i * 3 + j), __ldg(plane + (size_t)y * W + x), acc);
    }
  }
  out[idx] = relu ? fmaxf(acc, 0.f) : acc;
}

// stride 1, W % 4 == 0: a thread produces 4 horizontally adjacent outputs from one float4 + two halo scalars per input row
// (1.5 loads per output instead of 9); same FMA order per output as the kernel above (bias, then taps row by row).
__global__ void __launch_bounds__(256) dw3x3_f32_s1v4_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ out, int C, int H, int W, int relu, long long total4) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int W4 = W >> 2;
  const int x0 = (int)(idx % W4) * 4; long long t = idx / W4; const int y = (int)(t % H); t /= H; const int c = (int)(t % C);
  const float* plane = in + (size_t)t * H * W;                      // t = b*C + c
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + c * 9 + i);
  const float b0 = bias ? __ldg(bias + c) : 0.f;
  float acc[4] = {b0, b0, b0, b0};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int yy = y - 1 + i;
    if ((unsigned)yy >= (unsigned)H) continue;                     // zero padding: the taps of this row contribute nothing
    const float* r = plane + (size_t)yy * W + x0;
    const float4 q = __ldg((const float4*)r);
    const bool hl = x0 > 0, hr = x0 + 4 < W;
    const float v[6] = {hl ? __ldg(r - 1) : 0.f, q.x, q.y, q.z, q.w, hr ? __ldg(r + 4) : 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const bool ok = (o + j >= 1 || hl) && (o + j <= 4 || hr);   // out-of-image taps are skipped, as in the scalar kernel
        if (ok) acc[o] = fmaf(wk[i * 3 + j], v[o + j], acc[o]);
      }
  }
  float4 o4 = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if (relu) { o4.x = fmaxf(o4.x, 0.f); o4.y = fmaxf(o4.y, 0.f); o4.z = fmaxf(o4.z, 0.f); o4.w = fmaxf(o4.w, 0.f); }
  *(float4*)(out + (size_t)t * H * W + (size_t)y * W + x0) = o4;
}

// stride 2, W % 8 == 0: one thread = 4 adjacent outputs of a row from two aligned 16-byte loads + the left halo per input row
// (9 loads per 4 outputs instead of 36, 32-bit index arithmetic); same taps in the same order per output.
__global__ void __launch_bounds__(256) dw3x3_f32_s2v4_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ out, int C, int H, int W, int Ho, int Wo, int relu, unsigned total4) {
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total4) return;
  const unsigned Wo4 = (unsigned)Wo >> 2;
  const unsigned x4 = idx % Wo4; unsigned t = idx / Wo4; const unsigned ho = t % (unsigned)Ho; const unsigned pl = t / (unsigned)Ho;
  const int c = (int)(pl % (unsigned)C);
  const float* plane = in + (size_t)pl * H * W;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + c * 9 + i);
  const float b0 = bias ? __ldg(bias + c) : 0.f;
  float acc[4] = {b0, b0, b0, b0};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int y = (int)ho * 2 - 1 + i;
    const bool rv = (unsigned)y < (unsigned)H;
    const float* rp = plane + (size_t)(rv ? y : 0) * W + 8 * x4;
    float4 a = __ldg((const float4*)rp), b = __ldg((const float4*)rp + 1);
    if (!rv) { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
    const float v[9] = {(rv && x4 > 0) ? __ldg(rp - 1) : 0.f, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = fmaf(wk[i * 3 + j], v[2 * o + j], acc[o]);
  }
  float4 o4 = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if (relu) { o4.x = fmaxf(o4.x, 0.f); o4.y = fmaxf(o4.y, 0.f); o4.z = fmaxf(o4.z, 0.f); o4.w = fmaxf(o4.w, 0.f); }
  *(float4*)(out + ((size_t)pl * Ho + ho) * Wo + 4 * x4) = o4;
}

// stride 1 with H and W multiples of 4: one thread = 4 x 4 outputs.  Six row reads (one 16-byte load each; the two halo
// columns come from the neighbouring lanes by shuffle when a row's threads share a warp) feed 16 outputs, where the 4 x 1
// kernel above issues 9 loads per 4 outputs and three 64-bit index divisions per thread.  Same taps in the same order per
// output (rows top to bottom, columns left to right); taps outside the image add an exact zero instead of being skipped.
template <bool SHFL>
__global__ void __launch_bounds__(256) dw3x3_f32_s1r4_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ out, int C, int H, int W, int relu, unsigned total) {
  unsigned idx = blockIdx.x * 256u + threadIdx.x;
  const bool active = idx < total;
  if (!active) idx = total - 1;                                     // keeps the lane alive for the shuffles
  const unsigned W4 = (unsigned)W >> 2, H4 = (unsigned)H >> 2;
  const unsigned x4 = idx % W4; unsigned t = idx / W4; const unsigned ys = t % H4; const unsigned pl = t / H4; const int c = (int)(pl % (unsigned)C);
  const int x0 = (int)x4 * 4, y0 = (int)ys * 4;
  const float* plane = in + (size_t)pl * H * W;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + c * 9 + i);
  const float b0 = bias ? __ldg(bias + c) : 0.f;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[r][o] = b0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int yy = y0 - 1 + r;
    const bool rv = (unsigned)yy < (unsigned)H;
    const float* rp = plane + (size_t)(rv ? yy : 0) * W + x0;
    float4 q = __ldg((const float4*)rp);
    if (!rv) q = make_float4(0.f, 0.f, 0.f, 0.f);
    float left, right;
    if (SHFL) {
      left = __shfl_up_sync(0xffffffffu, q.w, 1); right = __shfl_down_sync(0xffffffffu, q.x, 1);
      if (x4 == 0) left = 0.f;
      if (x4 == W4 - 1) right = 0.f;
    } else {
      left = (rv && x0 > 0) ? __ldg(rp - 1) : 0.f; right = (rv && x0 + 4 < W) ? __ldg(rp + 4) : 0.f;
    }
    const float v[6] = {left, q.x, q.y, q.z, q.w, right};
#pragma unroll
    for (int orow = 0; orow < 4; ++orow) {
      const int i = r - orow;
      if (i >= 0 && i < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[orow][o] = fmaf(wk[i * 3 + j], v[o + j], acc[orow][o]);
      }
    }
  }
  if (!active) return;
  float* op = out + (size_t)pl * H * W + (size_t)y0 * W + x0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float4 o4 = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    if (relu) { o4.x = fmaxf(o4.x, 0.f); o4.y = fmaxf(o4.y, 0.f); o4.z = fmaxf(o4.z, 0.f); o4.w = fmaxf(o4.w, 0.f); }
    *(float4*)(op + (size_t)r * W) = o4;
  }
}

// depthwise 3x3 (pad 1, stride 1) over the NEAREST x2 UPSAMPLED map without ever writing it: out[Y][X] = b + sum_ij w[i][j] *
// a[(Y+i-1) >> 1][(X+j-1) >> 1] (zero outside), a = the low-resolution plane.  Used for the heads: their first 1x1 conv + ReLU
// commute with nearest upsampling exactly (pointwise), so conv(upsample(z)) is computed as upsample(conv(z)) at a quarter of
// the pixels and the 4x larger tensor is only ever produced as this kernel's output (shufflenetv2_dcn.py:244-271 after the last
// deconv block :286-300).  One thread = 4 low-resolution pixels of a row = 2 x 8 outputs; taps in the order of the plain kernel.
__global__ void __launch_bounds__(256) dw3x3_up2_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ out, int C, int h, int wd, int relu, unsigned total) {
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const unsigned W4 = (unsigned)wd >> 2;
  const unsigned x4 = idx % W4; unsigned t = idx / W4; const unsigned y = t % (unsigned)h; const unsigned pl = t / (unsigned)h;
  const int c = (int)(pl % (unsigned)C), x0 = (int)x4 * 4;
  const float* plane = in + (size_t)pl * h * wd;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + c * 9 + i);
  const float b0 = bias ? __ldg(bias + c) : 0.f;
  float v[3][6];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = (int)y - 1 + r;
    const bool rv = (unsigned)yy < (unsigned)h;
    const float* rp = plane + (size_t)(rv ? yy : 0) * wd + x0;
    const float4 q = __ldg((const float4*)rp);
    v[r][0] = (rv && x0 > 0) ? __ldg(rp - 1) : 0.f;
    v[r][1] = rv ? q.x : 0.f; v[r][2] = rv ? q.y : 0.f; v[r][3] = rv ? q.z : 0.f; v[r][4] = rv ? q.w : 0.f;
    v[r][5] = (rv && x0 + 4 < wd) ? __ldg(rp + 4) : 0.f;
  }
  float* op = out + (size_t)pl * (4 * h * wd) + (size_t)(2 * y) * (2 * wd) + 2 * x0;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float acc = b0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int rr = 1 + ((dy + i - 1) >> 1);                         // low row of tap row i, relative to y - 1 .. y + 1: floor((dy + i - 1) / 2)
#pragma unroll
          for (int j = 0; j < 3; ++j) acc = fmaf(wk[i * 3 + j], v[rr][1 + k + ((dx + j - 1) >> 1)], acc);
        }
        o[2 * k + dx] = relu ? fmaxf(acc, 0.f) : acc;
      }
    *(float4*)(op + (size_t)dy * (2 * wd)) = make_float4(o[0], o[1], o[2], o[3]);
    *(float4*)(op + (size_t)dy * (2 * wd) + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}
extern "C" int cdn_dw3x3_up2_f32(const float* input, const float* weight, const float* bias, float* output, int B, int C, int h, int w,
                                 int relu, cdn_stream_t stream) {
  CDN_CHECK(input && weight && output && C >= 1 && h >= 1 && w >= 4 && w % 4 == 0 && B >= 0, CDN_ERR_INVALID,
            "dw3x3_up2_f32: bad arguments (the low-resolution width must be a multiple of 4)");
  CDN_CHECK(((((uintptr_t)input) | ((uintptr_t)output)) & 15) == 0, CDN_ERR_INVALID, "dw3x3_up2_f32: tensors must be 16-byte aligned");
  const long long total = (long long)B * C * h * (w / 4);
  CDN_CHECK(total < (1ll << 32) - 256, CDN_ERR_INVALID, "dw3x3_up2_f32: tensor too large");
  if (total == 0) return 0;
  dw3x3_up2_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, weight, bias, output, C, h, w, relu, (unsigned)total);
  CDN_LAUNCH_CHECK("dw3x3_up2_f32_kernel");
  return 0;
}

extern "C" int cdn_dw3x3_f32(const float* input, const float* weight, const float* bias, float* output, int B, int C, int H, int W,
                             int stride, int relu, cdn_stream_t stream) {
  CDN_CHECK(input && weight && output && C >= 1 && (stride == 1 || stride == 2), CDN_ERR_INVALID, "dw3x3_f32: bad arguments");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = (long long)B * C * Ho * Wo;
  if (total == 0) return 0;
  if (stride == 1 && W % 4 == 0 && H % 4 == 0 && ((((uintptr_t)input) | ((uintptr_t)output)) & 15) == 0 && total / 16 < (1ll << 32) - 256) {
    const unsigned total16 = (unsigned)(total / 16);
    const int W4 = W / 4;
    if (W4 <= 32 && 32 % W4 == 0) dw3x3_f32_s1r4_kernel<true><<<(total16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(input, weight, bias, output, C, H, W, relu, total16);
    else dw3x3_f32_s1r4_kernel<false><<<(total16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(input, weight, bias, output, C, H, W, relu, total16);
    CDN_LAUNCH_CHECK("dw3x3_f32_s1r4_kernel");
    return 0;
  }
  if (stride == 1 && W % 4 == 0 && ((((uintptr_t)input) | ((uintptr_t)output)) & 15) == 0) {
    const long long total4 = total / 4;
    dw3x3_f32_s1v4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, weight, bias, output, C, H, W, relu, total4);
    CDN_LAUNCH_CHECK("dw3x3_f32_s1v4_kernel");
    return 0;
  }
  if (stride == 2 && W % 8 == 0 && ((((uintptr_t)input) | ((uintptr_t)output)) & 15) == 0 && total / 4 < (1ll << 32) - 256) {
    const unsigned total4 = (unsigned)(total / 4);
    dw3x3_f32_s2v4_kernel<<<(total4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(input, weight, bias, output, C, H, W, Ho, Wo, relu, total4);
    CDN_LAUNCH_CHECK("dw3x3_f32_s2v4_kernel");
    return 0;
  }
  dw3x3_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, weight, bias, output, C, H, W, Ho, Wo, stride, relu, total);
  CDN_LAUNCH_CHECK("dw3x3_f32_kernel");
  return 0;
}

// ---- 1x1 conv on channel slices with a strided output channel map (split / cat / channel_shuffle folded) --------------
//   out[b][out_coff + co*out_cstride][p] = act(bias[co] + sum_c W[co][c] * in[b][in_coff + c][p])
#define PWS_CO 8
#define PWS_PX 4
// One thread = PWS_PX consecutive pixels x PWS_CO output channels (register tile 4 x 8): every weight fetched from shared
// memory feeds 4 FMAs and every activation 8, so the loop is FMA-bound instead of LDS-bound.
__global__ void __launch_bounds__(128) pw_slice_f32_kernel(const float* __restrict__ in, int in_ctotal, int in_coff, int C,
                                                           const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                                           int out_ctotal, int out_coff, int out_cstride, int Co, int relu, int ppi,
                                                           long long total_groups) {
  extern __shared__ float sw[];                 // [C][PWS_CO] (channel-major: one LDS.128 x2 per input channel)
  const int co0 = blockIdx.y * PWS_CO;
  for (int i = threadIdx.x; i < PWS_CO * C; i += blockDim.x) {
    const int c = i / PWS_CO, r = i - c * PWS_CO;
    sw[i] = (co0 + r < Co) ? w[(size_t)(co0 + r) * C + c] : 0.f;
  }
  __syncthreads();
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // group of PWS_PX pixels inside one image
  if (gi >= total_groups) return;
  const int gpi = ppi / PWS_PX;                 // host guarantees ppi % PWS_PX == 0
  const long long b = gi / gpi; const int pi = (int)(gi - b * gpi) * PWS_PX;
  const float* x = in + ((size_t)b * in_ctotal + in_coff) * ppi + pi;
  // two-level accumulation: fp32 FMA over blocks of 32 input channels, block sums added in fp64 -- a plain fp32 chain
  // over K = 1024 loses ~1e-5 relative per layer, which the float path's accuracy contract cannot afford
  double tot[PWS_PX][PWS_CO];
#pragma unroll
  for (int q = 0; q < PWS_PX; ++q)
#pragma unroll
    for (int r = 0; r < PWS_CO; ++r) tot[q][r] = (bias && co0 + r < Co) ? (double)bias[co0 + r] : 0.0;
  for (int c0 = 0; c0 < C; c0 += 32) {
    float acc[PWS_PX][PWS_CO];
#pragma unroll
    for (int q = 0; q < PWS_PX; ++q)
#pragma unroll
      for (int r = 0; r < PWS_CO; ++r) acc[q][r] = 0.f;
    const int c1 = min(c0 + 32, C);
    for (int c = c0; c < c1; ++c) {
      const float4 v = __ldg((const float4*)(x + (size_t)c * ppi));
      const float4 wa = *(const float4*)(sw + c * PWS_CO), wb = *(const float4*)(sw + c * PWS_CO + 4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int q = 0; q < PWS_PX; ++q)
#pragma unroll
        for (int r = 0; r < PWS_CO; ++r) acc[q][r] = fmaf(ww[r], vv[q], acc[q][r]);
    }
#pragma unroll
    for (int q = 0; q < PWS_PX; ++q)
#pragma unroll
      for (int r = 0; r < PWS_CO; ++r) tot[q][r] += (double)acc[q][r];
  }
#pragma unroll
  for (int r = 0; r < PWS_CO; ++r)
    if (co0 + r < Co) {
      float4 o;
      o.x = (float)tot[0][r]; o.y = (float)tot[1][r]; o.z = (float)tot[2][r]; o.w = (float)tot[3][r];
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      *(float4*)(out + ((size_t)b * out_ctotal + out_coff + (size_t)(co0 + r) * out_cstride) * ppi + pi) = o;
    }
}

// ---- the same 1x1 conv, shared-memory tiled (the path taken whenever pixels_per_image % 128 == 0) -------------------
// CTA tile = 64 output channels x 128 pixels of one image, K in chunks of 16 input channels staged in shared memory
// (weights transposed to [k][co], activations [k][pixel]); a thread owns 8 output channels x 4 consecutive pixels.  All
// lanes of a warp share their 8 output channels, so the weight reads are broadcasts and the activation read is one
// contiguous 512-byte LDS.128: 3 shared-memory instructions feed 32 FMAs, and every activation is fetched from L2 once
// per 64 output channels instead of once per 8 (the register-tile kernel above is bound by those re-reads).
// Arithmetic is the one of the kernel above, bit for bit: fp32 FMA chains over blocks of 32 input channels in ascending
// order, block sums added in fp64.
#define PWT_CO 64
#define PWT_PX 128
#define PWT_K 16
template <bool F2>
__global__ void __launch_bounds__(256, 2) pw_tile_f32_kernel(const float* __restrict__ in, int in_ctotal, int in_coff, int C,
                                                          const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                                          int out_ctotal, int out_coff, int out_cstride, int Co, int relu, int ppi) {
  __shared__ __align__(16) float sA[2][PWT_K][PWT_CO];          // weights, [k][co]
  __shared__ __align__(16) float sB[2][PWT_K][(F2 ? 2 : 1) * PWT_PX];   // activations, [k][pixel]; F2: every value stored twice = the (v, v) operand of FFMA2
  const int tiles_per_img = ppi / PWT_PX;
  const int b = blockIdx.x / tiles_per_img, px0 = (blockIdx.x - b * tiles_per_img) * PWT_PX;
  const int co0 = blockIdx.y * PWT_CO;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 4 pixels tx*4.., 8 output channels ty*8..
  const float* x = in + ((size_t)b * in_ctotal + in_coff) * ppi + px0;
  // loader roles: weights -- thread -> (co row r = tid / 4, four consecutive k); activations -- two float4 per thread
  const int a_r = threadIdx.x >> 2, a_k = (threadIdx.x & 3) * 4;
  const int b_k = threadIdx.x >> 5, b_p = (threadIdx.x & 31) * 4;          // rows b_k and b_k + 8
  float ra[4]; float4 rb0, rb1;
  auto gload = [&](int c0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + a_k + i;
      ra[i] = (co0 + a_r < Co && c < C) ? __ldg(w + (size_t)(co0 + a_r) * C + c) : 0.f;
    }
    rb0 = (c0 + b_k < C) ? __ldg((const float4*)(x + (size_t)(c0 + b_k) * ppi + b_p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    rb1 = (c0 + b_k + 8 < C) ? __ldg((const float4*)(x + (size_t)(c0 + b_k + 8) * ppi + b_p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sA[buf][a_k + i][a_r] = ra[i];
    if constexpr (F2) {
      *(float4*)&sB[buf][b_k][2 * b_p] = make_float4(rb0.x, rb0.x, rb0.y, rb0.y);
      *(float4*)&sB[buf][b_k][2 * b_p + 4] = make_float4(rb0.z, rb0.z, rb0.w, rb0.w);
      *(float4*)&sB[buf][b_k + 8][2 * b_p] = make_float4(rb1.x, rb1.x, rb1.y, rb1.y);
      *(float4*)&sB[buf][b_k + 8][2 * b_p + 4] = make_float4(rb1.z, rb1.z, rb1.w, rb1.w);
    } else {
      *(float4*)&sB[buf][b_k][b_p] = rb0;
      *(float4*)&sB[buf][b_k + 8][b_p] = rb1;
    }
  };
  double tot[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const double bv = (bias && co0 + ty * 8 + r < Co) ? (double)bias[co0 + ty * 8 + r] : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) tot[r][q] = bv;
  }
  // F2 (debug bit 30, A/B only): two output channels per instruction with Blackwell's packed fp32 FMA -- per lane the same
  // operations in the same order as the scalar form, identical bits, half the FMA issue slots; acc2[rp][q] = (channel 2 rp,
  // channel 2 rp + 1) of pixel q.  Measured SLOWER (60.2 vs 42.5 ms per 128-image forward of the 2x COCO model): the (v, v)
  // operand doubles the activation tile's shared-memory reads and the kernel is bound there, not on FMA issue.
  unsigned long long acc2[4][4];
  float acc[8][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc2[r][q] = 0ull; acc[2 * r][q] = acc[2 * r + 1][q] = 0.f; }
  const int nchunks = (C + PWT_K - 1) / PWT_K;
  gload(0); sstore(0);
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) gload((ch + 1) * PWT_K);              // next chunk's global loads fly during this chunk's FMAs
#pragma unroll
    for (int k = 0; k < PWT_K; ++k) {
      if constexpr (F2) {
        const ulonglong2 wa = *(const ulonglong2*)&sA[buf][k][ty * 8], wb = *(const ulonglong2*)&sA[buf][k][ty * 8 + 4];
        const ulonglong2 va = *(const ulonglong2*)&sB[buf][k][tx * 8], vb = *(const ulonglong2*)&sB[buf][k][tx * 8 + 4];
        const unsigned long long ww[4] = {wa.x, wa.y, wb.x, wb.y};
        const unsigned long long vv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][q]) : "l"(ww[r]), "l"(vv[q]));
      } else {
        const float4 wa = *(const float4*)&sA[buf][k][ty * 8], wb = *(const float4*)&sA[buf][k][ty * 8 + 4];
        const float4 v = *(const float4*)&sB[buf][k][tx * 4];
        const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(ww[r], vv[q], acc[r][q]);
      }
    }
    if ((ch & 1) == 1 || ch + 1 == nchunks) {                   // a block of 32 input channels is complete
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if constexpr (F2) {
            tot[2 * r][q] += (double)__uint_as_float((uint32_t)acc2[r][q]);
            tot[2 * r + 1][q] += (double)__uint_as_float((uint32_t)(acc2[r][q] >> 32));
            acc2[r][q] = 0ull;
          } else {
            tot[2 * r][q] += (double)acc[2 * r][q];
            tot[2 * r + 1][q] += (double)acc[2 * r + 1][q];
            acc[2 * r][q] = acc[2 * r + 1][q] = 0.f;
          }
        }
    }
    if (ch + 1 < nchunks) sstore(buf ^ 1);                      // the other buffer was last read in the previous iteration
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int co = co0 + ty * 8 + r;
    if (co < Co) {
      float4 o;
      o.x = (float)tot[r][0]; o.y = (float)tot[r][1]; o.z = (float)tot[r][2]; o.w = (float)tot[r][3];
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      *(float4*)(out + ((size_t)b * out_ctotal + out_coff + (size_t)co * out_cstride) * ppi + px0 + tx * 4) = o;
    }
  }
}

extern "C" int cdn_pw_slice_f32(const float* input, int in_ctotal, int in_coff, int C, const float* weight, const float* bias,
                                float* output, int out_ctotal, int out_coff, int out_cstride, int Co, int relu, int B,
                                int pixels_per_image, cdn_stream_t stream) {
  CDN_CHECK(input && weight && output && C >= 1 && Co >= 1 && in_coff >= 0 && in_coff + C <= in_ctotal && out_cstride >= 1 &&
            out_coff >= 0 && out_coff + (Co - 1) * out_cstride < out_ctotal, CDN_ERR_INVALID, "pw_slice_f32: channel slice out of range");
  CDN_CHECK((size_t)PWS_CO * C * sizeof(float) <= 160 * 1024, CDN_ERR_INVALID, "pw_slice_f32: C=%d too large", C);
  CDN_CHECK(pixels_per_image % PWS_PX == 0 && (((uintptr_t)input | (uintptr_t)output) & 15) == 0, CDN_ERR_INVALID,
            "pw_slice_f32: pixels per image must be a multiple of %d and the tensors 16-byte aligned", PWS_PX);
  const long long total = (long long)B * (pixels_per_image / PWS_PX);
  if (total == 0) return 0;
  if (pixels_per_image % PWT_PX == 0 && (long long)B * (pixels_per_image / PWT_PX) < (1ll << 31)) {
    dim3 grid((unsigned)(B * (pixels_per_image / PWT_PX)), (unsigned)((Co + PWT_CO - 1) / PWT_CO));
    if (!(g_cdn_debug_flags & (1u << 30)))
      pw_tile_f32_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(input, in_ctotal, in_coff, C, weight, bias, output, out_ctotal,
                                                                       out_coff, out_cstride, Co, relu, pixels_per_image);
    else
      pw_tile_f32_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(input, in_ctotal, in_coff, C, weight, bias, output, out_ctotal,
                                                                      out_coff, out_cstride, Co, relu, pixels_per_image);
    CDN_LAUNCH_CHECK("pw_tile_f32_kernel");
    return 0;
  }
  static bool attr_set[64] = {};             // function attributes are per DEVICE
  if (cdn_first_on_device(attr_set)) CDN_CUDA(cudaFuncSetAttribute(pw_slice_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  dim3 grid((unsigned)((total + 127) / 128), (unsigned)((Co + PWS_CO - 1) / PWS_CO));
  pw_slice_f32_kernel<<<grid, 128, (size_t)PWS_CO * C * sizeof(float), (cudaStream_t)stream>>>(
      input, in_ctotal, in_coff, C, weight, bias, output, out_ctotal, out_coff, out_cstride, Co, relu, pixels_per_image, total);
  CDN_LAUNCH_CHECK("pw_slice_f32_kernel");
  return 0;
}

// ---- channel copy with the same output map (the pass-through half of a unit), max-pool 3/2/1, nearest x2 upsample -------
__global__ void copy_channels_f32_kernel(const float* __restrict__ in, int in_ctotal, int in_coff, float* __restrict__ out, int out_ctotal,
                                         int out_coff, int out_cstride, int n, int ppi, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pi = (int)(idx % ppi); long long t = idx / ppi; const int c = (int)(t % n); const long long b = t / n;
  out[((size_t)b * out_ctotal + out_coff + (size_t)c * out_cstride) * ppi + pi] = in[((size_t)b * in_ctotal + in_coff + c) * ppi + pi];
}
// the same with planes that are a multiple of 4 floats and 16-byte aligned tensors: grid = (plane quarter-blocks, channel, image),
// 16 bytes per thread and no index division (the scalar form spent its time in two 64-bit divisions per element)
__global__ void __launch_bounds__(256) copy_channels_f32_v4_kernel(const float4* __restrict__ in, int in_ctotal, int in_coff, float4* __restrict__ out,
                                                                   int out_ctotal, int out_coff, int out_cstride, int ppi4) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float4* src = in + ((size_t)b * in_ctotal + in_coff + c) * ppi4;
  float4* dst = out + ((size_t)b * out_ctotal + out_coff + (size_t)c * out_cstride) * ppi4;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < ppi4; i += gridDim.x * 256) dst[i] = __ldg(src + i);
}
extern "C" int cdn_copy_channels_f32(const float* input, int in_ctotal, int in_coff, float* output, int out_ctotal, int out_coff,
                                     int out_cstride, int n, int B, int pixels_per_image, cdn_stream_t stream) {
  CDN_CHECK(input && output && n >= 1 && in_coff >= 0 && in_coff + n <= in_ctotal && out_coff >= 0 &&
            out_coff + (n - 1) * out_cstride < out_ctotal, CDN_ERR_INVALID, "copy_channels_f32: slice out of range");
  const long long total = (long long)B * n * pixels_per_image;
  if (total == 0) return 0;
  if (pixels_per_image % 4 == 0 && (((uintptr_t)input | (uintptr_t)output) & 15) == 0 && n <= 65535 && B <= 65535) {
    const int ppi4 = pixels_per_image / 4;
    dim3 grid((unsigned)std::min((ppi4 + 255) / 256, 8), (unsigned)n, (unsigned)B);
    copy_channels_f32_v4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)input, in_ctotal, in_coff, (float4*)output, out_ctotal,
                                                                       out_coff, out_cstride, ppi4);
    CDN_LAUNCH_CHECK("copy_channels_f32_v4_kernel");
    return 0;
  }
  copy_channels_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, in_ctotal, in_coff, output, out_ctotal,
                                                                                             out_coff, out_cstride, n, pixels_per_image, total);
  CDN_LAUNCH_CHECK("copy_channels_f32_kernel");
  return 0;
}

__global__ void maxpool3s2_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo); long long t = idx / Wo; const int ho = (int)(t % Ho); t /= Ho;
  const float* plane = in + (size_t)t * H * W;
  float m = -3.402823466e38f;
  for (int i = 0; i < 3; ++i) {
    const int y = 2 * ho - 1 + i; if ((unsigned)y >= (unsigned)H) continue;
    for (int j = 0; j < 3; ++j) { const int x = 2 * wo - 1 + j; if ((unsigned)x < (unsigned)W) m = fmaxf(m, plane[(size_t)y * W + x]); }
  }
  out[idx] = m;
}
extern "C" int cdn_maxpool3s2_f32(const float* input, float* output, int planes, int H, int W, cdn_stream_t stream) {
  CDN_CHECK(input && output && planes >= 0 && H >= 1 && W >= 1, CDN_ERR_INVALID, "maxpool3s2_f32: bad arguments");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)planes * Ho * Wo;
  if (total == 0) return 0;
  maxpool3s2_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, output, H, W, Ho, Wo, total);
  CDN_LAUNCH_CHECK("maxpool3s2_f32_kernel");
  return 0;
}

__global__ void upsample2x_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int W2 = 2 * W, H2 = 2 * H;
  const int x = (int)(idx % W2); long long t = idx / W2; const int y = (int)(t % H2); t /= H2;
  out[idx] = in[((size_t)t * H + (y >> 1)) * W + (x >> 1)];
}
extern "C" int cdn_upsample2x_f32(const float* input, float* output, int planes, int H, int W, cdn_stream_t stream) {
  CDN_CHECK(input && output && planes >= 0 && H >= 1 && W >= 1, CDN_ERR_INVALID, "upsample2x_f32: bad arguments");
  const long long total = (long long)planes * 4 * H * W;
  if (total == 0) return 0;
  upsample2x_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(input, output, H, W, total);
  CDN_LAUNCH_CHECK("upsample2x_f32_kernel");
  return 0;
}
