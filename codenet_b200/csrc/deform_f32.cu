// General deformable convolution forward, fp32 NCHW: the reference's native plug-in point
// (deform_conv_forward_cuda, lib/models/external/src/dcn_deform_conv_cuda.cpp:151-258; kernel
// dcn_deform_conv_cuda_kernel.cu:189-242, bilinear :83-114).  Gather and MAC are fused: there is no im2col
// buffer, no per-group GEMM and no transposed copy.  One thread per output element; lanes run along W so the
// offset reads and the output store are coalesced.
#include "common.cuh"
#include <algorithm>

struct DefF32Params {
  const float* in; const float* w; const float* off; float* out;
  int B, C, H, W, Co, kH, kW, dH, dW, padH, padW, dilH, dilW, group, dg, Ho, Wo;
  long long total;
};

__device__ __forceinline__ float bilinear_ref(const float* img, int H, int W, float h, float w) {
  int hl = (int)floorf(h), wl = (int)floorf(w);
  int hh_ = hl + 1, wh_ = wl + 1;
  float lh = h - hl, lw = w - wl, hh = 1 - lh, hw = 1 - lw;
  float v1 = (hl >= 0 && wl >= 0) ? __ldg(img + hl * W + wl) : 0.f;
  float v2 = (hl >= 0 && wh_ <= W - 1) ? __ldg(img + hl * W + wh_) : 0.f;
  float v3 = (hh_ <= H - 1 && wl >= 0) ? __ldg(img + hh_ * W + wl) : 0.f;
  float v4 = (hh_ <= H - 1 && wh_ <= W - 1) ? __ldg(img + hh_ * W + wh_) : 0.f;
  return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

__global__ void deform_conv_f32_kernel(DefF32Params p) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total) return;
  int wo = (int)(idx % p.Wo); long long t = idx / p.Wo; int ho = (int)(t % p.Ho); t /= p.Ho;
  int co = (int)(t % p.Co); int b = (int)(t / p.Co);
  const int cpg = p.C / p.group, copg = p.Co / p.group, g = co / copg;
  const int cpdg = p.C / p.dg;
  const int KK = p.kH * p.kW;
  const size_t plane_o = (size_t)p.Ho * p.Wo;
  float acc = 0.f;
  for (int cl = 0; cl < cpg; ++cl) {
    const int c = g * cpg + cl;
    const float* img = p.in + ((size_t)b * p.C + c) * p.H * p.W;
    const float* offb = p.off + ((size_t)b * p.dg + c / cpdg) * 2 * KK * plane_o + (size_t)ho * p.Wo + wo;
    const float* wk = p.w + ((size_t)co * cpg + cl) * KK;
    for (int i = 0; i < p.kH; ++i)
      for (int j = 0; j < p.kW; ++j) {
        const int tap = i * p.kW + j;
        float oh = __ldg(offb + (size_t)(2 * tap) * plane_o), ow = __ldg(offb + (size_t)(2 * tap + 1) * plane_o);
        float h_im = (float)(ho * p.dH - p.padH + i * p.dilH) + oh;
        float w_im = (float)(wo * p.dW - p.padW + j * p.dilW) + ow;
        float val = 0.f;
        if (h_im > -1 && w_im > -1 && h_im < p.H && w_im < p.W) val = bilinear_ref(img, p.H, p.W, h_im, w_im);
        acc = fmaf(__ldg(wk + tap), val, acc);
      }
  }
  p.out[idx] = acc;
}

int deform_f32_launch(const DefF32Params& p, cudaStream_t st) {
  if (p.total == 0) return 0;
  deform_conv_f32_kernel<<<(unsigned)((p.total + 255) / 256), 256, 0, st>>>(p);
  CDN_LAUNCH_CHECK("deform_conv_f32_kernel");
  return 0;
}

extern "C" int cdn_deform_conv_forward_f32(const float* input, const float* weight, const float* offset, float* output,
                                           int B, int C, int H, int W, int Co, int kW, int kH, int dW, int dH,
                                           int padW, int padH, int dilW, int dilH, int group, int deformable_group,
                                           int im2col_step, cdn_stream_t stream) {
  (void)im2col_step;
  // shape_check, dcn_deform_conv_cuda.cpp:61-149
  CDN_CHECK(input && weight && offset && output, CDN_ERR_INVALID, "deform_conv: null tensor");
  CDN_CHECK(kW > 0 && kH > 0, CDN_ERR_INVALID, "kernel size should be greater than zero, but got kH: %d kW: %d", kH, kW);
  CDN_CHECK(dW > 0 && dH > 0, CDN_ERR_INVALID, "stride should be greater than zero, but got dH: %d dW: %d", dH, dW);
  CDN_CHECK(dilW > 0 && dilH > 0, CDN_ERR_INVALID, "dilation should be greater than 0, but got dilationH: %d dilationW: %d", dilH, dilW);
  CDN_CHECK(group > 0 && deformable_group > 0 && C % group == 0 && Co % group == 0, CDN_ERR_INVALID, "channels must divide groups");
  CDN_CHECK(C % deformable_group == 0, CDN_ERR_INVALID, "input channels must divide deformable group size");
  int Ho = (H + 2 * padH - (dilH * (kH - 1) + 1)) / dH + 1, Wo = (W + 2 * padW - (dilW * (kW - 1) + 1)) / dW + 1;
  CDN_CHECK(Ho >= 1 && Wo >= 1, CDN_ERR_INVALID, "Given input size: (%d x %d x %d). Calculated output size: (%d x %d x %d). Output size is too small",
            C, H, W, Co, Ho, Wo);
  CDN_CHECK(H >= kH && W >= kW, CDN_ERR_INVALID, "input image is smaller than kernel");
  DefF32Params p{input, weight, offset, output, B, C, H, W, Co, kH, kW, dH, dW, padH, padW, dilH, dilW, group,
                 deformable_group, Ho, Wo, (long long)B * Co * Ho * Wo};
  return deform_f32_launch(p, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------
// Fused co-designed deformable module, fp32 NCHW (row a1 of SURVEY.md 8(a)):
//   DeformConvWithOffsetScaleBoundPositive.forward, lib/models/external/modules/dcn_deform_conv.py:323-330 --
//   s = Hardtanh[-bound+1, bound](conv1x1_{C->1, stride}(x) + bias);  o = anchor * (s - 1);
//   y = deform_conv(x, o, W_dw[C,1,3,3], stride, pad 1, groups = C)
// in two launches: the scale scalar per output pixel (8 bytes per pixel of scratch), then the gather -- the 18-channel
// offset tensor, the im2col buffer and the per-channel GEMMs of the reference (dcn_deform_conv_cuda.cpp:196-245) never
// exist.  Lanes run along W (coalesced NCHW rows).
// ---------------------------------------------------------------------------------------------------------
struct DefDwF32Params {
  const float* in; const float* ws; const float* wdw; float* out;
  double* sd;                                   // [B][Ho*Wo] clamped scale per output pixel (phase A -> phase B)
  float bs, lo, hi;
  int B, C, H, W, Ho, Wo, stride;
  int up;                                       // 1: `in` is [B][C][H/2][W/2] and is read through a virtual nearest x2 upsampling
};

// Phase A: s = Hardtanh(w_scale . x + b_scale) per output pixel.  The scale scalar moves the sample positions, and the sampled
// features can vary by O(1) per pixel: an fp32 error in s is amplified into the output, so the C -> 1 reduction is kept in fp64.
// 32 pixels x 8 channel groups per CTA: a group sums channels g, g + 8, ... (coalesced 128-byte rows of the planes), the eight
// partial sums are added in a fixed order.
__global__ void __launch_bounds__(256) deform_scale_f32_kernel(DefDwF32Params p) {
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int npx = p.Ho * p.Wo;
  const int px = blockIdx.x * 32 + lane, b = blockIdx.y;
  double sd = 0.0;
  if (px < npx) {
    const int ho = px / p.Wo, wo = px - ho * p.Wo;
    const int Ws = p.W >> p.up;                                     // stored row length / plane of the (possibly low-resolution) input
    const size_t plane = (size_t)(p.H >> p.up) * Ws;
    const float* x = p.in + (size_t)b * p.C * plane + (size_t)((ho * p.stride) >> p.up) * Ws + ((wo * p.stride) >> p.up);   // conv_scale: kernel 1, padding 0
    // four independent chains per thread (the single chain left the loads waiting behind a dependent fp64 FMA each)
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = grp;
    for (; c + 24 < p.C; c += 32) {
      s0 = fma((double)__ldg(p.ws + c), (double)__ldg(x + (size_t)c * plane), s0);
      s1 = fma((double)__ldg(p.ws + c + 8), (double)__ldg(x + (size_t)(c + 8) * plane), s1);
      s2 = fma((double)__ldg(p.ws + c + 16), (double)__ldg(x + (size_t)(c + 16) * plane), s2);
      s3 = fma((double)__ldg(p.ws + c + 24), (double)__ldg(x + (size_t)(c + 24) * plane), s3);
    }
    for (; c < p.C; c += 8) s0 = fma((double)__ldg(p.ws + c), (double)__ldg(x + (size_t)c * plane), s0);
    sd = (s0 + s1) + (s2 + s3);
  }
  part[grp][lane] = sd;
  __syncthreads();
  if (grp == 0 && px < npx) {
    double t = (double)p.bs;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += part[g][lane];
    p.sd[(size_t)b * npx + px] = fmin(fmax(t, (double)p.lo), (double)p.hi);
  }
}

// Phase B: one thread = one output pixel x DDF_CG channels.  The 9 taps share their sample geometry across all channels: row /
// column floors and fractions are computed once per thread (the outer taps sit at (i-1)*(s-1) from the regular position, the
// centre row / column is integral); corners outside the image get weight 0 and a clamped address, which is the reference's
// "zero outside the image, per corner" (dcn_deform_conv_cuda_kernel.cu:83-114) without a select per load.  The grid runs over
// (pixel block, channel group, image): the 16 x 16 layer with 2153 channels gets 69 k CTAs where one thread per pixel looping
// over every channel (round 1) left the GPU with 220 threads per SM.
#define DDF_CG 16
__global__ void __launch_bounds__(128, 4) deform_gather_f32_kernel(DefDwF32Params p) {
  const int npx = p.Ho * p.Wo;
  const int px = blockIdx.x * 128 + threadIdx.x, b = blockIdx.z;
  if (px >= npx) return;
  const int ho = px / p.Wo, wo = px - ho * p.Wo;
  const int hc = ho * p.stride, wc = wo * p.stride;
  const double d = p.sd[(size_t)b * npx + px] - 1.0;
  // sample rows / columns of the three tap rows / columns: h_im = ho*stride - 1 + i + (i - 1)*d
  int ra[3], rb[3], ca[3], cb[3]; float wra[3], wrb[3], wca[3], wcb[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double him = (double)(hc - 1 + i) + (double)(i - 1) * d, wim = (double)(wc - 1 + i) + (double)(i - 1) * d;
    const double hf = floor(him), wf = floor(wim);
    const int y = (int)hf, x = (int)wf;
    const float lh = (float)(him - hf), lw = (float)(wim - wf);
    // outside the reference's range test (h_im > -1 && h_im < H): every corner is invalid or has weight zero already
    wra[i] = ((unsigned)y < (unsigned)p.H) ? 1.f - lh : 0.f;        wrb[i] = ((unsigned)(y + 1) < (unsigned)p.H) ? lh : 0.f;
    wca[i] = ((unsigned)x < (unsigned)p.W) ? 1.f - lw : 0.f;        wcb[i] = ((unsigned)(x + 1) < (unsigned)p.W) ? lw : 0.f;
    // addresses in the stored plane: with the virtual upsampling, sample (y, x) of the 2x map is element (y >> 1, x >> 1)
    ra[i] = (min(max(y, 0), p.H - 1) >> p.up) * (p.W >> p.up);  rb[i] = (min(max(y + 1, 0), p.H - 1) >> p.up) * (p.W >> p.up);
    ca[i] = min(max(x, 0), p.W - 1) >> p.up;                    cb[i] = min(max(x + 1, 0), p.W - 1) >> p.up;
  }
  const size_t plane = (size_t)(p.H >> p.up) * (p.W >> p.up);
  const int c_begin = blockIdx.y * DDF_CG, c_end = min(c_begin + DDF_CG, p.C);
  const float* img = p.in + ((size_t)b * p.C + c_begin) * plane;
  float* ob = p.out + ((size_t)b * p.C + c_begin) * npx + px;
  // The centre tap row / column sits at an integral position (its offset is 0 * d): its lower / right corners carry weight
  // exactly 0 and are not loaded -- 25 loads per channel instead of 36 (corner taps 4, edge taps 2, centre 1); a skipped term is
  // an exact +0 in the reference's sum.  The corner weights are per pixel, not per channel: hoisted out of the channel loop.
  float w4[3][3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      w4[i][j][0] = wra[i] * wca[j]; w4[i][j][1] = wra[i] * wcb[j]; w4[i][j][2] = wrb[i] * wca[j]; w4[i][j][3] = wrb[i] * wcb[j];
    }
  for (int c = c_begin; c < c_end; ++c, img += plane, ob += npx) {
    const float* wk = p.wdw + c * 9;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float* pa = img + ra[i];
      const float* pb = img + rb[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float val = w4[i][j][0] * __ldg(pa + ca[j]);
        if (j != 1) val += w4[i][j][1] * __ldg(pa + cb[j]);
        if (i != 1) val += w4[i][j][2] * __ldg(pb + ca[j]);
        if (i != 1 && j != 1) val += w4[i][j][3] * __ldg(pb + cb[j]);
        acc = fmaf(__ldg(wk + i * 3 + j), val, acc);
      }
    }
    *ob = acc;
  }
}

extern "C" size_t cdn_deform_dw_f32_ws_bytes(int B, int H, int W, int stride) {
  if (B < 0 || H < 1 || W < 1 || (stride != 1 && stride != 2)) return 0;
  const size_t Ho = (size_t)((H + 2 - 3) / stride + 1), Wo = (size_t)((W + 2 - 3) / stride + 1);
  return std::max<size_t>(B * Ho * Wo, 1) * sizeof(double);
}

static int deform_dw_f32_run(const float* input, const float* w_scale, float b_scale, int offset_bound, const float* w_dw,
                             float* output, int B, int C, int H, int W, int stride, int up, void* d_ws, size_t ws_bytes,
                             cdn_stream_t stream) {
  CDN_CHECK(input && w_scale && w_dw && output, CDN_ERR_INVALID, "deform_dw_f32: null tensor");
  CDN_CHECK(B >= 0 && B <= 65535 && C >= 1 && H >= 1 && W >= 1 && (stride == 1 || stride == 2) && offset_bound >= 1, CDN_ERR_INVALID,
            "deform_dw_f32: bad shape / stride / bound");
  CDN_CHECK(d_ws && ((uintptr_t)d_ws & 7) == 0 && ws_bytes >= cdn_deform_dw_f32_ws_bytes(B, H, W, stride), CDN_ERR_INVALID,
            "deform_dw_f32: workspace of %zu bytes (8-byte aligned) needed, got %zu", cdn_deform_dw_f32_ws_bytes(B, H, W, stride), ws_bytes);
  DefDwF32Params p;
  p.in = input; p.ws = w_scale; p.wdw = w_dw; p.out = output; p.bs = b_scale; p.sd = (double*)d_ws;
  p.lo = (float)(-offset_bound + 1); p.hi = (float)offset_bound;
  p.B = B; p.C = C; p.H = H; p.W = W; p.stride = stride; p.up = up;
  p.Ho = (H + 2 - 3) / stride + 1; p.Wo = (W + 2 - 3) / stride + 1;
  const int npx = p.Ho * p.Wo;
  if (B == 0) return 0;
  CDN_CHECK((C + DDF_CG - 1) / DDF_CG <= 65535, CDN_ERR_INVALID, "deform_dw_f32: too many channels");
  deform_scale_f32_kernel<<<dim3((unsigned)((npx + 31) / 32), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(p);
  CDN_LAUNCH_CHECK("deform_scale_f32_kernel");
  deform_gather_f32_kernel<<<dim3((unsigned)((npx + 127) / 128), (unsigned)((C + DDF_CG - 1) / DDF_CG), (unsigned)B), 128, 0, (cudaStream_t)stream>>>(p);
  CDN_LAUNCH_CHECK("deform_gather_f32_kernel");
  return 0;
}

extern "C" int cdn_deform_dw_f32_ws(const float* input, const float* w_scale, float b_scale, int offset_bound, const float* w_dw,
                                    float* output, int B, int C, int H, int W, int stride, void* d_ws, size_t ws_bytes,
                                    cdn_stream_t stream) {
  return deform_dw_f32_run(input, w_scale, b_scale, offset_bound, w_dw, output, B, C, H, W, stride, 0, d_ws, ws_bytes, stream);
}
// The module applied to the nearest x2 upsampling of `input` [B][C][h][w] (stride 1), which is never written: the block
// "ReLU -> nn.Upsample(2) -> DeformConvWithOffsetScaleBoundPositive" of the up path (shufflenetv2_dcn.py:286-300).  Samples of
// the 2h x 2w map are read as element (y >> 1, x >> 1); output [B][C][2h][2w]; workspace cdn_deform_dw_f32_ws_bytes(B, 2h, 2w, 1).
extern "C" int cdn_deform_dw_up2_f32_ws(const float* input, const float* w_scale, float b_scale, int offset_bound, const float* w_dw,
                                        float* output, int B, int C, int h, int w, void* d_ws, size_t ws_bytes, cdn_stream_t stream) {
  CDN_CHECK(h >= 1 && w >= 1 && h < (1 << 14) && w < (1 << 14), CDN_ERR_INVALID, "deform_dw_up2_f32: bad shape");
  return deform_dw_f32_run(input, w_scale, b_scale, offset_bound, w_dw, output, B, C, 2 * h, 2 * w, 1, 1, d_ws, ws_bytes, stream);
}

// The form without a workspace argument keeps one scratch buffer per device, grown on demand (a growth allocates and therefore
// must not happen inside a stream capture: call once eagerly first, or use cdn_deform_dw_f32_ws).
extern "C" int cdn_deform_dw_f32(const float* input, const float* w_scale, float b_scale, int offset_bound, const float* w_dw,
                                 float* output, int B, int C, int H, int W, int stride, cdn_stream_t stream) {
  static void* scratch[64] = {}; static size_t scratch_bytes[64] = {};
  int dev = 0;
  CDN_CUDA(cudaGetDevice(&dev));
  CDN_CHECK(dev >= 0 && dev < 64, CDN_ERR_INVALID, "deform_dw_f32: device index out of range");
  const size_t need = cdn_deform_dw_f32_ws_bytes(B, H, W, stride);
  CDN_CHECK(need > 0, CDN_ERR_INVALID, "deform_dw_f32: bad shape / stride");
  if (scratch_bytes[dev] < need) {
    CDN_CUDA(cudaDeviceSynchronize());
    if (scratch[dev]) cudaFree(scratch[dev]);
    scratch[dev] = nullptr; scratch_bytes[dev] = 0;
    CDN_CUDA(cudaMalloc(&scratch[dev], need));
    scratch_bytes[dev] = need;
  }
  return cdn_deform_dw_f32_ws(input, w_scale, b_scale, offset_bound, w_dw, output, B, C, H, W, stride, scratch[dev], scratch_bytes[dev], stream);
}

// ---------------------------------------------------------------------------------------------------------
// fp32 1x1 convolution, NCHW: out[b][co][p] = sum_c W[co][c] * in[b][c][p] (+ bias[co]).  Used by the fp32 module for its
// conv_channel (modules/dcn_deform_conv.py:311-317).  One thread = one pixel x PWF_CO output channels; the weight rows
// are broadcast from shared memory; pixels run along lanes (coalesced planes).  fp32 FMA accumulation (the 1e-4 contract
// of the float path rules TF32 out).
// ---------------------------------------------------------------------------------------------------------
#define PWF_CO 8
__global__ void __launch_bounds__(128) pw_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                     float* __restrict__ out, int C, int Co, int ppi, long long total_px) {
  extern __shared__ float sw[];                 // [PWF_CO][C]
  const int co0 = blockIdx.y * PWF_CO;
  for (int i = threadIdx.x; i < PWF_CO * C; i += blockDim.x) {
    const int r = i / C, c = i - r * C;
    sw[i] = (co0 + r < Co) ? w[(size_t)(co0 + r) * C + c] : 0.f;
  }
  __syncthreads();
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= total_px) return;
  const long long b = px / ppi; const int pi = (int)(px - b * ppi);
  const float* x = in + (size_t)b * C * ppi + pi;
  double tot[PWF_CO];                            // fp32 FMA over blocks of 32 channels, block sums in fp64
#pragma unroll
  for (int r = 0; r < PWF_CO; ++r) tot[r] = (bias && co0 + r < Co) ? (double)bias[co0 + r] : 0.0;
  for (int c0 = 0; c0 < C; c0 += 32) {
    float acc[PWF_CO];
#pragma unroll
    for (int r = 0; r < PWF_CO; ++r) acc[r] = 0.f;
    const int c1 = min(c0 + 32, C);
    for (int c = c0; c < c1; ++c) {
      const float v = __ldg(x + (size_t)c * ppi);
#pragma unroll
      for (int r = 0; r < PWF_CO; ++r) acc[r] = fmaf(sw[r * C + c], v, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < PWF_CO; ++r) tot[r] += (double)acc[r];
  }
#pragma unroll
  for (int r = 0; r < PWF_CO; ++r)
    if (co0 + r < Co) out[((size_t)b * Co + co0 + r) * ppi + pi] = (float)tot[r];
}

extern "C" int cdn_pw_f32(const float* input, const float* weight, const float* bias, float* output, int B, int C, int Co,
                          int pixels_per_image, cdn_stream_t stream) {
  CDN_CHECK(input && weight && output && B >= 0 && C >= 1 && Co >= 1 && pixels_per_image >= 1, CDN_ERR_INVALID, "pw_f32: bad arguments");
  CDN_CHECK((size_t)PWF_CO * C * sizeof(float) <= 160 * 1024, CDN_ERR_INVALID, "pw_f32: C=%d too large for the shared weight tile", C);
  const long long total = (long long)B * pixels_per_image;
  if (total == 0) return 0;
  const size_t smem = (size_t)PWF_CO * C * sizeof(float);
  static bool attr_set[64] = {};             // function attributes are per DEVICE
  if (cdn_first_on_device(attr_set)) CDN_CUDA(cudaFuncSetAttribute(pw_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  dim3 grid((unsigned)((total + 127) / 128), (unsigned)((Co + PWF_CO - 1) / PWF_CO));
  pw_f32_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(input, weight, bias, output, C, Co, pixels_per_image, total);
  CDN_LAUNCH_CHECK("pw_f32_kernel");
  return 0;
}
