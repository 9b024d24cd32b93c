// General deformable convolution forward, fp32 NCHW: the reference's native plug-in point
// (deform_conv_forward_cuda, lib/models/external/src/dcn_deform_conv_cuda.cpp:151-258; kernel
// dcn_deform_conv_cuda_kernel.cu:189-242, bilinear :83-114).  Gather and MAC are fused: there is no im2col
// buffer, no per-group GEMM and no transposed copy.  One thread per output element; lanes run along W so the
// offset reads and the output store are coalesced.
#include "common.cuh"

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
