// The steps either side of the path on the device:
//   * cdn_warp_affine_u8  -- the image transform of BaseDetector.pre_process (lib/detectors/base_detector.py:48-76):
//     cv2.warpAffine(resized, trans_input, (inp_w, inp_h), flags=cv2.INTER_LINEAR) on the raw uint8 HWC image, bit for bit
//     (OpenCV's fixed-point scheme: coordinates in 1/1024 pixel rounded to 1/32, 15-bit bilinear weights, constant border 0),
//     optionally with the mirrored copy --flip_test appends (:69-70).  At test scale 1 the cv2.resize in front of it is the
//     identity, so the raw frame goes to the device as it is and the normalisation happens in the stem kernel's 3 x 256 table.
//   * cdn_ctdet_group_by_class -- the per-class grouping of ctdet_post_process (lib/utils/post_process.py:86-103): detections
//     [B][K][6] reordered by class (stable, i.e. by descending score inside a class) + per-class counts, so the host only
//     slices views.
#include "common.cuh"

struct WarpParams { const uint8_t* src; uint8_t* dst; int H, W, dh, dw, flip; double m[6]; };

__global__ void warp_affine_u8_kernel(const WarpParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.dw) return;
  // dst -> src map in 1/1024 pixel, as OpenCV forms it: per-column and per-row terms are rounded SEPARATELY, then summed
  const int adelta = __double2int_rn(p.m[0] * x * 1024.0), bdelta = __double2int_rn(p.m[3] * x * 1024.0);
  const int X0 = __double2int_rn((p.m[1] * y + p.m[2]) * 1024.0) + 16, Y0 = __double2int_rn((p.m[4] * y + p.m[5]) * 1024.0) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
  const int fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  int acc[3] = {0, 0, 0};
  auto tap = [&](int yy, int xx, int w) {
    if ((unsigned)yy < (unsigned)p.H && (unsigned)xx < (unsigned)p.W) {
      const uint8_t* s = p.src + ((size_t)yy * p.W + xx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += (int)__ldg(s + c) * w;
    }
  };
  tap(sy, sx, w00); tap(sy, sx + 1, w01); tap(sy + 1, sx, w10); tap(sy + 1, sx + 1, w11);
  uint8_t v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = (uint8_t)min(max((acc[c] + (1 << 14)) >> 15, 0), 255);
  uint8_t* d = p.dst + ((size_t)y * p.dw + x) * 3;
  d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
  if (p.flip) {
    uint8_t* f = p.dst + (size_t)p.dh * p.dw * 3 + ((size_t)y * p.dw + (p.dw - 1 - x)) * 3;
    f[0] = v[0]; f[1] = v[1]; f[2] = v[2];
  }
}

extern "C" int cdn_warp_affine_u8(const uint8_t* d_src, int H, int W, const double* M6, uint8_t* d_dst, int dst_h, int dst_w,
                                  int flip_copy, cdn_stream_t stream) {
  CDN_CHECK(d_src && d_dst && M6 && H > 0 && W > 0 && dst_h > 0 && dst_w > 0, CDN_ERR_INVALID, "warp_affine: bad arguments");
  // invert the 2 x 3 matrix exactly as cv::warpAffine does (imgwarp.cpp), in double
  double m[6] = {M6[0], M6[1], M6[2], M6[3], M6[4], M6[5]};
  double D = m[0] * m[4] - m[1] * m[3];
  D = D != 0 ? 1. / D : 0;
  const double A11 = m[4] * D, A22 = m[0] * D;
  m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22;
  const double b1 = -m[0] * m[2] - m[1] * m[5], b2 = -m[3] * m[2] - m[4] * m[5];
  m[2] = b1; m[5] = b2;
  WarpParams p; p.src = d_src; p.dst = d_dst; p.H = H; p.W = W; p.dh = dst_h; p.dw = dst_w; p.flip = flip_copy ? 1 : 0;
  for (int i = 0; i < 6; ++i) p.m[i] = m[i];
  dim3 grid((unsigned)((dst_w + 127) / 128), (unsigned)dst_h);
  warp_affine_u8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  CDN_LAUNCH_CHECK("warp_affine_u8_kernel");
  return 0;
}

// one CTA per image; K <= 1024.  out[b][rank] = dets[b][t] with rank = (detections of smaller classes) + (earlier ones of
// the same class); counts[b][c] = detections of class c.  Classes outside [0, num_classes) go last and are not counted.
__global__ void ctdet_group_by_class_kernel(const float* __restrict__ dets, int K, int num_classes, float* __restrict__ out,
                                            int* __restrict__ counts) {
  extern __shared__ int s_cls[];                 // [K] classes, then [num_classes + 1] histogram / offsets
  int* s_hist = s_cls + K;
  const int b = blockIdx.x, t = threadIdx.x;
  const float* d = dets + (size_t)b * K * 6;
  for (int c = t; c <= num_classes; c += blockDim.x) s_hist[c] = 0;
  __syncthreads();
  int cls = num_classes;
  if (t < K) {
    const float cf = d[t * 6 + 5];
    const int ci = (int)cf;
    if (cf >= 0.f && ci < num_classes) cls = ci;
    s_cls[t] = cls;
    atomicAdd(&s_hist[cls], 1);
  }
  __syncthreads();
  if (t < K) {
    int rank = 0;
    for (int c = 0; c < cls; ++c) rank += s_hist[c];
    for (int j = 0; j < t; ++j) rank += s_cls[j] == cls;
    float* o = out + ((size_t)b * K + rank) * 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = d[t * 6 + i];
  }
  for (int c = t; c < num_classes; c += blockDim.x) counts[(size_t)b * num_classes + c] = s_hist[c];
}

extern "C" int cdn_ctdet_group_by_class(const float* d_dets, int batch, int K, int num_classes, float* d_out, int32_t* d_counts,
                                        cdn_stream_t stream) {
  CDN_CHECK(d_dets && d_out && d_counts && batch >= 0 && K >= 1 && K <= 1024 && num_classes >= 1 && num_classes <= 4096, CDN_ERR_INVALID,
            "group_by_class: bad arguments (K <= 1024)");
  CDN_CHECK(d_dets != d_out, CDN_ERR_INVALID, "group_by_class: in-place operation is not supported");
  if (batch == 0) return 0;
  const int threads = (std::max(K, 32) + 31) / 32 * 32;
  ctdet_group_by_class_kernel<<<batch, threads, (size_t)(K + num_classes + 1) * sizeof(int), (cudaStream_t)stream>>>(
      d_dets, K, num_classes, d_out, d_counts);
  CDN_LAUNCH_CHECK("ctdet_group_by_class_kernel");
  return 0;
}
