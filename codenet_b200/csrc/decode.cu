// ctdet decode: 3x3 max-pool NMS + top-K + gather + box assembly in ONE launch, one CTA per image.
// Restates lib/models/decode.py:474-505 (_nms :10-16, _topk :110-126, gathers lib/models/utils.py:14-29) on the
// heat-map LOGITS (sigmoid is monotone, so peaks and order are unchanged; the score written is sigmoid(logit)).
// The reference's two-stage top-K (per class, then over classes) selects the same set as one global top-K.
// Order / ties: logit descending, then class ascending, then spatial index ascending (torch.topk leaves ties open).
//
// Phase 1  every thread scans its share of the map, tests "equal to the max of the 3x3 neighbourhood" and appends
//          peaks as 64-bit composites (ordered key << 32 | ~flat_index) to a per-image list (warp-aggregated).
// Phase 2  radix select (11-bit digits, warp-shuffle scans) of the K-th largest composite over the peak list.
// Phase 3  the K survivors are bitonic-sorted in shared memory and turned into boxes.
#include "layers.cuh"

#define DEC_THREADS 1024
#define DEC_BINS 2048
#define DEC_MAXK 1024

__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

struct DecParams {
  const float* hm; const float* wh; const float* reg;
  long long hm_is, wh_is, reg_is;            // per-image strides (elements)
  int cat, H, W, K;
  int is_prob;                               // heat map holds probabilities (reference API) instead of logits
  unsigned long long* list;                  // [batch][cat*H*W] peak composites
  unsigned int* counts;                      // [batch] peaks found by ctdet_peaks_kernel (reset by the select kernel)
  float* dets; int32_t* inds;
};

// Phase 1 as its own grid-wide kernel: one thread per heat-map element (all images), warp-aggregated append to the
// image's peak list.  The list order is arbitrary; the selection below orders by the (unique) composite value.
__global__ void __launch_bounds__(256) ctdet_peaks_kernel(DecParams p, int batch) {
  const int HW = p.H * p.W;
  const unsigned n = (unsigned)p.cat * HW;                   // elements per image (< 2^31, checked on the host)
  const unsigned e = blockIdx.x * 256u + threadIdx.x;        // grid.y = image
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  bool peak = false; float v = 0.f;
  if (e < n) {
    const int sp = (int)(e % (unsigned)HW); const int y = sp / p.W, x = sp - y * p.W;
    const float* plane = p.hm + (size_t)b * p.hm_is + (e - sp);
    v = __ldg(plane + sp);
    peak = (v == v);                                         // NaN never is a peak
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy; if ((unsigned)yy >= (unsigned)p.H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = x + dx; if ((dx | dy) == 0 || (unsigned)xx >= (unsigned)p.W) continue;
        if (__ldg(plane + yy * p.W + xx) > v) peak = false;
      }
    }
  }
  // block-aggregated append: one global atomic per block (per-warp atomics on 256 counters serialise in L2)
  __shared__ unsigned s_wcnt[8], s_base;
  const unsigned m = __ballot_sync(0xffffffffu, peak);
  const int warp = threadIdx.x >> 5;
  if (lane == 0) s_wcnt[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const unsigned c = s_wcnt[w]; s_wcnt[w] = tot; tot += c; }
    s_base = tot ? atomicAdd(p.counts + b, tot) : 0u;
  }
  __syncthreads();
  if (peak) p.list[(size_t)b * n + s_base + s_wcnt[warp] + __popc(m & ((1u << lane) - 1u))] =
      ((unsigned long long)f2key(v) << 32) | (unsigned long long)(0xffffffffu - e);
}

// Row-strip variant of the peak kernel (W % 4 == 0): a thread owns 4 horizontally adjacent elements and walks down a
// strip of DEC_STRIP rows, keeping the horizontal 3-max of the previous two rows in registers, so every heat-map value
// is loaded ~1.4 times (one float4 + two halo scalars per row) instead of 9.
#define DEC_STRIP 16
__global__ void __launch_bounds__(256) ctdet_peaks_rows_kernel(DecParams p, int strips_per_plane) {
  pdl_launch_dependents();
  const int HW = p.H * p.W, W4 = p.W >> 2;
  const unsigned n = (unsigned)p.cat * HW;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned tid = blockIdx.x * 256u + threadIdx.x;
  const unsigned per_plane = (unsigned)strips_per_plane * W4;
  const unsigned plane_i = tid / per_plane, rem = tid - plane_i * per_plane;
  const bool live = plane_i < (unsigned)p.cat;
  const int strip = (int)(rem / W4), x0 = (int)(rem % W4) * 4;
  const float* plane = p.hm + (size_t)b * p.hm_is + (size_t)(live ? plane_i : 0) * HW;
  const int y0 = strip * DEC_STRIP, y1 = min(y0 + DEC_STRIP, p.H);
  const float NEG = -3.402823466e38f;
  // row record: the 4 values and the horizontal 3-max at each of the 4 positions
  auto load_row = [&](int y, float (&v)[4], float (&hm3)[4]) {
    if ((unsigned)y >= (unsigned)p.H || !live) { v[0] = v[1] = v[2] = v[3] = NEG; hm3[0] = hm3[1] = hm3[2] = hm3[3] = NEG; return; }
    const float* r = plane + (size_t)y * p.W + x0;
    const float4 q = __ldg((const float4*)r);
    const float l = x0 > 0 ? __ldg(r - 1) : NEG, rr = x0 + 4 < p.W ? __ldg(r + 4) : NEG;
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    hm3[0] = fmaxf(fmaxf(l, q.x), q.y); hm3[1] = fmaxf(fmaxf(q.x, q.y), q.z);
    hm3[2] = fmaxf(fmaxf(q.y, q.z), q.w); hm3[3] = fmaxf(fmaxf(q.z, q.w), rr);
  };
  float vm[4], hmm[4], vc[4], hmc[4], vp[4], hmp[4];
  load_row(y0 - 1, vm, hmm); load_row(y0, vc, hmc);
  __shared__ unsigned s_wcnt[8], s_base;
  unsigned long long mask = 0ull;                // bit (y - y0) * 4 + i: element is a peak (4 x DEC_STRIP = 64 elements)
  for (int y = y0; y < y1; ++y) {
    load_row(y + 1, vp, hmp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float m = fmaxf(fmaxf(hmm[i], hmc[i]), hmp[i]);
      if (live && vc[i] == m) mask |= 1ull << ((y - y0) * 4 + i);     // equal to the 3x3 maximum (NaN compares false)
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { vm[i] = vc[i]; hmm[i] = hmc[i]; vc[i] = vp[i]; hmc[i] = hmp[i]; }
  }
  const unsigned cnt = (unsigned)__popcll(mask);
  // block-aggregated append
  unsigned incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_wcnt[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const unsigned c = s_wcnt[w]; s_wcnt[w] = tot; tot += c; }
    s_base = tot ? atomicAdd(p.counts + b, tot) : 0u;
  }
  __syncthreads();
  unsigned long long* dst = p.list + (size_t)b * n + s_base + s_wcnt[warp] + (incl - cnt);
  while (mask) {
    const int bit = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const int y = y0 + (bit >> 2), x = x0 + (bit & 3);
    const float v = __ldg(plane + (size_t)y * p.W + x);            // L1 / L2 hit: the strip was just read
    const unsigned e = plane_i * HW + (unsigned)y * p.W + x;
    *dst++ = ((unsigned long long)f2key(v) << 32) | (unsigned long long)(0xffffffffu - e);
  }
}

__global__ void __launch_bounds__(DEC_THREADS) ctdet_decode_kernel(DecParams p) {
  __shared__ unsigned int hist[DEC_BINS];
  __shared__ unsigned long long sel[DEC_MAXK];
  __shared__ unsigned int s_count, s_nsel;
  __shared__ unsigned long long s_prefix, s_mask;
  __shared__ unsigned int s_krem, s_done;
  __shared__ unsigned int warp_sums[32];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = p.H * p.W;
  const long long n = (long long)p.cat * HW;
  const float* hm = p.hm + (size_t)b * p.hm_is;
  unsigned long long* list = p.list + (size_t)b * n;

  if (tid == 0) { s_count = p.counts[b]; s_nsel = 0; }
  __syncthreads();
  const unsigned int count = s_count;
  const unsigned int Ksel = min((unsigned)p.K, count);

  // ---- phase 2: radix select of the Ksel-th largest composite ------------------------------------------
  if (tid == 0) { s_prefix = 0ull; s_mask = 0ull; s_krem = Ksel; s_done = (Ksel == 0 || Ksel == count) ? 1u : 0u; }
  __syncthreads();
  for (int shift = 53; !s_done; shift = max(shift - 11, 0)) {
    const int width = (shift == 0) ? 9 : 11;   // 53,42,31,20,9 -> 11 bits each, then the last 9 bits
    const unsigned nb = 1u << width;
    for (int i = tid; i < DEC_BINS; i += DEC_THREADS) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix, mask = s_mask;
    for (unsigned i = tid; i < count; i += DEC_THREADS) {
      unsigned long long c = list[i];
      if ((c & mask) == prefix) atomicAdd(&hist[(unsigned)(c >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    // suffix counts from the top digit: each thread owns 2 bins (DEC_BINS / DEC_THREADS)
    {
      unsigned d0 = nb - 1 - 2 * tid, d1 = nb - 2 - 2 * tid;           // descending digit order
      unsigned c0 = (2 * tid < nb) ? hist[d0] : 0u, c1 = (2 * tid + 1 < nb) ? hist[d1] : 0u;
      unsigned mine = c0 + c1, incl = mine;
      for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (lane == 31) warp_sums[warp] = incl;
      __syncthreads();
      unsigned woff = 0;
      for (int w2 = 0; w2 < warp; ++w2) woff += warp_sums[w2];
      unsigned before = woff + incl - mine;                              // elements with a larger digit
      const unsigned krem = s_krem;
      __syncthreads();
      if (2 * tid < nb && before < krem && krem <= before + c0) {
        s_prefix = prefix | ((unsigned long long)d0 << shift);
        s_mask = mask | ((unsigned long long)(nb - 1) << shift);
        s_krem = krem - before;
        if (krem == before + c0 || shift == 0) s_done = 1u;               // whole bin taken (or fully resolved)
      } else if (2 * tid + 1 < nb && before + c0 < krem && krem <= before + c0 + c1) {
        s_prefix = prefix | ((unsigned long long)d1 << shift);
        s_mask = mask | ((unsigned long long)(nb - 1) << shift);
        s_krem = krem - before - c0;
        if (krem == before + c0 + c1 || shift == 0) s_done = 1u;
      }
      __syncthreads();
    }
    if (shift == 0) break;
  }
  __syncthreads();
  // threshold: every composite >= (prefix restricted to the resolved bits) is selected
  const unsigned long long thr = (Ksel == count) ? 0ull : s_prefix;
  for (unsigned i = tid; i < count && Ksel > 0; i += DEC_THREADS) {
    unsigned long long c = list[i];
    if (c >= thr) { unsigned pos = atomicAdd(&s_nsel, 1u); if (pos < DEC_MAXK) sel[pos] = c; }
  }
  __syncthreads();
  // ---- phase 3: sort (descending) and emit ------------------------------------------------------------
  unsigned nsel = min(s_nsel, (unsigned)DEC_MAXK);
  unsigned np2 = 1; while (np2 < nsel) np2 <<= 1;
  for (unsigned i = nsel + tid; i < np2; i += DEC_THREADS) sel[i] = 0ull;
  __syncthreads();
  for (unsigned k2 = 2; k2 <= np2; k2 <<= 1)
    for (unsigned j = k2 >> 1; j > 0; j >>= 1) {
      for (unsigned i = tid; i < np2; i += DEC_THREADS) {
        unsigned ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = sel[i], c = sel[ixj];
          bool desc = ((i & k2) == 0);
          if (desc ? (a < c) : (a > c)) { sel[i] = c; sel[ixj] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < p.K; i += DEC_THREADS) {
    float* d = p.dets + ((size_t)b * p.K + i) * 6;
    if ((unsigned)i < nsel && (unsigned)i < Ksel) {
      unsigned long long c = sel[i];
      uint32_t flat = 0xffffffffu - (uint32_t)(c & 0xffffffffull);
      float logit = key2f((uint32_t)(c >> 32));
      int cls = flat / HW, sp = flat - cls * HW;
      float xs = (float)(sp % p.W), ys = (float)(sp / p.W);
      if (p.reg) { xs += p.reg[(size_t)b * p.reg_is + sp]; ys += p.reg[(size_t)b * p.reg_is + HW + sp]; }
      else { xs += 0.5f; ys += 0.5f; }
      float w = p.wh[(size_t)b * p.wh_is + sp], h = p.wh[(size_t)b * p.wh_is + HW + sp];
      d[0] = xs - w / 2; d[1] = ys - h / 2; d[2] = xs + w / 2; d[3] = ys + h / 2;
      d[4] = p.is_prob ? logit : (float)(1.0 / (1.0 + exp(-(double)logit)));
      d[5] = (float)cls;
      if (p.inds) p.inds[(size_t)b * p.K + i] = (int32_t)flat;
    } else {
      d[0] = d[1] = d[2] = d[3] = d[4] = d[5] = 0.f;
      if (p.inds) p.inds[(size_t)b * p.K + i] = -1;
    }
  }
}

int decode_launch(const float* hm, long long hm_img_stride, const float* wh, long long wh_img_stride, const float* reg,
                  long long reg_img_stride, int batch, int cat, int H, int W, int K, int is_prob,
                  unsigned long long* scratch, unsigned int* counts, float* dets, int32_t* inds, cudaStream_t st) {
  CDN_CHECK(K >= 1 && K <= DEC_MAXK, CDN_ERR_INVALID, "decode: K=%d must be in 1..%d", K, DEC_MAXK);
  CDN_CHECK(cat >= 1 && H >= 1 && W >= 1 && (long long)cat * H * W < (1ll << 31), CDN_ERR_INVALID, "decode: bad shape");
  CDN_CHECK(hm && wh && dets && scratch && counts, CDN_ERR_INVALID, "decode: null pointer");
  CDN_CHECK(batch <= 65535, CDN_ERR_INVALID, "decode: batch %d exceeds 65535", batch);
  if (batch == 0) return 0;
  DecParams p{hm, wh, reg, hm_img_stride, wh_img_stride, reg_img_stride, cat, H, W, K, is_prob, scratch, counts, dets, inds};
  CDN_CUDA(cudaMemsetAsync(counts, 0, (size_t)batch * sizeof(unsigned int), st));
  const unsigned n = (unsigned)cat * H * W;
  if (W % 4 == 0 && (((uintptr_t)hm | (uintptr_t)(hm_img_stride * 4)) & 15) == 0) {
    const int strips = (H + DEC_STRIP - 1) / DEC_STRIP;
    const unsigned threads = (unsigned)cat * strips * (W / 4);
    ctdet_peaks_rows_kernel<<<dim3((threads + 255) / 256, batch), 256, 0, st>>>(p, strips);
    CDN_LAUNCH_CHECK("ctdet_peaks_rows_kernel");
  } else {
    ctdet_peaks_kernel<<<dim3((n + 255) / 256, batch), 256, 0, st>>>(p, batch);
    CDN_LAUNCH_CHECK("ctdet_peaks_kernel");
  }
  ctdet_decode_kernel<<<batch, DEC_THREADS, 0, st>>>(p);
  CDN_LAUNCH_CHECK("ctdet_decode_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// ctdet_post_process on the device (lib/utils/post_process.py:86-103 with transform_preds / affine_transform,
// lib/utils/image.py:14-21,58-61): both box corners of every detection go through the image's inverse affine map,
// evaluated like numpy does -- np.dot(t[float64 2x3], [x, y, 1][float32]) accumulated left to right in double, then
// stored into the float32 detections array.  In place on dets [batch][K][6]; trans: [batch][6] doubles (row major 2x3).
// ---------------------------------------------------------------------------------------------------------
__global__ void ctdet_post_affine_kernel(float* dets, const double* trans, int K, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double* t = trans + (i / K) * 6;
  float* d = dets + i * 6;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const double x = (double)d[2 * c], y = (double)d[2 * c + 1];
    const double nx = __dadd_rn(__dadd_rn(__dmul_rn(t[0], x), __dmul_rn(t[1], y)), t[2]);
    const double ny = __dadd_rn(__dadd_rn(__dmul_rn(t[3], x), __dmul_rn(t[4], y)), t[5]);
    d[2 * c] = (float)nx; d[2 * c + 1] = (float)ny;
  }
}

extern "C" int cdn_ctdet_post_affine(float* d_dets, int batch, int K, const double* h_trans, cdn_stream_t stream) {
  CDN_CHECK(d_dets && h_trans && batch >= 0 && K >= 1, CDN_ERR_INVALID, "post_affine: bad arguments");
  if (batch == 0) return 0;
  double* d_t = nullptr;
  CDN_CUDA(cudaMalloc(&d_t, (size_t)batch * 6 * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(d_t, h_trans, (size_t)batch * 6 * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  const long long total = (long long)batch * K;
  if (e == cudaSuccess) {
    ctdet_post_affine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_dets, d_t, K, total);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(d_t);
  CDN_CHECK(e == cudaSuccess, CDN_ERR_CUDA, "post_affine: %s", cudaGetErrorString(e));
  return 0;
}

// ---- flip-test merge (lib/detectors/ctdet.py:35-38) --------------------------------------------------------------
// out[i][c][y][x] = (in[2i][c][y][x] + in[2i+1][c][y][W-1-x]) / 2 in fp32: the reference's
// (hm[0:1] + flip_tensor(hm[1:2])) / 2 for the image and its mirror (one add, one exact halving).
__global__ void ctdet_flip_merge_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int H, int W, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % W); long long t = idx / W;
  const int y = (int)(t % H); t /= H;
  const int c = (int)(t % planes); const long long i = t / planes;
  const size_t plane = (size_t)H * W;
  const float a = in[((size_t)(2 * i) * planes + c) * plane + (size_t)y * W + x];
  const float b = in[((size_t)(2 * i + 1) * planes + c) * plane + (size_t)y * W + (W - 1 - x)];
  out[idx] = __fdiv_rn(__fadd_rn(a, b), 2.0f);
}

extern "C" int cdn_ctdet_flip_merge(const float* d_hm, const float* d_wh, int pairs, int cat, int H, int W, float* d_out_hm,
                                    float* d_out_wh, cdn_stream_t stream) {
  CDN_CHECK(d_hm && d_out_hm && pairs >= 0 && cat >= 1 && H >= 1 && W >= 1, CDN_ERR_INVALID, "flip_merge: bad arguments");
  CDN_CHECK((d_wh == nullptr) == (d_out_wh == nullptr), CDN_ERR_INVALID, "flip_merge: wh input and output must both be given or both be null");
  if (pairs == 0) return 0;
  long long total = (long long)pairs * cat * H * W;
  ctdet_flip_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_hm, d_out_hm, cat, H, W, total);
  CDN_LAUNCH_CHECK("ctdet_flip_merge_kernel");
  if (d_wh) {
    total = (long long)pairs * 2 * H * W;
    ctdet_flip_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_wh, d_out_wh, 2, H, W, total);
    CDN_LAUNCH_CHECK("ctdet_flip_merge_kernel");
  }
  return 0;
}
