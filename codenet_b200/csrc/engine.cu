// C ABI entry points: error plumbing, stand-alone operator calls, and the whole-network engine
// (a static plan of kernel launches over pre-allocated NHWC int8 activation buffers, replayed as a CUDA graph).
#include "layers.cuh"
#include <algorithm>
#include <mutex>

// ---- error plumbing -----------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
unsigned g_cdn_debug_flags = 0;

int cdn_fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
std::mutex& cdn_attr_mutex() { static std::mutex m; return m; }
int cdn_num_sms() {                            // SM count of the CURRENT device (cached per device)
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(cdn_attr_mutex());
  if (!sms[dev]) { cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev); if (sms[dev] <= 0) sms[dev] = 148; }
  return sms[dev];
}

int dev_requant_upload(DevRequant& d, const cdn_requant* rq, const int32_t* acc_bias_host, int n_pad) {
  CDN_CHECK(rq && rq->n <= n_pad, CDN_ERR_INVALID, "requant: n=%d exceeds padded size %d", rq ? rq->n : -1, n_pad);
  std::vector<float> Mh(n_pad, 0.f), Bh(n_pad, 0.f), thr(n_pad, 0.5f);
  std::vector<double> M(n_pad, 0.0), B(n_pad, 0.0);
  std::vector<int32_t> ab(n_pad, 0);
  for (int i = 0; i < rq->n; ++i) {
    CDN_CHECK(std::isfinite(rq->M[i]) && std::isfinite(rq->B[i]), CDN_ERR_INVALID, "requant: non-finite constant at channel %d", i);
    RqFast f = rq_fast_from(rq->M[i], rq->B[i]);
    Mh[i] = f.Mh; Bh[i] = f.Bh; thr[i] = f.thr; M[i] = rq->M[i]; B[i] = rq->B[i];
  }
  if (acc_bias_host) for (int i = 0; i < n_pad; ++i) ab[i] = acc_bias_host[i];
  if (dev_upload(&d.Mh, Mh.data(), n_pad)) return CDN_ERR_CUDA;
  if (dev_upload(&d.Bh, Bh.data(), n_pad)) return CDN_ERR_CUDA;
  if (dev_upload(&d.thr, thr.data(), n_pad)) return CDN_ERR_CUDA;
  if (dev_upload(&d.M, M.data(), n_pad)) return CDN_ERR_CUDA;
  if (dev_upload(&d.B, B.data(), n_pad)) return CDN_ERR_CUDA;
  if (dev_upload(&d.acc_bias, ab.data(), n_pad)) return CDN_ERR_CUDA;
  d.lo = std::max(-128, std::min(127, rq->lo)); d.n = n_pad;
  return 0;
}
void dev_requant_free(DevRequant& d) {
  cudaFree(d.Mh); cudaFree(d.Bh); cudaFree(d.thr); cudaFree(d.M); cudaFree(d.B); cudaFree(d.acc_bias);
  d = DevRequant();
}

// ---- integer requantisation solver (common.cuh, RqInt) ---------------------------------------------------------
// the definition of the result: rint(fl64(fl64(v*M) + B)), product and sum rounded separately (no contraction)
static inline long long rq_q64(long long v, double M, double B) {
  volatile double p = (double)v * M;
  volatile double t = p + B;
  return llrint((double)t);
}

bool rq_int_solve(double M, double B, int lo, long long vmin, long long vmax, RqInt* out) {
  if (!std::isfinite(M) || !std::isfinite(B) || vmin > vmax) return false;
  if (lo < -128) lo = -128;
  if (lo > 127) lo = 127;
  if (M == 0.0) {                              // padding channels: the result is the constant rint(B) (clamped by the caller's sat8 / lo)
    const double q = std::min(std::max(nearbyint(B), -1048576.0), 1048576.0);
    out->Mi = 0; out->sh = 0; out->Bi = (long long)q * (1ll << 32);
    return true;
  }
  if (!(M > 0.0)) return false;
  int e = 0; (void)frexp(M, &e);               // M in [2^(e-1), 2^e)  ->  M * 2^(31-e) in [2^30, 2^31)
  int S_nat = 31 - e;
  if (S_nat < 32) S_nat = 32;
  if (S_nat > 63) return false;
  // a[k]: smallest v with result >= k (vmax + 1 when none, vmin when all), k = lo+1 .. 127
  const int nk = 127 - lo;
  std::vector<long long> a((size_t)std::max(nk, 0));
  const long long qmin = rq_q64(vmin, M, B), qmax = rq_q64(vmax, M, B);
  for (int i = 0; i < nk; ++i) {
    const int k = lo + 1 + i;
    if (qmax < k) { a[i] = vmax + 1; continue; }
    if (qmin >= k) { a[i] = vmin; continue; }
    long long lo_v = vmin, hi_v = vmax;        // result(lo_v) < k <= result(hi_v)
    const double est = ceil(((double)k - 0.5 - B) / M);
    if (est > (double)vmin && est <= (double)vmax) {
      const long long v0 = (long long)est;
      for (long long c = v0 - 2; c <= v0 + 2; ++c) {
        if (c <= vmin || c > vmax) continue;
        if (rq_q64(c - 1, M, B) < k && rq_q64(c, M, B) >= k) { lo_v = c - 1; hi_v = c; break; }
      }
    }
    while (hi_v - lo_v > 1) {
      const long long mid = lo_v + (hi_v - lo_v) / 2;
      if (rq_q64(mid, M, B) >= k) hi_v = mid; else lo_v = mid;
    }
    a[i] = hi_v;
  }
  const __int128 one = 1;
  const __int128 vm = std::max(vmin < 0 ? -(__int128)vmin : (__int128)vmin, vmax < 0 ? -(__int128)vmax : (__int128)vmax);
  // Scale candidates: S = 32 first (shift 0: the device then needs no SHF after the IMAD.HI -- kernels whose layers solve with
  // shift 0 in every channel run a variant without it), then the scale that uses all 31 multiplier bits.
  for (int pass = 0; pass < 2; ++pass) {
    int S = pass == 0 ? 32 : S_nat;
    if (pass == 1 && S_nat == 32) break;
    double mi_d = ldexp(M, S);
    if (mi_d >= 2147483648.0 - 256.0) {        // keep room for the +-d search below
      if (S == 32) { if (pass == 0) continue; return false; }   // multiplier >= 0.5: the output grid is as fine as the accumulator's
      --S; mi_d *= 0.5;
    }
    const long long Mi0 = llrint(mi_d);
    const __int128 nominal = (__int128)ldexpl((long double)B + 0.5L, S);
    for (int t = 0; t <= 256; ++t) {
      const long long d = (t & 1) ? (t + 1) / 2 : -(long long)(t / 2);       // 0, 1, -1, 2, -2, ...
      const long long Mi = Mi0 + d;
      if (Mi <= 0 || Mi >= (1ll << 31)) continue;
      __int128 L = 0, U = 0; bool hasL = false, hasU = false;
      for (int i = 0; i < nk; ++i) {
        const __int128 K2 = (__int128)(lo + 1 + i) * (one << S);
        if (a[i] <= vmax) { const __int128 c = K2 - (__int128)a[i] * Mi; if (!hasL || c > L) L = c; hasL = true; }
        if (a[i] - 1 >= vmin) { const __int128 c = K2 - (__int128)(a[i] - 1) * Mi - 1; if (!hasU || c < U) U = c; hasU = true; }
      }
      if (hasL && hasU && L > U) continue;
      __int128 Bi;
      if (hasL && hasU) Bi = L + (U - L) / 2;
      else if (hasL) Bi = std::max(L, nominal);
      else if (hasU) Bi = std::min(U, nominal);
      else Bi = nominal;
      const __int128 absB = Bi < 0 ? -Bi : Bi;
      if (absB + vm * Mi >= (one << 62)) break;
      out->Mi = (int32_t)Mi; out->sh = S - 32; out->Bi = (long long)Bi;
      return true;
    }
  }
  return false;
}

bool rq_int_rebase(RqInt* r, long long acc_bias, long long amax) {
  const __int128 one = 1;
  const __int128 Bi = (__int128)r->Bi + (__int128)acc_bias * r->Mi;
  const __int128 absB = Bi < 0 ? -Bi : Bi;
  if (absB + (__int128)(amax < 0 ? -amax : amax) * r->Mi >= (one << 62)) return false;
  r->Bi = (long long)Bi;
  return true;
}

extern "C" int cdn_rq_int_solve(double M, double B, int lo, int64_t vmin, int64_t vmax, int32_t* Mi, int32_t* sh, int64_t* Bi) {
  CDN_CHECK(Mi && sh && Bi, CDN_ERR_INVALID, "rq_int_solve: null output pointer");
  RqInt r;
  if (!rq_int_solve(M, B, lo, (long long)vmin, (long long)vmax, &r))
    return cdn_fail(CDN_ERR_INVALID, "rq_int_solve: no exact fixed-point pair for M=%.17g B=%.17g on [%lld, %lld]", M, B,
                    (long long)vmin, (long long)vmax);
  *Mi = r.Mi; *sh = r.sh; *Bi = (int64_t)r.Bi;
  return 0;
}

extern "C" const char* cdn_last_error(void) { return g_err; }
extern "C" int cdn_version(void) { return 100; }
extern "C" int cdn_set_debug_flags(unsigned flags) { g_cdn_debug_flags = flags; return 0; }
extern "C" int cdn_check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return cdn_fail(CDN_ERR_NO_DEVICE, "no CUDA device visible; codenet_b200 has no CPU fallback"); }
  CDN_CHECK(device >= 0 && device < n, CDN_ERR_NO_DEVICE, "device %d out of range (%d visible)", device, n);
  int major = 0;
  CDN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CDN_CHECK(major == 10, CDN_ERR_NO_DEVICE, "device %d has compute capability %d.x; this library is built for sm_100a only", device, major);
  return 0;
}

// ---- stand-alone operator calls (tests / integration; they synchronise the stream before returning) -----------
extern "C" int cdn_stem_f32_i8(const float* d_img, int batch, int H, int W, int stride, int pool, const int8_t* wq, int C,
                               const cdn_requant* rq, int8_t* d_out, int out_pitch, cdn_stream_t stream) {
  StemDevice d{}; int8_t* tmp = nullptr;
  int r = stem_device_build(d, wq, C, rq);
  if (!r && pool) {
    size_t bytes = (size_t)batch * ((H - 1) / stride + 1) * ((W - 1) / stride + 1) * out_pitch;
    if (cudaMalloc(&tmp, bytes ? bytes : 16) != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "stem: scratch allocation failed");
  }
  if (!r) r = stem_launch(d, d_img, batch, H, W, stride, pool, d_out, out_pitch, tmp, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && e != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "stem: %s", cudaGetErrorString(e));
  cudaFree(tmp); stem_device_free(d);
  return r;
}

extern "C" int cdn_dw3x3_i8(const int8_t* d_in, int in_pitch, int batch, int H, int W, int in_shift, int stride,
                            const int8_t* wq, int C, int zx, const cdn_requant* rq, int8_t* d_out, int out_pitch,
                            cdn_stream_t stream) {
  DwDevice d{};
  int Cp = std::min(in_pitch, out_pitch);
  int r = dw_device_build(d, wq, nullptr, C, Cp, zx, rq);
  if (!r) r = dw_launch(d, d_in, in_pitch, d_out, out_pitch, batch, H, W, in_shift, stride, zx, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && e != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "dw3x3: %s", cudaGetErrorString(e));
  dw_device_free(d);
  return r;
}

extern "C" int cdn_deform_dw_w4a8(const int8_t* d_in, int in_pitch, int batch, int H, int W, int in_shift,
                                  const cdn_deform_scale* sc, const int8_t* wq, int C, int zx, const cdn_requant* rq,
                                  int8_t* d_out, int out_pitch, float* d_sval, cdn_stream_t stream) {
  CDN_CHECK(sc && sc->ws, CDN_ERR_INVALID, "deform: null scale descriptor");
  DwDevice d{};
  int Cp = std::min(in_pitch, out_pitch);
  int r = dw_device_build(d, wq, sc->ws, C, Cp, zx, rq);
  if (!r) r = deform_scale_build(d, sc, sc->ws, C, zx);
  if (!r) r = deform_launch(d, sc, d_in, in_pitch, d_out, out_pitch, batch, H, W, in_shift, zx, d_sval, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && e != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "deform: %s", cudaGetErrorString(e));
  dw_device_free(d);
  return r;
}

extern "C" int cdn_pw_gemm_i8(const int8_t* d_in, int in_pitch, int64_t pixels, const cdn_pw_desc* desc,
                              const int8_t* d_pass, int pass_pitch, int8_t* d_out, int out_pitch,
                              float* d_out_f32, int pixels_per_image, cdn_stream_t stream) {
  PwDevice d{};
  int r = pw_device_build(d, desc, pass_pitch);
  if (!r) r = pw_launch(d, d_in, in_pitch, pixels, d_pass, pass_pitch, d_out, out_pitch, d_out_f32, pixels_per_image,
                        nullptr, nullptr, nullptr, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && e != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "pw_gemm: %s", cudaGetErrorString(e));
  pw_device_free(d);
  return r;
}

extern "C" int cdn_shuffle_unit_i8(const int8_t* d_x, int x_pitch, int batch, int H, int W, int stride,
                                   const cdn_pw_desc* pw1, const int8_t* dw_wq, int dw_C, int dw_zx, const cdn_requant* dw_rq,
                                   const cdn_pw_desc* pw3, const int8_t* d_pass, int pass_pitch,
                                   int8_t* d_out, int out_pitch, cdn_stream_t stream) {
  CDN_CHECK(d_x && pw1 && dw_wq && dw_rq && pw3 && d_pass && d_out, CDN_ERR_INVALID, "shuffle unit: null argument");
  CDN_CHECK(stride == 1 || stride == 2, CDN_ERR_INVALID, "shuffle unit: stride must be 1 or 2");
  PwDevice a{}, c{}; DwDevice b{};
  const int mid_pitch = (pw1->N + 31) / 32 * 32;
  int r = pw_device_build(a, pw1, 0);
  if (!r) r = dw_device_build(b, dw_wq, nullptr, dw_C, mid_pitch, dw_zx, dw_rq);
  if (!r) r = pw_device_build(c, pw3, pass_pitch);
  if (!r) {
    const bool ok = stride == 1 ? (d_pass == d_x && unit_fused_ok(a, b, c, x_pitch, mid_pitch, out_pitch, H, W))
                                : unit_s2_fused_ok(a, b, c, x_pitch, mid_pitch, pass_pitch, out_pitch, H, W);
    if (!ok) r = cdn_fail(CDN_ERR_INVALID, "unit not fusable: stride %d, pitches %d / %d / %d, map %dx%d (DESIGN.md 4.1b lists what the fused kernels take)",
                          stride, x_pitch, mid_pitch, out_pitch, H, W);
    else if (stride == 1 && unit_fused_ws_ok(a, b, c, x_pitch, mid_pitch, out_pitch, H, W))
      r = unit_fused_ws_launch(a, b, c, d_x, d_out, batch, H, W, dw_zx, nullptr, nullptr, (cudaStream_t)stream);
    else if (stride == 1) r = unit_fused_launch(a, b, c, d_x, d_out, mid_pitch, batch, H, W, dw_zx, nullptr, nullptr, (cudaStream_t)stream);
    else r = unit_s2_fused_launch(a, b, c, d_x, d_pass, pass_pitch, d_out, batch, H, W, dw_zx, nullptr, nullptr, (cudaStream_t)stream);
  }
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && e != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "shuffle unit: %s", cudaGetErrorString(e));
  pw_device_free(a); pw_device_free(c); dw_device_free(b);
  return r;
}

// ---- persistent fused deformable layer (constants uploaded once; used by the layer sweep and by integrations that
// keep the reference's module structure) ------------------------------------------------------------------------------
struct cdn_deform_layer { DwDevice dw; cdn_deform_scale sc; std::vector<int8_t> ws; int C, Cp, zx, device; };

extern "C" int cdn_deform_layer_create(cdn_deform_layer** out, const cdn_deform_scale* sc, const int8_t* wq, int C, int pitch,
                                       int zx, const cdn_requant* rq) {
  CDN_CHECK(out && sc && sc->ws && wq && rq, CDN_ERR_INVALID, "deform layer: null argument");
  int dev = 0; CDN_CUDA(cudaGetDevice(&dev));
  if (int r = cdn_check_device(dev)) return r;
  cdn_deform_layer* L = new cdn_deform_layer();
  L->C = C; L->Cp = pitch; L->zx = zx; L->device = dev;
  L->sc = *sc; L->ws.assign(sc->ws, sc->ws + C); L->sc.ws = L->ws.data();
  if (int r = dw_device_build(L->dw, wq, sc->ws, C, pitch, zx, rq)) { delete L; return r; }
  if (int r = deform_scale_build(L->dw, &L->sc, L->ws.data(), C, zx)) { dw_device_free(L->dw); delete L; return r; }
  *out = L;
  return 0;
}
extern "C" int cdn_deform_layer_run(cdn_deform_layer* L, const int8_t* d_in, int in_pitch, int batch, int H, int W, int in_shift,
                                    int8_t* d_out, int out_pitch, float* d_sval, cdn_stream_t stream) {
  CDN_CHECK(L != nullptr, CDN_ERR_INVALID, "null deform layer");
  CDN_CHECK(std::min(in_pitch, out_pitch) >= L->Cp, CDN_ERR_INVALID, "deform layer: pitch smaller than the layer's %d", L->Cp);
  return deform_launch(L->dw, &L->sc, d_in, in_pitch, d_out, out_pitch, batch, H, W, in_shift, L->zx, d_sval, (cudaStream_t)stream);
}
extern "C" int cdn_deform_layer_destroy(cdn_deform_layer* L) {
  if (!L) return 0;
  cudaSetDevice(L->device);
  dw_device_free(L->dw);
  delete L;
  return 0;
}

static int decode_standalone(const float* d_hm, const float* d_wh, const float* d_reg, int batch, int cat, int H, int W,
                             int K, int is_prob, float* d_dets, int32_t* d_inds, cdn_stream_t stream) {
  unsigned long long* scratch = nullptr;
  CDN_CHECK(batch >= 0 && cat >= 1 && H >= 1 && W >= 1, CDN_ERR_INVALID, "decode: bad shape");
  size_t n = (size_t)batch * cat * H * W;
  CDN_CUDA(cudaMalloc(&scratch, (n ? n : 1) * sizeof(unsigned long long) + ((size_t)batch + 1) * sizeof(unsigned int)));
  long long hw = (long long)H * W;
  int r = decode_launch(d_hm, cat * hw, d_wh, 2 * hw, d_reg, 2 * hw, batch, cat, H, W, K, is_prob, scratch,
                        (unsigned int*)(scratch + (n ? n : 1)), d_dets, d_inds, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && e != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "decode: %s", cudaGetErrorString(e));
  cudaFree(scratch);
  return r;
}
extern "C" size_t cdn_ctdet_decode_ws_bytes(int batch, int cat, int H, int W) {
  if (batch < 0 || cat < 1 || H < 1 || W < 1) return 0;
  size_t n = (size_t)batch * cat * H * W;
  return (n ? n : 1) * sizeof(unsigned long long) + ((size_t)batch + 1) * sizeof(unsigned int);
}
extern "C" int cdn_ctdet_decode_ws(const float* d_hm, long long hm_img_stride, const float* d_wh, long long wh_img_stride,
                                   const float* d_reg, long long reg_img_stride, int batch, int cat, int H, int W,
                                   int K, int is_prob, float* d_dets, int32_t* d_inds, void* d_ws, size_t ws_bytes,
                                   cdn_stream_t stream) {
  CDN_CHECK(batch >= 0 && cat >= 1 && H >= 1 && W >= 1, CDN_ERR_INVALID, "decode: bad shape");
  CDN_CHECK(d_ws && ((uintptr_t)d_ws & 7) == 0 && ws_bytes >= cdn_ctdet_decode_ws_bytes(batch, cat, H, W), CDN_ERR_INVALID,
            "decode: workspace of %zu bytes (8-byte aligned) needed, got %zu", cdn_ctdet_decode_ws_bytes(batch, cat, H, W), ws_bytes);
  size_t n = (size_t)batch * cat * H * W;
  long long hw = (long long)H * W;
  CDN_CHECK((!hm_img_stride || hm_img_stride >= cat * hw) && (!wh_img_stride || wh_img_stride >= 2 * hw) &&
            (!reg_img_stride || reg_img_stride >= 2 * hw), CDN_ERR_INVALID, "decode: image stride smaller than one image");
  unsigned long long* scratch = (unsigned long long*)d_ws;
  return decode_launch(d_hm, hm_img_stride ? hm_img_stride : cat * hw, d_wh, wh_img_stride ? wh_img_stride : 2 * hw, d_reg,
                       reg_img_stride ? reg_img_stride : 2 * hw, batch, cat, H, W, K, is_prob, scratch,
                       (unsigned int*)(scratch + (n ? n : 1)), d_dets, d_inds, (cudaStream_t)stream);
}
extern "C" int cdn_ctdet_decode(const float* d_hm, const float* d_wh, const float* d_reg, int batch, int cat, int H, int W,
                                int K, float* d_dets, int32_t* d_inds, cdn_stream_t stream) {
  return decode_standalone(d_hm, d_wh, d_reg, batch, cat, H, W, K, 0, d_dets, d_inds, stream);
}
extern "C" int cdn_ctdet_decode_prob(const float* d_heat, const float* d_wh, const float* d_reg, int batch, int cat, int H,
                                     int W, int K, float* d_dets, int32_t* d_inds, cdn_stream_t stream) {
  return decode_standalone(d_heat, d_wh, d_reg, batch, cat, H, W, K, 1, d_dets, d_inds, stream);
}

// ---- engine -------------------------------------------------------------------------------------------------
struct EngTensor { int H, W, pitch; int8_t* ptr; };
struct EngOp {
  int kind;                                  // 0 stem, 1 dw, 2 deform, 3 pw
  int in_t, pass_t, out_t, in_shift, stride, pool, zx, H, W;
  StemDevice stem; DwDevice dw; PwDevice pw; cdn_deform_scale sc; std::vector<int8_t> ws_host;
  CUtensorMap tmA, tmP, tmO;
};

struct cdn_engine {
  int device = 0, max_batch = 0, finalized = 0;
  int in_H = 0, in_W = 0;
  std::vector<EngTensor> tensors;
  std::vector<EngOp*> ops;
  int cat = 0, hH = 0, hW = 0, K = 100, has_reg = 1, n_f32 = 0;
  float* heads = nullptr;                    // fp32 [batch][n_f32][hH*hW] logits / wh / reg
  unsigned long long* dec_scratch = nullptr;
  float* dets = nullptr; int32_t* inds = nullptr;
  int8_t* stem_tmp = nullptr;
  // host path
  float* d_img = nullptr; size_t d_img_bytes = 0;
  float* lut = nullptr;                      // 3 x 256 normalisation table for uint8 input
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  std::vector<cudaEvent_t> ev;
  // pipelined host path (submit / wait): two input + detection slots, so the H2D of step n+1 overlaps the compute of step n
  struct Slot {
    uint8_t* d_img = nullptr; float* dets = nullptr; int32_t* inds = nullptr;
    cudaEvent_t copied = nullptr, done = nullptr; int busy = 0, used = 0;
  } slots[2];
  int host_chunk = 64, use_graph = 1, hm_logits = 0;
  int fuse_heads = 1;                        // heads.dw2 + heads.out as one kernel when the pair is eligible (heads_fused.cu)
  int fuse_units = 1;                        // stride-1 ShuffleNetV2 units (pw1 -> dw2 -> pw3) as one kernel (unit_fused.cu); 2: also write the tensors in between
  // graph cache
  struct GraphKey { const void* a[6]; int batch; bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(*this)) == 0; } };
  std::vector<std::pair<GraphKey, cudaGraphExec_t>> graphs;
  int launches = 0;
};

__global__ void heads_copyout_kernel(const float* heads, int n_f32, int cat, int ppi, long long total,
                                     float* hm, float* wh, float* reg, int hm_logits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int pi = (int)(i % ppi); long long t = i / ppi; int n = (int)(t % n_f32); long long b = t / n_f32;
  float v = heads[i];
  if (n < cat) { if (hm) hm[((size_t)b * cat + n) * ppi + pi] = hm_logits ? v : 1.f / (1.f + __expf(-v)); }
  else if (n < cat + 2) { if (wh) wh[((size_t)b * 2 + (n - cat)) * ppi + pi] = v; }
  else if (n < cat + 4) { if (reg) reg[((size_t)b * 2 + (n - cat - 2)) * ppi + pi] = v; }
}

#define ENG_CHECK(e) CDN_CHECK((e) != nullptr, CDN_ERR_INVALID, "null engine")

extern "C" int cdn_engine_create(cdn_engine** out, int device) {
  CDN_CHECK(out != nullptr, CDN_ERR_INVALID, "null out pointer");
  if (int r = cdn_check_device(device)) return r;
  CDN_CUDA(cudaSetDevice(device));
  cdn_engine* e = new cdn_engine();
  e->device = device;
  CDN_CUDA(cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
  CDN_CUDA(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
  *out = e;
  return 0;
}

extern "C" int cdn_engine_destroy(cdn_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.second);
  for (auto* op : e->ops) { stem_device_free(op->stem); dw_device_free(op->dw); pw_device_free(op->pw); delete op; }
  for (auto& t : e->tensors) cudaFree(t.ptr);
  cudaFree(e->heads); cudaFree(e->dec_scratch); cudaFree(e->dets); cudaFree(e->inds); cudaFree(e->stem_tmp); cudaFree(e->d_img); cudaFree(e->lut);
  for (auto ev : e->ev) cudaEventDestroy(ev);
  for (auto& sl : e->slots) {
    cudaFree(sl.d_img); cudaFree(sl.dets); cudaFree(sl.inds);
    if (sl.copied) cudaEventDestroy(sl.copied);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  if (e->s_compute) cudaStreamDestroy(e->s_compute);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  delete e;
  return 0;
}

extern "C" int cdn_engine_set_option(cdn_engine* e, const char* name, int value) {
  ENG_CHECK(e);
  if (!strcmp(name, "host_chunk")) e->host_chunk = std::max(1, value);
  else if (!strcmp(name, "use_graph")) e->use_graph = value;
  else if (!strcmp(name, "hm_logits")) e->hm_logits = value ? 1 : 0;
  else if (!strcmp(name, "fuse_heads")) e->fuse_heads = value ? 1 : 0;
  else if (!strcmp(name, "fuse_units")) e->fuse_units = std::max(0, std::min(2, value));
  else return cdn_fail(CDN_ERR_INVALID, "unknown engine option '%s'", name);
  return 0;
}

extern "C" int cdn_engine_add_tensor(cdn_engine* e, int H, int W, int pitch) {
  ENG_CHECK(e);
  CDN_CHECK(!e->finalized, CDN_ERR_STATE, "engine already finalized");
  CDN_CHECK(H > 0 && W > 0 && pitch > 0 && pitch % 32 == 0, CDN_ERR_INVALID, "tensor %dx%d pitch %d: pitch must be a multiple of 32", H, W, pitch);
  e->tensors.push_back(EngTensor{H, W, pitch, nullptr});
  return (int)e->tensors.size() - 1;
}

static int check_t(cdn_engine* e, int t, const char* what) {
  CDN_CHECK(t >= 0 && t < (int)e->tensors.size(), CDN_ERR_INVALID, "%s: tensor id %d out of range", what, t);
  return 0;
}

extern "C" int cdn_engine_add_stem(cdn_engine* e, int out_t, int H, int W, int stride, int pool, const int8_t* wq, int C,
                                   const cdn_requant* rq) {
  ENG_CHECK(e);
  CDN_CHECK(!e->finalized, CDN_ERR_STATE, "engine already finalized");
  if (int r = check_t(e, out_t, "stem")) return r;
  CDN_CUDA(cudaSetDevice(e->device));
  EngOp* op = new EngOp();
  op->kind = 0; op->out_t = out_t; op->H = H; op->W = W; op->stride = stride; op->pool = pool;
  if (int r = stem_device_build(op->stem, wq, C, rq)) { delete op; return r; }
  int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  if (pool) { Ho = (Ho - 1) / 2 + 1; Wo = (Wo - 1) / 2 + 1; }
  const EngTensor& t = e->tensors[out_t];
  CDN_CHECK(t.H == Ho && t.W == Wo, CDN_ERR_INVALID, "stem: output tensor is %dx%d, expected %dx%d", t.H, t.W, Ho, Wo);
  e->in_H = H; e->in_W = W;
  e->ops.push_back(op);
  return 0;
}

extern "C" int cdn_engine_add_dw(cdn_engine* e, int in_t, int out_t, int in_shift, int stride, const int8_t* wq, int C,
                                 int zx, const cdn_requant* rq) {
  ENG_CHECK(e);
  CDN_CHECK(!e->finalized, CDN_ERR_STATE, "engine already finalized");
  if (int r = check_t(e, in_t, "dw")) return r;
  if (int r = check_t(e, out_t, "dw")) return r;
  CDN_CUDA(cudaSetDevice(e->device));
  const EngTensor &ti = e->tensors[in_t], &to = e->tensors[out_t];
  int H = ti.H << in_shift, W = ti.W << in_shift;
  CDN_CHECK(to.H == (H - 1) / stride + 1 && to.W == (W - 1) / stride + 1, CDN_ERR_INVALID, "dw: output tensor shape mismatch");
  EngOp* op = new EngOp();
  op->kind = 1; op->in_t = in_t; op->out_t = out_t; op->in_shift = in_shift; op->stride = stride; op->zx = zx; op->H = H; op->W = W;
  if (int r = dw_device_build(op->dw, wq, nullptr, C, std::min(ti.pitch, to.pitch), zx, rq)) { delete op; return r; }
  e->ops.push_back(op);
  return 0;
}

extern "C" int cdn_engine_add_deform(cdn_engine* e, int in_t, int out_t, int in_shift, const cdn_deform_scale* sc,
                                     const int8_t* wq, int C, int zx, const cdn_requant* rq) {
  ENG_CHECK(e);
  CDN_CHECK(!e->finalized, CDN_ERR_STATE, "engine already finalized");
  CDN_CHECK(sc && sc->ws, CDN_ERR_INVALID, "deform: null scale descriptor");
  if (int r = check_t(e, in_t, "deform")) return r;
  if (int r = check_t(e, out_t, "deform")) return r;
  CDN_CUDA(cudaSetDevice(e->device));
  const EngTensor &ti = e->tensors[in_t], &to = e->tensors[out_t];
  int H = ti.H << in_shift, W = ti.W << in_shift;
  CDN_CHECK(to.H == H && to.W == W, CDN_ERR_INVALID, "deform: output tensor shape mismatch");
  EngOp* op = new EngOp();
  op->kind = 2; op->in_t = in_t; op->out_t = out_t; op->in_shift = in_shift; op->stride = 1; op->zx = zx; op->H = H; op->W = W;
  op->sc = *sc; op->ws_host.assign(sc->ws, sc->ws + C); op->sc.ws = op->ws_host.data();
  if (int r = dw_device_build(op->dw, wq, sc->ws, C, std::min(ti.pitch, to.pitch), zx, rq)) { delete op; return r; }
  if (int r = deform_scale_build(op->dw, &op->sc, op->ws_host.data(), C, zx)) { dw_device_free(op->dw); delete op; return r; }
  e->ops.push_back(op);
  return 0;
}

extern "C" int cdn_engine_add_pw(cdn_engine* e, int in_t, int pass_t, int out_t, const cdn_pw_desc* desc) {
  ENG_CHECK(e);
  CDN_CHECK(!e->finalized, CDN_ERR_STATE, "engine already finalized");
  if (int r = check_t(e, in_t, "pw")) return r;
  if (pass_t >= 0) if (int r = check_t(e, pass_t, "pw pass")) return r;
  if (out_t >= 0) if (int r = check_t(e, out_t, "pw out")) return r;
  CDN_CUDA(cudaSetDevice(e->device));
  const EngTensor& ti = e->tensors[in_t];
  if (pass_t >= 0) CDN_CHECK(e->tensors[pass_t].H == ti.H && e->tensors[pass_t].W == ti.W, CDN_ERR_INVALID, "pw: pass tensor shape mismatch");
  if (out_t >= 0) CDN_CHECK(e->tensors[out_t].H == ti.H && e->tensors[out_t].W == ti.W, CDN_ERR_INVALID, "pw: output tensor shape mismatch");
  CDN_CHECK((out_t >= 0) != (desc && desc->n_f32 > 0), CDN_ERR_INVALID, "pw: exactly one of int8 output tensor / fp32 head output must be set");
  EngOp* op = new EngOp();
  op->kind = 3; op->in_t = in_t; op->pass_t = pass_t; op->out_t = out_t; op->H = ti.H; op->W = ti.W;
  if (int r = pw_device_build(op->pw, desc, pass_t >= 0 ? e->tensors[pass_t].pitch : 0)) { delete op; return r; }
  if (desc->n_f32 > 0) e->n_f32 = desc->n_f32;
  e->ops.push_back(op);
  return 0;
}

extern "C" int cdn_engine_set_heads(cdn_engine* e, int cat, int H, int W, int K, int has_reg) {
  ENG_CHECK(e);
  CDN_CHECK(cat > 0 && H > 0 && W > 0 && K > 0 && K <= 1024, CDN_ERR_INVALID, "heads: bad arguments");
  e->cat = cat; e->hH = H; e->hW = W; e->K = K; e->has_reg = has_reg;
  return 0;
}

extern "C" int cdn_engine_finalize(cdn_engine* e, int max_batch) {
  ENG_CHECK(e);
  CDN_CHECK(!e->finalized, CDN_ERR_STATE, "engine already finalized");
  CDN_CHECK(max_batch > 0 && !e->ops.empty() && e->cat > 0, CDN_ERR_INVALID, "finalize: empty plan, heads not set, or bad batch");
  CDN_CHECK(e->n_f32 == e->cat + 2 + (e->has_reg ? 2 : 0), CDN_ERR_INVALID, "finalize: head conv has %d fp32 planes, expected %d", e->n_f32, e->cat + 2 + (e->has_reg ? 2 : 0));
  CDN_CUDA(cudaSetDevice(e->device));
  e->max_batch = max_batch;
  for (auto& t : e->tensors) {
    size_t bytes = (size_t)max_batch * t.H * t.W * t.pitch;
    CDN_CUDA(cudaMalloc((void**)&t.ptr, bytes));
    CDN_CUDA(cudaMemset(t.ptr, 0, bytes));
  }
  size_t ppi = (size_t)e->hH * e->hW;
  CDN_CUDA(cudaMalloc((void**)&e->heads, (size_t)max_batch * e->n_f32 * ppi * sizeof(float)));
  CDN_CUDA(cudaMalloc((void**)&e->dec_scratch, (size_t)max_batch * e->cat * ppi * sizeof(unsigned long long) + ((size_t)max_batch + 1) * sizeof(unsigned int)));
  CDN_CUDA(cudaMalloc((void**)&e->dets, (size_t)max_batch * e->K * 6 * sizeof(float)));
  CDN_CUDA(cudaMalloc((void**)&e->inds, (size_t)max_batch * e->K * sizeof(int32_t)));
  for (auto* op : e->ops) {
    if (op->kind == 0 && op->pool) {
      size_t b = (size_t)max_batch * ((op->H - 1) / op->stride + 1) * ((op->W - 1) / op->stride + 1) * 32;
      CDN_CUDA(cudaMalloc((void**)&e->stem_tmp, b));
    }
    if (op->kind == 3) {
      const EngTensor& ti = e->tensors[op->in_t];
      uint64_t rows = (uint64_t)max_batch * ti.H * ti.W;
      if (int r = make_tmap_2d(&op->tmA, ti.ptr, ti.pitch, rows, ti.pitch, 128)) return r;
      if (op->pass_t >= 0) { const EngTensor& tp = e->tensors[op->pass_t]; if (int r = make_tmap_2d(&op->tmP, tp.ptr, tp.pitch, rows, tp.pitch, 128)) return r; }
      if (op->out_t >= 0) { const EngTensor& to = e->tensors[op->out_t]; if (int r = make_tmap_2d(&op->tmO, to.ptr, to.pitch, rows, to.pitch, 128)) return r; }
    }
  }
  CDN_CUDA(cudaDeviceSynchronize());
  e->finalized = 1;
  return 0;
}

// ops oi, oi+1 = depthwise conv through the virtual x2 upsample feeding the fp32 head conv, eligible for heads_fused.cu
static bool engine_fuses_heads(const cdn_engine* e, size_t oi) {
  if (!e->fuse_heads || (g_cdn_debug_flags & 256u) || oi + 1 >= e->ops.size()) return false;   // bit 8: never fuse (A/B)
  const EngOp *a = e->ops[oi], *b = e->ops[oi + 1];
  if (a->kind != 1 || b->kind != 3 || a->in_shift != 1 || a->stride != 1 || b->in_t != a->out_t || b->out_t >= 0 || b->pass_t >= 0) return false;
  const EngTensor &ti = e->tensors[a->in_t], &tm = e->tensors[a->out_t];
  return heads_fused_ok(a->dw, b->pw, ti.pitch, tm.pitch, ti.H, ti.W) && 2 * ti.H == e->hH && 2 * ti.W == e->hW;
}

// ops oi, oi+1, oi+2 = the branch of a stride-1 ShuffleNetV2 unit (1x1 conv on the second half of the stage tensor, depthwise
// 3x3, 1x1 conv interleaving its columns with the first half), eligible for unit_fused.cu
static bool engine_fuses_unit(const cdn_engine* e, size_t oi) {
  if (!e->fuse_units || (g_cdn_debug_flags & (1u << 19)) || oi + 2 >= e->ops.size()) return false;   // bit 19: never fuse (A/B)
  const EngOp *a = e->ops[oi], *b = e->ops[oi + 1], *c = e->ops[oi + 2];
  if (a->kind != 3 || b->kind != 1 || c->kind != 3) return false;
  if (a->out_t < 0 || a->pass_t >= 0 || b->in_t != a->out_t || b->in_shift != 0 || b->stride != 1 || c->in_t != b->out_t ||
      c->pass_t != a->in_t || c->out_t < 0) return false;
  const EngTensor &tx = e->tensors[a->in_t], &t1 = e->tensors[a->out_t], &t2 = e->tensors[b->out_t], &to = e->tensors[c->out_t];
  if (t1.pitch != t2.pitch) return false;
  return unit_fused_ok(a->pw, b->dw, c->pw, tx.pitch, t1.pitch, to.pitch, tx.H, tx.W);
}

// ops oi, oi+1, oi+2 = branch 2 of the first stride-2 unit (1x1 conv at full resolution, depthwise 3x3 stride 2, 1x1 conv
// interleaving its columns with branch 1's output), eligible for unit_s2_fused.cu
static bool engine_fuses_unit_s2(const cdn_engine* e, size_t oi) {
  if (!e->fuse_units || (g_cdn_debug_flags & (1u << 19)) || oi + 2 >= e->ops.size()) return false;
  const EngOp *a = e->ops[oi], *b = e->ops[oi + 1], *c = e->ops[oi + 2];
  if (a->kind != 3 || b->kind != 1 || c->kind != 3) return false;
  if (a->out_t < 0 || a->pass_t >= 0 || b->in_t != a->out_t || b->in_shift != 0 || b->stride != 2 || c->in_t != b->out_t ||
      c->pass_t < 0 || c->pass_t == a->in_t || c->out_t < 0) return false;
  const EngTensor &tx = e->tensors[a->in_t], &t1 = e->tensors[a->out_t], &t2 = e->tensors[b->out_t], &tp = e->tensors[c->pass_t],
                  &to = e->tensors[c->out_t];
  if (t1.pitch != t2.pitch || t1.H != tx.H || t1.W != tx.W || tp.H != to.H || tp.W != to.W || 2 * to.H != tx.H || 2 * to.W != tx.W) return false;
  return unit_s2_fused_ok(a->pw, b->dw, c->pw, tx.pitch, t1.pitch, tp.pitch, to.pitch, tx.H, tx.W);
}

// d_img: fp32 NCHW image, or (is_u8) uint8 NHWC image normalised in the stem through e->lut
static int engine_enqueue(cdn_engine* e, const void* d_img, int is_u8, int batch, float* d_hm, float* d_wh, float* d_reg,
                          float* d_dets, int32_t* d_inds, cudaStream_t st, std::vector<cudaEvent_t>* marks = nullptr) {
  int launches = 0;
  auto mark = [&]() { if (marks) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); marks->push_back(ev); } };
  mark();
  for (size_t oi = 0; oi < e->ops.size(); ++oi) {
    EngOp* op = e->ops[oi];
    int r = 0;
    if (engine_fuses_heads(e, oi)) {
      // depthwise conv through the upsample + fp32 head conv as one kernel; the int8 tensor between them is not written
      const EngOp* nx = e->ops[oi + 1];
      const EngTensor& ti = e->tensors[op->in_t];
      r = heads_fused_launch(op->dw, nx->pw, ti.ptr, ti.pitch, batch, ti.H, ti.W, op->zx, e->heads, st);
      if (r) return r;
      launches++;
      mark(); mark();                        // the second op of the pair takes no time of its own
      ++oi;
      continue;
    }
    if (engine_fuses_unit(e, oi)) {
      // a whole stride-1 unit as one kernel; the two int8 tensors inside it are written only on request (fuse_units = 2)
      const EngOp *dwo = e->ops[oi + 1], *p3 = e->ops[oi + 2];
      const EngTensor &tx = e->tensors[op->in_t], &t1 = e->tensors[op->out_t], &t2 = e->tensors[dwo->out_t], &to = e->tensors[p3->out_t];
      if (unit_fused_ws_ok(op->pw, dwo->dw, p3->pw, tx.pitch, t1.pitch, to.pitch, tx.H, tx.W))
        r = unit_fused_ws_launch(op->pw, dwo->dw, p3->pw, tx.ptr, to.ptr, batch, tx.H, tx.W, dwo->zx,
                                 e->fuse_units == 2 ? t1.ptr : nullptr, e->fuse_units == 2 ? t2.ptr : nullptr, st);
      else
        r = unit_fused_launch(op->pw, dwo->dw, p3->pw, tx.ptr, to.ptr, t1.pitch, batch, tx.H, tx.W, dwo->zx,
                              e->fuse_units == 2 ? t1.ptr : nullptr, e->fuse_units == 2 ? t2.ptr : nullptr, st);
      if (r) return r;
      launches++;
      mark(); mark(); mark();                // the other two ops of the unit take no time of their own
      oi += 2;
      continue;
    }
    if (engine_fuses_unit_s2(e, oi)) {
      const EngOp *dwo = e->ops[oi + 1], *p3 = e->ops[oi + 2];
      const EngTensor &tx = e->tensors[op->in_t], &t1 = e->tensors[op->out_t], &t2 = e->tensors[dwo->out_t], &tp = e->tensors[p3->pass_t],
                      &to = e->tensors[p3->out_t];
      r = unit_s2_fused_launch(op->pw, dwo->dw, p3->pw, tx.ptr, tp.ptr, tp.pitch, to.ptr, batch, tx.H, tx.W, dwo->zx,
                               e->fuse_units == 2 ? t1.ptr : nullptr, e->fuse_units == 2 ? t2.ptr : nullptr, st);
      if (r) return r;
      launches++;
      mark(); mark(); mark();
      oi += 2;
      continue;
    }
    switch (op->kind) {
      case 0: {
        const EngTensor& to = e->tensors[op->out_t];
        r = stem_launch(op->stem, is_u8 ? nullptr : (const float*)d_img, batch, op->H, op->W, op->stride, op->pool, to.ptr, to.pitch,
                        e->stem_tmp, st, is_u8 ? (const uint8_t*)d_img : nullptr, e->lut);
        launches += op->pool ? 2 : 1;
      } break;
      case 1: {
        const EngTensor &ti = e->tensors[op->in_t], &to = e->tensors[op->out_t];
        r = dw_launch(op->dw, ti.ptr, ti.pitch, to.ptr, to.pitch, batch, op->H, op->W, op->in_shift, op->stride, op->zx, st);
        launches++;
      } break;
      case 2: {
        const EngTensor &ti = e->tensors[op->in_t], &to = e->tensors[op->out_t];
        r = deform_launch(op->dw, &op->sc, ti.ptr, ti.pitch, to.ptr, to.pitch, batch, op->H, op->W, op->in_shift, op->zx, nullptr, st);
        launches++;
      } break;
      case 3: {
        const EngTensor& ti = e->tensors[op->in_t];
        const int8_t* pass = op->pass_t >= 0 ? e->tensors[op->pass_t].ptr : nullptr;
        int pass_pitch = op->pass_t >= 0 ? e->tensors[op->pass_t].pitch : 0;
        int8_t* out = op->out_t >= 0 ? e->tensors[op->out_t].ptr : nullptr;
        int out_pitch = op->out_t >= 0 ? e->tensors[op->out_t].pitch : 0;
        r = pw_launch(op->pw, ti.ptr, ti.pitch, (long long)batch * ti.H * ti.W, pass, pass_pitch, out, out_pitch,
                      op->out_t < 0 ? e->heads : nullptr, e->hH * e->hW, &op->tmA, op->pass_t >= 0 ? &op->tmP : nullptr,
                      op->out_t >= 0 ? &op->tmO : nullptr, st);
        launches++;
      } break;
    }
    if (r) return r;
    mark();
  }
  const long long ppi = (long long)e->hH * e->hW;
  if (d_hm || d_wh || d_reg) {
    long long total = (long long)batch * e->n_f32 * ppi;
    heads_copyout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->heads, e->n_f32, e->cat, (int)ppi, total, d_hm, d_wh, d_reg, e->hm_logits);
    CDN_LAUNCH_CHECK("heads_copyout_kernel");
    launches++;
  }
  mark();
  if (d_dets) {
    const float* hm = e->heads;
    const float* wh = e->heads + (size_t)e->cat * ppi;
    const float* reg = e->has_reg ? e->heads + (size_t)(e->cat + 2) * ppi : nullptr;
    long long is = (long long)e->n_f32 * ppi;
    if (int r = decode_launch(hm, is, wh, is, reg, is, batch, e->cat, e->hH, e->hW, e->K, 0, e->dec_scratch,
                              (unsigned int*)(e->dec_scratch + (size_t)e->max_batch * e->cat * ppi), d_dets, d_inds, st)) return r;
    launches += 2;
  }
  mark();
  e->launches = launches;
  return 0;
}

// Eager run with a CUDA event between consecutive launches: ms[i] = duration of op i (plan order), then the
// heads copy-out and the decode.  n must be >= number of ops + 2.  Used by bench.py for the per-kernel roofline.
extern "C" int cdn_engine_profile(cdn_engine* e, const float* d_img, int batch, float* ms, int n, cdn_stream_t stream) {
  ENG_CHECK(e);
  CDN_CHECK(e->finalized && batch > 0 && batch <= e->max_batch && d_img && ms, CDN_ERR_INVALID, "profile: bad arguments");
  CDN_CHECK(n >= (int)e->ops.size() + 2, CDN_ERR_INVALID, "profile: need room for %d timings", (int)e->ops.size() + 2);
  CDN_CUDA(cudaSetDevice(e->device));
  std::vector<cudaEvent_t> marks;
  int r = engine_enqueue(e, d_img, 0, batch, nullptr, nullptr, nullptr, e->dets, e->inds, (cudaStream_t)stream, &marks);
  cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
  if (!r && ce != cudaSuccess) r = cdn_fail(CDN_ERR_CUDA, "profile: %s", cudaGetErrorString(ce));
  for (int i = 0; i < n; ++i) ms[i] = 0.f;
  if (!r) for (size_t i = 0; i + 1 < marks.size() && (int)i < n; ++i) cudaEventElapsedTime(&ms[i], marks[i], marks[i + 1]);
  for (auto ev : marks) cudaEventDestroy(ev);
  return r;
}

static int engine_run_any(cdn_engine* e, const void* d_img, int is_u8, int batch, float* d_hm, float* d_wh, float* d_reg,
                          float* d_dets, int32_t* d_inds, cdn_stream_t stream) {
  ENG_CHECK(e);
  CDN_CHECK(e->finalized, CDN_ERR_STATE, "engine not finalized");
  CDN_CHECK(batch >= 0 && batch <= e->max_batch, CDN_ERR_INVALID, "batch %d exceeds the finalized maximum %d", batch, e->max_batch);
  CDN_CHECK(d_img != nullptr, CDN_ERR_INVALID, "null image pointer");
  CDN_CHECK(!is_u8 || e->lut != nullptr, CDN_ERR_STATE, "uint8 input needs cdn_engine_set_normalization first");
  if (batch == 0) return 0;
  CDN_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!e->use_graph || (g_cdn_debug_flags & 2u)) return engine_enqueue(e, d_img, is_u8, batch, d_hm, d_wh, d_reg, d_dets, d_inds, st);
  cdn_engine::GraphKey key; memset(&key, 0, sizeof(key));
  key.a[0] = d_img; key.a[1] = d_hm; key.a[2] = d_wh; key.a[3] = d_reg; key.a[4] = d_dets; key.a[5] = d_inds; key.batch = batch | (e->hm_logits << 30) | (is_u8 << 29) | (e->fuse_heads << 28) | (e->fuse_units << 26);
  for (auto& g : e->graphs) if (g.first == key) { CDN_CUDA(cudaGraphLaunch(g.second, st)); return 0; }
  // capture once per (pointers, batch)
  cudaStream_t cap = e->s_compute;
  CDN_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
  int r = engine_enqueue(e, d_img, is_u8, batch, d_hm, d_wh, d_reg, d_dets, d_inds, cap);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(cap, &graph);
  if (r) { if (graph) cudaGraphDestroy(graph); return r; }
  CDN_CHECK(ce == cudaSuccess && graph, CDN_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  CDN_CHECK(ce == cudaSuccess, CDN_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
  if (e->graphs.size() >= 16) { cudaGraphExecDestroy(e->graphs.front().second); e->graphs.erase(e->graphs.begin()); }
  e->graphs.push_back({key, exec});
  CDN_CUDA(cudaGraphLaunch(exec, st));
  return 0;
}

extern "C" int cdn_engine_run(cdn_engine* e, const float* d_img, int batch, float* d_hm, float* d_wh, float* d_reg,
                              float* d_dets, int32_t* d_inds, cdn_stream_t stream) {
  return engine_run_any(e, d_img, 0, batch, d_hm, d_wh, d_reg, d_dets, d_inds, stream);
}
extern "C" int cdn_engine_run_u8(cdn_engine* e, const uint8_t* d_img, int batch, float* d_hm, float* d_wh, float* d_reg,
                                 float* d_dets, int32_t* d_inds, cdn_stream_t stream) {
  return engine_run_any(e, d_img, 1, batch, d_hm, d_wh, d_reg, d_dets, d_inds, stream);
}

// LUT[c][u] = fl32((u / 255. - mean[c]) / std[c]) evaluated in double exactly as numpy evaluates
// ((inp_image / 255. - self.mean) / self.std).astype(np.float32), lib/detectors/base_detector.py:66
extern "C" int cdn_engine_set_normalization(cdn_engine* e, const float* mean3, const float* std3) {
  ENG_CHECK(e);
  CDN_CHECK(mean3 && std3, CDN_ERR_INVALID, "set_normalization: null pointer");
  CDN_CUDA(cudaSetDevice(e->device));
  std::vector<float> lut(768);
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) lut[c * 256 + u] = (float)(((double)u / 255.0 - (double)mean3[c]) / (double)std3[c]);
  if (!e->lut) CDN_CUDA(cudaMalloc((void**)&e->lut, 768 * sizeof(float)));
  CDN_CUDA(cudaMemcpy(e->lut, lut.data(), 768 * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

static int engine_run_host_any(cdn_engine* e, const void* h_img_v, int is_u8, int batch, float* h_dets, int32_t* h_inds) {
  const uint8_t* h_img = (const uint8_t*)h_img_v;
  ENG_CHECK(e);
  CDN_CHECK(e->finalized, CDN_ERR_STATE, "engine not finalized");
  CDN_CHECK(batch >= 0 && batch <= e->max_batch && h_img && h_dets, CDN_ERR_INVALID, "run_host: bad arguments");
  if (batch == 0) return 0;
  CDN_CUDA(cudaSetDevice(e->device));
  const size_t img_bytes = (size_t)3 * e->in_H * e->in_W * (is_u8 ? 1 : sizeof(float));   // per image
  if (!e->d_img) {
    e->d_img_bytes = (size_t)e->max_batch * 3 * e->in_H * e->in_W * sizeof(float);
    CDN_CUDA(cudaMalloc((void**)&e->d_img, e->d_img_bytes));
  }
  // Chunk schedule: the copy engine is faster than the compute (PCIe ~51 GB/s vs ~20 us per image), so chunks grow
  // geometrically (x1.6 from host_chunk/2): compute starts after a short first copy, every later chunk has just
  // arrived when the previous one finishes, and launches get larger (small launches use the GPU poorly).
  // H2D of chunk i+1 overlaps the compute of chunk i (two streams, one event per chunk).
  std::vector<int> sizes;
  {
    const int hc = std::max(2, e->host_chunk);
    int left = batch, c = std::max(1, hc / 2);
    while (left > 0) {
      int take = std::min(left, c);
      if (left - take > 0 && left - take < hc / 2) take = left;       // a short tail is merged into this chunk
      sizes.push_back(take); left -= take;
      c = std::max(8, (int)(c * 1.6 + 7) / 8 * 8);
    }
  }
  const int nchunks = (int)sizes.size();
  while ((int)e->ev.size() < nchunks) { cudaEvent_t ev; CDN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); e->ev.push_back(ev); }
  for (int c = 0, b0 = 0; c < nchunks; b0 += sizes[c], ++c) {
    CDN_CUDA(cudaMemcpyAsync((uint8_t*)e->d_img + (size_t)b0 * img_bytes, h_img + (size_t)b0 * img_bytes, (size_t)sizes[c] * img_bytes,
                             cudaMemcpyHostToDevice, e->s_copy));
    CDN_CUDA(cudaEventRecord(e->ev[c], e->s_copy));
  }
  for (int c = 0, b0 = 0; c < nchunks; b0 += sizes[c], ++c) {
    CDN_CUDA(cudaStreamWaitEvent(e->s_compute, e->ev[c], 0));
    if (int r = engine_run_any(e, (uint8_t*)e->d_img + (size_t)b0 * img_bytes, is_u8, sizes[c], nullptr, nullptr, nullptr,
                               e->dets + (size_t)b0 * e->K * 6, e->inds + (size_t)b0 * e->K, e->s_compute)) return r;
  }
  CDN_CUDA(cudaMemcpyAsync(h_dets, e->dets, (size_t)batch * e->K * 6 * sizeof(float), cudaMemcpyDeviceToHost, e->s_compute));
  if (h_inds) CDN_CUDA(cudaMemcpyAsync(h_inds, e->inds, (size_t)batch * e->K * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_compute));
  CDN_CUDA(cudaStreamSynchronize(e->s_compute));
  return 0;
}

extern "C" int cdn_engine_run_host(cdn_engine* e, const float* h_img, int batch, float* h_dets, int32_t* h_inds) {
  return engine_run_host_any(e, h_img, 0, batch, h_dets, h_inds);
}
extern "C" int cdn_engine_run_host_u8(cdn_engine* e, const uint8_t* h_img, int batch, float* h_dets, int32_t* h_inds) {
  return engine_run_host_any(e, h_img, 1, batch, h_dets, h_inds);
}

// ---- pipelined host path ------------------------------------------------------------------------------------------------
// submit(slot) enqueues H2D (copy stream) -> forward + decode (compute stream) -> D2H of the detections and returns at once;
// wait(slot) blocks until that step's detections are in the caller's host buffers.  With two slots in flight the copy of
// step n+1 runs while step n computes: the step time becomes max(copy, compute) instead of their (chunk-pipelined) sum.
// The activations are shared by both slots; steps are serialised on the compute stream.  Host buffers should be pinned.
static int engine_submit_any(cdn_engine* e, const void* h_img, int is_u8, int batch, float* h_dets, int32_t* h_inds, int slot) {
  ENG_CHECK(e);
  CDN_CHECK(e->finalized, CDN_ERR_STATE, "engine not finalized");
  CDN_CHECK(slot >= 0 && slot < 2, CDN_ERR_INVALID, "submit: slot must be 0 or 1");
  CDN_CHECK(batch > 0 && batch <= e->max_batch && h_img && h_dets, CDN_ERR_INVALID, "submit: bad arguments");
  CDN_CHECK(!is_u8 || e->lut != nullptr, CDN_ERR_STATE, "uint8 input needs cdn_engine_set_normalization first");
  CDN_CUDA(cudaSetDevice(e->device));
  cdn_engine::Slot& sl = e->slots[slot];
  CDN_CHECK(!sl.busy, CDN_ERR_STATE, "submit: slot %d is still in flight (call cdn_engine_wait first)", slot);
  if (!sl.d_img) {
    CDN_CUDA(cudaMalloc((void**)&sl.d_img, (size_t)e->max_batch * 3 * e->in_H * e->in_W * sizeof(float)));
    CDN_CUDA(cudaMalloc((void**)&sl.dets, (size_t)e->max_batch * e->K * 6 * sizeof(float)));
    CDN_CUDA(cudaMalloc((void**)&sl.inds, (size_t)e->max_batch * e->K * sizeof(int32_t)));
    CDN_CUDA(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
    CDN_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  }
  const size_t bytes = (size_t)batch * 3 * e->in_H * e->in_W * (is_u8 ? 1 : sizeof(float));
  if (sl.used) CDN_CUDA(cudaStreamWaitEvent(e->s_copy, sl.done, 0));     // the previous step of this slot has consumed its input
  CDN_CUDA(cudaMemcpyAsync(sl.d_img, h_img, bytes, cudaMemcpyHostToDevice, e->s_copy));
  CDN_CUDA(cudaEventRecord(sl.copied, e->s_copy));
  CDN_CUDA(cudaStreamWaitEvent(e->s_compute, sl.copied, 0));
  if (int r = engine_run_any(e, sl.d_img, is_u8, batch, nullptr, nullptr, nullptr, sl.dets, sl.inds, e->s_compute)) return r;
  CDN_CUDA(cudaMemcpyAsync(h_dets, sl.dets, (size_t)batch * e->K * 6 * sizeof(float), cudaMemcpyDeviceToHost, e->s_compute));
  if (h_inds) CDN_CUDA(cudaMemcpyAsync(h_inds, sl.inds, (size_t)batch * e->K * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_compute));
  CDN_CUDA(cudaEventRecord(sl.done, e->s_compute));
  sl.busy = 1; sl.used = 1;
  return 0;
}
extern "C" int cdn_engine_submit_host(cdn_engine* e, const float* h_img, int batch, float* h_dets, int32_t* h_inds, int slot) {
  return engine_submit_any(e, h_img, 0, batch, h_dets, h_inds, slot);
}
extern "C" int cdn_engine_submit_host_u8(cdn_engine* e, const uint8_t* h_img, int batch, float* h_dets, int32_t* h_inds, int slot) {
  return engine_submit_any(e, h_img, 1, batch, h_dets, h_inds, slot);
}
extern "C" int cdn_engine_wait(cdn_engine* e, int slot) {
  ENG_CHECK(e);
  CDN_CHECK(slot >= 0 && slot < 2, CDN_ERR_INVALID, "wait: slot must be 0 or 1");
  cdn_engine::Slot& sl = e->slots[slot];
  if (!sl.busy) return 0;
  CDN_CUDA(cudaSetDevice(e->device));
  sl.busy = 0;
  CDN_CUDA(cudaEventSynchronize(sl.done));
  return 0;
}

extern "C" int cdn_engine_read_tensor(cdn_engine* e, int tensor, int batch, int8_t* h_out) {
  ENG_CHECK(e);
  CDN_CHECK(e->finalized, CDN_ERR_STATE, "engine not finalized");
  if (int r = check_t(e, tensor, "read_tensor")) return r;
  CDN_CHECK(batch > 0 && batch <= e->max_batch && h_out, CDN_ERR_INVALID, "read_tensor: bad arguments");
  CDN_CUDA(cudaSetDevice(e->device));
  CDN_CUDA(cudaDeviceSynchronize());
  const EngTensor& t = e->tensors[tensor];
  CDN_CUDA(cudaMemcpy(h_out, t.ptr, (size_t)batch * t.H * t.W * t.pitch, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cdn_engine_read_heads(cdn_engine* e, int batch, float* h_out) {
  ENG_CHECK(e);
  CDN_CHECK(e->finalized && batch > 0 && batch <= e->max_batch && h_out, CDN_ERR_INVALID, "read_heads: bad arguments");
  CDN_CUDA(cudaSetDevice(e->device));
  CDN_CUDA(cudaDeviceSynchronize());
  CDN_CUDA(cudaMemcpy(h_out, e->heads, (size_t)batch * e->n_f32 * e->hH * e->hW * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cdn_engine_num_launches(cdn_engine* e) { return e ? e->launches : 0; }
extern "C" int cdn_engine_heads_fused(cdn_engine* e) {
  if (!e || !e->finalized) return 0;
  for (size_t oi = 0; oi < e->ops.size(); ++oi) if (engine_fuses_heads(e, oi)) return 1;
  return 0;
}
extern "C" int cdn_engine_units_fused(cdn_engine* e) {       // how many stride-1 units run as one kernel
  if (!e || !e->finalized) return 0;
  int n = 0;
  for (size_t oi = 0; oi < e->ops.size(); ++oi) if (engine_fuses_unit(e, oi) || engine_fuses_unit_s2(e, oi)) { ++n; oi += 2; }
  return n;
}
// how plan op i runs: 0 = a launch of its own, 1 = first op of the fused heads tail, 2 = first op of a fused stride-1 unit,
// 3 = first op of the fused stride-2 branch,
// -1 = folded into the launch of an earlier op
extern "C" int cdn_engine_op_fusion(cdn_engine* e, int i) {
  if (!e || !e->finalized || i < 0 || i >= (int)e->ops.size()) return 0;
  for (size_t oi = 0; oi < e->ops.size(); ++oi) {
    const bool s2 = engine_fuses_unit_s2(e, oi);
    const int span = engine_fuses_heads(e, oi) ? 2 : ((engine_fuses_unit(e, oi) || s2) ? 3 : 1);
    if ((size_t)i >= oi && (size_t)i < oi + span) return (size_t)i == oi ? (span == 2 ? 1 : (span == 3 ? (s2 ? 3 : 2) : 0)) : -1;
    oi += span - 1;
  }
  return 0;
}
extern "C" int cdn_engine_requant_stats(cdn_engine* e, int* int_layers, int* guarded_layers) {
  ENG_CHECK(e);
  int ni = 0, ng = 0;
  for (const EngOp* op : e->ops) {
    if (op->kind == 1 || (op->kind == 2 && op->sc.mode == 0)) { if (op->dw.use_int) ++ni; else ++ng; }
    else if (op->kind == 2) ++ng;                      // bilinear: non-integer accumulators, guarded fp32 by construction
    else if (op->kind == 3 && op->pw.n_f32 == 0) { if (op->pw.use_int) ++ni; else ++ng; }
  }
  if (int_layers) *int_layers = ni;
  if (guarded_layers) *guarded_layers = ng;
  return 0;
}
