// Shared device/host helpers of libcodenet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <string>
#include <mutex>
#include "../../include/codenet_b200.h"

// ---------------------------------------------------------------------------------------------------------
// error plumbing (no exceptions across the C ABI)
// ---------------------------------------------------------------------------------------------------------
int cdn_fail(int code, const char* fmt, ...);
#define CDN_CHECK(cond, code, ...) do { if (!(cond)) return cdn_fail(code, __VA_ARGS__); } while (0)
#define CDN_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    return cdn_fail(CDN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)
#define CDN_LAUNCH_CHECK(name) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) \
    return cdn_fail(CDN_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); } while (0)

extern unsigned g_cdn_debug_flags;
int cdn_num_sms();
// cudaFuncSetAttribute is per device: true the first time it is called for the current device with this flag array.
// Engines on different devices may be driven from different host threads: the flag tables are guarded by one mutex (a
// second caller may see "first" only after the first has returned from here, and setting an attribute twice is harmless).
std::mutex& cdn_attr_mutex();
inline bool cdn_first_on_device(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;     // unknown device: set the attributes again
  std::lock_guard<std::mutex> lk(cdn_attr_mutex());
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// Requantisation  q = clamp(rint(fl64(fl64(acc*M) + B)), lo, 127)   (DESIGN.md "requantisation")
//
// Fast path: t = fmaf((float)acc, Mh, Bh) in fp32, rounded half-to-even with the 1.5*2^23 trick.  |t - t_exact| is
// bounded per channel by `eps` (computed on the host from |B| and the int8 range), so whenever t is further than eps
// from a rounding boundary the fp32 result equals the fp64 one; otherwise the fp64 formula is evaluated.  The
// result is therefore bit-identical to the fp64 formula for every input.  Requires |acc| < 2^24 (checked on the host).
// ---------------------------------------------------------------------------------------------------------
struct RqFast { float Mh, Bh, thr; };        // thr = 0.5 - eps

// Host: derive the fast constants.
static inline RqFast rq_fast_from(double M, double B) {
  RqFast r;
  r.Mh = (float)M;
  r.Bh = (float)B;
  // |t_fp32 - t_fp64| <= 2^-24 (|acc*M| + |B| + |t|) with |acc*M| <= |t| + |B| and |t| <= 130 inside the clamp; x2 safety
  double eps = 2.0 * ldexp(1.0, -24) * (2.0 * fabs(B) + 2.0 * 130.0 + 1.0);
  if (eps > 0.49) eps = 0.49;
  r.thr = (float)(0.5 - eps);
  return r;
}

#define CDN_MAGIC_I_HOST 0x4B400000

// ---------------------------------------------------------------------------------------------------------
// Integer requantisation (the path the int8 kernels run):  q = sat8(max(((v*Mi + Bi) >> 32) >> sh, lo))
//
// v -> rint(fl64(fl64(v*M) + B)) is a monotone step function of the integer accumulator v, fully described by the 255
// accumulator values at which the result steps from k-1 to k.  The host finds those step positions with the fp64 formula
// (the definition of the result) and solves for a 31-bit multiplier Mi and a 64-bit offset Bi whose fixed-point line
// v*Mi + Bi crosses k * 2^(32+sh) at exactly the same integers, over the whole accumulator range the layer can produce.
// The 64-bit product is exact, so the device result equals the fp64 formula for EVERY accumulator in that range -- by
// construction, no guard band and no slow path: IMAD.HI + SHF (+ VIMNMX when lo > -128) + half an I2IP per element
// instead of the ~7 instructions of the guarded fp32 sequence below.  If no (Mi, Bi) exists for some channel (never
// observed; needs a multiplier >= 0.5 or step positions closer than 2^-31 relative) the layer keeps the guarded fp32
// sequence, which is exact by re-evaluation.
// ---------------------------------------------------------------------------------------------------------
struct RqInt { int32_t Mi; int32_t sh; long long Bi; };     // 16 bytes: one LDS.128 / LDG.128 per channel
// v in [vmin, vmax] (inclusive) is the accumulator INCLUDING acc_bias; returns false when no exact pair exists.
bool rq_int_solve(double M, double B, int lo, long long vmin, long long vmax, RqInt* out);
// the same constants re-based to a raw accumulator a = v - acc_bias (|a| <= amax); false on 64-bit overflow
bool rq_int_rebase(RqInt* r, long long acc_bias, long long amax);
#ifdef __CUDACC__
// Programmatic dependent launch: every engine kernel lets its successor start early (launch_dependents at entry); a
// successor launched with the programmatic-serialization attribute runs its prologue (barrier init, TMEM allocation,
// constant tables, weight loads) while this grid drains and calls pdl_wait() before touching activations.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define CDN_MAGIC_F 12582912.0f              // 1.5 * 2^23
#define CDN_MAGIC_I 0x4B400000

// ---- lean requantisation (v2 kernels) ----------------------------------------------------------------------
// One element costs IADD, FADD, FFMA, FMNMX, 3 FADD and two halves of 3-input FMNMX: the rounding-boundary guard
// and the upper clamp are accumulated as running maxima over a group of elements and tested ONCE per group
// (a serial predicate chain per element costs ~40% more issue cycles; tools/ubench_requant.cu).
//   v_magic = acc + acc_bias + MAGIC_I   (|acc + acc_bias| < 2^22, so the bits ARE the float 1.5*2^23 + value)
// Returns the float r = rint(clamp(t)) + 1.5*2^23 whose LOW BYTE is the int8 result, valid iff !rq_group_bad().
struct RqGuard { float d0, d1, tmax; };
__device__ __forceinline__ void rq_guard_init(RqGuard& g) { g.d0 = 0.f; g.d1 = 0.f; g.tmax = -3.0e38f; }
template <int PARITY>
__device__ __forceinline__ uint32_t rq_fast(int v_magic, float Mh, float Bh, float lo_f, RqGuard& g) {
  const float f = __fadd_rn(__int_as_float(v_magic), -CDN_MAGIC_F);
  float t = __fmaf_rn(f, Mh, Bh);
  t = fmaxf(t, lo_f);
  g.tmax = fmaxf(g.tmax, t);
  const float r = __fadd_rn(t, CDN_MAGIC_F);
  const float kk = __fadd_rn(r, -CDN_MAGIC_F);
  const float d = fabsf(__fadd_rn(t, -kk));
  if (PARITY) g.d1 = fmaxf(g.d1, d); else g.d0 = fmaxf(g.d0, d);
  return __float_as_uint(r);
}
// Two elements per FMA-pipe instruction with Blackwell's packed fp32 ops (FADD2 / FFMA2; `add/fma.rn.f32x2`): the same
// sequence on register pairs, 23% fewer issue cycles per element at saturation (tools/ubench_requant.cu V10 vs V8).
__device__ __forceinline__ float2 cdn_fadd2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 cdn_fmul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 cdn_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
// v0, v1: magic-biased accumulators of two elements; M2 = (Mh0, Mh1), B2 = (Bh0, Bh1).  Bit-identical to two rq_fast calls.
__device__ __forceinline__ void rq_fast2(int v0, int v1, float2 M2, float2 B2, float lo_f, RqGuard& g, uint32_t& r0, uint32_t& r1) {
  const float2 f = cdn_fadd2(make_float2(__int_as_float(v0), __int_as_float(v1)), make_float2(-CDN_MAGIC_F, -CDN_MAGIC_F));
  float2 t = cdn_ffma2(f, M2, B2);
  t.x = fmaxf(t.x, lo_f); t.y = fmaxf(t.y, lo_f);
  g.tmax = fmaxf(g.tmax, fmaxf(t.x, t.y));
  const float2 r = cdn_fadd2(t, make_float2(CDN_MAGIC_F, CDN_MAGIC_F));
  const float2 kk = cdn_fadd2(r, make_float2(-CDN_MAGIC_F, -CDN_MAGIC_F));
  const float2 d = cdn_ffma2(kk, make_float2(-1.0f, -1.0f), t);          // t - kk, exact (kk is an integer near t)
  g.d0 = fmaxf(g.d0, fabsf(d.x)); g.d1 = fmaxf(g.d1, fabsf(d.y));
  r0 = __float_as_uint(r.x); r1 = __float_as_uint(r.y);
}
// true when some element of the group was within eps of a rounding boundary or above the int8 range
__device__ __forceinline__ bool rq_group_bad(const RqGuard& g, float thr) {
  return fmaxf(g.d0, g.d1) > thr || g.tmax > 127.0f + thr;
}
// exact value of the same element: fl64(fl64(v*M) + B), clamped, rounded half-to-even
__device__ __forceinline__ uint32_t rq_exact(int v, double M, double B, float lo_f) {
  double td = __dadd_rn(__dmul_rn((double)v, M), B);
  td = fmin(fmax(td, (double)lo_f), 127.0);
  return __float_as_uint((float)__double2int_rn(td) + CDN_MAGIC_F);
}

// ---- integer requantisation (see RqInt above) ---------------------------------------------------------------
__device__ __forceinline__ int rq_int(int v, int Mi, int sh, long long Bi) {
  const long long x = (long long)v * Mi + Bi;              // IMAD.HI takes the 64-bit addend
  return (int)(x >> 32) >> sh;
}
// the same through an explicit mad.wide.s32: for code where the compiler expands the C form into a 64 x 64-bit multiply
// (six instructions; seen in the deformable kernel's lambdas).  IMAD.WIDE measured ~5% slower than IMAD.HI where both apply.
__device__ __forceinline__ int rq_int_wide(int v, int Mi, int sh, long long Bi) {
  long long x;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(x) : "r"(v), "r"(Mi), "l"(Bi));
  return (int)(x >> 32) >> sh;
}
// the same with the PTX spelled out as mul.wide.s32 + add.s64 + high half, the form ptxas turns into ONE IMAD.HI with a 64-bit
// addend: in loops where the front end hoists the sign extension of Mi it otherwise emits a 64-bit multiply (heads_fused.cu)
__device__ __forceinline__ int rq_int_hi(int v, int Mi, int sh, long long Bi) {
  int hi;
  asm("{\n\t.reg .b64 t;\n\t.reg .b32 lo;\n\tmul.wide.s32 t, %1, %2;\n\tadd.s64 t, t, %3;\n\tmov.b64 {lo, %0}, t;\n\t}"
      : "=r"(hi) : "r"(v), "r"(Mi), "l"(Bi));
  return hi >> sh;
}
// shift 0 (every channel of a layer whose DwDevice / PwDevice has sh0 set): the high word IS the result, one IMAD.HI
__device__ __forceinline__ int rq_int_hi0(int v, int Mi, long long Bi) {
  int hi;
  asm("{\n\t.reg .b64 t;\n\t.reg .b32 lo;\n\tmul.wide.s32 t, %1, %2;\n\tadd.s64 t, t, %3;\n\tmov.b64 {lo, %0}, t;\n\t}"
      : "=r"(hi) : "r"(v), "r"(Mi), "l"(Bi));
  return hi;
}
__device__ __forceinline__ int rq_int(int v, const int4& r) {
  return rq_int(v, r.x, r.y, (long long)(((unsigned long long)(uint32_t)r.w << 32) | (uint32_t)r.z));
}
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// element `off` (32-bit, unsigned) of a word array: base + 4*off as ONE IMAD.WIDE.U32
__device__ __forceinline__ const uint32_t* word_ptr(const uint32_t* base, uint32_t off) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"((unsigned long long)base));
  return (const uint32_t*)r;
}
__device__ __forceinline__ uint32_t* word_ptr(uint32_t* base, uint32_t off) {
  return const_cast<uint32_t*>(word_ptr((const uint32_t*)base, off));
}
// word `off` of the array whose 64-bit address is (hi:lo): one IMAD.WIDE.U32 on the FMA pipe
__device__ __forceinline__ const uint32_t* word_ptr64(uint32_t lo, uint32_t hi, uint32_t off) {
  unsigned long long b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(lo), "r"(hi));
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"(b));
  return (const uint32_t*)r;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
  uint4 r; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr)); return r;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t saddr) {
  uint2 r; asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr)); return r;
}
// four int32 -> four saturated int8 in one little-endian word (two I2IP.S8.S32.SAT)
__device__ __forceinline__ uint32_t pack_sat4(int q0, int q1, int q2, int q3) {
  uint32_t t, w;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(q3), "r"(q2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(q1), "r"(q0), "r"(t));
  return w;
}

// pack the low bytes of four rq_fast / rq_exact results into one little-endian word
__device__ __forceinline__ uint32_t pack4_lowbytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t ab = __byte_perm(a, b, 0x0040);   // [a.0, b.0, a.0, a.0] -> bytes0,1 used
  uint32_t cd = __byte_perm(c, d, 0x0040);
  return __byte_perm(ab, cd, 0x5410);
}

__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) {
  return __dp4a((int)a, (int)b, c);
}

// 4x4 byte transpose: in x0..x3 (word t = 4 channels of tap t); out y[c] = (x0[c], x1[c], x2[c], x3[c])
__device__ __forceinline__ void transpose4x4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                                             uint32_t& y0, uint32_t& y1, uint32_t& y2, uint32_t& y3) {
  uint32_t a = __byte_perm(x0, x1, 0x5140);  // x0.0 x1.0 x0.1 x1.1
  uint32_t b = __byte_perm(x0, x1, 0x7362);  // x0.2 x1.2 x0.3 x1.3
  uint32_t c = __byte_perm(x2, x3, 0x5140);
  uint32_t d = __byte_perm(x2, x3, 0x7362);
  y0 = __byte_perm(a, c, 0x5410);
  y1 = __byte_perm(a, c, 0x7632);
  y2 = __byte_perm(b, d, 0x5410);
  y3 = __byte_perm(b, d, 0x7632);
}
#endif

// ---------------------------------------------------------------------------------------------------------
// Device-side constant blocks owned by the library (uploaded once per layer)
// ---------------------------------------------------------------------------------------------------------
struct DevRequant {                          // arrays of length n (padded by the caller as needed)
  float* Mh = nullptr; float* Bh = nullptr; float* thr = nullptr;
  double* M = nullptr; double* B = nullptr; int32_t* acc_bias = nullptr;
  int lo = -128; int n = 0;
};
int dev_requant_upload(DevRequant& d, const cdn_requant* rq, const int32_t* acc_bias_host, int n_pad);
void dev_requant_free(DevRequant& d);

template <typename T> int dev_upload(T** dptr, const T* host, size_t n) {
  CDN_CUDA(cudaMalloc((void**)dptr, n * sizeof(T) > 0 ? n * sizeof(T) : 16));
  if (n) CDN_CUDA(cudaMemcpy(*dptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
