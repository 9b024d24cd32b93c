// Micro-benchmark (GPU box only): issue cost of candidate requantisation sequences, in SM cycles per element-step
// per SMSP.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_requant ubench_requant.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define MAGIC_F 12582912.0f
#define MAGIC_I 0x4B400000
#define ITER 2048
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
#define NV 8

template <int V>
__global__ void __launch_bounds__(256) k(const int* __restrict__ in, uint32_t* out, float Mh, float Bh, float lo, float thr, int ab) {
  int acc[NV];
  for (int i = 0; i < NV; ++i) acc[i] = in[threadIdx.x + 256 * i] ;
  uint32_t sk[NV]; for (int i = 0; i < NV; ++i) sk[i] = 0; uint32_t sink = 0; bool bad = false; float dm0 = 0.f, dm1 = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int a = acc[i] + it;     // 1 IADD of loop overhead per element (varies the input)
      if (V == 0) {            // magic int->float, float clamp, magic round, TwoSum guard
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        float r = __fadd_rn(t, MAGIC_F);
        float kk = __fadd_rn(r, -MAGIC_F);
        bad |= fabsf(__fadd_rn(t, -kk)) > thr;
        sk[i] += __float_as_uint(r);
      } else if (V == 1) {     // I2F instead of magic
        float f = (float)(a + ab);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        float r = __fadd_rn(t, MAGIC_F);
        float kk = __fadd_rn(r, -MAGIC_F);
        bad |= fabsf(__fadd_rn(t, -kk)) > thr;
        sk[i] += __float_as_uint(r);
      } else if (V == 2) {     // FRND for the guard
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        float kk = rintf(t);
        bad |= fabsf(__fadd_rn(t, -kk)) > thr;
        sk[i] += __float_as_uint(__fadd_rn(kk, MAGIC_F));
      } else if (V == 3) {     // F2I.S8 saturating + FRND guard, ReLU max in float
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = fmaxf(__fmaf_rn(f, Mh, Bh), lo);
        int q; asm("cvt.rni.s8.f32 %0, %1;" : "=r"(q) : "f"(t));
        float kk = rintf(t);
        bad |= fabsf(__fadd_rn(t, -kk)) > thr;
        sk[i] += (uint32_t)q;
      } else if (V == 4) {     // no guard at all (lower bound)
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        sk[i] += __float_as_uint(__fadd_rn(t, MAGIC_F));
      } else if (V == 5) {     // exact fp64: magic int->double, mul, add, magic round, int clamp
        double d = __hiloint2double(0x43300000, (a + ab) ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
        double t = __dadd_rn(__dmul_rn(d, (double)Mh), (double)Bh);
        double r = __dadd_rn(t, 6755399441055744.0);                                          // 1.5 * 2^52
        int q = __double2loint(r);
        q = min(max(q, (int)lo), 127);
        sk[i] += (uint32_t)q;
      } else if (V == 6) {     // fixed point with 8 fraction bits from one FFMA (Bh pre-biased), integer round/guard
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = __fmaf_rn(f, Mh, Bh + 49152.0f);                 // 1.5*2^15: ulp 2^-8
        int v = __float_as_int(t) + 128 + 2;
        bad |= (unsigned)(v & 255) <= 4u;
        int q = min(max((v - 2) >> 8, (int)lo), 127);
        sk[i] += (uint32_t)q;
      } else if (V == 8) {     // V0 with the guard accumulated as a running max (two chains), one compare at the end
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        float r = __fadd_rn(t, MAGIC_F);
        float kk = __fadd_rn(r, -MAGIC_F);
        float d = fabsf(__fadd_rn(t, -kk));
        if (i & 1) dm1 = fmaxf(dm1, d); else dm0 = fmaxf(dm0, d);
        sk[i] += __float_as_uint(r);
      } else if (V == 9) {     // V8 with I2F
        float f = (float)(a + ab);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        float r = __fadd_rn(t, MAGIC_F);
        float kk = __fadd_rn(r, -MAGIC_F);
        float d = fabsf(__fadd_rn(t, -kk));
        if (i & 1) dm1 = fmaxf(dm1, d); else dm0 = fmaxf(dm0, d);
        sk[i] += __float_as_uint(r);
      } else if (V == 10) {    // V8 on packed f32x2 (FADD2 / FFMA2): two elements per FMA-pipe instruction
        if ((i & 1) == 0) {
          int a1 = acc[i + 1] + it;
          float2 f = fadd2(make_float2(__int_as_float(a + ab), __int_as_float(a1 + ab)), make_float2(-MAGIC_F, -MAGIC_F));
          float2 t = ffma2(f, make_float2(Mh, Mh), make_float2(Bh, Bh));
          t.x = fminf(fmaxf(t.x, lo), 127.f); t.y = fminf(fmaxf(t.y, lo), 127.f);
          float2 r = fadd2(t, make_float2(MAGIC_F, MAGIC_F));
          float2 kk = fadd2(r, make_float2(-MAGIC_F, -MAGIC_F));
          float2 d = ffma2(kk, make_float2(-1.f, -1.f), t);
          dm0 = fmaxf(dm0, fabsf(d.x)); dm1 = fmaxf(dm1, fabsf(d.y));
          sk[i] += __float_as_uint(r.x); sk[i + 1] += __float_as_uint(r.y);
        }
      } else if (V == 7) {     // V0 but guard via second magic (9 fraction bits) + integer test
        float f = __fadd_rn(__int_as_float(a + ab), -MAGIC_F);
        float t = __fmaf_rn(f, Mh, Bh);
        t = fminf(fmaxf(t, lo), 127.f);
        float r = __fadd_rn(t, MAGIC_F);
        int g = __float_as_int(__fadd_rn(t, 24576.0f));             // 1.5*2^14: ulp 2^-9
        bad |= (unsigned)((g + 2 - 256) & 511) <= 4u;
        sk[i] += __float_as_uint(r);
      }
    }
  }
  long long t1 = clock64();
  for (int i = 0; i < NV; ++i) sink += sk[i];
  if (bad || fmaxf(dm0, dm1) > thr) sink ^= 0x55;
  out[blockIdx.x * 256 + threadIdx.x] = sink;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * 256] = (uint32_t)(t1 - t0);
}

template <int V> void run(const char* name, int* in, uint32_t* out, int warps_per_smsp) {
  int threads = 32 * 4 * warps_per_smsp; if (threads > 256) threads = 256;
  int blocks_per_sm = (32 * 4 * warps_per_smsp) / threads;
  int grid = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<grid, threads>>>(in, out, 0.0123f, -3.7f, -128.f, 0.4999f, 1234);
  cudaEventRecord(e0);
  k<V><<<grid, threads>>>(in, out, 0.0123f, -3.7f, -128.f, 0.4999f, 1234);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  uint32_t cyc; cudaMemcpy(&cyc, out + grid * 256, 4, cudaMemcpyDeviceToHost);
  double steps = (double)ITER * NV * warps_per_smsp;   // element-steps per SMSP
  printf("%-34s warps/SMSP=%d  cycles/elt-step/SMSP (wall, 1965 MHz) = %.2f  (ms %.3f, err %s)\n", name, warps_per_smsp, ms * 1.965e6 / steps, ms,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int* in; uint32_t* out;
  cudaMalloc(&in, 256 * NV * 4); cudaMemset(in, 1, 256 * NV * 4);
  cudaMalloc(&out, (148 * 16 * 256 + 16) * 4);
  for (int w : {1, 2, 4, 8}) {
    run<0>("V0 magic + TwoSum guard", in, out, w);
    run<1>("V1 I2F + TwoSum guard", in, out, w);
    run<2>("V2 magic + FRND guard", in, out, w);
    run<3>("V3 F2I.S8 sat + FRND guard", in, out, w);
    run<4>("V4 no guard (lower bound)", in, out, w);
    run<5>("V5 exact fp64", in, out, w);
    run<6>("V6 fixed-point 8 frac bits", in, out, w);
    run<7>("V7 magic + 2nd-magic int guard", in, out, w);
    run<8>("V8 magic + TwoSum, max-accum guard", in, out, w);
    run<9>("V9 I2F + TwoSum, max-accum guard", in, out, w);
    run<10>("V10 V8 on packed f32x2", in, out, w);
  }
  return 0;
}
