// Micro-benchmark (GPU box only): how fast can an epilogue warp fetch per-column (warp-uniform) constants?
// cycles per warp-instruction per SM for broadcast LDS.32/64/128, dynamic-index LDC.32/64 from __constant__ memory,
// and the integer requantisation fed from registers / shared memory / the constant bank.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_consts ubench_consts.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 512
#define NCOL 64
__constant__ int4 c_kc[1024];

__device__ __forceinline__ int rq_int(int v, int Mi, int sh, long long Bi) {
  long long x = (long long)v * Mi + Bi; return (int)(x >> 32) >> sh;
}
__device__ __forceinline__ uint32_t pack_sat4(int q0, int q1, int q2, int q3) {
  uint32_t t, w;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(q3), "r"(q2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(q1), "r"(q0), "r"(t));
  return w;
}

__device__ __forceinline__ int lds32(const void* p) { int r; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p))); return r; }
__device__ __forceinline__ int2 lds64(const void* p) { int2 r; asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(p))); return r; }
__device__ __forceinline__ int4 lds128(const void* p) { int4 r; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"((uint32_t)__cvta_generic_to_shared(p))); return r; }
// V: 0 LDS.32 bcast, 1 LDS.64 bcast, 2 LDS.128 bcast, 3 LDC.32 dyn, 4 LDC.64 dyn, 5 LDC.128(dyn, as compiled),
//    6 requant consts in regs, 7 requant consts via LDS.128, 8 requant consts via LDC(int4), 9 fp32 guarded w/ LDS.128+LDS.64 per pair (old)
template <int V>
__global__ void __launch_bounds__(512) k(const int* __restrict__ in, uint32_t* out, int colbase, int Mi, int sh, long long Bi) {
  __shared__ int4 s_kc[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_kc[i] = make_int4(Mi + i, sh, (int)Bi, (int)(Bi >> 32));
  __syncthreads();
  int acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = in[(threadIdx.x + 32 * i) & 1023];
  uint32_t sink = 0;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    const int col = (colbase + it * 16 + warp * 16) & 1008;     // warp-uniform, dynamic
    if (V == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sink += lds32(&s_kc[col + i]);
    } else if (V == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { int2 r = lds64(&s_kc[col + i]); sink += r.x ^ r.y; }
    } else if (V == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { int4 r = lds128(&s_kc[col + i]); sink += r.x ^ r.y ^ r.z ^ r.w; }
    } else if (V == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sink += ((const int*)c_kc)[(col + i) * 4 + (it & 1)];
    } else if (V == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { int2 r = ((const int2*)c_kc)[(col + i) * 2 + (it & 1)]; sink += r.x ^ r.y; }
    } else if (V == 5) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { int4 r = c_kc[(col + i + it) & 1023]; sink += r.x ^ r.y ^ r.z ^ r.w; }
    } else if (V == 6) {
      int q[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i] = rq_int(acc[i] + it, Mi + i, sh, Bi);
      sink += pack_sat4(q[0], q[1], q[2], q[3]) ^ pack_sat4(q[4], q[5], q[6], q[7]) ^ pack_sat4(q[8], q[9], q[10], q[11]) ^ pack_sat4(q[12], q[13], q[14], q[15]);
    } else if (V == 7) {
      int q[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { int4 r = s_kc[col + i]; q[i] = rq_int(acc[i] + it, r.x, r.y, ((long long)r.w << 32) | (uint32_t)r.z); }
      sink += pack_sat4(q[0], q[1], q[2], q[3]) ^ pack_sat4(q[4], q[5], q[6], q[7]) ^ pack_sat4(q[8], q[9], q[10], q[11]) ^ pack_sat4(q[12], q[13], q[14], q[15]);
    } else if (V == 8) {
      int q[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { int4 r = c_kc[col + i]; q[i] = rq_int(acc[i] + it, r.x, r.y, ((long long)r.w << 32) | (uint32_t)r.z); }
      sink += pack_sat4(q[0], q[1], q[2], q[3]) ^ pack_sat4(q[4], q[5], q[6], q[7]) ^ pack_sat4(q[8], q[9], q[10], q[11]) ^ pack_sat4(q[12], q[13], q[14], q[15]);
    } else if (V == 9) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        int4 a = s_kc[col + i]; int2 b = *(const int2*)&s_kc[col + i + 1];
        float t0_ = __fmaf_rn(__int_as_float(acc[i] + b.x + it), __int_as_float(a.x), __int_as_float(a.z));
        float t1_ = __fmaf_rn(__int_as_float(acc[i + 1] + b.y + it), __int_as_float(a.y), __int_as_float(a.w));
        sink += __float_as_int(t0_) ^ __float_as_int(t1_);
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (uint32_t)(t1 - t0);
}

template <int V> void run(const char* name, int* in, uint32_t* out, int per_iter) {
  const int threads = 512, grid = 148;           // 16 warps per SM, as the GEMM epilogue
  k<V><<<grid, threads>>>(in, out, 16, 1 << 30, 9, 123456789012345ll);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<V><<<grid, threads>>>(in, out, 16, 1 << 30, 9, 123456789012345ll);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  uint32_t cyc; cudaMemcpy(&cyc, out + grid * threads, 4, cudaMemcpyDeviceToHost);
  const double per_sm = (double)cyc / ((double)ITER * per_iter * 16);   // cycles per warp-level unit per SM
  printf("%-52s in-kernel cycles/unit/SM = %6.2f   (kernel %.3f ms, %s)\n", name, per_sm, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int* in; uint32_t* out;
  cudaMalloc(&in, 1024 * 4); cudaMemset(in, 1, 1024 * 4);
  cudaMalloc(&out, (148 * 512 + 16) * 4);
  int4 h[1024]; for (int i = 0; i < 1024; ++i) h[i] = make_int4((1 << 30) + i, 9, i, 28);
  cudaMemcpyToSymbol(c_kc, h, sizeof(h));
  run<0>("LDS.32 broadcast (unit = 1 load)", in, out, 16);
  run<1>("LDS.64 broadcast (unit = 1 load)", in, out, 16);
  run<2>("LDS.128 broadcast (unit = 1 load)", in, out, 16);
  run<3>("LDC.32 dynamic uniform index (unit = 1 load)", in, out, 16);
  run<4>("LDC.64 dynamic uniform index (unit = 1 load)", in, out, 16);
  run<5>("LDC int4 dynamic uniform index (unit = 1 int4)", in, out, 16);
  run<6>("int requant, consts in registers (unit = 1 column)", in, out, 16);
  run<7>("int requant, LDS.128 per column (unit = 1 column)", in, out, 16);
  run<8>("int requant, LDC int4 per column (unit = 1 column)", in, out, 16);
  run<9>("fp32 FFMA only, LDS.128+LDS.64 per pair (unit = 1 col)", in, out, 16);
  return 0;
}
