// Device helpers shared by the fused ShuffleNetV2 unit kernels (unit_fused.cu: stride 1, unit_s2_fused.cu: stride 2).
#pragma once
#include "layers.cuh"
#include "tc_ptx.cuh"

// `mid` tile: pixel p = r*ROWPIX + c at p*HP (ROWPIX * HP a multiple of 128), 16-byte unit u XOR-swizzled by the pixel's COLUMN so that (i) the epilogue's
// row-per-lane 16-byte stores and (ii) the stencil's pixel-per-(half-)warp word loads are both conflict-free, and (iii) a
// stencil thread's four pixel offsets are constants (the row advances by a multiple of 128 bytes).
template <int HP> __device__ __forceinline__ uint32_t uf_mid_swz(uint32_t c) {
  return HP == 128 ? (c & 7u) : ((((c >> 1) & 1u) << 2) | ((c >> 1) & 3u));   // HP = 64: bit 2 swaps the two 64-byte halves of a line
}
template <int HP> __device__ __forceinline__ uint32_t uf_mid_off(uint32_t p, uint32_t c, uint32_t u) {
  return (p * (uint32_t)HP + (u << 4)) ^ (uf_mid_swz<HP>(c) << 4);
}

// hi32(v * Mi + Bi): the requantisation of a channel whose shift is 0 (every channel of every CoDeNet layer: rq_int_solve
// tries the scale 2^32 first), one IMAD.HI
__device__ __forceinline__ int uf_rq_ns(int v, int Mi, long long Bi) {
  int hi;
  asm("{\n\t.reg .b64 t;\n\t.reg .b32 lo;\n\tmul.wide.s32 t, %1, %2;\n\tadd.s64 t, t, %3;\n\tmov.b64 {lo, %0}, t;\n\t}"
      : "=r"(hi) : "r"(v), "r"(Mi), "l"(Bi));
  return hi;
}
// FAST: shift 0 and no lower clamp (lo = -128 is the saturation); generic: RqInt with its shift, max(., lo)
template <bool FAST> __device__ __forceinline__ int uf_rq(int v, const uint4& k, int lo) {
  const long long Bi = (long long)(((unsigned long long)k.w << 32) | k.z);
  if (FAST) return uf_rq_ns(v, (int)k.x, Bi);
  return max(rq_int(v, (int)k.x, (int)k.y, Bi), lo);
}
template <bool FAST>
__device__ __forceinline__ uint32_t uf_rq_word(const int (&acc)[4], const int2 (&km)[4], const long long (&kb)[4]) {
  int q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = FAST ? uf_rq_ns(acc[c], km[c].x, kb[c]) : rq_int_hi(acc[c], km[c].x, km[c].y, kb[c]);
  return pack_sat4(q[0], q[1], q[2], q[3]);
}
__device__ __forceinline__ uint32_t uf_mask_word(uint32_t w, int rem) {
  return rem >= 4 ? w : (rem <= 0 ? 0u : (w & (0xffffffffu >> (8 * (4 - rem)))));
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}
__device__ __forceinline__ void sts_u128(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// mbarrier wait that suspends in hardware for up to ~20 us per try: all 256 threads wait for the tensor core / TMA here, and a
// plain try_wait loop cost 9 % of the kernel's issue slots in spin iterations
__device__ __forceinline__ void uf_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    if (done) break;
    if (++spins > (1u << 20)) { printf("cdn unit_fused: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}

// E2 of the fast path for the warps of group G (= first / second half of the unit's output channels): chunk j takes 8 new
// columns G*Gp + 8j and the 8 pass-through bytes G*PG + 8j and writes 16 interleaved bytes at G*HP + 16j; with PG (channels per
// group) a template constant every offset, shift and mask below is an immediate, and chunks go two at a time (one 16-column
// TMEM load, pass-through words shared between neighbours).
template <int HP, int PG, int G>
__device__ __forceinline__ void uf_e2_fast(uint32_t taddr, uint32_t prow, uint32_t srow, uint32_t x7s, uint32_t s_kc3) {
  constexpr int Gp = (PG + 7) & ~7, NCH = HP / 16;
#pragma unroll
  for (int j = 0; j < NCH; j += 2) {
    const int col = G * Gp + 8 * j, pass_off = G * PG + 8 * j, dst = G * HP + 16 * j;
    const int cnt0 = PG - 8 * j < 0 ? 0 : (PG - 8 * j > 8 ? 8 : PG - 8 * j), cnt1 = PG - 8 * j - 8 < 0 ? 0 : (PG - 8 * j - 8 > 8 ? 8 : PG - 8 * j - 8);
    uint32_t o[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (cnt0 > 0) {
      uint32_t acc[16];
      tmem_ld16(taddr + (uint32_t)col, acc);
      // 16 pass-through bytes from byte pass_off: aligned 8-byte words (each inside one 16-byte swizzle unit) + funnel shifts
      const int o8 = pass_off & ~7, b = pass_off & 7;
      uint32_t w[6];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        if (t == 2 && b == 0) { w[4] = 0u; w[5] = 0u; continue; }
        const int ob = o8 + 8 * t;
        const uint2 v = lds_u64(prow + ((((uint32_t)(ob >> 4)) << 4) ^ x7s) + (uint32_t)(ob & 8));
        w[2 * t] = v.x; w[2 * t + 1] = v.y;
      }
      uint32_t ps[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int wi = t + (b >> 2), sh = 8 * (b & 3);
        ps[t] = sh == 0 ? w[wi] : __funnelshift_r(w[wi], w[wi + 1], sh);
      }
      tmem_ld_wait();
      int v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i >= 8 && cnt1 == 0) { v[i] = 0; continue; }
        const uint4 k = lds_u128(s_kc3 + (uint32_t)(col + i) * 16u);
        v[i] = uf_rq<true>((int)acc[i], k, -128);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cnt = h ? cnt1 : cnt0;
        if (cnt == 0) continue;
        const uint32_t n_lo = pack_sat4(v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
        const uint32_t n_hi = pack_sat4(v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]);
        // out[2i] = pass[i], out[2i+1] = new[i]
        o[4 * h] = __byte_perm(ps[2 * h], n_lo, 0x5140); o[4 * h + 1] = __byte_perm(ps[2 * h], n_lo, 0x7362);
        o[4 * h + 2] = __byte_perm(ps[2 * h + 1], n_hi, 0x5140); o[4 * h + 3] = __byte_perm(ps[2 * h + 1], n_hi, 0x7362);
        if (cnt < 8) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[4 * h + t] = uf_mask_word(o[4 * h + t], 2 * cnt - 4 * t);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int d = dst + 16 * h;
      sts_u128(srow + (uint32_t)(d >> 7) * 16384u + ((((uint32_t)(d & 127) >> 4) << 4) ^ x7s), o[4 * h], o[4 * h + 1], o[4 * h + 2], o[4 * h + 3]);
    }
  }
}


// one lane of a converged warp (the lowest): single-thread tcgen05 / TMA issue without the per-instruction BRA.U.ANY loops the
// compiler wraps around them under a divergent `tid == 0` predicate
__device__ __forceinline__ bool uf_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// E1 of NB accumulator blocks at once: this lane's row of every block (TMEM address taddr[b], `mid` pixel address mpix[b] with
// swizzle term mswz[b]) is requantised with ONE fetch of the per-column constants -- the broadcast LDS.128 per column is a third of
// the epilogue's instructions when every block fetches its own.  Columns hf*16 + {0, 32, ...}; rows that are not `valid` are
// skipped, rows outside the image store the real zero `pad`; dump[b] (tests) is the row's pixel in the global copy of `mid` or null.
template <int HP, bool FAST, int NB>
__device__ __forceinline__ void uf_e1(const uint32_t (&taddr)[NB], const uint32_t (&mpix)[NB], const uint32_t (&mswz)[NB],
                                      const bool (&valid)[NB], const bool (&inside)[NB], int hf, uint32_t s_kc1, int lo, uint32_t pad,
                                      int8_t* const (&dump)[NB]) {
#pragma unroll
  for (int c0 = 0; c0 < HP; c0 += 32) {
    const uint32_t col = (uint32_t)(c0 + hf * 16);
    uint32_t acc[NB][16];
#pragma unroll
    for (int b = 0; b < NB; ++b) tmem_ld16(taddr[b] + col, acc[b]);
    tmem_ld_wait();
    const uint32_t kc = s_kc1 + col * 16u;
    uint32_t o[NB][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      int v[NB][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 k = lds_u128(kc + (uint32_t)(4 * g + i) * 16u);
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b][i] = uf_rq<FAST>((int)acc[b][4 * g + i], k, lo);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) o[b][g] = inside[b] ? pack_sat4(v[b][0], v[b][1], v[b][2], v[b][3]) : pad;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (valid[b]) {
        sts_u128((mpix[b] + col) ^ mswz[b], o[b][0], o[b][1], o[b][2], o[b][3]);
        if (dump[b]) *(uint4*)(dump[b] + col) = make_uint4(o[b][0], o[b][1], o[b][2], o[b][3]);
      }
  }
}
