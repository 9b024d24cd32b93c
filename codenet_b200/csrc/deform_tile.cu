// Fused co-designed deformable depthwise conv with the INPUT TILE + HALO STAGED IN SHARED MEMORY BY TMA (int8 NHWC).
//
//   QuantDeformConvWithOffsetScaleBoundPositive.forward (portable_quantizer/quant_modules.py:668-671), integer-offset mode:
//   scale conv C->1 + Hardtanh + QuantAct(s) + round  ->  9-tap gather at (h + (i-1)s, w + (j-1)s)  ->  3x3 depthwise MAC  ->
//   QuantAct requantisation, one kernel, int8 in / int8 out.
//
// Why shared memory: deform_int_v3_kernel (dw.cu) gathers its 9 taps per output word straight from L1/L2 and ncu shows it
// waiting for them (long-scoreboard on the first use of the taps, 34 % warps active, DRAM at 4-11 % of peak): the kernel is
// bound by gather LATENCY, not bytes.  |s| <= bound (Hardtanh), so the taps of an output row band lie inside the band plus a
// halo of `bound` rows: a work item = (image, channel slice of <= 256 bytes per pixel, band of R output rows) whose stored
// input rows -- for the three CoDeNet layers the WHOLE stored image of the slice: 16x16x256 B = 64 KB, or 13 rows of 32x128 B
// behind the virtual x2 upsample -- arrive with ONE 4-D cp.async.bulk.tensor.  All gathers then are LDS (30 cycles, 128 B per
// clock and SM, conflict-free: a warp reads whole pixels), and the input is read from HBM exactly once.
//
//   phase A  thread j owns stored pixel j of the band: dot product of its slice bytes with the scale-conv weights (LDS.128,
//            chunk order rotated by j so that a quarter warp hits 8 distinct bank groups; no shuffles).  Layers wider than one
//            slice run as a THREAD-BLOCK CLUSTER over the channel slices: every CTA writes its partial dot products into all
//            peers' shared memory (DSMEM) and one cluster barrier later each CTA owns the full sums.  The integer offset
//            scalar is then a count of host-computed thresholds (deform_scale_build, dw.cu), and the thread writes the 9 tap
//            PIXEL INDICES of its (1 or 2x2) output pixels into a shared table; out-of-image taps index a pad pixel that
//            holds q = -zx (real zero), so the gather loop has no selects and no border branch.
//   phase C  a lane owns V channel words (V = 2: 8 channels, LDS.64 taps) with their 9 tap weights and requantisation
//            constants in registers; per pixel 3 LDS.128 (table) + 9 IMAD (addresses, FMA pipe) + 9 LDS + per word two byte
//            transposes, 12 dp4a, 4 x (IMAD.HI + SHF), 2 I2IP + one store; the taps of the next pixel are in flight while
//            this one is computed.
#include "layers.cuh"
#include "dw_params.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define DTL_TAB_STRIDE 48                      // 9 x u32 tap pixel indices + the unit's output pixel / block flag, padded to 48 bytes
#ifndef DTL_DEFAULT_VARIANT
#define DTL_DEFAULT_VARIANT 0
#endif

struct DefTParams {
  int Hs, Ws, H, W, shift;                   // stored input size, logical input (= output) size
  int pitch_out_w, cw_total;                 // output pitch in words; channel words of the layer (Cp / 4)
  int SLB, ns, R, reach, tile_rows;          // slice bytes, slices, output rows per band, halo, stored rows in the TMA box
  int nst_max;                               // stored pixels whose s one band computes (R >> shift) * Ws
  int lpp, ppw;                              // lanes per pixel, pixels per warp (phase C)
  uint32_t off_tab, off_part, off_ws, off_thr, off_bar, tile_bytes;
  const uint32_t *wA, *wB, *wC; const int4* ki; int lo_i;
  const uint32_t* ws; const int* s_thr; int s_n, s_lo;
  uint32_t pad_word;
  uint32_t* out; float* sval;
  int no_dedup;                              // A/B: every output pixel gathers for itself
  int dbg;                                   // timing experiments only (results are wrong): 1 no phase C, 2 no tap loads, 4 no MAC
};

__device__ __forceinline__ bool g_dtl_no_dedup(const DefTParams& p) { return p.no_dedup != 0; }

template <int V, int RQ, int NT, int MINB, bool PIPE>
__global__ void __launch_bounds__(NT, MINB) deform_tile_int_kernel(const __grid_constant__ CUtensorMap tmI, const DefTParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t dtl_smem_raw[];
  // the dynamic window starts at the same offset in every CTA of the cluster, so aligned offsets agree across peers
  const uint32_t s0 = (smem_u32(dtl_smem_raw) + 127u) & ~127u;
  uint8_t* const g0 = dtl_smem_raw + (s0 - smem_u32(dtl_smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = blockIdx.x, band = blockIdx.y, b = blockIdx.z;
  const int r0 = band * p.R;
  const int sr0 = max(r0 - p.reach, 0) >> p.shift;                  // first stored row of the tile
  const int rows_here = min(p.R, p.H - r0);                         // logical rows of this band
  const uint32_t bar = s0 + p.off_bar;
  const int pad_idx = p.tile_rows * p.Ws;                            // pixel index of the pad pixel (right behind the tile)
  int* const n_units = reinterpret_cast<int*>(g0 + p.off_bar + 8);
  if (tid == 0) {
    // the tile is requested before anything else: the constants below are fetched while it is in flight
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *n_units = 0;
    pdl_wait();
    mbar_expect_tx(bar, p.tile_bytes);
    tma_load_4d(s0, &tmI, cs * p.SLB, 0, sr0, b, bar);
  }
  if (p.ns > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // peers are running before any DSMEM store
  // ---- constants ----
  const int wps = p.SLB >> 2;                                        // words per slice
  for (int i = tid; i < wps; i += NT) {
    const int gw = cs * wps + i;
    reinterpret_cast<uint32_t*>(g0 + p.off_ws)[i] = gw < p.cw_total ? __ldg(p.ws + gw) : 0u;
    reinterpret_cast<uint32_t*>(g0 + (uint32_t)pad_idx * p.SLB)[i] = p.pad_word;
  }
  for (int i = tid; i < 128; i += NT) reinterpret_cast<int*>(g0 + p.off_thr)[i] = i < p.s_n ? __ldg(p.s_thr + i) : 0x7fffffff;
  const int cl = lane % p.lpp, sub = lane / p.lpp;
  const int lw0 = cl * V;                                            // first word of this lane inside the slice
  const bool lane_on = lw0 < wps && sub < p.ppw;
  uint32_t wA[V][4], wB[V][4], wC[V][4]; int Mi[V][4], sh[V][4]; long long Bi[V][4];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int gw = min(cs * wps + lw0 + v, p.cw_total - 1);          // inactive lanes read valid constants and never store
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ch = gw * 4 + c;
      wA[v][c] = __ldg(p.wA + ch); wB[v][c] = __ldg(p.wB + ch); wC[v][c] = __ldg(p.wC + ch);
      const int2 ms = __ldg(reinterpret_cast<const int2*>(p.ki + ch));
      Mi[v][c] = ms.x; sh[v][c] = ms.y; Bi[v][c] = __ldg(reinterpret_cast<const long long*>(p.ki + ch) + 1);
    }
  }
  const bool word_on[2] = {lane_on && cs * wps + lw0 < p.cw_total, V == 2 && lane_on && lw0 + 1 < wps && cs * wps + lw0 + 1 < p.cw_total};
  __syncthreads();                                                   // barrier init, pad pixel, scale weights, thresholds
  pdl_wait();
  mbar_wait(bar, 0);
  if (p.ns > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");             // completes the arrive at kernel entry
  // ---------------- phase A: partial dot products of the band's stored pixels ----------------
  const int st_r0 = r0 >> p.shift;                                   // first stored row whose s this band needs
  const int nst = ((rows_here + (1 << p.shift) - 1) >> p.shift) * p.Ws;
  const int nchunk = p.SLB >> 4;
  int* const part_own = reinterpret_cast<int*>(g0 + p.off_part);
  for (int j = tid; j < nst; j += NT) {
    const int srow = j / p.Ws, scol = j - srow * p.Ws;
    const uint32_t px = s0 + (uint32_t)(((st_r0 + srow - sr0) * p.Ws + scol) * p.SLB);
    int acc = 0;
    for (int k = 0; k < nchunk; ++k) {
      int ch = j + k; ch -= (ch / nchunk) * nchunk;
      const uint4 x = lds_u128(px + 16u * ch), w = lds_u128(s0 + p.off_ws + 16u * ch);
      acc = dp4a_ss(x.x, w.x, acc); acc = dp4a_ss(x.y, w.y, acc); acc = dp4a_ss(x.z, w.z, acc); acc = dp4a_ss(x.w, w.w, acc);
    }
    if (p.ns > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      for (int r = 0; r < p.ns; ++r) cluster.map_shared_rank(part_own, r)[cs * p.nst_max + j] = acc;
    } else part_own[j] = acc;
  }
  if (p.ns > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // ---------------- offset scalar + work units ----------------
  // A unit = one output pixel, or (behind the virtual x2 upsample, s even) a 2x2 block of output pixels: the block shares
  // its stored neighbourhood, and with an even s all four pixels sample the same nine stored pixels -- one gather, four stores.
  const int* const thr = reinterpret_cast<const int*>(g0 + p.off_thr);
  for (int j = tid; j < nst; j += NT) {
    int v = 0;
    for (int r = 0; r < p.ns; ++r) v += part_own[r * p.nst_max + j];
    int cnt = 0;
#pragma unroll
    for (int step = 64; step > 0; step >>= 1) if (v >= thr[cnt + step - 1]) cnt += step;
    const int si = p.s_lo + cnt;
    const int srow = j / p.Ws, scol = j - srow * p.Ws;
    const int h0 = r0 + (srow << p.shift), w0 = scol << p.shift;
    const bool block = p.shift == 1 && !(si & 1) && h0 + 1 < p.H && w0 + 1 < p.W && !(g_dtl_no_dedup(p));
    const int nrep = block ? 1 : (1 << p.shift);
    const int ny = min(nrep, p.H - h0), nx = min(nrep, p.W - w0);
    int u = atomicAdd(n_units, ny * nx);
    for (int dy = 0; dy < ny; ++dy)
      for (int dx = 0; dx < nx; ++dx) {
        const int h = h0 + dy, w = w0 + dx;
        int yi[3], xi[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int y = h + (i - 1) * si, x = w + (i - 1) * si;
          yi[i] = (unsigned)y < (unsigned)p.H ? ((y >> p.shift) - sr0) * p.Ws : -1;
          xi[i] = (unsigned)x < (unsigned)p.W ? (x >> p.shift) : -1;
        }
        uint32_t* t = reinterpret_cast<uint32_t*>(g0 + p.off_tab + (uint32_t)u * DTL_TAB_STRIDE);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) t[i * 3 + jj] = (yi[i] < 0 || xi[jj] < 0) ? (uint32_t)pad_idx : (uint32_t)(yi[i] + xi[jj]);
        t[9] = (uint32_t)((h - r0) * p.W + w) | (block ? 0x80000000u : 0u);
        ++u;
      }
    if (p.sval != nullptr && cs == 0)
      for (int dy = 0; dy < (1 << p.shift); ++dy)
        for (int dx = 0; dx < (1 << p.shift); ++dx)
          if (h0 + dy < p.H && w0 + dx < p.W) p.sval[((size_t)b * p.H + h0 + dy) * p.W + w0 + dx] = (float)si;
  }
  __syncthreads();
  // ---------------- phase C: gather from shared memory, MAC, requantise, store ----------------
  if (!lane_on || (p.dbg & 1)) return;
  const int nu = *n_units;
  const uint32_t lane_base = s0 + (uint32_t)lw0 * 4u;
  const uint32_t tab0 = s0 + p.off_tab;
  uint32_t* const out_l = p.out + ((size_t)b * p.H + r0) * p.W * (size_t)p.pitch_out_w + cs * wps + lw0;
  const uint32_t slb = (uint32_t)p.SLB;
  auto fetch = [&](int u, uint32_t (&x)[V][9], uint32_t& meta) {
    const uint32_t ta = tab0 + (uint32_t)u * DTL_TAB_STRIDE;
    const uint4 q0 = lds_u128(ta), q1 = lds_u128(ta + 16);
    const uint2 q2 = lds_u64(ta + 32);
    const uint32_t idx[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
    meta = q2.y;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const uint32_t a = idx[t] * slb + lane_base;
      if (p.dbg & 2) { x[0][t] = a; x[V - 1][t] = a * 3u; continue; }
      if (V == 2) { const uint2 r = lds_u64(a); x[0][t] = r.x; x[V - 1][t] = r.y; }
      else x[0][t] = lds_u32(a);
    }
  };
  auto compute = [&](uint32_t meta, const uint32_t (&x)[V][9]) {
    uint32_t o[V];
    if (p.dbg & 4) {
#pragma unroll
      for (int v = 0; v < V; ++v) { o[v] = 0; for (int t = 0; t < 9; ++t) o[v] ^= x[v][t]; }
    } else
#pragma unroll
    for (int v = 0; v < V; ++v) {
      uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
      transpose4x4(x[v][0], x[v][1], x[v][2], x[v][3], a0, a1, a2, a3);
      transpose4x4(x[v][4], x[v][5], x[v][6], x[v][7], b0, b1, b2, b3);
      int acc[4], q[4];
      acc[0] = dp4a_ss(x[v][8], wC[v][0], dp4a_ss(b0, wB[v][0], dp4a_ss(a0, wA[v][0], 0)));
      acc[1] = dp4a_ss(x[v][8], wC[v][1], dp4a_ss(b1, wB[v][1], dp4a_ss(a1, wA[v][1], 0)));
      acc[2] = dp4a_ss(x[v][8], wC[v][2], dp4a_ss(b2, wB[v][2], dp4a_ss(a2, wA[v][2], 0)));
      acc[3] = dp4a_ss(x[v][8], wC[v][3], dp4a_ss(b3, wB[v][3], dp4a_ss(a3, wA[v][3], 0)));
#pragma unroll
      for (int c = 0; c < 4; ++c) { q[c] = RQ == 3 ? rq_int_hi0(acc[c], Mi[v][c], Bi[v][c]) : rq_int_hi(acc[c], Mi[v][c], sh[v][c], Bi[v][c]); if (RQ == 2) q[c] = max(q[c], p.lo_i); }   // RQ 3: shift 0 in every channel
      o[v] = pack_sat4(q[0], q[1], q[2], q[3]);
    }
    const uint32_t pl = meta & 0x7fffffffu;
    uint32_t* dst = word_ptr(out_l, pl * (uint32_t)p.pitch_out_w);
    auto put = [&](uint32_t* d) {
      if (V == 2 && word_on[1]) *reinterpret_cast<uint2*>(d) = make_uint2(o[0], o[V - 1]);
      else if (word_on[0]) *d = o[0];
    };
    put(dst);
    if (meta & 0x80000000u) {                  // 2x2 block with a common gather
      put(dst + p.pitch_out_w);
      uint32_t* d2 = word_ptr(dst, (uint32_t)p.W * (uint32_t)p.pitch_out_w);
      put(d2); put(d2 + p.pitch_out_w);
    }
  };
  const int pstep = (NT / 32) * p.ppw;
  int u = warp * p.ppw + sub;
  if (PIPE) {
    uint32_t xa[V][9], xb[V][9], ma = 0, mb = 0;
    if (u < nu) fetch(u, xa, ma);
#pragma unroll 1
    for (; u < nu; u += 2 * pstep) {
      const int u1 = u + pstep, u2 = u + 2 * pstep;
      if (u1 < nu) fetch(u1, xb, mb);
      compute(ma, xa);
      if (u2 < nu) fetch(u2, xa, ma);
      if (u1 < nu) compute(mb, xb);
    }
  } else {
    // no software pipeline (fewer registers, twice the resident warps): the other warps of the scheduler cover the LDS latency
#pragma unroll 1
    for (; u < nu; u += pstep) {
      uint32_t xa[V][9], ma;
      fetch(u, xa, ma);
      compute(ma, xa);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Bilinear offsets (the reference's default: fractional s, quant_modules.py:668-671) on the same staged tile.
//
// The 36 corners of the 9 taps lie on a 5 x 5 lattice of stored pixels (rows hl0, hl0+1, h, hl2, hl2+1 and the same for
// columns; the centre row / column is integral); the blend is separable and runs in fp32, two channels per instruction
// (FADD2 / FMUL2 / FFMA2), exactly the operation sequence of deform_dw_v2_kernel<1> (dw.cu), so the host-derived error
// bound `thr_bil` applies unchanged: words that come within it of a rounding boundary are queued and re-evaluated by all
// threads at the end of the tile with the fp64 code that follows dcn_deform_conv_cuda_kernel.cu:83-114,210-227 on exact
// integers (deform_bilinear_exact_word, from global memory) -- bit-exact against the fp64 oracle.
// The tile carries one extra pixel column and one extra pixel row holding q = -zx (real zero): out-of-image lattice rows /
// columns index them, so address = row offset + column offset with no validity test in the gather loop.
// ---------------------------------------------------------------------------------------------------------
#define DTB_ENT 80                             // bytes per unit: 5 row offsets, 5 column offsets, 4 fractions, pixel, s (double)
#define DTB_QCAP 1024

struct DefTBil { const float2* mb; float lo_f, thr_bil; double Ms, bs, ss, zs, u_lo, u_hi; long long acc_s_bias; };

template <int NT>
__global__ void __launch_bounds__(NT, 2) deform_tile_bil_kernel(const __grid_constant__ CUtensorMap tmI, const DefTParams p, const DefTBil f,
                                                                const DwParams dwp) {
  pdl_launch_dependents();
  extern __shared__ uint8_t dtl_smem_raw[];
  const uint32_t s0 = (smem_u32(dtl_smem_raw) + 127u) & ~127u;
  uint8_t* const g0 = dtl_smem_raw + (s0 - smem_u32(dtl_smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = blockIdx.x, band = blockIdx.y, b = blockIdx.z;
  const int r0 = band * p.R;
  const int sr0 = max(r0 - p.reach, 0) >> p.shift;
  const int rows_here = min(p.R, p.H - r0);
  const uint32_t bar = s0 + p.off_bar;
  const int rowpx = p.Ws + 1;                                         // pixels per staged row (the last one is the pad column)
  int* const q_n = reinterpret_cast<int*>(g0 + p.off_bar + 8);
  uint32_t* const q_ent = reinterpret_cast<uint32_t*>(g0 + p.off_thr);   // the exact-evaluation queue lives where the int kernel keeps its thresholds
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *q_n = 0;
    pdl_wait();
    mbar_expect_tx(bar, p.tile_bytes);
    tma_load_4d(s0, &tmI, cs * p.SLB, 0, sr0, b, bar);
  }
  if (p.ns > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  const int wps = p.SLB >> 2;
  for (int i = tid; i < wps; i += NT) {
    const int gw = cs * wps + i;
    reinterpret_cast<uint32_t*>(g0 + p.off_ws)[i] = gw < p.cw_total ? __ldg(p.ws + gw) : 0u;
  }
  const int cl = lane % p.lpp, sub = lane / p.lpp;
  const bool lane_on = cl < wps && sub < p.ppw && cs * wps + cl < p.cw_total;
  const int gwl = min(cs * wps + cl, p.cw_total - 1);
  float wf[4][9], Mh[4], Bh[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int ch = gwl * 4 + c;
    const uint32_t a = __ldg(p.wA + ch), bb = __ldg(p.wB + ch), cc = __ldg(p.wC + ch);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const uint32_t wword = t < 4 ? a : (t < 8 ? bb : cc);
      const int sh = t < 8 ? 8 * (t & 3) : 8 * c;
      wf[c][t] = (float)(int)(int8_t)((wword >> sh) & 0xff);
    }
    const float2 mb = __ldg(f.mb + ch);
    Mh[c] = mb.x; Bh[c] = mb.y;
  }
  __syncthreads();
  pdl_wait();
  mbar_wait(bar, 0);
  if (p.ns > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  // pad column of every staged row and the pad row behind the tile: q = -zx
  {
    const int npad = p.tile_rows + rowpx;                            // pixels to fill
    for (int i = tid; i < npad * wps; i += NT) {
      const int px = i / wps, wd = i - px * wps;
      const int pix = px < p.tile_rows ? px * rowpx + p.Ws : p.tile_rows * rowpx + (px - p.tile_rows);
      reinterpret_cast<uint32_t*>(g0 + (size_t)pix * p.SLB)[wd] = p.pad_word;
    }
  }
  // ---------------- phase A ----------------
  const int st_r0 = r0 >> p.shift;
  const int nst = ((rows_here + (1 << p.shift) - 1) >> p.shift) * p.Ws;
  const int nchunk = p.SLB >> 4;
  int* const part_own = reinterpret_cast<int*>(g0 + p.off_part);
  for (int j = tid; j < nst; j += NT) {
    const int srow = j / p.Ws, scol = j - srow * p.Ws;
    const uint32_t px = s0 + (uint32_t)(((st_r0 + srow - sr0) * rowpx + scol) * p.SLB);
    int acc = 0;
    for (int k = 0; k < nchunk; ++k) {
      int ch = j + k; ch -= (ch / nchunk) * nchunk;
      const uint4 x = lds_u128(px + 16u * ch), w = lds_u128(s0 + p.off_ws + 16u * ch);
      acc = dp4a_ss(x.x, w.x, acc); acc = dp4a_ss(x.y, w.y, acc); acc = dp4a_ss(x.z, w.z, acc); acc = dp4a_ss(x.w, w.w, acc);
    }
    if (p.ns > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      for (int r = 0; r < p.ns; ++r) cluster.map_shared_rank(part_own, r)[cs * p.nst_max + j] = acc;
    } else part_own[j] = acc;
  }
  if (p.ns > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t pad_row_off = (uint32_t)(p.tile_rows * rowpx) * (uint32_t)p.SLB, pad_col_off = (uint32_t)p.Ws * (uint32_t)p.SLB;
  for (int j = tid; j < nst; j += NT) {
    long long v = 0;
    for (int r = 0; r < p.ns; ++r) v += part_own[r * p.nst_max + j];
    // fp64, product and sum rounded separately as in the oracle (dcn_deform_conv.py:295-330, quant_modules.py:648-653)
    double u = __dadd_rn(__dmul_rn((double)(v + f.acc_s_bias), f.Ms), f.bs);
    u = fmin(fmax(u, f.u_lo), f.u_hi);
    const double qs = rint(__dsub_rn(__dmul_rn(f.ss, u), f.zs));
    const double sv = __ddiv_rn(__dadd_rn(qs, f.zs), f.ss);
    const double dd = __dsub_rn(sv, 1.0);
    const int srow = j / p.Ws, scol = j - srow * p.Ws;
    const int nrep = 1 << p.shift;
    for (int dy = 0; dy < nrep; ++dy)
      for (int dx = 0; dx < nrep; ++dx) {
        const int h = r0 + (srow << p.shift) + dy, w = (scol << p.shift) + dx;
        if (h >= p.H || w >= p.W) continue;
        const int pl = (h - r0) * p.W + w;
        uint32_t* t = reinterpret_cast<uint32_t*>(g0 + p.off_tab + (uint32_t)pl * DTB_ENT);
        int R[5], Cc[5];
        R[2] = h; Cc[2] = w;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double sg = e ? 1.0 : -1.0;
          const double him = __dadd_rn((double)(h - 1 + 2 * e), sg * dd), wim = __dadd_rn((double)(w - 1 + 2 * e), sg * dd);
          const double hf = floor(him), wfl = floor(wim);
          R[3 * e] = (int)hf; R[3 * e + 1] = (int)hf + 1; Cc[3 * e] = (int)wfl; Cc[3 * e + 1] = (int)wfl + 1;
          t[10 + e] = __float_as_uint((float)__dsub_rn(him, hf));
          t[12 + e] = __float_as_uint((float)__dsub_rn(wim, wfl));
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          t[k] = (unsigned)R[k] < (unsigned)p.H ? (uint32_t)(((R[k] >> p.shift) - sr0) * rowpx) * (uint32_t)p.SLB : pad_row_off;
          t[5 + k] = (unsigned)Cc[k] < (unsigned)p.W ? (uint32_t)(Cc[k] >> p.shift) * (uint32_t)p.SLB : pad_col_off;
        }
        t[14] = (uint32_t)pl;
        *reinterpret_cast<double*>(t + 16) = sv;
        if (p.sval != nullptr && cs == 0) p.sval[((size_t)b * p.H + h) * p.W + w] = (float)sv;
      }
  }
  __syncthreads();
  // ---------------- phase C ----------------
  const int npx = rows_here * p.W;
  const uint32_t lane_base = s0 + (uint32_t)cl * 4u;
  const uint32_t tab0 = s0 + p.off_tab;
  uint32_t* const out_l = p.out + ((size_t)b * p.H + r0) * p.W * (size_t)p.pitch_out_w + cs * wps + cl;
  const int zx = -(int)(int8_t)(p.pad_word & 0xff);
  const float unb = 8388608.0f + 128.0f - (float)zx;                 // as_float(0x4B000000 | (q ^ 0x80)) - unb = q + zx
  const float2 nunb = make_float2(-unb, -unb);
  const int pstep = (NT / 32) * p.ppw;
  if (lane_on) {
#pragma unroll 1
    for (int u = warp * p.ppw + sub; u < npx; u += pstep) {
      const uint32_t ta = tab0 + (uint32_t)u * DTB_ENT;
      const uint4 e0 = lds_u128(ta), e1 = lds_u128(ta + 16), e2 = lds_u128(ta + 32);
      const uint2 e3 = lds_u64(ta + 48);
      const uint32_t rowB[5] = {e0.x, e0.y, e0.z, e0.w, e1.x};
      const uint32_t colB[5] = {e1.y + lane_base, e1.z + lane_base, e1.w + lane_base, e2.x + lane_base, e2.y + lane_base};
      const float lh0 = __uint_as_float(e2.z), lh2 = __uint_as_float(e2.w), lw0 = __uint_as_float(e3.x), lw2 = __uint_as_float(e3.y);
      const float rho[5] = {1.0f - lh0, lh0, 1.0f, 1.0f - lh2, lh2};
      const float cwt[4] = {1.0f - lw0, lw0, 1.0f - lw2, lw2};
      float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        uint32_t xw[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) xw[q] = lds_u32(rowB[r] + colB[q]) ^ 0x80808080u;
        const int ti = r < 2 ? 0 : (r == 2 ? 1 : 2);                 // tap row fed by this lattice row
        const float2 rho2 = make_float2(rho[r], rho[r]);
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
          const int c = 2 * cp;
          float2 a[5];
#pragma unroll
          for (int q = 0; q < 5; ++q)
            a[q] = cdn_fadd2(make_float2(__uint_as_float(__byte_perm(xw[q], 0x4B000000u, 0x7650 + c)),
                                         __uint_as_float(__byte_perm(xw[q], 0x4B000000u, 0x7650 + c + 1))), nunb);
          const float2 u0 = cdn_ffma2(make_float2(cwt[1], cwt[1]), a[1], cdn_fmul2(make_float2(cwt[0], cwt[0]), a[0]));
          const float2 u2 = cdn_ffma2(make_float2(cwt[3], cwt[3]), a[4], cdn_fmul2(make_float2(cwt[2], cwt[2]), a[3]));
          const float2 w0 = make_float2(wf[c][ti * 3], wf[c + 1][ti * 3]), w1 = make_float2(wf[c][ti * 3 + 1], wf[c + 1][ti * 3 + 1]),
                       w2 = make_float2(wf[c][ti * 3 + 2], wf[c + 1][ti * 3 + 2]);
          const float2 t = cdn_ffma2(w2, u2, cdn_ffma2(w1, a[2], cdn_fmul2(w0, u0)));
          acc2[cp] = cdn_ffma2(rho2, t, acc2[cp]);
        }
      }
      const float acc[4] = {acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y};
      RqGuard gd; rq_guard_init(gd);
      uint32_t rr[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t = fmaf(acc[c], Mh[c], Bh[c]);
        t = fmaxf(t, f.lo_f);
        gd.tmax = fmaxf(gd.tmax, t);
        const float r_ = __fadd_rn(t, CDN_MAGIC_F), kk = __fadd_rn(r_, -CDN_MAGIC_F);
        gd.d0 = fmaxf(gd.d0, fabsf(__fadd_rn(t, -kk)));
        rr[c] = __float_as_uint(r_);
      }
      const uint32_t oword = pack4_lowbytes(rr[0], rr[1], rr[2], rr[3]);
      bool queued = false;
      if (rq_group_bad(gd, f.thr_bil)) {
        const int pos = atomicAdd(q_n, 1);
        if (pos < DTB_QCAP) { q_ent[pos] = ((uint32_t)u << 8) | (uint32_t)cl; queued = true; }
      }
      if (!queued) {
        uint32_t* dst = word_ptr(out_l, (uint32_t)u * (uint32_t)p.pitch_out_w);
        if (rq_group_bad(gd, f.thr_bil)) {                           // queue overflow: evaluate here
          const int h = r0 + u / p.W, w = u - (u / p.W) * p.W;
          const double sv = *reinterpret_cast<const double*>(g0 + p.off_tab + (uint32_t)u * DTB_ENT + 64);
          *dst = deform_bilinear_exact_word(dwp, dwp.in + (size_t)b * dwp.Hs * dwp.Ws * dwp.in_pitch_w + gwl, h, w, sv, gwl);
        } else *dst = oword;
      }
    }
  }
  __syncthreads();
  // ---------------- exact re-evaluation of the queued words ----------------
  const int nq = min(*q_n, DTB_QCAP);
  for (int i = tid; i < nq; i += NT) {
    const uint32_t e = q_ent[i];
    const int u = (int)(e >> 8), lw = (int)(e & 0xffu);
    const int gw = cs * wps + lw;
    const int h = r0 + u / p.W, w = u - (u / p.W) * p.W;
    const double sv = *reinterpret_cast<const double*>(g0 + p.off_tab + (uint32_t)u * DTB_ENT + 64);
    p.out[(((size_t)b * p.H + h) * p.W + w) * (size_t)p.pitch_out_w + gw] =
        deform_bilinear_exact_word(dwp, dwp.in + (size_t)b * dwp.Hs * dwp.Ws * dwp.in_pitch_w + gw, h, w, sv, gw);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// kernel variants (A/B through cdn_set_debug_flags): 0 = 8 channels per lane, 256 threads, 2 CTAs per SM, software pipeline;
// 1 = 4 channels per lane, 512 threads, 64 registers, no software pipeline (twice the warps); 2 = 4 channels per lane, 256
// threads, 3 CTAs per SM
struct DefTPlan { int ok, variant, V, NT, SLB, ns, R, nbands, tile_rows, lpp, ppw; uint32_t off_tab, off_part, off_ws, off_thr, off_bar, tile_bytes; size_t smem; };

// Chooses slice, band height and the shared-memory layout; ok = 0 when the shape does not fit (the caller keeps deform_int_v3_kernel).
static DefTPlan deform_tile_plan(int cw_total, int in_pitch, int H, int W, int in_shift, int bound, int batch, int mode) {
  const int reach = mode == 1 ? bound + 1 : bound;              // the high corner of a fractional tap reaches one pixel further
  DefTPlan P; memset(&P, 0, sizeof(P));
  const int Hs = H >> in_shift, Ws = W >> in_shift;
  const int cp = cw_total * 4;                                       // channel bytes the layer processes
  P.SLB = cp >= 256 ? 256 : cp;
  if (P.SLB % 16 || in_pitch % 16 || Ws + (mode == 1 ? 1 : 0) > 256 || Ws < 1) return P;
  P.ns = (cp + P.SLB - 1) / P.SLB;
  if (P.ns > 8) return P;                                            // portable cluster size
  P.variant = DTL_DEFAULT_VARIANT;
  if (g_cdn_debug_flags & 4096u) P.variant = 1;                     // bits 12 / 13: kernel variant (A/B)
  if (g_cdn_debug_flags & 8192u) P.variant = 2;
  if ((g_cdn_debug_flags & 12288u) == 12288u) P.variant = 0;
  if (P.SLB < 64 && P.variant == 0) P.variant = 2;
  if (mode == 1) P.variant = 3;                                     // bilinear kernel: 4 channels per lane, 256 threads, 2 CTAs per SM
  P.V = P.variant == 0 ? 2 : 1;
  P.NT = P.variant == 1 ? 512 : 256;
  const int wps = P.SLB / 4, lanes = (wps + P.V - 1) / P.V;
  P.lpp = 1; while (P.lpp < lanes) P.lpp <<= 1;
  if (P.lpp > 32) return P;
  P.ppw = 32 / P.lpp;
  const size_t budget = (P.variant == 2 ? 73 : 110) * 1024;
  const int per_sm = P.variant == 2 ? 3 : 2;
  const int step = 1 << in_shift;
  for (int k = 1; k <= H; ++k) {
    int R = (H + k - 1) / k; R = (R + step - 1) / step * step;
    if (k > 1 && R == P.R) continue;
    // stored rows a band can need: logical rows [r0 - reach, r0 + R - 1 + reach]
    int rows = ((R - 1 + 2 * reach) >> in_shift) + 2;
    rows = std::min(rows, Hs);
    if (rows > 256) continue;
    const size_t tile = (size_t)rows * (Ws + (mode == 1 ? 1 : 0)) * P.SLB;   // bytes the TMA box delivers
    size_t off = mode == 1 ? (size_t)(rows + 1) * (Ws + 1) * P.SLB : tile + P.SLB;   // + pad row / pad pixel
    off = (off + 15) / 16 * 16; const size_t off_tab = off; off += (size_t)R * W * (mode == 1 ? DTB_ENT : DTL_TAB_STRIDE);
    const int nst_max = (R >> in_shift) * Ws;
    const size_t off_part = off; off += (size_t)P.ns * nst_max * 4;
    off = (off + 15) / 16 * 16; const size_t off_ws = off; off += P.SLB;
    const size_t off_thr = off; off += mode == 1 ? DTB_QCAP * 4 : 128 * 4;
    const size_t off_bar = off; off += 16;
    P.R = R;
    if (off + 128 > budget) continue;
    P.nbands = (H + R - 1) / R; P.tile_rows = rows; P.tile_bytes = (uint32_t)tile;
    P.off_tab = (uint32_t)off_tab; P.off_part = (uint32_t)off_part; P.off_ws = (uint32_t)off_ws; P.off_thr = (uint32_t)off_thr; P.off_bar = (uint32_t)off_bar;
    P.smem = off + 128;
    // enough CTAs to fill the machine a few times over, as long as bands stay at least 4 rows (2 stored rows) high
    const long long ctas = (long long)P.ns * P.nbands * batch, want = 3ll * cdn_num_sms() * per_sm;
    if (ctas >= want || R <= 4 * step) { P.ok = 1; return P; }
    P.ok = 1;                                                        // fits; keep looking for a finer split
  }
  return P;
}

bool deform_tile_ok(const DwDevice& d, const cdn_deform_scale* sc, int in_pitch, int out_pitch, int batch, int H, int W, int in_shift) {
  if (g_cdn_debug_flags & 2048u) return false;                                              // bit 11: the LDG kernels of dw.cu (A/B)
  if (sc->mode == 0 && (!d.use_int || !d.ki || !d.s_mode0_ok)) return false;
  if (sc->mode == 1 && (!d.mb || (long long)H * W > (1 << 23))) return false;
  if (H >= 65536 || W >= 65536 || batch > 65535) return false;
  return deform_tile_plan(d.cw_total, in_pitch, H, W, in_shift, sc->bound, batch, sc->mode).ok != 0;
}

int deform_tile_launch(const DwDevice& d, const cdn_deform_scale* sc, const int8_t* in, int in_pitch, int8_t* out, int out_pitch,
                       int batch, int H, int W, int in_shift, int zx, float* sval, const DwParams* dwp, cudaStream_t st) {
  const DefTPlan P = deform_tile_plan(d.cw_total, in_pitch, H, W, in_shift, sc->bound, batch, sc->mode);
  CDN_CHECK(P.ok, CDN_ERR_STATE, "deform (tile): shape not eligible");
  DefTParams p; memset(&p, 0, sizeof(p));
  p.Hs = H >> in_shift; p.Ws = W >> in_shift; p.H = H; p.W = W; p.shift = in_shift;
  p.pitch_out_w = out_pitch / 4; p.cw_total = d.cw_total;
  p.SLB = P.SLB; p.ns = P.ns; p.R = P.R; p.reach = sc->mode == 1 ? sc->bound + 1 : sc->bound; p.tile_rows = P.tile_rows;
  p.nst_max = (P.R >> in_shift) * p.Ws; p.lpp = P.lpp; p.ppw = P.ppw;
  p.off_tab = P.off_tab; p.off_part = P.off_part; p.off_ws = P.off_ws; p.off_thr = P.off_thr; p.off_bar = P.off_bar; p.tile_bytes = P.tile_bytes;
  p.wA = d.wA; p.wB = d.wB; p.wC = d.wC; p.ki = (const int4*)d.ki; p.lo_i = d.rq.lo;
  p.ws = d.ws; p.s_thr = d.s_thr; p.s_n = d.s_n; p.s_lo = d.s_lo;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx) * 0x01010101u;
  p.out = (uint32_t*)out; p.sval = sval;
  p.no_dedup = (g_cdn_debug_flags & 16384u) ? 1 : 0;                // bit 14: no 2x2 block units (A/B)
  p.dbg = (int)((g_cdn_debug_flags >> 16) & 7u);                    // bits 16-18: timing experiments (wrong results)
  CUtensorMap tmI;
  if (int r = make_tmap_nhwc_box(&tmI, in, (uint64_t)in_pitch, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)batch, (uint32_t)P.SLB,
                                 (uint32_t)(p.Ws + (sc->mode == 1 ? 1 : 0)), (uint32_t)P.tile_rows)) return r;
  const bool lo_on = d.rq.lo > -128;
  typedef void (*Kern)(CUtensorMap, DefTParams);
  static const Kern kerns[3][3] = {
      {deform_tile_int_kernel<2, 1, 256, 2, true>, deform_tile_int_kernel<2, 2, 256, 2, true>, deform_tile_int_kernel<2, 3, 256, 2, true>},
      {deform_tile_int_kernel<1, 1, 512, 2, false>, deform_tile_int_kernel<1, 2, 512, 2, false>, deform_tile_int_kernel<1, 3, 512, 2, false>},
      {deform_tile_int_kernel<1, 1, 256, 3, true>, deform_tile_int_kernel<1, 2, 256, 3, true>, deform_tile_int_kernel<1, 3, 256, 3, true>}};
  const int rqv = lo_on ? 1 : ((d.sh0 && !(g_cdn_debug_flags & (1u << 22))) ? 2 : 0);       // requantisation variant: plain, lower clamp, shift 0
  static bool attr_set[10][64] = {};
  if (sc->mode == 1) {
    CDN_CHECK(dwp != nullptr, CDN_ERR_STATE, "deform (tile, bilinear): missing parameter block");
    if (cdn_first_on_device(attr_set[9])) {
      CDN_CUDA(cudaFuncSetAttribute(deform_tile_bil_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      CDN_CUDA(cudaFuncSetAttribute(deform_tile_bil_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
  } else if (cdn_first_on_device(attr_set[P.variant * 3 + rqv])) {
    CDN_CUDA(cudaFuncSetAttribute(kerns[P.variant][rqv], cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    CDN_CUDA(cudaFuncSetAttribute(kerns[P.variant][rqv], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)P.ns, (unsigned)P.nbands, (unsigned)batch); cfg.blockDim = dim3((unsigned)P.NT); cfg.stream = st;
  cfg.dynamicSmemBytes = P.smem;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_cdn_debug_flags & 64u) ? 0 : 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (P.ns > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)P.ns; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  if (sc->mode == 1) {
    DefTBil f; memset(&f, 0, sizeof(f));
    f.mb = d.mb; f.lo_f = (float)d.rq.lo; f.thr_bil = d.thr_bil;
    f.Ms = sc->Ms; f.bs = sc->bs; f.ss = sc->ss; f.zs = sc->zs; f.u_lo = (double)(-sc->bound + 1); f.u_hi = (double)sc->bound;
    f.acc_s_bias = d.acc_s_bias;
    CDN_CUDA(cudaLaunchKernelEx(&cfg, deform_tile_bil_kernel<256>, tmI, p, f, *dwp));
  } else CDN_CUDA(cudaLaunchKernelEx(&cfg, kerns[P.variant][rqv], tmI, p));
  CDN_LAUNCH_CHECK("deform_tile kernel");
  return 0;
}
