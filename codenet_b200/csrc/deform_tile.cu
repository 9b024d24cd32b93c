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
  if (!lane_on) return;
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
      if (V == 2) { const uint2 r = lds_u64(a); x[0][t] = r.x; x[V - 1][t] = r.y; }
      else x[0][t] = lds_u32(a);
    }
  };
  auto compute = [&](uint32_t meta, const uint32_t (&x)[V][9]) {
    uint32_t o[V];
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
      for (int c = 0; c < 4; ++c) { q[c] = rq_int_hi(acc[c], Mi[v][c], sh[v][c], Bi[v][c]); if (RQ == 2) q[c] = max(q[c], p.lo_i); }
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
// host side
// ---------------------------------------------------------------------------------------------------------
// kernel variants (A/B through cdn_set_debug_flags): 0 = 8 channels per lane, 256 threads, 2 CTAs per SM, software pipeline;
// 1 = 4 channels per lane, 512 threads, 64 registers, no software pipeline (twice the warps); 2 = 4 channels per lane, 256
// threads, 3 CTAs per SM
struct DefTPlan { int ok, variant, V, NT, SLB, ns, R, nbands, tile_rows, lpp, ppw; uint32_t off_tab, off_part, off_ws, off_thr, off_bar, tile_bytes; size_t smem; };

// Chooses slice, band height and the shared-memory layout; ok = 0 when the shape does not fit (the caller keeps deform_int_v3_kernel).
static DefTPlan deform_tile_plan(int cw_total, int in_pitch, int H, int W, int in_shift, int reach, int batch) {
  DefTPlan P; memset(&P, 0, sizeof(P));
  const int Hs = H >> in_shift, Ws = W >> in_shift;
  const int cp = cw_total * 4;                                       // channel bytes the layer processes
  P.SLB = cp >= 256 ? 256 : cp;
  if (P.SLB % 16 || in_pitch % 16 || Ws > 256 || Ws < 1) return P;
  P.ns = (cp + P.SLB - 1) / P.SLB;
  if (P.ns > 8) return P;                                            // portable cluster size
  P.variant = DTL_DEFAULT_VARIANT;
  if (g_cdn_debug_flags & 4096u) P.variant = 1;                     // bits 12 / 13: kernel variant (A/B)
  if (g_cdn_debug_flags & 8192u) P.variant = 2;
  if ((g_cdn_debug_flags & 12288u) == 12288u) P.variant = 0;
  if (P.SLB < 64 && P.variant == 0) P.variant = 2;
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
    const size_t tile = (size_t)rows * Ws * P.SLB;
    size_t off = tile + P.SLB;                                       // + pad pixel
    off = (off + 15) / 16 * 16; const size_t off_tab = off; off += (size_t)R * W * DTL_TAB_STRIDE;
    const int nst_max = (R >> in_shift) * Ws;
    const size_t off_part = off; off += (size_t)P.ns * nst_max * 4;
    off = (off + 15) / 16 * 16; const size_t off_ws = off; off += P.SLB;
    const size_t off_thr = off; off += 128 * 4;
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
  if (sc->mode != 0 || !d.use_int || !d.ki || !d.s_mode0_ok || (g_cdn_debug_flags & 2048u)) return false;   // bit 11: v3 kernel (A/B)
  if (H >= 65536 || W >= 65536 || batch > 65535) return false;
  return deform_tile_plan(d.cw_total, in_pitch, H, W, in_shift, sc->bound, batch).ok != 0;
}

int deform_tile_launch(const DwDevice& d, const cdn_deform_scale* sc, const int8_t* in, int in_pitch, int8_t* out, int out_pitch,
                       int batch, int H, int W, int in_shift, int zx, float* sval, cudaStream_t st) {
  const DefTPlan P = deform_tile_plan(d.cw_total, in_pitch, H, W, in_shift, sc->bound, batch);
  CDN_CHECK(P.ok, CDN_ERR_STATE, "deform (tile): shape not eligible");
  DefTParams p; memset(&p, 0, sizeof(p));
  p.Hs = H >> in_shift; p.Ws = W >> in_shift; p.H = H; p.W = W; p.shift = in_shift;
  p.pitch_out_w = out_pitch / 4; p.cw_total = d.cw_total;
  p.SLB = P.SLB; p.ns = P.ns; p.R = P.R; p.reach = sc->bound; p.tile_rows = P.tile_rows;
  p.nst_max = (P.R >> in_shift) * p.Ws; p.lpp = P.lpp; p.ppw = P.ppw;
  p.off_tab = P.off_tab; p.off_part = P.off_part; p.off_ws = P.off_ws; p.off_thr = P.off_thr; p.off_bar = P.off_bar; p.tile_bytes = P.tile_bytes;
  p.wA = d.wA; p.wB = d.wB; p.wC = d.wC; p.ki = (const int4*)d.ki; p.lo_i = d.rq.lo;
  p.ws = d.ws; p.s_thr = d.s_thr; p.s_n = d.s_n; p.s_lo = d.s_lo;
  p.pad_word = (uint32_t)(uint8_t)(int8_t)(-zx) * 0x01010101u;
  p.out = (uint32_t*)out; p.sval = sval;
  p.no_dedup = (g_cdn_debug_flags & 16384u) ? 1 : 0;                // bit 14: no 2x2 block units (A/B)
  CUtensorMap tmI;
  if (int r = make_tmap_nhwc_box(&tmI, in, (uint64_t)in_pitch, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)batch, (uint32_t)P.SLB,
                                 (uint32_t)p.Ws, (uint32_t)P.tile_rows)) return r;
  const bool lo_on = d.rq.lo > -128;
  typedef void (*Kern)(CUtensorMap, DefTParams);
  static const Kern kerns[3][2] = {
      {deform_tile_int_kernel<2, 1, 256, 2, true>, deform_tile_int_kernel<2, 2, 256, 2, true>},
      {deform_tile_int_kernel<1, 1, 512, 2, false>, deform_tile_int_kernel<1, 2, 512, 2, false>},
      {deform_tile_int_kernel<1, 1, 256, 3, true>, deform_tile_int_kernel<1, 2, 256, 3, true>}};
  const Kern kern = kerns[P.variant][lo_on ? 1 : 0];
  static bool attr_set[6][64] = {};
  if (cdn_first_on_device(attr_set[P.variant * 2 + (lo_on ? 1 : 0)])) {
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    CDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
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
  CDN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmI, p));
  CDN_LAUNCH_CHECK("deform_tile_int_kernel");
  return 0;
}
