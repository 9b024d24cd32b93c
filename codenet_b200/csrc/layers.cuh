// Per-layer device state and launchers shared between the translation units of libcodenet_b200.
#pragma once
#include "common.cuh"

struct DwDevice {                            // depthwise / deformable layer constants
  uint32_t *wA = nullptr, *wB = nullptr, *wC = nullptr, *ws = nullptr;
  uint32_t *wpk1 = nullptr, *wpk2 = nullptr, *wpku = nullptr;   // v2 packings: stride 1, stride 2, upsample-folded
  float2* mb = nullptr; int32_t* abm = nullptr; float thr = 0.5f, thr_bil = 0.01f; int u_ok = 1;
  RqInt* ki = nullptr; int use_int = 0;        // exact integer requantisation constants (acc_bias folded in)
  int sh0 = 0;                                 // use_int and every channel's shift is 0 (kernels may drop the SHF after the IMAD.HI)
  uint32_t* pad_px = nullptr;                  // integer-offset mode: one pixel of pad words (q = -zx), the target of out-of-image taps
  int32_t* s_thr = nullptr; int s_n = 0, s_lo = 0, s_mode0_ok = 0;   // integer-offset mode: thresholds of s over the scale conv's dot product
  DevRequant rq;
  int cw_total = 0;
  long long acc_s_bias = 0;
};
int dw_device_build(DwDevice& d, const int8_t* wq, const int8_t* ws, int C, int Cp, int zx, const cdn_requant* rq);
int deform_scale_build(DwDevice& d, const cdn_deform_scale* sc, const int8_t* ws, int C, int zx);   // after dw_device_build, deformable layers
void dw_device_free(DwDevice& d);
int dw_launch(const DwDevice& d, const int8_t* in, int in_pitch, int8_t* out, int out_pitch, int batch, int H, int W,
              int in_shift, int stride, int zx, cudaStream_t st);
// depthwise conv (stride 1 / 2) with the input tile staged by TMA (dw_tma.cu); dw_launch picks it when the shape is eligible
bool dw_tma_ok(const DwDevice& d, int in_pitch, int out_pitch, int H, int W, int in_shift, int stride);
int dw_tma_launch(const DwDevice& d, const int8_t* in, int in_pitch, int8_t* out, int out_pitch, int batch, int H, int W,
                  int stride, int zx, cudaStream_t st);
int deform_launch(const DwDevice& d, const cdn_deform_scale* sc, const int8_t* in, int in_pitch, int8_t* out,
                  int out_pitch, int batch, int H, int W, int in_shift, int zx, float* sval, cudaStream_t st);

// integer-offset layer with the input tile + halo staged in shared memory by TMA (deform_tile.cu); deform_launch picks it
// when the shape is eligible
bool deform_tile_ok(const DwDevice& d, const cdn_deform_scale* sc, int in_pitch, int out_pitch, int batch, int H, int W, int in_shift);
struct DwParams;
int deform_tile_launch(const DwDevice& d, const cdn_deform_scale* sc, const int8_t* in, int in_pitch, int8_t* out, int out_pitch,
                       int batch, int H, int W, int in_shift, int zx, float* sval, const DwParams* dwp, cudaStream_t st);
int make_tmap_nhwc_box(CUtensorMap* m, const void* base, uint64_t pitch, uint64_t W, uint64_t H, uint64_t batch,
                       uint32_t box_c, uint32_t box_w, uint32_t box_h);

struct StemDevice {
  double* w = nullptr; double* M = nullptr; double* B = nullptr; int C = 0; double lo = -128;
  std::vector<double> hw, hM, hB;            // host copies: passed to the kernel by value (constant bank)
};
int stem_device_build(StemDevice& d, const int8_t* wq, int C, const cdn_requant* rq);
void stem_device_free(StemDevice& d);
int stem_launch(const StemDevice& d, const float* img, int batch, int H, int W, int stride, int pool,
                int8_t* out, int out_pitch, int8_t* tmp, cudaStream_t st, const uint8_t* img_u8 = nullptr,
                const float* lut = nullptr);

struct PwDevice {
  int K = 0, k_off = 0, N = 0, Kp = 0, BN = 0, n_tiles = 0, num_k_blocks = 0, stages = 0;
  int has_pass = 0, pass_segs = 0, pass_bufs = 1, groups = 2, nbuf = 0, resident = 0, n_chunks = 0, n_segs = 0, n_f32 = 0;
  float thr = 0.5f;                          // layer-wide rounding-boundary guard (min over columns)
  int use_int = 0;                           // kc holds RqInt records (exact integer requantisation) instead of the fp32 pairs
  int sh0 = 0;                               // use_int and every column's shift is 0
  int il_hp = 0;                             // ... and il_hp bytes per group in the output pixel
  int il_pg = 0;                             // > 0: the chunk table is the canonical cat + channel_shuffle interleave with il_pg channels per group
  size_t smem_bytes = 0;
  int8_t* w = nullptr;                       // [BN*n_tiles][Kp]
  cdn_pw_chunk* chunks = nullptr; void* segs = nullptr; int* tile_seg = nullptr; void* kc = nullptr;
  double* Mf = nullptr; double* bf = nullptr;
  DevRequant rq;
  CUtensorMap tmB;
};
int pw_device_build(PwDevice& d, const cdn_pw_desc* desc, int pass_pitch);
void pw_device_free(PwDevice& d);
int pw_init_attrs();
int pw_launch(const PwDevice& d, const int8_t* in, int in_pitch, long long pixels, const int8_t* pass, int pass_pitch,
              int8_t* out, int out_pitch, float* out_f32, int ppi, const CUtensorMap* tmA, const CUtensorMap* tmP,
              const CUtensorMap* tmO, cudaStream_t st);
// heads tail: depthwise conv through the x2 upsample as the A-tile producer of the fp32 head conv (heads_fused.cu)
bool heads_fused_ok(const DwDevice& dw, const PwDevice& pw, int in_pitch, int mid_pitch, int Hs, int Ws);
int heads_fused_launch(const DwDevice& dw, const PwDevice& pw, const int8_t* in, int in_pitch, int batch, int Hs, int Ws,
                       int zx, float* out_f32, cudaStream_t st);
// a whole stride-1 ShuffleNetV2 unit (1x1 conv, depthwise 3x3, 1x1 conv + cat + channel shuffle) as one kernel (unit_fused.cu)
bool unit_fused_ok(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, int x_pitch, int mid_pitch, int out_pitch, int H, int W);
int unit_fused_launch(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, const int8_t* x, int8_t* out, int HP,
                      int batch, int H, int W, int zx_mid, int8_t* dump_c1, int8_t* dump_d2, cudaStream_t st);
// the same unit, warp-specialised (unit_fused_ws.cu): the phases of consecutive tiles on different warps, 116 / 122-channel halves
bool unit_fused_ws_ok(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, int x_pitch, int mid_pitch, int out_pitch, int H, int W);
int unit_fused_ws_launch(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, const int8_t* x, int8_t* out,
                         int batch, int H, int W, int zx_mid, int8_t* dump_c1, int8_t* dump_d2, cudaStream_t st);
int make_tmap_nhwc_swz(CUtensorMap* m, const void* base, uint64_t pitch, uint64_t W, uint64_t H, uint64_t batch,
                       uint32_t box_c, uint32_t box_w, uint32_t box_h);
// the branch of the FIRST stride-2 unit (1x1 conv at full resolution, depthwise 3x3 stride 2, 1x1 conv + cat + channel shuffle)
// as one kernel (unit_s2_fused.cu)
bool unit_s2_fused_ok(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, int x_pitch, int mid_pitch, int pass_pitch,
                      int out_pitch, int H, int W);
int unit_s2_fused_launch(const PwDevice& pw1, const DwDevice& dw, const PwDevice& pw3, const int8_t* x, const int8_t* pass, int pass_pitch,
                         int8_t* out, int batch, int H, int W, int zx_mid, int8_t* dump_c1, int8_t* dump_d2, cudaStream_t st);
int make_tmap_nhwc(CUtensorMap* m, const void* base, uint64_t pitch, uint64_t W, uint64_t H, uint64_t batch, uint32_t box_w, uint32_t box_h);
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_rows);

int decode_launch(const float* hm, long long hm_img_stride, const float* wh, long long wh_img_stride, const float* reg,
                  long long reg_img_stride, int batch, int cat, int H, int W, int K, int is_prob,
                  unsigned long long* scratch, unsigned int* counts, float* dets, int32_t* inds, cudaStream_t st);
