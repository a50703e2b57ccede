// Internal interfaces of the tile-classifier forward (K2): fp32 CUDA-core path and
// bf16 tcgen05 path share the model object defined in model.cu.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched at run time)

#include "common.cuh"

namespace cs {

// ---------------------------------------------------------------------------
// fp32 path (fwd_fp32.cu)
// ---------------------------------------------------------------------------
struct ConvF32Args {
  const float* in;  // element (n, ci, y, x) at n*in_sn + ci*in_sc + y*in_sy + x*in_sx
  int64_t in_sn, in_sc, in_sy, in_sx;
  int Hi, Wi, Cin, Ho, Wo, Cout, k, stride, pad;
  const float* w;         // [k*k*Cin][Cout], row index (dy*k+dx)*Cin + ci
  const float* bias;      // [Cout] (folded BN shift)
  const float* residual;  // nullable, NHWC like out
  float* out;             // [M][Cout], M = n*Ho*Wo  (NHWC)
  int64_t M;
  int relu;
};
int launch_conv_fp32(const ConvF32Args& a, cudaStream_t st);
int launch_maxpool_fp32(const float* in, float* out, int64_t n, int Hi, int Wi, int C,
                        cudaStream_t st);
int launch_bilinear_fp32(const float* in, float* out, int64_t n, int Hi, int Wi, int Ho, int Wo, int C,
                         cudaStream_t st);
int launch_head_fp32(const float* x4, int64_t n, int P, int C, const float* fc_w,
                     const float* fc_b, float* prob_out, float* logits_out, float* feat_out,
                     cudaStream_t st);

// ---------------------------------------------------------------------------
// bf16 tcgen05 path (stem_ts.cu, stem_win.cu, conv_ysum.cu, conv_halo.cu, conv_gemm.cu, fwd_tc.cu)
// ---------------------------------------------------------------------------
constexpr int kGemmBM = 128;       // UMMA M (rows per CTA tile)
constexpr int kGemmBK = 64;        // bf16 elements per K step = one 128-byte swizzle row
constexpr int kMaxSteps = 40;      // K steps per (layer, n-variant)
constexpr int kMaxVariants = 8;

// One K step of 64 channels, packed: which A box to fetch and which B columns it meets.
//   bits  0-9  : A coordinate 0 / 64 (channel chunk, or K-column chunk in 2-D mode)
//   bits 10-19 : B coordinate 0 / 64 (K-column chunk)
//   bits 20-21 : dx + 1, bits 22-23 : dy + 1   (4-D maps: pixel shift in {-1, 0, +1})
//   bits 24-25 : A tensor map index
struct KStep {
  uint32_t v;
  __host__ __device__ int a_c0() const { return (int)(v & 1023u) * 64; }
  __host__ __device__ int b_k() const { return (int)((v >> 10) & 1023u) * 64; }
  __host__ __device__ int dx() const { return (int)((v >> 20) & 3u) - 1; }
  __host__ __device__ int dy() const { return (int)((v >> 22) & 3u) - 1; }
  __host__ __device__ int map() const { return (int)((v >> 24) & 3u); }
  static KStep make(int a_c0, int b_k, int dx, int dy, int map) {
    KStep s;
    s.v = (uint32_t)(a_c0 / 64) | ((uint32_t)(b_k / 64) << 10) | ((uint32_t)(dx + 1) << 20) |
          ((uint32_t)(dy + 1) << 22) | ((uint32_t)map << 24);
    return s;
  }
};

struct GemmParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  KStep steps[kMaxVariants][kMaxSteps];
  int n_steps[kMaxVariants];
  int n_variants;     // 1: every n-tile walks steps[0]; else steps[n_tile]
  // more than kMaxVariants step lists (wide grouped convs): device table
  // [n_variants][kMaxSteps + 1] of u32, entry 0 = step count; nullptr otherwise
  const uint32_t* ext_steps;
  int a_mode;         // bit i set: A map i is 4-D {c, x, y, tile}; clear: 2-D {k, row}
  int units_per_mtile;  // 4-D mode: tiles (instances) per 128-row M tile
  int num_m_tiles, num_n_tiles;
  int reverse;        // walk work items from the last to the first (L2 snake order between layers)
  int stages;         // set by the launcher: TMA/MMA ring depth that fits shared memory
  uint32_t epi_set_bytes;  // set by the launcher: epilogue staging set, 16 KB (hi) or 32 KB (hi+lo)
  uint32_t epi_sets;       // set by the launcher: 2 staging sets, or a ring of 4 (residual GEMMs)
  int cluster;        // 1, or 2: CTA pairs share the B tile by TMA multicast (b_map box = BN/2 rows)
  int n_total;        // output row pitch, elements
  int64_t m_valid;    // rows that exist
  const float* bias;  // [n_total]
  const __nv_bfloat16* res_hi;  // nullable
  const __nv_bfloat16* res_lo;  // nullable
  __nv_bfloat16* out_hi;        // nullable
  __nv_bfloat16* out_lo;        // nullable
  float* out_f32;         // nullable: raw fp32 result (diagnostics)
  int relu;
  // TMA views of the four tensors above: {n_total, rows} with a {64, 128} box (make_mat_map_2d)
  CUtensorMap res_hi_map, res_lo_map, out_hi_map, out_lo_map;
};

// BN (the CTA's N tile) must be 64, 128 or 256.
int launch_conv_gemm(const GemmParams& p, int BN, cudaStream_t st);

// Host: 4-D activation map {C, W, H, T} with a {64, bw, bh, bt} box, 128-byte swizzle.
int make_act_map_4d(CUtensorMap* map, const void* base, int C, int W, int H, int64_t T,
                    int64_t stride_x_elems, int64_t stride_y_elems, int64_t stride_t_elems,
                    int box_w, int box_h, int box_t);
// Host: 2-D K-major matrix map {K, rows} with a {64, box_rows} box, 128-byte swizzle.
int make_mat_map_2d(CUtensorMap* map, const void* base, int64_t K, int64_t rows,
                    int64_t row_pitch_elems, int box_rows);

// Stride-1 3x3 conv with y-halo tiles (conv_halo.cu): 8x8x64 and 4x4x128 stages.
struct HaloParams {
  CUtensorMap a_map;  // make_act_map_halo
  CUtensorMap b_map;  // [Cout][9*Cin] K-major, box {64, Cout / cluster}
  int cluster;        // 1, or 2: CTA pairs (tcgen05 cta_group::2), each CTA holds half of B
  int num_m_tiles;    // ceil(instances / (128 / (W*W)))
  int reverse;        // walk M tiles from the last to the first
  int64_t n_inst;     // instances that exist
  const float* bias;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;
  int relu;
  // TMA views {C, W, T, H}, box {64, W, IMG, H} (make_act_map_halo with halo = 0)
  CUtensorMap res_hi_map, res_lo_map, out_hi_map, out_lo_map;
  // fused 1x1 stride-2 downsample of the block input (layer-2 entry): one extra K step whose A
  // box comes from ds_map (make_act_map_halo_ds) and whose weights are B columns [9*Cin, +64)
  CUtensorMap ds_map;
  int has_ds;
};
bool halo_supported(int W, int Cin, int Cout);
int make_act_map_halo_ds(CUtensorMap* map, const void* base, int C, int W, int H, int64_t T);
int launch_conv_halo(const HaloParams& p, int W, int Cin, cudaStream_t st);
int make_act_map_halo(CUtensorMap* map, const void* base, int C, int W, int H, int64_t T,
                      int halo /* 1: box H+2 rows for the A operand, 0: H rows for epilogue tiles */);

// Layer-1 convs in y-sum form (conv_ysum.cu): 8x8 images, 64 -> 64 channels, stride 1.
struct YsumParams {
  CUtensorMap a_map;  // make_act_map_4d {C, W, H, T}, box {64, 8, 8, 2}
  CUtensorMap a_box_map;  // same tensor, box {64, 10, 8, 2} (one-box form: fetched at x = -1)
  CUtensorMap b_map;  // [3*192][64] from pack_ysum_weights, box {64, 192 / cluster}
  int cluster;        // 1, or 2: CTA pairs (tcgen05 cta_group::2)
  int num_m_tiles;    // ceil(instances / 2) of this launch
  int tile_base;      // first M tile of this launch (sub-batches of a forward batch); multiple of `cluster`
  int reverse;
  int64_t n_inst;
  const float* bias;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;
  int relu;
  CUtensorMap res_hi_map, res_lo_map, out_hi_map, out_lo_map;   // [rows][64], box {64, 128}
};
// A whole layer-1 BasicBlock in one kernel (conv_block.cu): both convs in y-sum form, the
// intermediate tensor in shared memory, the residual taken from conv1's operand box.  CTA pairs.
struct YsumBlockParams {
  CUtensorMap x_box_map;        // block input {C, W, H, T}, box {64, 10, 8, 2} (fetched at x = -1)
  CUtensorMap b1_map, b2_map;   // [3*192][64] from pack_ysum_weights, box {64, 96}
  CUtensorMap out_map;          // [rows][64], box {64, 128}
  int num_m_tiles;              // ceil(instances / 2) of this launch
  int tile_base;                // first M tile of this launch; even
  int reverse;
  const float* bias1;
  const float* bias2;
};
int launch_ysum_block(const YsumBlockParams& p, cudaStream_t st);
bool ysum_supported(int W, int Cin, int Cout);
void pack_ysum_weights(const float* w_oihw, uint16_t* out /* [576 * 64] */);
int launch_conv_ysum(const YsumParams& p, cudaStream_t st);

struct StemArgs {
  // source A: u8 images + uniform grid
  const uint8_t* img;
  int H, W, tile, interval, grid_w;
  int64_t tiles_per_bag;
  int64_t inst_begin;
  // source B: materialised fp32 NCHW tiles (nullable)
  const float* x;
  int64_t count;          // instances in this batch
  const float* w;         // [147][64] fp32 (folded)
  const float* bias;      // [64]
  __nv_bfloat16* out_hi;  // [count][Hp*Wp][64]
  __nv_bfloat16* out_lo;
};
int launch_stem_bf16(const StemArgs& a, cudaStream_t st);

// Window-form tensor-core stem (stem_win.cu), tile 32 only: no im2col, the padded bf16 image is
// the UMMA operand.  w_packed_dev holds stem_win_weight_bytes() bytes from pack_stem_weights_win.
int stem_win_weight_bytes();
void pack_stem_weights_win(const float* w_oihw, uint16_t* out);
int launch_stem_win(const StemArgs& a, const void* w_packed_dev, const uint16_t* lut_bf16_dev,
                    cudaStream_t st);
// Weights-stationary stem (stem_ts.cu), tile 32 only, the default: the filters live in tensor
// memory as the A operand, the staged image is the B operand (half the shared-memory operand reads).
int stem_ts_weight_bytes();
void pack_stem_weights_ts(const float* w_oihw, uint16_t* out);
int launch_stem_ts(const StemArgs& a, const void* w_packed_dev, cudaStream_t st);

int launch_head_bf16(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int64_t n, int P,
                     int C, const float* fc_w, const float* fc_b, float* prob_out,
                     float* logits_out, float* feat_out, cudaStream_t st);

int get_norm_lut_host(float* dst768);

}  // namespace cs
