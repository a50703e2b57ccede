// Model object for the tile classifier: weight packing, per-layer GEMM planning and the
// forward drivers behind cs_model_forward_tiles / cs_model_forward_tensor.
//
// Reference networks: MILResNet with BasicBlock or Bottleneck (model/resnet.py:15-79, 81-127,
// 179-193, 234-269; ctors :336-361) and MILResNeXt with grouped Bottleneck
// (model/resnext.py:67-113, 305-340, 418-442).  The caller has folded eval-mode BN into every
// conv (fp32).  Activation layout on the device is [instance][pixel (row-major y,x)][channel];
// bf16 mode keeps a `hi` tensor (MMA operand) and, for block outputs, a `lo` tensor
// (value - hi) for the residual stream.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <vector>

#include "fwd.cuh"

namespace cs {
namespace {

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

bool env_is(const char* name, const char* value) {
  const char* e = getenv(name);
  return e != nullptr && strcmp(e, value) == 0;
}
const bool g_snake = !env_is("CELLSEG_SNAKE", "0");          // alternate the tile walk per layer
// CTA pairs (tcgen05 cta_group::2, M = 256): on by default for the N >= 128 kernels
// (layer-2 halo convs -10 %, dense / pointwise GEMMs -7 %); the N = 64 halo conv of layer 1 is
// 14 % slower in pair mode and stays single-CTA.  CELLSEG_CLUSTER=1 turns pairs off.
const int g_cluster = env_is("CELLSEG_CLUSTER", "1") ? 1 : 2;
const bool g_disable_halo = env_is("CELLSEG_HALO", "0");     // diagnostics: generic kernel only
const bool g_disable_halo_ds = env_is("CELLSEG_HALO_DS", "0");   // layer-2 entry conv2 in the generic kernel
// 3x3 convs with at most this many output pixels take the dense form (rows = instances, N =
// (output pixel, co), K = (input pixel, ci), all-zero weight blocks skipped): 16 = the 2x2 and
// the 4x4 stages.  For the 4x4 x 128 stage (layer 2) the dense form issues 17 % fewer MMA cycles
// than the halo kernel (only in-bounds taps: 1.97 M instead of 2.36 M MACs per instance and conv
// at N = 256) and streams whole 128-instance M tiles: +5.4 % on the whole step, conv stage
// 0.574 -> 0.606 of the sustained peak (gpurun r2d).  CELLSEG_DENSE_PO=4 restores the halo
// kernel for layer 2; =64 also sends the 8x8 stage through the dense form (more MMA work: 3.6 M
// against 2.36 M MACs per conv -- measured, not the default).
const int g_dense_max_po = env_is("CELLSEG_DENSE_PO", "4") ? 4 : (env_is("CELLSEG_DENSE_PO", "64") ? 64 : 16);
// N tile of the dense form: 256 (two 4x4-stage pixels per tile) by default; CELLSEG_DENSE_BN=128
// gives every output pixel of the 4x4 stage its own tile (exactly the in-bounds taps, smaller MMAs).
const int g_dense_bn = env_is("CELLSEG_DENSE_BN", "128") ? 128 : 256;
// Grouped 3x3 convs on 4x4 maps (ResNeXt layer 2) also take the dense form, with one (pixel,
// 64-channel block) per 64-wide N tile: only the in-bounds taps of the tile's own channel block
// are visited (1.64 M instead of 2.36 M MACs per instance and conv) on 128-instance M tiles;
// ResNeXt-50 2.95 -> 3.06 M instances/s (gpurun r2n).  CELLSEG_DENSE_GROUP_PO=4 restores the
// shifted-box form.
const int g_dense_group_po = env_is("CELLSEG_DENSE_GROUP_PO", "4") ? 4 : 16;
// Instances per stem + layer-1 sub-batch; 0 (default): whole forward batches.  4 736 = 16 y-sum M
// tiles per CTA: x, mid and y of a sub-batch are 3 x 39 MB and stay in L2.  Measured (gpurun r2j,
// one box): 0 -> 10.2 M instances/s at 1 390 MHz; 9 472 -> 10.08 M; 4 736 -> 9.93 M at 1 515 MHz;
// 2 368 -> 9.50 M at 1 612 MHz.  The saved HBM traffic shows as a HIGHER clock under the power cap,
// but the 7 extra launches per sub-batch cost more than it returns.  Experiment switch.
const int64_t g_l1_sub = []() -> int64_t {
  const char* e = getenv("CELLSEG_L1_SUB");
  return e != nullptr ? atoll(e) : 0;
}();
// Stem form for tile 32.  Default: weights-stationary (stem_ts.cu: filters in tensor memory, the
// staged image as the shared-memory operand).  CELLSEG_STEM=win restores the window form with
// both operands in shared memory (stem_win.cu).
const bool g_stem_win = env_is("CELLSEG_STEM", "win");
// Residual stream: bf16 by default.  CELLSEG_RESIDUAL=hilo carries a second bf16 tensor
// lo = value - bf16(value) between blocks (~16 mantissa bits); measured max|dp| moves by < 1.5e-3
// (ResNet-34 0.0101 -> 0.0098) while the block epilogues move twice the bytes (-11 % throughput).
const bool g_no_lo = !env_is("CELLSEG_RESIDUAL", "hilo");
const int g_xbufs = g_no_lo ? 2 : 4;
// Layer 1 in y-sum form (conv_ysum.cu: three vertical taps per N = 192 MMA, summed in the
// epilogue).  Halves the operand reads of the MMAs at the price of a three times larger
// accumulator read-out: 102 vs 108 us per 18 944 instances against the halo kernel and +2 %
// sustained throughput (less power).  CELLSEG_YSUM=0 puts layer 1 back on the halo kernel.
const bool g_disable_ysum = env_is("CELLSEG_YSUM", "0");
// y-sum in CTA pairs (M = 256 MMAs, each CTA holds half of every weight tile): with the 16-warp
// epilogue the kernel is no longer epilogue-bound and the halved B-operand reads pay, +1.5..2 %
// on the whole step.  CELLSEG_YSUM_PAIRS=0 keeps single-CTA MMAs.
const bool g_ysum_pairs = !env_is("CELLSEG_YSUM_PAIRS", "0");
// Layer-1 BasicBlocks as ONE kernel (conv_block.cu): conv1 -> ReLU -> conv2 -> + x -> ReLU with the
// intermediate tensor in shared memory (16 instead of 40 KB of HBM traffic per instance and block).
// Needs CTA pairs and the bf16 residual stream.  CELLSEG_BLOCK_FUSE=0 restores two y-sum launches.
const bool g_block_fuse = !env_is("CELLSEG_BLOCK_FUSE", "0");
const bool g_no_group_ysum = env_is("CELLSEG_GROUP_YSUM", "0");   // grouped 8x8 convs back in the generic kernel

struct ConvW {
  int cin = 0, cout = 0, k = 0, stride = 1, pad = 0, groups = 1;
  std::vector<float> w;  // dense OIHW [cout][cin][k][k], BN folded, zeros outside the groups
  std::vector<float> b;  // [cout]
  float* d_w32 = nullptr;  // [k*k*cin][cout]
  float* d_b = nullptr;
};

struct BlockDesc {
  bool bottleneck;
  int conv1, conv2, conv3, ds;  // indices into convs; conv3 / ds = -1 when absent
  int cin, width, cout, stride;
};

struct ConvGeom {
  int Hi, Wi, Cin, Ho, Wo, Cout, k, stride, pad, groups;
};

struct PlannedConv {
  GemmParams p;
  int BN = 0;
  bool dense = false;  // rows = instances (else rows = instances x output pixels)
  int Po = 0;          // output pixels per instance
  bool halo = false;   // stride-1 3x3 on 8x8x64 / 4x4x128: y-halo kernel (conv_halo.cu)
  HaloParams hp;
  int halo_W = 0, halo_Cin = 0;
  bool ysum = false;   // layer-1 y-sum kernel (conv_ysum.cu)
  bool io_final = false;  // residual / output maps were built by the planner (plan_ysum_block)
  YsumParams yp;
  bool block = false;  // this conv1 and the next entry (conv2) run as one fused BasicBlock (conv_block.cu)
  bool skip = false;   // conv2 of a fused block: launched by the entry before it
  YsumBlockParams bp;
  uint16_t* d_B2 = nullptr;
  __nv_bfloat16* d_B = nullptr;
  float* d_bias = nullptr;
  uint32_t* d_ext_steps = nullptr;
};

void free_planned(PlannedConv& pc) {
  if (pc.d_B) cudaFree(pc.d_B);
  if (pc.d_bias) cudaFree(pc.d_bias);
  if (pc.d_ext_steps) cudaFree(pc.d_ext_steps);
  if (pc.d_B2) cudaFree(pc.d_B2);
  pc.d_B2 = nullptr;
  pc.d_B = nullptr;
  pc.d_bias = nullptr;
  pc.d_ext_steps = nullptr;
}

// One candidate K step before it is known which N tiles need it.
struct StepCand {
  int a_c0, b_k, dx, dy, map;
};

// Builds the GEMM description of one convolution (optionally with a 1x1 downsample of the
// block input fused as extra K steps reading `ds_in_hi`).
//   in_hi     : conv input  [b_pad][Hi*Wi][Cin] bf16
//   ds_in_hi  : block input [b_pad][Hx*Wx][Cx]  bf16 (only when gds != nullptr)
// Forms:
//   3x3, >= 16 output pixels : shifted boxes, rows = (instance, oy, ox), 4-D A maps
//   3x3, <= 4 output pixels  : dense, rows = instances, N = (out pixel, co), K = (in pixel, ci)
//   1x1                      : pointwise, rows = (instance, pixel); stride 2 reads parity phase 0
// For every N tile only the K steps whose weight block is not identically zero are kept:
// padding-only taps of the dense form and the off-diagonal blocks of grouped convs cost nothing.
int plan_conv(const ConvGeom& g, const float* w_oihw, const float* bias, const ConvGeom* gds,
              const float* wds_oihw, const float* bds, const __nv_bfloat16* in_hi,
              const __nv_bfloat16* ds_in_hi, int64_t b_pad, PlannedConv* out) {
  PlannedConv pc;
  memset(&pc.p, 0, sizeof(pc.p));
  const int Po = g.Ho * g.Wo, Pi = g.Hi * g.Wi;
  pc.Po = Po;
  if (g.Cin % 64 != 0 || g.Cout % 64 != 0) {
    set_error("plan_conv: channels %d -> %d must be multiples of 64", g.Cin, g.Cout);
    return CS_ERR_UNSUPPORTED;
  }
  if (gds && (gds->k != 1 || gds->Ho != g.Ho || gds->Wo != g.Wo || gds->Cout != g.Cout ||
              gds->Cin % 64 != 0)) {
    set_error("plan_conv: unsupported downsample geometry");
    return CS_ERR_UNSUPPORTED;
  }
  std::vector<float> bias_full;
  std::vector<uint16_t> B;
  std::vector<StepCand> cands;
  int64_t K_cat = 0;
  int rc;
  const int kk = g.k * g.k;

  // (grouped 3x3 convs: dense form up to g_dense_group_po output pixels, 64-wide N tiles)
  if (g.k == 1 || Po > (g.groups > 1 ? g_dense_group_po : g_dense_max_po)) {
    // ---- rows = (instance, oy, ox): shifted boxes (3x3) or pointwise (1x1)
    const bool pointwise = g.k == 1;
    if (!pointwise && (kGemmBM % Po != 0 || g.k != 3 || g.pad != 1)) {
      set_error("plan_conv: unsupported shifted-box geometry Ho=%d Wo=%d k=%d", g.Ho, g.Wo, g.k);
      return CS_ERR_UNSUPPORTED;
    }
    // (a strided 1x1 only reads phase (0,0), so odd input sizes such as 1x1 -> 1x1 are fine)
    if ((g.stride != 1 && g.stride != 2) ||
        (g.stride == 2 && !pointwise && (g.Hi != 2 * g.Ho || g.Wi != 2 * g.Wo)) ||
        (g.stride == 2 && kGemmBM % Po != 0)) {
      set_error("plan_conv: unsupported stride geometry Hi=%d Ho=%d s=%d", g.Hi, g.Ho, g.stride);
      return CS_ERR_UNSUPPORTED;
    }
    pc.dense = false;
    pc.BN = (g.groups > 1 && !pointwise) ? 64 : (g.Cout >= 256 ? 256 : g.Cout);
    pc.p.n_total = g.Cout;
    const int K_main = kk * g.Cin;
    K_cat = K_main + (gds ? gds->Cin : 0);
    int n_maps = 0;
    if (pointwise && g.stride == 1) {
      // A = the activation matrix itself: [b_pad * Pi rows][Cin]
      rc = make_mat_map_2d(&pc.p.a_map[0], in_hi, g.Cin, b_pad * Pi, g.Cin, kGemmBM);
      if (rc != CS_OK) return rc;
      n_maps = 1;
      for (int c0 = 0; c0 < g.Cin; c0 += 64) cands.push_back({c0, c0, 0, 0, 0});
    } else if (pointwise) {
      // 1x1 stride 2: parity phase (0,0) of the input, no shift
      pc.p.a_mode |= 1;
      pc.p.units_per_mtile = kGemmBM / Po;
      rc = make_act_map_4d(&pc.p.a_map[0], in_hi, g.Cin, g.Wo, g.Ho, b_pad, 2 * g.Cin,
                           (int64_t)2 * g.Wi * g.Cin, (int64_t)Pi * g.Cin, g.Wo, g.Ho, kGemmBM / Po);
      if (rc != CS_OK) return rc;
      n_maps = 1;
      for (int c0 = 0; c0 < g.Cin; c0 += 64) cands.push_back({c0, c0, 0, 0, 0});
    } else if (g.stride == 1) {
      pc.p.a_mode |= 1;
      pc.p.units_per_mtile = kGemmBM / Po;
      rc = make_act_map_4d(&pc.p.a_map[0], in_hi, g.Cin, g.Wi, g.Hi, b_pad, g.Cin,
                           (int64_t)g.Wi * g.Cin, (int64_t)Pi * g.Cin, g.Wo, g.Ho, kGemmBM / Po);
      if (rc != CS_OK) return rc;
      n_maps = 1;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx)
          for (int c0 = 0; c0 < g.Cin; c0 += 64)
            cands.push_back({c0, (dy * 3 + dx) * g.Cin + c0, dx - 1, dy - 1, 0});
    } else {
      pc.p.a_mode |= 0xF;
      pc.p.units_per_mtile = kGemmBM / Po;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          rc = make_act_map_4d(&pc.p.a_map[a * 2 + b], in_hi + ((int64_t)a * g.Wi + b) * g.Cin,
                               g.Cin, g.Wo, g.Ho, b_pad, 2 * g.Cin, (int64_t)2 * g.Wi * g.Cin,
                               (int64_t)Pi * g.Cin, g.Wo, g.Ho, kGemmBM / Po);
          if (rc != CS_OK) return rc;
        }
      n_maps = 4;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx)
          for (int c0 = 0; c0 < g.Cin; c0 += 64) {
            // iy = 2*oy + dy - 1: dy=0 -> phase 1 shift -1; dy=1 -> phase 0; dy=2 -> phase 1
            const int pa = (dy == 1) ? 0 : 1, pb = (dx == 1) ? 0 : 1;
            cands.push_back({c0, (dy * 3 + dx) * g.Cin + c0, dx == 0 ? -1 : 0, dy == 0 ? -1 : 0,
                             pa * 2 + pb});
          }
    }
    if (gds) {
      if (n_maps >= 4) { set_error("plan_conv: no free A map for the downsample"); return CS_ERR_UNSUPPORTED; }
      if (gds->stride == 1) {
        rc = make_mat_map_2d(&pc.p.a_map[n_maps], ds_in_hi, gds->Cin, b_pad * gds->Hi * gds->Wi,
                             gds->Cin, kGemmBM);
      } else {
        pc.p.a_mode |= 1 << n_maps;
        pc.p.units_per_mtile = kGemmBM / Po;
        rc = make_act_map_4d(&pc.p.a_map[n_maps], ds_in_hi, gds->Cin, g.Wo, g.Ho, b_pad, 2 * gds->Cin,
                             (int64_t)2 * gds->Wi * gds->Cin, (int64_t)gds->Hi * gds->Wi * gds->Cin,
                             g.Wo, g.Ho, kGemmBM / Po);
      }
      if (rc != CS_OK) return rc;
      for (int c0 = 0; c0 < gds->Cin; c0 += 64) cands.push_back({c0, K_main + c0, 0, 0, n_maps});
    }
    // B[co][tap*Cin + ci] (+ [K_main + ci] for the downsample)
    B.assign((size_t)g.Cout * K_cat, 0);
    for (int co = 0; co < g.Cout; ++co) {
      for (int ci = 0; ci < g.Cin; ++ci)
        for (int t = 0; t < kk; ++t)
          B[(size_t)co * K_cat + t * g.Cin + ci] = f32_to_bf16_rn(w_oihw[((size_t)co * g.Cin + ci) * kk + t]);
      if (gds)
        for (int ci = 0; ci < gds->Cin; ++ci)
          B[(size_t)co * K_cat + K_main + ci] = f32_to_bf16_rn(wds_oihw[(size_t)co * gds->Cin + ci]);
    }
    bias_full.resize(g.Cout);
    for (int co = 0; co < g.Cout; ++co) bias_full[co] = bias[co] + (gds ? bds[co] : 0.f);
  } else {
    // ---- dense form: rows = instances, N = (out pixel, co), K = (in pixel, ci)
    pc.dense = true;
    const int N_total = Po * g.Cout;
    pc.BN = (N_total % 256 == 0 && (g_dense_bn == 256 || g.Cout >= 256)) ? 256 : (N_total % 128 == 0 ? 128 : 64);
    if (g.groups > 1 && Po > 4) pc.BN = 64;     // one (pixel, 64-channel block) per N tile
    pc.p.units_per_mtile = kGemmBM;
    pc.p.n_total = N_total;
    const int64_t K_main = (int64_t)Pi * g.Cin;
    const int64_t K_ds = gds ? (int64_t)gds->Hi * gds->Wi * gds->Cin : 0;
    K_cat = K_main + K_ds;
    if (K_cat > 65535) { set_error("plan_conv: dense K %lld too large", (long long)K_cat); return CS_ERR_UNSUPPORTED; }
    rc = make_mat_map_2d(&pc.p.a_map[0], in_hi, K_main, b_pad, K_main, kGemmBM);
    if (rc != CS_OK) return rc;
    if (gds) {
      rc = make_mat_map_2d(&pc.p.a_map[1], ds_in_hi, K_ds, b_pad, K_ds, kGemmBM);
      if (rc != CS_OK) return rc;
    }
    B.assign((size_t)N_total * K_cat, 0);
    for (int oy = 0; oy < g.Ho; ++oy)
      for (int ox = 0; ox < g.Wo; ++ox) {
        const int po = oy * g.Wo + ox;
        for (int dy = 0; dy < g.k; ++dy)
          for (int dx = 0; dx < g.k; ++dx) {
            int iy = oy * g.stride - g.pad + dy, ix = ox * g.stride - g.pad + dx;
            if (iy < 0 || iy >= g.Hi || ix < 0 || ix >= g.Wi) continue;
            const int pi = iy * g.Wi + ix;
            for (int co = 0; co < g.Cout; ++co) {
              uint16_t* dst = &B[((size_t)po * g.Cout + co) * K_cat + (size_t)pi * g.Cin];
              const float* src = &w_oihw[(size_t)co * g.Cin * kk + dy * g.k + dx];
              for (int ci = 0; ci < g.Cin; ++ci) dst[ci] = f32_to_bf16_rn(src[(size_t)ci * kk]);
            }
          }
        if (gds) {
          const int pi = (oy * gds->stride) * gds->Wi + ox * gds->stride;
          for (int co = 0; co < g.Cout; ++co) {
            uint16_t* dst = &B[((size_t)po * g.Cout + co) * K_cat + K_main + (size_t)pi * gds->Cin];
            for (int ci = 0; ci < gds->Cin; ++ci) dst[ci] = f32_to_bf16_rn(wds_oihw[(size_t)co * gds->Cin + ci]);
          }
        }
      }
    for (int64_t kb = 0; kb < K_cat; kb += 64)
      cands.push_back(kb < K_main ? StepCand{(int)kb, (int)kb, 0, 0, 0}
                                  : StepCand{(int)(kb - K_main), (int)kb, 0, 0, 1});
    bias_full.resize(N_total);
    for (int po = 0; po < Po; ++po)
      for (int co = 0; co < g.Cout; ++co) bias_full[(size_t)po * g.Cout + co] = bias[co] + (gds ? bds[co] : 0.f);
  }

  // ---- K steps per N tile: drop weight blocks that are identically zero
  pc.p.num_n_tiles = pc.p.n_total / pc.BN;
  std::vector<std::vector<KStep>> per_tile(pc.p.num_n_tiles);
  for (int nt = 0; nt < pc.p.num_n_tiles; ++nt) {
    for (const StepCand& c : cands) {
      bool any = false;
      for (int n = nt * pc.BN; n < (nt + 1) * pc.BN && !any; ++n) {
        const uint16_t* row = &B[(size_t)n * K_cat + c.b_k];
        for (int q = 0; q < 64; ++q)
          if ((row[q] & 0x7fff) != 0) { any = true; break; }
      }
      if (any) per_tile[nt].push_back(KStep::make(c.a_c0, c.b_k, c.dx, c.dy, c.map));
    }
    if (per_tile[nt].empty()) per_tile[nt].push_back(KStep::make(0, 0, 0, 0, 0));  // all-zero weights
    if ((int)per_tile[nt].size() > kMaxSteps) {
      set_error("plan_conv: %d K steps exceed %d", (int)per_tile[nt].size(), kMaxSteps);
      return CS_ERR_UNSUPPORTED;
    }
  }
  bool same = true;
  for (int nt = 1; nt < pc.p.num_n_tiles && same; ++nt) {
    same = per_tile[nt].size() == per_tile[0].size();
    for (size_t i = 0; same && i < per_tile[nt].size(); ++i) same = per_tile[nt][i].v == per_tile[0][i].v;
  }
  pc.p.n_variants = same ? 1 : pc.p.num_n_tiles;
  pc.p.ext_steps = nullptr;
  if (pc.p.n_variants > kMaxVariants) {
    // wide grouped convs: the per-tile step lists live in device memory
    std::vector<uint32_t> ext((size_t)pc.p.n_variants * (kMaxSteps + 1), 0u);
    for (int v = 0; v < pc.p.n_variants; ++v) {
      ext[(size_t)v * (kMaxSteps + 1)] = (uint32_t)per_tile[v].size();
      for (size_t i = 0; i < per_tile[v].size(); ++i) ext[(size_t)v * (kMaxSteps + 1) + 1 + i] = per_tile[v][i].v;
    }
    CS_CUDA(cudaMalloc(&pc.d_ext_steps, ext.size() * sizeof(uint32_t)));
    CS_CUDA(cudaMemcpy(pc.d_ext_steps, ext.data(), ext.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    pc.p.ext_steps = pc.d_ext_steps;
  } else {
    for (int v = 0; v < pc.p.n_variants; ++v) {
      pc.p.n_steps[v] = (int)per_tile[v].size();
      for (size_t i = 0; i < per_tile[v].size(); ++i) pc.p.steps[v][i] = per_tile[v][i];
    }
  }

  CS_CUDA(cudaMalloc(&pc.d_B, B.size() * sizeof(uint16_t)));
  CS_CUDA(cudaMemcpy(pc.d_B, B.data(), B.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  CS_CUDA(cudaMalloc(&pc.d_bias, bias_full.size() * sizeof(float)));
  CS_CUDA(cudaMemcpy(pc.d_bias, bias_full.data(), bias_full.size() * sizeof(float), cudaMemcpyHostToDevice));
  pc.p.cluster = pc.BN >= 128 ? g_cluster : 1;   // pairs do not pay off at N = 64
  rc = make_mat_map_2d(&pc.p.b_map, pc.d_B, K_cat, pc.p.n_total, K_cat, pc.BN / pc.p.cluster);
  if (rc != CS_OK) { free_planned(pc); return rc; }
  pc.p.bias = pc.d_bias;
  if (!pc.dense && !gds && g.k == 3 && g.groups == 1 && g.stride == 1 && g.Hi == g.Wi &&
      ysum_supported(g.Wi, g.Cin, g.Cout) && !g_disable_ysum) {
    // layer 1: the shifted-box A map of this plan + weights regrouped per kernel column
    std::vector<uint16_t> B2(576 * 64);
    pack_ysum_weights(w_oihw, B2.data());
    CS_CUDA(cudaMalloc(&pc.d_B2, B2.size() * sizeof(uint16_t)));
    CS_CUDA(cudaMemcpy(pc.d_B2, B2.data(), B2.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    pc.yp.a_map = pc.p.a_map[0];
    rc = make_act_map_4d(&pc.yp.a_box_map, in_hi, g.Cin, g.Wi, g.Hi, b_pad, g.Cin, (int64_t)g.Wi * g.Cin,
                         (int64_t)Pi * g.Cin, g.Wi + 2, g.Hi, kGemmBM / Po);
    if (rc != CS_OK) { free_planned(pc); return rc; }
    pc.yp.cluster = g_ysum_pairs ? g_cluster : 1;
    rc = make_mat_map_2d(&pc.yp.b_map, pc.d_B2, 64, 576, 64, 192 / pc.yp.cluster);
    if (rc != CS_OK) { free_planned(pc); return rc; }
    pc.yp.bias = pc.d_bias;
    pc.ysum = true;
  } else if (!pc.dense && g.k == 3 && g.groups == 1 && g.stride == 1 && g.Hi == g.Wi &&
      halo_supported(g.Wi, g.Cin, g.Cout) && !g_disable_halo &&
      (!gds || (!g_disable_halo_ds && g.Wi == 4 && gds->Cin == 64 && gds->stride == 2 && gds->Hi == 8 && gds->Wi == 8))) {
    rc = make_act_map_halo(&pc.hp.a_map, in_hi, g.Cin, g.Wi, g.Hi, b_pad, 1);
    if (rc != CS_OK) { free_planned(pc); return rc; }
    pc.hp.has_ds = 0;
    if (gds) {   // BasicBlock shortcut of the layer-2 entry as one more K step of the halo kernel
      rc = make_act_map_halo_ds(&pc.hp.ds_map, ds_in_hi, gds->Cin, g.Wi, g.Hi, b_pad);
      if (rc != CS_OK) { free_planned(pc); return rc; }
      pc.hp.has_ds = 1;
    }
    pc.hp.cluster = pc.BN >= 128 ? g_cluster : 1;
    rc = make_mat_map_2d(&pc.hp.b_map, pc.d_B, K_cat, pc.p.n_total, K_cat, pc.BN / pc.hp.cluster);
    if (rc != CS_OK) { free_planned(pc); return rc; }
    pc.hp.bias = pc.d_bias;
    pc.halo = true;
    pc.halo_W = g.Wi;
    pc.halo_Cin = g.Cin;
  }
  *out = pc;
  return CS_OK;
}

// One 64-channel block of a grouped stride-1 3x3 conv on 8x8 maps, through the y-sum kernel.
// A grouped conv whose groups do not straddle 64-channel blocks (ResNeXt layer 1: 32 groups of 4
// or 8 channels) is C/64 independent 64 -> 64 convolutions on channel slices of the same tensors:
// exactly the layer-1 shape of the BasicBlock nets, so each slice runs in conv_ysum_kernel on
// strided views (channel extent 64, pixel pitch C) instead of the generic shifted-box kernel
// (nine boxes per M tile; 774 us per 37 888 instances at 21 % tensor pipe, L2 -> SM bound, ncu
// r02_resnext50_batch).  w_oihw is the DENSE [C][C][3][3] weight, blk the block index; in / out are
// the full [b_pad][64][C] tensors.  The output maps are final (no finalize_io_maps).
int plan_ysum_block(int C, int blk, const float* w_oihw, const float* bias, const __nv_bfloat16* in_hi,
                    __nv_bfloat16* out_hi, int relu, int64_t b_pad, PlannedConv* out) {
  PlannedConv pc;
  memset(&pc.p, 0, sizeof(pc.p));
  memset(&pc.yp, 0, sizeof(pc.yp));
  pc.Po = 64;
  pc.BN = 64;
  pc.ysum = true;
  pc.io_final = true;
  std::vector<float> wb((size_t)64 * 64 * 9);
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < 64; ++ci)
      memcpy(&wb[((size_t)co * 64 + ci) * 9], &w_oihw[((size_t)(blk * 64 + co) * C + blk * 64 + ci) * 9], 9 * sizeof(float));
  std::vector<uint16_t> B2(576 * 64);
  pack_ysum_weights(wb.data(), B2.data());
  CS_CUDA(cudaMalloc(&pc.d_B2, B2.size() * sizeof(uint16_t)));
  CS_CUDA(cudaMemcpy(pc.d_B2, B2.data(), B2.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  CS_CUDA(cudaMalloc(&pc.d_bias, 64 * sizeof(float)));
  CS_CUDA(cudaMemcpy(pc.d_bias, bias + blk * 64, 64 * sizeof(float), cudaMemcpyHostToDevice));
  const __nv_bfloat16* in = in_hi + blk * 64;
  __nv_bfloat16* o = out_hi + blk * 64;
  int rc = make_act_map_4d(&pc.yp.a_map, in, 64, 8, 8, b_pad, C, (int64_t)8 * C, (int64_t)64 * C, 8, 8, 2);
  if (rc == CS_OK)
    rc = make_act_map_4d(&pc.yp.a_box_map, in, 64, 8, 8, b_pad, C, (int64_t)8 * C, (int64_t)64 * C, 10, 8, 2);
  pc.yp.cluster = g_ysum_pairs ? g_cluster : 1;
  if (rc == CS_OK) rc = make_mat_map_2d(&pc.yp.b_map, pc.d_B2, 64, 576, 64, 192 / pc.yp.cluster);
  if (rc == CS_OK) rc = make_mat_map_2d(&pc.p.out_hi_map, o, 64, b_pad * 64, C, kGemmBM);
  if (rc != CS_OK) { free_planned(pc); return rc; }
  pc.yp.bias = pc.d_bias;
  pc.p.bias = pc.d_bias;
  pc.p.out_hi = o;
  pc.p.relu = relu;
  *out = pc;
  return CS_OK;
}

// Builds the TMA views of the residual / output tensors once their pointers are known.
int finalize_io_maps(PlannedConv& pc, int64_t b_pad) {
  int rc = CS_OK;
  if (pc.io_final) return CS_OK;
  if (pc.halo) {
    const int W = pc.halo_W, C = pc.p.n_total;
    if (pc.p.res_hi && (rc = make_act_map_halo(&pc.hp.res_hi_map, pc.p.res_hi, C, W, W, b_pad, 0))) return rc;
    if (pc.p.res_lo && (rc = make_act_map_halo(&pc.hp.res_lo_map, pc.p.res_lo, C, W, W, b_pad, 0))) return rc;
    if (pc.p.out_hi && (rc = make_act_map_halo(&pc.hp.out_hi_map, pc.p.out_hi, C, W, W, b_pad, 0))) return rc;
    if (pc.p.out_lo && (rc = make_act_map_halo(&pc.hp.out_lo_map, pc.p.out_lo, C, W, W, b_pad, 0))) return rc;
    return CS_OK;
  }
  const int64_t rows = pc.dense ? b_pad : b_pad * pc.Po;
  const int64_t N = pc.p.n_total;
  if (pc.p.res_hi && (rc = make_mat_map_2d(&pc.p.res_hi_map, pc.p.res_hi, N, rows, N, kGemmBM))) return rc;
  if (pc.p.res_lo && (rc = make_mat_map_2d(&pc.p.res_lo_map, pc.p.res_lo, N, rows, N, kGemmBM))) return rc;
  if (pc.p.out_hi && (rc = make_mat_map_2d(&pc.p.out_hi_map, pc.p.out_hi, N, rows, N, kGemmBM))) return rc;
  if (pc.p.out_lo && (rc = make_mat_map_2d(&pc.p.out_lo_map, pc.p.out_lo, N, rows, N, kGemmBM))) return rc;
  return CS_OK;
}

// Launches one planned convolution for `count` instances (pointers / relu taken from pc.p).
// inst_base (y-sum layers only): the launch covers instances [inst_base, inst_base + count) of
// the planned buffers -- a sub-batch of the forward batch.
// conv1 / conv2 of a BasicBlock without downsample were just pushed: if both run in the y-sum
// kernel on CTA pairs with the bf16 stream only, the pair becomes one fused launch.
void fuse_basic_block(std::vector<PlannedConv>& layers, const __nv_bfloat16* x_in) {
  if (!g_block_fuse || g_l1_sub > 0 || layers.size() < 2) return;
  PlannedConv& c1 = layers[layers.size() - 2];
  PlannedConv& c2 = layers[layers.size() - 1];
  if (!c1.ysum || !c2.ysum || c1.io_final || c2.io_final) return;
  if (c2.p.res_hi != x_in) return;       // the residual must be the block input
  if (c1.yp.cluster != 2 || c2.yp.cluster != 2) return;
  if (c1.p.res_hi || c1.p.res_lo || c1.p.out_lo || !c1.p.relu) return;
  if (!c2.p.res_hi || c2.p.res_lo || c2.p.out_lo || !c2.p.out_hi || !c2.p.relu) return;
  c1.bp.x_box_map = c1.yp.a_box_map;     // the block input is conv1's operand and conv2's residual
  c1.bp.b1_map = c1.yp.b_map;
  c1.bp.b2_map = c2.yp.b_map;
  c1.bp.out_map = c2.p.out_hi_map;
  c1.bp.bias1 = c1.yp.bias;
  c1.bp.bias2 = c2.yp.bias;
  c1.block = true;
  c2.skip = true;
}

int launch_planned(const PlannedConv& pc, int64_t count, float* out_f32, cudaStream_t st,
                   int reverse = 0, int64_t inst_base = 0) {
  if (pc.block) {
    YsumBlockParams bp = pc.bp;
    bp.reverse = reverse;
    bp.tile_base = (int)(inst_base / 2);
    bp.num_m_tiles = (int)ceil_div<int64_t>(count, 2);
    return launch_ysum_block(bp, st);
  }
  if (pc.ysum) {
    YsumParams yp = pc.yp;
    yp.res_hi = pc.p.res_hi; yp.res_lo = pc.p.res_lo;
    yp.out_hi = pc.p.out_hi; yp.out_lo = pc.p.out_lo;
    yp.res_hi_map = pc.p.res_hi_map; yp.res_lo_map = pc.p.res_lo_map;
    yp.out_hi_map = pc.p.out_hi_map; yp.out_lo_map = pc.p.out_lo_map;
    yp.out_f32 = out_f32;
    yp.relu = pc.p.relu;
    yp.n_inst = inst_base + count;
    yp.reverse = reverse;
    yp.tile_base = (int)(inst_base / 2);
    yp.num_m_tiles = (int)ceil_div<int64_t>(count, 2);
    return launch_conv_ysum(yp, st);
  }
  if (pc.halo) {
    HaloParams hp = pc.hp;
    hp.res_hi = pc.p.res_hi; hp.res_lo = pc.p.res_lo;
    hp.out_hi = pc.p.out_hi; hp.out_lo = pc.p.out_lo;
    hp.out_f32 = out_f32;
    hp.relu = pc.p.relu;
    hp.n_inst = count;
    hp.reverse = reverse;
    hp.num_m_tiles = (int)ceil_div<int64_t>(count, kGemmBM / pc.Po);
    return launch_conv_halo(hp, pc.halo_W, pc.halo_Cin, st);
  }
  GemmParams p = pc.p;
  int64_t rows = pc.dense ? count : count * pc.Po;
  p.m_valid = rows;
  p.num_m_tiles = (int)ceil_div<int64_t>(rows, kGemmBM);
  p.out_f32 = out_f32;
  p.reverse = reverse;
  return launch_conv_gemm(p, pc.BN, st);
}

struct TcPlan {
  int tile = 0;
  int64_t max_batch = 0, b_pad = 0;
  void* ws = nullptr;
  std::vector<PlannedConv> layers;  // every conv launch of the encoder in order
  __nv_bfloat16* x_hi[2] = {nullptr, nullptr};   // block input / output (ping-pong)
  __nv_bfloat16* x_lo[2] = {nullptr, nullptr};
  __nv_bfloat16* mid[2] = {nullptr, nullptr};    // conv1 / conv2 outputs inside a block
  const __nv_bfloat16* x4_hi = nullptr;
  const __nv_bfloat16* x4_lo = nullptr;
  int P4 = 1, C4 = 512;
  uint16_t* d_stem_w2 = nullptr; // window-form stem weights (stem_win.cu)
  uint16_t* d_stem_w3 = nullptr; // weights-stationary stem weights (stem_ts.cu)
  uint16_t* d_lut = nullptr;     // [3][256] bf16 normalisation LUT
  ~TcPlan() {
    for (auto& l : layers) free_planned(l);
    if (d_stem_w2) cudaFree(d_stem_w2);
    if (d_stem_w3) cudaFree(d_stem_w3);
    if (d_lut) cudaFree(d_lut);
  }
};

int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

constexpr int64_t kFp32Chunk = 4096;

}  // namespace
}  // namespace cs

using namespace cs;

struct cs_model {
  int arch = 0;
  std::vector<ConvW> convs;
  std::vector<BlockDesc> blocks;
  int feat_dim = 512;
  float* d_fc_w = nullptr;
  float* d_fc_b = nullptr;
  std::unique_ptr<TcPlan> plan;
  int64_t last_launches = 0;
};

namespace cs {
namespace {

// Largest block-level activation (elements per instance) at this tile size: x buffers hold
// block inputs / outputs, mid buffers the conv1 / conv2 outputs.
// Output size of the stem conv (7x7, stride 2, pad 3) and of the 3x3 / 2 max pool behind it
// (model/resnet.py:236-239): 32 -> 16 -> 8 for tiles, 299 -> 150 -> 75 for whole images.
int stem_conv_size(int S) { return (S + 6 - 7) / 2 + 1; }
int stem_pool_size(int S) { return (stem_conv_size(S) + 2 - 3) / 2 + 1; }

void act_sizes(const cs_model* m, int tile, int64_t* x_elems, int64_t* mid_elems) {
  int H = stem_pool_size(tile);
  int64_t xe = (int64_t)H * H * 64, me = 0;
  for (const BlockDesc& b : m->blocks) {
    const int Ho = (H + 2 - 3) / b.stride + 1;
    // conv1 keeps the input resolution in a Bottleneck (stride sits on conv2), halves it in a BasicBlock
    const int64_t m1 = b.bottleneck ? (int64_t)H * H * b.width : (int64_t)Ho * Ho * b.width;
    const int64_t m2 = (int64_t)Ho * Ho * b.width;
    if (m1 > me) me = m1;
    if (m2 > me) me = m2;
    const int64_t y = (int64_t)Ho * Ho * b.cout;
    if (y > xe) xe = y;
    H = Ho;
  }
  *x_elems = xe;
  *mid_elems = me;
}

int64_t fp32_floats_per_inst(const cs_model* m, int tile) {
  int64_t xe, me;
  act_sizes(m, tile, &xe, &me);
  // input tile + stem conv output + x, y, ds + two mids
  const int64_t hc = stem_conv_size(tile);
  return (int64_t)3 * tile * tile + hc * hc * 64 + 3 * xe + 2 * me;
}

int build_tc_plan(cs_model* m, int tile, int64_t max_batch, void* ws, int64_t ws_bytes) {
  auto plan = std::make_unique<TcPlan>();
  plan->tile = tile;
  plan->max_batch = max_batch;
  plan->b_pad = round_up(max_batch, kGemmBM);
  plan->ws = ws;
  int64_t xe, me;
  act_sizes(m, tile, &xe, &me);
  const int64_t xb = round_up(plan->b_pad * xe * 2, 1024), mb = round_up(plan->b_pad * me * 2, 1024);
  if (ws_bytes < g_xbufs * xb + 2 * mb + 1024) {
    set_error("bf16 workspace too small: %lld < %lld", (long long)ws_bytes,
              (long long)(g_xbufs * xb + 2 * mb + 1024));
    return CS_ERR_WORKSPACE;
  }
  uintptr_t base = round_up((int64_t)(uintptr_t)ws, 1024);
  for (int i = 0; i < 2; ++i) {
    plan->x_hi[i] = reinterpret_cast<__nv_bfloat16*>(base + i * xb);
    plan->x_lo[i] = g_no_lo ? nullptr : reinterpret_cast<__nv_bfloat16*>(base + (2 + i) * xb);
    plan->mid[i] = reinterpret_cast<__nv_bfloat16*>(base + g_xbufs * xb + i * mb);
  }
  const int64_t bp = plan->b_pad;
  auto push = [&](PlannedConv& pc, const __nv_bfloat16* rh, const __nv_bfloat16* rl, __nv_bfloat16* oh,
                  __nv_bfloat16* ol, int relu) {
    pc.p.res_hi = rh; pc.p.res_lo = g_no_lo ? nullptr : rl; pc.p.out_hi = oh;
    pc.p.out_lo = g_no_lo ? nullptr : ol; pc.p.relu = relu;
    int rc = finalize_io_maps(pc, bp);
    if (rc != CS_OK) { free_planned(pc); return rc; }
    plan->layers.push_back(pc);
    return (int)CS_OK;
  };

  int xi = 0;   // x lives in (x_hi[xi], x_lo[xi]); the block writes (x_hi[1-xi], x_lo[1-xi])
  bool x_has_lo = false;   // the stem writes the bf16 stream only (see stem_win.cu)
  int H = tile / 4, W = tile / 4, C = 64;
  for (const BlockDesc& b : m->blocks) {
    const int Ho = (H + 2 - 3) / b.stride + 1, Wo = (W + 2 - 3) / b.stride + 1;
    const ConvW& c1 = m->convs[b.conv1];
    const ConvW& c2 = m->convs[b.conv2];
    PlannedConv p;
    int rc;
    if (!b.bottleneck) {
      ConvGeom g1{H, W, C, Ho, Wo, b.cout, 3, b.stride, 1, 1};
      ConvGeom g2{Ho, Wo, b.cout, Ho, Wo, b.cout, 3, 1, 1, 1};
      rc = plan_conv(g1, c1.w.data(), c1.b.data(), nullptr, nullptr, nullptr, plan->x_hi[xi], nullptr, bp, &p);
      if (rc != CS_OK) return rc;
      if ((rc = push(p, nullptr, nullptr, plan->mid[0], nullptr, 1)) != CS_OK) return rc;
      if (b.ds >= 0) {   // downsample fused into conv2 as extra K steps
        const ConvW& cd = m->convs[b.ds];
        ConvGeom gd{H, W, C, Ho, Wo, b.cout, 1, b.stride, 0, 1};
        rc = plan_conv(g2, c2.w.data(), c2.b.data(), &gd, cd.w.data(), cd.b.data(), plan->mid[0],
                       plan->x_hi[xi], bp, &p);
        if (rc != CS_OK) return rc;
        rc = push(p, nullptr, nullptr, plan->x_hi[1 - xi], plan->x_lo[1 - xi], 1);
      } else {
        rc = plan_conv(g2, c2.w.data(), c2.b.data(), nullptr, nullptr, nullptr, plan->mid[0], nullptr, bp, &p);
        if (rc != CS_OK) return rc;
        rc = push(p, plan->x_hi[xi], x_has_lo ? plan->x_lo[xi] : nullptr, plan->x_hi[1 - xi],
                  plan->x_lo[1 - xi], 1);
        if (rc == CS_OK) fuse_basic_block(plan->layers, plan->x_hi[xi]);
      }
      if (rc != CS_OK) return rc;
    } else {
      const ConvW& c3 = m->convs[b.conv3];
      ConvGeom g1{H, W, C, H, W, b.width, 1, 1, 0, 1};
      ConvGeom g2{H, W, b.width, Ho, Wo, b.width, 3, b.stride, 1, c2.groups};
      ConvGeom g3{Ho, Wo, b.width, Ho, Wo, b.cout, 1, 1, 0, 1};
      rc = plan_conv(g1, c1.w.data(), c1.b.data(), nullptr, nullptr, nullptr, plan->x_hi[xi], nullptr, bp, &p);
      if (rc != CS_OK) return rc;
      if ((rc = push(p, nullptr, nullptr, plan->mid[0], nullptr, 1)) != CS_OK) return rc;
      const int cin_g = b.width / c2.groups;
      if (c2.groups > 1 && H == 8 && W == 8 && b.stride == 1 && b.width % 64 == 0 && 64 % cin_g == 0 &&
          !g_disable_ysum && !g_no_group_ysum) {
        // grouped 3x3 of ResNeXt layer 1: one y-sum conv per 64-channel block
        for (int blk = 0; blk < b.width / 64; ++blk) {
          rc = plan_ysum_block(b.width, blk, c2.w.data(), c2.b.data(), plan->mid[0], plan->mid[1], 1, bp, &p);
          if (rc != CS_OK) return rc;
          plan->layers.push_back(p);
        }
      } else {
        rc = plan_conv(g2, c2.w.data(), c2.b.data(), nullptr, nullptr, nullptr, plan->mid[0], nullptr, bp, &p);
        if (rc != CS_OK) return rc;
        if ((rc = push(p, nullptr, nullptr, plan->mid[1], nullptr, 1)) != CS_OK) return rc;
      }
      if (b.ds >= 0) {   // downsample as its own launch into y, then conv3 adds y in place
        const ConvW& cd = m->convs[b.ds];
        ConvGeom gd{H, W, C, Ho, Wo, b.cout, 1, b.stride, 0, 1};
        rc = plan_conv(gd, cd.w.data(), cd.b.data(), nullptr, nullptr, nullptr, plan->x_hi[xi], nullptr, bp, &p);
        if (rc != CS_OK) return rc;
        if ((rc = push(p, nullptr, nullptr, plan->x_hi[1 - xi], plan->x_lo[1 - xi], 0)) != CS_OK) return rc;
        rc = plan_conv(g3, c3.w.data(), c3.b.data(), nullptr, nullptr, nullptr, plan->mid[1], nullptr, bp, &p);
        if (rc != CS_OK) return rc;
        rc = push(p, plan->x_hi[1 - xi], plan->x_lo[1 - xi], plan->x_hi[1 - xi], plan->x_lo[1 - xi], 1);
      } else {
        rc = plan_conv(g3, c3.w.data(), c3.b.data(), nullptr, nullptr, nullptr, plan->mid[1], nullptr, bp, &p);
        if (rc != CS_OK) return rc;
        rc = push(p, plan->x_hi[xi], x_has_lo ? plan->x_lo[xi] : nullptr, plan->x_hi[1 - xi],
                  plan->x_lo[1 - xi], 1);
      }
      if (rc != CS_OK) return rc;
    }
    xi = 1 - xi;
    x_has_lo = true;
    H = Ho; W = Wo; C = b.cout;
  }
  plan->x4_hi = plan->x_hi[xi];
  plan->x4_lo = g_no_lo ? nullptr : plan->x_lo[xi];
  plan->P4 = H * W;
  plan->C4 = C;
  if (tile == 32) {
    std::vector<uint16_t> lut(768);
    float lut_f[768];
    get_norm_lut_host(lut_f);
    for (int i = 0; i < 768; ++i) lut[i] = f32_to_bf16_rn(lut_f[i]);
    std::vector<uint16_t> sw2(stem_win_weight_bytes() / 2);
    pack_stem_weights_win(m->convs[0].w.data(), sw2.data());
    CS_CUDA(cudaMalloc(&plan->d_stem_w2, sw2.size() * 2));
    CS_CUDA(cudaMemcpy(plan->d_stem_w2, sw2.data(), sw2.size() * 2, cudaMemcpyHostToDevice));
    std::vector<uint16_t> sw3(stem_ts_weight_bytes() / 2);
    pack_stem_weights_ts(m->convs[0].w.data(), sw3.data());
    CS_CUDA(cudaMalloc(&plan->d_stem_w3, sw3.size() * 2));
    CS_CUDA(cudaMemcpy(plan->d_stem_w3, sw3.data(), sw3.size() * 2, cudaMemcpyHostToDevice));
    CS_CUDA(cudaMalloc(&plan->d_lut, lut.size() * 2));
    CS_CUDA(cudaMemcpy(plan->d_lut, lut.data(), lut.size() * 2, cudaMemcpyHostToDevice));
  }
  m->plan = std::move(plan);
  return CS_OK;
}

int run_tc_batch(cs_model* m, const StemArgs& stem_in, int64_t count, float* prob_out,
                 float* logits_out, float* feat_out, cudaStream_t st) {
  TcPlan& pl = *m->plan;
  StemArgs sa = stem_in;
  sa.count = count;
  sa.w = m->convs[0].d_w32;
  sa.bias = m->convs[0].d_b;
  sa.out_hi = pl.x_hi[0];
  sa.out_lo = nullptr;   // bf16 stream only; the first block's residual add reads x_hi alone
  int rc;
  // Optional (CELLSEG_L1_SUB, off by default -- see g_l1_sub): layer 1 in sub-batches.  A forward
  // batch of 75 776 instances keeps 621 MB per 8x8x64 tensor: the stem output and the six layer-1
  // convs round-trip through HBM (9.6 of the 21.5 GB a batch moves, the residual convs at
  // 4.9 TB/s = 75 % of the copy peak; ncu r02_full_batch).  The convs of layer 1 are independent
  // per instance, so stem + layer 1 can run for g_l1_sub instances at a time -- three live tensors
  // of 39 MB stay in the 126 MB L2 -- with only the layer-1 output going to HBM.
  size_t n_ysum = 0;
  while (n_ysum < pl.layers.size() && pl.layers[n_ysum].ysum) ++n_ysum;
  const bool sub_batched = g_l1_sub > 0 && pl.tile == 32 && n_ysum > 0 && count > g_l1_sub;
  const int64_t step = sub_batched ? g_l1_sub : count;
  for (int64_t i0 = 0; i0 < count; i0 += step) {
    const int64_t cnt = count - i0 < step ? count - i0 : step;
    StemArgs ss = sa;
    ss.count = cnt;
    ss.inst_begin = sa.inst_begin + i0;
    if (sa.x) ss.x = sa.x + i0 * 3 * pl.tile * pl.tile;
    ss.out_hi = sa.out_hi + i0 * (int64_t)(pl.tile / 4) * (pl.tile / 4) * 64;
    if (pl.tile == 32)
      rc = g_stem_win ? launch_stem_win(ss, pl.d_stem_w2, pl.d_lut, st) : launch_stem_ts(ss, pl.d_stem_w3, st);
    else
      rc = launch_stem_bf16(ss, st);
    if (rc != CS_OK) return rc;
    m->last_launches++;
    if (!sub_batched) break;
    for (size_t li = 0; li < n_ysum; ++li) {
      rc = launch_planned(pl.layers[li], cnt, nullptr, st, g_snake ? (int)((li + 1) & 1) : 0, i0);
      if (rc != CS_OK) return rc;
      m->last_launches++;
    }
  }
  int li = 0, launches = 0;
  for (PlannedConv& pc : pl.layers) {
    // snake order: odd launches walk their tiles backwards, starting with the rows the previous
    // launch wrote last (still L2 resident when a batch's activations exceed L2)
    ++li;
    if (sub_batched && (size_t)li <= n_ysum) continue;     // done above, per sub-batch
    if (pc.skip) continue;                                  // conv2 of a fused BasicBlock
    ++launches;
    rc = launch_planned(pc, count, nullptr, st, g_snake ? (launches & 1) : 0);
    if (rc != CS_OK) return rc;
    m->last_launches++;
  }
  rc = launch_head_bf16(pl.x4_hi, pl.x4_lo, count, pl.P4, pl.C4, m->d_fc_w, m->d_fc_b, prob_out,
                        logits_out, feat_out, st);
  if (rc != CS_OK) return rc;
  m->last_launches++;
  return CS_OK;
}

// fp32 path for `count` instances whose normalised NCHW tiles are at x_in.
// maps_out (nullable): four device pointers receiving the NHWC fp32 outputs of layer1..layer4
// (x1..x4 of resnet_forward(return_intermediate=True), model/resnet.py:240-246).
int run_fp32_batch(cs_model* m, const float* x_in, int tile, int64_t count, float* ws_f,
                   float* prob_out, float* logits_out, float* feat_out, cudaStream_t st,
                   float* const* maps_out = nullptr) {
  const int S = tile;
  const int Hc = stem_conv_size(S), Hp = stem_pool_size(S);
  int64_t xe, me;
  act_sizes(m, tile, &xe, &me);
  float* c1 = ws_f;
  float* q = c1 + count * (int64_t)Hc * Hc * 64;
  float* xb[3];   // x, y, ds
  for (int i = 0; i < 3; ++i) { xb[i] = q; q += count * xe; }
  float* mid[2];
  for (int i = 0; i < 2; ++i) { mid[i] = q; q += count * me; }
  const ConvW& stem = m->convs[0];
  ConvF32Args a{};
  a.in = x_in; a.in_sn = (int64_t)3 * S * S; a.in_sc = (int64_t)S * S; a.in_sy = S; a.in_sx = 1;
  a.Hi = S; a.Wi = S; a.Cin = 3; a.Ho = Hc; a.Wo = Hc; a.Cout = 64; a.k = 7; a.stride = 2; a.pad = 3;
  a.w = stem.d_w32; a.bias = stem.d_b; a.residual = nullptr; a.out = c1;
  a.M = count * Hc * Hc; a.relu = 1;
  int rc = launch_conv_fp32(a, st);
  if (rc != CS_OK) return rc;
  rc = launch_maxpool_fp32(c1, xb[0], count, Hc, Hc, 64, st);
  if (rc != CS_OK) return rc;
  m->last_launches += 2;
  auto conv = [&](const ConvW& w, const float* in, int H, int W, int Cin, int Ho, int Wo, const float* res,
                  float* out, int relu) {
    ConvF32Args c{};
    c.in = in; c.in_sn = (int64_t)H * W * Cin; c.in_sc = 1; c.in_sy = (int64_t)W * Cin; c.in_sx = Cin;
    c.Hi = H; c.Wi = W; c.Cin = Cin; c.Ho = Ho; c.Wo = Wo; c.Cout = w.cout; c.k = w.k; c.stride = w.stride;
    c.pad = w.pad; c.w = w.d_w32; c.bias = w.d_b; c.residual = res; c.out = out; c.M = count * Ho * Wo;
    c.relu = relu;
    m->last_launches++;
    return launch_conv_fp32(c, st);
  };
  int xi = 0, H = Hp, W = Hp, C = 64;
  size_t bi = 0;
  int layer = 0;
  for (const BlockDesc& b : m->blocks) {
    if (maps_out && bi > 0 && b.stride == 2) {   // a stride-2 block opens the next layer: x is x_{layer+1}
      if (maps_out[layer])
        CS_CUDA(cudaMemcpyAsync(maps_out[layer], xb[xi], sizeof(float) * count * H * W * C,
                                cudaMemcpyDeviceToDevice, st));
      ++layer;
    }
    ++bi;
    const int Ho = (H + 2 - 3) / b.stride + 1, Wo = (W + 2 - 3) / b.stride + 1;
    float* x = xb[xi];
    float* y = xb[1 - xi];
    const float* res = x;
    if (b.ds >= 0) {
      if ((rc = conv(m->convs[b.ds], x, H, W, C, Ho, Wo, nullptr, xb[2], 0)) != CS_OK) return rc;
      res = xb[2];
    }
    if (!b.bottleneck) {
      if ((rc = conv(m->convs[b.conv1], x, H, W, C, Ho, Wo, nullptr, mid[0], 1)) != CS_OK) return rc;
      if ((rc = conv(m->convs[b.conv2], mid[0], Ho, Wo, b.cout, Ho, Wo, res, y, 1)) != CS_OK) return rc;
    } else {
      if ((rc = conv(m->convs[b.conv1], x, H, W, C, H, W, nullptr, mid[0], 1)) != CS_OK) return rc;
      if ((rc = conv(m->convs[b.conv2], mid[0], H, W, b.width, Ho, Wo, nullptr, mid[1], 1)) != CS_OK) return rc;
      if ((rc = conv(m->convs[b.conv3], mid[1], Ho, Wo, b.width, Ho, Wo, res, y, 1)) != CS_OK) return rc;
    }
    xi = 1 - xi; H = Ho; W = Wo; C = b.cout;
  }
  if (maps_out && maps_out[3])
    CS_CUDA(cudaMemcpyAsync(maps_out[3], xb[xi], sizeof(float) * count * H * W * C,
                            cudaMemcpyDeviceToDevice, st));
  rc = launch_head_fp32(xb[xi], count, H * W, C, m->d_fc_w, m->d_fc_b, prob_out, logits_out,
                        feat_out, st);
  if (rc != CS_OK) return rc;
  m->last_launches++;
  return CS_OK;
}

int check_forward_args(const char* fn, const cs_model* m, int tile, int precision, void* ws,
                       int64_t ws_bytes, int64_t max_batch) {
  CS_REQUIRE(m != nullptr, "%s: model is NULL", fn);
  CS_REQUIRE(precision == CS_PREC_FP32 || precision == CS_PREC_BF16, "%s: bad precision %d", fn, precision);
  CS_REQUIRE(tile == 16 || tile == 32 || (precision == CS_PREC_FP32 && tile >= 16 && tile <= 2048),
             "%s: tile %d unsupported (bf16: 16 or 32; fp32: 16 .. 2048)", fn, tile);
  CS_REQUIRE(ws != nullptr && max_batch > 0, "%s: workspace is NULL or max_batch <= 0", fn);
  int64_t need = cs_model_workspace_bytes(m, tile, max_batch, precision);
  if (ws_bytes < need) {
    set_error("%s: workspace %lld B < required %lld B", fn, (long long)ws_bytes, (long long)need);
    return CS_ERR_WORKSPACE;
  }
  return CS_OK;
}

int ensure_plan(cs_model* m, int tile, int64_t max_batch, void* ws, int64_t ws_bytes) {
  if (m->plan && m->plan->tile == tile && m->plan->max_batch == max_batch && m->plan->ws == ws)
    return CS_OK;
  m->plan.reset();
  return build_tc_plan(m, tile, max_batch, ws, ws_bytes);
}

}  // namespace
}  // namespace cs

extern "C" {

int cs_model_create(int arch, int n_convs, const float* const* conv_w_host,
                    const float* const* conv_b_host, const float* fc_w_host,
                    const float* fc_b_host, cs_model** out) {
  CS_REQUIRE(out != nullptr, "cs_model_create: out is NULL");
  *out = nullptr;
  CS_REQUIRE(conv_w_host && conv_b_host && fc_w_host && fc_b_host, "cs_model_create: NULL weights");
  std::vector<int> layers;
  bool bottleneck = false;
  int groups = 1, width_per_group = 64;
  switch (arch) {
    case CS_ARCH_RESNET18: layers = {2, 2, 2, 2}; break;
    case CS_ARCH_RESNET34: layers = {3, 4, 6, 3}; break;
    case CS_ARCH_RESNET50: layers = {3, 4, 6, 3}; bottleneck = true; break;
    case CS_ARCH_RESNEXT50_32X4D: layers = {3, 4, 6, 3}; bottleneck = true; groups = 32; width_per_group = 4; break;
    case CS_ARCH_RESNEXT101_32X8D: layers = {3, 4, 23, 3}; bottleneck = true; groups = 32; width_per_group = 8; break;
    default: set_error("cs_model_create: unknown arch %d", arch); return CS_ERR_UNSUPPORTED;
  }
  int rc = cs_check_device();
  if (rc != CS_OK) return rc;

  auto m = std::make_unique<cs_model>();
  m->arch = arch;
  auto add_conv = [&](int cin, int cout, int k, int stride, int pad, int grp) {
    ConvW c; c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.pad = pad; c.groups = grp;
    m->convs.push_back(std::move(c));
    return (int)m->convs.size() - 1;
  };
  add_conv(3, 64, 7, 2, 3, 1);
  int inplanes = 64;
  const int planes[4] = {64, 128, 256, 512};
  const int expansion = bottleneck ? 4 : 1;
  for (int L = 0; L < 4; ++L)
    for (int bi = 0; bi < layers[L]; ++bi) {
      int stride = (bi == 0 && L > 0) ? 2 : 1;
      BlockDesc b;
      b.bottleneck = bottleneck;
      b.cin = inplanes; b.cout = planes[L] * expansion; b.stride = stride;
      b.conv3 = -1;
      if (!bottleneck) {
        b.width = planes[L];
        b.conv1 = add_conv(inplanes, planes[L], 3, stride, 1, 1);
        b.conv2 = add_conv(planes[L], planes[L], 3, 1, 1, 1);
      } else {
        // model/resnext.py:81: width = int(planes * (base_width / 64.)) * groups; stride on conv2
        b.width = (planes[L] * width_per_group / 64) * groups;
        b.conv1 = add_conv(inplanes, b.width, 1, 1, 0, 1);
        b.conv2 = add_conv(b.width, b.width, 3, stride, 1, groups);
        b.conv3 = add_conv(b.width, b.cout, 1, 1, 0, 1);
      }
      b.ds = (stride != 1 || inplanes != b.cout) ? add_conv(inplanes, b.cout, 1, stride, 0, 1) : -1;
      inplanes = b.cout;
      m->blocks.push_back(b);
    }
  CS_REQUIRE((int)m->convs.size() == n_convs, "cs_model_create: arch %d has %d convs, got %d", arch,
             (int)m->convs.size(), n_convs);
  for (int i = 0; i < n_convs; ++i) {
    ConvW& c = m->convs[i];
    CS_REQUIRE(conv_w_host[i] && conv_b_host[i], "cs_model_create: conv %d weights NULL", i);
    const int kk = c.k * c.k;
    const int cin_g = c.cin / c.groups, cout_g = c.cout / c.groups;
    // torch layout [cout][cin/groups][k][k] -> dense [cout][cin][k][k] (zeros outside the group)
    c.w.assign((size_t)c.cout * c.cin * kk, 0.f);
    for (int co = 0; co < c.cout; ++co) {
      const int grp = co / cout_g;
      for (int cl = 0; cl < cin_g; ++cl)
        memcpy(&c.w[((size_t)co * c.cin + grp * cin_g + cl) * kk],
               &conv_w_host[i][((size_t)co * cin_g + cl) * kk], kk * sizeof(float));
    }
    c.b.assign(conv_b_host[i], conv_b_host[i] + c.cout);
    // fp32 device layout [(dy*k+dx)*cin + ci][cout]
    const size_t nw = c.w.size();
    std::vector<float> t(nw);
    for (int co = 0; co < c.cout; ++co)
      for (int ci = 0; ci < c.cin; ++ci)
        for (int tp = 0; tp < kk; ++tp)
          t[((size_t)tp * c.cin + ci) * c.cout + co] = c.w[((size_t)co * c.cin + ci) * kk + tp];
    CS_CUDA(cudaMalloc(&c.d_w32, nw * sizeof(float)));
    CS_CUDA(cudaMemcpy(c.d_w32, t.data(), nw * sizeof(float), cudaMemcpyHostToDevice));
    CS_CUDA(cudaMalloc(&c.d_b, c.cout * sizeof(float)));
    CS_CUDA(cudaMemcpy(c.d_b, c.b.data(), c.cout * sizeof(float), cudaMemcpyHostToDevice));
  }
  m->feat_dim = 512 * expansion;
  CS_CUDA(cudaMalloc(&m->d_fc_w, 2 * m->feat_dim * sizeof(float)));
  CS_CUDA(cudaMemcpy(m->d_fc_w, fc_w_host, 2 * m->feat_dim * sizeof(float), cudaMemcpyHostToDevice));
  CS_CUDA(cudaMalloc(&m->d_fc_b, 2 * sizeof(float)));
  CS_CUDA(cudaMemcpy(m->d_fc_b, fc_b_host, 2 * sizeof(float), cudaMemcpyHostToDevice));
  *out = m.release();
  return CS_OK;
}

int cs_model_destroy(cs_model* m) {
  if (!m) return CS_OK;
  m->plan.reset();
  for (ConvW& c : m->convs) {
    if (c.d_w32) cudaFree(c.d_w32);
    if (c.d_b) cudaFree(c.d_b);
  }
  if (m->d_fc_w) cudaFree(m->d_fc_w);
  if (m->d_fc_b) cudaFree(m->d_fc_b);
  delete m;
  return CS_OK;
}

int cs_model_feature_dim(const cs_model* m) { return m ? m->feat_dim : 0; }

int cs_model_set_fc(cs_model* m, const float* fc_w_host, const float* fc_b_host) {
  CS_REQUIRE(m && fc_w_host && fc_b_host, "cs_model_set_fc: NULL pointer");
  CS_CUDA(cudaMemcpy(m->d_fc_w, fc_w_host, 2 * m->feat_dim * sizeof(float), cudaMemcpyHostToDevice));
  CS_CUDA(cudaMemcpy(m->d_fc_b, fc_b_host, 2 * sizeof(float), cudaMemcpyHostToDevice));
  return CS_OK;
}

int64_t cs_model_workspace_bytes(const cs_model* m, int tile, int64_t max_batch, int precision) {
  if (!m || max_batch <= 0) return 0;
  if (precision != CS_PREC_FP32 && tile != 16 && tile != 32) return 0;
  if (tile < 16 || tile > 2048) return 0;
  if (precision == CS_PREC_FP32) {
    int64_t chunk = max_batch < kFp32Chunk ? max_batch : kFp32Chunk;
    return chunk * fp32_floats_per_inst(m, tile) * 4 + 4096;
  }
  int64_t xe, me;
  act_sizes(m, tile, &xe, &me);
  int64_t b_pad = round_up(max_batch, kGemmBM);
  return g_xbufs * round_up(b_pad * xe * 2, 1024) + 2 * round_up(b_pad * me * 2, 1024) + 4096;
}

int cs_model_forward_tiles(cs_model* m, const uint8_t* img, int n_bags, int H, int W, int tile,
                           int interval, int64_t inst_begin, int64_t inst_count, int precision,
                           float* prob_out, float* feat_out, void* workspace,
                           int64_t workspace_bytes, int64_t max_batch, void* stream) {
  int rc = check_forward_args("cs_model_forward_tiles", m, tile, precision, workspace,
                              workspace_bytes, max_batch);
  if (rc != CS_OK) return rc;
  CS_REQUIRE(img && prob_out, "cs_model_forward_tiles: NULL pointer");
  int gh = grid_count(H, tile, interval), gw = grid_count(W, tile, interval);
  CS_REQUIRE(gh > 0 && gw > 0, "cs_model_forward_tiles: bad geometry H=%d W=%d tile=%d interval=%d",
             H, W, tile, interval);
  const int64_t T = (int64_t)gh * gw;
  CS_REQUIRE(inst_begin >= 0 && inst_count >= 0 && inst_begin + inst_count <= (int64_t)n_bags * T,
             "cs_model_forward_tiles: instance range [%lld,+%lld) outside %d bags x %lld tiles",
             (long long)inst_begin, (long long)inst_count, n_bags, (long long)T);
  cudaStream_t st = as_stream(stream);
  m->last_launches = 0;
  if (precision == CS_PREC_BF16) {
    rc = ensure_plan(m, tile, max_batch, workspace, workspace_bytes);
    if (rc != CS_OK) return rc;
    StemArgs sa{};
    sa.img = img; sa.H = H; sa.W = W; sa.tile = tile; sa.interval = interval; sa.grid_w = gw;
    sa.tiles_per_bag = T; sa.x = nullptr;
    for (int64_t done = 0; done < inst_count; done += max_batch) {
      int64_t cnt = inst_count - done < max_batch ? inst_count - done : max_batch;
      sa.inst_begin = inst_begin + done;
      rc = run_tc_batch(m, sa, cnt, prob_out + done, nullptr,
                        feat_out ? feat_out + done * m->feat_dim : nullptr, st);
      if (rc != CS_OK) return rc;
    }
  } else {
    int64_t chunk = max_batch < kFp32Chunk ? max_batch : kFp32Chunk;
    float* x_in = reinterpret_cast<float*>(round_up((int64_t)(uintptr_t)workspace, 256));
    float* rest = x_in + chunk * 3 * tile * tile;
    for (int64_t done = 0; done < inst_count; done += chunk) {
      int64_t cnt = inst_count - done < chunk ? inst_count - done : chunk;
      rc = cs_unfold_normalize(img, n_bags, H, W, tile, interval, inst_begin + done, cnt, x_in, stream);
      if (rc != CS_OK) return rc;
      m->last_launches++;
      rc = run_fp32_batch(m, x_in, tile, cnt, rest, prob_out + done, nullptr,
                          feat_out ? feat_out + done * m->feat_dim : nullptr, st);
      if (rc != CS_OK) return rc;
    }
  }
  return CS_OK;
}

int cs_model_forward_tensor(cs_model* m, const float* x, int64_t n, int tile, int precision,
                            float* logits_out, float* feat_out, void* workspace,
                            int64_t workspace_bytes, int64_t max_batch, void* stream) {
  int rc = check_forward_args("cs_model_forward_tensor", m, tile, precision, workspace,
                              workspace_bytes, max_batch);
  if (rc != CS_OK) return rc;
  CS_REQUIRE(x && n >= 0, "cs_model_forward_tensor: NULL input or n < 0");
  CS_REQUIRE(logits_out || feat_out, "cs_model_forward_tensor: no output requested");
  cudaStream_t st = as_stream(stream);
  m->last_launches = 0;
  const int64_t per = (int64_t)3 * tile * tile;
  if (precision == CS_PREC_BF16) {
    rc = ensure_plan(m, tile, max_batch, workspace, workspace_bytes);
    if (rc != CS_OK) return rc;
    StemArgs sa{};
    sa.tile = tile;
    for (int64_t done = 0; done < n; done += max_batch) {
      int64_t cnt = n - done < max_batch ? n - done : max_batch;
      sa.x = x + done * per;
      rc = run_tc_batch(m, sa, cnt, nullptr, logits_out ? logits_out + done * 2 : nullptr,
                        feat_out ? feat_out + done * m->feat_dim : nullptr, st);
      if (rc != CS_OK) return rc;
    }
  } else {
    int64_t chunk = max_batch < kFp32Chunk ? max_batch : kFp32Chunk;
    float* x_unused = reinterpret_cast<float*>(round_up((int64_t)(uintptr_t)workspace, 256));
    float* rest = x_unused + chunk * 3 * tile * tile;
    for (int64_t done = 0; done < n; done += chunk) {
      int64_t cnt = n - done < chunk ? n - done : chunk;
      rc = run_fp32_batch(m, x + done * per, tile, cnt, rest, nullptr,
                          logits_out ? logits_out + done * 2 : nullptr,
                          feat_out ? feat_out + done * m->feat_dim : nullptr, st);
      if (rc != CS_OK) return rc;
    }
  }
  return CS_OK;
}

// N4: the encoder on whole images (Stage-1 "image" mode, model/resnet.py:271-278, and the
// encoder half of "segment" mode, :259-260).  fp32 CUDA-core path, any square size.
int cs_model_forward_image(cs_model* m, const float* x, int64_t n, int size, float* feat_out,
                           float* x1_out, float* x2_out, float* x3_out, float* x4_out,
                           void* workspace, int64_t workspace_bytes, int64_t max_batch, void* stream) {
  int rc = check_forward_args("cs_model_forward_image", m, size, CS_PREC_FP32, workspace, workspace_bytes,
                              max_batch);
  if (rc != CS_OK) return rc;
  CS_REQUIRE(x && n >= 0, "cs_model_forward_image: NULL input or n < 0");
  CS_REQUIRE(feat_out || x1_out || x2_out || x3_out || x4_out, "cs_model_forward_image: no output requested");
  cudaStream_t st = as_stream(stream);
  m->last_launches = 0;
  const int64_t per = (int64_t)3 * size * size;
  const int64_t chunk = max_batch < kFp32Chunk ? max_batch : kFp32Chunk;
  float* x_unused = reinterpret_cast<float*>(round_up((int64_t)(uintptr_t)workspace, 256));
  float* rest = x_unused + chunk * per;
  // per-instance element counts of x1..x4 (layer outputs)
  int64_t map_elems[4];
  {
    int H = stem_pool_size(size), li = -1;
    size_t bi = 0;
    for (const BlockDesc& b : m->blocks) {
      if (bi == 0 || b.stride == 2) ++li;
      H = (H + 2 - 3) / b.stride + 1;
      map_elems[li] = (int64_t)H * H * b.cout;
      ++bi;
    }
  }
  float* outs[4] = {x1_out, x2_out, x3_out, x4_out};
  for (int64_t done = 0; done < n; done += chunk) {
    const int64_t cnt = n - done < chunk ? n - done : chunk;
    float* maps[4];
    for (int i = 0; i < 4; ++i) maps[i] = outs[i] ? outs[i] + done * map_elems[i] : nullptr;
    rc = run_fp32_batch(m, x + done * per, size, cnt, rest, nullptr, nullptr,
                        feat_out ? feat_out + done * m->feat_dim : nullptr, st, maps);
    if (rc != CS_OK) return rc;
  }
  return CS_OK;
}

// N4: one conv2d on NHWC fp32 maps (Stage-3 decoder layers upconv1..8 / seg_out_conv with their
// BatchNorm folded by the caller, model/resnet.py:194-199, 280-303): w_dev is [k*k*Cin][Cout]
// (row index (dy*k+dx)*Cin + ci), bias_dev [Cout]; Cout % 4 == 0.
int cs_conv2d_nhwc_f32(const float* in, int64_t n, int H, int W, int Cin, const float* w_dev,
                       const float* bias_dev, int Cout, int k, int stride, int pad, int relu,
                       float* out, void* stream) {
  CS_REQUIRE(in && w_dev && bias_dev && out, "cs_conv2d_nhwc_f32: NULL pointer");
  CS_REQUIRE(n >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && k > 0 && stride > 0 && pad >= 0,
             "cs_conv2d_nhwc_f32: bad geometry");
  if (n == 0) return CS_OK;
  ConvF32Args c{};
  c.in = in; c.in_sn = (int64_t)H * W * Cin; c.in_sc = 1; c.in_sy = (int64_t)W * Cin; c.in_sx = Cin;
  c.Hi = H; c.Wi = W; c.Cin = Cin;
  c.Ho = (H + 2 * pad - k) / stride + 1; c.Wo = (W + 2 * pad - k) / stride + 1;
  c.Cout = Cout; c.k = k; c.stride = stride; c.pad = pad;
  c.w = w_dev; c.bias = bias_dev; c.residual = nullptr; c.out = out;
  c.M = n * c.Ho * c.Wo; c.relu = relu;
  return launch_conv_fp32(c, as_stream(stream));
}

// F.interpolate(x, size, mode="bilinear", align_corners=True) on NHWC fp32 maps.
int cs_resize_bilinear_nhwc_f32(const float* in, int64_t n, int Hi, int Wi, int C, int Ho, int Wo,
                                float* out, void* stream) {
  CS_REQUIRE(in && out, "cs_resize_bilinear_nhwc_f32: NULL pointer");
  CS_REQUIRE(n >= 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0, "cs_resize_bilinear_nhwc_f32: bad geometry");
  return launch_bilinear_fp32(in, out, n, Hi, Wi, Ho, Wo, C, as_stream(stream));
}

int64_t cs_model_last_launch_count(const cs_model* m) { return m ? m->last_launches : 0; }

// ---------------------------------------------------------------------------
// Diagnostics (used by tests/): run the production planner + tcgen05 kernel on one
// convolution or one plain GEMM and return the raw fp32 result.
// ---------------------------------------------------------------------------

// out_f32[M][N] = A[M][K] . B[N][K]^T + bias[N];  A, B bf16 (device), K multiple of 64 and
// <= 2560, N multiple of bn, bn in {64,128,256}.
int cs_debug_gemm_bf16(const void* a_bf16, const void* b_bf16, int64_t M, int N, int K,
                       const float* bias, int bn, float* out_f32, void* stream) {
  CS_REQUIRE(a_bf16 && b_bf16 && bias && out_f32, "cs_debug_gemm_bf16: NULL pointer");
  CS_REQUIRE(K % 64 == 0 && K / 64 <= kMaxSteps && N % bn == 0 && M > 0,
             "cs_debug_gemm_bf16: bad shape M=%lld N=%d K=%d bn=%d", (long long)M, N, K, bn);
  int rc = cs_check_device();
  if (rc != CS_OK) return rc;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  rc = make_mat_map_2d(&p.a_map[0], a_bf16, K, M, K, kGemmBM);
  if (rc != CS_OK) return rc;
  p.cluster = g_cluster;
  rc = make_mat_map_2d(&p.b_map, b_bf16, K, N, K, bn / p.cluster);
  if (rc != CS_OK) return rc;
  p.n_variants = 1;
  p.ext_steps = nullptr;
  p.n_steps[0] = K / 64;
  for (int s = 0; s < K / 64; ++s) p.steps[0][s] = KStep::make(s * 64, s * 64, 0, 0, 0);
  p.a_mode = 0;
  p.units_per_mtile = kGemmBM;
  p.num_m_tiles = (int)ceil_div<int64_t>(M, kGemmBM);
  p.num_n_tiles = N / bn;
  p.n_total = N;
  p.m_valid = M;
  p.bias = bias;
  p.out_f32 = out_f32;
  return launch_conv_gemm(p, bn, as_stream(stream));
}

// One k x k conv (k = 3: pad 1; k = 1: pad 0; stride 1|2; `groups` groups with the torch weight
// layout [Cout][Cin/groups][k][k]) through plan_conv: in_hi bf16 [n][Hi*Wi][Cin] (device),
// w_host / bias_host fp32 (host) -> out_f32 [n][Ho*Wo][Cout] (device), no ReLU.  n % 128 == 0.
// reverse: walk the work items backwards (the snake order of odd layers).
int cs_debug_conv_bf16(const void* in_hi, int64_t n, int Hi, int Wi, int Cin, int Cout, int k,
                       int stride, int groups, const float* w_host, const float* bias_host,
                       int reverse, float* out_f32, void* stream) {
  CS_REQUIRE(in_hi && w_host && bias_host && out_f32, "cs_debug_conv_bf16: NULL pointer");
  CS_REQUIRE(n > 0 && n % kGemmBM == 0, "cs_debug_conv_bf16: n must be a positive multiple of 128");
  CS_REQUIRE((k == 1 || k == 3) && groups >= 1 && Cin % groups == 0 && Cout % groups == 0,
             "cs_debug_conv_bf16: bad k / groups");
  int rc = cs_check_device();
  if (rc != CS_OK) return rc;
  const int pad = k == 3 ? 1 : 0, kk = k * k;
  const int Ho = (Hi + 2 * pad - k) / stride + 1, Wo = (Wi + 2 * pad - k) / stride + 1;
  std::vector<float> dense((size_t)Cout * Cin * kk, 0.f);
  const int cin_g = Cin / groups, cout_g = Cout / groups;
  for (int co = 0; co < Cout; ++co)
    for (int cl = 0; cl < cin_g; ++cl)
      memcpy(&dense[((size_t)co * Cin + (co / cout_g) * cin_g + cl) * kk],
             &w_host[((size_t)co * cin_g + cl) * kk], kk * sizeof(float));
  ConvGeom g{Hi, Wi, Cin, Ho, Wo, Cout, k, stride, pad, groups};
  PlannedConv pc;
  rc = plan_conv(g, dense.data(), bias_host, nullptr, nullptr, nullptr,
                 reinterpret_cast<const __nv_bfloat16*>(in_hi), nullptr, n, &pc);
  if (rc != CS_OK) return rc;
  pc.p.relu = 0;
  rc = launch_planned(pc, n, out_f32, as_stream(stream), reverse != 0);
  cudaError_t e = cudaStreamSynchronize(as_stream(stream));
  free_planned(pc);
  if (rc != CS_OK) return rc;
  CS_CUDA(e);
  return CS_OK;
}

// The tile-32 tensor-core stem in isolation (default form, or stem_win.cu with CELLSEG_STEM=win):
// img u8 [n_bags][H][W][3] (device), folded conv1 weights [64][3][7][7] / bias [64] fp32 (host)
// -> out_bf16 [inst_count][8*8][64] (device) = maxpool3x3/2(relu(conv7x7/2(normalise(tile)) + bias)).
int cs_debug_stem_bf16(const uint8_t* img, int n_bags, int H, int W, int interval, int64_t inst_begin,
                       int64_t inst_count, const float* w_host, const float* bias_host, void* out_bf16,
                       void* stream) {
  CS_REQUIRE(img && w_host && bias_host && out_bf16, "cs_debug_stem_bf16: NULL pointer");
  CS_REQUIRE(n_bags > 0 && H >= 32 && W >= 32 && interval > 0 && inst_count > 0 && inst_begin >= 0,
             "cs_debug_stem_bf16: bad geometry");
  int rc = cs_check_device();
  if (rc != CS_OK) return rc;
  const int gh = cs_grid_count(H, 32, interval), gw = cs_grid_count(W, 32, interval);
  CS_REQUIRE(inst_begin + inst_count <= (int64_t)n_bags * gh * gw, "cs_debug_stem_bf16: instance range");
  const int nw = g_stem_win ? stem_win_weight_bytes() : stem_ts_weight_bytes();
  std::vector<uint16_t> packed(nw / 2);
  if (g_stem_win) pack_stem_weights_win(w_host, packed.data());
  else pack_stem_weights_ts(w_host, packed.data());
  uint16_t* d_w = nullptr;
  float* d_b = nullptr;
  CS_CUDA(cudaMalloc(&d_w, nw));
  CS_CUDA(cudaMalloc(&d_b, 64 * sizeof(float)));
  CS_CUDA(cudaMemcpy(d_w, packed.data(), nw, cudaMemcpyHostToDevice));
  CS_CUDA(cudaMemcpy(d_b, bias_host, 64 * sizeof(float), cudaMemcpyHostToDevice));
  StemArgs sa{};
  sa.img = img; sa.H = H; sa.W = W; sa.tile = 32; sa.interval = interval; sa.grid_w = gw;
  sa.tiles_per_bag = (int64_t)gh * gw;
  sa.inst_begin = inst_begin;
  sa.x = nullptr;
  sa.count = inst_count;
  sa.w = nullptr;
  sa.bias = d_b;
  sa.out_hi = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  sa.out_lo = nullptr;
  rc = g_stem_win ? launch_stem_win(sa, d_w, nullptr, as_stream(stream)) : launch_stem_ts(sa, d_w, as_stream(stream));
  cudaError_t e = cudaStreamSynchronize(as_stream(stream));
  cudaFree(d_w);
  cudaFree(d_b);
  if (rc != CS_OK) return rc;
  CS_CUDA(e);
  return CS_OK;
}

// One layer-1 BasicBlock (8x8 x 64 channels, no downsample) through the production planner:
// y = relu(conv2(relu(conv1(x) + b1)) + b2 + x), one fused launch (conv_block.cu) or, with
// CELLSEG_BLOCK_FUSE=0, two y-sum launches (*launches_out says which).  in_hi / out_bf16 bf16
// [n][64][64] (device), weights [64][64][3][3] / biases [64] fp32 (host).  n % 128 == 0.  Synchronises.
int cs_debug_basic_block_bf16(const void* in_hi, int64_t n, const float* w1_host, const float* b1_host,
                              const float* w2_host, const float* b2_host, int reverse, void* out_bf16,
                              int* launches_out, void* stream) {
  CS_REQUIRE(in_hi && w1_host && b1_host && w2_host && b2_host && out_bf16, "cs_debug_basic_block_bf16: NULL pointer");
  CS_REQUIRE(n > 0 && n % kGemmBM == 0, "cs_debug_basic_block_bf16: n must be a positive multiple of 128");
  int rc = cs_check_device();
  if (rc != CS_OK) return rc;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(in_hi);
  __nv_bfloat16* mid = nullptr;
  CS_CUDA(cudaMalloc(&mid, (size_t)n * 64 * 64 * sizeof(__nv_bfloat16)));
  ConvGeom g{8, 8, 64, 8, 8, 64, 3, 1, 1, 1};
  std::vector<PlannedConv> layers(2);
  rc = plan_conv(g, w1_host, b1_host, nullptr, nullptr, nullptr, x, nullptr, n, &layers[0]);
  if (rc == CS_OK) {
    layers[0].p.res_hi = nullptr; layers[0].p.res_lo = nullptr; layers[0].p.out_hi = mid;
    layers[0].p.out_lo = nullptr; layers[0].p.relu = 1;
    rc = finalize_io_maps(layers[0], n);
  }
  if (rc == CS_OK) rc = plan_conv(g, w2_host, b2_host, nullptr, nullptr, nullptr, mid, nullptr, n, &layers[1]);
  if (rc == CS_OK) {
    layers[1].p.res_hi = x; layers[1].p.res_lo = nullptr;
    layers[1].p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    layers[1].p.out_lo = nullptr; layers[1].p.relu = 1;
    rc = finalize_io_maps(layers[1], n);
  }
  if (rc == CS_OK) {
    fuse_basic_block(layers, x);
    int launches = 0;
    for (PlannedConv& pc : layers) {
      if (pc.skip) continue;
      rc = launch_planned(pc, n, nullptr, as_stream(stream), (reverse != 0) ^ (launches & 1));
      ++launches;
      if (rc != CS_OK) break;
    }
    if (launches_out) *launches_out = launches;
  }
  cudaError_t e = cudaStreamSynchronize(as_stream(stream));
  free_planned(layers[0]);
  free_planned(layers[1]);
  cudaFree(mid);
  if (rc != CS_OK) return rc;
  CS_CUDA(e);
  return CS_OK;
}

}  // extern "C"
