// K2 (bf16 mode): implicit-GEMM convolutions on tcgen05 tensor cores.
//
// Reference: BasicBlock.forward (model/resnet.py:28-43) and resnet_forward (:234-248)
// under eval-mode BN folded into the conv (see model.cu for the folding/packing).
//
// Every 3x3 / 1x1 convolution of the encoder is one launch of conv_gemm_kernel:
//   out[r][n] = act( bias[n] + sum_steps A_step[r][0:64] . B[n][b_k : b_k+64]  (+ residual) )
// A is never materialised as im2col.  A K step is a TMA box:
//   4-D mode  rows = (instance, oy, ox): box {64 ch, W, H, instances} fetched at the
//             tap's pixel shift (dx, dy); out-of-image pixels are zero-filled by TMA,
//             which is exactly the conv's zero padding.  Stride-2 convs read one of
//             four parity-phase maps of the input.
//   2-D mode  rows = instances, columns = (pixel, channel): small maps (<= 2x2
//             outputs) become dense GEMMs whose all-zero K blocks were dropped on the
//             host, so taps that only ever see padding cost nothing.
// Operands land in shared memory in the 128-byte-swizzled K-major layout UMMA
// expects; accumulators live in TMEM (two stages, so the epilogue of tile i overlaps
// the MMAs of tile i+1); the epilogue adds the folded-BN bias and the residual,
// applies ReLU and writes bf16 `hi` (next layer's operand) plus bf16 `lo`
// (= value - hi) so the residual stream keeps ~16 mantissa bits.
//
// Warp roles (320 threads, one CTA per SM, persistent over (m_tile, n_tile)):
//   warp 0 : TMA producer (one lane)          warp 1 : TMEM alloc + MMA issuer (one lane)
//   warps 2-9 : epilogue; warp % 4 = TMEM lane quadrant, (warp-2)/4 = column half.  The
//   residual of the next 32-column chunk is fetched (256-bit loads) before the accumulator
//   wait / while the current chunk is converted and stored.
#include "fwd.cuh"
#include "gemm_epilogue.cuh"

namespace cs {
namespace {

// ----------------------------------------------------------------------------
// The GEMM kernel
// ----------------------------------------------------------------------------
template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr uint32_t kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr uint32_t kBBytes = BN * kGemmBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kBarOffset = kStages * kStageBytes;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256 + 1024;  // barriers + align slack
  static constexpr uint32_t kTmemCols = 2 * BN;                    // two accumulator stages
};

constexpr int kGemmThreads = 320;  // TMA warp, MMA warp, 8 epilogue warps

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bar_base = base + Cfg::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::kBarOffset + 8 * (2 * Cfg::kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    fence_barrier_init();
    prefetch_tmap(&p.b_map);
    prefetch_tmap(&p.a_map[0]);
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_work = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int wi = p.reverse ? num_work - 1 - w : w;
        const int m_tile = wi / p.num_n_tiles, n_tile = wi - m_tile * p.num_n_tiles;
        const int var = p.n_variants > 1 ? n_tile : 0;
        const int ns = p.n_steps[var];
        for (int s = 0; s < ns; ++s) {
          const KStep st = p.steps[var][s];
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          if (p.a_mode == 0)
            tma_load_2d(a_dst, &p.a_map[st.map], full_bar(stage), st.a_c0, m_tile * kGemmBM);
          else
            tma_load_4d(a_dst, &p.a_map[st.map], full_bar(stage), st.a_c0, st.dx, st.dy,
                        m_tile * p.units_per_mtile);
          tma_load_2d(b_dst, &p.b_map, full_bar(stage), st.b_k, n_tile * BN);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int wi = p.reverse ? num_work - 1 - w : w;
        const int m_tile = wi / p.num_n_tiles, n_tile = wi - m_tile * p.num_n_tiles;
        const int var = p.n_variants > 1 ? n_tile : 0;
        const int ns = p.n_steps[var];
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int s = 0; s < ns; ++s) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = umma_desc_sw128(a_src);
          const uint64_t b_desc = umma_desc_sw128(a_src + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // +32 B per K=16 slice inside the 128-byte swizzle row: +2 in (addr >> 4)
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                      (s > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem stage when these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // 8 epilogue warps: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kColsPerWarp = BN / 2;
    const int row_in_tile = quad * 32 + lane;
    const EpiArgs ea{p.bias, p.res_hi, p.res_lo, p.out_hi, p.out_lo, p.out_f32, p.relu};
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int wi = p.reverse ? num_work - 1 - w : w;
      const int m_tile = wi / p.num_n_tiles, n_tile = wi - m_tile * p.num_n_tiles;
      const int64_t row = (int64_t)m_tile * kGemmBM + row_in_tile;
      const int col0 = n_tile * BN + half * kColsPerWarp;
      epilogue_warp<kColsPerWarp / 32>(
          ea, row < p.m_valid, row * (int64_t)p.n_total + col0, col0,
          tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * kColsPerWarp),
          tfull_bar(acc), acc_phase);
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------
// Stem (v1, CUDA cores): unfold + normalise + conv7x7/2 + bias + ReLU + maxpool3x3/2.
// Reference: dataset/dataset.py:409-416 (crop, ToTensor, Normalize) and
// model/resnet.py:236-239.  One CTA walks tiles; thread = 4 output pixels x 16 channels.
// ----------------------------------------------------------------------------
constexpr int kStemK = 147;

template <int S>
struct StemGeom {
  static constexpr int kIn = S + 6;        // zero-padded input side (pad 3)
  static constexpr int kC = S / 2;         // conv output side
  static constexpr int kP = S / 4;         // pooled output side
  static constexpr int kInElems = kIn * kIn * 3;
  static constexpr int kConvPix = kC * kC;
  static constexpr size_t kSmem =
      (size_t)(kStemK * 64 + kInElems + kConvPix * 64 + 3 * 256) * sizeof(float);
};

__constant__ float c_stem_lut[3 * 256];

template <int S>
__global__ void __launch_bounds__(256, 1)
stem_bf16_kernel(StemArgs a) {
  using G = StemGeom<S>;
  extern __shared__ float sm[];
  float* w_s = sm;                         // [147][64]
  float* in_s = w_s + kStemK * 64;         // [kIn][kIn][3]
  float* conv_s = in_s + G::kInElems;      // [kC*kC][64]
  float* lut_s = conv_s + G::kConvPix * 64;
  const int tid = threadIdx.x;
  for (int i = tid; i < kStemK * 64; i += 256) w_s[i] = a.w[i];
  for (int i = tid; i < 768; i += 256) lut_s[i] = c_stem_lut[i];
  for (int i = tid; i < G::kInElems; i += 256) in_s[i] = 0.f;  // border stays zero
  __syncthreads();

  // thread -> (pixel group, channel group)
  constexpr int kGroupsPerRow = G::kC / 4;
  constexpr int kPixGroups = G::kConvPix / 4;  // 64 (S=32) or 16 (S=16)
  const int cg = tid & 3;
  const int pg = tid >> 2;  // 0..63

  for (int64_t t = blockIdx.x; t < a.count; t += gridDim.x) {
    // ---- load + normalise the S x S x 3 tile into the padded buffer
    if (a.x == nullptr) {
      int64_t inst = a.inst_begin + t;
      int64_t bag = inst / a.tiles_per_bag;
      int tl = (int)(inst - bag * a.tiles_per_bag);
      int gy = tl / a.grid_w, gx = tl - gy * a.grid_w;
      int row0 = grid_coord(gy, a.H, S, a.interval), col0 = grid_coord(gx, a.W, S, a.interval);
      const uint8_t* src = a.img + ((bag * a.H + row0) * (int64_t)a.W + col0) * 3;
      for (int e = tid; e < S * S * 3; e += 256) {
        int y = e / (S * 3), r = e - y * (S * 3);  // r = x*3 + c
        int c = r % 3;
        uint8_t u = src[(int64_t)y * a.W * 3 + r];
        in_s[((y + 3) * G::kIn + 3) * 3 + r] = lut_s[c * 256 + u];
      }
    } else {
      const float* src = a.x + t * (int64_t)(3 * S * S);  // NCHW
      for (int e = tid; e < S * S * 3; e += 256) {
        int c = e / (S * S), r = e - c * (S * S);
        int y = r / S, x = r - y * S;
        in_s[((y + 3) * G::kIn + (x + 3)) * 3 + c] = src[e];
      }
    }
    __syncthreads();

    // ---- conv 7x7 stride 2 (+bias, ReLU) into conv_s
    if (pg < kPixGroups) {
      const int oy = pg / kGroupsPerRow, ox0 = (pg - oy * kGroupsPerRow) * 4;
      float acc[4][16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
      for (int dy = 0; dy < 7; ++dy) {
        const float* in_row = in_s + ((2 * oy + dy) * G::kIn + 2 * ox0) * 3;
        for (int dxc = 0; dxc < 21; ++dxc) {  // dxc = dx*3 + c
          const float4* wr =
              reinterpret_cast<const float4*>(w_s + (dy * 21 + dxc) * 64 + cg * 16);
          float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
          const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float xv = in_row[i * 6 + dxc];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* dst = conv_s + ((oy * G::kC + ox0 + i) * 64 + cg * 16);
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = fmaxf(acc[i][j] + a.bias[cg * 16 + j], 0.f);
      }
    }
    __syncthreads();

    // ---- maxpool 3x3 stride 2 pad 1 -> hi/lo bf16, [pixel][64]
    for (int o = tid; o < G::kP * G::kP * 4; o += 256) {
      const int cq = o & 3, pp = o >> 2;
      const int py = pp / G::kP, px = pp - py * G::kP;
      float m[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) m[j] = 0.f;  // inputs are >= 0 after ReLU
      for (int dy = 0; dy < 3; ++dy) {
        int iy = 2 * py - 1 + dy;
        if (iy < 0 || iy >= G::kC) continue;
        for (int dx = 0; dx < 3; ++dx) {
          int ix = 2 * px - 1 + dx;
          if (ix < 0 || ix >= G::kC) continue;
          const float* srcp = conv_s + ((iy * G::kC + ix) * 64 + cq * 16);
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = fmaxf(m[j], srcp[j]);
        }
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = pack_bf16x2(m[2 * j], m[2 * j + 1]);
        lo[j] = pack_bf16x2(m[2 * j] - bf16_lo_f(hi[j]), m[2 * j + 1] - bf16_hi_f(hi[j]));
      }
      int64_t off = (t * (G::kP * G::kP) + pp) * 64 + cq * 16;
      uint4* oh = reinterpret_cast<uint4*>(a.out_hi + off);
      oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      oh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      if (a.out_lo) {
        uint4* ol = reinterpret_cast<uint4*>(a.out_lo + off);
        ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    __syncthreads();
  }
}

// Tile head on the hi/lo stream: avgpool+maxpool -> Linear(C,2) -> softmax[:,1].
// Reference: model/resnet.py:266-267, inference.py:24-27.  One warp per instance.
__global__ void __launch_bounds__(256)
head_bf16_kernel(const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                 int64_t n, int P, int C, const float* __restrict__ fc_w,
                 const float* __restrict__ fc_b, float* __restrict__ prob_out,
                 float* __restrict__ logits_out, float* __restrict__ feat_out) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const __nv_bfloat16* xh = x_hi + warp * (int64_t)P * C;
  const __nv_bfloat16* xl = x_lo ? x_lo + warp * (int64_t)P * C : nullptr;
  float z0 = 0.f, z1 = 0.f;
  for (int c = lane; c < C; c += 32) {
    float s = 0.f, mx = -INFINITY;
    for (int q = 0; q < P; ++q) {
      float v = __bfloat162float(xh[(int64_t)q * C + c]);
      if (xl) v += __bfloat162float(xl[(int64_t)q * C + c]);
      s += v;
      mx = fmaxf(mx, v);
    }
    float f = s / (float)P + mx;
    if (feat_out) feat_out[warp * C + c] = f;
    z0 = fmaf(f, fc_w[c], z0);
    z1 = fmaf(f, fc_w[C + c], z1);
  }
  for (int o = 16; o > 0; o >>= 1) {
    z0 += __shfl_xor_sync(0xffffffffu, z0, o);
    z1 += __shfl_xor_sync(0xffffffffu, z1, o);
  }
  if (lane == 0) {
    z0 += fc_b[0];
    z1 += fc_b[1];
    if (logits_out) { logits_out[warp * 2] = z0; logits_out[warp * 2 + 1] = z1; }
    if (prob_out) {
      float mx = fmaxf(z0, z1);
      float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
      prob_out[warp] = e1 / (e0 + e1);
    }
  }
}

// ----------------------------------------------------------------------------
// Host helpers
// ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encoder(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return CS_ERR_CUDA;
    }
    cached = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  *fn = cached;
  return CS_OK;
}

template <int BN>
int launch_gemm_bn(const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg::kSmemBytes));
    if (dev < 64) attr_done[dev] = true;
  }
  int work = p.num_m_tiles * p.num_n_tiles;
  int grid = work < kNumSMs ? work : kNumSMs;
  conv_gemm_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(p);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // namespace

int launch_conv_gemm(const GemmParams& p, int BN, cudaStream_t st) {
  if (p.num_m_tiles <= 0 || p.num_n_tiles <= 0) return CS_OK;
  switch (BN) {
    case 64: return launch_gemm_bn<64>(p, st);
    case 128: return launch_gemm_bn<128>(p, st);
    case 256: return launch_gemm_bn<256>(p, st);
    default:
      set_error("launch_conv_gemm: unsupported N tile %d", BN);
      return CS_ERR_UNSUPPORTED;
  }
}

int make_act_map_4d(CUtensorMap* map, const void* base, int C, int W, int H, int64_t T,
                    int64_t stride_x_elems, int64_t stride_y_elems, int64_t stride_t_elems,
                    int box_w, int box_h, int box_t) {
  EncodeTiledFn enc;
  int rc = get_encoder(&enc);
  if (rc != CS_OK) return rc;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
  cuuint64_t strides[3] = {(cuuint64_t)stride_x_elems * 2, (cuuint64_t)stride_y_elems * 2,
                           (cuuint64_t)stride_t_elems * 2};
  cuuint32_t box[4] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_w, (cuuint32_t)box_h,
                       (cuuint32_t)box_t};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d C=%d W=%d H=%d T=%lld box=%dx%dx%d) failed: %d", C, W, H,
              (long long)T, box_w, box_h, box_t, (int)r);
    return CS_ERR_CUDA;
  }
  return CS_OK;
}

int make_mat_map_2d(CUtensorMap* map, const void* base, int64_t K, int64_t rows,
                    int64_t row_pitch_elems, int box_rows) {
  EncodeTiledFn enc;
  int rc = get_encoder(&enc);
  if (rc != CS_OK) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d K=%lld rows=%lld pitch=%lld box_rows=%d) failed: %d",
              (long long)K, (long long)rows, (long long)row_pitch_elems, box_rows, (int)r);
    return CS_ERR_CUDA;
  }
  return CS_OK;
}

int launch_stem_bf16(const StemArgs& a, cudaStream_t st) {
  static bool lut_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !lut_done[dev]) {
    float lut[768];
    get_norm_lut_host(lut);
    CS_CUDA(cudaMemcpyToSymbol(c_stem_lut, lut, sizeof(lut)));
    CS_CUDA(cudaFuncSetAttribute(stem_bf16_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)StemGeom<32>::kSmem));
    CS_CUDA(cudaFuncSetAttribute(stem_bf16_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)StemGeom<16>::kSmem));
    if (dev < 64) lut_done[dev] = true;
  }
  if (a.count <= 0) return CS_OK;
  int grid = (int)(a.count < (int64_t)kNumSMs * 4 ? a.count : (int64_t)kNumSMs * 4);
  if (a.tile == 32) {
    grid = (int)(a.count < kNumSMs ? a.count : kNumSMs);
    stem_bf16_kernel<32><<<grid, 256, StemGeom<32>::kSmem, st>>>(a);
  } else if (a.tile == 16) {
    stem_bf16_kernel<16><<<grid, 256, StemGeom<16>::kSmem, st>>>(a);
  } else {
    set_error("bf16 stem: tile %d unsupported (16 or 32)", a.tile);
    return CS_ERR_UNSUPPORTED;
  }
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int launch_head_bf16(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int64_t n, int P, int C,
                     const float* fc_w, const float* fc_b, float* prob_out, float* logits_out,
                     float* feat_out, cudaStream_t st) {
  if (n <= 0) return CS_OK;
  int64_t blocks = ceil_div<int64_t>(n * 32, 256);
  head_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x_hi, x_lo, n, P, C, fc_w, fc_b, prob_out,
                                                    logits_out, feat_out);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // namespace cs
