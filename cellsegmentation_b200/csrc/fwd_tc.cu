// K2 (bf16 mode) support code: CUDA-core stem (tile 16), tile head, TMA descriptor builders.
// The tensor-core kernels live in stem_win.cu, conv_ysum.cu,
// conv_halo.cu and conv_gemm.cu.
#include "fwd.cuh"
#include "gemm_epilogue.cuh"

namespace cs {
namespace {

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
// Reference: model/resnet.py:266-267, inference.py:24-27.  One warp per instance; a lane owns
// groups of 8 consecutive channels (one 16-byte load per pixel, C % 256 == 0), so a 512-channel
// 1x1 map is two independent 16-byte loads per lane and the launch streams x4 at HBM speed.
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  v[0] = bf16_lo_f(u.x); v[1] = bf16_hi_f(u.x); v[2] = bf16_lo_f(u.y); v[3] = bf16_hi_f(u.y);
  v[4] = bf16_lo_f(u.z); v[5] = bf16_hi_f(u.z); v[6] = bf16_lo_f(u.w); v[7] = bf16_hi_f(u.w);
}

// P1 = true: the 1x1 final map of tile sizes 16 and 32 (avgpool = maxpool = the value itself, so the
// pooled feature is v + v exactly) with the lane's fc_tile columns held in registers across the
// instances it walks.  The generic form below it (any P, C) spent ~1 050 warp instructions per
// instance on IEEE divisions and 64-bit index arithmetic and was issue-bound at 0.8 TB/s
// (ncu r02: 73 % issue-active, 101 us per 75 776 instances).
template <bool P1>
__global__ void __launch_bounds__(256)
head_bf16_kernel(const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                 int64_t n, int P, int C, const float* __restrict__ fc_w,
                 const float* __restrict__ fc_b, float* __restrict__ prob_out,
                 float* __restrict__ logits_out, float* __restrict__ feat_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float b0 = fc_b[0], b1 = fc_b[1];
  if (P1) {
    // C == 512: two groups of 8 channels per lane
    float w0[2][8], w1[2][8];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int c = lane * 8 + 256 * g;
#pragma unroll
      for (int e = 0; e < 8; ++e) { w0[g][e] = fc_w[c + e]; w1[g][e] = fc_w[C + c + e]; }
    }
    for (int64_t inst = warp0; inst < n; inst += n_warps) {
      const uint4* xh = reinterpret_cast<const uint4*>(x_hi + inst * 512) + lane;
      uint4 q[2] = {__ldg(xh), __ldg(xh + 32)};
      float f[2][8];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float v[8];
        unpack8(q[g], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[g][e] = v[e];
      }
      if (x_lo) {
        const uint4* xl = reinterpret_cast<const uint4*>(x_lo + inst * 512) + lane;
        const uint4 l[2] = {__ldg(xl), __ldg(xl + 32)};
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float v[8];
          unpack8(l[g], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[g][e] += v[e];
        }
      }
      float z0 = 0.f, z1 = 0.f;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[g][e] = f[g][e] / 1.0f + f[g][e];          // s / P + max with P = 1 (the division is exact)
          z0 = fmaf(f[g][e], w0[g][e], z0);
          z1 = fmaf(f[g][e], w1[g][e], z1);
        }
        if (feat_out) {
          float4* fo = reinterpret_cast<float4*>(feat_out + inst * 512 + lane * 8 + 256 * g);
          fo[0] = make_float4(f[g][0], f[g][1], f[g][2], f[g][3]);
          fo[1] = make_float4(f[g][4], f[g][5], f[g][6], f[g][7]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        z0 += __shfl_xor_sync(0xffffffffu, z0, o);
        z1 += __shfl_xor_sync(0xffffffffu, z1, o);
      }
      if (lane == 0) {
        z0 += b0;
        z1 += b1;
        if (logits_out) { logits_out[inst * 2] = z0; logits_out[inst * 2 + 1] = z1; }
        if (prob_out) {
          const float mx = fmaxf(z0, z1);
          const float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
          prob_out[inst] = e1 / (e0 + e1);
        }
      }
    }
    return;
  }
  const float fp = (float)P;
  for (int64_t inst = warp0; inst < n; inst += n_warps) {
    const __nv_bfloat16* xh = x_hi + inst * (int64_t)P * C;
    const __nv_bfloat16* xl = x_lo ? x_lo + inst * (int64_t)P * C : nullptr;
    float z0 = 0.f, z1 = 0.f;
#pragma unroll 2
    for (int c = lane * 8; c < C; c += 256) {
      float s[8], mx[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] = 0.f; mx[e] = -INFINITY; }
      for (int q = 0; q < P; ++q) {
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(xh + (int64_t)q * C + c)), v);
        if (xl) {
          float w[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(xl + (int64_t)q * C + c)), w);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += w[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += v[e]; mx[e] = fmaxf(mx[e], v[e]); }
      }
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = s[e] / fp + mx[e];
      if (feat_out) {
        float4* fo = reinterpret_cast<float4*>(feat_out + inst * C + c);
        fo[0] = make_float4(f[0], f[1], f[2], f[3]);
        fo[1] = make_float4(f[4], f[5], f[6], f[7]);
      }
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(fc_w + c)), a1 = __ldg(reinterpret_cast<const float4*>(fc_w + c + 4));
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(fc_w + C + c)), c1 = __ldg(reinterpret_cast<const float4*>(fc_w + C + c + 4));
      z0 = fmaf(f[0], a0.x, z0); z0 = fmaf(f[1], a0.y, z0); z0 = fmaf(f[2], a0.z, z0); z0 = fmaf(f[3], a0.w, z0);
      z0 = fmaf(f[4], a1.x, z0); z0 = fmaf(f[5], a1.y, z0); z0 = fmaf(f[6], a1.z, z0); z0 = fmaf(f[7], a1.w, z0);
      z1 = fmaf(f[0], c0.x, z1); z1 = fmaf(f[1], c0.y, z1); z1 = fmaf(f[2], c0.z, z1); z1 = fmaf(f[3], c0.w, z1);
      z1 = fmaf(f[4], c1.x, z1); z1 = fmaf(f[5], c1.y, z1); z1 = fmaf(f[6], c1.z, z1); z1 = fmaf(f[7], c1.w, z1);
    }
    for (int o = 16; o > 0; o >>= 1) {
      z0 += __shfl_xor_sync(0xffffffffu, z0, o);
      z1 += __shfl_xor_sync(0xffffffffu, z1, o);
    }
    if (lane == 0) {
      z0 += b0;
      z1 += b1;
      if (logits_out) { logits_out[inst * 2] = z0; logits_out[inst * 2 + 1] = z1; }
      if (prob_out) {
        float mx = fmaxf(z0, z1);
        float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
        prob_out[inst] = e1 / (e0 + e1);
      }
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

}  // namespace

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
    CS_CUDA(cudaFuncSetAttribute(stem_bf16_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)StemGeom<16>::kSmem));
    if (dev < 64) lut_done[dev] = true;
  }
  if (a.count <= 0) return CS_OK;
  int grid = (int)(a.count < (int64_t)num_sms() * 4 ? a.count : (int64_t)num_sms() * 4);
  if (a.tile == 16) {   // tile 32 runs the window-form tensor-core stem (stem_win.cu)
    stem_bf16_kernel<16><<<grid, 256, StemGeom<16>::kSmem, st>>>(a);
  } else {
    set_error("CUDA-core bf16 stem: tile %d unsupported (16 only)", a.tile);
    return CS_ERR_UNSUPPORTED;
  }
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int launch_head_bf16(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int64_t n, int P, int C,
                     const float* fc_w, const float* fc_b, float* prob_out, float* logits_out,
                     float* feat_out, cudaStream_t st) {
  if (n <= 0) return CS_OK;
  CS_REQUIRE(C % 256 == 0, "launch_head_bf16: feature width %d must be a multiple of 256", C);
  int64_t blocks = ceil_div<int64_t>(n * 32, 256);
  const int64_t cap = (int64_t)num_sms() * 8;                 // 64 warps per SM, each walking instances
  if (blocks > cap) blocks = cap;
  if (P == 1 && C == 512)
    CS_CUDA(launch_pdl(head_bf16_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, st, 1, x_hi, x_lo, n, P, C,
                       fc_w, fc_b, prob_out, logits_out, feat_out));
  else
    CS_CUDA(launch_pdl(head_bf16_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, st, 1, x_hi, x_lo, n, P, C,
                       fc_w, fc_b, prob_out, logits_out, feat_out));
  return CS_OK;
}

}  // namespace cs
