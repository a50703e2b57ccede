// K1+K2a fused on tensor cores: unfold + normalise + conv7x7/2 + BN bias + ReLU + maxpool3x3/2
// for 32x32 tiles (the test_tile.py / train_tile.py geometry).
//
// Reference: crop/ToTensor/Normalize (dataset/dataset.py:409-416, 78-83) and
// conv1 -> bn1 -> relu -> maxpool (model/resnet.py:236-239), eval-mode BN folded.
//
// GEMM view per instance: D[256 conv pixels][64] = A[256][K] . W[64][K]^T with
// K = 7 rows (dy) x 24 (21 = 7 dx * 3 channels, padded with zero weights) + 24 zero = 192.
// For a fixed (pixel, dy) the 21 inputs are contiguous in the zero-padded, normalised
// bf16 copy of the tile, so four producer warps build the im2col rows straight into the
// 128-byte-swizzled K-major layout UMMA reads (one thread = one pixel row).  A 128-pixel
// half instance is one UMMA M tile; a ring of two such A tiles lets the producers run ahead
// of the single MMA-issuing thread, and four TMEM accumulators let the MMAs run ahead of the
// epilogue warps, which add the bias, apply ReLU, park the 16x16x64 fp32 map in shared
// memory, max-pool it 3x3/2 and write the bf16 hi/lo stream [instance][8*8][64].
//
// Warp roles (288 threads, one CTA per SM, persistent over instances):
//   warps 0-3 producers   warps 4-7 epilogue + pooling   warp 8 weights TMA + MMA issue
#include "fwd.cuh"
#include "tc_ptx.cuh"

namespace cs {
namespace {

constexpr int kS = 32;                 // tile side
constexpr int kPad = kS + 6;           // zero-padded side (pad 3)
constexpr int kInPitch = 288;          // bytes per padded input row (>= 38*3*2 = 228; 72 words: conflict-free)
constexpr int kInBytes = kPad * kInPitch;
constexpr int kK = 192;                // padded GEMM K
constexpr int kChunkA = 128 * 128;     // one [128 rows][64 k] bf16 K chunk
constexpr int kATile = 3 * kChunkA;    // 48 KB per 128-pixel M tile
constexpr int kChunkB = 64 * 128;
constexpr int kThreads = 288;

struct Smem {
  static constexpr uint32_t a = 0;                          // 2 x 48 KB ring
  static constexpr uint32_t b = 2 * kATile;                 // 24 KB weights
  static constexpr uint32_t conv = b + 3 * kChunkB;         // 64 KB fp32 [256 px][64]
  static constexpr uint32_t in = conv + 256 * 64 * 4;       // 2 x staged input
  static constexpr uint32_t lut = in + 2 * kInBytes;        // 768 bf16
  static constexpr uint32_t bias = lut + 768 * 2;           // 64 fp32
  static constexpr uint32_t bars = bias + 64 * 4;           // mbarriers
  static constexpr uint32_t total = bars + 128;
};

struct StemTcParams {
  CUtensorMap w_map;         // [64][192] bf16, box {64, 64}
  StemArgs a;
  const uint16_t* lut_bf16;  // [3][256]
};

__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
stem_tc_kernel(const __grid_constant__ StemTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t bars = base + Smem::bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };          // 2
  auto empty_bar = [&](int s) { return bars + 8u * (2 + s); };   // 2
  auto tfull_bar = [&](int a) { return bars + 8u * (4 + a); };   // 4
  auto tempty_bar = [&](int a) { return bars + 8u * (8 + a); };  // 4
  const uint32_t wfull_bar = bars + 8u * 12;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bp + Smem::bars + 8 * 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const StemArgs& a = p.a;

  // one-time shared-memory init: zero both staged inputs (border stays zero), the K tail of
  // both A tiles (16-byte units 21..23 of every row are never rewritten), LUT, bias
  for (int i = tid; i < 2 * kInBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(bp + Smem::in)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * 128 * 3; i += kThreads) {
    int tile = i / 384, r = (i / 3) % 128, u = 21 + i % 3;
    uint32_t off = Smem::a + tile * kATile + (u >> 3) * kChunkA + r * 128 + (((u & 7) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(bp + off) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < 768; i += kThreads)
    reinterpret_cast<uint16_t*>(bp + Smem::lut)[i] = p.lut_bf16[i];
  for (int i = tid; i < 64; i += kThreads) reinterpret_cast<float*>(bp + Smem::bias)[i] = a.bias[i];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
    for (int q = 0; q < 4; ++q) { mbar_init(tfull_bar(q), 1); mbar_init(tempty_bar(q), 128); }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32((const void*)tmem_slot), 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // the output buffers may still be read by the previous batch's kernels

  const int64_t n_inst = a.count;

  if (warp < 4) {
    // ================= producers: stage input, build im2col rows =================
    const uint16_t* lut = reinterpret_cast<const uint16_t*>(bp + Smem::lut);
    // Raw inputs of the NEXT instance are prefetched into registers (u8 bytes, or fp32 bits
    // for the tensor path) and only converted when they are staged, one iteration later, so
    // the global-load latency is hidden behind the im2col build of the current instance.
    uint32_t pre[24];
    const bool from_img = a.x == nullptr;
    const int sy = tid >> 2, sq = tid & 3;
    auto prefetch = [&](int64_t t) {
      if (from_img) {
        int64_t inst = a.inst_begin + t;
        int64_t bag = inst / a.tiles_per_bag;
        int tl = (int)(inst - bag * a.tiles_per_bag);
        int gy = tl / a.grid_w, gx = tl - gy * a.grid_w;
        int row0 = grid_coord(gy, a.H, kS, a.interval), col0 = grid_coord(gx, a.W, kS, a.interval);
        const uint8_t* src = a.img + ((bag * a.H + row0) * (int64_t)a.W + col0) * 3;
        // thread -> tile row sy and a 24-element quarter sq of that row (8 pixels x 3 channels)
        const uint8_t* rowp = src + (int64_t)sy * a.W * 3 + 24 * sq;
#pragma unroll
        for (int i = 0; i < 24; ++i) pre[i] = __ldg(rowp + i);
      } else {
        const float* src = a.x + t * (int64_t)(3 * kS * kS) + sy * kS + 8 * sq;  // NCHW fp32
#pragma unroll
        for (int i = 0; i < 24; ++i)   // element i = pixel i/3, channel i%3
          pre[i] = __float_as_uint(__ldg(src + (i % 3) * kS * kS + i / 3));
      }
    };
    auto staged_value = [&](int i) -> uint16_t {
      if (from_img) return lut[(i % 3) * 256 + pre[i]];
      __nv_bfloat16 h = __float2bfloat16_rn(__uint_as_float(pre[i]));
      return *reinterpret_cast<uint16_t*>(&h);
    };
    int it = 0;
    int stage = 0;
    uint32_t phase = 0;
    int64_t t = blockIdx.x;
    if (t < n_inst) prefetch(t);
    for (; t < n_inst; t += gridDim.x, ++it) {
      const int buf = it & 1;
      uint8_t* in_g = bp + Smem::in + buf * kInBytes;
      {
        // padded row sy + 3, elements 9 + 24*sq .. (3 pad pixels * 3 channels = 9)
        uint16_t* dstp = reinterpret_cast<uint16_t*>(in_g + (sy + 3) * kInPitch) + 9 + 24 * sq;
#pragma unroll
        for (int i = 0; i < 24; ++i) dstp[i] = staged_value(i);
      }
      named_bar(1, 128);
      if (t + gridDim.x < n_inst) prefetch(t + gridDim.x);
      const uint32_t in_s = base + Smem::in + buf * kInBytes;
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const int r = tid;                       // row of the M tile
        const int pix = m * 128 + r;
        const int oy = pix >> 4, ox = pix & 15;
        const uint32_t a_tile = base + Smem::a + stage * kATile + r * 128;
        const uint32_t src0 = in_s + (2 * oy) * kInPitch + 12 * ox;
        const int sw = r & 7;
#pragma unroll
        for (int dy = 0; dy < 7; ++dy) {
          const uint32_t s = src0 + dy * kInPitch;
          uint32_t w[12];
#pragma unroll
          for (int j = 0; j < 12; ++j) w[j] = lds32(s + 4 * j);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int u = 3 * dy + q;            // 16-byte unit of the flat 384-byte row
            const uint32_t dst = a_tile + (u >> 3) * kChunkA + (((u & 7) ^ sw) << 4);
            sts128(dst, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          }
        }
        fence_async_smem();
        mbar_arrive(full_bar(stage));
        if (++stage == 2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 8) {
    // ================= weights TMA + MMA issue =================
    if (lane == 0) {
      mbar_expect_tx(wfull_bar, 3 * kChunkB);
      for (int kc = 0; kc < 3; ++kc)
        tma_load_2d(base + Smem::b + kc * kChunkB, &p.w_map, wfull_bar, kc * 64, 0);
      mbar_wait(wfull_bar, 0);
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t t = blockIdx.x; t < n_inst; t += gridDim.x) {
        for (int m = 0; m < 2; ++m) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
#pragma unroll
          for (int kc = 0; kc < 3; ++kc) {
            const uint64_t ad = umma_desc_sw128(base + Smem::a + stage * kATile + kc * kChunkA);
            const uint64_t bd = umma_desc_sw128(base + Smem::b + kc * kChunkB);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc,
                        (kc > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(acc));
          if (++stage == 2) { stage = 0; phase ^= 1u; }
          if (++acc == 4) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ================= epilogue: bias + ReLU -> smem, maxpool -> hi/lo =================
    const int etid = tid - 128;               // 0..127
    const int quad = warp & 3;
    const float* bias_s = reinterpret_cast<const float*>(bp + Smem::bias);
    float* conv_s = reinterpret_cast<float*>(bp + Smem::conv);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < n_inst; t += gridDim.x) {
      for (int m = 0; m < 2; ++m) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const int pix = m * 128 + quad * 32 + lane;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 64 + h * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c4 = h * 8 + j;
            float4 v;
            v.x = fmaxf(__uint_as_float(r[4 * j + 0]) + bias_s[4 * c4 + 0], 0.f);
            v.y = fmaxf(__uint_as_float(r[4 * j + 1]) + bias_s[4 * c4 + 1], 0.f);
            v.z = fmaxf(__uint_as_float(r[4 * j + 2]) + bias_s[4 * c4 + 2], 0.f);
            v.w = fmaxf(__uint_as_float(r[4 * j + 3]) + bias_s[4 * c4 + 3], 0.f);
            *reinterpret_cast<float4*>(conv_s + pix * 64 + ((c4 ^ (pix & 15)) << 2)) = v;
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        if (++acc == 4) { acc = 0; acc_phase ^= 1u; }
      }
      named_bar(2, 128);
      // maxpool 3x3 / 2, pad 1 over the 16x16 map.  Values are >= 0 after ReLU, so the padding
      // is "max with 0" and clamping a window index re-reads an element already in the window.
      // Each thread owns one (output column px, channel quad c4) and slides down the 16 conv
      // rows: 3 loads per row, every row shared by two vertically adjacent outputs.
      {
        const int c4 = etid & 15, px = etid >> 4;
        const int ixa = (2 * px - 1) < 0 ? 0 : 2 * px - 1, ixb = 2 * px, ixc = 2 * px + 1;
        const float* pa = conv_s + ixa * 64 + ((c4 ^ ixa) << 2);
        const float* pb = conv_s + ixb * 64 + ((c4 ^ ixb) << 2);
        const float* pc = conv_s + ixc * 64 + ((c4 ^ ixc) << 2);
        auto hmax = [&](int iy) {
          const float4 u = *reinterpret_cast<const float4*>(pa + iy * 16 * 64);
          const float4 v = *reinterpret_cast<const float4*>(pb + iy * 16 * 64);
          const float4 w = *reinterpret_cast<const float4*>(pc + iy * 16 * 64);
          return make_float4(fmaxf(fmaxf(u.x, v.x), w.x), fmaxf(fmaxf(u.y, v.y), w.y),
                             fmaxf(fmaxf(u.z, v.z), w.z), fmaxf(fmaxf(u.w, v.w), w.w));
        };
        float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int py = 0; py < 8; ++py) {
          const float4 h0 = hmax(2 * py), h1 = hmax(2 * py + 1);
          float4 mx;
          mx.x = fmaxf(fmaxf(prev.x, h0.x), h1.x); mx.y = fmaxf(fmaxf(prev.y, h0.y), h1.y);
          mx.z = fmaxf(fmaxf(prev.z, h0.z), h1.z); mx.w = fmaxf(fmaxf(prev.w, h0.w), h1.w);
          prev = h1;
          const uint32_t h0p = pack_bf16x2(mx.x, mx.y), h1p = pack_bf16x2(mx.z, mx.w);
          const int64_t off = (t * 64 + (py * 8 + px)) * 64 + c4 * 4;
          *reinterpret_cast<uint2*>(a.out_hi + off) = make_uint2(h0p, h1p);
          if (a.out_lo) {
            const uint32_t l0 = pack_bf16x2(mx.x - bf16_lo_f(h0p), mx.y - bf16_hi_f(h0p));
            const uint32_t l1 = pack_bf16x2(mx.z - bf16_lo_f(h1p), mx.w - bf16_hi_f(h1p));
            *reinterpret_cast<uint2*>(a.out_lo + off) = make_uint2(l0, l1);
          }
        }
      }
      named_bar(2, 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// Host: packs the folded stem weights [64][3][7][7] into bf16 [64][192] (k = dy*24 + dx*3 + c).
void pack_stem_weights_bf16(const float* w_oihw, uint16_t* out) {
  auto rn = [](float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
  };
  for (int co = 0; co < 64; ++co) {
    for (int k = 0; k < kK; ++k) out[co * kK + k] = 0;
    for (int c = 0; c < 3; ++c)
      for (int dy = 0; dy < 7; ++dy)
        for (int dx = 0; dx < 7; ++dx)
          out[co * kK + dy * 24 + dx * 3 + c] = rn(w_oihw[((co * 3 + c) * 7 + dy) * 7 + dx]);
  }
}

int launch_stem_tc(const StemArgs& a, const void* w_bf16_dev, const uint16_t* lut_bf16_dev,
                   cudaStream_t st) {
  if (a.tile != kS) {
    set_error("tensor-core stem supports tile 32 only (got %d)", a.tile);
    return CS_ERR_UNSUPPORTED;
  }
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(Smem::total + 1024)));
    if (dev < 64) attr_done[dev] = true;
  }
  if (a.count <= 0) return CS_OK;
  StemTcParams p;
  int rc = make_mat_map_2d(&p.w_map, w_bf16_dev, kK, 64, kK, 64);
  if (rc != CS_OK) return rc;
  p.a = a;
  p.lut_bf16 = lut_bf16_dev;
  int grid = (int)(a.count < num_sms() ? a.count : num_sms());
  CS_CUDA(launch_pdl(stem_tc_kernel, dim3((unsigned)grid), dim3(kThreads), Smem::total + 1024, st, 1, p));
  return CS_OK;
}

}  // namespace cs
