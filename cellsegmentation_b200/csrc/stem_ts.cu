// K1+K2a fused on tensor cores, weights-stationary form: unfold + normalise + conv7x7/2 + BN bias +
// ReLU + maxpool3x3/2 for 32x32 tiles with the WEIGHTS as the tensor-memory operand.
//
// Reference: crop/ToTensor/Normalize (dataset/dataset.py:409-416, 78-83) and
// conv1 -> bn1 -> relu -> maxpool (model/resnet.py:236-239), eval-mode BN folded.
//
// stem_win.cu reads both operands of every 128x64x16 MMA from shared memory: 6 KB for 32 cycles
// of math.  Operands stream at ~66 B/clk, so its 28 MMAs per instance take ~86 cycles each and
// the kernel sits at the shared-memory bandwidth (ncu r02: 1 343 tensor + 608 LSU wavefronts per
// instance in 2 415 cycles, 38 % tensor pipe).  Here the roles are swapped:
//     D[m][n] = sum_k A[m][k] * B[n][k]
//     A (tensor memory, written once per CTA): m = (16-channel group q, column parity b, channel)
//       = 128 rows: the 64 filters twice, the copy for odd conv columns shifted by two pixels
//     B (shared memory): the staged image itself, n = 8 * oy + j <-> conv pixels (oy, 2j + b)
// The staged tile is the same zero-padded, normalised bf16 image of 4-channel pixels as in
// stem_win.cu (rows of kPitch bytes, 16-byte chunks at their SWIZZLE_32B addresses).  For kernel
// row ky the inputs of BOTH conv columns 2j and 2j+1 lie in the 12 padded pixels from column 4j
// (K = 12 px x 4 ch = 48 = three K steps): kernel column kx sits in slot kx + 1 for the even and
// kx + 3 for the odd column.  Rows n and n + 1 are 4 pixels = 32 bytes apart, 8-row groups two
// image rows apart (SBO = 2 * kPitch), exactly the K-major SWIZZLE_32B operand of stem_win.cu.
// Per instance: 7 ky x 3 K steps = 21 MMAs of 128x128x16 (64 cycles of math each) that read only
// the 4 KB B slice from shared memory: 672 instead of 1 343 operand wavefronts.
//
// TMEM lane m of the accumulator holds ONE channel at one column parity, its 128 columns the
// conv pixels (oy, j).  Read with tcgen05.ld.32x32b a thread owns whole conv rows of its channel:
// the vertical window is register arithmetic on packed bf16x2 pairs, the horizontal one needs the
// other parity, which lives 16 lanes away in the same warp (one shuffle per pair).  A warp
// covers pooled rows 4h .. 4h+3 (h = 0 | 1) and re-reads conv row 7 instead of exchanging it.
// Bias is added in fp32 and the sums rounded to bf16 before the window max (rounding is
// monotonic, so this equals rounding last).  The [64 px][64 ch] tile leaves through a swizzled
// staging buffer (2-byte stores, conflict free) and one TMA store.
//
// Warp roles (416 threads, one CTA per SM, persistent over instances):
//   warps 0-3 producers   warps 4-11 epilogue (TMEM lane quarter = warp & 3, pooled-row half =
//   (warp - 4) >> 2; warps 4-7 also write the weights to tensor memory)   warp 12 MMA issue
#include "fwd.cuh"
#include "tc_ptx.cuh"

namespace cs {
namespace {

constexpr int kS = 32;                 // tile side
constexpr int kRows = kS + 6;          // padded rows (3 above, 3 below)
constexpr int kPitch = 320;            // bytes per padded row: 40 px x 8 B (4 left, 32, 4 right)
constexpr int kImgBytes = 12288;       // kRows * kPitch = 12160, rounded to 256
constexpr int kStages = 6;             // staged images in flight
constexpr int kAccs = 2;               // TMEM accumulators (one instance = 128 columns)
constexpr int kSlices = 21;            // 7 ky x 3 K steps
constexpr int kAccCol0 = 256;          // weights in columns 0 .. 167, accumulators at 256 and 384
constexpr int kWBytes = kSlices * 128 * 32;   // [slice][m 128][16 k] bf16
constexpr int kOutTile = 64 * 128;     // [64 px][64 ch] bf16, SWIZZLE_128B
constexpr int kEpiWarps = 8;
constexpr int kMmaWarp = 4 + kEpiWarps;
constexpr int kThreads = 32 * (kMmaWarp + 1);
static_assert(kRows * kPitch <= kImgBytes, "staged image");

struct Smem {
  static constexpr uint32_t img = 0;
  static constexpr uint32_t out = img + kStages * kImgBytes;    // 2 staging sets
  static constexpr uint32_t bars = out + 2 * kOutTile;
  static constexpr uint32_t total = bars + 256;
};
static_assert(Smem::out % 1024 == 0, "swizzled regions must stay aligned");

struct StemTsParams {
  CUtensorMap hi_map;           // [count*64 rows][64 ch] bf16, box {64, 64}, SWIZZLE_128B
  StemArgs a;
  const uint16_t* w_packed;     // kWBytes / 2 bf16
  float norm_a[3], norm_b[3];   // normalised value of byte v in channel c = fma(v, a[c], b[c])
};

__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// SWIZZLE_32B: 16-byte chunk bit (4) ^= address bit 7
__device__ __forceinline__ uint32_t swz32(uint32_t addr) { return addr ^ ((addr >> 3) & 16u); }

__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;                       // LBO (unused: K extent = one swizzle row)
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                       // SWIZZLE_32B
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]^T: lane m of the 8 columns at a_tmem holds row m of A, two
// consecutive k per 32-bit column.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint4& lo, const uint4& hi) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
stem_ts_kernel(const __grid_constant__ StemTsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t bars = base + Smem::bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };                     // kStages
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };        // kStages
  auto tfull_bar = [&](int q) { return bars + 8u * (2 * kStages + q); };    // kAccs
  auto tempty_bar = [&](int q) { return bars + 8u * (2 * kStages + kAccs + q); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(bp + Smem::bars + 8 * (2 * kStages + 2 * kAccs));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const StemArgs& a = p.a;

  // one-time init: zero the staged images (borders stay zero)
  for (int i = tid; i < kStages * kImgBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(bp + Smem::img)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
    for (int q = 0; q < kAccs; ++q) { mbar_init(tfull_bar(q), 1); mbar_init(tempty_bar(q), 32 * kEpiWarps); }
    fence_barrier_init();
    prefetch_tmap(&p.hi_map);
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
  fence_async_shared();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // weights -> tensor memory: warp 4 + q writes lanes 32q .. 32q+31, eight columns per slice
  if (warp >= 4 && warp < 8) {
    const int m = (warp & 3) * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(p.w_packed) + m * 2;
#pragma unroll 3
    for (int s = 0; s < kSlices; ++s) {
      const uint4 lo = __ldg(src + s * 256), hi = __ldg(src + s * 256 + 1);
      tmem_st_x8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(s * 8), lo, hi);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();   // the output buffers may still be read by the previous batch's kernels

  const int64_t n_inst = a.count;

  if (warp < 4) {
    // ================= producers: u8 / fp32 tile -> padded bf16x4 image =================
    // Image source: a thread owns 24 bytes (8 px) of one tile row and fetches the 7 aligned
    // words that cover them; the bytes are normalised with one FMA each, which rounds to the same
    // bf16 as the reference's (v/255 - mean)/std for all 3 x 256 inputs (checked on the host at
    // launch).  The raw words of the next kDepth instances are kept in flight in registers.
    constexpr int kDepth = 4;
    uint32_t pre[kDepth][7];
    uint32_t sh[kDepth];
    uint32_t pref[24];
    const bool from_img = a.x == nullptr;
    const int sy = tid >> 2, sq = tid & 3;   // tile row, 8-pixel quarter of that row
    auto prefetch_img = [&](uint32_t (&w)[7], uint32_t& shift, int64_t t) {
      int64_t inst = a.inst_begin + t;
      int64_t bag = inst / a.tiles_per_bag;
      int tl = (int)(inst - bag * a.tiles_per_bag);
      int gy = tl / a.grid_w, gx = tl - gy * a.grid_w;
      int row0 = grid_coord(gy, a.H, kS, a.interval), col0 = grid_coord(gx, a.W, kS, a.interval);
      const uint8_t* rowp = a.img + ((bag * a.H + row0 + sy) * (int64_t)a.W + col0 + 8 * sq) * 3;
      const uintptr_t pa = reinterpret_cast<uintptr_t>(rowp);
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(pa & ~(uintptr_t)3);
      shift = (uint32_t)(pa & 3) * 8u;
#pragma unroll
      for (int k = 0; k < 6; ++k) w[k] = __ldg(wp + k);
      w[6] = shift ? __ldg(wp + 6) : 0u;
    };
    auto prefetch_x = [&](int64_t t) {
      const float* src = a.x + t * (int64_t)(3 * kS * kS) + sy * kS + 8 * sq;  // NCHW fp32
#pragma unroll
      for (int i = 0; i < 24; ++i)   // element i = pixel i/3, channel i%3
        pref[i] = __float_as_uint(__ldg(src + (i % 3) * kS * kS + i / 3));
    };
    int stage = 0;
    uint32_t phase = 0;
    int64_t t = blockIdx.x;
    const int64_t step = gridDim.x;
    if (from_img) {
#pragma unroll
      for (int u = 0; u < kDepth; ++u)
        if (t + u * step < n_inst) prefetch_img(pre[u], sh[u], t + u * step);
    } else if (t < n_inst) {
      prefetch_x(t);
    }
    // padded row sy + 3, padded pixel 4 + 8*sq: four 16-byte chunks (2 pixels each)
    const uint32_t dst_off = (uint32_t)((sy + 3) * kPitch + (4 + 8 * sq) * 8);
    while (t < n_inst) {
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        if (t >= n_inst) break;
        float f[24];
        if (from_img) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const uint32_t word = __funnelshift_r(pre[u][k], pre[u][k + 1], sh[u]);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int i = 4 * k + m;
              f[i] = fmaf((float)((word >> (8 * m)) & 0xffu), p.norm_a[i % 3], p.norm_b[i % 3]);
            }
          }
          if (t + kDepth * step < n_inst) prefetch_img(pre[u], sh[u], t + kDepth * step);
        } else {
#pragma unroll
          for (int i = 0; i < 24; ++i) f[i] = __uint_as_float(pref[i]);
          if (t + step < n_inst) prefetch_x(t + step);
        }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t dst = base + Smem::img + stage * kImgBytes + dst_off;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack_bf16x2(f[6 * c + 0], f[6 * c + 1]);
          v.y = pack_bf16x2(f[6 * c + 2], 0.f);
          v.z = pack_bf16x2(f[6 * c + 3], f[6 * c + 4]);
          v.w = pack_bf16x2(f[6 * c + 5], 0.f);
          sts128v(swz32(dst + 16u * c), v);
        }
        fence_async_shared();
        mbar_arrive(full_bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
        t += step;
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issue =================
    // The whole warp walks the loop and waits; one elected lane issues.  All 21 descriptors of an
    // instance are the first one plus a constant (the start address field counts 16-byte units).
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int64_t t = blockIdx.x; t < n_inst; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t bd0 = desc_sw32(base + Smem::img + stage * kImgBytes, 2 * kPitch);
        const uint32_t d_tmem = tmem_base + (uint32_t)(kAccCol0 + acc * 128);
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
          for (int s = 0; s < 3; ++s)
            umma_bf16_ts(d_tmem, tmem_base + (uint32_t)((ky * 3 + s) * 8),
                         bd0 + (uint64_t)((ky * kPitch + s * 32) >> 4), idesc, (ky > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
      if (++acc == kAccs) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================= epilogue: bias, bf16 pack, pool in registers, ReLU, TMA store ========
    const int q = warp & 3;                    // TMEM lanes 32q ..: channels 16q .. 16q+15, both parities
    const int h = (warp - 4) >> 2;             // pooled rows 4h .. 4h+3
    const int etid = (warp - 4) * 32 + lane;
    const int b = lane >> 4;                   // column parity of this lane's conv pixels
    const int co = 16 * q + (lane & 15);
    const bool odd = b != 0;
    constexpr uint32_t kNegInf2 = 0xff80ff80u;
    const float bz = a.bias[co];
    // staging offsets: this thread writes pooled rows 4h + p, columns 4b + x (x = 0..3) of channel co
    uint32_t offx[4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
      offx[x] = (uint32_t)((4 * h * 8 + 4 * b + x) * 128) + ((((uint32_t)(co >> 3)) ^ (uint32_t)(4 * b + x)) << 4) +
                (uint32_t)(co & 7) * 2u;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_inst; t += gridDim.x, ++it) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      // conv rows 8h - 1 .. 8h + 7 of this lane's channel and parity: V[v][j], v = 0 is the row
      // above the warp's first window (row 7 for h = 1, outside the image for h = 0)
      uint32_t r_lo[32], r_hi[32], r_up[8];
      const uint32_t c0 = lane_addr + (uint32_t)(kAccCol0 + acc * 128 + 64 * h);
      tmem_ld32(c0, r_lo);
      tmem_ld32(c0 + 32u, r_hi);
      if (h) tmem_ld8(c0 - 8u, r_up);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == kAccs) { acc = 0; acc_phase ^= 1u; }
      // P[v][jj]: bf16x2 of (conv + bias) at conv row 8h - 1 + v, columns 2(2jj) + b and 2(2jj+1) + b
      uint32_t P[9][4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        P[0][jj] = h ? pack_bf16x2(__uint_as_float(r_up[2 * jj]) + bz, __uint_as_float(r_up[2 * jj + 1]) + bz)
                     : kNegInf2;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          P[1 + v][jj] = pack_bf16x2(__uint_as_float(r_lo[8 * v + 2 * jj]) + bz,
                                     __uint_as_float(r_lo[8 * v + 2 * jj + 1]) + bz);
          P[5 + v][jj] = pack_bf16x2(__uint_as_float(r_hi[8 * v + 2 * jj]) + bz,
                                     __uint_as_float(r_hi[8 * v + 2 * jj + 1]) + bz);
        }
      }
      // the staging set of this instance must have been read by the TMA store issued two
      // instances ago (thread 0 issues every store; only the latest group may stay in flight)
      const int set = it & 1;
      if (etid == 0) bulk_wait_read_1();
      named_bar(2, 32 * kEpiWarps);
      const uint32_t out_s = base + Smem::out + (uint32_t)(set * kOutTile);
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        // vertical window of pooled row 4h + pr: conv rows 2(4h+pr) - 1 .. + 1 = V[2pr .. 2pr+2]
        uint32_t vm[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) vm[jj] = max3_bf16x2(P[2 * pr][jj], P[2 * pr + 1][jj], P[2 * pr + 2][jj]);
        // horizontal window of pooled column px: conv columns 2px - 1 (odd, j = px - 1), 2px (even,
        // j = px), 2px + 1 (odd, j = px).  The even lane finishes px 0..3, the odd lane px 4..7.
        const uint32_t s0 = odd ? vm[0] : vm[2], s1 = odd ? vm[1] : vm[3];
        const uint32_t g0 = __shfl_xor_sync(0xffffffffu, s0, 16), g1 = __shfl_xor_sync(0xffffffffu, s1, 16);
        const uint32_t own0 = odd ? vm[2] : vm[0], own1 = odd ? vm[3] : vm[1];
        const uint32_t o0 = odd ? own0 : g0, o1 = odd ? own1 : g1;          // odd-parity pairs
        const uint32_t prev = odd ? vm[1] : kNegInf2;
        const uint32_t sh0 = __funnelshift_r(prev, o0, 16);                  // (odd[2jj-1], odd[2jj])
        const uint32_t sh1 = __funnelshift_r(o0, o1, 16);
        const uint32_t w0 = max_bf16x2(max3_bf16x2(own0, g0, sh0), 0u);     // window max, ReLU
        const uint32_t w1 = max_bf16x2(max3_bf16x2(own1, g1, sh1), 0u);
        const uint32_t row = out_s + (uint32_t)(pr * 1024);
        sts16(row + offx[0], w0);
        sts16(row + offx[1], w0 >> 16);
        sts16(row + offx[2], w1);
        sts16(row + offx[3], w1 >> 16);
      }
      fence_async_shared();
      named_bar(2, 32 * kEpiWarps);
      if (etid == 0) {
        tma_store_2d(&p.hi_map, out_s, 0, (int)(t * 64));
        bulk_commit_group();
      }
    }
    if (etid == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

uint16_t bf16_rn_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

struct NormConst { float a[3], b[3]; bool ok; };
NormConst make_norm_const() {
  const double mean[3] = {(double)0.485f, (double)0.456f, (double)0.406f};
  const double stdv[3] = {(double)0.229f, (double)0.224f, (double)0.225f};
  NormConst nc;
  float lut[768];
  get_norm_lut_host(lut);
  nc.ok = true;
  for (int c = 0; c < 3; ++c) {
    nc.a[c] = (float)(1.0 / (255.0 * stdv[c]));
    nc.b[c] = (float)(-mean[c] / stdv[c]);
    for (int v = 0; v < 256; ++v)
      if (bf16_rn_host(fmaf((float)v, nc.a[c], nc.b[c])) != bf16_rn_host(lut[c * 256 + v])) nc.ok = false;
  }
  return nc;
}

}  // namespace

// Host: packs the folded stem weights [64][3][7][7] into bf16 [ky 7][kstep 3][m 128][16]:
// row m = 32q + 16b + c' is filter 16q + c' for conv columns of parity b; k = slot * 4 + c
// with kx = 4 * kstep + slot - 1 - 2b (slots outside 0 <= kx < 7 and channel 3 stay zero).
void pack_stem_weights_ts(const float* w_oihw, uint16_t* out) {
  memset(out, 0, kWBytes);
  for (int ky = 0; ky < 7; ++ky)
    for (int s = 0; s < 3; ++s)
      for (int m = 0; m < 128; ++m) {
        const int q = m >> 5, b = (m >> 4) & 1, co = 16 * q + (m & 15);
        for (int slot = 0; slot < 4; ++slot) {
          const int kx = 4 * s + slot - 1 - 2 * b;
          if (kx < 0 || kx > 6) continue;
          for (int c = 0; c < 3; ++c)
            out[((ky * 3 + s) * 128 + m) * 16 + slot * 4 + c] =
                bf16_rn_host(w_oihw[((co * 3 + c) * 7 + ky) * 7 + kx]);
        }
      }
}

int stem_ts_weight_bytes() { return kWBytes; }

int launch_stem_ts(const StemArgs& a, const void* w_packed_dev, cudaStream_t st) {
  if (a.tile != kS) {
    set_error("tensor-core stem supports tile 32 only (got %d)", a.tile);
    return CS_ERR_UNSUPPORTED;
  }
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(stem_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(Smem::total + 1024)));
    if (dev < 64) attr_done[dev] = true;
  }
  if (a.count <= 0) return CS_OK;
  StemTsParams p;
  int rc = make_mat_map_2d(&p.hi_map, a.out_hi, 64, a.count * 64, 64, 64);
  if (rc != CS_OK) return rc;
  if (a.out_lo) {
    set_error("stem_ts writes the bf16 stream only (out_lo must be NULL)");
    return CS_ERR_INVALID_ARG;
  }
  p.a = a;
  p.w_packed = reinterpret_cast<const uint16_t*>(w_packed_dev);
  {
    // fma(v, a, b) must round to the same bf16 as the reference's (v/255 - mean)/std (the LUT)
    static const NormConst nc = make_norm_const();
    if (!nc.ok) {
      set_error("stem_ts: FMA normalisation does not reproduce the bf16 LUT");
      return CS_ERR_UNSUPPORTED;
    }
    for (int c = 0; c < 3; ++c) { p.norm_a[c] = nc.a[c]; p.norm_b[c] = nc.b[c]; }
  }
  int grid = (int)(a.count < num_sms() ? a.count : num_sms());
  CS_CUDA(launch_pdl(stem_ts_kernel, dim3((unsigned)grid), dim3(kThreads), Smem::total + 1024, st, 1, p));
  return CS_OK;
}

}  // namespace cs
