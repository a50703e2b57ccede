// K2d: a whole layer-1 BasicBlock (8x8 x 64 channels) in one kernel:
//     y = relu(conv2(relu(conv1(x) + b1)) + b2 + x)
// with both convolutions in the y-sum form of conv_ysum.cu and the intermediate tensor kept in
// shared memory.
//
// Reference: BasicBlock.forward of layer1 (model/resnet.py:28-43).
//
// An 8x8 map is a whole instance, so a 3x3 convolution needs no pixels from other M tiles: the
// epilogue of conv1 can write its bf16 result straight into a shared-memory tile in the one-box
// operand layout of conv_ysum.cu ({64 ch, x = -1..8, 8 rows, 2 images} = 160 rows of 128 B, the
// columns x = -1 and 8 zero), where the MMAs of conv2 read it; the residual of conv2 is the
// block input, which is still resident as conv1's operand box.  Per instance and block HBM sees
// 8 KB in and 8 KB out instead of 40 KB (conv1: 8 + 8, conv2: 8 + 8 residual + 8), and shared
// memory sees neither the residual tile (TMA write + read) nor the store / reload of the
// intermediate tile.
//
// Schedule.  Jobs run in the order c1(0), [c1(k), c2(k-1)] for k = 1 .. n-1, c2(n-1) and
// alternate between the two accumulator stages, so while the epilogue of c1(k) builds the
// intermediate tile the tensor core works on c2(k-1), and nothing waits for a job's own epilogue.
// An x box lives from its load until the epilogue of c2 has read the residual (four stages);
// an intermediate tile is free again once c1 of the tile after next has completed (two slots:
// tcgen05.commit covers every earlier MMA of the issuing thread).
//
// CTA pairs only (cta_group::2, as in conv_ysum.cu): each CTA holds its own boxes and 96 of the
// 192 rows of every weight tile -- with both convolutions' weights resident (72 KB per CTA) the
// single-CTA form does not fit.
//
// Warp roles: warp 0 TMA producer   warp 1 TMEM alloc + MMA issue (leader CTA)   warps 2-17
// epilogue (TMEM lane quarter = warp % 4, 16-channel group = (warp - 2) / 4)   warp 18: output DMA
#include <stdlib.h>
#include <string.h>

#include "fwd.cuh"
#include "tc_ptx.cuh"

namespace cs {
namespace {

constexpr int EW = 16;                             // epilogue warps
constexpr int NI = 32 / EW;                        // 8-channel groups per epilogue thread
constexpr int CW = 8 * NI;                         // channels per epilogue warp
constexpr int kXStages = 4;
constexpr uint32_t kBoxBytes = 160 * 128;          // 2 images x 8 rows x 10 pixels x 64 ch bf16
constexpr uint32_t kBTile = 96 * 128;              // this CTA's half of one [192][64] weight tile
constexpr uint32_t kOutBytes = 128 * 128;          // [128 rows][64 ch] bf16, SWIZZLE_128B
constexpr uint32_t kXchBytes = 2 * EW * 64 * NI * 4;   // 2 job parities x EW warps x 64 NI floats
constexpr uint32_t kOffMid = kXStages * kBoxBytes;
constexpr uint32_t kOffOut = kOffMid + 2 * kBoxBytes;
constexpr uint32_t kOffW = kOffOut + kOutBytes;    // [conv 2][dx 3] weight tiles
constexpr uint32_t kOffXch = kOffW + 6 * kBTile;
constexpr uint32_t kOffBars = kOffXch + kXchBytes;
constexpr uint32_t kSmemBytes = kOffBars + 256 + 1024;
constexpr uint32_t kAccCols = 256;                 // accumulator stage pitch in TMEM columns
constexpr int kThreads = (EW + 3) * 32;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kOffMid % 1024 == 0 && kOffOut % 1024 == 0 && kOffW % 1024 == 0, "swizzled regions");

__device__ __forceinline__ void named_bar(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr));
}
__global__ void __launch_bounds__(kThreads, 1)
ysum_block_kernel(const __grid_constant__ YsumBlockParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bar_base = base + kOffBars;
  auto xfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto xempty_bar = [&](int s) { return bar_base + 8u * (kXStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kXStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kXStages + 2 + a); };
  auto midfull_bar = [&](int s) { return bar_base + 8u * (2 * kXStages + 4 + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * kXStages + 6);
  const uint32_t outready_bar = bar_base + 8u * (2 * kXStages + 7);
  const uint32_t outfree_bar = bar_base + 8u * (2 * kXStages + 8);
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBars + 8 * (2 * kXStages + 9));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  constexpr uint16_t kMask = 3;
  // a cluster walks groups of two neighbouring M tiles; CTA `rank` owns tile group * 2 + rank
  const int m_groups = (p.num_m_tiles + 1) / 2;
  const int first = blockIdx.x / 2, step_g = gridDim.x / 2;
  auto tile_of = [&](int gi) { return p.tile_base + (p.reverse ? m_groups - 1 - gi : gi) * 2 + rank; };
  const int n_tiles = first < m_groups ? (m_groups - 1 - first) / step_g + 1 : 0;

  // the intermediate tiles: the pad columns (box pixels 0 and 9) stay zero for the whole kernel
  for (int i = threadIdx.x; i < (int)(2 * kBoxBytes / 16); i += kThreads)
    reinterpret_cast<uint4*>(base_ptr + kOffMid)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kXStages; ++s) { mbar_init(xfull_bar(s), 1); mbar_init(xempty_bar(s), EW); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * EW);
      mbar_init(midfull_bar(a), 2 * EW);
    }
    mbar_init(w_bar, 1);
    mbar_init(outready_bar, EW * 32);
    mbar_init(outfree_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&p.x_box_map);
    prefetch_tmap(&p.b1_map);
    prefetch_tmap(&p.b2_map);
    prefetch_tmap(&p.out_map);
  }
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_slot), 512);
  fence_async_shared();          // the zeroed tiles are read by the async proxy (MMA operand)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the leader's barriers exist before the peer's loads signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  const uint32_t wres = base + kOffW;

  if (warp == 0) {
    if (lane == 0) {
      // weights are constants: fetch them while the previous layer still drains
      if (rank == 0) mbar_expect_tx(w_bar, 2 * 6 * kBTile);
      for (int dx = 0; dx < 3; ++dx) {
        tma_load_2d_pair(wres + dx * kBTile, &p.b1_map, w_bar, 0, dx * 192 + rank * 96);
        tma_load_2d_pair(wres + (3 + dx) * kBTile, &p.b2_map, w_bar, 0, dx * 192 + rank * 96);
      }
      pdl_wait();
      for (int k = 0; k < n_tiles; ++k) {
        const int s = k % kXStages;
        mbar_wait(xempty_bar(s), (uint32_t)(((k / kXStages) & 1) ^ 1));
        if (rank == 0) mbar_expect_tx(xfull_bar(s), 2 * kBoxBytes);
        tma_load_4d_pair(base + s * kBoxBytes, &p.x_box_map, xfull_bar(s), 0, -1, 0,
                         tile_of(first + k * step_g) * 2);
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop and waits; one elected lane issues a job's twelve MMAs
    // (tc_ptx.cuh: elect_one_sync keeps the issue sequence on the uniform datapath).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 192);
      mbar_wait(w_bar, 0);
      int acc = 0;
      uint32_t acc_phase = 0;
      auto job = [&](uint32_t box, int conv) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
          const uint64_t a0 = umma_desc_sw128_sbo(box, 1280u);
          const uint64_t b0 = umma_desc_sw128(wres + conv * 3 * kBTile);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(d_tmem, a0 + (uint64_t)(dx * 8 + 2 * k), b0 + (uint64_t)(dx * (kBTile >> 4) + 2 * k),
                             idesc, (dx > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(tfull_bar(acc), kMask);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      };
      for (int k = 0; k <= n_tiles; ++k) {
        if (k < n_tiles) {
          const int s = k % kXStages;
          mbar_wait(xfull_bar(s), (uint32_t)((k / kXStages) & 1));
          job(base + s * kBoxBytes, 0);
        }
        if (k >= 1) {
          const int t = k - 1;
          mbar_wait(midfull_bar(t & 1), (uint32_t)((t >> 1) & 1));
          job(base + kOffMid + (t & 1) * kBoxBytes, 1);
        }
      }
    }
  } else if (warp < 2 + EW) {
    pdl_wait();
    const int quad = warp & 3;                 // TMEM lanes 32*quad ..: image quad/2, rows y0 .. y0+3
    const int cg = (warp - 2) >> 2;            // channels CW*cg .. CW*cg + CW-1
    const int g = lane >> 2, t = lane & 3;     // x, channel-pair slot
    const bool upper = (quad & 1) == 0;        // owns image rows 0..3 (else 4..7)
    const int pair_bar = 1 + (quad >> 1) * (EW / 4) + cg;     // named barrier of the two quarters
    const int partner = (warp - 2) ^ 1;                        // same image, same channel group
    float2 bias1[NI], bias2[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      bias1[i] = make_float2(p.bias1[cg * CW + 8 * i + 2 * t], p.bias1[cg * CW + 8 * i + 2 * t + 1]);
      bias2[i] = make_float2(p.bias2[cg * CW + 8 * i + 2 * t], p.bias2[cg * CW + 8 * i + 2 * t + 1]);
    }
    // byte offsets of (row yy, channel group i): in the 128-row output tile, and in an operand box
    // (image quad/2, row 4*(quad&1) + yy, box pixel g + 1); both are 128-byte-swizzled rows
    uint32_t off_out[4][NI], off_box[4][NI];
#pragma unroll
    for (int yy = 0; yy < 4; ++yy)
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const uint32_t chunk = (uint32_t)(cg * NI + i);
        const uint32_t r_out = (uint32_t)(quad * 32 + yy * 8 + g);
        const uint32_t r_box = (uint32_t)((quad >> 1) * 80 + ((quad & 1) * 4 + yy) * 10 + g + 1);
        off_out[yy][i] = r_out * 128u + 4u * t + ((chunk ^ (r_out & 7u)) << 4);
        off_box[yy][i] = r_box * 128u + 4u * t + ((chunk ^ (r_box & 7u)) << 4);
      }
    constexpr int kXw = 64 * NI;               // floats per warp slot
    float* const xch = reinterpret_cast<float*>(base_ptr + kOffXch);
    float* const xw0 = xch + (warp - 2) * kXw + lane * 2 * NI;
    const float* const xr0 = xch + partner * kXw + lane * 2 * NI;
    const float2 zero2 = make_float2(0.f, 0.f);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t jobs = 0;
    // One job's accumulators -> the four summed rows of this thread, S[yy][i] (+ bias)
    auto sums = [&](const float2 (&bias)[NI], float2 (&S)[4][NI]) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      // R[dy][w][4i + 2h + e]: image row y0 + 2w + h, column x = g, channel CW*cg + 8i + 2t + e
      uint32_t R[3][2][4 * NI];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int w = 0; w < 2; ++w)
          tmem_ld_16x256b_x2(tmem_base + ((uint32_t)(quad * 32 + w * 16) << 16) +
                                 (uint32_t)(acc * kAccCols + dy * 64 + cg * CW),
                             R[dy][w]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      auto D2 = [&](int dy, int yy, int i) -> float2 {
        return make_float2(__uint_as_float(R[dy][yy >> 1][4 * i + 2 * (yy & 1)]),
                           __uint_as_float(R[dy][yy >> 1][4 * i + 2 * (yy & 1) + 1]));
      };
      // rows 3 | 4 of an image live in different lane quarters: exchange them (fp32, 2*NI per thread)
      const int par = (int)(jobs & 1);
      ++jobs;
      float* xw = xw0 + par * EW * kXw;
      const float* xr = xr0 + par * EW * kXw;
      {
        // D_0 of row 3 feeds row 4 of the lower quarter; D_2 of row 4 feeds row 3 of the upper one
        const float2 a = upper ? D2(0, 3, 0) : D2(2, 0, 0), b = upper ? D2(0, 3, 1) : D2(2, 0, 1);
        *reinterpret_cast<float4*>(xw) = make_float4(a.x, a.y, b.x, b.y);
      }
      named_bar(pair_bar, 64);
      float2 imp[NI];
      {
        const float4 v4 = *reinterpret_cast<const float4*>(xr);
        imp[0] = make_float2(v4.x, v4.y);
        imp[1] = make_float2(v4.z, v4.w);
      }
#pragma unroll
      for (int yy = 0; yy < 4; ++yy)
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          // D_0 of row y-1 + D_1 of row y + D_2 of row y+1 + bias, in that order
          const float2 above = yy > 0 ? D2(0, yy - 1, i) : (upper ? zero2 : imp[i]);
          const float2 below = yy < 3 ? D2(2, yy + 1, i) : (upper ? imp[i] : zero2);
          S[yy][i] = __fadd2_rn(__fadd2_rn(__fadd2_rn(above, D2(1, yy, i)), below), bias[i]);
        }
    };
    static_assert(NI == 2, "the exchange moves one float4 per thread");
    for (int k = 0; k <= n_tiles; ++k) {
      if (k < n_tiles) {
        // ---- epilogue of conv1: relu(sum + b1) -> bf16 -> intermediate tile k & 1 ----
        float2 S[4][NI];
        sums(bias1, S);
        const uint32_t mid = base + kOffMid + (uint32_t)(k & 1) * kBoxBytes;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy)
#pragma unroll
          for (int i = 0; i < NI; ++i)
            sts32(mid + off_box[yy][i], pack_bf16x2(fmaxf(S[yy][i].x, 0.f), fmaxf(S[yy][i].y, 0.f)));
        fence_async_shared();
        __syncwarp();
        // Plain (CTA-scope release) remote arrive, as for the accumulator barriers: the tile was
        // written with st.shared + fence.proxy.async by this CTA's own threads and is read by this
        // SM's tensor core; a .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR and cost
        // 27 % of the kernel's stall samples (ncu r02w).
        if (lane == 0) mbar_arrive_leader(midfull_bar(k & 1));
      }
      if (k >= 1) {
        // ---- epilogue of conv2: relu(sum + b2 + x) -> bf16 -> staging tile -> TMA store ----
        const int tl = k - 1, s = tl % kXStages;
        float2 S[4][NI];
        sums(bias2, S);
        const uint32_t xbox = base + (uint32_t)s * kBoxBytes;
        uint32_t res[4][NI];
#pragma unroll
        for (int yy = 0; yy < 4; ++yy)
#pragma unroll
          for (int i = 0; i < NI; ++i) res[yy][i] = lds32(xbox + off_box[yy][i]);
        mbar_wait(outfree_bar, (uint32_t)((tl & 1) ^ 1));
        const uint32_t stg = base + kOffOut;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy)
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const float2 v = __fadd2_rn(S[yy][i], make_float2(bf16_lo_f(res[yy][i]), bf16_hi_f(res[yy][i])));
            sts32(stg + off_out[yy][i], pack_bf16x2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)));
          }
        fence_async_shared();
        mbar_arrive(outready_bar);
        __syncwarp();
        if (lane == 0) mbar_arrive(xempty_bar(s));   // the residual has been read: the box is free
      }
    }
  } else if (lane == 0) {
    pdl_wait();
    for (int k = 0; k < n_tiles; ++k) {
      mbar_wait(outready_bar, (uint32_t)(k & 1));
      tma_store_2d(&p.out_map, base + kOffOut, 0, tile_of(first + k * step_g) * 128);
      bulk_commit_group();
      bulk_wait_read_all();
      mbar_arrive(outfree_bar);
    }
    bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its tiles
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

int launch_ysum_block(const YsumBlockParams& p, cudaStream_t st) {
  if (p.num_m_tiles <= 0) return CS_OK;
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(ysum_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kSmemBytes));
    if (dev < 64) attr_done[dev] = true;
  }
  const int sms = num_sms();
  const int groups = (p.num_m_tiles + 1) / 2;
  const int clusters = groups < sms / 2 ? groups : sms / 2;
  CS_CUDA(launch_pdl(ysum_block_kernel, dim3((unsigned)(clusters * 2)), dim3(kThreads), kSmemBytes, st, 2, p));
  return CS_OK;
}

}  // namespace cs
