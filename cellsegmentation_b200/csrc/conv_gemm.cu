// K2 (bf16 mode): the generic implicit-GEMM convolution kernel on tcgen05 tensor cores.
//
// Reference: BasicBlock.forward (model/resnet.py:28-43) and resnet_forward (:234-248) under
// eval-mode BN folded into the conv (model.cu does the folding/packing/planning).
//
//   out[r][n] = act( bias[n] + sum_steps A_step[r][0:64] . B[n][b_k : b_k+64]  (+ residual) )
//
// A is never materialised as im2col.  A K step is a TMA box:
//   4-D mode  rows = (instance, oy, ox): box {64 ch, W, H, instances} fetched at the tap's pixel
//             shift (dx, dy); out-of-image pixels are zero-filled by TMA = the conv's padding.
//             Stride-2 convs read one of four parity-phase maps of the input.
//   2-D mode  rows = instances, columns = (pixel, channel): maps with <= 2x2 outputs become dense
//             GEMMs whose all-zero K blocks were dropped on the host.
// Operands land in shared memory in the 128-byte-swizzled K-major layout UMMA expects;
// accumulators live in TMEM (two stages: the epilogue of tile i overlaps the MMAs of tile i+1).
//
// CTA pairs (CL = 2, tcgen05 cta_group::2): with both operands in shared memory a single-CTA
// MMA is bound by operand reads (~66 B/clk measured: 40-60 % tensor-pipe utilisation).  Two CTAs
// of a cluster take neighbouring M tiles of the same N tile; each loads its own A tile and HALF
// of the B tile, the leader (rank 0) issues one M = 256 MMA per K slice for both, and every CTA
// finds its 128 x N accumulator rows in its own TMEM.  Per CTA the B bytes read per MMA halve.
// All operand loads complete on the leader's "full" barrier; a multicast commit releases the
// stage in both CTAs and hands the accumulator to both epilogues; both epilogues return the
// accumulator stage on the leader's "tempty" barrier.
//
// Warp roles (352 threads, one CTA per SM, persistent):
//   warp 0 : TMA producer (one lane)          warp 1 : TMEM alloc + MMA issuer (one lane)
//   warps 2-9 : epilogue; warp % 4 = TMEM lane quadrant, (warp-2)/4 = column half.
//   warp 10 : epilogue DMA (residual tiles in, result tiles out by TMA; gemm_epilogue.cuh)
#include <stdlib.h>

#include "fwd.cuh"
#include "gemm_epilogue.cuh"

namespace cs {
namespace {

template <int BN, int CL>
struct GemmCfg {
  static constexpr int kMaxStages = 8;
  static constexpr uint32_t kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr uint32_t kBBytes = (BN / CL) * kGemmBK * 2;   // a pair member holds half of B
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kSmemLimit = 227 * 1024;
  // Shared memory: [stages x (A|B)] [staging: 2 sets of hi(+lo) tiles] [barriers]; the stage
  // count is whatever fits once the epilogue staging (16 or 32 KB per set) is reserved.
  static constexpr int stages_for(uint32_t staging_bytes) {
    int s = (int)((kSmemLimit - 1024 - 256 - staging_bytes) / kStageBytes);
    return s > kMaxStages ? kMaxStages : s;
  }
  static constexpr uint32_t smem_bytes(int stages, uint32_t staging_bytes) {
    return stages * kStageBytes + staging_bytes + 256 + 1024;
  }
  static constexpr int kChunks = BN / 64;                          // 64-column epilogue chunks
  static constexpr uint32_t kTmemCols = 2 * BN;                    // two accumulator stages
};

constexpr int kGemmThreads = 352;  // TMA warp, MMA warp, 8 epilogue warps, epilogue DMA warp

template <int BN, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CL>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int n_stages = p.stages;
  const uint32_t set_bytes = p.epi_set_bytes;                 // 16 KB (hi only) or 32 KB (hi + lo)
  const uint32_t staging = base + n_stages * Cfg::kStageBytes;
  const int n_sets = (int)p.epi_sets;                         // 2, or a ring of 4 (gemm_epilogue.cuh)
  const uint32_t bar_off = n_stages * Cfg::kStageBytes + n_sets * set_bytes;
  const uint32_t bar_base = base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kMaxStages + 2 + a); };
  EpiBars ebars;
  for (int s = 0; s < kEpiMaxSets; ++s) {
    ebars.res_full[s] = bar_base + 8u * (2 * Cfg::kMaxStages + 4 + s);
    ebars.out_ready[s] = bar_base + 8u * (2 * Cfg::kMaxStages + 4 + kEpiMaxSets + s);
  }
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
      base_ptr + bar_off + 8 * (2 * Cfg::kMaxStages + 4 + 2 * kEpiMaxSets));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      // single CTA: every epilogue thread arrives; pair: one elected lane per epilogue warp of
      // both CTAs arrives on the leader's barrier
      mbar_init(tempty_bar(a), CL == 1 ? kEpiThreads : CL * kEpiWarps);
    }
    epi_bars_init(ebars, kEpiThreads, n_sets);
    fence_barrier_init();
    prefetch_tmap(&p.b_map);
    prefetch_tmap(&p.a_map[0]);
  }
  if (warp == 1) {
    if (CL == 1) tmem_alloc(smem_u32((const void*)tmem_slot), Cfg::kTmemCols);
    else tmem_alloc_pair(smem_u32((const void*)tmem_slot), Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next layer may start its prologue on SMs that free up
  pdl_wait();                // ... and this one touches activations only after its producer is done

  // Work items of a cluster: (group of CL consecutive M tiles, N tile); CTA `rank` owns M tile
  // group*CL + rank.  CTAs whose M tile does not exist still run the protocol (zero-filled A,
  // stores predicated off).
  const int m_groups = (p.num_m_tiles + CL - 1) / CL;
  const int num_work = m_groups * p.num_n_tiles;
  const int first = blockIdx.x / CL, step_w = gridDim.x / CL;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = first; w < num_work; w += step_w) {
        const int wi = p.reverse ? num_work - 1 - w : w;
        const int m_tile = (wi / p.num_n_tiles) * CL + rank, n_tile = wi % p.num_n_tiles;
        const int var = p.n_variants > 1 ? n_tile : 0;
        const uint32_t* ext = p.ext_steps ? p.ext_steps + (size_t)var * (kMaxSteps + 1) : nullptr;
        const int ns = ext ? (int)__ldg(ext) : p.n_steps[var];
        for (int s = 0; s < ns; ++s) {
          KStep st;
          if (ext) st.v = __ldg(ext + 1 + s); else st = p.steps[var][s];
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          const int mp = st.map();
          if (CL == 1) {
            mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            if (((p.a_mode >> mp) & 1) == 0)
              tma_load_2d(a_dst, &p.a_map[mp], full_bar(stage), st.a_c0(), m_tile * kGemmBM);
            else
              tma_load_4d(a_dst, &p.a_map[mp], full_bar(stage), st.a_c0(), st.dx(), st.dy(),
                          m_tile * p.units_per_mtile);
            tma_load_2d(b_dst, &p.b_map, full_bar(stage), st.b_k(), n_tile * BN);
          } else {
            // both CTAs' boxes complete on the leader's barrier, which expects the pair's bytes
            if (rank == 0) mbar_expect_tx(full_bar(stage), CL * Cfg::kStageBytes);
            if (((p.a_mode >> mp) & 1) == 0)
              tma_load_2d_pair(a_dst, &p.a_map[mp], full_bar(stage), st.a_c0(), m_tile * kGemmBM);
            else
              tma_load_4d_pair(a_dst, &p.a_map[mp], full_bar(stage), st.a_c0(), st.dx(), st.dy(),
                               m_tile * p.units_per_mtile);
            tma_load_2d_pair(b_dst, &p.b_map, full_bar(stage), st.b_k(), n_tile * BN + rank * (BN / CL));
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop and waits; one elected lane issues the MMAs and commits of a
    // K step (tc_ptx.cuh: elect_one_sync keeps the issue sequence on the uniform datapath).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM * CL, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = first; w < num_work; w += step_w) {
        const int wi = p.reverse ? num_work - 1 - w : w;
        const int n_tile = wi % p.num_n_tiles;
        const int var = p.n_variants > 1 ? n_tile : 0;
        const int ns = p.ext_steps ? (int)__ldg(p.ext_steps + (size_t)var * (kMaxSteps + 1)) : p.n_steps[var];
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int s = 0; s < ns; ++s) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_src = base + stage * Cfg::kStageBytes;
            const uint64_t a_desc = umma_desc_sw128(a_src);
            const uint64_t b_desc = umma_desc_sw128(a_src + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kGemmBK / 16; ++k) {
              // +32 B per K=16 slice inside the 128-byte swizzle row: +2 in (addr >> 4)
              if (CL == 1)
                umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                          (s > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                               (s > 0 || k > 0) ? 1u : 0u);
            }
            // frees the stage when these MMAs retire -- in both CTAs of a pair
            if (CL == 1) umma_commit(empty_bar(stage));
            else umma_commit_pair(empty_bar(stage), kMask);
            if (s == ns - 1) {   // accumulator complete -> epilogue(s)
              if (CL == 1) umma_commit(tfull_bar(acc));
              else umma_commit_pair(tfull_bar(acc), kMask);
            }
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // 8 epilogue warps: TMEM lane quadrant = warp % 4, column half of each 64-column chunk =
    // (warp - 2) / 4.  They only touch TMEM and the shared-memory staging sets.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const EpiArgs ea{p.bias, p.res_hi != nullptr, p.res_lo != nullptr, p.out_hi != nullptr,
                     p.out_lo != nullptr, p.out_f32, p.relu};
    int acc = 0;
    uint32_t acc_phase = 0;
    int64_t q = 0;
    for (int w = first; w < num_work; w += step_w) {
      const int wi = p.reverse ? num_work - 1 - w : w;
      const int m_tile = (wi / p.num_n_tiles) * CL + rank, n_tile = wi % p.num_n_tiles;
      const int64_t row = (int64_t)m_tile * kGemmBM + r;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < Cfg::kChunks; ++c, ++q) {
        const int s = (int)(q & (n_sets - 1));
        mbar_wait(ebars.res_full[s], (uint32_t)((q >> (n_sets >> 1)) & 1));   // q / n_sets for 2 | 4
        const int col = n_tile * BN + c * 64 + half * 32;
        epi_chunk(ea, staging + s * set_bytes, r, half,
                  tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c * 64 + half * 32),
                  col, row < p.m_valid, row * (int64_t)p.n_total + col);
        fence_async_shared();
        mbar_arrive(ebars.out_ready[s]);
      }
      tc_fence_before();
      if (CL == 1) {
        mbar_arrive(tempty_bar(acc));
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (lane == 0) {
    // epilogue DMA thread: residual tiles in, result tiles out, one 64-column chunk at a time
    int n_items = 0;
    for (int w = first; w < num_work; w += step_w) ++n_items;
    auto coords = [&](int64_t q, int* col, int* row) {
      const int item = (int)(q / Cfg::kChunks), c = (int)(q % Cfg::kChunks);
      const int w = first + item * step_w;
      const int wi = p.reverse ? num_work - 1 - w : w;
      const int m_tile = (wi / p.num_n_tiles) * CL + rank, n_tile = wi % p.num_n_tiles;
      *col = n_tile * BN + c * 64;
      *row = m_tile * kGemmBM;
    };
    const bool rh = p.res_hi != nullptr, rl = p.res_lo != nullptr;
    const bool oh = p.out_hi != nullptr, ol = p.out_lo != nullptr;
    epi_dma_loop(
        (int64_t)n_items * Cfg::kChunks, staging, set_bytes, ebars,
        (rh ? kEpiTileBytes : 0u) + (rl ? kEpiTileBytes : 0u), oh || ol,
        [&](int64_t q, uint32_t set, uint32_t bar) {
          int col, row;
          coords(q, &col, &row);
          if (rh) tma_load_2d(set, &p.res_hi_map, bar, col, row);
          if (rl) tma_load_2d(set + kEpiTileBytes, &p.res_lo_map, bar, col, row);
        },
        [&](int64_t q, uint32_t set) {
          int col, row;
          coords(q, &col, &row);
          if (oh) tma_store_2d(&p.out_hi_map, set, col, row);
          if (ol) tma_store_2d(&p.out_lo_map, set + kEpiTileBytes, col, row);
        },
        n_sets);
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still write into it
  if (warp == 1) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// CELLSEG_EPI_RING=1: a ring of four staging sets for the residual GEMMs (gemm_epilogue.cuh).  Off by
// default: it removes the res_full stalls of those launches under ncu, but the step is power-bound
// and the same-box A/B is a tie (12.08-12.12 vs 12.12 M instances/s at 1 237 MHz, gpurun r2y).
const bool g_epi_ring = []() {
  const char* e = getenv("CELLSEG_EPI_RING");
  return e != nullptr && e[0] == '1' && e[1] == 0;
}();

template <int BN, int CL>
int launch_gemm_bn(const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CL>;
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, CL>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemLimit));
    if (dev < 64) attr_done[dev] = true;
  }
  GemmParams q = p;
  q.epi_set_bytes = (p.out_lo || p.res_lo) ? kEpiSetBytes : kEpiTileBytes;
  // optionally a ring of four 16 KB sets when a residual is loaded (one operand stage less)
  q.epi_sets = (g_epi_ring && p.res_hi && q.epi_set_bytes == kEpiTileBytes) ? 4u : 2u;
  q.stages = Cfg::stages_for(q.epi_sets * q.epi_set_bytes);
  const uint32_t smem = Cfg::smem_bytes(q.stages, q.epi_sets * q.epi_set_bytes);
  const int m_groups = (p.num_m_tiles + CL - 1) / CL;
  const int work = m_groups * p.num_n_tiles;
  int clusters = work < num_sms() / CL ? work : num_sms() / CL;
  CS_CUDA(launch_pdl(conv_gemm_kernel<BN, CL>, dim3((unsigned)(clusters * CL)), dim3(kGemmThreads), smem, st,
                     CL, q));
  return CS_OK;
}

}  // namespace

int launch_conv_gemm(const GemmParams& p, int BN, cudaStream_t st) {
  if (p.num_m_tiles <= 0 || p.num_n_tiles <= 0) return CS_OK;
  const int cl = p.cluster > 1 ? 2 : 1;
  switch (BN * 10 + cl) {
    case 641: return launch_gemm_bn<64, 1>(p, st);
    case 1281: return launch_gemm_bn<128, 1>(p, st);
    case 2561: return launch_gemm_bn<256, 1>(p, st);
    case 642: return launch_gemm_bn<64, 2>(p, st);
    case 1282: return launch_gemm_bn<128, 2>(p, st);
    case 2562: return launch_gemm_bn<256, 2>(p, st);
    default:
      set_error("launch_conv_gemm: unsupported N tile %d", BN);
      return CS_ERR_UNSUPPORTED;
  }
}

}  // namespace cs
