// K3 fast path, warp-per-bag form (experiment, off by default -- see g_enable_warp for the
// measurement): adaptive top-k of one bag by ONE WARP, no block barrier.
//
// Reference: sample() inference.py:31-42 (np.lexsort + per-position predicate) and the
// pseudo-label rule dataset/dataset.py:168-169.  Same results as the exact shared-memory sort in
// select_topk.cu, which stays the fallback for every bag this path declines.
//
// Why: the CTA-per-bag kernel (select_reg.cu) is bound by the latency of a bag's dependent chain
// (loads -> three block barriers -> warp-0 sort -> ranking, ~4 us) times the bags per resident CTA
// (ten 128-thread CTAs per SM = ten bags in flight per SM, ncu r2i: 54.6 us for 20 000 bags of
// 3025, 1 450 warp instructions per bag at 55 % issue-active).  Here a warp owns a bag:
//   0. a lane pulls NV 16-byte vectors (vector v = lane + 32 j of the aligned superset of the bag)
//      straight into registers -- 24 vectors = 96 registers for a 3025-instance bag, sixteen bags
//      in flight per SM; the previous bag's words in vector 0 are zeroed in registers, the next
//      bag's words in the last vector have their marks masked off in step 2
//   1. vector maxima fold into four group maxima per lane; tau = a value with at least n of the
//      64 (n <= 64) or 128 (n <= 128) group maxima >= tau, found by a 23-step binary search on the
//      bit pattern (one REDUX per step; the low 8 mantissa bits are left open: tau is then at most
//      2^-15 below the n-th largest group maximum, which only admits a stray candidate)
//   2. each lane marks its elements >= tau in a 96-bit register mask (no branches), a shuffle
//      scan gives every lane its slot range in the warp's candidate list, and the few marked
//      elements are re-read (L1 / L2 hits) into the list as (bits << 32 | index)
//   3. candidates are ranked by counting against the list (shared-memory broadcasts); the n best
//      go straight to their output slots in ascending (prob, index) order (ties keep the larger
//      indices, like the stable lexsort)
// No shared atomics, no __syncthreads: ~900 warp instructions per bag.
// Declined (listed for the exact kernel): kept set not the plain suffix of the order (wrap-around
// cases), n > 128, a negative / NaN / -0.0 probability, tau of zero (padding words would qualify),
// more candidates than the list holds (heavy ties).  Bags longer than 3069 instances keep the
// CTA-per-bag kernels.
#include <stdlib.h>

#include "common.cuh"
#include "select_common.cuh"

namespace cs {
namespace {

constexpr uint32_t kInf = 0x7f800000u;
constexpr int kWarpsPerCta = 4;
constexpr int kCandW = 256;          // candidate slots per warp (2 KB)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <int NV>
__device__ __forceinline__ void bag_by_warp(const Segs& segs, const float* __restrict__ prob, const EmitArgs& ea,
                                            int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list,
                                            unsigned long long* cand, int b, int lane) {
  // groups of vectors per lane whose maxima tau is picked from: 4 (2 for the two-vector form)
  constexpr int NG = NV >= 4 ? 4 : 2, VPG = NV / NG;
  static_assert(NV % NG == 0, "whole groups of vectors per lane");

  const int64_t s = segs.start(b);
  const int T = (int)(segs.start(b + 1) - s);
  if (T <= 0) return;                                        // warp-uniform, like every exit below
  const float* src = prob + s;
  const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);   // words before the bag in its first vector
  const int nvec = (mis + T + 3) >> 2;

  // the kept range of this bag under the literal predicate (closed form; the 64-bit arithmetic is
  // done before the bag occupies the registers)
  const int32_t label = ea.labels[b];
  int n;
  {
    const int64_t k = label == 0 ? (int64_t)ea.topk_neg : (int64_t)label * (int64_t)ea.tiles_per_pos;
    const Kept kr = kept_ranges(segs.gstart(b), T, segs.gtotal(), k);
    const int n1 = kr.b1 - kr.a1, n2 = kr.b2 - kr.a2;
    n = n1 + n2;
    if (n == 0) return;                                      // nothing kept
    const bool suffix = (n2 == 0 && kr.b1 == T) || (n1 == 0 && kr.b2 == T);
    if (!(suffix && n <= 32 * NG && nvec <= NV * 32)) {
      if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = b;    // (fb_count was zeroed two stream ops earlier)
      return;
    }
  }

  // 0. the bag, as raw bits
  const uint4* vsrc = reinterpret_cast<const uint4*>(src - mis) + lane;
  uint4 x[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    x[j] = make_uint4(0u, 0u, 0u, 0u);
    if (lane + 32 * j < nvec) x[j] = ldg_stream(vsrc + 32 * j);
  }

  // foreign words: the previous bag's, in vector 0 of lane 0, are zeroed here; the next bag's, in
  // the last vector, stay (patching a register picked by a run-time index costs more than the whole
  // step) -- their marks are masked off in step 2 and the candidate count re-checked
  if (lane == 0) {
    if (mis > 0) x[0].x = 0u;
    if (mis > 1) x[0].y = 0u;
    if (mis > 2) x[0].z = 0u;
  }

  // 1. group maxima (negative / NaN inputs have bit patterns above +inf and surface in every maximum)
  uint32_t gm[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    uint32_t m = 0;
#pragma unroll
    for (int j = g * VPG; j < (g + 1) * VPG; ++j)
      m = max(m, max(max(x[j].x, x[j].y), max(x[j].z, x[j].w)));
    gm[g] = m;
  }
  const uint32_t top = __reduce_max_sync(0xffffffffu, max(max(gm[0], gm[1]), max(gm[2], gm[3])));
  uint32_t tau = 0;
  if (n <= 64) {
    // 64 columns: the two halves of a lane's vectors (the two vectors themselves when NV = 2)
    const uint32_t h0 = NG == 4 ? max(gm[0], gm[1]) : gm[0], h1 = NG == 4 ? max(gm[2], gm[3]) : gm[1];
#pragma unroll 1
    for (int bit = 30; bit >= 8; --bit) {
      const uint32_t c = tau | (1u << bit);
      const int cnt = __reduce_add_sync(0xffffffffu, (h0 >= c ? 1 : 0) + (h1 >= c ? 1 : 0));
      if (cnt >= n) tau = c;
    }
  } else {
#pragma unroll 1
    for (int bit = 30; bit >= 8; --bit) {
      const uint32_t c = tau | (1u << bit);
      const int cnt = __reduce_add_sync(0xffffffffu, (gm[0] >= c ? 1 : 0) + (gm[1] >= c ? 1 : 0) +
                                                         (gm[2] >= c ? 1 : 0) + (gm[3] >= c ? 1 : 0));
      if (cnt >= n) tau = c;
    }
  }
  // bad input (negative / NaN / -0.0), or tau 0 (the zeroed padding words would qualify): the
  // exact kernel takes the bag
  if (top > kInf || tau == 0u) {
    if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = b;
    return;
  }

  // 2. per-lane marks (element 4 j + c of the lane -> bit 4 j + c), slot ranges, candidate list
  constexpr int kMaskWords = (4 * NV + 31) / 32;
  uint32_t mk[kMaskWords];
#pragma unroll
  for (int w = 0; w < kMaskWords; ++w) mk[w] = 0u;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int w = (4 * j) >> 5, sh = (4 * j) & 31;
    if (x[j].x >= tau) mk[w] |= 1u << sh;
    if (x[j].y >= tau) mk[w] |= 2u << sh;
    if (x[j].z >= tau) mk[w] |= 4u << sh;
    if (x[j].w >= tau) mk[w] |= 8u << sh;
  }
  {
    // the lane that holds the bag's last vector keeps its first 4 jl + end marks (end = valid words
    // of that vector, 1..4); every other mark of the warp is an element of the bag or a zero
    const int last = nvec - 1;
    const int end = mis + T - 4 * last;
    const int vb = (last & 31) == lane ? 4 * (last >> 5) + end : 4 * NV;
#pragma unroll
    for (int w = 0; w < kMaskWords; ++w) {
      const int lo = vb - 32 * w;
      mk[w] &= lo >= 32 ? 0xffffffffu : (lo <= 0 ? 0u : (1u << lo) - 1u);
    }
  }
  int mine = 0;
#pragma unroll
  for (int w = 0; w < kMaskWords; ++w) mine += __popc(mk[w]);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const int count = __shfl_sync(0xffffffffu, incl, 31);
  // heavy ties around the threshold, or a foreign word among the group maxima left fewer than n
  if (count > kCandW || count < n) {
    if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = b;
    return;
  }
  int off = incl - mine;
  const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
  for (int w = 0; w < kMaskWords; ++w) {
    uint32_t mm = mk[w];
    while (mm) {
      const int slot = 32 * w + __ffs(mm) - 1;               // = 4 j + c
      mm &= mm - 1;
      const int e = 4 * (lane + 32 * (slot >> 2)) + (slot & 3) - mis;   // index in the bag: 0 .. T-1
      const uint32_t bits = __ldg(wsrc + e);                 // L1 / L2 hit: the bag was just streamed
      cand[off++] = ((unsigned long long)bits << 32) | (unsigned)e;
    }
  }
  __syncwarp();

  // The output offsets come from the scan kernel this one was launched behind (programmatic
  // dependent launch; everything above overlapped with it).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t o0 = ea.out_offsets[b];
  const uint8_t pl = label == 0 ? 0 : 1;

  // 3. rank by counting; the n largest go to slots o0 + (n-1-rank): ascending (prob, index)
  for (int j = lane; j < count; j += 32) {
    const unsigned long long me = cand[j];
    int above = 0;
#pragma unroll 4
    for (int i = 0; i < count; ++i) above += cand[i] > me ? 1 : 0;
    if (above < n) {
      const int64_t p = o0 + (n - 1 - above);
      if (p < ea.capacity) {
        ea.idx_out[p] = (int32_t)(s + (int64_t)(unsigned)(me & 0xffffffffull));
        ea.label_out[p] = pl;
      }
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(32 * kWarpsPerCta, NV <= 8 ? 8 : 4)
select_warp_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea,
                   int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list) {
  __shared__ unsigned long long cand_s[kWarpsPerCta][kCandW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b < segs.n_bags) bag_by_warp<NV>(segs, prob, ea, fb_count, fb_list, cand_s[warp], b, lane);
  // every thread orders itself behind the scan kernel before it exits, whichever way its bag went:
  // the grid must not complete (and release the exact kernel behind it) ahead of the offsets
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// CELLSEG_SELECT_WARP=1 selects this kernel; the default stays the CTA-per-bag kernel of select_reg.cu.
// Measured at 20 000 bags x 3025 (gpurun r2aa, same box, outputs identical): 75.8 us against 54.6 us
// under ncu (27.8 M warp instructions at 45 % issue-active, 13 resident warps per SM), whole call
// 0.124 against 0.070 ms.  A bag takes ~7.4 us in one warp against ~4 us in a four-warp CTA: with
// 127 registers per thread only 3-4 warps share a scheduler, too few to cover the dependent chains
// of the threshold search (23 REDUX round trips), the mark words and the ranking loop.  Second
// cost: the ~0.2 % of bags whose next-bag words beat their own threshold (count < n, small n) are
// declined, and a single declined 3025-instance bag costs the exact kernel ~40 us of latency.
const bool g_enable_warp = []() {
  const char* e = getenv("CELLSEG_SELECT_WARP");
  return e != nullptr && e[0] == '1';
}();

}  // namespace

// Warp-per-bag fast path for bags of up to 3069 instances; *handled = false for longer bags (or
// when not switched on): the caller goes on to launch_select_reg / launch_select_fast.  fb_count must
// be zero on entry; declined bags are appended to fb_list[0 .. *fb_count).
int launch_select_warp(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                       int32_t* fb_count, int32_t* fb_list, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!g_enable_warp) return CS_OK;
  const int64_t words = max_T + 3;                      // worst-case misalignment
  const dim3 grid((unsigned)ceil_div(segs.n_bags, kWarpsPerCta)), block(32 * kWarpsPerCta);
  cudaError_t e;
  if (words <= 2 * 32 * 4) e = launch_pdl(select_warp_kernel<2>, grid, block, 0, st, 1, segs, prob, ea, fb_count, fb_list);
  else if (words <= 8 * 32 * 4) e = launch_pdl(select_warp_kernel<8>, grid, block, 0, st, 1, segs, prob, ea, fb_count, fb_list);
  else if (words <= 24 * 32 * 4) e = launch_pdl(select_warp_kernel<24>, grid, block, 0, st, 1, segs, prob, ea, fb_count, fb_list);
  else return CS_OK;
  CS_CUDA(e);
  *handled = true;
  return CS_OK;
}

}  // namespace cs
