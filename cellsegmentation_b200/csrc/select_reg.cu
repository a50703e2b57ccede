// K3 fast path, register-resident form: adaptive top-k of one bag by one small CTA without a
// shared-memory copy of the bag.
//
// Reference: sample() inference.py:31-42 (np.lexsort + per-position predicate) and the
// pseudo-label rule dataset/dataset.py:168-169.  Same results as the exact shared-memory sort in
// select_topk.cu, which stays the fallback for every bag this path declines.
//
// The roofline of this stage is the 4 B/instance read (12.1 KB per 3025-instance bag).  The first
// fast path (select_fast.cu) staged the bag in shared memory and walked it three times with LDS:
// ~4 800 warp instructions per bag, i.e. issue-bound at 22 % of the HBM roofline (ncu r01_p:
// 41 % issue-active at 12 CTAs/SM).  Here
//   0. every thread pulls its NV 16-byte vectors of the bag straight into registers (aligned
//      superset of the bag; the up-to-three foreign words at either end are zeroed) -- all
//      loads of a bag are in flight at once and nothing is staged
//   1. the raw bit patterns (non-negative floats order like unsigned integers) reduce to thread
//      maxima.  tau = the n-th largest of a set of group maxima, so at least n instances are
//      >= tau: for small n the 32 column maxima, ranked by warp 0 with shuffles; for n >= 20 the
//      128 thread maxima themselves (every warp sorts its 32 in registers, every thread then
//      ranks its own maximum by binary searches in the four sorted lists) -- the finer groups
//      cut the candidates from ~3.3 n to ~1.15 n, and step 3 is quadratic in them
//   2. one pass over the registers marks the instances >= tau (a few per cent) in a per-thread
//      bit mask (two instructions per element); the marked ones are re-read (L1 hits) and
//      appended to a shared candidate list through a shared atomic
//   3. candidates are ranked by counting, the n best go straight to their output slots in
//      ascending (prob, index) order (ties keep the larger indices, like the stable lexsort)
// Declined (handled by the exact kernel through the fallback list): kept set not the plain
// suffix of the order (wrap-around cases), n > 128 (= THREADS), a negative / NaN / -0.0 probability, tau of
// +0.0 (padding words would qualify), more than 512 candidates (heavy ties), bags longer than
// the register budget (select_fast.cu takes those).
#include "common.cuh"
#include "select_common.cuh"

namespace cs {
namespace {

constexpr int kMaxCand = 512;
constexpr uint32_t kInf = 0x7f800000u;
constexpr int kFineN = 20;      // kept counts from here on take tau from the 128 thread maxima

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <int NV, int THREADS>
__global__ void __launch_bounds__(THREADS)
select_reg_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea,
                  int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list) {
  constexpr int kWarps = THREADS / 32;
  __shared__ unsigned long long cand[kMaxCand];
  __shared__ uint32_t tmax[THREADS];
  __shared__ uint32_t sorted_max[kWarps][32];
  __shared__ int s_count, s_bad;
  __shared__ uint32_t s_tau;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int64_t s = segs.start(b);
  const int T = (int)(segs.start(b + 1) - s);
  if (T <= 0) return;
  const Kept kr = kept_ranges(segs.gstart(b), T, segs.gtotal(),
                              bag_k(ea.labels, b, ea.tiles_per_pos, ea.topk_neg));
  const int n1 = kr.b1 - kr.a1, n2 = kr.b2 - kr.a2, n = n1 + n2;
  if (n == 0) return;                                     // all conditions block-uniform
  auto decline = [&]() {
    if (tid == 0) fb_list[atomicAdd(fb_count, 1)] = b;   // (fb_count was zeroed two stream ops earlier)
  };
  const float* src = prob + s;
  const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);   // words before the bag in its first vector
  const int nvec = (mis + T + 3) >> 2;
  const bool suffix = (n2 == 0 && kr.b1 == T) || (n1 == 0 && kr.b2 == T);
  if (!suffix || n > 128 || nvec > NV * THREADS) {
    decline();
    return;
  }

  // 0. the bag, as raw bits: vector v = tid + THREADS*j holds elements 4v - mis .. 4v - mis + 3
  const uint4* vsrc = reinterpret_cast<const uint4*>(src - mis);
  uint32_t x[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = tid + THREADS * j;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (v < nvec) q = ldg_stream(vsrc + v);
    x[j][0] = q.x; x[j][1] = q.y; x[j][2] = q.z; x[j][3] = q.w;
  }
  if (tid == 0) {                                          // words of the previous bag
    s_count = 0;
    s_bad = 0;
    s_tau = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c < mis) x[0][c] = 0u;
  }
  {                                                        // words of the next bag
    const int last = nvec - 1, end = mis + T - 4 * last;   // valid words in the last vector: 1..4
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (tid + THREADS * j == last) {
#pragma unroll
        for (int c = 1; c < 4; ++c)
          if (c >= end) x[j][c] = 0u;
      }
  }

  // 1. thread maxima -> column maxima -> tau (negative / NaN inputs have bit patterns above +inf
  // and surface in every maximum)
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) m = max(max(m, max(x[j][0], x[j][1])), max(x[j][2], x[j][3]));
  tmax[tid] = m;
  uint32_t top_w = m;                                      // warp-wide maximum: bad inputs surface here
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) top_w = max(top_w, __shfl_xor_sync(0xffffffffu, top_w, o));
  if (n >= kFineN) {
    // every warp sorts its 32 thread maxima (descending, bitonic network in registers) ...
    uint32_t v = m;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, v, j);
        const bool up = ((lane & k) == 0) == ((lane & j) == 0);   // keep the larger of the pair
        v = up ? max(v, other) : min(v, other);
      }
    }
    sorted_max[warp][lane] = v;                            // sorted_max[w][0] is warp w's largest
  }
  __syncthreads();
  if (n >= kFineN) {
    // ... and every thread ranks its own maximum among all THREADS maxima: elements greater than
    // it, plus equal ones held by lower thread ids (a strict total order, exactly one rank n-1)
    const unsigned same = __match_any_sync(0xffffffffu, m);
    const int eq_lower_lanes = __popc(same & ((1u << lane) - 1u));   // ties inside the own warp
    int rank = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      // descending list: count of entries > m (and >= m) by binary search
      int lo = 0, hi = 32;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_max[w][mid] > m) lo = mid + 1; else hi = mid; }
      int gt = lo;
      lo = gt; hi = 32;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_max[w][mid] >= m) lo = mid + 1; else hi = mid; }
      const int ge = lo;
      // equal entries of warp w: those with a lower thread id than this one come first
      const int eq_before = w < warp ? ge - gt : (w == warp ? eq_lower_lanes : 0);
      rank += gt + eq_before;
    }
    if (rank == n - 1) s_tau = m;
    if (lane == 0 && top_w > kInf) s_bad = 1;
  } else if (warp == 0) {
    uint32_t cm = tmax[lane];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) cm = max(cm, tmax[32 * w + lane]);
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint32_t mj = __shfl_sync(0xffffffffu, cm, j);
      rank += (mj > cm || (mj == cm && j < lane)) ? 1 : 0;
    }
    const unsigned pick = __ballot_sync(0xffffffffu, rank == n - 1);
    const unsigned first = __ballot_sync(0xffffffffu, rank == 0);
    const uint32_t t32 = __shfl_sync(0xffffffffu, cm, __ffs(pick) - 1);
    const uint32_t top = __shfl_sync(0xffffffffu, cm, __ffs(first) - 1);
    if (lane == 0) {
      s_tau = t32;
      if (top > kInf) s_bad = 1;
    }
  }
  __syncthreads();
  const uint32_t tau = s_tau;
  // bad input (negative / NaN / -0.0), or tau 0 (= +0.0: the zeroed padding words would qualify):
  // leave the bag to the exact kernel
  if (s_bad != 0 || tau == 0u) {
    decline();
    return;
  }

  // 2. candidates: (bits << 32 | index in the bag), appended in any order.  Mark first (ISETP +
  // predicated OR per element), then visit the few marked elements.
  if (m >= tau) {
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (x[j][c] >= tau) mask |= 1u << (4 * j + c);
    }
    while (mask) {
      const int q = __ffs(mask) - 1;
      mask &= mask - 1;
      const int e = 4 * (tid + THREADS * (q >> 2)) + (q & 3) - mis;
      const uint32_t bits = __float_as_uint(__ldg(src + e));   // L1 / L2 hit: the bag was just streamed
      const int pos = atomicAdd(&s_count, 1);
      if (pos < kMaxCand) cand[pos] = ((unsigned long long)bits << 32) | (unsigned)e;
    }
  }
  __syncthreads();
  const int count = s_count;
  if (count > kMaxCand) {   // heavy ties around the threshold
    decline();
    return;
  }

  // 3. rank by counting; the n largest go to slots o0 + (n-1-rank): ascending (prob, index).
  // The offsets come from the scan kernel this one was launched behind (programmatic dependent
  // launch): everything above overlapped with it.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t o0 = ea.out_offsets[b];
  const uint8_t pl = ea.labels[b] == 0 ? 0 : 1;
  for (int j = tid; j < count; j += THREADS) {
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

template <int NV, int THREADS>
cudaError_t launch_reg(const Segs& segs, const float* prob, const EmitArgs& ea, int32_t* fb_count,
                       int32_t* fb_list, cudaStream_t st) {
  return launch_pdl(select_reg_kernel<NV, THREADS>, dim3((unsigned)segs.n_bags), dim3(THREADS), 0, st, 1,
                    segs, prob, ea, fb_count, fb_list);
}

}  // namespace

// Register-resident fast path for bags of up to 4093 instances; *handled = false for longer
// bags (the caller falls back to launch_select_fast).  fb_count must be zero on entry; declined
// bags are appended to fb_list[0 .. *fb_count).
int launch_select_reg(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                      int32_t* fb_count, int32_t* fb_list, cudaStream_t st, bool* handled) {
  *handled = false;
  const int64_t words = max_T + 3;                      // worst-case misalignment
  cudaError_t e;
  if (words <= 2 * 128 * 4) e = launch_reg<2, 128>(segs, prob, ea, fb_count, fb_list, st);
  else if (words <= 6 * 128 * 4) e = launch_reg<6, 128>(segs, prob, ea, fb_count, fb_list, st);
  else if (words <= 8 * 128 * 4) e = launch_reg<8, 128>(segs, prob, ea, fb_count, fb_list, st);
  else return CS_OK;
  CS_CUDA(e);
  *handled = true;
  return CS_OK;
}

}  // namespace cs
