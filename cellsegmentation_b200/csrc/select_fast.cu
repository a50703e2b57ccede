// K3 fast path: adaptive top-k of one bag by one small CTA.
//
// Reference: sample() inference.py:31-42 (np.lexsort + per-position predicate) and the
// pseudo-label rule dataset/dataset.py:168-169.  Same results as the exact shared-memory sort
// in select_topk.cu, which stays as the fallback for every bag this path declines.
//
// The roofline of this stage is the 4 B/instance read (12.1 KB per 3025-instance bag); a full
// per-bag sort costs O(T log^2 T) shared-memory work and ran at 1 % of it.  Here a CTA of 128
// threads
//   0. copies the bag into shared memory with 16-byte cp.async (no registers held while the
//      bytes are in flight: ~40 registers/thread, 12 resident CTAs per SM hide the HBM latency)
//   1. reduces monotone keys (raw bits + 1; 0 = padding) to 32 column maxima when n <= 32
//      instances are kept, or to the 128 thread maxima when n <= 128.  The n-th largest group
//      maximum is a threshold tau with at least n instances >= tau
//   2. appends the instances >= tau to a <= 512-entry list of (key << 32 | index): per-thread
//      counts, one warp scan and ONE shared atomic per warp reserve the slots
//   3. ranks every candidate by counting the candidates above it -- no sort -- and writes the
//      n best straight to their output slots in ascending (prob, index) order (ties keep the
//      larger indices, like the stable lexsort)
// Declined (appended to the fallback list, handled by the exact kernel): kept set not the
// plain suffix [T-n, T) of the order (wrap-around cases), n > 128, any negative / NaN / -0.0
// probability, more than 512 candidates (heavy ties).
#include "common.cuh"
#include "select_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 128;
constexpr int kMaxCand = 512;
constexpr uint32_t kBadKey = 0xffffffffu;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// key of element j of the bag staged at `bits`: raw bits + 1 for p in [+0, +inf], kBadKey for
// negative (incl. -0.0) / NaN inputs.
__device__ __forceinline__ uint32_t key_of(uint32_t bits) {
  return bits > 0x7f800000u ? kBadKey : bits + 1u;
}

__global__ void __launch_bounds__(kThreads)
select_fast_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea, int max_T,
                   int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list) {
  extern __shared__ __align__(16) uint32_t stage[];   // max_T + 4 words
  __shared__ unsigned long long cand[kMaxCand];
  __shared__ uint32_t tmax[kThreads];
  __shared__ int s_count;
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
  const bool suffix = (n2 == 0 && kr.b1 == T) || (n1 == 0 && kr.b2 == T);
  if (!suffix || n > kThreads || T > max_T) {
    decline();
    return;
  }

  // 0. stage the bag: element j lands at stage[a + j], a = word misalignment of the source so
  // that source and destination share their 16-byte phase.
  const float* src = prob + s;
  const int a = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
  const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(stage);
  const int head = (4 - a) & 3;                       // elements before the first aligned 16 B
  const int n_head = head < T ? head : T;
  const int n_vec = (T - n_head) >> 2;
  const int tail0 = n_head + 4 * n_vec;
  if (tid < n_head) cp_async4(dst0 + 4 * (a + tid), src + tid);
  for (int v = tid; v < n_vec; v += kThreads)
    cp_async16(dst0 + 4 * (a + n_head + 4 * v), src + n_head + 4 * v);
  if (tid < T - tail0) cp_async4(dst0 + 4 * (a + tail0 + tid), src + tail0 + tid);
  if (tid == 0) s_count = 0;
  cp_async_wait_all();
  __syncthreads();
  const uint32_t* bits = stage + a;

  // 1. thread maxima (a bad key is the largest value, so it surfaces in every reduction)
  uint32_t m = 0;
  for (int j = tid; j < T; j += kThreads) m = max(m, key_of(bits[j]));
  tmax[tid] = m;
  __syncthreads();
  uint32_t tau;
  if (n <= 32) {   // every warp ranks the 32 column maxima itself: no extra barrier
    uint32_t cm = max(max(tmax[lane], tmax[32 + lane]), max(tmax[64 + lane], tmax[96 + lane]));
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint32_t mj = __shfl_sync(0xffffffffu, cm, j);
      rank += (mj > cm || (mj == cm && j < lane)) ? 1 : 0;
    }
    const unsigned pick = __ballot_sync(0xffffffffu, rank == n - 1);
    const unsigned top = __ballot_sync(0xffffffffu, rank == 0);
    tau = __shfl_sync(0xffffffffu, cm, __ffs(pick) - 1);
    if (__shfl_sync(0xffffffffu, cm, __ffs(top) - 1) == kBadKey) {
      decline();
      return;
    }
  } else {
    int rank = 0;
    uint32_t mx = 0;
    for (int j = 0; j < kThreads; ++j) {
      const uint32_t mj = tmax[j];
      mx = max(mx, mj);
      rank += (mj > m || (mj == m && j < tid)) ? 1 : 0;
    }
    if (mx == kBadKey) {
      decline();
      return;
    }
    if (rank == n - 1) s_tau = m;   // exactly one thread holds the n-th largest maximum
    __syncthreads();
    tau = s_tau;
  }

  // 2. candidates (key >= tau; tau >= 1 so padding never qualifies): count, reserve, write
  int mine = 0;
  for (int j = tid; j < T; j += kThreads) mine += key_of(bits[j]) >= tau ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  int base = 0;
  if (lane == 31) base = atomicAdd(&s_count, incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  int pos = base + incl - mine;
  for (int j = tid; j < T; j += kThreads) {
    const uint32_t k = key_of(bits[j]);
    if (k >= tau) {
      if (pos < kMaxCand) cand[pos] = ((unsigned long long)k << 32) | (unsigned)j;
      ++pos;
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
  for (int j = tid; j < count; j += kThreads) {
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

}  // namespace

// Launches the fast path; *handled = false when max_T does not fit shared memory.  fb_count
// must be zero on entry; declined bags are appended to fb_list[0 .. *fb_count).
int launch_select_fast(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                       int32_t* fb_count, int32_t* fb_list, cudaStream_t st, bool* handled) {
  *handled = false;
  if (max_T > 40000) return CS_OK;                 // 160 KB of staging: beyond that use the exact kernel
  const size_t smem = (size_t)(max_T + 4) * 4;
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(select_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (40000 + 4) * 4));
    if (dev < 64) attr_done[dev] = true;
  }
  CS_CUDA(launch_pdl(select_fast_kernel, dim3((unsigned)segs.n_bags), dim3(kThreads), smem, st, 1, segs, prob,
                     ea, (int)max_T, fb_count, fb_list));
  *handled = true;
  return CS_OK;
}

}  // namespace cs
