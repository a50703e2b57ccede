// K3 fast path, register-resident form: adaptive top-k of one bag by one small CTA without a
// shared-memory copy of the bag.
//
// Reference: sample() inference.py:31-42 (np.lexsort + per-position predicate) and the
// pseudo-label rule dataset/dataset.py:168-169.  Same results as the exact shared-memory sort in
// select_topk.cu, which stays the fallback for every bag this path declines.
//
// The roofline of this stage is the 4 B/instance read (12.1 KB per 3025-instance bag), but the
// kernels are INSTRUCTION-ISSUE bound: the staged fast path of round 1 (select_fast.cu) spent
// ~4 800 warp instructions per bag (22 % of the HBM roofline), the first register-resident
// version 2 400 (72 us for 20 000 bags at 64 % issue-active, ncu r2d/r2f).  This version is
// built to issue as little as possible per element:
//   0. every thread pulls its NV 16-byte vectors of the bag straight into registers (aligned
//      superset of the bag; the up-to-three foreign words at either end are zeroed by the two
//      threads that hold them) -- all loads of a bag are in flight at once, nothing is staged
//   1. raw bit patterns (non-negative floats order like unsigned integers) reduce to one maximum
//      per vector (kept) and one per thread; warp 0 folds the thread maxima into 64 column
//      maxima, sorts them with a register bitonic network (two per lane) and publishes
//      tau = the n-th largest: at least n instances are >= tau, and only ~1.1-1.35 n are
//   2. a thread marks its vectors whose maximum reaches tau (six compares); the few marked
//      vectors are re-read (L1 hits) and their elements >= tau appended to a shared candidate
//      list through a shared atomic
//   3. candidates are ranked by counting, the n best go straight to their output slots in
//      ascending (prob, index) order (ties keep the larger indices, like the stable lexsort)
// The bag's kept range (64-bit closed form) is worked out by thread 0 alone while the loads fly.
// (Kept counts above 64 -- count labels that large are rare -- take tau from the 128 thread
// maxima by plain counting.)
// Declined (handled by the exact kernel through the fallback list): kept set not the plain
// suffix of the order (wrap-around cases), n > 128, a negative / NaN / -0.0 probability, tau of
// +0.0 (padding words would qualify), more than 512 candidates (heavy ties), bags longer than
// the register budget (select_fast.cu takes those).
#include <stdlib.h>

#include "common.cuh"
#include "select_common.cuh"

namespace cs {
namespace {

constexpr int kMaxCand = 512;
constexpr uint32_t kInf = 0x7f800000u;
constexpr int kCols = 64;       // column maxima tau is picked from; also the largest n served

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// One compare-exchange stage of a 64-element bitonic network held two per lane (element i < 32
// in `a` of lane i, element i >= 32 in `b` of lane i - 32); final order descending.
template <int K, int J>
__device__ __forceinline__ void bitonic64_stage(uint32_t& a, uint32_t& b, int lane) {
  if (J == 32) {                      // partner = the lane's other register (only in the K = 64 merge)
    const uint32_t hi = max(a, b), lo = min(a, b);
    a = hi; b = lo;                   // (i & 64) == 0 for all: the lower index keeps the larger
  } else {
    const uint32_t oa = __shfl_xor_sync(0xffffffffu, a, J), ob = __shfl_xor_sync(0xffffffffu, b, J);
    const bool lower = (lane & J) == 0;
    const bool desc_a = (lane & K) == 0, desc_b = ((lane + 32) & K) == 0;
    a = (desc_a == lower) ? max(a, oa) : min(a, oa);
    b = (desc_b == lower) ? max(b, ob) : min(b, ob);
  }
}

// The same for 32 elements, one per lane; final order descending (rank r in lane r).
template <int K, int J>
__device__ __forceinline__ void bitonic32_stage(uint32_t& a, int lane) {
  const uint32_t o = __shfl_xor_sync(0xffffffffu, a, J);
  const bool lower = (lane & J) == 0, desc = (lane & K) == 0;
  a = (desc == lower) ? max(a, o) : min(a, o);
}

// Per-bag shared state; two copies alternate between consecutive bags of a persistent CTA (a
// thread is never more than one bag ahead of the slowest: every bag has block barriers).
struct SelState {
  unsigned long long cand[kMaxCand];
  uint32_t tmax[128];
  int count, n;                       // n: kept count, or -1 = this path declines the bag
  uint32_t tau;
  int pseudo_label;                   // labels[b] != 0
};

struct BagView {
  const float* src;                   // first instance of the bag
  const uint4* vsrc;                  // 16-byte aligned superset
  int64_t s;
  int T, mis, nvec;
  bool fits;
};

template <int NV, int THREADS>
__device__ __forceinline__ BagView bag_view(const Segs& segs, const float* __restrict__ prob, int b) {
  BagView v;
  v.s = segs.start(b);
  v.T = (int)(segs.start(b + 1) - v.s);
  v.src = prob + v.s;
  v.mis = (int)((reinterpret_cast<uintptr_t>(v.src) >> 2) & 3);   // words before the bag in its first vector
  v.nvec = (v.mis + v.T + 3) >> 2;
  v.fits = v.T > 0 && v.nvec <= NV * THREADS;
  v.vsrc = reinterpret_cast<const uint4*>(v.src - v.mis);
  return v;
}

// 0. the bag, as raw bits: vector v = tid + THREADS*j holds elements 4v - mis .. 4v - mis + 3
template <int NV, int THREADS>
__device__ __forceinline__ void bag_load(const BagView& bv, uint4 (&x)[NV], int tid) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = tid + THREADS * j;
    x[j] = make_uint4(0u, 0u, 0u, 0u);
    if (bv.fits && v < bv.nvec) x[j] = ldg_stream(bv.vsrc + v);
  }
}

// Steps 1-3 for the bag whose vectors are in x[].  Block-uniform control flow (every early
// return is taken by all threads), three block barriers.
template <int NV, int THREADS>
__device__ __forceinline__ void bag_process(const Segs& segs, const EmitArgs& ea, int b, const BagView& bv,
                                            uint4 (&x)[NV], SelState& st, int32_t* __restrict__ fb_count,
                                            int32_t* __restrict__ fb_list) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = bv.T, mis = bv.mis;
  if (tid == 0) {
    // the kept range of this bag under the literal predicate, while the loads are in flight
    const Kept kr = kept_ranges(segs.gstart(b), T, segs.gtotal(),
                                bag_k(ea.labels, b, ea.tiles_per_pos, ea.topk_neg));
    const int n1 = kr.b1 - kr.a1, n2 = kr.b2 - kr.a2, n = n1 + n2;
    const bool suffix = (n2 == 0 && kr.b1 == T) || (n1 == 0 && kr.b2 == T);
    st.n = (n == 0) ? 0 : ((suffix && n <= THREADS && bv.fits) ? n : -1);
    st.count = 0;
    st.tau = 0;
    st.pseudo_label = ea.labels[b] == 0 ? 0 : 1;
    if (mis > 0) x[0].x = 0u;                              // words of the previous bag
    if (mis > 1) x[0].y = 0u;
    if (mis > 2) x[0].z = 0u;
  }
  {                                                        // words of the next bag
    const int last = bv.nvec - 1;
    if ((last & (THREADS - 1)) == tid) {
      const int end = mis + T - 4 * last;                  // valid words in the last vector: 1..4
      const int jl = last / THREADS;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (j == jl) {
          if (end < 2) x[j].y = 0u;
          if (end < 3) x[j].z = 0u;
          if (end < 4) x[j].w = 0u;
        }
    }
  }

  // 1. vector maxima -> thread maximum -> 64 column maxima -> tau (negative / NaN inputs have bit
  // patterns above +inf and surface in every maximum)
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) m = max(m, max(max(x[j].x, x[j].y), max(x[j].z, x[j].w)));
  st.tmax[tid] = m;
  __syncthreads();
  const int n = st.n;
  if (n == 0) return;                                      // nothing kept (block-uniform)
  if (n < 0) {                                             // wrap-around ranges, n > 128, bag too long
    if (tid == 0) fb_list[atomicAdd(fb_count, 1)] = b;     // (fb_count was zeroed two stream ops earlier)
    return;
  }
  if (n > kCols) {
    // rare (count labels above 64): tau = the n-th largest of the 128 thread maxima, every thread
    // ranks its own by counting (ties: lower thread id first, so exactly one has rank n - 1)
    int rank = 0;
    for (int j = 0; j < THREADS; ++j) {
      const uint32_t mj = st.tmax[j];
      rank += (mj > m || (mj == m && j < tid)) ? 1 : 0;
    }
    if (m > kInf) st.tau = 0xffffffffu;                    // bad input: decline (wins over the store below)
    __syncthreads();
    if (rank == n - 1 && st.tau != 0xffffffffu) st.tau = m == 0u ? 0xffffffffu : m;
  } else if (warp == 0 && n <= 16 && ea.small_n_cols32) {
    // small kept counts (most positive bags): the n-th largest of 32 column maxima (a column = the
    // four / two threads lane, lane + 32, ...) is as sharp a threshold for n <= 16 as 64 columns are
    // for n <= 32 (~1.15-1.4 n candidates) and its network has 15 single-register stages, not 21
    // double ones
    uint32_t a = max(st.tmax[lane], st.tmax[lane + 32]);
    if (THREADS == 128) a = max(a, max(st.tmax[lane + 64], st.tmax[lane + 96]));
    bitonic32_stage<2, 1>(a, lane);
    bitonic32_stage<4, 2>(a, lane);   bitonic32_stage<4, 1>(a, lane);
    bitonic32_stage<8, 4>(a, lane);   bitonic32_stage<8, 2>(a, lane);   bitonic32_stage<8, 1>(a, lane);
    bitonic32_stage<16, 8>(a, lane);  bitonic32_stage<16, 4>(a, lane);  bitonic32_stage<16, 2>(a, lane);
    bitonic32_stage<16, 1>(a, lane);
    bitonic32_stage<32, 16>(a, lane); bitonic32_stage<32, 8>(a, lane);  bitonic32_stage<32, 4>(a, lane);
    bitonic32_stage<32, 2>(a, lane);  bitonic32_stage<32, 1>(a, lane);
    const uint32_t top = __shfl_sync(0xffffffffu, a, 0);
    const uint32_t t = __shfl_sync(0xffffffffu, a, n - 1);
    if (lane == 0) st.tau = (top > kInf || t == 0u) ? 0xffffffffu : t;
  } else if (warp == 0) {
    uint32_t a, c;
    if (THREADS == 128) {
      a = max(st.tmax[lane], st.tmax[lane + 64]);                // column lane
      c = max(st.tmax[lane + 32], st.tmax[lane + 96]);           // column lane + 32
    } else {                                                     // 64 threads: a thread is a column
      a = st.tmax[lane];
      c = st.tmax[lane + 32];
    }
    bitonic64_stage<2, 1>(a, c, lane);
    bitonic64_stage<4, 2>(a, c, lane);  bitonic64_stage<4, 1>(a, c, lane);
    bitonic64_stage<8, 4>(a, c, lane);  bitonic64_stage<8, 2>(a, c, lane);  bitonic64_stage<8, 1>(a, c, lane);
    bitonic64_stage<16, 8>(a, c, lane); bitonic64_stage<16, 4>(a, c, lane); bitonic64_stage<16, 2>(a, c, lane);
    bitonic64_stage<16, 1>(a, c, lane);
    bitonic64_stage<32, 16>(a, c, lane); bitonic64_stage<32, 8>(a, c, lane); bitonic64_stage<32, 4>(a, c, lane);
    bitonic64_stage<32, 2>(a, c, lane);  bitonic64_stage<32, 1>(a, c, lane);
    bitonic64_stage<64, 32>(a, c, lane); bitonic64_stage<64, 16>(a, c, lane); bitonic64_stage<64, 8>(a, c, lane);
    bitonic64_stage<64, 4>(a, c, lane);  bitonic64_stage<64, 2>(a, c, lane);  bitonic64_stage<64, 1>(a, c, lane);
    // descending: rank r sits in `a` of lane r (r < 32) or in `c` of lane r - 32
    const uint32_t top = __shfl_sync(0xffffffffu, a, 0);
    const uint32_t ta = __shfl_sync(0xffffffffu, a, (n - 1) & 31), tc = __shfl_sync(0xffffffffu, c, (n - 1) & 31);
    const uint32_t t = n <= 32 ? ta : tc;
    // bad input (negative / NaN / -0.0), or tau 0 (= +0.0: the zeroed padding words would
    // qualify): 0xffffffff tells everyone to leave the bag to the exact kernel
    if (lane == 0) st.tau = (top > kInf || t == 0u) ? 0xffffffffu : t;
  }
  __syncthreads();
  const uint32_t tau = st.tau;
  if (tau == 0xffffffffu) {
    if (tid == 0) fb_list[atomicAdd(fb_count, 1)] = b;
    return;
  }

  // The output offsets come from the scan kernel this one was launched behind (programmatic
  // dependent launch; everything above overlapped with it).  Ask for this bag's offset now, so
  // the L2 round trip hides behind steps 2 and 3.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t o0 = ea.out_offsets[b];
  const uint8_t pl = (uint8_t)st.pseudo_label;

  // 2. candidates: (bits << 32 | index in the bag), appended in any order
  if (m >= tau) {
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j)           // (the vector maxima are recomputed: two instructions each)
      if (max(max(x[j].x, x[j].y), max(x[j].z, x[j].w)) >= tau) mask |= 1u << j;
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const int v = tid + THREADS * j;
      const uint4 q = __ldg(bv.vsrc + v);                  // L1 / L2 hit: the bag was just streamed
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int e = 4 * v + c - mis;
        if (w[c] >= tau && e >= 0 && e < T) {
          const int pos = atomicAdd(&st.count, 1);
          if (pos < kMaxCand) st.cand[pos] = ((unsigned long long)w[c] << 32) | (unsigned)e;
        }
      }
    }
  }
  __syncthreads();
  const int count = st.count;
  if (count > kMaxCand) {   // heavy ties around the threshold
    if (tid == 0) fb_list[atomicAdd(fb_count, 1)] = b;
    return;
  }

  // 3. rank by counting; the n largest go to slots o0 + (n-1-rank): ascending (prob, index)
  for (int j = tid; j < count; j += THREADS) {
    const unsigned long long me = st.cand[j];
    int above = 0;
#pragma unroll 4
    for (int i = 0; i < count; ++i) above += st.cand[i] > me ? 1 : 0;
    if (above < n) {
      const int64_t p = o0 + (n - 1 - above);
      if (p < ea.capacity) {
        ea.idx_out[p] = (int32_t)(bv.s + (int64_t)(unsigned)(me & 0xffffffffull));
        ea.label_out[p] = pl;
      }
    }
  }
}

// One CTA per bag.  OCC = resident CTAs per SM the register allocation is held to (NV <= 6: 10 ->
// 48 registers, 12 -> 40 registers with four spilled words, 16 -> 32 registers with ~30).
template <int NV, int THREADS, int OCC>
__global__ void __launch_bounds__(THREADS, OCC)
select_reg_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea,
                  int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list) {
  static_assert(THREADS == 128 || THREADS == 64, "two threads, or one, per column");
  // The exact clean-up pass is launched behind this grid as a programmatic dependent: it may become
  // resident once every CTA of this grid has started, i.e. in the slots the last wave leaves free,
  // and waits there for this grid to complete.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ SelState st;
  const int b = blockIdx.x;
  const BagView bv = bag_view<NV, THREADS>(segs, prob, b);
  if (bv.T > 0) {                                         // block-uniform
    uint4 x[NV];
    bag_load<NV, THREADS>(bv, x, threadIdx.x);
    bag_process<NV, THREADS>(segs, ea, b, bv, x, st, fb_count, fb_list);
  }
  // every thread orders itself behind the offsets kernel before it exits, whichever way its bag
  // went: this grid must not complete (and release the clean-up pass) ahead of the offsets
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Persistent CTAs (experiment, off by default -- see g_persist): bag b + k * gridDim.x in turn, with
// the NEXT bag's vectors requested before the current one is worked on.
template <int NV, int THREADS>
__global__ void __launch_bounds__(THREADS, NV <= 6 ? 7 : 5)
select_reg_persistent_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea,
                             int32_t* __restrict__ fb_count, int32_t* __restrict__ fb_list) {
  static_assert(THREADS == 128, "two threads per column");
  __shared__ SelState st[2];
  int b = blockIdx.x;
  if (b < segs.n_bags) {
    BagView bv = bag_view<NV, THREADS>(segs, prob, b);
    uint4 x[NV];
    bag_load<NV, THREADS>(bv, x, threadIdx.x);
    for (int k = 0; b < segs.n_bags; ++k) {
      const int bn = b + (int)gridDim.x;
      BagView bvn = bv;
      uint4 xn[NV];
      if (bn < segs.n_bags) {
        bvn = bag_view<NV, THREADS>(segs, prob, bn);
        bag_load<NV, THREADS>(bvn, xn, threadIdx.x);
      }
      if (bv.T > 0) bag_process<NV, THREADS>(segs, ea, b, bv, x, st[k & 1], fb_count, fb_list);
      b = bn;
      bv = bvn;
#pragma unroll
      for (int j = 0; j < NV; ++j) x[j] = xn[j];
    }
  }
  // like select_reg_kernel: no thread leaves before the offsets kernel has completed
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// CELLSEG_SELECT_PERSIST=1 selects the persistent kernel.  Measured at 20 000 bags x 3025 (ncu r2i):
// 79.5 us against 54.6 us for one CTA per bag -- a bag's chain of three block barriers, the
// warp-0 sort and the ranking loop takes ~4 us however early its loads were issued, so what hides
// it is the NUMBER of resident CTAs (10 of 48 registers against 7 of 72), not the prefetch.
const bool g_persist = []() {
  const char* e = getenv("CELLSEG_SELECT_PERSIST");
  return e != nullptr && e[0] == '1';
}();

// CELLSEG_SELECT_OCC=12 | 16: more bags in flight per SM at the price of a tighter register budget
// (see select_reg_kernel).  Default 10.
const int g_occ = []() {
  const char* e = getenv("CELLSEG_SELECT_OCC");
  const int v = e != nullptr ? atoi(e) : 10;
  return v == 12 || v == 16 ? v : 10;
}();

// CELLSEG_SELECT_CTA=64: two warps per bag (a thread holds twice the vectors and is a column of
// its own), more bags in flight per SM and ~30 % fewer warp instructions per bag.  Bags keeping
// 65..128 instances go to the exact kernel in this form.  CELLSEG_SELECT_OCC64 = 12 | 14 | 16
// resident CTAs per SM for the 12-vector form (85 / 73 / 64 registers).
const int g_cta = []() {
  const char* e = getenv("CELLSEG_SELECT_CTA");
  return e != nullptr && atoi(e) == 64 ? 64 : 128;
}();
const int g_occ64 = []() {
  const char* e = getenv("CELLSEG_SELECT_OCC64");
  const int v = e != nullptr ? atoi(e) : 12;
  return v == 14 || v == 16 ? v : 12;
}();

template <int NV, int THREADS, int OCC>
cudaError_t launch_one(const Segs& segs, const float* prob, const EmitArgs& ea, int32_t* fb_count,
                       int32_t* fb_list, cudaStream_t st) {
  return launch_pdl(select_reg_kernel<NV, THREADS, OCC>, dim3((unsigned)segs.n_bags), dim3(THREADS), 0, st, 1,
                    segs, prob, ea, fb_count, fb_list);
}

template <int NV>
cudaError_t launch_reg(const Segs& segs, const float* prob, const EmitArgs& ea, int32_t* fb_count,
                       int32_t* fb_list, cudaStream_t st) {
  const int per_sm = NV <= 6 ? 7 : 5;
  if (g_persist && segs.n_bags > num_sms() * per_sm)
    return launch_pdl(select_reg_persistent_kernel<NV, 128>, dim3((unsigned)(num_sms() * per_sm)), dim3(128),
                      0, st, 1, segs, prob, ea, fb_count, fb_list);
  if (NV > 6) return launch_one<NV, 128, 8>(segs, prob, ea, fb_count, fb_list, st);
  if (g_occ == 12) return launch_one<NV, 128, 12>(segs, prob, ea, fb_count, fb_list, st);
  if (g_occ == 16) return launch_one<NV, 128, 16>(segs, prob, ea, fb_count, fb_list, st);
  return launch_one<NV, 128, 10>(segs, prob, ea, fb_count, fb_list, st);
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
  if (g_cta == 64 && !g_persist) {
    if (words <= 2 * 64 * 4) e = launch_one<2, 64, 16>(segs, prob, ea, fb_count, fb_list, st);
    else if (words <= 4 * 64 * 4) e = launch_one<4, 64, 16>(segs, prob, ea, fb_count, fb_list, st);
    else if (words <= 12 * 64 * 4) {
      if (g_occ64 == 16) e = launch_one<12, 64, 16>(segs, prob, ea, fb_count, fb_list, st);
      else if (g_occ64 == 14) e = launch_one<12, 64, 14>(segs, prob, ea, fb_count, fb_list, st);
      else e = launch_one<12, 64, 12>(segs, prob, ea, fb_count, fb_list, st);
    } else if (words <= 16 * 64 * 4) e = launch_one<16, 64, 9>(segs, prob, ea, fb_count, fb_list, st);
    else return CS_OK;
  } else if (words <= 2 * 128 * 4) e = launch_reg<2>(segs, prob, ea, fb_count, fb_list, st);
  else if (words <= 6 * 128 * 4) e = launch_reg<6>(segs, prob, ea, fb_count, fb_list, st);
  else if (words <= 8 * 128 * 4) e = launch_reg<8>(segs, prob, ea, fb_count, fb_list, st);
  else return CS_OK;
  CS_CUDA(e);
  *handled = true;
  return CS_OK;
}

}  // namespace cs
