// K3: segmented (per-bag) ordering, adaptive top-k selection, threshold ranking.
//
// Reference:
//   sample()            inference.py:31-42   order = np.lexsort((probs, groups));
//                                            index[i] = groups[i] != groups[(i+k_i) % N]
//   make_train_data()   dataset/dataset.py:168-169   pseudo-label = labels[bag] != 0
//   rank()              test_tile.py:63-79   keep prob > threshold in lexsort order
//
// groups (tileIDX) is non-decreasing, so lexsort((probs, groups)) is an
// independent stable sort of every bag's probabilities: ascending prob, ties by
// ascending instance index, NaN last.  One CTA owns one bag: the bag's keys are
// loaded once from HBM (coalesced), ordered in shared memory by a bitonic network
// on (sortable-u32 key, u16 local index) pairs, and only the kept entries are
// written back.  No host round trip: per-bag counts -> device scan -> emit.
//
// Literal predicate.  With s = segment start, T = segment size, k' = k mod N and
// local sorted rank j, u = s + j + k':
//   u <  N : kept iff j >= T - k'
//   u >= N : kept iff j <  N - k'          (wrap-around into an earlier bag)
// i.e. ranks [max(0,T-k'), min(T,J0)) U [max(0,J0), min(T,N-k')) with J0 = N-k'-s.
// Normal case: the last min(k,T) ranks.  k' = 0 keeps nothing (also k = N).
//
// Algorithmic bytes: 4 B/instance read + 5 B per kept instance + 4 B/bag label.
#include <stdlib.h>

#include "common.cuh"
#include "select_common.cuh"

using namespace cs;

namespace {

constexpr int kThreads = 512;
constexpr int kMaxSegPow2 = 32768;  // 6 B * 32768 = 192 KB of shared memory

__global__ void __launch_bounds__(256)
rank_count_kernel(Segs segs, const float* __restrict__ prob, float thr,
                  int64_t* __restrict__ counts) {
  int b = blockIdx.x;
  int64_t s = segs.start(b), e = segs.start(b + 1);
  int c = 0;
  for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) c += (prob[i] > thr) ? 1 : 0;
  __shared__ int wsum[8];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w];
    counts[b] = t;
  }
}

// In-place exclusive scan of counts[0..n) into offsets[0..n]; offsets[n] = total.
// One CTA of 32 warps; warp w owns the contiguous slice [w*per, (w+1)*per) and walks it 32
// elements at a time (coalesced loads, shuffle scan, running carry); the 32 slice totals are
// scanned by warp 0 and added in a second coalesced pass.
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(int64_t* __restrict__ data, int n) {
  // dependents (the emit kernels) only need the offsets for their final stores: let them start
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ int64_t warp_tot[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // thread t owns the contiguous entries [t*per, (t+1)*per): independent loads, one block scan of
  // the 1024 partial sums, then the running prefixes are written back
  const int per = (n + 1023) / 1024;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  constexpr int kRegs = 24;                               // entries kept in registers (n <= 24 576)
  int64_t v[kRegs];
  int64_t sum = 0;
  if (per <= kRegs) {                                     // all loads in flight at once
#pragma unroll
    for (int k = 0; k < kRegs; ++k) v[k] = lo + k < hi ? data[lo + k] : 0;
#pragma unroll
    for (int k = 0; k < kRegs; ++k) sum += v[k];
  } else {
    for (int i = lo; i < hi; ++i) sum += data[i];
  }
  int64_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const int64_t w = warp_tot[lane];
    int64_t xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    warp_tot[lane] = xs - w;                              // exclusive prefix of the warp totals
    if (lane == 31) data[n] = xs;
  }
  __syncthreads();
  int64_t run = warp_tot[warp] + x - sum;                 // exclusive prefix of this thread's entries
  if (per <= kRegs) {
#pragma unroll
    for (int k = 0; k < kRegs; ++k) {
      if (lo + k < hi) data[lo + k] = run;
      run += v[k];
    }
  } else {
    for (int i = lo; i < hi; ++i) {
      const int64_t t = data[i];
      data[i] = run;
      run += t;
    }
  }
}

// Per-bag kept counts AND their exclusive scan in one launch (the counts are closed-form in the
// count labels): offsets[b] = sum of kept(b') for b' < b, offsets[n] = total.
// Every CTA works out the counts of 1024 bags (the 64-bit range formula is ~80 instructions per
// bag: a single CTA needed 25 us for 20 000 bags, and the selection kernel launched behind this
// one waits for the offsets before its stores); the CTA that takes the last ticket scans them in
// place, in chunks staged through shared memory (coalesced loads -> every thread sums a
// contiguous slice -> block scan of the 1024 partial sums -> coalesced stores).  `ticket` must be
// zero on entry.
constexpr int kOffChunk = 24576;     // 96 KB of dynamic shared memory: 20 000 bags are one chunk

__global__ void __launch_bounds__(1024)
select_offsets_kernel(Segs segs, const int32_t* __restrict__ labels, int32_t tiles_per_pos,
                      int32_t topk_neg, int64_t* __restrict__ offsets, int32_t* __restrict__ ticket) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ int32_t cnt[];      // kOffChunk entries (only the scanning CTA touches them)
  __shared__ int32_t warp_tot[32];
  __shared__ int s_last;
  const int n = segs.n_bags;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const int b = blockIdx.x * 1024 + tid;
    if (b < n) {
      const int64_t s = segs.start(b), e = segs.start(b + 1);
      offsets[b] = kept_ranges(segs.gstart(b), e - s, segs.gtotal(),
                               bag_k(labels, b, tiles_per_pos, topk_neg)).count();
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int64_t carry = 0;
  for (int base = 0; base < n; base += kOffChunk) {
    const int m = min(kOffChunk, n - base);
    // other CTAs' stores: read through L2; all of a thread's loads in flight before the first use
#pragma unroll
    for (int q0 = 0; q0 < kOffChunk / 1024; q0 += 8) {
      int64_t v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = tid + 1024 * (q0 + q);
        v[q] = i < m ? __ldcg(offsets + base + i) : 0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = tid + 1024 * (q0 + q);
        if (i < m) cnt[i] = (int32_t)v[q];
      }
    }
    __syncthreads();
    const int per = (m + 1023) / 1024;
    const int lo = min(tid * per, m), hi = min(lo + per, m);
    int32_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += cnt[i];
    int32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const int32_t w = warp_tot[lane];
      int32_t xs = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, xs, o);
        if (lane >= o) xs += y;
      }
      warp_tot[lane] = xs - w;                              // exclusive prefix of the warp totals
    }
    __syncthreads();
    int32_t run = warp_tot[warp] + x - sum;                 // exclusive prefix of this thread's slice
    for (int i = lo; i < hi; ++i) {
      const int32_t c = cnt[i];
      cnt[i] = run;
      run += c;
    }
    __syncthreads();
    for (int i = tid; i < m; i += 1024) offsets[base + i] = carry + cnt[i];
    const int last_t = (m - 1) / per;
    __syncthreads();
    if (tid == last_t) warp_tot[0] = run;                   // run == chunk total for the owner of the last slice
    __syncthreads();
    carry += warp_tot[0];
    __syncthreads();
  }
  if (tid == 0) offsets[n] = carry;
}

// The same for up to 31 x 1024 bags without the last-block pass: every CTA scans the counts of its
// 1024 bags, publishes the block total (bit 63 = ready) and adds the totals of the CTAs before it as
// soon as they show up (at most 30 words, read by one warp).  The whole grid is co-resident (<= 31
// CTAs) and a CTA only waits for lower-numbered ones, so the wait cannot deadlock.  One L2 round
// trip instead of ticket -> reload of all counts -> second scan: the selection kernel launched
// behind this one waits that much less for its offsets.  `flags` (gridDim.x words) must be zero on entry.
constexpr int kLookbackMaxBlocks = 31;

__global__ void __launch_bounds__(1024)
select_offsets_lookback_kernel(Segs segs, const int32_t* __restrict__ labels, int32_t tiles_per_pos,
                               int32_t topk_neg, int64_t* __restrict__ offsets,
                               unsigned long long* __restrict__ flags) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ int32_t warp_tot[32];
  __shared__ long long s_base;
  const int n = segs.n_bags;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x * 1024 + tid;
  int32_t cnt = 0;
  if (b < n) {
    const int64_t s = segs.start(b), e = segs.start(b + 1);
    cnt = kept_ranges(segs.gstart(b), e - s, segs.gtotal(), bag_k(labels, b, tiles_per_pos, topk_neg)).count();
  }
  int32_t x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const int32_t w = warp_tot[lane];
    int32_t xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    warp_tot[lane] = xs - w;                                // exclusive prefix of the warp totals
    const long long block_total = __shfl_sync(0xffffffffu, xs, 31);
    volatile unsigned long long* vf = flags;
    if (lane == 0) vf[blockIdx.x] = (1ull << 63) | (unsigned long long)block_total;
    long long before = 0;
    if (lane < (int)blockIdx.x) {
      unsigned long long f;
      do { f = vf[lane]; } while ((f >> 63) == 0ull);
      before = (long long)(f & ~(1ull << 63));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) {
      s_base = before;
      if (blockIdx.x == gridDim.x - 1) offsets[n] = before + block_total;
    }
  }
  __syncthreads();
  if (b < n) offsets[b] = s_base + (int64_t)(warp_tot[warp] + x - cnt);
}

// Offsets without any inter-CTA traffic (experiment, CELLSEG_SELECT_OFFSETS=recount): the kept counts are
// closed-form and cheap (~20 instructions for a bag no wrap-around reaches), so CTA i simply
// recomputes the counts of the i x 1024 bags before its own (i per thread) instead of waiting for
// the CTAs that own them.  No flags, hence no zero-fill launch in front: CTA 0 clears the
// declined-bag counter itself BEFORE it releases the dependent selection kernel (a dependent grid
// starts only once every CTA of this one has triggered).  Measured: no gain over memset + look-back.
__global__ void __launch_bounds__(1024)
select_offsets_recount_kernel(Segs segs, const int32_t* __restrict__ labels, int32_t tiles_per_pos,
                              int32_t topk_neg, int64_t* __restrict__ offsets, int32_t* __restrict__ fb_count) {
  if (blockIdx.x == 0) {                                    // block-uniform
    if (threadIdx.x == 0) {
      fb_count[0] = 0;
      __threadfence();
    }
    __syncthreads();                                         // no thread of CTA 0 triggers before the store is out
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ int32_t warp_tot[32];
  __shared__ long long warp_before[32];
  __shared__ long long s_base;
  const int n = segs.n_bags;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // uniform bags only (the host checks): a count needs nothing but the bag's label, and all the
  // labels a thread needs are requested before the first one is used
  const int64_t T = segs.uniform_T, N = segs.gtotal();
  auto count_of = [&](int b, int32_t label) -> int32_t {
    const int64_t k = label == 0 ? (int64_t)topk_neg : (int64_t)label * (int64_t)tiles_per_pos;
    return kept_ranges((int64_t)b * T + segs.g_off, T, N, k).count();
  };
  const int b = blockIdx.x * 1024 + tid;
  const int32_t my_label = b < n ? labels[b] : 0;
  int32_t lab[kLookbackMaxBlocks - 1];
#pragma unroll
  for (int q = 0; q < kLookbackMaxBlocks - 1; ++q) lab[q] = q < (int)blockIdx.x ? labels[q * 1024 + tid] : 0;
  const int32_t cnt = b < n ? count_of(b, my_label) : 0;
  long long before = 0;                                      // bags tid, tid + 1024, ... of the earlier blocks
#pragma unroll
  for (int q = 0; q < kLookbackMaxBlocks - 1; ++q)
    if (q < (int)blockIdx.x) before += count_of(q * 1024 + tid, lab[q]);
  int32_t x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 31) warp_tot[warp] = x;
  if (lane == 0) warp_before[warp] = before;
  __syncthreads();
  if (warp == 0) {
    const int32_t w = warp_tot[lane];
    int32_t xs = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    warp_tot[lane] = xs - w;                                // exclusive prefix of the warp totals
    long long base = warp_before[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    if (lane == 31) {
      s_base = base;
      if (blockIdx.x == gridDim.x - 1) offsets[n] = base + xs;
    }
  }
  __syncthreads();
  if (b < n) offsets[b] = s_base + (int64_t)(warp_tot[warp] + x - cnt);
}

// ---- per-bag sort + emit ----------------------------------------------------
enum Mode { kLexsort = 0, kSelect = 1, kRank = 2 };

template <int kMode>
__global__ void __launch_bounds__(kThreads)
seg_sort_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // Behind a fast-path kernel this grid is launched with the programmatic-dependent-launch
  // attribute: its CTAs become resident while the fast path's last wave drains and wait here for
  // that grid to complete (a no-op for a plain launch).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // Either every bag (one CTA each) or, after the fast path, only the bags it declined.
  const int n_items = ea.fb_list != nullptr ? *ea.fb_count : segs.n_bags;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int b = ea.fb_list != nullptr ? ea.fb_list[item] : item;
  const int64_t s = segs.start(b);
  const int T = (int)(segs.start(b + 1) - s);
  if (T <= 0) continue;
  if (kMode == kRank && ea.rank_fast_cap > 0 &&
      ea.out_offsets[b + 1] - ea.out_offsets[b] <= ea.rank_fast_cap)
    continue;                                   // emitted by rank_fast_kernel
  const int P = pow2_ceil(T);
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_raw);
  uint16_t* idx = reinterpret_cast<uint16_t*>(smem_raw + (size_t)P * 4);

  for (int i = threadIdx.x; i < P; i += kThreads) {
    key[i] = i < T ? cs::float_sort_key(prob[s + i]) : 0xffffffffu;
    idx[i] = (uint16_t)(i < T ? i : 0xffff);
  }
  __syncthreads();

  // Bitonic network on (key, idx); padding (0xffffffff, 0xffff) sorts after every
  // real entry, including NaNs (their idx is < 0xffff because T <= 32768).
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += kThreads) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo | j;
        uint32_t ka = key[lo], kb = key[hi];
        uint16_t ia = idx[lo], ib = idx[hi];
        bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
        bool asc = (lo & k) == 0;
        if (a_gt_b == asc) {
          key[lo] = kb; key[hi] = ka;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }

  if (kMode == kLexsort) {
    for (int r = threadIdx.x; r < T; r += kThreads) ea.idx_out[s + r] = (int32_t)(s + idx[r]);
  } else if (kMode == kSelect) {
    const Kept kr = kept_ranges(segs.gstart(b), T, segs.gtotal(),
                                bag_k(ea.labels, b, ea.tiles_per_pos, ea.topk_neg));
    const int64_t o0 = ea.out_offsets[b];
    const uint8_t pl = ea.labels[b] == 0 ? 0 : 1;
    const int n1 = kr.b1 - kr.a1, n = kr.count();
    for (int q = threadIdx.x; q < n; q += kThreads) {
      int r = q < n1 ? kr.a1 + q : kr.a2 + (q - n1);
      int64_t pos = o0 + q;
      if (pos < ea.capacity) {
        ea.idx_out[pos] = (int32_t)(s + idx[r]);
        ea.label_out[pos] = pl;
      }
    }
  } else {  // kRank: prob > thr is a suffix of the ascending order, NaN excluded
    const int64_t o0 = ea.out_offsets[b];
    const int n = (int)(ea.out_offsets[b + 1] - o0);
    // first rank whose prob > thr: kept entries are contiguous up to the first NaN
    const uint32_t kthr = cs::float_sort_key(ea.thr);
    // binary search for first key > kthr (thr NaN -> nothing is kept, n == 0)
    int lo = 0, hi = T;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (key[mid] > kthr) hi = mid; else lo = mid + 1;
    }
    for (int q = threadIdx.x; q < n; q += kThreads) {
      int r = lo + q;
      int64_t pos = o0 + q;
      if (pos < ea.capacity) {
        int64_t gi = s + idx[r];
        ea.idx_out[pos] = (int32_t)gi;
        if (ea.prob_out) ea.prob_out[pos] = prob[gi];
      }
    }
  }
  __syncthreads();   // shared memory is reused by the next item
  }
}

// rank() fast path: with a threshold like 0.95 a bag keeps a few per cent of its instances, so
// instead of ordering the whole bag the CTA streams it once, appends the instances > thr to a
// shared list (their number is already known from the offsets) and ranks those by counting:
// ascending (prob, index), exactly the suffix of the lexsort order the exact kernel would emit.
// Bags that keep more than kRankCap entries are left to the exact kernel.
constexpr int kRankCap = 512;

__global__ void __launch_bounds__(128)
rank_fast_kernel(Segs segs, const float* __restrict__ prob, EmitArgs ea) {
  __shared__ unsigned long long cand[kRankCap];
  __shared__ int s_count;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int64_t o0 = ea.out_offsets[b];
  const int n = (int)(ea.out_offsets[b + 1] - o0);
  if (n <= 0 || n > kRankCap) return;
  const int64_t s = segs.start(b);
  const int T = (int)(segs.start(b + 1) - s);
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int j0 = 0; j0 < T; j0 += 128) {
    const int j = j0 + tid;
    const float p = j < T ? prob[s + j] : 0.f;
    const bool in = j < T && p > ea.thr;           // NaN compares false, like the reference
    const unsigned mask = __ballot_sync(0xffffffffu, in);
    int base = 0;
    if (lane == 0 && mask) base = atomicAdd(&s_count, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (in) {
      const int pos = base + __popc(mask & ((1u << lane) - 1u));
      if (pos < kRankCap) cand[pos] = ((unsigned long long)cs::float_sort_key(p) << 32) | (unsigned)j;
    }
  }
  __syncthreads();
  const int count = s_count < kRankCap ? s_count : kRankCap;     // == n
  for (int j = tid; j < count; j += 128) {
    const unsigned long long me = cand[j];
    int below = 0;
#pragma unroll 4
    for (int i = 0; i < count; ++i) below += cand[i] < me ? 1 : 0;
    const int64_t pos = o0 + below;
    if (pos < ea.capacity) {
      const int64_t gi = s + (int64_t)(unsigned)(me & 0xffffffffull);
      ea.idx_out[pos] = (int32_t)gi;
      if (ea.prob_out) ea.prob_out[pos] = prob[gi];
    }
  }
}

// Diagnostics: CELLSEG_SELECT_FAST=0 routes every bag through the exact sort kernel.
const bool g_disable_fast = []() {
  const char* e = getenv("CELLSEG_SELECT_FAST");
  return e != nullptr && e[0] == '0';
}();

// CELLSEG_SELECT_FAST=staged keeps the round-1 shared-memory fast path for every bag size.
const bool g_staged_fast = []() {
  const char* e = getenv("CELLSEG_SELECT_FAST");
  return e != nullptr && e[0] == 's';
}();

// Offsets kernel: 1 = look-back scan (default), 0 = CELLSEG_SELECT_OFFSETS=recount (uniform bags
// only; a tie with the look-back scan: 0.0685 vs 0.0675 ms per 20 000 bags, gpurun r2af), 2 = =ticket
// (the last-block scan; also what runs for more than 31 744 bags).
const int g_offsets_mode = []() {
  const char* e = getenv("CELLSEG_SELECT_OFFSETS");
  return e == nullptr ? 1 : (e[0] == 'r' ? 0 : (e[0] == 't' ? 2 : 1));
}();

// CELLSEG_SELECT_SORT_PDL=0: plain stream-ordered launch of the exact clean-up pass.
const bool g_sort_plain_launch = []() {
  const char* e = getenv("CELLSEG_SELECT_SORT_PDL");
  return e != nullptr && e[0] == '0';
}();

// CELLSEG_SELECT_COL32=0: always take the threshold from 64 column maxima (select_reg.cu).
const bool g_cols32 = []() {
  const char* e = getenv("CELLSEG_SELECT_COL32");
  return !(e != nullptr && e[0] == '0');
}();

struct SegHostInfo {
  int max_pow2;
};

// Largest segment decides the shared-memory footprint.  Uniform segments are known
// on the host; ragged ones are bounded by the caller through `uniform_T` (used as
// the maximum segment size when offsets are given).
int seg_smem_bytes(int64_t max_T, size_t* bytes) {
  CS_REQUIRE(max_T > 0 && max_T <= kMaxSegPow2,
             "segment size %lld outside (0, %d]: unsupported by the per-bag sort",
             (long long)max_T, kMaxSegPow2);
  int p = 1;
  while (p < max_T) p <<= 1;
  *bytes = (size_t)p * 6;
  return CS_OK;
}

template <int kMode>
int launch_sort(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                cudaStream_t st) {
  size_t smem = 0;
  int rc = seg_smem_bytes(max_T, &smem);
  if (rc != CS_OK) return rc;
  static bool attr_set[64][3] = {{false}};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_set[dev][kMode]) {
    CS_CUDA(cudaFuncSetAttribute(seg_sort_kernel<kMode>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSegPow2 * 6));
    if (dev < 64) attr_set[dev][kMode] = true;
  }
  const int grid = ea.fb_list != nullptr ? (segs.n_bags < cs::num_sms() * 4 ? segs.n_bags : cs::num_sms() * 4) : segs.n_bags;
  if (ea.fb_list != nullptr && ea.sort_pdl && !g_sort_plain_launch) {
    // clean-up pass behind a fast-path kernel: launch + CTA start-up overlap that kernel's tail
    CS_CUDA(cs::launch_pdl(seg_sort_kernel<kMode>, dim3((unsigned)grid), dim3(kThreads), smem, st, 1, segs, prob, ea));
    return CS_OK;
  }
  seg_sort_kernel<kMode><<<grid, kThreads, smem, st>>>(segs, prob, ea);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

int check_segs(const char* fn, const float* prob, const int64_t* seg_offsets, int64_t uniform_T,
               int n_bags) {
  CS_REQUIRE(prob != nullptr, "%s: prob is NULL", fn);
  CS_REQUIRE(n_bags > 0, "%s: n_bags = %d", fn, n_bags);
  CS_REQUIRE(uniform_T > 0,
             "%s: uniform_T must be > 0 (the segment size, or the maximum segment size when "
             "seg_offsets is given)", fn);
  if (!seg_offsets)
    CS_REQUIRE((int64_t)n_bags * uniform_T < (int64_t)INT32_MAX,
               "%s: %lld instances do not fit int32 indices", fn,
               (long long)n_bags * (long long)uniform_T);
  return CS_OK;
}

}  // namespace

extern "C" {

int cs_lexsort_segments(const float* prob, const int64_t* seg_offsets, int64_t uniform_T,
                        int n_bags, int32_t* order_out, void* stream) {
  int rc = check_segs("cs_lexsort_segments", prob, seg_offsets, uniform_T, n_bags);
  if (rc != CS_OK) return rc;
  CS_REQUIRE(order_out != nullptr, "cs_lexsort_segments: order_out is NULL");
  Segs segs{seg_offsets, uniform_T, n_bags, 0, 0};
  EmitArgs ea{};
  ea.idx_out = order_out;
  return launch_sort<kLexsort>(segs, prob, ea, uniform_T, cs::as_stream(stream));
}

int cs_select_topk(const float* prob, const int64_t* seg_offsets, int64_t uniform_T, int n_bags,
                   const int32_t* labels, int32_t tiles_per_pos, int32_t topk_neg,
                   int32_t* sel_idx_out, uint8_t* sel_label_out, int64_t* sel_offsets_out,
                   int64_t capacity, void* workspace, int64_t workspace_bytes, void* stream) {
  return cs_select_topk_shard(prob, seg_offsets, uniform_T, n_bags, labels, tiles_per_pos, topk_neg, 0, 0,
                              sel_idx_out, sel_label_out, sel_offsets_out, capacity, workspace,
                              workspace_bytes, stream);
}

int cs_select_topk_shard(const float* prob, const int64_t* seg_offsets, int64_t uniform_T, int n_bags,
                         const int32_t* labels, int32_t tiles_per_pos, int32_t topk_neg,
                         int64_t global_tile_offset, int64_t global_total_tiles,
                         int32_t* sel_idx_out, uint8_t* sel_label_out, int64_t* sel_offsets_out,
                         int64_t capacity, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_segs("cs_select_topk", prob, seg_offsets, uniform_T, n_bags);
  CS_REQUIRE(global_tile_offset >= 0 && global_total_tiles >= 0,
             "cs_select_topk_shard: global offset / total must be >= 0");
  CS_REQUIRE(workspace != nullptr && workspace_bytes >= cs_select_workspace_bytes(n_bags),
             "cs_select_topk: workspace too small (need %lld bytes)",
             (long long)cs_select_workspace_bytes(n_bags));
  if (rc != CS_OK) return rc;
  CS_REQUIRE(labels && sel_idx_out && sel_label_out && sel_offsets_out,
             "cs_select_topk: NULL pointer");
  CS_REQUIRE(tiles_per_pos >= 0 && topk_neg >= 0 && capacity >= 0,
             "cs_select_topk: tiles_per_pos, topk_neg and capacity must be >= 0");
  cudaStream_t st = cs::as_stream(stream);
  Segs segs{seg_offsets, uniform_T, n_bags, global_tile_offset, global_total_tiles};
  int32_t* fb_count = static_cast<int32_t*>(workspace);
  int32_t* fb_list = fb_count + 64;
  int32_t* ticket = fb_count + 1;
  // workspace: [0] declined-bag count, [1] ticket, bytes 8..255 look-back words, 256.. declined list
  const int off_blocks = cs::ceil_div(n_bags, 1024);
  const bool small_grid = off_blocks <= kLookbackMaxBlocks && off_blocks <= cs::num_sms();
  const bool recount = g_offsets_mode == 0 && small_grid && seg_offsets == nullptr;
  const bool lookback = !recount && g_offsets_mode <= 1 && small_grid && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0;
  if (!recount) CS_CUDA(cudaMemsetAsync(fb_count, 0, lookback ? 256 : 2 * sizeof(int32_t), st));
  if (recount) {
    select_offsets_recount_kernel<<<off_blocks, 1024, 0, st>>>(segs, labels, tiles_per_pos, topk_neg,
                                                              sel_offsets_out, fb_count);
    CS_LAUNCH_CHECK();
  } else if (lookback) {
    select_offsets_lookback_kernel<<<off_blocks, 1024, 0, st>>>(
        segs, labels, tiles_per_pos, topk_neg, sel_offsets_out,
        reinterpret_cast<unsigned long long*>(fb_count + 2));
    CS_LAUNCH_CHECK();
  } else {
    static bool attr_done[64] = {false};
    int dev = 0;
    CS_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_done[dev]) {
      CS_CUDA(cudaFuncSetAttribute(select_offsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kOffChunk * (int)sizeof(int32_t)));
      if (dev < 64) attr_done[dev] = true;
    }
    select_offsets_kernel<<<off_blocks, 1024, kOffChunk * sizeof(int32_t), st>>>(
        segs, labels, tiles_per_pos, topk_neg, sel_offsets_out, ticket);
    CS_LAUNCH_CHECK();
  }
  EmitArgs ea{};
  ea.labels = labels;
  ea.tiles_per_pos = tiles_per_pos;
  ea.topk_neg = topk_neg;
  ea.idx_out = sel_idx_out;
  ea.label_out = sel_label_out;
  ea.out_offsets = sel_offsets_out;
  ea.capacity = capacity;
  ea.small_n_cols32 = g_cols32 ? 1 : 0;
  // Fast paths first: a register-resident CTA per bag (select_reg.cu) for bags of up to 4093
  // instances (with CELLSEG_SELECT_WARP=1 a warp per bag, select_warp.cu, up to 3069),
  // shared-memory staged (select_fast.cu) beyond; bags they decline are listed and ordered exactly
  bool handled = false, ordered = false;
  if (!g_disable_fast) {
    if (!g_staged_fast) {
      rc = launch_select_warp(segs, prob, ea, uniform_T, fb_count, fb_list, st, &handled);
      if (rc != CS_OK) return rc;
      if (!handled) rc = launch_select_reg(segs, prob, ea, uniform_T, fb_count, fb_list, st, &handled);
      if (rc != CS_OK) return rc;
      // every thread of these kernels waits for the offsets kernel before it exits, so a clean-up
      // pass that is their programmatic dependent is ordered behind the offsets too
      ordered = handled;
    }
    if (!handled) {
      rc = launch_select_fast(segs, prob, ea, uniform_T, fb_count, fb_list, st, &handled);
      if (rc != CS_OK) return rc;
    }
  }
  if (handled) {
    ea.fb_count = fb_count;
    ea.fb_list = fb_list;
    ea.sort_pdl = ordered ? 1 : 0;
  }
  return launch_sort<kSelect>(segs, prob, ea, uniform_T, st);
}

int64_t cs_select_workspace_bytes(int n_bags) { return n_bags > 0 ? 256 + 4 * (int64_t)n_bags : 0; }

int cs_rank_threshold(const float* prob, const int64_t* seg_offsets, int64_t uniform_T, int n_bags,
                      float threshold, int32_t* sel_idx_out, float* sel_prob_out,
                      int64_t* sel_offsets_out, int64_t capacity, void* stream) {
  int rc = check_segs("cs_rank_threshold", prob, seg_offsets, uniform_T, n_bags);
  if (rc != CS_OK) return rc;
  CS_REQUIRE(sel_idx_out && sel_offsets_out, "cs_rank_threshold: NULL pointer");
  CS_REQUIRE(capacity >= 0, "cs_rank_threshold: capacity < 0");
  cudaStream_t st = cs::as_stream(stream);
  Segs segs{seg_offsets, uniform_T, n_bags, 0, 0};
  rank_count_kernel<<<n_bags, 256, 0, st>>>(segs, prob, threshold, sel_offsets_out);
  CS_LAUNCH_CHECK();
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(sel_offsets_out, n_bags);
  CS_LAUNCH_CHECK();
  EmitArgs ea{};
  ea.thr = threshold;
  ea.idx_out = sel_idx_out;
  ea.prob_out = sel_prob_out;
  ea.out_offsets = sel_offsets_out;
  ea.capacity = capacity;
  if (!g_disable_fast) {
    rank_fast_kernel<<<n_bags, 128, 0, st>>>(segs, prob, ea);
    CS_LAUNCH_CHECK();
    ea.rank_fast_cap = kRankCap;
  }
  return launch_sort<kRank>(segs, prob, ea, uniform_T, st);
}

}  // extern "C"
