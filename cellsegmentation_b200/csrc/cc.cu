// N1: connected-component clean-up of the refined masks on the GPU.
//
// Reference: remove_small_regions (utils/image_processing.py:14-17) =
//   skimage.morphology.remove_small_objects(img, min_size)        (scikit-image 0.19.0)
//   skimage.morphology.remove_small_holes(img, area_threshold)
// both with the default connectivity 1 (4 neighbours): label the components, drop those with
// size < min_size; holes = the same on the complement (background components with
// size < area_threshold become foreground, whether or not they touch the border).
//
// Run-based labelling in shared memory.  One CTA per image (persistent over the batch):
//   0. the mask becomes a bit image (warp ballots of coalesced byte loads): 12 KB for 299 x 299
//   1. every row is cut into runs of set bits (start / end masks, popc + ffs); a block scan of
//      the per-row counts numbers the runs (~2 000 per LYSTO-like mask, ~9 000 for pure noise)
//   2. a run is united with the runs of the row above that it overlaps in x (4-connectivity):
//      union-find on RUNS with atomicMin linking, all in shared memory
//   3. run lengths are added up at the roots; runs of components below the threshold are
//      flipped in the bit image
// once for the objects and once, on the inverted bits, for the holes; then the bits go back to
// bytes.  Integer work, bit-exact against the scipy restatement in oracle/masks.py.  Images
// with more than kMaxRuns runs in a pass (pathological speckle) or wider than kMaxW take the
// pixel-level union-find in global memory that was the first version of this kernel (workspace:
// two int32 per pixel per resident CTA).
#include "common.cuh"

namespace {

// Two table sizes.  LYSTO-like masks have ~2 000 runs: the first launch gives every image a
// 4 096-run table (63 KB of shared memory, 512 threads: three CTAs per SM); the few images that
// overflow it (pure per-pixel noise at 299 x 299 has ~9 000 runs) are listed and redone by a
// second launch with 16 384-run tables (196 KB, one CTA per SM), which also owns the
// pixel-level fallback.  With one table size for all (round 1) every image paid for the noise
// case: one CTA per SM, 1.3 M masks/s.
constexpr int kSmallRuns = 4096, kSmallThreads = 512;
constexpr int kBigRuns = 16384, kBigThreads = 1024;
constexpr int kMaxW = 320;              // 10 words per row
constexpr int kMaxH = 320;
constexpr int kWords = kMaxW / 32;

template <int kMaxRuns, int kThreads>
struct SmemT {
  uint32_t bits[kMaxH * kWords];        // row-major bit image
  uint32_t row_first[kMaxH + 1];        // first run id of every row (exclusive scan of counts)
  uint16_t run_s[kMaxRuns], run_e[kMaxRuns];   // inclusive x range of a run
  uint32_t parent[kMaxRuns];
  uint32_t csize[kMaxRuns];
  uint32_t scan_tmp[kThreads / 32];
};

// Root of x with path halving.  Links always point to smaller ids and only roots are re-linked
// (atomicMin in unite_s), so overwriting a non-root's parent with its grandparent races benignly.
__device__ __forceinline__ uint32_t find_root_s(volatile uint32_t* parent, uint32_t x) {
  uint32_t p = parent[x];
  while (p != x) {
    const uint32_t gp = parent[p];
    if (gp != p) parent[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ uint32_t find_root_ro(const volatile uint32_t* parent, uint32_t x) {
  uint32_t p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}
__device__ __forceinline__ void unite_s(uint32_t* parent, uint32_t a, uint32_t b) {
  while (true) {
    a = find_root_s(parent, a);
    b = find_root_s(parent, b);
    if (a == b) return;
    if (a < b) { uint32_t t = a; a = b; b = t; }   // link the larger root under the smaller
    const uint32_t old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}

// word j of row r of the (optionally inverted) image; bits >= W are always 0
template <class Smem>
__device__ __forceinline__ uint32_t row_word(const Smem& sm, int r, int j, int W, bool invert) {
  if (j < 0 || j >= kWords) return 0u;
  uint32_t w = sm.bits[r * kWords + j];
  if (invert) {
    w = ~w;
    const int valid = W - 32 * j;
    if (valid <= 0) w = 0u;
    else if (valid < 32) w &= (1u << valid) - 1u;
  }
  return w;
}

// Removes (flips) the 4-connected components of set bits (of cleared bits when invert) whose
// size is < thresh.  Returns false (nothing changed) if the image has too many runs.
template <int kMaxRuns, int kThreads>
__device__ bool prune_runs(SmemT<kMaxRuns, kThreads>& sm, int H, int W, bool invert, uint32_t thresh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1a. runs per row
  for (int r0 = 0; r0 < H; r0 += kThreads) {
    const int r = r0 + tid;
    uint32_t cnt = 0;
    if (r < H) {
      uint32_t prev_msb = 0;
      for (int j = 0; j < kWords; ++j) {
        const uint32_t w = row_word(sm, r, j, W, invert);
        cnt += __popc(w & ~((w << 1) | prev_msb));
        prev_msb = w >> 31;
      }
    }
    // block exclusive scan of cnt, continuing from the total of the previous chunk of rows
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) sm.scan_tmp[warp] = x;
    __syncthreads();
    uint32_t base = r0 == 0 ? 0u : sm.row_first[r0];
    for (int w2 = 0; w2 < warp; ++w2) base += sm.scan_tmp[w2];
    if (r < H) sm.row_first[r] = base + x - cnt;
    if (tid == kThreads - 1) sm.row_first[min(r0 + kThreads, H)] = base + x;   // running total
    __syncthreads();
  }
  const uint32_t n_runs = sm.row_first[H];
  if (n_runs > (uint32_t)kMaxRuns) return false;
  // 1b. run extents: the k-th start bit of a row pairs with its k-th end bit
  for (int r = tid; r < H; r += kThreads) {
    uint32_t id_s = sm.row_first[r], id_e = id_s;
    uint32_t prev_msb = 0;
    for (int j = 0; j < kWords; ++j) {
      const uint32_t w = row_word(sm, r, j, W, invert);
      const uint32_t next_lsb = row_word(sm, r, j + 1, W, invert) & 1u;
      uint32_t starts = w & ~((w << 1) | prev_msb);
      uint32_t ends = w & ~((w >> 1) | (next_lsb << 31));
      prev_msb = w >> 31;
      while (starts) { const int b = __ffs(starts) - 1; starts &= starts - 1; sm.run_s[id_s++] = (uint16_t)(32 * j + b); }
      while (ends) { const int b = __ffs(ends) - 1; ends &= ends - 1; sm.run_e[id_e++] = (uint16_t)(32 * j + b); }
    }
  }
  for (uint32_t i = tid; i < n_runs; i += kThreads) { sm.parent[i] = i; sm.csize[i] = 0; }
  __syncthreads();
  // 2. unions with the overlapping runs of the row above.  A thread owns a CONTIGUOUS range of
  // runs: its row index and its cursor into the row above only move forward (two-pointer merge),
  // so the phase costs O(runs) instead of O(runs x runs per row).
  const uint32_t per = (n_runs + kThreads - 1) / kThreads;
  const uint32_t lo_i = min((uint32_t)tid * per, n_runs), hi_i = min(lo_i + per, n_runs);
  {
    int r = 0;
    uint32_t j = 0, j_end = 0;                                // cursor in row r - 1
    for (uint32_t i = lo_i; i < hi_i; ++i) {
      bool moved = i == lo_i;
      while (sm.row_first[r + 1] <= i) { ++r; moved = true; }  // row of run i
      if (r == 0) continue;                                   // row 0 has nothing above
      if (moved) { j = sm.row_first[r - 1]; j_end = sm.row_first[r]; }
      const uint32_t s = sm.run_s[i], e = sm.run_e[i];
      while (j < j_end && sm.run_e[j] < s) ++j;               // runs entirely to the left
      for (uint32_t k = j; k < j_end && sm.run_s[k] <= e; ++k) unite_s(sm.parent, i, k);
    }
  }
  __syncthreads();
  // 3. sizes at the roots, then flip the runs of small components
  // flatten: every run points straight at its root.  Read-only walks here -- a concurrent path
  // halving store could overwrite a finished entry with a non-root ancestor.
  for (uint32_t i = tid; i < n_runs; i += kThreads) {
    const uint32_t root = find_root_ro(sm.parent, i);
    if (root != i) sm.parent[i] = root;                    // (roots keep parent == self)
  }
  __syncthreads();
  // every thread adds up the lengths of its contiguous runs per root before touching the shared
  // counter (a mask with one giant component would otherwise serialise thousands of atomics)
  {
    uint32_t cur = 0xffffffffu, sum = 0;
    for (uint32_t i = lo_i; i < hi_i; ++i) {
      const uint32_t root = sm.parent[i];
      if (root != cur) {
        if (sum) atomicAdd(&sm.csize[cur], sum);
        cur = root;
        sum = 0;
      }
      sum += (uint32_t)(sm.run_e[i] - sm.run_s[i] + 1);
    }
    if (sum) atomicAdd(&sm.csize[cur], sum);
  }
  __syncthreads();
  {
    int r = 0;
    for (uint32_t i = lo_i; i < hi_i; ++i) {
      while (sm.row_first[r + 1] <= i) ++r;
      if (sm.csize[sm.parent[i]] >= thresh) continue;
      const int s = sm.run_s[i], e = sm.run_e[i];
      for (int j = s >> 5; j <= (e >> 5); ++j) {
        const int lo = max(s - 32 * j, 0), hi = min(e - 32 * j, 31);
        const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        if (invert) atomicOr(&sm.bits[r * kWords + j], m);
        else atomicAnd(&sm.bits[r * kWords + j], ~m);
      }
    }
  }
  __syncthreads();
  return true;
}

// ---- pixel-level fallback in global memory (first version of this kernel) ----------------------
__device__ __forceinline__ int find_root(volatile int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}
__device__ __forceinline__ void unite(int* parent, int a, int b) {
  while (true) {
    a = find_root(parent, a);
    b = find_root(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}
__device__ void prune_components(uint8_t* img, int H, int W, int target, int thresh, int* parent,
                                 int* size) {
  const int P = H * W;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    parent[p] = (img[p] != 0) == (target != 0) ? p : -1;
    size[p] = 0;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    if (parent[p] < 0) continue;
    const int x = p % W;
    if (x > 0 && parent[p - 1] >= 0) unite(parent, p, p - 1);
    if (p >= W && parent[p - W] >= 0) unite(parent, p, p - W);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    if (parent[p] < 0) continue;
    const int r = find_root(parent, p);
    atomicAdd(&size[r], 1);
    if (r != p) atomicExch(&parent[p], r);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int q = parent[p];
    if (q < 0) continue;
    const int r = find_root(parent, q);
    if (size[r] < thresh) img[p] = (uint8_t)(1 - target);
  }
  __syncthreads();
}

template <int kMaxRuns, int kThreads>
__device__ void load_bits(SmemT<kMaxRuns, kThreads>& sm, const uint8_t* img, int H, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kRowsPerIter = 4;   // 40 byte loads in flight per lane hide the L2 latency
  for (int r0 = warp * kRowsPerIter; r0 < H; r0 += (kThreads / 32) * kRowsPerIter) {
    uint8_t v[kRowsPerIter][kWords];
#pragma unroll
    for (int q = 0; q < kRowsPerIter; ++q) {
      const uint8_t* row = img + (size_t)(r0 + q) * W;
#pragma unroll
      for (int j = 0; j < kWords; ++j) {
        const int x = 32 * j + lane;
        v[q][j] = (r0 + q < H && x < W) ? __ldg(row + x) : (uint8_t)0;
      }
    }
#pragma unroll
    for (int q = 0; q < kRowsPerIter; ++q) {
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < kWords; ++j) {
        const uint32_t w = __ballot_sync(0xffffffffu, v[q][j] != 0);
        mine = lane == j ? w : mine;
      }
      if (r0 + q < H && lane < kWords) sm.bits[(r0 + q) * kWords + lane] = mine;
    }
  }
  __syncthreads();
}

template <int kMaxRuns, int kThreads>
__device__ void store_bits(const SmemT<kMaxRuns, kThreads>& sm, uint8_t* img, int H, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < H; r += kThreads / 32) {
    uint8_t* row = img + (size_t)r * W;
#pragma unroll
    for (int j = 0; j < kWords; ++j) {
      const int x = 32 * j + lane;
      if (x < W) row[x] = (uint8_t)((sm.bits[r * kWords + j] >> lane) & 1u);
    }
  }
  __syncthreads();
}

// FIRST = true: every image, small tables; images that overflow them (or do not fit the bit
// image at all) are appended to ovf_list.  FIRST = false: the listed images only, big tables,
// then the pixel-level path in global memory.
template <int kMaxRuns, int kThreads, bool FIRST>
__global__ void __launch_bounds__(kThreads)
remove_small_regions_kernel(uint8_t* mask, int n_bags, int H, int W, int min_object,
                            int hole_area, int* __restrict__ ovf_count, int* __restrict__ ovf_list,
                            int* __restrict__ ws) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  using Smem = SmemT<kMaxRuns, kThreads>;
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int P = H * W;
  int* parent = FIRST ? nullptr : ws + (size_t)blockIdx.x * 2 * P;
  int* size = FIRST ? nullptr : parent + P;
  const bool fits = W <= kMaxW && H <= kMaxH;
  const int n_items = FIRST ? n_bags : *ovf_count;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = FIRST ? item : ovf_list[item];
    uint8_t* img = mask + (size_t)b * P;
    bool done = false;
    if (fits) {
      load_bits(sm, img, H, W);
      bool ok = true;
      if (min_object > 0) ok = prune_runs(sm, H, W, false, (uint32_t)min_object);
      if (ok && hole_area > 0) ok = prune_runs(sm, H, W, true, (uint32_t)hole_area);
      if (ok) {
        store_bits(sm, img, H, W);
        done = true;
      }
      __syncthreads();
    }
    if (done) continue;
    if (FIRST) {   // the bytes are untouched: the second launch redoes this image
      if (threadIdx.x == 0) ovf_list[atomicAdd(ovf_count, 1)] = b;
    } else {       // pixel-level path on the untouched bytes
      for (int p = threadIdx.x; p < P; p += blockDim.x) img[p] = img[p] != 0 ? 1 : 0;
      __syncthreads();
      if (min_object > 0) prune_components(img, H, W, 1, min_object, parent, size);
      if (hole_area > 0) prune_components(img, H, W, 0, hole_area, parent, size);
    }
  }
}

int cc_grid(int n_bags) {
  int g = cs::num_sms();      // second launch: the big run tables take an SM's shared memory
  return n_bags < g ? n_bags : g;
}

}  // namespace

extern "C" {

int64_t cs_cc_workspace_bytes(int n_bags, int H, int W) {
  if (n_bags <= 0 || H <= 0 || W <= 0) return 0;
  // overflow counter + list, then two int32 per pixel per CTA of the second launch
  return 256 + 4 * (int64_t)n_bags + 256 + (int64_t)cc_grid(n_bags) * 2 * H * W * (int64_t)sizeof(int) + 256;
}

int cs_remove_small_regions(uint8_t* mask, int n_bags, int H, int W, int min_object_size,
                            int hole_area_threshold, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  CS_REQUIRE(mask != nullptr && workspace != nullptr, "cs_remove_small_regions: NULL pointer");
  CS_REQUIRE(n_bags > 0 && H > 0 && W > 0 && (int64_t)H * W < (1 << 30),
             "cs_remove_small_regions: bad geometry");
  CS_REQUIRE(min_object_size >= 0 && hole_area_threshold >= 0,
             "cs_remove_small_regions: thresholds must be >= 0");
  if (workspace_bytes < cs_cc_workspace_bytes(n_bags, H, W)) {
    cs::set_error("cs_remove_small_regions: workspace %lld B < required %lld B",
                  (long long)workspace_bytes, (long long)cs_cc_workspace_bytes(n_bags, H, W));
    return CS_ERR_WORKSPACE;
  }
  using SmallSmem = SmemT<kSmallRuns, kSmallThreads>;
  using BigSmem = SmemT<kBigRuns, kBigThreads>;
  auto small_k = remove_small_regions_kernel<kSmallRuns, kSmallThreads, true>;
  auto big_k = remove_small_regions_kernel<kBigRuns, kBigThreads, false>;
  static bool attr_done[64] = {false};
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    CS_CUDA(cudaFuncSetAttribute(small_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem)));
    CS_CUDA(cudaFuncSetAttribute(big_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem)));
    if (dev < 64) attr_done[dev] = true;
  }
  cudaStream_t st = cs::as_stream(stream);
  int* base = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  int* ovf_count = base;
  int* ovf_list = base + 64;
  int* ws = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(ovf_list + n_bags) + 255) & ~(uintptr_t)255);
  CS_CUDA(cudaMemsetAsync(ovf_count, 0, sizeof(int), st));
  const int per_sm = (int)((227 * 1024) / (sizeof(SmallSmem) + 1024));
  const int small_grid = n_bags < cs::num_sms() * per_sm ? n_bags : cs::num_sms() * per_sm;
  small_k<<<small_grid, kSmallThreads, sizeof(SmallSmem), st>>>(mask, n_bags, H, W, min_object_size,
                                                               hole_area_threshold, ovf_count, ovf_list, nullptr);
  CS_LAUNCH_CHECK();
  big_k<<<cc_grid(n_bags), kBigThreads, sizeof(BigSmem), st>>>(mask, n_bags, H, W, min_object_size,
                                                              hole_area_threshold, ovf_count, ovf_list, ws);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // extern "C"
