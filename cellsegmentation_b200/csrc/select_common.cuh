// Types shared by the exact per-bag sort (select_topk.cu) and the register-resident fast path
// (select_fast.cu).
#pragma once
#include "common.cuh"

namespace cs {

struct Segs {
  const int64_t* offsets;  // nullptr -> uniform
  int64_t uniform_T;
  int n_bags;
  // The literal predicate groups[i] != groups[(i + k) % N] looks at GLOBAL positions.  When the
  // bags are one shard of a larger set, g_off is the global index of this shard's first tile and
  // g_total the global tile count N (0: the shard is the whole set).
  int64_t g_off, g_total;
  __device__ __forceinline__ int64_t gstart(int b) const { return start(b) + g_off; }
  __device__ __forceinline__ int64_t gtotal() const { return g_total > 0 ? g_total : total(); }
  __device__ __forceinline__ int64_t start(int b) const {
    return offsets ? offsets[b] : (int64_t)b * uniform_T;
  }
  __device__ __forceinline__ int64_t total() const {
    return offsets ? offsets[n_bags] : (int64_t)n_bags * uniform_T;
  }
};

__device__ __forceinline__ int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Kept ranks of one bag under the literal predicate: [a1,b1) U [a2,b2).
struct Kept {
  int a1, b1, a2, b2;
  __device__ __forceinline__ int count() const { return (b1 - a1) + (b2 - a2); }
};

__device__ __forceinline__ Kept kept_ranges(int64_t s, int64_t T, int64_t N, int64_t k) {
  Kept r{0, 0, 0, 0};
  if (T <= 0 || N <= 0) return r;
  int64_t kp = k < N ? k : k % N;   // (the 64-bit division is ~100 instructions; k < N almost always)
  if (s + T + kp <= N) {            // no wrap-around reaches this bag: the top min(k', T) ranks
    const int t = (int)T, c = kp < T ? (int)kp : t;
    r.a1 = t - c; r.b1 = t; r.a2 = t; r.b2 = t;
    return r;
  }
  int64_t J0 = N - kp - s;
  int64_t a1 = T - kp > 0 ? T - kp : 0;
  int64_t b1 = T < J0 ? T : J0;
  int64_t a2 = J0 > 0 ? J0 : 0;
  int64_t b2 = T < N - kp ? T : N - kp;
  if (b1 < a1) b1 = a1;
  if (b2 < a2) b2 = a2;
  r.a1 = (int)a1; r.b1 = (int)b1; r.a2 = (int)a2; r.b2 = (int)b2;
  return r;
}

__device__ __forceinline__ int64_t bag_k(const int32_t* labels, int b, int32_t tiles_per_pos,
                                         int32_t topk_neg) {
  int32_t c = labels[b];
  return c == 0 ? (int64_t)topk_neg : (int64_t)c * (int64_t)tiles_per_pos;
}

struct EmitArgs {
  const int32_t* labels;
  int32_t tiles_per_pos, topk_neg;
  float thr;
  int32_t* idx_out;
  uint8_t* label_out;
  float* prob_out;
  const int64_t* out_offsets;  // [n_bags+1]
  int64_t capacity;
  const int32_t* fb_count;      // exact kernel after the fast path: number of declined bags ...
  const int32_t* fb_list;       // ... and their indices (nullptr: process every bag)
  int sort_pdl;                 // host side: the clean-up pass may be launched as a programmatic dependent
  int small_n_cols32;           // select_reg.cu: threshold from 32 column maxima when a bag keeps <= 16
  int rank_fast_cap;            // kRank: bags with <= this many kept entries were emitted by the
                                // fast kernel and are skipped by the exact one (0: none)
};


int launch_select_fast(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                       int32_t* fb_count, int32_t* fb_list, cudaStream_t st, bool* handled);
int launch_select_warp(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                       int32_t* fb_count, int32_t* fb_list, cudaStream_t st, bool* handled);
int launch_select_reg(const Segs& segs, const float* prob, const EmitArgs& ea, int64_t max_T,
                      int32_t* fb_count, int32_t* fb_list, cudaStream_t st, bool* handled);

}  // namespace cs
