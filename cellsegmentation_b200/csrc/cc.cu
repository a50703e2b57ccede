// N1: connected-component clean-up of the refined masks on the GPU.
//
// Reference: remove_small_regions (utils/image_processing.py:14-17) =
//   skimage.morphology.remove_small_objects(img, min_size)        (scikit-image 0.19.0)
//   skimage.morphology.remove_small_holes(img, area_threshold)
// both with the default connectivity 1 (4 neighbours): label the components, drop those with
// size < min_size; holes = the same on the complement (background components with
// size < area_threshold become foreground, whether or not they touch the border).
//
// One CTA per image, persistent over the images of the batch.  Labels live in a per-CTA slice
// of the caller's workspace (two int32 per pixel: parent / size), which stays L2 resident
// (715 KB per 299x299 image).  Union-find with atomicMin linking (larger root -> smaller root),
// then path flattening, size histogram by atomicAdd on the roots, and the threshold pass.
// Integer work, bit-exact against the scipy restatement in oracle/masks.py.
#include "common.cuh"

namespace {

constexpr int kThreads = 1024;

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
    if (a < b) { int t = a; a = b; b = t; }   // link the larger root under the smaller
    int old = atomicMin(&parent[a], b);
    if (old == a) return;                     // a was still a root: linked
    a = old;                                  // somebody linked it first: retry from there
  }
}

// Removes 4-connected components of pixels equal to `target` whose size is < thresh by
// flipping them to 1 - target.
__device__ void prune_components(uint8_t* img, int H, int W, int target, int thresh, int* parent,
                                 int* size) {
  const int P = H * W;
  for (int p = threadIdx.x; p < P; p += kThreads) {
    parent[p] = (img[p] != 0) == (target != 0) ? p : -1;
    size[p] = 0;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += kThreads) {
    if (parent[p] < 0) continue;
    const int x = p % W;
    if (x > 0 && parent[p - 1] >= 0) unite(parent, p, p - 1);
    if (p >= W && parent[p - W] >= 0) unite(parent, p, p - W);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += kThreads) {
    if (parent[p] < 0) continue;
    const int r = find_root(parent, p);
    atomicAdd(&size[r], 1);
    // remember the root in place: roots keep parent[r] == r, others point straight at it
    if (r != p) atomicExch(&parent[p], r);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += kThreads) {
    const int q = parent[p];
    if (q < 0) continue;
    const int r = find_root(parent, q);
    if (size[r] < thresh) img[p] = (uint8_t)(1 - target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
remove_small_regions_kernel(uint8_t* mask, int n_bags, int H, int W, int min_object,
                            int hole_area, int* __restrict__ ws) {
  const int P = H * W;
  int* parent = ws + (size_t)blockIdx.x * 2 * P;
  int* size = parent + P;
  for (int b = blockIdx.x; b < n_bags; b += gridDim.x) {
    uint8_t* img = mask + (size_t)b * P;
    // masks are 0/1 by contract; normalise so that "1 - target" flips are well defined
    for (int p = threadIdx.x; p < P; p += kThreads) img[p] = img[p] != 0 ? 1 : 0;
    __syncthreads();
    if (min_object > 0) prune_components(img, H, W, 1, min_object, parent, size);
    if (hole_area > 0) prune_components(img, H, W, 0, hole_area, parent, size);
  }
}

int cc_grid(int n_bags) {
  int g = cs::kNumSMs * 2;
  return n_bags < g ? n_bags : g;
}

}  // namespace

extern "C" {

int64_t cs_cc_workspace_bytes(int n_bags, int H, int W) {
  if (n_bags <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)cc_grid(n_bags) * 2 * H * W * (int64_t)sizeof(int) + 256;
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
  int* ws = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  remove_small_regions_kernel<<<cc_grid(n_bags), kThreads, 0, cs::as_stream(stream)>>>(
      mask, n_bags, H, W, min_object_size, hole_area_threshold, ws);
  CS_LAUNCH_CHECK();
  return CS_OK;
}

}  // extern "C"
