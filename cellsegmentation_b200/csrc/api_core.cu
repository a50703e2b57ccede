// Version, error text, device check and the host-side tile-grid helpers.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cs {
static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// SM count of the current device (queried once per device; 148 on a B200).  Every persistent
// grid is sized from it.
int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}
}  // namespace cs

extern "C" {

int cs_version(void) { return 0 * 10000 + 1 * 100 + 0; }

const char* cs_last_error(void) { return cs::last_error(); }

int cs_check_device(void) {
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CS_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    cs::set_error("device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev,
                  prop.name, prop.major, prop.minor);
    return CS_ERR_DEVICE;
  }
  return CS_OK;
}

// dataset/dataset.py:728-740
int cs_grid_count(int dim, int tile, int interval) { return cs::grid_count(dim, tile, interval); }

// dataset/dataset.py:718-742 (row-major over (x = row, y = col))
int cs_grid_cover_host(int c, int dim, int tile, int interval, int32_t* lo, int32_t* hi) {
  CS_REQUIRE(lo != nullptr && hi != nullptr, "cs_grid_cover_host: NULL pointer");
  const int cnt = cs::grid_count(dim, tile, interval);
  CS_REQUIRE(cnt > 0 && c >= 0 && c < dim, "cs_grid_cover_host: bad geometry dim=%d tile=%d interval=%d c=%d", dim,
             tile, interval, c);
  int l, h;
  cs::grid_cover(c, dim, tile, interval, cnt, &l, &h);
  *lo = l;
  *hi = h;
  return CS_OK;
}

int cs_grid_coords_host(int H, int W, int tile, int interval, int32_t* xy_host,
                        int64_t capacity_tiles) {
  CS_REQUIRE(xy_host != nullptr, "cs_grid_coords_host: xy_host is NULL");
  int ch = cs::grid_count(H, tile, interval), cw = cs::grid_count(W, tile, interval);
  CS_REQUIRE(ch > 0 && cw > 0, "cs_grid_coords_host: bad geometry H=%d W=%d tile=%d interval=%d",
             H, W, tile, interval);
  CS_REQUIRE((int64_t)ch * cw <= capacity_tiles,
             "cs_grid_coords_host: capacity %lld < %lld tiles", (long long)capacity_tiles,
             (long long)ch * cw);
  int64_t t = 0;
  for (int gy = 0; gy < ch; ++gy)
    for (int gx = 0; gx < cw; ++gx, ++t) {
      xy_host[2 * t] = cs::grid_coord(gy, H, tile, interval);
      xy_host[2 * t + 1] = cs::grid_coord(gx, W, tile, interval);
    }
  return CS_OK;
}

}  // extern "C"
