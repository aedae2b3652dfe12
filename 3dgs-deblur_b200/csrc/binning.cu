// Tile binning: inclusive scan of tiles-per-Gaussian, (tile | depth) key emission, stable radix sort,
// per-tile [start, end) ranges.  Semantics of /root/reference/gsplat/gsplat/utils.py:106-182 and
// /root/reference/gsplat/gsplat/cuda/csrc/forward.cu:116-180; the scan and the sort are cub device
// primitives restricted to the key bits that can be set instead of a generic 64-bit torch.sort + gather.
#include <cub/cub.cuh>

#include "common.cuh"

namespace b200 {

// One warp per Gaussian: lanes stride over the Gaussian's tile rectangle, so a screen-filling splat
// (thousands of tiles) is emitted with coalesced 8-byte / 4-byte stores instead of one thread looping.
// Slots reserved by the scan but not covered by the int-radius bbox (the reference's "phantom"
// entries, forward.cu:90-104 vs :130-135) are written as key 0 / id 0, which is what the reference's
// zero-initialised buffers (bindings.cu:381-384) contain there.
__global__ void __launch_bounds__(256) map_intersects_kernel(int n, const float2 *__restrict__ xys,
                                                             const float *__restrict__ depths,
                                                             const int32_t *__restrict__ radii,
                                                             const int32_t *__restrict__ cum, int tbx, int tby,
                                                             float bw, int num_intersects,
                                                             int64_t *__restrict__ isect_ids,
                                                             int32_t *__restrict__ gaussian_ids) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int idx = warp;
    const int start = idx == 0 ? 0 : cum[idx - 1];
    const int end = min(cum[idx], num_intersects);
    int emitted = 0;
    const int r = radii[idx];
    if (r > 0) {
        const float2 c = xys[idx];
        int x0, y0, x1, y1;
        tile_bbox(c.x, c.y, (float)r, tbx, tby, bw, x0, y0, x1, y1);
        const int w = x1 - x0;
        emitted = max(0, w * (y1 - y0));
        emitted = min(emitted, max(0, end - start));
        const int64_t depth_id = (int64_t)__float_as_int(depths[idx]);
        for (int k = lane; k < emitted; k += 32) {
            const int ty = y0 + k / w, tx = x0 + k % w;
            const int64_t tile = (int64_t)ty * tbx + tx;
            isect_ids[start + k] = (tile << 32) | depth_id;
            gaussian_ids[start + k] = idx;
        }
    }
    for (int k = start + emitted + lane; k < end; k += 32) {
        isect_ids[k] = 0;
        gaussian_ids[k] = 0;
    }
}

__global__ void __launch_bounds__(256) tile_bin_edges_kernel(int m, const int64_t *__restrict__ sorted,
                                                             int2 *__restrict__ bins) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int cur = (int)(sorted[i] >> 32);
    if (i == 0) bins[cur].x = 0;
    if (i == m - 1) bins[cur].y = m;
    if (i == 0) return;
    const int prev = (int)(sorted[i - 1] >> 32);
    if (prev != cur) {
        bins[prev].y = i;
        bins[cur].x = i;
    }
}

static int key_end_bit(int num_tiles) {
    int bits = 0;
    while ((1ll << bits) < (long long)num_tiles) ++bits;
    return 32 + bits;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_scan_temp_bytes(int num_points) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr, num_points);
    return bytes;
}

extern "C" int b200_cumulative_intersects(int num_points, const int32_t *num_tiles_hit, int32_t *cum_tiles_hit,
                                          void *temp, size_t temp_bytes, int32_t *total_host_pinned, void *stream) {
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(num_tiles_hit && cum_tiles_hit && temp, "null pointer");
    cudaStream_t st = as_stream(stream);
    B200_CUDA(cub::DeviceScan::InclusiveSum(temp, temp_bytes, num_tiles_hit, cum_tiles_hit, num_points, st));
    count_launch(2);  // decoupled look-back scan: init + scan kernels
    if (total_host_pinned)
        B200_CUDA(cudaMemcpyAsync(total_host_pinned, cum_tiles_hit + (num_points - 1), sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, st));
    return B200_OK;
}

extern "C" int b200_map_gaussian_to_intersects(int num_points, int num_intersects, const float *xys,
                                               const float *depths, const int32_t *radii,
                                               const int32_t *cum_tiles_hit, unsigned tiles_x, unsigned tiles_y,
                                               unsigned block_width, int64_t *isect_ids, int32_t *gaussian_ids,
                                               void *stream) {
    B200_REQUIRE(num_points >= 1 && num_intersects >= 0, "bad sizes");
    if (num_intersects == 0) return B200_OK;
    B200_REQUIRE(xys && depths && radii && cum_tiles_hit && isect_ids && gaussian_ids, "null pointer");
    B200_REQUIRE(block_width > 1 && block_width <= 16, "block_width must be between 2 and 16");
    const long long threads = 32ll * num_points;
    const int blocks = (int)((threads + 255) / 256);
    map_intersects_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
        num_points, reinterpret_cast<const float2 *>(xys), depths, radii, cum_tiles_hit, (int)tiles_x, (int)tiles_y,
        (float)block_width, num_intersects, isect_ids, gaussian_ids);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" size_t b200_sort_temp_bytes(int num_intersects) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs((void *)nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, num_intersects, 0, 64);
    return bytes;
}

extern "C" int b200_sort_intersects(int num_intersects, int num_tiles, const int64_t *isect_ids,
                                    const int32_t *gaussian_ids, int64_t *isect_ids_sorted,
                                    int32_t *gaussian_ids_sorted, void *temp, size_t temp_bytes, void *stream) {
    B200_REQUIRE(num_intersects >= 0 && num_tiles >= 1, "bad sizes");
    if (num_intersects == 0) return B200_OK;
    B200_REQUIRE(isect_ids && gaussian_ids && isect_ids_sorted && gaussian_ids_sorted && temp, "null pointer");
    // keys are non-negative, so unsigned order == signed order; only [0, 32 + log2(tiles)) bits vary
    B200_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, reinterpret_cast<const uint64_t *>(isect_ids),
                                              reinterpret_cast<uint64_t *>(isect_ids_sorted), gaussian_ids,
                                              gaussian_ids_sorted, num_intersects, 0, key_end_bit(num_tiles),
                                              as_stream(stream)));
    count_launch(1 + (key_end_bit(num_tiles) + 7) / 8);  // onesweep: histogram + one pass per 8-bit digit
    return B200_OK;
}

extern "C" int b200_get_tile_bin_edges(int num_intersects, int num_tiles, const int64_t *isect_ids_sorted,
                                       int32_t *tile_bins, void *stream) {
    B200_REQUIRE(num_intersects >= 0 && num_tiles >= 1, "bad sizes");
    B200_REQUIRE(tile_bins, "null pointer");
    cudaStream_t st = as_stream(stream);
    B200_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)num_tiles, st));
    if (num_intersects == 0) return B200_OK;
    B200_REQUIRE(isect_ids_sorted, "null pointer");
    tile_bin_edges_kernel<<<ceil_div(num_intersects, 256), 256, 0, st>>>(num_intersects, isect_ids_sorted,
                                                                        reinterpret_cast<int2 *>(tile_bins));
    B200_LAUNCH_CHECK();
    return B200_OK;
}
