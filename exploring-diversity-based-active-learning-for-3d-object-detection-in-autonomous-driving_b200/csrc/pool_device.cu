// pool_device.cu -- device-pointer variants of the input entry points, for callers whose pool arrays already live in
// HBM (torch.ops.adsel.*, adsel_b200/torch_ops.py): no host round trip of the inputs.
//
//   adsel_set_pool_device        = adsel_set_pool (api.cu) with the validation, the bounding box, the group sizes and the
//                                  frames-by-group table computed by kernels; only a 64-byte summary comes back to the host
//   adsel_bev_set_points_device  = adsel_bev_set_points with a device-to-device copy of the points
//   adsel_copy_selected_device   = the selected indices of the last adsel_select, copied device-to-device
//
// The frames-by-group table must list every group's frames in ascending index order (the counting sort of
// adsel_set_pool): a stable radix sort of (group id, frame index) pairs -- cub::DeviceRadixSort, a library call on this
// set-up step (the toolkit's header-only CUB), like the exclusive scan of the group sizes.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <vector>

#include "common.cuh"

namespace adsel {

struct PoolSummary {
    int max_group_id, bad_group, bad_boxes, bad_loc, max_group, contiguous;
    int pad[2];
    unsigned long long box[4];   // ordered keys of xmin, ymin, xmax, ymax
};

__device__ __forceinline__ unsigned long long ordered_key(double v) {   // monotone map double -> u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static double from_ordered_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

__global__ void pool_stats_kernel(const double2 *__restrict__ loc, const int32_t *__restrict__ group_id,
                                  const int32_t *__restrict__ n_boxes, int n, PoolSummary *s) {
    int mg = -1, bg = 0, bb = 0, bl = 0;
    unsigned long long lo_x = ~0ull, lo_y = ~0ull, hi_x = 0ull, hi_y = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int g = group_id[i];
        mg = max(mg, g);
        bg |= g < 0;
        bb |= n_boxes[i] < 0;
        const double2 p = loc[i];
        bl |= !(p.x == p.x && p.y == p.y);
        const unsigned long long kx = ordered_key(p.x), ky = ordered_key(p.y);
        lo_x = min(lo_x, kx);
        hi_x = max(hi_x, kx);
        lo_y = min(lo_y, ky);
        hi_y = max(hi_y, ky);
    }
    atomicMax(&s->max_group_id, mg);
    if (bg) atomicExch(&s->bad_group, 1);
    if (bb) atomicExch(&s->bad_boxes, 1);
    if (bl) atomicExch(&s->bad_loc, 1);
    atomicMin(&s->box[0], lo_x);
    atomicMin(&s->box[1], lo_y);
    atomicMax(&s->box[2], hi_x);
    atomicMax(&s->box[3], hi_y);
}

__global__ void group_count_kernel(const int32_t *__restrict__ group_id, int n, int32_t *counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[group_id[i]], 1);
}

__global__ void iota_kernel(int32_t *p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void group_summary_kernel(const int32_t *__restrict__ counts, int n_groups, const int32_t *__restrict__ members, int n,
                                     PoolSummary *s) {
    int mx = 0, broken = 0;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x) mx = max(mx, counts[g]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) broken |= members[i] != i;
    atomicMax(&s->max_group, mx);
    if (broken) atomicExch(&s->contiguous, 0);
}

void update_shard(adsel_handle *h);

int set_pool_device(adsel_handle *h, int32_t n, const double *d_loc_xy, const int32_t *d_group_id, const int32_t *d_n_boxes) {
    ADSEL_REQUIRE(n > 0 && d_loc_xy && d_group_id && d_n_boxes, "adsel_set_pool_device: null/empty argument");
    ADSEL_TRY(h->loc.reserve(n));
    ADSEL_TRY(h->group_id.reserve(n));
    ADSEL_TRY(h->n_boxes.reserve(n));
    ADSEL_TRY(h->group_members.reserve(n));
    ADSEL_TRY(h->d_result.reserve(16));
    ADSEL_CUDA(cudaMemcpyAsync(h->loc.p, d_loc_xy, sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, h->stream));
    ADSEL_CUDA(cudaMemcpyAsync(h->group_id.p, d_group_id, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, h->stream));
    ADSEL_CUDA(cudaMemcpyAsync(h->n_boxes.p, d_n_boxes, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, h->stream));
    PoolSummary init;
    memset(&init, 0, sizeof(init));
    init.max_group_id = -1;
    init.contiguous = 1;
    init.box[0] = init.box[1] = ~0ull;
    PoolSummary *ds = reinterpret_cast<PoolSummary *>(h->d_result.p);
    static_assert(sizeof(PoolSummary) <= 16 * sizeof(double), "summary fits the result block");
    ADSEL_CUDA(cudaMemcpyAsync(ds, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    pool_stats_kernel<<<h->num_sms * 2, 256, 0, h->stream>>>(h->loc.p, h->group_id.p, h->n_boxes.p, n, ds);
    PoolSummary s;
    ADSEL_CUDA(cudaMemcpyAsync(&s, ds, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    ADSEL_REQUIRE(!s.bad_group, "adsel_set_pool: negative group id");
    ADSEL_REQUIRE(!s.bad_boxes, "adsel_set_pool: negative box count");
    ADSEL_REQUIRE(!s.bad_loc, "adsel_set_pool: NaN ego location");
    const int n_groups = s.max_group_id + 1;
    for (int q = 0; q < 4; ++q) h->bbox[q] = from_ordered_key(s.box[q]);
    // group sizes -> offsets; frames by group in ascending index order (stable sort of (group, index))
    ADSEL_TRY(h->group_ptr.reserve(n_groups + 1));
    DevBuf<int32_t> counts, iota, keys_out;
    DevBuf<unsigned char> tmp;
    ADSEL_TRY(counts.reserve(n_groups + 1));
    ADSEL_TRY(iota.reserve(n));
    ADSEL_TRY(keys_out.reserve(n));
    ADSEL_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(int32_t) * (n_groups + 1), h->stream));
    group_count_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->group_id.p, n, counts.p);
    iota_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(iota.p, n);
    size_t scan_bytes = 0, sort_bytes = 0;
    ADSEL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts.p, h->group_ptr.p, n_groups + 1, h->stream));
    int bits = 1;
    while ((1ll << bits) < n_groups) ++bits;
    ADSEL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, h->group_id.p, keys_out.p, iota.p, h->group_members.p, n, 0, bits, h->stream));
    ADSEL_TRY(tmp.reserve(scan_bytes > sort_bytes ? scan_bytes : sort_bytes));
    ADSEL_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, counts.p, h->group_ptr.p, n_groups + 1, h->stream));
    ADSEL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, sort_bytes, h->group_id.p, keys_out.p, iota.p, h->group_members.p, n, 0, bits, h->stream));
    group_summary_kernel<<<h->num_sms * 2, 256, 0, h->stream>>>(counts.p, n_groups, h->group_members.p, n, ds);
    ADSEL_CUDA(cudaMemcpyAsync(&s, ds, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    h->launches += 7;
    counts.release();
    iota.release();
    keys_out.release();
    tmp.release();
    h->max_group = s.max_group;
    h->groups_contiguous = s.contiguous != 0;
    h->n = n;
    h->n_groups = n_groups;
    update_shard(h);
    h->graph_ready = false;
    h->lists_mode = -1;
    return ADSEL_OK;
}

int bev_layout(adsel_handle *h, int32_t n_frames, const int64_t *frame_off_host);
int greedy_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                  int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);

__global__ void index_range_kernel(const int32_t *__restrict__ idx, int m, int n, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m && (idx[i] < 0 || idx[i] >= n)) atomicExch(bad, 1);
}

}  // namespace adsel

using namespace adsel;

extern "C" {

int adsel_set_pool_device(adsel_handle *h, int32_t n, const double *d_loc_xy, const int32_t *d_group_id, const int32_t *d_n_boxes) {
    ADSEL_REQUIRE(h, "adsel_set_pool_device: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return set_pool_device(h, n, d_loc_xy, d_group_id, d_n_boxes);
}

int adsel_bev_set_points_device(adsel_handle *h, const float *d_points, const int64_t *frame_off_host, int32_t n_frames) {
    ADSEL_REQUIRE(h && d_points, "adsel_bev_set_points_device: null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_TRY(bev_layout(h, n_frames, frame_off_host));
    ADSEL_CUDA(cudaMemcpyAsync(h->bev_pts.p, d_points, sizeof(float) * 5 * (size_t)h->bev_points, cudaMemcpyDeviceToDevice, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    return ADSEL_OK;
}

int adsel_select_device(adsel_handle *h, const adsel_select_params *p, const int32_t *d_sampled, int32_t n_sampled,
                        int32_t first_index, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h && out_m && h->n > 0, "adsel_select_device: null argument / no pool");
    ADSEL_CUDA(cudaSetDevice(h->device));
    if (n_sampled > 0) {
        ADSEL_REQUIRE(d_sampled, "adsel_select_device: null sampled list");
        ADSEL_TRY(h->d_result.reserve(16));
        int *bad = reinterpret_cast<int *>(h->d_result.p), host_bad = 0;
        ADSEL_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
        index_range_kernel<<<(n_sampled + 255) / 256, 256, 0, h->stream>>>(d_sampled, n_sampled, h->n, bad);
        ADSEL_CUDA(cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ADSEL_CUDA(cudaStreamSynchronize(h->stream));
        ADSEL_REQUIRE(!host_bad, "adsel_select: sampled index out of range");
    }
    std::vector<int32_t> scratch((size_t)h->n);     // the ABI of adsel_select returns the list on the host as well
    h->sampled_is_device = true;
    const int status = greedy_select(h, p, d_sampled, n_sampled, first_index, scratch.data(), h->n, out_m, out_cost);
    h->sampled_is_device = false;
    return status;
}

int adsel_copy_selected_device(adsel_handle *h, int32_t *d_dst, int32_t m) {
    ADSEL_REQUIRE(h && (m == 0 || d_dst) && m >= 0 && (size_t)m <= h->d_out.cap, "adsel_copy_selected_device: bad argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    if (m > 0) ADSEL_CUDA(cudaMemcpyAsync(d_dst, h->d_out.p, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    return ADSEL_OK;
}

int adsel_bev_descriptors_device(adsel_handle *h, int32_t **d_counts, uint16_t **d_levels, int32_t **d_norm2, int32_t *frames, int32_t *cells) {
    ADSEL_REQUIRE(h && h->bev_cells > 0, "adsel_bev_descriptors_device: call adsel_bev_build first");
    if (d_counts) *d_counts = h->bev_counts.p;
    if (d_levels) *d_levels = h->desc_levels;
    if (d_norm2) *d_norm2 = h->desc_norm2;
    if (frames) *frames = h->bev_frames;
    if (cells) *cells = h->bev_cells;
    return ADSEL_OK;
}

}  // extern "C"
