// api.cu -- extern "C" surface of libadsel_b200.so (include/adsel_b200.h).
#include <vector>

#include "common.cuh"

namespace adsel {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

int begin_stage(adsel_handle *h) {
    ADSEL_CUDA(cudaEventRecord(h->ev[0], h->stream));
    return ADSEL_OK;
}
int end_stage(adsel_handle *h, int stage) {
    ADSEL_CUDA(cudaEventRecord(h->ev[1], h->stream));
    ADSEL_CUDA(cudaEventSynchronize(h->ev[1]));
    float ms = 0.f;
    ADSEL_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stage_ms[stage] = ms;
    return ADSEL_OK;
}

int build_knn_graph(adsel_handle *h, int32_t k);
int build_knn_graph_sharded(adsel_handle *h, int32_t k);
void update_shard(adsel_handle *h);
int knn_prepare(adsel_handle *h, int32_t k, bool own_tables);
int greedy_prepare(adsel_handle *h, const adsel_select_params *p, int32_t n_sampled, int32_t out_cap, bool sharded);
void comm_destroy(adsel_handle *h);
int greedy_select_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                                int32_t first_index, int32_t out_cap, bool replicate);
int greedy_select_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);

static int refresh_nnz(adsel_handle *h) {
    if (h->nnz >= 0) return ADSEL_OK;
    int32_t nnz32 = 0;
    ADSEL_CUDA(cudaMemcpyAsync(&nnz32, h->row_ptr.p + h->n, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    h->nnz = nnz32;
    return ADSEL_OK;
}
int geodesic_rows(adsel_handle *h, const int32_t *sources, int32_t m, double radius, double *out_host);
int geodesic_max_finite(adsel_handle *h, double *out);
int greedy_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                  int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);

int fused_feature_select(adsel_handle *h, const adsel_select_params *p, double lambda_f, int32_t seed_spatial_only,
                         const int32_t *sampled, int32_t n_sampled, int32_t first_index, int32_t *out_idx, int32_t out_cap,
                         int32_t *out_m, double *out_cost);
int bev_generate(adsel_handle *h, int32_t n_frames, const int64_t *frame_off_host, int32_t n_sweeps, uint64_t seed);
int bev_set_points(adsel_handle *h, const float *pts_host, const int64_t *frame_off_host, int32_t n_frames);
int bev_build(adsel_handle *h, const adsel_bev_params *p, int aggregate_atomics, int32_t frame_base, int32_t total_frames);

int pairdist_build(adsel_handle *h);
int pairdist_set_descriptors(adsel_handle *h, const uint16_t *levels_host, int32_t n, int32_t dim);
int pairdist_build_sharded(adsel_handle *h);
int feature_distance_build(adsel_handle *h, const float *feats_host, int32_t n, int32_t dim, int32_t p);
int dense_prepare(adsel_handle *h, int32_t n_sampled, int32_t out_cap);
int dense_select_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *n_boxes_host, const int32_t *sampled,
                               int32_t n_sampled, int32_t first_index, int32_t out_cap);
int dense_select_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);
int dense_select(adsel_handle *h, const adsel_select_params *p, const int32_t *n_boxes_host, const int32_t *sampled,
                 int32_t n_sampled, int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);

__global__ void exp_f64_kernel(const double *__restrict__ x, long long n, double *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = np_exp_f64(x[i]);
}

__global__ void exp_f32_kernel(const float *__restrict__ x, long long n, float *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = np_exp_f32(x[i]);
}

}  // namespace adsel

using namespace adsel;

extern "C" {

const char *adsel_version(void) { return "adsel_b200 0.1.0 (sm_100a)"; }
const char *adsel_last_error(void) { return g_last_error.c_str(); }

int adsel_device_count(int *count) {
    ADSEL_REQUIRE(count, "adsel_device_count: null argument");
    ADSEL_CUDA(cudaGetDeviceCount(count));
    return ADSEL_OK;
}

int adsel_create(int device, adsel_handle **out) {
    ADSEL_REQUIRE(out, "adsel_create: null argument");
    int count = 0;
    ADSEL_CUDA(cudaGetDeviceCount(&count));
    ADSEL_REQUIRE(device >= 0 && device < count, "adsel_create: no such CUDA device (this library has no CPU fallback)");
    ADSEL_CUDA(cudaSetDevice(device));
    adsel_handle *h = new adsel_handle();
    h->device = device;
    ADSEL_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    ADSEL_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    ADSEL_CUDA(cudaEventCreate(&h->ev[0]));
    ADSEL_CUDA(cudaEventCreate(&h->ev[1]));
    ADSEL_CUDA(cudaEventCreate(&h->timer_ev[0]));
    ADSEL_CUDA(cudaEventCreate(&h->timer_ev[1]));
    *out = h;
    return ADSEL_OK;
}

int adsel_destroy(adsel_handle *h) {
    if (!h) return ADSEL_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    comm_destroy(h);
    h->loc.release(); h->group_id.release(); h->n_boxes.release(); h->group_ptr.release(); h->group_members.release();
    h->knn_idx.release(); h->knn_dist.release(); h->grid_cell.release(); h->grid_idx.release(); h->grid_start.release(); h->grid_loc.release(); h->row_ptr.release(); h->col.release(); h->deg_own.release();
    h->deg_rev.release(); h->w.release();
    h->sc_dist.release(); h->sc_stamp.release(); h->sc_q0.release(); h->sc_q1.release(); h->sc_touched.release();
    h->list_off.release(); h->list_len.release(); h->list_idx.release(); h->list_val.release(); h->cursor.release();
    h->fps.release(); h->frame_rec.release(); h->is_selected.release(); h->tlut.release();
    h->d_sampled.release(); h->d_out.release(); h->d_result.release();
    h->pd_d2.release(); h->feat.release(); h->pd_tiles.release(); h->pd_norm.release(); h->pd_fps.release(); h->pd_boxes.release();
    h->bev_pts.release(); h->bev_off.release(); h->bev_counts.release(); h->bev_norm2.release(); h->bev_levels.release();
    cudaEventDestroy(h->ev[0]);
    cudaEventDestroy(h->ev[1]);
    cudaEventDestroy(h->timer_ev[0]);
    cudaEventDestroy(h->timer_ev[1]);
    cudaStreamDestroy(h->stream);
    delete h;
    return ADSEL_OK;
}

int adsel_get_stream(adsel_handle *h, void **stream) {
    ADSEL_REQUIRE(h && stream, "adsel_get_stream: null argument");
    *stream = (void *)h->stream;
    return ADSEL_OK;
}

int adsel_synchronize(adsel_handle *h) {
    ADSEL_REQUIRE(h, "adsel_synchronize: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    return ADSEL_OK;
}

int adsel_timer_start(adsel_handle *h) {
    ADSEL_REQUIRE(h, "adsel_timer_start: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    ADSEL_CUDA(cudaEventRecord(h->timer_ev[0], h->stream));
    return ADSEL_OK;
}

int adsel_timer_stop(adsel_handle *h, float *elapsed_ms) {
    ADSEL_REQUIRE(h && elapsed_ms, "adsel_timer_stop: null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaEventRecord(h->timer_ev[1], h->stream));
    ADSEL_CUDA(cudaEventSynchronize(h->timer_ev[1]));
    ADSEL_CUDA(cudaEventElapsedTime(elapsed_ms, h->timer_ev[0], h->timer_ev[1]));
    return ADSEL_OK;
}

int adsel_set_pool(adsel_handle *h, int32_t n, const double *loc_xy, const int32_t *group_id, const int32_t *n_boxes) {
    ADSEL_REQUIRE(h && n > 0 && loc_xy && group_id && n_boxes, "adsel_set_pool: null/empty argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    int32_t n_groups = 0;
    for (int i = 0; i < n; ++i) {
        ADSEL_REQUIRE(group_id[i] >= 0, "adsel_set_pool: negative group id");
        ADSEL_REQUIRE(n_boxes[i] >= 0, "adsel_set_pool: negative box count");
        if (group_id[i] + 1 > n_groups) n_groups = group_id[i] + 1;
    }
    // frames by group, ascending index inside a group (counting sort)
    std::vector<int32_t> ptr(n_groups + 1, 0), members(n);
    for (int i = 0; i < n; ++i) ptr[group_id[i] + 1]++;
    for (int g = 0; g < n_groups; ++g) ptr[g + 1] += ptr[g];
    {
        std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
        for (int i = 0; i < n; ++i) members[fill[group_id[i]]++] = i;
    }
    h->bbox[0] = h->bbox[2] = loc_xy[0];
    h->bbox[1] = h->bbox[3] = loc_xy[1];
    for (int i = 0; i < n; ++i) {
        const double x = loc_xy[2 * i], y = loc_xy[2 * i + 1];
        ADSEL_REQUIRE(x == x && y == y, "adsel_set_pool: NaN ego location");
        h->bbox[0] = x < h->bbox[0] ? x : h->bbox[0];
        h->bbox[1] = y < h->bbox[1] ? y : h->bbox[1];
        h->bbox[2] = x > h->bbox[2] ? x : h->bbox[2];
        h->bbox[3] = y > h->bbox[3] ? y : h->bbox[3];
    }
    h->max_group = 0;
    for (int g = 0; g < n_groups; ++g) h->max_group = ptr[g + 1] - ptr[g] > h->max_group ? ptr[g + 1] - ptr[g] : h->max_group;
    h->groups_contiguous = true;   // consecutive runs: frames sorted by group are already in index order
    for (int i = 0; i < n; ++i) h->groups_contiguous = h->groups_contiguous && members[i] == i;
    h->n = n;
    h->n_groups = n_groups;
    update_shard(h);
    h->graph_ready = false;
    h->lists_mode = -1;
    ADSEL_TRY(h->loc.reserve(n));
    ADSEL_TRY(h->group_id.reserve(n));
    ADSEL_TRY(h->n_boxes.reserve(n));
    ADSEL_TRY(h->group_ptr.reserve(n_groups + 1));
    ADSEL_TRY(h->group_members.reserve(n));
    ADSEL_CUDA(cudaMemcpyAsync(h->loc.p, loc_xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaMemcpyAsync(h->group_id.p, group_id, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaMemcpyAsync(h->n_boxes.p, n_boxes, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaMemcpyAsync(h->group_ptr.p, ptr.data(), sizeof(int32_t) * (n_groups + 1), cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaMemcpyAsync(h->group_members.p, members.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));   // host vectors go out of scope
    return ADSEL_OK;
}

int adsel_build_knn_graph(adsel_handle *h, int32_t k) {
    ADSEL_REQUIRE(h, "adsel_build_knn_graph: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return build_knn_graph(h, k);
}

int adsel_get_knn(adsel_handle *h, int32_t *knn_idx, double *knn_dist) {
    ADSEL_REQUIRE(h && h->graph_ready && knn_idx && knn_dist, "adsel_get_knn: graph not built / null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    const size_t cells = (size_t)h->n * (h->k + 1);
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    ADSEL_CUDA(cudaMemcpy(knn_idx, h->knn_idx_p, sizeof(int32_t) * cells, cudaMemcpyDeviceToHost));
    ADSEL_CUDA(cudaMemcpy(knn_dist, h->knn_dist_p, sizeof(double) * cells, cudaMemcpyDeviceToHost));
    return ADSEL_OK;
}

int adsel_get_csr(adsel_handle *h, int32_t *row_ptr, int32_t *col, double *w, int64_t *nnz) {
    ADSEL_REQUIRE(h && h->graph_ready && nnz, "adsel_get_csr: graph not built / null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_TRY(refresh_nnz(h));
    *nnz = h->nnz;
    if (row_ptr) ADSEL_CUDA(cudaMemcpy(row_ptr, h->row_ptr.p, sizeof(int32_t) * (h->n + 1), cudaMemcpyDeviceToHost));
    if (col) ADSEL_CUDA(cudaMemcpy(col, h->col.p, sizeof(int32_t) * h->nnz, cudaMemcpyDeviceToHost));
    if (w) ADSEL_CUDA(cudaMemcpy(w, h->w.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost));
    return ADSEL_OK;
}

int adsel_geodesic_rows(adsel_handle *h, const int32_t *sources, int32_t m, double radius, double *out) {
    ADSEL_REQUIRE(h, "adsel_geodesic_rows: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return geodesic_rows(h, sources, m, radius, out);
}

int adsel_geodesic_max_finite(adsel_handle *h, double *out) {
    ADSEL_REQUIRE(h && out, "adsel_geodesic_max_finite: null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return geodesic_max_finite(h, out);
}

int adsel_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                 int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h, "adsel_select: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return greedy_select(h, p, sampled, n_sampled, first_index, out_idx, out_cap, out_m, out_cost);
}

int adsel_build_knn_graph_sharded(adsel_handle *h, int32_t k) {
    ADSEL_REQUIRE(h, "adsel_build_knn_graph_sharded: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return build_knn_graph_sharded(h, k);
}

int adsel_shard_prepare(adsel_handle *h, int32_t k, const adsel_select_params *p, int32_t n_sampled, int32_t out_cap) {
    ADSEL_REQUIRE(h && p && h->n > 0 && out_cap > 0, "adsel_shard_prepare: set the pool first / null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    if (p->spatial == ADSEL_SPATIAL_GEODESIC) ADSEL_TRY(knn_prepare(h, k, false));
    ADSEL_TRY(greedy_prepare(h, p, n_sampled, out_cap, true));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    return ADSEL_OK;
}

int adsel_select_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                               int32_t first_index, int32_t out_cap) {
    ADSEL_REQUIRE(h, "adsel_select_sharded_begin: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return greedy_select_sharded_begin(h, p, sampled, n_sampled, first_index, out_cap, false);
}

int adsel_select_replicated_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                                  int32_t first_index, int32_t out_cap) {
    ADSEL_REQUIRE(h, "adsel_select_replicated_begin: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return greedy_select_sharded_begin(h, p, sampled, n_sampled, first_index, out_cap, true);
}

int adsel_select_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h, "adsel_select_sharded_end: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    if (h->knn_pending_timing) {   // the graph build enqueued before this selection has completed by now
        h->knn_pending_timing = false;
        ADSEL_CUDA(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        ADSEL_CUDA(cudaEventElapsedTime(&ms, h->knn_ev[0], h->knn_ev[1]));
        h->stage_ms[0] = ms;
        h->stage_ms[1] = 0.f;
    }
    return greedy_select_sharded_end(h, out_idx, out_cap, out_m, out_cost);
}

int adsel_get_fps(adsel_handle *h, double *fps) {
    ADSEL_REQUIRE(h && fps && h->fps.p, "adsel_get_fps: no selection has run");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaMemcpy(fps, h->fps.p, sizeof(double) * h->n, cudaMemcpyDeviceToHost));
    return ADSEL_OK;
}

int adsel_get_stage_ms(adsel_handle *h, float *ms8) {
    ADSEL_REQUIRE(h && ms8, "adsel_get_stage_ms: null argument");
    memcpy(ms8, h->stage_ms, sizeof(h->stage_ms));
    return ADSEL_OK;
}

int adsel_get_counters(adsel_handle *h, int64_t *c4) {
    ADSEL_REQUIRE(h && c4, "adsel_get_counters: null argument");
    if (h->graph_ready) {
        ADSEL_CUDA(cudaSetDevice(h->device));
        ADSEL_TRY(refresh_nnz(h));
    }
    c4[0] = h->launches;
    c4[1] = h->last_steps;
    c4[2] = h->list_total;
    c4[3] = h->nnz;
    return ADSEL_OK;
}

int adsel_get_greedy_info(adsel_handle *h, int32_t *info2) {
    ADSEL_REQUIRE(h && info2, "adsel_get_greedy_info: null argument");
    info2[0] = h->greedy_variant;
    info2[1] = h->greedy_cluster;
    return ADSEL_OK;
}

int adsel_bev_set_points(adsel_handle *h, const float *points, const int64_t *frame_off, int32_t n_frames) {
    ADSEL_REQUIRE(h, "adsel_bev_set_points: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return bev_set_points(h, points, frame_off, n_frames);
}

int adsel_bev_generate(adsel_handle *h, int32_t n_frames, const int64_t *frame_off, int32_t n_sweeps, uint64_t seed) {
    ADSEL_REQUIRE(h, "adsel_bev_generate: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return bev_generate(h, n_frames, frame_off, n_sweeps, seed);
}

int adsel_bev_get_points(adsel_handle *h, float *points) {
    ADSEL_REQUIRE(h && points && h->bev_batch > 0, "adsel_bev_get_points: no points / null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaMemcpy(points, h->bev_pts.p, sizeof(float) * 5 * (size_t)h->bev_points, cudaMemcpyDeviceToHost));
    return ADSEL_OK;
}

int adsel_bev_build(adsel_handle *h, const adsel_bev_params *p, int32_t aggregate_atomics, int32_t frame_base,
                    int32_t total_frames) {
    ADSEL_REQUIRE(h && p, "adsel_bev_build: null argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return bev_build(h, p, aggregate_atomics, frame_base, total_frames);
}

int adsel_bev_get_descriptors(adsel_handle *h, int32_t *counts, uint16_t *levels, int32_t *norm2, int32_t *cells) {
    ADSEL_REQUIRE(h && h->bev_cells > 0, "adsel_bev_get_descriptors: call adsel_bev_build first");
    ADSEL_CUDA(cudaSetDevice(h->device));
    const size_t tot = (size_t)h->bev_frames * h->bev_cells;
    if (cells) *cells = h->bev_cells;
    if (counts) ADSEL_CUDA(cudaMemcpy(counts, h->bev_counts.p, sizeof(int32_t) * tot, cudaMemcpyDeviceToHost));
    if (levels) ADSEL_CUDA(cudaMemcpy(levels, h->desc_levels, sizeof(uint16_t) * tot, cudaMemcpyDeviceToHost));
    if (norm2) ADSEL_CUDA(cudaMemcpy(norm2, h->desc_norm2, sizeof(int32_t) * h->bev_frames, cudaMemcpyDeviceToHost));
    return ADSEL_OK;
}

int adsel_pairdist_set_descriptors(adsel_handle *h, const uint16_t *levels_bf16, int32_t n, int32_t dim) {
    ADSEL_REQUIRE(h, "adsel_pairdist_set_descriptors: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return pairdist_set_descriptors(h, levels_bf16, n, dim);
}

int adsel_pairdist_build(adsel_handle *h) {
    ADSEL_REQUIRE(h, "adsel_pairdist_build: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return pairdist_build(h);
}

int adsel_pairdist_get_rows(adsel_handle *h, int32_t row0, int32_t rows, int32_t *out) {
    ADSEL_REQUIRE(h && out && h->pd_n > 0, "adsel_pairdist_get_rows: call adsel_pairdist_build first");
    ADSEL_REQUIRE(row0 >= 0 && rows > 0 && row0 + rows <= h->pd_n, "adsel_pairdist_get_rows: rows out of range");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    // out is [rows, pd_cols]: the columns this handle holds (all of them unless the matrix is a column shard)
    if (h->pd_cols > 0)
        ADSEL_CUDA(cudaMemcpy2D(out, sizeof(int32_t) * h->pd_cols, h->d2_p + (size_t)row0 * h->pd_ld, sizeof(int32_t) * h->pd_ld,
                                sizeof(int32_t) * h->pd_cols, rows, cudaMemcpyDeviceToHost));
    return ADSEL_OK;
}

int adsel_select_dense(adsel_handle *h, const adsel_select_params *p, const int32_t *n_boxes, const int32_t *sampled,
                       int32_t n_sampled, int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m,
                       double *out_cost) {
    ADSEL_REQUIRE(h, "adsel_select_dense: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return dense_select(h, p, n_boxes, sampled, n_sampled, first_index, out_idx, out_cap, out_m, out_cost);
}

int adsel_select_fused_feature(adsel_handle *h, const adsel_select_params *p, double lambda_f, int32_t seed_spatial_only,
                               const int32_t *sampled, int32_t n_sampled, int32_t first_index, int32_t *out_idx,
                               int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h, "adsel_select_fused_feature: null handle");
    ADSEL_REQUIRE(lambda_f >= 0.0, "adsel_select_fused_feature: lambda_f must be >= 0");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return fused_feature_select(h, p, lambda_f, seed_spatial_only, sampled, n_sampled, first_index, out_idx, out_cap, out_m, out_cost);
}

int adsel_feature_distance_build(adsel_handle *h, const float *feats, int32_t n, int32_t dim, int32_t p) {
    ADSEL_REQUIRE(h, "adsel_feature_distance_build: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return feature_distance_build(h, feats, n, dim, p);
}

int adsel_pairdist_build_sharded(adsel_handle *h) {
    ADSEL_REQUIRE(h, "adsel_pairdist_build_sharded: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return pairdist_build_sharded(h);
}

int adsel_select_dense_prepare(adsel_handle *h, int32_t n_sampled, int32_t out_cap) {
    ADSEL_REQUIRE(h && out_cap > 0, "adsel_select_dense_prepare: bad argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return dense_prepare(h, n_sampled, out_cap);
}

int adsel_select_dense_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *n_boxes, const int32_t *sampled,
                                     int32_t n_sampled, int32_t first_index, int32_t out_cap) {
    ADSEL_REQUIRE(h, "adsel_select_dense_sharded_begin: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return dense_select_sharded_begin(h, p, n_boxes, sampled, n_sampled, first_index, out_cap);
}

int adsel_select_dense_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h, "adsel_select_dense_sharded_end: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return dense_select_sharded_end(h, out_idx, out_cap, out_m, out_cost);
}

int adsel_exp_f64(adsel_handle *h, const double *x, int64_t n, double *out) {
    ADSEL_REQUIRE(h && x && out && n > 0, "adsel_exp_f64: null/empty argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    double *dx = nullptr, *dy = nullptr;
    ADSEL_CUDA(cudaMalloc(&dx, sizeof(double) * n));
    ADSEL_CUDA(cudaMalloc(&dy, sizeof(double) * n));
    ADSEL_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    exp_f64_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(dx, n, dy);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_CUDA(cudaMemcpyAsync(out, dy, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(dx);
    cudaFree(dy);
    return ADSEL_OK;
}

int adsel_exp_f32(adsel_handle *h, const float *x, int64_t n, float *out) {
    ADSEL_REQUIRE(h && x && out && n > 0, "adsel_exp_f32: null/empty argument");
    ADSEL_CUDA(cudaSetDevice(h->device));
    float *dx = nullptr, *dy = nullptr;
    ADSEL_CUDA(cudaMalloc(&dx, sizeof(float) * n));
    if (cudaMalloc(&dy, sizeof(float) * n) != cudaSuccess) {
        cudaFree(dx);
        ADSEL_CUDA(cudaErrorMemoryAllocation);
    }
    ADSEL_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    exp_f32_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(dx, n, dy);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_CUDA(cudaMemcpyAsync(out, dy, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(dx);
    cudaFree(dy);
    return ADSEL_OK;
}

}  // extern "C"
