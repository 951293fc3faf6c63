/*
 * adsel_b200.h -- C ABI of the B200-native diversity-based active-selection hot path.
 *
 * Drop-in boundary.  The reference has no native code on this path: the whole of
 * det3d/selectors/spatial_temporal_selector.py:59-193 is host numpy/scipy.  Each entry point below
 * names the reference lines it replaces.  Plain pointers and sizes only; no torch types.  Host
 * arrays are owned by the caller, device buffers by the handle.  One handle per CUDA device / rank.
 * Every function returns 0 on success and a negative adsel_status otherwise (never throws across
 * the ABI); adsel_last_error() returns the message of the calling thread's last failure.
 *
 * There is NO CPU fallback: every compute entry point launches sm_100a kernels and fails with
 * ADSEL_ERR_CUDA when no device is present.
 */
#ifndef ADSEL_B200_H
#define ADSEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct adsel_handle adsel_handle;

enum adsel_status {
    ADSEL_OK = 0,
    ADSEL_ERR_INVALID = -1,    /* bad argument / call order */
    ADSEL_ERR_CUDA = -2,       /* CUDA runtime failure (message has the cudaError string) */
    ADSEL_ERR_CAPACITY = -3,   /* an output buffer supplied by the caller is too small */
    ADSEL_ERR_RESELECTED = -4, /* greedy argmax landed on an already selected frame:
                                  the reference's AssertionError (spatial_temporal_selector.py:182) */
    ADSEL_ERR_COMM = -5,       /* sharded path: timed out waiting for a peer rank (mailbox / barrier) */
    ADSEL_ERR_PEER_ABORT = -6  /* sharded path: a peer rank aborted the selection; every rank returns an error
                                  from the same call, so a caller may retry collectively */
};

/* spatial term of a row of the distance map */
enum adsel_spatial {
    ADSEL_SPATIAL_NONE = 0,      /* TemporalSelector (temporal_selector.py:41-104) */
    ADSEL_SPATIAL_GEODESIC = 1,  /* kNN-graph Dijkstra geodesic (spatial_temporal_selector.py:93-104) */
    ADSEL_SPATIAL_EUCLID = 2     /* 2-D Euclid inside a group, margin 1e6 outside
                                    (euclidean_spatial_selector.py:93-101) */
};
enum adsel_normalize { ADSEL_NORM_RAW = 0, ADSEL_NORM_EXP = 1, ADSEL_NORM_LINEAR = 2 };
enum adsel_aggregate { ADSEL_AGG_SUM = 0, ADSEL_AGG_MIN = 1, ADSEL_AGG_MAX = 2 };

typedef struct adsel_select_params {
    int32_t spatial;          /* adsel_spatial */
    int32_t temporal;         /* 0: no temporal term; 1: |i-j| inside the frame's group else 1e6
                                 (spatial_temporal_selector.py:110-134, temporal_selector.py:50-63) */
    int32_t normalize;        /* adsel_normalize (spatial_temporal_selector.py:138-146) */
    int32_t aggregate;        /* adsel_aggregate (spatial_temporal_selector.py:148-155) */
    double lambda_t;          /* weight of the temporal term for ADSEL_AGG_SUM (:151) */
    double max_spatial;       /* ADSEL_NORM_LINEAR divisor, from adsel_geodesic_max_finite (:139) */
    double max_temporal;      /* ADSEL_NORM_LINEAR divisor "max_temporal_distance" (:127-129,141) */
    double cost_b;            /* per-box annotation cost (base_selector.py:24,85) */
    double cost_f;            /* per-frame annotation cost (base_selector.py:25,82) */
    double cost_start;        /* get_cost_amount() of the frames already in the buffer (base_selector.py:78-86) */
    double budget;            /* int(current_budget) (spatial_temporal_selector.py:186) */
    int32_t forbid_sampled;   /* 1: also assert against the buffer's frames (spatial_selector.py:124) */
    int32_t reserved;
} adsel_select_params;

/* ---- library ------------------------------------------------------------------------------ */
const char *adsel_version(void);
const char *adsel_last_error(void);
int adsel_device_count(int *count);

/* ---- handle ------------------------------------------------------------------------------- */
int adsel_create(int device, adsel_handle **out);
int adsel_destroy(adsel_handle *h);
/* cudaStream_t all of the handle's work is enqueued on (returned as void*). */
int adsel_get_stream(adsel_handle *h, void **stream);
int adsel_synchronize(adsel_handle *h);
/* Device timer: CUDA events recorded on the handle's stream (bench.py's timed region). */
int adsel_timer_start(adsel_handle *h);
int adsel_timer_stop(adsel_handle *h, float *elapsed_ms);

/* ---- pool (replaces the per-frame python loops, spatial_temporal_selector.py:78-91,114-129) ---
 * loc_xy   [n,2] float64  ego x,y per frame (the host keeps the reference's own expression
 *                         -(cal[:3,3].T @ cal[:3,:3])[:2]; BLAS may fuse, so it is not redone here)
 * group_id [n]   int32    temporal group per frame: consecutive-run id (SpatialTemporalSelector),
 *                         logfile-name id (TemporalSelector) or city id (EuSpatialSelector)
 * n_boxes  [n]   int32    len(info["gt_names"]) per frame (cost model)
 */
int adsel_set_pool(adsel_handle *h, int32_t n, const double *loc_xy, const int32_t *group_id,
                   const int32_t *n_boxes);
/* Every frame's ego location redone on `device` in the reference's literal order (spatial_temporal_selector.py:83-91):
 * pose9 [m,9] float64 = {c00,c01,c03, c10,c11,c13, c20,c21,c23} of car_from_global, loc_xy [m,2] what the host extraction
 * produced, fused = which rounding numpy's vector-matrix product was found to use on this host (1: fma chain, 0: separately
 * rounded products).  *n_bad = frames whose bit patterns differ (NaN == NaN), *first_bad = the first of them or -1.
 * Host pointers; needs no handle (it runs before the pool exists). */
int adsel_check_ego_locations(int device, const double *pose9, const double *loc_xy, int64_t m, int fused, int64_t *n_bad,
                              int64_t *first_bad);

/* ---- spatial graph (replaces cKDTree.query + dense scatter + dense->CSR,
 *      spatial_temporal_selector.py:93-101 and scipy's csgraph_from_dense |w|<=1e-8 drop) -------- */
/* Device-pointer variants for callers whose arrays already live in HBM on the handle's device (torch.ops.adsel.*): the
 * same semantics and errors as adsel_set_pool / adsel_select / adsel_bev_set_points, no host round trip of the inputs
 * (validation, bounding box, group sizes and the frames-by-group table are computed by kernels; only a 64-byte summary
 * and the result scalars cross to the host).  Pointers are read on the handle's stream: the caller orders its own
 * producer streams before the call.  adsel_copy_selected_device copies the m indices of the last adsel_select* call. */
int adsel_set_pool_device(adsel_handle *h, int32_t n, const double *d_loc_xy, const int32_t *d_group_id,
                          const int32_t *d_n_boxes);
int adsel_select_device(adsel_handle *h, const adsel_select_params *p, const int32_t *d_sampled, int32_t n_sampled,
                        int32_t first_index, int32_t *out_m, double *out_cost);
int adsel_copy_selected_device(adsel_handle *h, int32_t *d_dst, int32_t m);
int adsel_bev_set_points_device(adsel_handle *h, const float *d_points, const int64_t *frame_off_host, int32_t n_frames);
/* device pointers of the descriptors built by adsel_bev_build (owned by the handle, valid until the next build) */
int adsel_bev_descriptors_device(adsel_handle *h, int32_t **d_counts, uint16_t **d_levels, int32_t **d_norm2, int32_t *frames,
                                 int32_t *cells);

int adsel_build_knn_graph(adsel_handle *h, int32_t k);
/* test/inspection read-backs (host buffers): knn_idx [n,k+1] int32, knn_dist [n,k+1] float64 */
int adsel_get_knn(adsel_handle *h, int32_t *knn_idx, double *knn_dist);
/* CSR of the symmetrised graph: row_ptr [n+1] int32; col/w [nnz] (pass NULL to query nnz only) */
int adsel_get_csr(adsel_handle *h, int32_t *row_ptr, int32_t *col, double *w, int64_t *nnz);

/* ---- geodesic rows (replaces csgraph.shortest_path(..., method="D"), :103-104) ---------------
 * Dense rows for m sources: out [m,n] float64, inf when unreachable.  radius < inf truncates the
 * search: entries whose geodesic distance is >= radius are reported as inf. */
int adsel_geodesic_rows(adsel_handle *h, const int32_t *sources, int32_t m, double radius, double *out);
/* max over all finite entries of the full geodesic map (linear normalisation, :139). */
int adsel_geodesic_max_finite(adsel_handle *h, double *out);

/* ---- selection (replaces :110-193: temporal map, normalise, aggregate, greedy k-center loop,
 *      cost-budget stop).  sampled [n_sampled]: frames already in the buffer; first_index: the
 *      caller's random.choice(range(n)) used when n_sampled == 0 (:168), ignored otherwise.
 *      out_idx [out_cap] receives the NEW frames in selection order, *out_m their count,
 *      *out_cost the accumulated cost including the frame that overflowed the budget. ---------- */
int adsel_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                 int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);
/* Read back the final min-distance-to-selected vector fps [n] float64 (tests). */
int adsel_get_fps(adsel_handle *h, double *fps);
/* Device-time of the stages of the last build/select calls, milliseconds (CUDA events on the
 * handle's stream): [0] knn, [1] csr build, [2] geodesic lists, [3] greedy loop, [4] geodesic max,
 * [5] bev histogram, [6] pairwise distance tiles, [7] dense greedy loop. */
int adsel_get_stage_ms(adsel_handle *h, float *ms8);
/* counters of the last select: [0] kernels launched since adsel_create, [1] greedy steps,
 * [2] total neighbour-list entries, [3] csr nnz */
int adsel_get_counters(adsel_handle *h, int64_t *c4);
/* which greedy kernel the last selection ran: info2[0] = 0 global-memory loop (pool beyond a cluster's shared memory, column
 * shards), 1 multi-warp resident kernel (one CTA or a thread-block cluster: greedy_fast_kernel), 2 generic resident kernel
 * (Euclid term, non-contiguous groups: greedy_resident_kernel), 3 lazy rows (one search per selected frame), 4 fused
 * feature loop, 5 one warp on resident columns with one slot load per step (greedy_warp_kernel: geodesic rows with the exp
 * normalisation on pools that fit one CTA's shared memory), 6 the same loop with the columns spread over the shared memory
 * of a 2 / 4 / 8-CTA cluster of passive CTAs (greedy_warp_cluster_kernel, up to ~229 k frames); info2[1] = CTAs of the cluster */
int adsel_get_greedy_info(adsel_handle *h, int32_t *info2);

/* ---- pool sharded over the GPUs of one NVSwitch box -------------------------------
 * The reference issues no collective here: under torch.distributed.launch every rank recomputes the
 * whole selection and rank 0 writes (base_selector.py:57, tools/active_select.py:94-101).  In this
 * library one handle per rank owns a contiguous, 256-aligned slice of the pool's columns; the kNN rows
 * are computed per slice and stored straight into every rank's table over NVLink, and every greedy
 * step ends with ONE exchange of the ranks' local (value, index) maxima, done inside the persistent
 * greedy kernel through peer-memory mailboxes (np.argmax's first-maximum rule holds across shards).
 *
 * Call order on every rank: adsel_create, adsel_comm_init (-> 64-byte CUDA IPC handle), exchange the
 * handles by any means (torch.distributed all_gather in this repo), adsel_comm_attach, adsel_set_pool,
 * adsel_build_knn_graph_sharded, adsel_select_sharded_begin, adsel_select_sharded_end.  The *_sharded
 * calls only enqueue work (no host synchronisation, no allocation after a first call of the same
 * shape), so one host thread may also drive several handles (adsel_comm_attach_local: ranks that live in
 * one process, on one or several devices).  All ranks must make the same sequence of sharded calls. */
/* n, k, dim, list_cap size the peer-visible arena (kNN tables [n,k+1], descriptors [n,dim], list_cap
 * neighbour-list entries; 0 = default 256 n / world + 16 n); identical on all ranks. */
int adsel_comm_init(adsel_handle *h, int32_t rank, int32_t world, int32_t n, int32_t k, int32_t dim, int64_t list_cap,
                    void *ipc_handle_out /* 64 bytes, may be NULL */);
int adsel_comm_attach(adsel_handle *h, const void *ipc_handles /* world x 64 bytes, rank order */);
int adsel_comm_detach(adsel_handle *h);   /* unmap the peers' arenas; every rank does this before any rank re-sizes or frees */
int adsel_comm_attach_local(adsel_handle *h, adsel_handle *const *handles /* world handles, rank order */);
int adsel_comm_set_timeout(adsel_handle *h, double seconds);   /* device-side wait limit (default 20 s) */
int adsel_comm_barrier(adsel_handle *h);                       /* enqueue a device-side barrier over all ranks */
int adsel_comm_check(adsel_handle *h);                         /* synchronise; ADSEL_ERR_COMM if a barrier timed out */
int adsel_shard_range(adsel_handle *h, int32_t *lo, int32_t *hi);   /* this rank's frames [lo, hi) */
/* optional: performs every allocation of the following build/select calls now, so that they only enqueue
 * (needed when one host thread drives several ranks: cudaFree may wait for a rank's spinning kernel) */
int adsel_shard_prepare(adsel_handle *h, int32_t k, const adsel_select_params *p, int32_t n_sampled, int32_t out_cap);
int adsel_build_knn_graph_sharded(adsel_handle *h, int32_t k);
/* as adsel_select, split in two: `begin` enqueues lists + greedy loop, `end` waits and returns the result
 * (identical on every rank).  The neighbour lists are searched from the sources of the rank's slice and every
 * entry is appended, over NVLink, to the pool of the rank owning its column.  ADSEL_ERR_CAPACITY /
 * ADSEL_ERR_PEER_ABORT from `end` on all ranks means a rank's list region was too small: re-create the arenas
 * with list_cap >= the largest adsel_get_counters()[2] over the ranks, rebuild the graph and call again. */
int adsel_select_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled,
                               int32_t n_sampled, int32_t first_index, int32_t out_cap);
int adsel_select_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);
/* The default multi-GPU schedule since the greedy loop became resident in one SM's shared memory (a step costs less
 * than one NVLink round trip): only the BUILD is sharded.  Every rank searches from the sources of its slice and stores
 * each list entry into EVERY rank's arena over NVLink (size the arenas with list_cap >= the total entry count, e.g.
 * 288 n); then each rank runs the whole greedy loop on the complete lists -- no exchange per step, the same result on
 * every rank, which is what the reference does per rank (base_selector.py:57) minus the redundant build.  Collect with
 * adsel_select_sharded_end (ADSEL_ERR_CAPACITY: the same on every rank, re-create the arenas larger and repeat). */
int adsel_select_replicated_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled,
                                  int32_t n_sampled, int32_t first_index, int32_t out_cap);

/* ---- BEV density descriptor from LiDAR points (north_star kernel 1; no reference function -- the
 *      binning semantics are those of det3d/ops/point_cloud/bev_ops.py:27-53, see csrc/bev.cu) -------
 * points: float32 [P,5] = x,y,z,intensity,dt (det3d/datasets/pipelines/loading.py:121-126), frames
 * concatenated; frame_off int64 [n_frames+1] point offsets (ragged frames allowed, frame_off[0] = 0). */
typedef struct adsel_bev_params {
    float range[6];   /* xmin,ymin,zmin,xmax,ymax,zmax  (examples/active/cbgs_spatial_temporal.py:280) */
    float voxel[3];   /* cell size; grid = round((max-min)/voxel) (bev_ops.py:27-30) */
    int32_t reserved;
} adsel_bev_params;
int adsel_bev_set_points(adsel_handle *h, const float *points, const int64_t *frame_off, int32_t n_frames);
/* synthetic multi-sweep LiDAR generated directly in HBM (counter-based hash keyed by frame/sweep);
 * seed: low 32 bits = generator seed, high 32 bits = pool index of the batch's first frame, so that a frame's
 * points do not depend on how the pool is cut into batches or shards */
int adsel_bev_generate(adsel_handle *h, int32_t n_frames, const int64_t *frame_off, int32_t n_sweeps, uint64_t seed);
int adsel_bev_get_points(adsel_handle *h, float *points);
/* histogram + density level + squared norm for every frame of the resident point batch;
 * aggregate_atomics: 1 = warp-aggregated shared atomics.  The batch's descriptors are written at frame
 * `frame_base` of a pool of `total_frames` frames (total_frames <= 0: the batch is the whole pool), so a
 * pool larger than HBM's share for points can be streamed batch by batch (first batch: frame_base 0). */
int adsel_bev_build(adsel_handle *h, const adsel_bev_params *p, int32_t aggregate_atomics, int32_t frame_base,
                    int32_t total_frames);
/* read-backs: counts int32 [n_frames,cells], levels uint16 (bf16 bits) [n_frames,cells], norm2 int32 [n_frames];
 * any pointer may be NULL; *cells receives ny*nx */
int adsel_bev_get_descriptors(adsel_handle *h, int32_t *counts, uint16_t *levels, int32_t *norm2, int32_t *cells);

/* ---- pairwise squared distances of the descriptors + dense greedy (north_star kernel 3; nearest
 *      reference code: euclidean_spatial_selector.py:96, feature_selector.py:96-105, the unused
 *      det3d/ops/distance/src/cdist_gpu.cu:31-50) -------------------------------------------------
 * descriptors: bf16 bit patterns [n,dim] of integers 0..31 (adsel_bev_build leaves them on the device;
 * adsel_pairdist_set_descriptors uploads caller-supplied ones). */
int adsel_pairdist_set_descriptors(adsel_handle *h, const uint16_t *levels_bf16, int32_t n, int32_t dim);
/* d2[i,j] = |x_i|^2 + |x_j|^2 - 2 x_i.x_j as exact int32, Gram tiles on tcgen05 (TMEM accumulators, TMA loads) */
int adsel_pairdist_build(adsel_handle *h);
/* read back rows [row0,row0+rows) of d2 as int32 [rows,n] (tests) */
int adsel_pairdist_get_rows(adsel_handle *h, int32_t row0, int32_t rows, int32_t *out);
/* greedy k-center over the rows of d2 with the reference's cost-budget loop (spatial_selector.py:101-138);
 * only the cost fields and forbid_sampled of the params are used; n_boxes [n] host array. */
int adsel_select_dense(adsel_handle *h, const adsel_select_params *p, const int32_t *n_boxes, const int32_t *sampled,
                       int32_t n_sampled, int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m,
                       double *out_cost);

/* ---- feature-distance map of the detector-feature selectors (replaces the row loop of
 *      det3d/selectors/feature_selector.py:87-109, same in spatial_feature / spatial_temporal_feature / badge /
 *      uwe selectors).  feats: float32 [n, dim] host array (the reference's mean-pooled FPN features);
 *      p = 1: sum |a-b|; p = 2: the reference's sqrt((a-b)**2).sum(1) (also an L1 distance).  The float32 map
 *      replaces the descriptor distance matrix of the handle: adsel_pairdist_get_rows returns the float32 bit
 *      patterns (non-negative floats order like their int32 bits) and adsel_select_dense runs the greedy loop. */
int adsel_feature_distance_build(adsel_handle *h, const float *feats, int32_t n, int32_t dim, int32_t p);
/* Greedy loop of the selectors that add the feature term to the geodesic (and temporal) terms:
 * det3d/selectors/spatial_feature_selector.py:186-237 (p->temporal = 0, p->aggregate = sum | min | max, seed_spatial_only = 1:
 * the min-distance vector is seeded from the normalised SPATIAL rows only, :201-209) and
 * spatial_temporal_feature_selector.py:208-258 (p->temporal = 1, sum with p->lambda_t and lambda_f, seed_spatial_only = 0).
 * p->spatial must be geodesic, p->normalize exp; the feature term is normalised with numpy's float32 exp and, in the
 * three-term sum, multiplied by float32(lambda_f) in float32, as numpy does.  Needs adsel_build_knn_graph and
 * adsel_feature_distance_build (whose map is the reference's bit for bit when the feature dimension is a multiple of 32). */
int adsel_select_fused_feature(adsel_handle *h, const adsel_select_params *p, double lambda_f, int32_t seed_spatial_only,
                               const int32_t *sampled, int32_t n_sampled, int32_t first_index, int32_t *out_idx,
                               int32_t out_cap, int32_t *out_m, double *out_cost);

/* ---- the BEV / dense path sharded over ranks (after adsel_comm_init with dim = cells, n = pool frames) ----
 * adsel_bev_build then stores each frame's descriptor into EVERY rank's arena (the all-gather is fused into
 * the histogram kernel's epilogue); every rank builds the frames of its own slice [adsel_shard_range), at
 * frame_base = their pool index.  adsel_pairdist_build_sharded (enqueue only): device barrier, the distance
 * tiles of this rank's 256-aligned column slice d2[:, lo:hi), device barrier.  adsel_pairdist_get_rows then
 * returns [rows, hi-lo].  The dense greedy loop runs on the column slice with one mailbox exchange per step. */
int adsel_pairdist_build_sharded(adsel_handle *h);
int adsel_select_dense_prepare(adsel_handle *h, int32_t n_sampled, int32_t out_cap);   /* allocations only */
int adsel_select_dense_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *n_boxes,
                                     const int32_t *sampled, int32_t n_sampled, int32_t first_index, int32_t out_cap);
int adsel_select_dense_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);

/* ---- element-wise helpers (tests) ----------------------------------------------------------- */
/* out[i] = exp(x[i]) with the arithmetic of numpy's AVX512_SKX float64 exp (np.exp in
 * spatial_temporal_selector.py:142-143); host buffers, computed on the device. */
int adsel_exp_f64(adsel_handle *h, const double *x, int64_t n, double *out);
/* out[i] = exp(x[i]) with the arithmetic of numpy's float32 exp (its own AVX512F / AVX2 kernel): the
 * normalisation `1 - np.exp(-feature_distance_map)` of the float32 feature map in
 * det3d/selectors/spatial_feature_selector.py:190 and spatial_temporal_feature_selector.py:214. */
int adsel_exp_f32(adsel_handle *h, const float *x, int64_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif /* ADSEL_B200_H */
