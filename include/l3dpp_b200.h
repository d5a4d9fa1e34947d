/* l3dpp_b200.h -- C ABI of libl3dpp_b200.so
 *
 * B200-native (sm_100a) implementation of the Line3D++ multi-view 2-D segment
 * matching -> two-view triangulation -> multi-view scoring -> affinity-matrix stage that
 * BTREE-C802/3DLine-SLAM runs inside Line3D::matchImages / Line3D::reconstruct3Dlines.
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface it
 * replaces (file:line relative to the reference tree).  Plain pointers and sizes only; all
 * functions return 0 on success or a negative code (l3d_last_error() has the text).  There is no
 * CPU fallback: without a CUDA device every compute call fails with L3D_ERR_CUDA.
 *
 * Two levels:
 *   (1) cudawrapper level -- l3d_match_lines / l3d_score_matches take HOST buffers exactly like
 *       the DataArray arguments of L3DPP::match_lines_GPU / score_matches_GPU
 *       (include/cudawrapper.h:63-81) and are blocking, like DataArray::upload()
 *       (include/dataArray.h:199-218).
 *   (2) Line3D level -- l3d_scene_* / l3d_match_images / l3d_affinity keep every table resident in
 *       HBM and run the whole of Line3D::computeMatches (src/line3D.cc:846-930) and
 *       Line3D::computingAffinityMatrix (src/line3D.cc:2275-2402) on the device.
 */
#ifndef L3DPP_B200_H_
#define L3DPP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L3D_OK 0
#define L3D_ERR_ARG (-1)
#define L3D_ERR_CUDA (-2)
#define L3D_ERR_STATE (-3)
#define L3D_ERR_CAPACITY (-4)

typedef struct l3d_ctx l3d_ctx;

/* One potential match: the POD image of L3DPP::Match (include/commons.h:197-219). */
typedef struct {
    uint32_t src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap_score, score3D;
    float depth_p1, depth_p2, depth_q1, depth_q2;
    uint32_t flags; /* bit0 = match_orientation_ */
} l3d_match;

/* One list entry of matches_[cam][seg] as returned by l3d_get_view_lists (36 bytes). */
typedef struct {
    uint32_t tgt_cam, tgt_seg;
    float overlap_score, score3D;
    float depth_p1, depth_p2, depth_q1, depth_q2;
    uint32_t flags;
} l3d_list_rec;

/* One row of estimated_position3D_ (src/line3D.cc:1957-1967): best match + its 3-D segment. */
typedef struct {
    uint32_t src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap_score, score3D;
    float depth_p1, depth_p2, depth_q1, depth_q2;
    float length;
    uint32_t pad;
    double P1[3], P2[3], dir[3];
} l3d_entry;

/* Arguments of Line3D::addImage (src/line3D.cc:117-121) with the image reduced to its size. */
typedef struct {
    uint32_t cam_id;
    uint32_t width, height;
    uint32_t num_segs;
    double K[9], R[9], t[3]; /* row-major; camera model x = K [R|t] X */
    float median_depth;
} l3d_view;

/* Arguments of Line3D::matchImages (src/line3D.cc:496-498) + the constructor's max_img_width. */
typedef struct {
    float sigma_p;
    float sigma_a;
    uint32_t num_neighbors;
    float epipolar_overlap;
    int32_t knn;
    float const_reg_depth;
    int32_t max_image_width; /* Line3D::max_image_width_, used by the bounds test line3D.cc:1142-1148 */
    int32_t filter_mode;     /* 0: FP32 guard-banded pre-filter (default); 1: none (every pair exact) */
    int32_t keep_scored;     /* 1: keep the pre-filter lists of every view (parity tests) */
    int32_t shard_rank;      /* multi-GPU: this process owns the shard_rank-th of shard_world contiguous view slices */
    int32_t shard_world;     /* 0 or 1: no sharding */
} l3d_params;

typedef struct {
    uint64_t pair_tests;      /* sum over matched pairs of N_src * N_tgt (this shard) */
    uint64_t candidates;      /* survivors of the FP32 pre-filter (this shard) */
    uint64_t forward_matches; /* matches kept after kNN + orientation filter (all shards once gathered) */
    uint64_t scored_entries;  /* sum of list lengths at scoring time */
    uint64_t sim_evals;       /* sibling pairs visited by the scoring kernel */
    uint64_t filtered_entries;
    uint64_t pair_tests_run;  /* of pair_tests, the ones the FP32 kernel evaluated; the others were skipped with their
                                 whole warp of source rows because the target lay outside the warp's epipolar wedge */
    uint32_t num_views, num_pairs, num_pairs_local;
    uint32_t num_entries, num_edges, num_local_ids, num_clusters;
    uint32_t gpu_launches;    /* kernels launched since the last l3d_reset_counters */
} l3d_counts;

/* indices into the array filled by l3d_get_timings (milliseconds, CUDA events on the ctx stream) */
enum {
    L3D_T_PREP = 0,    /* per-segment descriptors + rays */
    L3D_T_PAIRTEST,    /* K1: FP32 pair test + compaction */
    L3D_T_EXACT,       /* K2: exact re-test, triangulation, kNN, orientation filter */
    L3D_T_SCORE,       /* K3: potential lists, similarities, data-flow scoring, filtering */
    L3D_T_AFFINITY,    /* K4 */
    L3D_T_TOTAL,
    L3D_T_K1_KERNEL,   /* sum of K1 kernel launches alone */
    L3D_T_K1_LAUNCHES,
    L3D_T_COUNT
};

const char* l3d_last_error(void);
const char* l3d_version(void);

/* replaces: implicit CUDA context of the reference GPU path. device < 0: current device. */
int l3d_ctx_create(l3d_ctx** out, int device);
void l3d_ctx_destroy(l3d_ctx* ctx);
/* All kernels of ctx are launched on `cuda_stream` (a cudaStream_t; NULL = legacy default). */
int l3d_ctx_set_stream(l3d_ctx* ctx, void* cuda_stream);

/* ------------------------------------------------------------------------------------------
 * (1) cudawrapper level, host buffers in / host buffers out, blocking
 * ------------------------------------------------------------------------------------------ */

/* replaces L3DPP::match_lines_GPU (include/cudawrapper.h:63-71, call site src/line3D.cc:1257-1261)
 * with the results of Line3D::matchingCPU (src/line3D.cc:1097-1212): for every source segment
 * the <= kNN matches with the highest epipolar overlap (all of them if kNN <= 0) that pass the
 * bounds test and have four positive triangulated depths, in the order the reference appends them
 * to matches_[src][r] (rows ascending, priority-queue pop order inside a row).
 * lines_*: n x (x1,y1,x2,y2) float, F/RtKinv_*: row-major 3x3 double, C_*: double[3].
 * out_row_off: n_src+1 offsets into out (may be NULL). Returns L3D_ERR_CAPACITY if cap is too small
 * (*out_count then holds the required size). */
int l3d_match_lines(l3d_ctx* ctx, const float* lines_src, uint32_t n_src, const float* lines_tgt,
                    uint32_t n_tgt, const double* F, const double* RtKinv_src,
                    const double* RtKinv_tgt, const double* C_src, const double* C_tgt,
                    uint32_t src_cam, uint32_t tgt_cam, float epi_overlap, int32_t knn,
                    int32_t max_image_width, int32_t filter_mode, l3d_match* out, uint64_t cap,
                    uint64_t* out_count, uint32_t* out_row_off);

/* replaces L3DPP::score_matches_GPU (include/cudawrapper.h:74-81, call site src/line3D.cc:1633-1635)
 * with the arithmetic of Line3D::scoringCPU's new-match branch (src/line3D.cc:1513-1547):
 * matches[i] = {srcSeg, tgtCam, depth_p1, depth_p2} (float4, src/line3D.cc:1616-1617),
 * ranges[s] = {first,last} inclusive or {-1,-1} (int2, src/line3D.cc:1582-1596),
 * regularizers_tgt[i] = {sigma_tgt(P1), sigma_tgt(P2)} (float2, src/line3D.cc:1619-1620),
 * scores[i] receives score3D_. */
int l3d_score_matches(l3d_ctx* ctx, const float* lines, uint32_t n_lines, const float* matches,
                      uint32_t n_matches, const int32_t* ranges, float* scores,
                      const float* regularizers_tgt, const double* RtKinv, const double* C,
                      float two_sigA_sqr, float k, float min_similarity);

/* ------------------------------------------------------------------------------------------
 * (2) Line3D level, tables resident in HBM
 * ------------------------------------------------------------------------------------------ */

/* replaces Line3D::addImage + UpdataImage bookkeeping (src/line3D.cc:117-227, 433-487) for a batch
 * of views: begin, add every active view (explicit neighbour lists = the
 * neighbors_by_worldpoints=false path, src/line3D.cc:604-616), commit (uploads the tables). */
int l3d_scene_begin(l3d_ctx* ctx);
int l3d_scene_add_view(l3d_ctx* ctx, const l3d_view* view, const float* segs_xyxy,
                       const uint32_t* neighbor_cam_ids, uint32_t num_neighbors);
int l3d_scene_commit(l3d_ctx* ctx);
/* the three calls above for n_views views at once: segs_concat holds the views' segments back to
 * back (views[i].num_segs each), nbrs_concat their neighbour camera ids (nbr_counts[i] each). */
int l3d_scene_set(l3d_ctx* ctx, const l3d_view* views, uint32_t n_views, const float* segs_concat,
                  const uint32_t* neighbor_cam_ids_concat, const uint32_t* nbr_counts);
/* neighbors_by_worldpoints = true (Line3D::Line3D src/line3D.cc:60-75, processWPlist :230-241): the
 * lists hold the world-point ids each view observes (VisualSfM .nvm input); l3d_match_images then
 * chooses the visual neighbours like Line3D::findVisualNeighborsFromWPs (src/line3D.cc:723-843).
 * All views of a scene use the same kind of list. */
int l3d_scene_add_view_wps(l3d_ctx* ctx, const l3d_view* view, const float* segs_xyxy,
                           const uint32_t* worldpoint_ids, uint32_t num_worldpoints);
int l3d_scene_set_wps(l3d_ctx* ctx, const l3d_view* views, uint32_t n_views, const float* segs_concat,
                      const uint32_t* worldpoint_ids_concat, const uint32_t* wp_counts);
/* host-only (no device): the neighbours the call above would choose; out_cam_ids has
 * n_views x num_neighbors slots, out_counts[i] of row i are filled (camera ids, ascending) */
int l3d_neighbors_from_worldpoints(const l3d_view* views, uint32_t n_views, const uint32_t* wps_concat,
                                   const uint32_t* wp_counts, uint32_t num_neighbors,
                                   uint32_t* out_cam_ids, uint32_t* out_counts);
/* visual neighbours of a view as used by the last l3d_match_images (camera ids, ascending) */
int l3d_get_neighbors(l3d_ctx* ctx, uint32_t cam_id, uint32_t* out, uint32_t cap, uint32_t* count);

/* replaces L3DPP::find_collinear_segments_GPU (include/cudawrapper.h:84-86) as View::findCollinGPU
 * calls it (src/view.cc:203-236): the N x N byte table of a view's collinear segments, computed like
 * View::findCollinCPU (src/view.cc:238-293).  buffer[r * row_stride_bytes + c] = 1 iff c is collinear to
 * r (the reference reads buffer->dataCPU(c, r)).  Host arrays, blocking.  (The affinity stage with
 * collinearity_t > 0, src/line3D.cc:2328-2396, is not part of this build: the reference runs with -1.) */
int l3d_find_collinear(l3d_ctx* ctx, const float* lines_xyxy, uint32_t n, float dist_t, char* buffer,
                       uint64_t row_stride_bytes);

/* ---- incremental (key-frame stream) mode: the calls L3DPPing::Run (src/L3DPPing.cpp:98-236) makes
 * on its Line3D object between two reconstructions.  The context keeps what Line3D keeps from one
 * matchImages to the next: matched_ (a view pair is matched once), processed_, the filtered match
 * lists with their scores, and Add_camID_ / Delete_camID_ for the score deltas of
 * Line3D::scoringCPU (src/line3D.cc:1439-1512).  After l3d_stream_begin, l3d_match_images,
 * l3d_affinity, l3d_cluster and the getters act on the stream state (one GPU).
 *   l3d_stream_begin        new Line3D object; neighbors_by_worldpoints as in its constructor
 *   l3d_stream_begin_cycle  the resets of src/L3DPPing.cpp:98-103 (world-point maps, Add/Delete sets)
 *   l3d_stream_add_image    Line3D::addImage   (src/line3D.cc:117-227); camera ids ascending
 *   l3d_stream_delete_image Line3D::deleteImage (src/line3D.cc:396-430)
 *   l3d_stream_update_image Line3D::UpdataImage (src/line3D.cc:433-487): pose + world points / neighbours */
int l3d_stream_begin(l3d_ctx* ctx, int neighbors_by_worldpoints);
int l3d_stream_begin_cycle(l3d_ctx* ctx);
int l3d_stream_add_image(l3d_ctx* ctx, const l3d_view* view, const float* segs_xyxy,
                         const uint32_t* wps_or_nbrs, uint32_t n_list);
int l3d_stream_delete_image(l3d_ctx* ctx, uint32_t cam_id);
int l3d_stream_update_image(l3d_ctx* ctx, uint32_t cam_id, const double* R, const double* t,
                            float median_depth, const uint32_t* wps_or_nbrs, uint32_t n_list);

/* replaces Line3D::matchImages (src/line3D.cc:496-640): translate(), spatial regularisers,
 * computeMatches() (matching, orientation filter, scoring, inverse matches, filtering) and the
 * estimated_position3D_ table, all on the device.  l3d_match_stage12 runs matching only (this
 * shard's pairs), l3d_match_stage3 the scoring; see the multi-GPU section for sharded runs. */
int l3d_match_images(l3d_ctx* ctx, const l3d_params* params);
int l3d_match_stage12(l3d_ctx* ctx, const l3d_params* params);
int l3d_match_stage3(l3d_ctx* ctx);

/* replaces Line3D::computingAffinityMatrix (src/line3D.cc:2275-2402) incl. the median scene depth
 * (src/line3D.cc:2074-2091): builds A_ (edge list with first-touch local IDs) on the device. */
int l3d_affinity(l3d_ctx* ctx);

/* replaces SparseMatrix::SparseMatrix (src/sparsematrix.cc:8-61), the device layout of A_ that
 * Line3D::performRDD hands to the GPU (src/line3D.cc:2453): entries float4{i, j, w / normalization, 0}
 * sorted by (column, row) -- or (row, column) with sort_by_row -- and per column (row) the index of
 * its first entry, -1 if it has none (num_local_ids of them).  Built and kept on the device
 * (l3d_get_sparse_device: device pointers, valid until the next l3d_affinity* call); the host copies
 * are made when the output pointers are not NULL. */
int l3d_affinity_sparse(l3d_ctx* ctx, int sort_by_row, float normalization_factor, float* entries_xyzw,
                        int32_t* start_indices, uint32_t cap_entries, uint32_t cap_rows);
int l3d_get_sparse_device(l3d_ctx* ctx, const void** entries_float4, const void** start_indices_int);

/* Unchanged consumer, provided for convenience: Felzenszwalb-Huttenlocher clustering of A_
 * exactly as L3DPP::performClustering (src/clustering.cc:7-48, include/universe.h:59-117), on
 * the host. */
int l3d_cluster(l3d_ctx* ctx);
/* stand-alone form of the same routine (edges: ne x (i,j), weights: ne, out: n root ids) */
int l3d_cluster_edges(const int32_t* edges_ij, const float* weights, uint32_t ne, uint32_t n,
                      int32_t* out_root);

/* replaces the rest of Line3D::reconstruct3Dlines after the clustering (src/line3D.cc:2115-2141): for every cluster
 * seen by >= visibility_t cameras (at least 3) Line3D::get3DlineFromCluster (src/line3D.cc:2578-2641),
 * Line3D::findCollinearSegments_return (:2763-2870, with project2DsegmentOnto3Dline :2644-2687),
 * Line3D::filterTinySegments (:2724-2760) and the translation back, on the device (one thread per cluster).
 * Needs l3d_cluster.  The lines are kept in the context:
 *   l3d_get_lines3D_counts  counts3 = {lines, 3-D segments in total, 2-D residuals in total}
 *   l3d_get_lines3D         seg_off[lines+1], segs6[6 x segments] (P1, P2), res_off[lines+1],
 *                           res2[2 x residuals] (camera id, segment id), ref_cam[lines] (LineCluster3D::reference_view)
 *   l3d_save_lines3D_txt    Line3D::save3DLinesAsTXT (src/line3D.cc:3122-3178) into the file `path` */
int l3d_lines3D(l3d_ctx* ctx, uint32_t visibility_t);
int l3d_get_lines3D_counts(l3d_ctx* ctx, uint32_t* counts3);
int l3d_get_lines3D(l3d_ctx* ctx, uint32_t* seg_off, double* segs6, uint32_t* res_off, uint32_t* res2, uint32_t* ref_cam);
int l3d_save_lines3D_txt(l3d_ctx* ctx, const char* path);

/* results (host pointers) */
int l3d_get_counts(l3d_ctx* ctx, l3d_counts* out);
int l3d_reset_counters(l3d_ctx* ctx);
int l3d_get_timings(l3d_ctx* ctx, float* ms, uint32_t n);
int l3d_get_pairs(l3d_ctx* ctx, uint32_t* src_tgt_cam_ids, uint32_t cap_pairs);
/* which: 0 = lists as they were right after scoring (needs keep_scored), 1 = filtered lists
 * (= matches_[cam] after filterMatches). row_off: num_segs+1. Returns L3D_ERR_CAPACITY with the
 * needed record count in *out_count if cap is too small. */
int l3d_get_view_lists(l3d_ctx* ctx, uint32_t cam_id, int which, uint32_t* row_off,
                       l3d_list_rec* recs, uint64_t cap, uint64_t* out_count);
int l3d_get_entries(l3d_ctx* ctx, l3d_entry* out, uint32_t cap);
int l3d_get_edges(l3d_ctx* ctx, int32_t* edges_ij, float* weights, uint32_t cap);
int l3d_get_local2global(l3d_ctx* ctx, uint32_t* cam_seg, uint32_t cap);
int l3d_get_cluster_ids(l3d_ctx* ctx, int32_t* out, uint32_t cap);
/* info[0..2] = camera centre C (current, untranslated), kmm = {k, median_depth, median_sigma} */
int l3d_get_view_info(l3d_ctx* ctx, uint32_t cam_id, double* C, float* kmm);
int l3d_get_med_scene_depth_lines(l3d_ctx* ctx, float* out);

/* Multi-GPU (one process per GPU; the collective itself is torch.distributed / NCCL).  Reference
 * views are split into `shard_world` contiguous slices (l3d_params); a rank matches the pairs whose
 * source view it owns, builds / finishes the scoring rows and the affinity edges of its views, and
 * four exchanges make the results whole again on every rank:
 *   l3d_match_stage12 -> FORWARD (+ forward records all-to-all) -> l3d_score_build -> PROGRAMS -> l3d_score_fold -> HYPOTHESES
 *   -> l3d_affinity_edges -> EDGES -> l3d_affinity_ids (-> l3d_cluster).
 * Per exchange: every rank calls l3d_shard_blob_size + l3d_shard_export, the blobs are all-gathered
 * `stride` bytes apart (stride >= the largest size, multiple of 32), every rank calls
 * l3d_shard_import with all blobs and all sizes.  `device_ptr` != 0: device pointers (they must
 * stay valid until the next exchange of the same kind).  With one rank l3d_match_stage3 /
 * l3d_affinity run the same phases back to back. */
enum { L3D_X_FORWARD = 0, L3D_X_PROGRAMS = 1, L3D_X_HYPOTHESES = 2, L3D_X_EDGES = 3 };
int l3d_score_build(l3d_ctx* ctx);
int l3d_score_fold(l3d_ctx* ctx);
int l3d_affinity_edges(l3d_ctx* ctx);
int l3d_affinity_ids(l3d_ctx* ctx);
int l3d_shard_blob_size(l3d_ctx* ctx, int kind, uint64_t* bytes);
int l3d_shard_export(l3d_ctx* ctx, int kind, void* dst, uint64_t cap_bytes, int device_ptr);
int l3d_shard_import(l3d_ctx* ctx, int kind, const void* all_blobs, uint64_t stride_bytes, int world,
                     const uint64_t* sizes, int device_ptr);
/* FORWARD carries the per-row match counts only.  The match RECORDS of a boundary pair (target view in another
 * slice) are needed by one rank, the owner of the target view, so they travel all-to-all right after the FORWARD
 * exchange: every rank calls l3d_shard_forward_plan (records it sends to / receives from every peer), packs its
 * records grouped by destination rank, the groups are exchanged (all-to-all-v, 32 bytes per record), and every rank
 * unpacks what it received (grouped by source rank, ascending). */
int l3d_shard_forward_plan(l3d_ctx* ctx, uint64_t* send_records, uint64_t* recv_records);
int l3d_shard_forward_pack(l3d_ctx* ctx, void* dst, uint64_t cap_bytes, int device_ptr);
int l3d_shard_forward_unpack(l3d_ctx* ctx, const void* src, uint64_t bytes, int device_ptr);
/* Steady-state variant without the size exchange (device pointers only): the blob carries a 32-byte
 * header with its own size, written by the device, so the sender never waits for its cursors; the
 * stride comes from the previous step.  *redo != 0 after the import: a blob did not fit -- nothing
 * was imported and every rank repeats the exchange with the three calls above. */
int l3d_shard_export_hdr(l3d_ctx* ctx, int kind, void* dst, uint64_t stride_bytes);
int l3d_shard_import_hdr(l3d_ctx* ctx, int kind, const void* all_blobs, uint64_t stride_bytes, int world,
                         uint64_t* sizes_out, int* redo);

/* deterministic device math exposed for parity tests (n values, host pointers) */
int l3d_test_expf(l3d_ctx* ctx, const float* x, float* y, uint32_t n);
int l3d_test_acos(l3d_ctx* ctx, const double* x, double* y, uint32_t n);
/* peak-FP32 micro-benchmark used by bench.py for the roofline denominator (TFLOP/s) */
int l3d_bench_fp32_peak(l3d_ctx* ctx, float* tflops);
/* the same for the FP64 pipe (DFMA chains): the denominator of the K2 roofline */
int l3d_bench_fp64_peak(l3d_ctx* ctx, float* tflops);

#ifdef __cplusplus
}
#endif
#endif /* L3DPP_B200_H_ */
