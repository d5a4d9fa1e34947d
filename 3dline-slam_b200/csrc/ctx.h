// ctx.h -- the context object behind the C ABI (private to libl3dpp_b200.so)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/l3dpp_b200.h"
#include "host_geom.h"
#include "internal.h"

namespace l3d {
// launchers defined in the other translation units
int k1_rows_per_cta();
int launch_k2_exact(const PairDev*, const K1Cta*, uint32_t, uint32_t, uint32_t, const float4*, const SegRays*,
                    const double*, const SegPlane*, const SegV32*, const SegDesc*, const RowEpi32*, const uint32_t*, const uint32_t*,
                    const ViewDev*, const uint32_t*, const uint32_t*, unsigned long long*, FwdRec*, FwdRec*, uint32_t*, uint32_t*, uint32_t*,
                    uint2*, uint32_t*, float*, float, int, int, int, int, int*, cudaStream_t);
int launch_k2_compact(const uint32_t*, const uint32_t*, const uint32_t*, uint32_t, const FwdRec*, FwdRec*, uint32_t*,
                      uint32_t, const uint32_t*, cudaStream_t);
int launch_k3_score(uint32_t, const uint32_t*, ListRec*, const ListGeo*, float, float, void*, cudaStream_t);
int launch_k3_inv_capacity(const PairDev*, uint32_t, uint32_t, const uint32_t*, const uint32_t*, const FwdRec*,
                           uint32_t*, uint32_t*, uint32_t, uint32_t, cudaStream_t);
int launch_fwd_bmask(const PairDev*, uint32_t, const uint32_t*, uint32_t, uint32_t, uint32_t*, cudaStream_t);
int launch_fwd_bgather(const uint32_t*, const uint32_t*, const uint32_t*, uint32_t, uint32_t, const FwdRec*, FwdRec*,
                       cudaStream_t);
int launch_fwd_place(const unsigned char*, uint64_t, int, const uint32_t*, int, uint32_t, const uint32_t*,
                     const uint32_t*, const uint32_t*, const uint32_t*, const FwdRec*, const uint32_t*, FwdRec*,
                     cudaStream_t);
int launch_rec_blocks(const uint4*, const uint32_t*, const uint32_t*, uint32_t, const FwdRec*, FwdRec*, cudaStream_t);
int launch_fwd_move(const uint32_t*, uint32_t, uint32_t, const uint32_t*, const FwdRec*, const uint32_t*, FwdRec*,
                    cudaStream_t);
int launch_k3_list_capacity(const ViewDev*, const uint32_t*, uint32_t, const IncDev*, const uint32_t*, const PairDev*,
                            const uint32_t*, const uint32_t*, uint32_t*, void*, cudaStream_t);
int launch_k3_records(const PairDev*, uint32_t, uint32_t, const uint32_t*, const FwdRec*, float*, const uint32_t*,
                      uint32_t*, uint2*, uint32_t, uint32_t, cudaStream_t);
size_t k3_wf_stats_bytes();
size_t k3_sib_bytes();
int k3_wf_max_inc();
int k3_max_staged();
size_t k3_stats_bytes();
int launch_k4_has(const EntryDev*, uint32_t, uint32_t*, cudaStream_t);
int launch_k4_median(ViewDev*, uint32_t, const EntryDev*, uint32_t*, cudaStream_t);
int launch_k4_edges_count(const ViewDev*, const uint32_t*, const EntryDev*, uint32_t, const uint32_t*, const uint32_t*,
                          const ListRec*, float, float, float*, uint32_t*, unsigned long long*, uint32_t, uint32_t,
                          cudaStream_t);
int launch_k4_edges_write(const ViewDev*, uint32_t, const uint32_t*, const uint32_t*, const ListRec*, const float*,
                          const uint32_t*, void*, uint32_t, uint32_t, cudaStream_t);
int launch_k4_ids(const void*, uint32_t, uint32_t*, uint32_t*, uint32_t*, uint32_t*, size_t, int2*, float*,
                  uint32_t*, cudaStream_t);
size_t k4_edge_bytes();
int launch_k5_collinear(const float4*, uint32_t, float, char*, size_t, cudaStream_t);
int launch_k4_sparse(const int2*, const float*, uint32_t, uint32_t, int, float, uint32_t*, uint32_t*, uint32_t*, uint2*,
                     uint32_t*, size_t, float4*, int*, cudaStream_t);
int launch_test_expf(const float*, float*, uint32_t, cudaStream_t);
int launch_test_acos(const double*, double*, uint32_t, cudaStream_t);
int launch_fp32_peak(float*, int, int, cudaStream_t);
int launch_fp64_peak(double*, int, int, cudaStream_t);
int launch_score_prep(const float4*, uint32_t, const float4*, const float2*, const double*, const double*, float,
                      const uint32_t*, ListRec*, ListGeo*, cudaStream_t);
}  // namespace l3d

using namespace l3d;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(L3D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

// growable device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // capacity >= n; contents are NOT preserved unless keep (number of elements to keep) is given
    cudaError_t ensure(size_t n, size_t keep = 0, cudaStream_t st = 0)
    {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 2);
        T* np = nullptr;
        cudaError_t e = cudaMalloc((void**)&np, std::max<size_t>(ncap, 1) * sizeof(T));
        if (getenv("L3D_MEM_DEBUG") && ncap * sizeof(T) >= ((size_t)256 << 20))   // where the HBM goes (capacity planning)
            fprintf(stderr, "[mem] %.2f GB = %zu x %zu B (%s)\n", ncap * sizeof(T) / 1073741824.0, ncap, sizeof(T),
                    __PRETTY_FUNCTION__);
        if (e != cudaSuccess) return e;
        if (keep && p) {
            e = cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
        }
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
};

// stream mode: the tables grow a little every cycle; reallocate rarely (twice the need, never below
// `floor` elements) so that a steady-state cycle does not stall in cudaMalloc / cudaFree
template <typename T>
static inline cudaError_t ensure_roomy(DevBuf<T>& b, size_t n, size_t floor, size_t keep = 0, cudaStream_t st = 0)
{
    if (n <= b.cap) return cudaSuccess;
    return b.ensure(std::max(2 * n, floor), keep, st);
}

struct HostView {
    l3d_view v;
    std::vector<float> segs;
    const float* ext_segs = nullptr;  // l3d_scene_set: the caller's buffer, valid until the commit inside the same call
    std::vector<uint32_t> nbrs;
    std::vector<uint32_t> nb_views;  // neighbours as ascending view indices (l3d_scene_commit / plan_pairs)
    std::vector<uint32_t> wps;       // observed world points (neighbors_by_worldpoints mode)
    hg::Camera cam;
    float k = 0.0f, median_depth = 0.0f, median_sigma = 0.0f;
    uint32_t seg_off = 0;
    // key-frame stream mode (stream.cu): views are never removed from the table (Line3D::deleteImage
    // keeps views_[camID], src/line3D.cc:396-430)
    bool current = true;       // in view_order_ / views_reserved_
    bool processed = false;    // processed_[camID]
    bool uploaded = false;     // segments resident on the device
    uint32_t filt_total = 0;   // entries that survived the view's last filterMatches
    uint32_t num_wps = 0;      // num_worldpoints_[camID] (kept across cycles)
    bool has_fixed = false;    // fixed_visual_neighbors_ holds the view (set by UpdataImage)
    bool paired_after_delete = false;  // a later key frame named this deleted view as a neighbour
};

struct HostPair {
    uint32_t src, tgt;  // view indices
    uint32_t batch;
    uint64_t fwd_total = 0;  // forward records of this pair
    uint32_t rec_start = 0;  // first forward record (canonical layout)
    bool local = true;       // matched by this shard
};

struct Batch {
    uint32_t pair0, pair1;  // [pair0, pair1)
    uint32_t row0, n_rows;
    uint32_t cta0, n_ctas;
    uint64_t mask_words;
    uint32_t max_tgt;  // largest target view of the batch's pairs (chooses the K2 variant)
};

struct StageTimer {
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
    float ms[L3D_T_COUNT] = {0};
    ~StageTimer() { reset(); }  // events of a step that returned early
    void reset()
    {
        for (auto& e : ev) {
            cudaEventDestroy(e.second.first);
            cudaEventDestroy(e.second.second);
        }
        ev.clear();
        memset(ms, 0, sizeof(ms));
    }
    cudaEvent_t begin(int id, cudaStream_t st)
    {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        ev.push_back({id, {a, b}});
        return b;
    }
    void end(cudaEvent_t b, cudaStream_t st) { cudaEventRecord(b, st); }
    void collect()
    {
        for (auto& e : ev) {
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, e.second.first, e.second.second) == cudaSuccess) ms[e.first] += t;
            cudaEventDestroy(e.second.first);
            cudaEventDestroy(e.second.second);
        }
        ev.clear();
    }
};

#define L3D_MAX_WORLD_C 16
struct l3d_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    l3d_params prm{};
    bool have_params = false;
    bool fixed3D = false;  // sigma_p given in world units (fixed3Dregularizer_, src/line3D.cc:525-530)

    // host scene
    std::vector<HostView> views;  // sorted by cam id at commit
    std::map<uint32_t, uint32_t> cam2view;
    bool committed = false;
    bool by_worldpoints = false;  // neighbours are chosen from world-point lists at match time
    uint32_t S = 0;  // total segments
    hg::V3 translation{0, 0, 0};
    bool translated = false;  // the host cameras currently carry -translation (enter_/leave_translated)
    float two_sigA_sqr = 200.0f, epi_overlap = 0.25f;
    float med_scene_depth_lines = 0.0f;

    std::vector<HostPair> pairs;
    std::vector<PairDev> pairs_h;
    std::vector<Batch> batches;
    std::vector<K1Cta> ctas_h;
    std::vector<uint32_t> inc_off_h;  // per view
    std::vector<IncDev> inc_h;
    uint32_t total_rows = 0, total_tgt_rows = 0;
    uint64_t total_fwd = 0;
    int stage = 0;  // 0: scene, 1: stage12 done, 2: stage3 done, 3: affinity done, 4: clustered
    bool raw_mode = false;        // l3d_match_lines: cameras given as (RtKinv, C), F given, no translation
    double F_override[9] = {0};

    void* pinned = nullptr;  // pinned host staging of the segment upload
    size_t pinned_cap = 0;
    // pinned read-back scratch: the small device->host copies a step synchronises on (counts, cursors,
    // per-pair totals, the view table) land here -- a pageable target costs an extra staging hop each time
    unsigned char* rb = nullptr;  // allocated by l3d_ctx_create; a context built on the stack (l3d_match_lines)
    size_t rb_cap = 0;            // reads back into the pageable fallback below
    unsigned char rb_fallback[4096] = {0};
    enum { RB_NCAND = 0, RB_NFIN = 8, RB_DEVMAX = 16, RB_SMALL = 32, RB_NENT = 64, RB_STATS = 128, RB_NEDGES = 512,
           RB_NLOCAL = 520, RB_K1RUN = 528, RB_BIG = 4096 };
    bool k1_run_pending = false;
    template <typename T>
    T* rb_at(size_t off) { return reinterpret_cast<T*>((rb ? rb : rb_fallback) + off); }
    bool rb_fits(size_t bytes) const { return rb && RB_BIG + bytes <= rb_cap; }
    l3d_ctx* pair_scratch = nullptr;  // two-view scratch context of l3d_match_lines (owned)
    ~l3d_ctx()
    {
        delete pair_scratch;
        if (pinned) cudaFreeHost(pinned);
        if (rb) cudaFreeHost(rb);
    }

    // device tables
    DevBuf<float4> d_segs;
    DevBuf<uint32_t> d_seg_view;
    DevBuf<SegDesc> d_desc;
    DevBuf<SegRays> d_rays;
    DevBuf<double> d_midray;
    DevBuf<SegPlane> d_planes;
    DevBuf<SegV32> d_v32;          // FP32 image of rays / planes (K2's certified depth-sign test)
    DevBuf<RowEpi32> d_row_epi;    // K1's per-row epipolar lines of the current batch, in K1's sorted row order
    DevBuf<RowEpi32> d_row_epi_nat;  // the same in natural row order (staging of the row sort)
    DevBuf<float2> d_row_key;      // direction keys of the row's two lines (sorted row order)
    DevBuf<uint32_t> d_perm, d_iperm;  // sorted position -> natural row and back (batch-local)
    DevBuf<uint2> d_l2g_cs;  // local2global_ as (camera id, segment) pairs (l3d_get_local2global)
    DevBuf<unsigned long long> d_k1_run;  // pair tests K1 evaluated (l3d_counts::pair_tests_run)
    DevBuf<uint32_t> d_ncont, d_k2ctr;  // K2: contenders per batch row; {work items, fallback rows}
    DevBuf<uint2> d_fb_rows;       // K2: rows handed to the literal row kernel
    DevBuf<uint32_t> d_row_pair;   // K2: pair of every batch row
    DevBuf<float> d_row_T;         // K2: kNN-th largest certain lower bound of every batch row
    int n_sm = 0;
    DevBuf<float> d_view_xb;
    DevBuf<ViewDev> d_views;
    DevBuf<PairDev> d_pairs;
    DevBuf<K1Cta> d_ctas;
    DevBuf<IncDev> d_inc;
    // stage 1/2 scratch
    DevBuf<uint32_t> d_mask, d_cand_cnt, d_cand_off, d_fin_cnt, d_fin_off, d_scan, d_k3_cls;
    DevBuf<unsigned long long> d_heap;
    DevBuf<FwdRec> d_cand_rec, d_fin_rec;
    // forward store
    DevBuf<FwdRec> d_fwd_rec;
    DevBuf<uint32_t> d_fwd_off, d_fwd_cnt;
    // stage 3
    DevBuf<uint32_t> d_inv_cap, d_inv_fill, d_inv_off, d_scan_tmp, d_inc_off, d_view_max;
    DevBuf<uint2> d_inv_ent;
    DevBuf<uint32_t> d_L_ub, d_L_off, d_L_cnt;
    DevBuf<ListRec> d_L_rec;
    DevBuf<unsigned char> d_L_sib;
    DevBuf<double> d_L_dir;
    DevBuf<float2> d_L_reg;
    DevBuf<uint32_t> d_L_f, d_L_c, d_L_h, d_fwd_row, d_prog_off, d_prog_nh;
    DevBuf<unsigned char> d_L_meta, d_prog;
    DevBuf<float> d_L_score, d_fwd_score;
    uint64_t prog_cap = 0;  // fold-program store, 16-byte units (grown on overflow)
    uint64_t L_total = 0, filt_cap = 0;
    uint32_t k3_maxm = 0;
    uint32_t k3_list_max = 0;   // longest potential list of the committed scene (0: not read yet)
    bool force_list_max = false;
    bool k3_big_rows = false;
    int stage3_phase = 0, stage4_phase = 0;
    cudaEvent_t ev_total3 = nullptr, ev_total4 = nullptr;

    // sharding: contiguous view slices (see plan_pairs), buffers adopted from the exchanges
    int world = 1, rank = 0;
    std::vector<uint32_t> slice_view{0, 0}, slice_g{0, 0}, slice_row{0, 0};
    std::vector<uint32_t> view_needed;  // per view: some pair incident to this rank's slice touches it
    const void* prog_all = nullptr;      // all-gathered fold programs (caller-owned until the next exchange)
    const ListRec* filt_all = nullptr;   // filtered lists of every slice (d_filt_all)
    const void* edges_all = nullptr;     // edges of every slice in traversal order (d_edges_all)
    DevBuf<ListRec> d_filt_all;
    DevBuf<unsigned char> d_edges_all;
    DevBuf<FwdRec> d_fwd_alt, d_bx_rec;
    DevBuf<uint32_t> d_bx_cnt, d_bx_off, d_bx_cnt_all, d_bx_off_all, d_fwd_off_local;
    uint64_t pair_total_sum = 0;  // forward records over all pairs (refresh_pair_totals)
    uint32_t n_edges_local = 0, n_edges_all = 0;
    uint64_t local_fwd = 0;      // forward records produced by this rank's pairs
    uint64_t xchg_var[4] = {0, 0, 0, 0};
    uint64_t fwd_send[L3D_MAX_WORLD_C] = {0}, fwd_recv[L3D_MAX_WORLD_C] = {0};  // FORWARD records per peer (forward_plan)
    bool fwd_planned = false;
    DevBuf<uint4> d_blk_items;
    DevBuf<uint32_t> d_blk_chunks;
    DevBuf<unsigned char> d_xchg_stage[4];
    DevBuf<uint32_t> d_slice_g;
    uint64_t shard_sim_evals = 0, shard_scored = 0, shard_filtered = 0;
    std::vector<uint64_t> L_cap_h;   // per-view list capacity
    DevBuf<ListRec> d_filt_rec;
    DevBuf<uint32_t> d_filt_off, d_filt_cnt, d_small;  // d_small: [0]=filt_total [1]=err [2]=median overflow
    DevBuf<unsigned char> d_stats;                     // 2 x ScoreStats
    DevBuf<EntryDev> d_entries;
    DevBuf<uint32_t> d_has, d_entry_idx;
    std::vector<uint64_t> L_base_h;  // per-view base into the list buffers
    // stage 4
    DevBuf<float> d_filt_sim;
    DevBuf<uint32_t> d_E_cnt, d_E_off, d_first_touch, d_flags, d_flag_scan, d_l2g;
    DevBuf<unsigned char> d_edges;
    DevBuf<int2> d_A_ij;
    DevBuf<float> d_A_w;
    DevBuf<unsigned long long> d_tests;
    // SparseMatrix layout of A_ (l3d_affinity_sparse)
    DevBuf<uint32_t> d_sp_hist, d_sp_off, d_sp_fill;
    DevBuf<uint2> d_sp_tmp;
    DevBuf<float4> d_sp_entries;
    DevBuf<int> d_sp_start;
    bool sparse_ready = false;
    DevBuf<char> d_collin;
    DevBuf<float4> d_collin_lines;
    // host results
    std::vector<int32_t> cluster_ids;
    // cluster -> 3-D line tail (l3d_lines3D): the cameras as Line3D::reconstruct3Dlines sees them between translate()
    // and untranslate(), the translation they were shifted by, and the final lines (flat, reference order)
    std::vector<TailView> tail_views;
    hg::V3 tail_translation{0, 0, 0};
    std::vector<uint32_t> l3_seg_off, l3_res_off, l3_res, l3_ref_cam;
    std::vector<double> l3_segs;
    bool lines_ready = false;
    DevBuf<uint32_t> d_t_cl_off, d_t_members, d_t_ord, d_t_camtab, d_t_out_n, d_t_out_ref;
    DevBuf<double> d_t_L, d_t_LC, d_t_pts, d_t_out_seg;
    DevBuf<float> d_t_dist;
    DevBuf<unsigned char> d_t_ok;
    DevBuf<TailView> d_t_views;

    // key-frame stream mode (stream.cu)
    bool stream_mode = false;
    std::set<uint32_t> st_add, st_del;                   // Add_camID_ / Delete_camID_ as view indices
    std::set<std::pair<uint32_t, uint32_t>> st_matched;  // matched_ (unordered view pairs)
    uint32_t st_cycle = 0;
    uint64_t st_w_extent = 0, st_f_extent = 0;  // extents of this cycle's working / filtered arenas
    DevBuf<ListRec> d_st_filt_old, d_st_W_rec;
    DevBuf<unsigned char> d_st_W_geo, d_st_pairs, d_st_vflag, d_st_stats;
    DevBuf<uint32_t> d_st_W_row, d_st_I_cnt, d_st_I_off, d_st_I_fill, d_st_I_key, d_st_W_cnt, d_st_W_off, d_st_F_cnt,
        d_st_F_off, d_st_best, d_st_view_total, d_st_row_g, d_st_vout;

    l3d_counts cnt{};
    StageTimer tm;
};


// shared between ctx.cu and abi.cu
int refresh_pair_totals(l3d_ctx* ctx);
int plan_pairs(l3d_ctx* ctx);
void plan_batches(l3d_ctx* ctx);
int run_stage12_batches(l3d_ctx* ctx);
void compute_translation(l3d_ctx* ctx);
void apply_translation(l3d_ctx* ctx, double sign);
void enter_translated(l3d_ctx* ctx);
void leave_translated(l3d_ctx* ctx);
int stream_match_images(l3d_ctx* ctx, const l3d_params* params);
int upload_views(l3d_ctx* ctx);
int set_params(l3d_ctx* ctx, const l3d_params* params);
int score_rebuild(l3d_ctx* ctx, uint32_t needed_units);
int score_hypotheses_ready(l3d_ctx* ctx, bool* prog_overflow, uint32_t* prog_needed);
#define L3D_MAX_WORLD 16
