// internal.h -- device-side table layouts and kernel launchers shared by the translation units of
// libl3dpp_b200.so.  Data layout in HBM:
//   segs      float4[S]        x1,y1,x2,y2 of every 2-D segment, views concatenated       16 B/seg
//   desc      SegDesc[S]       K1 target-side descriptor (1-D parametrisation)           32 B/seg
//   rays      SegRays[S]       normalised viewing rays of both endpoints (double)        48 B/seg
//   midray    double[3][S]     normalised ray through the 2-D midpoint (AoS double3)      24 B/seg
//   planes    SegPlane[S]      normal of the plane (centre, r1, r2) and n.C                  32 B/seg
//   views     ViewDev[V]
//   pairs     PairDev[P]       matched view pairs in reference order (src asc, tgt asc)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l3d {

struct __align__(16) SegDesc {
    float q1x, q1y, ux, uy;  // first endpoint and unit direction of the segment
    float L, slo, shi, g;    // length, admissible parameter interval (image bounds), guard term
};
static_assert(sizeof(SegDesc) == 32, "SegDesc must be 32 bytes");

struct __align__(16) SegRays {  // 16-byte aligned: gathered with three 128-bit loads per lane
    double r1[3], r2[3];
};

// plane through the camera centre spanned by the two endpoint rays (Line3D::triangulationDepths,
// src/line3D.cc:1373-1378): unit normal and its dot product with the (translated) camera centre
struct __align__(16) SegPlane {
    double n[3];
    double cn;
};

// FP32 image of the triangulation tables of one segment (K2's certified depth-sign test): plane normal and
// n.C, endpoint rays; rounded once from the double tables of k0_prep
struct __align__(16) SegV32 {
    float nx, ny, nz, cn;
    float r1x, r1y, r1z, r2x;
    float r2y, r2z, pad0, pad1;
};
static_assert(sizeof(SegV32) == 48, "SegV32 must be 48 bytes");

// what K1 knows about a source row and K2's FP32 ranking re-uses: the normalised epipolar lines of both
// endpoints, the numerator error bound and the smaller of the two line norms (batch-local row index)
struct __align__(16) RowEpi32 {
    float A1, B1, C1, A2;
    float B2, C2, cN, nmin;
};
static_assert(sizeof(RowEpi32) == 32, "RowEpi32 must be 32 bytes");

struct ViewDev {
    double C[3];
    double RtKinv[9];
    float k;
    float median_depth;
    uint32_t seg_off, n_seg;
    uint32_t cam_id;
    float xb;  // max |x1|+|y1| over the view's segments (K1 guard)
    uint32_t order;
    uint32_t needed;  // 0: no pair of this rank touches the view (its per-segment tables are skipped)
};

struct PairDev {
    double F[9];
    uint32_t src_view, tgt_view;
    uint32_t src_off, n_src;
    uint32_t tgt_off, n_tgt;
    uint32_t row_base;   // global row index of (pair,0): prefix sum of n_src over pairs
    uint32_t tgt_base;   // prefix sum of n_tgt over pairs (inverse-match CSR rows)
    uint32_t words;      // ceil(n_tgt/32)
    uint32_t emit_inverse;  // tgt view is processed after src view (line3D.cc:1994)
    uint64_t mask_base;  // word offset of this pair's bit mask inside the batch buffer
    uint32_t batch_row0; // first row of the batch this pair belongs to
    uint32_t xflag;      // multi-GPU: 0, or 1 + the rank owning the target view when it is in another rank's slice
};

// forward match record, 32 B (matches_ entries produced by matching, line3D.cc:1169-1181)
struct __align__(16) FwdRec {
    uint32_t c;       // tgt segment
    float overlap;
    float d_p1, d_p2, d_q1, d_q2;
    float score;
    uint32_t flags;
};
static_assert(sizeof(FwdRec) == 32, "FwdRec must be 32 bytes");

// list entry at scoring / after filtering, 40 B (8-byte aligned: five 64-bit loads per record)
struct __align__(8) ListRec {
    uint32_t tgt_view, tgt_seg;
    float overlap, score;
    float d_p1, d_p2, d_q1, d_q2;
    uint32_t flags;    // bit0 orientation flag (inverse matches), bit1: inverse
    uint32_t src_idx;  // forward-record index this entry came from (score write-back), or ~0u
};
static_assert(sizeof(ListRec) == 40, "ListRec must be 40 bytes");

// per-entry geometry staged for scoring, 48 B
struct ListGeo {
    double dir[3];
    float reg1, reg2;
    float length;
    uint32_t run;  // 1 if this entry starts a new target-camera run
    uint32_t pad0, pad1;
};

// best hypothesis of a segment (estimated_position3D_ row), indexed by global segment id
struct EntryDev {
    double P1[3], P2[3], dir[3];
    float length;
    uint32_t tgt_view, tgt_seg;
    float overlap, score;
    float d_p1, d_p2, d_q1, d_q2;
    uint32_t has;  // 1 if the segment has a best match with score > 0.75
};

// a view as the cluster -> 3-D line tail needs it (k6_lines3d.cu): K, R, t and centre in the translated frame
// of Line3D::reconstruct3Dlines, View::min_line_length_ (src/view.cc:33: diagonal * 0.005)
struct TailView {
    double K[9], R[9], t[3], C[3];
    float min_line_length;
    uint32_t cam_id;
};

struct K1Cta {
    uint32_t pair, tile;
};

struct IncDev {  // one incident pair of a view, in list order
    uint32_t pair;
    uint32_t inverse;  // 1: this view is the pair's target (entries come from the inverse CSR)
};

// device tables of the scoring stage (K3), filled from the context by ctx.cu
struct K3Tables {
    const ViewDev* views; const uint32_t* seg_view; const PairDev* pairs; const IncDev* inc; const uint32_t* inc_off;
    const SegRays* rays; const uint32_t* fwd_off; const uint32_t* fwd_cnt; FwdRec* fwd_rec; float* fwd_score; const uint32_t* fwd_row;
    const uint32_t* inv_off; const uint32_t* inv_fill; const uint2* inv_ent;
    const uint32_t* L_off; uint32_t* L_f; unsigned char* L_meta; float* L_score;
    void* L_sib; double* L_dir; float2* L_reg; uint32_t* L_c; uint32_t* L_h;
    uint32_t* prog_off; uint32_t* prog_nh; void* prog; uint32_t prog_cap;
    uint32_t* L_cnt; ListRec* L_rec; uint32_t* view_max; ListRec* filt_rec; uint32_t filt_cap;
    uint32_t* filt_off; uint32_t* filt_cnt; EntryDev* entries; void* stats;
    uint32_t S, maxm; float two_sigA_sqr;
    uint32_t g_lo, g_hi;  // rows (global segment indices) this rank builds and finishes
    uint32_t* cls;        // build: 16 counters + 3 row lists (k3_class_words(g_hi - g_lo) words)
};
size_t k3_class_words(uint32_t rows);
int launch_k3_build(const K3Tables& t, cudaStream_t st, int* err);
int launch_k3_fold(const K3Tables& t, cudaStream_t st, int* err);
int launch_k3_finish(const K3Tables& t, cudaStream_t st);
int launch_k3_adopt_programs(const void* all, uint64_t stride, int world, const uint32_t* slice_g, uint32_t S,
                             uint32_t* prog_off, uint32_t* prog_nh, float* fwd_score, cudaStream_t st);

// key-frame stream mode (k3_stream.cu): a new view pair as seen from one of its views -- incoming
// (the view is the pair's target and receives inverse matches) or outgoing (the view is the source)
struct StreamPair {
    uint32_t rec_start, rec_cnt;  // forward records of the pair
    uint32_t row_base, n_src;     // pair rows
    uint32_t other;               // the other view
    uint32_t pad;
};

// one step of the key-frame stream walk (k3_stream.cu): the rows of all previously scored views, or
// the rows of one new view
struct StreamStepArgs {
    uint32_t n, n_in, in0, in_total, w_base, f_base, w_cap, f_cap;
    const uint32_t* row_g; const uint32_t* seg_view;
    const void* pairs;  // StreamPair[] on the device (all descriptors of the cycle)
    const uint32_t* vout0; const uint32_t* vnout;
    const FwdRec* fwd_rec; float* fwd_score; const uint32_t* fwd_off; const uint32_t* fwd_cnt;
    uint32_t* I_off; uint32_t* I_cnt; uint32_t* I_fill; uint32_t* I_key;
    uint32_t* scan; size_t scan_words;
    uint32_t* filt_off; uint32_t* filt_cnt; ListRec* filt_old; ListRec* filt_new;
    uint32_t* W_cnt; uint32_t* W_off; ListRec* W_rec; uint32_t* W_row; ListGeo* W_geo;
    uint32_t* L_off; uint32_t* L_cnt;
    const ViewDev* views; const SegRays* rays; const double* midray; const unsigned char* vflag;
    uint32_t* view_max; uint32_t* F_cnt; uint32_t* F_off; uint32_t* best_e;
    EntryDev* entries; uint32_t* view_total; void* stats;
    float two_sigA_sqr;
};
int launch_stream_step(const StreamStepArgs& a, cudaStream_t st);
int launch_stream_update_entries(uint32_t S, const uint32_t* seg_view, const ViewDev* views, const SegRays* rays,
                                 const SegPlane* planes, EntryDev* entries, cudaStream_t st);
size_t stream_stats_bytes();

// ---------------- launchers (each returns the number of kernels launched) ----------------
int launch_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch,
                    size_t scratch_words, cudaStream_t st);
size_t scan_scratch_words(uint32_t n);

int launch_k0_prep(const float4* segs, const uint32_t* seg_view, const ViewDev* views, uint32_t S,
                   int max_image_width, SegDesc* desc, SegRays* rays, double* midray, SegPlane* planes,
                   SegV32* v32, float* view_xb,
                   cudaStream_t st);

int launch_k6_lines3d(uint32_t n_clusters, const uint32_t* cl_off, const uint32_t* members, const EntryDev* entries,
                      const float4* segs, const SegRays* rays, const uint32_t* seg_view, const TailView* views,
                      const double* t3, double* Lbuf, double* LCbuf, double* pts, float* dist, uint32_t* ord,
                      unsigned char* okflag, uint32_t* camtab, uint32_t* out_n, uint32_t* out_ref, double* out_seg,
                      cudaStream_t st);

int launch_l2g_camseg(const uint32_t* l2g, uint32_t n, const uint32_t* seg_view, const ViewDev* views, uint2* out,
                      cudaStream_t st);

// K1: rows of every pair sorted by the direction of their epipolar lines (k1_pairtest.cu); epi_rho / key_rho /
// the mask are in sorted ("rho") order, perm[rho] = natural row, iperm[natural row] = rho (batch-local indices)
int launch_k1_rowsort(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, const float4* segs,
                      const float* view_xb, RowEpi32* epi_nat, RowEpi32* epi_rho, float2* key_rho, uint32_t* perm,
                      uint32_t* iperm, cudaStream_t st);
int launch_k1_pairtest(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, const SegDesc* desc,
                       const RowEpi32* epi_rho, const float2* key_rho, const uint32_t* perm, uint32_t* mask,
                       uint32_t* cand_cnt, float thr, int filter_mode, unsigned long long* tests_run,
                       cudaStream_t st);

}  // namespace l3d
