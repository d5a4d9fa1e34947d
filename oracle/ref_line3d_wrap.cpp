// oracle/ref_line3d_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's own Line3D++ sources -- src/view.cc, src/line3D.cc, src/clustering.cc and the headers they
// include -- compiled UNMODIFIED from where they lie under /root/reference into oracle/_ref/libref_line3d*.so
// (oracle/Makefile, target ref_line3d), behind the same C entry points as the CPU restatement
// (oracle/l3d_oracle.cpp: orc_create / orc_add_image / ... ), so that tests can run the two side by side on
// the same scenes.  Eigen, Boost and OpenCV are not installed in this image: oracle/standin/ supplies the small
// part of their interfaces these sources touch (see the headers there for what that does and does not pin).
// What runs here IS the reference's control flow, thresholds, containers, list handling and OpenMP-free serial
// order; what is the stand-in's: 3-vector arithmetic order, JacobiSVD, and (in the _det build) the libm calls.
//
// Private members are read directly (#define private public): the class layout is unaffected because the whole
// library is this single translation unit.
#define private public
#define protected public
#include "line3D.h"
#undef private
#undef protected

#include "/root/reference/src/clustering.cc"

// ---- hook: A_, local2global_ and the cluster roots exist only inside Line3D::clusterSegments
// (src/line3D.cc:2502-2575: performClustering sorts A_ in place, the maps are cleared right after) ----
namespace L3DPP {   // the call site is written L3DPP::performClustering(...)
namespace l3d_hook {
struct Snapshot {
    std::vector<L3DPP::CLEdge> A;
    std::vector<std::pair<unsigned, unsigned>> local2global;
    std::vector<int> cluster_ids;
};
static Snapshot g_snap;
static L3DPP::Line3D* g_current = nullptr;
static inline L3DPP::CLUniverse* hooked_clustering(std::list<L3DPP::CLEdge>& edges, int n, float c)
{
    g_snap.A.assign(edges.begin(), edges.end());
    g_snap.local2global.clear();
    if (g_current)
        for (int i = 0; i < n; ++i) {
            const L3DPP::Segment2D& s = g_current->local2global_[i];
            g_snap.local2global.push_back({s.camID(), s.segID()});
        }
    L3DPP::CLUniverse* u = (L3DPP::performClustering)(edges, n, c);
    // CLUniverse::find compresses only the queried node and never changes a root: reading the ids is harmless
    g_snap.cluster_ids.resize(n);
    for (int i = 0; i < n; ++i) g_snap.cluster_ids[i] = u->find(i);
    return u;
}
}  // namespace l3d_hook
}  // namespace L3DPP
namespace l3d_hook = L3DPP::l3d_hook;
#define performClustering(A, n, c) l3d_hook::hooked_clustering(A, n, c)

#define private public
#include "/root/reference/src/view.cc"
#include "/root/reference/src/line3D.cc"
#undef private
#undef performClustering

#include <cstdint>
#include <cstring>
#ifdef L3D_REF_OPENMP
#include <omp.h>
#endif

#include "standin/l3d_standin_offpath.h"

namespace {
struct Ref {
    L3DPP::Line3D* L;
    std::map<unsigned, unsigned> nseg;
    l3d_hook::Snapshot snap;
    bool by_wps;
    uint64_t pair_tests = 0;
    std::map<unsigned, std::set<unsigned>> matched_before;
    std::vector<std::pair<unsigned, unsigned>> pair_log;  // (src, tgt) in the order computeMatches visited them
};
struct OrcRec {
    uint32_t tgt_cam, tgt_seg;
    float overlap, score, d_p1, d_p2, d_q1, d_q2;
    uint32_t flags;
};
struct OrcEntry {
    uint32_t src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap, score, d_p1, d_p2, d_q1, d_q2, length;
    uint32_t pad;
    double P1[3], P2[3], dir[3];
};
Eigen::Matrix3d toM3(const double* m)
{
    Eigen::Matrix3d M;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M(i, j) = m[3 * i + j];
    return M;
}
// silence the reference's progress output (std::cout) while its methods run
struct Quiet {
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// Line3D::Line3D(output_folder, load_segments, max_img_width, max_line_segments, neighbors_by_worldpoints, use_GPU)
// with the arguments L3DPPing::Run passes (src/L3DPPing.cpp:67-69) except the neighbour mode and the width
void* orc_create(int max_img_width, int neighbors_by_worldpoints)
{
    Quiet q;
    Ref* r = new Ref;
    r->by_wps = neighbors_by_worldpoints != 0;
    r->L = new L3DPP::Line3D("/tmp/l3d_ref_out", false, max_img_width, 100000, r->by_wps, false);
    r->L->matched_.clear();  // src/L3DPPing.cpp:72
    return r;
}
void orc_destroy(void* h)
{
    Ref* r = (Ref*)h;
    // Line3D::~Line3D deletes its views; nothing else to release
    delete r->L;
    delete r;
}
#ifdef L3D_REF_OPENMP
void orc_set_threads(int n) { omp_set_num_threads(n > 0 ? n : omp_get_num_procs()); }
int orc_max_threads() { return omp_get_max_threads(); }
#else
void orc_set_threads(int) {}
int orc_max_threads() { return 1; }
#endif
void orc_set_snapshot(void*, int) {}

int orc_add_image(void* h, uint32_t camID, const double* K, const double* R, const double* t, unsigned w, unsigned hh,
                  float median_depth, const uint32_t* wn, int nwn, const float* segs, int nsegs)
{
    Quiet q;
    Ref* r = (Ref*)h;
    std::list<unsigned int> l(wn, wn + nwn);
    std::vector<cv::Vec4f> s(nsegs);
    for (int i = 0; i < nsegs; ++i) s[i] = cv::Vec4f(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]);
    cv::Mat image((int)hh, (int)w, CV_8U);  // only its size is read when segments are given (src/line3D.cc:124,209)
    const size_t before = r->L->views_.size();
    r->L->addImage(camID, image, toM3(K), toM3(R), Eigen::Vector3d(t[0], t[1], t[2]), median_depth, l, s);
    r->nseg[camID] = (unsigned)nsegs;
    return r->L->views_.size() == before + 1 ? 0 : -1;
}
int orc_update_image(void* h, uint32_t camID, const double* R, const double* t, float median_depth, const uint32_t* wn,
                     int nwn)
{
    Quiet q;
    Ref* r = (Ref*)h;
    std::list<unsigned int> l(wn, wn + nwn);
    r->L->UpdataImage(camID, toM3(R), Eigen::Vector3d(t[0], t[1], t[2]), median_depth, l);
    return 0;
}
int orc_delete_image(void* h, uint32_t camID)
{
    Quiet q;
    return ((Ref*)h)->L->deleteImage(camID) ? 1 : 0;
}
// the per-cycle resets of L3DPPing::Run (src/L3DPPing.cpp:98-103)
void orc_begin_cycle(void* h)
{
    L3DPP::Line3D* L = ((Ref*)h)->L;
    L->views2worldpoints_.clear();
    L->worldpoints2views_.clear();
    L->Delete_camID_.clear();
    L->Add_camID_.clear();
}
void orc_match_images(void* h, float sp, float sa, unsigned nn, float eo, int knn, float crd)
{
    Quiet q;
    Ref* r = (Ref*)h;
    r->L->matchImages(sp, sa, nn, eo, knn, crd);
    // The pairs this call matched, in the order Line3D::computeMatches visits them (src/line3D.cc:848-887: views
    // ascending, their visual neighbours ascending, skipping pairs in matched_): read back from the sets the
    // reference filled -- visual_neighbors_ (this call's neighbours) and matched_ (grown by both directions).
    for (auto& vn : r->L->visual_neighbors_) {
        const unsigned src = vn.first;
        for (unsigned tgt : vn.second) {
            if (r->matched_before[src].count(tgt)) continue;
            if (!r->L->matched_.count(src) || !r->L->matched_[src].count(tgt)) continue;  // not matched by the reference
            r->pair_log.push_back({src, tgt});
            r->pair_tests += (uint64_t)r->nseg[src] * r->nseg[tgt];
            r->matched_before[src].insert(tgt);
            r->matched_before[tgt].insert(src);
        }
    }
}
int orc_num_pairs(void* h) { return (int)((Ref*)h)->pair_log.size(); }
void orc_get_pairs(void* h, uint32_t* out)
{
    auto& v = ((Ref*)h)->pair_log;
    for (size_t i = 0; i < v.size(); ++i) { out[2 * i] = v[i].first; out[2 * i + 1] = v[i].second; }
}
// Line3D::reconstruct3Dlines(visibility_t = 3, diffusion off, collinearity off, CERES off): src/L3DPPing.cpp:227
void orc_reconstruct(void* h)
{
    Quiet q;
    Ref* r = (Ref*)h;
    l3d_hook::g_current = r->L;
    l3d_hook::g_snap = l3d_hook::Snapshot();
    r->L->reconstruct3Dlines(3, false, -1.0f, false);
    r->snap = l3d_hook::g_snap;
    l3d_hook::g_current = nullptr;
}
void orc_reconstruct_collin(void* h, float collinearity_t)
{
    Quiet q;
    Ref* r = (Ref*)h;
    l3d_hook::g_current = r->L;
    l3d_hook::g_snap = l3d_hook::Snapshot();
    r->L->reconstruct3Dlines(3, false, collinearity_t, false);
    r->snap = l3d_hook::g_snap;
    l3d_hook::g_current = nullptr;
}
uint64_t orc_pair_tests(void* h) { return ((Ref*)h)->pair_tests; }

// which: only 1 (the current lists = matches_ after filterMatches) exists in the reference
uint64_t orc_list_total(void* h, uint32_t cam, int which)
{
    Ref* r = (Ref*)h;
    if (which != 1 || !r->L->matches_.count(cam)) return 0;
    uint64_t n = 0;
    for (auto& l : r->L->matches_[cam]) n += l.size();
    return n;
}
int orc_get_lists(void* h, uint32_t cam, int which, uint32_t* row_off, OrcRec* out)
{
    Ref* r = (Ref*)h;
    if (which != 1 || !r->L->matches_.count(cam)) return -1;
    auto& v = r->L->matches_[cam];
    uint32_t n = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        row_off[i] = n;
        for (const L3DPP::Match& m : v[i]) {
            out[n].tgt_cam = m.tgt_camID_;
            out[n].tgt_seg = m.tgt_segID_;
            out[n].overlap = m.overlap_score_;
            out[n].score = m.score3D_;
            out[n].d_p1 = m.depth_p1_;
            out[n].d_p2 = m.depth_p2_;
            out[n].d_q1 = m.depth_q1_;
            out[n].d_q2 = m.depth_q2_;
            out[n].flags = m.match_orientation_ ? 1u : 0u;
            ++n;
        }
    }
    row_off[v.size()] = n;
    return 0;
}
int orc_num_entries(void* h) { return (int)((Ref*)h)->L->estimated_position3D_.size(); }
void orc_get_entries(void* h, OrcEntry* out)
{
    auto& est = ((Ref*)h)->L->estimated_position3D_;
    for (size_t i = 0; i < est.size(); ++i) {
        const L3DPP::Segment3D& s = est[i].first;
        const L3DPP::Match& m = est[i].second;
        OrcEntry& o = out[i];
        o.src_cam = m.src_camID_; o.src_seg = m.src_segID_; o.tgt_cam = m.tgt_camID_; o.tgt_seg = m.tgt_segID_;
        o.overlap = m.overlap_score_; o.score = m.score3D_;
        o.d_p1 = m.depth_p1_; o.d_p2 = m.depth_p2_; o.d_q1 = m.depth_q1_; o.d_q2 = m.depth_q2_;
        o.length = s.length();
        o.pad = 0;
        const Eigen::Vector3d P1 = s.P1(), P2 = s.P2(), d = s.dir();
        for (int k = 0; k < 3; ++k) { o.P1[k] = P1(k); o.P2[k] = P2(k); o.dir[k] = d(k); }
    }
}
int orc_num_edges(void* h) { return (int)((Ref*)h)->snap.A.size(); }
void orc_get_edges(void* h, int* ij, float* w)
{
    auto& A = ((Ref*)h)->snap.A;
    for (size_t i = 0; i < A.size(); ++i) { ij[2 * i] = A[i].i_; ij[2 * i + 1] = A[i].j_; w[i] = A[i].w_; }
}
int orc_num_local(void* h) { return (int)((Ref*)h)->snap.local2global.size(); }
void orc_get_local2global(void* h, uint32_t* cam_seg)
{
    auto& v = ((Ref*)h)->snap.local2global;
    for (size_t i = 0; i < v.size(); ++i) { cam_seg[2 * i] = v[i].first; cam_seg[2 * i + 1] = v[i].second; }
}
int orc_get_cluster_ids(void* h, int* out)
{
    auto& v = ((Ref*)h)->snap.cluster_ids;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}
int orc_get_view_info(void* h, uint32_t cam, double* C, float* kmm)
{
    L3DPP::Line3D* L = ((Ref*)h)->L;
    if (!L->views_.count(cam)) return -1;
    L3DPP::View* v = L->views_[cam];
    const Eigen::Vector3d c = v->C();
    C[0] = c(0); C[1] = c(1); C[2] = c(2);
    kmm[0] = v->k(); kmm[1] = v->median_depth(); kmm[2] = v->median_sigma();
    return 0;
}
int orc_get_neighbors(void* h, uint32_t cam, uint32_t* out, int cap)
{
    L3DPP::Line3D* L = ((Ref*)h)->L;
    if (!L->visual_neighbors_.count(cam)) return -1;
    int n = 0;
    for (unsigned v : L->visual_neighbors_[cam]) { if (n < cap) out[n] = v; ++n; }
    return n;
}
float orc_med_scene_depth_lines(void* h) { return ((Ref*)h)->L->med_scene_depth_lines_; }

// the final 3-D lines (Line3D::lines3D_ after reconstruct3Dlines): per line the collinear 3-D segments and the
// 2-D residuals of the underlying cluster.  Two-call protocol: sizes first, then the flat arrays.
//   counts[0] = lines, [1] = 3-D segments in total, [2] = residuals in total
void ref_lines3D_counts(void* h, uint32_t* counts)
{
    auto& v = ((Ref*)h)->L->lines3D_;
    counts[0] = (uint32_t)v.size();
    counts[1] = counts[2] = 0;
    for (auto& l : v) {
        counts[1] += (uint32_t)l.collinear3Dsegments_.size();
        counts[2] += (uint32_t)l.underlyingCluster_.residuals()->size();
    }
}
// seg_off[lines+1], segs[6 * n_segs] (P1, P2), res_off[lines+1], res[2 * n_res] (cam, seg), ref_view[lines]
void ref_lines3D_get(void* h, uint32_t* seg_off, double* segs, uint32_t* res_off, uint32_t* res, uint32_t* ref_view)
{
    auto& v = ((Ref*)h)->L->lines3D_;
    uint32_t ns = 0, nr = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        seg_off[i] = ns;
        res_off[i] = nr;
        ref_view[i] = v[i].underlyingCluster_.reference_view();
        for (const L3DPP::Segment3D& s : v[i].collinear3Dsegments_) {
            const Eigen::Vector3d P1 = s.P1(), P2 = s.P2();
            for (int k = 0; k < 3; ++k) { segs[6 * ns + k] = P1(k); segs[6 * ns + 3 + k] = P2(k); }
            ++ns;
        }
        for (const L3DPP::Segment2D& s : *v[i].underlyingCluster_.residuals()) {
            res[2 * nr] = s.camID();
            res[2 * nr + 1] = s.segID();
            ++nr;
        }
    }
    seg_off[v.size()] = ns;
    res_off[v.size()] = nr;
}
// Line3D::save3DLinesAsTXT (src/line3D.cc:3122-3178) into `folder`: the reference's own writer
void ref_save_txt(void* h, const char* folder)
{
    Quiet q;
    ((Ref*)h)->L->save3DLinesAsTXT(folder);
}

}  // extern "C"
