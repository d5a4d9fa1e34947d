// oracle/l3d_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Dependency-free CPU restatement of the Line3D++ matching -> triangulation -> scoring ->
// affinity -> clustering path of BTREE-C802/3DLine-SLAM (reference files cited per function as
// file:line relative to the reference root).  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library.
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN CODE: the reference ships no tests, golden vectors or runnable inputs
// for this path, and Eigen / Boost / OpenCV are absent here -- but src/line3D.cc, src/view.cc and src/clustering.cc
// compile unmodified against small stand-ins for the part of those libraries they touch (oracle/standin/,
// oracle/ref_line3d_wrap.cpp -> oracle/_ref/libref_line3d_{det,libm}.so, `make -C oracle ref`).  This restatement
// reproduces that build BIT FOR BIT -- filtered match lists, estimated_position3D_, A_ (order, ids, weights),
// local2global_, cluster roots, k, median depths -- on batch scenes, on the world-point (.nvm) path and cycle by
// cycle on a key-frame stream with deleted views (tests/test_ref_line3d.py); the CUDA path is compared with the
// same build directly (tests/test_ref_line3d_gpu.py).  What the stand-ins define and the reference does not pin:
// the order of 3-vector sums, and (det build) the libm calls; the libm build agrees within 1e-4 with identical
// match sets and cluster roots.  Further pins: hand-derived known-answer tests (tests/test_oracle_kat.py), the
// reference's clustering alone (oracle/_ref/libref_clustering.so, tests/test_ref_clustering.py) and a second
// independent reading in numpy (tests/test_oracle_second_reading.py).
//
// Canonical arithmetic: IEEE double/float exactly where the reference
// uses them, scalar left-to-right sums, row-major 3x3*v, true divisions, no FMA contraction
// (-ffp-contract=off), serial result order (= the reference with OMP_NUM_THREADS=1), and the
// deterministic transcendentals of oracle/detmath.h in place of libm.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "detmath.h"

namespace orc {

static const double EPS = 1e-12;          // L3D_EPS, commons.h:105
static const float PI_1_32 = 0.098174771f;  // commons.h:109
static const float PI_31_32 = 3.043417886f; // commons.h:110
static const float MIN_SIM_3D = 0.50f;      // commons.h:69
static const float MIN_BEST_3D = 0.75f;     // commons.h:70
static const float MIN_BEST_PERC = 0.10f;   // commons.h:71
static const float MIN_AFFINITY = 0.50f;    // commons.h:78

struct V3 {
    double x, y, z;
};
struct M3 {
    double m[3][3];
};

static inline V3 add(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 sub(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 scale(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(const V3& a, const V3& b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }
static inline V3 normalized(const V3& a)
{
    const double n = norm(a);
    return {a.x / n, a.y / n, a.z / n};
}
static inline V3 mul(const M3& A, const V3& v)
{
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
            A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
static inline M3 matmul(const M3& A, const M3& B)
{
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
static inline M3 transpose(const M3& A)
{
    M3 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T.m[i][j] = A.m[j][i];
    return T;
}
// adjugate / determinant inverse (the closed form Eigen uses for fixed 3x3)
static inline M3 inverse(const M3& A)
{
    auto cof = [&](int i, int j) {
        return A.m[(i + 1) % 3][(j + 1) % 3] * A.m[(i + 2) % 3][(j + 2) % 3] -
               A.m[(i + 1) % 3][(j + 2) % 3] * A.m[(i + 2) % 3][(j + 1) % 3];
    };
    const double det = cof(0, 0) * A.m[0][0] + cof(1, 0) * A.m[1][0] + cof(2, 0) * A.m[2][0];
    const double invdet = 1.0 / det;
    M3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i][j] = cof(j, i) * invdet;
    return R;
}

// commons.h:197-219
struct Match {
    uint32_t src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap, score3D;
    float d_p1, d_p2, d_q1, d_q2;
    bool orient, valid;
};
// commons.h:233-241
struct MatchKNN {
    bool operator()(const Match& a, const Match& b) const { return a.overlap < b.overlap; }
};

// segment3D.h:46-128
struct Seg3D {
    V3 P1{0, 0, 0}, P2{0, 0, 0}, dir{0, 0, 0};
    float length = 0.0f;
    bool valid = false;
    Seg3D() {}
    Seg3D(const V3& a, const V3& b)
    {
        length = (float)norm(sub(a, b));
        if (length > EPS) {
            P1 = a;
            P2 = b;
            dir = normalized(sub(b, a));
            valid = true;
        } else {
            length = 0.0f;
        }
    }
    // segment3D.h:80-84: P1 + (dir * (P-P1)^T) * dir, evaluated as outer product then mat*vec
    float distPointLine(const V3& P) const
    {
        const V3 d = sub(P, P1);
        const double dv[3] = {dir.x, dir.y, dir.z};
        const double df[3] = {d.x, d.y, d.z};
        double h[3];
        for (int i = 0; i < 3; ++i)
            h[i] = (dv[i] * df[0]) * dv[0] + (dv[i] * df[1]) * dv[1] + (dv[i] * df[2]) * dv[2];
        const V3 hp = {P1.x + h[0], P1.y + h[1], P1.z + h[2]};
        return (float)norm(sub(hp, P));
    }
};

struct Seg2f {
    float x1, y1, x2, y2;
};

// view.cc:6-44, view.h:122-135
struct View {
    uint32_t id = 0;
    std::vector<Seg2f> lines;
    M3 K, Kinv, R, Rt, RtKinv;
    V3 t, C, pp;
    unsigned width = 0, height = 0;
    float k = 0.0f, median_depth = 0.0f, median_sigma = 0.0f, initial_median_depth = 0.0f;
    V3 C_match{0, 0, 0};  // test hook: the (translated) centre matchImages worked with

    void init(uint32_t id_, const M3& K_, const M3& R_, const V3& t_, unsigned w, unsigned h,
              float med_depth)
    {
        id = id_;
        K = K_;
        R = R_;
        t = t_;
        width = w;
        height = h;
        initial_median_depth = (float)std::fmax(std::fabs((double)med_depth), EPS);
        pp = {K.m[0][2], K.m[1][2], 1.0};
        Kinv = inverse(K);
        Rt = transpose(R);
        RtKinv = matmul(Rt, Kinv);
        C = mul(Rt, scale(t, -1.0));
        k = 0.0f;
        median_depth = 0.0f;
        median_sigma = 0.0f;
    }
    // view.cc:62-87
    void update(const M3& R_, const V3& t_, float med_depth)
    {
        R = R_;
        Rt = transpose(R);
        RtKinv = matmul(Rt, Kinv);
        t = t_;
        C = mul(Rt, scale(t, -1.0));
        if (initial_median_depth == 0.0f)
            initial_median_depth = (float)std::fmax(std::fabs((double)med_depth), EPS);
    }
    // view.cc:539-543
    void translate(const V3& tv)
    {
        C = add(C, tv);
        const V3 rc = mul(R, C);
        t = {-rc.x, -rc.y, -rc.z};
    }
    // view.cc:346-350
    V3 ray(const V3& p) const { return normalized(mul(RtKinv, p)); }
    // view.cc:336-343
    float specificSpatialReg(float r) const
    {
        const V3 pps = {pp.x + (double)r, pp.y + 0.0, pp.z + 0.0};
        const V3 a = ray(pp), b = ray(pps);
        const double alpha = orc_acos(std::fmin(std::fmax(dot(a, b), -1.0), 1.0));
        return (float)orc_sin(alpha);
    }
    // view.cc:385-400
    Seg3D unproject(uint32_t seg, float d1, float d2) const
    {
        if (seg >= lines.size()) return Seg3D();
        const Seg2f& l = lines[seg];
        const V3 p1 = {(double)l.x1, (double)l.y1, 1.0}, p2 = {(double)l.x2, (double)l.y2, 1.0};
        return Seg3D(add(C, scale(ray(p1), (double)d1)), add(C, scale(ray(p2), (double)d2)));
    }
    // view.cc:474-477
    float regularizerFrom3D(const V3& P) const { return (float)(norm(sub(P, C)) * (double)k); }
    // view.cc:495-513
    double segmentQualityAngle(const Seg3D& s, uint32_t seg) const
    {
        if (seg >= lines.size()) return 0.0;
        const Seg2f& l = lines[seg];
        const V3 p = {0.5 * ((double)l.x1 + (double)l.x2), 0.5 * ((double)l.y1 + (double)l.y2), 1.0};
        const V3 r1 = ray(p);
        return orc_acos(std::fmin(std::fmax(dot(r1, s.dir), -1.0), 1.0));
    }
    // view.cc:486-492
    double opticalAxesAngle(const View& v) const
    {
        return orc_acos(std::fmin(std::fmax(dot(ray(pp), v.ray(v.pp)), -1.0), 1.0));
    }
    // view.cc:516-530
    float distanceVisualNeighborScore(const View& v) const
    {
        const V3 rc = mul(R, v.C);
        const V3 c = add(rc, t);
        const float d1 = (float)std::fabs(1.0 * c.x + 0.0 * c.y + 0.0 * c.z);
        const float d2 = (float)std::fabs(0.0 * c.x + 1.0 * c.y + 0.0 * c.z);
        return d1 + d2;
    }
    float baseLine(const View& v) const { return (float)norm(sub(C, v.C)); }
    // collinear segments (view.cc:180-200, 238-318)
    std::vector<std::list<unsigned>> collin;
    float collin_t = -1.0f;
    static bool pointOnSegment2D(const V3& p1, const V3& p2, const V3& x)  // view.cc:321-327
    {
        const double v1x = p1.x - x.x, v1y = p1.y - x.y, v2x = p2.x - x.x, v2y = p2.y - x.y;
        return (v1x * v2x + v1y * v2y) < EPS;
    }
    static float distPoint2Line2D(const V3& l, const V3& p)  // view.cc:296-299
    {
        return (float)std::fabs((l.x * p.x + l.y * p.y + l.z) / sqrtf((float)(l.x * l.x + l.y * l.y)));
    }
    static bool collinearPair(const Seg2f& a, const Seg2f& b, float dist_t)  // body of view.cc:256-291
    {
        const V3 p0 = {(double)a.x1, (double)a.y1, 1.0}, p1 = {(double)a.x2, (double)a.y2, 1.0};
        const V3 q0 = {(double)b.x1, (double)b.y1, 1.0}, q1 = {(double)b.x2, (double)b.y2, 1.0};
        const V3 line1 = cross(p0, p1), line2 = cross(q0, q1);
        if (pointOnSegment2D(p0, p1, q0) || pointOnSegment2D(p0, p1, q1) || pointOnSegment2D(q0, q1, p0) ||
            pointOnSegment2D(q0, q1, p1))
            return false;
        const float d1 = (float)std::fmax((double)distPoint2Line2D(line1, q0), (double)distPoint2Line2D(line1, q1));
        const float d2 = (float)std::fmax((double)distPoint2Line2D(line2, p0), (double)distPoint2Line2D(line2, p1));
        return (float)std::fmax((double)d1, (double)d2) < dist_t;
    }
    void findCollinearSegments(float dist_t)
    {
        if (std::fabs((double)(dist_t - collin_t)) < EPS) return;  // already computed
        if (!(dist_t > EPS)) return;
        collin_t = dist_t;
        collin.assign(lines.size(), std::list<unsigned>());
        for (size_t r = 0; r < lines.size(); ++r)
            for (size_t c = 0; c < lines.size(); ++c)
                if (r != c && collinearPair(lines[r], lines[c], collin_t)) collin[r].push_back((unsigned)c);
    }
    std::list<unsigned> collinearSegments(unsigned seg) const  // view.cc:312-318
    {
        if (collin.size() == lines.size() && seg < lines.size()) return collin[seg];
        return std::list<unsigned>();
    }
    // view.h:122-135
    void updateMedianDepth(float d, float sigmaP, float med_scene_depth)
    {
        median_depth = d;
        if (sigmaP > 0.0f) k = sigmaP / med_scene_depth;
        median_sigma = k * median_depth;
    }
};

struct Edge {
    int i, j;
    float w;
};
struct Entry {
    Seg3D seg;
    Match m;
};
typedef std::pair<uint32_t, uint32_t> Seg2D;  // (camID, segID), ordered like commons.h:126-128

struct Timers {
    double match = 0, orient = 0, score = 0, inverse = 0, filter = 0, update = 0, affinity = 0,
           cluster = 0, match_images = 0, reconstruct = 0;
};

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

class Line3D {
  public:
    int max_image_width;
    bool neighbors_by_worldpoints;
    bool snapshot = true;
    // params (line3D.cc:17-34)
    unsigned num_neighbors = 10;
    float epipolar_overlap = 0.25f;
    int kNN = 10;
    float sigma_p = 2.5f, sigma_a = 10.0f, two_sigA_sqr = 200.0f;
    float const_reg_depth = -1.0f;
    bool fixed3Dreg = false;
    float med_scene_depth = (float)EPS, med_scene_depth_lines = 0.0f;
    V3 translation{0, 0, 0};
    uint64_t num_lines_total = 0;

    std::map<uint32_t, View*> views;
    std::set<uint32_t> views_reserved;
    std::vector<uint32_t> view_order;
    std::map<uint32_t, std::vector<std::list<Match>>> matches;
    std::map<uint32_t, unsigned> num_matches;
    std::map<uint32_t, bool> processed;
    std::map<uint32_t, std::set<uint32_t>> visual_neighbors;
    std::map<uint32_t, std::list<uint32_t>> fixed_visual_neighbors;
    std::map<uint32_t, std::set<uint32_t>> matched;
    std::map<uint32_t, std::map<uint32_t, M3>> fundamentals;
    std::map<uint32_t, float> views_avg_depths;
    std::map<uint32_t, std::list<uint32_t>> worldpoints2views, views2worldpoints;
    std::map<uint32_t, unsigned> num_worldpoints;
    std::set<uint32_t> delete_cams, add_cams;

    std::vector<Entry> est3D;
    std::map<Seg2D, size_t> entry_map;
    std::list<Edge> A;
    std::vector<Edge> A_snapshot;
    std::map<Seg2D, int> global2local;
    std::map<int, Seg2D> local2global;
    std::vector<Seg2D> local2global_snapshot;
    std::vector<int> cluster_ids;
    std::map<Seg2D, std::set<Seg2D>> used;
    int localID = 0;

    // snapshots for parity checks (lists as they are right after scoring, before filtering)
    std::map<uint32_t, std::vector<std::list<Match>>> snap_scored;
    std::vector<std::pair<uint32_t, uint32_t>> pair_log;  // (src,tgt) in matching order
    uint64_t pair_tests = 0;
    Timers tm;

    Line3D(int max_w, bool by_wps) : max_image_width(max_w), neighbors_by_worldpoints(by_wps) {}
    ~Line3D()
    {
        for (auto& kv : views) delete kv.second;
    }

    // line3D.cc:117-227 (segments always given; image reduced to its size)
    int addImage(uint32_t camID, const M3& K, const M3& R, const V3& t, unsigned w, unsigned h,
                 float median_depth, const std::list<uint32_t>& wps_or_nbrs,
                 const std::vector<Seg2f>& segs)
    {
        if (std::max(w, h) < 400) return -1;  // L3D_DEF_MIN_IMG_WIDTH, line3D.cc:124
        if (views_reserved.count(camID)) return -2;
        views_reserved.insert(camID);
        add_cams.insert(camID);
        if (wps_or_nbrs.empty()) return -3;
        if (segs.empty()) return -4;
        View* v = new View();
        v->lines = segs;
        v->init(camID, K, R, t, w, h, median_depth);
        views[camID] = v;
        view_order.push_back(camID);
        matches[camID] = std::vector<std::list<Match>>(segs.size());
        num_matches[camID] = 0;
        processed[camID] = false;
        visual_neighbors[camID] = std::set<uint32_t>();
        num_lines_total += segs.size();
        views_avg_depths[camID] = (float)std::fmax((double)median_depth, EPS);
        return 0;
    }

    // line3D.cc:396-430
    bool deleteImage(uint32_t camID)
    {
        if (!views_reserved.count(camID)) return false;
        num_lines_total -= views[camID]->lines.size();
        view_order.erase(std::find(view_order.begin(), view_order.end(), camID));
        views_avg_depths.erase(camID);
        num_matches.erase(camID);
        processed.erase(camID);
        visual_neighbors.erase(camID);
        num_worldpoints.erase(camID);
        views2worldpoints.erase(camID);
        views_reserved.erase(camID);
        delete_cams.insert(camID);
        return true;
    }

    // line3D.cc:433-487
    int updateImage(uint32_t camID, const M3& R, const V3& t, float median_depth,
                    const std::list<uint32_t>& wps_or_nbrs)
    {
        auto it = views.find(camID);
        if (it == views.end()) return 0;
        views_avg_depths[camID] = (float)std::fmax((double)median_depth, EPS);
        it->second->update(R, t, median_depth);
        if (!views_reserved.count(camID)) return -1;
        if (neighbors_by_worldpoints) {
            if (wps_or_nbrs.empty()) return -3;
            // processWPlist, line3D.cc:230-241
            for (uint32_t wp : wps_or_nbrs) worldpoints2views[wp].push_back(camID);
            num_worldpoints[camID] = (unsigned)wps_or_nbrs.size();
            views2worldpoints[camID] = wps_or_nbrs;
        } else {
            fixed_visual_neighbors[camID] = wps_or_nbrs;  // line3D.cc:244-247
        }
        return 0;
    }

    // line3D.cc:643-720
    void performTranslation(const V3& tv)
    {
        for (uint32_t id : view_order) views[id]->translate(tv);
    }
    void translate()
    {
        if (views.empty()) return;
        translation = {0, 0, 0};
        double* tr[3] = {&translation.x, &translation.y, &translation.z};
        for (int i = 0; i < 3; ++i) {
            std::vector<double> c;
            for (uint32_t id : view_order) {
                const V3& C = views[id]->C;
                const double val = (i == 0) ? C.x : (i == 1 ? C.y : C.z);
                if (std::fabs(val) > EPS) c.push_back(val);
            }
            if (!c.empty()) {
                std::sort(c.begin(), c.end());
                *tr[i] = c[c.size() / 2];
            }
        }
        performTranslation({-translation.x, -translation.y, -translation.z});
    }
    void untranslate() { performTranslation(translation); }

    // line3D.cc:723-843
    void findVisualNeighborsFromWPs(uint32_t camID)
    {
        if (!visual_neighbors.count(camID)) return;
        visual_neighbors[camID].clear();
        std::map<uint32_t, unsigned> common;
        for (uint32_t wp : views2worldpoints[camID])
            for (uint32_t vID : worldpoints2views[wp])
                if (vID != camID) ++common[vID];
        if (common.empty()) return;
        struct VN {
            uint32_t cam;
            float score, axisAngle, distScore;
        };
        std::list<VN> nb;
        View* v = views[camID];
        for (auto& kv : common) {
            VN vn;
            vn.cam = kv.first;
            vn.score = 2.0f * float(kv.second) /
                       float(num_worldpoints[camID] + num_worldpoints[kv.first]);
            vn.axisAngle = (float)v->opticalAxesAngle(*views[kv.first]);
            vn.distScore = v->distanceVisualNeighborScore(*views[kv.first]);
            if (vn.axisAngle < 1.571f && kv.second > 4) nb.push_back(vn);
        }
        nb.sort([](const VN& a, const VN& b) { return a.score > b.score; });
        if (nb.size() > num_neighbors) {
            std::list<VN> tmp = nb;
            const float score_t = 0.80f * nb.front().score;
            unsigned bigger = 0;
            for (auto it = nb.begin(); it != nb.end() && it->score > score_t; ++it) ++bigger;
            nb.resize(bigger);
            nb.sort([](const VN& a, const VN& b) { return a.distScore > b.distScore; });
            if (nb.size() > num_neighbors / 2) nb.resize(num_neighbors / 2);
            nb.splice(nb.end(), tmp);
        }
        std::set<uint32_t> usedn;
        const float min_baseline = v->specificSpatialReg(0.5f) * v->median_depth;
        for (auto it = nb.begin(); it != nb.end() && usedn.size() < num_neighbors; ++it) {
            View* v2 = views[it->cam];
            if (!usedn.count(it->cam) && v->baseLine(*v2) > min_baseline) {
                bool ok = true;
                for (uint32_t u : usedn)
                    if (!(v->baseLine(*views[u]) > min_baseline)) {
                        ok = false;
                        break;
                    }
                if (ok) usedn.insert(it->cam);
            }
        }
        visual_neighbors[camID] = usedn;
    }

    // line3D.cc:1058-1094
    M3 fundamental(View* s, View* t)
    {
        auto& fs = fundamentals[s->id];
        if (fs.count(t->id)) return fs[t->id];
        auto& ft = fundamentals[t->id];
        if (ft.count(s->id)) return transpose(ft[s->id]);
        const M3 R = matmul(t->R, transpose(s->R));
        const V3 Rt1 = mul(R, s->t);
        const V3 tt = sub(t->t, Rt1);
        M3 T;
        T.m[0][0] = 0.0;   T.m[0][1] = -tt.z; T.m[0][2] = tt.y;
        T.m[1][0] = tt.z;  T.m[1][1] = 0.0;   T.m[1][2] = -tt.x;
        T.m[2][0] = -tt.y; T.m[2][1] = tt.x;  T.m[2][2] = 0.0;
        const M3 E = matmul(T, R);
        const M3 F = matmul(matmul(inverse(transpose(t->K)), E), inverse(s->K));
        fundamentals[s->id][t->id] = F;
        return F;
    }

    // line3D.cc:1274-1280
    static bool pointOnSegment(const V3& x, const V3& p1, const V3& p2)
    {
        const double v1x = p1.x - x.x, v1y = p1.y - x.y, v2x = p2.x - x.x, v2y = p2.y - x.y;
        return (v1x * v2x + v1y * v2y) < EPS;
    }
    // line3D.cc:1283-1362
    static float mutualOverlap(const V3* pt)
    {
        float overlap = 0.0f;
        if (pointOnSegment(pt[0], pt[2], pt[3]) || pointOnSegment(pt[1], pt[2], pt[3]) ||
            pointOnSegment(pt[2], pt[0], pt[1]) || pointOnSegment(pt[3], pt[0], pt[1])) {
            float max_dist = 0.0f;
            size_t o1 = 0, i1 = 1, i2 = 2, o2 = 3;
            for (size_t i = 0; i < 3; ++i)
                for (size_t j = i + 1; j < 4; ++j) {
                    const float d = (float)norm(sub(pt[i], pt[j]));
                    if (d > max_dist) {
                        max_dist = d;
                        o1 = i;
                        o2 = j;
                    }
                }
            if (max_dist < 1.0f) return 0.0f;
            if (o1 == 0) {
                if (o2 == 1) { i1 = 2; i2 = 3; }
                else if (o2 == 2) { i1 = 1; i2 = 3; }
                else { i1 = 1; i2 = 2; }
            } else if (o1 == 1) {
                i1 = 0;
                i2 = (o2 == 2) ? 3 : 2;
            } else {
                i1 = 0;
                i2 = 1;
            }
            overlap = (float)(norm(sub(pt[i1], pt[i2])) / (double)max_dist);
        }
        return overlap;
    }
    // line3D.cc:1365-1390
    static void triangulationDepths(const View* vs, const V3& p1, const V3& p2, const View* vt,
                                    const V3& q1, const V3& q2, double& d1, double& d2)
    {
        const V3 C1 = vs->C, r1 = vs->ray(p1), r2 = vs->ray(p2);
        const V3 C2 = vt->C;
        const V3 n = normalized(cross(vt->ray(q1), vt->ray(q2)));
        if (std::fabs(dot(r1, n)) < EPS || std::fabs(dot(r2, n)) < EPS) {
            d1 = -1;
            d2 = -1;
            return;
        }
        d1 = (dot(C2, n) - dot(n, C1)) / dot(n, r1);
        d2 = (dot(C2, n) - dot(n, C1)) / dot(n, r2);
    }

    // line3D.cc:1097-1212
    void matchingCPU(uint32_t src, uint32_t tgt, const M3& F)
    {
        View* vs = views[src];
        View* vt = views[tgt];
        const std::vector<Seg2f>& ls = vs->lines;
        const std::vector<Seg2f>& lt = vt->lines;
        std::vector<std::list<Match>>& out = matches[src];
        unsigned total = 0;
        const double W = (double)max_image_width;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
        for (int r = 0; r < (int)ls.size(); ++r) {
            int new_matches = 0;
            const V3 p1 = {(double)ls[r].x1, (double)ls[r].y1, 1.0};
            const V3 p2 = {(double)ls[r].x2, (double)ls[r].y2, 1.0};
            const V3 e1 = mul(F, p1), e2 = mul(F, p2);
            std::priority_queue<Match, std::vector<Match>, MatchKNN> pq;
            for (size_t c = 0; c < lt.size(); ++c) {
                const V3 q1 = {(double)lt[c].x1, (double)lt[c].y1, 1.0};
                const V3 q2 = {(double)lt[c].x2, (double)lt[c].y2, 1.0};
                const V3 l2 = cross(q1, q2);
                V3 a = cross(l2, e1), b = cross(l2, e2);
                if (std::fabs(a.z) > EPS && std::fabs(b.z) > EPS) {
                    a = {a.x / a.z, a.y / a.z, a.z / a.z};
                    b = {b.x / b.z, b.y / b.z, b.z / b.z};
                    if (a.x < 0 || a.x > W || a.y < 0 || a.y > W || b.x < 0 || b.x > W ||
                        b.y < 0 || b.y > W)
                        continue;
                    const V3 pts[4] = {a, b, q1, q2};
                    const float score = mutualOverlap(pts);
                    if (score > epipolar_overlap) {
                        double ds1, ds2, dt1, dt2;
                        triangulationDepths(vs, p1, p2, vt, q1, q2, ds1, ds2);
                        triangulationDepths(vt, q1, q2, vs, p1, p2, dt1, dt2);
                        if (ds1 > EPS && ds2 > EPS && dt1 > EPS && dt2 > EPS) {
                            Match M;
                            M.src_cam = src;
                            M.src_seg = (uint32_t)r;
                            M.tgt_cam = tgt;
                            M.tgt_seg = (uint32_t)c;
                            M.overlap = score;
                            M.score3D = 0.0f;
                            M.d_p1 = (float)ds1;
                            M.d_p2 = (float)ds2;
                            M.d_q1 = (float)dt1;
                            M.d_q2 = (float)dt2;
                            M.orient = false;
                            M.valid = false;
                            if (kNN > 0)
                                pq.push(M);
                            else {
                                out[r].push_back(M);
                                ++new_matches;
                            }
                        }
                    }
                }
            }
            if (kNN > 0)
                while (new_matches < kNN && !pq.empty()) {
                    out[r].push_back(pq.top());
                    pq.pop();
                    ++new_matches;
                }
            total += new_matches;
        }
        num_matches[src] += total;
        pair_tests += (uint64_t)ls.size() * lt.size();
    }

    // line3D.cc:1826-1838
    Seg3D unprojectMatch(const Match& m, bool src = true)
    {
        if (src) return views[m.src_cam]->unproject(m.src_seg, m.d_p1, m.d_p2);
        return views[m.tgt_cam]->unproject(m.tgt_seg, m.d_q1, m.d_q2);
    }

    // line3D.cc:962-1014
    void checkMatchOrientation(uint32_t src)
    {
        if (!matches.count(src)) return;
        std::vector<std::list<Match>>& ml = matches[src];
        unsigned total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
        for (int i = 0; i < (int)ml.size(); ++i) {
            std::list<Match> remaining;
            for (auto it = ml[i].begin(); it != ml[i].end(); ++it) {
                if (!it->orient) {
                    Match m = *it;
                    const Seg3D s = unprojectMatch(m);
                    const double ang = views[m.src_cam]->segmentQualityAngle(s, m.src_seg);
                    if (ang > PI_1_32 && ang < PI_31_32) {
                        it->orient = true;  // the stored copy keeps orient=false, as in the reference
                        remaining.push_back(m);
                    }
                } else {
                    remaining.push_back(*it);
                }
            }
            ml[i] = remaining;
            total += (unsigned)ml[i].size();
        }
        num_matches[src] = total;
    }

    // line3D.cc:1016-1055
    void updateMatch(uint32_t src)
    {
        if (!matches.count(src)) return;
        std::vector<std::list<Match>>& ml = matches[src];
        unsigned total = 0;
        for (size_t i = 0; i < ml.size(); ++i) {
            if (!delete_cams.empty()) {
                std::list<Match> remaining;
                for (const Match& m : ml[i])
                    if (!delete_cams.count(m.tgt_cam)) remaining.push_back(m);
                ml[i] = remaining;
            }
            total += (unsigned)ml[i].size();
        }
        num_matches[src] = total;
    }

    // line3D.cc:1841-1853
    static float angleBetweenSeg3D(const Seg3D& s1, const Seg3D& s2, bool undirected)
    {
        const float dot_p = (float)dot(s1.dir, s2.dir);
        float angle =
            (float)((double)orc_acosf(std::fmax(std::fmin(dot_p, 1.0f), -1.0f)) / M_PI * 180.0f);
        if (undirected && angle > 90.0f) angle = 180.0f - angle;
        return angle;
    }

    // line3D.cc:1685-1716
    float similarityForScoring(const Match& m1, const Match& m2, const Seg3D& seg1, float reg1,
                               float reg2)
    {
        const Seg3D seg2 = unprojectMatch(m2, true);
        if (seg1.length < EPS || seg2.length < EPS) return 0.0f;
        float sim_p = 0.0f;
        if (m1.src_cam == m2.src_cam && m1.src_seg == m2.src_seg) {
            const float d1 = m1.d_p1 - m2.d_p1;
            const float d2 = m1.d_p2 - m2.d_p2;
            sim_p = std::fmin(orc_expf(-d1 * d1 / reg1), orc_expf(-d2 * d2 / reg2));
        } else
            return 0.0f;
        const float angle = angleBetweenSeg3D(seg1, seg2, true);
        const float sim_a = orc_expf(-angle * angle / two_sigA_sqr);
        const float sim = std::fmin(sim_a, sim_p);
        return (sim > MIN_SIM_3D) ? sim : 0.0f;
    }

    // line3D.cc:1405-1562
    void scoringCPU(uint32_t src, float& valid_f)
    {
        valid_f = 0.0f;
        View* v = views[src];
        const float k = v->k;
        unsigned num_valid = 0;
        std::vector<std::list<Match>>& ml = matches[src];
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : num_valid)
        for (int i = 0; i < (int)ml.size(); ++i) {
            bool valid_exists = false;
            for (auto it = ml[i].begin(); it != ml[i].end(); ++it) {
                const Match M = *it;
                if (delete_cams.count(M.tgt_cam)) continue;
                const Seg3D M3D = v->unproject(M.src_seg, M.d_p1, M.d_p2);
                float reg1, reg2;
                const float sig1 = M.d_p1 * k;
                const float sig2 = M.d_p2 * k;
                reg1 = 2.0f * sig1 * sig1;
                reg2 = 2.0f * sig2 * sig2;
                const float sig1_t = views[M.tgt_cam]->regularizerFrom3D(M3D.P1);
                const float sig2_t = views[M.tgt_cam]->regularizerFrom3D(M3D.P2);
                reg1 = 0.5f * (reg1 + 2.0f * sig1_t * sig1_t);
                reg2 = 0.5f * (reg2 + 2.0f * sig2_t * sig2_t);
                if (it->score3D != 0) {
                    // already scored in an earlier cycle: add / subtract per-camera maxima of the
                    // cameras added / deleted since (line3D.cc:1439-1512)
                    std::map<uint32_t, float> adds, dels;
                    for (auto it2 = ml[i].begin(); it2 != ml[i].end(); ++it2) {
                        const Match M2 = *it2;
                        if (M.tgt_cam == M2.tgt_cam) continue;
                        if (!add_cams.empty() && add_cams.count(M2.tgt_cam)) {
                            const float sim = similarityForScoring(M, M2, M3D, reg1, reg2);
                            auto f = adds.find(M2.tgt_cam);
                            if (f != adds.end()) {
                                if (sim > f->second) f->second = sim;
                            } else
                                adds[M2.tgt_cam] = sim;
                        }
                        if (!delete_cams.empty() && delete_cams.count(M2.tgt_cam)) {
                            const float sim = similarityForScoring(M, M2, M3D, reg1, reg2);
                            auto f = dels.find(M2.tgt_cam);
                            if (f != dels.end()) {
                                if (sim > f->second) f->second = sim;
                            } else
                                dels[M2.tgt_cam] = sim;
                        }
                    }
                    for (auto& a : adds) it->score3D += a.second;
                    for (auto& d : dels) it->score3D -= d.second;
                    if (it->score3D > MIN_BEST_3D) valid_exists = true;
                } else {
                    std::map<uint32_t, float> per_cam;
                    for (auto it2 = ml[i].begin(); it2 != ml[i].end(); ++it2) {
                        const Match M2 = *it2;
                        if (M.tgt_cam == M2.tgt_cam) continue;
                        if (delete_cams.count(M2.tgt_cam)) continue;
                        const float sim = similarityForScoring(M, M2, M3D, reg1, reg2);
                        auto f = per_cam.find(M2.tgt_cam);
                        if (f != per_cam.end()) {
                            if (sim > f->second) {
                                it->score3D -= f->second;
                                it->score3D += sim;
                                f->second = sim;
                            }
                        } else {
                            it->score3D += sim;
                            per_cam[M2.tgt_cam] = sim;
                        }
                    }
                    if (it->score3D > MIN_BEST_3D) valid_exists = true;
                }
            }
            if (valid_exists) ++num_valid;
        }
        valid_f = float(num_valid) / float(v->lines.size());
    }

    // line3D.cc:1986-2015
    void storeInverseMatches(uint32_t src)
    {
        std::vector<std::list<Match>>& ml = matches[src];
        for (size_t i = 0; i < ml.size(); ++i)
            for (const Match& m : ml[i]) {
                if (m.score3D > 0.0f && !processed[m.tgt_cam]) {
                    Match inv = m;
                    inv.src_cam = m.tgt_cam;
                    inv.src_seg = m.tgt_seg;
                    inv.tgt_cam = m.src_cam;
                    inv.tgt_seg = m.src_seg;
                    inv.d_p1 = m.d_q1;
                    inv.d_p2 = m.d_q2;
                    inv.d_q1 = m.d_p1;
                    inv.d_q2 = m.d_p2;
                    inv.score3D = 0.0f;
                    inv.orient = true;
                    inv.valid = false;
                    matches[m.tgt_cam][m.tgt_seg].push_back(inv);
                    ++num_matches[m.tgt_cam];
                }
            }
    }

    // line3D.cc:1911-1983 (serial order)
    void filterMatches(uint32_t src)
    {
        std::vector<float> depths;
        std::vector<std::list<Match>>& ml = matches[src];
        float max_score = 0.0f;
        for (size_t i = 0; i < ml.size(); ++i)
            for (const Match& m : ml[i]) max_score = std::fmax(max_score, m.score3D);
        const float score_lim = MIN_BEST_PERC * max_score;
        for (size_t i = 0; i < ml.size(); ++i) {
            Match best;
            best.score3D = 0.0f;
            std::list<Match> all = ml[i];
            ml[i].clear();
            for (const Match& m : all) {
                if (m.score3D > 0.0f && m.score3D > score_lim) {
                    ml[i].push_back(m);
                    if (m.score3D > best.score3D) best = m;
                }
            }
            if (best.score3D > MIN_BEST_3D) {
                const Seg3D s = unprojectMatch(best, true);
                entry_map[Seg2D(src, (uint32_t)i)] = est3D.size();
                est3D.push_back({s, best});
                depths.push_back(best.d_p1);
                depths.push_back(best.d_p2);
            }
        }
        float med = (float)EPS;
        if (!depths.empty()) {
            std::sort(depths.begin(), depths.end());
            med = depths[depths.size() / 2];
        }
        if (!fixed3Dreg)
            views[src]->updateMedianDepth(med, -1.0f, med_scene_depth);
        else
            views[src]->updateMedianDepth(med, sigma_p, med_scene_depth);
    }

    // line3D.cc:846-930
    void computeMatches()
    {
        for (auto it = visual_neighbors.begin(); it != visual_neighbors.end(); ++it) {
            const uint32_t src = it->first;
            double t0 = now();
            for (uint32_t tgt : it->second) {
                if (!matched[src].count(tgt)) {
                    const M3 F = fundamental(views[src], views[tgt]);
                    matchingCPU(src, tgt, F);
                    matched[src].insert(tgt);
                    matched[tgt].insert(src);
                    pair_log.push_back({src, tgt});
                }
            }
            double t1 = now();
            tm.match += t1 - t0;
            checkMatchOrientation(src);
            double t2 = now();
            tm.orient += t2 - t1;
            float valid_f;
            scoringCPU(src, valid_f);
            double t3 = now();
            tm.score += t3 - t2;
            updateMatch(src);
            if (snapshot) snap_scored[src] = matches[src];
            storeInverseMatches(src);
            double t4 = now();
            tm.inverse += t4 - t3;
            filterMatches(src);
            tm.filter += now() - t4;
            processed[src] = true;
        }
    }

    // line3D.cc:1857-1908
    void updateMatchesAndEst3D()
    {
        entry_map.clear();
        std::vector<Entry> upd;
        for (size_t i = 0; i < est3D.size(); ++i) {
            Match m = est3D[i].m;
            if (views.count(m.src_cam) && views.count(m.tgt_cam) && m.score3D > MIN_BEST_3D) {
                View* vs = views[m.src_cam];
                View* vt = views[m.tgt_cam];
                const Seg2f& ls = vs->lines[m.src_seg];
                const Seg2f& lt = vt->lines[m.tgt_seg];
                const V3 p1 = {(double)ls.x1, (double)ls.y1, 1.0}, p2 = {(double)ls.x2, (double)ls.y2, 1.0};
                const V3 q1 = {(double)lt.x1, (double)lt.y1, 1.0}, q2 = {(double)lt.x2, (double)lt.y2, 1.0};
                double ds1, ds2, dt1, dt2;
                triangulationDepths(vs, p1, p2, vt, q1, q2, ds1, ds2);
                triangulationDepths(vt, q1, q2, vs, p1, p2, dt1, dt2);
                if (ds1 > EPS && ds2 > EPS && dt1 > EPS && dt2 > EPS) {
                    m.d_p1 = (float)ds1;
                    m.d_p2 = (float)ds2;
                    m.d_q1 = (float)dt1;
                    m.d_q2 = (float)dt2;
                    const Seg3D s = unprojectMatch(m, true);
                    entry_map[Seg2D(m.src_cam, m.src_seg)] = upd.size();
                    upd.push_back({s, m});
                }
            }
        }
        est3D = upd;
    }

    // line3D.cc:496-640
    void matchImages(float sigma_position, float sigma_angle, unsigned nnbrs, float epi_overlap,
                     int knn, float const_depth)
    {
        const double T0 = now();
        if (views.empty()) return;
        num_neighbors = (unsigned)std::max(int(nnbrs), 2);
        sigma_p = sigma_position;
        sigma_a = (float)std::fmin(std::fabs((double)sigma_angle), 90.0);
        two_sigA_sqr = 2.0f * sigma_a * sigma_a;
        epipolar_overlap = (float)std::fmin(std::fabs((double)epi_overlap), (double)0.99f);
        kNN = knn;
        const_reg_depth = const_depth;
        if (sigma_p < 0.0f) {
            fixed3Dreg = true;
            sigma_p = std::fabs(sigma_p);
        } else {
            fixed3Dreg = false;
            sigma_p = (float)std::fmax((double)0.1f, (double)sigma_p);
        }
        est3D.clear();
        entry_map.clear();
        med_scene_depth = const_reg_depth;
        // (metric-sigma median-depth quirk of line3D.cc:557-565 not restated: SURVEY.md App. B)
        translate();
        for (uint32_t camID : view_order) views[camID]->C_match = views[camID]->C;
        for (uint32_t camID : view_order) {
            if (!fixed3Dreg)
                views[camID]->k = views[camID]->specificSpatialReg(sigma_p);
            else
                views[camID]->k = sigma_p / med_scene_depth;
            if (!matches.count(camID)) {
                matches[camID] = std::vector<std::list<Match>>(views[camID]->lines.size());
                num_matches[camID] = 0;
                processed[camID] = false;
            }
        }
        for (uint32_t camID : view_order) {
            if (fixed_visual_neighbors.count(camID)) {
                if (visual_neighbors[camID].empty())
                    for (uint32_t n : fixed_visual_neighbors[camID])
                        if (views.count(n)) visual_neighbors[camID].insert(n);
            } else {
                findVisualNeighborsFromWPs(camID);
            }
        }
        computeMatches();
        const double T1 = now();
        updateMatchesAndEst3D();
        tm.update += now() - T1;
        untranslate();
        tm.match_images += now() - T0;
    }

    // line3D.cc:1737-1823
    float similarity(const Seg3D& s1, const Match& m1, const Seg2D& seg2, bool truncate)
    {
        auto f = entry_map.find(seg2);
        if (f == entry_map.end()) return 0.0f;
        const Seg3D& s2 = est3D[f->second].seg;
        const Match& m2 = est3D[f->second].m;
        if (s1.length < EPS || s2.length < EPS) return 0.0f;
        View* v1 = views[m1.src_cam];
        View* v2 = views[m2.src_cam];
        const float angle = angleBetweenSeg3D(s1, s2, true);
        const float sim_a = orc_expf(-angle * angle / two_sigA_sqr);
        float cutoff1 = v1->median_depth;
        float cutoff2 = v2->median_depth;
        if (med_scene_depth_lines > EPS) {
            cutoff1 = std::fmin(cutoff1, med_scene_depth_lines);
            cutoff2 = std::fmin(cutoff2, med_scene_depth_lines);
        }
        const float d11 = s2.distPointLine(s1.P1);
        const float d12 = s2.distPointLine(s1.P2);
        const float d21 = s1.distPointLine(s2.P1);
        const float d22 = s1.distPointLine(s2.P2);
        const float sig11 = (m1.d_p1 > cutoff1) ? cutoff1 * v1->k : m1.d_p1 * v1->k;
        const float sig12 = (m1.d_p2 > cutoff1) ? cutoff1 * v1->k : m1.d_p2 * v1->k;
        const float reg11 = 2.0f * sig11 * sig11;
        const float reg12 = 2.0f * sig12 * sig12;
        const float sig21 = (m2.d_p1 > cutoff2) ? cutoff2 * v2->k : m2.d_p1 * v2->k;
        const float sig22 = (m2.d_p2 > cutoff2) ? cutoff2 * v2->k : m2.d_p2 * v2->k;
        const float reg21 = 2.0f * sig21 * sig21;
        const float reg22 = 2.0f * sig22 * sig22;
        const float sim_p1 = std::fmin(orc_expf(-d11 * d11 / reg11), orc_expf(-d12 * d12 / reg12));
        const float sim_p2 = std::fmin(orc_expf(-d21 * d21 / reg21), orc_expf(-d22 * d22 / reg22));
        const float sim_p = std::fmin(sim_p1, sim_p2);
        const float sim = std::fmin(sim_a, sim_p);
        if (truncate) return (sim > MIN_SIM_3D) ? sim : 0.0f;
        return sim;
    }

    // line3D.cc:2405-2446
    bool unused(const Seg2D& a, const Seg2D& b)
    {
        auto& ua = used[a];
        if (ua.count(b)) return false;
        ua.insert(b);
        used[b].insert(a);
        return true;
    }
    int getLocalID(const Seg2D& s)
    {
        auto f = global2local.find(s);
        if (f != global2local.end()) return f->second;
        const int id = localID++;
        global2local[s] = id;
        local2global[id] = s;
        return id;
    }

    // line3D.cc:2275-2402 (serial order), including the links to collinear segments of line3D.cc:2328-2396
    // when collinearity_t > 0 (the product builds the collinearity-off case only)
    float collinearity_t = -1.0f;
    void computingAffinityMatrix()
    {
        A.clear();
        global2local.clear();
        local2global.clear();
        localID = 0;
        used.clear();
        const bool collin_on = collinearity_t > EPS;
        for (size_t i = 0; i < est3D.size(); ++i) {
            const Seg3D& s = est3D[i].seg;
            const Match m = est3D[i].m;
            const Seg2D seg(m.src_cam, m.src_seg);
            int id1 = -1;
            bool found_aff = false;
            const std::list<Match>& ml = matches[m.src_cam][m.src_seg];
            for (const Match& m2 : ml) {
                const Seg2D seg2(m2.tgt_cam, m2.tgt_seg);
                const float sim = similarity(s, m, seg2, false);
                if (sim > MIN_AFFINITY && unused(seg, seg2)) {
                    if (id1 < 0) id1 = getLocalID(seg);
                    const int id2 = getLocalID(seg2);
                    A.push_back({id1, id2, sim});
                    A.push_back({id2, id1, sim});
                    found_aff = true;
                    if (collin_on) {  // links to the segments collinear to the target (line3D.cc:2328-2360)
                        for (unsigned c : views[seg2.first]->collinearSegments(seg2.second)) {
                            const Seg2D seg2c(seg2.first, c);
                            const float simc = similarity(s, m, seg2c, false);
                            if (simc > MIN_AFFINITY && unused(seg, seg2c)) {
                                const int id2c = getLocalID(seg2c);
                                A.push_back({id1, id2c, simc});
                                A.push_back({id2c, id1, simc});
                            }
                        }
                    }
                }
            }
            if (found_aff && id1 >= 0 && collin_on) {  // line3D.cc:2364-2396
                for (unsigned c : views[seg.first]->collinearSegments(seg.second)) {
                    const Seg2D segc(seg.first, c);
                    const float simc = similarity(s, m, segc, false);
                    if (simc > MIN_AFFINITY && unused(seg, segc)) {
                        const int idc = getLocalID(segc);
                        A.push_back({id1, idc, simc});
                        A.push_back({idc, id1, simc});
                    }
                }
            }
        }
        used.clear();
    }

    // clustering.cc:7-48 + universe.h:59-117 (Felzenszwalb-Huttenlocher union-find, c=3)
    void clusterSegments()
    {
        cluster_ids.clear();
        const int n = (int)global2local.size();
        if (A.empty()) return;
        A.sort([](const Edge& a, const Edge& b) { return a.w < b.w; });
        struct El {
            int rank, id, size;
        };
        std::vector<El> el(n);
        for (int i = 0; i < n; ++i) el[i] = {0, i, 1};
        auto find = [&](int x) {
            int y = x;
            while (y != el[y].id) y = el[y].id;
            el[x].id = y;
            return y;
        };
        const float c = 3.0f;
        std::vector<float> thr(n, c);
        for (const Edge& e : A) {
            int a = find(e.i), b = find(e.j);
            if (a != b && e.w <= thr[a] && e.w <= thr[b]) {
                if (el[a].rank > el[b].rank) {
                    el[b].id = a;
                    el[a].size += el[b].size;
                } else {
                    el[a].id = b;
                    el[b].size += el[a].size;
                    if (el[a].rank == el[b].rank) el[b].rank++;
                }
                a = find(a);
                thr[a] = e.w + c / (float)el[a].size;
            }
        }
        cluster_ids.resize(n);
        for (auto& kv : local2global) cluster_ids[kv.first] = find(kv.first);  // line3D.cc:2522-2535
    }

    // line3D.cc:2018-2118 (up to and including clustering; the 3-D line tail is out of scope)
    void reconstruct(float collin_t = -1.0f)
    {
        const double T0 = now();
        A_snapshot.clear();
        local2global_snapshot.clear();
        cluster_ids.clear();
        if (est3D.empty()) return;
        const float prev_collin_t = collinearity_t;  // line3D.cc:2041-2042
        collinearity_t = collin_t;
        translate();
        // line3D.cc:2068-2072, 2250-2270: every view in views_ (a view added later keeps an empty table
        // until the threshold changes, as in the reference)
        if (collinearity_t > EPS && (prev_collin_t < EPS || std::fabs((double)(prev_collin_t - collinearity_t)) > EPS))
            for (auto& kv : views) kv.second->findCollinearSegments(collinearity_t);
        std::vector<float> sd;
        for (auto& kv : views) {
            const bool active =
                std::find(view_order.begin(), view_order.end(), kv.second->id) != view_order.end();
            if (kv.second->median_depth > EPS && active) sd.push_back(kv.second->median_depth);
        }
        if (!sd.empty()) {
            std::sort(sd.begin(), sd.end());
            med_scene_depth_lines = sd[sd.size() / 2];
        } else
            med_scene_depth_lines = 0.0f;
        const double T1 = now();
        computingAffinityMatrix();
        tm.affinity += now() - T1;
        A_snapshot.assign(A.begin(), A.end());
        local2global_snapshot.resize(local2global.size());
        for (auto& kv : local2global) local2global_snapshot[kv.first] = kv.second;
        const double T2 = now();
        clusterSegments();
        tm.cluster += now() - T2;
        A.clear();
        untranslate();
        tm.reconstruct += now() - T0;
    }
};

}  // namespace orc

// ------------------------------------------------------------------------------------------
// flat C interface for ctypes (tests / bench cpu_baseline only)
// ------------------------------------------------------------------------------------------
using namespace orc;

struct OrcRec {  // one list entry, 36 bytes
    uint32_t tgt_cam, tgt_seg;
    float overlap, score, d_p1, d_p2, d_q1, d_q2;
    uint32_t flags;  // bit0: orientation flag (true for inverse matches)
};
struct OrcEntry {  // estimated_position3D_ row
    uint32_t src_cam, src_seg, tgt_cam, tgt_seg;
    float overlap, score, d_p1, d_p2, d_q1, d_q2;
    float length;
    uint32_t pad;
    double P1[3], P2[3], dir[3];
};

static M3 toM3(const double* a)
{
    M3 m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m.m[i][j] = a[i * 3 + j];
    return m;
}

extern "C" {

void* orc_create(int max_img_width, int neighbors_by_worldpoints)
{
    return new Line3D(max_img_width, neighbors_by_worldpoints != 0);
}
void orc_destroy(void* h) { delete (Line3D*)h; }
void orc_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}
int orc_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_snapshot(void* h, int on) { ((Line3D*)h)->snapshot = on != 0; }

int orc_add_image(void* h, uint32_t camID, const double* K, const double* R, const double* t,
                  unsigned w, unsigned hh, float median_depth, const uint32_t* wn, int nwn,
                  const float* segs, int nsegs)
{
    std::list<uint32_t> l(wn, wn + nwn);
    std::vector<Seg2f> s(nsegs);
    for (int i = 0; i < nsegs; ++i) s[i] = {segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]};
    return ((Line3D*)h)->addImage(camID, toM3(K), toM3(R), {t[0], t[1], t[2]}, w, hh, median_depth, l, s);
}
int orc_update_image(void* h, uint32_t camID, const double* R, const double* t, float median_depth,
                     const uint32_t* wn, int nwn)
{
    std::list<uint32_t> l(wn, wn + nwn);
    return ((Line3D*)h)->updateImage(camID, toM3(R), {t[0], t[1], t[2]}, median_depth, l);
}
int orc_delete_image(void* h, uint32_t camID) { return ((Line3D*)h)->deleteImage(camID) ? 1 : 0; }
// mirrors the per-cycle resets L3DPPing::Run performs before delete/add (L3DPPing.cpp:98-103)
void orc_begin_cycle(void* h)
{
    Line3D* L = (Line3D*)h;
    L->views2worldpoints.clear();
    L->worldpoints2views.clear();
    L->delete_cams.clear();
    L->add_cams.clear();
}
void orc_match_images(void* h, float sp, float sa, unsigned nn, float eo, int knn, float crd)
{
    ((Line3D*)h)->matchImages(sp, sa, nn, eo, knn, crd);
}
void orc_reconstruct(void* h) { ((Line3D*)h)->reconstruct(); }
void orc_reconstruct_collin(void* h, float collinearity_t) { ((Line3D*)h)->reconstruct(collinearity_t); }

int orc_num_pairs(void* h) { return (int)((Line3D*)h)->pair_log.size(); }
void orc_get_pairs(void* h, uint32_t* out)
{
    Line3D* L = (Line3D*)h;
    for (size_t i = 0; i < L->pair_log.size(); ++i) {
        out[2 * i] = L->pair_log[i].first;
        out[2 * i + 1] = L->pair_log[i].second;
    }
}
uint64_t orc_pair_tests(void* h) { return ((Line3D*)h)->pair_tests; }

static const std::vector<std::list<Match>>* pick(Line3D* L, uint32_t cam, int which)
{
    auto& mp = which == 0 ? L->snap_scored : L->matches;
    auto f = mp.find(cam);
    return f == mp.end() ? nullptr : &f->second;
}
// which: 0 = lists right after scoring (pre-filter), 1 = current (filtered) lists
uint64_t orc_list_total(void* h, uint32_t cam, int which)
{
    auto* v = pick((Line3D*)h, cam, which);
    uint64_t n = 0;
    if (v)
        for (auto& l : *v) n += l.size();
    return n;
}
int orc_get_lists(void* h, uint32_t cam, int which, uint32_t* row_off, OrcRec* out)
{
    auto* v = pick((Line3D*)h, cam, which);
    if (!v) return -1;
    uint32_t n = 0;
    for (size_t i = 0; i < v->size(); ++i) {
        row_off[i] = n;
        for (const Match& m : (*v)[i]) {
            out[n].tgt_cam = m.tgt_cam;
            out[n].tgt_seg = m.tgt_seg;
            out[n].overlap = m.overlap;
            out[n].score = m.score3D;
            out[n].d_p1 = m.d_p1;
            out[n].d_p2 = m.d_p2;
            out[n].d_q1 = m.d_q1;
            out[n].d_q2 = m.d_q2;
            out[n].flags = m.orient ? 1u : 0u;
            ++n;
        }
    }
    row_off[v->size()] = n;
    return 0;
}
int orc_num_entries(void* h) { return (int)((Line3D*)h)->est3D.size(); }
void orc_get_entries(void* h, OrcEntry* out)
{
    Line3D* L = (Line3D*)h;
    for (size_t i = 0; i < L->est3D.size(); ++i) {
        const Entry& e = L->est3D[i];
        OrcEntry& o = out[i];
        o.src_cam = e.m.src_cam;
        o.src_seg = e.m.src_seg;
        o.tgt_cam = e.m.tgt_cam;
        o.tgt_seg = e.m.tgt_seg;
        o.overlap = e.m.overlap;
        o.score = e.m.score3D;
        o.d_p1 = e.m.d_p1;
        o.d_p2 = e.m.d_p2;
        o.d_q1 = e.m.d_q1;
        o.d_q2 = e.m.d_q2;
        o.length = e.seg.length;
        o.pad = 0;
        o.P1[0] = e.seg.P1.x; o.P1[1] = e.seg.P1.y; o.P1[2] = e.seg.P1.z;
        o.P2[0] = e.seg.P2.x; o.P2[1] = e.seg.P2.y; o.P2[2] = e.seg.P2.z;
        o.dir[0] = e.seg.dir.x; o.dir[1] = e.seg.dir.y; o.dir[2] = e.seg.dir.z;
    }
}
int orc_num_edges(void* h) { return (int)((Line3D*)h)->A_snapshot.size(); }
void orc_get_edges(void* h, int* ij, float* w)
{
    Line3D* L = (Line3D*)h;
    for (size_t i = 0; i < L->A_snapshot.size(); ++i) {
        ij[2 * i] = L->A_snapshot[i].i;
        ij[2 * i + 1] = L->A_snapshot[i].j;
        w[i] = L->A_snapshot[i].w;
    }
}
// SparseMatrix::SparseMatrix (sparsematrix.cc:8-61) over arbitrary (i,j,w) triples: std::list::sort with
// sortCLEdgesByCol / sortCLEdgesByRow (clustering.h:70-78), float4{i,j,w/norm,0}, start index of every
// row/column or -1.  Stand-alone so that the known-answer tests can feed hand-made edge lists.
void orc_sparse_matrix(const int* ij, const float* w, int ne, int n, float norm, int sort_by_row, float* entries4,
                       int* start_indices)
{
    struct E {
        int i, j;
        float w;
    };
    std::list<E> l;
    for (int e = 0; e < ne; ++e) l.push_back({ij[2 * e], ij[2 * e + 1], w[e]});
    if (sort_by_row)
        l.sort([](const E& a, const E& b) { return (a.i < b.i) || (a.i == b.i && a.j < b.j); });
    else
        l.sort([](const E& a, const E& b) { return (a.j < b.j) || (a.j == b.j && a.i < b.i); });
    for (int r = 0; r < n; ++r) start_indices[r] = -1;
    int pos = 0, current = -1;
    for (const E& e : l) {
        entries4[4 * pos + 0] = (float)e.i;
        entries4[4 * pos + 1] = (float)e.j;
        entries4[4 * pos + 2] = e.w / norm;
        entries4[4 * pos + 3] = 0.0f;
        const int rc = sort_by_row ? e.i : e.j;
        if (current != rc) {
            start_indices[rc] = pos;
            current = rc;
        }
        ++pos;
    }
}
// View::findCollinCPU (view.cc:238-293) with View::pointOnSegment (view.cc:321-327) and
// View::distance_point2line_2D (view.cc:296-299): out[r*n + c] = 1 iff segment c is collinear to r
// (no mutual overlap, all four point-to-line distances below dist_t).
void orc_find_collinear(const float* lines, int n, float dist_t, char* out)
{
    auto on_seg = [](const V3& p1, const V3& p2, const V3& x) {
        const double v1x = p1.x - x.x, v1y = p1.y - x.y, v2x = p2.x - x.x, v2y = p2.y - x.y;
        return (v1x * v2x + v1y * v2y) < EPS;
    };
    auto dist = [](const V3& l, const V3& p) {
        return (float)std::fabs((l.x * p.x + l.y * p.y + l.z) / sqrtf((float)(l.x * l.x + l.y * l.y)));
    };
#pragma omp parallel for schedule(dynamic, 8)
    for (int r = 0; r < n; ++r) {
        const V3 p0 = {(double)lines[4 * r], (double)lines[4 * r + 1], 1.0};
        const V3 p1 = {(double)lines[4 * r + 2], (double)lines[4 * r + 3], 1.0};
        const V3 line1 = cross(p0, p1);
        for (int c = 0; c < n; ++c) {
            out[(size_t)r * n + c] = 0;
            if (r == c) continue;
            const V3 q0 = {(double)lines[4 * c], (double)lines[4 * c + 1], 1.0};
            const V3 q1 = {(double)lines[4 * c + 2], (double)lines[4 * c + 3], 1.0};
            const V3 line2 = cross(q0, q1);
            if (on_seg(p0, p1, q0) || on_seg(p0, p1, q1) || on_seg(q0, q1, p0) || on_seg(q0, q1, p1)) continue;
            const float d1 = (float)std::fmax((double)dist(line1, q0), (double)dist(line1, q1));
            const float d2 = (float)std::fmax((double)dist(line2, p0), (double)dist(line2, p1));
            if ((float)std::fmax((double)d1, (double)d2) < dist_t) out[(size_t)r * n + c] = 1;
        }
    }
}
int orc_num_local(void* h) { return (int)((Line3D*)h)->local2global_snapshot.size(); }
void orc_get_local2global(void* h, uint32_t* cam_seg)
{
    Line3D* L = (Line3D*)h;
    for (size_t i = 0; i < L->local2global_snapshot.size(); ++i) {
        cam_seg[2 * i] = L->local2global_snapshot[i].first;
        cam_seg[2 * i + 1] = L->local2global_snapshot[i].second;
    }
}
int orc_get_cluster_ids(void* h, int* out)
{
    Line3D* L = (Line3D*)h;
    for (size_t i = 0; i < L->cluster_ids.size(); ++i) out[i] = L->cluster_ids[i];
    return (int)L->cluster_ids.size();
}
// info: C[3] (current, untranslated), k, median_depth, median_sigma
int orc_get_view_info(void* h, uint32_t cam, double* C, float* kmm)
{
    Line3D* L = (Line3D*)h;
    auto f = L->views.find(cam);
    if (f == L->views.end()) return -1;
    C[0] = f->second->C.x; C[1] = f->second->C.y; C[2] = f->second->C.z;
    kmm[0] = f->second->k;
    kmm[1] = f->second->median_depth;
    kmm[2] = f->second->median_sigma;
    return 0;
}
// test hook: the camera exactly as the last matchImages saw it (RtKinv row-major, translated centre)
int orc_get_match_camera(void* h, uint32_t cam, double* RtKinv9, double* C3)
{
    Line3D* L = (Line3D*)h;
    auto f = L->views.find(cam);
    if (f == L->views.end()) return -1;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) RtKinv9[3 * i + j] = f->second->RtKinv.m[i][j];
    C3[0] = f->second->C_match.x; C3[1] = f->second->C_match.y; C3[2] = f->second->C_match.z;
    return 0;
}
// test hook: the fundamental matrix the last matchImages cached for the ordered pair (src, tgt), row-major
int orc_get_fundamental(void* h, uint32_t src, uint32_t tgt, double* F9)
{
    Line3D* L = (Line3D*)h;
    auto fs = L->fundamentals.find(src);
    if (fs == L->fundamentals.end() || !fs->second.count(tgt)) return -1;
    const M3& F = fs->second[tgt];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) F9[3 * i + j] = F.m[i][j];
    return 0;
}
int orc_get_neighbors(void* h, uint32_t cam, uint32_t* out, int cap)
{
    Line3D* L = (Line3D*)h;
    auto f = L->visual_neighbors.find(cam);
    if (f == L->visual_neighbors.end()) return -1;
    int n = 0;
    for (uint32_t v : f->second) {
        if (n < cap) out[n] = v;
        ++n;
    }
    return n;
}
float orc_med_scene_depth_lines(void* h) { return ((Line3D*)h)->med_scene_depth_lines; }
void orc_get_translation(void* h, double* t)
{
    Line3D* L = (Line3D*)h;
    t[0] = L->translation.x; t[1] = L->translation.y; t[2] = L->translation.z;
}
// timers: match, orient, score, inverse, filter, update, affinity, cluster, match_images, reconstruct
void orc_get_timers(void* h, double* out)
{
    const Timers& t = ((Line3D*)h)->tm;
    const double v[10] = {t.match, t.orient, t.score, t.inverse, t.filter,
                          t.update, t.affinity, t.cluster, t.match_images, t.reconstruct};
    memcpy(out, v, sizeof(v));
}

// ---- primitives exposed for known-answer tests ----
float orc_kat_expf(float x) { return orc_expf(x); }
double orc_kat_acos(double x) { return orc_acos(x); }
float orc_kat_acosf(float x) { return orc_acosf(x); }
double orc_kat_sin(double x) { return orc_sin(x); }
float orc_kat_mutual_overlap(const double* p12)
{
    V3 pts[4];
    for (int i = 0; i < 4; ++i) pts[i] = {p12[3 * i], p12[3 * i + 1], p12[3 * i + 2]};
    return Line3D::mutualOverlap(pts);
}
void orc_kat_fundamental(const double* K1, const double* R1, const double* t1, const double* K2,
                         const double* R2, const double* t2, double* F)
{
    Line3D L(640, false);
    View a, b;
    a.init(0, toM3(K1), toM3(R1), {t1[0], t1[1], t1[2]}, 640, 480, 1.0f);
    b.init(1, toM3(K2), toM3(R2), {t2[0], t2[1], t2[2]}, 640, 480, 1.0f);
    const M3 f = L.fundamental(&a, &b);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) F[i * 3 + j] = f.m[i][j];
}
void orc_kat_inverse3(const double* A, double* out)
{
    const M3 r = inverse(toM3(A));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = r.m[i][j];
}
float orc_kat_angle(const double* d1, const double* d2)
{
    Seg3D a, b;
    a.dir = {d1[0], d1[1], d1[2]};
    b.dir = {d2[0], d2[1], d2[2]};
    return Line3D::angleBetweenSeg3D(a, b, true);
}
float orc_kat_dist_point_line(const double* P1, const double* P2, const double* P)
{
    Seg3D s({P1[0], P1[1], P1[2]}, {P2[0], P2[1], P2[2]});
    return s.distPointLine({P[0], P[1], P[2]});
}
// F-H clustering on an arbitrary edge list (known-answer tests of clustering.cc semantics)
int orc_kat_cluster(const int* ij, const float* w, int ne, int n, int* out)
{
    Line3D L(640, false);
    for (int i = 0; i < ne; ++i) L.A.push_back({ij[2 * i], ij[2 * i + 1], w[i]});
    for (int i = 0; i < n; ++i) {
        L.global2local[Seg2D(0, (uint32_t)i)] = i;
        L.local2global[i] = Seg2D(0, (uint32_t)i);
    }
    L.clusterSegments();
    for (size_t i = 0; i < L.cluster_ids.size(); ++i) out[i] = L.cluster_ids[i];
    return (int)L.cluster_ids.size();
}


// ---- cudawrapper-level restatements (tests of l3d_match_lines / l3d_score_matches) ----
// Runs translate() + getFundamentalMatrix + matchingCPU(src,tgt) only (no orientation filter, no
// scoring): what L3DPP::match_lines_GPU has to deliver (src/line3D.cc:1237-1271).  Returns the
// matrices the GPU entry point receives (row-major F, RtKinv of both views, translated centres).
int orc_match_only(void* h, uint32_t src, uint32_t tgt, float epi_overlap, int knn, double* F9, double* Ms9,
                   double* Mt9, double* Cs3, double* Ct3)
{
    Line3D* L = (Line3D*)h;
    if (!L->views.count(src) || !L->views.count(tgt)) return -1;
    L->epipolar_overlap = (float)std::fmin(std::fabs((double)epi_overlap), (double)0.99f);
    L->kNN = knn;
    L->translate();
    View* vs = L->views[src];
    View* vt = L->views[tgt];
    const M3 F = L->fundamental(vs, vt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            F9[3 * i + j] = F.m[i][j];
            Ms9[3 * i + j] = vs->RtKinv.m[i][j];
            Mt9[3 * i + j] = vt->RtKinv.m[i][j];
        }
    Cs3[0] = vs->C.x; Cs3[1] = vs->C.y; Cs3[2] = vs->C.z;
    Ct3[0] = vt->C.x; Ct3[1] = vt->C.y; Ct3[2] = vt->C.z;
    L->matchingCPU(src, tgt, F);
    return 0;
}

// Line3D::scoringCPU new-match branch (src/line3D.cc:1513-1547) over the packed buffers of
// Line3D::scoringGPU (src/line3D.cc:1582-1623): matches = {srcSeg, tgtCam, d_p1, d_p2} per entry,
// ranges = {first,last} per segment, regs_tgt = {sigma_tgt(P1), sigma_tgt(P2)}.
void orc_score_packed(const float* lines, uint32_t n_lines, const float* matches, uint32_t n_matches,
                      const int32_t* ranges, const float* regs_tgt, const double* RtKinv9, const double* C3,
                      float two_sigA_sqr, float k, float min_sim, float* scores)
{
    Line3D L(640, false);
    L.two_sigA_sqr = two_sigA_sqr;
    View v;
    v.RtKinv = toM3(RtKinv9);
    v.C = {C3[0], C3[1], C3[2]};
    v.lines.resize(n_lines);
    for (uint32_t i = 0; i < n_lines; ++i) v.lines[i] = {lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]};
    for (uint32_t i = 0; i < n_matches; ++i) scores[i] = 0.0f;
    for (uint32_t s = 0; s < n_lines; ++s) {
        const int a = ranges[2 * s], b = ranges[2 * s + 1];
        if (a < 0) continue;
        for (int e = a; e <= b; ++e) {
            const uint32_t seg = (uint32_t)matches[4 * e];
            const uint32_t cam = (uint32_t)matches[4 * e + 1];
            const float d1 = matches[4 * e + 2], d2 = matches[4 * e + 3];
            const Seg3D M3D = v.unproject(seg, d1, d2);
            const float sig1 = d1 * k, sig2 = d2 * k;
            float reg1 = 2.0f * sig1 * sig1, reg2 = 2.0f * sig2 * sig2;
            const float s1t = regs_tgt[2 * e], s2t = regs_tgt[2 * e + 1];
            reg1 = 0.5f * (reg1 + 2.0f * s1t * s1t);
            reg2 = 0.5f * (reg2 + 2.0f * s2t * s2t);
            std::map<uint32_t, float> per_cam;
            float score = 0.0f;
            for (int e2 = a; e2 <= b; ++e2) {
                const uint32_t cam2 = (uint32_t)matches[4 * e2 + 1];
                if (cam2 == cam) continue;
                const Seg3D S2 = v.unproject((uint32_t)matches[4 * e2], matches[4 * e2 + 2], matches[4 * e2 + 3]);
                float sim = 0.0f;
                if (!(M3D.length < EPS || S2.length < EPS)) {
                    const float dd1 = d1 - matches[4 * e2 + 2], dd2 = d2 - matches[4 * e2 + 3];
                    const float sim_p = std::fmin(orc_expf(-dd1 * dd1 / reg1), orc_expf(-dd2 * dd2 / reg2));
                    const float angle = Line3D::angleBetweenSeg3D(M3D, S2, true);
                    const float sim_a = orc_expf(-angle * angle / two_sigA_sqr);
                    const float sm = std::fmin(sim_a, sim_p);
                    sim = (sm > min_sim) ? sm : 0.0f;
                }
                auto f = per_cam.find(cam2);
                if (f != per_cam.end()) {
                    if (sim > f->second) {
                        score -= f->second;
                        score += sim;
                        f->second = sim;
                    }
                } else {
                    score += sim;
                    per_cam[cam2] = sim;
                }
            }
            scores[e] = score;
        }
    }
}

}  // extern "C"
