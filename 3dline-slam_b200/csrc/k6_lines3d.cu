// k6_lines3d.cu -- K6: the cluster -> 3-D line tail of Line3D::reconstruct3Dlines (exact TU, -fmad=false).
//
// After the (host, unchanged) graph clustering the reference turns every cluster that is seen by at least
// visibility_t cameras into 3-D line segments:
//   Line3D::get3DlineFromCluster       src/line3D.cc:2578-2641  centre of gravity of the hypotheses' end points,
//                                      principal direction of their 3x3 scatter matrix (Eigen::JacobiSVD),
//                                      reference view = camera of the longest 2-D segment
//   Line3D::project2DsegmentOnto3Dline src/line3D.cc:2644-2687  closest points between the 3-D line and the viewing
//                                      rays of a member's 2-D end points
//   Line3D::findCollinearSegments_return src/line3D.cc:2763-2870 sort the projected points along the line, sweep:
//                                      a 3-D segment wherever members of >= 3 cameras overlap
//   Line3D::filterTinySegments         src/line3D.cc:2724-2760  drop segments that project shorter than
//                                      View::min_line_length_ into the reference view (View::projectedLongEnough,
//                                      View::project, src/view.cc:403-456)
//   Line3D::performTranslation         src/line3D.cc:697-720    back to the untranslated frame
// One thread per cluster, in the reference's own order of operations (clusters are small: a handful of members),
// so that the result can be compared bit for bit with the reference's sources compiled in oracle/_ref; the
// eigen-solver is the cyclic Jacobi iteration of oracle/standin/l3d_standin_eigen.h (any SVD gives the same
// direction up to rounding and sign, and the sign does not reach the output).
#include "exact.cuh"
#include "internal.h"

namespace l3d {

#define L3D_EPS 1e-12

struct V3d {
    double x, y, z;
};
__device__ __forceinline__ V3d v3(double x, double y, double z) { return V3d{x, y, z}; }
__device__ __forceinline__ V3d vadd(const V3d& a, const V3d& b) { return V3d{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3d vsub(const V3d& a, const V3d& b) { return V3d{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3d vscale(double s, const V3d& a) { return V3d{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double vdot(const V3d& a, const V3d& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double vnorm(const V3d& a) { return sqrt(vdot(a, a)); }
__device__ __forceinline__ V3d vnormalized(const V3d& a)
{
    const double n = vnorm(a);
    return V3d{a.x / n, a.y / n, a.z / n};
}

// Segment3D::Segment3D(P1, P2) (include/segment3D.h:58-77): a segment shorter than 1e-12 is the null segment
struct Seg3 {
    V3d P1, P2, dir;
    bool valid;
};
__device__ __forceinline__ Seg3 make_seg3(const V3d& a, const V3d& b)
{
    Seg3 s;
    const float len = (float)vnorm(vsub(a, b));
    if (len > L3D_EPS) {
        s.P1 = a;
        s.P2 = b;
        s.dir = vnormalized(vsub(b, a));
        s.valid = true;
    } else {
        s.P1 = s.P2 = s.dir = v3(0, 0, 0);
        s.valid = false;
    }
    return s;
}

// symmetric 3x3 -> eigenvector of the largest |eigenvalue| (JacobiSVD(Scat).matrixU().col(argmax S))
__device__ void principal_direction(const double A[3][3], double dir[3])
{
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[i][j] = 0.5 * (A[i][j] + A[j][i]);
            v[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 3; ++q) off += a[p][q] * a[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    // columns by descending |eigenvalue| (exchange sort of the stand-in), then S.maxCoeff(&pos): first maximum
    int order[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (fabs(a[order[j]][order[j]]) > fabs(a[order[i]][order[i]])) {
                const int tmp = order[i];
                order[i] = order[j];
                order[j] = tmp;
            }
    int at = 0;
    for (int i = 1; i < 3; ++i)
        if (fabs(a[order[i]][order[i]]) > fabs(a[order[at]][order[at]])) at = i;
    for (int i = 0; i < 3; ++i) dir[i] = v[i][order[at]];
}

// View::project (src/view.cc:403-421)
__device__ __forceinline__ void project_view(const TailView& V, const V3d& P, double& px, double& py)
{
    double q0 = V.R[0] * P.x + V.R[1] * P.y + V.R[2] * P.z + V.t[0];
    double q1 = V.R[3] * P.x + V.R[4] * P.y + V.R[5] * P.z + V.t[1];
    double q2 = V.R[6] * P.x + V.R[7] * P.y + V.R[8] * P.z + V.t[2];
    const double xn = (1.0 * q0 + 0.0 * q2) / q2;
    const double yn = (1.0 * q1 + 0.0 * q2) / q2;
    q0 = V.K[0] * xn + V.K[1] * yn + V.K[2] * 1.0;
    q1 = V.K[3] * xn + V.K[4] * yn + V.K[5] * 1.0;
    q2 = V.K[6] * xn + V.K[7] * yn + V.K[8] * 1.0;
    px = q0 / q2;
    py = q1 / q2;
}

__global__ void __launch_bounds__(64) k6_lines3d_kernel(
    uint32_t n_clusters, const uint32_t* __restrict__ cl_off, const uint32_t* __restrict__ members,
    const EntryDev* __restrict__ entries, const float4* __restrict__ segs, const SegRays* __restrict__ rays,
    const uint32_t* __restrict__ seg_view, const TailView* __restrict__ views, double tx, double ty, double tz,
    double* __restrict__ Lbuf, double* __restrict__ LCbuf, double* __restrict__ pts, float* __restrict__ dist,
    uint32_t* __restrict__ ord, unsigned char* __restrict__ okflag, uint32_t* __restrict__ camtab,
    uint32_t* __restrict__ out_n, uint32_t* __restrict__ out_ref, double* __restrict__ out_seg)
{
    const uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_clusters) return;
    const uint32_t m0 = cl_off[ci], m = cl_off[ci + 1] - m0, n = 2 * m;
    double* L = Lbuf + 6 * (size_t)m0;    // 3 x n, row-major
    double* LC = LCbuf + 6 * (size_t)m0;  // 3 x n
    double* pt = pts + 6 * (size_t)m0;    // n projected points
    float* dd_ = dist + 2 * (size_t)m0;
    uint32_t* od = ord + 2 * (size_t)m0;
    unsigned char* ok = okflag + (size_t)m0;  // per member: projection succeeded; later: line open
    uint32_t* ct = camtab + 2 * (size_t)m0;   // (view, open count) pairs

    // ---- Line3D::get3DlineFromCluster ----
    V3d P = v3(0, 0, 0);
    uint32_t ref_view = views[0].cam_id * 0u;  // reference_cam = 0 unless a segment is longer than 0
    bool have_ref = false;
    float max_len = 0.0f;
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t g = members[m0 + i];
        const EntryDev& e = entries[g];
        P = vadd(P, v3(e.P1[0], e.P1[1], e.P1[2]));
        P = vadd(P, v3(e.P2[0], e.P2[1], e.P2[2]));
        L[0 * n + 2 * i] = e.P1[0]; L[1 * n + 2 * i] = e.P1[1]; L[2 * n + 2 * i] = e.P1[2];
        L[0 * n + 2 * i + 1] = e.P2[0]; L[1 * n + 2 * i + 1] = e.P2[1]; L[2 * n + 2 * i + 1] = e.P2[2];
        const float4 c = segs[g];
        const float length_sqr = fa(fm(fs(c.x, c.z), fs(c.x, c.z)), fm(fs(c.y, c.w), fs(c.y, c.w)));
        if (length_sqr > max_len) {
            max_len = length_sqr;
            ref_view = seg_view[g];
            have_ref = true;
        }
    }
    const double nd = (double)(int)n;
    P = v3(P.x / nd, P.y / nd, P.z / nd);
    // Scat = L * (I - (1/n) 1 1^T) * L^T, evaluated as (L * C) * L^T with left-to-right sums
    const double inv_n = 1.0 / nd;
    const double c_diag = 1.0 - inv_n * 1.0, c_off = 0.0 - inv_n * 1.0;
    for (int r = 0; r < 3; ++r)
        for (uint32_t j = 0; j < n; ++j) {
            double acc = L[r * n + 0] * (j == 0 ? c_diag : c_off);
            for (uint32_t k = 1; k < n; ++k) acc = acc + L[r * n + k] * (k == j ? c_diag : c_off);
            LC[r * n + j] = acc;
        }
    double Scat[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double acc = LC[r * n + 0] * L[c * n + 0];
            for (uint32_t j = 1; j < n; ++j) acc = acc + LC[r * n + j] * L[c * n + j];
            Scat[r][c] = acc;
        }
    double dv[3];
    principal_direction(Scat, dv);
    const V3d dir = vnormalized(v3(dv[0], dv[1], dv[2]));
    const Seg3 line = make_seg3(vsub(P, dir), vadd(P, dir));

    // ---- Line3D::findCollinearSegments_return ----
    const V3d COG = vscale(0.5, vadd(line.P1, line.P2));
    float distToCOG = 0.0f;
    V3d border = v3(0, 0, 0);
    uint32_t n_ok = 0;
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t g = members[m0 + i];
        const TailView& V = views[seg_view[g]];
        // Line3D::project2DsegmentOnto3Dline
        const V3d Pp = line.P1, u = line.dir, Q = v3(V.C[0], V.C[1], V.C[2]);
        const SegRays rr = rays[g];
        const V3d v1 = v3(rr.r1[0], rr.r1[1], rr.r1[2]), v2 = v3(rr.r2[0], rr.r2[1], rr.r2[2]);
        const V3d w = vsub(Pp, Q);
        const double a = vdot(u, u), b1 = vdot(u, v1), b2 = vdot(u, v2), c1 = vdot(v1, v1), c2 = vdot(v2, v2);
        const double d = vdot(u, w), e1 = vdot(v1, w), e2 = vdot(v2, w);
        const double denom1 = a * c1 - b1 * b1, denom2 = a * c2 - b2 * b2;
        ok[i] = 0;
        if (fabs(denom1) > L3D_EPS && fabs(denom2) > L3D_EPS) {
            const double s1 = (b1 * e1 - c1 * d) / denom1, s2 = (b2 * e2 - c2 * d) / denom2;
            const Seg3 proj = make_seg3(vadd(Pp, vscale(s1, u)), vadd(Pp, vscale(s2, u)));
            ok[i] = 1;
            ++n_ok;
            pt[6 * i + 0] = proj.P1.x; pt[6 * i + 1] = proj.P1.y; pt[6 * i + 2] = proj.P1.z;
            pt[6 * i + 3] = proj.P2.x; pt[6 * i + 4] = proj.P2.y; pt[6 * i + 5] = proj.P2.z;
            float dc = (float)vnorm(vsub(proj.P1, COG));
            if (dc > distToCOG) {
                distToCOG = dc;
                border = proj.P1;
            }
            dc = (float)vnorm(vsub(proj.P2, COG));
            if (dc > distToCOG) {
                distToCOG = dc;
                border = proj.P2;
            }
        }
    }
    uint32_t n_out = 0;
    if (2 * n_ok >= 6) {
        // linePoints: the end points of the successful members in member order; stable sort by distance to the border
        uint32_t np = 0;
        for (uint32_t i = 0; i < m; ++i) {
            if (!ok[i]) continue;
            for (uint32_t e = 0; e < 2; ++e) {
                const uint32_t pid = 2 * i + e;
                const V3d p = v3(pt[3 * pid], pt[3 * pid + 1], pt[3 * pid + 2]);
                const float db = (float)vnorm(vsub(p, border));
                // insertion keeps equal distances in insertion order, like std::list::sort
                uint32_t pos = np;
                while (pos > 0 && dd_[pos - 1] > db) {
                    dd_[pos] = dd_[pos - 1];
                    od[pos] = od[pos - 1];
                    --pos;
                }
                dd_[pos] = db;
                od[pos] = pid;
                ++np;
            }
        }
        // sweep: a member's first point opens its line, the second closes it; `open` counts lines per camera
        for (uint32_t i = 0; i < m; ++i) ok[i] = 0;  // now: line i is open
        uint32_t n_cams = 0;                         // cameras with an open line
        uint32_t tab = 0;                            // used entries of the (view, count) table
        bool opened = false;
        V3d start = v3(0, 0, 0);
        for (uint32_t q = 0; q < np; ++q) {
            const uint32_t pid = od[q], li = pid >> 1;
            const uint32_t view = seg_view[members[m0 + li]];
            uint32_t slot = 0;
            while (slot < tab && ct[2 * slot] != view) ++slot;
            if (slot == tab) {
                ct[2 * tab] = view;
                ct[2 * tab + 1] = 0;
                ++tab;
            }
            if (!ok[li]) {
                ok[li] = 1;
                if (ct[2 * slot + 1]++ == 0) ++n_cams;
            } else {
                ok[li] = 0;
                if (--ct[2 * slot + 1] == 0) --n_cams;
            }
            const V3d p = v3(pt[3 * pid], pt[3 * pid + 1], pt[3 * pid + 2]);
            if (opened && n_cams < 3) {
                const Seg3 l = make_seg3(start, p);
                double* o = out_seg + 6 * ((size_t)m0 + n_out);
                o[0] = l.P1.x; o[1] = l.P1.y; o[2] = l.P1.z; o[3] = l.P2.x; o[4] = l.P2.y; o[5] = l.P2.z;
                ++n_out;
                opened = false;
            } else if (!opened && n_cams >= 3) {
                start = p;
                opened = true;
            }
        }
    }
    // ---- Line3D::filterTinySegments (in the reference view), then back to the untranslated frame ----
    uint32_t kept = 0;
    if (n_out) {
        const TailView& V = views[have_ref ? ref_view : 0u];
        for (uint32_t k = 0; k < n_out; ++k) {
            double* o = out_seg + 6 * ((size_t)m0 + k);
            const V3d A = v3(o[0], o[1], o[2]), B = v3(o[3], o[4], o[5]);
            double ax, ay, bx, by;
            project_view(V, A, ax, ay);
            project_view(V, B, bx, by);
            const double dx = ax - bx, dy = ay - by;
            if (sqrt(dx * dx + dy * dy) > (double)V.min_line_length) {
                double* w = out_seg + 6 * ((size_t)m0 + kept);
                w[0] = A.x + tx; w[1] = A.y + ty; w[2] = A.z + tz;
                w[3] = B.x + tx; w[4] = B.y + ty; w[5] = B.z + tz;
                ++kept;
            }
        }
    }
    out_n[ci] = kept;
    out_ref[ci] = have_ref ? ref_view : 0xffffffffu;
}

int launch_k6_lines3d(uint32_t n_clusters, const uint32_t* cl_off, const uint32_t* members, const EntryDev* entries,
                      const float4* segs, const SegRays* rays, const uint32_t* seg_view, const TailView* views,
                      const double* t3, double* Lbuf, double* LCbuf, double* pts, float* dist, uint32_t* ord,
                      unsigned char* okflag, uint32_t* camtab, uint32_t* out_n, uint32_t* out_ref, double* out_seg,
                      cudaStream_t st)
{
    if (n_clusters == 0) return 0;
    k6_lines3d_kernel<<<(n_clusters + 63) / 64, 64, 0, st>>>(n_clusters, cl_off, members, entries, segs, rays, seg_view,
                                                             views, t3[0], t3[1], t3[2], Lbuf, LCbuf, pts, dist, ord,
                                                             okflag, camtab, out_n, out_ref, out_seg);
    return 1;
}

}  // namespace l3d
