// host_geom.h -- host-side double geometry of the product: camera set-up (View::View,
// src/view.cc:6-44), View::translate (src/view.cc:539-543), the spatial regulariser
// (src/view.cc:336-343) and Line3D::getFundamentalMatrix (src/line3D.cc:1058-1094), in the canonical
// operation order of SURVEY.md Appendix A.  Built with -ffp-contract=off (no FMA).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace l3d {
namespace hg {

struct V3 {
    double x, y, z;
};
struct M3 {
    double m[9];  // row-major
};

inline V3 operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

inline V3 mul(const M3& A, const V3& v)
{
    return V3{A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
              A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
inline M3 matmul(const M3& A, const M3& B)
{
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
inline M3 transpose(const M3& A)
{
    M3 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T.m[3 * i + j] = A.m[3 * j + i];
    return T;
}
// adjugate over determinant (first-column expansion), entries scaled by 1/det
inline M3 inverse(const M3& A)
{
    double cf[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            cf[3 * i + j] = A.m[3 * i1 + j1] * A.m[3 * i2 + j2] - A.m[3 * i1 + j2] * A.m[3 * i2 + j1];
        }
    const double det = cf[0] * A.m[0] + cf[3] * A.m[3] + cf[6] * A.m[6];
    const double invdet = 1.0 / det;
    M3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[3 * i + j] = cf[3 * j + i] * invdet;
    return R;
}
inline V3 normalized(const V3& a)
{
    const double n = std::sqrt(dot(a, a));
    return V3{a.x / n, a.y / n, a.z / n};
}

// deterministic acos / sin (same published series as the device side, detmath.cuh)
double det_acos(double x);
double det_sin(double x);

struct Camera {
    M3 K, Kinv, R, Rt, RtKinv;
    V3 t, C, pp;
    void init(const double* K9, const double* R9, const double* t3);
    void update(const double* R9, const double* t3);  // View::UpdateView (src/view.cc:62-87)
    void translate(const V3& tv);
    float spatial_regularizer(float r) const;
};

M3 fundamental(const Camera& src, const Camera& tgt);

// Line3D::findVisualNeighborsFromWPs (src/line3D.cc:723-843) for every view: wps[v] = the world
// points view v observes (processWPlist, src/line3D.cc:230-241), median_depth[v] = View::median_depth()
// (0 for a freshly added view, src/view.cc:13).  out[v] = ascending view indices of its neighbours.
void visual_neighbors_from_worldpoints(const std::vector<const Camera*>& cams, const std::vector<float>& median_depth,
                                       const std::vector<std::vector<uint32_t>>& wps, unsigned num_neighbors,
                                       std::vector<std::vector<uint32_t>>& out);

}  // namespace hg
}  // namespace l3d
