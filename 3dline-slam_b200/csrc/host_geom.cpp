// host_geom.cpp -- see host_geom.h.  Compile with -ffp-contract=off.
#include "host_geom.h"

#include <algorithm>
#include <unordered_map>

#include <cstring>

namespace l3d {
namespace hg {

namespace {
// asin(x)/x = sum c_n x^(2n), c_n = (2n)!/(4^n (n!)^2 (2n+1)), n = 0..29
const double kAsin[30] = {
    0x1.0000000000000p+0,  0x1.5555555555555p-3,  0x1.3333333333333p-4,  0x1.6db6db6db6db7p-5,
    0x1.f1c71c71c71c7p-6,  0x1.6e8ba2e8ba2e9p-6,  0x1.1c4ec4ec4ec4fp-6,  0x1.c99999999999ap-7,
    0x1.7a87878787878p-7,  0x1.3fde50d79435ep-7,  0x1.12ef3cf3cf3cfp-7,  0x1.df3bd37a6f4dfp-8,
    0x1.a6863d70a3d71p-8,  0x1.782dda12f684cp-8,  0x1.51ba308d3dcb1p-8,  0x1.31683bdef7bdfp-8,
    0x1.15ee9d45d1746p-8,  0x1.fcaf8fb6db6dbp-9,  0x1.d3d2a8e0dd67dp-9,  0x1.b026f57b13b14p-9,
    0x1.90cb77f60c7cep-9,  0x1.750de64d7d05fp-9,  0x1.5c5f56efaaaabp-9,  0x1.464c0950f7d47p-9,
    0x1.3275586c5f2f0p-9,  0x1.208d3570ae5a6p-9,  0x1.1052bc5fa960ap-9,  0x1.018f963c229bfp-9,
    0x1.e82be60d9127ep-10, 0x1.cf7dea5b6e830p-10};
// (-1)^n/(2n+1)! and (-1)^n/(2n)!
const double kSin[12] = {0x1.0000000000000p+0,   -0x1.5555555555555p-3,  0x1.1111111111111p-7,
                         -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19,  -0x1.ae64567f544e4p-26,
                         0x1.6124613a86d09p-33,  -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49,
                         -0x1.2f49b46814157p-57, 0x1.71b8ef6dcf572p-66,  -0x1.761b41316381ap-75};
const double kCos[12] = {0x1.0000000000000p+0,   -0x1.0000000000000p-1,  0x1.5555555555555p-5,
                         -0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-16,  -0x1.27e4fb7789f5cp-22,
                         0x1.1eed8eff8d898p-29,  -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45,
                         -0x1.6827863b97d97p-53, 0x1.e542ba4020225p-62,  -0x1.0ce396db7f853p-70};
const double kPi = 0x1.921fb54442d18p+1, kPi2 = 0x1.921fb54442d18p+0, kPi4 = 0x1.921fb54442d18p-1;

double horner(const double* c, int n, double z)
{
    double s = c[n - 1];
    for (int i = n - 2; i >= 0; --i) s = s * z + c[i];
    return s;
}
}  // namespace

double det_acos(double x)
{
    const double ax = std::fabs(x);
    if (!(ax <= 1.0)) return NAN;
    if (ax <= 0.5) return kPi2 - x * horner(kAsin, 30, x * x);
    const double z = (1.0 - ax) * 0.5;
    const double a = 2.0 * (std::sqrt(z) * horner(kAsin, 30, z));
    return x > 0.0 ? a : kPi - a;
}

double det_sin(double x)
{
    const double xr = x > kPi2 ? kPi - x : x;
    if (xr <= kPi4) return xr * horner(kSin, 12, xr * xr);
    const double y = kPi2 - xr;
    return horner(kCos, 12, y * y);
}

void Camera::init(const double* K9, const double* R9, const double* t3)
{
    std::memcpy(K.m, K9, sizeof(K.m));
    std::memcpy(R.m, R9, sizeof(R.m));
    t = V3{t3[0], t3[1], t3[2]};
    pp = V3{K.m[2], K.m[5], 1.0};
    Kinv = inverse(K);
    Rt = transpose(R);
    RtKinv = matmul(Rt, Kinv);
    C = mul(Rt, V3{t.x * -1.0, t.y * -1.0, t.z * -1.0});
}

void Camera::update(const double* R9, const double* t3)
{
    std::memcpy(R.m, R9, sizeof(R.m));
    Rt = transpose(R);
    RtKinv = matmul(Rt, Kinv);
    t = V3{t3[0], t3[1], t3[2]};
    C = mul(Rt, V3{t.x * -1.0, t.y * -1.0, t.z * -1.0});
}

void Camera::translate(const V3& tv)
{
    C = C + tv;
    const V3 rc = mul(R, C);
    t = V3{-rc.x, -rc.y, -rc.z};
}

float Camera::spatial_regularizer(float r) const
{
    const V3 pps = V3{pp.x + (double)r, pp.y + 0.0, pp.z + 0.0};
    const V3 a = normalized(mul(RtKinv, pp)), b = normalized(mul(RtKinv, pps));
    const double alpha = det_acos(std::fmin(std::fmax(dot(a, b), -1.0), 1.0));
    return (float)det_sin(alpha);
}

M3 fundamental(const Camera& s, const Camera& tg)
{
    const M3 R = matmul(tg.R, transpose(s.R));
    const V3 tt = tg.t - mul(R, s.t);
    M3 T;
    T.m[0] = 0.0;   T.m[1] = -tt.z; T.m[2] = tt.y;
    T.m[3] = tt.z;  T.m[4] = 0.0;   T.m[5] = -tt.x;
    T.m[6] = -tt.y; T.m[7] = tt.x;  T.m[8] = 0.0;
    const M3 E = matmul(T, R);
    return matmul(matmul(inverse(transpose(tg.K)), E), inverse(s.K));
}

void visual_neighbors_from_worldpoints(const std::vector<const Camera*>& cams, const std::vector<float>& median_depth,
                                       const std::vector<std::vector<uint32_t>>& wps, unsigned num_neighbors,
                                       std::vector<std::vector<uint32_t>>& out)
{
    const uint32_t V = (uint32_t)cams.size();
    out.assign(V, std::vector<uint32_t>());
    // worldpoints2views_ as a flat inverted index: the observations (world point, view) grouped by
    // world point; a view's walk over its own list and the groups of its points counts what the
    // reference's two nested list walks count (src/line3D.cc:733-752), without a hash map.
    size_t n_obs = 0;
    uint32_t max_wp = 0;
    for (uint32_t v = 0; v < V; ++v) {
        n_obs += wps[v].size();
        for (uint32_t wp : wps[v]) max_wp = std::max(max_wp, wp);
    }
    if (!n_obs) return;
    std::vector<uint32_t> by_wp;         // views, grouped by world point (counting walk only)
    std::vector<uint32_t> start;         // group g = by_wp[start[g] .. start[g+1])
    std::vector<uint32_t> group_wp;      // sparse ids only: the world point of every group, ascending
    const bool dense = (size_t)max_wp < 8 * n_obs + 4096;
    // small scenes with dense ids (a key-frame window): one bit per (view, world point); the shared
    // points of two views are a popcount over the AND of their bit rows.  Only valid when no view lists a
    // point twice (the counting walk below multiplies the multiplicities, like the reference does).
    const size_t words = ((size_t)max_wp + 64) / 64;
    bool use_bits = dense && (size_t)V * V * words < ((size_t)1 << 22);
    std::vector<uint64_t> bits;
    if (use_bits) {
        bits.assign((size_t)V * words, 0ull);
        for (uint32_t v = 0; v < V && use_bits; ++v)
            for (uint32_t wp : wps[v]) {
                uint64_t& w = bits[(size_t)v * words + (wp >> 6)];
                const uint64_t m = 1ull << (wp & 63);
                if (w & m) {
                    use_bits = false;  // duplicate observation
                    break;
                }
                w |= m;
            }
    }
    if (!use_bits) {
        by_wp.resize(n_obs);
        if (dense) {  // counting sort, group index = world point id
            start.assign((size_t)max_wp + 2, 0u);
            for (uint32_t v = 0; v < V; ++v)
                for (uint32_t wp : wps[v]) ++start[(size_t)wp + 1];
            for (size_t i = 1; i < start.size(); ++i) start[i] += start[i - 1];
            std::vector<uint32_t> fill(start.begin(), start.end() - 1);
            for (uint32_t v = 0; v < V; ++v)
                for (uint32_t wp : wps[v]) by_wp[fill[wp]++] = v;
        } else {  // sort the (world point, view) keys
            std::vector<uint64_t> keys;
            keys.reserve(n_obs);
            for (uint32_t v = 0; v < V; ++v)
                for (uint32_t wp : wps[v]) keys.push_back(((uint64_t)wp << 32) | v);
            std::sort(keys.begin(), keys.end());
            start.push_back(0u);
            for (size_t i = 0; i < keys.size(); ++i) {
                by_wp[i] = (uint32_t)keys[i];
                if (i + 1 == keys.size() || (keys[i + 1] >> 32) != (keys[i] >> 32)) {
                    start.push_back((uint32_t)i + 1);
                    group_wp.push_back((uint32_t)(keys[i] >> 32));
                }
            }
        }
    }
    struct VN {
        uint32_t view;
        float score, axis_angle, dist_score;
    };
    std::vector<uint32_t> common(V);
    for (uint32_t v = 0; v < V; ++v) {
        std::fill(common.begin(), common.end(), 0u);
        bool any = false;
        if (use_bits) {
            const uint64_t* bv = &bits[(size_t)v * words];
            for (uint32_t o = 0; o < V; ++o) {
                if (o == v) continue;
                const uint64_t* bo = &bits[(size_t)o * words];
                uint32_t c = 0;
                for (size_t w = 0; w < words; ++w) c += (uint32_t)__builtin_popcountll(bv[w] & bo[w]);
                common[o] = c;
                any |= c != 0;
            }
        } else {
            for (uint32_t wp : wps[v]) {
                const size_t g = dense ? (size_t)wp
                                       : (size_t)(std::lower_bound(group_wp.begin(), group_wp.end(), wp) - group_wp.begin());
                for (uint32_t i = start[g]; i < start[g + 1]; ++i)
                    if (by_wp[i] != v) {
                        ++common[by_wp[i]];
                        any = true;
                    }
            }
        }
        if (!any) continue;
        const Camera& cv = *cams[v];
        const V3 ray_v = normalized(mul(cv.RtKinv, cv.pp));
        std::vector<VN> nb;  // candidates in ascending view order (std::map iteration order)
        for (uint32_t o = 0; o < V; ++o) {
            if (!common[o]) continue;
            const Camera& co = *cams[o];
            VN vn;
            vn.view = o;
            vn.score = 2.0f * float(common[o]) / float(wps[v].size() + wps[o].size());
            // View::opticalAxesAngle (src/view.cc:486-492)
            vn.axis_angle = (float)det_acos(std::fmin(std::fmax(dot(ray_v, normalized(mul(co.RtKinv, co.pp))), -1.0), 1.0));
            // View::distanceVisualNeighborScore (src/view.cc:516-530): |x| + |y| of the other centre in this camera
            const V3 c = mul(cv.R, co.C) + cv.t;
            const float d1 = (float)std::fabs(1.0 * c.x + 0.0 * c.y + 0.0 * c.z);
            const float d2 = (float)std::fabs(0.0 * c.x + 1.0 * c.y + 0.0 * c.z);
            vn.dist_score = d1 + d2;
            if (vn.axis_angle < 1.571f && common[o] > 4) nb.push_back(vn);
        }
        // std::list::sort is a stable merge sort
        std::stable_sort(nb.begin(), nb.end(), [](const VN& a, const VN& b) { return a.score > b.score; });
        if (nb.size() > num_neighbors) {
            const std::vector<VN> tmp = nb;
            const float score_t = 0.80f * nb.front().score;
            size_t bigger = 0;
            while (bigger < nb.size() && nb[bigger].score > score_t) ++bigger;
            nb.resize(bigger);
            std::stable_sort(nb.begin(), nb.end(), [](const VN& a, const VN& b) { return a.dist_score > b.dist_score; });
            if (nb.size() > num_neighbors / 2) nb.resize(num_neighbors / 2);
            nb.insert(nb.end(), tmp.begin(), tmp.end());
        }
        auto baseline = [&](uint32_t o) {
            const V3 d = cv.C - cams[o]->C;
            return (float)std::sqrt(dot(d, d));
        };
        const float min_baseline = cv.spatial_regularizer(0.5f) * median_depth[v];
        std::vector<uint32_t> used;  // kept ascending (std::set)
        for (size_t i = 0; i < nb.size() && used.size() < num_neighbors; ++i) {
            const uint32_t o = nb[i].view;
            if (std::binary_search(used.begin(), used.end(), o) || !(baseline(o) > min_baseline)) continue;
            bool valid = true;
            for (uint32_t u : used)
                if (!(baseline(u) > min_baseline)) {
                    valid = false;
                    break;
                }
            if (valid) used.insert(std::lower_bound(used.begin(), used.end(), o), o);
        }
        out[v] = used;
    }
}

}  // namespace hg
}  // namespace l3d
