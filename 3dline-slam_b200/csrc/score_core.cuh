// score_core.cuh -- Line3D::similarityForScoring (src/line3D.cc:1685-1716) with
// Line3D::angleBetweenSeg3D (src/line3D.cc:1841-1853) in the canonical float/double sequence,
// shared by the data-flow scoring kernels and the l3d_score_matches kernel (exact TUs only).
#pragma once
#include "detmath.cuh"
#include <math.h>

#include "exact.cuh"

namespace l3d {

// compact per-entry data the scoring loop reads for every sibling (16 B, one LDS/LDG.128)
struct __align__(16) Sib {
    float d_p1, d_p2;
    uint32_t cam;    // target view of the entry
    uint32_t flags;  // bit0: starts a new camera run, bit1: 3-D segment valid (length >= eps)
};

// dotcut: cos of (sqrt(0.70 two_sigA_sqr) + 0.01) degrees when min_sim >= 0.5, else -1 (disabled).
// returns sim (0 if truncated).  xcut/pcut implement the exact early exits (see k3 kernels):
// with min_sim >= 0.5, exp(x) <= 0.4966 < min_sim for x < -0.70, so the result is 0 either way.
__device__ __forceinline__ float sim_for_scoring(float Md1, float Md2, float reg1, float reg2, bool Mvalid,
                                                 const D3& dirM, const Sib& s2, const double* __restrict__ dir2,
                                                 float two_sigA_sqr, float min_sim, float xcut, float pcut,
                                                 float dotcut)
{
    if (!Mvalid || !(s2.flags & 2u)) return 0.0f;
    const float d1 = fs(Md1, s2.d_p1);
    const float d2 = fs(Md2, s2.d_p2);
    const float n1 = fm(-d1, d1), n2 = fm(-d2, d2);
    // cheap certain reject: -d^2 < -0.75 reg  =>  -d^2/reg < -0.70 (reg > 0), no division needed
    if (xcut > -1.0f && (n1 < fm(-0.75f, reg1) || n2 < fm(-0.75f, reg2)) && reg1 > 0.0f && reg2 > 0.0f) return 0.0f;
    const float x1 = fd(n1, reg1);
    const float x2 = fd(n2, reg2);
    if (x1 < xcut || x2 < xcut) return 0.0f;
    // the direction test before the two exponentials of the position term: both orders return 0 for the same
    // pairs (each early exit is certain on its own), and the dot product is the cheaper one
    const float dot_p = (float)dot3(dirM, d3(dir2[0], dir2[1], dir2[2]));
    // |dot| < cos(angle_cut + 0.01 deg) => angle > angle_cut => -angle^2/two_sigA_sqr < -0.70 => sim_a < 0.4966
    if (fabsf(dot_p) < dotcut) return 0.0f;
    const float sim_p = fminf(det_expf(x1), det_expf(x2));
    if (sim_p <= pcut) return 0.0f;
    float angle = (float)dm(dd((double)det_acosf(fmaxf(fminf(dot_p, 1.0f), -1.0f)), L3D_PI), (double)180.0f);
    if (angle > 90.0f) angle = fs(180.0f, angle);
    const float sim_a = det_expf(fd(fm(-angle, angle), two_sigA_sqr));
    const float s = fminf(sim_a, sim_p);
    return (s > min_sim) ? s : 0.0f;
}

// host: conservative |dot| threshold below which the angular similarity is certainly < 0.4966
static inline float score_dotcut(float two_sigA_sqr, float min_sim)
{
    if (!(min_sim >= 0.5f)) return -1.0f;
    const double cut_deg = sqrt(0.70 * (double)two_sigA_sqr) + 0.01;
    if (cut_deg >= 89.9) return -1.0f;
    return (float)cos(cut_deg * 3.14159265358979323846 / 180.0);
}

__device__ __forceinline__ uint32_t float_ordered(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

}  // namespace l3d
