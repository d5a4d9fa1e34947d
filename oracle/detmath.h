/* oracle/detmath.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Deterministic expf / acos / acosf / sin built from IEEE-754 double +,-,*,/ and sqrt only
 * (no FMA contraction: build with -ffp-contract=off).  The reference calls libm
 * (expf: line3D.cc:1702,1709,1759,1808-1809; acos: line3D.cc:1845, view.cc:341,509;
 * sin: view.cc:342); glibc's last-ulp behaviour is CPU-dependent (ifunc FMA variants) and
 * cannot be reproduced on a GPU, so the oracle DEFINES these functions (SURVEY.md section 7,
 * "Transcendentals on decision paths") and the CUDA side implements the same sequence
 * independently (3dline-slam_b200/csrc/detmath.cuh).  tests/test_detmath.py pins them
 * against libm to <= 1 ulp.
 */
#ifndef L3D_ORACLE_DETMATH_H_
#define L3D_ORACLE_DETMATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

static const double ORC_T32[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

/* asin(x)/x = sum_n c_n x^(2n),  c_n = (2n)! / (4^n (n!)^2 (2n+1)) */
static const double ORC_ASIN_C[30] = {
    0x1.0000000000000p+0,  0x1.5555555555555p-3,  0x1.3333333333333p-4,  0x1.6db6db6db6db7p-5,
    0x1.f1c71c71c71c7p-6,  0x1.6e8ba2e8ba2e9p-6,  0x1.1c4ec4ec4ec4fp-6,  0x1.c99999999999ap-7,
    0x1.7a87878787878p-7,  0x1.3fde50d79435ep-7,  0x1.12ef3cf3cf3cfp-7,  0x1.df3bd37a6f4dfp-8,
    0x1.a6863d70a3d71p-8,  0x1.782dda12f684cp-8,  0x1.51ba308d3dcb1p-8,  0x1.31683bdef7bdfp-8,
    0x1.15ee9d45d1746p-8,  0x1.fcaf8fb6db6dbp-9,  0x1.d3d2a8e0dd67dp-9,  0x1.b026f57b13b14p-9,
    0x1.90cb77f60c7cep-9,  0x1.750de64d7d05fp-9,  0x1.5c5f56efaaaabp-9,  0x1.464c0950f7d47p-9,
    0x1.3275586c5f2f0p-9,  0x1.208d3570ae5a6p-9,  0x1.1052bc5fa960ap-9,  0x1.018f963c229bfp-9,
    0x1.e82be60d9127ep-10, 0x1.cf7dea5b6e830p-10};

static const double ORC_SIN_C[12] = {
    0x1.0000000000000p+0,   -0x1.5555555555555p-3,  0x1.1111111111111p-7,  -0x1.a01a01a01a01ap-13,
    0x1.71de3a556c734p-19,  -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41,
    0x1.952c77030ad4ap-49,  -0x1.2f49b46814157p-57, 0x1.71b8ef6dcf572p-66, -0x1.761b41316381ap-75};
static const double ORC_COS_C[12] = {
    0x1.0000000000000p+0,   -0x1.0000000000000p-1,  0x1.5555555555555p-5,  -0x1.6c16c16c16c17p-10,
    0x1.a01a01a01a01ap-16,  -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37,
    0x1.ae7f3e733b81fp-45,  -0x1.6827863b97d97p-53, 0x1.e542ba4020225p-62, -0x1.0ce396db7f853p-70};

#define ORC_PI 0x1.921fb54442d18p+1
#define ORC_PI_2 0x1.921fb54442d18p+0
#define ORC_PI_4 0x1.921fb54442d18p-1

/* exp(x) for float x, evaluated in double:  x = (32 e + j) ln2/32 + r,  exp = 2^e 2^(j/32) e^r */
static inline float orc_expf(float x)
{
    if (x != x) return x;
    if (x > 88.8f) return INFINITY;
    if (x < -150.0f) return 0.0f;
    const double xd = (double)x;
    const double kd = rint(xd * 0x1.71547652b82fep+5);
    const double r = (xd - kd * 0x1.62e42fee00000p-6) - kd * 0x1.a39ef35793c76p-38;
    const int k = (int)kd;
    const int j = k & 31;
    const int e = k >> 5; /* arithmetic shift: floor division */
    /* e^r - 1 ~ r + r^2/2 + r^3/6 + r^4/24 + r^5/120,  |r| <= ln2/64 */
    double p = 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r;
    const double s = ORC_T32[j];
    const double y = s + s * p;
    uint64_t bits = (uint64_t)(e + 1023) << 52;
    double scale;
    memcpy(&scale, &bits, 8);
    return (float)(y * scale);
}

static inline double orc_asin_series(double z)
{
    double s = ORC_ASIN_C[29];
    for (int n = 28; n >= 0; --n) s = s * z + ORC_ASIN_C[n];
    return s;
}

/* acos(x), x in [-1,1] (callers clamp, as the reference does) */
static inline double orc_acos(double x)
{
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return NAN;
    if (ax <= 0.5) {
        const double z = x * x;
        return ORC_PI_2 - x * orc_asin_series(z);
    }
    const double z = (1.0 - ax) * 0.5;
    const double r = sqrt(z);
    const double a = 2.0 * (r * orc_asin_series(z));
    return (x > 0.0) ? a : (ORC_PI - a);
}

/* acosf: float in / float out.  Same range reduction as orc_acos, but the asin series is cut at
 * 16 terms (relative error < 5e-12, far below a float ulp) and evaluated with Estrin's scheme:
 * pairs (c0+c1 z), ..., then z^2, z^4, z^8 combinations -- a short dependency chain for the GPU. */
static inline double orc_asin_series16(double z)
{
    const double* c = ORC_ASIN_C;
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double p0 = c[0] + c[1] * z, p1 = c[2] + c[3] * z, p2 = c[4] + c[5] * z, p3 = c[6] + c[7] * z;
    const double p4 = c[8] + c[9] * z, p5 = c[10] + c[11] * z, p6 = c[12] + c[13] * z, p7 = c[14] + c[15] * z;
    const double q0 = p0 + p1 * z2, q1 = p2 + p3 * z2, q2 = p4 + p5 * z2, q3 = p6 + p7 * z2;
    const double r0 = q0 + q1 * z4, r1 = q2 + q3 * z4;
    return r0 + r1 * z8;
}
static inline float orc_acosf(float xf)
{
    const double x = (double)xf;
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return NAN;
    if (ax <= 0.5) return (float)(ORC_PI_2 - x * orc_asin_series16(x * x));
    const double z = (1.0 - ax) * 0.5;
    const double a = 2.0 * (sqrt(z) * orc_asin_series16(z));
    return (float)((x > 0.0) ? a : (ORC_PI - a));
}

/* sin(x) for x in [0, pi] (only use: View::getSpecificSpatialReg, view.cc:341-342) */
static inline double orc_sin(double x)
{
    double xr = (x > ORC_PI_2) ? (ORC_PI - x) : x;
    if (xr <= ORC_PI_4) {
        const double z = xr * xr;
        double s = ORC_SIN_C[11];
        for (int n = 10; n >= 0; --n) s = s * z + ORC_SIN_C[n];
        return xr * s;
    }
    const double y = ORC_PI_2 - xr;
    const double z = y * y;
    double s = ORC_COS_C[11];
    for (int n = 10; n >= 0; --n) s = s * z + ORC_COS_C[n];
    return s;
}

#endif /* L3D_ORACLE_DETMATH_H_ */
