// detmath.cuh -- deterministic expf / acos / acosf for the exact (decision-path) kernels.
//
// Same evaluation sequence as the CPU oracle defines for the reference's libm calls
// (expf: src/line3D.cc:1702,1709,1759,1808; acos: src/line3D.cc:1845, src/view.cc:509), written
// here with explicit round-to-nearest intrinsics so that no FMA can be formed whatever the
// compile flags are.  Only IEEE +,-,*,/ and sqrt in double: bit-identical on host and device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l3d {

// per-lane index: keep it in global memory (L1-cached), not the constant bank
static __device__ const double c_T32[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

static __constant__ double c_ASIN[30] = {
    0x1.0000000000000p+0,  0x1.5555555555555p-3,  0x1.3333333333333p-4,  0x1.6db6db6db6db7p-5,
    0x1.f1c71c71c71c7p-6,  0x1.6e8ba2e8ba2e9p-6,  0x1.1c4ec4ec4ec4fp-6,  0x1.c99999999999ap-7,
    0x1.7a87878787878p-7,  0x1.3fde50d79435ep-7,  0x1.12ef3cf3cf3cfp-7,  0x1.df3bd37a6f4dfp-8,
    0x1.a6863d70a3d71p-8,  0x1.782dda12f684cp-8,  0x1.51ba308d3dcb1p-8,  0x1.31683bdef7bdfp-8,
    0x1.15ee9d45d1746p-8,  0x1.fcaf8fb6db6dbp-9,  0x1.d3d2a8e0dd67dp-9,  0x1.b026f57b13b14p-9,
    0x1.90cb77f60c7cep-9,  0x1.750de64d7d05fp-9,  0x1.5c5f56efaaaabp-9,  0x1.464c0950f7d47p-9,
    0x1.3275586c5f2f0p-9,  0x1.208d3570ae5a6p-9,  0x1.1052bc5fa960ap-9,  0x1.018f963c229bfp-9,
    0x1.e82be60d9127ep-10, 0x1.cf7dea5b6e830p-10};

#define L3D_PI 0x1.921fb54442d18p+1
#define L3D_PI_2 0x1.921fb54442d18p+0

// exp(x), float in / float out, evaluated in double: x = (32 e + j) ln2/32 + r
__device__ __forceinline__ float det_expf(float x)
{
    if (x != x) return x;
    if (x > 88.8f) return __int_as_float(0x7f800000);
    if (x < -150.0f) return 0.0f;
    const double xd = (double)x;
    const double kd = rint(__dmul_rn(xd, 0x1.71547652b82fep+5));
    const double r = __dsub_rn(__dsub_rn(xd, __dmul_rn(kd, 0x1.62e42fee00000p-6)),
                               __dmul_rn(kd, 0x1.a39ef35793c76p-38));
    const int k = (int)kd;
    const int j = k & 31;
    const int e = k >> 5;
    double p = 1.0 / 120.0;
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 24.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 6.0);
    p = __dadd_rn(__dmul_rn(p, r), 0.5);
    p = __dadd_rn(__dmul_rn(p, r), 1.0);
    p = __dmul_rn(p, r);
    const double s = c_T32[j];
    const double y = __dadd_rn(s, __dmul_rn(s, p));
    const double scale = __longlong_as_double((long long)(e + 1023) << 52);
    return (float)__dmul_rn(y, scale);
}

__device__ __forceinline__ double det_asin_series(double z)
{
    double s = c_ASIN[29];
#pragma unroll
    for (int n = 28; n >= 0; --n) s = __dadd_rn(__dmul_rn(s, z), c_ASIN[n]);
    return s;
}

// acos(x), x in [-1,1]
__device__ __forceinline__ double det_acos(double x)
{
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return __longlong_as_double(0x7ff8000000000000LL);
    if (ax <= 0.5) {
        const double z = __dmul_rn(x, x);
        return __dsub_rn(L3D_PI_2, __dmul_rn(x, det_asin_series(z)));
    }
    const double z = __dmul_rn(__dsub_rn(1.0, ax), 0.5);
    const double r = __dsqrt_rn(z);
    const double a = __dmul_rn(2.0, __dmul_rn(r, det_asin_series(z)));
    return (x > 0.0) ? a : __dsub_rn(L3D_PI, a);
}

// acosf: float in / float out; 16-term asin series, Estrin evaluation (short dependency chain)
__device__ __forceinline__ double det_asin_series16(double z)
{
    const double z2 = __dmul_rn(z, z), z4 = __dmul_rn(z2, z2), z8 = __dmul_rn(z4, z4);
#define L3D_PAIR(a, b) __dadd_rn(c_ASIN[a], __dmul_rn(c_ASIN[b], z))
    const double p0 = L3D_PAIR(0, 1), p1 = L3D_PAIR(2, 3), p2 = L3D_PAIR(4, 5), p3 = L3D_PAIR(6, 7);
    const double p4 = L3D_PAIR(8, 9), p5 = L3D_PAIR(10, 11), p6 = L3D_PAIR(12, 13), p7 = L3D_PAIR(14, 15);
#undef L3D_PAIR
    const double q0 = __dadd_rn(p0, __dmul_rn(p1, z2)), q1 = __dadd_rn(p2, __dmul_rn(p3, z2));
    const double q2 = __dadd_rn(p4, __dmul_rn(p5, z2)), q3 = __dadd_rn(p6, __dmul_rn(p7, z2));
    const double r0 = __dadd_rn(q0, __dmul_rn(q1, z4)), r1 = __dadd_rn(q2, __dmul_rn(q3, z4));
    return __dadd_rn(r0, __dmul_rn(r1, z8));
}
__device__ __forceinline__ float det_acosf(float xf)
{
    const double x = (double)xf;
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return __int_as_float(0x7fc00000);
    if (ax <= 0.5) return (float)__dsub_rn(L3D_PI_2, __dmul_rn(x, det_asin_series16(__dmul_rn(x, x))));
    const double z = __dmul_rn(__dsub_rn(1.0, ax), 0.5);
    const double a = __dmul_rn(2.0, __dmul_rn(__dsqrt_rn(z), det_asin_series16(z)));
    return (float)((x > 0.0) ? a : __dsub_rn(L3D_PI, a));
}

}  // namespace l3d
