// exact.cuh -- double-precision 3-vector helpers for the decision-path ("exact") kernels.
// Every operation is an explicit round-to-nearest intrinsic, so the sequence is the canonical one
// of SURVEY.md Appendix A (left-to-right sums, true divisions, no FMA) whatever the compile flags.
#pragma once
#include <cuda_runtime.h>

namespace l3d {

struct D3 {
    double x, y, z;
};

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dd(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 sub3(const D3& a, const D3& b) { return D3{ds(a.x, b.x), ds(a.y, b.y), ds(a.z, b.z)}; }
__device__ __forceinline__ D3 add3(const D3& a, const D3& b) { return D3{da(a.x, b.x), da(a.y, b.y), da(a.z, b.z)}; }
__device__ __forceinline__ D3 scale3(const D3& a, double s) { return D3{dm(a.x, s), dm(a.y, s), dm(a.z, s)}; }
__device__ __forceinline__ double dot3(const D3& a, const D3& b)
{
    return da(da(dm(a.x, b.x), dm(a.y, b.y)), dm(a.z, b.z));
}
__device__ __forceinline__ D3 cross3(const D3& a, const D3& b)
{
    return D3{ds(dm(a.y, b.z), dm(a.z, b.y)), ds(dm(a.z, b.x), dm(a.x, b.z)), ds(dm(a.x, b.y), dm(a.y, b.x))};
}
__device__ __forceinline__ double norm3(const D3& a) { return __dsqrt_rn(dot3(a, a)); }
__device__ __forceinline__ D3 normalized3(const D3& a)
{
    const double n = norm3(a);
    return D3{dd(a.x, n), dd(a.y, n), dd(a.z, n)};
}
// row-major 3x3 * v
__device__ __forceinline__ D3 mul33(const double* __restrict__ M, const D3& v)
{
    return D3{da(da(dm(M[0], v.x), dm(M[1], v.y)), dm(M[2], v.z)),
              da(da(dm(M[3], v.x), dm(M[4], v.y)), dm(M[5], v.z)),
              da(da(dm(M[6], v.x), dm(M[7], v.y)), dm(M[8], v.z))};
}

// float helpers with explicit rounding (no FFMA contraction)
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fd(float a, float b) { return __fdiv_rn(a, b); }

}  // namespace l3d
