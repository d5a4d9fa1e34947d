// k0_prep.cu -- per-segment tables computed once per scene (exact TU: built with -fmad=false).
//
//  * SegDesc: the target-side descriptor of the K1 pair test.  All four points the reference
//    feeds to Line3D::mutualOverlap (src/line3D.cc:1150-1156) lie on the target segment's line,
//    so the test is one-dimensional in the arc-length parameter s along the unit direction u from
//    q1 (q1 -> 0, q2 -> L); the image-bounds test (src/line3D.cc:1142-1148) becomes an interval
//    [slo, shi] of s.  Computed in double, rounded outwards to float.
//  * SegRays / midray: View::getNormalizedRay (src/view.cc:346-350) of both endpoints and of the
//    2-D midpoint (View::segmentQualityAngle, src/view.cc:500-506), in the canonical double sequence.
#include "exact.cuh"
#include "internal.h"

namespace l3d {

__global__ void __launch_bounds__(256, 5) k0_prep_kernel(const float4* __restrict__ segs,
                                                      const uint32_t* __restrict__ seg_view,
                                                      const ViewDev* __restrict__ views, uint32_t S, double W,
                                                      SegDesc* __restrict__ desc, SegRays* __restrict__ rays,
                                                      double* __restrict__ midray, SegPlane* __restrict__ planes,
                                                      SegV32* __restrict__ v32, float* __restrict__ view_xb)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const uint32_t v = seg_view[i];
    if (!views[v].needed) return;  // multi-GPU: a view none of this rank's pairs touches
    const float4 sg = segs[i];
    const double x1 = (double)sg.x, y1 = (double)sg.y, x2 = (double)sg.z, y2 = (double)sg.w;

    // ---- rays (exact) ----
    const double* M = views[v].RtKinv;
    const D3 r1 = normalized3(mul33(M, d3(x1, y1, 1.0)));
    const D3 r2 = normalized3(mul33(M, d3(x2, y2, 1.0)));
    SegRays rr;
    rr.r1[0] = r1.x; rr.r1[1] = r1.y; rr.r1[2] = r1.z;
    rr.r2[0] = r2.x; rr.r2[1] = r2.y; rr.r2[2] = r2.z;
    rays[i] = rr;
    // Eigen: p = 0.5*(p1+p2) on Vector2d
    const D3 rm = normalized3(mul33(M, d3(dm(0.5, da(x1, x2)), dm(0.5, da(y1, y2)), 1.0)));
    midray[3 * (size_t)i + 0] = rm.x;
    midray[3 * (size_t)i + 1] = rm.y;
    midray[3 * (size_t)i + 2] = rm.z;
    // n = (r1 x r2).normalized() and n.C, the per-segment part of Line3D::triangulationDepths
    const D3 n = normalized3(cross3(r1, r2));
    const double* Cv = views[v].C;
    SegPlane pl;
    pl.n[0] = n.x; pl.n[1] = n.y; pl.n[2] = n.z;
    pl.cn = dot3(d3(Cv[0], Cv[1], Cv[2]), n);
    planes[i] = pl;
    // FP32 image of the same tables for K2's certified depth-sign test (not on the decision path)
    SegV32 f;
    f.nx = (float)n.x; f.ny = (float)n.y; f.nz = (float)n.z; f.cn = (float)pl.cn;
    f.r1x = (float)r1.x; f.r1y = (float)r1.y; f.r1z = (float)r1.z;
    f.r2x = (float)r2.x; f.r2y = (float)r2.y; f.r2z = (float)r2.z;
    f.pad0 = f.pad1 = 0.0f;
    v32[i] = f;

    // ---- K1 descriptor (conservative, not on the decision path) ----
    const double dx = x2 - x1, dy = y2 - y1;
    const double L = sqrt(dx * dx + dy * dy);
    SegDesc d;
    d.q1x = sg.x;
    d.q1y = sg.y;
    if (L > 0.0) {
        const double ux = dx / L, uy = dy / L;
        double lo = -1e300, hi = 1e300;
        if (ux > 0.0) { lo = fmax(lo, (0.0 - x1) / ux); hi = fmin(hi, (W - x1) / ux); }
        else if (ux < 0.0) { lo = fmax(lo, (W - x1) / ux); hi = fmin(hi, (0.0 - x1) / ux); }
        else if (x1 < 0.0 || x1 > W) { lo = 1e300; hi = -1e300; }
        if (uy > 0.0) { lo = fmax(lo, (0.0 - y1) / uy); hi = fmin(hi, (W - y1) / uy); }
        else if (uy < 0.0) { lo = fmax(lo, (W - y1) / uy); hi = fmin(hi, (0.0 - y1) / uy); }
        else if (y1 < 0.0 || y1 > W) { lo = 1e300; hi = -1e300; }
        d.ux = (float)ux;
        d.uy = (float)uy;
        d.L = (float)L;
        d.slo = __double2float_rd(lo);
        d.shi = __double2float_ru(hi);
        // guard for the float roundings of L, slo, shi (a few ulps of their magnitude)
        const double mag = L + fmin(fabs(lo), 1e9) + fmin(fabs(hi), 1e9);
        d.g = __double2float_ru(mag * 4.0 * 5.9604644775390625e-08);
    } else {
        // zero-length segment: l2 = 0 in the reference, never a match; K1 marks it a candidate
        // (D == 0 path) and the exact kernel rejects it.
        d.ux = 0.0f; d.uy = 0.0f; d.L = 0.0f; d.slo = -3.0e38f; d.shi = 3.0e38f; d.g = 0.0f;
    }
    desc[i] = d;
    atomicMax((int*)&view_xb[v], __float_as_int(fabsf(sg.x) + fabsf(sg.y) + 1.0f));
}

int launch_k0_prep(const float4* segs, const uint32_t* seg_view, const ViewDev* views, uint32_t S,
                   int max_image_width, SegDesc* desc, SegRays* rays, double* midray, SegPlane* planes,
                   SegV32* v32, float* view_xb, cudaStream_t st)
{
    if (S == 0) return 0;
    k0_prep_kernel<<<(S + 255) / 256, 256, 0, st>>>(segs, seg_view, views, S, (double)max_image_width, desc,
                                                     rays, midray, planes, v32, view_xb);
    return 1;
}

}  // namespace l3d
