// k5_collinear.cu -- the per-view collinearity table behind l3d_find_collinear, the drop-in for
// L3DPP::find_collinear_segments_GPU (include/cudawrapper.h:84-86): View::findCollinCPU
// (src/view.cc:238-293) with View::pointOnSegment (src/view.cc:321-327) and
// View::distance_point2line_2D (src/view.cc:296-299).  Exact TU: every operation is the
// reference's, in its order and width (double cross products and dot tests, sqrtf of the float-
// rounded squared norm, double division, float maxima and threshold test).
//
// An N x N byte table per view, HBM-write bound (1 B per pair, 16 B of segment per row/column
// element amortised): one thread per pair, the column index fastest so that a warp stores 32
// consecutive bytes; the row's segment is a broadcast load.
#include "exact.cuh"
#include "internal.h"

namespace l3d {

#define L3D_EPS 1e-12

__device__ __forceinline__ bool on_seg2d(double p1x, double p1y, double p2x, double p2y, double xx, double xy)
{
    const double v1x = ds(p1x, xx), v1y = ds(p1y, xy), v2x = ds(p2x, xx), v2y = ds(p2y, xy);
    return da(dm(v1x, v2x), dm(v1y, v2y)) < L3D_EPS;
}

__device__ __forceinline__ float dist_p2l(const D3& l, double px, double py)
{
    const double num = da(da(dm(l.x, px), dm(l.y, py)), l.z);
    const float den = __fsqrt_rn((float)da(dm(l.x, l.x), dm(l.y, l.y)));
    return (float)fabs(dd(num, (double)den));
}

__global__ void __launch_bounds__(256) k5_collinear_kernel(const float4* __restrict__ lines, uint32_t n, float dist_t,
                                                           char* __restrict__ out, size_t row_stride)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = blockIdx.y;
    if (c >= n) return;
    char res = 0;
    if (r != c) {
        const float4 a = lines[r], b = lines[c];
        const D3 p0 = d3((double)a.x, (double)a.y, 1.0), p1 = d3((double)a.z, (double)a.w, 1.0);
        const D3 q0 = d3((double)b.x, (double)b.y, 1.0), q1 = d3((double)b.z, (double)b.w, 1.0);
        if (!(on_seg2d(p0.x, p0.y, p1.x, p1.y, q0.x, q0.y) || on_seg2d(p0.x, p0.y, p1.x, p1.y, q1.x, q1.y) ||
              on_seg2d(q0.x, q0.y, q1.x, q1.y, p0.x, p0.y) || on_seg2d(q0.x, q0.y, q1.x, q1.y, p1.x, p1.y))) {
            const D3 line1 = cross3(p0, p1), line2 = cross3(q0, q1);
            const float d1 = fmaxf(dist_p2l(line1, q0.x, q0.y), dist_p2l(line1, q1.x, q1.y));
            const float d2 = fmaxf(dist_p2l(line2, p0.x, p0.y), dist_p2l(line2, p1.x, p1.y));
            res = fmaxf(d1, d2) < dist_t ? 1 : 0;
        }
    }
    out[(size_t)r * row_stride + c] = res;
}

int launch_k5_collinear(const float4* lines, uint32_t n, float dist_t, char* out, size_t row_stride, cudaStream_t st)
{
    if (!n) return 0;
    dim3 grid((n + 255) / 256, n);
    k5_collinear_kernel<<<grid, 256, 0, st>>>(lines, n, dist_t, out, row_stride);
    return 1;
}

}  // namespace l3d
