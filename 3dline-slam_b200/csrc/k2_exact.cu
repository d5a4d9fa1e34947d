// k2_exact.cu -- K2: exact re-evaluation of the K1 candidates in the reference's double sequence,
// two-view triangulation, kNN selection and the orientation filter (exact TU, -fmad=false; every
// arithmetic step is also an explicit _rn intrinsic).
//
// Three launches per batch:
//   expand : the row's candidate bits -> a list of target indices in ascending order, i.e. the order
//            Line3D::matchingCPU pushes matches (src/line3D.cc:1124-1196);
//   K2a    : ONE THREAD PER CANDIDATE (flat, no divergent per-row loops): repeats the pair test
//            exactly (src/line3D.cc:1131-1158, mutualOverlap :1283-1362), triangulates both
//            directions (Line3D::triangulationDepths, src/line3D.cc:1365-1390) and evaluates the
//            orientation test of the would-be match (Line3D::checkMatchOrientation
//            src/line3D.cc:962-1014, View::segmentQualityAngle src/view.cc:495-513);
//   K2b    : one thread per row: pushes the valid matches on a binary max-heap keyed by overlap
//            (std::priority_queue<Match, vector, Match_kNN>, include/commons.h:233-244; the sift-up /
//            sift-down steps follow the textbook algorithm libstdc++ uses, so equal overlaps pop in
//            the same order), pops min(kNN, n) (src/line3D.cc:1198-1206), drops the ones that fail
//            the orientation test and leaves the survivors, in list order, in fin_rec.
#include "detmath.cuh"
#include "exact.cuh"
#include "internal.h"

namespace l3d {

static constexpr int K2_ROWS = 256;
#define L3D_EPS 1e-12

__device__ __forceinline__ bool point_on_segment(const D3& x, const D3& p1, const D3& p2)
{
    const double v1x = ds(p1.x, x.x), v1y = ds(p1.y, x.y);
    const double v2x = ds(p2.x, x.x), v2y = ds(p2.y, x.y);
    return da(dm(v1x, v2x), dm(v1y, v2y)) < L3D_EPS;
}

// Line3D::mutualOverlap, src/line3D.cc:1283-1362
__device__ __forceinline__ float mutual_overlap(const D3* pt)
{
    if (!(point_on_segment(pt[0], pt[2], pt[3]) || point_on_segment(pt[1], pt[2], pt[3]) ||
          point_on_segment(pt[2], pt[0], pt[1]) || point_on_segment(pt[3], pt[0], pt[1])))
        return 0.0f;
    float max_dist = 0.0f;
    int o1 = 0, o2 = 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            const float d = (float)norm3(sub3(pt[i], pt[j]));
            if (d > max_dist) {
                max_dist = d;
                o1 = i;
                o2 = j;
            }
        }
    if (max_dist < 1.0f) return 0.0f;
    int i1, i2;
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
    // select without dynamic indexing of the register array
    D3 a = pt[0], b = pt[1];
    if (i1 == 1) a = pt[1];
    if (i1 == 2) a = pt[2];
    if (i2 == 1) b = pt[1];
    if (i2 == 2) b = pt[2];
    if (i2 == 3) b = pt[3];
    return (float)dd(norm3(sub3(a, b)), (double)max_dist);
}

__device__ __forceinline__ float key_overlap(unsigned long long k) { return __uint_as_float((uint32_t)(k >> 32)); }

// std::push_heap step: value already stored at index n-1 conceptually; sift it up
__device__ __forceinline__ void heap_push(unsigned long long* h, uint32_t n_before, unsigned long long value)
{
    uint32_t hole = n_before;
    const float v = key_overlap(value);
    while (hole > 0) {
        const uint32_t parent = (hole - 1) >> 1;
        const unsigned long long pk = h[parent];
        if (!(key_overlap(pk) < v)) break;
        h[hole] = pk;
        hole = parent;
    }
    h[hole] = value;
}

// std::pop_heap on [0,n): afterwards the heap is [0,n-1)
__device__ __forceinline__ void heap_pop(unsigned long long* h, uint32_t n)
{
    if (n <= 1) return;
    const unsigned long long value = h[n - 1];
    const uint32_t len = n - 1;
    uint32_t hole = 0, second = 0;
    while ((int)second < ((int)len - 1) / 2) {
        second = 2 * (second + 1);
        if (key_overlap(h[second]) < key_overlap(h[second - 1])) --second;
        h[hole] = h[second];
        hole = second;
    }
    if ((len & 1u) == 0 && (int)second == ((int)len - 2) / 2) {
        second = 2 * (second + 1);
        h[hole] = h[second - 1];
        hole = second - 1;
    }
    // __push_heap(first, hole, top=0, value)
    const float v = key_overlap(value);
    while (hole > 0) {
        const uint32_t parent = (hole - 1) >> 1;
        const unsigned long long pk = h[parent];
        if (!(key_overlap(pk) < v)) break;
        h[hole] = pk;
        hole = parent;
    }
    h[hole] = value;
}

__device__ __forceinline__ D3 ld3(const double* p) { return D3{p[0], p[1], p[2]}; }

static constexpr uint32_t K2_INVALID = 0xffffffffu;

// ---- K2 expand: bit mask -> per-row candidate lists (ascending target index) ----
__global__ void __launch_bounds__(K2_ROWS) k2_expand_kernel(const PairDev* __restrict__ pairs,
                                                            const K1Cta* __restrict__ ctas,
                                                            const uint32_t* __restrict__ mask,
                                                            const uint32_t* __restrict__ cand_off,
                                                            uint32_t* __restrict__ cand_c,
                                                            uint32_t* __restrict__ cand_row,
                                                            uint32_t* __restrict__ row_pair)
{
    const K1Cta cta = ctas[blockIdx.x];
    const PairDev& P = pairs[cta.pair];
    const uint32_t r = cta.tile * K2_ROWS + threadIdx.x;
    if (r >= P.n_src) return;
    const uint32_t lrow = P.row_base - P.batch_row0 + r;
    row_pair[lrow] = cta.pair;
    uint32_t pos = cand_off[lrow];
    const uint32_t* __restrict__ mrow = mask + P.mask_base + r;
    for (uint32_t w = 0; w < P.words; ++w) {
        uint32_t m = mrow[(size_t)w * P.n_src];
        while (m) {
            const uint32_t j = __ffs(m) - 1;
            m &= m - 1;
            cand_c[pos] = w * 32 + j;
            cand_row[pos] = lrow;
            ++pos;
        }
    }
}

// ---- K2a: one thread per candidate: exact pair test, both triangulations, orientation test ----
__global__ void __launch_bounds__(128) k2a_exact_kernel(
    const PairDev* __restrict__ pairs, const uint32_t* __restrict__ row_pair, const uint32_t* __restrict__ cand_c,
    const uint32_t* __restrict__ cand_row, uint32_t n_cand, const float4* __restrict__ segs,
    const SegRays* __restrict__ rays, const double* __restrict__ midray, const ViewDev* __restrict__ views,
    FwdRec* __restrict__ cand_rec, float thr, double W)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cand) return;
    const uint32_t lrow = cand_row[t];
    const uint32_t c = cand_c[t];
    const PairDev& P = pairs[row_pair[lrow]];
    const uint32_t r = lrow - (P.row_base - P.batch_row0);
    FwdRec rec;
    rec.c = c;
    rec.flags = K2_INVALID;
    rec.overlap = 0.0f;
    rec.d_p1 = rec.d_p2 = rec.d_q1 = rec.d_q2 = rec.score = 0.0f;

    const float4 sg = segs[P.src_off + r];
    const D3 p1 = d3((double)sg.x, (double)sg.y, 1.0), p2 = d3((double)sg.z, (double)sg.w, 1.0);
    const D3 e1 = mul33(P.F, p1), e2 = mul33(P.F, p2);
    const float4 tg = segs[P.tgt_off + c];
    const D3 q1 = d3((double)tg.x, (double)tg.y, 1.0), q2 = d3((double)tg.z, (double)tg.w, 1.0);
    const D3 l2 = cross3(q1, q2);
    D3 a = cross3(l2, e1), b = cross3(l2, e2);
    bool ok = fabs(a.z) > L3D_EPS && fabs(b.z) > L3D_EPS;
    float score = 0.0f;
    if (ok) {
        a = d3(dd(a.x, a.z), dd(a.y, a.z), dd(a.z, a.z));
        b = d3(dd(b.x, b.z), dd(b.y, b.z), dd(b.z, b.z));
        ok = !(a.x < 0 || a.x > W || a.y < 0 || a.y > W || b.x < 0 || b.x > W || b.y < 0 || b.y > W);
    }
    if (ok) {
        const D3 pts[4] = {a, b, q1, q2};
        score = mutual_overlap(pts);
        ok = score > thr;
    }
    if (ok) {
        const ViewDev& vs = views[P.src_view];
        const ViewDev& vt = views[P.tgt_view];
        const SegRays sr = rays[P.src_off + r];
        const D3 rp1 = ld3(sr.r1), rp2 = ld3(sr.r2);
        const D3 Cs = ld3(vs.C), Ct = ld3(vt.C);
        const SegRays tr = rays[P.tgt_off + c];
        const D3 rq1 = ld3(tr.r1), rq2 = ld3(tr.r2);
        double ds1 = -1.0, ds2 = -1.0, dt1 = -1.0, dt2 = -1.0;
        {  // triangulationDepths(src,p | tgt,q)
            const D3 nA = normalized3(cross3(rq1, rq2));
            const double a1 = dot3(rp1, nA), a2 = dot3(rp2, nA);
            if (!(fabs(a1) < L3D_EPS || fabs(a2) < L3D_EPS)) {
                const double num = ds(dot3(Ct, nA), dot3(nA, Cs));
                ds1 = dd(num, dot3(nA, rp1));
                ds2 = dd(num, dot3(nA, rp2));
            }
        }
        {  // triangulationDepths(tgt,q | src,p)
            const D3 nB = normalized3(cross3(rp1, rp2));
            const double b1 = dot3(rq1, nB), b2 = dot3(rq2, nB);
            if (!(fabs(b1) < L3D_EPS || fabs(b2) < L3D_EPS)) {
                const double numB = ds(dot3(Cs, nB), dot3(nB, Ct));
                dt1 = dd(numB, dot3(nB, rq1));
                dt2 = dd(numB, dot3(nB, rq2));
            }
        }
        if (ds1 > L3D_EPS && ds2 > L3D_EPS && dt1 > L3D_EPS && dt2 > L3D_EPS) {
            rec.overlap = score;
            rec.d_p1 = (float)ds1;
            rec.d_p2 = (float)ds2;
            rec.d_q1 = (float)dt1;
            rec.d_q2 = (float)dt2;
            // orientation test of the would-be match (checkMatchOrientation, src/line3D.cc:962-1014):
            // flags = 0 keep, 1 = dropped by the orientation filter if it survives the kNN selection
            const D3 P1 = add3(Cs, scale3(rp1, (double)rec.d_p1));
            const D3 P2 = add3(Cs, scale3(rp2, (double)rec.d_p2));
            const float len = (float)norm3(sub3(P1, P2));
            D3 dir = d3(0.0, 0.0, 0.0);
            if (len > L3D_EPS) dir = normalized3(sub3(P2, P1));
            const D3 rmid = ld3(midray + 3 * (size_t)(P.src_off + r));
            const double ang = det_acos(fmin(fmax(dot3(rmid, dir), -1.0), 1.0));
            rec.flags = (ang > (double)0.098174771f && ang < (double)3.043417886f) ? 0u : 1u;
        }
    }
    cand_rec[t] = rec;
}

// ---- K2b: one thread per row: priority-queue order, kNN pops, orientation filter ----
__global__ void __launch_bounds__(K2_ROWS) k2b_select_kernel(const uint32_t* __restrict__ cand_off, uint32_t n_rows,
                                                             unsigned long long* __restrict__ heap,
                                                             const FwdRec* __restrict__ cand_rec,
                                                             FwdRec* __restrict__ fin_rec,
                                                             uint32_t* __restrict__ fin_cnt, int knn, int apply_orient)
{
    const uint32_t lrow = blockIdx.x * blockDim.x + threadIdx.x;
    if (lrow >= n_rows) return;
    const uint32_t base = cand_off[lrow];
    const uint32_t cap = cand_off[lrow + 1] - base;
    unsigned long long* __restrict__ hp = heap + base;
    const FwdRec* __restrict__ crec = cand_rec + base;
    FwdRec* __restrict__ frec = fin_rec + base;
    uint32_t nout = 0;
    if (knn > 0) {
        uint32_t n = 0;
        for (uint32_t i = 0; i < cap; ++i) {  // push order = ascending target index
            const FwdRec rc = crec[i];
            if (rc.flags == K2_INVALID) continue;
            heap_push(hp, n, ((unsigned long long)__float_as_uint(rc.overlap) << 32) | i);
            ++n;
        }
        const uint32_t npop = min((uint32_t)knn, n);
        uint32_t hn = n;
        for (uint32_t t = 0; t < npop; ++t) {
            const uint32_t idx = (uint32_t)(hp[0] & 0xffffffffu);
            heap_pop(hp, hn);
            --hn;
            FwdRec rc = crec[idx];
            if (rc.flags == 0u || !apply_orient) {
                rc.flags = 0u;
                frec[nout++] = rc;
            }
        }
    } else {
        for (uint32_t i = 0; i < cap; ++i) {
            FwdRec rc = crec[i];
            if (rc.flags == 0u || (!apply_orient && rc.flags != K2_INVALID)) {
                rc.flags = 0u;
                frec[nout++] = rc;
            }
        }
    }
    fin_cnt[lrow] = nout;
}

__global__ void __launch_bounds__(256) k2_compact_kernel(const uint32_t* __restrict__ cand_off,
                                                         const uint32_t* __restrict__ fin_cnt,
                                                         const uint32_t* __restrict__ fin_off, uint32_t rec_base,
                                                         const FwdRec* __restrict__ fin_rec,
                                                         FwdRec* __restrict__ fwd_rec,
                                                         uint32_t* __restrict__ fwd_off_global, uint32_t n_rows)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t src = cand_off[row];
    const uint32_t dst = rec_base + fin_off[row];
    const uint32_t n = fin_cnt[row];
    fwd_off_global[row] = dst;
    for (uint32_t i = 0; i < n; ++i) fwd_rec[dst + i] = fin_rec[src + i];
}

int launch_k2_exact(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, uint32_t n_rows, uint32_t n_cand,
                    const float4* segs, const SegRays* rays, const double* midray, const ViewDev* views,
                    const uint32_t* mask, const uint32_t* cand_off, uint32_t* cand_c, uint32_t* cand_row,
                    uint32_t* row_pair, unsigned long long* heap, FwdRec* cand_rec, FwdRec* fin_rec, uint32_t* fin_cnt,
                    float thr, int knn, int max_image_width, int apply_orient, cudaStream_t st)
{
    if (n_ctas == 0) return 0;
    int launches = 0;
    k2_expand_kernel<<<n_ctas, K2_ROWS, 0, st>>>(pairs, ctas, mask, cand_off, cand_c, cand_row, row_pair);
    ++launches;
    if (n_cand) {
        k2a_exact_kernel<<<(n_cand + 127) / 128, 128, 0, st>>>(pairs, row_pair, cand_c, cand_row, n_cand, segs, rays,
                                                                midray, views, cand_rec, thr, (double)max_image_width);
        ++launches;
    }
    k2b_select_kernel<<<(n_rows + K2_ROWS - 1) / K2_ROWS, K2_ROWS, 0, st>>>(cand_off, n_rows, heap, cand_rec, fin_rec,
                                                                            fin_cnt, knn, apply_orient);
    ++launches;
    return launches;
}

int launch_k2_compact(const uint32_t* cand_off, const uint32_t* fin_cnt, const uint32_t* fin_off,
                      uint32_t rec_base, const FwdRec* fin_rec, FwdRec* fwd_rec, uint32_t* fwd_off_global,
                      uint32_t n_rows, cudaStream_t st)
{
    if (n_rows == 0) return 0;
    k2_compact_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(cand_off, fin_cnt, fin_off, rec_base, fin_rec,
                                                             fwd_rec, fwd_off_global, n_rows);
    return 1;
}

}  // namespace l3d
