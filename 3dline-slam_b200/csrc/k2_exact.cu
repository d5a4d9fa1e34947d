// k2_exact.cu -- K2: exact re-evaluation of the K1 candidates in the reference's double sequence,
// two-view triangulation, kNN selection and the orientation filter (exact TU, -fmad=false; every
// arithmetic step is also an explicit _rn intrinsic).
//
// One kernel per batch, one warp per source row (k2_row_kernel): the row's candidate bits -> target
// indices in ascending order, i.e. the order Line3D::matchingCPU pushes matches
// (src/line3D.cc:1124-1196); the sign of the four triangulated depths (Line3D::triangulationDepths,
// src/line3D.cc:1365-1390); the exact pair test of the survivors (src/line3D.cc:1131-1158,
// mutualOverlap :1283-1362); the kNN selection in the pop order of
// std::priority_queue<Match, vector, Match_kNN> (include/commons.h:233-244, src/line3D.cc:1198-1206:
// rank by overlap, and on equal overlaps a replay of the textbook sift-up / sift-down steps libstdc++
// uses); depths and the orientation test (Line3D::checkMatchOrientation src/line3D.cc:962-1014,
// View::segmentQualityAngle src/view.cc:495-513) of the popped matches only.  k2_compact_kernel then
// packs the rows' survivors into the forward-match store.
#include <cstdlib>

#include "detmath.cuh"
#include "exact.cuh"
#include "internal.h"
#include "tma.cuh"

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


// Line3D::mutualOverlap (src/line3D.cc:1283-1362) for the four collinear points a, b, q1, q2 (all with
// z == 1 exactly: x/x = 1 in IEEE arithmetic, so every z difference is exactly 0 and drops out of
// the 3-D norms).  The reference takes the maximum of the six float-rounded distances with a strict
// '>', i.e. the FIRST pair whose float distance equals the maximum.  float(sqrt(.)) is monotone in
// the squared distance, so when the runner-up squared distance is below the maximum by more than
// a relative 2^-21 it cannot round to the same float and the first double maximum is that pair:
// one sqrt instead of six.  Otherwise (and for NaN) the reference sequence is evaluated literally.
__device__ __forceinline__ float mutual_overlap_xy(double ax, double ay, double bx, double by, double q1x, double q1y,
                                                   double q2x, double q2y)
{
    const D3 pt[4] = {d3(ax, ay, 1.0), d3(bx, by, 1.0), d3(q1x, q1y, 1.0), d3(q2x, q2y, 1.0)};
    if (!(point_on_segment(pt[0], pt[2], pt[3]) || point_on_segment(pt[1], pt[2], pt[3]) ||
          point_on_segment(pt[2], pt[0], pt[1]) || point_on_segment(pt[3], pt[0], pt[1])))
        return 0.0f;
    double d2[6];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i + 1; j < 4; ++j) {
                const double dx = ds(pt[i].x, pt[j].x), dy = ds(pt[i].y, pt[j].y);
                d2[k++] = da(dm(dx, dx), dm(dy, dy));  // + dz*dz with dz == 0 exactly
            }
    }
    double m2 = d2[0];
    int im = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k)
        if (d2[k] > m2) {
            m2 = d2[k];
            im = k;
        }
    // runner-up: the largest squared distance strictly below the maximum
    const double lim = dm(m2, 1.0 - 0x1p-21);
    bool clear = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) clear &= (d2[k] == m2) | (d2[k] <= lim);
    if (clear) {
        const float max_dist = (float)__dsqrt_rn(m2);
        if (!(max_dist > 0.0f) || max_dist < 1.0f) return 0.0f;
        double in2 = d2[5];  // inner pair = complement of the outer pair: index 5 - im
        if (im == 1) in2 = d2[4];
        if (im == 2) in2 = d2[3];
        if (im == 3) in2 = d2[2];
        if (im == 4) in2 = d2[1];
        if (im == 5) in2 = d2[0];
        return (float)dd(__dsqrt_rn(in2), (double)max_dist);
    }
    return mutual_overlap(pt);
}

// RN(num / den) > 1e-12, decided without the division whenever the quotient is not within 1e-10
// (relative) of the threshold: rounding is monotone and doubles near 1e-12 are 2e-28 apart
__device__ __forceinline__ bool depth_positive(double num, double den)
{
    if (num == 0.0 || ((num < 0.0) != (den < 0.0))) return false;
    const double an = fabs(num), p = L3D_EPS * fabs(den);
    if (an > p * 1.0000000001) return true;
    if (an < p * 0.9999999999) return false;
    return dd(num, den) > L3D_EPS;
}

static constexpr int K2_WARPS = 4;     // warps per CTA, one row per warp at a time
static constexpr int K2_SUB = 64;      // rows per CTA (a quarter of a K1 tile)
static constexpr int K2_CHUNK = 1024;  // target segments per enumeration chunk (32 mask words)

struct __align__(16) K2WarpSmem {
    float ps[K2_CHUNK];        // exact overlap of the candidates that pass the overlap test
    unsigned short cl[K2_CHUNK];  // chunk-local target index: candidates, then (in place) the passing ones
};

// ---- K2: one warp per source row: enumerate the K1 candidates (ascending target index = the order
// Line3D::matchingCPU pushes matches, src/line3D.cc:1124-1196), phase A: exact pair test of every
// candidate, phase B: triangulation + orientation test of the ones that pass (dense lanes again),
// then the kNN selection in priority-queue pop order and the orientation filter.  Only matches with
// four positive depths are ever written to memory. ----
__global__ void __launch_bounds__(K2_WARPS * 32) k2_row_kernel(
    const PairDev* __restrict__ pairs, const K1Cta* __restrict__ ctas, const uint32_t* __restrict__ mask,
    const uint32_t* __restrict__ cand_off, const float4* __restrict__ segs, const SegRays* __restrict__ rays,
    const double* __restrict__ midray, const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views,
    unsigned long long* __restrict__ heap, FwdRec* __restrict__ cand_rec, FwdRec* __restrict__ fin_rec, uint32_t* __restrict__ fin_cnt, float thr, double W,
    int knn, int apply_orient)
{
    __shared__ K2WarpSmem wsm[K2_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    K2WarpSmem& sm = wsm[warp];
    const K1Cta cta = ctas[blockIdx.x / (K2_ROWS / K2_SUB)];
    const PairDev& P = pairs[cta.pair];
    const uint32_t n_src = P.n_src, n_tgt = P.n_tgt, words = P.words;
    const uint32_t row0 = cta.tile * K2_ROWS + (blockIdx.x % (K2_ROWS / K2_SUB)) * K2_SUB;
    const ViewDev& vs = views[P.src_view];
    const ViewDev& vt = views[P.tgt_view];
    const D3 Cs = ld3(vs.C), Ct = ld3(vt.C);

    for (uint32_t rr = warp; rr < K2_SUB; rr += K2_WARPS) {
        const uint32_t r = row0 + rr;
        if (r >= n_src) break;  // warp-uniform
        const uint32_t lrow = P.row_base - P.batch_row0 + r;
        const uint32_t base = cand_off[lrow];
        // scratch of the row, sized by its K1 candidate count: staging keys (overlap << 32 | target),
        // heap replay keys and the pop order
        unsigned long long* __restrict__ stage = heap + base;
        unsigned long long* __restrict__ hscr = reinterpret_cast<unsigned long long*>(cand_rec + base);
        uint32_t* __restrict__ gsel = reinterpret_cast<uint32_t*>(hscr + (cand_off[lrow + 1] - base));
        FwdRec* __restrict__ frec = fin_rec + base;
        const uint32_t* __restrict__ mrow = mask + P.mask_base + r;

        // row constants (src/line3D.cc:1113-1121)
        const float4 sg = segs[P.src_off + r];
        const D3 p1 = d3((double)sg.x, (double)sg.y, 1.0), p2 = d3((double)sg.z, (double)sg.w, 1.0);
        const D3 e1 = mul33(P.F, p1), e2 = mul33(P.F, p2);
        // triangulation constants of the row: n = (r1 x r2).normalized() and n.C come from k0_prep
        const SegRays sr = rays[P.src_off + r];
        const D3 rp1 = ld3(sr.r1), rp2 = ld3(sr.r2);
        const SegPlane plB = planes[P.src_off + r];
        const D3 nB = ld3(plB.n);
        const double numB = ds(plB.cn, dot3(nB, Ct));
        const D3 rmid = ld3(midray + 3 * (size_t)(P.src_off + r));
        uint32_t n_valid = 0;

        // Candidates of several mask chunks are collected into one list before the two phases run, so
        // that a row of a large target view (C4: 3 chunks, ~18 candidates each) fills the lanes of one
        // pass instead of three sparse ones.  cl holds (target - cl_base) as 16 bits: chunks are batched
        // while the view has <= 65536 segments, otherwise every chunk is flushed on its own.
        const bool batch_chunks = n_tgt <= 65536u;
        uint32_t ncl = 0, cl_base = 0;
        auto flush = [&]() {
            // ---- phase V: which candidates triangulate to four positive depths (src/line3D.cc:1160-1168,
            // 1365-1390)?  A match is pushed iff it passes the pair test AND has four positive depths; the
            // depth signs are the cheaper half (no division: see depth_positive), so they go first and
            // the pair test only runs on the ~half of the candidates that survive.  The divisions are
            // done for the matches that survive the kNN selection. ----
            const uint32_t total = ncl;
            uint32_t nval = 0;
            for (uint32_t k0 = 0; k0 < total; k0 += 32) {
                const uint32_t k = k0 + lane;
                const bool active = k < total;
                const uint32_t cidx = active ? (uint32_t)sm.cl[k] : 0u;
                bool valid = false;
                if (active) {
                    const uint32_t c = cl_base + cidx;
                    const SegRays tr = rays[P.tgt_off + c];
                    const SegPlane plA = planes[P.tgt_off + c];
                    const D3 nA = ld3(plA.n);
                    const double a1 = dot3(rp1, nA), a2 = dot3(rp2, nA);
                    const double b1 = dot3(ld3(tr.r1), nB), b2 = dot3(ld3(tr.r2), nB);
                    if (!(fabs(a1) < L3D_EPS || fabs(a2) < L3D_EPS || fabs(b1) < L3D_EPS || fabs(b2) < L3D_EPS)) {
                        const double num = ds(plA.cn, dot3(nA, Cs));
                        valid = depth_positive(num, a1) && depth_positive(num, a2) && depth_positive(numB, b1) &&
                                depth_positive(numB, b2);
                    }
                }
                __syncwarp();  // every lane has read its cl[k] before the in-place compaction
                const uint32_t bal = __ballot_sync(0xffffffffu, valid);
                if (valid) sm.cl[nval + __popc(bal & lt_mask)] = (unsigned short)cidx;
                nval += __popc(bal);
                __syncwarp();
            }
            // ---- phase A: the pair test in the reference's double sequence (src/line3D.cc:1131-1158) ----
            for (uint32_t k0 = 0; k0 < nval; k0 += 32) {
                const uint32_t k = k0 + lane;
                uint32_t c = 0;
                float score = 0.0f;
                bool pass = false;
                if (k < nval) {
                    c = cl_base + (uint32_t)sm.cl[k];
                    const float4 tg = segs[P.tgt_off + c];
                    const D3 q1 = d3((double)tg.x, (double)tg.y, 1.0), q2 = d3((double)tg.z, (double)tg.w, 1.0);
                    const D3 l2 = cross3(q1, q2);
                    const D3 a = cross3(l2, e1), b = cross3(l2, e2);
                    if (fabs(a.z) > L3D_EPS && fabs(b.z) > L3D_EPS) {
                        const double ax = dd(a.x, a.z), ay = dd(a.y, a.z), bx = dd(b.x, b.z), by = dd(b.y, b.z);
                        if (!(ax < 0 || ax > W || ay < 0 || ay > W || bx < 0 || bx > W || by < 0 || by > W)) {
                            score = mutual_overlap_xy(ax, ay, bx, by, q1.x, q1.y, q2.x, q2.y);
                            pass = score > thr;
                        }
                    }
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                if (pass) stage[n_valid + __popc(bal & lt_mask)] = ((unsigned long long)__float_as_uint(score) << 32) | c;
                n_valid += __popc(bal);
            }
            __syncwarp();
            ncl = 0;
        };

        for (uint32_t cb = 0; cb < n_tgt; cb += K2_CHUNK) {
            // ---- enumerate the candidates of this chunk in ascending target order ----
            const uint32_t w = (cb >> 5) + lane;
            uint32_t bits = (w < words) ? mrow[(size_t)w * n_src] : 0u;
            const uint32_t cnt = __popc(bits);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;  // warp-uniform
            if (ncl && (!batch_chunks || ncl + total > (uint32_t)K2_CHUNK)) flush();
            if (!batch_chunks) cl_base = cb;
            uint32_t off = ncl + incl - cnt;
            while (bits) {
                const uint32_t j = __ffs(bits) - 1;
                bits &= bits - 1;
                sm.cl[off++] = (unsigned short)(cb - cl_base + lane * 32 + j);
            }
            ncl += total;
            __syncwarp();
        }
        if (ncl) flush();

        // ---- selection: std::priority_queue pop order (include/commons.h:233-244, src/line3D.cc:1198-1206):
        // sel[t] = staging index of the t-th popped match ----
        __syncwarp();
        uint32_t npop = n_valid;   // kNN <= 0: every match, ascending target order
        int sel_mode = 0;          // 0: identity, 1: sm.cl (shared), 2: gsel (global)
        if (knn > 0 && n_valid) {
            npop = min((uint32_t)knn, n_valid);
            bool fast = n_valid <= K2_CHUNK;
            if (fast) {
                // distinct overlaps pop in descending order: rank = number of strictly larger overlaps;
                // equal overlaps among the popped ones show up as fewer than npop distinct ranks below npop
                for (uint32_t k = lane; k < n_valid; k += 32) sm.ps[k] = key_overlap(stage[k]);
                if (lane < 4 && n_valid + lane < K2_CHUNK) sm.ps[n_valid + lane] = -1.0f;  // pad: never larger
                __syncwarp();
                uint32_t seen = 0, rankbits = 0;
                const uint32_t n4 = (n_valid + 3u) >> 2;
                for (uint32_t k = lane; k < n_valid; k += 32) {
                    const float ov = sm.ps[k];
                    uint32_t rank = 0;
                    if (n4 * 4 <= K2_CHUNK) {
                        for (uint32_t j = 0; j < n4; ++j) {
                            const float4 o = reinterpret_cast<const float4*>(sm.ps)[j];
                            rank += (o.x > ov) + (o.y > ov) + (o.z > ov) + (o.w > ov);
                        }
                    } else {
                        for (uint32_t j = 0; j < n_valid; ++j) rank += (sm.ps[j] > ov) ? 1u : 0u;
                    }
                    if (rank < npop) {
                        sm.cl[rank] = (unsigned short)k;  // colliding ranks are detected below
                        rankbits |= 1u << (rank & 31u);
                        ++seen;
                    }
                }
                if (npop <= 32) {  // npop distinct ranks below npop <=> no two of them are equal
                    fast = (uint32_t)__popc(__reduce_or_sync(0xffffffffu, rankbits)) == npop;
                } else {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, d);
                    fast = seen == npop;
                    if (fast) {  // a collision leaves a slot stale: every slot must hold its own rank
                        uint32_t ok = 1;
                        for (uint32_t t = lane; t < npop; t += 32) {
                            const float ov = sm.ps[sm.cl[t]];
                            uint32_t rank = 0;
                            for (uint32_t j = 0; j < n_valid; ++j) rank += (sm.ps[j] > ov) ? 1u : 0u;
                            ok &= (rank == t) ? 1u : 0u;
                        }
                        fast = __all_sync(0xffffffffu, ok != 0u);
                    }
                }
                __syncwarp();
                sel_mode = 1;
            }
            if (!fast) {
                // equal overlaps: replay the binary heap (push in ascending target order, pop kNN)
                sel_mode = 2;
                if (lane == 0) {
                    for (uint32_t i = 0; i < n_valid; ++i)
                        heap_push(hscr, i, (stage[i] & 0xffffffff00000000ull) | i);
                    uint32_t hn = n_valid;
                    for (uint32_t t = 0; t < npop; ++t) {
                        gsel[t] = (uint32_t)(hscr[0] & 0xffffffffu);
                        heap_pop(hscr, hn);
                        --hn;
                    }
                }
                __syncwarp();
            }
        }

        // ---- the popped matches: depths (src/line3D.cc:1365-1390), the orientation test of the would-be
        // match (checkMatchOrientation, src/line3D.cc:962-1014), output in pop order ----
        uint32_t nout = 0;
        for (uint32_t t0 = 0; t0 < npop; t0 += 32) {
            const uint32_t t = t0 + lane;
            bool keep = false;
            FwdRec rec;
            rec.flags = 0u;
            rec.score = 0.0f;
            if (t < npop) {
                const uint32_t k = sel_mode == 0 ? t : (sel_mode == 1 ? (uint32_t)sm.cl[t] : gsel[t]);
                const unsigned long long key = stage[k];
                const uint32_t c = (uint32_t)(key & 0xffffffffu);
                const SegRays tr = rays[P.tgt_off + c];
                const SegPlane plA = planes[P.tgt_off + c];
                const D3 nA = ld3(plA.n);
                const double num = ds(plA.cn, dot3(nA, Cs));
                // n.ray(p1): the same products in the same order as ray(p1).n
                const double ds1 = dd(num, dot3(rp1, nA)), ds2 = dd(num, dot3(rp2, nA));
                const double dt1 = dd(numB, dot3(ld3(tr.r1), nB)), dt2 = dd(numB, dot3(ld3(tr.r2), nB));
                rec.c = c;
                rec.overlap = key_overlap(key);
                rec.d_p1 = (float)ds1;
                rec.d_p2 = (float)ds2;
                rec.d_q1 = (float)dt1;
                rec.d_q2 = (float)dt2;
                keep = true;
                if (apply_orient) {
                    // The test is acos(x) in (0.0982, 3.0434) with x = ray(mid) . dir, dir = (P2-P1)/|P2-P1|,
                    // i.e. |x| < 0.99518...: when (ray(mid).(P2-P1))^2 < 0.9951^2 |P2-P1|^2 the exact x
                    // (relative error ~1e-15) is inside by a margin of 8e-5 and no sqrt/div/acos is needed.
                    const D3 P1 = add3(Cs, scale3(rp1, (double)rec.d_p1));
                    const D3 P2 = add3(Cs, scale3(rp2, (double)rec.d_p2));
                    const D3 vv = sub3(P2, P1);
                    const double v2 = dot3(vv, vv), sv = dot3(rmid, vv);
                    if (!(v2 > 1e-20 && sv * sv < 0.99022401 * v2)) {
                        const float len = (float)norm3(sub3(P1, P2));
                        D3 dir = d3(0.0, 0.0, 0.0);
                        if (len > L3D_EPS) dir = normalized3(vv);
                        const double ang = det_acos(fmin(fmax(dot3(rmid, dir), -1.0), 1.0));
                        keep = ang > (double)0.098174771f && ang < (double)3.043417886f;
                    }
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) frec[nout + __popc(bal & lt_mask)] = rec;
            nout += __popc(bal);
        }
        if (lane == 0) fin_cnt[lrow] = nout;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K2, tile variant: one CTA = 128 source rows of one pair (8 warps, 16 rows each) and the target
// tables of the pair -- endpoint rays and plane records of up to 1024 target segments -- staged in
// shared memory with two 1-D TMA bulk copies per chunk.  The gathers of phase V and of the finishing
// pass (the stalls of the row kernel) become shared-memory reads.  Same phases, same arithmetic.
// ------------------------------------------------------------------------------------------
static constexpr int K2T_WARPS = 8;
static constexpr int K2T_ROWS = 128;    // rows per CTA (half a K1 tile)
static constexpr int K2T_TCH = 1024;    // target segments staged per chunk (32 mask words)
static constexpr int K2T_CL = 256;      // candidates enumerated per pass
static constexpr int K2T_KEEP = 256;    // matches of one row kept in shared memory for the selection

struct __align__(16) K2TWarp {
    float ps[K2T_KEEP];           // overlap of the row's matches (ascending target order)
    uint32_t sc[K2T_KEEP];        // their target segments
    unsigned short cl[K2T_CL];    // chunk-local candidate indices (compacted in place); later the pop order
};

struct K2TSmem {
    SegRays rays[K2T_TCH];
    SegPlane planes[K2T_TCH];
    K2TWarp w[K2T_WARPS];
    uint32_t nv[K2T_ROWS];  // matches per row so far (pairs with more than one chunk)
    uint64_t bar;
};

__global__ void __launch_bounds__(K2T_WARPS * 32, 2) k2_tile_kernel(
    const PairDev* __restrict__ pairs, const K1Cta* __restrict__ ctas, const uint32_t* __restrict__ mask,
    const uint32_t* __restrict__ cand_off, const float4* __restrict__ segs, const SegRays* __restrict__ rays,
    const double* __restrict__ midray, const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views,
    unsigned long long* __restrict__ heap, FwdRec* __restrict__ cand_rec, FwdRec* __restrict__ fin_rec,
    uint32_t* __restrict__ fin_cnt, float thr, double W, int knn, int apply_orient)
{
    extern __shared__ __align__(128) unsigned char k2t_raw[];
    K2TSmem& S = *reinterpret_cast<K2TSmem*>(k2t_raw);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    K2TWarp& sm = S.w[warp];
    const K1Cta cta = ctas[blockIdx.x / (K2_ROWS / K2T_ROWS)];
    const PairDev& P = pairs[cta.pair];
    const uint32_t n_src = P.n_src, n_tgt = P.n_tgt, words = P.words;
    const uint32_t row0 = cta.tile * K2_ROWS + (blockIdx.x % (K2_ROWS / K2T_ROWS)) * K2T_ROWS;
    if (row0 >= n_src) return;  // uniform
    const ViewDev& vs = views[P.src_view];
    const ViewDev& vt = views[P.tgt_view];
    const D3 Cs = ld3(vs.C), Ct = ld3(vt.C);
    const bool single = n_tgt <= (uint32_t)K2T_TCH;

    if (threadIdx.x == 0) {
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < (uint32_t)K2T_ROWS; i += K2T_WARPS * 32) S.nv[i] = 0u;
    __syncthreads();

    uint32_t chunk_no = 0;
    for (uint32_t cb = 0; cb < n_tgt; cb += K2T_TCH, ++chunk_no) {
        const uint32_t tcnt = min((uint32_t)K2T_TCH, n_tgt - cb);
        if (threadIdx.x == 0) {
            mbar_expect_tx(&S.bar, tcnt * (uint32_t)(sizeof(SegRays) + sizeof(SegPlane)));
            tma_load_1d(S.rays, rays + P.tgt_off + cb, tcnt * (uint32_t)sizeof(SegRays), &S.bar);
            tma_load_1d(S.planes, planes + P.tgt_off + cb, tcnt * (uint32_t)sizeof(SegPlane), &S.bar);
        }
        mbar_wait(&S.bar, chunk_no & 1u);

        for (uint32_t rr = warp; rr < (uint32_t)K2T_ROWS; rr += K2T_WARPS) {
            const uint32_t r = row0 + rr;
            if (r >= n_src) break;  // warp-uniform
            const uint32_t lrow = P.row_base - P.batch_row0 + r;
            const uint32_t base = cand_off[lrow];
            unsigned long long* __restrict__ stage = heap + base;
            unsigned long long* __restrict__ hscr = reinterpret_cast<unsigned long long*>(cand_rec + base);
            uint32_t* __restrict__ gsel = reinterpret_cast<uint32_t*>(hscr + (cand_off[lrow + 1] - base));
            FwdRec* __restrict__ frec = fin_rec + base;
            const uint32_t* __restrict__ mrow = mask + P.mask_base + r;

            // row constants (src/line3D.cc:1113-1121); plane normal and n.C come from k0_prep
            const float4 sg = segs[P.src_off + r];
            const D3 p1 = d3((double)sg.x, (double)sg.y, 1.0), p2 = d3((double)sg.z, (double)sg.w, 1.0);
            const D3 e1 = mul33(P.F, p1), e2 = mul33(P.F, p2);
            const SegRays sr = rays[P.src_off + r];
            const D3 rp1 = ld3(sr.r1), rp2 = ld3(sr.r2);
            const SegPlane plB = planes[P.src_off + r];
            const D3 nB = ld3(plB.n);
            const double numB = ds(plB.cn, dot3(nB, Ct));
            uint32_t n_valid = single ? 0u : S.nv[rr];

            // ---- the chunk's mask words of this row; candidates are enumerated <= K2T_CL at a time ----
            const uint32_t w = (cb >> 5) + lane;
            const uint32_t allbits = (w < words) ? mrow[(size_t)w * n_src] : 0u;
            uint32_t tot_all = __popc(allbits);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) tot_all += __shfl_xor_sync(0xffffffffu, tot_all, d);
            if (tot_all) {
                const uint32_t wstep = tot_all <= (uint32_t)K2T_CL ? 32u : 8u;  // 8 words hold <= 256 candidates
                for (uint32_t w0 = 0; w0 < 32u; w0 += wstep) {
                    uint32_t bits = (lane >= w0 && lane < w0 + wstep) ? allbits : 0u;
                    const uint32_t cnt = __popc(bits);
                    uint32_t incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                        if ((int)lane >= d) incl += t;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total == 0) continue;  // warp-uniform
                    uint32_t off = incl - cnt;
                    while (bits) {
                        const uint32_t j = __ffs(bits) - 1;
                        bits &= bits - 1;
                        sm.cl[off++] = (unsigned short)(lane * 32 + j);
                    }
                    __syncwarp();

                    // ---- phase V: four positive depths?  (see k2_row_kernel) target records from shared memory ----
                    uint32_t nval = 0;
                    for (uint32_t k0 = 0; k0 < total; k0 += 32) {
                        const uint32_t k = k0 + lane;
                        const bool active = k < total;
                        const uint32_t cidx = active ? (uint32_t)sm.cl[k] : 0u;
                        bool valid = false;
                        if (active) {
                            const SegRays& tr = S.rays[cidx];
                            const SegPlane& plA = S.planes[cidx];
                            const D3 nA = ld3(plA.n);
                            const double a1 = dot3(rp1, nA), a2 = dot3(rp2, nA);
                            const double b1 = dot3(ld3(tr.r1), nB), b2 = dot3(ld3(tr.r2), nB);
                            if (!(fabs(a1) < L3D_EPS || fabs(a2) < L3D_EPS || fabs(b1) < L3D_EPS || fabs(b2) < L3D_EPS)) {
                                const double num = ds(plA.cn, dot3(nA, Cs));
                                valid = depth_positive(num, a1) && depth_positive(num, a2) &&
                                        depth_positive(numB, b1) && depth_positive(numB, b2);
                            }
                        }
                        __syncwarp();  // every lane has read its cl[k] before the in-place compaction
                        const uint32_t bal = __ballot_sync(0xffffffffu, valid);
                        if (valid) sm.cl[nval + __popc(bal & lt_mask)] = (unsigned short)cidx;
                        nval += __popc(bal);
                        __syncwarp();
                    }

                    // ---- phase A: the pair test in the reference's double sequence (src/line3D.cc:1131-1158) ----
                    for (uint32_t k0 = 0; k0 < nval; k0 += 32) {
                        const uint32_t k = k0 + lane;
                        uint32_t c = 0;
                        float score = 0.0f;
                        bool pass = false;
                        if (k < nval) {
                            c = cb + (uint32_t)sm.cl[k];
                            const float4 tg = segs[P.tgt_off + c];
                            const D3 q1 = d3((double)tg.x, (double)tg.y, 1.0), q2 = d3((double)tg.z, (double)tg.w, 1.0);
                            const D3 l2 = cross3(q1, q2);
                            const D3 a = cross3(l2, e1), b = cross3(l2, e2);
                            if (fabs(a.z) > L3D_EPS && fabs(b.z) > L3D_EPS) {
                                const double ax = dd(a.x, a.z), ay = dd(a.y, a.z), bx = dd(b.x, b.z), by = dd(b.y, b.z);
                                if (!(ax < 0 || ax > W || ay < 0 || ay > W || bx < 0 || bx > W || by < 0 || by > W)) {
                                    score = mutual_overlap_xy(ax, ay, bx, by, q1.x, q1.y, q2.x, q2.y);
                                    pass = score > thr;
                                }
                            }
                        }
                        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                        if (pass) {
                            const uint32_t pos = n_valid + __popc(bal & lt_mask);
                            if (single && pos < (uint32_t)K2T_KEEP) {
                                sm.ps[pos] = score;
                                sm.sc[pos] = c;
                            } else {
                                stage[pos] = ((unsigned long long)__float_as_uint(score) << 32) | c;
                            }
                        }
                        n_valid += __popc(bal);
                    }
                    __syncwarp();
                }
            }
            if (!single) {
                if (lane == 0) S.nv[rr] = n_valid;
                if (cb + K2T_TCH < n_tgt) continue;  // more chunks to come for this row
                // last chunk: bring the row's matches into shared memory for the selection
                __syncwarp();
                for (uint32_t k = lane; k < n_valid && k < (uint32_t)K2T_KEEP; k += 32) {
                    const unsigned long long key = stage[k];
                    sm.ps[k] = key_overlap(key);
                    sm.sc[k] = (uint32_t)(key & 0xffffffffu);
                }
            }
            __syncwarp();
            auto match_key = [&](uint32_t k) -> unsigned long long {
                return k < (uint32_t)K2T_KEEP ? (((unsigned long long)__float_as_uint(sm.ps[k]) << 32) | sm.sc[k]) : stage[k];
            };

            // ---- selection: std::priority_queue pop order (include/commons.h:233-244, src/line3D.cc:1198-1206) ----
            uint32_t npop = n_valid;   // kNN <= 0: every match, ascending target order
            int sel_mode = 0;          // 0: identity, 1: sm.cl (shared), 2: gsel (global)
            if (knn > 0 && n_valid) {
                npop = min((uint32_t)knn, n_valid);
                bool fast = n_valid + 4 <= (uint32_t)K2T_KEEP && npop <= (uint32_t)K2T_CL;
                if (fast) {
                    if (lane < 4) sm.ps[n_valid + lane] = -1.0f;  // pad to a multiple of 4: never larger
                    __syncwarp();
                    uint32_t seen = 0, rankbits = 0;
                    const uint32_t n4 = (n_valid + 3u) >> 2;
                    for (uint32_t k = lane; k < n_valid; k += 32) {
                        const float ov = sm.ps[k];
                        uint32_t rank = 0;
                        for (uint32_t j = 0; j < n4; ++j) {
                            const float4 o = reinterpret_cast<const float4*>(sm.ps)[j];
                            rank += (o.x > ov) + (o.y > ov) + (o.z > ov) + (o.w > ov);
                        }
                        if (rank < npop) {
                            sm.cl[rank] = (unsigned short)k;  // colliding ranks are detected below
                            rankbits |= 1u << (rank & 31u);
                            ++seen;
                        }
                    }
                    if (npop <= 32) {  // npop distinct ranks below npop <=> no two of them are equal
                        fast = (uint32_t)__popc(__reduce_or_sync(0xffffffffu, rankbits)) == npop;
                    } else {
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, d);
                        fast = seen == npop;
                        __syncwarp();
                        if (fast) {  // a collision leaves a slot stale: every slot must hold its own rank
                            uint32_t ok = 1;
                            for (uint32_t t = lane; t < npop; t += 32) {
                                const uint32_t kk = sm.cl[t];
                                uint32_t rank = 0xffffffffu;
                                if (kk < n_valid) {
                                    const float ov = sm.ps[kk];
                                    rank = 0;
                                    for (uint32_t j = 0; j < n_valid; ++j) rank += (sm.ps[j] > ov) ? 1u : 0u;
                                }
                                ok &= (rank == t) ? 1u : 0u;
                            }
                            fast = __all_sync(0xffffffffu, ok != 0u);
                        }
                    }
                    __syncwarp();
                    sel_mode = 1;
                }
                if (!fast) {
                    // equal overlaps (or a very long row): replay the binary heap (push in ascending target
                    // order, pop kNN)
                    sel_mode = 2;
                    if (lane == 0) {
                        for (uint32_t i = 0; i < n_valid; ++i)
                            heap_push(hscr, i, (match_key(i) & 0xffffffff00000000ull) | i);
                        uint32_t hn = n_valid;
                        for (uint32_t t = 0; t < npop; ++t) {
                            gsel[t] = (uint32_t)(hscr[0] & 0xffffffffu);
                            heap_pop(hscr, hn);
                            --hn;
                        }
                    }
                    __syncwarp();
                }
            }

            // ---- the popped matches: depths + orientation test (src/line3D.cc:1365-1390, 962-1014), output in
            // pop order; the target records come from shared memory when the pair has a single chunk ----
            const D3 rmid = ld3(midray + 3 * (size_t)(P.src_off + r));
            uint32_t nout = 0;
            for (uint32_t t0 = 0; t0 < npop; t0 += 32) {
                const uint32_t t = t0 + lane;
                bool keep = false;
                FwdRec rec;
                rec.flags = 0u;
                rec.score = 0.0f;
                if (t < npop) {
                    const uint32_t k = sel_mode == 0 ? t : (sel_mode == 1 ? (uint32_t)sm.cl[t] : gsel[t]);
                    const unsigned long long key = match_key(k);
                    const uint32_t c = (uint32_t)(key & 0xffffffffu);
                    const SegRays tr = single ? S.rays[c] : rays[P.tgt_off + c];
                    const SegPlane plA = single ? S.planes[c] : planes[P.tgt_off + c];
                    const D3 nA = ld3(plA.n);
                    const double num = ds(plA.cn, dot3(nA, Cs));
                    // n.ray(p1): the same products in the same order as ray(p1).n
                    const double ds1 = dd(num, dot3(rp1, nA)), ds2 = dd(num, dot3(rp2, nA));
                    const double dt1 = dd(numB, dot3(ld3(tr.r1), nB)), dt2 = dd(numB, dot3(ld3(tr.r2), nB));
                    rec.c = c;
                    rec.overlap = key_overlap(key);
                    rec.d_p1 = (float)ds1;
                    rec.d_p2 = (float)ds2;
                    rec.d_q1 = (float)dt1;
                    rec.d_q2 = (float)dt2;
                    keep = true;
                    if (apply_orient) {
                        // |x| < 0.9951 (thresholds: |x| < 0.99518): no sqrt/div/acos needed, see k2_row_kernel
                        const D3 P1 = add3(Cs, scale3(rp1, (double)rec.d_p1));
                        const D3 P2 = add3(Cs, scale3(rp2, (double)rec.d_p2));
                        const D3 vv = sub3(P2, P1);
                        const double v2 = dot3(vv, vv), sv = dot3(rmid, vv);
                        if (!(v2 > 1e-20 && sv * sv < 0.99022401 * v2)) {
                            const float len = (float)norm3(sub3(P1, P2));
                            D3 dir = d3(0.0, 0.0, 0.0);
                            if (len > L3D_EPS) dir = normalized3(vv);
                            const double ang = det_acos(fmin(fmax(dot3(rmid, dir), -1.0), 1.0));
                            keep = ang > (double)0.098174771f && ang < (double)3.043417886f;
                        }
                    }
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (keep) frec[nout + __popc(bal & lt_mask)] = rec;
                nout += __popc(bal);
            }
            if (lane == 0) fin_cnt[lrow] = nout;
            __syncwarp();
        }
        __syncthreads();  // everyone is done with the staged tables before the next chunk overwrites them
    }
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
                    const float4* segs, const SegRays* rays, const double* midray, const SegPlane* planes,
                    const ViewDev* views, const uint32_t* mask, const uint32_t* cand_off, unsigned long long* heap, FwdRec* cand_rec,
                    FwdRec* fin_rec, uint32_t* fin_cnt, float thr, int knn, int max_image_width, int apply_orient,
                    uint32_t max_tgt, cudaStream_t st)
{
    if (n_ctas == 0) return 0;
    (void)n_rows;
    (void)n_cand;
    // test / tuning hook: 0 = row kernel always, 1 = by target-view size (default), 2 = tile kernel always
    const char* ev = getenv("L3D_K2_VARIANT");
    const int variant = ev ? atoi(ev) : 1;
    // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
    cudaFuncSetAttribute(k2_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2TSmem));
    // the tile kernel keeps a pair's whole target table in shared memory when it has <= 1024 segments
    // (C2: 0.93 vs 1.00 ms); with several chunks per pair the row kernel is the faster one (C4 shape)
    if (variant == 2 || (variant == 1 && max_tgt <= (uint32_t)K2T_TCH)) {
        k2_tile_kernel<<<n_ctas * (K2_ROWS / K2T_ROWS), K2T_WARPS * 32, sizeof(K2TSmem), st>>>(
            pairs, ctas, mask, cand_off, segs, rays, midray, planes, views, heap, cand_rec, fin_rec, fin_cnt, thr,
            (double)max_image_width, knn, apply_orient);
        return 1;
    }
    // 128 registers, 4 CTAs = 16 warps per SM: compiling for more resident warps (80 / 64 registers)
    // spills and was measured slower (1.11 ms -> 1.11 / 1.39 ms on C2)
    k2_row_kernel<<<n_ctas * (K2_ROWS / K2_SUB), K2_WARPS * 32, 0, st>>>(pairs, ctas, mask, cand_off, segs, rays, midray,
                                                                        planes, views, heap, cand_rec, fin_rec, fin_cnt,
                                                                        thr, (double)max_image_width, knn, apply_orient);
    return 1;
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
