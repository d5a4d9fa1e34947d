// k2_exact.cu -- K2: exact re-evaluation of the K1 candidates in the reference's double sequence,
// two-view triangulation, kNN selection and the orientation filter (exact TU, -fmad=false; every
// arithmetic step is also an explicit _rn intrinsic).
//
// The fast path is k2a_filter / k2b_exact / k2c_select below ("certified contenders"); k2_row_kernel is the
// literal evaluation of one row by one warp, used for the rows the fast path hands back and as the whole-batch
// variant L3D_K2_VARIANT=0: the row's candidate bits -> target
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
template <bool LIST>
__global__ void __launch_bounds__(K2_WARPS * 32) k2_row_kernel(
    const PairDev* __restrict__ pairs, const K1Cta* __restrict__ ctas, const uint32_t* __restrict__ mask,
    const uint32_t* __restrict__ cand_off, const float4* __restrict__ segs, const SegRays* __restrict__ rays,
    const double* __restrict__ midray, const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views,
    unsigned long long* __restrict__ heap, FwdRec* __restrict__ cand_rec, FwdRec* __restrict__ fin_rec, uint32_t* __restrict__ fin_cnt, float thr, double W,
    int knn, int apply_orient, const uint2* __restrict__ fb_rows, const uint32_t* __restrict__ ctr,
    const uint32_t* __restrict__ iperm)
{
    __shared__ K2WarpSmem wsm[K2_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    K2WarpSmem& sm = wsm[warp];
    // LIST: the rows named by fb_rows[0 .. ctr[1]) (pair, row), dealt to the warps of the grid; otherwise the
    // 64 rows of this CTA's quarter of a K1 tile
    K1Cta cta = K1Cta{0u, 0u};
    uint32_t row0 = 0;
    if (!LIST) {
        cta = ctas[blockIdx.x / (K2_ROWS / K2_SUB)];
        row0 = cta.tile * K2_ROWS + (blockIdx.x % (K2_ROWS / K2_SUB)) * K2_SUB;
    }
    const uint32_t it_end = LIST ? ctr[1] : (uint32_t)K2_SUB;
    const uint32_t it_step = LIST ? gridDim.x * K2_WARPS : (uint32_t)K2_WARPS;

    for (uint32_t it = LIST ? blockIdx.x * K2_WARPS + warp : warp; it < it_end; it += it_step) {
        uint32_t pair_idx = cta.pair, r = row0 + it;
        if (LIST) {
            const uint2 f = fb_rows[it];
            pair_idx = f.x;
            r = f.y;
        }
        const PairDev& P = pairs[pair_idx];
        const uint32_t n_src = P.n_src, n_tgt = P.n_tgt, words = P.words;
        if (r >= n_src) break;  // warp-uniform (never taken in LIST mode)
        const ViewDev& vs = views[P.src_view];
        const ViewDev& vt = views[P.tgt_view];
        const D3 Cs = ld3(vs.C), Ct = ld3(vt.C);
        const uint32_t lrow = P.row_base - P.batch_row0 + r;
        const uint32_t base = cand_off[lrow];
        // scratch of the row, sized by its K1 candidate count: staging keys (overlap << 32 | target),
        // heap replay keys and the pop order
        unsigned long long* __restrict__ stage = heap + base;
        unsigned long long* __restrict__ hscr = reinterpret_cast<unsigned long long*>(cand_rec + base);
        uint32_t* __restrict__ gsel = reinterpret_cast<uint32_t*>(hscr + (cand_off[lrow + 1] - base));
        FwdRec* __restrict__ frec = fin_rec + base;
        const uint32_t* __restrict__ mrow = mask + P.mask_base + iperm[lrow];  // the mask is in K1's sorted row order

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
// K2, fast path ("certified contenders").  Per source row K1 leaves ~50 candidates, ~27 of them triangulate to
// four positive depths, and only the kNN = 10 best overlaps are popped.  The exact double sequence is therefore
// evaluated only where it decides something, and the per-item steps run one thread per item on dense lanes:
//
//   k2_front    (thread per row, target tables in shared memory)  walks the row's mask bits in ascending target
//                                      order (the order Line3D::matchingCPU pushes matches) and, per candidate, in
//                                      FP32: (i) the four depth signs from the FP32 image of the triangulation
//                                      tables -- certainly positive / certainly not / unknown; (ii) a bracket
//                                      [Lb, U] of the overlap the reference would compute, with the error bounds of
//                                      K1's formulation.  Certain failures are dropped, and so is a candidate whose
//                                      upper bound lies below the lower bounds of kNN candidates that CERTAINLY are
//                                      matches: it can never be popped.  The rest -- the contenders -- are queued.
//   k2b_exact   (thread per contender) the reference's pair test (src/line3D.cc:1131-1158) and, for "unknown"
//                                      depth signs, the exact sign test.
//   k2_rank     (thread per row)       the contenders that are matches, in priority-queue pop order.
//   k2_finish   (thread per popped match) depths and orientation test (src/line3D.cc:1365-1390, 962-1014).
//   k2_row_kernel over a row list      rows the fast path cannot decide -- equal overlaps among the popped
//                                      matches (the heap replay needs every match of the row), more than
//                                      K2_MAXC contenders -- are redone from scratch by the literal row kernel.
//
// Measured on the way here (C2, 1.19e7 candidates): one warp per row for the front and the rank/finish steps:
// 0.38 + 0.35 ms against 0.05 ms for the exact kernel -- per-row latency chains, not arithmetic, were the cost;
// flat kernels gathering the target tables from global memory: expand 0.16 + filter 0.18 + select 0.21 ms, bound
// by L1 tag cycles (one per lane and load for scattered 16-byte reads) and by scattered 8-byte stores.
//
// Exactness: pruning only removes candidates that are strictly beaten by kNN certain matches, so the popped
// set and its order are those of the full list whenever the popped overlaps are distinct; ties are detected
// on the contenders (a pruned candidate cannot tie with a popped one) and sent to the row kernel.  Error
// bounds are stated next to the code.
// ------------------------------------------------------------------------------------------
static constexpr int K2_MAXC = 64;      // queued survivors of one row the rank kernel holds per thread
static constexpr int K2_TOPK = 16;      // largest kNN the select kernel prunes for (register-resident)
#define K2_ROW_FALLBACK 0xffffffffu

__device__ __forceinline__ float rcp_approx_ftz(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// exclusive prefix of v over the warp and the warp total
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t& total)
{
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

// FP32 constants of one source row for the certified tests
struct RowF32 {
    RowEpi32 e;                                 // K1's normalised epipolar lines, error bound, smaller line norm
    float rp1x, rp1y, rp1z, rp2x, rp2y, rp2z;   // endpoint rays
    float nBx, nBy, nBz;                        // normal of the row's own plane
    float Csx, Csy, Csz, CsL1;                  // source centre
    float sB;                                   // exact sign of the row's plane against the target centre: +-1, 0 = unknown
};

// One K1 candidate (row constants R, target tables T / d0 / d1) in FP32:
//  * depth signs (Line3D::triangulationDepths, src/line3D.cc:1365-1390): rays and normals are unit vectors rounded
//    once to float, |error(a)| <= 6u < eps_a / 2; num = cn - n.Cs has |error| <= 4u|cn| + 6u|Cs|_1.  Equal classes =>
//    quotient > eps/2 >> 1e-12 (valid depth); opposite classes => quotient < 0 (no match); else unknown.
//  * a bracket [Lb, U] of the overlap the reference would compute, K1's formulation and error terms.
// Returns false if the candidate CERTAINLY is no match.  Lb > -inf only if it CERTAINLY is one.
__device__ __forceinline__ bool certify_candidate(const RowF32& R, const SegV32& T, const float4 d0, const float4 d1, float thr,
                                                  float& U, float& Lb, bool& unk)
{
    const float u = 5.9604645e-08f;  // 2^-24
    const float ninf = __int_as_float(0xff800000), pinf = __int_as_float(0x7f800000);
    const float a1 = fmaf(R.rp1x, T.nx, fmaf(R.rp1y, T.ny, R.rp1z * T.nz));
    const float a2 = fmaf(R.rp2x, T.nx, fmaf(R.rp2y, T.ny, R.rp2z * T.nz));
    const float b1 = fmaf(T.r1x, R.nBx, fmaf(T.r1y, R.nBy, T.r1z * R.nBz));
    const float b2 = fmaf(T.r2x, R.nBx, fmaf(T.r2y, R.nBy, T.r2z * R.nBz));
    const float num = T.cn - fmaf(T.nx, R.Csx, fmaf(T.ny, R.Csy, T.nz * R.Csz));
    const float eps_n = fmaf(16.0f * u, fabsf(T.cn) + R.CsL1, 1e-9f);
    const float eps_a = 1e-6f;
    // with sn = sign(num), sB = sign(numB) (0 when unknown): a quotient is certainly positive iff the signed
    // denominator exceeds eps, certainly negative iff it is below -eps
    const float sn = fabsf(num) > eps_n ? copysignf(1.0f, num) : 0.0f;
    const float sB = R.sB;
    const float q1 = a1 * sn, q2 = a2 * sn, q3 = b1 * sB, q4 = b2 * sB;
    const bool vpos = fminf(fminf(q1, q2), fminf(q3, q4)) > eps_a;
    const bool vneg = fminf(fminf(q1, q2), fminf(q3, q4)) < -eps_a;
    // ---- overlap bracket (k1_pairtest.cu) ----
    const RowEpi32& e = R.e;
    const float cD = 32.0f * u;
    const float k2thr = 2.0f * (1.0f + thr);
    const float N1 = fmaf(e.A1, d0.x, fmaf(e.B1, d0.y, e.C1));
    const float D1 = fmaf(e.A1, d0.z, e.B1 * d0.w);
    const float N2 = fmaf(e.A2, d0.x, fmaf(e.B2, d0.y, e.C2));
    const float D2 = fmaf(e.A2, d0.z, e.B2 * d0.w);
    const float r1 = rcp_approx_ftz(D1), r2 = rcp_approx_ftz(D2);
    const float s1 = -N1 * r1, s2 = -N2 * r2;
    const float lo = fminf(s1, s2), hi = fmaxf(s1, s2);
    const float Dd = fmaf(fmaxf(fabsf(lo), fabsf(hi)), cD, e.cN) * fmaxf(fabsf(r1), fabsf(r2));
    const float L = d1.x, slo = d1.y, shi = d1.z, g = d1.w;
    const bool rej_bounds = (hi - Dd > shi) | (lo + Dd < slo);
    const float inner = fminf(hi, L) - fmaxf(lo, 0.0f);
    const float outer = fmaxf(hi, L) - fminf(lo, 0.0f);
    // |error(inner)|, |error(outer)| <= E (two parameters with error Dd each, float rounding of the differences,
    // descriptor roundings g: the terms of K1's guard G)
    const float E = fmaf(2.0f, Dd, fmaf(outer, 1.0e-6f, g));
    const float margin = fmaf(-thr, outer, inner);
    const bool rej_thr = margin < -fmaf(Dd, k2thr, fmaf(outer, 1.0e-6f, g));
    // the reference's overlap is fl32(inner / fl32(outer)) of the true values (relative 1.3e-7) and may pick another
    // outer pair when two float distances tie (<= 6e-8): 5e-7 absolute slack
    U = pinf;
    Lb = ninf;
    const float den_lo = outer - E, den_hi = outer + E;
    if (den_lo > 0.0f) {
        U = fmaf((inner + E) * rcp_approx_ftz(den_lo), 1.0f + 2.0e-6f, 5.0e-7f);
        Lb = fmaf((inner - E) * rcp_approx_ftz(den_hi), 1.0f - 2.0e-6f, -5.0e-7f);
    }
    // a CERTAIN match: positive depths; both intersections inside the image by more than the error; |a.z| = L n |D|
    // far above 1e-12; some point strictly inside the other segment (pointOnSegment, src/line3D.cc:1274-1280, true
    // with a margin); outer distance above one pixel; overlap above the threshold.  Every comparison is false for
    // NaN, so NaN never certifies.
    const float m = Dd + g + 1.0e-4f;
    const bool inside = ((lo > m) & (lo < L - m)) | ((hi > m) & (hi < L - m)) | ((lo < -m) & (hi > m)) |
                        ((lo < L - m) & (hi > L + m));
    const bool cert = vpos & (lo - Dd > slo + g) & (hi + Dd < shi - g) & (fabsf(D1) > 1.0e-4f) & (fabsf(D2) > 1.0e-4f) &
                      (L * e.nmin > 1.0e-5f) & inside & (L > 1.001f) & (Lb > thr);
    if (!cert) Lb = ninf;
    unk = !vpos;
    // certain failures: a depth certainly not positive, an intersection certainly outside the image, overlap
    // certainly not above the threshold (NaN fails none of these comparisons: kept)
    return !(vneg | rej_bounds | rej_thr | (U <= thr));
}

// ---- K2 front: one thread per source row (one CTA = the 256 rows of a K1 tile), the pair's target tables -- FP32
// triangulation tables and K1 descriptors, 80 B per segment -- staged in shared memory in chunks of K2F_TCH
// segments by two 1-D TMA bulk copies.  Gathering them from global memory costs one L1 tag cycle per lane and
// load (measured: 0.18 ms for the FP32 tests alone on C2); out of shared memory the same reads are bank conflicts
// only.  Every lane walks the set bits of its own row's mask words (ascending target order), certifies the
// candidate, and keeps a running kNN-th largest certain lower bound T in a register insertion network: a
// candidate whose upper bound is below T can be dropped at once (T only grows).  Survivors go to the row's
// slots (row, target | unknown-depth flag) and (U); a second pass drops the ones below the final T, packs the
// contenders to the front of the row and queues them for the exact kernel. ----
static constexpr int K2F_TCH = 512;
static constexpr int K2F_Q = 64;  // per-lane queue of chunk-local candidates (flushed when a lane passes 32)
struct K2FSmem {
    SegV32 v32[K2F_TCH];
    SegDesc desc[K2F_TCH];
    unsigned short q[K2F_Q][K2_ROWS];  // [slot][thread]: conflict-free for a lane walking its own queue
    uint64_t bar;
};

// KT: size of the insertion network (the smallest of 4 / 8 / 10 / 12 / 16 that holds kNN; 0 = no pruning)
template <int KT>
__global__ void __launch_bounds__(K2_ROWS, 3) k2_front_kernel(
    const PairDev* __restrict__ pairs, const K1Cta* __restrict__ ctas, const uint32_t* __restrict__ mask,
    const uint32_t* __restrict__ cand_off, const SegV32* __restrict__ v32, const SegDesc* __restrict__ desc,
    const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views, const RowEpi32* __restrict__ row_epi,
    const uint32_t* __restrict__ perm, float thr, int knn, uint2* __restrict__ cand_rc, float* __restrict__ cand_u,
    uint32_t* __restrict__ ncont, uint32_t* __restrict__ row_pair, float* __restrict__ row_T, uint32_t* __restrict__ work,
    uint32_t* __restrict__ ctr)
{
    extern __shared__ __align__(128) unsigned char k2f_raw[];
    K2FSmem& S = *reinterpret_cast<K2FSmem*>(k2f_raw);
    const K1Cta cta = ctas[blockIdx.x];
    const PairDev& P = pairs[cta.pair];
    const uint32_t n_src = P.n_src, n_tgt = P.n_tgt, words = P.words;
    const uint32_t tid = threadIdx.x;
    // thread = position rho in K1's sorted row order (the mask and the line records are in that order: coalesced,
    // and neighbouring lanes see similar candidate sets); r = the natural row everything else is indexed by
    const uint32_t rho = cta.tile * K2_ROWS + tid;
    const uint32_t lrho = P.row_base - P.batch_row0 + (rho < n_src ? rho : 0u);
    const uint32_t r = rho < n_src ? perm[lrho] : n_src;
    const uint32_t lane = tid & 31;
    const float ninf = __int_as_float(0xff800000), pinf = __int_as_float(0x7f800000);
    constexpr bool prune = KT > 0;
    constexpr int KN = KT > 0 ? KT : 1;

    if (tid == 0) {
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- row set-up ----
    uint32_t lrow = 0, base = 0, ncand = 0;
    RowF32 R;
    memset(&R, 0, sizeof(R));
    if (r < n_src) {
        lrow = P.row_base - P.batch_row0 + r;
        base = cand_off[lrow];
        ncand = cand_off[lrow + 1] - base;
    }
    if (ncand) {
        const ViewDev& vs = views[P.src_view];
        const D3 Ct = ld3(views[P.tgt_view].C);
        const SegV32 sv = v32[P.src_off + r];
        // the row's own plane against the target centre, in the exact sequence: its sign is known exactly
        const SegPlane plB = planes[P.src_off + r];
        const double numB = ds(plB.cn, dot3(ld3(plB.n), Ct));
        R.e = row_epi[lrho];
        R.rp1x = sv.r1x; R.rp1y = sv.r1y; R.rp1z = sv.r1z; R.rp2x = sv.r2x; R.rp2y = sv.r2y; R.rp2z = sv.r2z;
        R.nBx = sv.nx; R.nBy = sv.ny; R.nBz = sv.nz;
        R.Csx = (float)vs.C[0]; R.Csy = (float)vs.C[1]; R.Csz = (float)vs.C[2];
        R.CsL1 = fabsf(R.Csx) + fabsf(R.Csy) + fabsf(R.Csz);
        R.sB = numB > 1e-9 ? 1.0f : (numB < -1e-9 ? -1.0f : 0.0f);
        row_pair[lrow] = cta.pair;
    }
    const uint32_t* __restrict__ mrow = mask + P.mask_base + (rho < n_src ? rho : 0u);

    // sorted insertion network in registers; the first KT - kNN entries are +inf, so its last entry is the
    // kNN-th largest certain lower bound seen so far
    float top[KN];
#pragma unroll
    for (int q = 0; q < KN; ++q) top[q] = (q < KN - knn) ? pinf : ninf;
    float T = ninf;       // = top[KT - 1]
    uint32_t ns = 0;      // survivors written to the row's slots
    uint32_t nq = 0;      // candidates waiting in this lane's queue

    // Evaluate the queued candidates of all 32 lanes in lockstep: iteration k takes the k-th entry of every lane's
    // queue (lanes with shorter queues idle).  No lane runs ahead, so the warp never splits into single-lane paths
    // (the first version let every lane walk its own bits with early `continue`s: 13 active lanes per instruction).
    auto flush = [&](uint32_t cb) {
        uint32_t kmax = nq;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        for (uint32_t k = 0; k < kmax; ++k) {
            const bool act = k < nq;
            const uint32_t cl = act ? (uint32_t)S.q[k][tid] : 0u;
            const float4* tp = reinterpret_cast<const float4*>(&S.v32[cl]);
            const float4 t0 = tp[0], t1 = tp[1], t2 = tp[2];
            SegV32 Tt;
            Tt.nx = t0.x; Tt.ny = t0.y; Tt.nz = t0.z; Tt.cn = t0.w;
            Tt.r1x = t1.x; Tt.r1y = t1.y; Tt.r1z = t1.z; Tt.r2x = t1.w;
            Tt.r2y = t2.x; Tt.r2z = t2.y;
            const float4* dp = reinterpret_cast<const float4*>(&S.desc[cl]);
            const float4 d0 = dp[0], d1 = dp[1];
            float U, Lb;
            bool unk;
            bool keep = certify_candidate(R, Tt, d0, d1, thr, U, Lb, unk) & act;
            keep &= !(U < T);  // kNN certain matches already lie strictly above it
            if (prune && keep && Lb > T) {  // a certain match that raises T
                float v = Lb;
#pragma unroll
                for (int q = 0; q < KN; ++q) {
                    const float hi = fmaxf(top[q], v);
                    v = fminf(top[q], v);
                    top[q] = hi;
                }
                T = top[KN - 1];
            }
            if (keep) {
                cand_rc[base + ns] = make_uint2(lrow, (cb + cl) | (unk ? 0x80000000u : 0u));
                cand_u[base + ns] = U;
                ++ns;
            }
        }
        nq = 0;
        __syncwarp();
    };

    uint32_t chunk_no = 0;
    for (uint32_t cb = 0; cb < n_tgt; cb += K2F_TCH, ++chunk_no) {
        const uint32_t tcnt = min((uint32_t)K2F_TCH, n_tgt - cb);
        if (tid == 0) {
            mbar_expect_tx(&S.bar, tcnt * (uint32_t)(sizeof(SegV32) + sizeof(SegDesc)));
            tma_load_1d(S.v32, v32 + P.tgt_off + cb, tcnt * (uint32_t)sizeof(SegV32), &S.bar);
            tma_load_1d(S.desc, desc + P.tgt_off + cb, tcnt * (uint32_t)sizeof(SegDesc), &S.bar);
        }
        mbar_wait(&S.bar, chunk_no & 1u);

        // the chunk's mask words, the same word index on every lane (coalesced loads, ascending target order)
        const uint32_t w0 = cb >> 5, w_end = min(words, (cb + K2F_TCH) >> 5);
        // eight words are fetched at a time: one load per word with the walk of its bits hanging on it was a chain of
        // exposed global-memory latencies (ncu: 4.6 warps per issue waiting on the long scoreboard)
        for (uint32_t wg = w0; wg < w_end; wg += 8) {
            uint32_t wb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) wb[i] = (ncand && wg + i < w_end) ? mrow[(size_t)(wg + i) * n_src] : 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (wg + i >= w_end) break;  // warp-uniform
                uint32_t bits = wb[i];
                const uint32_t lbase = (wg + i - w0) << 5;
                while (bits) {
                    const uint32_t j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    S.q[nq++][tid] = (unsigned short)(lbase + j);
                }
                if (__any_sync(0xffffffffu, nq > 32u)) flush(cb);
            }
        }
        flush(cb);
        __syncthreads();  // everyone is done with the staged tables before the next chunk overwrites them
    }

    // ---- every survivor is queued; the exact kernel skips the ones below the row's FINAL T (the running T only
    // grew), so no second pass over the row is needed here ----
    if (r < n_src) {
        ncont[lrow] = ns;
        if (ncand) row_T[lrow] = T;
    }
    uint32_t total = 0;
    const uint32_t ex = warp_excl_scan(ns, lane, total);
    uint32_t wq = 0;
    if (lane == 0 && total) wq = atomicAdd(&ctr[0], total);
    wq = __shfl_sync(0xffffffffu, wq, 0) + ex;
    for (uint32_t jq = 0; jq < ns; ++jq) work[wq + jq] = base + jq;
}

// ---- K2b: the reference's double sequence for one contender per thread.  The queue holds every survivor of the
// front kernel; the ones below their row's FINAL T (kNN certain matches lie strictly above them) are dropped
// here, and the rest are re-packed through a small per-warp ring so that the double-precision work runs on
// full warps ----
__device__ __forceinline__ unsigned long long exact_contender(
    const PairDev* __restrict__ pairs, const float4* __restrict__ segs, const SegRays* __restrict__ rays,
    const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views, const uint32_t* __restrict__ row_pair,
    const uint2 rc, float thr, double W)
{
    const PairDev& P = pairs[row_pair[rc.x]];
    const uint32_t r = rc.x - (P.row_base - P.batch_row0), c = rc.y & 0x7fffffffu;
    if (rc.y >> 31) {
        // depth signs the FP32 test could not certify: the exact test of k2_row_kernel's phase V
        const D3 Cs = ld3(views[P.src_view].C), Ct = ld3(views[P.tgt_view].C);
        const SegRays sr = rays[P.src_off + r];
        const SegPlane plB = planes[P.src_off + r];
        const D3 nB = ld3(plB.n);
        const double numB = ds(plB.cn, dot3(nB, Ct));
        const SegRays tr = rays[P.tgt_off + c];
        const SegPlane plA = planes[P.tgt_off + c];
        const D3 nA = ld3(plA.n);
        const double a1 = dot3(ld3(sr.r1), nA), a2 = dot3(ld3(sr.r2), nA);
        const double b1 = dot3(ld3(tr.r1), nB), b2 = dot3(ld3(tr.r2), nB);
        if (fabs(a1) < L3D_EPS || fabs(a2) < L3D_EPS || fabs(b1) < L3D_EPS || fabs(b2) < L3D_EPS) return ~0ull;
        const double num = ds(plA.cn, dot3(nA, Cs));
        if (!(depth_positive(num, a1) && depth_positive(num, a2) && depth_positive(numB, b1) && depth_positive(numB, b2)))
            return ~0ull;
    }
    // src/line3D.cc:1113-1158
    const float4 sg = segs[P.src_off + r];
    const D3 p1 = d3((double)sg.x, (double)sg.y, 1.0), p2 = d3((double)sg.z, (double)sg.w, 1.0);
    const D3 e1 = mul33(P.F, p1), e2 = mul33(P.F, p2);
    const float4 tg = segs[P.tgt_off + c];
    const D3 q1 = d3((double)tg.x, (double)tg.y, 1.0), q2 = d3((double)tg.z, (double)tg.w, 1.0);
    const D3 l2 = cross3(q1, q2);
    const D3 a = cross3(l2, e1), b = cross3(l2, e2);
    if (fabs(a.z) > L3D_EPS && fabs(b.z) > L3D_EPS) {
        const double ax = dd(a.x, a.z), ay = dd(a.y, a.z), bx = dd(b.x, b.z), by = dd(b.y, b.z);
        if (!(ax < 0 || ax > W || ay < 0 || ay > W || bx < 0 || bx > W || by < 0 || by > W)) {
            const float score = mutual_overlap_xy(ax, ay, bx, by, q1.x, q1.y, q2.x, q2.y);
            if (score > thr) return ((unsigned long long)__float_as_uint(score) << 32) | c;
        }
    }
    return ~0ull;  // not a match
}

__global__ void __launch_bounds__(256, 4) k2b_exact_kernel(
    const PairDev* __restrict__ pairs, const float4* __restrict__ segs, const SegRays* __restrict__ rays,
    const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views, const uint32_t* __restrict__ work,
    const uint2* __restrict__ cand_rc, const float* __restrict__ cand_u, const uint32_t* __restrict__ row_pair,
    const float* __restrict__ row_T, const uint32_t* __restrict__ ctr, unsigned long long* __restrict__ stage, float thr,
    double W)
{
    __shared__ uint32_t ring[8][64];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* pend = ring[warp];
    uint32_t np = 0;  // pending slots of this warp (warp-uniform)
    const uint32_t n_work = ctr[0];
    const uint32_t stride = gridDim.x * blockDim.x;
    // warp-uniform trip count: every lane of a warp runs the same number of iterations
    const uint32_t first = blockIdx.x * blockDim.x + warp * 32;
    for (uint32_t i0 = first; i0 < n_work; i0 += stride) {
        const uint32_t i = i0 + lane;
        bool want = false;
        uint32_t slot = 0;
        if (i < n_work) {
            slot = work[i];
            want = !(cand_u[slot] < row_T[cand_rc[slot].x]);
            if (!want) stage[slot] = ~0ull;  // never popped
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, want);
        if (want) pend[np + __popc(bal & lt_mask)] = slot;
        np += __popc(bal);
        __syncwarp();
        if (np >= 32u) {
            const uint32_t s = pend[lane];
            stage[s] = exact_contender(pairs, segs, rays, planes, views, row_pair, cand_rc[s], thr, W);
            __syncwarp();
            const uint32_t rest = np - 32u;
            const uint32_t mv = lane < rest ? pend[32 + lane] : 0u;
            __syncwarp();
            if (lane < rest) pend[lane] = mv;
            np = rest;
            __syncwarp();
        }
    }
    if (lane < np) {
        const uint32_t s = pend[lane];
        stage[s] = exact_contender(pairs, segs, rays, planes, views, row_pair, cand_rc[s], thr, W);
    }
}

// ---- pop order of the contenders that are matches: std::priority_queue<Match, vector, Match_kNN>
// (include/commons.h:233-244, src/line3D.cc:1198-1206) pops distinct overlaps in descending order ----
__global__ void __launch_bounds__(K2_ROWS) k2_rank_kernel(
    const PairDev* __restrict__ pairs, const K1Cta* __restrict__ ctas, const uint32_t* __restrict__ cand_off,
    uint32_t* __restrict__ ncont, const unsigned long long* __restrict__ stage, unsigned long long* __restrict__ pop_key,
    uint32_t* __restrict__ fin_cnt, int knn, uint32_t* __restrict__ fwork, uint32_t* __restrict__ ctr,
    uint2* __restrict__ fb_rows)
{
    const K1Cta cta = ctas[blockIdx.x];
    const PairDev& P = pairs[cta.pair];
    const uint32_t r = cta.tile * K2_ROWS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool row_ok = r < P.n_src;
    uint32_t lrow = 0, base = 0, npop = 0;
    if (row_ok) {
        lrow = P.row_base - P.batch_row0 + r;
        base = cand_off[lrow];
        const uint32_t nc = ncont[lrow];
        bool fallback = nc > (uint32_t)K2_MAXC;
        if (nc && !fallback) {
            unsigned long long key[K2_MAXC];
            uint32_t nv = 0;
            for (uint32_t i = 0; i < nc; ++i) {
                const unsigned long long k = stage[base + i];
                if (k != ~0ull) key[nv++] = k;
            }
            if (knn <= 0) {  // every match, ascending target order
                npop = nv;
                for (uint32_t t = 0; t < nv; ++t) pop_key[base + t] = key[t];
            } else {
                npop = min((uint32_t)knn, nv);
                for (uint32_t i = 0; i < nv && !fallback; ++i) {
                    const float ov = key_overlap(key[i]);
                    uint32_t rank = 0, equal = 0;
                    for (uint32_t j = 0; j < nv; ++j) {
                        const float oj = key_overlap(key[j]);
                        rank += oj > ov ? 1u : 0u;
                        equal += oj == ov ? 1u : 0u;
                    }
                    if (rank < npop) {
                        // equal overlaps at a popped rank: the pop order is the heap's, which needs every match
                        if (equal > 1u) fallback = true;
                        pop_key[base + rank] = key[i];
                    }
                }
            }
        }
        if (fallback) {
            npop = 0;
            ncont[lrow] = K2_ROW_FALLBACK;
            fb_rows[atomicAdd(&ctr[1], 1u)] = make_uint2(cta.pair, r);
        } else {
            ncont[lrow] = npop;  // from here on: the number of popped matches
        }
        fin_cnt[lrow] = 0u;
    }
    uint32_t total = 0;
    const uint32_t ex = warp_excl_scan(npop, lane, total);
    uint32_t f0 = 0;
    if (lane == 0 && total) f0 = atomicAdd(&ctr[2], total);
    f0 = __shfl_sync(0xffffffffu, f0, 0) + ex;
    for (uint32_t t = 0; t < npop; ++t) fwork[f0 + t] = base + t;
}

// ---- depths (src/line3D.cc:1365-1390) and orientation test (checkMatchOrientation, src/line3D.cc:962-1014)
// of one popped match per thread; a match that fails the test keeps its slot, flagged, and is skipped by the
// compaction ----
__global__ void __launch_bounds__(256, 4) k2_finish_kernel(
    const PairDev* __restrict__ pairs, const SegRays* __restrict__ rays, const double* __restrict__ midray,
    const SegPlane* __restrict__ planes, const ViewDev* __restrict__ views, const uint32_t* __restrict__ fwork,
    const uint2* __restrict__ cand_rc, const uint32_t* __restrict__ row_pair, const unsigned long long* __restrict__ pop_key,
    const uint32_t* __restrict__ ctr, FwdRec* __restrict__ fin_rec, uint32_t* __restrict__ fin_cnt, int apply_orient)
{
    const uint32_t n = ctr[2];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t p = fwork[i];
        const uint32_t lrow = cand_rc[p].x;  // every slot of a row's region names the row
        const PairDev& P = pairs[row_pair[lrow]];
        const uint32_t r = lrow - (P.row_base - P.batch_row0);
        const unsigned long long key = pop_key[p];
        const uint32_t c = (uint32_t)(key & 0xffffffffu);
        const D3 Cs = ld3(views[P.src_view].C), Ct = ld3(views[P.tgt_view].C);
        const SegRays sr = rays[P.src_off + r];
        const D3 rp1 = ld3(sr.r1), rp2 = ld3(sr.r2);
        const SegPlane plB = planes[P.src_off + r];
        const D3 nB = ld3(plB.n);
        const double numB = ds(plB.cn, dot3(nB, Ct));
        const SegRays tr = rays[P.tgt_off + c];
        const SegPlane plA = planes[P.tgt_off + c];
        const D3 nA = ld3(plA.n);
        const double num = ds(plA.cn, dot3(nA, Cs));
        // n.ray(p1): the same products in the same order as ray(p1).n
        const double ds1 = dd(num, dot3(rp1, nA)), ds2 = dd(num, dot3(rp2, nA));
        const double dt1 = dd(numB, dot3(ld3(tr.r1), nB)), dt2 = dd(numB, dot3(ld3(tr.r2), nB));
        FwdRec rec;
        rec.c = c;
        rec.overlap = key_overlap(key);
        rec.d_p1 = (float)ds1;
        rec.d_p2 = (float)ds2;
        rec.d_q1 = (float)dt1;
        rec.d_q2 = (float)dt2;
        rec.score = 0.0f;
        rec.flags = 0u;
        bool keep = true;
        if (apply_orient) {
            // The test is acos(x) in (0.0982, 3.0434) with x = ray(mid) . dir, dir = (P2-P1)/|P2-P1|, i.e.
            // |x| < 0.99518...: when (ray(mid).(P2-P1))^2 < 0.9951^2 |P2-P1|^2 the exact x (relative error ~1e-15)
            // is inside by a margin of 8e-5 and no sqrt/div/acos is needed.
            const D3 rmid = ld3(midray + 3 * (size_t)(P.src_off + r));
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
        if (keep) atomicAdd(&fin_cnt[lrow], 1u);
        else rec.flags = 0x80000000u;
        fin_rec[p] = rec;
    }
}

// packs the rows' matches into the forward-match store.  ncont == NULL (row-kernel variant) or a row handed to the
// row kernel: the row's fin_cnt records are contiguous; otherwise the row holds ncont popped records of which the
// ones that failed the orientation test are flagged
__global__ void __launch_bounds__(256) k2_compact_kernel(const uint32_t* __restrict__ cand_off,
                                                         const uint32_t* __restrict__ fin_cnt,
                                                         const uint32_t* __restrict__ fin_off, uint32_t rec_base,
                                                         const FwdRec* __restrict__ fin_rec,
                                                         FwdRec* __restrict__ fwd_rec,
                                                         uint32_t* __restrict__ fwd_off_global, uint32_t n_rows,
                                                         const uint32_t* __restrict__ ncont)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t src = cand_off[row];
    const uint32_t dst = rec_base + fin_off[row];
    const uint32_t n = fin_cnt[row];
    fwd_off_global[row] = dst;
    const uint32_t npop = ncont ? ncont[row] : K2_ROW_FALLBACK;
    if (npop == K2_ROW_FALLBACK) {
        for (uint32_t i = 0; i < n; ++i) fwd_rec[dst + i] = fin_rec[src + i];
        return;
    }
    uint32_t o = 0;
    for (uint32_t t = 0; t < npop; ++t) {
        const FwdRec rec = fin_rec[src + t];
        if (!(rec.flags & 0x80000000u)) fwd_rec[dst + o++] = rec;
    }
}

// K2 over one batch.  Scratch, all sized by the batch's candidate count n_cand: heap (8 B per candidate: the exact
// kernel's keys), cand_rec (32 B per candidate, carved into the survivor list, their upper bounds, the popped keys
// and the work queue), fin_rec.  ncont / row_pair / fb_rows: one entry per batch row; ctr: four words.
// The row kernel re-uses heap and cand_rec for its rows after the flat kernels are done with them.
// variant 0 (L3D_K2_VARIANT=0): the literal row kernel over every row.  Returns the number of launches;
// *uses_ncont tells the compaction whether ncont is meaningful.
int launch_k2_exact(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, uint32_t n_rows, uint32_t n_cand,
                    const float4* segs, const SegRays* rays, const double* midray, const SegPlane* planes,
                    const SegV32* v32, const SegDesc* desc, const RowEpi32* row_epi, const uint32_t* perm, const uint32_t* iperm,
                    const ViewDev* views, const uint32_t* mask, const uint32_t* cand_off, unsigned long long* heap, FwdRec* cand_rec,
                    FwdRec* fin_rec, uint32_t* fin_cnt, uint32_t* ncont, uint32_t* ctr, uint2* fb_rows, uint32_t* row_pair, float* row_T, float thr, int knn,
                    int max_image_width, int apply_orient, int n_sm, int* uses_ncont, cudaStream_t st)
{
    *uses_ncont = 0;
    if (n_ctas == 0) return 0;
    (void)n_rows;
    const char* ev = getenv("L3D_K2_VARIANT");
    const int variant = ev ? atoi(ev) : 1;
    if (variant == 0) {
        k2_row_kernel<false><<<n_ctas * (K2_ROWS / K2_SUB), K2_WARPS * 32, 0, st>>>(
            pairs, ctas, mask, cand_off, segs, rays, midray, planes, views, heap, cand_rec, fin_rec, fin_cnt, thr,
            (double)max_image_width, knn, apply_orient, nullptr, nullptr, iperm);
        return 1;
    }
    *uses_ncont = 1;
    cudaMemsetAsync(ctr, 0, 4 * sizeof(uint32_t), st);
    // cand_rec (32 B per candidate) carved into flat arrays of n_cand elements
    unsigned char* scratch = reinterpret_cast<unsigned char*>(cand_rec);
    const size_t n8 = ((size_t)n_cand + 1) * 8;
    uint2* cand_rc = reinterpret_cast<uint2*>(scratch);
    unsigned long long* pop_key = reinterpret_cast<unsigned long long*>(scratch + n8);
    uint32_t* work = reinterpret_cast<uint32_t*>(scratch + 2 * n8);
    float* cand_u = reinterpret_cast<float*>(scratch + 2 * n8 + n8 / 2);
    // the front kernel with the smallest insertion network that holds kNN (0: kNN <= 0 or > 16, no pruning)
    auto front = [&](auto kernel) {
        // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2FSmem));
        kernel<<<n_ctas, K2_ROWS, sizeof(K2FSmem), st>>>(pairs, ctas, mask, cand_off, v32, desc, planes, views, row_epi, perm, thr, knn,
                                                          cand_rc, cand_u, ncont, row_pair, row_T, work, ctr);
    };
    if (knn <= 0 || knn > K2_TOPK) front(k2_front_kernel<0>);
    else if (knn <= 4) front(k2_front_kernel<4>);
    else if (knn <= 8) front(k2_front_kernel<8>);
    else if (knn <= 10) front(k2_front_kernel<10>);
    else if (knn <= 12) front(k2_front_kernel<12>);
    else front(k2_front_kernel<16>);
    // the number of contenders is known on the device only: a grid that fills the machine, striding over the queue
    k2b_exact_kernel<<<n_sm * 8, 256, 0, st>>>(pairs, segs, rays, planes, views, work, cand_rc, cand_u, row_pair, row_T, ctr, heap, thr,
                                               (double)max_image_width);
    k2_rank_kernel<<<n_ctas, K2_ROWS, 0, st>>>(pairs, ctas, cand_off, ncont, heap, pop_key, fin_cnt, knn, work, ctr, fb_rows);
    k2_finish_kernel<<<n_sm * 8, 256, 0, st>>>(pairs, rays, midray, planes, views, work, cand_rc, row_pair, pop_key, ctr,
                                               fin_rec, fin_cnt, apply_orient);
    // rows the fast path could not decide (ties among the popped matches, very long rows): usually none
    k2_row_kernel<true><<<n_sm, K2_WARPS * 32, 0, st>>>(pairs, ctas, mask, cand_off, segs, rays, midray, planes, views, heap,
                                                       cand_rec, fin_rec, fin_cnt, thr, (double)max_image_width, knn,
                                                       apply_orient, fb_rows, ctr, iperm);
    return 5;
}

int launch_k2_compact(const uint32_t* cand_off, const uint32_t* fin_cnt, const uint32_t* fin_off,
                      uint32_t rec_base, const FwdRec* fin_rec, FwdRec* fwd_rec, uint32_t* fwd_off_global,
                      uint32_t n_rows, const uint32_t* ncont, cudaStream_t st)
{
    if (n_rows == 0) return 0;
    k2_compact_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(cand_off, fin_cnt, fin_off, rec_base, fin_rec,
                                                             fwd_rec, fwd_off_global, n_rows, ncont);
    return 1;
}

}  // namespace l3d
