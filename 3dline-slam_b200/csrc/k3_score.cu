// k3_score.cu -- K3: per-view list assembly, multi-view scoring, inverse-match propagation and
// filtering (exact TU).  One call sequence per view, in ascending camera-ID order, because
// Line3D::storeInverseMatches (src/line3D.cc:1986-2015) makes the lists of later views depend on
// the scores of earlier ones.
//
//   assemble: list of segment i of view A = [inverse matches from earlier views, in the order
//             they were appended] ++ [forward matches per target camera ascending, each in
//             priority-queue pop order]           (src/line3D.cc:846-930)
//   score   : Line3D::scoringCPU new-match branch (src/line3D.cc:1513-1547) with
//             Line3D::similarityForScoring (src/line3D.cc:1685-1716); one lane per match M walks
//             its siblings in list order, folding the per-camera running maximum into score3D_.
//             Entries of one target camera are contiguous in every list the reference can build
//             (whole per-camera blocks are appended), so the std::map<camID,float> of the
//             reference reduces to "current run" state.
//   inverse : storeInverseMatches -> per-pair CSR over target segments, entries ordered by the
//             forward-record index (= source row ascending, list order inside a row).
//   filter  : Line3D::filterMatches (src/line3D.cc:1911-1983): keep score>0 && >0.1*max, first
//             strict maximum = best, best>0.75 -> estimated_position3D_ row.
#include "detmath.cuh"
#include "exact.cuh"
#include "internal.h"

namespace l3d {

#define L3D_EPS 1e-12
static constexpr uint32_t NOIDX = 0xffffffffu;

__device__ __forceinline__ D3 ld3g(const double* p) { return D3{p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------
// assemble: counts
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_count_kernel(const IncDev* __restrict__ inc, uint32_t n_inc,
                                                       const PairDev* __restrict__ pairs,
                                                       const uint32_t* __restrict__ fwd_cnt,
                                                       const uint32_t* __restrict__ inv_cnt, uint32_t n_rows,
                                                       uint32_t* __restrict__ L_cnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint32_t m = 0;
    for (uint32_t q = 0; q < n_inc; ++q) {
        const PairDev& P = pairs[inc[q].pair];
        if (inc[q].inverse)
            m += inv_cnt[P.tgt_base + i];
        else
            m += fwd_cnt[P.row_base + i];
    }
    L_cnt[i] = m;
}

// ------------------------------------------------------------------------------------------
// assemble: gather records + per-entry geometry (one thread per row, list order)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k3_gather_kernel(
    uint32_t view, const IncDev* __restrict__ inc, uint32_t n_inc, const PairDev* __restrict__ pairs,
    const ViewDev* __restrict__ views, const SegRays* __restrict__ rays, const uint32_t* __restrict__ fwd_off,
    const uint32_t* __restrict__ fwd_cnt, const FwdRec* __restrict__ fwd_rec, const uint32_t* __restrict__ inv_off,
    const uint32_t* __restrict__ inv_cnt, const uint2* __restrict__ inv_ent, const uint32_t* __restrict__ L_off, ListRec* __restrict__ L_rec,
    ListGeo* __restrict__ L_geo, uint32_t* __restrict__ err_flag)
{
    const ViewDev& va = views[view];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= va.n_seg) return;
    const SegRays sr = rays[va.seg_off + i];
    const D3 r1 = ld3g(sr.r1), r2 = ld3g(sr.r2);
    const D3 Ca = ld3g(va.C);
    const float k = va.k;
    uint32_t pos = L_off[i];
    uint32_t prev_cam = NOIDX;
    const uint32_t first = pos;

    for (uint32_t q = 0; q < n_inc; ++q) {
        const uint32_t pi = inc[q].pair;
        const PairDev& P = pairs[pi];
        const bool inv = inc[q].inverse != 0;
        uint32_t b, n, other;
        if (inv) {
            b = inv_off[P.tgt_base + i];
            n = inv_cnt[P.tgt_base + i];
            other = P.src_view;
        } else {
            b = fwd_off[P.row_base + i];
            n = fwd_cnt[P.row_base + i];
            other = P.tgt_view;
        }
        if (n == 0) continue;
        const ViewDev& vo = views[other];
        const D3 Co = ld3g(vo.C);
        const float ko = vo.k;
        for (uint32_t e = 0; e < n; ++e) {
            ListRec L;
            if (inv) {
                const uint2 ie = inv_ent[b + e];  // x: forward record index, y: source row in the other view
                const FwdRec f = fwd_rec[ie.x];
                L.tgt_view = other;
                L.tgt_seg = ie.y;
                L.overlap = f.overlap;
                L.score = 0.0f;
                L.d_p1 = f.d_q1;
                L.d_p2 = f.d_q2;
                L.d_q1 = f.d_p1;
                L.d_q2 = f.d_p2;
                L.flags = 3u;
                L.src_idx = NOIDX;
            } else {
                const FwdRec f = fwd_rec[b + e];
                L.tgt_view = other;
                L.tgt_seg = f.c;
                L.overlap = f.overlap;
                L.score = 0.0f;
                L.d_p1 = f.d_p1;
                L.d_p2 = f.d_p2;
                L.d_q1 = f.d_q1;
                L.d_q2 = f.d_q2;
                L.flags = 0u;
                L.src_idx = b + e;
            }
            // M3D = View::unprojectSegment (src/view.cc:385-400)
            D3 P1 = add3(Ca, scale3(r1, (double)L.d_p1));
            D3 P2 = add3(Ca, scale3(r2, (double)L.d_p2));
            float len = (float)norm3(sub3(P1, P2));
            D3 dir = d3(0.0, 0.0, 0.0);
            if (len > L3D_EPS) {
                dir = normalized3(sub3(P2, P1));
            } else {
                P1 = d3(0.0, 0.0, 0.0);
                P2 = d3(0.0, 0.0, 0.0);
                len = 0.0f;
            }
            // regularisers (src/line3D.cc:1429-1438, src/view.cc:474-477)
            const float sig1 = fm(L.d_p1, k), sig2 = fm(L.d_p2, k);
            float reg1 = fm(fm(2.0f, sig1), sig1);
            float reg2 = fm(fm(2.0f, sig2), sig2);
            const float s1t = (float)dm(norm3(sub3(P1, Co)), (double)ko);
            const float s2t = (float)dm(norm3(sub3(P2, Co)), (double)ko);
            reg1 = fm(0.5f, fa(reg1, fm(fm(2.0f, s1t), s1t)));
            reg2 = fm(0.5f, fa(reg2, fm(fm(2.0f, s2t), s2t)));
            ListGeo G;
            G.dir[0] = dir.x; G.dir[1] = dir.y; G.dir[2] = dir.z;
            G.reg1 = reg1;
            G.reg2 = reg2;
            G.length = len;
            G.run = (other != prev_cam) ? 1u : 0u;
            G.pad0 = pi;
            G.pad1 = 0u;
            if (G.run) {
                // a camera must not re-appear after another one (never happens in the reference flow)
                for (uint32_t z = first; z < pos; ++z)
                    if (L_rec[z].tgt_view == other) atomicExch(err_flag, 1u);
            }
            prev_cam = other;
            L_rec[pos] = L;
            L_geo[pos] = G;
            ++pos;
        }
    }
}

// ------------------------------------------------------------------------------------------
// score: one warp per row
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_ordered(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct ScoreStats {
    unsigned long long sim_evals;
    unsigned long long scored;
    uint32_t max_score_ord;  // ordered-uint encoding of the float maximum
    uint32_t num_valid;
};

__global__ void __launch_bounds__(256) k3_score_kernel(uint32_t n_rows, const uint32_t* __restrict__ L_off,
                                                       ListRec* __restrict__ L_rec,
                                                       const ListGeo* __restrict__ L_geo,
                                                       FwdRec* __restrict__ fwd_rec, float two_sigA_sqr,
                                                       float min_sim, ScoreStats* __restrict__ stats)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    const uint32_t b = L_off[warp];
    const uint32_t m = L_off[warp + 1] - b;
    if (m == 0) return;
    float wmax = 0.0f;
    bool any_valid = false;
    unsigned long long evals = 0;
    // exp(x) <= 0.4966 for x < -0.70, so with min_sim >= 0.5 the result would be truncated to 0
    // anyway; for smaller thresholds the early exits are disabled
    const float xcut = (min_sim >= 0.5f) ? -0.70f : -3.0e38f;
    const float pcut = (min_sim >= 0.5f) ? 0.5f : -1.0f;

    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t me = base + lane;
        const bool act = me < m;
        ListRec M;
        ListGeo G;
        if (act) {
            M = L_rec[b + me];
            G = L_geo[b + me];
        } else {
            M.tgt_view = NOIDX; M.d_p1 = M.d_p2 = 0.0f;
            G.dir[0] = G.dir[1] = G.dir[2] = 0.0; G.reg1 = G.reg2 = 1.0f; G.length = 0.0f;
        }
        const D3 dirM = d3(G.dir[0], G.dir[1], G.dir[2]);
        float score = 0.0f;  // new matches start at score3D_ = 0 (src/line3D.cc:1175)
        float stored = 0.0f;
        bool in_run = false;  // a map entry for the current run's camera exists
        for (uint32_t j = 0; j < m; ++j) {
            // uniform (broadcast) loads of the sibling
            const ListRec M2 = L_rec[b + j];
            const ListGeo G2 = L_geo[b + j];
            if (G2.run) in_run = false;
            if (!act || M2.tgt_view == M.tgt_view) continue;
            ++evals;
            // similarityForScoring (src/line3D.cc:1685-1716)
            float sim = 0.0f;
            if (!(G.length < L3D_EPS || G2.length < L3D_EPS)) {
                const float d1 = fs(M.d_p1, M2.d_p1);
                const float d2 = fs(M.d_p2, M2.d_p2);
                const float x1 = fd(fm(-d1, d1), G.reg1);
                const float x2 = fd(fm(-d2, d2), G.reg2);
                if (!(x1 < xcut || x2 < xcut)) {
                    const float sim_p = fminf(det_expf(x1), det_expf(x2));
                    if (!(sim_p <= pcut)) {
                        // angleBetweenSeg3D (src/line3D.cc:1841-1853)
                        const float dot_p = (float)dot3(dirM, d3(G2.dir[0], G2.dir[1], G2.dir[2]));
                        float angle = (float)dm(dd((double)det_acosf(fmaxf(fminf(dot_p, 1.0f), -1.0f)), L3D_PI),
                                                (double)180.0f);
                        if (angle > 90.0f) angle = fs(180.0f, angle);
                        const float sim_a = det_expf(fd(fm(-angle, angle), two_sigA_sqr));
                        const float s = fminf(sim_a, sim_p);
                        sim = (s > min_sim) ? s : 0.0f;
                    }
                }
            }
            // per-camera running maximum folded into the score (src/line3D.cc:1527-1540)
            if (in_run) {
                if (sim > stored) {
                    score = fs(score, stored);
                    score = fa(score, sim);
                    stored = sim;
                }
            } else {
                score = fa(score, sim);
                stored = sim;
                in_run = true;
            }
        }
        if (act) {
            L_rec[b + me].score = score;
            if (M.src_idx != NOIDX && fwd_rec) fwd_rec[M.src_idx].score = score;
            wmax = fmaxf(wmax, score);
            any_valid |= (score > 0.75f);
        }
    }
    // warp-level reductions (max is exact and order-free)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
        evals += __shfl_xor_sync(0xffffffffu, evals, d);
    }
    const bool valid = __any_sync(0xffffffffu, any_valid);
    if (lane == 0) {
        atomicMax(&stats->max_score_ord, float_ordered(wmax));
        atomicAdd(&stats->sim_evals, evals);
        atomicAdd(&stats->scored, (unsigned long long)m);
        if (valid) atomicAdd(&stats->num_valid, 1u);
    }
}

// ------------------------------------------------------------------------------------------
// inverse matches: count / fill / sort
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_inv_count_kernel(uint32_t n_rows, const uint32_t* __restrict__ L_off,
                                                           const ListRec* __restrict__ L_rec,
                                                           const ListGeo* __restrict__ L_geo,
                                                           const PairDev* __restrict__ pairs,
                                                           uint32_t* __restrict__ inv_cnt)
{
    const uint32_t total = L_off[n_rows];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const ListRec L = L_rec[e];
        if (L.src_idx == NOIDX || !(L.score > 0.0f)) continue;
        const PairDev& P = pairs[L_geo[e].pad0];
        if (!P.emit_inverse) continue;
        atomicAdd(&inv_cnt[P.tgt_base + L.tgt_seg], 1u);
    }
}

__global__ void __launch_bounds__(256) k3_inv_offsets_kernel(const uint32_t* __restrict__ scan, uint32_t first_row,
                                                             uint32_t n, uint32_t rec_base,
                                                             uint32_t* __restrict__ inv_off)
{
    // start offsets only: the row after the range belongs to the next view (counts live in inv_cnt)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv_off[first_row + i] = rec_base + scan[i];
}

__global__ void __launch_bounds__(256) k3_inv_fill_kernel(uint32_t n_rows, const uint32_t* __restrict__ L_off,
                                                          const ListRec* __restrict__ L_rec,
                                                          const ListGeo* __restrict__ L_geo,
                                                          const PairDev* __restrict__ pairs,
                                                          const uint32_t* __restrict__ inv_off,
                                                          uint32_t* __restrict__ inv_fill, uint2* __restrict__ inv_ent)
{
    const uint32_t total = L_off[n_rows];
    // row of entry e: binary search in L_off
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const ListRec L = L_rec[e];
        if (L.src_idx == NOIDX || !(L.score > 0.0f)) continue;
        const PairDev& P = pairs[L_geo[e].pad0];
        if (!P.emit_inverse) continue;
        uint32_t lo = 0, hi = n_rows;  // largest row with L_off[row] <= e
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (L_off[mid] <= e) lo = mid; else hi = mid;
        }
        const uint32_t tr = P.tgt_base + L.tgt_seg;
        const uint32_t slot = atomicAdd(&inv_fill[tr], 1u);
        inv_ent[inv_off[tr] + slot] = make_uint2(L.src_idx, lo);
    }
}

__global__ void __launch_bounds__(256) k3_inv_sort_kernel(uint32_t first_row, uint32_t n,
                                                          const uint32_t* __restrict__ inv_off,
                                                          const uint32_t* __restrict__ inv_cnt,
                                                          uint2* __restrict__ inv_ent)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = inv_off[first_row + i];
    const uint32_t cnt = inv_cnt[first_row + i];
    for (uint32_t a = 1; a < cnt; ++a) {  // insertion sort by forward-record index
        const uint2 v = inv_ent[b + a];
        uint32_t z = a;
        while (z > 0 && inv_ent[b + z - 1].x > v.x) {
            inv_ent[b + z] = inv_ent[b + z - 1];
            --z;
        }
        inv_ent[b + z] = v;
    }
}

// ------------------------------------------------------------------------------------------
// filter: count kept + best (one warp per row), then write
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ordered_to_float(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(256) k3_filter_count_kernel(
    uint32_t view, const ViewDev* __restrict__ views, const SegRays* __restrict__ rays,
    const uint32_t* __restrict__ L_off, const ListRec* __restrict__ L_rec, const ScoreStats* __restrict__ stats,
    uint32_t* __restrict__ F_cnt, EntryDev* __restrict__ entries)
{
    const ViewDev& va = views[view];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= va.n_seg) return;
    const float max_score = fmaxf(0.0f, ordered_to_float(stats->max_score_ord));
    const float lim = fm(0.10f, max_score);
    const uint32_t b = L_off[warp];
    const uint32_t m = L_off[warp + 1] - b;
    uint32_t kept = 0;
    float best = 0.0f;
    uint32_t best_idx = NOIDX;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t e = base + lane;
        float s = 0.0f;
        bool keep = false;
        if (e < m) {
            s = L_rec[b + e].score;
            keep = (s > 0.0f) && (s > lim);
        }
        kept += __popc(__ballot_sync(0xffffffffu, keep));
        // first strict maximum in list order
        float cs = keep ? s : 0.0f;
        uint32_t ci = keep ? e : NOIDX;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, cs, d);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, ci, d);
            if (os > cs || (os == cs && oi < ci)) { cs = os; ci = oi; }
        }
        if (ci != NOIDX && cs > best) { best = cs; best_idx = ci; }
    }
    if (lane == 0) {
        F_cnt[warp] = kept;
        EntryDev E;
        E.has = 0u;
        if (best_idx != NOIDX && best > 0.75f) {
            const ListRec B = L_rec[b + best_idx];
            const SegRays sr = rays[va.seg_off + warp];
            const D3 Ca = ld3g(va.C);
            D3 P1 = add3(Ca, scale3(ld3g(sr.r1), (double)B.d_p1));
            D3 P2 = add3(Ca, scale3(ld3g(sr.r2), (double)B.d_p2));
            float len = (float)norm3(sub3(P1, P2));
            D3 dir = d3(0.0, 0.0, 0.0);
            if (len > L3D_EPS) dir = normalized3(sub3(P2, P1));
            else { P1 = d3(0, 0, 0); P2 = d3(0, 0, 0); len = 0.0f; }
            E.P1[0] = P1.x; E.P1[1] = P1.y; E.P1[2] = P1.z;
            E.P2[0] = P2.x; E.P2[1] = P2.y; E.P2[2] = P2.z;
            E.dir[0] = dir.x; E.dir[1] = dir.y; E.dir[2] = dir.z;
            E.length = len;
            E.tgt_view = B.tgt_view;
            E.tgt_seg = B.tgt_seg;
            E.overlap = B.overlap;
            E.score = B.score;
            E.d_p1 = B.d_p1; E.d_p2 = B.d_p2; E.d_q1 = B.d_q1; E.d_q2 = B.d_q2;
            E.has = 1u;
            entries[va.seg_off + warp] = E;
        } else {
            entries[va.seg_off + warp].has = 0u;
        }
    }
}

// filt_off_global[seg_off + i] = filt_base + F_off[i]; records compacted in list order
__global__ void __launch_bounds__(256) k3_filter_write_kernel(
    uint32_t view, const ViewDev* __restrict__ views, const uint32_t* __restrict__ L_off,
    const ListRec* __restrict__ L_rec, const ScoreStats* __restrict__ stats, const uint32_t* __restrict__ F_off,
    const uint32_t* __restrict__ filt_total_in, ListRec* __restrict__ filt_rec, uint32_t* __restrict__ filt_off,
    uint32_t* __restrict__ filt_cnt)
{
    const ViewDev& va = views[view];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= va.n_seg) return;
    const float max_score = fmaxf(0.0f, ordered_to_float(stats->max_score_ord));
    const float lim = fm(0.10f, max_score);
    const uint32_t b = L_off[warp];
    const uint32_t m = L_off[warp + 1] - b;
    const uint32_t dst0 = *filt_total_in + F_off[warp];
    uint32_t w = 0;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t e = base + lane;
        bool keep = false;
        ListRec L;
        if (e < m) {
            L = L_rec[b + e];
            keep = (L.score > 0.0f) && (L.score > lim);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep) filt_rec[dst0 + w + __popc(bal & ((1u << lane) - 1u))] = L;
        w += __popc(bal);
    }
    if (lane == 0) {
        filt_off[va.seg_off + warp] = dst0;
        filt_cnt[va.seg_off + warp] = w;
    }
}

__global__ void k3_advance_total_kernel(uint32_t* __restrict__ filt_total, const uint32_t* __restrict__ F_off,
                                        uint32_t n_rows, ScoreStats* __restrict__ stats,
                                        ScoreStats* __restrict__ accum)
{
    *filt_total += F_off[n_rows];
    // fold the per-view stats into the accumulators and reset for the next view
    accum->sim_evals += stats->sim_evals;
    accum->scored += stats->scored;
    accum->num_valid += stats->num_valid;
    stats->sim_evals = 0;
    stats->scored = 0;
    stats->num_valid = 0;
    stats->max_score_ord = 0;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_k3_count(const IncDev* inc, uint32_t n_inc, const PairDev* pairs, const uint32_t* fwd_cnt,
                    const uint32_t* inv_cnt, uint32_t n_rows, uint32_t* L_cnt, cudaStream_t st)
{
    k3_count_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(inc, n_inc, pairs, fwd_cnt, inv_cnt, n_rows, L_cnt);
    return 1;
}

int launch_k3_gather(uint32_t view, uint32_t n_rows, const IncDev* inc, uint32_t n_inc, const PairDev* pairs,
                     const ViewDev* views, const SegRays* rays, const uint32_t* fwd_off, const uint32_t* fwd_cnt,
                     const FwdRec* fwd_rec, const uint32_t* inv_off, const uint32_t* inv_cnt, const uint2* inv_ent,
                     const uint32_t* L_off, ListRec* L_rec, ListGeo* L_geo, uint32_t* err_flag, cudaStream_t st)
{
    k3_gather_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(view, inc, n_inc, pairs, views, rays, fwd_off, fwd_cnt,
                                                            fwd_rec, inv_off, inv_cnt, inv_ent, L_off, L_rec, L_geo, err_flag);
    return 1;
}

int launch_k3_score(uint32_t n_rows, const uint32_t* L_off, ListRec* L_rec, const ListGeo* L_geo, FwdRec* fwd_rec,
                    float two_sigA_sqr, float min_sim, void* stats, cudaStream_t st)
{
    if (!n_rows) return 0;
    const uint32_t warps_per_block = 8;
    k3_score_kernel<<<(n_rows + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(
        n_rows, L_off, L_rec, L_geo, fwd_rec, two_sigA_sqr, min_sim, (ScoreStats*)stats);
    return 1;
}

int launch_k3_inverse(uint32_t n_rows, const uint32_t* L_off, const ListRec* L_rec, const ListGeo* L_geo,
                      const PairDev* pairs, uint32_t first_tgt_row, uint32_t n_tgt_rows, uint32_t rec_base,
                      uint32_t* inv_cnt, uint32_t* inv_fill, uint32_t* inv_off, uint2* inv_ent, uint32_t* scan_tmp,
                      uint32_t* scan_scratch, size_t scan_scratch_words_, uint32_t grid, cudaStream_t st)
{
    if (n_tgt_rows == 0) return 0;
    int launches = 0;
    k3_inv_count_kernel<<<grid, 256, 0, st>>>(n_rows, L_off, L_rec, L_geo, pairs, inv_cnt);
    ++launches;
    launches += launch_scan_u32(inv_cnt + first_tgt_row, scan_tmp, n_tgt_rows, scan_scratch, scan_scratch_words_, st);
    k3_inv_offsets_kernel<<<(n_tgt_rows + 255) / 256, 256, 0, st>>>(scan_tmp, first_tgt_row, n_tgt_rows, rec_base,
                                                                         inv_off);
    ++launches;
    k3_inv_fill_kernel<<<grid, 256, 0, st>>>(n_rows, L_off, L_rec, L_geo, pairs, inv_off, inv_fill, inv_ent);
    ++launches;
    k3_inv_sort_kernel<<<(n_tgt_rows + 255) / 256, 256, 0, st>>>(first_tgt_row, n_tgt_rows, inv_off, inv_cnt, inv_ent);
    ++launches;
    return launches;
}

int launch_k3_filter(uint32_t view, uint32_t n_rows, const ViewDev* views, const SegRays* rays,
                     const uint32_t* L_off, const ListRec* L_rec, void* stats, void* accum, uint32_t* F_cnt,
                     uint32_t* F_off, EntryDev* entries, uint32_t* filt_total, ListRec* filt_rec, uint32_t* filt_off,
                     uint32_t* filt_cnt, uint32_t* scan_scratch, size_t scan_scratch_words_, cudaStream_t st)
{
    int launches = 0;
    const uint32_t blocks = (n_rows + 7) / 8;
    k3_filter_count_kernel<<<blocks, 256, 0, st>>>(view, views, rays, L_off, L_rec, (const ScoreStats*)stats, F_cnt,
                                                    entries);
    ++launches;
    launches += launch_scan_u32(F_cnt, F_off, n_rows, scan_scratch, scan_scratch_words_, st);
    k3_filter_write_kernel<<<blocks, 256, 0, st>>>(view, views, L_off, L_rec, (const ScoreStats*)stats, F_off,
                                                    filt_total, filt_rec, filt_off, filt_cnt);
    ++launches;
    k3_advance_total_kernel<<<1, 1, 0, st>>>(filt_total, F_off, n_rows, (ScoreStats*)stats, (ScoreStats*)accum);
    ++launches;
    return launches;
}

size_t k3_stats_bytes() { return sizeof(ScoreStats); }

}  // namespace l3d
