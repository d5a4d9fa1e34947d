// k3_stream.cu -- scoring for the incremental (key-frame stream) mode: the match lists of a view
// persist from one matchImages cycle to the next, views are added / deleted / re-posed in between
// (L3DPPing::Run, src/L3DPPing.cpp:98-236), and every cycle walks the current views in ascending
// camera id (Line3D::computeMatches, src/line3D.cc:846-930).  Exact TU.
//
// The walk is kept view by view (a view's list holds the inverse matches of the views before it),
// each step a short chain of kernels over the rows of ONE view:
//   st_inv_count / st_inv_scatter   inverse matches this view receives (storeInverseMatches,
//                                   src/line3D.cc:1986-2015): forward records of the new pairs whose
//                                   source was scored earlier in the cycle and whose score is > 0
//   st_row_count / st_fill          list of a row = persisted survivors of the last filter, then the
//                                   inverse matches in push order, then the new forward matches
//   st_geo                          3-D segment and regularisers of every entry at the CURRENT poses
//   st_score                        Line3D::scoringCPU (src/line3D.cc:1405-1562), both branches: a
//                                   fresh entry (score 0) is scored against all siblings, an entry
//                                   scored in an earlier cycle gains / loses the per-camera maxima
//                                   of the cameras added / deleted since
//   st_post                         updateMatch (src/line3D.cc:1016-1055), score write-back, view max
//   st_filter_count / st_filter_write   Line3D::filterMatches (src/line3D.cc:1911-1983)
// and, once per cycle, st_update_entries = Line3D::update_Matches_and_Estimated_position3D
// (src/line3D.cc:1857-1908): the best matches are re-triangulated with the current poses.
#include "internal.h"
#include "score_core.cuh"

namespace l3d {

#define L3D_EPS 1e-12
static constexpr uint32_t NOIDX = 0xffffffffu;
static constexpr uint32_t VF_ADD = 1u, VF_DEL = 2u;
static constexpr uint32_t LF_DEAD = 4u;  // ListRec.flags: target camera deleted this cycle
static constexpr uint32_t LF_DROP = 8u;  // ListRec.flags: failed this cycle's orientation test

__device__ __forceinline__ D3 ld3s(const double* p) { return D3{p[0], p[1], p[2]}; }

// One step of the walk = the rows of a set of views that do not feed each other: all views scored in
// an earlier cycle go in ONE step (they receive no inverse matches any more, so they are independent),
// then every new view is a step of its own, in ascending camera id (it receives the inverse matches of
// the views before it).
struct StreamStep {
    uint32_t n;                  // rows of the step
    uint32_t n_in, in_total;     // incoming pair descriptors / their forward records (single-view steps only)
    uint32_t w_base, f_base;     // bases of the step's regions in the working / filtered arenas
    uint32_t w_cap, f_cap;
    const uint32_t* row_g;       // global segment index of every step row (views ascending, segments ascending)
    const uint32_t* seg_view;
    const ViewDev* views;
    const StreamPair* pairs;     // all descriptors of the cycle
    const uint32_t* vout0;       // per view: first outgoing descriptor, and their number
    const uint32_t* vnout;
};

// ---- inverse matches received by the view ----
__global__ void __launch_bounds__(256) st_inv_kernel(StreamStep s, const StreamPair* __restrict__ in,
                                                     const FwdRec* __restrict__ fwd_rec,
                                                     const float* __restrict__ fwd_score,
                                                     const uint32_t* __restrict__ I_off, uint32_t* __restrict__ I_cnt,
                                                     uint32_t* __restrict__ I_key, int scatter)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.in_total) return;
    uint32_t q = 0;
    while (q + 1 < s.n_in && t >= in[q].rec_cnt) {
        t -= in[q].rec_cnt;
        ++q;
    }
    const uint32_t f = in[q].rec_start + t;
    if (!(fwd_score[f] > 0.0f)) return;
    const uint32_t c = fwd_rec[f].c;
    const uint32_t slot = atomicAdd(&I_cnt[c], 1u);
    if (scatter) I_key[I_off[c] + slot] = f;  // sorted per row by st_fill: f grows with (source order, row, position)
}

// one thread per row.  Line3D::checkMatchOrientation (src/line3D.cc:962-1014) runs every cycle and, as
// the reference stores the copy taken BEFORE it sets match_orientation_, a forward match is tested
// again with the current pose of its source view each time (inverse matches carry the flag and are
// not); the ones that fail now are dropped before scoring.
__global__ void __launch_bounds__(128) st_row_count_kernel(StreamStep s,
                                                           const uint32_t* __restrict__ filt_off,
                                                           const uint32_t* __restrict__ filt_cnt,
                                                           ListRec* __restrict__ filt_old,
                                                           const ViewDev* __restrict__ views,
                                                           const SegRays* __restrict__ rays,
                                                           const double* __restrict__ midray,
                                                           const uint32_t* __restrict__ I_cnt,
                                                           const uint32_t* __restrict__ fwd_cnt,
                                                           uint32_t* __restrict__ W_cnt)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s.n) return;
    const uint32_t g = s.row_g[r];
    const uint32_t v = s.seg_view[g];
    const uint32_t seg = g - views[v].seg_off;
    const StreamPair* out = s.pairs + s.vout0[v];
    const uint32_t n_out = s.vnout[v];
    const uint32_t pn = filt_cnt[g], pb = filt_off[g];
    uint32_t m = s.n_in ? I_cnt[r] : 0u;
    if (pn) {
        const D3 C = ld3s(views[v].C);
        const SegRays sr = rays[g];
        const D3 r1 = ld3s(sr.r1), r2 = ld3s(sr.r2), rmid = ld3s(midray + 3 * (size_t)g);
        for (uint32_t e = 0; e < pn; ++e) {
            ListRec& L = filt_old[pb + e];
            bool keep = true;
            if (!(L.flags & 1u)) {
                const D3 P1 = add3(C, scale3(r1, (double)L.d_p1));
                const D3 P2 = add3(C, scale3(r2, (double)L.d_p2));
                const float len = (float)norm3(sub3(P1, P2));
                D3 dir = d3(0.0, 0.0, 0.0);
                if (len > L3D_EPS) dir = normalized3(sub3(P2, P1));
                const double ang = det_acos(fmin(fmax(dot3(rmid, dir), -1.0), 1.0));
                keep = ang > (double)0.098174771f && ang < (double)3.043417886f;
            }
            if (keep) ++m;
            else L.flags |= LF_DROP;
        }
    }
    for (uint32_t q = 0; q < n_out; ++q) m += fwd_cnt[out[q].row_base + seg];
    W_cnt[r] = m;
}

// one thread per row: the list in reference order
__global__ void __launch_bounds__(128) st_fill_kernel(StreamStep s, const StreamPair* __restrict__ in,
                                                      const uint32_t* __restrict__ filt_off,
                                                      const uint32_t* __restrict__ filt_cnt,
                                                      const ListRec* __restrict__ filt_old,
                                                      const uint32_t* __restrict__ I_off,
                                                      const uint32_t* __restrict__ I_cnt, uint32_t* __restrict__ I_key,
                                                      const uint32_t* __restrict__ fwd_off,
                                                      const uint32_t* __restrict__ fwd_cnt,
                                                      const FwdRec* __restrict__ fwd_rec,
                                                      const uint32_t* __restrict__ W_off, ListRec* __restrict__ W_rec,
                                                      uint32_t* __restrict__ W_row, uint32_t* __restrict__ L_off,
                                                      uint32_t* __restrict__ L_cnt, uint32_t* __restrict__ err)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s.n) return;
    const uint32_t g = s.row_g[r];
    const uint32_t v = s.seg_view[g];
    const uint32_t seg = g - s.views[v].seg_off;
    const StreamPair* out = s.pairs + s.vout0[v];
    const uint32_t n_out = s.vnout[v];
    const uint32_t lo = W_off[r], m = W_off[r + 1] - lo;
    L_off[g] = s.w_base + lo;
    L_cnt[g] = m;
    if (!m) return;
    if ((uint64_t)lo + m > s.w_cap) {
        atomicOr(err, 1u);
        L_cnt[g] = 0;
        return;
    }
    ListRec* dst = W_rec + s.w_base + lo;
    uint32_t* drow = W_row + s.w_base + lo;
    uint32_t w = 0;
    // survivors of the previous cycle's filter, in their order
    const uint32_t pn = filt_cnt[g], pb = filt_off[g];
    for (uint32_t e = 0; e < pn; ++e) {
        ListRec L = filt_old[pb + e];
        if (L.flags & LF_DROP) continue;
        L.src_idx = NOIDX;
        dst[w] = L;
        drow[w++] = r;
    }
    // inverse matches: ascending forward-record index = push order
    const uint32_t in_n = s.n_in ? I_cnt[r] : 0u, ib = s.n_in ? I_off[r] : 0u;
    for (uint32_t a = 1; a < in_n; ++a) {
        const uint32_t key = I_key[ib + a];
        uint32_t b = a;
        while (b > 0 && I_key[ib + b - 1] > key) {
            I_key[ib + b] = I_key[ib + b - 1];
            --b;
        }
        I_key[ib + b] = key;
    }
    for (uint32_t a = 0; a < in_n; ++a) {
        const uint32_t f = I_key[ib + a];
        uint32_t q = 0;
        while (q + 1 < s.n_in && f >= in[q].rec_start + in[q].rec_cnt) ++q;
        // source row of the record: last pair row whose first record is <= f
        uint32_t a0 = 0, a1 = in[q].n_src;
        while (a1 - a0 > 1) {
            const uint32_t mid = (a0 + a1) >> 1;
            if (fwd_off[in[q].row_base + mid] <= f) a0 = mid; else a1 = mid;
        }
        const FwdRec R = fwd_rec[f];
        ListRec L;
        L.tgt_view = in[q].other;
        L.tgt_seg = a0;
        L.overlap = R.overlap;
        L.score = 0.0f;
        L.d_p1 = R.d_q1; L.d_p2 = R.d_q2; L.d_q1 = R.d_p1; L.d_q2 = R.d_p2;
        L.flags = 3u;
        L.src_idx = NOIDX;
        dst[w] = L;
        drow[w++] = r;
    }
    // new forward matches, targets ascending, kNN pop order
    for (uint32_t q = 0; q < n_out; ++q) {
        const uint32_t row = out[q].row_base + seg;
        const uint32_t fb = fwd_off[row], fn = fwd_cnt[row];
        for (uint32_t e = 0; e < fn; ++e) {
            const FwdRec R = fwd_rec[fb + e];
            ListRec L;
            L.tgt_view = out[q].other;
            L.tgt_seg = R.c;
            L.overlap = R.overlap;
            L.score = 0.0f;
            L.d_p1 = R.d_p1; L.d_p2 = R.d_p2; L.d_q1 = R.d_q1; L.d_q2 = R.d_q2;
            L.flags = 0u;  // passed K2's orientation test now; tested again next cycle (see st_row_count_kernel)
            L.src_idx = fb + e;
            dst[w] = L;
            drow[w++] = r;
        }
    }
}

// one thread per list entry: View::unprojectSegment + the regularisers of scoringCPU at the current poses
__global__ void __launch_bounds__(128) st_geo_kernel(StreamStep s, const uint32_t* __restrict__ W_off,
                                                     const ViewDev* __restrict__ views,
                                                     const SegRays* __restrict__ rays,
                                                     const ListRec* __restrict__ W_rec,
                                                     const uint32_t* __restrict__ W_row, ListGeo* __restrict__ W_geo)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = min(W_off[s.n], s.w_cap);
    if (e >= total) return;
    const ListRec L = W_rec[s.w_base + e];
    const uint32_t r = W_row[s.w_base + e];
    const uint32_t g = s.row_g[r];
    const ViewDev& va = views[s.seg_view[g]];
    const ViewDev& vo = views[L.tgt_view];
    const SegRays sr = rays[g];
    const D3 C = ld3s(va.C), Co = ld3s(vo.C);
    // View::unprojectSegment (src/view.cc:385-400) + Segment3D ctor (include/segment3D.h:58-77)
    D3 P1 = add3(C, scale3(ld3s(sr.r1), (double)L.d_p1));
    D3 P2 = add3(C, scale3(ld3s(sr.r2), (double)L.d_p2));
    float len = (float)norm3(sub3(P1, P2));
    D3 dir = d3(0.0, 0.0, 0.0);
    if (len > L3D_EPS) {
        dir = normalized3(sub3(P2, P1));
    } else {
        P1 = d3(0.0, 0.0, 0.0);
        P2 = d3(0.0, 0.0, 0.0);
        len = 0.0f;
    }
    // src/line3D.cc:1429-1438, src/view.cc:474-477
    const float sig1 = fm(L.d_p1, va.k), sig2 = fm(L.d_p2, va.k);
    float reg1 = fm(fm(2.0f, sig1), sig1);
    float reg2 = fm(fm(2.0f, sig2), sig2);
    const float s1t = (float)dm(norm3(sub3(P1, Co)), (double)vo.k);
    const float s2t = (float)dm(norm3(sub3(P2, Co)), (double)vo.k);
    reg1 = fm(0.5f, fa(reg1, fm(fm(2.0f, s1t), s1t)));
    reg2 = fm(0.5f, fa(reg2, fm(fm(2.0f, s2t), s2t)));
    ListGeo G;
    G.dir[0] = dir.x; G.dir[1] = dir.y; G.dir[2] = dir.z;
    G.reg1 = reg1;
    G.reg2 = reg2;
    G.length = len;
    // a new target-camera run starts here
    G.run = (e == W_off[r] || W_rec[s.w_base + e - 1].tgt_view != L.tgt_view) ? 1u : 0u;
    G.pad0 = G.pad1 = 0u;
    W_geo[s.w_base + e] = G;
}

struct StreamStats {
    unsigned long long sim_evals, scored;
    uint32_t num_valid, err;
    uint32_t filtered, pad;
};

// one warp per row; lanes own matches M, siblings are walked in list order
__global__ void __launch_bounds__(256) st_score_kernel(StreamStep s, const uint32_t* __restrict__ W_off,
                                                       const ViewDev* __restrict__ views,
                                                       const unsigned char* __restrict__ vflag,
                                                       ListRec* __restrict__ W_rec, const ListGeo* __restrict__ W_geo,
                                                       float two_sigA_sqr, float dotcut,
                                                       StreamStats* __restrict__ stats)
{
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r >= s.n) return;
    const uint32_t lo = W_off[r];
    const uint32_t m = W_off[r + 1] - lo;
    if (m == 0 || (uint64_t)lo + m > s.w_cap) return;
    ListRec* rec = W_rec + s.w_base + lo;
    const ListGeo* geo = W_geo + s.w_base + lo;
    const float min_sim = 0.5f, xcut = -0.70f, pcut = 0.5f;
    unsigned long long evals = 0;
    bool valid = false;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t me = base + lane;
        if (me >= m) continue;
        const ListRec M = rec[me];
        if (vflag[M.tgt_view] & VF_DEL) continue;  // src/line3D.cc:1424-1426
        const ListGeo G = geo[me];
        const D3 dirM = d3(G.dir[0], G.dir[1], G.dir[2]);
        const bool Mvalid = !(G.length < L3D_EPS);
        auto sim_to = [&](uint32_t j) {
            const ListRec M2 = rec[j];
            const ListGeo G2 = geo[j];
            Sib s2;
            s2.d_p1 = M2.d_p1;
            s2.d_p2 = M2.d_p2;
            s2.cam = M2.tgt_view;
            s2.flags = (G2.length < L3D_EPS) ? 0u : 2u;
            ++evals;
            return sim_for_scoring(M.d_p1, M.d_p2, G.reg1, G.reg2, Mvalid, dirM, s2, G2.dir, two_sigA_sqr, min_sim,
                                   xcut, pcut, dotcut);
        };
        float score = M.score;
        if (score != 0.0f) {
            // scored in an earlier cycle (src/line3D.cc:1439-1512): per-camera maxima of the cameras added
            // since are added, those of the cameras being deleted are subtracted, each set in ascending
            // camera id (std::map order)
            for (uint32_t pass = 0; pass < 2; ++pass) {
                const uint32_t bit = pass == 0 ? VF_ADD : VF_DEL;
                long long last = -1;
                for (;;) {
                    long long best = 0x7fffffffffffffffll;
                    uint32_t best_view = NOIDX;
                    for (uint32_t j = 0; j < m; ++j) {
                        const uint32_t tv = rec[j].tgt_view;
                        if (tv == M.tgt_view || !(vflag[tv] & bit)) continue;
                        const long long cam = (long long)views[tv].cam_id;
                        if (cam > last && cam < best) {
                            best = cam;
                            best_view = tv;
                        }
                    }
                    if (best_view == NOIDX) break;
                    bool first = true;
                    float mx = 0.0f;
                    for (uint32_t j = 0; j < m; ++j) {
                        if (rec[j].tgt_view != best_view) continue;
                        const float sim = sim_to(j);
                        if (first || sim > mx) mx = sim;
                        first = false;
                    }
                    score = pass == 0 ? fa(score, mx) : fs(score, mx);
                    last = best;
                }
            }
        } else {
            // new match (src/line3D.cc:1513-1547)
            float stored = 0.0f;
            bool in_run = false;
            for (uint32_t j = 0; j < m; ++j) {
                const uint32_t tv = rec[j].tgt_view;
                if (geo[j].run) in_run = false;
                if (tv == M.tgt_view || (vflag[tv] & VF_DEL)) continue;
                const float sim = sim_to(j);
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
        }
        rec[me].score = score;
        valid |= score > 0.75f;
    }
    valid = __any_sync(0xffffffffu, valid);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, d);
    if (lane == 0) {
        atomicAdd(&stats->sim_evals, evals);
        if (valid) atomicAdd(&stats->num_valid, 1u);
    }
}

// one thread per entry: updateMatch (entries to cameras deleted this cycle die), forward-score
// write-back (read by the later views' st_inv_kernel), maximum score of the view
__global__ void __launch_bounds__(256) st_post_kernel(StreamStep s, const uint32_t* __restrict__ W_off,
                                                      const unsigned char* __restrict__ vflag,
                                                      ListRec* __restrict__ W_rec, const uint32_t* __restrict__ W_row,
                                                      float* __restrict__ fwd_score,
                                                      uint32_t* __restrict__ view_max, StreamStats* __restrict__ stats)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = min(W_off[s.n], s.w_cap);
    bool live = false;
    uint32_t v = NOIDX, key = 0u;
    if (e < total) {
        ListRec& L = W_rec[s.w_base + e];
        if (vflag[L.tgt_view] & VF_DEL) {
            L.flags |= LF_DEAD;
        } else {
            if (L.src_idx != NOIDX) fwd_score[L.src_idx] = L.score;
            live = true;
            v = s.seg_view[s.row_g[W_row[s.w_base + e]]];
            key = float_ordered(L.score);
        }
    }
    // one atomic per (warp, view): the entries of a view are contiguous
    const uint32_t grp = __match_any_sync(0xffffffffu, v);
    const uint32_t mx = __reduce_max_sync(grp, key);
    const uint32_t lane = threadIdx.x & 31;
    if (live && lane == (uint32_t)(__ffs(grp) - 1)) atomicMax(&view_max[v], mx);
    const uint32_t n_live = __popc(__ballot_sync(0xffffffffu, live));
    if (lane == 0 && n_live) atomicAdd(&stats->scored, (unsigned long long)n_live);
}

__global__ void __launch_bounds__(128) st_filter_count_kernel(StreamStep s, const uint32_t* __restrict__ W_off,
                                                              const ListRec* __restrict__ W_rec,
                                                              const uint32_t* __restrict__ view_max,
                                                              uint32_t* __restrict__ F_cnt, uint32_t* __restrict__ best_e)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s.n) return;
    const uint32_t lo = W_off[r];
    uint32_t m = W_off[r + 1] - lo;
    if ((uint64_t)lo + m > s.w_cap) m = 0;
    const float max_score = fmaxf(0.0f, ordered_to_float(view_max[s.seg_view[s.row_g[r]]]));
    const float lim = fm(0.10f, max_score);
    const ListRec* rec = W_rec + s.w_base + lo;
    uint32_t kept = 0, bi = NOIDX;
    float best = 0.0f;
    for (uint32_t e = 0; e < m; ++e) {
        const ListRec L = rec[e];
        if (L.flags & LF_DEAD) continue;
        if (L.score > 0.0f && L.score > lim) {
            ++kept;
            if (L.score > best) {  // first strict maximum in list order
                best = L.score;
                bi = e;
            }
        }
    }
    F_cnt[r] = kept;
    best_e[r] = (bi != NOIDX && best > 0.75f) ? bi : NOIDX;
}

__global__ void __launch_bounds__(128) st_filter_write_kernel(StreamStep s, const uint32_t* __restrict__ W_off,
                                                              const ListRec* __restrict__ W_rec,
                                                              const uint32_t* __restrict__ view_max,
                                                              const uint32_t* __restrict__ F_off,
                                                              const uint32_t* __restrict__ best_e,
                                                              const ViewDev* __restrict__ views,
                                                              const SegRays* __restrict__ rays,
                                                              ListRec* __restrict__ filt_new,
                                                              uint32_t* __restrict__ filt_off,
                                                              uint32_t* __restrict__ filt_cnt,
                                                              EntryDev* __restrict__ entries,
                                                              uint32_t* __restrict__ view_total,
                                                              StreamStats* __restrict__ stats)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= s.n) return;
    const uint32_t g = s.row_g[r];
    const uint32_t v = s.seg_view[g];
    const uint32_t lo = W_off[r];
    uint32_t m = W_off[r + 1] - lo;
    if ((uint64_t)lo + m > s.w_cap) m = 0;
    const float max_score = fmaxf(0.0f, ordered_to_float(view_max[v]));
    const float lim = fm(0.10f, max_score);
    const ListRec* rec = W_rec + s.w_base + lo;
    const uint32_t fo = F_off[r], kept = F_off[r + 1] - fo;
    const bool fits = (uint64_t)fo + kept <= s.f_cap;
    if (!fits) atomicOr(&stats->err, 2u);
    uint32_t w = 0;
    if (fits && kept)
        for (uint32_t e = 0; e < m; ++e) {
            const ListRec L = rec[e];
            if (L.flags & LF_DEAD) continue;
            if (L.score > 0.0f && L.score > lim) filt_new[s.f_base + fo + w++] = L;
        }
    filt_off[g] = s.f_base + fo;
    filt_cnt[g] = fits ? kept : 0u;
    if (kept) {
        atomicAdd(&view_total[v], kept);
        atomicAdd(&stats->filtered, kept);
    }
    EntryDev& E = entries[g];
    const uint32_t bi = best_e[r];
    if (bi == NOIDX) {
        E.has = 0u;
        return;
    }
    // unprojectMatch(best,true) (src/line3D.cc:1826-1838)
    const ListRec B = rec[bi];
    const SegRays sr = rays[g];
    const D3 Ca = ld3s(views[v].C);
    D3 P1 = add3(Ca, scale3(ld3s(sr.r1), (double)B.d_p1));
    D3 P2 = add3(Ca, scale3(ld3s(sr.r2), (double)B.d_p2));
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
}

// Line3D::update_Matches_and_Estimated_position3D (src/line3D.cc:1857-1908): every best match is
// triangulated again (Line3D::triangulationDepths, src/line3D.cc:1365-1390) with the current poses;
// the entry is dropped unless all four depths are positive.  The plane normals and endpoint rays
// are K0's per-segment tables (same arithmetic as the matching stage).
__global__ void __launch_bounds__(128) st_update_entries_kernel(uint32_t S, const uint32_t* __restrict__ seg_view,
                                                                const ViewDev* __restrict__ views,
                                                                const SegRays* __restrict__ rays,
                                                                const SegPlane* __restrict__ planes,
                                                                EntryDev* __restrict__ entries)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= S) return;
    EntryDev& E = entries[g];
    if (!E.has) return;
    const ViewDev& vs = views[seg_view[g]];
    const ViewDev& vt = views[E.tgt_view];
    const uint32_t t = vt.seg_off + E.tgt_seg;
    const SegRays rs = rays[g], rt = rays[t];
    const SegPlane ps = planes[g], pt = planes[t];
    const D3 Cs = ld3s(vs.C), Ct = ld3s(vt.C);
    const D3 nt = ld3s(pt.n), ns = ld3s(ps.n);
    const D3 s1 = ld3s(rs.r1), s2 = ld3s(rs.r2), t1 = ld3s(rt.r1), t2 = ld3s(rt.r2);
    double ds1 = -1.0, ds2 = -1.0, dt1 = -1.0, dt2 = -1.0;
    {
        const double a = dot3(s1, nt), b = dot3(s2, nt);
        if (!(fabs(a) < L3D_EPS || fabs(b) < L3D_EPS)) {
            const double num = ds(pt.cn, dot3(nt, Cs));
            ds1 = dd(num, a);
            ds2 = dd(num, b);
        }
    }
    {
        const double a = dot3(t1, ns), b = dot3(t2, ns);
        if (!(fabs(a) < L3D_EPS || fabs(b) < L3D_EPS)) {
            const double num = ds(ps.cn, dot3(ns, Ct));
            dt1 = dd(num, a);
            dt2 = dd(num, b);
        }
    }
    if (!(ds1 > L3D_EPS && ds2 > L3D_EPS && dt1 > L3D_EPS && dt2 > L3D_EPS)) {
        E.has = 0u;
        return;
    }
    E.d_p1 = (float)ds1; E.d_p2 = (float)ds2; E.d_q1 = (float)dt1; E.d_q2 = (float)dt2;
    D3 P1 = add3(Cs, scale3(s1, (double)E.d_p1));
    D3 P2 = add3(Cs, scale3(s2, (double)E.d_p2));
    float len = (float)norm3(sub3(P1, P2));
    D3 dir = d3(0.0, 0.0, 0.0);
    if (len > L3D_EPS) dir = normalized3(sub3(P2, P1));
    else { P1 = d3(0, 0, 0); P2 = d3(0, 0, 0); len = 0.0f; }
    E.P1[0] = P1.x; E.P1[1] = P1.y; E.P1[2] = P1.z;
    E.P2[0] = P2.x; E.P2[1] = P2.y; E.P2[2] = P2.z;
    E.dir[0] = dir.x; E.dir[1] = dir.y; E.dir[2] = dir.z;
    E.length = len;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
size_t stream_stats_bytes() { return sizeof(StreamStats); }

int launch_stream_step(const StreamStepArgs& a, cudaStream_t st)
{
    StreamStep s;
    s.n = a.n; s.n_in = a.n_in; s.in_total = a.in_total;
    s.w_base = a.w_base; s.f_base = a.f_base; s.w_cap = a.w_cap; s.f_cap = a.f_cap;
    s.row_g = a.row_g; s.seg_view = a.seg_view; s.views = a.views;
    s.pairs = (const StreamPair*)a.pairs; s.vout0 = a.vout0; s.vnout = a.vnout;
    const StreamPair* in = (const StreamPair*)a.pairs + a.in0;
    StreamStats* stats = (StreamStats*)a.stats;
    const uint32_t n = a.n;
    if (!n) return 0;
    int launches = 0;
    if (a.n_in) {  // inverse matches received (a new view)
        cudaMemsetAsync(a.I_cnt, 0, ((size_t)n + 1) * 4, st);
        if (a.in_total) {
            st_inv_kernel<<<(a.in_total + 255) / 256, 256, 0, st>>>(s, in, a.fwd_rec, a.fwd_score, a.I_off, a.I_cnt,
                                                                     a.I_key, 0);
            ++launches;
        }
        launches += launch_scan_u32(a.I_cnt, a.I_off, n, a.scan, a.scan_words, st);
        if (a.in_total) {
            cudaMemsetAsync(a.I_fill, 0, ((size_t)n + 1) * 4, st);
            st_inv_kernel<<<(a.in_total + 255) / 256, 256, 0, st>>>(s, in, a.fwd_rec, a.fwd_score, a.I_off, a.I_fill,
                                                                     a.I_key, 1);
            ++launches;
        }
    }
    st_row_count_kernel<<<(n + 127) / 128, 128, 0, st>>>(s, a.filt_off, a.filt_cnt, a.filt_old, a.views, a.rays,
                                                          a.midray, a.I_cnt, a.fwd_cnt, a.W_cnt);
    ++launches;
    launches += launch_scan_u32(a.W_cnt, a.W_off, n, a.scan, a.scan_words, st);
    st_fill_kernel<<<(n + 127) / 128, 128, 0, st>>>(s, in, a.filt_off, a.filt_cnt, a.filt_old, a.I_off, a.I_cnt,
                                                     a.I_key, a.fwd_off, a.fwd_cnt, a.fwd_rec, a.W_off, a.W_rec,
                                                     a.W_row, a.L_off, a.L_cnt, &stats->err);
    ++launches;
    if (a.w_cap) {
        st_geo_kernel<<<(a.w_cap + 127) / 128, 128, 0, st>>>(s, a.W_off, a.views, a.rays, a.W_rec, a.W_row, a.W_geo);
        st_score_kernel<<<(n + 7) / 8, 256, 0, st>>>(s, a.W_off, a.views, a.vflag, a.W_rec, a.W_geo, a.two_sigA_sqr,
                                                      score_dotcut(a.two_sigA_sqr, 0.5f), stats);
        st_post_kernel<<<(a.w_cap + 255) / 256, 256, 0, st>>>(s, a.W_off, a.vflag, a.W_rec, a.W_row, a.fwd_score,
                                                               a.view_max, stats);
        launches += 3;
    }
    st_filter_count_kernel<<<(n + 127) / 128, 128, 0, st>>>(s, a.W_off, a.W_rec, a.view_max, a.F_cnt, a.best_e);
    ++launches;
    launches += launch_scan_u32(a.F_cnt, a.F_off, n, a.scan, a.scan_words, st);
    st_filter_write_kernel<<<(n + 127) / 128, 128, 0, st>>>(s, a.W_off, a.W_rec, a.view_max, a.F_off, a.best_e, a.views,
                                                             a.rays, a.filt_new, a.filt_off, a.filt_cnt, a.entries,
                                                             a.view_total, stats);
    ++launches;
    return launches;
}

int launch_stream_update_entries(uint32_t S, const uint32_t* seg_view, const ViewDev* views, const SegRays* rays,
                                 const SegPlane* planes, EntryDev* entries, cudaStream_t st)
{
    if (!S) return 0;
    st_update_entries_kernel<<<(S + 127) / 128, 128, 0, st>>>(S, seg_view, views, rays, planes, entries);
    return 1;
}

}  // namespace l3d
