// k3_wavefront.cu -- K3: the whole scoring wavefront of Line3D::computeMatches
// (src/line3D.cc:846-930) as ONE persistent cooperative kernel (exact TU).
//
// Views are processed in ascending camera-ID order because Line3D::storeInverseMatches
// (src/line3D.cc:1986-2015) feeds the scored matches of a view into the lists of its not yet
// processed neighbours.  Phase k of the kernel does, separated from phase k+1 by one grid sync:
//   task A (view k), one CTA per source segment i, all in one pass over the row:
//     assemble  list = [inverse matches from earlier views, in append order] ++ [forward matches
//               per target camera ascending, in priority-queue pop order]  (SURVEY.md App. A.6);
//               inverse entries of one pair are ordered by forward-record index (= source row
//               ascending, list order inside the row) with a rank sort
//     geometry  View::unprojectSegment + regularisers (src/view.cc:385-400, src/line3D.cc:1426-1438)
//     score     Line3D::scoringCPU new-match branch (src/line3D.cc:1513-1547): thread per match M
//               walks its siblings in list order from shared memory; entries of one target camera
//               are contiguous (whole per-camera blocks are appended), so the reference's
//               std::map<camID,float> reduces to "current run" state
//     inverse   storeInverseMatches: score>0 forward matches are appended (atomic slot) to the
//               pre-sized CSR slot of their target segment
//   task B (view k-1): Line3D::filterMatches (src/line3D.cc:1911-1983), one warp per segment:
//               keep score>0 && >0.1*max, first strict maximum = best, best>0.75 ->
//               estimated_position3D_ row.  It needs the view-wide maximum, hence the one-phase lag.
// List storage uses offsets computed BEFORE the wavefront from upper bounds (every forward record
// of an inverse-emitting pair may become an inverse match), so no prefix sum is needed inside.
#include <cooperative_groups.h>

#include "internal.h"
#include "score_core.cuh"

namespace cg = cooperative_groups;

namespace l3d {

#define L3D_EPS 1e-12
static constexpr uint32_t NOIDX = 0xffffffffu;
static constexpr int WF_THREADS = 64;
static constexpr int WF_MAXM = 256;   // list entries staged in shared memory
static constexpr int WF_MASKW = WF_MAXM / 32;
static constexpr int WF_LCAP = 2048;  // flagged (M, sibling) pairs evaluated in the flat pass
static constexpr int WF_MAXINC = 64;  // incident pairs staged in shared memory

struct WfStats {
    unsigned long long sim_evals;
    unsigned long long scored;
    uint32_t num_valid;
    uint32_t filt_cursor;  // bump allocator of the filtered-record store
    uint32_t err;          // bit0: incident list too long, bit1: filtered store overflow
    uint32_t pad;
};

// per forward record: 3-D direction and regularisers of the match seen from its source view
// (G_fwd) and, for inverse-emitting pairs, seen from its target view (G_inv); computed before
// the wavefront so that no FP64 geometry sits on its critical path
struct GeoRec {
    double dir[3];
    float reg1, reg2;
    uint32_t valid;
    uint32_t pad;
};

struct WfArgs {
    const ViewDev* views;
    const PairDev* pairs;
    const IncDev* inc;
    const uint32_t* inc_off;  // [V+1]
    const SegRays* rays;
    const uint32_t* fwd_off;
    const uint32_t* fwd_cnt;
    FwdRec* fwd_rec;
    const GeoRec* G_fwd;
    const GeoRec* G_inv;
    const uint32_t* inv_off;  // start of the CSR slot of every (pair, tgt segment)
    uint32_t* inv_fill;       // entries appended so far
    uint2* inv_ent;           // x: forward record index, y: source row
    const uint32_t* L_off;    // [S+1] upper-bound offsets (global segment order)
    const uint64_t* L_base;   // [V] base of the view's list region inside L_rec/L_sib/L_dir
    uint32_t* L_cnt;          // [S] actual list lengths
    ListRec* L_rec;
    Sib* L_sib;
    double* L_dir;            // 3 doubles per entry
    float2* L_reg;            // regularisers of the entry as M
    uint32_t* view_max;       // [V] ordered-uint maximum score of the view
    ListRec* filt_rec;
    uint32_t filt_cap;
    uint32_t* filt_off;       // [S]
    uint32_t* filt_cnt;       // [S]
    EntryDev* entries;        // [S]
    WfStats* stats;
    uint32_t V;
    float two_sigA_sqr;
    float dotcut;  // see score_core.cuh
    uint32_t* dbg;  // optional per-row cycle counters (L3D_WF_DEBUG), 4 words per segment
};

__device__ __forceinline__ D3 ld3w(const double* p) { return D3{p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------
// task A: one CTA assembles, scores and propagates one row
// ------------------------------------------------------------------------------------------
struct WfSmem {
    Sib sib[WF_MAXM];
    double dir[3 * WF_MAXM];
    float2 reg[WF_MAXM];
    unsigned char runid[WF_MAXM];       // camera block (incident pair slot) of every entry
    uint32_t mask[WF_MAXM * WF_MASKW];  // per match M: siblings that need the full similarity
    uint32_t cnt[WF_MAXM + 1];          // exclusive prefix of the per-M counts
    float sims[WF_LCAP];                // results of the flat slow-path pass, grouped by M, ascending j
    uint32_t blk_b[WF_MAXINC], blk_n[WF_MAXINC], blk_pos[WF_MAXINC + 1];
    uint32_t blk_other[WF_MAXINC];      // the other view of the pair
    uint32_t blk_flags[WF_MAXINC];      // bit0: inverse block, bit1: pair emits inverse matches
    uint32_t blk_tbase[WF_MAXINC];      // tgt_base of the pair (inverse emission)
    uint32_t total;
};

// cheap certain reject of similarityForScoring: -d^2 < -0.75 reg  =>  -d^2/reg < -0.70  =>  sim_p < 0.4966
__device__ __forceinline__ bool sim_needs_full(const Sib& M, float thr1, float thr2, bool regs_ok, const Sib& S2)
{
    if (!(M.flags & 2u) || !(S2.flags & 2u)) return false;  // invalid 3-D segment: sim = 0
    const float d1 = fs(M.d_p1, S2.d_p1), d2 = fs(M.d_p2, S2.d_p2);
    const float n1 = fm(-d1, d1), n2 = fm(-d2, d2);
    return !(regs_ok && (n1 < thr1 || n2 < thr2));
}

__device__ void wf_score_row(const WfArgs& a, uint32_t v, uint32_t i, WfSmem& sm)
{
    const ViewDev& va = a.views[v];
    const uint32_t i0 = a.inc_off[v], n_inc = a.inc_off[v + 1] - i0;
    const uint32_t g = va.seg_off + i;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const long long t_start = a.dbg ? clock64() : 0;

    // block table (n_inc <= WF_MAXINC is checked on the host)
    if (tid < (int)n_inc) {
        const IncDev q = a.inc[i0 + tid];
        const PairDev& P = a.pairs[q.pair];
        if (q.inverse) {
            sm.blk_b[tid] = a.inv_off[P.tgt_base + i];
            sm.blk_n[tid] = a.inv_fill[P.tgt_base + i];
            sm.blk_other[tid] = P.src_view;
            sm.blk_flags[tid] = 1u;
        } else {
            sm.blk_b[tid] = a.fwd_off[P.row_base + i];
            sm.blk_n[tid] = a.fwd_cnt[P.row_base + i];
            sm.blk_other[tid] = P.tgt_view;
            sm.blk_flags[tid] = P.emit_inverse ? 2u : 0u;
        }
        sm.blk_tbase[tid] = P.tgt_base;
    }
    __syncthreads();
    if (tid < 32) {  // exclusive prefix of the block sizes (n_inc <= 64: two values per lane)
        const uint32_t n0 = (lane < n_inc) ? sm.blk_n[lane] : 0u;
        const uint32_t n1 = (lane + 32 < n_inc) ? sm.blk_n[lane + 32] : 0u;
        uint32_t x0 = n0, x1 = n1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t0 = __shfl_up_sync(0xffffffffu, x0, d), t1 = __shfl_up_sync(0xffffffffu, x1, d);
            if ((int)lane >= d) { x0 += t0; x1 += t1; }
        }
        const uint32_t tot0 = __shfl_sync(0xffffffffu, x0, 31);
        if (lane < n_inc) sm.blk_pos[lane] = x0 - n0;
        if (lane + 32 < n_inc) sm.blk_pos[lane + 32] = tot0 + x1 - n1;
        const uint32_t tot = tot0 + __shfl_sync(0xffffffffu, x1, 31);
        if (lane == 0) sm.blk_pos[n_inc] = tot;
    }
    __syncthreads();
    const uint32_t m = sm.blk_pos[n_inc];
    if (tid == 0) a.L_cnt[g] = m;
    if (m == 0) return;  // uniform

    const size_t lbase = (size_t)a.L_base[v] + (a.L_off[g] - a.L_off[va.seg_off]);
    ListRec* __restrict__ Lr = a.L_rec + lbase;
    Sib* __restrict__ Ls = a.L_sib + lbase;
    double* __restrict__ Ld = a.L_dir + 3 * lbase;
    float2* __restrict__ Lg = a.L_reg + lbase;
    const bool in_smem = m <= WF_MAXM;

    // ---- assemble: one thread per list entry (geometry comes from the pre-pass) ----
    for (uint32_t e = tid; e < m; e += WF_THREADS) {
        uint32_t q = 0;
        while (sm.blk_pos[q + 1] <= e) ++q;  // n_inc is small
        uint32_t j = e - sm.blk_pos[q];
        const uint32_t b = sm.blk_b[q], n = sm.blk_n[q];
        ListRec L;
        uint32_t dst = e;
        const GeoRec* gp;
        L.tgt_view = sm.blk_other[q];
        if (sm.blk_flags[q] & 1u) {
            const uint2 ie = a.inv_ent[b + j];
            // append order of the reference = ascending forward-record index: rank sort
            uint32_t rank = 0;
            for (uint32_t z = 0; z < n; ++z) rank += (a.inv_ent[b + z].x < ie.x) ? 1u : 0u;
            dst = sm.blk_pos[q] + rank;
            j = rank;
            const FwdRec f = a.fwd_rec[ie.x];
            gp = a.G_inv + ie.x;
            L.tgt_seg = ie.y;
            L.overlap = f.overlap;
            L.d_p1 = f.d_q1;
            L.d_p2 = f.d_q2;
            L.d_q1 = f.d_p1;
            L.d_q2 = f.d_p2;
            L.flags = 3u;
            L.src_idx = NOIDX;
        } else {
            const FwdRec f = a.fwd_rec[b + j];
            gp = a.G_fwd + (b + j);
            L.tgt_seg = f.c;
            L.overlap = f.overlap;
            L.d_p1 = f.d_p1;
            L.d_p2 = f.d_p2;
            L.d_q1 = f.d_q1;
            L.d_q2 = f.d_q2;
            L.flags = 0u;
            L.src_idx = b + j;
        }
        L.score = 0.0f;
        const GeoRec G = *gp;
        Sib sb;
        sb.d_p1 = L.d_p1;
        sb.d_p2 = L.d_p2;
        sb.cam = L.tgt_view;
        sb.flags = ((j == 0) ? 1u : 0u) | (G.valid ? 2u : 0u);
        Lr[dst] = L;
        if (in_smem) {
            sm.runid[dst] = (unsigned char)q;
            sm.sib[dst] = sb;
            sm.dir[3 * dst + 0] = G.dir[0];
            sm.dir[3 * dst + 1] = G.dir[1];
            sm.dir[3 * dst + 2] = G.dir[2];
            sm.reg[dst] = make_float2(G.reg1, G.reg2);
        } else {
            Ls[dst] = sb;
            Ld[3 * dst + 0] = G.dir[0];
            Ld[3 * dst + 1] = G.dir[1];
            Ld[3 * dst + 2] = G.dir[2];
            Lg[dst] = make_float2(G.reg1, G.reg2);
        }
    }
    __threadfence_block();
    __syncthreads();
    const long long t_gather = a.dbg ? clock64() : 0;

    const Sib* __restrict__ sib = in_smem ? sm.sib : Ls;
    const double* __restrict__ dirs = in_smem ? sm.dir : Ld;
    const float2* __restrict__ regs = in_smem ? sm.reg : Lg;

    // ---- pass 1 (rows that fit in shared memory): cheap test, bit mask + counts per match M ----
    bool flat = in_smem;
    if (flat) {
        const uint32_t words = (m + 31) >> 5;
        for (uint32_t e = tid; e < m; e += WF_THREADS) {
            const Sib M = sib[e];
            const float2 rg = regs[e];
            const bool regs_ok = rg.x > 0.0f && rg.y > 0.0f;
            const float thr1 = fm(-0.75f, rg.x), thr2 = fm(-0.75f, rg.y);
            uint32_t c = 0;
            const bool Mok = (M.flags & 2u) != 0;
            for (uint32_t w = 0; w < words; ++w) {
                uint32_t bits = 0;
                const uint32_t jend = min(32u, m - (w << 5));
                const Sib* __restrict__ sw = sib + (w << 5);
#pragma unroll 4
                for (uint32_t jj = 0; jj < jend; ++jj) {
                    const Sib S2 = sw[jj];
                    // branch-free form of sim_needs_full() && different camera
                    const float d1 = fs(M.d_p1, S2.d_p1), d2 = fs(M.d_p2, S2.d_p2);
                    const float n1 = fm(-d1, d1), n2 = fm(-d2, d2);
                    const bool rej = regs_ok & ((n1 < thr1) | (n2 < thr2));
                    const bool need = Mok & ((S2.flags & 2u) != 0) & (S2.cam != M.cam) & !rej;
                    bits |= (need ? 1u : 0u) << jj;
                }
                sm.mask[e * WF_MASKW + w] = bits;
                c += __popc(bits);
            }
            sm.cnt[e] = c;
        }
        __syncthreads();
        if (tid < 32) {  // exclusive prefix over m <= 256 counts: 8 per lane
            uint32_t loc[WF_MAXM / 32];
            uint32_t sum = 0;
#pragma unroll
            for (int z = 0; z < WF_MAXM / 32; ++z) {
                const uint32_t idx = lane * (WF_MAXM / 32) + z;
                loc[z] = (idx < m) ? sm.cnt[idx] : 0u;
                sum += loc[z];
            }
            uint32_t x = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, x, d);
                if ((int)lane >= d) x += t;
            }
            uint32_t run = x - sum;
#pragma unroll
            for (int z = 0; z < WF_MAXM / 32; ++z) {
                const uint32_t idx = lane * (WF_MAXM / 32) + z;
                if (idx < m) sm.cnt[idx] = run;
                run += loc[z];
            }
            if (lane == 31) sm.total = x;
        }
        __syncthreads();
        const uint32_t T = sm.total;
        if (tid == 0) sm.cnt[m] = T;
        flat = T <= WF_LCAP;  // uniform
        if (flat) {
            __syncthreads();
            // ---- pass 2: the flagged (M, j) pairs, one per thread: full similarity ----
            for (uint32_t t = tid; t < T; t += WF_THREADS) {
                uint32_t lo = 0, hi = m;  // largest e with cnt[e] <= t
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (sm.cnt[mid] <= t) lo = mid; else hi = mid;
                }
                const uint32_t e = lo;
                uint32_t k = t - sm.cnt[e];
                uint32_t w = 0, bits = sm.mask[e * WF_MASKW];
                while (k >= (uint32_t)__popc(bits)) {
                    k -= __popc(bits);
                    bits = sm.mask[e * WF_MASKW + (++w)];
                }
                const uint32_t j = (w << 5) + __fns(bits, 0, k + 1);
                const Sib M = sib[e];
                const float2 rg = regs[e];
                const D3 dirM = d3(dirs[3 * e], dirs[3 * e + 1], dirs[3 * e + 2]);
                sm.sims[t] = sim_for_scoring(M.d_p1, M.d_p2, rg.x, rg.y, true, dirM, sib[j], dirs + 3 * j, a.two_sigA_sqr,
                                             0.5f, -0.70f, 0.5f, a.dotcut);
            }
            __syncthreads();
        }
    }

    // ---- fold: one thread per match M, siblings in list order (src/line3D.cc:1515-1543) ----
    float wmax = 0.0f;
    uint32_t evals = 0;
    bool any_valid = false;
    for (uint32_t e = tid; e < m; e += WF_THREADS) {
        const Sib M = sib[e];
        float score = 0.0f, stored = 0.0f;
        bool in_run = false;
        if (flat) {
            // Only flagged siblings can have sim != 0.  A sibling with sim == 0 never changes the
            // score (x + 0 = x; 0 > stored is false; and a later sim s > 0 in the same run gives
            // (score - 0) + s, the same value as a first add), so the fold of the reference
            // (src/line3D.cc:1527-1540) visits the flagged ones in list order, one run per camera block.
            const uint32_t words = (m + 31) >> 5;
            uint32_t next = sm.cnt[e];
            uint32_t cur_run = 0xffffffffu;
            for (uint32_t w = 0; w < words; ++w) {
                uint32_t bits = sm.mask[e * WF_MASKW + w];
                while (bits) {
                    const uint32_t j = (w << 5) + (__ffs(bits) - 1);
                    bits &= bits - 1;
                    const float sim = sm.sims[next++];
                    const uint32_t r = sm.runid[j];
                    if (r != cur_run) {
                        score = fa(score, sim);
                        stored = sim;
                        cur_run = r;
                    } else if (sim > stored) {
                        score = fs(score, stored);
                        score = fa(score, sim);
                        stored = sim;
                    }
                }
            }
            evals += m - sm.blk_n[sm.runid[e]];
        } else {
            const float2 rg = regs[e];
            const D3 dirM = d3(dirs[3 * e], dirs[3 * e + 1], dirs[3 * e + 2]);
            const bool Mvalid = (M.flags & 2u) != 0;
            for (uint32_t j = 0; j < m; ++j) {
                const Sib S2 = sib[j];
                if (S2.flags & 1u) in_run = false;
                if (S2.cam == M.cam) continue;
                ++evals;
                const float sim = sim_for_scoring(M.d_p1, M.d_p2, rg.x, rg.y, Mvalid, dirM, S2, dirs + 3 * j,
                                                  a.two_sigA_sqr, 0.5f, -0.70f, 0.5f, a.dotcut);
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
        const ListRec L = Lr[e];
        Lr[e].score = score;
        wmax = fmaxf(wmax, score);
        any_valid |= (score > 0.75f);
        if (L.src_idx != NOIDX) {
            a.fwd_rec[L.src_idx].score = score;
            // storeInverseMatches (src/line3D.cc:1986-2015)
            if (score > 0.0f) {
                uint32_t q = 0;
                while (sm.blk_pos[q + 1] <= e) ++q;
                if (sm.blk_flags[q] & 2u) {
                    const uint32_t tr = sm.blk_tbase[q] + L.tgt_seg;
                    const uint32_t slot = atomicAdd(&a.inv_fill[tr], 1u);
                    a.inv_ent[a.inv_off[tr] + slot] = make_uint2(L.src_idx, i);
                }
            }
        }
    }
    // block reductions: maximum (exact, order-free), counters
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
        evals += __shfl_xor_sync(0xffffffffu, evals, d);
    }
    if (lane == 0) {
        if (wmax > 0.0f) atomicMax(&a.view_max[v], float_ordered(wmax));
        if (evals) atomicAdd(&a.stats->sim_evals, (unsigned long long)evals);
    }
    const int row_valid = __syncthreads_or(any_valid ? 1 : 0);
    if (a.dbg && tid == 0) {
        const long long t_end = clock64();
        a.dbg[4 * g + 0] = (uint32_t)(t_gather - t_start);
        a.dbg[4 * g + 1] = (uint32_t)(t_end - t_gather);
        a.dbg[4 * g + 2] = m;
        a.dbg[4 * g + 3] = flat ? sm.total : 0xffffffffu;
    }
    if (tid == 0) {
        atomicAdd(&a.stats->scored, (unsigned long long)m);
        if (row_valid) atomicAdd(&a.stats->num_valid, 1u);
    }
}

// ------------------------------------------------------------------------------------------
// task B: one warp filters one row of the previous view
// ------------------------------------------------------------------------------------------
__device__ void wf_filter_row(const WfArgs& a, uint32_t v, uint32_t i, uint32_t lane)
{
    const ViewDev& va = a.views[v];
    const uint32_t g = va.seg_off + i;
    const uint32_t m = a.L_cnt[g];
    const size_t lbase = (size_t)a.L_base[v] + (a.L_off[g] - a.L_off[va.seg_off]);
    const ListRec* __restrict__ Lr = a.L_rec + lbase;
    const float max_score = fmaxf(0.0f, ordered_to_float(a.view_max[v]));
    const float lim = fm(0.10f, max_score);
    uint32_t kept = 0;
    float best = 0.0f;
    uint32_t best_idx = NOIDX;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t e = base + lane;
        float s = 0.0f;
        bool keep = false;
        if (e < m) {
            s = Lr[e].score;
            keep = (s > 0.0f) && (s > lim);
        }
        kept += __popc(__ballot_sync(0xffffffffu, keep));
        float cs = keep ? s : 0.0f;
        uint32_t ci = keep ? e : NOIDX;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, cs, d);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, ci, d);
            if (os > cs || (os == cs && oi < ci)) {
                cs = os;
                ci = oi;
            }
        }
        if (ci != NOIDX && cs > best) {  // first strict maximum in list order
            best = cs;
            best_idx = ci;
        }
    }
    uint32_t dst0 = 0;
    if (lane == 0 && kept) dst0 = atomicAdd(&a.stats->filt_cursor, kept);
    dst0 = __shfl_sync(0xffffffffu, dst0, 0);
    const bool fits = (kept == 0) || ((uint64_t)dst0 + kept <= a.filt_cap);
    if (!fits && lane == 0) atomicOr(&a.stats->err, 2u);
    uint32_t w = 0;
    if (kept && fits)
        for (uint32_t base = 0; base < m; base += 32) {
            const uint32_t e = base + lane;
            bool keep = false;
            ListRec L;
            if (e < m) {
                L = Lr[e];
                keep = (L.score > 0.0f) && (L.score > lim);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) a.filt_rec[dst0 + w + __popc(bal & ((1u << lane) - 1u))] = L;
            w += __popc(bal);
        }
    if (lane == 0) {
        a.filt_off[g] = dst0;
        a.filt_cnt[g] = fits ? kept : 0u;
        EntryDev& E = a.entries[g];
        if (best_idx != NOIDX && best > 0.75f) {
            const ListRec B = Lr[best_idx];
            const SegRays sr = a.rays[g];
            const D3 Ca = ld3w(va.C);
            D3 P1 = add3(Ca, scale3(ld3w(sr.r1), (double)B.d_p1));
            D3 P2 = add3(Ca, scale3(ld3w(sr.r2), (double)B.d_p2));
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
        } else {
            E.has = 0u;
        }
    }
}

__global__ void __launch_bounds__(WF_THREADS) k3_wavefront_kernel(const WfArgs a)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char wf_smem_raw[];
    WfSmem& sm = *reinterpret_cast<WfSmem*>(wf_smem_raw);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warps_per_cta = WF_THREADS / 32;
    for (uint32_t k = 0; k <= a.V; ++k) {
        if (k < a.V) {
            const uint32_t n_rows = a.views[k].n_seg;
            for (uint32_t i = blockIdx.x; i < n_rows; i += gridDim.x) {
                __syncthreads();  // shared staging is reused row after row
                wf_score_row(a, k, i, sm);
            }
        }
        if (k >= 1) {
            const uint32_t n_rows = a.views[k - 1].n_seg;
            for (uint32_t i = blockIdx.x * warps_per_cta + warp; i < n_rows; i += gridDim.x * warps_per_cta)
                wf_filter_row(a, k - 1, i, lane);
        }
        grid.sync();
    }
}

// ------------------------------------------------------------------------------------------
// pre-pass: sizes of the inverse CSR slots and of the lists (upper bounds)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_inv_capacity_kernel(const PairDev* __restrict__ pairs, uint32_t P,
                                                              uint32_t n_rows, const uint32_t* __restrict__ fwd_off,
                                                              const uint32_t* __restrict__ fwd_cnt,
                                                              const FwdRec* __restrict__ fwd_rec,
                                                              uint32_t* __restrict__ inv_cap)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t n = fwd_cnt[row];
    if (!n) return;
    uint32_t lo = 0, hi = P;  // pair of this row: largest p with row_base <= row
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pairs[mid].row_base <= row) lo = mid; else hi = mid;
    }
    const PairDev& D = pairs[lo];
    if (!D.emit_inverse) return;
    const uint32_t b = fwd_off[row];
    for (uint32_t e = 0; e < n; ++e) atomicAdd(&inv_cap[D.tgt_base + fwd_rec[b + e].c], 1u);
}

__global__ void __launch_bounds__(256) k3_list_capacity_kernel(const ViewDev* __restrict__ views,
                                                               const uint32_t* __restrict__ seg_view, uint32_t S,
                                                               const IncDev* __restrict__ inc,
                                                               const uint32_t* __restrict__ inc_off,
                                                               const PairDev* __restrict__ pairs,
                                                               const uint32_t* __restrict__ fwd_cnt,
                                                               const uint32_t* __restrict__ inv_cap,
                                                               uint32_t* __restrict__ L_ub)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= S) return;
    const uint32_t v = seg_view[g];
    const uint32_t i = g - views[v].seg_off;
    uint32_t m = 0;
    for (uint32_t q = inc_off[v]; q < inc_off[v + 1]; ++q) {
        const PairDev& P = pairs[inc[q].pair];
        m += inc[q].inverse ? inv_cap[P.tgt_base + i] : fwd_cnt[P.row_base + i];
    }
    L_ub[g] = m;
}

int launch_k3_inv_capacity(const PairDev* pairs, uint32_t P, uint32_t n_rows, const uint32_t* fwd_off,
                           const uint32_t* fwd_cnt, const FwdRec* fwd_rec, uint32_t* inv_cap, cudaStream_t st)
{
    if (!n_rows || !P) return 0;
    k3_inv_capacity_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(pairs, P, n_rows, fwd_off, fwd_cnt, fwd_rec, inv_cap);
    return 1;
}

int launch_k3_list_capacity(const ViewDev* views, const uint32_t* seg_view, uint32_t S, const IncDev* inc,
                            const uint32_t* inc_off, const PairDev* pairs, const uint32_t* fwd_cnt,
                            const uint32_t* inv_cap, uint32_t* L_ub, cudaStream_t st)
{
    if (!S) return 0;
    k3_list_capacity_kernel<<<(S + 255) / 256, 256, 0, st>>>(views, seg_view, S, inc, inc_off, pairs, fwd_cnt, inv_cap,
                                                              L_ub);
    return 1;
}


// ------------------------------------------------------------------------------------------
// pre-pass: 3-D geometry of every forward record, seen from both views
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ GeoRec make_geo(const D3& C, const D3& r1, const D3& r2, float d1, float d2, float k,
                                           const D3& Co, float ko)
{
    // M3D = View::unprojectSegment (src/view.cc:385-400) + Segment3D ctor (include/segment3D.h:58-77)
    D3 P1 = add3(C, scale3(r1, (double)d1));
    D3 P2 = add3(C, scale3(r2, (double)d2));
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
    const float sig1 = fm(d1, k), sig2 = fm(d2, k);
    float reg1 = fm(fm(2.0f, sig1), sig1);
    float reg2 = fm(fm(2.0f, sig2), sig2);
    const float s1t = (float)dm(norm3(sub3(P1, Co)), (double)ko);
    const float s2t = (float)dm(norm3(sub3(P2, Co)), (double)ko);
    reg1 = fm(0.5f, fa(reg1, fm(fm(2.0f, s1t), s1t)));
    reg2 = fm(0.5f, fa(reg2, fm(fm(2.0f, s2t), s2t)));
    GeoRec G;
    G.dir[0] = dir.x; G.dir[1] = dir.y; G.dir[2] = dir.z;
    G.reg1 = reg1;
    G.reg2 = reg2;
    G.valid = (len < L3D_EPS) ? 0u : 1u;
    G.pad = 0u;
    return G;
}

__global__ void __launch_bounds__(128) k3_geom_kernel(const PairDev* __restrict__ pairs, uint32_t P, uint32_t n_rows,
                                                      const ViewDev* __restrict__ views,
                                                      const SegRays* __restrict__ rays,
                                                      const uint32_t* __restrict__ fwd_off,
                                                      const uint32_t* __restrict__ fwd_cnt,
                                                      const FwdRec* __restrict__ fwd_rec, GeoRec* __restrict__ G_fwd,
                                                      GeoRec* __restrict__ G_inv)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t n = fwd_cnt[row];
    if (!n) return;
    uint32_t lo = 0, hi = P;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pairs[mid].row_base <= row) lo = mid; else hi = mid;
    }
    const PairDev& D = pairs[lo];
    const uint32_t i = row - D.row_base;
    const ViewDev& vs = views[D.src_view];
    const ViewDev& vt = views[D.tgt_view];
    const D3 Cs = ld3w(vs.C), Ct = ld3w(vt.C);
    const SegRays sr = rays[D.src_off + i];
    const D3 rs1 = ld3w(sr.r1), rs2 = ld3w(sr.r2);
    const uint32_t b = fwd_off[row];
    for (uint32_t e = 0; e < n; ++e) {
        const FwdRec f = fwd_rec[b + e];
        G_fwd[b + e] = make_geo(Cs, rs1, rs2, f.d_p1, f.d_p2, vs.k, Ct, vt.k);
        if (D.emit_inverse) {
            const SegRays tr = rays[D.tgt_off + f.c];
            G_inv[b + e] = make_geo(Ct, ld3w(tr.r1), ld3w(tr.r2), f.d_q1, f.d_q2, vt.k, Cs, vs.k);
        }
    }
}

int launch_k3_geom(const PairDev* pairs, uint32_t P, uint32_t n_rows, const ViewDev* views, const SegRays* rays,
                   const uint32_t* fwd_off, const uint32_t* fwd_cnt, const FwdRec* fwd_rec, void* G_fwd, void* G_inv,
                   cudaStream_t st)
{
    if (!n_rows || !P) return 0;
    k3_geom_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(pairs, P, n_rows, views, rays, fwd_off, fwd_cnt, fwd_rec,
                                                          (GeoRec*)G_fwd, (GeoRec*)G_inv);
    return 1;
}
size_t k3_geo_bytes() { return sizeof(GeoRec); }

size_t k3_wf_stats_bytes() { return sizeof(WfStats); }
size_t k3_sib_bytes() { return sizeof(Sib); }
int k3_wf_max_inc() { return WF_MAXINC; }

// cooperative launch of the wavefront; returns the number of launches or a negative CUDA error
int launch_k3_wavefront(const ViewDev* views, const PairDev* pairs, const IncDev* inc, const uint32_t* inc_off,
                        const SegRays* rays, const uint32_t* fwd_off, const uint32_t* fwd_cnt, FwdRec* fwd_rec,
                        const void* G_fwd, const void* G_inv, const uint32_t* inv_off, uint32_t* inv_fill, uint2* inv_ent, const uint32_t* L_off,
                        const uint64_t* L_base, uint32_t* L_cnt, ListRec* L_rec, void* L_sib, double* L_dir,
                        float2* L_reg, uint32_t* view_max, ListRec* filt_rec, uint32_t filt_cap, uint32_t* filt_off,
                        uint32_t* filt_cnt, EntryDev* entries, void* stats, uint32_t V, uint32_t max_rows,
                        float two_sigA_sqr, uint32_t* dbg, cudaStream_t st, int* err)
{
    WfArgs a;
    a.views = views; a.pairs = pairs; a.inc = inc; a.inc_off = inc_off; a.rays = rays;
    a.fwd_off = fwd_off; a.fwd_cnt = fwd_cnt; a.fwd_rec = fwd_rec;
    a.G_fwd = (const GeoRec*)G_fwd; a.G_inv = (const GeoRec*)G_inv;
    a.inv_off = inv_off; a.inv_fill = inv_fill; a.inv_ent = inv_ent;
    a.L_off = L_off; a.L_base = L_base; a.L_cnt = L_cnt; a.L_rec = L_rec; a.L_sib = (Sib*)L_sib; a.L_dir = L_dir; a.L_reg = L_reg;
    a.view_max = view_max; a.filt_rec = filt_rec; a.filt_cap = filt_cap; a.filt_off = filt_off; a.filt_cnt = filt_cnt;
    a.entries = entries; a.stats = (WfStats*)stats; a.V = V; a.two_sigA_sqr = two_sigA_sqr;
    a.dotcut = score_dotcut(two_sigA_sqr, 0.5f);
    a.dbg = dbg;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(WfSmem);
    cudaFuncSetAttribute(k3_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k3_wavefront_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_wavefront_kernel, WF_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) {
        *err = (int)e;
        return -1;
    }
    // one CTA per row of the largest view if they all fit, otherwise every resident slot
    uint32_t grid = (uint32_t)(sms * per_sm);
    const uint32_t want = ((max_rows + (uint32_t)sms - 1) / (uint32_t)sms) * (uint32_t)sms;
    if (want < grid) grid = want > 0 ? want : (uint32_t)sms;
    void* args[] = {(void*)&a};
    e = cudaLaunchCooperativeKernel((void*)k3_wavefront_kernel, dim3(grid), dim3(WF_THREADS), args, smem, st);
    if (e != cudaSuccess) {
        *err = (int)e;
        return -1;
    }
    return 1;
}

}  // namespace l3d
