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
static constexpr int WF_THREADS = 128;
static constexpr int WF_MAXM = 256;   // list entries staged in shared memory
static constexpr int WF_MAXINC = 64;  // incident pairs staged in shared memory

struct WfStats {
    unsigned long long sim_evals;
    unsigned long long scored;
    uint32_t num_valid;
    uint32_t filt_cursor;  // bump allocator of the filtered-record store
    uint32_t err;          // bit0: incident list too long, bit1: filtered store overflow
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
};

__device__ __forceinline__ D3 ld3w(const double* p) { return D3{p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------
// task A: one CTA assembles, scores and propagates one row
// ------------------------------------------------------------------------------------------
__device__ void wf_score_row(const WfArgs& a, uint32_t v, uint32_t i, Sib* s_sib, double* s_dir, float2* s_reg,
                             uint32_t* s_blk)
{
    // s_blk: [0..WF_MAXINC) start b, [WF_MAXINC..2*WF_MAXINC) count n, [2*WF_MAXINC..3*WF_MAXINC] list position
    uint32_t* blk_b = s_blk;
    uint32_t* blk_n = s_blk + WF_MAXINC;
    uint32_t* blk_pos = s_blk + 2 * WF_MAXINC;
    const ViewDev& va = a.views[v];
    const uint32_t i0 = a.inc_off[v], n_inc = a.inc_off[v + 1] - i0;
    const uint32_t g = va.seg_off + i;
    const int tid = threadIdx.x;

    // block table (n_inc <= WF_MAXINC is checked on the host)
    if (tid < (int)n_inc) {
        const IncDev q = a.inc[i0 + tid];
        const PairDev& P = a.pairs[q.pair];
        if (q.inverse) {
            blk_b[tid] = a.inv_off[P.tgt_base + i];
            blk_n[tid] = a.inv_fill[P.tgt_base + i];
        } else {
            blk_b[tid] = a.fwd_off[P.row_base + i];
            blk_n[tid] = a.fwd_cnt[P.row_base + i];
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (uint32_t q = 0; q < n_inc; ++q) {
            blk_pos[q] = run;
            run += blk_n[q];
        }
        blk_pos[n_inc] = run;
    }
    __syncthreads();
    const uint32_t m = blk_pos[n_inc];
    if (tid == 0) a.L_cnt[g] = m;
    if (m == 0) return;  // uniform

    const size_t lbase = (size_t)a.L_base[v] + (a.L_off[g] - a.L_off[va.seg_off]);
    ListRec* __restrict__ Lr = a.L_rec + lbase;
    Sib* __restrict__ Ls = a.L_sib + lbase;
    double* __restrict__ Ld = a.L_dir + 3 * lbase;
    float2* __restrict__ Lg = a.L_reg + lbase;
    const bool in_smem = m <= WF_MAXM;

    const SegRays sr = a.rays[g];
    const D3 r1 = ld3w(sr.r1), r2 = ld3w(sr.r2);
    const D3 Ca = ld3w(va.C);
    const float k = va.k;

    // ---- assemble + geometry: one thread per list entry ----
    for (uint32_t e = tid; e < m; e += WF_THREADS) {
        uint32_t q = 0;
        while (blk_pos[q + 1] <= e) ++q;  // n_inc is small
        uint32_t j = e - blk_pos[q];
        const IncDev iq = a.inc[i0 + q];
        const PairDev& P = a.pairs[iq.pair];
        const uint32_t b = blk_b[q], n = blk_n[q];
        ListRec L;
        uint32_t dst = e;
        if (iq.inverse) {
            const uint2 ie = a.inv_ent[b + j];
            // append order of the reference = ascending forward-record index: rank sort
            uint32_t rank = 0;
            for (uint32_t z = 0; z < n; ++z) rank += (a.inv_ent[b + z].x < ie.x) ? 1u : 0u;
            dst = blk_pos[q] + rank;
            j = rank;
            const FwdRec f = a.fwd_rec[ie.x];
            L.tgt_view = P.src_view;
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
            L.tgt_view = P.tgt_view;
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
        const ViewDev& vo = a.views[L.tgt_view];
        const D3 Co = ld3w(vo.C);
        const float sig1 = fm(L.d_p1, k), sig2 = fm(L.d_p2, k);
        float reg1 = fm(fm(2.0f, sig1), sig1);
        float reg2 = fm(fm(2.0f, sig2), sig2);
        const float s1t = (float)dm(norm3(sub3(P1, Co)), (double)vo.k);
        const float s2t = (float)dm(norm3(sub3(P2, Co)), (double)vo.k);
        reg1 = fm(0.5f, fa(reg1, fm(fm(2.0f, s1t), s1t)));
        reg2 = fm(0.5f, fa(reg2, fm(fm(2.0f, s2t), s2t)));
        Sib sb;
        sb.d_p1 = L.d_p1;
        sb.d_p2 = L.d_p2;
        sb.cam = L.tgt_view;
        sb.flags = ((j == 0) ? 1u : 0u) | ((len < L3D_EPS) ? 0u : 2u);
        Lr[dst] = L;
        Ls[dst] = sb;
        Ld[3 * dst + 0] = dir.x;
        Ld[3 * dst + 1] = dir.y;
        Ld[3 * dst + 2] = dir.z;
        Lg[dst] = make_float2(reg1, reg2);
        if (in_smem) {
            s_sib[dst] = sb;
            s_dir[3 * dst + 0] = dir.x;
            s_dir[3 * dst + 1] = dir.y;
            s_dir[3 * dst + 2] = dir.z;
            s_reg[dst] = make_float2(reg1, reg2);
        }
    }
    __threadfence_block();
    __syncthreads();

    // ---- score: one thread per match M, siblings in list order ----
    const Sib* __restrict__ sib = in_smem ? s_sib : Ls;
    const double* __restrict__ dirs = in_smem ? s_dir : Ld;
    float wmax = 0.0f;
    uint32_t evals = 0;
    bool any_valid = false;
    for (uint32_t e = tid; e < m; e += WF_THREADS) {
        const Sib M = sib[e];
        const float2 rg = in_smem ? s_reg[e] : Lg[e];
        const float reg1 = rg.x, reg2 = rg.y;
        const D3 dirM = d3(dirs[3 * e], dirs[3 * e + 1], dirs[3 * e + 2]);
        const bool Mvalid = (M.flags & 2u) != 0;
        float score = 0.0f, stored = 0.0f;
        bool in_run = false;
        for (uint32_t j = 0; j < m; ++j) {
            const Sib S2 = sib[j];
            if (S2.flags & 1u) in_run = false;
            if (S2.cam == M.cam) continue;
            ++evals;
            const float sim = sim_for_scoring(M.d_p1, M.d_p2, reg1, reg2, Mvalid, dirM, S2, dirs + 3 * j,
                                              a.two_sigA_sqr, 0.5f, -0.70f, 0.5f);
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
        const ListRec L = Lr[e];
        Lr[e].score = score;
        wmax = fmaxf(wmax, score);
        any_valid |= (score > 0.75f);
        if (L.src_idx != NOIDX) {
            a.fwd_rec[L.src_idx].score = score;
            // storeInverseMatches (src/line3D.cc:1986-2015)
            if (score > 0.0f) {
                uint32_t q = 0;
                while (blk_pos[q + 1] <= e) ++q;
                const PairDev& P = a.pairs[a.inc[i0 + q].pair];
                if (P.emit_inverse) {
                    const uint32_t tr = P.tgt_base + L.tgt_seg;
                    const uint32_t slot = atomicAdd(&a.inv_fill[tr], 1u);
                    a.inv_ent[a.inv_off[tr] + slot] = make_uint2(L.src_idx, i);
                }
            }
        }
    }
    // block reductions: maximum (exact, order-free), counters
    const uint32_t lane = tid & 31;
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
    __shared__ Sib s_sib[WF_MAXM];
    __shared__ double s_dir[3 * WF_MAXM];
    __shared__ float2 s_reg[WF_MAXM];
    __shared__ uint32_t s_blk[3 * WF_MAXINC + 2];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warps_per_cta = WF_THREADS / 32;
    for (uint32_t k = 0; k <= a.V; ++k) {
        if (k < a.V) {
            const uint32_t n_rows = a.views[k].n_seg;
            for (uint32_t i = blockIdx.x; i < n_rows; i += gridDim.x) {
                __syncthreads();  // shared staging is reused row after row
                wf_score_row(a, k, i, s_sib, s_dir, s_reg, s_blk);
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

size_t k3_wf_stats_bytes() { return sizeof(WfStats); }
size_t k3_sib_bytes() { return sizeof(Sib); }
int k3_wf_max_inc() { return WF_MAXINC; }

// cooperative launch of the wavefront; returns the number of launches or a negative CUDA error
int launch_k3_wavefront(const ViewDev* views, const PairDev* pairs, const IncDev* inc, const uint32_t* inc_off,
                        const SegRays* rays, const uint32_t* fwd_off, const uint32_t* fwd_cnt, FwdRec* fwd_rec,
                        const uint32_t* inv_off, uint32_t* inv_fill, uint2* inv_ent, const uint32_t* L_off,
                        const uint64_t* L_base, uint32_t* L_cnt, ListRec* L_rec, void* L_sib, double* L_dir,
                        float2* L_reg, uint32_t* view_max, ListRec* filt_rec, uint32_t filt_cap, uint32_t* filt_off,
                        uint32_t* filt_cnt, EntryDev* entries, void* stats, uint32_t V, uint32_t max_rows,
                        float two_sigA_sqr, cudaStream_t st, int* err)
{
    WfArgs a;
    a.views = views; a.pairs = pairs; a.inc = inc; a.inc_off = inc_off; a.rays = rays;
    a.fwd_off = fwd_off; a.fwd_cnt = fwd_cnt; a.fwd_rec = fwd_rec;
    a.inv_off = inv_off; a.inv_fill = inv_fill; a.inv_ent = inv_ent;
    a.L_off = L_off; a.L_base = L_base; a.L_cnt = L_cnt; a.L_rec = L_rec; a.L_sib = (Sib*)L_sib; a.L_dir = L_dir; a.L_reg = L_reg;
    a.view_max = view_max; a.filt_rec = filt_rec; a.filt_cap = filt_cap; a.filt_off = filt_off; a.filt_cnt = filt_cnt;
    a.entries = entries; a.stats = (WfStats*)stats; a.V = V; a.two_sigA_sqr = two_sigA_sqr;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_wavefront_kernel, WF_THREADS, 0);
    if (e != cudaSuccess || per_sm < 1) {
        *err = (int)e;
        return -1;
    }
    // one CTA per row of the largest view if they all fit, otherwise every resident slot
    uint32_t grid = (uint32_t)(sms * per_sm);
    const uint32_t want = ((max_rows + (uint32_t)sms - 1) / (uint32_t)sms) * (uint32_t)sms;
    if (want < grid) grid = want > 0 ? want : (uint32_t)sms;
    void* args[] = {(void*)&a};
    e = cudaLaunchCooperativeKernel((void*)k3_wavefront_kernel, dim3(grid), dim3(WF_THREADS), args, 0, st);
    if (e != cudaSuccess) {
        *err = (int)e;
        return -1;
    }
    return 1;
}

}  // namespace l3d
