// k3_dataflow.cu -- K3: scoring of Line3D::computeMatches (src/line3D.cc:846-930) as a data-flow
// computation instead of a per-view wavefront (exact TU).
//
// The reference processes the views in ascending camera-ID order because
// Line3D::storeInverseMatches (src/line3D.cc:1986-2015) appends every forward match of view A whose
// score3D_ is > 0 to the list of its target segment in a not yet processed view B.  What flows from
// A to B is only the PRESENCE of those entries: the similarity of two list entries
// (Line3D::similarityForScoring, src/line3D.cc:1685-1716) depends on their depths and 3-D
// directions, which are known as soon as matching is done.  So:
//
//   build  (all rows of all views at once, one CTA per segment): assemble the POTENTIAL list of the
//          row = [every forward match pointing at this segment from an earlier view, in append
//          order] ++ [its own forward matches per target camera]; evaluate the similarity of every
//          (M, sibling) pair that can be non-zero; emit the row's "fold program": per match M
//          with at least one such sibling a head record and the sibling records (record index whose
//          score decides presence, similarity, camera block).
//   fold   (one persistent kernel, rows dealt round-robin to the resident warps in (view, segment)
//          order): Line3D::scoringCPU's accumulation (src/line3D.cc:1515-1543) over the siblings
//          that are present; a sibling that is an inverse match waits (spins) on the score of its
//          forward record, which a row of an earlier view produces.  Every warp takes its rows in
//          ascending order and a row only ever waits on rows of earlier views, so the smallest
//          unfinished row never waits: no deadlock, no grid-wide barrier, and the critical path is
//          the longest dependency chain instead of the number of views.
//   finish (all rows at once, one warp per segment): which potential entries exist, the scored
//          lists (optional), Line3D::filterMatches (src/line3D.cc:1911-1983) and the
//          estimated_position3D_ row.
//
// A sibling with similarity 0 never changes score3D_ (x + 0 = x; 0 > stored is false; a later
// s > 0 of the same camera gives (score - 0) + s, the same value as a first add), so only pairs
// that pass the cheap certain-reject test are evaluated and folded.
#include <algorithm>
#include <cstdlib>

#include "internal.h"
#include "score_core.cuh"

namespace l3d {

#define L3D_EPS 1e-12
static constexpr uint32_t NOIDX = 0xffffffffu;
static constexpr int DF_MAXINC = 64;      // incident pairs of one view
static constexpr int DF_MAXM_CAP = 2048;  // potential entries staged in shared memory at most
static constexpr uint32_t DF_SPIN_LIMIT = 1u << 22;

struct WfStats {
    unsigned long long sim_evals;
    unsigned long long scored;
    uint32_t num_valid;
    uint32_t filt_cursor;  // bump allocator of the filtered-record store
    uint32_t err;          // bit1: filtered store overflow, bit2: program store overflow, bit3: dependency timeout
    uint32_t prog_cursor;  // bump allocator of the fold programs (16-byte units)
    uint32_t ticket;
    uint32_t max_list;     // longest potential list
    uint32_t pad[2];
};

// 3-D direction and regularisers of a list entry (a forward match seen from its source view, or an
// inverse match seen from its target view), computed while the row is assembled
struct GeoRec {
    double dir[3];
    float reg1, reg2;
    uint32_t valid;
    uint32_t pad;
};

__device__ __forceinline__ D3 ld3w(const double* p) { return D3{p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------
// pre-pass
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pair_of_row(const PairDev* __restrict__ pairs, uint32_t P, uint32_t row)
{
    uint32_t lo = 0, hi = P;  // largest p with row_base <= row
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pairs[mid].row_base <= row) lo = mid; else hi = mid;
    }
    return lo;
}

// the pair of each of 32 ascending consecutive rows: lane 0 searches, the others walk on from its answer (a search
// per thread was fourteen dependent loads, the whole cost of the two kernels below)
__device__ __forceinline__ uint32_t pair_of_row_warp(const PairDev* __restrict__ pairs, uint32_t P, uint32_t row,
                                                     uint32_t first_row)
{
    uint32_t p = 0;
    if ((threadIdx.x & 31) == 0) p = pair_of_row(pairs, P, first_row);
    p = __shfl_sync(0xffffffffu, p, 0);
    while (p + 1 < P && pairs[p + 1].row_base <= row) ++p;
    return p;
}

// per (pair, source row): the row and the pair of every forward record, and how many records point at each
// (pair, target segment) = the capacity of its inverse-match slot
__global__ void __launch_bounds__(256) k3_inv_capacity_kernel(const PairDev* __restrict__ pairs, uint32_t P,
                                                              uint32_t n_rows, const uint32_t* __restrict__ fwd_off,
                                                              const uint32_t* __restrict__ fwd_cnt,
                                                              const FwdRec* __restrict__ fwd_rec,
                                                              uint32_t* __restrict__ fwd_row,
                                                              uint32_t* __restrict__ inv_cap, uint32_t v_lo,
                                                              uint32_t v_hi)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t rc = min(row, n_rows - 1);
    const uint32_t pidx = pair_of_row_warp(pairs, P, rc, min(row - (threadIdx.x & 31), n_rows - 1));
    if (row >= n_rows) return;
    const uint32_t n = fwd_cnt[row];
    if (!n) return;
    const PairDev& D = pairs[pidx];
    const uint32_t b = fwd_off[row];
    // inverse slots are only needed (and, in a sharded run, the records only present) for the
    // target views [v_lo, v_hi) this rank builds
    const bool slots = D.emit_inverse && D.tgt_view >= v_lo && D.tgt_view < v_hi;
    for (uint32_t e = 0; e < n; ++e) {
        fwd_row[b + e] = row;
        if (slots) atomicAdd(&inv_cap[D.tgt_base + fwd_rec[b + e].c], 1u);
    }
}

__global__ void __launch_bounds__(256) k3_list_capacity_kernel(const ViewDev* __restrict__ views,
                                                               const uint32_t* __restrict__ seg_view, uint32_t S,
                                                               const IncDev* __restrict__ inc,
                                                               const uint32_t* __restrict__ inc_off,
                                                               const PairDev* __restrict__ pairs,
                                                               const uint32_t* __restrict__ fwd_cnt,
                                                               const uint32_t* __restrict__ inv_cap,
                                                               uint32_t* __restrict__ L_ub, WfStats* __restrict__ stats)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0;
    if (g < S) {
        const uint32_t v = seg_view[g];
        const uint32_t i = g - views[v].seg_off;
        for (uint32_t q = inc_off[v]; q < inc_off[v + 1]; ++q) {
            const PairDev& P = pairs[inc[q].pair];
            m += inc[q].inverse ? inv_cap[P.tgt_base + i] : fwd_cnt[P.row_base + i];
        }
        L_ub[g] = m;
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&stats->max_list, m);
}

__device__ __forceinline__ GeoRec make_geo(const D3& C, const D3& r1, const D3& r2, float d1, float d2, float k,
                                           const D3& Co, float ko)
{
    // View::unprojectSegment (src/view.cc:385-400) + Segment3D ctor (include/segment3D.h:58-77)
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

// one thread per forward record: score reset and the (static) inverse-match slot of its target
// segment (only for the target views [v_lo, v_hi) this rank builds)
__global__ void __launch_bounds__(256) k3_record_kernel(const PairDev* __restrict__ pairs, uint32_t P, uint32_t F,
                                                        const uint32_t* __restrict__ fwd_row,
                                                        const FwdRec* __restrict__ fwd_rec,
                                                        float* __restrict__ fwd_score,
                                                        const uint32_t* __restrict__ inv_off,
                                                        uint32_t* __restrict__ inv_fill, uint2* __restrict__ inv_ent,
                                                        uint32_t v_lo, uint32_t v_hi)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    // records are in row order: the rows of a warp's 32 records ascend
    const uint32_t row = fwd_row[min(f, F - 1)];
    const uint32_t pidx = pair_of_row_warp(pairs, P, row, __shfl_sync(0xffffffffu, row, 0));
    if (f >= F) return;
    fwd_score[f] = 0.0f;
    const PairDev& D = pairs[pidx];
    if (!(D.emit_inverse && D.tgt_view >= v_lo && D.tgt_view < v_hi)) return;
    const uint32_t tr_row = D.tgt_base + fwd_rec[f].c;
    const uint32_t slot = atomicAdd(&inv_fill[tr_row], 1u);
    inv_ent[inv_off[tr_row] + slot] = make_uint2(f, row - D.row_base);
}

// ------------------------------------------------------------------------------------------
// build: potential list + fold program of one row
// ------------------------------------------------------------------------------------------
struct BuildArgs {
    const ViewDev* views;
    const uint32_t* seg_view;
    const PairDev* pairs;
    const IncDev* inc;
    const uint32_t* inc_off;  // [V+1]
    const uint32_t* fwd_off;
    const uint32_t* fwd_cnt;
    const FwdRec* fwd_rec;
    float* fwd_score;  // score3D_ of every forward record, dense (doubles as the fold's ready flag)
    const SegRays* rays;
    const uint32_t* inv_off;   // start of the slot of every (pair, tgt segment)
    const uint32_t* inv_fill;  // entries in the slot
    const uint2* inv_ent;      // x: forward record index, y: source row
    const uint32_t* L_off;     // [S+1] offsets of the potential lists (global segment order)
    uint32_t* L_f;             // per potential entry: forward record index
    unsigned char* L_meta;     // per potential entry: incident-pair slot | inverse << 7
    // global staging for rows longer than maxm (may be NULL when the host knows every row fits)
    Sib* L_sib;
    double* L_dir;
    float2* L_reg;
    uint32_t* L_c;
    uint32_t* L_h;
    uint32_t* prog_off;  // [S] first 16-byte unit of the row's program
    uint32_t* prog_nh;   // [S] number of head records (0: nothing to fold)
    uint4* prog;
    uint32_t prog_cap;
    WfStats* stats;
    uint32_t S;
    uint32_t g_lo, g_hi;  // rows built by this rank (global segment indices)
    uint32_t maxm;
    float two_sigA_sqr;
    float dotcut;  // see score_core.cuh
};

struct BlockTab {
    uint32_t b[DF_MAXINC], n[DF_MAXINC], pos[DF_MAXINC + 1];
    uint32_t other[DF_MAXINC];  // the other view of the pair
    uint32_t inv[DF_MAXINC];    // 1: inverse block
};

// cheap certain reject of similarityForScoring: |d| > t >= sqrt(0.75 reg) (1 + 1e-5)  =>  -d^2/reg < -0.70
// after every rounding of the exact sequence  =>  exp(.) < 0.4966 < 0.5  =>  the similarity is truncated
// to 0.  The test only selects which pairs get the full evaluation; it never decides a result.
__device__ __forceinline__ float reject_threshold(float reg)
{
    return (reg > 0.0f) ? __fsqrt_ru(0.75f * reg) * 1.00001f : __int_as_float(0x7f800000);
}
__device__ __forceinline__ bool pair_flagged(const Sib& M, bool Mok, float t1, float t2, const Sib& S2)
{
    const float d1 = fs(M.d_p1, S2.d_p1), d2 = fs(M.d_p2, S2.d_p2);
    const bool rej = (fabsf(d1) > t1) | (fabsf(d2) > t2);
    return Mok & ((S2.flags & 2u) != 0) & (S2.cam != M.cam) & !rej;
}

static constexpr int DF_MASKM = 256;  // rows up to this length keep their flag bits in shared memory
                                      // between the count and emit passes (maskw words per entry)
static constexpr int DF_SMALL = 128;  // rows up to this length are built by a launch with small staging

// bitonic sort of 64 keys held two per lane, element i = 32 r + lane (a: r = 0, b: r = 1), ascending in i
__device__ __forceinline__ void warp_sort64(uint32_t& a, uint32_t& b, uint32_t lane)
{
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {  // k == 64: the partner of element i is the lane's other key
                const uint32_t lo = min(a, b), hi = max(a, b);
                a = lo;
                b = hi;
            } else {
                const uint32_t oa = __shfl_xor_sync(0xffffffffu, a, j), ob = __shfl_xor_sync(0xffffffffu, b, j);
                const bool lower = (lane & j) == 0;
                const bool up_a = (lane & k) == 0, up_b = ((lane + 32u) & k) == 0;
                a = (lower == up_a) ? min(a, oa) : max(a, oa);
                b = (lower == up_b) ? min(b, ob) : max(b, ob);
            }
        }
    }
}

// first position of the ascending run[0 .. 64) whose key is not below `key`
__device__ __forceinline__ uint32_t run_lower_bound(const uint32_t* __restrict__ run, uint32_t key)
{
    uint32_t lo = 0;
#pragma unroll
    for (int s2 = 32; s2 > 0; s2 >>= 1)
        if (run[lo + s2 - 1] < key) lo += s2;
    return lo + ((lo == 63u && run[63] < key) ? 1u : 0u);
}

// one row (global segment g) by one CTA of NT threads; every early return is CTA-uniform
template <int NT>
__device__ __forceinline__ void build_row(const BuildArgs& a, const uint32_t g, unsigned char* df_smem, BlockTab& bt,
                                          uint32_t* s_tot, uint32_t& s_base)
{
    constexpr int DF_THREADS = NT;
    const uint32_t v = a.seg_view[g];
    const ViewDev& va = a.views[v];
    const uint32_t i = g - va.seg_off;
    const uint32_t i0 = a.inc_off[v], n_inc = a.inc_off[v + 1] - i0;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31;

    // carve the dynamic shared memory
    const uint32_t maxm = a.maxm;
    double* sm_dir = reinterpret_cast<double*>(df_smem);
    Sib* sm_sib = reinterpret_cast<Sib*>(sm_dir + 3 * (size_t)maxm);
    float2* sm_reg = reinterpret_cast<float2*>(sm_sib + maxm);
    uint32_t* sm_c = reinterpret_cast<uint32_t*>(sm_reg + maxm);
    uint32_t* sm_h = sm_c + (maxm + 2);
    uint32_t* sm_mask = sm_h + (maxm + 2);  // maskm x maskw words (16-byte aligned: maxm is a multiple of 4)
    const uint32_t maskm = maxm < (uint32_t)DF_MASKM ? maxm : (uint32_t)DF_MASKM;
    const uint32_t maskw = (maskm + 31u) >> 5;
    uint32_t* sm_ukey = sm_mask + (((size_t)maskm * maskw + 3) & ~(size_t)3);  // depth keys, list order
    uint32_t* sm_skey = sm_ukey + ((maskm + 3u) & ~3u);                          // sorted ascending
    uint32_t* sm_runs = sm_skey + ((maskm + 3u) & ~3u);                          // 2 NT keys: the warps' sorted runs

    // block table (n_inc <= DF_MAXINC is checked on the host)
    if (tid < (int)n_inc) {
        const IncDev q = a.inc[i0 + tid];
        const PairDev& P = a.pairs[q.pair];
        if (q.inverse) {
            bt.b[tid] = a.inv_off[P.tgt_base + i];
            bt.n[tid] = a.inv_fill[P.tgt_base + i];
            bt.other[tid] = P.src_view;
            bt.inv[tid] = 1u;
        } else {
            bt.b[tid] = a.fwd_off[P.row_base + i];
            bt.n[tid] = a.fwd_cnt[P.row_base + i];
            bt.other[tid] = P.tgt_view;
            bt.inv[tid] = 0u;
        }
    }
    __syncthreads();
    if (tid < 32) {  // exclusive prefix of the block sizes (n_inc <= 64: two values per lane)
        const uint32_t n0 = (lane < n_inc) ? bt.n[lane] : 0u;
        const uint32_t n1 = (lane + 32 < n_inc) ? bt.n[lane + 32] : 0u;
        uint32_t x0 = n0, x1 = n1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t0 = __shfl_up_sync(0xffffffffu, x0, d), t1 = __shfl_up_sync(0xffffffffu, x1, d);
            if ((int)lane >= d) { x0 += t0; x1 += t1; }
        }
        const uint32_t tot0 = __shfl_sync(0xffffffffu, x0, 31);
        if (lane < n_inc) bt.pos[lane] = x0 - n0;
        if (lane + 32 < n_inc) bt.pos[lane + 32] = tot0 + x1 - n1;
        const uint32_t tot = tot0 + __shfl_sync(0xffffffffu, x1, 31);
        if (lane == 0) bt.pos[n_inc] = tot;
    }
    __syncthreads();
    const uint32_t m = bt.pos[n_inc];
    if (m == 0) {
        if (tid == 0) a.prog_nh[g] = 0u;
        return;  // uniform
    }
    const size_t lbase = a.L_off[g];
    const bool in_smem = m <= maxm;
    if (!in_smem && a.L_sib == nullptr) {  // the host sized maxm from an upper bound: cannot happen
        if (tid == 0) {
            a.prog_nh[g] = 0u;
            atomicOr(&a.stats->err, 4u);
        }
        return;
    }
    Sib* __restrict__ sib = in_smem ? sm_sib : a.L_sib + lbase;
    double* __restrict__ dirs = in_smem ? sm_dir : a.L_dir + 3 * lbase;
    float2* __restrict__ regs = in_smem ? sm_reg : a.L_reg + lbase;
    uint32_t* __restrict__ cnt = in_smem ? sm_c : a.L_c + lbase + g;   // m + 1 values per row
    uint32_t* __restrict__ hof = in_smem ? sm_h : a.L_h + lbase + g;
    uint32_t* __restrict__ Lf = a.L_f + lbase;
    unsigned char* __restrict__ Lm = a.L_meta + lbase;

    // ---- assemble: one thread per potential entry (View::unprojectSegment + regularisers) ----
    const SegRays sra = a.rays[g];
    const D3 Ca = ld3w(va.C), ra1 = ld3w(sra.r1), ra2 = ld3w(sra.r2);
    for (uint32_t e = tid; e < m; e += DF_THREADS) {
        uint32_t q = 0, qh = n_inc;  // the block of entry e: the last one that starts at or before it
        while (qh - q > 1) {
            const uint32_t mid = (q + qh) >> 1;
            if (bt.pos[mid] <= e) q = mid; else qh = mid;
        }
        const uint32_t j = e - bt.pos[q];
        const uint32_t b = bt.b[q], n = bt.n[q];
        uint32_t dst = e, f;
        Sib sb;
        if (bt.inv[q]) {
            const uint2 ie = a.inv_ent[b + j];
            // append order of the reference = ascending forward-record index: rank sort
            uint32_t rank = 0;
            for (uint32_t z = 0; z < n; ++z) rank += (a.inv_ent[b + z].x < ie.x) ? 1u : 0u;
            dst = bt.pos[q] + rank;
            f = ie.x;
            const FwdRec r = a.fwd_rec[f];
            sb.d_p1 = r.d_q1;  // storeInverseMatches swaps the views and the depths
            sb.d_p2 = r.d_q2;
        } else {
            f = b + j;
            const FwdRec r = a.fwd_rec[f];
            sb.d_p1 = r.d_p1;
            sb.d_p2 = r.d_p2;
        }
        // either way the entry is a 3-D segment through THIS row's 2-D segment at the entry's depths
        const ViewDev& vo = a.views[bt.other[q]];
        const GeoRec G = make_geo(Ca, ra1, ra2, sb.d_p1, sb.d_p2, va.k, ld3w(vo.C), vo.k);
        sb.cam = bt.other[q];
        sb.flags = G.valid ? 2u : 0u;
        Lf[dst] = f;
        Lm[dst] = (unsigned char)(q | (bt.inv[q] << 7));
        sib[dst] = sb;
        dirs[3 * dst + 0] = G.dir[0];
        dirs[3 * dst + 1] = G.dir[1];
        dirs[3 * dst + 2] = G.dir[2];
        regs[dst] = make_float2(G.reg1, G.reg2);
    }
    __threadfence_block();
    __syncthreads();

    // ---- find the (M, sibling) pairs that need the full similarity ----
    // Rows staged in shared memory: the entries are ranked by their first depth, so the siblings that
    // can pass |d_p1 - d_p1'| <= t1 form a window of the sorted order (found by binary search); only
    // they get the complete cheap test.  Flag bits are kept for the emit pass.
    const bool use_mask = m <= (uint32_t)DF_MASKM && m <= maxm;
    if (use_mask) {
        // key = bits of the (positive) first depth with the list position in the low byte: unique,
        // ordered like the depths up to 256 ulps (the window below is widened by more than that);
        // invalid 3-D segments sort last
        for (uint32_t e = tid; e < m; e += DF_THREADS) {
            const Sib sb = sm_sib[e];
            const uint32_t kb = (sb.flags & 2u) ? (__float_as_uint(fmaxf(sb.d_p1, 0.0f)) & 0xffffff00u) : 0x7f800000u;
            sm_ukey[e] = kb | e;
        }
        __syncthreads();
        if (m > 64u && m <= 2u * NT) {
            // every warp sorts 64 keys in registers (two per lane, shuffles only), the sorted runs go to shared
            // memory, and a key's rank is its position in its own run plus its lower bounds in the other runs (the
            // keys are unique).  The all-pairs rank sort below was a quarter (rows up to 128 entries) to a third
            // (up to 256) of the kernel's instructions.
            uint32_t ka = (uint32_t)tid < m ? sm_ukey[tid] : 0xffffffffu;  // pads sort last and are never written
            uint32_t kb2 = (uint32_t)tid + NT < m ? sm_ukey[tid + NT] : 0xffffffffu;
            warp_sort64(ka, kb2, lane);
            const uint32_t wrp = (uint32_t)tid >> 5;
            sm_runs[wrp * 64u + lane] = ka;
            sm_runs[wrp * 64u + 32u + lane] = kb2;
            __syncthreads();
            uint32_t ra = lane, rb = 32u + lane;
#pragma unroll
            for (uint32_t w2 = 0; w2 < NT / 32; ++w2)
                if (w2 != wrp) {
                    ra += run_lower_bound(sm_runs + w2 * 64u, ka);
                    rb += run_lower_bound(sm_runs + w2 * 64u, kb2);
                }
            if (ka != 0xffffffffu) sm_skey[ra] = ka;
            if (kb2 != 0xffffffffu) sm_skey[rb] = kb2;
        } else {
            for (uint32_t e = tid; e < m; e += DF_THREADS) {
                const uint32_t ke = sm_ukey[e];
                uint32_t r = 0;
                uint32_t j = 0;
                for (; j + 4 <= m; j += 4) {  // sm_ukey is 16-byte aligned
                    const uint4 k4 = *reinterpret_cast<const uint4*>(sm_ukey + j);
                    r += (k4.x < ke) + (k4.y < ke) + (k4.z < ke) + (k4.w < ke);
                }
                for (; j < m; ++j) r += sm_ukey[j] < ke;
                sm_skey[r] = ke;
            }
        }
        __syncthreads();
    }
    for (uint32_t e = tid; e < m; e += DF_THREADS) {
        const Sib M = sib[e];
        const float2 rg = regs[e];
        const float t1 = reject_threshold(rg.x), t2 = reject_threshold(rg.y);
        const bool Mok = (M.flags & 2u) != 0;
        uint32_t c = 0;
        if (use_mask) {
            const uint32_t words = (m + 31) >> 5;
            for (uint32_t w = 0; w < words; ++w) sm_mask[e * maskw + w] = 0u;
            if (Mok) {
                // window of first depths that can pass, widened beyond every rounding of the test and
                // of the key (256 ulps = 1.6e-5 relative)
                const float slack = fmaf(1.0e-4f, t1, 1.0e-4f * fabsf(M.d_p1));
                const float flo = fmaxf(M.d_p1 - t1 - slack, 0.0f), fhi = M.d_p1 + t1 + slack;
                const uint32_t klo = __float_as_uint(flo) & 0xffffff00u;
                const uint32_t khi = (fhi >= 0.0f ? __float_as_uint(fhi) : 0u) | 0xffu;  // +inf: everything
                uint32_t lo = 0, hi = m;  // first sorted position with key >= klo
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (sm_skey[mid] < klo) lo = mid + 1; else hi = mid;
                }
                for (uint32_t r = lo; r < m && sm_skey[r] <= khi; ++r) {
                    const uint32_t j = sm_skey[r] & 0xffu;
                    if (pair_flagged(M, true, t1, t2, sm_sib[j])) {
                        sm_mask[e * maskw + (j >> 5)] |= 1u << (j & 31u);
                        ++c;
                    }
                }
            }
        } else if (Mok) {  // an invalid 3-D segment has similarity 0 with everything
#pragma unroll 4
            for (uint32_t j = 0; j < m; ++j) c += pair_flagged(M, Mok, t1, t2, sib[j]) ? 1u : 0u;
        }
        cnt[e] = c;
    }
    __threadfence_block();
    __syncthreads();
    if (tid < 32) {  // exclusive prefixes: pair offsets and head indices
        uint32_t run_c = 0, run_h = 0;
        for (uint32_t base = 0; base < m; base += 32) {
            const uint32_t e = base + lane;
            const uint32_t c = (e < m) ? cnt[e] : 0u;
            const uint32_t h = c ? 1u : 0u;
            uint32_t xc = c, xh = h;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t tc = __shfl_up_sync(0xffffffffu, xc, d), th = __shfl_up_sync(0xffffffffu, xh, d);
                if ((int)lane >= d) { xc += tc; xh += th; }
            }
            if (e < m) {
                cnt[e] = run_c + xc - c;
                hof[e] = h ? (run_h + xh - h) : NOIDX;
            }
            run_c += __shfl_sync(0xffffffffu, xc, 31);
            run_h += __shfl_sync(0xffffffffu, xh, 31);
        }
        if (lane == 0) {
            cnt[m] = run_c;
            s_tot[0] = run_c;
            s_tot[1] = run_h;
            uint32_t base = NOIDX;
            if (run_c) {
                base = atomicAdd(&a.stats->prog_cursor, 1u + run_c + run_h);
                if ((uint64_t)base + 1u + run_c + run_h > a.prog_cap) {
                    atomicOr(&a.stats->err, 4u);
                    base = NOIDX;
                }
            }
            s_base = base;
            a.prog_off[g] = base;
            a.prog_nh[g] = (base == NOIDX) ? 0u : run_h;
        }
    }
    __threadfence_block();
    __syncthreads();
    const uint32_t T = s_tot[0], NH = s_tot[1];
    if (T == 0 || s_base == NOIDX) return;  // uniform
    // program of the row: header {heads, siblings, segment}, head records, sibling records
    if (tid == 0) a.prog[s_base] = make_uint4(NH, T, g, 0u);
    uint4* __restrict__ heads = a.prog + s_base + 1;
    uint4* __restrict__ prs = heads + NH;

    // ---- emit: head + sibling records of every match with flagged siblings ----
    for (uint32_t e = tid; e < m; e += DF_THREADS) {
        const uint32_t start = cnt[e], c = cnt[e + 1] - start;
        if (!c) continue;
        const uint32_t f = Lf[e];
        const uint32_t inv = Lm[e] >> 7;
        heads[hof[e]] = make_uint4(e | (inv << 31), start, c, f);
        if (!inv) a.fwd_score[f] = -1.0f;  // pending: the fold kernel publishes the score
        uint32_t k = start;
        if (use_mask) {
            const uint32_t words = (m + 31) >> 5;
            for (uint32_t w = 0; w < words; ++w) {
                uint32_t bits = sm_mask[e * maskw + w];
                while (bits) {
                    const uint32_t j = (w << 5) + (__ffs(bits) - 1);
                    bits &= bits - 1;
                    const unsigned char mj = Lm[j];
                    prs[k++] = make_uint4((mj >> 7) ? Lf[j] : NOIDX, 0u, e, j | ((uint32_t)(mj & 63u) << 24));
                }
            }
        } else {
            const Sib M = sib[e];
            const float2 rg = regs[e];
            const float t1 = reject_threshold(rg.x), t2 = reject_threshold(rg.y);
            for (uint32_t j = 0; j < m; ++j)
                if (pair_flagged(M, true, t1, t2, sib[j])) {
                    const unsigned char mj = Lm[j];
                    prs[k++] = make_uint4((mj >> 7) ? Lf[j] : NOIDX, 0u, e, j | ((uint32_t)(mj & 63u) << 24));
                }
        }
    }
    __threadfence_block();
    __syncthreads();

    // ---- the flagged pairs, one per thread: full similarity ----
    for (uint32_t t = tid; t < T; t += DF_THREADS) {
        const uint4 pr = prs[t];
        const uint32_t e = pr.z, j = pr.w & 0xffffffu;
        const Sib M = sib[e];
        const float2 rg = regs[e];
        const D3 dirM = d3(dirs[3 * e], dirs[3 * e + 1], dirs[3 * e + 2]);
        const float sim = sim_for_scoring(M.d_p1, M.d_p2, rg.x, rg.y, true, dirM, sib[j], dirs + 3 * j, a.two_sigA_sqr,
                                          0.5f, -0.70f, 0.5f, a.dotcut);
        prs[t].y = __float_as_uint(sim);
    }
}

// rows by length class (the launcher gives every class its own staging size and CTA width); empty rows
// are closed here
__global__ void __launch_bounds__(256) k3_classify_kernel(const uint32_t* __restrict__ L_off, uint32_t g_lo,
                                                          uint32_t g_hi, uint32_t b0, uint32_t b1,
                                                          uint32_t* __restrict__ cls_cnt, uint32_t* __restrict__ cls_rows,
                                                          uint32_t* __restrict__ prog_nh)
{
    const uint32_t g = g_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, rows = g_hi - g_lo;
    int cls = -1;
    if (g < g_hi) {
        const uint32_t len = L_off[g + 1] - L_off[g];
        if (len == 0) prog_nh[g] = 0u;
        else cls = len <= b0 ? 0 : (len <= b1 ? 1 : 2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t bal = __ballot_sync(0xffffffffu, cls == c);
        if (!bal) continue;
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(bal) - 1)) base = atomicAdd(&cls_cnt[c], (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (cls == c) cls_rows[(size_t)c * rows + base + __popc(bal & ((1u << lane) - 1u))] = g;
    }
}

// persistent CTAs: CTA b takes the rows b, b + gridDim.x, ... of its class
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k3_build_kernel(const BuildArgs a, const uint32_t* __restrict__ cls_cnt,
                                                            const uint32_t* __restrict__ cls_rows)
{
    extern __shared__ __align__(16) unsigned char df_smem[];
    __shared__ BlockTab bt;
    __shared__ uint32_t s_tot[2], s_base;
    const uint32_t n = *cls_cnt;
    for (uint32_t x = blockIdx.x; x < n; x += gridDim.x) {
        build_row<NT>(a, cls_rows[x], df_smem, bt, s_tot, s_base);
        __syncthreads();  // the shared tables are reused by the next row
    }
}

// ------------------------------------------------------------------------------------------
// fold: scoringCPU's accumulation over the present siblings, in dependency order
// ------------------------------------------------------------------------------------------
struct FoldArgs {
    const uint32_t* seg_view;
    const uint32_t* L_off;
    float* L_score;  // per potential entry: score3D_ of the entry as M (0 unless folded)
    float* fwd_score;
    const uint32_t* prog_off;
    const uint32_t* prog_nh;
    const uint4* prog;
    uint32_t* view_max;  // [V] ordered-uint maximum score of the view
    WfStats* stats;
    uint32_t S;
    uint32_t g_lo, g_hi;  // L_score is kept for the rows this rank finishes
    uint32_t sleep_ns;    // back-off between two polls of a pending score
};

__device__ __forceinline__ float wait_score(float* fwd_score, uint32_t f, WfStats* stats, uint32_t sleep_ns)
{
    const volatile float* p = fwd_score + f;
    float s = *p;
    uint32_t spins = 0;
    while (s < 0.0f) {
        if (sleep_ns) __nanosleep(sleep_ns);
        s = *p;
        if (++spins > DF_SPIN_LIMIT) {  // cannot happen (see the header); never hang the device
            atomicOr(&stats->err, 8u);
            return 0.0f;
        }
    }
    return s;
}

static constexpr int FOLD_WARPS = 8;
static constexpr int FOLD_CAP = 128;  // program records of one row staged in shared memory

// the accumulation of src/line3D.cc:1515-1543 over the sibling records [first, first + n)
__device__ __forceinline__ float fold_siblings(const uint4* prs, uint32_t first, uint32_t n)
{
    float score = 0.0f, stored = 0.0f;
    uint32_t cur_run = NOIDX;
    for (uint32_t t = first; t < first + n; ++t) {
        const uint4 pr = prs[t];
        const float sim = __uint_as_float(pr.y);
        if (!(sim > 0.0f)) continue;  // similarity 0 or sibling absent
        const uint32_t run = pr.w >> 24;
        if (run != cur_run) {  // first sibling of this camera (src/line3D.cc:1536-1540)
            score = fa(score, sim);
            stored = sim;
            cur_run = run;
        } else if (sim > stored) {  // src/line3D.cc:1527-1534
            score = fs(score, stored);
            score = fa(score, sim);
            stored = sim;
        }
    }
    return score;
}

__global__ void __launch_bounds__(FOLD_WARPS * 32) k3_fold_kernel(const FoldArgs a)
{
    __shared__ uint4 sprog[FOLD_WARPS][FOLD_CAP];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* sp = sprog[warp];
    // Rows are dealt round-robin to the resident warps, each warp takes its rows in ascending order.
    // The smallest unfinished row of the grid is then always the current row of its warp and all the
    // rows it waits on are finished: it never blocks, so the grid always makes progress.
    const uint32_t n_warps = gridDim.x * FOLD_WARPS;
    for (uint32_t g = blockIdx.x * FOLD_WARPS + warp; g < a.S; g += n_warps) {
        const uint32_t NH = a.prog_nh[g];
        if (NH == 0) continue;  // warp-uniform
        {
            const uint4* __restrict__ prog = a.prog + a.prog_off[g];
            const size_t lbase = a.L_off[g];
            const uint32_t view = a.seg_view[g];
            const bool mine = g >= a.g_lo && g < a.g_hi;
            const uint32_t T = prog[0].y;
            const uint4* __restrict__ heads = prog + 1;
            const uint4* __restrict__ prs = heads + NH;
            float wmax = 0.0f;
            if (NH + T <= (uint32_t)FOLD_CAP) {
                // stage the program, resolve every presence in parallel, then fold from shared memory:
                // the time between the last dependency and the published score stays short
                __syncwarp();
                for (uint32_t x = lane; x < NH + T; x += 32) {
                    uint4 r = heads[x];
                    if (x < NH) {
                        // an inverse match exists iff its forward record scored > 0 (src/line3D.cc:1994-1996)
                        if ((r.x >> 31) && !(wait_score(a.fwd_score, r.w, a.stats, a.sleep_ns) > 0.0f)) r.z = 0xffffffffu;
                    } else if (__uint_as_float(r.y) > 0.0f && r.x != NOIDX) {
                        if (!(wait_score(a.fwd_score, r.x, a.stats, a.sleep_ns) > 0.0f)) r.y = 0u;
                    }
                    sp[x] = r;
                }
                __syncwarp();
                for (uint32_t h = lane; h < NH; h += 32) {
                    const uint4 H = sp[h];
                    if (H.z == 0xffffffffu) continue;  // absent inverse match
                    const float score = fold_siblings(sp + NH, H.y, H.z);
                    if (!(H.x >> 31)) *(volatile float*)(a.fwd_score + H.w) = score;
                    if (mine) a.L_score[lbase + (H.x & 0x7fffffffu)] = score;
                    wmax = fmaxf(wmax, score);
                }
            } else {
                for (uint32_t h = lane; h < NH; h += 32) {
                    const uint4 H = heads[h];
                    const uint32_t e = H.x & 0x7fffffffu, inv = H.x >> 31;
                    if (inv && !(wait_score(a.fwd_score, H.w, a.stats, a.sleep_ns) > 0.0f)) continue;
                    float score = 0.0f, stored = 0.0f;
                    uint32_t cur_run = NOIDX;
                    for (uint32_t t = H.y; t < H.y + H.z; ++t) {
                        const uint4 pr = prs[t];
                        const float sim = __uint_as_float(pr.y);
                        if (!(sim > 0.0f)) continue;
                        if (pr.x != NOIDX && !(wait_score(a.fwd_score, pr.x, a.stats, a.sleep_ns) > 0.0f)) continue;
                        const uint32_t run = pr.w >> 24;
                        if (run != cur_run) {
                            score = fa(score, sim);
                            stored = sim;
                            cur_run = run;
                        } else if (sim > stored) {
                            score = fs(score, stored);
                            score = fa(score, sim);
                            stored = sim;
                        }
                    }
                    if (!inv) *(volatile float*)(a.fwd_score + H.w) = score;
                    if (mine) a.L_score[lbase + e] = score;
                    wmax = fmaxf(wmax, score);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
            if (lane == 0 && wmax > 0.0f) atomicMax(&a.view_max[view], float_ordered(wmax));
        }
    }
}

// ------------------------------------------------------------------------------------------
// finish: existing entries, scored lists, filterMatches, estimated_position3D_
// ------------------------------------------------------------------------------------------
struct FinishArgs {
    const ViewDev* views;
    const uint32_t* seg_view;
    const PairDev* pairs;
    const IncDev* inc;
    const uint32_t* inc_off;
    const SegRays* rays;
    const FwdRec* fwd_rec;
    const float* fwd_score;
    const uint32_t* fwd_row;
    const uint32_t* L_off;
    const uint32_t* L_f;
    const unsigned char* L_meta;
    const float* L_score;
    uint32_t* L_cnt;   // [S] actual list lengths
    ListRec* L_rec;    // scored lists (NULL unless keep_scored), row g at L_off[g]
    const uint32_t* view_max;
    ListRec* filt_rec;
    uint32_t filt_cap;
    uint32_t* filt_off;  // [S]
    uint32_t* filt_cnt;  // [S]
    EntryDev* entries;   // [S]
    WfStats* stats;
    uint32_t S;
    uint32_t g_lo, g_hi;  // rows finished by this rank
};

__device__ __forceinline__ ListRec make_list_rec(const FinishArgs& a, uint32_t i0, uint32_t f, uint32_t meta, float score)
{
    const FwdRec r = a.fwd_rec[f];
    const PairDev& P = a.pairs[a.inc[i0 + (meta & 63u)].pair];
    ListRec L;
    L.overlap = r.overlap;
    L.score = score;
    if (meta >> 7) {  // storeInverseMatches: views and depths swapped, orientation flag set
        L.tgt_view = P.src_view;
        L.tgt_seg = a.fwd_row[f] - P.row_base;
        L.d_p1 = r.d_q1;
        L.d_p2 = r.d_q2;
        L.d_q1 = r.d_p1;
        L.d_q2 = r.d_p2;
        L.flags = 3u;
        L.src_idx = NOIDX;
    } else {
        L.tgt_view = P.tgt_view;
        L.tgt_seg = r.c;
        L.d_p1 = r.d_p1;
        L.d_p2 = r.d_p2;
        L.d_q1 = r.d_q1;
        L.d_q2 = r.d_q2;
        L.flags = 0u;
        L.src_idx = f;
    }
    return L;
}

static constexpr int FIN_WARPS = 8;

__global__ void __launch_bounds__(FIN_WARPS * 32, 6) k3_finish_kernel(const FinishArgs a)
{
    __shared__ uint32_t blkcnt[FIN_WARPS][DF_MAXINC];
    // the rows of a CTA reserve their filtered entries and add their counters with ONE atomic per CTA and
    // address (one per row serialised 50 000 same-address atomics at C2 size)
    __shared__ uint32_t s_kept[FIN_WARPS], s_valid[FIN_WARPS], s_base;
    __shared__ unsigned long long s_scored[FIN_WARPS], s_evals[FIN_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = a.g_lo + blockIdx.x * FIN_WARPS + warp;
    const bool in_range = g < a.g_hi;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const size_t lbase = in_range ? a.L_off[g] : 0;
    const uint32_t m = in_range ? a.L_off[g + 1] - a.L_off[g] : 0u;
    const uint32_t v = in_range ? a.seg_view[g] : 0u;
    const ViewDev& va = a.views[v];
    const uint32_t i0 = a.inc_off[v];
    if (in_range && m == 0 && lane == 0) {
        a.L_cnt[g] = 0u;
        a.filt_off[g] = 0u;
        a.filt_cnt[g] = 0u;
        a.entries[g].has = 0u;
    }
    blkcnt[warp][lane] = 0u;
    blkcnt[warp][lane + 32] = 0u;
    __syncwarp();
    const float max_score = fmaxf(0.0f, ordered_to_float(a.view_max[v]));
    const float lim = fm(0.10f, max_score);
    uint32_t n_present = 0, kept = 0;
    float best = 0.0f;
    uint32_t best_idx = NOIDX;
    bool any_valid = false;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t x = base + lane;
        bool present = false, keep = false;
        float s = 0.0f;
        uint32_t f = 0, meta = 0;
        if (x < m) {
            f = a.L_f[lbase + x];
            meta = a.L_meta[lbase + x];
            present = !(meta >> 7) || a.fwd_score[f] > 0.0f;
            if (present) {
                s = a.L_score[lbase + x];
                keep = (s > 0.0f) && (s > lim);
                any_valid |= (s > 0.75f);
                atomicAdd(&blkcnt[warp][meta & 63u], 1u);
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, present);
        if (a.L_rec && present) a.L_rec[lbase + n_present + __popc(bal & lt_mask)] = make_list_rec(a, i0, f, meta, s);
        n_present += __popc(bal);
        kept += __popc(__ballot_sync(0xffffffffu, keep));
        float cs = keep ? s : 0.0f;
        uint32_t ci = keep ? x : NOIDX;
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
    __syncwarp();
    // sibling pairs the reference visits: present entries of other cameras
    unsigned long long sq = (unsigned long long)blkcnt[warp][lane] * blkcnt[warp][lane] +
                            (unsigned long long)blkcnt[warp][lane + 32] * blkcnt[warp][lane + 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    const bool row_valid = __any_sync(0xffffffffu, any_valid);

    if (lane == 0) {
        s_kept[warp] = kept;
        s_scored[warp] = n_present;
        s_evals[warp] = (unsigned long long)n_present * n_present - sq;
        s_valid[warp] = row_valid ? 1u : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tk = 0, tv = 0;
        unsigned long long ts = 0, te = 0;
#pragma unroll
        for (int w2 = 0; w2 < FIN_WARPS; ++w2) {
            tk += s_kept[w2];
            tv += s_valid[w2];
            ts += s_scored[w2];
            te += s_evals[w2];
        }
        s_base = tk ? atomicAdd(&a.stats->filt_cursor, tk) : 0u;
        if (ts) atomicAdd(&a.stats->scored, ts);
        if (te) atomicAdd(&a.stats->sim_evals, te);
        if (tv) atomicAdd(&a.stats->num_valid, tv);
    }
    __syncthreads();
    if (!in_range || m == 0) return;
    uint32_t dst0 = s_base;
    for (uint32_t w2 = 0; w2 < warp; ++w2) dst0 += s_kept[w2];
    const bool fits = (kept == 0) || ((uint64_t)dst0 + kept <= a.filt_cap);
    if (!fits && lane == 0) atomicOr(&a.stats->err, 2u);
    uint32_t w = 0;
    if (kept && fits)
        for (uint32_t base = 0; base < m; base += 32) {
            const uint32_t x = base + lane;
            bool keep = false;
            float s = 0.0f;
            uint32_t f = 0, meta = 0;
            if (x < m) {
                f = a.L_f[lbase + x];
                meta = a.L_meta[lbase + x];
                if (!(meta >> 7) || a.fwd_score[f] > 0.0f) {
                    s = a.L_score[lbase + x];
                    keep = (s > 0.0f) && (s > lim);
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) a.filt_rec[dst0 + w + __popc(bal & lt_mask)] = make_list_rec(a, i0, f, meta, s);
            w += __popc(bal);
        }
    if (lane == 0) {
        a.L_cnt[g] = n_present;
        a.filt_off[g] = dst0;
        a.filt_cnt[g] = fits ? kept : 0u;
        EntryDev& E = a.entries[g];
        if (best_idx != NOIDX && best > 0.75f) {
            const ListRec B = make_list_rec(a, i0, a.L_f[lbase + best_idx], a.L_meta[lbase + best_idx], best);
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

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_k3_inv_capacity(const PairDev* pairs, uint32_t P, uint32_t n_rows, const uint32_t* fwd_off,
                           const uint32_t* fwd_cnt, const FwdRec* fwd_rec, uint32_t* fwd_row, uint32_t* inv_cap,
                           uint32_t v_lo, uint32_t v_hi, cudaStream_t st)
{
    if (!n_rows || !P) return 0;
    k3_inv_capacity_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(pairs, P, n_rows, fwd_off, fwd_cnt, fwd_rec, fwd_row,
                                                                  inv_cap, v_lo, v_hi);
    return 1;
}

int launch_k3_list_capacity(const ViewDev* views, const uint32_t* seg_view, uint32_t S, const IncDev* inc,
                            const uint32_t* inc_off, const PairDev* pairs, const uint32_t* fwd_cnt,
                            const uint32_t* inv_cap, uint32_t* L_ub, void* stats, cudaStream_t st)
{
    if (!S) return 0;
    k3_list_capacity_kernel<<<(S + 255) / 256, 256, 0, st>>>(views, seg_view, S, inc, inc_off, pairs, fwd_cnt, inv_cap,
                                                              L_ub, (WfStats*)stats);
    return 1;
}

int launch_k3_records(const PairDev* pairs, uint32_t P, uint32_t F, const uint32_t* fwd_row, const FwdRec* fwd_rec,
                      float* fwd_score, const uint32_t* inv_off, uint32_t* inv_fill, uint2* inv_ent, uint32_t v_lo, uint32_t v_hi,
                      cudaStream_t st)
{
    if (!F || !P) return 0;
    k3_record_kernel<<<(F + 255) / 256, 256, 0, st>>>(pairs, P, F, fwd_row, fwd_rec, fwd_score, inv_off, inv_fill,
                                                       inv_ent, v_lo, v_hi);
    return 1;
}
size_t k3_wf_stats_bytes() { return sizeof(WfStats); }
size_t k3_sib_bytes() { return sizeof(Sib); }
int k3_wf_max_inc() { return DF_MAXINC; }
int k3_max_staged() { return DF_MAXM_CAP; }

static size_t build_smem_bytes(uint32_t maxm, uint32_t nt)
{
    const size_t maskm = maxm < (uint32_t)DF_MASKM ? maxm : (uint32_t)DF_MASKM, maskw = (maskm + 31) / 32;
    return (size_t)maxm * (24 + sizeof(Sib) + sizeof(float2)) + 2 * ((size_t)maxm + 2) * 4 +
           (((maskm * maskw + 3) & ~(size_t)3) + 2 * ((maskm + 3) & ~(size_t)3) + 2 * (size_t)nt) * 4 + 16;
}

static BuildArgs build_args(const K3Tables& t, uint32_t maxm)
{
    BuildArgs b;
    b.views = t.views; b.seg_view = t.seg_view; b.pairs = t.pairs; b.inc = t.inc; b.inc_off = t.inc_off;
    b.fwd_off = t.fwd_off; b.fwd_cnt = t.fwd_cnt; b.fwd_rec = t.fwd_rec; b.fwd_score = t.fwd_score;
    b.rays = t.rays;
    b.inv_off = t.inv_off; b.inv_fill = t.inv_fill; b.inv_ent = t.inv_ent;
    b.L_off = t.L_off; b.L_f = t.L_f; b.L_meta = t.L_meta;
    b.L_sib = (Sib*)t.L_sib; b.L_dir = t.L_dir; b.L_reg = t.L_reg; b.L_c = t.L_c; b.L_h = t.L_h;
    b.prog_off = t.prog_off; b.prog_nh = t.prog_nh; b.prog = (uint4*)t.prog; b.prog_cap = t.prog_cap;
    b.stats = (WfStats*)t.stats; b.S = t.S; b.g_lo = t.g_lo; b.g_hi = t.g_hi; b.maxm = maxm;
    b.two_sigA_sqr = t.two_sigA_sqr;
    b.dotcut = score_dotcut(t.two_sigA_sqr, 0.5f);
    return b;
}

// build: potential lists and fold programs of the rows [g_lo, g_hi).  The rows are sorted into three length
// classes; every class is built by persistent CTAs with its own staging size and width: the many short
// rows get a small shared-memory staging and 64 threads (more rows in flight per SM: a row is a chain of
// dependent gathers), the long ones more threads per row (their per-entry loops are the critical path).
template <int NT, int MINB>
static int launch_build_class(const K3Tables& t, uint32_t mm, uint32_t max_smem_m, int cls, uint32_t rows, int n_sm,
                              cudaStream_t st, int* err)
{
    const BuildArgs b = build_args(t, mm);
    const size_t smem = build_smem_bytes(mm, NT);
    cudaError_t e = cudaFuncSetAttribute(k3_build_kernel<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)build_smem_bytes(max_smem_m, NT));
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_build_kernel<NT, MINB>, NT, smem);
    if (e != cudaSuccess || per_sm < 1) {
        *err = (int)(e != cudaSuccess ? e : cudaErrorLaunchOutOfResources);
        return -1;
    }
    uint32_t grid = (uint32_t)(n_sm * per_sm);
    if (grid > rows) grid = rows;
    k3_build_kernel<NT, MINB><<<grid, NT, smem, st>>>(b, t.cls + cls, t.cls + 16 + (size_t)cls * rows);
    return 1;
}

size_t k3_class_words(uint32_t rows) { return 16 + 3 * (size_t)rows; }

int launch_k3_build(const K3Tables& t, cudaStream_t st, int* err)
{
    if (t.g_hi <= t.g_lo) return 0;
    uint32_t maxm = t.maxm;
    if (maxm > (uint32_t)DF_MAXM_CAP) maxm = DF_MAXM_CAP;
    maxm = (maxm + 3u) & ~3u;  // keeps the shared-memory arrays 16-byte aligned
    if (maxm < 4) maxm = 4;
    static int b0 = -1, b1 = -1, n_sm = 0;
    if (b0 < 0) {
        const char* e0 = getenv("L3D_K3_B0");  // tuning hooks: the class bounds
        const char* e1 = getenv("L3D_K3_B1");
        b0 = e0 ? atoi(e0) : DF_SMALL;
        b1 = e1 ? atoi(e1) : DF_MASKM;
        b0 = (std::max(b0, 4) + 3) & ~3;
        b1 = (std::max(b1, b0) + 3) & ~3;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint32_t rows = t.g_hi - t.g_lo;
    const uint32_t m0 = std::min<uint32_t>(maxm, (uint32_t)b0), m1 = std::min<uint32_t>(maxm, (uint32_t)b1);
    // a row longer than the staging of the last class is built in the global staging (t.L_sib)
    const bool only0 = t.maxm <= m0 && t.L_sib == nullptr, only01 = t.maxm <= m1 && t.L_sib == nullptr;
    cudaMemsetAsync(t.cls, 0, 16 * sizeof(uint32_t), st);
    k3_classify_kernel<<<(rows + 255) / 256, 256, 0, st>>>(t.L_off, t.g_lo, t.g_hi, only0 ? 0xffffffffu : m0,
                                                           only01 ? 0xffffffffu : m1, t.cls, t.cls + 16, t.prog_nh);
    int launches = 1, r;
    if ((r = launch_build_class<64, 16>(t, m0, maxm, 0, rows, n_sm, st, err)) < 0) return -1;
    launches += r;
    if (!only0) {
        if ((r = launch_build_class<128, 8>(t, m1, maxm, 1, rows, n_sm, st, err)) < 0) return -1;
        launches += r;
    }
    if (!only01) {
        if ((r = launch_build_class<256, 4>(t, maxm, maxm, 2, rows, n_sm, st, err)) < 0) return -1;
        launches += r;
    }
    return launches;
}

// fold: every row of the scene (replicated on every rank: the scores of all views are needed)
int launch_k3_fold(const K3Tables& t, cudaStream_t st, int* err)
{
    if (!t.S) return 0;
    FoldArgs f;
    f.seg_view = t.seg_view; f.L_off = t.L_off; f.L_score = t.L_score; f.fwd_score = t.fwd_score;
    f.prog_off = t.prog_off; f.prog_nh = t.prog_nh; f.prog = (const uint4*)t.prog; f.view_max = t.view_max;
    f.stats = (WfStats*)t.stats; f.S = t.S; f.g_lo = t.g_lo; f.g_hi = t.g_hi;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_fold_kernel, FOLD_WARPS * 32, 0);
    if (e != cudaSuccess || per_sm < 1) {
        *err = (int)e;
        return -1;
    }
    // every CTA must be resident (the progress argument needs running warps): at most one wave
    uint32_t grid = (uint32_t)(sms * per_sm);
    const uint32_t want = (t.S + FOLD_WARPS - 1) / FOLD_WARPS;
    if (want < grid) grid = want ? want : 1u;
    static int sleep_ns = -1;
    if (sleep_ns < 0) {
        const char* ev = getenv("L3D_FOLD_SLEEP_NS");  // tuning hook
        sleep_ns = ev ? atoi(ev) : 40;
    }
    f.sleep_ns = (uint32_t)sleep_ns;
    // cooperative launch: the runtime guarantees that all CTAs are co-resident (or fails), which the
    // progress argument of the fold relies on even if another stream is using part of the device
    void* args[] = {(void*)&f};
    e = cudaLaunchCooperativeKernel((void*)k3_fold_kernel, dim3(grid), dim3(FOLD_WARPS * 32), args, 0, st);
    if (e != cudaSuccess) {
        *err = (int)e;
        return -1;
    }
    return 1;
}

// finish: lists, filterMatches and hypotheses of the rows [g_lo, g_hi)
int launch_k3_finish(const K3Tables& t, cudaStream_t st)
{
    if (t.g_hi <= t.g_lo) return 0;
    FinishArgs c;
    c.views = t.views; c.seg_view = t.seg_view; c.pairs = t.pairs; c.inc = t.inc; c.inc_off = t.inc_off;
    c.rays = t.rays; c.fwd_rec = t.fwd_rec; c.fwd_score = t.fwd_score; c.fwd_row = t.fwd_row; c.L_off = t.L_off; c.L_f = t.L_f;
    c.L_meta = t.L_meta; c.L_score = t.L_score; c.L_cnt = t.L_cnt; c.L_rec = t.L_rec; c.view_max = t.view_max;
    c.filt_rec = t.filt_rec; c.filt_cap = t.filt_cap; c.filt_off = t.filt_off; c.filt_cnt = t.filt_cnt;
    c.entries = t.entries; c.stats = (WfStats*)t.stats; c.S = t.S; c.g_lo = t.g_lo; c.g_hi = t.g_hi;
    k3_finish_kernel<<<(t.g_hi - t.g_lo + FIN_WARPS - 1) / FIN_WARPS, FIN_WARPS * 32, 0, st>>>(c);
    return 1;
}

// multi-GPU: adopt the all-gathered fold programs.  Blob of rank q (stride bytes apart):
//   u32 nh[rows_pad_q] | u32 off[rows_pad_q] | uint4 records[]   (rows_pad = rows of the slice rounded up to 4)
// prog_off becomes an index into the gathered buffer; the forward heads are marked pending.
__global__ void __launch_bounds__(256) k3_adopt_programs_kernel(const unsigned char* __restrict__ all, uint64_t stride,
                                                                int world, const uint32_t* __restrict__ slice_g,
                                                                uint32_t S, uint32_t* __restrict__ prog_off,
                                                                uint32_t* __restrict__ prog_nh,
                                                                float* __restrict__ fwd_score)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= S) return;
    int q = 0;
    while (q + 1 < world && slice_g[q + 1] <= g) ++q;
    const uint32_t rows = slice_g[q + 1] - slice_g[q], rows_pad = (rows + 3u) & ~3u;
    const unsigned char* blob = all + (uint64_t)q * stride;
    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(blob);
    const uint32_t nh = hdr[g - slice_g[q]];
    const uint32_t off = hdr[rows_pad + (g - slice_g[q])];
    const uint64_t rec0 = ((uint64_t)q * stride + 8ull * rows_pad) / 16ull;  // first record of the blob, 16-byte units
    prog_nh[g] = nh;
    prog_off[g] = nh ? (uint32_t)(rec0 + off) : 0u;
    if (nh) {
        const uint4* heads = reinterpret_cast<const uint4*>(all) + rec0 + off + 1;
        for (uint32_t h = 0; h < nh; ++h) {
            const uint4 H = heads[h];
            if (!(H.x >> 31)) fwd_score[H.w] = -1.0f;
        }
    }
}

int launch_k3_adopt_programs(const void* all, uint64_t stride, int world, const uint32_t* slice_g, uint32_t S,
                             uint32_t* prog_off, uint32_t* prog_nh, float* fwd_score, cudaStream_t st)
{
    if (!S) return 0;
    k3_adopt_programs_kernel<<<(S + 255) / 256, 256, 0, st>>>((const unsigned char*)all, stride, world, slice_g, S,
                                                               prog_off, prog_nh, fwd_score);
    return 1;
}

}  // namespace l3d
