// k4_affinity.cu -- K4: per-view median depths, the sparse affinity matrix A_ and its first-touch
// local IDs (exact TU).
//
// Replaces Line3D::computingAffinityMatrix (src/line3D.cc:2275-2402) with
// Line3D::similarity(Segment3D,Match,Segment2D,bool) (src/line3D.cc:1737-1823),
// Segment3D::distance_Point2Line (include/segment3D.h:80-84), Line3D::unused
// (src/line3D.cc:2405-2425) and Line3D::getLocalID (src/line3D.cc:2428-2446).
//
// The reference walks estimated_position3D_ serially; an edge (i -> j) is "unused" unless the
// unordered pair was already inserted, which can only have happened from entry j < i whose own
// filtered list holds i (the similarity is bitwise symmetric), so the de-duplication is a lookup
// in j's list.  Local IDs are handed out at first touch in traversal order: with kept edge k
// touching its source at position 2k and its target at 2k+1, the ID of a segment is the rank of
// its smallest touch position among all first touches (atomicMin + prefix sum).
#include "detmath.cuh"
#include "exact.cuh"
#include "internal.h"

namespace l3d {

#define L3D_EPS 1e-12

__global__ void __launch_bounds__(256) k4_has_kernel(const EntryDev* __restrict__ entries, uint32_t S,
                                                     uint32_t* __restrict__ has)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) has[i] = entries[i].has;
}

// View::update_median_depth input (src/line3D.cc:1964-1982): median of the best matches' depths
static constexpr int MED_CAP = 16384;
__global__ void __launch_bounds__(1024) k4_median_kernel(ViewDev* __restrict__ views,
                                                         const EntryDev* __restrict__ entries,
                                                         uint32_t* __restrict__ overflow)
{
    extern __shared__ float sd[];
    __shared__ uint32_t cnt;
    ViewDev& v = views[blockIdx.x];
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < v.n_seg; i += blockDim.x) {
        const EntryDev& E = entries[v.seg_off + i];
        if (E.has) {
            const uint32_t p = atomicAdd(&cnt, 2u);
            if (p + 1 < MED_CAP) {
                sd[p] = E.d_p1;
                sd[p + 1] = E.d_p2;
            }
        }
    }
    __syncthreads();
    const uint32_t n = cnt;
    if (n > MED_CAP) {
        if (threadIdx.x == 0) atomicExch(overflow, 1u);
        return;
    }
    if (n == 0) {
        if (threadIdx.x == 0) v.median_depth = (float)L3D_EPS;
        return;
    }
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (uint32_t i = n + threadIdx.x; i < np2; i += blockDim.x) sd[i] = __int_as_float(0x7f800000);
    __syncthreads();
    for (uint32_t k = 2; k <= np2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const float a = sd[i], b = sd[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        sd[i] = b;
                        sd[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) v.median_depth = sd[n / 2];
}

__device__ __forceinline__ D3 ld3e(const double* p) { return D3{p[0], p[1], p[2]}; }

// include/segment3D.h:80-84, evaluated as (dir * diff^T) * dir
__device__ __forceinline__ float dist_point_line(const D3& P1, const D3& dir, const D3& P)
{
    const D3 d = sub3(P, P1);
    const double hx = da(da(dm(dm(dir.x, d.x), dir.x), dm(dm(dir.x, d.y), dir.y)), dm(dm(dir.x, d.z), dir.z));
    const double hy = da(da(dm(dm(dir.y, d.x), dir.x), dm(dm(dir.y, d.y), dir.y)), dm(dm(dir.y, d.z), dir.z));
    const double hz = da(da(dm(dm(dir.z, d.x), dir.x), dm(dm(dir.z, d.y), dir.y)), dm(dm(dir.z, d.z), dir.z));
    const D3 hp = d3(da(P1.x, hx), da(P1.y, hy), da(P1.z, hz));
    return (float)norm3(sub3(hp, P));
}

// returns the similarity if it exceeds 0.5, otherwise any value <= 0.5 (early exits are safe
// because the caller only tests sim > L3D_DEF_MIN_AFFINITY)
__device__ __forceinline__ float affinity_sim(const EntryDev& e1, const ViewDev& v1, const EntryDev& e2,
                                              const ViewDev& v2, float two_sigA_sqr, float msdl)
{
    if (e1.length < L3D_EPS || e2.length < L3D_EPS) return 0.0f;
    const D3 dir1 = ld3e(e1.dir), dir2 = ld3e(e2.dir);
    const float dot_p = (float)dot3(dir1, dir2);
    float angle = (float)dm(dd((double)det_acosf(fmaxf(fminf(dot_p, 1.0f), -1.0f)), L3D_PI), (double)180.0f);
    if (angle > 90.0f) angle = fs(180.0f, angle);
    const float xa = fd(fm(-angle, angle), two_sigA_sqr);
    if (xa < -0.70f) return 0.0f;
    const float sim_a = det_expf(xa);
    float cutoff1 = v1.median_depth, cutoff2 = v2.median_depth;
    if (msdl > L3D_EPS) {
        cutoff1 = fminf(cutoff1, msdl);
        cutoff2 = fminf(cutoff2, msdl);
    }
    const D3 A1 = ld3e(e1.P1), A2 = ld3e(e1.P2), B1 = ld3e(e2.P1), B2 = ld3e(e2.P2);
    const float d11 = dist_point_line(B1, dir2, A1);
    const float d12 = dist_point_line(B1, dir2, A2);
    const float d21 = dist_point_line(A1, dir1, B1);
    const float d22 = dist_point_line(A1, dir1, B2);
    const float sig11 = (e1.d_p1 > cutoff1) ? fm(cutoff1, v1.k) : fm(e1.d_p1, v1.k);
    const float sig12 = (e1.d_p2 > cutoff1) ? fm(cutoff1, v1.k) : fm(e1.d_p2, v1.k);
    const float sig21 = (e2.d_p1 > cutoff2) ? fm(cutoff2, v2.k) : fm(e2.d_p1, v2.k);
    const float sig22 = (e2.d_p2 > cutoff2) ? fm(cutoff2, v2.k) : fm(e2.d_p2, v2.k);
    const float reg11 = fm(fm(2.0f, sig11), sig11), reg12 = fm(fm(2.0f, sig12), sig12);
    const float reg21 = fm(fm(2.0f, sig21), sig21), reg22 = fm(fm(2.0f, sig22), sig22);
    const float x11 = fd(fm(-d11, d11), reg11), x12 = fd(fm(-d12, d12), reg12);
    const float x21 = fd(fm(-d21, d21), reg21), x22 = fd(fm(-d22, d22), reg22);
    if (x11 < -0.70f || x12 < -0.70f || x21 < -0.70f || x22 < -0.70f) return 0.0f;
    const float sim_p1 = fminf(det_expf(x11), det_expf(x12));
    const float sim_p2 = fminf(det_expf(x21), det_expf(x22));
    const float sim_p = fminf(sim_p1, sim_p2);
    return fminf(sim_a, sim_p);
}

// pass 1a: similarity of every filtered-list entry of every hypothesis + de-duplication, one thread
// per (row, list entry) -- a row has only a handful of entries, one thread per row left the SMs idle
__global__ void __launch_bounds__(128) k4_edges_sim_kernel(
    const ViewDev* __restrict__ views, const uint32_t* __restrict__ seg_view, const EntryDev* __restrict__ entries,
    const uint32_t* __restrict__ filt_off, const uint32_t* __restrict__ filt_cnt,
    const ListRec* __restrict__ filt_rec, float two_sigA_sqr, float msdl, float* __restrict__ filt_sim,
    unsigned long long* __restrict__ tests, uint32_t g_lo, uint32_t g_hi, uint32_t max_cnt)
{
    const uint32_t g = g_lo + blockIdx.x * (blockDim.x / 8) + threadIdx.x / 8;  // 8 threads share a row
    if (g >= g_hi) return;
    const uint32_t n = filt_cnt[g];
    if (!n) return;
    const uint32_t b = filt_off[g];
    const bool has = entries[g].has != 0u;
    if (has && (threadIdx.x & 7u) == 0) atomicAdd(tests, (unsigned long long)n);
    const uint32_t v1i = seg_view[g];
    const ViewDev& v1 = views[v1i];
    const uint32_t my_seg = g - v1.seg_off;
    for (uint32_t z = threadIdx.x & 7u; z < n; z += 8) {
        float sim = -1.0f;
        if (has) {
            const ListRec m2 = filt_rec[b + z];
            const ViewDev& v2 = views[m2.tgt_view];
            const uint32_t t = v2.seg_off + m2.tgt_seg;
            if (entries[t].has) {
                sim = affinity_sim(entries[g], v1, entries[t], v2, two_sigA_sqr, msdl);
                if (sim > 0.5f && t < g) {
                    // Line3D::unused: inserted before by entry t if its list holds this segment
                    const uint32_t tb = filt_off[t], tn = filt_cnt[t];
                    for (uint32_t y = 0; y < tn; ++y) {
                        const ListRec o = filt_rec[tb + y];
                        if (o.tgt_view == v1i && o.tgt_seg == my_seg) {
                            sim = -2.0f;
                            break;
                        }
                    }
                }
            }
        }
        filt_sim[b + z] = sim;
    }
    (void)max_cnt;
}

// pass 1b: edges kept per row
__global__ void __launch_bounds__(256) k4_edges_count_kernel(const uint32_t* __restrict__ filt_off,
                                                             const uint32_t* __restrict__ filt_cnt,
                                                             const float* __restrict__ filt_sim,
                                                             uint32_t* __restrict__ E_cnt, uint32_t g_lo, uint32_t g_hi)
{
    const uint32_t g = g_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_hi) return;
    const uint32_t n = filt_cnt[g], b = filt_off[g];
    uint32_t kept = 0;
    for (uint32_t z = 0; z < n; ++z) kept += filt_sim[b + z] > 0.5f ? 1u : 0u;
    E_cnt[g] = kept;
}

struct EdgeDev {
    uint32_t src, tgt;  // global segment ids
    float w;
};

// pass 2: edges in traversal order + first-touch positions
__global__ void __launch_bounds__(128) k4_edges_write_kernel(
    const ViewDev* __restrict__ views, uint32_t S, const uint32_t* __restrict__ filt_off,
    const uint32_t* __restrict__ filt_cnt, const ListRec* __restrict__ filt_rec, const float* __restrict__ filt_sim,
    const uint32_t* __restrict__ E_off, EdgeDev* __restrict__ edges, uint32_t g_lo, uint32_t g_hi)
{
    const uint32_t g = g_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_hi) return;
    uint32_t k = E_off[g];
    if (E_off[g + 1] == k) return;
    const uint32_t n = filt_cnt[g], b = filt_off[g];
    for (uint32_t z = 0; z < n; ++z) {
        const float sim = filt_sim[b + z];
        if (sim > 0.5f) {
            const ListRec m2 = filt_rec[b + z];
            const uint32_t t = views[m2.tgt_view].seg_off + m2.tgt_seg;
            edges[k] = EdgeDev{g, t, sim};
            ++k;
        }
    }
}

// first touch position of every segment: edge k touches its source at 2k and its target at 2k+1
__global__ void __launch_bounds__(256) k4_touch_kernel(const EdgeDev* __restrict__ edges, uint32_t n_edges,
                                                       uint32_t* __restrict__ first_touch)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edges) return;
    const EdgeDev e = edges[k];
    atomicMin(&first_touch[e.src], 2 * k);
    atomicMin(&first_touch[e.tgt], 2 * k + 1);
}

// flag[p] = 1 iff touch position p is the first touch of its segment
__global__ void __launch_bounds__(256) k4_touch_flags_kernel(const EdgeDev* __restrict__ edges, uint32_t n_edges,
                                                             const uint32_t* __restrict__ first_touch,
                                                             uint32_t* __restrict__ flags)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 2 * n_edges) return;
    const EdgeDev e = edges[p >> 1];
    const uint32_t s = (p & 1) ? e.tgt : e.src;
    flags[p] = (first_touch[s] == p) ? 1u : 0u;
}

// A_: (id1,id2,w),(id2,id1,w) per kept edge; local2global[id] = global segment
__global__ void __launch_bounds__(256) k4_emit_kernel(const EdgeDev* __restrict__ edges, uint32_t n_edges,
                                                      const uint32_t* __restrict__ first_touch,
                                                      const uint32_t* __restrict__ flag_scan, int2* __restrict__ A_ij,
                                                      float* __restrict__ A_w, uint32_t* __restrict__ local2global)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edges) return;
    const EdgeDev e = edges[k];
    const uint32_t fs_ = first_touch[e.src], ft = first_touch[e.tgt];
    const int id1 = (int)flag_scan[fs_], id2 = (int)flag_scan[ft];
    A_ij[2 * k] = make_int2(id1, id2);
    A_ij[2 * k + 1] = make_int2(id2, id1);
    A_w[2 * k] = e.w;
    A_w[2 * k + 1] = e.w;
    if (fs_ == 2 * k) local2global[id1] = e.src;
    if (ft == 2 * k + 1) local2global[id2] = e.tgt;
}

int launch_k4_has(const EntryDev* entries, uint32_t S, uint32_t* has, cudaStream_t st)
{
    if (!S) return 0;
    k4_has_kernel<<<(S + 255) / 256, 256, 0, st>>>(entries, S, has);
    return 1;
}

int launch_k4_median(ViewDev* views, uint32_t V, const EntryDev* entries, uint32_t* overflow, cudaStream_t st)
{
    if (!V) return 0;
    // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
    cudaFuncSetAttribute(k4_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MED_CAP * 4);
    k4_median_kernel<<<V, 1024, MED_CAP * 4, st>>>(views, entries, overflow);
    return 1;
}

int launch_k4_edges_count(const ViewDev* views, const uint32_t* seg_view, const EntryDev* entries, uint32_t S,
                          const uint32_t* filt_off, const uint32_t* filt_cnt, const ListRec* filt_rec,
                          float two_sigA_sqr, float msdl, float* filt_sim, uint32_t* E_cnt,
                          unsigned long long* tests, uint32_t g_lo, uint32_t g_hi, cudaStream_t st)
{
    (void)S;
    if (g_hi <= g_lo) return 0;
    const uint32_t rows = g_hi - g_lo;
    k4_edges_sim_kernel<<<(rows + 15) / 16, 128, 0, st>>>(views, seg_view, entries, filt_off, filt_cnt, filt_rec,
                                                            two_sigA_sqr, msdl, filt_sim, tests, g_lo, g_hi, 0u);
    k4_edges_count_kernel<<<(rows + 255) / 256, 256, 0, st>>>(filt_off, filt_cnt, filt_sim, E_cnt, g_lo, g_hi);
    return 2;
}

int launch_k4_edges_write(const ViewDev* views, uint32_t S, const uint32_t* filt_off, const uint32_t* filt_cnt,
                          const ListRec* filt_rec, const float* filt_sim, const uint32_t* E_off, void* edges,
                          uint32_t g_lo, uint32_t g_hi, cudaStream_t st)
{
    if (g_hi <= g_lo) return 0;
    k4_edges_write_kernel<<<(g_hi - g_lo + 127) / 128, 128, 0, st>>>(views, S, filt_off, filt_cnt, filt_rec, filt_sim,
                                                                      E_off, (EdgeDev*)edges, g_lo, g_hi);
    return 1;
}

int launch_k4_ids(const void* edges, uint32_t n_edges, uint32_t* first_touch, uint32_t* flags,
                  uint32_t* flag_scan, uint32_t* scan_scratch, size_t scan_words, int2* A_ij, float* A_w,
                  uint32_t* local2global, cudaStream_t st)
{
    if (!n_edges) return 0;
    int launches = 0;
    k4_touch_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>((const EdgeDev*)edges, n_edges, (uint32_t*)first_touch);
    ++launches;
    k4_touch_flags_kernel<<<(2 * n_edges + 255) / 256, 256, 0, st>>>((const EdgeDev*)edges, n_edges, first_touch,
                                                                      flags);
    ++launches;
    launches += launch_scan_u32(flags, flag_scan, 2 * n_edges, scan_scratch, scan_words, st);
    k4_emit_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>((const EdgeDev*)edges, n_edges, first_touch, flag_scan,
                                                           A_ij, A_w, local2global);
    ++launches;
    return launches;
}

size_t k4_edge_bytes() { return sizeof(EdgeDev); }

// ------------------------------------------------------------------------------------------
// SparseMatrix::SparseMatrix (src/sparsematrix.cc:8-61): A_ sorted by (column,row) or (row,column)
// (sortCLEdgesByCol / sortCLEdgesByRow, include/clustering.h:70-78), stored as float4{i,j,w/norm,0}
// with the start index of every column (row), -1 where it is empty.  A directed pair occurs once in
// A_ (Line3D::unused), so the order is total: counting sort by the primary key, then every key's
// handful of entries is ordered by the secondary key.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k4_sparse_hist_kernel(const int2* __restrict__ A_ij, uint32_t E, int by_row,
                                                             uint32_t* __restrict__ hist)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int2 ij = A_ij[e];
    atomicAdd(&hist[by_row ? ij.x : ij.y], 1u);
}

__global__ void __launch_bounds__(256) k4_sparse_scatter_kernel(const int2* __restrict__ A_ij,
                                                                const float* __restrict__ A_w, uint32_t E, int by_row,
                                                                const uint32_t* __restrict__ off,
                                                                uint32_t* __restrict__ fill, uint2* __restrict__ tmp)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int2 ij = A_ij[e];
    const uint32_t key = by_row ? ij.x : ij.y, other = by_row ? ij.y : ij.x;
    tmp[off[key] + atomicAdd(&fill[key], 1u)] = make_uint2(other, __float_as_uint(A_w[e]));
}

__global__ void __launch_bounds__(128) k4_sparse_finish_kernel(uint32_t n, const uint32_t* __restrict__ off,
                                                               uint2* __restrict__ tmp, float norm, int by_row,
                                                               float4* __restrict__ entries, int* __restrict__ start)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t b = off[k], m = off[k + 1] - b;
    start[k] = m ? (int)b : -1;
    for (uint32_t a = 1; a < m; ++a) {
        const uint2 x = tmp[b + a];
        uint32_t c = a;
        while (c > 0 && tmp[b + c - 1].x > x.x) {
            tmp[b + c] = tmp[b + c - 1];
            --c;
        }
        tmp[b + c] = x;
    }
    for (uint32_t a = 0; a < m; ++a) {
        const uint2 x = tmp[b + a];
        const float w = fd(__uint_as_float(x.y), norm);
        entries[b + a] = by_row ? make_float4((float)k, (float)x.x, w, 0.0f) : make_float4((float)x.x, (float)k, w, 0.0f);
    }
}

int launch_k4_sparse(const int2* A_ij, const float* A_w, uint32_t E, uint32_t n, int by_row, float norm, uint32_t* hist,
                     uint32_t* off, uint32_t* fill, uint2* tmp, uint32_t* scan_scratch, size_t scan_words,
                     float4* entries, int* start, cudaStream_t st)
{
    if (!E || !n) return 0;
    int launches = 0;
    cudaMemsetAsync(hist, 0, ((size_t)n + 1) * 4, st);
    cudaMemsetAsync(fill, 0, ((size_t)n + 1) * 4, st);
    k4_sparse_hist_kernel<<<(E + 255) / 256, 256, 0, st>>>(A_ij, E, by_row, hist);
    ++launches;
    launches += launch_scan_u32(hist, off, n, scan_scratch, scan_words, st);
    k4_sparse_scatter_kernel<<<(E + 255) / 256, 256, 0, st>>>(A_ij, A_w, E, by_row, off, fill, tmp);
    k4_sparse_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, off, tmp, norm, by_row, entries, start);
    return launches + 2;
}

}  // namespace l3d
