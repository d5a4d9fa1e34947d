// misc.cu -- device-math test hooks, the FP32 peak probe and the list preparation of
// l3d_score_matches (exact TU).
#include "detmath.cuh"
#include "exact.cuh"
#include "internal.h"

namespace l3d {

__global__ void test_expf_kernel(const float* __restrict__ x, float* __restrict__ y, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = det_expf(x[i]);
}
__global__ void test_acos_kernel(const double* __restrict__ x, double* __restrict__ y, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = det_acos(x[i]);
}

int launch_test_expf(const float* x, float* y, uint32_t n, cudaStream_t st)
{
    if (!n) return 0;
    test_expf_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, y, n);
    return 1;
}
int launch_test_acos(const double* x, double* y, uint32_t n, cudaStream_t st)
{
    if (!n) return 0;
    test_acos_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, y, n);
    return 1;
}

// 16 independent FFMA chains per thread: measures the FP32 pipe peak the K1 roofline divides by
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* __restrict__ sink, int iters)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
    const float b = 0.999f, c = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], b, c);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) sink[(blockIdx.x * blockDim.x + threadIdx.x) & ((1 << 20) - 1)] = s;
}
int launch_fp32_peak(float* sink, int blocks, int iters, cudaStream_t st)
{
    fp32_peak_kernel<<<blocks, 256, 0, st>>>(sink, iters);
    return 1;
}

// the same with DFMA: the FP64 pipe peak the K2 roofline divides by
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* __restrict__ sink, int iters)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-3 * (double)(threadIdx.x + i);
    const double b = 0.999, c = 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __fma_rn(a[i], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) sink[(blockIdx.x * blockDim.x + threadIdx.x) & ((1 << 19) - 1)] = s;
}
int launch_fp64_peak(double* sink, int blocks, int iters, cudaStream_t st)
{
    fp64_peak_kernel<<<blocks, 256, 0, st>>>(sink, iters);
    return 1;
}

// list preparation for l3d_score_matches: one thread per source segment walks its range
// (src/line3D.cc:1582-1623 packs the same data on the host)
__global__ void __launch_bounds__(128) score_prep_kernel(const float4* __restrict__ lines, uint32_t n_lines,
                                                         const float4* __restrict__ matches,
                                                         const float2* __restrict__ regs_tgt,
                                                         const double* __restrict__ RtKinv,
                                                         const double* __restrict__ C, float k,
                                                         const uint32_t* __restrict__ off,
                                                         ListRec* __restrict__ L_rec, ListGeo* __restrict__ L_geo)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const uint32_t b = off[i], e = off[i + 1];
    if (b == e) return;
    const D3 Cc = d3(C[0], C[1], C[2]);
    uint32_t prev = 0xffffffffu;
    for (uint32_t z = b; z < e; ++z) {
        const float4 m = matches[z];
        const uint32_t seg = (uint32_t)m.x;
        const uint32_t cam = (uint32_t)m.y;
        ListRec L;
        L.tgt_view = cam;
        L.tgt_seg = 0;
        L.overlap = 0.0f;
        L.score = 0.0f;
        L.d_p1 = m.z;
        L.d_p2 = m.w;
        L.d_q1 = L.d_q2 = 0.0f;
        L.flags = 0;
        L.src_idx = 0xffffffffu;
        ListGeo G;
        D3 dir = d3(0, 0, 0);
        float len = 0.0f;
        if (seg < n_lines) {
            const float4 sg = lines[seg];
            const D3 r1 = normalized3(mul33(RtKinv, d3((double)sg.x, (double)sg.y, 1.0)));
            const D3 r2 = normalized3(mul33(RtKinv, d3((double)sg.z, (double)sg.w, 1.0)));
            const D3 P1 = add3(Cc, scale3(r1, (double)m.z));
            const D3 P2 = add3(Cc, scale3(r2, (double)m.w));
            len = (float)norm3(sub3(P1, P2));
            if (len > 1e-12) dir = normalized3(sub3(P2, P1));
            else len = 0.0f;
        }
        const float sig1 = fm(m.z, k), sig2 = fm(m.w, k);
        const float2 rt = regs_tgt[z];
        G.reg1 = fm(0.5f, fa(fm(fm(2.0f, sig1), sig1), fm(fm(2.0f, rt.x), rt.x)));
        G.reg2 = fm(0.5f, fa(fm(fm(2.0f, sig2), sig2), fm(fm(2.0f, rt.y), rt.y)));
        G.dir[0] = dir.x; G.dir[1] = dir.y; G.dir[2] = dir.z;
        G.length = len;
        G.run = (cam != prev) ? 1u : 0u;
        G.pad0 = G.pad1 = 0;
        prev = cam;
        L_rec[z] = L;
        L_geo[z] = G;
    }
}

int launch_score_prep(const float4* lines, uint32_t n_lines, const float4* matches, const float2* regs_tgt,
                      const double* RtKinv, const double* C, float k, const uint32_t* off, ListRec* L_rec,
                      ListGeo* L_geo, cudaStream_t st)
{
    if (!n_lines) return 0;
    score_prep_kernel<<<(n_lines + 127) / 128, 128, 0, st>>>(lines, n_lines, matches, regs_tgt, RtKinv, C, k, off,
                                                              L_rec, L_geo);
    return 1;
}

// ---- multi-GPU FORWARD exchange (abi.cu): only the records of boundary pairs (target view in
// another rank's slice) travel; every rank gets all per-row counts ----
__device__ __forceinline__ uint32_t pair_of_row_m(const PairDev* __restrict__ pairs, uint32_t P, uint32_t row)
{
    uint32_t lo = 0, hi = P;  // largest p with row_base <= row
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pairs[mid].row_base <= row) lo = mid; else hi = mid;
    }
    return lo;
}
// bcnt[row - row_lo] = records of the row that travel
__global__ void __launch_bounds__(256) fwd_bmask_kernel(const PairDev* __restrict__ pairs, uint32_t P,
                                                        const uint32_t* __restrict__ fwd_cnt, uint32_t row_lo,
                                                        uint32_t row_hi, uint32_t* __restrict__ bcnt)
{
    const uint32_t row = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    const uint32_t n = fwd_cnt[row];
    bcnt[row - row_lo] = (n && pairs[pair_of_row_m(pairs, P, row)].xflag) ? n : 0u;
}
// sender: boundary records of the rows [row_lo, row_hi) (local layout) -> contiguous export buffer
__global__ void __launch_bounds__(256) fwd_bgather_kernel(const uint32_t* __restrict__ bcnt,
                                                          const uint32_t* __restrict__ boff,
                                                          const uint32_t* __restrict__ fwd_off_local, uint32_t row_lo,
                                                          uint32_t row_hi, const FwdRec* __restrict__ fwd_rec,
                                                          FwdRec* __restrict__ out)
{
    const uint32_t row = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    const uint32_t n = bcnt[row - row_lo];
    if (!n) return;
    const FwdRec* src = fwd_rec + fwd_off_local[row];
    FwdRec* dst = out + boff[row - row_lo];
    for (uint32_t e = 0; e < n; ++e) dst[e] = src[e];
}
// ---- local2global_ as (camera id, segment) pairs (l3d_get_local2global): the host used to do a binary search
// over the views per id after the copy ----
__global__ void __launch_bounds__(256) l2g_camseg_kernel(const uint32_t* __restrict__ l2g, uint32_t n,
                                                         const uint32_t* __restrict__ seg_view,
                                                         const ViewDev* __restrict__ views, uint2* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = l2g[i];
    const ViewDev& v = views[seg_view[g]];
    out[i] = make_uint2(v.cam_id, g - v.seg_off);
}
int launch_l2g_camseg(const uint32_t* l2g, uint32_t n, const uint32_t* seg_view, const ViewDev* views, uint2* out,
                      cudaStream_t st)
{
    if (!n) return 0;
    l2g_camseg_kernel<<<(n + 255) / 256, 256, 0, st>>>(l2g, n, seg_view, views, out);
    return 1;
}

// ---- FORWARD records (abi.cu): own rows into the canonical layout; pair blocks for the all-to-all ----
// copies the rows' records: src_off / dst_off index the two stores, n from cnt (0 = skip)
__global__ void __launch_bounds__(256) fwd_move_kernel(const uint32_t* __restrict__ cnt, uint32_t row_lo, uint32_t row_hi,
                                                       const uint32_t* __restrict__ src_off, const FwdRec* __restrict__ src,
                                                       const uint32_t* __restrict__ dst_off, FwdRec* __restrict__ dst)
{
    const uint32_t row = row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_hi) return;
    const uint32_t n = cnt[row - row_lo];
    if (!n) return;
    const FwdRec* a = src + src_off[row];
    FwdRec* b = dst + dst_off[row];
    for (uint32_t e = 0; e < n; ++e) b[e] = a[e];
}
// copies record blocks: item i = {src record, dst record, count}; one CTA per (item, 128-record chunk), two threads per 32-byte record
__global__ void __launch_bounds__(256) rec_blocks_kernel(const uint4* __restrict__ items, const uint32_t* __restrict__ chunk_item,
                                                         const uint32_t* __restrict__ chunk_first,
                                                         const FwdRec* __restrict__ src, FwdRec* __restrict__ dst)
{
    const uint32_t it = chunk_item[blockIdx.x];
    const uint4 b = items[it];
    const uint32_t k = (blockIdx.x - chunk_first[it]) * 128u + (threadIdx.x >> 1);
    if (k >= b.z) return;
    // 32-byte records as two 16-byte halves per thread pair
    const uint4* s4 = reinterpret_cast<const uint4*>(src + b.x + k) + (threadIdx.x & 1);
    uint4* d4 = reinterpret_cast<uint4*>(dst + b.y + k) + (threadIdx.x & 1);
    *d4 = *s4;
}
int launch_rec_blocks(const uint4* items, const uint32_t* chunk_item, const uint32_t* chunk_first, uint32_t n_chunks,
                      const FwdRec* src, FwdRec* dst, cudaStream_t st)
{
    if (!n_chunks) return 0;
    rec_blocks_kernel<<<n_chunks, 256, 0, st>>>(items, chunk_item, chunk_first, src, dst);
    return 1;
}
int launch_fwd_move(const uint32_t* cnt, uint32_t row_lo, uint32_t row_hi, const uint32_t* src_off, const FwdRec* src,
                    const uint32_t* dst_off, FwdRec* dst, cudaStream_t st)
{
    if (row_hi <= row_lo) return 0;
    fwd_move_kernel<<<(row_hi - row_lo + 255) / 256, 256, 0, st>>>(cnt, row_lo, row_hi, src_off, src, dst_off, dst);
    return 1;
}

// receiver: own records (local layout) and the boundary records of the other ranks -> canonical layout
struct SliceRows {
    uint32_t row[17];
};
__global__ void __launch_bounds__(256) fwd_place_kernel(const unsigned char* __restrict__ all, uint64_t stride, int world,
                                                        SliceRows sl, int rank, uint32_t n_rows,
                                                        const uint32_t* __restrict__ fwd_cnt,
                                                        const uint32_t* __restrict__ fwd_off,
                                                        const uint32_t* __restrict__ bcnt_all,
                                                        const uint32_t* __restrict__ boff_all,
                                                        const FwdRec* __restrict__ own_rec,
                                                        const uint32_t* __restrict__ own_off_local,
                                                        FwdRec* __restrict__ out)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t n = fwd_cnt[row];
    if (!n) return;
    int q = 0;
    while (q + 1 < world && sl.row[q + 1] <= row) ++q;
    const FwdRec* src = nullptr;
    if (q == rank) {
        src = own_rec + own_off_local[row];
    } else if (bcnt_all[row]) {
        const uint32_t rows_q = sl.row[q + 1] - sl.row[q];
        const unsigned char* blob = all + (uint64_t)q * stride + (((uint64_t)rows_q + 7ull) & ~7ull) * 4ull;
        src = reinterpret_cast<const FwdRec*>(blob) + (boff_all[row] - boff_all[sl.row[q]]);
    }
    if (!src) return;
    FwdRec* dst = out + fwd_off[row];
    for (uint32_t e = 0; e < n; ++e) dst[e] = src[e];
}
int launch_fwd_bmask(const PairDev* pairs, uint32_t P, const uint32_t* fwd_cnt, uint32_t row_lo, uint32_t row_hi,
                     uint32_t* bcnt, cudaStream_t st)
{
    if (row_hi <= row_lo) return 0;
    fwd_bmask_kernel<<<(row_hi - row_lo + 255) / 256, 256, 0, st>>>(pairs, P, fwd_cnt, row_lo, row_hi, bcnt);
    return 1;
}
int launch_fwd_bgather(const uint32_t* bcnt, const uint32_t* boff, const uint32_t* fwd_off_local, uint32_t row_lo,
                       uint32_t row_hi, const FwdRec* fwd_rec, FwdRec* out, cudaStream_t st)
{
    if (row_hi <= row_lo) return 0;
    fwd_bgather_kernel<<<(row_hi - row_lo + 255) / 256, 256, 0, st>>>(bcnt, boff, fwd_off_local, row_lo, row_hi,
                                                                       fwd_rec, out);
    return 1;
}
int launch_fwd_place(const unsigned char* all, uint64_t stride, int world, const uint32_t* slice_row, int rank,
                     uint32_t n_rows, const uint32_t* fwd_cnt, const uint32_t* fwd_off, const uint32_t* bcnt_all,
                     const uint32_t* boff_all, const FwdRec* own_rec, const uint32_t* own_off_local, FwdRec* out,
                     cudaStream_t st)
{
    if (!n_rows) return 0;
    SliceRows sl;
    for (int q = 0; q <= world && q < 17; ++q) sl.row[q] = slice_row[q];
    fwd_place_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(all, stride, world, sl, rank, n_rows, fwd_cnt, fwd_off,
                                                           bcnt_all, boff_all, own_rec, own_off_local, out);
    return 1;
}

// header of a self-describing exchange blob (abi.cu): payload size from the device-side cursor
__global__ void blob_hdr_kernel(unsigned long long* dst, uint64_t fixed_bytes, uint64_t elem_bytes,
                                const uint32_t* n_dev, uint64_t n_imm, const uint32_t* flags_dev, uint32_t kind)
{
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_imm;
    dst[0] = fixed_bytes + n * elem_bytes;
    dst[1] = (uint64_t)(flags_dev ? *flags_dev : 0u) | ((uint64_t)kind << 32);
    dst[2] = 0;
    dst[3] = 0;
}
int launch_blob_hdr(void* dst, uint64_t fixed_bytes, uint64_t elem_bytes, const uint32_t* n_dev, uint64_t n_imm,
                    const uint32_t* flags_dev, uint32_t kind, cudaStream_t st)
{
    blob_hdr_kernel<<<1, 1, 0, st>>>((unsigned long long*)dst, fixed_bytes, elem_bytes, n_dev, n_imm, flags_dev, kind);
    return 1;
}

// multi-GPU: adopt the all-gathered hypotheses (abi.cu, L3D_X_HYPOTHESES): entries, filtered-list
// counts and offsets of every row, and the filtered records of every slice packed into one store
struct HypHdr32 {
    unsigned long long a, b;
    uint32_t c, d, e, f;
};
__global__ void __launch_bounds__(128) hyp_adopt_kernel(const unsigned char* __restrict__ all, uint64_t stride, int world,
                                                        const uint32_t* __restrict__ slice_g,
                                                        const uint32_t* __restrict__ fbase, uint32_t S,
                                                        EntryDev* __restrict__ entries, uint32_t* __restrict__ filt_cnt,
                                                        uint32_t* __restrict__ filt_off, ListRec* __restrict__ filt_all)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= S) return;
    int q = 0;
    while (q + 1 < world && slice_g[q + 1] <= g) ++q;
    const uint32_t rows = slice_g[q + 1] - slice_g[q], rp = (rows + 3u) & ~3u, i = g - slice_g[q];
    const unsigned char* blob = all + (uint64_t)q * stride + sizeof(HypHdr32);
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(blob);
    const uint32_t* off = cnt + rp;
    const EntryDev* ent = reinterpret_cast<const EntryDev*>(blob + 8ull * rp);
    const ListRec* rec = reinterpret_cast<const ListRec*>(blob + 8ull * rp + (uint64_t)rows * sizeof(EntryDev));
    const uint32_t n = cnt[i], o = off[i];
    entries[g] = ent[i];
    filt_cnt[g] = n;
    filt_off[g] = fbase[q] + o;
    for (uint32_t z = 0; z < n; ++z) filt_all[fbase[q] + o + z] = rec[o + z];
}
int launch_hyp_adopt(const void* all, uint64_t stride, int world, const uint32_t* slice_g, const uint32_t* fbase,
                     uint32_t S, EntryDev* entries, uint32_t* filt_cnt, uint32_t* filt_off, ListRec* filt_all,
                     cudaStream_t st)
{
    if (!S) return 0;
    hyp_adopt_kernel<<<(S + 127) / 128, 128, 0, st>>>((const unsigned char*)all, stride, world, slice_g, fbase, S,
                                                       entries, filt_cnt, filt_off, filt_all);
    return 1;
}

}  // namespace l3d
